"""Per-step trace of one CTA of the TS gather-GEMM (ud3d_debug_set_trace) + timing with parts disabled
(ud3d_debug_set_flags: 1 no MMA, 16 no epilogue stores, 32 no gathers (all rows read the zero row))."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200 import ops, _lib  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict  # noqa: E402
from tools.ts_probe import timed  # noqa: E402,F401

lib = _lib.load()
lib.ud3d_debug_set_flags.argtypes = [C.c_int]
lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]

cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
x, inv = model.collate(torch.cat(pts), offs, len(pts))
for level in [int(v) for v in os.environ.get("LEVELS", "0,1").split(",")]:
    lv = x.pyramid.levels[level]
    c = cfg["backbone"]["num_planes"][level]
    xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
    xs = ops.act_split(xin, relu=False)
    w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
    act = torch.empty_like(xin)
    one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
    tb, tm, pm = lv.subm_conv
    xs_il = ops.operand_form_interleave(xs)
    run = lambda: ops.gemm(xs_il, w, table=tb, tile_mask=tm, in_split=2, no_raw=True, acts=[(act, one, zero)], row_perm=pm)
    for fl in (0,):
        lib.ud3d_debug_set_flags(fl if fl > 2 else 0)
        print(f"level {level} c={c} flags {fl:3d}: {timed(run):8.1f} us", flush=True)
    lib.ud3d_debug_set_flags(int(os.environ.get("TRACE_FLAGS", "0")))
    trace = torch.zeros(4096, dtype=torch.int64, device="cuda")
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -3)
    run(); torch.cuda.synchronize()
    tt = trace.cpu().numpy()[:4 * 148].reshape(148, 4)
    tt = tt[tt[:, 0] > 0]
    t0g = tt[:, 0].min()
    dur = (tt[:, 1] - tt[:, 0]) / 1e3
    print(f"  {len(tt)} CTAs on {len(set(tt[:, 2]))} SMs: start spread {(tt[:, 0].max() - t0g) / 1e3:.1f} us, duration min/median/max "
          f"{dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f} us, span {(tt[:, 1].max() - t0g) / 1e3:.1f} us")
    trace.zero_()
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), 7)
    run(); torch.cuda.synchronize()
    lib.ud3d_debug_set_trace(None, 0)
    lib.ud3d_debug_set_flags(0)
    t = trace.cpu().numpy()
    st = t[:2048].reshape(256, 8).astype(np.float64)
    tl = t[2048:2048 + 256].reshape(64, 4).astype(np.float64)
    t0 = st[st > 0].min()
    print("  step:  mma_full  mma_issued | prod_wait  prod_has  published  st_issued | loop_top  rows_issued   (cycles from first event)")
    for g in range(0, 72):
        r = st[g]
        print("  %3d: " % g + " ".join("%9.0f" % (v - t0) if v > 0 else "        -" for v in r[:8]))
    print("  tile: epi_sees_acc  epi_done  mma_start")
    for i in range(0, 10):
        r = tl[i]
        print("  %3d: " % i + " ".join("%9.0f" % (v - t0) if v > 0 else "        -" for v in r))
    full = st[:, 0]; ok = full > 0
    d = np.diff(full[ok])
    print(f"  MMA-warp step interval: median {np.median(d):.0f} cycles, mean {d.mean():.0f}; issue time per step (full->issued) median {np.median((st[:,1]-st[:,0])[ok]):.0f}")
