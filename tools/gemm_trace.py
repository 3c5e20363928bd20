"""Timeline of the gather-GEMM (gemm.cu) on one backbone level: per-CTA phases from %globaltimer
(ud3d_debug_set_trace(buf, -2)) and the per-step clock64 trace of one CTA (ud3d_debug_set_trace(buf, cta))."""
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
from tools.ts_probe import timed  # noqa: E402

lib = _lib.load()
lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]

cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
x, inv = model.collate(torch.cat(pts), offs, len(pts))
def cta_timeline(run, n_ctas, label):
    trace = torch.zeros(8 * n_ctas + 64, dtype=torch.int64, device="cuda")
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
    run(); torch.cuda.synchronize()
    lib.ud3d_debug_set_trace(None, 0)
    tl = trace.cpu().numpy()[:8 * n_ctas].reshape(n_ctas, 8).astype(np.float64)
    tl = tl[tl[:, 0] > 0]
    t0 = tl[:, 0].min()
    start, end, smid, pro, main = tl[:, 0] - t0, tl[:, 1] - t0, tl[:, 2].astype(int), tl[:, 3] - t0, tl[:, 4] - t0
    print(f"  {label}: {len(tl)} CTAs (blockIdx.y == 0) on {len(set(smid))} SMs, span {end.max() / 1e3:.1f} us; start spread {start.max() / 1e3:.1f} us; per CTA (us): "
          f"prologue {np.median(pro - start) / 1e3:.2f}, main loop {np.median(main - pro) / 1e3:.2f}, epilogue {np.median(end - main) / 1e3:.2f}, "
          f"life {np.median(end - start) / 1e3:.2f}")


if os.environ.get("DENSE", "1") == "1":
    T = 13447
    for (ci, co, actf, res) in ((256, 768, None, False), (256, 256, None, True), (256, 1024, "gelu", False), (1024, 256, None, True)):
        xin = torch.randn(T, ci, device="cuda")
        xs = ops.act_split(xin, relu=False)
        w = ops.PackedWeight(torch.randn(co, ci, device="cuda") * 0.05)
        b = torch.randn(co, device="cuda")
        r = torch.randn(T, co, device="cuda") if res else None
        act = torch.zeros(T, co, device="cuda")
        run = lambda: ops.gemm(xs, w, in_split=True, bias=b, act=actf, residual=r, no_raw=not res, acts=[(act, None, None, False)])
        print(f"dense {ci}->{co} {actf} res={res}: {timed(run):8.1f} us/launch", flush=True)
        cta_timeline(run, (T + 127) // 128, "timeline")

for level in [int(v) for v in os.environ.get("LEVELS", "0,1").split(",") if v != ""]:
    lv = x.pyramid.levels[level]
    c = cfg["backbone"]["num_planes"][level]
    xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
    xs = ops.act_split(xin, relu=False)
    w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
    act = torch.empty_like(xin); raw = torch.empty_like(xin); res = torch.randn_like(xin)
    one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
    tb, tm, pm = lv.subm_conv
    run = lambda: ops.gemm(xs, w, table=tb, tile_mask=tm, in_split=1, residual=res, out=raw, acts=[(act, one, zero)], row_perm=pm)
    print(f"level {level} n={lv.n} c={c}: {timed(run):8.1f} us/launch (graph replay)", flush=True)
    n_tiles = (lv.n + 127) // 128
    masks = tm.cpu().numpy().astype(np.uint32)
    nact = np.array([bin(int(m)).count("1") for m in masks])
    print(f"  {n_tiles} tiles, active offsets per tile: mean {nact.mean():.1f} min {nact.min()} max {nact.max()}")
    trace = torch.zeros(8 * n_tiles + 64, dtype=torch.int64, device="cuda")
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
    run(); torch.cuda.synchronize()
    tl = trace.cpu().numpy()[:8 * n_tiles].reshape(n_tiles, 8).astype(np.float64)
    t0 = tl[:, 0].min()
    start, end, smid, pro, main = tl[:, 0] - t0, tl[:, 1] - t0, tl[:, 2].astype(int), tl[:, 3] - t0, tl[:, 4] - t0
    print(f"  kernel span {end.max() / 1e3:.1f} us on {len(set(smid))} SMs; per CTA (us): prologue {np.median(pro - start) / 1e3:.2f}, "
          f"main loop {np.median(main - pro) / 1e3:.2f}, epilogue {np.median(end - main) / 1e3:.2f}, life {np.median(end - start) / 1e3:.2f}")
    per_step = (main - pro) / np.maximum(nact * ((c + 31) // 32), 1)
    print(f"  main loop per step: median {np.median(per_step):.0f} ns; CTAs per SM (mean) {n_tiles / len(set(smid)):.1f}; "
          f"sum of CTA lives / (SMs x span) = {np.sum(end - start) / (len(set(smid)) * end.max()):.2f} resident CTAs")
    # occupancy over time of one SM
    sm0 = smid[0]
    sel = np.where(smid == sm0)[0]
    order = sel[np.argsort(start[sel])]
    print(f"  SM {sm0}: CTA start / prologue / main / end (us)")
    for i in order[:12]:
        print(f"    tile {i:5d} nact {nact[i]:2d}: {start[i] / 1e3:7.2f} {pro[i] / 1e3:7.2f} {main[i] / 1e3:7.2f} {end[i] / 1e3:7.2f}")
    trace.zero_()
    cta = int(order[4]) if len(order) > 4 else 0
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), cta)
    run(); torch.cuda.synchronize()
    lib.ud3d_debug_set_trace(None, 0)
    t = trace.cpu().numpy().astype(np.float64)
    ns = int(t[1023])
    st = t[:8 * 64].reshape(64, 8)
    c0 = t[1019]
    print(f"  CTA {cta}: {ns} steps; cycles from tile start: producers start {t[1022] - c0:.0f}, epilogue starts {t[1021] - c0:.0f}, end {t[1020] - c0:.0f}")
    print(f"   entry {t[1014] - c0:.0f}, barriers+TMEM ready {t[1015] - c0:.0f}, table slice in smem {t[1017] - c0:.0f}, offsets listed {t[1018] - c0:.0f}")
    print("   step: prod_top prod_has_stage prod_issued | mma_full mma_issued")
    for g in range(min(ns, 30)):
        r = st[g]
        print("   %3d: " % g + " ".join("%8.0f" % (r[k] - c0) if r[k] > 0 else "       -" for k in (0, 1, 2, 4, 6)))
