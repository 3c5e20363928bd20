"""Debug: per-step clock64 trace of one CTA of the level-1 SubM3 conv (operand-form input)."""
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

cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
x, inv = model.collate(torch.cat(pts), offs, len(pts))
lv = x.pyramid.levels[int(os.environ.get("LEVEL", 0))]
c = cfg["backbone"]["num_planes"][int(os.environ.get("LEVEL", 0))]
xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
xs = ops.act_split(xin, relu=False)
w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
act = torch.empty_like(xin)
one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
lib = _lib.load()
lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
lib.ud3d_debug_set_flags.argtypes = [C.c_int]
lib.ud3d_debug_set_flags(int(os.environ.get("FLAGS", 0)))
for _ in range(3):
    ops.gemm(xs, w, table=lv.subm, tile_mask=lv.subm_mask, in_split=True, no_raw=True, acts=[(act, one, zero)])
torch.cuda.synchronize()
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
blk = int(os.environ.get("BLOCK", 100))
lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), blk)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.gemm(xs, w, table=lv.subm, tile_mask=lv.subm_mask, in_split=True, no_raw=True, acts=[(act, one, zero)])
e1.record()
torch.cuda.synchronize()
lib.ud3d_debug_set_trace(None, 0)
t = trace.cpu().numpy()
n = int(t[1023]); t0 = t[1019]
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us; block {blk}: nsteps={n}; prologue->first step {t[1022]-t0} cyc; "
      f"acc_full at {t[1021]-t0}; epilogue end {t[1020]-t0}")
print("step: empty_ok  issued  landed  published | a_full  b_full  committed   (cycles since kernel start of this CTA)")
for i in range(min(n, 40)):
    r = t[i * 8:(i + 1) * 8] - t0
    print(f"{i:3d}: {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} | {r[4]:8d} {r[5]:8d} {r[6]:8d}")
