"""Debug: time the level-l SubM3 conv (operand-form in/out) with parts of the kernel disabled
(ud3d_debug_set_flags bits: 1 no MMA, 2 no weight copies, 4 no zero fill, 8 no proxy fence, 16 no epilogue
stores, 32 no gather)."""
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
lib = _lib.load()
lib.ud3d_debug_set_flags.argtypes = [C.c_int]
lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
flag_sets = [int(f) for f in os.environ.get("FLAGS", "0,4,8,16,2,1,3,32,33,35,51,63").split(",")]
for level in [int(v) for v in os.environ.get("LEVELS", "0,1").split(",")]:
    lv = x.pyramid.levels[level]
    c = cfg["backbone"]["num_planes"][level]
    xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
    xs = ops.act_split(xin, relu=False)
    w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
    act = torch.empty_like(xin)
    one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
    run = lambda: ops.gemm(xs, w, table=lv.subm, tile_mask=lv.subm_mask, in_split=True, no_raw=True, acts=[(act, one, zero)])
    for fl in flag_sets:
        lib.ud3d_debug_set_flags(fl)
        run()
        torch.cuda.synchronize()
        # GPU-only time: REPS launches replayed from a CUDA graph (no host launch latency between them)
        REPS = 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(REPS):
                run()
        ts = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / REPS)
        # host-issued: one launch between two events (includes the Python / ctypes launch latency)
        th = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            th.append(e0.elapsed_time(e1) * 1e3)
        print(f"level {level} n={lv.n} c={c} flags={fl:5d}: graph {np.median(ts[2:]):8.1f} us/launch   host-issued {np.median(th[2:]):8.1f} us")
        if os.environ.get("TIMELINE"):
            trace = torch.zeros(8 * 8192, dtype=torch.int64, device="cuda")
            lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
            run(); torch.cuda.synchronize(); lib.ud3d_debug_set_trace(None, 0)
            t = trace.cpu().numpy().reshape(-1, 8); t = t[t[:, 0] > 0]
            t0 = t[:, 0].min()
            rel = lambda col: (t[:, col] - t0) / 1e3
            names = {3: "prologue", 4: "main loop", 5: "park", 6: "cluster barrier", 7: "reduce+store", 1: "end"}
            prev = rel(0)
            line = f"    {len(t)} CTAs, span {rel(1).max():.1f} us, start p50/p100 {np.percentile(prev, 50):.1f}/{prev.max():.1f}:"
            for col in (3, 4, 5, 6, 7, 1):
                if (t[:, col] > 0).all():
                    cur = rel(col)
                    line += f" {names[col]} +{np.mean(cur - prev):.2f} (max {np.max(cur - prev):.2f})"
                    prev = cur
            print(line)
    lib.ud3d_debug_set_flags(0)
