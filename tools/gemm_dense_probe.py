"""Debug: graph-timed encoder-shaped dense GEMMs (operand-form in/out) + CTA phase timeline."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unidet3d_b200 import ops, _lib  # noqa: E402

lib = _lib.load()
lib.ud3d_debug_set_flags.argtypes = [C.c_int]
lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
T = int(os.environ.get("T", 13447))
for name, cin, cout, act, raw in [("qkv 256->768", 256, 768, None, False), ("out 256->256 (+res)", 256, 256, None, True),
                                  ("ffn1 256->1024 gelu", 256, 1024, "gelu", False), ("ffn2 1024->256 (+res)", 1024, 256, None, True)]:
    x = torch.randn(T, cin, device="cuda")
    xs = ops.act_split(x, relu=False)
    w = ops.PackedWeight(torch.randn(cout, 1, cin, device="cuda") * 0.05)
    bias = torch.randn(cout, device="cuda")
    res = torch.randn(T, cout, device="cuda") if raw else None
    o_s = torch.empty(T, cout, device="cuda")
    out = torch.empty(T, cout, device="cuda")
    if raw:
        run = lambda: ops.gemm(xs, w, bias=bias, residual=res, in_split=True, out=out)
    else:
        run = lambda: ops.gemm(xs, w, bias=bias, act=act, in_split=True, no_raw=True, out=out, acts=[(o_s, None, None, False)])
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            run()
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 100)
    fl = 2.0 * T * cin * cout
    print(f"{name:24s}: {np.median(ts):7.1f} us/launch  {fl / np.median(ts) / 1e6:6.1f} TFLOP/s (x3 on the tensor pipe)")
    trace = torch.zeros(8 * 8192, dtype=torch.int64, device="cuda")
    lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
    run(); torch.cuda.synchronize(); lib.ud3d_debug_set_trace(None, 0)
    t = trace.cpu().numpy().reshape(-1, 8); t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = lambda col: (t[:, col] - t0) / 1e3
    prev = rel(0)
    line = f"    {len(t)} CTAs (column tile 0), span {rel(1).max():.1f} us, start p50/p100 {np.percentile(prev, 50):.1f}/{prev.max():.1f}:"
    for col, nm in ((3, "prologue"), (4, "main loop"), (1, "epilogue")):
        if (t[:, col] > 0).all():
            cur = rel(col)
            line += f" {nm} +{np.mean(cur - prev):.2f} (max {np.max(cur - prev):.2f})"
            prev = cur
    print(line)
