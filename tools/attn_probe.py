"""Time / cross-check the two attention kernels (mma.sync operand-form kernel vs tcgen05 kernel with P in TMEM) on the
bench's token distribution (8 scenes x ~1680 superpoints, 8 heads) and at T = 4096."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from opform import split_encode, split_decode  # noqa: E402
from unidet3d_b200 import ops  # noqa: E402


def timed(run, reps=6):
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run()
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return float(np.median(ts[2:]))


for lens in ([1700, 1650, 1688, 1702, 1671, 1690, 1660, 1686], [4096], [4096] * 4, [300, 37, 1]):
    g = torch.Generator().manual_seed(1)
    H, d = 8, 256
    qkv = torch.randn(sum(lens), 3 * d, generator=g) * float(os.environ.get("SCALE", "1.5"))
    cu = torch.tensor(np.cumsum([0] + lens), dtype=torch.int32).cuda()
    ref = []
    for i, T in enumerate(lens):
        s = qkv[int(cu[i]):int(cu[i + 1])].double()
        q, k, v = [s[:, j * d:(j + 1) * d].view(T, H, 32).transpose(0, 1) for j in range(3)]
        a = torch.softmax(q @ k.transpose(1, 2) / 32 ** 0.5, -1) @ v
        ref.append(a.transpose(0, 1).reshape(T, d))
    ref = torch.cat(ref)
    x = split_encode(qkv).cuda()
    for name, tc in (("mma.sync", False), ("tcgen05", True)):
        run = lambda: ops.attention(x, cu, max(lens), H, split_in=True, tcgen05=tc)
        try:
            o = run(); torch.cuda.synchronize()
            err = float((split_decode(o.cpu()).double() - ref).abs().max() / ref.abs().max())
            print(f"lens {str(lens)[:40]:40s} {name:9s}: {timed(run):8.1f} us   rel err {err:.2e}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"lens {lens} {name}: FAILED {e}")
