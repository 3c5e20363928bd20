"""Time the attention kernels on the bench workload shape (8 scenes x ~1680 tokens, 8 heads)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unidet3d_b200 import ops
lens = [1822, 1586, 1700, 1650, 1733, 1690, 1610, 1656]
T = sum(lens)
cu = torch.tensor(np.cumsum([0] + lens), dtype=torch.int32, device="cuda")
qkv = torch.randn(T, 768, device="cuda")
qs = ops.act_split(qkv, relu=False)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
a = ops.attention(qs, cu, max(lens), 8, split_in=True)
b = ops.attention(qs, cu, max(lens), 8, split_in=True, tcgen05=True)
print("max diff tc vs mma (operand-form bits as fp32 views are not comparable; decode):")
def dec(s):
    N, C = s.shape
    v = s.contiguous().view(torch.bfloat16).view(N, C // 32, 64).float()
    return (v[:, :, :32] + v[:, :, 32:]).reshape(N, C)
print(float((dec(a) - dec(b)).abs().max()), float(dec(a).abs().max()))
print("mma.sync :", round(t(lambda: ops.attention(qs, cu, max(lens), 8, split_in=True)), 1), "us")
print("tcgen05  :", round(t(lambda: ops.attention(qs, cu, max(lens), 8, split_in=True, tcgen05=True)), 1), "us")
