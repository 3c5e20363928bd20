"""Host-side encode / decode of the "operand form" of a feature map, shared by the GPU tests.

Per row, per 32-channel chunk: 128 bytes = 8 pieces of 16 bytes.  Layout 1 (INTERLEAVED = False): pieces 0..3 = bf16
hi of channels 0..31, pieces 4..7 = bf16 lo.  Layout 2 (INTERLEAVED = False): piece 2q = bf16 hi of channels 8q..8q+7,
piece 2q+1 = bf16 lo of the same channels (x ~= hi + lo): one 32-byte load = one thread's share of a tcgen05.st."""
import torch

INTERLEAVED = False


def split_encode(x, interleaved=None):
    """fp32 [N,C] -> operand form [N,C] (viewed as fp32)."""
    N, C = x.shape
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    if INTERLEAVED if interleaved is None else interleaved:
        packed = torch.stack([hi.view(N, C // 32, 4, 8), lo.view(N, C // 32, 4, 8)], dim=3).contiguous()   # [N, chunk, q, hi|lo, 8]
    else:
        packed = torch.cat([hi.view(N, C // 32, 32), lo.view(N, C // 32, 32)], dim=2).contiguous()
    return packed.view(N, C * 2).view(torch.float32).view(N, C)


def split_decode(s):
    N, C = s.shape
    b = s.contiguous().view(torch.bfloat16)
    if INTERLEAVED:
        b = b.view(N, C // 32, 4, 2, 8).float()
        return (b[:, :, :, 0] + b[:, :, :, 1]).reshape(N, C)
    b = b.view(N, C // 32, 64).float()
    return (b[:, :, :32] + b[:, :, 32:]).reshape(N, C)
