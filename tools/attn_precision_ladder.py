"""CPU study (no GPU needed): which of the bf16 split terms of the attention products does the 1e-3 gate on the
encoder's final logits / boxes actually need?  Emulates the tensor-core operand rounding inside the oracle's
self-attention (everything else stays fp32) on the full-size encoder (d = 256, 8 heads, 6 layers).
Terms: x = hi + lo (two bf16).  QK: which of {hi.hi, lo.hi, hi.lo} are kept; PV likewise."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import encoder as oenc  # noqa: E402
from unidet3d_b200 import configs  # noqa: E402


def split(x):
    hi = x.bfloat16().float()
    return hi, (x - hi).bfloat16().float()


def prod(a, b, terms):
    """a @ b with the given split terms: 'hh', 'lh' (a_lo b_hi), 'hl' (a_hi b_lo); 'f' = fp32."""
    if terms == "f":
        return a @ b
    ah, al = split(a)
    bh, bl = split(b)
    out = 0
    for t in terms.split("+"):
        out = out + {"hh": ah @ bh, "lh": al @ bh, "hl": ah @ bl}[t]
    return out


def run(mode_qk, mode_pv, T=(1700, 1500), seed=0, sharpen=1.0):
    cfg = configs.model_cfg(("scannet",))
    ocfg = configs.oracle_cfg(cfg)["encoder"]
    n_union = len(set(sum(cfg["decoder"]["datasets_classes"], []))) + 1
    sd = oenc.make_encoder_state_dict(6, 32, 256, 1024, n_union, seed)
    if sharpen != 1.0:      # emulate a trained encoder: larger q / k projections -> peaked softmax (a few keys dominate)
        for k in list(sd):
            if k.endswith("attn.in_proj_weight"):
                w = sd[k].clone()
                w[:512] *= sharpen
                sd[k] = w
    g = torch.Generator().manual_seed(1)
    x = [torch.randn(t, 32, generator=g) for t in T]
    c = [torch.randn(t, 3, generator=g) for t in T]
    orig = oenc.self_attention

    def attn(sd_, p, xx, num_heads):
        d = xx.shape[1]
        hd = d // num_heads
        qkv = xx @ sd_[p + ".attn.in_proj_weight"].t() + sd_[p + ".attn.in_proj_bias"]
        q, k, v = qkv.split(d, dim=1)
        Tt = xx.shape[0]
        q = q.view(Tt, num_heads, hd).transpose(0, 1)
        k = k.view(Tt, num_heads, hd).transpose(0, 1)
        v = v.view(Tt, num_heads, hd).transpose(0, 1)
        s = prod(q, k.transpose(1, 2), mode_qk) / math.sqrt(hd)
        pm = torch.exp(s - s.max(-1, keepdim=True).values)          # unnormalised, like the online softmax
        den = pm.sum(-1, keepdim=True)
        if mode_pv != "f" and "lh" not in mode_pv:      # P enters as a single bf16 term: the row sum uses the same rounded values
            den = pm.bfloat16().float().sum(-1, keepdim=True)
        o = prod(pm, v, mode_pv) / den
        o = o.transpose(0, 1).reshape(Tt, d)
        z = o @ sd_[p + ".attn.out_proj.weight"].t() + sd_[p + ".attn.out_proj.bias"] + xx
        return F.layer_norm(z, (d,), sd_[p + ".norm.weight"], sd_[p + ".norm.bias"], 1e-5)

    ref = oenc.encoder_forward(sd, ocfg, x, c, ["scannet"] * len(T), all_heads=False)
    oenc.self_attention = attn
    try:
        out = oenc.encoder_forward(sd, ocfg, x, c, ["scannet"] * len(T), all_heads=False)
    finally:
        oenc.self_attention = orig
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    return max(rel(out["cls_preds"][i], ref["cls_preds"][i]) for i in range(len(T))), max(
        rel(out["bboxes"][i], ref["bboxes"][i]) for i in range(len(T)))


if __name__ == "__main__":
    torch.set_num_threads(8)
    sharpen = float(os.environ.get("SHARPEN", "1"))
    for qk, pv in [("hh+lh+hl", "hh+lh+hl"), ("hh+lh+hl", "hh+hl"), ("hh+lh+hl", "hh+lh"), ("hh+lh+hl", "hh"),
                   ("hh+lh", "hh+lh+hl"), ("hh", "hh+lh+hl"), ("hh+lh", "hh"), ("hh", "hh")]:
        e = run(qk, pv, sharpen=sharpen)
        print(f"sharpen {sharpen}: QK {qk:9s} PV {pv:9s}: logits rel err {e[0]:.2e}  boxes {e[1]:.2e}", flush=True)
