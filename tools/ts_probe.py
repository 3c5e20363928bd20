"""Time / cross-check the two gather-GEMM kernels (A operand through shared memory vs registers -> TMEM) on the level-l
SubM3 convs of the bench workload, in canonical and regrouped (ud3d_subm3_tile_order) row order, and on encoder-shaped
dense GEMMs.  FLAGS bit 4096 of ud3d_debug_set_flags forces the shared-memory kernel."""
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

lib = _lib.load()
lib.ud3d_debug_set_flags.argtypes = [C.c_int]


def timed(run, reps=10):
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


def main():
    cfg, scenes, names, preset = bench.make_workload(os.environ.get("WORKLOAD", "scannet_b8"), 0)
    model = u.MODELS.build(cfg).eval()
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.cuda()
    pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
    x, inv = model.collate(torch.cat(pts), offs, len(pts))
    for level in [int(v) for v in os.environ.get("LEVELS", "0,1,2").split(",")]:
        lv = x.pyramid.levels[level]
        c = cfg["backbone"]["num_planes"][level]
        xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
        xs = ops.act_split(xin, relu=False)
        w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
        res = torch.randn(lv.n, c, device="cuda")
        one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
        tb, tm, pm = lv.subm_conv
        outs = {}
        for order, (t_, m_, p_) in (("canonical", (lv.subm, lv.subm_mask, None)), ("regrouped", (tb, tm, pm))):
            for path, fl in (("smem", 1), ("tmem", 2)):
                lib.ud3d_debug_set_flags(fl if fl > 2 else 0)
                act = torch.zeros_like(xin); raw = torch.zeros_like(xin)
                xin_ = xs if fl == 1 else ops.operand_form_interleave(xs)
                run = lambda: ops.gemm(xin_, w, table=t_, tile_mask=m_, in_split=fl, residual=res, out=raw, acts=[(act, one, zero)], row_perm=p_)
                try:
                    us = timed(run)
                except Exception as e:  # noqa: BLE001
                    print(f"level {level} {order} {path}: FAILED {e}")
                    continue
                outs[(order, path)] = (raw.clone(), act.clone())
                print(f"level {level} n={lv.n} c={c} {order:9s} {path}: {us:8.1f} us/launch", flush=True)
        lib.ud3d_debug_set_flags(0)
        ref = outs.get(("canonical", "smem"))
        for k, v in outs.items():
            if ref is not None and k != ("canonical", "smem"):
                e0 = float((v[0] - ref[0]).abs().max() / ref[0].abs().max())
                e1 = float((v[1].view(torch.int32) != ref[1].view(torch.int32)).float().mean())
                print(f"   {k}: raw rel err vs canonical/smem {e0:.2e}, operand-form words differing {e1:.2e}")
    # encoder-shaped dense GEMMs (the gather-GEMM's identity-gather path), checked against fp32 torch
    T = 13447
    for (ci, co, actf, res) in ((256, 768, None, False), (256, 256, None, True), (256, 1024, "gelu", False), (1024, 256, None, True),
                                (256, 100, None, False)):
        xin = torch.randn(T, ci, device="cuda")
        xs = ops.act_split(xin, relu=False)
        wt = torch.randn(co, ci, device="cuda") * 0.05
        w = ops.PackedWeight(wt)
        b = torch.randn(co, device="cuda")
        r = torch.randn(T, co, device="cuda") if res else None
        if co % 32 == 0:
            act = torch.zeros(T, co, device="cuda")
            run = lambda: ops.gemm(xs, w, in_split=True, bias=b, act=actf, residual=r, no_raw=False, acts=[(act, None, None, False)])
        else:
            run = lambda: ops.gemm(xs, w, in_split=True, bias=b, act=actf)
        us = timed(run)
        o = run(); torch.cuda.synchronize()
        ref = xin.double() @ wt.double().t() + b.double()
        if actf == "gelu":
            ref = torch.nn.functional.gelu(ref)
        if res:
            ref = ref + r.double()
        e = float((o.double() - ref).abs().max() / ref.abs().max())
        print(f"dense {ci}->{co} {actf} res={res}: {us:8.1f} us/launch   rel err vs fp64 {e:.2e}", flush=True)


if __name__ == "__main__":
    main()
