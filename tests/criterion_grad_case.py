"""Shared by the CPU and GPU tests of the criterion's GRADIENTS against tests/golden/criterion_grad_ref.npz -- the
reference's own criterion.py (+ axis_aligned_iou_loss.py) differentiated with torch.autograd in the build container
(tests/golden/make_golden.py::gen_criterion_grad).  ``check(run)`` feeds the fixture's three layers to ``run`` and
compares what comes back; the CPU test passes the oracle under autograd, the GPU test the library's kernels."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
N_LAYERS, N_SCENES = 3, 4


def load():
    fx = np.load(os.path.join(HERE, "golden", "criterion_grad_ref.npz"))
    names = [str(n) for n in fx["names"]]
    layers = [dict(cls_preds=[torch.as_tensor(fx[f"l{l}_cls{i}"]) for i in range(N_SCENES)],
                   bboxes=[torch.as_tensor(fx[f"l{l}_box{i}"]) for i in range(N_SCENES)]) for l in range(N_LAYERS)]
    gts = [dict(labels=torch.as_tensor(fx[f"gt_labels{i}"]).long(), boxes=torch.as_tensor(fx[f"gt_boxes{i}"]),
                query_masks=torch.as_tensor(fx[f"qmask{i}"])) for i in range(N_SCENES)]
    cfg = dict(datasets=[str(d) for d in fx["datasets"]], datasets_weights=[float(w) for w in fx["datasets_weights"]],
               topk=[int(k) for k in fx["topk"]], loss_weight=[0.5, 1.0], non_object_weight=0.1, w_cls=0.5, w_box=2.0, iter_matcher=True)
    return fx, names, layers, gts, cfg


def ref_match(fx, l, i, T, G):
    m = torch.zeros((T, G), dtype=torch.bool)
    if G:
        m[torch.as_tensor(fx[f"l{l}_iq{i}"]).long(), torch.as_tensor(fx[f"l{l}_ig{i}"]).long()] = True
    return m


def check(run, rtol=5e-4, atol=2e-6, loss_rtol=1e-4, min_compared=10):
    """``run(names, layers, gts, cfg) -> (loss, d_cls, d_box, matches)``: layers[0] is the final layer, layers[1:] the
    auxiliary ones (the order the reference's ``pred`` / ``aux_outputs`` had); d_cls / d_box / matches are indexed
    [layer][scene] in the same order (a missing gradient may be None).  A (layer, scene) whose matching differs from the
    reference's (a flipped near-tie) is left out of the gradient comparison; at most two may be."""
    fx, names, layers, gts, cfg = load()
    loss, d_cls, d_box, matches = run(names, layers, gts, cfg)
    compared, all_same = 0, True
    for l in range(N_LAYERS):
        for i in range(N_SCENES):
            T, G = layers[l]["cls_preds"][i].shape[0], gts[i]["labels"].numel()
            got_m = torch.as_tensor(matches[l][i]).cpu().bool().reshape(T, G)
            if not torch.equal(got_m, ref_match(fx, l, i, T, G)):
                all_same = False
                continue
            compared += 1
            dc = torch.as_tensor(d_cls[l][i]).detach().cpu()
            want_c = torch.as_tensor(fx[f"l{l}_dcls{i}"])
            assert torch.allclose(dc, want_c, rtol=rtol, atol=atol), (l, i, float((dc - want_c).abs().max()))
            want_b = torch.as_tensor(fx[f"l{l}_dbox{i}"])
            db = torch.zeros_like(want_b) if d_box[l][i] is None else torch.as_tensor(d_box[l][i]).detach().cpu()
            assert torch.allclose(db, want_b, rtol=rtol, atol=atol), (l, i, float((db - want_b).abs().max()))
    assert compared >= min_compared, compared
    if all_same:
        ref = float(fx["det_loss"])
        assert abs(float(loss) - ref) < loss_rtol * abs(ref), (float(loss), ref)
    return compared
