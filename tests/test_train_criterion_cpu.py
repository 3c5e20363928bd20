"""CPU: the differentiable criterion of unidet3d_b200/train.py (loss of a layer given the match matrix) against
torch.autograd through the oracle's criterion (pinned to the reference's criterion.py): value and gradients w.r.t. the
logits and boxes of every head.  The matcher is the oracle's here (the GPU matcher is checked in the GPU tests)."""
import types

import numpy as np
import torch

from oracle import criterion as ocrit


def test_criterion_backward_matches_the_oracle_autograd():
    from unidet3d_b200 import train
    from unidet3d_b200.criterion import UniDet3DCriterion
    g = torch.Generator().manual_seed(5)
    datasets = ["scannet", "s3dis"]
    crit = UniDet3DCriterion(matcher=dict(costs=[dict(type="QueryClassificationCost", weight=0.5), dict(type="BboxCostJointTraining", weight=2.0)]),
                             loss_weight=[0.5, 1.0], non_object_weight=0.1, iter_matcher=True, bbox_loss_simple=dict(mode="diou"),
                             bbox_loss_rotated=dict(mode="diou"), datasets=datasets, datasets_weights=[1.0, 0.7], topk=[3, 2])
    names = ["scannet", "s3dis", "scannet"]
    Ts, Gs, Cs = [60, 45, 30], [5, 0, 3], [6, 4, 6]
    gts, insts = [], []
    for T, G, C in zip(Ts, Gs, Cs):
        labels = torch.randint(0, C, (G,), generator=g)
        boxes = torch.cat((torch.rand(G, 3, generator=g) * 4, torch.rand(G, 3, generator=g) + 0.3), 1)
        qm = torch.rand(G, T, generator=g) < 0.7
        gts.append(dict(labels=labels, boxes=boxes, query_masks=qm))
        inst = types.SimpleNamespace(labels_3d=labels, query_masks=qm,
                                     bboxes_3d=types.SimpleNamespace(gravity_center=boxes[:, :3], tensor=boxes, with_yaw=False))
        insts.append(inst)

    def head():
        return dict(cls_preds=[torch.randn(T, C + 1, generator=g) for T, C in zip(Ts, Cs)],
                    bboxes=[torch.cat((torch.rand(T, 3, generator=g) * 4, torch.rand(T, 3, generator=g) + 0.2), 1) for T in Ts])

    final, aux = head(), [head(), head()]
    out = dict(cls_preds=final["cls_preds"], bboxes=final["bboxes"], aux_outputs=aux)
    cfg = dict(datasets=datasets, datasets_weights=[1.0, 0.7], topk=[3, 2], loss_weight=[0.5, 1.0], non_object_weight=0.1,
               w_cls=0.5, w_box=2.0, iter_matcher=True)

    def match_fn(cp, pb, boxes, labels, qm, topk):
        iq, ig = ocrit.uni_matcher(cp, pb, labels, boxes, qm, topk, 0.5, 2.0)
        m = torch.zeros((cp.shape[0], labels.numel()), dtype=torch.bool)
        m[iq, ig] = True
        return m

    loss, d_cls, d_box = train.criterion_backward(crit, out, insts, names, match_fn=match_fn)
    # oracle with autograd
    leaves = []

    def req(hd):
        r = dict(cls_preds=[t.clone().requires_grad_(True) for t in hd["cls_preds"]], bboxes=[t.clone().requires_grad_(True) for t in hd["bboxes"]])
        leaves.append(r)
        return r

    aux_r = [req(a) for a in aux]
    fin_r = req(final)
    ref = ocrit.criterion(dict(cls_preds=fin_r["cls_preds"], bboxes=fin_r["bboxes"], aux_outputs=aux_r), gts, names, cfg)
    ref.backward()
    assert abs(float(loss) - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
    for hd, dc, db in zip(leaves, d_cls, d_box):            # aux heads first, final head last
        for t, gt_ in zip(hd["cls_preds"], dc):
            assert torch.allclose(gt_, t.grad, atol=1e-6, rtol=1e-4)
        for t, gb in zip(hd["bboxes"], db):
            want = t.grad if t.grad is not None else torch.zeros_like(t)
            got = gb if gb is not None else torch.zeros_like(t)
            assert torch.allclose(got, want, atol=1e-6, rtol=1e-4)
