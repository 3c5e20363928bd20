"""CPU: the oracle's criterion under torch.autograd against the gradients of the REFERENCE's own criterion.py
(tests/golden/criterion_grad_ref.npz): the function the GPU gradient kernels are compared with is the reference's."""
import torch

from oracle import criterion as oc
import criterion_grad_case as case


def _oracle_run(names, layers, gts, cfg):
    leaves = [dict(cls_preds=[t.clone().requires_grad_(True) for t in lay["cls_preds"]],
                   bboxes=[t.clone().requires_grad_(True) for t in lay["bboxes"]]) for lay in layers]
    pred = dict(leaves[0], aux_outputs=leaves[1:])
    loss = oc.criterion(pred, gts, names, cfg)
    loss.backward()
    matches = []
    for lay in layers:
        ms = []
        for i, g in enumerate(gts):
            T, G = lay["cls_preds"][i].shape[0], g["labels"].numel()
            m = torch.zeros((T, G), dtype=torch.bool)
            if G:
                iq, ig = oc.uni_matcher(lay["cls_preds"][i], lay["bboxes"][i], g["labels"], g["boxes"], g["query_masks"],
                                        cfg["topk"][cfg["datasets"].index(names[i])])
                m[iq, ig] = True
            ms.append(m)
        matches.append(ms)
    return (loss.detach(), [[t.grad for t in lay["cls_preds"]] for lay in leaves], [[t.grad for t in lay["bboxes"]] for lay in leaves],
            matches)


def test_oracle_criterion_gradients_match_the_reference_criterion_under_autograd():
    compared = case.check(_oracle_run, rtol=1e-4, atol=1e-7, min_compared=12)
    assert compared == 12          # every (layer, scene): the oracle's matcher reproduces the reference's matches exactly


def test_fixture_covers_the_interesting_cases():
    fx, names, layers, gts, cfg = case.load()
    assert [g["labels"].numel() for g in gts] == [6, 5, 0, 3] and set(names) == {"scannet", "s3dis"}
    assert not gts[3]["query_masks"][1].any()                                     # a GT no query may match
    assert all(abs(fx[f"l{l}_dbox2"]).max() == 0 for l in range(3))               # the scene without GT has no box gradient
    assert all(abs(fx[f"l{l}_dbox0"]).max() > 0 and abs(fx[f"l{l}_dcls2"]).max() > 0 for l in range(3))
    multi = [(case.ref_match(fx, l, i, layers[l]["cls_preds"][i].shape[0], gts[i]["labels"].numel()).sum(1) > 1).any()
             for l in range(3) for i in (0, 1, 3)]
    print("queries matched to several GTs present:", any(multi))
