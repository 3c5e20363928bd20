"""GPU: the evaluator (unidet3d_b200/evaluate.py, csrc/eval.cu) against the fixture of the reference's own indoor_eval and
against the oracle on a larger random result set.  AP is a float32 rounded from a double sum: tolerance 1e-6."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import evaluate as oev

DEV = "cuda"


def _annos(g, n_img):
    gt_annos, dt_annos = [], []
    for i in range(n_img):
        gm, dm = g["gt_img"] == i, g["det_img"] == i
        gt_annos.append(dict(gt_bboxes_3d=torch.as_tensor(g["gt_boxes"][gm]).to(DEV), gt_labels_3d=torch.as_tensor(g["gt_labels"][gm])))
        dt_annos.append(dict(bboxes_3d=torch.as_tensor(g["det_boxes"][dm]).to(DEV), scores_3d=torch.as_tensor(g["det_scores"][dm]).to(DEV),
                             labels_3d=torch.as_tensor(g["det_labels"][dm]).to(DEV)))
    return gt_annos, dt_annos


def _same(out, ref, tol=1e-6):
    assert list(out.keys()) == list(ref.keys())
    for k in ref:
        if np.isnan(ref[k]):
            assert np.isnan(out[k]), k
        else:
            assert abs(out[k] - ref[k]) <= tol * max(1.0, abs(ref[k])), (k, out[k], ref[k])


def test_indoor_eval_vs_reference_fixture(golden_dir):
    from unidet3d_b200.evaluate import indoor_eval
    g = np.load(os.path.join(golden_dir, "evaluate_ref.npz"))
    ref = {str(k): float(v) for k, v in zip(g["keys"], g["values"])}
    gt_annos, dt_annos = _annos(g, 6)
    out = indoor_eval(gt_annos, dt_annos, [float(t) for t in g["metric"]], {i: f"c{i}" for i in range(7)})
    _same(out, ref)


def test_indoor_eval_vs_oracle_large():
    from unidet3d_b200.evaluate import indoor_eval
    rng = np.random.default_rng(8)
    n_img, n_cls = 40, 18
    d = dict(gt_boxes=[], gt_labels=[], gt_img=[], det_boxes=[], det_scores=[], det_labels=[], det_img=[])
    for img in range(n_img):
        m = int(rng.integers(0, 25))
        gb = np.concatenate([rng.uniform(0, 8, (m, 3)), rng.uniform(0.3, 2.0, (m, 3)), np.zeros((m, 1))], 1).astype(np.float32)
        if img % 3 == 0:
            gb[:, 6] = rng.uniform(-3, 3, m)
        gl = rng.integers(0, n_cls - 1, m)
        k = int(rng.integers(0, 400))
        src = rng.integers(0, max(m, 1), k)
        db = (gb[src] if m else np.zeros((k, 7), np.float32)) + rng.normal(0, 0.15, (k, 7)).astype(np.float32)
        db[:, 3:6] = np.abs(db[:, 3:6]) + 0.05
        if img % 3:
            db[:, 6] = 0
        dl = gl[src] if m else rng.integers(0, n_cls, k)
        flip = rng.random(k) < 0.3
        dl = np.where(flip, rng.integers(0, n_cls, k), dl)
        ds = rng.random(k).astype(np.float32)            # distinct with probability ~1: the order is well defined
        d["gt_boxes"].append(gb); d["gt_labels"].append(gl); d["gt_img"].append(np.full(m, img))
        d["det_boxes"].append(db.astype(np.float32)); d["det_scores"].append(ds); d["det_labels"].append(dl); d["det_img"].append(np.full(k, img))
    g = {k: np.concatenate(v) for k, v in d.items()}
    label2cat = {i: f"c{i}" for i in range(n_cls)}
    ref = oev.indoor_eval(g["det_boxes"], g["det_scores"], g["det_labels"], g["det_img"], g["gt_boxes"], g["gt_labels"], g["gt_img"],
                          [0.25, 0.5, 0.7], label2cat)
    gt_annos, dt_annos = _annos(g, n_img)
    out = indoor_eval(gt_annos, dt_annos, [0.25, 0.5, 0.7], label2cat)
    _same(out, ref, tol=2e-6)
    assert 0.0 < ref["mAP_0.25"] < 1.0 and ref["mAP_0.70"] < ref["mAP_0.25"]
