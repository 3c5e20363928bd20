"""CPU: the oracle's restatement of the reference evaluator (oracle/evaluate.py) against the fixture produced by the
reference's own indoor_eval (tests/golden/make_golden.py: gen_evaluate)."""
import os

import numpy as np

from oracle import evaluate as oev


def _load(golden_dir):
    g = np.load(os.path.join(golden_dir, "evaluate_ref.npz"))
    ref = {str(k): float(v) for k, v in zip(g["keys"], g["values"])}
    return g, ref


def test_indoor_eval_matches_the_reference(golden_dir):
    g, ref = _load(golden_dir)
    label2cat = {i: f"c{i}" for i in range(7)}
    out = oev.indoor_eval(g["det_boxes"], g["det_scores"], g["det_labels"], g["det_img"], g["gt_boxes"], g["gt_labels"],
                          g["gt_img"], [float(t) for t in g["metric"]], label2cat)
    assert list(out.keys()) == list(ref.keys())          # same classes, same order (dict insertion order of the reference)
    for k in ref:
        if np.isnan(ref[k]):
            assert np.isnan(out[k]), k
        else:
            assert out[k] == ref[k], (k, out[k], ref[k])
    # the fixture exercises the conventions: a class with detections but no ground truth is nan (skipped by nanmean),
    # a class with ground truth but no detection counts as AP 0
    assert np.isnan(ref["c5_AP_0.25"]) and ref["c6_AP_0.25"] == 0.0


def test_overlaps_3d_known_answers():
    a = np.array([[0, 0, 0, 2, 2, 2, 0]], np.float32)
    b = np.array([[1, 0, 0, 2, 2, 2, 0], [0, 0, 0, 2, 2, 2, 0], [5, 5, 5, 1, 1, 1, 0], [0, 0, 1, 2, 2, 2, np.pi / 2]], np.float32)
    iou = oev.overlaps_3d(a, b)[0]
    assert np.allclose(iou, [4 / 12, 1.0, 0.0, 4 / 12], atol=1e-5)
    # a 45-degree square inside: intersection = octagon-free diamond area 2 (side sqrt2) x height 2 -> 4 / (8 + 4 - 4)
    c = np.array([[0, 0, 0, np.sqrt(2), np.sqrt(2), 2, np.pi / 4]], np.float32)
    assert abs(float(oev.overlaps_3d(a, c)[0, 0]) - 4 / 8) < 1e-4
