"""CPU: the oracle's restatement of ElasticTransfrom / PointSample_ against fixtures produced by the reference's own
classes (tests/golden/make_golden.py: gen_augment, gen_gt_prep)."""
import os

import numpy as np

from oracle import augment as oaug


def test_elastic_transform_matches_the_reference_bit_for_bit(golden_dir):
    g = np.load(os.path.join(golden_dir, "augment_ref.npz"))
    pts = g["points"]
    applied = 0
    for tag in "abc":
        g0, g1, m0, m1, vs, p, seed = g[f"{tag}_cfg"]
        np.random.seed(int(seed))
        out = oaug.elastic_transform(pts[:, :3], vs, [int(g0), int(g1)], [int(m0), int(m1)], p)
        ref = g[f"{tag}_out"]
        assert out.dtype == ref.dtype and np.array_equal(out, ref), tag
        applied += int(ref.dtype == np.float64)
    assert applied == 2          # case b draws rand() > p: the coordinates pass through as float32


def test_point_sample_matches_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "gt_prep_ref.npz"))
    out = oaug.point_sample(g["ps_choices"], pts_instance_mask=g["ps_in_inst"], pts_semantic_mask=g["sn_pts_semantic_mask"],
                            sp_pts_mask=g["sn_sp_pts_mask"])
    assert np.array_equal(out["pts_instance_mask"], g["ps_out_inst"])
    assert np.array_equal(out["pts_semantic_mask"], g["ps_out_sem"])
    assert np.array_equal(out["sp_pts_mask"], g["ps_out_sp"])
