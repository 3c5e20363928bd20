"""GPU: ElasticTransfrom / PointSample_ (unidet3d_b200/augment.py, csrc/augment.cu) against the reference fixtures and
the oracle.  Elastic coordinates are doubles like the reference's: tolerance 1e-9 voxel units (rounding order only)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import augment as oaug

DEV = "cuda"


def test_elastic_transform_vs_reference_fixture(golden_dir):
    from unidet3d_b200.augment import ElasticTransfrom
    g = np.load(os.path.join(golden_dir, "augment_ref.npz"))
    pts = torch.as_tensor(g["points"]).to(DEV)
    for tag in "abc":
        g0, g1, m0, m1, vs, p, seed = g[f"{tag}_cfg"]
        np.random.seed(int(seed))
        t = ElasticTransfrom(gran=[int(g0), int(g1)], mag=[int(m0), int(m1)], voxel_size=float(vs), p=float(p))
        out = t.transform(dict(points=pts))["elastic_coords"].cpu().numpy()
        ref = g[f"{tag}_out"].astype(np.float64)
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= 1e-9, (tag, np.abs(out - ref).max())


def test_elastic_blur_is_bit_exact_and_out_of_grid_points_stay():
    from unidet3d_b200 import augment
    rng = np.random.RandomState(5)
    noise = [rng.randn(9, 14, 6).astype("float32") for _ in range(3)]
    ref = np.stack(oaug.blur_noise([n.copy() for n in noise]))
    out = augment.elastic_blur(torch.from_numpy(np.stack(noise)).to(DEV)).cpu().numpy()
    assert np.array_equal(out, ref)
    # points outside the noise grid (fill_value 0) are not displaced; a point exactly on the last grid node is inside
    gran, mag = 6.0, 40.0
    x = np.array([[48.0, 10.0, 3.0], [48.0001, 10.0, 3.0], [-48.0, -78.0, -30.0], [0.0, 0.0, 31.0]], dtype=np.float64)
    import scipy.interpolate
    ax = [np.linspace(-(b - 1) * gran, (b - 1) * gran, b) for b in (9, 14, 6)]
    want = x + np.hstack([scipy.interpolate.RegularGridInterpolator(ax, n, bounds_error=0, fill_value=0)(x)[:, None] for n in ref]) * mag
    got = augment.elastic_apply(torch.from_numpy(x).to(DEV), torch.from_numpy(ref).to(DEV), gran, mag).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-12
    assert np.array_equal(got[1], x[1]) and np.array_equal(got[3], x[3])


def test_point_sample_vs_reference_fixture_and_oracle(golden_dir):
    from unidet3d_b200.augment import PointSample_, compact_ids
    g = np.load(os.path.join(golden_dir, "gt_prep_ref.npz"))
    choices = g["ps_choices"]
    ps = PointSample_(num_points=len(choices))
    ps._choices = lambda n: choices
    d = dict(points=torch.zeros((len(g["ps_in_inst"]), 6), device=DEV), pts_instance_mask=torch.as_tensor(g["ps_in_inst"]).to(DEV),
             pts_semantic_mask=torch.as_tensor(g["sn_pts_semantic_mask"]).to(DEV), sp_pts_mask=torch.as_tensor(g["sn_sp_pts_mask"]).to(DEV))
    out = ps.transform(d)
    assert np.array_equal(out["pts_instance_mask"].cpu().numpy(), g["ps_out_inst"])
    assert np.array_equal(out["pts_semantic_mask"].cpu().numpy(), g["ps_out_sem"])
    assert np.array_equal(out["sp_pts_mask"].cpu().numpy(), g["ps_out_sp"])
    assert out["points"].shape[0] == len(choices)
    # random ids incl. -1, sampling with replacement (the reference's np.random.choice default), large id range
    rng = np.random.default_rng(4)
    n = 200000
    inst = rng.integers(-1, 300, n)
    sp = rng.integers(0, 70000, n) * 3
    ch = rng.integers(0, n, 150000)
    want = oaug.point_sample(ch, pts_instance_mask=inst, sp_pts_mask=sp)
    gi, ni = compact_ids(torch.as_tensor(inst[ch]).to(DEV), int(inst.max()))
    gs, ns = compact_ids(torch.as_tensor(sp[ch]).to(DEV), int(sp.max()))
    assert np.array_equal(gi.cpu().numpy(), want["pts_instance_mask"])
    assert np.array_equal(gs.cpu().numpy(), want["sp_pts_mask"])
    assert int(ni) == len(np.unique(inst[ch][inst[ch] >= 0])) and int(ns) == len(np.unique(sp[ch]))
    # no ids at all / only negative ids
    e, ne = compact_ids(torch.full((5,), -1, dtype=torch.int64, device=DEV), 0)
    assert e.cpu().tolist() == [-1] * 5 and int(ne) == 0
