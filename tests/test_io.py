"""CPU: the input wire format readers (reference: unidet3d/loading.py, tools/scannet_data_utils.py)."""
import numpy as np

from unidet3d_b200 import io
from unidet3d_b200.synthetic import make_scene


def test_bin_roundtrip_and_color_normalisation(tmp_path):
    pts, sp = make_scene(0, 2000, 2.0, 0.2)
    raw = pts.copy()
    raw[:, 3:] = raw[:, 3:] * 127.5 + 127.5          # back to 0..255 like the files on disk
    raw.astype(np.float32).tofile(tmp_path / "scene.bin")
    sp.astype(np.int64).tofile(tmp_path / "scene_sp.bin")
    p = io.load_points_bin(str(tmp_path / "scene.bin"))
    s = io.load_superpoints_bin(str(tmp_path / "scene_sp.bin"))
    assert p.shape == (2000, 6) and p.dtype == np.float32 and s.dtype == np.int64 and np.array_equal(s, sp)
    n = io.normalize_points_color(p)
    assert np.allclose(n[:, :3], pts[:, :3]) and np.allclose(n[:, 3:], pts[:, 3:], atol=1e-5)
    assert np.array_equal(io.load_points_bin(str(tmp_path / "scene.bin"), use_dim=(0, 1, 2))[:, :3], p[:, :3])


def test_gt_prep_matches_reference_transforms(golden_dir):
    """unidet3d_b200.gt_prep vs the reference's own PointDetClassMappingScanNet / PointDetClassMappingS3DIS /
    PointSample_ (tests/golden/gt_prep_ref.npz): relabelled instance masks, labels and superpoint masks bit-exact."""
    import os
    import numpy as np
    from unidet3d_b200 import gt_prep
    g = np.load(os.path.join(golden_dir, "gt_prep_ref.npz"))
    inst, labels, spm = gt_prep.scannet_gt(g["sn_pts_instance_mask"], g["sn_pts_semantic_mask"], g["sn_sp_pts_mask"], 20, [0, 1])
    assert np.array_equal(inst, g["sn_out_inst"]) and np.array_equal(labels, g["sn_out_labels"])
    assert spm.dtype == bool and np.array_equal(spm, g["sn_out_sp_masks"])
    inst, labels, spm = gt_prep.s3dis_gt(g["s3_pts_instance_mask"], g["s3_pts_semantic_mask"], g["s3_sp_pts_mask"], [7, 8, 9, 10, 11])
    assert np.array_equal(inst, g["s3_out_inst"]) and np.array_equal(labels, g["s3_out_labels"])
    assert np.array_equal(spm, g["s3_out_sp_masks"])
    i2, s2, p2 = gt_prep.point_sample(g["ps_choices"], g["ps_in_inst"], g["sn_pts_semantic_mask"], g["sn_sp_pts_mask"])
    assert np.array_equal(i2, g["ps_out_inst"]) and np.array_equal(s2, g["ps_out_sem"]) and np.array_equal(p2, g["ps_out_sp"])


def test_mask_bin_roundtrip(tmp_path):
    import numpy as np
    from unidet3d_b200 import gt_prep
    m = np.random.default_rng(0).integers(-1, 50, 1000).astype(np.int64)
    p = tmp_path / "instance_mask.bin"
    m.tofile(p)
    assert np.array_equal(gt_prep.load_mask_bin(str(p)), m)
