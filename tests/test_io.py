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
