"""Oracle (test infrastructure only): CPU restatement of the two training-pipeline transforms in front of the hot path.

* ``elastic_transform`` follows ``ElasticTransfrom.transform`` / ``.elastic`` (reference unidet3d/transforms_3d.py:12-83)
  call for call -- the same numpy / scipy functions in the same order, so that with the same ``np.random`` state it
  reproduces the reference bit for bit (pinned by tests/golden/augment_ref.npz, generated from the reference class).
* ``point_sample`` follows ``PointSample_.transform`` (transforms_3d.py:233-295) given the sampled ``choices``.
Parity: pinned against the reference's own classes executed in the build container (tests/golden/make_golden.py).
"""
import numpy as np
import scipy.interpolate
import scipy.ndimage


def elastic_noise_dims(x, gran):
    """transforms_3d.py:64."""
    return np.abs(x).max(0).astype(np.int32) // gran + 3


def draw_noise(noise_dim, rng=np.random):
    """transforms_3d.py:65-68: three float32 standard-normal grids, drawn in this order."""
    return [rng.randn(noise_dim[0], noise_dim[1], noise_dim[2]).astype('float32') for _ in range(3)]


def blur_noise(noise):
    """transforms_3d.py:60-62,70-74."""
    blur0 = np.ones((3, 1, 1)).astype('float32') / 3
    blur1 = np.ones((1, 3, 1)).astype('float32') / 3
    blur2 = np.ones((1, 1, 3)).astype('float32') / 3
    for blur in [blur0, blur1, blur2, blur0, blur1, blur2]:
        noise = [scipy.ndimage.convolve(n, blur, mode='constant', cval=0) for n in noise]
    return noise


def elastic(x, gran, mag, noise=None, rng=np.random):
    """transforms_3d.py:46-83.  ``noise``: the three raw (unblurred) grids, drawn here when None."""
    noise_dim = elastic_noise_dims(x, gran)
    if noise is None:
        noise = draw_noise(noise_dim, rng)
    noise = blur_noise(noise)
    ax = [np.linspace(-(b - 1) * gran, (b - 1) * gran, b) for b in noise_dim]
    interp = [scipy.interpolate.RegularGridInterpolator(ax, n, bounds_error=0, fill_value=0) for n in noise]
    return x + np.hstack([i(x)[:, None] for i in interp]) * mag


def elastic_transform(points_xyz, voxel_size, gran, mag, p=1.0, rng=np.random):
    """transforms_3d.py:28-44: -> elastic_coords (float64 when applied, float32 otherwise)."""
    coords = np.asarray(points_xyz, dtype=np.float32) / float(voxel_size)     # python float: the quotient stays float32
    if rng.rand() < p:
        coords = elastic(coords, gran[0], mag[0], rng=rng)
        coords = elastic(coords, gran[1], mag[1], rng=rng)
    return coords


def point_sample(choices, pts_instance_mask=None, pts_semantic_mask=None, sp_pts_mask=None):
    """transforms_3d.py:262-295 after ``choices`` are drawn."""
    out = {}
    if pts_instance_mask is not None:
        m = pts_instance_mask[choices]
        idxs = np.unique(m)
        mapping = np.zeros(np.max(idxs) + 2, dtype=int)
        new_idxs = np.arange(len(idxs))
        mapping[idxs] = new_idxs - 1 if idxs[0] == -1 else new_idxs
        out['pts_instance_mask'] = mapping[m]
    if pts_semantic_mask is not None:
        out['pts_semantic_mask'] = pts_semantic_mask[choices]
    if sp_pts_mask is not None:
        out['sp_pts_mask'] = np.unique(sp_pts_mask[choices], return_inverse=True)[1]
    return out
