"""Host-side ground-truth preparation of the training pipeline (the step BEFORE
``UniDet3D.loss``).  Plain numpy mirrors of the reference transforms that turn the on-disk masks into what the loss
consumes; integer / boolean work, bit-exact against the reference classes (tests/golden/gt_prep_ref.npz).

Wire format (mmdet3d ``LoadAnnotations3D`` + unidet3d/loading.py:23-52): ``instance_mask/<scene>.bin`` and
``semantic_mask/<scene>.bin`` int64 [N], ``super_points/<scene>.bin`` int64 [N].

* ``scannet_gt``   -- ``PointDetClassMappingScanNet.transform``   unidet3d/transforms_3d.py:146-228
* ``s3dis_gt``     -- ``PointDetClassMappingS3DIS.transform``     unidet3d/transforms_3d.py:85-145
* ``point_sample`` -- ``PointSample_.transform`` (re-indexing after the random choice)  unidet3d/transforms_3d.py:230-295
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def load_mask_bin(path: str) -> np.ndarray:
    """int64 [N] per-point mask file (instance / semantic / superpoint ids)."""
    return np.fromfile(path, dtype=np.int64)


def _superpoint_majority(inst_onehot: np.ndarray, sp: np.ndarray) -> np.ndarray:
    """``scatter_mean(inst_onehot.float(), sp, dim=-1) > 0.5``: bool [G, S], S = max(sp) + 1 -- instance g owns
    superpoint s when more than half of the superpoint's points belong to g."""
    G = inst_onehot.shape[0]
    S = int(sp.max()) + 1 if sp.size else 0
    cnt = np.bincount(sp, minlength=S).astype(np.float32)
    out = np.zeros((G, S), dtype=bool)
    for g in range(G):
        inside = np.bincount(sp, weights=inst_onehot[g].astype(np.float32), minlength=S).astype(np.float32)
        out[g] = inside / np.maximum(cnt, 1.0) > 0.5
    return out


def scannet_gt(pts_instance_mask, pts_semantic_mask, sp_pts_mask, num_classes: int, stuff_classes: Sequence[int]):
    """-> (instance mask relabelled to -1 / 0..G-1, gt_labels int64 [G], gt_sp_masks bool [G, S]).

    Points of the unlabelled class (== num_classes) and of the stuff classes lose their instance (-1); the remaining
    instance ids are compacted in ascending order; label = semantic class of the instance's first point minus the
    number of stuff classes."""
    inst = np.asarray(pts_instance_mask, np.int64).copy()
    sem = np.asarray(pts_semantic_mask, np.int64)
    sp = np.asarray(sp_pts_mask, np.int64)
    inst[sem == num_classes] = -1
    for c in stuff_classes:
        inst[sem == c] = -1
    idxs = np.unique(inst)
    if idxs[0] != -1:
        raise ValueError("scannet_gt: the reference asserts that at least one point has no instance")
    mapping = np.zeros(int(idxs.max()) + 2, dtype=np.int64)
    mapping[idxs] = np.arange(len(idxs)) - 1           # (idxs == -1 writes the last slot, exactly like the reference)
    inst = mapping[inst]
    G = len(idxs) - 1
    onehot = inst[None, :] == np.arange(G)[:, None]
    sp_masks = _superpoint_majority(onehot, sp) if G else np.zeros((0, int(sp.max()) + 1), dtype=bool)
    labels = np.zeros(G, dtype=np.int64)
    for g in range(G):
        labels[g] = sem[inst == g][0] - len(stuff_classes)
    return inst, labels, sp_masks


def s3dis_gt(pts_instance_mask, pts_semantic_mask, sp_pts_mask, classes: Sequence[int]):
    """-> (instance mask -1 / 0..G-1 over the kept instances, gt_labels int64 [G], gt_sp_masks bool [G, S]).

    Instances whose class (semantic id of their first point) is not in ``classes`` are dropped; labels are the
    positions in ``classes``."""
    inst = np.asarray(pts_instance_mask, np.int64).copy()
    sem = np.asarray(pts_semantic_mask, np.int64)
    sp = np.asarray(sp_pts_mask, np.int64)
    if np.unique(inst)[0] == 1:
        inst -= 1
    idxs = np.unique(inst)
    labels = np.array([sem[inst == i][0] for i in idxs], dtype=np.int64)
    keep = np.isin(labels, np.asarray(classes))
    n_oh = int(inst.max()) + 1
    onehot = (inst[None, :] == np.arange(n_oh)[:, None])[idxs[keep]] if keep.any() else np.zeros((0, inst.size), bool)
    # (one_hot(...).T[mask]: row r of the one-hot matrix is instance id r; the reference indexes it with a mask over the
    #  SORTED UNIQUE ids, which coincides with the ids when they are 0..n-1 -- the S3DIS files' convention)
    labels = labels[keep]
    mapping = np.zeros(max(classes) + 1, dtype=np.int64)
    for j, c in enumerate(classes):
        mapping[c] = j
    labels = mapping[labels]
    sp_masks = _superpoint_majority(onehot, sp) if len(labels) else np.zeros((0, int(sp.max()) + 1), dtype=bool)
    new_inst = onehot.argmax(0) if len(labels) else np.zeros(inst.size, np.int64)
    new_inst = np.where(onehot.sum(0) == 0, -1, new_inst) if len(labels) else np.full(inst.size, -1, np.int64)
    return new_inst.astype(np.int64), labels, sp_masks


def point_sample(choices: np.ndarray, pts_instance_mask: Optional[np.ndarray] = None,
                 pts_semantic_mask: Optional[np.ndarray] = None, sp_pts_mask: Optional[np.ndarray] = None
                 ) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], Optional[np.ndarray]]:
    """Masks after sub-sampling the points with ``choices``: instance ids re-compacted (keeping -1), superpoint ids
    re-compacted to 0..S'-1."""
    inst = sem = sp = None
    if pts_instance_mask is not None:
        inst = np.asarray(pts_instance_mask)[choices]
        idxs = np.unique(inst)
        mapping = np.zeros(int(idxs.max()) + 2, dtype=int)
        new = np.arange(len(idxs))
        mapping[idxs] = new - 1 if idxs[0] == -1 else new
        inst = mapping[inst]
    if pts_semantic_mask is not None:
        sem = np.asarray(pts_semantic_mask)[choices]
    if sp_pts_mask is not None:
        sp = np.unique(np.asarray(sp_pts_mask)[choices], return_inverse=True)[1]
    return inst, sem, sp
