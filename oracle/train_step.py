"""Oracle: the training loss of the detector END TO END (reference: ``UniDet3D.loss``, unidet3d/unidet3d.py:277-364), as a
differentiable torch-CPU composition of the oracle's pieces -- TEST INFRASTRUCTURE ONLY (tests/, never the product path).

    GT side     per scene: shift = min xyz (:293-299); GT boxes by instance masks (:311-315, get_bboxes_by_masks) or the
                annotated boxes moved by -shift (:316-328); superpoint centres = scatter_mean(xyz - shift) (:330-331);
                query masks = the loader's sp_masks, or get_targets by centre distance (:332-337)
    network     collate / voxelise (:344-346) -> input conv + SpConvUNet + output BN/ReLU in TRAIN mode (batch statistics)
                -> superpoint mean-pool (:348-351) -> queries = all superpoints (:353-354 with len <= query_thr)
                -> encoder with every head (:355) -> criterion (:356)

Pinned by tests/golden/train_step_ref.npz: the reference's own ``UniDet3D.loss`` executed in the build container (over the
dense spconv stand-in and a numpy MinkowskiEngine stand-in) and differentiated with torch.autograd -- loss value and the
gradient of every parameter (tests/test_oracle_golden.py).
"""
import numpy as np
import torch

from . import criterion as oc, encoder as oenc, spconv as ospconv, unet as ounet, voxelize as ovox
from .pool import scatter_mean, superpoint_pool


def gt_of_scene(points, superpoints, spec, target_topk):
    """-> dict(labels, boxes, query_masks) of one scene.  ``spec``: dict(labels, inst, sp_masks) for datasets whose boxes come
    from the instance masks, or dict(labels, boxes) (annotated boxes in the scene's original frame) for datasets whose
    query masks come from the centre distances."""
    xyz = points[:, :3] - points[:, :3].min(0)
    labels = torch.as_tensor(spec["labels"]).long()
    if "inst" in spec:
        return dict(labels=labels, boxes=oc.bboxes_by_masks(spec["inst"], xyz), query_masks=torch.as_tensor(spec["sp_masks"]).bool())
    boxes = torch.as_tensor(np.asarray(spec["boxes"], np.float32)).clone()
    boxes[:, :3] -= torch.as_tensor(points[:, :3].min(0))
    centers = scatter_mean(torch.as_tensor(xyz), torch.as_tensor(superpoints))
    return dict(labels=labels, boxes=boxes, query_masks=oc.targets_by_distance(centers, boxes, target_topk))


def training_loss(sd, cfg, crit_cfg, points, superpoints, datasets_names, gt_specs, target_topk):
    """det_loss of ``UniDet3D.loss`` in train mode.  ``sd``: the detector's state dict (``input_conv.*``, ``unet.*``,
    ``output_layer.*``, ``decoder.*``; tensors may require grad; the running statistics are updated in place like torch);
    ``cfg``: dict(voxel_size, min_spatial_shape, encoder=...) as unidet3d_b200.configs.oracle_cfg builds it."""
    det_sd = {k: t for k, t in sd.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: t for k, t in sd.items() if k.startswith("decoder.")}
    gts = [gt_of_scene(p, s, g, target_topk) for p, s, g in zip(points, superpoints, gt_specs)]
    n_sps = [int(np.asarray(s).max()) + 1 for s in superpoints]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    coords, feats, inverse, shape = ovox.voxelize(points, cfg["voxel_size"], cfg["min_spatial_shape"])
    prev = ospconv.TRAIN_MODE
    ospconv.TRAIN_MODE = True
    try:
        x, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    finally:
        ospconv.TRAIN_MODE = prev
    pooled = superpoint_pool(x, inverse, np.concatenate([np.asarray(s) + o for s, o in zip(superpoints, sp_off[:-1])]), int(sp_off[-1]))
    xs = [pooled[sp_off[i]:sp_off[i + 1]] for i in range(len(points))]
    ctrs = [scatter_mean(torch.as_tensor(p[:, :3] - p[:, :3].min(0)), torch.as_tensor(s)) for p, s in zip(points, superpoints)]
    pred = oenc.encoder_forward(enc_sd, cfg["encoder"], xs, ctrs, datasets_names, all_heads=True)
    return oc.criterion(pred, gts, datasets_names, crit_cfg), gts
