"""Oracle: the whole forward hot path for a batch of scenes (reference:
unidet3d/unidet3d.py:411-538 ``predict`` + ``predict_by_feat``), CPU.  This is also the CPU
baseline that bench.py times (``cpu_baseline`` / ``--impl reference``): the reference's own native
stack (spconv / MinkowskiEngine / torch_scatter / mmcv) cannot be installed offline, so the same
math runs as gather -> torch.mm -> index_add_ per kernel offset, F.batch_norm, explicit
multi-head attention and greedy NMS on all host cores."""
import numpy as np
import torch

from . import voxelize as ovox, unet as ounet, encoder as oenc, postprocess as opost
from .pool import scatter_mean, superpoint_pool
from .spconv import bn_relu


def forward_scenes(det_sd, enc_sd, cfg, points, superpoints, datasets_names, stages=None):
    """det_sd: detector-layout backbone state_dict (input_conv / unet.* / output_layer);
    enc_sd: encoder state_dict; cfg: dict(voxel_size, min_spatial_shape, encoder=dict(...),
    test_cfg=dict(...), fast_nms=[...], use_superpoints=[...]).
    Returns list of (boxes, labels, scores); fills ``stages`` (dict) with intermediates if given."""
    pts = [np.asarray(p, np.float32) for p in points]
    sps = [np.asarray(s, np.int64) for s in superpoints]
    n_sps = [int(s.max()) + 1 for s in sps]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    sp_centers = [scatter_mean(torch.as_tensor(p[:, :3]), torch.as_tensor(s)) for p, s in zip(pts, sps)]
    coords, feats, inverse, shape = ovox.voxelize(pts, cfg["voxel_size"], cfg["min_spatial_shape"])
    x, levels = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    sp_all = np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])])
    pooled = superpoint_pool(x, inverse, sp_all, int(sp_off[-1]))
    xs = [pooled[sp_off[i]:sp_off[i + 1]] for i in range(len(pts))]
    out = oenc.encoder_forward(enc_sd, cfg["encoder"], xs, sp_centers, datasets_names, all_heads=False)
    if stages is not None:
        stages.update(coords=coords, vox_feats=feats, inverse=inverse, shape=shape, backbone=x, pooled=pooled,
                      sp_centers=sp_centers, cls_preds=out["cls_preds"], bboxes=out["bboxes"], levels=levels)
    tc = cfg["test_cfg"]
    results = []
    for i, name in enumerate(datasets_names):
        ds = cfg["encoder"]["datasets"].index(name)
        results.append(opost.predict_by_feat(out["cls_preds"][i], out["bboxes"][i], torch.as_tensor(sps[i]),
                                             torch.as_tensor(pts[i][:, :3]), topk_insts=tc["topk_insts"],
                                             fast_nms=cfg["fast_nms"][ds], iou_thr=tc["iou_thr"][ds],
                                             use_superpoints=cfg["use_superpoints"][ds], low_sp_thr=tc["low_sp_thr"],
                                             up_sp_thr=tc["up_sp_thr"], score_thr=tc["score_thr"]))
    return results
