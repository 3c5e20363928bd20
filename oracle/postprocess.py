"""Oracle: predict_by_feat / multiclass NMS / superpoint trimming
(reference: unidet3d/unidet3d.py:475-677, unidet3d/criterion.py:180-198), PyTorch CPU."""
import numpy as np
import torch
import torch.nn.functional as F

from . import nms as _nms
from .pool import scatter_mean


def bbox_to_corners(bbox):
    """criterion.py:180-198 ``_bbox_to_loss`` for 6-dim boxes."""
    if bbox.shape[-1] != 6:
        return bbox
    return torch.stack((bbox[..., 0] - bbox[..., 3] / 2, bbox[..., 1] - bbox[..., 4] / 2,
                        bbox[..., 2] - bbox[..., 5] / 2, bbox[..., 0] + bbox[..., 3] / 2,
                        bbox[..., 1] + bbox[..., 4] / 2, bbox[..., 2] + bbox[..., 5] / 2), dim=-1)


def topk_candidates(cls_preds, pred_bboxes, topk_insts):
    """unidet3d.py:504-515."""
    scores = F.softmax(cls_preds, dim=-1)[:, :-1]
    num_classes = scores.shape[1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(len(cls_preds), 1).flatten(0, 1)
    scores, topk_idx = scores.flatten(0, 1).topk(topk_insts, sorted=True)
    labels = labels[topk_idx]
    query = torch.div(topk_idx, num_classes, rounding_mode="floor")
    return pred_bboxes[query], scores, labels, query


def multiclass_nms(bboxes, scores, labels, fast_nms, iou_thr, score_thr=0.0):
    """unidet3d.py:595-650.  Returns (boxes, scores, labels, kept indices into input)."""
    with_yaw = bboxes.shape[1] == 7
    out_b, out_s, out_l, out_i = [], [], [], []
    for c in labels.unique():
        sel = (labels == c).nonzero(as_tuple=True)[0]
        sel = sel[scores[sel] > score_thr]
        if len(sel) == 0:
            continue
        cb, cs, cl = bboxes[sel], scores[sel], labels[sel]
        if with_yaw:
            ids = _nms.nms3d(cb.numpy(), cs.numpy(), iou_thr)
        elif fast_nms:
            cb = torch.cat((cb, torch.zeros_like(cb[:, :1])), dim=1)   # :629-631 (yaw-0 padded copy is returned)
            ids = _nms.nms3d_normal(cb.numpy(), cs.numpy(), iou_thr)
        else:
            ids = _nms.aligned_3d_nms(bbox_to_corners(cb).numpy(), cs.numpy(), cl.numpy(), iou_thr)
        ids = torch.as_tensor(ids, dtype=torch.long)
        out_b.append(cb[ids]), out_s.append(cs[ids]), out_l.append(cl[ids]), out_i.append(sel[ids])
    if out_b:
        return torch.cat(out_b), torch.cat(out_s), torch.cat(out_l), torch.cat(out_i)
    return (bboxes.new_zeros((0, bboxes.shape[1])), bboxes.new_zeros((0,)),
            bboxes.new_zeros((0,)), torch.zeros(0, dtype=torch.long))


def face_distances_inside(points, boxes7):
    """unidet3d.py:652-677 + :566-567: inside[b, p] = min face distance > 0.
    points [N,3], boxes7 [M,7] (x,y,z,dx,dy,dz,yaw) -> bool [M,N]."""
    out = []
    for b in boxes7:
        sh = points - b[None, :3]
        c, s = torch.cos(-b[6]), torch.sin(-b[6])           # rotation_3d_in_axis(shift, -yaw, axis=2)
        rx = sh[:, 0] * c - sh[:, 1] * s
        ry = sh[:, 0] * s + sh[:, 1] * c
        cen = torch.stack((b[0] + rx, b[1] + ry, b[2] + sh[:, 2]), -1)
        d = torch.stack((cen[:, 0] - b[0] + b[3] / 2, b[0] + b[3] / 2 - cen[:, 0],
                         cen[:, 1] - b[1] + b[4] / 2, b[1] + b[4] / 2 - cen[:, 1],
                         cen[:, 2] - b[2] + b[5] / 2, b[2] + b[5] / 2 - cen[:, 2]), -1)
        out.append(d.min(-1).values > 0)
    return torch.stack(out) if out else torch.zeros((0, len(points)), dtype=torch.bool)


def trim_bboxes_by_superpoints(sp_pts_mask, point, bboxes, low_sp_thr, up_sp_thr):
    """unidet3d.py:540-593 -> trimmed boxes [M,6] (centre, size)."""
    if bboxes.shape[1] == 6:
        bboxes = torch.cat((bboxes, torch.zeros_like(bboxes[:, :1])), dim=1)
    inside = face_distances_inside(point, bboxes)                      # [M,N]
    sp = torch.as_tensor(sp_pts_mask, dtype=torch.long)
    n_sp = int(sp.max()) + 1
    sp_inside = scatter_mean(inside.float().t().contiguous(), sp, n_sp).t()   # [M,S]
    inside = inside.clone()
    inside[(sp_inside < low_sp_thr)[:, sp]] = False
    inside[(sp_inside > up_sp_thr)[:, sp]] = True
    out = []
    for m in range(len(bboxes)):
        sel = point[inside[m]]
        if len(sel):
            mx, mn = sel.max(0).values, sel.min(0).values
        else:
            mx = point.new_full((3,), float("-inf")); mn = point.new_full((3,), float("inf"))
        out.append(torch.cat(((mx + mn) / 2, mx - mn)))
    return torch.stack(out) if out else point.new_zeros((0, 6))


def predict_by_feat(cls_preds, pred_bboxes, sp_pts_mask, point, *, topk_insts, fast_nms, iou_thr,
                    use_superpoints, low_sp_thr, up_sp_thr, score_thr=0.0):
    """unidet3d.py:475-538 for one scene -> (boxes, labels, scores)."""
    b, s, l, _ = topk_candidates(cls_preds, pred_bboxes, topk_insts)
    nb, ns, nl, _ = multiclass_nms(b, s, l, fast_nms, iou_thr, score_thr)
    if use_superpoints:
        nb = trim_bboxes_by_superpoints(sp_pts_mask, point, nb, low_sp_thr, up_sp_thr)
    return nb, nl, ns
