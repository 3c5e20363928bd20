"""Oracle: training-side targets, matcher and loss VALUES -- torch-CPU fp32.

Restates, for the forward value only (no gradients):

* ``UniDet3D.get_bboxes_by_masks`` / ``get_gt_inst_masks``    unidet3d/unidet3d.py:220-275
* ``UniDet3D.get_targets``                                    unidet3d/unidet3d.py:371-409
* ``QueryClassificationCost``, ``BboxCostJointTraining``      unidet3d/criterion.py:200-284
* ``UniMatcher``                                              unidet3d/criterion.py:286-320
* ``UniDet3DCriterion.get_layer_loss`` / ``__call__``         unidet3d/criterion.py:44-178
* ``axis_aligned_diou_loss``                                  unidet3d/axis_aligned_iou_loss.py:14-53
* ``diff_diou_rotated_3d`` / ``rotated_diou_3d_loss``         unidet3d/rotated_iou_loss.py:14-82

Pinned by tests/golden/criterion_ref.npz: the reference's own criterion.py / axis_aligned_iou_loss.py /
rotated_iou_loss.py / unidet3d.py functions executed in the build container (tests/golden/make_golden.py).
Third-party pieces restated from their published definitions: mmdet3d ``AxisAlignedBboxOverlaps3D``
(is_aligned, eps 1e-6), mmdet ``weighted_loss`` (reduction 'none' = identity), mmcv ``box2corners``; the rotated
rectangle intersection area (mmcv ``oriented_box_intersection_2d``, exact up to 1e-6) is the oracle's vertex-sorting
``box_overlap_rotated`` (oracle/nms.py) with its corner tolerance set to 1e-6, checked through the golden fixture against
an independent float64 Sutherland-Hodgman clipper.

Reference quirks that are kept because they change the numbers:
* axis-aligned DIoU used as a MATCHING COST on [T, G, 6] inputs adds ``(r2 / c2)[:, 0]`` -- the centre-distance penalty
  of GT 0 -- to every column (axis_aligned_iou_loss.py:51); as a LOSS on [N, 6] inputs it is the per-pair penalty;
* rotated DIoU takes ``r2`` over (x, y, w) of the BEV boxes, not (x, y, z) (rotated_iou_loss.py:24-25,61);
* a query matched to several GTs gets the label of the LARGEST GT index (criterion.py:96, CPU index_put order).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .nms import box_overlap_rotated

INF = 1e8


# ----------------------------------------------------------------------------- GT preparation
def bboxes_by_masks(inst_ids, points):
    """unidet3d.py:220-275: instance ids int64 [n] (-1 = no instance) + points [n,3] -> boxes [n_inst, 6]
    (centre, size) = tight AABB of every instance's points (ids 0..max)."""
    inst_ids = torch.as_tensor(inst_ids, dtype=torch.long)
    points = torch.as_tensor(points, dtype=torch.float32)
    n_inst = int(inst_ids.max()) + 1 if (inst_ids >= 0).any() else 0
    boxes = []
    for i in range(n_inst):
        p = points[inst_ids == i]
        lo, hi = p.min(0).values, p.max(0).values
        boxes.append(torch.cat(((hi + lo) / 2, hi - lo)))
    return torch.stack(boxes) if boxes else points.new_zeros((0, 6))


def targets_by_distance(sp_centers, gt_boxes, topk):
    """unidet3d.py:371-409: superpoint centres [S,3], boxes [G, 6|7] (gravity centre first) -> bool [G, S]:
    every box keeps its ``topk`` nearest centres (strictly below the (topk+1)-th distance), every centre goes to
    the nearest box among those that kept it."""
    pts = torch.as_tensor(sp_centers, dtype=torch.float32)
    boxes = torch.as_tensor(gt_boxes, dtype=torch.float32)
    S, G = len(pts), len(boxes)
    d = ((boxes[None, :, :3] - pts[:, None, :]) ** 2).sum(-1)                   # [S, G]
    kth = torch.topk(d, min(topk + 1, S), largest=False, dim=0).values[-1]
    d = torch.where(d < kth[None], d, torch.tensor(INF))
    mv, mi = d.min(1)
    mi = torch.where(mv < INF, mi, torch.tensor(G))
    return F.one_hot(mi, G + 1)[:, :-1].bool().T


# ----------------------------------------------------------------------------- box losses
def bbox_to_loss(b):
    """criterion.py:180-198."""
    if b.shape[-1] != 6:
        return b
    return torch.stack((b[..., 0] - b[..., 3] / 2, b[..., 1] - b[..., 4] / 2, b[..., 2] - b[..., 5] / 2,
                        b[..., 0] + b[..., 3] / 2, b[..., 1] + b[..., 4] / 2, b[..., 2] + b[..., 5] / 2), -1)


def _aligned_iou(p, t, eps=1e-6):
    """mmdet3d AxisAlignedBboxOverlaps3D(is_aligned=True) on corner boxes [..., 6]."""
    a1 = (p[..., 3] - p[..., 0]) * (p[..., 4] - p[..., 1]) * (p[..., 5] - p[..., 2])
    a2 = (t[..., 3] - t[..., 0]) * (t[..., 4] - t[..., 1]) * (t[..., 5] - t[..., 2])
    wh = (torch.minimum(p[..., 3:], t[..., 3:]) - torch.maximum(p[..., :3], t[..., :3])).clamp(min=0)
    ov = wh[..., 0] * wh[..., 1] * wh[..., 2]
    union = torch.clamp(a1 + a2 - ov, min=eps)
    return ov / union


def aligned_diou_loss(pred, target):
    """axis_aligned_iou_loss.py:14-53 on corner boxes [..., 6]; keeps the ``[:, 0]`` of line 51."""
    iou_loss = 1 - _aligned_iou(pred, target)
    pc = (pred[..., :3] + pred[..., 3:]) / 2
    tc = (target[..., :3] + target[..., 3:]) / 2
    r2 = ((pc - tc) ** 2).sum(-1, keepdim=True)
    lo = torch.minimum(pred[..., :3], target[..., :3])
    hi = torch.maximum(pred[..., 3:], target[..., 3:])
    c2 = ((lo - hi) ** 2).sum(-1, keepdim=True)
    return iou_loss + (r2 / c2)[:, 0]


def _corners_xy(b5):
    """mmcv box2corners: (x, y, w, h, alpha) [..., 5] -> x, y [..., 4]."""
    x4 = torch.tensor([0.5, -0.5, -0.5, 0.5]) * b5[..., 2:3]
    y4 = torch.tensor([0.5, 0.5, -0.5, -0.5]) * b5[..., 3:4]
    s, c = torch.sin(b5[..., 4:5]), torch.cos(b5[..., 4:5])
    return x4 * c - y4 * s + b5[..., 0:1], x4 * s + y4 * c + b5[..., 1:2]


def rotated_diou_loss(pred, target):
    """rotated_iou_loss.py:14-82 on (x,y,z,w,h,l,alpha) boxes [..., 7] -> [...]."""
    shape = pred.shape[:-1]
    p, t = pred.reshape(-1, 7).float(), target.reshape(-1, 7).float()
    inter = torch.as_tensor(box_overlap_rotated(p.numpy(), t.numpy(), margin=1e-6))
    zmax1, zmin1 = p[:, 2] + p[:, 5] * 0.5, p[:, 2] - p[:, 5] * 0.5
    zmax2, zmin2 = t[:, 2] + t[:, 5] * 0.5, t[:, 2] - t[:, 5] * 0.5
    z_ov = (torch.minimum(zmax1, zmax2) - torch.maximum(zmin1, zmin2)).clamp(min=0)
    inter3 = inter * z_ov
    union3 = p[:, 3] * p[:, 4] * p[:, 5] + t[:, 3] * t[:, 4] * t[:, 5] - inter3
    b1, b2 = p[:, [0, 1, 3, 4, 6]], t[:, [0, 1, 3, 4, 6]]
    x1, y1 = _corners_xy(b1)
    x2, y2 = _corners_xy(b2)
    x_max = torch.maximum(x1.max(1).values, x2.max(1).values)
    x_min = torch.minimum(x1.min(1).values, x2.min(1).values)
    y_max = torch.maximum(y1.max(1).values, y2.max(1).values)
    y_min = torch.minimum(y1.min(1).values, y2.min(1).values)
    z_max, z_min = torch.maximum(zmax1, zmax2), torch.minimum(zmin1, zmin2)
    r2 = ((b1[:, :3] - b2[:, :3]) ** 2).sum(-1)
    c2 = (x_min - x_max) ** 2 + (y_min - y_max) ** 2 + (z_min - z_max) ** 2
    return (1 - (inter3 / union3 - r2 / c2)).reshape(shape)


def box_loss(pred, target):
    """per-pair DIoU loss of (centre, size[, yaw]) boxes [N, 6|7] (criterion.py:127-134)."""
    if target.shape[-1] == 7:
        return rotated_diou_loss(pred, target)
    return aligned_diou_loss(bbox_to_loss(pred), bbox_to_loss(target))


# ----------------------------------------------------------------------------- matcher
def match_cost(cls_pred, pred_boxes, gt_labels, gt_boxes, w_cls=0.5, w_box=2.0):
    """criterion.py:200-284: [T, G] = -w_cls * softmax(cls)[:, labels] + w_box * DIoU(pred_q, gt_g)."""
    T, G = len(cls_pred), len(gt_labels)
    cost_cls = -cls_pred.softmax(-1)[:, gt_labels] * w_cls
    pb = pred_boxes.unsqueeze(1).repeat(1, G, 1)
    gb = gt_boxes.unsqueeze(0).repeat(T, 1, 1)
    if gt_boxes.shape[1] == 7:
        cost_box = rotated_diou_loss(pb, gb)
    else:
        cost_box = aligned_diou_loss(bbox_to_loss(pb), bbox_to_loss(gb))
    return cost_cls + cost_box * w_box


def uni_matcher(cls_pred, pred_boxes, gt_labels, gt_boxes, query_masks, topk, w_cls=0.5, w_box=2.0):
    """criterion.py:286-320 -> (query ids, gt ids), argwhere order (query-major)."""
    gt_labels = torch.as_tensor(gt_labels, dtype=torch.long)
    if len(gt_labels) == 0:
        return gt_labels.new_empty((0,)), gt_labels.new_empty((0,))
    cost = match_cost(cls_pred, pred_boxes, gt_labels, gt_boxes, w_cls, w_box)
    cost = torch.where(torch.as_tensor(query_masks).bool().T, cost, torch.tensor(INF))
    kth = torch.topk(cost, topk + 1, dim=0, sorted=True, largest=False).values[-1:, :]
    ids = torch.argwhere(cost < kth)
    return ids[:, 0], ids[:, 1]


# ----------------------------------------------------------------------------- criterion
def layer_loss(cls_preds, pred_boxes, gts, datasets_names, cfg, indices=None):
    """criterion.py:44-142.  gts: list of dicts(labels int64 [G], boxes [G, 6|7], query_masks bool [G, T]);
    cfg: dict(datasets, datasets_weights, topk, loss_weight, non_object_weight, w_cls, w_box).
    Returns (loss, indices)."""
    if indices is None:
        indices = []
        for i, g in enumerate(gts):
            idx = cfg["datasets"].index(datasets_names[i])
            indices.append(uni_matcher(cls_preds[i], pred_boxes[i], g["labels"], g["boxes"], g["query_masks"],
                                       cfg["topk"][idx], cfg.get("w_cls", 0.5), cfg.get("w_box", 2.0)))
    cls_losses, bbox_losses = [], []
    for name, cp, pb, g, (iq, ig) in zip(datasets_names, cls_preds, pred_boxes, gts, indices):
        w = cfg["datasets_weights"][cfg["datasets"].index(name)]
        C = cp.shape[1] - 1
        tgt = torch.full((len(cp),), C, dtype=torch.long)
        tgt[iq] = torch.as_tensor(g["labels"], dtype=torch.long)[ig]
        cw = torch.tensor([1.0] * C + [cfg["non_object_weight"]])
        cls_losses.append(w * F.cross_entropy(cp, tgt, cw))
        if len(g["labels"]) == 0 or len(iq) == 0:
            continue
        bbox_losses.append(w * box_loss(pb[iq], torch.as_tensor(g["boxes"], dtype=torch.float32)[ig]).mean())
    cls_loss = torch.stack(cls_losses).mean()
    bbox_loss = torch.stack(bbox_losses).mean() if bbox_losses else torch.tensor(0.0)
    return cfg["loss_weight"][0] * cls_loss + cfg["loss_weight"][1] * bbox_loss, indices


def criterion(pred, gts, datasets_names, cfg):
    """criterion.py:144-178 -> det_loss (iter_matcher: every layer is matched on its own predictions)."""
    loss, indices = layer_loss(pred["cls_preds"], pred["bboxes"], gts, datasets_names, cfg)
    for aux in pred.get("aux_outputs", []):
        l, _ = layer_loss(aux["cls_preds"], aux["bboxes"], gts, datasets_names, cfg,
                          None if cfg.get("iter_matcher", True) else indices)
        loss = loss + l
    return loss
