"""Oracle: 3D NMS variants used by unidet3d/unidet3d.py:595-650 (numpy fp32).

Restates (third-party, absent -- SURVEY.md appendix A4):
* ``mmcv.ops.nms3d``        (mmcv @780ffed, iou3d kernels): rotated BEV IoU, EPS 1e-8,
  ``S_inter`` = polygon intersection of the two rotated rectangles (16 edge-edge
  crossings + corners-inside tests with MARGIN 1e-2, bubble-sorted by atan2 angle
  about the centroid, shoelace area); box j suppressed by kept i<j iff IoU > thr.
* ``mmcv.ops.nms3d_normal``: same with the axis-aligned BEV overlap (yaw ignored).
* ``mmdet3d...aligned_3d_nms`` (mmdet3d 1.4.0): greedy true-3D IoU on corner boxes,
  class-aware, survivors are ``iou <= thr``.
All return indices into the input, in descending-score order.
"""
import numpy as np

F = np.float32
EPS = F(1e-8)
MARGIN = F(1e-2)


def _cross3(p1x, p1y, p2x, p2y, p0x, p0y):
    return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y)


def _corners(box):
    """box [...,7] -> corners x,y arrays [...,5] (closed polygon)."""
    x, y, dx, dy, ang = box[..., 0], box[..., 1], box[..., 3], box[..., 4], box[..., 6]
    hx, hy = dx / F(2), dy / F(2)
    cx = np.stack([x - hx, x + hx, x + hx, x - hx], -1)
    cy = np.stack([y - hy, y - hy, y + hy, y + hy], -1)
    c, s = np.cos(ang).astype(F)[..., None], np.sin(ang).astype(F)[..., None]
    rx = (cx - x[..., None]) * c + (cy - y[..., None]) * (-s) + x[..., None]
    ry = (cx - x[..., None]) * s + (cy - y[..., None]) * c + y[..., None]
    rx = np.concatenate([rx, rx[..., :1]], -1)
    ry = np.concatenate([ry, ry[..., :1]], -1)
    return rx.astype(F), ry.astype(F)


def _in_box(box, px, py, margin=MARGIN):
    cx, cy = box[..., 0], box[..., 1]
    c, s = np.cos(-box[..., 6]).astype(F), np.sin(-box[..., 6]).astype(F)
    rx = (px - cx) * c + (py - cy) * (-s)
    ry = (px - cx) * s + (py - cy) * c
    return (np.abs(rx) < box[..., 3] / F(2) + F(margin)) & (np.abs(ry) < box[..., 4] / F(2) + F(margin))


def box_overlap_rotated(a, b, margin=MARGIN):
    """a, b float32 [P,7] -> intersection area [P] (mmcv iou3d ``box_overlap``; ``margin`` = its corner-in-box
    tolerance MARGIN = 1e-2, which over-estimates the area by up to ~1 %.  oracle/criterion.py passes 1e-6, the
    tolerance of mmcv's exact ``diff_iou_rotated`` intersection that the reference's rotated DIoU loss is built on)."""
    a, b = a.astype(F), b.astype(F)
    P = len(a)
    ax, ay = _corners(a)
    bx, by = _corners(b)
    px, py, valid = [], [], []
    with np.errstate(all="ignore"):
        for i in range(4):
            for j in range(4):
                p1x, p1y, p0x, p0y = ax[:, i + 1], ay[:, i + 1], ax[:, i], ay[:, i]
                q1x, q1y, q0x, q0y = bx[:, j + 1], by[:, j + 1], bx[:, j], by[:, j]
                rect = ((np.minimum(p0x, p1x) <= np.maximum(q0x, q1x)) & (np.minimum(q0x, q1x) <= np.maximum(p0x, p1x))
                        & (np.minimum(p0y, p1y) <= np.maximum(q0y, q1y)) & (np.minimum(q0y, q1y) <= np.maximum(p0y, p1y)))
                s1 = _cross3(q0x, q0y, p1x, p1y, p0x, p0y)
                s2 = _cross3(p1x, p1y, q1x, q1y, p0x, p0y)
                s3 = _cross3(p0x, p0y, q1x, q1y, q0x, q0y)
                s4 = _cross3(q1x, q1y, p1x, p1y, q0x, q0y)
                ok = rect & (s1 * s2 > 0) & (s3 * s4 > 0)
                s5 = _cross3(q1x, q1y, p1x, p1y, p0x, p0y)
                gen = np.abs(s5 - s1) > EPS
                x1 = (s5 * q0x - s1 * q1x) / (s5 - s1)
                y1 = (s5 * q0y - s1 * q1y) / (s5 - s1)
                a0, b0, c0 = p0y - p1y, p1x - p0x, p0x * p1y - p1x * p0y
                a1, b1, c1 = q0y - q1y, q1x - q0x, q0x * q1y - q1x * q0y
                D = a0 * b1 - a1 * b0
                x2 = (b0 * c1 - b1 * c0) / D
                y2 = (a1 * c0 - a0 * c1) / D
                px.append(np.where(gen, x1, x2)), py.append(np.where(gen, y1, y2)), valid.append(ok)
        for k in range(4):
            px.append(bx[:, k]), py.append(by[:, k]), valid.append(_in_box(a, bx[:, k], by[:, k], margin))
            px.append(ax[:, k]), py.append(ay[:, k]), valid.append(_in_box(b, ax[:, k], ay[:, k], margin))
        px, py, valid = np.stack(px, 1).astype(F), np.stack(py, 1).astype(F), np.stack(valid, 1)
        cnt = valid.sum(1)
        cxs, cys = np.zeros(P, F), np.zeros(P, F)
        for t in range(px.shape[1]):
            cxs = np.where(valid[:, t], cxs + px[:, t], cxs).astype(F)
            cys = np.where(valid[:, t], cys + py[:, t], cys).astype(F)
        cxs, cys = cxs / cnt.astype(F), cys / cnt.astype(F)
        ang = np.arctan2(py - cys[:, None], px - cxs[:, None]).astype(F)
        ang = np.where(valid, ang, np.inf)
        order = np.argsort(ang, axis=1, kind="stable")
        sx = np.take_along_axis(px, order, 1)
        sy = np.take_along_axis(py, order, 1)
        area = np.zeros(P, F)
        for k in range(px.shape[1] - 1):
            term = (sx[:, k] - sx[:, 0]) * (sy[:, k + 1] - sy[:, 0]) - (sy[:, k] - sy[:, 0]) * (sx[:, k + 1] - sx[:, 0])
            area = np.where(k < cnt - 1, area + term, area).astype(F)
    return (np.abs(area) / F(2)).astype(F)


def iou_bev_rotated(a, b):
    sa, sb = a[:, 3] * a[:, 4], b[:, 3] * b[:, 4]
    so = box_overlap_rotated(a, b)
    return (so / np.maximum(sa + sb - so, EPS)).astype(F)


def iou_bev_normal(a, b):
    a, b = a.astype(F), b.astype(F)
    left = np.maximum(a[:, 0] - a[:, 3] / F(2), b[:, 0] - b[:, 3] / F(2))
    right = np.minimum(a[:, 0] + a[:, 3] / F(2), b[:, 0] + b[:, 3] / F(2))
    top = np.maximum(a[:, 1] - a[:, 4] / F(2), b[:, 1] - b[:, 4] / F(2))
    bottom = np.minimum(a[:, 1] + a[:, 4] / F(2), b[:, 1] + b[:, 4] / F(2))
    w, h = np.maximum(right - left, F(0)), np.maximum(bottom - top, F(0))
    inter = w * h
    return (inter / np.maximum(a[:, 3] * a[:, 4] + b[:, 3] * b[:, 4] - inter, EPS)).astype(F)


def _greedy_from_matrix(iou, thr):
    n = len(iou)
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        removed |= iou[i] > thr
        removed[: i + 1] = removed[: i + 1]  # earlier entries are irrelevant
    return keep


def _nms_bev(boxes, scores, thr, iou_fn):
    boxes = np.asarray(boxes, F)
    scores = np.asarray(scores, F)
    n = len(boxes)
    if n == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(-scores, kind="stable")
    b = boxes[order]
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    iou = iou_fn(b[ii.ravel()], b[jj.ravel()]).reshape(n, n)
    iou = np.where(jj > ii, iou, F(0))
    keep = _greedy_from_matrix(iou, F(thr))
    return order[np.asarray(keep, np.int64)]


def nms3d(boxes, scores, thr):
    """mmcv.ops.nms3d: boxes [N,7] (x,y,z,dx,dy,dz,yaw)."""
    return _nms_bev(boxes, scores, thr, iou_bev_rotated)


def nms3d_normal(boxes, scores, thr):
    """mmcv.ops.nms3d_normal: boxes [N,7], yaw ignored."""
    return _nms_bev(boxes, scores, thr, iou_bev_normal)


def aligned_3d_nms(corner_boxes, scores, classes, thr):
    """mmdet3d aligned_3d_nms: corner boxes [N,6] (x1,y1,z1,x2,y2,z2)."""
    bx = np.asarray(corner_boxes, F)
    scores = np.asarray(scores, F)
    classes = np.asarray(classes)
    if len(bx) == 0:
        return np.zeros(0, np.int64)
    area = (bx[:, 3] - bx[:, 0]) * (bx[:, 4] - bx[:, 1]) * (bx[:, 5] - bx[:, 2])
    idx = np.argsort(scores, kind="stable")
    pick = []
    while len(idx):
        i = idx[-1]
        pick.append(i)
        r = idx[:-1]
        x1 = np.maximum(bx[i, 0], bx[r, 0]); y1 = np.maximum(bx[i, 1], bx[r, 1]); z1 = np.maximum(bx[i, 2], bx[r, 2])
        x2 = np.minimum(bx[i, 3], bx[r, 3]); y2 = np.minimum(bx[i, 4], bx[r, 4]); z2 = np.minimum(bx[i, 5], bx[r, 5])
        inter = np.maximum(F(0), x2 - x1) * np.maximum(F(0), y2 - y1) * np.maximum(F(0), z2 - z1)
        with np.errstate(all="ignore"):
            iou = inter / (area[i] + area[r] - inter)
        iou = iou * (classes[i] == classes[r]).astype(F)
        idx = r[iou <= F(thr)]
    return np.asarray(pick, np.int64)
