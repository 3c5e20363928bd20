"""Oracle (test infrastructure only): CPU restatement of the reference's detection evaluator.

Follows unidet3d/indoor_eval.py -- ``average_precision`` (:8-53, mode 'area'), ``eval_det_cls`` (:56-160),
``eval_map_recall`` (:163-202), ``indoor_eval`` (:205-300) -- on flat numpy arrays instead of dicts of box objects, and
restates the third-party IoU it calls (``DepthInstance3DBoxes.overlaps`` of mmdet3d 1.4.0 = ``BaseInstance3DBoxes.overlaps``:
height overlap x rotated BEV overlap recovered from mmcv ``box_iou_rotated``; absent from /root/reference).

Pinned: the evaluation logic (class bookkeeping, TP / FP marking, AP, the nan conventions of classes without ground truth)
against the reference's own ``indoor_eval`` executed with stub box objects (tests/golden/evaluate_ref.npz).  The IoU
itself is third-party: parity unpinned for rotated boxes (restated; axis-aligned boxes reduce to interval overlaps).
"""
import numpy as np

from . import nms as onms


def overlaps_3d(b1, b2):
    """b1 [n,7], b2 [m,7] = (cx, cy, cz, dx, dy, dz, yaw), gravity centres -> IoU [n,m] (float32 arithmetic like torch)."""
    b1 = np.asarray(b1, np.float32).reshape(-1, 7)
    b2 = np.asarray(b2, np.float32).reshape(-1, 7)
    n, m = len(b1), len(b2)
    if n * m == 0:
        return np.zeros((n, m), np.float32)
    top1, bot1 = b1[:, 2] + b1[:, 5] / 2, b1[:, 2] - b1[:, 5] / 2
    top2, bot2 = b2[:, 2] + b2[:, 5] / 2, b2[:, 2] - b2[:, 5] / 2
    oh = np.clip(np.minimum(top1[:, None], top2[None]) - np.maximum(bot1[:, None], bot2[None]), 0, None)
    d1 = np.clip(b1[:, 3:5], 1e-4, None)
    d2 = np.clip(b2[:, 3:5], 1e-4, None)
    A = np.repeat(np.concatenate([b1[:, :3], d1, b1[:, 5:]], 1), m, 0)
    B = np.tile(np.concatenate([b2[:, :3], d2, b2[:, 5:]], 1), (n, 1))
    aligned = (A[:, 6] == 0) & (B[:, 6] == 0)
    inter = np.zeros(n * m, np.float32)
    if aligned.any():
        a, b = A[aligned], B[aligned]
        wx = np.clip(np.minimum(a[:, 0] + a[:, 3] / 2, b[:, 0] + b[:, 3] / 2) - np.maximum(a[:, 0] - a[:, 3] / 2, b[:, 0] - b[:, 3] / 2), 0, None)
        wy = np.clip(np.minimum(a[:, 1] + a[:, 4] / 2, b[:, 1] + b[:, 4] / 2) - np.maximum(a[:, 1] - a[:, 4] / 2, b[:, 1] - b[:, 4] / 2), 0, None)
        inter[aligned] = wx * wy
    if (~aligned).any():
        inter[~aligned] = onms.box_overlap_rotated(A[~aligned], B[~aligned], margin=1e-5)
    inter = inter.reshape(n, m)
    o3 = inter * oh
    v1 = (b1[:, 3] * b1[:, 4] * b1[:, 5])[:, None]
    v2 = (b2[:, 3] * b2[:, 4] * b2[:, 5])[None]
    return (o3 / np.clip(v1 + v2 - o3, 1e-8, None)).astype(np.float32)


def average_precision(recalls, precisions):
    """indoor_eval.py:8-53, mode 'area', single scale.  -> float32 scalar."""
    mrec = np.hstack((0.0, recalls, 1.0))
    mpre = np.hstack((0.0, precisions, 0.0))
    for i in range(len(mpre) - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return np.float32(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def eval_det_cls(det_boxes, det_scores, det_img, gt_boxes, gt_img, iou_thr, order=None):
    """One class (indoor_eval.py:56-160).  -> list over thresholds of (recall[], precision[], ap)."""
    npos = len(gt_boxes)
    nd = len(det_boxes)
    iou_max = np.full(nd, -np.inf)
    jmax = np.zeros(nd, np.int64)
    for img in np.unique(det_img):
        d = np.where(det_img == img)[0]
        g = np.where(gt_img == img)[0]
        if len(g):
            iou = overlaps_3d(det_boxes[d], gt_boxes[g])
            iou_max[d] = iou.max(1)
            jmax[d] = g[iou.argmax(1)]            # first maximum, like the strict '>' scan
    if order is None:
        order = np.argsort(-det_scores, kind="stable")
    ret = []
    for thr in iou_thr:
        taken = np.zeros(npos, bool)
        gmap = {}
        tp = np.zeros(nd)
        fp = np.zeros(nd)
        for r, d in enumerate(order):
            if iou_max[d] > thr:
                j = jmax[d]
                if j not in gmap:
                    gmap[j] = True
                    tp[r] = 1.0
                else:
                    fp[r] = 1.0
            else:
                fp[r] = 1.0
        fpc, tpc = np.cumsum(fp), np.cumsum(tp)
        with np.errstate(divide="ignore", invalid="ignore"):
            recall = tpc / float(npos)
        precision = tpc / np.maximum(tpc + fpc, np.finfo(np.float64).eps)
        ret.append((recall, precision, average_precision(recall, precision)))
    return ret


def indoor_eval(det_boxes, det_scores, det_labels, det_img, gt_boxes, gt_labels, gt_img, metric, label2cat):
    """indoor_eval.py:205-300 on flat arrays -> the reference's ret_dict (AP / recall per class, mAP, mAR per threshold)."""
    det_labels = np.asarray(det_labels, np.int64)
    gt_labels = np.asarray(gt_labels, np.int64)
    # gt.keys(): classes in first-appearance order over images (detections of an image first, then its ground truth)
    keys = []
    n_img = int(max(det_img.max(initial=-1), gt_img.max(initial=-1))) + 1
    for img in range(n_img):
        for l in det_labels[det_img == img]:
            if l not in keys:
                keys.append(int(l))
        for l in gt_labels[gt_img == img]:
            if l not in keys:
                keys.append(int(l))
    pred_keys = set(int(l) for l in det_labels)
    rec = [dict() for _ in metric]
    ap = [dict() for _ in metric]
    for label in keys:
        if label in pred_keys:
            dm, gm = det_labels == label, gt_labels == label
            r = eval_det_cls(det_boxes[dm], det_scores[dm], det_img[dm], gt_boxes[gm], gt_img[gm], metric)
            for i in range(len(metric)):
                rec[i][label], ap[i][label] = r[i][0], np.array([r[i][2]], np.float32)
        else:
            for i in range(len(metric)):
                rec[i][label], ap[i][label] = np.zeros(1), np.zeros(1)
    ret = {}
    with np.errstate(invalid="ignore"):
        for i, thr in enumerate(metric):
            for label in ap[i]:
                ret[f'{label2cat[label]}_AP_{thr:.2f}'] = float(ap[i][label][0])
            ret[f'mAP_{thr:.2f}'] = float(np.nanmean(list(ap[i].values())))
            rl = []
            for label in rec[i]:
                ret[f'{label2cat[label]}_rec_{thr:.2f}'] = float(rec[i][label][-1])
                rl.append(rec[i][label][-1])
            ret[f'mAR_{thr:.2f}'] = float(np.nanmean(rl))
    return ret
