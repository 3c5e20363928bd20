"""Detection evaluator on the GPU, same call surface as the reference's ``indoor_eval`` (unidet3d/indoor_eval.py:205-300)
minus the ASCII table: boxes stay on the device between ``predict`` and the metric.

    ret = indoor_eval(gt_annos, dt_annos, metric=[0.25, 0.5], label2cat=classes)

``dt_annos[i]``: dict(bboxes_3d = CUDA tensor [n, 6 | 7] (cx, cy, cz, dx, dy, dz[, yaw]; gravity centre -- what
``UniDet3D.predict`` returns), scores_3d [n], labels_3d [n]); ``gt_annos[i]``: dict(gt_bboxes_3d [m, 6 | 7], gt_labels_3d [m]).
Returns the reference's dict: ``{cat}_AP_{thr}``, ``{cat}_rec_{thr}``, ``mAP_{thr}``, ``mAR_{thr}`` with the same
conventions (classes keyed in first-appearance order; a class with ground truth but no detection counts 0; a class with
detections but no ground truth is nan and skipped by the means).  Sorting (a two-key stable sort) and the concatenations
use torch; matching, TP / FP marking and the AP integration are ``ud3d_eval_detections`` (csrc/eval.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream, check


def _boxes7(b: torch.Tensor) -> torch.Tensor:
    b = b.float().reshape(-1, b.shape[-1]) if b.numel() else b.float().reshape(0, 7)
    if b.shape[1] == 6:
        b = torch.cat((b, b.new_zeros((b.shape[0], 1))), 1)
    return b.contiguous()


def eval_detections(det_boxes, det_scores, det_labels, det_img, gt_boxes, gt_labels, gt_img_offsets, n_cls: int,
                    thresholds: Sequence[float]):
    """Flat device arrays -> (ap float32 [n_cls, n_thr], final recall float64 [n_cls, n_thr], npos int32 [n_cls],
    detections per class int64 [n_cls])."""
    dev = det_boxes.device
    if not det_boxes.is_cuda:
        raise _lib.Ud3dError("eval_detections: expected CUDA tensors (unidet3d_b200 has no CPU path)")
    lib = _lib.load()
    D, G, n_thr = det_boxes.shape[0], gt_boxes.shape[0], len(thresholds)
    # (label ascending, score descending): stable sort by score, then stable sort by label
    idx = torch.sort(det_scores, descending=True, stable=True)[1]
    order = idx[torch.sort(det_labels[idx], stable=True)[1]].int().contiguous()
    counts = torch.bincount(det_labels.long(), minlength=n_cls)[:n_cls]
    class_offsets = torch.cat((counts.new_zeros(1), counts.cumsum(0))).int().contiguous()
    ap = torch.empty((n_cls, n_thr), dtype=torch.float32, device=dev)
    rec = torch.empty((n_cls, n_thr), dtype=torch.float64, device=dev)
    npos = torch.empty(n_cls, dtype=torch.int32, device=dev)
    wsb = int(lib.ud3d_eval_workspace_bytes(D, G, n_thr))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    thr = (C.c_float * n_thr)(*[float(t) for t in thresholds])
    check(lib.ud3d_eval_detections(_p(det_boxes), _p(det_labels), _p(det_img), D, _p(order), _p(class_offsets), n_cls,
                                   _p(gt_boxes), _p(gt_labels), _p(gt_img_offsets), G, gt_img_offsets.numel() - 1, thr, n_thr,
                                   _p(ap), _p(rec), _p(npos), _p(ws), wsb, _stream()), "ud3d_eval_detections")
    return ap, rec, npos, counts


def indoor_eval(gt_annos: List[dict], dt_annos: List[dict], metric: Sequence[float], label2cat) -> Dict[str, float]:
    assert len(dt_annos) == len(gt_annos)
    dev = None
    for a in dt_annos:
        dev = a['bboxes_3d'].device
        break
    n_img = len(dt_annos)
    det_boxes = torch.cat([_boxes7(a['bboxes_3d']) for a in dt_annos])
    det_scores = torch.cat([a['scores_3d'].float().reshape(-1) for a in dt_annos]).contiguous()
    det_labels = torch.cat([a['labels_3d'].reshape(-1) for a in dt_annos]).int().contiguous()
    det_img = torch.cat([torch.full((len(a['labels_3d']),), i, dtype=torch.int32, device=dev) for i, a in enumerate(dt_annos)])
    gt_boxes = torch.cat([_boxes7(torch.as_tensor(a['gt_bboxes_3d'], device=dev)) for a in gt_annos])
    gt_labels = torch.cat([torch.as_tensor(a['gt_labels_3d'], device=dev).reshape(-1) for a in gt_annos]).int().contiguous()
    gt_counts = [int(torch.as_tensor(a['gt_labels_3d']).numel()) for a in gt_annos]
    gt_img_offsets = torch.tensor(np.cumsum([0] + gt_counts), dtype=torch.int32, device=dev)
    n_cls = len(label2cat)
    ap, rec, npos, counts = eval_detections(det_boxes, det_scores, det_labels, det_img, gt_boxes, gt_labels, gt_img_offsets,
                                            n_cls, metric)
    # ONE read-back: the per-class results and, for the reference's key order, the first image each class appears in
    first_det = torch.full((n_cls,), n_img, dtype=torch.int64, device=dev).scatter_reduce(0, det_labels.long(), det_img.long(), "amin")
    gt_img = torch.repeat_interleave(torch.arange(n_img, device=dev), torch.tensor(gt_counts, device=dev))
    first_gt = torch.full((n_cls,), n_img, dtype=torch.int64, device=dev).scatter_reduce(0, gt_labels.long(), gt_img, "amin")
    ap_h, rec_h, cnt_h, fd, fg = ap.cpu().numpy(), rec.cpu().numpy(), counts.cpu().numpy(), first_det.cpu().numpy(), first_gt.cpu().numpy()
    # gt.keys() of the reference: insertion order while walking the images (an image's detections, then its ground truth)
    keys = _key_order(dt_annos, gt_annos, fd, fg, n_img)
    ret: Dict[str, float] = {}
    with np.errstate(invalid="ignore"):
        for i, thr in enumerate(metric):
            aps, recs = [], []
            for label in keys:
                a = float(ap_h[label, i]) if cnt_h[label] > 0 else 0.0
                ret[f'{label2cat[label]}_AP_{thr:.2f}'] = a
                aps.append(a)
            ret[f'mAP_{thr:.2f}'] = float(np.nanmean(aps)) if aps else float('nan')
            for label in keys:
                r = float(rec_h[label, i]) if cnt_h[label] > 0 else 0.0
                ret[f'{label2cat[label]}_rec_{thr:.2f}'] = r
                recs.append(r)
            ret[f'mAR_{thr:.2f}'] = float(np.nanmean(recs)) if recs else float('nan')
    return ret


def _key_order(dt_annos, gt_annos, first_det, first_gt, n_img):
    """Classes in the insertion order of the reference's ``gt`` dict.  Only the order inside one image needs the labels
    themselves; classes are bucketed by the first image they appear in (device-side minima) and ordered inside the
    bucket by walking that image's detections, then its ground truth."""
    first = np.minimum(first_det, first_gt)
    keys: List[int] = []
    for img in np.unique(first[first < n_img]):
        img = int(img)
        cand = set(np.where(first == img)[0].tolist())
        for seq in (dt_annos[img]['labels_3d'], gt_annos[img]['gt_labels_3d']):
            for l in torch.as_tensor(seq).reshape(-1).tolist():
                l = int(l)
                if l in cand:
                    keys.append(l)
                    cand.discard(l)
    return keys
