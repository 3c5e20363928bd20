"""UniDet3DCriterion: matcher + loss values on the GPU (reference: unidet3d/criterion.py:7-178, 200-320).

Same registry name and constructor arguments as the reference class; ``__call__(pred, insts, datasets_names)``
returns ``{'det_loss': tensor}`` like criterion.py:144-178.  One C-ABI call per (decoder layer, scene)
(``ud3d_criterion_layer``: UniMatcher cost / top-k threshold / match, weighted cross-entropy terms, DIoU box-loss
terms); the handful of scalar combinations (per-scene weights, means over scenes, loss weights, sum over layers) are
done on the 4-float results with torch.

The returned tensor does not carry an autograd graph: the gradients w.r.t. the logits and boxes come from
``ud3d_criterion_layer_grad`` (``unidet3d_b200.train.criterion_backward``), which the training step feeds into the
encoder's and the backbone's backward passes.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import ops
from .registry import register_model


def _cfg_weight(costs, type_name, default):
    for c in costs or ():
        if c.get("type") == type_name:
            return float(c.get("weight", default))
    return default


@register_model
class UniDet3DCriterion:
    def __init__(self, matcher, loss_weight, non_object_weight, iter_matcher, bbox_loss_simple, bbox_loss_rotated,
                 datasets, datasets_weights, topk):
        for cfg in (bbox_loss_simple, bbox_loss_rotated):
            if cfg is not None and cfg.get("mode", "diou") != "diou":
                raise NotImplementedError("only the DIoU box losses the reference configs use are implemented")
        costs = (matcher or {}).get("costs", [])
        self.w_cls = _cfg_weight(costs, "QueryClassificationCost", 0.5)
        self.w_box = _cfg_weight(costs, "BboxCostJointTraining", 2.0)
        self.loss_weight = [float(w) for w in loss_weight]
        self.non_object_weight = float(non_object_weight)
        self.iter_matcher = bool(iter_matcher)
        self.datasets = list(datasets)
        self.datasets_weights = [float(w) for w in datasets_weights]
        self.topk = [int(k) for k in topk]

    @staticmethod
    def _gt(inst):
        """(labels int64 [G], boxes fp32 [G, 6|7] = gravity centre + size (+ yaw), query_masks [G, T])."""
        b = inst.bboxes_3d
        boxes = torch.cat((b.gravity_center, b.tensor[:, 3:] if b.with_yaw else b.tensor[:, 3:6]), dim=1)
        return inst.labels_3d.long().contiguous(), boxes.float().contiguous(), inst.query_masks

    def layer_terms(self, aux_outputs, insts, datasets_names):
        """-> per scene (match bool [T, G], sums [4]) of one decoder layer."""
        out = []
        for cls_pred, bbox, inst, name in zip(aux_outputs["cls_preds"], aux_outputs["bboxes"], insts, datasets_names):
            labels, boxes, qm = self._gt(inst)
            idx = self.datasets.index(name)
            out.append(ops.criterion_layer(cls_pred, bbox.contiguous(), boxes, labels, qm, self.topk[idx], self.w_cls,
                                           self.w_box, self.non_object_weight))
        return out

    def get_layer_loss(self, aux_outputs, insts, datasets_names, indices=None):
        """criterion.py:44-142.  ``indices`` (a fixed matching shared by all layers, iter_matcher=False) is not
        supported: every reference config sets iter_matcher=True."""
        if indices is not None:
            raise NotImplementedError("iter_matcher=False")
        terms = self.layer_terms(aux_outputs, insts, datasets_names)
        sums = torch.stack([s for _, s in terms])                              # [B, 4]
        return self.combine(sums, self.scene_weights(sums, datasets_names))

    def scene_weights(self, sums, datasets_names):
        """dataset weight of every scene of the batch as a device vector (cached per batch composition)."""
        key = (tuple(datasets_names), sums.device)
        cache = self.__dict__.setdefault("_w_cache", {})
        if key not in cache:
            if len(cache) > 64:
                cache.clear()
            cache[key] = sums.new_tensor([self.datasets_weights[self.datasets.index(n)] for n in datasets_names])
        return cache[key]

    def combine(self, sums, w):
        """criterion.py:111,136-142: per-scene sums [B, 4] + dataset weights [B] -> the layer's loss (device scalar)."""
        cls_loss = (w * sums[:, 0] / sums[:, 1]).mean()
        has = sums[:, 3] > 0
        n_has = has.sum()
        box_each = torch.where(has, w * sums[:, 2] / sums[:, 3].clamp(min=1), sums.new_zeros(()))
        bbox_loss = torch.where(n_has > 0, box_each.sum() / n_has.clamp(min=1), sums.new_zeros(()))
        return self.loss_weight[0] * cls_loss + self.loss_weight[1] * bbox_loss

    def __call__(self, pred, insts, datasets_names):
        """criterion.py:144-178."""
        loss = self.get_layer_loss(pred, insts, datasets_names)
        if "aux_outputs" in pred:
            if not self.iter_matcher:
                raise NotImplementedError("iter_matcher=False")
            for aux in pred["aux_outputs"]:
                loss = loss + self.get_layer_loss(aux, insts, datasets_names)
        return {"det_loss": loss}
