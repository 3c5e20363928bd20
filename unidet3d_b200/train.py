"""Backward pass of the backbone in train mode: gradients of every parameter of ``input_conv`` / ``SpConvUNet`` /
``output_layer`` (reference unidet3d/unidet3d.py:95-134, spconv_unet.py:13-240 under torch.autograd) from the
gradient of the pooled superpoint features, on the library's kernels:

* sparse conv: weight gradient ``ud3d_conv_wgrad``; input gradient = the forward gather-GEMM on dY with the transposed
  weight over the transposed rulebook (SubM3: same table, offsets reversed; k2s2 conv <-> its inverse conv's table);
* train-mode (Sync)BatchNorm + ReLU: ``ud3d_bn_backward_sums`` / ``_apply`` (one all-reduce of 2C + 1 doubles per
  BatchNorm under torch.distributed);
* superpoint mean-pool: ``ud3d_segmented_mean_backward``.

The forward is the train-mode executor of ``SpConvUNet`` re-run op by op while a tape records one closure per op; the
backward replays the tape in reverse, accumulating gradients per tensor (a residual input or the skip concat has two
consumers).  Activations after BatchNorm + ReLU are recomputed, not stored.  Further down: the encoder's tape, the
criterion's gradients and the training step (``loss_backward`` / ``train_step``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .spconv_unet import SparseConvWeight, SpConvUNet


class Tape:
    def __init__(self, group=None):
        self.steps, self.grads, self.keep, self.group = [], {}, [], group
        self.named = {}                       # a few intermediate tensors by name (tests / debugging)

    def grad(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        return self.grads.get(id(t))

    def add(self, t: torch.Tensor, g: torch.Tensor):
        self.keep.append(t)                       # ids stay unique while the tape is alive
        cur = self.grads.get(id(t))
        self.grads[id(t)] = g if cur is None else cur + g

    def backward(self):
        for step in reversed(self.steps):
            step()


def _acc_param(p: torch.nn.Parameter, g: torch.Tensor):
    g = g.reshape(p.shape).to(p.dtype)
    p.grad = g if p.grad is None else p.grad + g


def conv(tape: Tape, x: torch.Tensor, conv_mod: SparseConvWeight, K: int, table, mask, n_out: int, table_t, reverse: bool,
         bn=None, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         need_input_grad: bool = True) -> torch.Tensor:
    """y = sparse_conv(relu(bn_train(x)) if bn else x) (+ residual); records its backward."""
    s = h = mean = invstd = None
    if bn is not None:
        s, h, mean, invstd = ops.bn_train(x, bn, group=tape.group)
    y = ops.gemm(x, ops.PackedWeight(conv_mod.weight, ts=False), table=table, tile_mask=mask, n_out=n_out, in_scale=s, in_shift=h,
                 in_relu=bn is not None, residual=residual, out=out)
    n_in = x.shape[0]

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        if residual is not None:
            tape.add(residual, dy)
        a = ops.bn_relu_apply(x, s, h) if bn is not None else x
        _acc_param(conv_mod.weight, ops.conv_wgrad(a, dy, K, table))
        if not need_input_grad:
            return
        da = ops.conv_dgrad(dy, conv_mod.weight.detach(), table_t, n_in, reverse_offsets=reverse)
        if bn is None:
            tape.add(x, da)
            return
        dx, dgamma, dbeta = ops.bn_relu_backward(x, da, s, h, mean, invstd, group=tape.group)
        _acc_param(bn.weight, dgamma)
        _acc_param(bn.bias, dbeta)
        tape.add(x, dx)

    tape.steps.append(bwd)
    return y


def _block(tape, blk, x, lv):
    cb = blk.conv_branch
    if isinstance(blk.i_branch[0], SparseConvWeight):
        identity = conv(tape, x, blk.i_branch[0], 1, None, None, x.shape[0], None, False)
    else:
        identity = x
    y = conv(tape, x, cb[2], 27, lv.subm, lv.subm_mask, lv.n, lv.subm, True, bn=cb[0])
    return conv(tape, y, cb[5], 27, lv.subm, lv.subm_mask, lv.n, lv.subm, True, bn=cb[3], residual=identity)


def unet_forward(tape: Tape, unet: SpConvUNet, x: torch.Tensor, pyr, l: int = 0) -> torch.Tensor:
    """spconv_unet.py:205-240 in train mode (the op sequence of SpConvUNet._forward_level_train) with the tape."""
    lv = pyr.levels[l]
    c = unet.num_planes[0]
    for blk in unet.blocks:
        x = _block(tape, blk, x, lv)
    if len(unet.num_planes) > 1:
        nxt = pyr.levels[l + 1]
        cat = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=x.device)
        cat[:, :c] = x
        x_in = x
        d = conv(tape, x, unet.conv[2], 8, lv.child, lv.child_mask, nxt.n, lv.up, False, bn=unet.conv[0])
        d = unet_forward(tape, unet.u, d, pyr, l + 1)
        up = conv(tape, d, unet.deconv[2], 8, lv.up, lv.up_mask, lv.n, lv.child, False, bn=unet.deconv[0], out=cat[:, c:])

        def split():                                   # gradient of the skip concat [identity | decoder]
            dc = tape.grad(cat)
            if dc is not None:
                tape.add(x_in, dc[:, :c])
                tape.add(up, dc[:, c:])

        tape.steps.append(split)
        x = cat
        for blk in unet.blocks_tail:
            x = _block(tape, blk, x, lv)
    return x


def backbone_forward(model, x, superpoints: torch.Tensor, inverse_mapping: torch.Tensor, n_superpoints: int, group=None):
    """``UniDet3D.extract_feat`` in train mode with a tape.  -> (pooled [n_sp, C], tape).  ``model.training`` must be True
    (running statistics are updated like in any train-mode forward)."""
    tape = Tape(group)
    lv0 = x.pyramid.levels[0]
    f = conv(tape, x.features, model.input_conv[0], 27, lv0.subm, lv0.subm_mask, lv0.n, lv0.subm, True, need_input_grad=False)
    y = unet_forward(tape, model.unet, f, x.pyramid)
    tape.named.update(stem=f, unet_out=y)
    bn = model.output_layer[0]
    sc, sh, mean, invstd = ops.bn_train(y, bn, group=group)
    pooled = ops.segmented_mean(y, superpoints, n_superpoints, gather=inverse_mapping, scale=sc, shift=sh, relu=True)

    def bwd():
        dp = tape.grad(pooled)
        da = ops.segmented_mean_backward(dp.contiguous(), superpoints, y.shape[0], gather=inverse_mapping)
        dy, dgamma, dbeta = ops.bn_relu_backward(y, da, sc, sh, mean, invstd, group=group)
        _acc_param(bn.weight, dgamma)
        _acc_param(bn.bias, dbeta)
        tape.add(y, dy)

    tape.steps.append(bwd)
    return pooled, tape


def backbone_backward(tape: Tape, pooled: torch.Tensor, d_pooled: torch.Tensor):
    """Fills ``.grad`` of every backbone parameter (accumulating into existing gradients like torch)."""
    tape.add(pooled, d_pooled)
    with torch.no_grad():
        tape.backward()


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20):
    """Data-parallel gradient exchange (tools/train.py:49-52 runs the reference under DDP): average the gradients over
    the ranks in buckets of ``bucket_bytes`` (one flat buffer per bucket, one all-reduce each)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    n_coll, i = 0, 0
    while i < len(grads):
        j, size = i, 0
        while j < len(grads) and (j == i or size + grads[j].numel() * 4 <= bucket_bytes):
            size += grads[j].numel() * 4
            j += 1
        flat = torch.cat([g.reshape(-1) for g in grads[i:j]])
        dist.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for g in grads[i:j]:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        i = j
    return n_coll


# ---------------------------------------------------------------------------------------------------------------------
# Encoder (reference unidet3d/encoder.py:113-239 in train mode: all num_layers + 1 heads are evaluated, :219-229)
def _linear(tape: Tape, x: torch.Tensor, lin: torch.nn.Linear, act: Optional[str] = None, residual: Optional[torch.Tensor] = None,
            weight=None, bias=None) -> torch.Tensor:
    """y = act(x W^T + b) (+ residual); the pre-activation map is kept for the backward pass."""
    W = lin.weight if weight is None else weight
    b = lin.bias if bias is None else bias
    pre = ops.gemm(x, ops.PackedWeight(W, ts=False), bias=b.detach().float().contiguous(), residual=residual if act is None else None)
    y = pre if act is None else ops.activation_forward(pre, act)
    if act is not None and residual is not None:
        raise NotImplementedError("activation + residual does not occur in the encoder")

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        dy = dy.contiguous()
        if residual is not None:
            tape.add(residual, dy)
        dpre = dy if act is None else ops.activation_backward(pre, dy, act)
        Wd = W.detach()
        _acc_param(W, ops.conv_wgrad(x, dpre, 1))
        _acc_param(b, ops.bn_batch_sums(dpre)[0].float())
        tape.add(x, ops.conv_dgrad(dpre, Wd.reshape(Wd.shape[0], 1, Wd.shape[1]), None, x.shape[0], reverse_offsets=False))

    tape.steps.append(bwd)
    return y


def _layernorm(tape: Tape, x: torch.Tensor, ln: torch.nn.LayerNorm) -> torch.Tensor:
    g = ln.weight.detach().float().contiguous()
    y = ops.layernorm(x, g, ln.bias.detach().float().contiguous(), eps=ln.eps)

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        dx, dgamma, dbeta = ops.layernorm_backward(x, dy.contiguous(), g, ln.eps)
        _acc_param(ln.weight, dgamma)
        _acc_param(ln.bias, dbeta)
        tape.add(x, dx)

    tape.steps.append(bwd)
    return y


def _head(tape: Tape, enc, Hq: torch.Tensor, centers: torch.Tensor, bounds, ds_idx):
    """encoder.py:165-201: out_norm -> class MLP (union of classes, per-dataset column gather) and box Linear + decode."""
    plan = enc._get_plan()
    nq = _layernorm(tape, Hq, enc.out_norm)
    h = _linear(tape, nq, enc.outs_cls[0], act="relu")
    logits = _linear(tape, h, enc.outs_cls[2])
    raw = _linear(tape, nq, enc.out_bboxes.linear)
    cls_preds, bboxes = [], []
    for i, j in enumerate(ds_idx):
        a, b = bounds[i], bounds[i + 1]
        cls_preds.append(ops.gather_columns(logits[a:b], plan["cols"][j]))
        bboxes.append(ops.bbox_decode(raw[a:b], centers[a:b], bool(enc.angles[j])))

    def bwd():
        grads = [(tape.grad(c), tape.grad(b)) for c, b in zip(cls_preds, bboxes)]
        if all(dc is None and db is None for dc, db in grads):
            return
        d_logits = torch.empty_like(logits)
        d_raw = torch.empty_like(raw)
        for i, j in enumerate(ds_idx):
            a, b = bounds[i], bounds[i + 1]
            dc, db = grads[i]
            ops.head_backward(raw[a:b], None if db is None else db.contiguous(), bool(enc.angles[j]),
                              None if dc is None else dc.contiguous(), plan["cols"][j], d_raw[a:b], d_logits[a:b])
        tape.add(logits, d_logits)
        tape.add(raw, d_raw)

    tape.steps.append(bwd)
    return cls_preds, bboxes


def encoder_forward(enc, X: torch.Tensor, centers: torch.Tensor, bounds, datasets_names, tape: Optional[Tape] = None):
    """``UniDet3DEncoder.forward`` on packed rows with a tape: -> (dict(cls_preds, bboxes, aux_outputs) exactly as the
    module returns it with all heads, tape).  fp32 dataflow (pre-activation maps are kept for the backward pass)."""
    tape = Tape() if tape is None else tape
    ds_idx = [enc.datasets.index(n) for n in datasets_names]
    cu = torch.tensor(bounds, dtype=torch.int32).to(X.device)
    max_T = max(b - a for a, b in zip(bounds[:-1], bounds[1:]))
    cls_all, box_all = [], []
    Hq = _linear(tape, X, enc.input_proj[0], act="relu")
    Hq = _linear(tape, Hq, enc.input_proj[2])
    c, b = _head(tape, enc, Hq, centers, bounds, ds_idx)
    cls_all.append(c), box_all.append(b)
    d = enc.d_model
    for sa, ff in zip(enc.self_attn_layers, enc.ffn_layers):
        qkv = _linear(tape, Hq, None, weight=sa.attn.in_proj_weight, bias=sa.attn.in_proj_bias)
        A = ops.attention(qkv, cu, max_T, enc.num_heads)

        def attn_bwd(qkv=qkv, A=A):
            dA = tape.grad(A)
            if dA is not None:
                tape.add(qkv, ops.attention_backward(qkv, cu, enc.num_heads, A, dA.contiguous()))

        tape.steps.append(attn_bwd)
        Z = _linear(tape, A, sa.attn.out_proj, residual=Hq)
        Hq = _layernorm(tape, Z, sa.norm)
        F1 = _linear(tape, Hq, ff.net[0], act=enc.activation_fn)
        Z = _linear(tape, F1, ff.net[3], residual=Hq)
        Hq = _layernorm(tape, Z, ff.norm)
        c, b = _head(tape, enc, Hq, centers, bounds, ds_idx)
        cls_all.append(c), box_all.append(b)
    aux = [dict(cls_preds=c, bboxes=b) for c, b in zip(cls_all[:-1], box_all[:-1])]
    return dict(cls_preds=cls_all[-1], bboxes=box_all[-1], aux_outputs=aux), tape


def encoder_backward(tape: Tape, outputs, d_cls, d_boxes):
    """d_cls / d_boxes: per head (aux heads first, final head last, like ``aux_outputs + [final]``) lists over scenes of
    gradients w.r.t. cls_preds / bboxes (None = no gradient).  Fills ``.grad`` of the encoder's parameters; returns the
    tape so that ``tape.grad(X)`` gives the gradient w.r.t. the pooled input features."""
    heads = outputs["aux_outputs"] + [dict(cls_preds=outputs["cls_preds"], bboxes=outputs["bboxes"])]
    for hd, dc, db in zip(heads, d_cls, d_boxes):
        for t, g in zip(hd["cls_preds"], dc):
            if g is not None:
                tape.add(t, g)
        for t, g in zip(hd["bboxes"], db):
            if g is not None:
                tape.add(t, g)
    with torch.no_grad():
        tape.backward()
    return tape


# ---------------------------------------------------------------------------------------------------------------------
# Criterion gradients (reference: torch.autograd through criterion.py:44-178).  Per (head, scene): ud3d_criterion_layer
# (UniMatcher + loss sums) and ud3d_criterion_layer_grad (softmax - one-hot with the class weights; forward-mode DIoU
# derivatives of the matched pairs, axis-aligned and rotated).  The scalar bookkeeping between the two -- dataset
# weights, means over scenes, loss weights -- is a handful of ops on the [B, 4] sums and stays on the device (no host
# synchronisation: whether a scene contributes a box term depends on its number of matched pairs).
def criterion_backward(crit, outputs, insts, datasets_names, debug: Optional[dict] = None):
    """-> (det_loss value, d_cls, d_boxes) with d_* per head (aux heads first, final head last) lists over scenes, in
    the layout ``encoder_backward`` takes."""
    if not crit.iter_matcher:
        raise NotImplementedError("iter_matcher=False")
    heads = outputs["aux_outputs"] + [dict(cls_preds=outputs["cls_preds"], bboxes=outputs["bboxes"])]
    gts = [crit._gt(inst) for inst in insts]
    B = len(datasets_names)
    total = None
    d_cls, d_box = [], []
    for hd in heads:
        terms = crit.layer_terms(hd, insts, datasets_names)
        sums = torch.stack([s for _, s in terms])                               # [B, 4]
        w = crit.scene_weights(sums, datasets_names)
        loss = crit.combine(sums, w)
        total = loss if total is None else total + loss
        n_has = (sums[:, 3] > 0).sum().clamp(min=1)
        scales = torch.stack((crit.loss_weight[0] * w / B, crit.loss_weight[1] * w / n_has), dim=1).contiguous()   # [B, 2]
        dc, db = [], []
        for i, (cp, pb, (labels, boxes, _), (match, s)) in enumerate(zip(hd["cls_preds"], hd["bboxes"], gts, terms)):
            g_c, g_b = ops.criterion_layer_grad(cp, pb.contiguous(), boxes, labels, match, s, scales[i], crit.non_object_weight)
            dc.append(g_c), db.append(g_b)
        d_cls.append(dc), d_box.append(db)
        if debug is not None:
            debug.setdefault("matches", []).append([m for m, _ in terms])
    return total, d_cls, d_box


# ---------------------------------------------------------------------------------------------------------------------
# The training step (reference: UniDet3D.loss, unidet3d/unidet3d.py:277-364, under torch.autograd + DDP + AdamW,
# tools/train.py:49-52, configs/unidet3d_1xb8_scannet.py optim_wrapper)
class _Marks:
    """CUDA-event marks around the stages of one step (``profile_step``)."""

    def __init__(self):
        self.ev = []

    def __call__(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append((name, e))

    def result(self):
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.ev[:-1], self.ev[1:]):
            out[name + "_ms"] = out.get(name + "_ms", 0.0) + a.elapsed_time(b)
        return out


def profile_step(model, batch_inputs_dict, batch_data_samples, group=None):
    """One ``loss_backward`` with CUDA events between its stages -> {stage_ms}: targets + collate, backbone forward,
    encoder forward, matcher + criterion gradients, encoder backward, backbone backward."""
    marks = _Marks()
    for p in model.parameters():
        p.grad = None
    loss_backward(model, batch_inputs_dict, batch_data_samples, group=group, marks=marks)
    return marks.result()


def loss_backward(model, batch_inputs_dict, batch_data_samples, group=None, debug: Optional[dict] = None, marks=None):
    """Forward of ``UniDet3D.loss`` in train mode with a tape, then the backward pass: fills ``.grad`` of every parameter
    of the detector (accumulating like torch) and returns ``{'det_loss': value}``.  ``model.training`` must be True
    (batch-statistics BatchNorm; under a torch.distributed ``group`` of several ranks its sums are all-reduced in both
    directions).  Query selection uses ``torch.randperm`` on the host like the reference when a scene has more than
    ``query_thr`` superpoints."""
    if not model.training:
        raise RuntimeError("loss_backward: call model.train() first (the backward pass is the train-mode executor's)")
    mark = marks if marks is not None else (lambda name: None)
    with torch.no_grad():
        mark("start")
        li = model._loss_inputs(batch_inputs_dict, batch_data_samples)
        B, names, sp_off = li["B"], li["names"], li["sp_off"]
        x, inverse = model.collate(li["pts"], li["offs"], B, li["el"])
        mark("targets_collate")
        pooled, tape = backbone_forward(model, x, li["sp_b"], inverse, int(sp_off[-1]), group=group)
        mark("backbone_fwd")
        tape.steps.append(lambda: mark("encoder_bwd"))       # replayed in reverse: the encoder's steps end here
        # query selection (unidet3d.py:182-218): rows of `pooled` are scene-contiguous already
        sel, bounds, centers, qmasks = [], [0], [], []
        for i in range(B):
            a, b = int(sp_off[i]), int(sp_off[i + 1])
            ids = torch.arange(a, b, device=pooled.device)
            c, m = li["sp_centers"][i], li["sp_masks"][i]
            if b - a > model.query_thr:
                keep = torch.randperm(b - a)[:model.query_thr].to(pooled.device)
                ids, c, m = ids[keep], c[keep], m[:, keep]
            sel.append(ids), centers.append(c), qmasks.append(m)
            bounds.append(bounds[-1] + int(ids.numel()))
        for g, m in zip(li["gt_insts"], qmasks):
            g.query_masks = m
        if bounds[-1] == pooled.shape[0]:
            X = pooled                                   # nothing dropped: the encoder reads the pooled rows directly
        else:
            idx = torch.cat(sel)
            X = pooled.index_select(0, idx)

            def sel_bwd():
                dX = tape.grad(X)
                if dX is not None:
                    tape.add(pooled, torch.zeros_like(pooled).index_add_(0, idx, dX))

            tape.steps.append(sel_bwd)
        out, _ = encoder_forward(model.decoder, X, torch.cat(centers).contiguous(), bounds, names, tape=tape)
        mark("encoder_fwd")
        loss, d_cls, d_box = criterion_backward(model.criterion, out, li["gt_insts"], names, debug=debug)
        mark("criterion")
        if debug is not None:
            debug.update(outputs=out, pooled=pooled, gt_insts=li["gt_insts"])
        encoder_backward(tape, out, d_cls, d_box)        # replays the whole tape: encoder steps, then the backbone's
        mark("backbone_bwd")
        # the closures hold the tape and the tape holds the closures: break the cycle now, or the activations of this step
        # stay allocated until Python's cycle collector happens to run
        tape.steps.clear(), tape.grads.clear(), tape.keep.clear(), tape.named.clear()
    return {"det_loss": loss}


def train_step(model, optimizer, batch_inputs_dict, batch_data_samples, group=None, clip_grad_norm: Optional[float] = 10.0):
    """zero_grad -> loss_backward -> gradient all-reduce (data parallel) -> clip (configs: max_norm 10) -> optimizer.step().
    Packed weight images / folded BatchNorms are invalidated because the parameters change."""
    optimizer.zero_grad(set_to_none=True)
    out = loss_backward(model, batch_inputs_dict, batch_data_samples, group=group)
    params = [p for p in model.parameters() if p.grad is not None]
    allreduce_gradients(params, group=group)
    if clip_grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(params, clip_grad_norm)
    optimizer.step()
    model.unet.invalidate_plan()
    model.decoder.invalidate_plan()
    if hasattr(model, "_plan"):
        model._plan = None
    return out
