"""Oracle: UniDet3DEncoder forward (reference: unidet3d/encoder.py:1-283), PyTorch CPU.

Functional restatement over a reference-layout ``state_dict``
(``input_proj.{0,2}``, ``self_attn_layers.{i}.attn.{in_proj_weight,in_proj_bias,
out_proj.weight,out_proj.bias}``, ``self_attn_layers.{i}.norm``,
``ffn_layers.{i}.net.{0,3}``, ``ffn_layers.{i}.norm``, ``out_norm``,
``outs_cls.{0,2}``, ``out_bboxes.linear``).  Pinned against the reference module
itself by tests/golden/encoder_ref.npz (see tests/golden/make_golden.py).
"""
import itertools
import math

import torch
import torch.nn.functional as F


def class_index_lists(datasets_classes):
    """encoder.py:151-161: union class list (+ 'no_obj') and per-dataset column ids."""
    unique_cls = sorted(list(set(itertools.chain.from_iterable(datasets_classes)))) + ["no_obj"]
    idxs = []
    for classes in datasets_classes:
        idxs.append([unique_cls.index(c) for c in classes] + [len(unique_cls) - 1])
    return unique_cls, idxs


def bbox_pred_to_bbox(points, bbox_pred):
    """encoder.py:241-283."""
    if bbox_pred.shape[0] == 0:
        return bbox_pred
    xc = points[:, 0] + (bbox_pred[:, 1] - bbox_pred[:, 0]) / 2
    yc = points[:, 1] + (bbox_pred[:, 3] - bbox_pred[:, 2]) / 2
    zc = points[:, 2] + (bbox_pred[:, 5] - bbox_pred[:, 4]) / 2
    base = torch.stack([xc, yc, zc, bbox_pred[:, 0] + bbox_pred[:, 1],
                        bbox_pred[:, 2] + bbox_pred[:, 3], bbox_pred[:, 4] + bbox_pred[:, 5]], -1)
    if bbox_pred.shape[1] == 6:
        return base
    scale = bbox_pred[:, 0] + bbox_pred[:, 1] + bbox_pred[:, 2] + bbox_pred[:, 3]
    q = torch.exp(torch.sqrt(bbox_pred[:, 6] ** 2 + bbox_pred[:, 7] ** 2))
    alpha = 0.5 * torch.atan2(bbox_pred[:, 6], bbox_pred[:, 7])
    return torch.stack((xc, yc, zc, scale / (1 + q), scale / (1 + q) * q,
                        bbox_pred[:, 5] + bbox_pred[:, 4], alpha), dim=-1)


def self_attention(sd, p, x, num_heads):
    """encoder.py:24-41 (nn.MultiheadAttention on an unbatched [T,d] input, post-norm)."""
    d = x.shape[1]
    hd = d // num_heads
    qkv = x @ sd[p + ".attn.in_proj_weight"].t() + sd[p + ".attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=1)
    T = x.shape[0]
    q = q.view(T, num_heads, hd).transpose(0, 1)
    k = k.view(T, num_heads, hd).transpose(0, 1)
    v = v.view(T, num_heads, hd).transpose(0, 1)
    attn = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(hd), dim=-1)
    o = (attn @ v).transpose(0, 1).reshape(T, d)
    z = o @ sd[p + ".attn.out_proj.weight"].t() + sd[p + ".attn.out_proj.bias"]
    z = z + x
    return F.layer_norm(z, (d,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)


def ffn(sd, p, x, activation_fn="gelu"):
    """encoder.py:63-80."""
    h = x @ sd[p + ".net.0.weight"].t() + sd[p + ".net.0.bias"]
    h = torch.relu(h) if activation_fn == "relu" else F.gelu(h)
    z = h @ sd[p + ".net.3.weight"].t() + sd[p + ".net.3.bias"]
    z = z + x
    return F.layer_norm(z, (x.shape[1],), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5)


def forward_head(sd, feats, sp_centers, class_idxs, angle):
    """encoder.py:165-201 for one scene."""
    d = feats.shape[1]
    nq = F.layer_norm(feats, (d,), sd["out_norm.weight"], sd["out_norm.bias"], 1e-5)
    h = torch.relu(nq @ sd["outs_cls.0.weight"].t() + sd["outs_cls.0.bias"])
    cls = (h @ sd["outs_cls.2.weight"].t() + sd["outs_cls.2.bias"])[:, torch.as_tensor(class_idxs)]
    b = nq @ sd["out_bboxes.linear.weight"].t() + sd["out_bboxes.linear.bias"]
    b = torch.hstack((torch.exp(b[:, :6]), b[:, 6:]))          # PredBBox, encoder.py:109-111
    if not angle:
        b = b[:, :6]
    return cls, bbox_pred_to_bbox(sp_centers, b)


def encoder_forward(sd, cfg, x, sp_centers, datasets_names, all_heads=True):
    """encoder.py:203-239.  cfg: dict(num_layers,num_heads,activation_fn,datasets,
    datasets_classes,angles).  Returns dict(cls_preds, bboxes, aux_outputs)."""
    _, cls_idxs = class_index_lists(cfg["datasets_classes"])
    ds = [cfg["datasets"].index(n) for n in datasets_names]
    feats = [torch.relu(y @ sd["input_proj.0.weight"].t() + sd["input_proj.0.bias"])
             @ sd["input_proj.2.weight"].t() + sd["input_proj.2.bias"] for y in x]

    def heads(fs):
        out = [forward_head(sd, f, c, cls_idxs[j], cfg["angles"][j]) for f, c, j in zip(fs, sp_centers, ds)]
        return [o[0] for o in out], [o[1] for o in out]

    cls_preds, bboxes = [], []
    if all_heads:
        c, b = heads(feats)
        cls_preds.append(c), bboxes.append(b)
    for i in range(cfg["num_layers"]):
        feats = [self_attention(sd, f"self_attn_layers.{i}", f, cfg["num_heads"]) for f in feats]
        feats = [ffn(sd, f"ffn_layers.{i}", f, cfg.get("activation_fn", "gelu")) for f in feats]
        if all_heads or i == cfg["num_layers"] - 1:
            c, b = heads(feats)
            cls_preds.append(c), bboxes.append(b)
    aux = [dict(cls_preds=c, bboxes=b) for c, b in zip(cls_preds[:-1], bboxes[:-1])]
    return dict(cls_preds=cls_preds[-1], bboxes=bboxes[-1], aux_outputs=aux, feats=feats)


from unidet3d_b200.synthetic import make_encoder_state_dict  # noqa: E402,F401  (shared synthetic weights)
