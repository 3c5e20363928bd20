"""Oracle: sparse convolution forward + BN(eval)+ReLU (PyTorch CPU fp32).

Restates spconv 2.3.6 conv forward (third-party, absent)
as used by unidet3d/spconv_unet.py:37-72,148-191 and unidet3d/unidet3d.py:96-103:
``out[o,:] = sum_k in[table[k,o],:] @ W_k`` with ``W_k = weight[:, kx,ky,kz, :].T``
(weight parameter layout ``[C_out, k0, k1, k2, C_in]``), fp32, no bias.
This gather -> mm -> index_add form is also the CPU baseline timed by bench.py.
"""
import torch


def weight_to_koc(weight):
    """[C_out, k0,k1,k2, C_in] -> [K, C_in, C_out] (K row-major over k0,k1,k2)."""
    co, k0, k1, k2, ci = weight.shape
    return weight.reshape(co, k0 * k1 * k2, ci).permute(1, 2, 0).contiguous()


def sparse_conv(feats, table, w_koc, n_out=None):
    """feats [N_in, C_in], table int [K, N_out], w_koc [K, C_in, C_out] -> [N_out, C_out]."""
    table = torch.as_tensor(table, dtype=torch.long)
    K, n = table.shape
    n_out = n if n_out is None else n_out
    out = feats.new_zeros((n_out, w_koc.shape[2]))
    for k in range(K):
        idx = table[k]
        m = idx >= 0
        if m.any():
            rows = m.nonzero(as_tuple=True)[0]
            out.index_add_(0, rows, feats[idx[rows]] @ w_koc[k])
    return out


TRAIN_MODE = False      # True: batch statistics (momentum 0.1, running statistics updated in place), like module.train()


def bn_relu(x, bn, eps=1e-4, relu=True):
    """BatchNorm1d(eps=1e-4, momentum=0.1) + ReLU (spconv_unet.py:119-124, unidet3d.py:104-111): eval mode by default,
    train mode (statistics over all rows = all active voxels of the batch) when ``TRAIN_MODE`` is set."""
    w, b, mean, var = bn
    y = torch.nn.functional.batch_norm(x, mean, var, w, b, TRAIN_MODE, 0.1 if TRAIN_MODE else 0.0, eps)
    return torch.relu(y) if relu else y
