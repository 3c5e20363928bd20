"""Oracle: ``torch_scatter.scatter_mean`` (2.1.2; third-party, absent -- SURVEY.md
appendix A3) at the reference call sites unidet3d/unidet3d.py:130 (superpoint
feature pooling) and :446-447 (superpoint centres): rows = max(index)+1,
out = sum / clamp(count, 1); ids that never occur -> zero rows."""
import torch


def scatter_mean(src, index, n_rows=None):
    index = torch.as_tensor(index, dtype=torch.long)
    n = int(index.max()) + 1 if n_rows is None else n_rows
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    out.index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


def superpoint_pool(vox_feats, inverse_mapping, superpoints, n_rows=None):
    """unidet3d.py:130: scatter_mean(x.features[inverse_mapping], superpoints, dim=0)."""
    inv = torch.as_tensor(inverse_mapping, dtype=torch.long)
    return scatter_mean(vox_feats[inv], superpoints, n_rows)
