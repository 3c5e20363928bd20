import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unidet3d_b200 import ops
from oracle import criterion as oc
DEV = "cuda"
rng = np.random.default_rng(3)
for dim, C, topk in [(6, 18, 6), (7, 17, 6)]:
    T, G = 3000, 40
    gt = np.concatenate([rng.uniform(0.5, 7.5, (G, 3)), rng.uniform(0.3, 2.0, (G, 3))] + ([rng.uniform(-3, 3, (G, 1))] if dim == 7 else []), 1).astype(np.float32)
    pb = np.concatenate([rng.uniform(0.5, 7.5, (T, 3)), rng.uniform(0.3, 2.0, (T, 3))] + ([rng.uniform(-3, 3, (T, 1))] if dim == 7 else []), 1).astype(np.float32)
    pb[:G * 8] = np.repeat(gt, 8, 0) + 0.08 * rng.standard_normal((G * 8, dim)).astype(np.float32)
    pb[:, 3:6] = np.abs(pb[:, 3:6]) + 0.05
    cls = (rng.standard_normal((T, C + 1)) * 2).astype(np.float32)
    labels = rng.integers(0, C, G)
    qm = rng.random((G, T)) < 0.3
    match, sums = ops.criterion_layer(torch.as_tensor(cls).to(DEV), torch.as_tensor(pb).to(DEV), torch.as_tensor(gt).to(DEV),
                                      torch.as_tensor(labels).to(DEV), torch.as_tensor(qm).to(DEV), topk, 0.5, 2.0, 0.1)
    cost = oc.match_cost(torch.as_tensor(cls), torch.as_tensor(pb), torch.as_tensor(labels), torch.as_tensor(gt))
    cost = torch.where(torch.as_tensor(qm).T, cost, torch.tensor(1e8))
    kth = torch.topk(cost, topk + 1, dim=0, largest=False).values
    ref = cost < kth[-1:]
    m = match.cpu()
    diff = torch.argwhere(m != ref)
    print(f"dim {dim}: gpu matches {int(m.sum())} ref {int(ref.sum())} differing {len(diff)}")
    for q, g in diff[:10].tolist():
        col = cost[:, g]
        print(f"  (q={q}, g={g}) gpu={bool(m[q, g])} ref={bool(ref[q, g])} cost={float(col[q]):.9g} kth values={[f'{float(x):.9g}' for x in kth[:, g]]}")
    per_col_gpu = m.sum(0).tolist(); per_col_ref = ref.sum(0).tolist()
    print("  per-column counts gpu", per_col_gpu[:20], "ref", per_col_ref[:20])
