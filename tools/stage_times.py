"""Stage-level device times of one forward step (CUDA events, warm), bench workload."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200 import ops  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict  # noqa: E402

cfg, scenes, names, preset = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
sps = [torch.as_tensor(s[1]).cuda() for s in scenes]
n_sps = [int(s[1].max()) + 1 for s in scenes]
B = len(pts)
for _ in range(3):
    model.forward_scenes(pts, sps, names, n_sps)
torch.cuda.synchronize()


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, float(np.median(ts))


P = torch.cat(pts)
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
sp_off = np.concatenate([[0], np.cumsum(n_sps)])
sp_b = torch.cat([s + int(o) for s, o in zip(sps, sp_off[:-1])])
(_, t_cent) = timed(lambda: ops.segmented_mean(P, sp_b, int(sp_off[-1]), channels=3))
(xi, t_collate) = timed(lambda: model.collate(P, offs, B))
x, inv = xi
(pooled, t_feat) = timed(lambda: model.extract_feat(x, sp_b, inv, sp_off))
cent = ops.segmented_mean(P, sp_b, int(sp_off[-1]), channels=3)
(out, t_enc) = timed(lambda: model.decoder.forward_packed(pooled, cent, [int(v) for v in sp_off], names))
pt_off = np.cumsum([0] + [len(p) for p in pts])


def post():
    res = []
    for i in range(B):
        a, b = int(pt_off[i]), int(pt_off[i + 1])
        res.append(model.predict_by_feat_scene(out["cls_preds"][i], out["bboxes"][i], P[a:b], sps[i], n_sps[i], 0))
    return res


(_, t_post) = timed(post)
ds_idx = [0] * B
(_, t_post_ms) = timed(lambda: model.postprocess_batch(out, P, sp_b, pt_off, sp_off, n_sps, ds_idx))
(_, t_all) = timed(lambda: model.forward_scenes(pts, sps, names, n_sps))
print(f"sp_centers {t_cent:.3f} ms | collate+rulebooks {t_collate:.3f} | input conv+unet+pool {t_feat:.3f} | encoder {t_enc:.3f} | "
      f"post (serial, 1 stream) {t_post:.3f} | post (8 streams) {t_post_ms:.3f} | whole forward_scenes {t_all:.3f} ms")
