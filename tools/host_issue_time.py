"""Host issue time vs GPU time of the stages of one forward step: a stage whose host time is close to its GPU time is
launch-bound (the GPU waits for Python)."""
import os
import sys
import time

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
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
allp = torch.cat(pts)
x, inv = model.collate(allp, offs, B)
torch.cuda.synchronize()


def stage(name, fn, reps=5):
    hs, gs, ls = [], [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        ops.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        r = fn()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        hs.append((t1 - t0) * 1e3)
        gs.append(e0.elapsed_time(e1))
        ls.append(ops.launch_count())
    print(f"{name:28s} host issue {np.median(hs):7.3f} ms   gpu {np.median(gs):7.3f} ms   launches {int(np.median(ls))}")
    return r


sp_off = np.concatenate([[0], np.cumsum(n_sps)]).astype(np.int64)
pt_off = np.concatenate([[0], np.cumsum([len(p) for p in pts])]).astype(np.int64)
sp_b = torch.cat([s + int(o) for s, o in zip(sps, sp_off[:-1])])
ds_idx = [model.decoder.datasets.index(n) for n in names]
sp_centers = ops.segmented_mean(allp, sp_b, int(sp_off[-1]), channels=3)
stage("collate+rulebooks", lambda: model.collate(allp, offs, B))
pooled = stage("input conv + unet + pool", lambda: model.extract_feat(x, sp_b, inv, sp_off))
out = stage("encoder", lambda: model.decoder.forward_packed(pooled, sp_centers, [int(v) for v in sp_off], names))
stage("post (8 streams)", lambda: model.postprocess_batch(out, allp, sp_b, pt_off, sp_off, n_sps, ds_idx))
stage("whole forward_scenes", lambda: model.forward_scenes(pts, sps, names, n_sps))
