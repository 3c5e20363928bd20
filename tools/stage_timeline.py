"""GPU-timeline (CUDA events, no extra synchronisation) and host-timeline of the stages inside forward_scenes: shows where
the GPU waits for the host and where the host runs ahead."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict  # noqa: E402

cfg, scenes, names, preset = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
host = len(sys.argv) > 2 and sys.argv[2] == "host"
if host:
    pts = [torch.as_tensor(s[0]).pin_memory() for s in scenes]
    sps = [torch.as_tensor(s[1]).pin_memory() for s in scenes]
else:
    pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
    sps = [torch.as_tensor(s[1]).cuda() for s in scenes]
n_sps = [int(s[1].max()) + 1 for s in scenes]
for _ in range(3):
    model.forward_scenes(pts, sps, names, n_sps)
torch.cuda.synchronize()
G, H, W = [], [], []
for _ in range(10):
    model.stage_events = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    model.forward_scenes(pts, sps, names, n_sps)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ev = model.stage_events
    G.append([e0.elapsed_time(e) for _, e, _ in ev])
    H.append([(t - t0) * 1e3 for _, _, t in ev])
    W.append((t1 - t0) * 1e3)
model.stage_events = None
G, H = np.median(np.array(G), 0), np.median(np.array(H), 0)
print(f"inputs: {'pinned host' if host else 'device'}; whole call (host wall) {np.median(W):.3f} ms")
print("stage        host reaches at   GPU finishes at   (ms since call start)")
for (name, _, _), g, h in zip(ev, G, H):
    print(f"{name:12s} {h:12.3f} {g:16.3f}")
