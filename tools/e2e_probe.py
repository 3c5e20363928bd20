"""Where does the end-to-end (host buffers) arm spend its extra time?  wall-clock per phase."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import unidet3d_b200 as u
from unidet3d_b200.synthetic import make_model_state_dict
cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval(); model.load_state_dict(make_model_state_dict(cfg, 0), strict=False); model.cuda()
pts = [s[0] for s in scenes]; sps = [s[1] for s in scenes]
h_pts = [torch.as_tensor(p).pin_memory() for p in pts]; h_sps = [torch.as_tensor(s).pin_memory() for s in sps]
d_pts = [t.cuda() for t in h_pts]; d_sps = [t.cuda() for t in h_sps]; n_sps = [int(s.max()) + 1 for s in sps]
def wall(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("device inputs + n_sps :", round(wall(lambda: model.forward_scenes(d_pts, d_sps, names, n_sps)), 2), "ms")
print("device inputs         :", round(wall(lambda: model.forward_scenes(d_pts, d_sps, names)), 2), "ms")
print("pinned host inputs    :", round(wall(lambda: model.forward_scenes(h_pts, h_sps, names)), 2), "ms")
print("pinned host + n_sps   :", round(wall(lambda: model.forward_scenes(h_pts, h_sps, names, n_sps)), 2), "ms")
print("numpy inputs          :", round(wall(lambda: model.forward_scenes(pts, sps, names)), 2), "ms")
