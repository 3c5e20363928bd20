"""cProfile of the host side of forward_scenes (which Python frames the launch-bound stages spend their time in)."""
import cProfile
import io
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict  # noqa: E402

cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
sps = [torch.as_tensor(s[1]).cuda() for s in scenes]
n_sps = [int(s[1].max()) + 1 for s in scenes]
for _ in range(3):
    model.forward_scenes(pts, sps, names, n_sps)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    model.forward_scenes(pts, sps, names, n_sps)
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
