"""Where the host spends its time in forward_pipelined: per batch, the wall time of submit_scenes and the time blocked in
collect(), for device-resident and pinned-host inputs (same loop as bench.py's run_throughput)."""
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

cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
n_sps = [int(s.max()) + 1 for s in sps]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
K = 30
t_sub, t_col = [], []
orig_submit, orig_collect = model.submit_scenes, model.collect


def submit(*a, **k):
    t = time.perf_counter(); r = orig_submit(*a, **k); t_sub.append(time.perf_counter() - t); return r


def collect(h):
    t = time.perf_counter(); r = orig_collect(h); t_col.append(time.perf_counter() - t); return r


model.submit_scenes, model.collect = submit, collect
parts = {}


def timed_method(obj, name, key):
    orig = getattr(obj, name)

    def f(*a, **k):
        t = time.perf_counter(); r = orig(*a, **k); parts.setdefault(key, []).append(time.perf_counter() - t); return r
    setattr(obj, name, f)


timed_method(model, "collate", "collate")
timed_method(model, "extract_feat", "backbone")
timed_method(model.decoder, "forward_packed", "encoder")
timed_method(model, "postprocess_batch", "post")
from unidet3d_b200 import rulebook as _rb  # noqa: E402
_orig_bp = _rb.build_pyramid
import unidet3d_b200.detector as _det  # noqa: E402


def _bp(*a, **k):
    t = time.perf_counter(); r = _orig_bp(*a, **k); parts.setdefault("  build_pyramid (in collate)", []).append(time.perf_counter() - t); return r


_det.build_pyramid = _bp
for label, P, S, nsp in (("device", [torch.as_tensor(p).cuda() for p in pts], [torch.as_tensor(s).cuda() for s in sps], n_sps),
                         ) + () if os.environ.get("ALL") != "1" else (("pinned host", [torch.as_tensor(p).pin_memory() for p in pts], [torch.as_tensor(s).pin_memory() for s in sps], None),
                         ("pinned host, n_superpoints given", [torch.as_tensor(p).pin_memory() for p in pts], [torch.as_tensor(s).pin_memory() for s in sps], n_sps)):
    for depth in (2, 3):
        for _ in model.forward_pipelined(((P, S, names, nsp) for _ in range(6)), depth=depth):
            pass
        torch.cuda.synchronize()
        t_sub.clear(); t_col.clear(); parts.clear()
        t0 = time.perf_counter()
        for _ in model.forward_pipelined(((P, S, names, nsp) for _ in range(K)), depth=depth, pre_submit=flush.zero_):
            pass
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / K * 1e3
        print(f"{label:34s} depth {depth}: {wall:6.3f} ms/batch wall; submit {np.mean(t_sub) * 1e3:6.3f} ms (max {np.max(t_sub) * 1e3:6.3f}), "
              f"blocked in collect {np.mean(t_col) * 1e3:6.3f} ms", flush=True)
        print("      " + ", ".join(f"{k} {np.mean(v) * 1e3:.3f}" for k, v in parts.items()), flush=True)
