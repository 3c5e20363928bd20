"""One forward step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--workload scannet_b8]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="scannet_b8")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--backbone-only", action="store_true")
    a = ap.parse_args()
    import unidet3d_b200 as u
    from unidet3d_b200.synthetic import make_model_state_dict
    cfg, scenes, names, preset = bench.make_workload(a.workload, 0)
    model = u.MODELS.build(cfg).eval()
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.cuda()
    pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
    sps = [torch.as_tensor(s[1]).cuda() for s in scenes]
    n_sps = [int(s[1].max()) + 1 for s in scenes]
    for _ in range(a.warmup):
        model.forward_scenes(pts, sps, names, n_sps)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.forward_scenes(pts, sps, names, n_sps)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
