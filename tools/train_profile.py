"""One training step of bench.py's train_b8 workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> python tools/train_profile.py
(the launch list of a training step: which kernels the step's time goes to).  Without ncu it prints the stage split."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "train_b8"
    import unidet3d_b200 as u
    from unidet3d_b200 import train
    from unidet3d_b200.structures import Det3DDataSample, InstanceData, PointData
    from unidet3d_b200.synthetic import make_model_state_dict, make_scannet_gt
    dev = torch.device("cuda", 0)
    cfg, scenes, names, preset = bench.make_workload(workload, 0)
    model = u.MODELS.build(cfg)
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05)
    P, samples = [], []
    for i, (pts, sp) in enumerate(scenes):
        labels, sp_masks, inst = make_scannet_gt(sp, bench.TRAIN_INSTANCES, i)
        t = [torch.as_tensor(a).to(dev) for a in (pts, sp, inst, labels, sp_masks)]
        P.append(t[0])
        samples.append(Det3DDataSample(lidar_path="data/scannet/points/x.bin", gt_pts_seg=PointData(sp_pts_mask=t[1], pts_instance_mask=t[2]),
                                       gt_instances_3d=InstanceData(labels_3d=t[3], sp_masks=t[4])))
    inputs = dict(points=P)
    for _ in range(2):
        train.train_step(model, opt, inputs, samples)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = train.train_step(model, opt, inputs, samples)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(json.dumps({"loss": float(out["det_loss"]), "stages_ms": train.profile_step(model, inputs, samples)}))
    if "--cprofile" in sys.argv:
        import cProfile
        import io
        import pstats
        import time
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.enable()
        train.train_step(model, opt, inputs, samples)
        pr.disable()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
        print(f"host issue time of one step {1e3 * (t1 - t0):.1f} ms (under cProfile), GPU drained {1e3 * (t2 - t1):.1f} ms later")
        print(buf.getvalue())


if __name__ == "__main__":
    main()
