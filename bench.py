#!/usr/bin/env python
"""bench.py -- scenes/sec of the UniDet3D forward hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload scannet_b8|s3dis_1|joint_b8]

A "step" is one pass of the hot path (voxelise -> input conv -> SpConvUNet -> superpoint pool ->
encoder -> last head -> top-k -> NMS -> superpoint trim) over ONE BATCH of synthetic scenes.
Default workload = BASELINE.json configs[1]: unidet3d_1xb8_scannet, 8 synthetic 100k-point
ScanNet-shaped scenes per batch (voxel 0.02 m, ~33k voxels/scene, 19-way head), random-init
weights of the reference architecture (no datasets / checkpoints offline).

* ``value``    : scenes/s with the batch already resident in HBM (device-timed, CUDA events): EXACTLY K batches through
                 ``UniDet3D.forward_pipelined`` (``--pipeline-depth`` batches in flight, default 2: the host-bound
                 voxelisation / launch work of batch i+1 overlaps the kernels of batch i);
* ``e2e``      : the same loop through the same public API with HOST buffers -- pinned H2D of points+superpoints
                 and D2H of the detections of every step inside the timed region;
                 ``config.latency_ms_per_batch`` = one batch at a time (``forward_scenes``), both input kinds;
* ``roofline`` : the dominant kernel (tcgen05 gather-GEMM, the 49 sparse convs): algorithmic bytes
                 (BASELINE.md section 3 model) / measured launch time vs measured HBM peak;
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle (oracle/detector.py) on the host cores.
Multi-GPU: scenes shard one batch per GPU, no data-path collective (weak scaling); one process per
GPU under torchrun, device-timed, max over ranks.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, batch, datasets)
    "scannet_b8": ("scannet100k", 8, ("scannet",)),
    "s3dis_1": ("s3dis500k", 1, ("scannet",)),
    "small_1": ("small20k", 1, ("scannet",)),
    # BASELINE.json configs[3]: training step (fwd + bwd + matcher + criterion + AdamW), batch 8 per GPU
    "train_b8": ("scannet100k", 8, ("scannet",)),
    "train_small": ("small20k", 2, ("scannet",)),
}
TRAIN_INSTANCES = 24      # GT instances per synthetic scene (ScanNet scenes hold ~10-40)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="scannet_b8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=2,
                    help="batches in flight in the timed loops (UniDet3D.forward_pipelined); 1 = one batch at a time")
    return ap.parse_args()


def host_threads():
    """Threads for the CPU arm: the cores this process may run on, capped at 32 -- beyond that the
    oracle's many small torch ops (per-offset mm / index_add_) slow down from OpenMP fork/join on the
    shared 128-core GPU hosts (measured: 128 threads -> ~1000x slower than 8)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    return max(1, min(n, int(os.environ.get("UD3D_CPU_THREADS", 32))))


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def make_workload(name, rank):
    from unidet3d_b200 import configs
    from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS
    if name == "joint_b8":
        preset, batch, datasets = "scannet100k", 8, configs.JOINT
    else:
        preset, batch, datasets = WORKLOADS[name]
    n, voxel, area, sp_cell = SCENE_PRESETS[preset]
    scenes = [make_scene(rank * batch + i, n, area, sp_cell) for i in range(batch)]
    cfg = configs.model_cfg(datasets, voxel_size=voxel)
    names = [datasets[i % len(datasets)] for i in range(batch)]
    return cfg, scenes, names, preset


class ClockSampler:
    """nvidia-smi clocks/throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/ud3d_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])), mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def conv_work_model(pyr, planes, in_channels=6):
    """Algorithmic FLOPs / bytes of the 49 sparse convs for THIS input:
    FLOPs = 2*P*Cin*Cout; bytes = 4*Nin*Cin + 4*Nout*Cout + 4*K*Cin*Cout + 4*K*Nout (tables)."""
    flops = bytes_ = 0
    launches = 0
    L = len(planes)

    def add(P, n_in, n_out, cin, cout, K):
        nonlocal flops, bytes_, launches
        flops += 2 * P * cin * cout
        bytes_ += 4 * n_in * cin + 4 * n_out * cout + 4 * K * cin * cout + 4 * K * n_out
        launches += 1

    Ps = [int((lv.subm >= 0).sum().item()) for lv in pyr.levels]
    Pd = [int((lv.child >= 0).sum().item()) if lv.child is not None else 0 for lv in pyr.levels]
    n = [lv.n for lv in pyr.levels]
    add(Ps[0], n[0], n[0], in_channels, planes[0], 27)
    for l in range(L):
        c = planes[l]
        for _ in range(4):
            add(Ps[l], n[l], n[l], c, c, 27)                     # blocks: 2 x (SubM3, SubM3)
        if l < L - 1:
            c1 = planes[l + 1]
            add(Pd[l], n[l], n[l + 1], c, c1, 8)                 # down
            add(Pd[l], n[l + 1], n[l], c1, c, 8)                 # up (same pairs reversed)
            add(n[l], n[l], n[l], 2 * c, c, 1)                   # tail block0 i_branch (1x1)
            add(Ps[l], n[l], n[l], 2 * c, c, 27)                 # tail block0 conv0
            for _ in range(3):
                add(Ps[l], n[l], n[l], c, c, 27)                 # tail block0 conv1, block1 x2
    return flops, bytes_, launches


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of the reference forward on the host cores (the reference's own
    native stack -- spconv / MinkowskiEngine / mmcv -- cannot be installed offline)."""
    if rank != 0:
        return
    if args.workload.startswith("train"):
        return run_reference_train(args)
    from oracle import detector as odet
    from unidet3d_b200 import configs
    from unidet3d_b200.synthetic import make_model_state_dict
    cfg, scenes, names, preset = make_workload(args.workload, 0)
    torch.set_num_threads(host_threads())
    sd = make_model_state_dict(cfg, 0)
    det_sd = {k: v for k, v in sd.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    ocfg = configs.oracle_cfg(cfg)
    sample = 1                                  # scenes per step: bounded sample of the batch
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    times = []
    for it in range(args.warmup + args.steps):
        i = it % len(scenes)
        t0 = time.perf_counter()
        odet.forward_scenes(det_sd, enc_sd, ocfg, pts[i:i + sample], sps[i:i + sample], names[i:i + sample])
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    val = sample * len(times) / total
    line = {"impl": "reference", "metric": "scenes/sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "preset": preset, "sample": f"{sample} scene/step (bounded sample of the batch)"},
            "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(times)} x {sample} scene of {preset}"},
            "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_reference_train(args):
    """CPU arm of the training workload: one training step (train-mode forward, matcher, criterion, torch.autograd
    backward, AdamW) of the oracle restatement on ONE scene per step on the host cores."""
    from oracle import criterion as oc, encoder as oenc, spconv as ospconv, unet as ounet, voxelize as ovox
    from oracle.pool import scatter_mean, superpoint_pool
    from unidet3d_b200 import configs
    from unidet3d_b200.synthetic import make_model_state_dict, make_scannet_gt
    cfg, scenes, names, preset = make_workload(args.workload, 0)
    torch.set_num_threads(host_threads())
    sd = {k: t.clone().float() for k, t in make_model_state_dict(cfg, 0).items()}
    params = [t.requires_grad_(True) for k, t in sd.items()
              if t.is_floating_point() and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    det_sd = {k: v for k, v in sd.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    ocfg = configs.oracle_cfg(cfg)
    cc = cfg["criterion"]
    ccfg = dict(datasets=list(cfg["decoder"]["datasets"]), datasets_weights=cc["datasets_weights"], topk=cc["topk"],
                loss_weight=cc["loss_weight"], non_object_weight=cc["non_object_weight"], iter_matcher=True)
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05)
    times = []
    for it in range(args.warmup + args.steps):
        pts, sp = scenes[it % len(scenes)]
        labels, sp_masks, inst = make_scannet_gt(sp, TRAIN_INSTANCES, it % len(scenes))
        t0 = time.perf_counter()
        xyz = pts[:, :3] - pts[:, :3].min(0)
        gt = dict(labels=torch.as_tensor(labels), boxes=oc.bboxes_by_masks(inst, xyz), query_masks=torch.as_tensor(sp_masks))
        coords, feats, inverse, shape = ovox.voxelize([pts], ocfg["voxel_size"], ocfg["min_spatial_shape"])
        ospconv.TRAIN_MODE = True
        try:
            x, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
        finally:
            ospconv.TRAIN_MODE = False
        pooled = superpoint_pool(x, inverse, sp, int(sp.max()) + 1)
        ctr = scatter_mean(torch.as_tensor(xyz), torch.as_tensor(sp))
        pred = oenc.encoder_forward(enc_sd, ocfg["encoder"], [pooled], [ctr], names[:1], all_heads=True)
        loss = oc.criterion(pred, [gt], names[:1], ccfg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        opt.step()
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    val = len(times) / total
    line = {"impl": "reference", "metric": "training scenes/sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "preset": preset, "sample": "1 scene/step (bounded sample of the batch)"},
            "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(times)} training steps x 1 scene of {preset}"},
            "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_train(args, rank, world, local):
    """BASELINE.json configs[3]: one training step per "step" -- ``train.train_step``: collate, train-mode backbone
    ((Sync)BatchNorm batch statistics), pooling, encoder with seven heads, GPU matcher + criterion, backward through all of
    it, gradient all-reduce over NCCL (N > 1), gradient clipping, AdamW -- on a batch of 8 synthetic 100k-point scenes
    per GPU."""
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    import unidet3d_b200 as u
    from unidet3d_b200 import ops, sharding, train
    from unidet3d_b200.structures import Det3DDataSample, InstanceData, PointData
    from unidet3d_b200.synthetic import make_model_state_dict, make_scannet_gt
    cfg, scenes, names, preset = make_workload(args.workload, rank)
    model = u.MODELS.build(cfg)
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05)     # configs/unidet3d_1xb8_scannet.py optim_wrapper
    batch = len(scenes)
    W, K = max(args.warmup, 3), args.steps
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def put(t, where):
        t = torch.as_tensor(t)
        return t.to(dev) if where == "device" else t.pin_memory()

    def make_inputs(where):
        P, samples, nbytes = [], [], 0
        for i, (pts, sp) in enumerate(scenes):
            labels, sp_masks, inst = make_scannet_gt(sp, TRAIN_INSTANCES, rank * batch + i)
            ts = [put(pts, where), put(sp, where), put(inst, where), put(labels, where), put(sp_masks, where)]
            nbytes += sum(t.numel() * t.element_size() for t in ts)
            P.append(ts[0])
            samples.append(Det3DDataSample(lidar_path="data/scannet/points/x.bin",
                                           gt_pts_seg=PointData(sp_pts_mask=ts[1], pts_instance_mask=ts[2]),
                                           gt_instances_3d=InstanceData(labels_3d=ts[3], sp_masks=ts[4])))
        return dict(points=P), samples, nbytes

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def run(inputs, samples, steps, read_loss):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        for _ in range(steps):
            flush_buf.zero_()
            out = train.train_step(model, opt, inputs, samples, group=group)
            last = float(out["det_loss"]) if read_loss else out["det_loss"]
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / 1e3, float(last)

    d_in, d_samples, _ = make_inputs("device")
    run(d_in, d_samples, W, False)
    gc.collect()
    gc.freeze()        # see main(): long-lived objects leave the collector's generations after the warm-up
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    ops.launch_count(reset=True)
    t_dev, loss_dev = run(d_in, d_samples, K, False)
    launches = ops.launch_count()
    h_in, h_samples, h2d = make_inputs("host")
    run(h_in, h_samples, 1, True)
    t_e2e, loss_e2e = run(h_in, h_samples, K, True)
    clocks = sampler.stop() if rank == 0 else None

    # ---- stage split of one step (events around the stages of train.loss_backward, one extra untimed-region step)
    stages = {}
    try:
        stages = train.profile_step(model, d_in, d_samples, group=group)
    except Exception as e:  # noqa: BLE001
        stages = {"error": repr(e)[:200]}
    # ---- dominant kernels: the sparse convolutions, forward + input gradient + weight gradient
    offs = torch.tensor(np.cumsum([0] + [len(s[0]) for s in scenes]), dtype=torch.int32, device=dev)
    with torch.no_grad():
        x, inverse = model.collate(torch.cat(d_in["points"]), offs, batch)
    flops, abytes, n_conv = conv_work_model(x.pyramid, cfg["backbone"]["num_planes"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roof = None
    if "backbone_fwd_ms" in stages and "backbone_bwd_ms" in stages:
        t_bb = (stages["backbone_fwd_ms"] + stages["backbone_bwd_ms"]) / 1e3
        # forward: read rows + write rows (conv_work_model); input gradient: the same traffic over the transposed rulebook;
        # weight gradient: reads both feature maps again -> ~3x the forward's algorithmic bytes
        achieved = 3 * abytes / t_bb / 1e9
        roof = {"bound": "hbm", "kernel": "sparse convolutions of the backbone: forward (gather_gemm_tc_kernel), input gradient "
                                          "(the same kernel over the transposed rulebook) and weight gradient (conv_wgrad_kernel)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "algorithmic_bytes_per_step": 3 * abytes, "backbone_fwd_bwd_ms": 1e3 * t_bb,
                "note": "time = CUDA events around the backbone's forward and backward stages of one step (BatchNorm, pooling "
                        "and residual kernels included); the backward kernels are correctness-first (weight gradient on the CUDA cores)"}
    t_dev, t_e2e = sharding.aggregate_times([t_dev, t_e2e], device=dev)
    if rank == 0:
        value = sharding.whole_job_throughput(batch * K, world, t_dev)
        e2e = sharding.whole_job_throughput(batch * K, world, t_e2e)
        n_param = sum(p.numel() for p in model.parameters())
        line = {"metric": "training scenes/sec", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "preset": preset, "batch_per_gpu": batch, "points_per_scene": int(scenes[0][0].shape[0]),
                           "voxels_per_level": [lv.n for lv in x.pyramid.levels], "gt_instances_per_scene": TRAIN_INSTANCES,
                           "datasets": list(cfg["decoder"]["datasets"]),
                           "step": "zero_grad, train-mode forward (batch-statistics BatchNorm), 7 heads, GPU matcher + criterion, "
                                   "backward (criterion, encoder, backbone), gradient all-reduce, clip 10, AdamW(lr 1e-4, wd 0.05)",
                           "collectives": ("none (1 rank)" if world == 1 else
                                           f"NCCL: SyncBatchNorm sums (45 forward + 45 backward all-reduces of 2C+1 doubles) and "
                                           f"{4 * n_param / 1e6:.1f} MB of gradients in 32 MB buckets"),
                           "precision": "fp32 storage; forward convs / Linear layers bf16 hi/lo 3-term split on tcgen05; backward fp32",
                           "l2": "flushed before every step (256 MiB memset inside the timed region)",
                           "parallelism": f"data-parallel dp{world}", "loss_last_step": loss_dev, "stages_ms": stages},
                "e2e": {"value": e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": 1e3 * t_e2e / K, "loss_last_step": loss_e2e},
                "gpu_launches": launches, "roofline": roof, "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, scenes, names, preset, args.workload, budget_s=400)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    args = parse()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (unidet3d_b200 has no CPU path; use --impl reference for the CPU arm)")
    if args.workload.startswith("train"):
        run_train(args, rank, world, local)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import unidet3d_b200 as u
    from unidet3d_b200 import ops
    from unidet3d_b200.synthetic import make_model_state_dict

    cfg, scenes, names, preset = make_workload(args.workload, rank)
    model = u.MODELS.build(cfg).eval()
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.to(dev)
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    batch = len(scenes)
    W, K = max(args.warmup, 3), args.steps

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def flush_l2():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    depth = max(1, args.pipeline_depth)

    def run_latency(P, S, n_sps_arg, reps):
        """one batch at a time (submit -> wait -> next): ms per batch"""
        evs = []
        for _ in range(reps):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.forward_scenes(P, S, names, n_sps_arg)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / reps

    def run_throughput(P, S, n_sps_arg):
        """EXACTLY K batches through the public pipelined API (`depth` batches in flight, each on its own stream; the L2
        flush of every step is issued on the step's stream inside the timed region); device time from the first
        submit to the last result, seconds"""
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        n_done = 0
        for res in model.forward_pipelined(((P, S, names, n_sps_arg) for _ in range(K)), depth=depth, pre_submit=flush_l2):
            n_done += 1
        for st in model._pipe_streams:
            cur.wait_stream(st)
        e1.record()
        barrier()
        assert n_done == K
        return e0.elapsed_time(e1) / 1e3

    # ---------------------------------------------------------------- device-resident arm (value)
    d_pts = [torch.as_tensor(p).to(dev) for p in pts]
    d_sps = [torch.as_tensor(s).to(dev) for s in sps]
    n_sps = [int(s.max()) + 1 for s in sps]
    for _ in range(W):
        model.forward_scenes(d_pts, d_sps, names, n_sps)
    for _ in model.forward_pipelined(((d_pts, d_sps, names, n_sps) for _ in range(2 * depth)), depth=depth):
        pass
    barrier()
    gc.collect()
    gc.freeze()        # the model / plans / cached buffers leave the collector's generations (what a serving process does
    #                    after start-up): later collections only scan the step's own short-lived objects
    lat_dev = run_latency(d_pts, d_sps, n_sps, max(3, min(K, 5)))
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    ops.launch_count(reset=True)
    t_dev = run_throughput(d_pts, d_sps, n_sps)
    launches = ops.launch_count()

    # ---------------------------------------------------------------- end-to-end arm (host buffers)
    h_pts = [torch.as_tensor(p).pin_memory() for p in pts]
    h_sps = [torch.as_tensor(s).pin_memory() for s in sps]
    for _ in range(2):
        model.forward_scenes(h_pts, h_sps, names)
    barrier()
    t_e2e = run_throughput(h_pts, h_sps, None)
    d2h = model.last_d2h_bytes       # packed per-scene result buffers + the small count / extent read-backs
    clocks = sampler.stop() if rank == 0 else None
    lat_e2e = run_latency(h_pts, h_sps, None, max(3, min(K, 5)))
    h2d = model.last_h2d_bytes       # points fp32 [n,6] + superpoint ids int64 [n] + scene offsets

    # ---------------------------------------------------------------- dominant kernel: the 49 sparse convs
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=dev)
    x, inverse = model.collate(torch.cat(d_pts), offs, batch)
    planes = cfg["backbone"]["num_planes"]
    flops, abytes, n_conv = conv_work_model(x.pyramid, planes)
    plan = model._get_plan()
    lv0 = x.pyramid.levels[0]

    def backbone_only():
        bn0 = model.unet.first_bn()
        f_act = torch.empty((lv0.n, plan["w_in"].c_out), dtype=torch.float32, device=dev)
        vox_s = ops.act_split(x.features, relu=False)
        tb, tm, pm = lv0.subm_conv
        f = ops.gemm(vox_s, plan["w_in"], table=tb, tile_mask=tm, in_split=True, acts=[(f_act, bn0[0], bn0[1])], row_perm=pm)
        y = x.replace_feature(f)
        y.features_act = f_act
        return model.unet(y)

    for _ in range(3):
        backbone_only()
    torch.cuda.synchronize()
    ops.launch_count(reset=True)
    reps = max(3, min(K, 10))
    evs = []
    for _ in range(reps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        backbone_only()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    conv_launches = ops.launch_count() // reps
    t_conv = sum(a.elapsed_time(b) for a, b in evs) / 1e3 / reps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # per-launch figures are over the n_conv gather-GEMM launches (the backbone sequence also contains a few tiny
    # act_split launches after split-K layers; their time is inside t_conv, i.e. charged to the conv kernel)
    achieved = (abytes / n_conv) / (t_conv / n_conv) / 1e9
    traffic, tinfo = None, {}
    try:   # DRAM bytes per launch of the same kernel from the committed ncu capture (profiles/)
        tfile = [f for f in ("r4_gemm_traffic.json", "r2_gemm_traffic.json") if os.path.exists(os.path.join(ROOT, "profiles", f))][0]
        tinfo = json.load(open(os.path.join(ROOT, "profiles", tfile)))
        traffic = tinfo["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    roof = {"bound": "hbm", "kernel": "gather_gemm_tc_kernel (input conv + the 49 sparse convs of the U-Net)",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
            "traffic_read": tinfo.get("dram_read_bytes_per_launch"), "traffic_write": tinfo.get("dram_write_bytes_per_launch"),
            "l2_to_sm_bytes_per_launch": tinfo.get("l2_to_sm_bytes_per_launch"),
            "peak_source": peak_src, "launches_per_step": n_conv, "avg_launch_us": 1e6 * t_conv / n_conv,
            "algorithmic_bytes_per_launch": abytes / n_conv,
            "algorithmic_tflops": flops / t_conv / 1e12,
            "tensor_frac_bf16x3": 3 * flops / t_conv / 1e12 / tf_peak,
            "backbone_ms_per_batch": 1e3 * t_conv,
            "note": "HBM is the roof the contract asks for, but not what binds this kernel: each input row is re-read from "
                    "L2 once per active kernel offset (L2->SM bytes ~7x the algorithmic bytes, profiles/r4_gemm_traffic.json) "
                    "and the main loop runs at ~45 B/cycle/SM of L2->SM traffic, the measured L2 throughput cap "
                    "(profiles/r4_summary.md); DRAM traffic is 1.1x algorithmic"}

    # ---------------------------------------------------------------- aggregate over ranks
    from unidet3d_b200 import sharding
    t_dev, t_e2e = sharding.aggregate_times([t_dev, t_e2e], device=dev)
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    value = sharding.whole_job_throughput(batch * K, world, t_dev)
    e2e = sharding.whole_job_throughput(batch * K, world, t_e2e)
    line = {"metric": "scenes/sec", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "preset": preset, "batch_per_gpu": batch,
                       "points_per_scene": int(pts[0].shape[0]), "voxels_per_level": [lv.n for lv in x.pyramid.levels],
                       "superpoints": int(sum(int(s.max()) + 1 for s in sps)), "datasets": list(cfg["decoder"]["datasets"]),
                       "precision": "fp32 storage; bf16 hi/lo 3-term split on tcgen05, fp32 accumulate",
                       "l2": "flushed before every step (256 MiB memset on the step's stream, inside the timed region)",
                       "parallelism": f"scene-sharded dp{world}", "pipeline_depth": depth,
                       "host": "gc.collect() + gc.freeze() after the warm-up (long-lived objects leave the collector)",
                       "latency_ms_per_batch": {"device_resident": lat_dev, "host_buffers": lat_e2e}},
            "e2e": {"value": e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / K},
            "gpu_launches": launches, "roofline": roof, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, scenes, names, preset, args.workload)
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(cfg, scenes, names, preset, workload="scannet_b8", budget_s=150):
    """Oracle ("port" of the reference math) on the host cores, in a child process with a hard time
    budget (a pathological host must not stall the bench): `--impl reference` with 2 steps of 1 scene."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--workload", workload]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "unit": "scenes/s", "cores": host_threads(), "kind": "port",
                "sample": "failed: " + (r.stderr.strip().splitlines() or ["no output"])[-1][:200]}
    except subprocess.TimeoutExpired:
        return {"value": None, "unit": "scenes/s", "cores": host_threads(), "kind": "port",
                "sample": f"timed out after {budget_s}s (3 x 1 scene of {preset})"}


if __name__ == "__main__":
    main()
