"""Debug: per-CTA start / end (globaltimer) and SM id of the level-l SubM3 conv."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, unidet3d_b200 as u
from unidet3d_b200 import ops, _lib
from unidet3d_b200.synthetic import make_model_state_dict
cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval(); model.load_state_dict(make_model_state_dict(cfg, 0), strict=False); model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
x, inv = model.collate(torch.cat(pts), offs, len(pts))
level = int(os.environ.get("LEVEL", 0))
lv = x.pyramid.levels[level]; c = cfg["backbone"]["num_planes"][level]
xin = torch.relu(torch.randn(lv.n, c, device="cuda")); xs = ops.act_split(xin, relu=False)
w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
act = torch.empty_like(xin); one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
lib = _lib.load(); lib.ud3d_debug_set_trace.argtypes = [C.c_void_p, C.c_int]; lib.ud3d_debug_set_flags.argtypes = [C.c_int]; lib.ud3d_debug_set_flags(int(os.environ.get("FLAGS", 0)))
run = lambda: ops.gemm(xs, w, table=lv.subm, tile_mask=lv.subm_mask, in_split=True, no_raw=True, acts=[(act, one, zero)])
for _ in range(3): run()
torch.cuda.synchronize()
trace = torch.zeros(8 * 8192, dtype=torch.int64, device="cuda")
lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
run(); torch.cuda.synchronize(); lib.ud3d_debug_set_trace(None, 0)
t = trace.cpu().numpy().reshape(-1, 8); t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = lambda col: (t[:, col] - t0) / 1e3
st, en, sm = rel(0), rel(1), t[:, 2]
print(f"level {level}: {len(t)} CTAs; kernel span {en.max():.1f} us; CTA lifetime mean {np.mean(en - st):.1f} us (min {np.min(en - st):.1f}, max {np.max(en - st):.1f})")
print("start times (us) percentiles 0/25/50/75/100:", np.percentile(st, [0, 25, 50, 75, 100]).round(1))
names = {3: "prologue done", 4: "main loop done", 5: "partial parked", 6: "cluster barrier 1", 7: "reduced+stored", 1: "end"}
prev = st
for col in (3, 4, 5, 6, 7, 1):
    if (t[:, col] > 0).all():
        cur = rel(col)
        d = cur - prev
        print(f"  -> {names[col]:18s}: +{np.mean(d):6.2f} us mean (min {np.min(d):.2f}, max {np.max(d):.2f}); at {np.mean(cur):.1f} us mean")
        prev = cur
print("SMs used", len(set(sm.tolist())))
