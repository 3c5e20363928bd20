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
trace = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
lib.ud3d_debug_set_trace(C.c_void_p(trace.data_ptr()), -2)
run(); torch.cuda.synchronize(); lib.ud3d_debug_set_trace(None, 0)
t = trace.cpu().numpy().reshape(-1, 4); t = t[t[:, 0] > 0]
t0 = t[:, 0].min(); st = (t[:, 0] - t0) / 1e3; en = (t[:, 1] - t0) / 1e3; sm = t[:, 2]
print(f"{len(t)} CTAs; kernel span {en.max():.1f} us; CTA lifetime mean {np.mean(en - st):.1f} us (min {np.min(en - st):.1f}, max {np.max(en - st):.1f})")
print("start times (us) percentiles 0/25/50/75/100:", np.percentile(st, [0, 25, 50, 75, 100]).round(1))
per_sm = {}
for a, b, s_ in zip(st, en, sm): per_sm.setdefault(int(s_), []).append((a, b))
occ = []
for s_, iv in per_sm.items():
    busy = sum(b - a for a, b in iv); occ.append(busy / en.max())
print(f"SMs used {len(per_sm)}; mean concurrent CTAs per SM {np.mean(occ):.2f} (min {np.min(occ):.2f}, max {np.max(occ):.2f}); CTAs per SM min {min(len(v) for v in per_sm.values())} max {max(len(v) for v in per_sm.values())}")
iv = sorted(per_sm[int(sm[0])]); print("SM", int(sm[0]), "intervals:", [(round(a, 1), round(b, 1)) for a, b in iv[:16]])
