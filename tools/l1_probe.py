"""Level-0 SubM3 conv: cp.async.cg vs .ca (L1-allocating) gathers under different shared-memory carve-outs / CTAs per SM."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200 import ops, _lib  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict  # noqa: E402
from tools.ts_probe import timed  # noqa: E402

lib = _lib.load()
lib.ud3d_debug_set_flags.argtypes = [C.c_int]
cfg, scenes, names, preset = bench.make_workload("scannet_b8", 0)
model = u.MODELS.build(cfg).eval()
model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
model.cuda()
pts = [torch.as_tensor(s[0]).cuda() for s in scenes]
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device="cuda")
x, inv = model.collate(torch.cat(pts), offs, len(pts))
for level in [int(v) for v in os.environ.get("LEVELS", "0").split(",")]:
    lv = x.pyramid.levels[level]
    c = cfg["backbone"]["num_planes"][level]
    xin = torch.relu(torch.randn(lv.n, c, device="cuda"))
    xs = ops.act_split(xin, relu=False)
    w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda") * 0.05)
    act = torch.empty_like(xin); raw = torch.empty_like(xin); res = torch.randn_like(xin)
    one = torch.ones(c, device="cuda"); zero = torch.zeros(c, device="cuda")
    tb, tm, pm = lv.subm_conv
    run = lambda: ops.gemm(xs, w, table=tb, tile_mask=tm, in_split=1, residual=res, out=raw, acts=[(act, one, zero)], row_perm=pm)
    for extra_kb in (0, 40):
        for carve in (0, 50, 66, 75, 85):
            for ca in (0, 128):
                lib.ud3d_debug_set_flags(ca | (carve << 16) | (extra_kb << 24))
                try:
                    print(f"level {level} extra smem {extra_kb} KB carve-out {carve or 'max'} {'ca' if ca else 'cg'}: {timed(run):8.1f} us", flush=True)
                except Exception as e:  # noqa: BLE001
                    print("failed", e)
    lib.ud3d_debug_set_flags(0)
