"""Rulebook pyramid: every neighbour table the backbone needs, built once per batch on the GPU
(the role of spconv's ``indice_dict`` cache: ``subm{l}`` shared by all 3x3x3 SubM convs of level l,
``spconv{l}`` shared by the strided conv and its inverse; reference unidet3d/spconv_unet.py:138,154,
183,200 and unidet3d/unidet3d.py:103)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import ops


@dataclass
class Level:
    coords: torch.Tensor                 # int32 [N,4]
    shape: List[int]                     # spconv spatial_shape of this level
    n: int
    subm: torch.Tensor                   # int32 [27,N]
    subm_mask: torch.Tensor              # [ceil(N/128)]
    child: Optional[torch.Tensor] = None       # int32 [8,N_next]   (strided conv gather table)
    child_mask: Optional[torch.Tensor] = None
    up: Optional[torch.Tensor] = None          # int32 [8,N]        (inverse conv gather table)
    up_mask: Optional[torch.Tensor] = None
    # rows regrouped by neighbourhood pattern (ops.subm3_tile_order): fewer active offsets per 128-row tile
    perm: Optional[torch.Tensor] = None        # int32 [N]   position -> row
    subm_p: Optional[torch.Tensor] = None      # int32 [27,N] = subm[:, perm]
    subm_mask_p: Optional[torch.Tensor] = None

    order_ready: Optional[torch.cuda.Event] = None   # the regrouping runs on a side stream

    @property
    def subm_conv(self):
        """(table, tile_mask, row_perm) to run a SubM3 convolution of this level with."""
        if self.perm is not None:
            if self.order_ready is not None:
                torch.cuda.current_stream().wait_event(self.order_ready)
                self.order_ready = None
            return self.subm_p, self.subm_mask_p, self.perm
        return self.subm, self.subm_mask, None


# levels with at least this many voxels get the regrouped SubM3 tile order (smaller levels are split-K launches whose
# time is not in the main loop; the regrouping pass would cost more than it saves)
TILE_ORDER_MIN_ROWS = 16384


_SIDE = {}


def _side_stream(dev) -> torch.cuda.Stream:
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


class Pyramid:
    def __init__(self, levels: List[Level]):
        self.levels = levels
        self._c_tables = None

    def c_tables(self, n_levels: int):
        """``ud3d_unet_tables[n_levels]`` (ctypes) for ud3d_unet_forward; orders the current stream after the side-stream
        regrouping of each level (``Level.subm_conv``)."""
        from . import _lib
        tabs = [lv.subm_conv for lv in self.levels[:n_levels]]      # (waits for order_ready on the current stream)
        if self._c_tables is None or len(self._c_tables) != n_levels:
            T = (_lib.UnetTables * n_levels)()
            for l, lv in enumerate(self.levels[:n_levels]):
                tb, tm, pm = tabs[l]
                t = T[l]
                t.n = lv.n
                t.subm, t.subm_mask = tb.data_ptr(), tm.data_ptr()
                t.row_perm = pm.data_ptr() if pm is not None else None
                if l + 1 < n_levels:
                    t.child, t.child_mask = lv.child.data_ptr(), lv.child_mask.data_ptr()
                    t.up, t.up_mask = lv.up.data_ptr(), lv.up_mask.data_ptr()
            self._c_tables = T
        return self._c_tables

    def __len__(self):
        return len(self.levels)


def build_pyramid(coords: Optional[torch.Tensor], spatial_shape: Sequence[int], batch_size: int, n_levels: int,
                  canonical: bool = False, extents: Optional[Sequence[int]] = None,
                  grid: Optional[ops.Grid] = None, seed_coords: Optional[torch.Tensor] = None) -> Pyramid:
    """coords int32 [N,4] on the GPU.  ``extents`` = max coord + 1 per axis (bounds the occupancy
    grid; computed here when not given).  ``grid``: an occupancy grid already built on ``coords``."""
    if coords is not None:
        if coords.dtype != torch.int32:
            coords = coords.to(torch.int32)
        coords = coords.contiguous()
    else:
        assert seed_coords is not None and extents is not None, "coords=None needs seed_coords and extents"
    dev = (coords if coords is not None else seed_coords).device
    if extents is None:
        extents = (coords[:, 1:].amax(0) + 1).tolist() if coords.shape[0] else [1, 1, 1]
    dims = [int(batch_size)] + [max(1, int(e)) for e in extents]
    shape = [int(s) for s in spatial_shape]
    count0 = None
    if grid is None:
        grid = ops.Grid(dims, dev)
        count0 = grid.build(coords if coords is not None else seed_coords)
    # the occupancy grid of every coarser level is folded out of the finer level's bitmap back to back (k=2,s=2 with the
    # per-level drop rule), so the voxel counts of every level come back in ONE host synchronisation
    shapes, dimss, grids, counts = [list(shape)], [list(dims)], [grid], []
    for l in range(1, n_levels):
        out_shape = [(s - 2) // 2 + 1 for s in shapes[-1]]
        d = [dimss[-1][0]] + [max(1, min(o, (x + 1) // 2)) for o, x in zip(out_shape, dimss[-1][1:])]
        g = ops.Grid(d, dev)
        counts.append(g.build_from_finer(grids[-1]))     # a pass over the finer bitmap; the points are not touched again
        shapes.append(out_shape), dimss.append(d), grids.append(g)
    if coords is None:
        # level 1 itself comes from the seed (voxelisation): its count rides on the same read-back
        ns = torch.cat([count0] + counts).cpu().tolist()                                # the only host sync here
        coords = grid.coords(ns[0])
        canonical = True
    else:
        ns = [coords.shape[0]] + (torch.cat(counts).cpu().tolist() if counts else [])  # the only host sync here
    levels: List[Level] = []
    c = coords
    for l in range(n_levels):
        table, mask = ops.rulebook_subm3(c, grids[l], canonical=(canonical or l > 0))
        lv = Level(coords=c, shape=list(shapes[l]), n=ns[l], subm=table, subm_mask=mask)
        if ns[l] >= TILE_ORDER_MIN_ROWS:
            # regroup on a side stream: it overlaps the rulebooks of the deeper levels (and, for level 2+, the
            # convolutions of the levels above); the first SubM3 conv of the level waits for `order_ready`
            cur = torch.cuda.current_stream()
            side = _side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                lv.perm, lv.subm_p, lv.subm_mask_p = ops.subm3_tile_order(table)
                lv.order_ready = torch.cuda.Event()
                lv.order_ready.record(side)
            for t in (table, lv.perm, lv.subm_p, lv.subm_mask_p):
                t.record_stream(cur)
            table.record_stream(side)
        levels.append(lv)
        if l + 1 == n_levels:
            break
        parents = ops.down2_parents(c, shapes[l])
        cc = grids[l + 1].coords(ns[l + 1])
        lv.child, lv.up, lv.child_mask, lv.up_mask = ops.rulebook_down2(c, parents, ns[l + 1], grids[l + 1])
        c = cc
    pyr = Pyramid(levels)
    pyr.grid0 = grid              # occupancy grid of level 1 (rank() = inverse mapping of the voxelisation)
    return pyr
