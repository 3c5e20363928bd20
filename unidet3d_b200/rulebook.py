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


class Pyramid:
    def __init__(self, levels: List[Level]):
        self.levels = levels

    def __len__(self):
        return len(self.levels)


def build_pyramid(coords: torch.Tensor, spatial_shape: Sequence[int], batch_size: int, n_levels: int,
                  canonical: bool = False, extents: Optional[Sequence[int]] = None,
                  grid: Optional[ops.Grid] = None) -> Pyramid:
    """coords int32 [N,4] on the GPU.  ``extents`` = max coord + 1 per axis (bounds the occupancy
    grid; computed here when not given).  ``grid``: an occupancy grid already built on ``coords``."""
    if coords.dtype != torch.int32:
        coords = coords.to(torch.int32)
    coords = coords.contiguous()
    dev = coords.device
    if extents is None:
        extents = (coords[:, 1:].amax(0) + 1).tolist() if coords.shape[0] else [1, 1, 1]
    dims = [int(batch_size)] + [max(1, int(e)) for e in extents]
    shape = [int(s) for s in spatial_shape]
    if grid is None:
        grid = ops.Grid(dims, dev)
        grid.build(coords)
    levels: List[Level] = []
    c, n = coords, coords.shape[0]
    for l in range(n_levels):
        table, mask = ops.rulebook_subm3(c, grid, canonical=(canonical or l > 0))
        lv = Level(coords=c, shape=list(shape), n=n, subm=table, subm_mask=mask)
        levels.append(lv)
        if l + 1 == n_levels:
            break
        parents = ops.down2_parents(c, shape)
        out_shape = [(s - 2) // 2 + 1 for s in shape]
        dims = [dims[0]] + [max(1, min(o, (d + 1) // 2)) for o, d in zip(out_shape, dims[1:])]
        cgrid = ops.Grid(dims, dev)
        n_next = int(cgrid.build(parents).item())            # one host sync per level
        cc = cgrid.coords(n_next)
        lv.child, lv.up, lv.child_mask, lv.up_mask = ops.rulebook_down2(c, parents, n_next, cgrid)
        c, n, shape, grid = cc, n_next, out_shape, cgrid
    return Pyramid(levels)
