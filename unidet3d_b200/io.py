"""Input wire format and host pipeline of the forward path: the step BEFORE the
hot path.  Plain numpy readers for the reference's on-disk format and a double-buffered pinned-memory stager that
overlaps the H2D copy of batch i+1 with the compute of batch i.

Wire format (reference: tools/scannet_data_utils.py:187-241 writes, unidet3d/loading.py:23-52 and mmdet3d
``LoadPointsFromFile`` read):
  * ``points/<scene>.bin``        float32 [N, 6]  (x, y, z, r, g, b), colours 0..255
  * ``super_points/<scene>.bin``  int64   [N]     superpoint id per point
Colour normalisation: ``(c - color_mean) / color_std`` with mean 127.5, std 127.5
(unidet3d/loading.py:70-106, configs/unidet3d_1xb8_scannet.py:181-183).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


def load_points_bin(path: str, load_dim: int = 6, use_dim: Sequence[int] = (0, 1, 2, 3, 4, 5)) -> np.ndarray:
    """float32 [N, load_dim] point file -> [N, len(use_dim)]."""
    pts = np.fromfile(path, dtype=np.float32)
    if pts.size % load_dim:
        raise ValueError(f"{path}: {pts.size} floats is not a multiple of load_dim={load_dim}")
    return np.ascontiguousarray(pts.reshape(-1, load_dim)[:, list(use_dim)])


def load_superpoints_bin(path: str) -> np.ndarray:
    """int64 [N] superpoint ids (unidet3d/loading.py:38-43)."""
    return np.fromfile(path, dtype=np.int64)


def normalize_points_color(points: np.ndarray, color_mean=127.5, color_std=127.5) -> np.ndarray:
    """NormalizePointsColor_ (unidet3d/loading.py:70-106) on a [N,6] array; returns a new float32 array."""
    out = points.astype(np.float32, copy=True)
    if color_mean is not None:
        out[:, 3:6] -= np.float32(color_mean) if np.isscalar(color_mean) else np.asarray(color_mean, np.float32)
    if color_std is not None:
        out[:, 3:6] /= np.float32(color_std) if np.isscalar(color_std) else np.asarray(color_std, np.float32)
    return out


class PinnedBatchStager:
    """Double-buffered host->device staging of scene batches.

    ``for pts, sps, n_sps in PinnedBatchStager(batches, device): model.forward_scenes(pts, sps, names, n_sps)``
    yields device tensors whose H2D copies ran on a side stream while the previous batch was being processed.
    ``batches`` is an iterable of (list of float32 [N_i,6] arrays, list of int64 [N_i] arrays).
    """

    def __init__(self, batches: Iterable, device, depth: int = 2):
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._pinned = {}
        self._slot_events = {}     # slot -> event of the last H2D copies issued from the slot's pinned buffers

    def _pin(self, key, shape, dtype):
        buf = self._pinned.get(key)
        if buf is None or buf.shape != torch.Size(shape) or buf.dtype != dtype:
            buf = self._pinned[key] = torch.empty(shape, dtype=dtype).pin_memory()
        return buf

    def _stage(self, slot: int, batch):
        pts, sps = batch
        d_pts, d_sps, n_sps = [], [], []
        # the pinned buffers of this slot may still be the source of an asynchronous H2D copy issued earlier: wait for
        # it ON THE HOST before overwriting them (a device-side wait_event does not order the host writes)
        prev = self._slot_events.get(slot)
        if prev is not None:
            prev.synchronize()
        with torch.cuda.stream(self.stream):
            for i, (p, s) in enumerate(zip(pts, sps)):
                p = np.asarray(p, np.float32)
                s = np.asarray(s, np.int64)
                hp = self._pin((slot, i, "p"), p.shape, torch.float32)
                hs = self._pin((slot, i, "s"), s.shape, torch.int64)
                hp.copy_(torch.from_numpy(p))
                hs.copy_(torch.from_numpy(s))
                n_sps.append(int(s.max()) + 1)
                d_pts.append(hp.to(self.device, non_blocking=True))
                d_sps.append(hs.to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot_events[slot] = ev
        return d_pts, d_sps, n_sps, ev

    def __iter__(self) -> Iterator[Tuple[List[torch.Tensor], List[torch.Tensor], List[int]]]:
        queue = []
        slot = 0
        for batch in self.batches:
            queue.append(self._stage(slot % self.depth, batch))
            slot += 1
            if len(queue) >= self.depth:
                yield self._pop(queue)
        while queue:
            yield self._pop(queue)

    def _pop(self, queue):
        d_pts, d_sps, n_sps, ev = queue.pop(0)
        torch.cuda.current_stream().wait_event(ev)
        for t in d_pts + d_sps:
            t.record_stream(torch.cuda.current_stream())
        return d_pts, d_sps, n_sps
