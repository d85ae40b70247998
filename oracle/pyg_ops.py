"""Restatement of the third-party graph ops on the reference's hot path.  TEST INFRASTRUCTURE.

The reference calls (models/pointcloud/pointnet2.py:23-35,48):
    gnn.fps(pos, batch, ratio)
    gnn.radius(x, y, r, batch_x=, batch_y=)           # max_num_neighbors left at default 32
    gnn.PointConv(local_nn=mlp)(x, (pos, pos_sub), edge_index)
    gnn.global_max_pool(x, batch)
from torch_geometric==1.7.2 (requirements.txt:15-18), which forwards to torch-cluster==1.6.0
and torch-scatter==2.0.9.  None of that source is under /root/reference and none of it is
installed here, so this file restates the published algorithms.  Choices those libraries
leave open are pinned here and are part of the parity contract (DESIGN.md "oracle pins"):

  FPS_RANDOM_START   False  start at each object's local index 0 (PyG's default is a random
                            start, which makes the reference itself non-deterministic).
  fps tie rule       first maximum (lowest index), distances d = (dx*dx + dy*dy) + dz*dz in
                            fp32 with no fused multiply-add.
  radius rule        torch-cluster's CUDA kernel: scan the object's points in ascending
                            index, keep the first 32 with d < r*r (strict), r*r evaluated in
                            double then rounded to fp32.
  PointConv          add_self_loops=True (the 1.7.2 default the reference does not override):
                            remove_self_loops on raw indices, then add_self_loops(num_nodes =
                            number of centroids); message = local_nn([x_j, pos_j - pos_i]);
                            aggregation max.
"""
from __future__ import annotations

import math
import re

import numpy as np
import torch
import torch.nn as nn

FPS_RANDOM_START = False
MAX_NUM_NEIGHBORS = 32
# torch-cluster's CUDA kernels accumulate `dist += tmp * tmp` per coordinate; nvcc's default -fmad=true would contract that
# into fma(dz, dz, fma(dy, dy, dx*dx)).  Which form the authors' wheel ran cannot be checked offline, so it is a switch:
# False (default) = one rounding per operation, True = the contracted form.  The engine's twin is T2L_DIST_FMA=1.
DIST_FMA = False


def _sqdist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(dx*dx + dy*dy) + dz*dz, one rounding per op (torch eager never contracts to FMA); with DIST_FMA the
    contracted sum, each fma evaluated in float64 (the product of two fp32 values is exact there) and rounded once
    to fp32 -- identical to a hardware fma except for double-rounding cases of probability ~2^-29 per operation."""
    d = a - b
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    if not DIST_FMA:
        return (dx * dx + dy * dy) + dz * dz
    r = dx * dx
    r = (dy.double() * dy.double() + r.double()).float()
    return (dz.double() * dz.double() + r.double()).float()


def _segments(batch: torch.Tensor):
    """(start, length) of each contiguous run of equal ids in a sorted batch vector."""
    b = batch.cpu().numpy()
    if len(b) == 0:
        return []
    cuts = np.flatnonzero(np.diff(b)) + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(b)]])
    return list(zip(starts.tolist(), (ends - starts).tolist()))


def fps(pos: torch.Tensor, batch: torch.Tensor = None, ratio: float = 0.5, random_start: bool = None):
    """Farthest point sampling per object; returns global indices grouped by object in
    selection order (torch-cluster fps: ceil(ratio * n) picks, dist initialised to +inf,
    dist = min(dist, |p - p_last|^2), next = argmax)."""
    if random_start is None:
        random_start = FPS_RANDOM_START
    assert not random_start, "oracle pins FPS start index 0"
    if batch is None:
        batch = torch.zeros(len(pos), dtype=torch.long)
    out = []
    for start, n in _segments(batch):
        p = pos[start:start + n].float()
        k = int(math.ceil(ratio * n))
        dist = torch.full((n,), float("inf"))
        cur = 0
        sel = [0]
        for _ in range(k - 1):
            dist = torch.minimum(dist, _sqdist(p, p[cur]))
            cur = int(torch.argmax(dist))  # first maximum
            sel.append(cur)
        out.append(torch.tensor(sel, dtype=torch.long) + start)
    return torch.cat(out)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors: int = MAX_NUM_NEIGHBORS):
    """For each y_i: up to max_num_neighbors points x_j of the same object with
    |x_j - y_i|^2 < r^2, lowest j first.  Returns stack([row=i, col=j])."""
    if batch_x is None:
        batch_x = torch.zeros(len(x), dtype=torch.long)
    if batch_y is None:
        batch_y = torch.zeros(len(y), dtype=torch.long)
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float64).float()
    seg_x = {int(batch_x[s]): (s, n) for s, n in _segments(batch_x)}
    rows, cols = [], []
    for ys, yn in _segments(batch_y):
        xs, xn = seg_x[int(batch_y[ys])]
        d = _sqdist(x[xs:xs + xn].float()[None, :, :], y[ys:ys + yn].float()[:, None, :])  # [yn, xn]
        mask = d < r2
        keep = mask & (torch.cumsum(mask.to(torch.int32), dim=1) <= max_num_neighbors)
        i, j = torch.nonzero(keep, as_tuple=True)  # row-major: ascending i, then ascending j
        rows.append(i + ys)
        cols.append(j + xs)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def global_max_pool(x: torch.Tensor, batch: torch.Tensor) -> torch.Tensor:
    return torch.stack([x[s:s + n].max(dim=0)[0] for s, n in _segments(batch)])


class PointConv(nn.Module):
    """PyG 1.7.2 PointConv(local_nn, global_nn=None, add_self_loops=True), aggr='max'."""

    def __init__(self, local_nn=None, global_nn=None, add_self_loops: bool = True):
        super().__init__()
        self.local_nn = local_nn
        self.global_nn = global_nn
        self.add_self_loops = add_self_loops

    def forward(self, x, pos, edge_index):
        if isinstance(pos, torch.Tensor):
            pos = (pos, pos)
        x_src = x[0] if isinstance(x, tuple) else x
        src, dst = edge_index[0], edge_index[1]
        if self.add_self_loops:
            keep = src != dst  # remove_self_loops compares RAW indices
            src, dst = src[keep], dst[keep]
            n = min(pos[0].size(0), pos[1].size(0))  # add_self_loops(num_nodes=min(...))
            loop = torch.arange(n, dtype=src.dtype)
            src, dst = torch.cat([src, loop]), torch.cat([dst, loop])
        msg = pos[0][src] - pos[1][dst]
        if x_src is not None:
            msg = torch.cat([x_src[src], msg], dim=1)
        msg = self.local_nn(msg)
        m = pos[1].size(0)
        out = torch.full((m, msg.size(1)), float("-inf"), dtype=msg.dtype)
        out = out.scatter_reduce(0, dst[:, None].expand_as(msg), msg, reduce="amax", include_self=True)
        if self.global_nn is not None:
            out = self.global_nn(out)
        return out


# ---- data / transforms surface (dataloading/kitti360pose/utils.py:126-146, evaluation/coarse.py:95-98)

class Data:
    def __init__(self, x=None, pos=None, **kw):
        self.x = x
        self.pos = pos
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.pos.size(0)

    def to(self, device):
        return self

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        b = cls(
            x=torch.cat([d.x for d in data_list]),
            pos=torch.cat([d.pos for d in data_list]),
        )
        b.batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long) for i, d in enumerate(data_list)])
        b.num_graphs = len(data_list)
        return b


class FixedPoints:
    """PyG FixedPoints(num, replace=True): np.random.choice(num_nodes, num, replace=True) from
    the GLOBAL numpy RNG (unseeded in the reference; tests seed it)."""

    def __init__(self, num, replace=True, allow_duplicates=False):
        assert replace
        self.num = num

    def __call__(self, data):
        n = data.num_nodes
        choice = torch.from_numpy(np.random.choice(n, self.num, replace=True)).long()
        for k in data.keys():
            v = getattr(data, k)
            if re.search("edge", k):
                continue
            if torch.is_tensor(v) and v.size(0) == n:
                setattr(data, k, v[choice])
        return data


class NormalizeScale:
    """Centre on the mean, scale into (-1, 1)."""

    def __call__(self, data):
        data.pos = data.pos - data.pos.mean(dim=-2, keepdim=True)
        scale = (1 / data.pos.abs().max()) * 0.999999
        data.pos = data.pos * scale
        return data


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data


class RandomRotate:  # training-only augmentation; never on the eval path
    def __init__(self, *a, **k):
        pass

    def __call__(self, data):
        raise NotImplementedError("training augmentation is out of scope for the oracle")
