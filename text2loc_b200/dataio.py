"""Host-side containers and packing for the coarse path's inputs.

Mirrors the part of the reference's data plumbing whose OUTPUT LAYOUT is the hot path's
input contract:

  dataloading/kitti360pose/utils.py:134-146   batch_object_points -> one point batch per cell
  evaluation/coarse.py:95-98                  T.FixedPoints(256) [+ T.NormalizeScale()]
  models/object_encoder.py:122-145            per-object mean rgb, centre, raw point count
  dataloading/kitti360pose/base.py:83-87      dict-of-lists collate

The reference uses PyG's Data/Batch; the engine only needs `.x` (rgb), `.pos` (xyz) and equal
256-point objects, so any object exposing those (a real PyG Batch included) is accepted.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

NUM_POINTS = 256


class PointsBatch:
    """x = rgb [n*P, 3], pos = xyz [n*P, 3], batch = object id per point (PyG Batch surface)."""

    def __init__(self, x: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor = None, num_graphs: int = None):
        self.x = x
        self.pos = pos
        self.batch = batch
        self.num_graphs = num_graphs

    @property
    def num_nodes(self):
        return self.pos.shape[0]

    def to(self, device):
        return self

    @classmethod
    def from_data_list(cls, data_list: Sequence["PointsBatch"]):
        x = torch.cat([d.x for d in data_list])
        pos = torch.cat([d.pos for d in data_list])
        batch = torch.cat([torch.full((d.num_nodes,), i, dtype=torch.long) for i, d in enumerate(data_list)])
        return cls(x, pos, batch, len(data_list))


class FixedPoints:
    """T.FixedPoints(num): `num` indices with replacement from the global numpy RNG."""

    def __init__(self, num: int = NUM_POINTS):
        self.num = num

    def __call__(self, data: PointsBatch) -> PointsBatch:
        choice = torch.from_numpy(np.random.choice(data.num_nodes, self.num, replace=True)).long()
        return PointsBatch(data.x[choice], data.pos[choice])


class NormalizeScale:
    """T.NormalizeScale(): centre on the mean and scale into (-1, 1)."""

    def __call__(self, data: PointsBatch) -> PointsBatch:
        pos = data.pos - data.pos.mean(dim=-2, keepdim=True)
        pos = pos * ((1 / pos.abs().max()) * 0.999999)
        return PointsBatch(data.x, pos)


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data


class PaddingObject:
    """Object3d.create_padding() (datapreparation/kitti360pose/imports.py:74-83): eight near-zero points, black, label "pad"."""

    def __init__(self):
        self.id = self.instance_id = -1
        self.xyz = np.random.rand(8, 3) * 0.001
        self.rgb = np.zeros((8, 3))
        self.label = "pad"

    def get_color_rgb(self):
        return np.mean(self.rgb, axis=0)

    def get_center(self):
        return np.mean(self.xyz, axis=0)


def batch_object_points(objects, transform) -> PointsBatch:
    """One point batch for the objects of a single cell (utils.py:134-146, default branch)."""
    data_list = [
        transform(PointsBatch(torch.tensor(obj.rgb, dtype=torch.float), torch.tensor(obj.xyz, dtype=torch.float)))
        for obj in objects
    ]
    assert len(data_list) >= 1
    return PointsBatch.from_data_list(data_list)


def collate_fn(data):
    """Kitti360BaseDataset.collate_fn (base.py:83-87): dict of lists."""
    return {key: [d[key] for d in data] for key in data[0].keys()}


def pack_cells(objects: List[list], object_points: Sequence) -> tuple:
    """Flatten (objects, object_points) of a batch of cells into the engine layout:

      pts      f32 [n_total, 256, 6]   xyz ‖ rgb
      meta     f32 [n_total, 7]        mean rgb ‖ centre ‖ raw count  (object_encoder.py:122-145;
                                       float64 numpy means narrowed to f32 as torch.tensor(..., dtype=float) does)
      cell_ptr i32 [B+1]
    """
    assert len(objects) == len(object_points)
    pts, meta, ptr = [], [], [0]
    for objs, pb in zip(objects, object_points):
        n = len(objs)
        pos, x = torch.as_tensor(pb.pos), torch.as_tensor(pb.x)
        if pos.shape[0] != n * NUM_POINTS:
            raise ValueError(
                f"cell has {n} objects but {pos.shape[0]} points; the engine requires exactly "
                f"{NUM_POINTS} points per object (pointnet_numpoints, evaluation/args.py:58)"
            )
        pts.append(torch.cat([pos.float(), x.float()], dim=1).reshape(n, NUM_POINTS, 6))
        for obj in objs:
            meta.append(np.concatenate([obj.get_color_rgb(), obj.get_center(), [len(obj.xyz)]]))
        ptr.append(ptr[-1] + n)
    return (
        torch.cat(pts).contiguous(),
        torch.from_numpy(np.asarray(meta, dtype=np.float64).astype(np.float32)),
        torch.tensor(ptr, dtype=torch.int32),
    )


class PackedCells:
    """A cell database in the engine's input layout, packed once and reusable across evaluations."""

    def __init__(self, pts: torch.Tensor, meta: torch.Tensor, cell_ptr: torch.Tensor, cell_ids):
        self.pts, self.meta, self.cell_ptr, self.cell_ids = pts, meta, cell_ptr, list(cell_ids)

    def __len__(self):
        return len(self.cell_ids)


def pack_cell_database(cells, num_points: int = NUM_POINTS, normalize_scale: bool = False, rng=None) -> PackedCells:
    """Vectorised packing of a whole database.

    The reference builds one PyG batch per cell per evaluation -- a Python loop over objects doing T.FixedPoints and
    three numpy reductions each (dataloading/kitti360pose/utils.py:134-146, models/object_encoder.py:122-145), ~1.7 ms per
    cell, which is ~300x the engine's device time per cell.  Here all raw points of all objects are concatenated once,
    the `num_points` samples of every object are drawn in one call (uniform with replacement, as T.FixedPoints does
    for every object), the per-object mean colour / centre come from two segmented sums, and the result is kept.
    normalize_scale: apply T.NormalizeScale to each object's sample (the `no_pc_augment=False` transform,
    evaluation/coarse.py:95-98).  The draw uses `rng` (numpy Generator; default: seeded from numpy's global state), so the
    sample differs from the reference's per-object np.random.choice sequence -- the reference itself is unseeded there."""
    rng = rng or np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))
    objs = [o for c in cells for o in c.objects]
    assert all(len(c.objects) >= 1 for c in cells)  # dataloading/kitti360pose/cells.py:202
    counts = np.array([len(o.xyz) for o in objs], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(counts)])
    xyz = np.concatenate([np.asarray(o.xyz, dtype=np.float64) for o in objs])
    rgb = np.concatenate([np.asarray(o.rgb, dtype=np.float64) for o in objs])
    idx = off[:-1, None] + np.minimum((rng.random((len(objs), num_points)) * counts[:, None]).astype(np.int64), counts[:, None] - 1)
    pos = xyz[idx].astype(np.float32)  # torch.tensor(obj.xyz, dtype=torch.float)[choice]
    col = rgb[idx].astype(np.float32)
    if normalize_scale:
        pos = pos - pos.mean(axis=1, keepdims=True)
        pos = pos * ((1.0 / np.abs(pos).max(axis=(1, 2), keepdims=True)) * 0.999999).astype(np.float32)
    meta = np.empty((len(objs), 7), dtype=np.float64)
    meta[:, 0:3] = np.add.reduceat(rgb, off[:-1], axis=0) / counts[:, None]  # Object3d.get_color_rgb
    meta[:, 3:6] = np.add.reduceat(xyz, off[:-1], axis=0) / counts[:, None]  # Object3d.get_center
    meta[:, 6] = counts                                                      # len(obj.xyz)
    ptr = np.concatenate([[0], np.cumsum([len(c.objects) for c in cells])]).astype(np.int32)
    return PackedCells(torch.from_numpy(np.concatenate([pos, col], axis=2)).contiguous(), torch.from_numpy(meta.astype(np.float32)),
                       torch.from_numpy(ptr), [c.id for c in cells])
