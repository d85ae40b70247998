"""Seeded synthetic weights and inputs for the coarse cell-retrieval path.

There are no checkpoints, no KITTI360Pose data and no T5 weights offline, so
parity tests, the golden-vector generator and ``bench.py`` all draw weights and
inputs from here.  Everything is generated with ``numpy.random.default_rng``
(PCG64), which is bit-stable across machines, so the GPU box regenerates
exactly what the golden vectors were made from.

Key names and shapes follow the reference's ``CellRetrievalNetwork.state_dict()``
(minus ``llm_model.*``, which the reference itself never saves:
training/coarse.py:327-332):

  models/pointcloud/pointnet2.py:57-63   sa1..sa3, ga, lin1, lin2, classifiers
  models/object_encoder.py:33-64         embeddings, pos/color/num encoders, mlp_pointnet, mlp_merge
  models/cell_retrieval.py:35            obj_inter_module.{0,1}
  models/language_encoder.py:98-103      intra_module.0, inter_mlp.0, inter_module.0
"""
from __future__ import annotations

import numpy as np

# Shapes fixed by evaluation/args.py defaults.
EMBED_DIM = 256
T5_DIM = 1024
NUM_POINTS = 256
OBJECT_SIZE = 28
NUM_CLASSES = 22  # len(KNOWN_CLASS), datapreparation/kitti360pose/utils.py:48-69
NUM_COLORS = 8  # len(COLOR_NAMES), datapreparation/kitti360pose/utils.py:231
NUM_MEAN = 1826.6844940968194  # models/object_encoder.py:44
NUM_STD = 2516.8905096993817  # models/object_encoder.py:45


def _linear(rng, sd, prefix, n_out, n_in, gain=1.0):
    sd[prefix + ".weight"] = (rng.standard_normal((n_out, n_in)) * (gain / np.sqrt(n_in))).astype(np.float32)
    sd[prefix + ".bias"] = (rng.standard_normal(n_out) * 0.05).astype(np.float32)


def _batchnorm(rng, sd, prefix, n):
    sd[prefix + ".weight"] = rng.uniform(0.8, 1.2, n).astype(np.float32)
    sd[prefix + ".bias"] = (rng.standard_normal(n) * 0.1).astype(np.float32)
    sd[prefix + ".running_mean"] = (rng.standard_normal(n) * 0.1).astype(np.float32)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, n).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(1000, dtype=np.int64)


def _mlp(rng, sd, prefix, channels):
    """get_mlp / get_mlp2 layout: <prefix>.<i>.0 = Linear, <prefix>.<i>.1 = BatchNorm1d
    (models/language_encoder.py:16-74)."""
    for i in range(1, len(channels)):
        _linear(rng, sd, f"{prefix}.{i - 1}.0", channels[i], channels[i - 1], gain=1.4)
        _batchnorm(rng, sd, f"{prefix}.{i - 1}.1", channels[i])


def _encoder_layer(rng, sd, prefix, d, ffn):
    """nn.TransformerEncoderLayer(d, H, dim_feedforward=ffn) key layout."""
    sd[prefix + ".self_attn.in_proj_weight"] = (rng.standard_normal((3 * d, d)) / np.sqrt(d)).astype(np.float32)
    sd[prefix + ".self_attn.in_proj_bias"] = (rng.standard_normal(3 * d) * 0.05).astype(np.float32)
    _linear(rng, sd, prefix + ".self_attn.out_proj", d, d)
    _linear(rng, sd, prefix + ".linear1", ffn, d, gain=1.4)
    _linear(rng, sd, prefix + ".linear2", d, ffn)
    for n in ("norm1", "norm2"):
        sd[f"{prefix}.{n}.weight"] = rng.uniform(0.9, 1.1, d).astype(np.float32)
        sd[f"{prefix}.{n}.bias"] = (rng.standard_normal(d) * 0.05).astype(np.float32)


def make_state_dict(seed: int = 0) -> dict:
    """Random 'trained-like' weights under the reference's checkpoint key names
    (SURVEY.md Appendix B): unit-ish activation scale, non-trivial BN running stats."""
    rng = np.random.default_rng([seed, 0x7E27])
    sd: dict = {}
    pn = "object_encoder.pointnet"
    _mlp(rng, sd, f"{pn}.sa1.point_conv.local_nn", [3 + 3, 32, 64])
    _mlp(rng, sd, f"{pn}.sa2.point_conv.local_nn", [64 + 3, 128, 128])
    _mlp(rng, sd, f"{pn}.sa3.point_conv.local_nn", [128 + 3, 256, 256])
    _mlp(rng, sd, f"{pn}.ga.mlp", [256 + 3, 512, 1024])
    _linear(rng, sd, f"{pn}.lin1", 512, 1024, gain=1.4)
    _linear(rng, sd, f"{pn}.lin2", 256, 512, gain=1.4)
    _linear(rng, sd, f"{pn}.class_classifier", NUM_CLASSES, 256)
    _linear(rng, sd, f"{pn}.color_classifier", NUM_COLORS, 256)
    oe = "object_encoder"
    sd[f"{oe}.class_embedding.weight"] = rng.standard_normal((NUM_CLASSES + 1, EMBED_DIM)).astype(np.float32)
    sd[f"{oe}.color_embedding.weight"] = rng.standard_normal((NUM_COLORS, EMBED_DIM)).astype(np.float32)
    _mlp(rng, sd, f"{oe}.pos_encoder", [3, 64, EMBED_DIM])
    _mlp(rng, sd, f"{oe}.color_encoder", [3, 64, EMBED_DIM])
    _mlp(rng, sd, f"{oe}.num_encoder", [1, 64, EMBED_DIM])
    _mlp(rng, sd, f"{oe}.mlp_pointnet", [256, EMBED_DIM])
    _mlp(rng, sd, f"{oe}.mlp_merge", [4 * EMBED_DIM, EMBED_DIM])
    for i in range(2):
        _encoder_layer(rng, sd, f"obj_inter_module.{i}", EMBED_DIM, 2 * EMBED_DIM)
    le = "language_encoder"
    _encoder_layer(rng, sd, f"{le}.intra_module.0", T5_DIM, 4 * T5_DIM)
    _mlp(rng, sd, f"{le}.inter_mlp", [T5_DIM, EMBED_DIM])
    _encoder_layer(rng, sd, f"{le}.inter_module.0", EMBED_DIM, 4 * EMBED_DIM)
    return sd


FINE_EMBED_DIM = 128  # fine_embed_dim, evaluation/args.py:41
FINE_PAD_SIZE = 16  # pad_size, evaluation/args.py:45


def _decoder_layer(rng, sd, prefix, d, ffn):
    """nn.TransformerDecoderLayer(d, H, dim_feedforward=ffn) key layout (models/cross_matcher.py:66-72)."""
    for att in ("self_attn", "multihead_attn"):
        sd[f"{prefix}.{att}.in_proj_weight"] = (rng.standard_normal((3 * d, d)) / np.sqrt(d)).astype(np.float32)
        sd[f"{prefix}.{att}.in_proj_bias"] = (rng.standard_normal(3 * d) * 0.05).astype(np.float32)
        _linear(rng, sd, f"{prefix}.{att}.out_proj", d, d)
    _linear(rng, sd, prefix + ".linear1", ffn, d, gain=1.4)
    _linear(rng, sd, prefix + ".linear2", d, ffn)
    for n in ("norm1", "norm2", "norm3"):
        sd[f"{prefix}.{n}.weight"] = rng.uniform(0.9, 1.1, d).astype(np.float32)
        sd[f"{prefix}.{n}.bias"] = (rng.standard_normal(d) * 0.05).astype(np.float32)


def make_fine_state_dict(seed: int = 0, n_decoder_layers: int = 2) -> dict:
    """Random 'trained-like' weights under the key names of the reference's fine-stage model
    ``CrossMatch.state_dict()`` (models/cross_matcher.py:39-78): ObjectEncoder and
    LanguageEncoder(is_fine=True) at d = 128, the cascaded cross-attention decoder layers and the offset MLP."""
    rng = np.random.default_rng([seed, 0xF17E])
    d = FINE_EMBED_DIM
    sd: dict = {}
    pn = "object_encoder.pointnet"
    _mlp(rng, sd, f"{pn}.sa1.point_conv.local_nn", [3 + 3, 32, 64])
    _mlp(rng, sd, f"{pn}.sa2.point_conv.local_nn", [64 + 3, 128, 128])
    _mlp(rng, sd, f"{pn}.sa3.point_conv.local_nn", [128 + 3, 256, 256])
    _mlp(rng, sd, f"{pn}.ga.mlp", [256 + 3, 512, 1024])
    _linear(rng, sd, f"{pn}.lin1", 512, 1024, gain=1.4)
    _linear(rng, sd, f"{pn}.lin2", 256, 512, gain=1.4)
    _linear(rng, sd, f"{pn}.class_classifier", NUM_CLASSES, 256)
    _linear(rng, sd, f"{pn}.color_classifier", NUM_COLORS, 256)
    oe = "object_encoder"
    sd[f"{oe}.class_embedding.weight"] = rng.standard_normal((NUM_CLASSES + 1, d)).astype(np.float32)
    sd[f"{oe}.color_embedding.weight"] = rng.standard_normal((NUM_COLORS, d)).astype(np.float32)
    _mlp(rng, sd, f"{oe}.pos_encoder", [3, 64, d])
    _mlp(rng, sd, f"{oe}.color_encoder", [3, 64, d])
    _mlp(rng, sd, f"{oe}.num_encoder", [1, 64, d])
    _mlp(rng, sd, f"{oe}.mlp_pointnet", [256, d])
    _mlp(rng, sd, f"{oe}.mlp_merge", [4 * d, d])
    le = "language_encoder"
    _encoder_layer(rng, sd, f"{le}.intra_module.0", T5_DIM, 4 * T5_DIM)
    _mlp(rng, sd, f"{le}.inter_mlp", [T5_DIM, d])
    for i in range(n_decoder_layers):
        _decoder_layer(rng, sd, f"cross_hints.{i}", d, 4 * d)
        _decoder_layer(rng, sd, f"cross_objects.{i}", d, 4 * d)
    _linear(rng, sd, "mlp_offsets.0", d // 2, d, gain=1.4)
    _linear(rng, sd, "mlp_offsets.2", 2, d // 2)
    return sd


def pointnet_state_dict(sd: dict) -> dict:
    """The sub-dict ObjectEncoder.__init__ loads from args.pointnet_path (models/object_encoder.py:50)."""
    p = "object_encoder.pointnet."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


# ---------------------------------------------------------------------------
# Inputs
# ---------------------------------------------------------------------------

class SynthObject:
    """Duck-typed stand-in for datapreparation/kitti360pose/imports.py::Object3d (fields the
    hot path reads: xyz, rgb, label, get_color_rgb, get_center; object_encoder.py:122-145)."""

    __slots__ = ("id", "instance_id", "xyz", "rgb", "label")

    def __init__(self, id, xyz, rgb, label="object"):
        self.id = id
        self.instance_id = id
        self.xyz = xyz
        self.rgb = rgb
        self.label = label

    def get_color_rgb(self):
        return np.mean(self.rgb, axis=0)

    def get_center(self):
        return np.mean(self.xyz, axis=0)


def make_cell_objects(seed: int, n_cells: int, n_obj, max_raw: int = 5000):
    """Raw (un-sampled) objects per cell, SURVEY.md §8(d) config-1 recipe.

    n_obj: int, or a sequence of per-cell object counts (ragged cells).
    Returns list[list[SynthObject]] with float64 xyz / rgb like the reference pickles.
    """
    rng = np.random.default_rng([seed, 0xCE11])
    counts = [n_obj] * n_cells if np.isscalar(n_obj) else list(n_obj)
    assert len(counts) == n_cells
    cells = []
    for c in range(n_cells):
        objs = []
        for o in range(counts[c]):
            centre = np.concatenate([rng.uniform(0, 1, 2), rng.uniform(0, 0.2, 1)])
            extent = rng.uniform(0.02, 0.3, 3) * np.array([1.0, 1.0, 0.3])
            n_raw = int(rng.integers(30, max_raw + 1))
            xyz = centre + (rng.uniform(-0.5, 0.5, (n_raw, 3)) * extent)
            colour = rng.uniform(0, 1, 3)
            rgb = np.clip(colour + rng.standard_normal((n_raw, 3)) * 0.05, 0.0, 1.0)
            objs.append(SynthObject(o, xyz, rgb))
        cells.append(objs)
    return cells


def sample_fixed_points(rng, n_raw: int, num: int = NUM_POINTS) -> np.ndarray:
    """T.FixedPoints(num) with PyG's default replace=True: `num` indices drawn uniformly
    with replacement (dataloading/kitti360pose/utils.py:141-142 applies it per object)."""
    return rng.integers(0, n_raw, num)


def pack_cells(cells, seed: int):
    """Host-side packing of raw objects into the engine's input layout:

      pts      f32 [n_total, 256, 6]  xyz ‖ rgb of the 256-sample
      meta     f32 [n_total, 7]       mean rgb (3) ‖ centre (3) ‖ raw point count (1)
      cell_ptr i32 [B+1]
    """
    rng = np.random.default_rng([seed, 0xF1ED])
    n_total = sum(len(c) for c in cells)
    pts = np.empty((n_total, NUM_POINTS, 6), np.float32)
    meta = np.empty((n_total, 7), np.float32)
    cell_ptr = np.zeros(len(cells) + 1, np.int32)
    k = 0
    for ci, objs in enumerate(cells):
        for obj in objs:
            idx = sample_fixed_points(rng, len(obj.xyz))
            pts[k, :, 0:3] = obj.xyz[idx]
            pts[k, :, 3:6] = obj.rgb[idx]
            meta[k, 0:3] = obj.get_color_rgb()
            meta[k, 3:6] = obj.get_center()
            meta[k, 6] = len(obj.xyz)
            k += 1
        cell_ptr[ci + 1] = k
    return pts, meta, cell_ptr


def make_packed_cells(seed: int, n_cells: int, n_obj: int):
    """Vectorised generator straight into the packed layout for bench-sized configs (the
    raw objects are never materialised).  Same distributions as make_cell_objects; points
    drawn with equal sample index coincide, as they do under FixedPoints' replacement."""
    rng = np.random.default_rng([seed, 0xB16])
    n = n_cells * n_obj
    centre = np.concatenate([rng.uniform(0, 1, (n, 2)), rng.uniform(0, 0.2, (n, 1))], axis=1)
    extent = rng.uniform(0.02, 0.3, (n, 3)) * np.array([1.0, 1.0, 0.3])
    n_raw = rng.integers(30, 5001, n)
    colour = rng.uniform(0, 1, (n, 3))
    pts = np.empty((n, NUM_POINTS, 6), np.float32)
    u = rng.random((n, NUM_POINTS, 3), dtype=np.float32) - 0.5
    g = rng.standard_normal((n, NUM_POINTS, 3), dtype=np.float32) * 0.05
    # objects with fewer raw points than samples repeat points: fold the sample index
    idx = (rng.random((n, NUM_POINTS)) * n_raw[:, None]).astype(np.int64)
    small = n_raw < NUM_POINTS
    if small.any():
        rows = np.nonzero(small)[0]
        src = idx[rows] % NUM_POINTS
        u[rows] = np.take_along_axis(u[rows], src[:, :, None].repeat(3, 2), axis=1)
        g[rows] = np.take_along_axis(g[rows], src[:, :, None].repeat(3, 2), axis=1)
    pts[:, :, 0:3] = centre[:, None, :] + u * extent[:, None, :]
    pts[:, :, 3:6] = np.clip(colour[:, None, :] + g, 0.0, 1.0)
    meta = np.empty((n, 7), np.float32)
    sig = 1.0 / np.sqrt(n_raw)[:, None]
    meta[:, 0:3] = np.clip(colour + rng.standard_normal((n, 3)) * 0.05 * sig, 0, 1)
    meta[:, 3:6] = centre + rng.standard_normal((n, 3)) * extent * 0.2887 * sig
    meta[:, 6] = n_raw
    cell_ptr = (np.arange(n_cells + 1) * n_obj).astype(np.int32)
    return pts, meta, cell_ptr


def make_t5_features(seed: int, n_queries: int, n_sent: int = 6, n_tok: int = 12) -> np.ndarray:
    """Stand-in for T5EncoderModel(...).last_hidden_state (models/language_encoder.py:122-125):
    f32 [n_queries * n_sent, n_tok, 1024], N(0,1)*0.2."""
    rng = np.random.default_rng([seed, 0x75])
    return (rng.standard_normal((n_queries * n_sent, n_tok, T5_DIM), dtype=np.float32) * 0.2)


def make_unit_rows(seed: int, n: int, d: int = EMBED_DIM) -> np.ndarray:
    """Random unit fp32 rows (search-only configs)."""
    rng = np.random.default_rng([seed, 0xD8])
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


# ---------------------------------------------------------------------------
# Duck-typed dataset with the surface eval_epoch / run_coarse touch
# (training/coarse.py:63-157, evaluation/coarse.py:40-84, dataloading/kitti360pose/cells.py:96-213)
# ---------------------------------------------------------------------------

DIRECTIONS = ("north", "south", "east", "west", "on-top")
COLOR_WORDS = ("dark-green", "gray", "gray-green", "bright-gray", "black", "green", "beige", "red")
CLASS_WORDS = ("building", "pole", "traffic light", "traffic sign", "parking", "sidewalk", "vegetation", "terrain",
               "road", "wall", "garage", "fence", "bridge", "tunnel", "box", "trash bin", "lamp", "smallpole",
               "guard rail", "vending machine", "stop", "bus stop")


class SynthCell:
    """Fields of datapreparation/kitti360pose/imports.py::Cell the coarse path reads."""

    def __init__(self, idx, scene_name, objects, cell_size, bbox_w):
        self.scene_name = scene_name
        self.id = f"{scene_name}_{idx:05.0f}"
        self.objects = objects
        self.cell_size = cell_size
        self.bbox_w = bbox_w

    def get_center(self):
        return 1 / 2 * (self.bbox_w[0:3] + self.bbox_w[3:6])


class SynthPose:
    def __init__(self, pose_w, cell_id, scene_name, text):
        self.pose_w = pose_w
        self.cell_id = cell_id
        self.scene_name = scene_name
        self.text = text


class SynthCellOnlyDataset:
    """Kitti360CoarseCellOnlyDataset (cells.py:190-213)."""

    def __init__(self, cells, transform):
        self.cells = cells
        self.transform = transform

    def __getitem__(self, idx):
        from text2loc_b200.dataio import batch_object_points

        cell = self.cells[idx]
        assert len(cell.objects) >= 1
        return {
            "cells": cell,
            "cell_ids": cell.id,
            "objects": cell.objects,
            "object_points": batch_object_points(cell.objects, self.transform),
        }

    def __len__(self):
        return len(self.cells)


class SynthCoarseDataset:
    """Kitti360CoarseDatasetMulti surface: one item per pose, all_cells / all_poses,
    get_cell_dataset() (cells.py:119-187).  Seeded; ragged object counts per cell."""

    def __init__(self, seed: int, n_cells: int, n_poses: int, n_obj=8, transform=None, max_raw: int = 5000,
                 scene_name: str = "0000", cell_size: float = 30.0, n_hints: int = 6):
        from text2loc_b200.dataio import FixedPoints

        self.transform = transform or FixedPoints(NUM_POINTS)
        rng = np.random.default_rng([seed, 0xDA7A])
        if np.isscalar(n_obj):
            n_obj = [n_obj] * n_cells
        objs = make_cell_objects(seed, n_cells, n_obj, max_raw=max_raw)
        side = int(np.ceil(np.sqrt(n_cells)))
        self.all_cells = []
        for i in range(n_cells):
            x0, y0 = (i % side) * cell_size / 2, (i // side) * cell_size / 2
            bbox = np.array([x0, y0, 0.0, x0 + cell_size, y0 + cell_size, cell_size])
            self.all_cells.append(SynthCell(i, scene_name, objs[i], cell_size, bbox))
        self.all_poses = []
        for _ in range(n_poses):
            cell = self.all_cells[int(rng.integers(n_cells))]
            pose_w = cell.bbox_w[0:3] + rng.uniform(0.2, 0.8, 3) * cell_size
            hints = [
                f"The pose is {DIRECTIONS[rng.integers(5)]} of a {COLOR_WORDS[rng.integers(8)]} {CLASS_WORDS[rng.integers(22)]}."
                for _ in range(n_hints)
            ]
            self.all_poses.append(SynthPose(pose_w, cell.id, scene_name, " ".join(hints)))
        self._cells_dict = {c.id: c for c in self.all_cells}

    def __getitem__(self, idx):
        from text2loc_b200.dataio import batch_object_points

        pose = self.all_poses[idx]
        cell = self._cells_dict[pose.cell_id]
        return {
            "poses": pose,
            "cells": cell,
            "objects": cell.objects,
            "object_points": batch_object_points(cell.objects, self.transform),
            "texts": pose.text,
            "cell_ids": pose.cell_id,
            "scene_names": pose.scene_name,
        }

    def __len__(self):
        return len(self.all_poses)

    def get_cell_dataset(self):
        return SynthCellOnlyDataset(self.all_cells, self.transform)
