"""numpy restatement of the on-device synthetic cell generator (text2loc_b200/csrc/synthgen.cu).  TEST INFRASTRUCTURE.

Bit-exact: a 64-bit integer hash per (seed, global object, slot), 24 random bits -> fp32 in [0, 1), then fp32
multiplies and adds rounded one at a time (numpy float32 arithmetic never fuses)."""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SLOT_POINT0 = 16


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def u01(seed: int, obj, slot):
    """uniform [0, 1) fp32 with 24 random bits; obj, slot broadcast."""
    with np.errstate(over="ignore"):
        key = np.asarray(obj, dtype=np.uint64) * np.uint64(4096) + np.asarray(slot, dtype=np.uint64)
        h = splitmix64(np.uint64(seed) ^ splitmix64(key))
    return (h >> np.uint64(40)).astype(np.uint32).astype(np.float32) * np.float32(2.0 ** -24)


def synth_cells(seed: int, first_cell: int, n_cells: int, obj_per_cell: int):
    """-> pts f32 [n, 256, 6], meta f32 [n, 7], cell_ptr i32 [n_cells + 1] for cells [first_cell, first_cell + n_cells)."""
    f = np.float32
    n = n_cells * obj_per_cell
    g = (np.arange(n, dtype=np.uint64) + np.uint64(first_cell * obj_per_cell))[:, None]  # [n, 1]
    c3 = np.arange(3, dtype=np.uint64)[None, :]
    zs = np.array([1.0, 1.0, 0.2], dtype=f)
    centre = u01(seed, g, c3) * zs
    extent = (f(0.02) + u01(seed, g, 3 + c3) * f(0.28)) * np.array([1.0, 1.0, 0.3], dtype=f)
    colour = u01(seed, g, 6 + c3)
    n_raw = 30 + (u01(seed, g[:, 0], 9) * f(4971.0)).astype(np.int64)
    k = np.arange(256, dtype=np.uint64)[None, :]
    s_small = (u01(seed, g, SLOT_POINT0 + 8 * k + 6) * n_raw[:, None].astype(f)).astype(np.uint64)
    s = np.where(n_raw[:, None] < 256, s_small, k)  # [n, 256]
    slot = (SLOT_POINT0 + 8 * s)[:, :, None] + c3[None, :, :]  # [n, 256, 3]
    g3 = g[:, :, None]
    u = u01(seed, g3, slot) + f(-0.5)
    xyz = centre[:, None, :] + u * extent[:, None, :]
    nz = (u01(seed, g3, slot + np.uint64(3)) + f(-0.5)) * f(0.17320508)
    rgb = np.clip(colour[:, None, :] + nz, f(0.0), f(1.0))
    pts = np.concatenate([xyz, rgb], axis=2).astype(f)
    meta = np.concatenate([colour, centre, n_raw[:, None].astype(f)], axis=1).astype(f)
    cell_ptr = (np.arange(n_cells + 1) * obj_per_cell).astype(np.int32)
    return pts, meta, cell_ptr
