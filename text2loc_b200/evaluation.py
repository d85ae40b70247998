"""Drop-ins for training/coarse.py::eval_epoch and evaluation/coarse.py::run_coarse.

Same signatures and return types as the reference.  What changes is where the work happens:

* query and cell encodings stay on the GPU;
* the per-query float64 GEMV + full argsort of training/coarse.py:119-125 is one batched
  t2l_search_topk call whose result is, by construction, the fp64 (score desc, row asc) order;
* the accuracy bookkeeping (training/coarse.py:131-150, evaluation/coarse.py:61-82,
  evaluation/utils.py:31-54) -- per-query Python loops in the reference -- is one
  t2l_topk_accuracy launch per function over ALL queries.  Cell ids
  are strings at the API (training/coarse.py:82,128); they are mapped to database rows once, by a
  sort, and back with one fancy-index.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import dataio


def rows_of_ids(db_ids: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Database row of every id in `ids` (any shape), -1 where the id is not in `db_ids`.
    db_ids are unique (dataloading/kitti360pose/cells.py:149-150)."""
    order = np.argsort(db_ids, kind="stable")
    sorted_ids = db_ids[order]
    flat = np.asarray(ids).reshape(-1)
    pos = np.clip(np.searchsorted(sorted_ids, flat), 0, len(sorted_ids) - 1)
    rows = np.where(sorted_ids[pos] == flat, order[pos], -1).astype(np.int64)
    return rows.reshape(np.shape(ids))


def scene_names(ids) -> np.ndarray:
    """Scene prefix of each id (`id.split("_")[0]`, evaluation/utils.py:45-46);
    compared by equality there)."""
    names = np.array([str(i).split("_")[0] for i in ids])
    return names


def _cached_database(model, cells_dataset):
    """model.cache_packed_cells = True opts in: the cell database is packed once per (model, dataset) with
    dataio.pack_cell_database and reused by every later evaluation; otherwise the reference's per-batch traversal runs."""
    if not getattr(model, "cache_packed_cells", False):
        return None
    cache = model.__dict__.setdefault("_packed_cells", {})
    key = (id(cells_dataset.cells), len(cells_dataset.cells))
    if key not in cache:
        num, normalize_scale = _transform_spec(cells_dataset.transform)
        cache[key] = dataio.pack_cell_database(cells_dataset.cells, num, normalize_scale=normalize_scale)
    return cache[key]


def _transform_spec(transform):
    """(points per object, NormalizeScale?) of a point transform the vectorised packer can stand in for: the two the
    reference's eval scripts build (evaluation/coarse.py:95-98, evaluation/pipeline.py:216-224)."""
    stages = list(getattr(transform, "transforms", [transform]))
    names = [type(t).__name__ for t in stages]
    if not names or names[0] != "FixedPoints" or any(n not in ("FixedPoints", "NormalizeScale") for n in names):
        raise ValueError(f"vectorised packing supports FixedPoints [+ NormalizeScale] transforms, got {names}")
    return getattr(stages[0], "num", dataio.NUM_POINTS), "NormalizeScale" in names


@torch.no_grad()
def eval_epoch(model, dataloader, args, return_encodings=False, return_distance=False):
    """training/coarse.py:63-157.  Returns (accuracies{k}, accuracies_close{k},
    top_retrievals{query_idx: ndarray<U32>[max(top_k)]}) [+ cell_encodings, text_encodings f64]
    [+ dists, scores]."""
    assert args.ranking_loss != "triplet"  # training/coarse.py:65
    model.eval()
    dataset = dataloader.dataset
    cells_dataset = dataset.get_cell_dataset()
    cells_dataloader = DataLoader(cells_dataset, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    n_q, n_c, d = len(dataset), len(cells_dataset), model.embed_dim
    assert n_c == len(dataset.all_cells)  # training/coarse.py:122 (`len(scores) == len(all_cells)`)
    dev = model.device
    engine = model.engine

    # ---- encode: queries first, then cells (the reference's traversal order, :91-113), rows stay on the GPU
    text_enc = torch.zeros((n_q, d), dtype=torch.float32, device=dev)
    cell_enc = torch.zeros((n_c, d), dtype=torch.float32, device=dev)
    query_cell_ids, db_cell_ids = [], []
    off = 0
    for batch in dataloader:
        enc = model.encode_text(batch["texts"])
        text_enc[off:off + len(enc)] = enc
        query_cell_ids.extend(batch["cell_ids"])
        off += len(enc)
    packed = _cached_database(model, cells_dataset)
    if packed is not None:
        # the database was packed once (vectorised) and is encoded in one engine call
        cell_enc[:] = model.encode_cells_packed(packed.pts, packed.meta, packed.cell_ptr)
        db_cell_ids = list(packed.cell_ids)
    else:
        off = 0
        for batch in cells_dataloader:
            enc = model.encode_objects(batch["objects"], batch["object_points"])
            cell_enc[off:off + len(enc)] = enc
            db_cell_ids.extend(batch["cell_ids"])
            off += len(enc)
    query_cell_ids = np.array(query_cell_ids, dtype="<U32")
    db_cell_ids = np.array(db_cell_ids, dtype="<U32")

    # ---- search (:119-125)
    k_max = min(int(np.max(args.top_k)), n_c)
    engine.db_build(cell_enc)
    idx_dev, score_dev, _ = engine.search_topk(text_enc, k_max)

    # ---- bookkeeping for every query at once (:127-150)
    top_k = sorted(int(k) for k in args.top_k)
    cell_size = cells_dataset.cells[0].cell_size
    cell_centres = np.array([cell.get_center()[0:2] for cell in cells_dataset.cells], dtype=np.float64)  # database row order
    query_xy = np.array([pose.pose_w[0:2] for pose in dataset.all_poses], dtype=np.float64)
    hit, close, dists = engine.topk_accuracy(
        idx_dev, query_xy, cell_centres, top_k, threshs=[cell_size / 2], target_row=rows_of_ids(db_cell_ids, query_cell_ids),
        want_dists=return_distance)
    # count / n in float64, which is what np.mean of the reference's per-query boolean lists evaluates to
    hit_rate = hit.sum(dim=0, dtype=torch.int64).cpu().numpy() / n_q
    close_rate = close[:, :, 0].sum(dim=0, dtype=torch.int64).cpu().numpy() / n_q
    accuracies = {k: hit_rate[top_k.index(int(k))] for k in args.top_k}
    accuracies_close = {k: close_rate[top_k.index(int(k))] for k in args.top_k}
    retrieved_ids = db_cell_ids[idx_dev.cpu().numpy()]  # [n_q, k_max] strings, best first (:128)
    top_retrievals = dict(enumerate(retrieved_ids))

    if return_encodings or return_distance:
        cell_encodings = cell_enc.cpu().numpy().astype(np.float64)  # f32 values in f64 buffers, as :81,84
        text_encodings = text_enc.cpu().numpy().astype(np.float64)
    if return_encodings:
        return accuracies, accuracies_close, top_retrievals, cell_encodings, text_encodings
    elif return_distance:
        return accuracies, accuracies_close, top_retrievals, cell_encodings, text_encodings, dists.cpu().numpy(), score_dev.cpu().numpy()
    return accuracies, accuracies_close, top_retrievals


def localisation_accuracies(engine, poses, all_cells, retrievals, pos_in_cells, top_k, threshs):
    """evaluation/utils.py:31-54 (calc_sample_accuracies) for all samples in one launch.

    retrievals [n, k] cell-id strings; pos_in_cells [n, k, 2] predicted position inside each retrieved
    cell as a fraction of the cell size.  Returns {k: {t: mean accuracy}}."""
    retrievals = np.asarray(retrievals)
    n, k = retrievals.shape
    assert k == max(top_k) == pos_in_cells.shape[1]  # evaluation/utils.py:33
    all_ids = np.array([cell.id for cell in all_cells], dtype="<U32")
    rows = rows_of_ids(all_ids, retrievals)
    assert (rows >= 0).all(), "retrieved an id that is not in dataset.all_cells"
    origin = np.array([cell.bbox_w[0:2] for cell in all_cells], dtype=np.float64)
    size = np.array([cell.cell_size for cell in all_cells], dtype=np.float64)
    # the credited position depends on (sample, slot), not only on the cell: one pseudo-row per retrieved slot
    pred_w = origin[rows] + np.asarray(pos_in_cells, dtype=np.float64) * size[rows][..., None]  # :37-42
    _, codes = np.unique(np.concatenate([scene_names(all_ids), scene_names([p.cell_id for p in poses])]), return_inverse=True)
    cell_scene, pose_scene = codes[:len(all_ids)], codes[len(all_ids):]
    top_k = sorted(int(x) for x in top_k)
    _, within, _ = engine.topk_accuracy(
        np.arange(n * k, dtype=np.int64).reshape(n, k), np.array([p.pose_w[0:2] for p in poses], dtype=np.float64),
        pred_w.reshape(n * k, 2), top_k, threshs=list(threshs), query_scene=pose_scene.astype(np.int32),
        cell_scene=cell_scene[rows].reshape(-1).astype(np.int32))
    rate = within.sum(dim=0, dtype=torch.int64).cpu().numpy() / n
    return {kk: {t: rate[top_k.index(int(kk)), j] for j, t in enumerate(threshs)} for kk in top_k}


@torch.no_grad()
def run_coarse(model, dataloader, args, verbose: bool = True):
    """evaluation/coarse.py:40-84 / evaluation/pipeline.py:40-87: text-to-cell retrieval ->
    (retrievals: list of ndarray<U32>[max(top_k)], best first; accuracies {k: {thresh: float}})."""
    model.eval()
    retrieval_accuracies, retrieval_accuracies_close, retrievals = eval_epoch(model, dataloader, args)
    retrievals = [retrievals[idx] for idx in range(len(retrievals))]  # Dict -> list
    if verbose:
        print("Retrieval Accs:")
        print(retrieval_accuracies)
        print("Retrieval Accs Close:")
        print(retrieval_accuracies_close)
    dataset = dataloader.dataset
    assert len(retrievals) == len(dataset.all_poses)  # evaluation/coarse.py:66
    pos_in_cells = np.full((len(retrievals), len(retrievals[0]), 2), 0.5)  # predict cell centres (:74)
    accuracies = localisation_accuracies(model.engine, dataset.all_poses, dataset.all_cells, np.stack(retrievals), pos_in_cells,
                                         args.top_k, args.threshs)
    return retrievals, {k: accuracies[int(k)] for k in args.top_k}


def _hint_text(pose) -> str:
    """The description run_fine feeds the model (dataloading/kitti360pose/eval.py:162-164, base.py:60-68)."""
    if hasattr(pose, "descriptions"):
        return " ".join(f"The pose is {d.direction} of a {d.object_color_text} {d.object_label}." for d in pose.descriptions)
    return pose.text


def _padded_objects(cell, pad_size: int):
    """Cut / pad a cell's object list to pad_size (dataloading/kitti360pose/eval.py:147-160); padding objects are 8 near-zero
    points, black (datapreparation/kitti360pose/imports.py:74-83), created by the objects' own class when it knows how."""
    objects = list(cell.objects)[:pad_size]
    maker = getattr(type(objects[0]), "create_padding", None) if objects else None
    while len(objects) < pad_size:
        objects.append(maker() if maker else dataio.PaddingObject())
    return objects


class _CellView:
    """A padded object list seen as a cell by dataio.pack_cell_database."""

    def __init__(self, objects, cell_id=None):
        self.objects, self.id = objects, cell_id


@torch.no_grad()
def run_fine(model, retrievals, dataloader, args, transform_fine, return_offsets: bool = False):
    """evaluation/pipeline.py:91-204: offsets of every query against each of its max(top_k) retrieved cells, then the
    thresholded localisation accuracy at the predicted in-cell positions -> {k: {thresh: accuracy}}.

    The reference calls the model once per query on max(top_k) re-padded, re-sampled copies of the retrieved cells
    (a Python loop over queries, "using a dataloader does not make it much faster").  Here the object branch does not
    depend on the query, so every DISTINCT retrieved cell is padded, sampled and encoded once
    (t2l_fine_encode_objects), every query's hints once (t2l_fine_encode_hints), and all query x cell pairs go through
    the cross-attention stack in one batched call (t2l_fine_match)."""
    model.eval()
    dataset = dataloader.dataset
    poses, all_cells = dataset.all_poses, dataset.all_cells
    retrievals = np.asarray(retrievals)
    n_q, k = retrievals.shape
    assert n_q == len(poses) and k == max(args.top_k)  # dataloading/kitti360pose/eval.py:133-134
    engine = model.engine
    pad = args.pad_size
    all_ids = np.array([cell.id for cell in all_cells], dtype="<U32")
    rows = rows_of_ids(all_ids, retrievals)  # [n_q, k] rows of all_cells
    assert (rows >= 0).all()
    used, pair_cell = np.unique(rows.reshape(-1), return_inverse=True)  # distinct retrieved cells, pair -> its slot

    # ---- object branch, once per distinct cell, in batches
    obj_emb = torch.empty((len(used) * pad, engine.FINE_DIM), dtype=torch.float32, device=engine.device)
    batch = max(1, 4096 // pad)
    for b0 in range(0, len(used), batch):
        cells = [all_cells[int(r)] for r in used[b0:b0 + batch]]
        objects = [_padded_objects(c, pad) for c in cells]
        if getattr(model, "vectorised_packing", False):
            # opt-in: all objects of the batch sampled and reduced in one vectorised pass instead of
            # one T.FixedPoints call + three numpy reductions per object (the draw differs from the per-object sequence)
            num, normalize_scale = _transform_spec(transform_fine)
            packed = dataio.pack_cell_database([_CellView(o) for o in objects], num, normalize_scale=normalize_scale)
            pts, meta, cell_ptr = packed.pts, packed.meta, packed.cell_ptr
        else:
            points = [dataio.batch_object_points(o, transform_fine) for o in objects]
            pts, meta, cell_ptr = dataio.pack_cells(objects, points)
        obj_emb[b0 * pad:(b0 + len(cells)) * pad] = engine.fine_encode_objects(pts, meta, cell_ptr)

    # ---- textual branch, once per query
    hints, n_hints = [], None
    qb = 256
    for q0 in range(0, n_q, qb):
        rows, nh = model.encode_hints([_hint_text(p) for p in poses[q0:q0 + qb]])
        assert n_hints in (None, nh), "every description must have the same number of hints (language_encoder.py:114)"
        n_hints = nh
        hints.append(rows)
    hints = torch.cat(hints)

    # ---- all query x retrieved-cell pairs
    pair_query = np.repeat(np.arange(n_q, dtype=np.int32), k)
    offsets = engine.fine_match(obj_emb, pair_cell.astype(np.int32), hints, pair_query, pad, n_hints).view(n_q, k, 2)
    offsets_np = offsets.cpu().numpy().astype(np.float64)
    acc = localisation_accuracies(engine, poses, all_cells, retrievals, offsets_np, args.top_k, args.threshs)
    acc = {kk: acc[int(kk)] for kk in args.top_k}
    return (acc, offsets_np) if return_offsets else acc
