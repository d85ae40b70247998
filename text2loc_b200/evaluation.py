"""Drop-ins for training/coarse.py::eval_epoch and evaluation/coarse.py::run_coarse.

Same signatures and return types as the reference.  What changes is where the work happens:
query and cell encodings stay on the GPU, and the per-query float64 GEMV + full argsort of
training/coarse.py:119-125 becomes one batched t2l_search_topk call whose result is, by
construction, the fp64 (score desc, row asc) order.  The accuracy bookkeeping
(training/coarse.py:127-150, evaluation/coarse.py:61-82) is O(nq*k) host work and stays in
Python, line for line.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import dataio


def calc_sample_accuracies(pose, top_cells, pos_in_cells, top_k, threshs):
    """evaluation/utils.py:31-54."""
    pose_w = pose.pose_w
    assert len(top_cells) == max(top_k) == len(pos_in_cells)
    pred_w = np.array([top_cells[i].bbox_w[0:2] + pos_in_cells[i, :] * top_cells[i].cell_size for i in range(len(top_cells))])
    dists = np.linalg.norm(pose_w[0:2] - pred_w, axis=1)
    pose_scene_name = pose.cell_id.split("_")[0]
    cell_scene_names = np.array([cell.id.split("_")[0] for cell in top_cells])
    dists[pose_scene_name != cell_scene_names] = np.inf
    return {k: {t: np.min(dists[0:k]) <= t for t in threshs} for k in top_k}


@torch.no_grad()
def eval_epoch(model, dataloader, args, return_encodings=False, return_distance=False):
    """training/coarse.py:63-157.  Returns (accuracies{k}, accuracies_close{k},
    top_retrievals{query_idx: ndarray<U32>[max(top_k)]}) [+ cell_encodings, text_encodings f64]
    [+ dists, scores]."""
    assert args.ranking_loss != "triplet"
    model.eval()
    accuracies = {k: [] for k in args.top_k}
    accuracies_close = {k: [] for k in args.top_k}

    cells_dataset = dataloader.dataset.get_cell_dataset()
    cells_dataloader = DataLoader(cells_dataset, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    cells_dict = {cell.id: cell for cell in cells_dataset.cells}
    cell_size = cells_dataset.cells[0].cell_size

    n_q, n_c, d = len(dataloader.dataset), len(cells_dataset), model.embed_dim
    dev = model.device
    cell_enc_dev = torch.zeros((n_c, d), dtype=torch.float32, device=dev)
    text_enc_dev = torch.zeros((n_q, d), dtype=torch.float32, device=dev)
    db_cell_ids = np.zeros(n_c, dtype="<U32")
    query_cell_ids = np.zeros(n_q, dtype="<U32")
    query_poses_w = np.array([pose.pose_w[0:2] for pose in dataloader.dataset.all_poses])

    # Encode the query side (same traversal order as the reference: queries first, then cells)
    index_offset = 0
    for batch in dataloader:
        text_enc = model.encode_text(batch["texts"])
        bs = len(text_enc)
        text_enc_dev[index_offset:index_offset + bs] = text_enc
        query_cell_ids[index_offset:index_offset + bs] = np.array(batch["cell_ids"])
        index_offset += bs

    # Encode the database side
    index_offset = 0
    for batch in cells_dataloader:
        cell_enc = model.encode_objects(batch["objects"], batch["object_points"])
        bs = len(cell_enc)
        cell_enc_dev[index_offset:index_offset + bs] = cell_enc
        db_cell_ids[index_offset:index_offset + bs] = np.array(batch["cell_ids"])
        index_offset += bs

    # Search: scores = D @ q in float64, order high -> low, first max(top_k)  (:119-125)
    k_max = int(np.max(args.top_k))
    assert n_c == len(dataloader.dataset.all_cells)
    engine = model.engine
    engine.db_build(cell_enc_dev)
    idx_dev, score_dev, _ = engine.search_topk(text_enc_dev, min(k_max, n_c))
    sorted_idx = idx_dev.cpu().numpy()
    sorted_scores = score_dev.cpu().numpy()

    top_retrievals = {}
    dists_list, scores_list = [], []
    for query_idx in range(n_q):
        retrieved_cell_ids = db_cell_ids[sorted_idx[query_idx]]
        target_cell_id = query_cell_ids[query_idx]
        for k in args.top_k:
            accuracies[k].append(target_cell_id in retrieved_cell_ids[0:k])
        top_retrievals[query_idx] = retrieved_cell_ids
        # Close-by accuracy
        target_pose_w = query_poses_w[query_idx]
        retrieved_cell_poses = [cells_dict[cell_id].get_center()[0:2] for cell_id in retrieved_cell_ids]
        dists = np.linalg.norm(target_pose_w - retrieved_cell_poses, axis=1)
        if return_distance:
            dists_list.append(dists[0:max(args.top_k)])
            scores_list.append(sorted_scores[query_idx])
        for k in args.top_k:
            accuracies_close[k].append(np.any(dists[0:k] <= cell_size / 2))

    for k in args.top_k:
        accuracies[k] = np.mean(accuracies[k])
        accuracies_close[k] = np.mean(accuracies_close[k])

    if return_encodings or return_distance:
        cell_encodings = cell_enc_dev.cpu().numpy().astype(np.float64)  # f32 values in f64 buffers, as :81,84
        text_encodings = text_enc_dev.cpu().numpy().astype(np.float64)
    if return_encodings:
        return accuracies, accuracies_close, top_retrievals, cell_encodings, text_encodings
    elif return_distance:
        return accuracies, accuracies_close, top_retrievals, cell_encodings, text_encodings, np.stack(dists_list), np.stack(scores_list)
    return accuracies, accuracies_close, top_retrievals


@torch.no_grad()
def run_coarse(model, dataloader, args, verbose: bool = True):
    """evaluation/coarse.py:40-84 / evaluation/pipeline.py:40-87: text-to-cell retrieval ->
    (retrievals: list of ndarray<U32>[max(top_k)], best first; accuracies {k: {thresh: float}})."""
    model.eval()
    all_cells_dict = {cell.id: cell for cell in dataloader.dataset.all_cells}
    retrieval_accuracies, retrieval_accuracies_close, retrievals = eval_epoch(model, dataloader, args)
    retrievals = [retrievals[idx] for idx in range(len(retrievals))]  # Dict -> list
    if verbose:
        print("Retrieval Accs:")
        print(retrieval_accuracies)
        print("Retrieval Accs Close:")
        print(retrieval_accuracies_close)
    assert len(retrievals) == len(dataloader.dataset.all_poses)

    accuracies = {k: {t: [] for t in args.threshs} for k in args.top_k}
    for i_sample in range(len(retrievals)):
        pose = dataloader.dataset.all_poses[i_sample]
        top_cells = [all_cells_dict[cell_id] for cell_id in retrievals[i_sample]]
        pos_in_cells = 0.5 * np.ones((len(top_cells), 2))  # Predict cell-centers
        accs = calc_sample_accuracies(pose, top_cells, pos_in_cells, args.top_k, args.threshs)
        for k in args.top_k:
            for t in args.threshs:
                accuracies[k][t].append(accs[k][t])
    for k in args.top_k:
        for t in args.threshs:
            accuracies[k][t] = np.mean(accuracies[k][t])
    return retrievals, accuracies
