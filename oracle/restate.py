"""Standalone CPU restatement of the coarse cell-retrieval path.  TEST INFRASTRUCTURE.

Dense, batched torch-fp32 / numpy-fp64 restatement that needs neither /root/reference nor
PyG, so it travels to the GPU box.  Each function cites the reference lines it follows.  It
is pinned (tests/test_oracle.py) against tests/golden/*.npz, which oracle/make_golden.py
produced by running the reference's own modules (oracle/reference_run.py).

Inputs are the engine's packed layout:
    pts f32 [n, 256, 6] (xyz ‖ rgb of the 256-sample), meta f32 [n, 7] (mean rgb ‖ centre ‖
    raw count), cell_ptr i32 [B+1]; text: t5 f32 [nq*S, L, 1024].
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .pyg_ops import MAX_NUM_NEIGHBORS, _sqdist

NUM_MEAN = 1826.6844940968194  # models/object_encoder.py:44
NUM_STD = 2516.8905096993817  # models/object_encoder.py:45
SA_CONFIG = ((0.5, 0.2), (0.5, 0.3), (0.5, 0.4))  # models/pointcloud/pointnet2.py:57-59


def _t(sd, key):
    return torch.as_tensor(np.asarray(sd[key]))


# ---- building blocks -------------------------------------------------------------------

def mlp(sd, prefix, x, n_layers, last_relu=True):
    """get_mlp / get_mlp2 in eval mode: Linear -> BatchNorm1d(running stats) -> ReLU
    (models/language_encoder.py:16-74; trailing ReLU for get_mlp)."""
    for i in range(n_layers):
        x = F.linear(x, _t(sd, f"{prefix}.{i}.0.weight"), _t(sd, f"{prefix}.{i}.0.bias"))
        x = F.batch_norm(
            x, _t(sd, f"{prefix}.{i}.1.running_mean"), _t(sd, f"{prefix}.{i}.1.running_var"),
            _t(sd, f"{prefix}.{i}.1.weight"), _t(sd, f"{prefix}.{i}.1.bias"), training=False, eps=1e-5,
        )
        if last_relu or i < n_layers - 1:
            x = F.relu(x)
    return x


def encoder_layer(sd, prefix, x, n_heads):
    """nn.TransformerEncoderLayer defaults (post-norm, ReLU, eps 1e-5, eval), input [S, B, d],
    NO mask (models/cell_retrieval.py:35,101-103; models/language_encoder.py:98,103,130-131,144-145)."""
    S, B, d = x.shape
    hd = d // n_heads
    qkv = F.linear(x, _t(sd, prefix + ".self_attn.in_proj_weight"), _t(sd, prefix + ".self_attn.in_proj_bias"))
    q, k, v = qkv.split(d, dim=-1)

    def heads(t):  # [S, B, d] -> [B, H, S, hd]
        return t.reshape(S, B, n_heads, hd).permute(1, 2, 0, 3)

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(S, B, d)
    o = F.linear(o, _t(sd, prefix + ".self_attn.out_proj.weight"), _t(sd, prefix + ".self_attn.out_proj.bias"))
    x = F.layer_norm(x + o, (d,), _t(sd, prefix + ".norm1.weight"), _t(sd, prefix + ".norm1.bias"), 1e-5)
    f = F.linear(F.relu(F.linear(x, _t(sd, prefix + ".linear1.weight"), _t(sd, prefix + ".linear1.bias"))),
                 _t(sd, prefix + ".linear2.weight"), _t(sd, prefix + ".linear2.bias"))
    return F.layer_norm(x + f, (d,), _t(sd, prefix + ".norm2.weight"), _t(sd, prefix + ".norm2.bias"), 1e-5)


# ---- PointNet++ (models/pointcloud/pointnet2.py:18-104) ------------------------------------

def fps_dense(pos: torch.Tensor, m: int) -> torch.Tensor:
    """pos [n, P, 3] -> local indices [n, m]; start 0, first-max ties (oracle/pyg_ops.py::fps)."""
    n, P, _ = pos.shape
    dist = torch.full((n, P), float("inf"))
    idx = torch.zeros((n, m), dtype=torch.long)
    cur = torch.zeros(n, dtype=torch.long)
    ar = torch.arange(n)
    for s in range(1, m):
        dist = torch.minimum(dist, _sqdist(pos, pos[ar, cur][:, None, :]))
        cur = torch.argmax(dist, dim=1)  # first maximum
        idx[:, s] = cur
    return idx


def ball_query_dense(pos: torch.Tensor, cpos: torch.Tensor, r: float):
    """First <=32 in-range points per centroid, ascending index, strict d < r*r with r*r rounded
    from double (oracle/pyg_ops.py::radius).  Returns nbr [n, M, 32] (local idx, -1 pad)."""
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float64).float()
    d = _sqdist(pos[:, None, :, :], cpos[:, :, None, :])  # [n, M, P]
    mask = d < r2
    rank = torch.cumsum(mask.to(torch.int32), dim=2)
    keep = mask & (rank <= MAX_NUM_NEIGHBORS)
    n, M, P = keep.shape
    nbr = torch.full((n, M, MAX_NUM_NEIGHBORS), -1, dtype=torch.long)
    ni, mi, pi = torch.nonzero(keep, as_tuple=True)
    nbr[ni, mi, (rank[ni, mi, pi] - 1).long()] = pi
    return nbr


def set_abstraction(sd, prefix, x, pos, loop, ratio, r):
    """SetAbstractionLayer.forward (pointnet2.py:25-37) for all objects at once.

    x [n, P, C], pos [n, P, 3].  loop = (src_obj, local_obj) from make_loop_src: src_obj[o] is
    the flat index of the object whose dense point feeds the re-added self loop of object o
    (PyG add_self_loops quirk on per-cell global indices): object b of a cell
    takes dense point (b%2)*M + m of object b//2 of the same cell.
    """
    loop_src_obj, local_obj = loop
    n, P, C = x.shape
    M = int(math.ceil(ratio * P))
    idx = fps_dense(pos, M)
    ar = torch.arange(n)[:, None]
    cpos = pos[ar, idx]  # [n, M, 3]
    nbr = ball_query_dense(pos, cpos, r)  # [n, M, 32]
    valid = nbr >= 0
    g = nbr.clamp(min=0)
    xj = x[ar[:, :, None], g]  # [n, M, 32, C]
    pj = pos[ar[:, :, None], g]
    msg = torch.cat([xj, pj - cpos[:, :, None, :]], dim=-1)
    h = mlp(sd, prefix, msg.reshape(-1, C + 3), 2).reshape(n, M, MAX_NUM_NEIGHBORS, -1)
    h = torch.where(valid[..., None], h, torch.full_like(h, float("-inf"))).max(dim=2)[0]
    # the re-added "self loop": dense point with the same per-cell global index as the centroid
    sp = (local_obj % 2)[:, None] * M + torch.arange(M)[None, :]  # [n, M]
    so = loop_src_obj[:, None].expand(n, M)
    msg2 = torch.cat([x[so, sp], pos[so, sp] - cpos], dim=-1)
    h2 = mlp(sd, prefix, msg2.reshape(-1, C + 3), 2).reshape(n, M, -1)
    return torch.maximum(h, h2), cpos, idx, nbr


def make_loop_src(cell_ptr: np.ndarray):
    """(src_obj [n], local_obj [n]) for packed objects delimited by cell_ptr."""
    cp = np.asarray(cell_ptr, dtype=np.int64)
    n = int(cp[-1])
    cell_of = np.repeat(np.arange(len(cp) - 1), np.diff(cp))
    base = cp[cell_of]
    local = np.arange(n) - base
    return torch.from_numpy(base + local // 2), torch.from_numpy(local)


def pointnet2_features2(sd, pts: torch.Tensor, cell_ptr, return_aux=False, chunk=64):
    """PointNet2.forward(...).features2 (pointnet2.py:80-90) for packed objects; the cell is the
    unit of independent work, so chunks are cut on cell boundaries."""
    pn = "object_encoder.pointnet"
    cp = np.asarray(cell_ptr, dtype=np.int64)
    outs, aux = [], []
    c0 = 0
    while c0 < len(cp) - 1:
        c1 = c0 + 1
        while c1 < len(cp) - 1 and cp[c1 + 1] - cp[c0] <= chunk:
            c1 += 1
        o0, o1 = int(cp[c0]), int(cp[c1])
        p = pts[o0:o1]
        loop = make_loop_src(cp[c0:c1 + 1] - cp[c0])
        pos, x = p[:, :, 0:3].contiguous(), p[:, :, 3:6].contiguous()  # data.pos = xyz, data.x = rgb
        a = {}
        for li, (ratio, r) in enumerate(SA_CONFIG):
            x, pos, idx, nbr = set_abstraction(sd, f"{pn}.sa{li + 1}.point_conv.local_nn", x, pos, loop, ratio, r)
            a[f"fps{li + 1}"], a[f"nbr{li + 1}"] = idx, nbr
        n, M, C = x.shape
        # GlobalAbstractionLayer (pointnet2.py:45-49): mlp(cat(x, pos)) then per-object max
        g = mlp(sd, f"{pn}.ga.mlp", torch.cat([x, pos], dim=-1).reshape(n * M, C + 3), 2).reshape(n, M, -1).max(dim=1)[0]
        f1 = F.relu(F.linear(g, _t(sd, f"{pn}.lin1.weight"), _t(sd, f"{pn}.lin1.bias")))
        f2 = F.relu(F.linear(f1, _t(sd, f"{pn}.lin2.weight"), _t(sd, f"{pn}.lin2.bias")))
        outs.append(f2)
        aux.append(a)
        c0 = c1
    f2 = torch.cat(outs)
    if return_aux:
        keys = aux[0].keys()
        return f2, {k: torch.cat([a[k] for a in aux]) for k in keys}
    return f2


# ---- object encoder + cell aggregation ---------------------------------------------------

def object_embeddings(sd, features2: torch.Tensor, meta: torch.Tensor) -> torch.Tensor:
    """ObjectEncoder.forward, branch class_embed=color_embed=False, 4 features
    (models/object_encoder.py:98-149)."""
    oe = "object_encoder"
    feats = [
        F.normalize(mlp(sd, f"{oe}.mlp_pointnet", features2, 1), dim=-1),
        F.normalize(mlp(sd, f"{oe}.color_encoder", meta[:, 0:3], 2), dim=-1),
        F.normalize(mlp(sd, f"{oe}.pos_encoder", meta[:, 3:6], 2), dim=-1),
        F.normalize(mlp(sd, f"{oe}.num_encoder", (meta[:, 6:7] - NUM_MEAN) / NUM_STD, 2), dim=-1),
    ]
    return mlp(sd, f"{oe}.mlp_merge", torch.cat(feats, dim=-1), 1)


def aggregate_cells(sd, emb: torch.Tensor, cell_ptr, object_size=28, n_heads=4, n_layers=2) -> torch.Tensor:
    """CellRetrievalNetwork.encode_objects after the object encoder (cell_retrieval.py:85-108):
    normalise, first <=28 objects into a zero-padded [B,28,256], 2 unmasked encoder layers,
    max over slots, normalise."""
    cp = np.asarray(cell_ptr, dtype=np.int64)
    B = len(cp) - 1
    emb = F.normalize(emb, dim=-1)
    x = torch.zeros(B, object_size, emb.shape[1])
    for c in range(B):
        k = min(int(cp[c + 1] - cp[c]), object_size)
        x[c, :k] = emb[cp[c]:cp[c] + k]
    x = x.permute(1, 0, 2).contiguous()
    for i in range(n_layers):
        x = encoder_layer(sd, f"obj_inter_module.{i}", x, n_heads)
    return F.normalize(x.max(dim=0)[0])


@torch.no_grad()
def encode_cells(sd, pts, meta, cell_ptr, return_aux=False):
    pts = torch.as_tensor(np.asarray(pts), dtype=torch.float32)
    meta = torch.as_tensor(np.asarray(meta), dtype=torch.float32)
    f2, aux = pointnet2_features2(sd, pts, cell_ptr, return_aux=True)
    emb = object_embeddings(sd, f2, meta)
    out = aggregate_cells(sd, emb, cell_ptr)
    if return_aux:
        aux.update(features2=f2, object_emb=emb)
        return out, aux
    return out


# ---- text head (models/language_encoder.py:125-148, models/cell_retrieval.py:57-63) -------

@torch.no_grad()
def encode_text(sd, t5, n_sent: int, n_heads=4, chunk=256):
    """t5 f32 [nq*S, L, 1024] (last_hidden_state) -> unit rows [nq, 256]."""
    t5 = torch.as_tensor(np.asarray(t5), dtype=torch.float32)
    le = "language_encoder"
    pooled = []
    for i in range(0, t5.shape[0], chunk):
        x = t5[i:i + chunk].permute(1, 0, 2)  # [L, B*S, 1024]
        x = encoder_layer(sd, f"{le}.intra_module.0", x, n_heads)
        pooled.append(x.permute(1, 0, 2).max(dim=1)[0])  # max over tokens, pads included
    x = mlp(sd, f"{le}.inter_mlp", torch.cat(pooled), 1, last_relu=False)  # get_mlp2: Linear+BN, no ReLU
    nq = x.shape[0] // n_sent
    x = x.view(nq, n_sent, -1).permute(1, 0, 2)  # [S, nq, 256]
    x = x + encoder_layer(sd, f"{le}.inter_module.0", x, n_heads)  # `x += layer(x)` (:144-145)
    return F.normalize(x.max(dim=0)[0])


# ---- fine stage (models/cross_matcher.py:83-129), SURVEY.md section 8f row 1 -------------------------------

def _mha(sd, prefix, q_in, kv_in, n_heads):
    """nn.MultiheadAttention forward (no masks, eval): q_in [Lq, B, d], kv_in [Lk, B, d]."""
    Lq, B, d = q_in.shape
    Lk = kv_in.shape[0]
    hd = d // n_heads
    w, b = _t(sd, prefix + ".in_proj_weight"), _t(sd, prefix + ".in_proj_bias")
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, w[2 * d:], b[2 * d:])
    q = q.reshape(Lq, B, n_heads, hd).permute(1, 2, 0, 3)
    k = k.reshape(Lk, B, n_heads, hd).permute(1, 2, 0, 3)
    v = v.reshape(Lk, B, n_heads, hd).permute(1, 2, 0, 3)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(Lq, B, d)
    return F.linear(o, _t(sd, prefix + ".out_proj.weight"), _t(sd, prefix + ".out_proj.bias"))


def decoder_layer(sd, prefix, tgt, memory, n_heads):
    """nn.TransformerDecoderLayer defaults (post-norm, ReLU, eps 1e-5, eval, no masks), inputs [L, B, d]
    (models/cross_matcher.py:66-72, :113-115)."""
    d = tgt.shape[-1]
    ln = lambda x, n: F.layer_norm(x, (d,), _t(sd, f"{prefix}.{n}.weight"), _t(sd, f"{prefix}.{n}.bias"), 1e-5)
    x = ln(tgt + _mha(sd, prefix + ".self_attn", tgt, tgt, n_heads), "norm1")
    x = ln(x + _mha(sd, prefix + ".multihead_attn", x, memory, n_heads), "norm2")
    f = F.linear(F.relu(F.linear(x, _t(sd, prefix + ".linear1.weight"), _t(sd, prefix + ".linear1.bias"))),
                 _t(sd, prefix + ".linear2.weight"), _t(sd, prefix + ".linear2.bias"))
    return ln(x + f, "norm3")


@torch.no_grad()
def fine_offsets(sd, pts, meta, cell_ptr, t5, n_hints: int, n_heads=4, n_layers=2):
    """CrossMatch.forward (models/cross_matcher.py:83-129) on packed inputs: every cell holds the same number of
    (padded) objects; t5 f32 [B*n_hints, L, 1024] are the hint sentences' T5 states.  Returns offsets [B, 2]."""
    pts = torch.as_tensor(np.asarray(pts), dtype=torch.float32)
    meta = torch.as_tensor(np.asarray(meta), dtype=torch.float32)
    t5 = torch.as_tensor(np.asarray(t5), dtype=torch.float32)
    cp = np.asarray(cell_ptr, dtype=np.int64)
    B = len(cp) - 1
    n_obj = int(cp[1] - cp[0])
    assert (np.diff(cp) == n_obj).all(), "the fine stage pads every cell to pad_size objects"
    # textual branch: LanguageEncoder(is_fine=True) = token layer, max over tokens, inter_mlp (language_encoder.py:130-140)
    le = "language_encoder"
    x = encoder_layer(sd, f"{le}.intra_module.0", t5.permute(1, 0, 2), n_heads)
    hints = mlp(sd, f"{le}.inter_mlp", x.permute(1, 0, 2).max(dim=1)[0], 1, last_relu=False).view(B, n_hints, -1)
    # 3D branch: ObjectEncoder at d = 128, per-object normalise (:98-108)
    f2 = pointnet2_features2(sd, pts, cell_ptr)
    obj = F.normalize(object_embeddings(sd, f2, meta), dim=-1).view(B, n_obj, -1)
    # cascaded cross-attention (:113-121)
    desc0, desc1 = obj.transpose(0, 1), hints.transpose(0, 1)
    for i in range(n_layers):
        desc0 = decoder_layer(sd, f"cross_objects.{i}", desc0, desc1, n_heads)
        desc1 = decoder_layer(sd, f"cross_hints.{i}", desc1, desc0, n_heads)
    h = desc1.max(dim=0)[0]
    h = F.relu(F.linear(h, _t(sd, "mlp_offsets.0.weight"), _t(sd, "mlp_offsets.0.bias")))
    return F.linear(h, _t(sd, "mlp_offsets.2.weight"), _t(sd, "mlp_offsets.2.bias"))


# ---- search (training/coarse.py:81-125) -----------------------------------------------------

def search_topk(cell_enc, text_enc, k: int):
    """scores = D_f64 @ q_f64 per query; order by (score desc, row index asc) -- the
    reference's np.argsort(-scores) default kind leaves tie order unspecified, the oracle pins
    it with a stable sort.  Returns (idx int64 [nq,k], score f64 [nq,k]).

    The product is evaluated row by row (einsum, no BLAS): a threaded dgemv may sum identical
    database rows in different orders (seen on the 16-core GPU host: 301 duplicate rows got
    scores 1 ulp apart), which would break exact ties by rounding noise instead of by row index."""
    D = np.asarray(cell_enc, dtype=np.float64)  # f32 values widened, as np.zeros(...) does (:81,84)
    Q = np.asarray(text_enc, dtype=np.float64)
    k = min(k, D.shape[0])
    idx = np.empty((Q.shape[0], k), np.int64)
    sc = np.empty((Q.shape[0], k), np.float64)
    for q in range(Q.shape[0]):
        s = np.einsum("ij,j->i", D, Q[q])
        o = np.argsort(-1.0 * s, kind="stable")[:k]
        idx[q], sc[q] = o, s[o]
    return idx, sc


def search_topk_reference_loop(cell_enc64: np.ndarray, text_enc64: np.ndarray, k: int):
    """The reference loop verbatim in behaviour (training/coarse.py:119-125), default argsort
    kind; used as the timed CPU baseline and to cross-check search_topk on tie-free data."""
    out = np.empty((len(text_enc64), k), np.int64)
    for query_idx in range(len(text_enc64)):
        scores = cell_enc64[:] @ text_enc64[query_idx]
        sorted_indices = np.argsort(-1.0 * scores)
        out[query_idx] = sorted_indices[0:k]
    return out


# ---- eval_epoch / run_coarse (training/coarse.py:63-157, evaluation/coarse.py:40-84) --------

def eval_epoch(sd, dataloader, args, frontend, return_encodings=False):
    """Oracle mirror of eval_epoch: same dataloader traversal order (so the unseeded global
    numpy RNG behind FixedPoints is consumed identically), same f64 buffers, same outputs."""
    from torch.utils.data import DataLoader

    from text2loc_b200 import dataio

    cells_dataset = dataloader.dataset.get_cell_dataset()
    cells_loader = DataLoader(cells_dataset, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    cells_dict = {cell.id: cell for cell in cells_dataset.cells}
    cell_size = cells_dataset.cells[0].cell_size
    d = 256
    cell_encodings = np.zeros((len(cells_dataset), d))
    db_cell_ids = np.zeros(len(cells_dataset), dtype="<U32")
    text_encodings = np.zeros((len(dataloader.dataset), d))
    query_cell_ids = np.zeros(len(dataloader.dataset), dtype="<U32")
    query_poses_w = np.array([pose.pose_w[0:2] for pose in dataloader.dataset.all_poses])
    off = 0
    for batch in dataloader:
        feat, n_sent = frontend(batch["texts"])
        enc = encode_text(sd, feat, n_sent).numpy()
        text_encodings[off:off + len(enc)] = enc
        query_cell_ids[off:off + len(enc)] = np.array(batch["cell_ids"])
        off += len(enc)
    off = 0
    for batch in cells_loader:
        pts, meta, cell_ptr = dataio.pack_cells(batch["objects"], batch["object_points"])
        enc = encode_cells(sd, pts, meta, cell_ptr).numpy()
        cell_encodings[off:off + len(enc)] = enc
        db_cell_ids[off:off + len(enc)] = np.array(batch["cell_ids"])
        off += len(enc)
    k_max = int(np.max(args.top_k))
    idx, _ = search_topk(cell_encodings, text_encodings, k_max)
    accuracies = {k: [] for k in args.top_k}
    accuracies_close = {k: [] for k in args.top_k}
    top_retrievals = {}
    for q in range(len(text_encodings)):
        ids = db_cell_ids[idx[q]]
        for k in args.top_k:
            accuracies[k].append(query_cell_ids[q] in ids[0:k])
        top_retrievals[q] = ids
        poses = [cells_dict[c].get_center()[0:2] for c in ids]
        dists = np.linalg.norm(query_poses_w[q] - poses, axis=1)
        for k in args.top_k:
            accuracies_close[k].append(np.any(dists[0:k] <= cell_size / 2))
    for k in args.top_k:
        accuracies[k] = np.mean(accuracies[k])
        accuracies_close[k] = np.mean(accuracies_close[k])
    if return_encodings:
        return accuracies, accuracies_close, top_retrievals, cell_encodings, text_encodings
    return accuracies, accuracies_close, top_retrievals


# ---- accuracy bookkeeping, the reference's per-query loops (checkers of t2l_topk_accuracy) -------------------

def retrieval_accuracies(retrieved_ids, query_cell_ids, query_poses_w, cells_dict, cell_size, top_k):
    """training/coarse.py:127-150 given the retrieved id lists: ({k: hit rate}, {k: close-by rate}, dists [nq, k_max])."""
    accuracies = {k: [] for k in top_k}
    accuracies_close = {k: [] for k in top_k}
    all_dists = []
    for q in range(len(retrieved_ids)):
        ids = retrieved_ids[q]
        for k in top_k:
            accuracies[k].append(query_cell_ids[q] in ids[0:k])
        poses = [cells_dict[c].get_center()[0:2] for c in ids]
        dists = np.linalg.norm(query_poses_w[q] - poses, axis=1)
        all_dists.append(dists)
        for k in top_k:
            accuracies_close[k].append(np.any(dists[0:k] <= cell_size / 2))
    return ({k: np.mean(v) for k, v in accuracies.items()}, {k: np.mean(v) for k, v in accuracies_close.items()}, np.stack(all_dists))


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


def localisation_accuracies(poses, all_cells, retrievals, pos_in_cells, top_k, threshs):
    """evaluation/coarse.py:69-82: per-sample calc_sample_accuracies, averaged."""
    cells_dict = {cell.id: cell for cell in all_cells}
    acc = {k: {t: [] for t in threshs} for k in top_k}
    for i in range(len(retrievals)):
        a = calc_sample_accuracies(poses[i], [cells_dict[c] for c in retrievals[i]], pos_in_cells[i], top_k, threshs)
        for k in top_k:
            for t in threshs:
                acc[k][t].append(a[k][t])
    return {k: {t: np.mean(acc[k][t]) for t in threshs} for k in top_k}
