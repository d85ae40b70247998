"""Pin the CPU oracle: restatement vs the golden vectors produced by the reference's own
Python (oracle/make_golden.py), plus hand-checkable micro-cases for the pinned PyG choices."""
import argparse
import hashlib

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle import fake_t5, pyg_ops, reference_run, restate
import synth
from text2loc_b200 import dataio


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_cells_golden(state_dict, golden):
    g = golden("cells_small.npz")
    out, aux = restate.encode_cells(state_dict, g["pts"], g["meta"], g["cell_ptr"], return_aux=True)
    assert np.abs(out.numpy() - g["cell_emb"]).max() < 2e-6  # reference's own encode_objects
    for i in (1, 2, 3):
        assert (aux[f"fps{i}"].numpy() == g[f"fps{i}"]).all()
        assert digest(aux[f"nbr{i}"].numpy().astype(np.int16)) == str(g["nbr_digest"][i - 1])
    assert np.abs(aux["features2"].numpy() - g["features2"]).max() < 1e-5
    # the two NormalizeScale'd cells must exercise the below-cap branch of the ball query
    n_obj = int(g["cell_ptr"][-1])
    assert int(g["nbr_count"][0]) < n_obj * 128 * 32


def test_text_golden(state_dict, golden):
    g = golden("text_small.npz")
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    assert digest(feat.numpy()) == str(g["t5_digest"]) and n_sent == int(g["n_sent"])
    out = restate.encode_text(state_dict, feat, n_sent)
    assert np.abs(out.numpy() - g["text_emb"]).max() < 2e-6


def test_search_golden(golden):
    g = golden("search_small.npz")
    D = synth.make_unit_rows(int(g["d_seed"]), int(g["n"]))
    Q = synth.make_unit_rows(int(g["q_seed"]), int(g["nq"]))
    idx, sc = restate.search_topk(D, Q, 10)
    assert (idx == g["idx"]).all()  # the reference's own argsort loop (tie-free data)
    # the golden scores are the reference's BLAS dgemv; the oracle sums row by row (see search_topk): same fp64
    # products, different summation order
    assert np.abs(sc - g["score"]).max() < 1e-14


def test_fine_stage_golden(golden):
    """Groundwork for SURVEY.md section 8f row 1 (the fine stage): the restatement of CrossMatch.forward is pinned to
    the reference's own module (oracle/make_golden.py fine); no engine code consumes it yet."""
    g = golden("fine_small.npz")
    sd = synth.make_fine_state_dict(int(g["weight_seed"]))
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    assert digest(feat.numpy()) == str(g["t5_digest"]) and n_sent == int(g["n_sent"])
    out = restate.fine_offsets(sd, g["pts"], g["meta"], g["cell_ptr"], feat, n_sent).numpy()
    assert out.shape == g["offsets"].shape == (5, 2)
    assert np.abs(out - g["offsets"]).max() < 2e-6


def test_eval_epoch_golden(state_dict, golden):
    from oracle.make_golden import e2e_dataset

    g = golden("eval_e2e.npz")
    args = argparse.Namespace(top_k=[int(k) for k in g["top_k"]], batch_size=int(g["batch_size"]), ranking_loss="pairwise")
    ds = e2e_dataset()
    loader = DataLoader(ds, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    np.random.seed(int(g["np_seed"]))
    acc, acc_close, retr, cell_enc, text_enc = restate.eval_epoch(
        state_dict, loader, args, fake_t5.FakeFrontend(int(g["fake_t5_seed"])), return_encodings=True)
    assert np.abs(cell_enc - g["cell_enc"]).max() < 2e-6
    assert np.abs(text_enc - g["text_enc"]).max() < 2e-6
    got = np.stack([retr[i] for i in range(len(ds))])
    # identical except where the reference's own k/k+1 score gap is below the 2e-6 encoding noise
    same = (got == g["retrievals"]).all(axis=1)
    assert same.mean() > 0.95
    assert np.allclose([acc[k] for k in args.top_k], g["acc"], atol=0.03)


# ---- micro-cases for the pinned third-party behaviour --------------------------------------

def test_fps_three_points_and_ties():
    pos = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [0.4, 0, 0], [1.0, 0, 0]])
    # start 0; farthest is index 1 (tie with 3 -> lowest index); then 2 (0.16 vs 0.36 -> min 0.16) vs 3 (0)
    assert pyg_ops.fps(pos, None, 0.75).tolist() == [0, 1, 2]
    dense = restate.fps_dense(pos[None], 3)
    assert dense[0].tolist() == [0, 1, 2]


def test_fps_all_duplicates_returns_zero():
    pos = torch.ones(8, 3) * 0.3
    assert pyg_ops.fps(pos, None, 0.5).tolist() == [0, 0, 0, 0]


def test_radius_cap_first_32_ascending_and_strict():
    x = torch.zeros(40, 3)
    x[:, 0] = torch.arange(40) * 1e-3
    y = torch.zeros(1, 3)
    e = pyg_ops.radius(x, y, 0.2)
    assert e.shape[1] == 32 and e[1].tolist() == list(range(32))
    # strict inequality: a point at distance exactly r (in fp32 arithmetic) is excluded
    r = 0.5
    x2 = torch.tensor([[0.5, 0, 0], [0.25, 0, 0]])
    assert pyg_ops.radius(x2, y, r)[1].tolist() == [1]
    nbr = restate.ball_query_dense(x[None], y[None], 0.2)
    assert nbr[0, 0].tolist() == list(range(32))


def test_pointconv_self_loop_quirk_two_objects():
    """Centroid i always receives the dense point with the same per-cell global index i,
    which for object 1 is a point of object 0."""
    torch.manual_seed(0)
    nn = torch.nn.Linear(3 + 3, 4)
    conv = pyg_ops.PointConv(local_nn=nn)
    pos = torch.rand(8, 3)  # two objects of 4 points
    x = torch.rand(8, 3)
    sub = torch.tensor([0, 2, 4, 6])  # centroids: 2 per object
    # no radius edges at all: the output must be exactly the self-loop messages
    out = conv(x, (pos, pos[sub]), torch.zeros(2, 0, dtype=torch.long))
    want = nn(torch.cat([x[:4], pos[:4] - pos[sub]], dim=1))  # dense points 0..3, all of object 0
    assert torch.allclose(out, want)


def test_search_ties_resolve_by_index():
    D = synth.make_unit_rows(1, 50)
    D[7] = D[3]
    D[20] = D[3]
    Q = D[3:4].copy()
    idx, sc = restate.search_topk(D, Q, 5)
    assert idx[0, :3].tolist() == [3, 7, 20] and sc[0, 0] == sc[0, 1] == sc[0, 2]


def test_search_k_larger_than_db():
    D = synth.make_unit_rows(1, 4)
    idx, _ = restate.search_topk(D, D[:2], 10)
    assert idx.shape == (2, 4)


def test_cell_with_more_than_28_objects_ignores_the_tail(state_dict):
    """Objects beyond object_size are encoded then dropped (cell_retrieval.py:94-98)."""
    cells = synth.make_cell_objects(5, 1, [30], max_raw=300)
    pts, meta, ptr = synth.pack_cells(cells, 5)
    f2 = restate.pointnet2_features2(state_dict, torch.from_numpy(pts), ptr)
    emb = restate.object_embeddings(state_dict, f2, torch.from_numpy(meta))
    full = restate.aggregate_cells(state_dict, emb, ptr)
    cut = restate.aggregate_cells(state_dict, emb[:28], np.array([0, 28]))
    assert torch.equal(full, cut)


@pytest.mark.skipif(not reference_run.available(), reason="reference tree not present (GPU box)")
def test_restatement_matches_reference_modules_other_seed():
    sd = synth.make_state_dict(7)
    model = reference_run.build_model(sd)
    cells = synth.make_cell_objects(3, 3, [2, 5, 1], max_raw=300)
    np.random.seed(1)
    batches = [dataio.batch_object_points(o, dataio.FixedPoints(256)) for o in cells]
    with torch.no_grad():
        want = model.encode_objects(cells, batches)
    pts, meta, ptr = dataio.pack_cells(cells, batches)
    got = restate.encode_cells(sd, pts, meta, ptr)
    assert (got - want).abs().max() < 2e-6


@pytest.mark.skipif(not reference_run.available(), reason="reference tree not present (GPU box)")
def test_restatement_text_head_matches_reference_other_shapes():
    """The reference's own CellRetrievalNetwork.encode_text (sentence split, fake tokenizer / T5, LanguageEncoder) on
    descriptions with 1, 3 and 6 sentences of different lengths (so n_sent and the padded token count vary), another
    weight seed: the restatement on the same T5 states agrees to fp32 rounding."""
    from oracle import fake_t5

    sd = synth.make_state_dict(7)
    model = reference_run.build_model(sd)
    rng = np.random.default_rng(12)
    for n_sent in (1, 3, 6):
        texts = [" ".join(f"The pose is {synth.DIRECTIONS[rng.integers(5)]} of a {synth.COLOR_WORDS[rng.integers(8)]} "
                          f"{synth.CLASS_WORDS[rng.integers(22)]}." for _ in range(n_sent)) for _ in range(5)]
        with torch.no_grad():
            want = model.encode_text(texts)
        feats, ns = fake_t5.FakeFrontend(0)(texts)
        assert ns == n_sent and want.shape == (5, 256)
        got = restate.encode_text(sd, feats, ns)
        assert (got - want).abs().max() < 2e-6, n_sent


@pytest.mark.skipif(not reference_run.available(), reason="reference tree not present (GPU box)")
def test_restatement_fine_stage_matches_reference_other_seed():
    """The reference's own CrossMatch.forward with another weight seed, 3 cells padded to 16 objects, 6 hints each."""
    from oracle import fake_t5

    sd = synth.make_fine_state_dict(5)
    model = reference_run.build_fine_model(sd)
    cells = synth.make_cell_objects(41, 3, [16] * 3, max_raw=300)
    np.random.seed(3)
    batches = [dataio.batch_object_points(o, dataio.FixedPoints(256)) for o in cells]
    rng = np.random.default_rng(8)
    texts = [" ".join(f"The pose is {synth.DIRECTIONS[rng.integers(5)]} of a {synth.COLOR_WORDS[rng.integers(8)]} "
                      f"{synth.CLASS_WORDS[rng.integers(22)]}." for _ in range(6)) for _ in range(3)]
    with torch.no_grad():
        want = model(cells, texts, batches)
    pts, meta, ptr = dataio.pack_cells(cells, batches)
    feats, n_hints = fake_t5.FakeFrontend(0)(texts)
    got = restate.fine_offsets(sd, pts, meta, ptr, feats, n_hints)
    assert want.shape == (3, 2) and (got - want).abs().max() < 5e-6


# ---- precision plan: operand rounding emulated on the CPU -----------------------------

def _emulated_text_error(state_dict, rounder, monkeypatch):
    """Row-relative error of the text embeddings when both operands of every Linear of the TOKEN layer are passed
    through `rounder` (what a reduced-precision tensor-core GEMM with fp32 accumulation does); everything else fp32."""
    import torch.nn.functional as F_

    t5 = synth.make_t5_features(5, 16, 6, 12)
    want = restate.encode_text(state_dict, t5, 6).numpy()
    real_linear = F_.linear
    real_layer = restate.encoder_layer

    def rounded_linear(x, w, b=None):
        return real_linear(rounder(x), rounder(w), b)

    def layer(sd, prefix, x, n_heads):
        if "intra_module" in prefix:  # the token layer carries 99 % of the text head's FLOPs
            monkeypatch.setattr(restate.F, "linear", rounded_linear)
            try:
                return real_layer(sd, prefix, x, n_heads)
            finally:
                monkeypatch.setattr(restate.F, "linear", real_linear)
        return real_layer(sd, prefix, x, n_heads)

    monkeypatch.setattr(restate, "encoder_layer", layer)
    got = restate.encode_text(state_dict, t5, 6).numpy()
    monkeypatch.setattr(restate, "encoder_layer", real_layer)
    return float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())


def test_precision_plan_token_layer_operand_rounding(state_dict, monkeypatch):
    """fp16 and tf32 (round-to-nearest) operands have the same 11-bit significand: both keep the text embeddings well
    inside the 1e-3 tolerance; bf16 operands (8 bits) do not -- which is why the engine's token layer is fp16/tf32."""
    def tf32_rna(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    e_f16 = _emulated_text_error(state_dict, lambda t: t.half().float(), monkeypatch)
    e_tf32 = _emulated_text_error(state_dict, tf32_rna, monkeypatch)
    e_bf16 = _emulated_text_error(state_dict, lambda t: t.bfloat16().float(), monkeypatch)
    print(f"\ntoken-layer operand rounding, text embedding error: fp16 {e_f16:.2e}, tf32 {e_tf32:.2e}, bf16 {e_bf16:.2e}")
    assert e_f16 < 5e-4 and e_tf32 < 5e-4
    assert abs(e_f16 - e_tf32) < 2e-4
    assert e_bf16 > 1e-3


def test_precision_plan_token_layer_fp16_residual_stream(state_dict, monkeypatch):
    """The engine's default token layer also carries the residual stream as fp16 rows: fp16 T5 states, q/k/v stored fp16,
    y = r16(x + attn), x1 = r16(LN(y)), y2 = r16(x1 + ffn), statistics and sums in fp32.  Emulating exactly those rounding
    points keeps the text embeddings inside the 1e-3 tolerance with a 2x margin (about 1.5x the error of the fp32 stream)."""
    import math

    import torch.nn.functional as F_

    from oracle.restate import _t

    def r16(t):
        return t.half().float()

    real_layer = restate.encoder_layer

    def make_layer(stream16):
        def layer(sd, prefix, x, n_heads):
            if "intra_module" not in prefix:
                return real_layer(sd, prefix, x, n_heads)
            S, B, d = x.shape
            hd = d // n_heads

            def lin(a, w, b):
                return F_.linear(r16(a), r16(_t(sd, prefix + w)), _t(sd, prefix + b))

            x = r16(x) if stream16 else x
            qkv = r16(lin(x, ".self_attn.in_proj_weight", ".self_attn.in_proj_bias"))
            q, k, v = (t.reshape(S, B, n_heads, hd).permute(1, 2, 0, 3) for t in qkv.split(d, dim=-1))
            att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
            o = lin((att @ v).permute(2, 0, 1, 3).reshape(S, B, d), ".self_attn.out_proj.weight", ".self_attn.out_proj.bias")
            y = r16(x + o) if stream16 else x + o
            x1 = F_.layer_norm(y, (d,), _t(sd, prefix + ".norm1.weight"), _t(sd, prefix + ".norm1.bias"), 1e-5)
            x1 = r16(x1) if stream16 else x1
            f = lin(r16(F_.relu(lin(x1, ".linear1.weight", ".linear1.bias"))), ".linear2.weight", ".linear2.bias")
            y2 = r16(x1 + f) if stream16 else x1 + f
            return F_.layer_norm(y2, (d,), _t(sd, prefix + ".norm2.weight"), _t(sd, prefix + ".norm2.bias"), 1e-5)
        return layer

    t5 = synth.make_t5_features(5, 16, 6, 12)
    want = restate.encode_text(state_dict, t5, 6).numpy()
    errs = {}
    for stream16 in (False, True):
        monkeypatch.setattr(restate, "encoder_layer", make_layer(stream16))
        got = restate.encode_text(state_dict, t5, 6).numpy()
        monkeypatch.setattr(restate, "encoder_layer", real_layer)
        errs[stream16] = float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())
    print(f"\ntoken layer residual stream, text embedding error: fp32 {errs[False]:.2e}, fp16 {errs[True]:.2e}")
    assert errs[False] < 4e-4
    assert errs[True] < 6e-4


def _emulated_cell_error(state_dict, r16, monkeypatch):
    """Row-relative error of the cell embeddings when the set-abstraction and global-abstraction MLPs use the engine's
    rounding points (sa_obj.cu / GA on fp16 operands): Px = r16(W1x x + b1), edge activation r16(relu(Px + W1p.d)),
    W2 r16, fp32 accumulation, layer outputs rounded to tf32; everything else fp32."""
    from text2loc_b200 import weights

    fw = {k: torch.from_numpy(v) for k, v in weights.engine_weights(state_dict).items()}
    real_mlp = restate.mlp

    def tf32(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    def mlp(sd, prefix, x, n_layers, last_relu=True):
        if "point_conv.local_nn" in prefix:
            lvl = prefix.split(".sa")[1][0]
            w1x, w1p, b1 = fw[f"sa{lvl}.w1x"], fw[f"sa{lvl}.w1p"], fw[f"sa{lvl}.b1"][0]
            w2, b2 = fw[f"sa{lvl}.w2"], fw[f"sa{lvl}.b2"][0]
            c = w1x.shape[1]
            xin = x[..., :c] if lvl == "1" else tf32(x[..., :c])  # levels 2/3: tf32 GEMM on the previous level's output
            px = r16(xin @ (w1x if lvl == "1" else tf32(w1x)).T + b1)
            a = r16(torch.relu(px + x[..., c:] @ w1p.T))
            return tf32(torch.relu(a @ r16(w2).T + b2))
        if ".ga.mlp" in prefix:
            g1 = r16(torch.relu(r16(x) @ r16(fw["ga.w1"]).T + fw["ga.b1"][0]))
            return tf32(torch.relu(g1 @ r16(fw["ga.w2"]).T + fw["ga.b2"][0]))
        return real_mlp(sd, prefix, x, n_layers, last_relu)

    cells = synth.make_cell_objects(21, 6, [3, 8, 1, 5, 12, 2], max_raw=400)
    pts, meta, ptr = synth.pack_cells(cells, 21)
    want = restate.encode_cells(state_dict, pts, meta, ptr).numpy()
    monkeypatch.setattr(restate, "mlp", mlp)
    got = restate.encode_cells(state_dict, pts, meta, ptr).numpy()
    monkeypatch.setattr(restate, "mlp", real_mlp)
    return float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())


def test_precision_plan_set_abstraction_operand_rounding(state_dict, monkeypatch):
    """fp16 Px / edge activations / W2 (the object-resident PointConv kernel) keep the cell embeddings inside the
    tolerance with a wide margin; the same rounding points in bf16 cost an order of magnitude more."""
    e_f16 = _emulated_cell_error(state_dict, lambda t: t.half().float(), monkeypatch)
    e_bf16 = _emulated_cell_error(state_dict, lambda t: t.bfloat16().float(), monkeypatch)
    print(f"\nset-abstraction operand rounding, cell embedding error: fp16 {e_f16:.2e}, bf16 {e_bf16:.2e}")
    assert e_f16 < 5e-4
    assert e_bf16 > 3 * e_f16


def test_precision_plan_per_point_first_linear(state_dict, monkeypatch):
    """sa_obj2.cu applies the first Linear of a PointConv once per point and once per centroid (Qx[j] - v[i], both fp16,
    relative to the object's own origin) instead of once per edge in fp32.  Emulating exactly those rounding points on the
    CPU keeps the cell embeddings as close to the fp32 oracle as round 1's per-edge form did."""
    import math

    from oracle.restate import MAX_NUM_NEIGHBORS, ball_query_dense, fps_dense
    from text2loc_b200 import weights

    fw = {k: torch.from_numpy(v) for k, v in weights.engine_weights(state_dict).items()}
    r16 = lambda t: t.half().float()

    def tf32(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    origin = {}

    def make_sa(per_point):
        def set_abstraction(sd_, prefix, x, pos, loop, ratio, r):
            lvl = prefix.split(".sa")[1][0]
            w1x, w1p, b1 = fw[f"sa{lvl}.w1x"], fw[f"sa{lvl}.w1p"], fw[f"sa{lvl}.b1"][0]
            w2, b2 = fw[f"sa{lvl}.w2"], fw[f"sa{lvl}.b2"][0]
            loop_src_obj, local_obj = loop
            n, P, C = x.shape
            M = int(math.ceil(ratio * P))
            idx = fps_dense(pos, M)
            ar = torch.arange(n)[:, None]
            cpos = pos[ar, idx]
            nbr = ball_query_dense(pos, cpos, r)
            valid, g = nbr >= 0, nbr.clamp(min=0)
            if lvl == "1":
                origin["o"] = pos[:, 0, :].clone()  # the object's point 0 = centroid 0 of every level
            o = origin["o"]
            px32 = (x if lvl == "1" else tf32(x)) @ (w1x if lvl == "1" else tf32(w1x)).T + b1
            sp = (local_obj % 2)[:, None] * M + torch.arange(M)[None, :]
            so = loop_src_obj[:, None].expand(n, M)
            if not per_point:  # round 1: Px fp16, exact fp32 position term per edge, one rounding of the sum
                px = r16(px32)
                a = r16(torch.relu(px[ar[:, :, None], g] + (pos[ar[:, :, None], g] - cpos[:, :, None, :]) @ w1p.T))
                a2 = r16(torch.relu(px[so, sp] + (pos[so, sp] - cpos) @ w1p.T))
            else:
                d = pos - o[:, None, :]
                if lvl == "1":
                    posterm = d @ w1p.T
                else:  # tf32 GEMM columns [hi | lo] against [W1p | W1p]
                    dh = tf32(d)
                    posterm = dh @ tf32(w1p).T + tf32(d - dh) @ tf32(w1p).T
                qx = r16(px32 + posterm)
                v = r16((cpos - o[:, None, :]) @ w1p.T)
                a = torch.relu(r16(qx[ar[:, :, None], g] - v[:, :, None, :]))  # fma.rn.relu.f16x2(1, Qx, -v)
                vs = r16((cpos - o[so[:, 0]][:, None, :]) @ w1p.T)  # self loops: relative to the SOURCE object's origin
                a2 = torch.relu(r16(qx[so, sp] - vs))
            h = tf32(torch.relu(a @ r16(w2).T + b2))
            h = torch.where(valid[..., None], h, torch.full_like(h, float("-inf"))).max(dim=2)[0]
            return torch.maximum(h, tf32(torch.relu(a2 @ r16(w2).T + b2))), cpos, idx, nbr
        return set_abstraction

    real_mlp, real_sa = restate.mlp, restate.set_abstraction

    def mlp(sd_, prefix, x, n_layers, last_relu=True):
        if ".ga.mlp" in prefix:
            g1 = r16(torch.relu(r16(x) @ r16(fw["ga.w1"]).T + fw["ga.b1"][0]))
            return tf32(torch.relu(g1 @ r16(fw["ga.w2"]).T + fw["ga.b2"][0]))
        return real_mlp(sd_, prefix, x, n_layers, last_relu)

    cells = synth.make_cell_objects(21, 6, [3, 8, 1, 5, 12, 2], max_raw=400)
    pts, meta, ptr = synth.pack_cells(cells, 21)
    want = restate.encode_cells(state_dict, pts, meta, ptr).numpy()
    errs = {}
    monkeypatch.setattr(restate, "mlp", mlp)
    for name, per_point in (("per edge (round 1)", False), ("per point + per centroid", True)):
        monkeypatch.setattr(restate, "set_abstraction", make_sa(per_point))
        got = restate.encode_cells(state_dict, pts, meta, ptr).numpy()
        errs[name] = float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())
    monkeypatch.setattr(restate, "mlp", real_mlp)
    monkeypatch.setattr(restate, "set_abstraction", real_sa)
    print("\nfirst-Linear formulation, cell embedding error:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["per point + per centroid"] < 3e-4
    assert errs["per point + per centroid"] < 2 * errs["per edge (round 1)"] + 5e-5


def test_padded_slots_are_identical_rows_and_one_weighted_row_stands_for_them(state_dict):
    """The engine's intra-cell layers carry min(n, 28) object rows + ONE padding row per cell instead of the reference's 28
    zero-padded slots (cell_retrieval.py:81-103: zero padding, no mask).  CPU proof on the oracle: (1) after each of the two
    layers the padded slots of a cell are identical rows; (2) a layer evaluated on the reduced rows, with the padding row's
    softmax weight multiplied by its multiplicity as a KEY, equals the full layer; (3) the cell embeddings agree."""
    import math

    import torch.nn.functional as F_

    from oracle.restate import _t

    slots, d, heads = 28, 256, 4
    g = torch.Generator().manual_seed(3)
    counts = [8, 1, 27, 28, 16]
    full = torch.zeros(len(counts), slots, d)
    for c, n in enumerate(counts):
        full[c, :n] = F_.normalize(torch.randn(n, d, generator=g), dim=-1)

    def reduced_layer(prefix, x, mult):  # x [rows, d] of ONE cell; mult = weight of the last row as a key
        hd = d // heads
        qkv = F_.linear(x, _t(state_dict, prefix + ".self_attn.in_proj_weight"), _t(state_dict, prefix + ".self_attn.in_proj_bias"))
        q, k, v = (t.reshape(-1, heads, hd).permute(1, 0, 2) for t in qkv.split(d, dim=-1))
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        e = torch.exp(sc - sc.max(dim=-1, keepdim=True)[0])
        e[..., -1] *= mult
        o = ((e / e.sum(dim=-1, keepdim=True)) @ v).permute(1, 0, 2).reshape(-1, d)
        o = F_.linear(o, _t(state_dict, prefix + ".self_attn.out_proj.weight"), _t(state_dict, prefix + ".self_attn.out_proj.bias"))
        x = F_.layer_norm(x + o, (d,), _t(state_dict, prefix + ".norm1.weight"), _t(state_dict, prefix + ".norm1.bias"), 1e-5)
        f = F_.linear(F_.relu(F_.linear(x, _t(state_dict, prefix + ".linear1.weight"), _t(state_dict, prefix + ".linear1.bias"))),
                      _t(state_dict, prefix + ".linear2.weight"), _t(state_dict, prefix + ".linear2.bias"))
        return F_.layer_norm(x + f, (d,), _t(state_dict, prefix + ".norm2.weight"), _t(state_dict, prefix + ".norm2.bias"), 1e-5)

    with torch.no_grad():
        x = full.permute(1, 0, 2).contiguous()
        for i in range(2):
            x = restate.encoder_layer(state_dict, f"obj_inter_module.{i}", x, heads)
            for c, n in enumerate(counts):
                if n < slots - 1:
                    assert (x[n:, c] - x[n, c]).abs().max() < 1e-6, "padded slots of a cell must stay identical rows"
        want = F_.normalize(x.max(dim=0)[0])
        for c, n in enumerate(counts):
            r = full[c, :n + 1] if n < slots else full[c]
            mult = float(slots - n) if n < slots else 1.0
            for i in range(2):
                r = reduced_layer(f"obj_inter_module.{i}", r, mult)
            got = F_.normalize(r.max(dim=0)[0], dim=0)
            assert (got - want[c]).abs().max() < 2e-6, f"cell {c} ({n} objects): reduced rows differ from the padded tensor"
