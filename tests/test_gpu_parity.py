"""GPU parity of the path through the C ABI and the drop-in API against the golden vectors
(outputs of the reference's own Python) and the oracle.

Tolerances (north_star): top-k indices bit-exact; embeddings and similarities within 1e-3
relative.  For unit-norm embeddings "relative" is taken as ||got - want||_2 / ||want||_2 per row.
"""
import argparse

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

EMB_TOL = 1e-3


@pytest.fixture(scope="module")
def eng(state_dict):
    from text2loc_b200.engine import Engine

    e = Engine("cuda:0")
    e.load_state_dict(state_dict)
    return e


def row_rel_err(got, want):
    return float((np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)).max())


# ---- encoders ----------------------------------------------------------------------------------

def test_encode_cells_golden(eng, golden):
    g = golden("cells_small.npz")
    out = eng.encode_cells(g["pts"], g["meta"], g["cell_ptr"]).cpu().numpy()
    err = row_rel_err(out, g["cell_emb"])
    print(f"\ncell embedding max row-relative error vs reference: {err:.3e}")
    assert out.shape == g["cell_emb"].shape
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    assert err < EMB_TOL


def test_encode_cells_independent_of_batching(eng, golden):
    """Cells are independent units: encoding them one at a time or all at once is the same."""
    g = golden("cells_small.npz")
    cp = g["cell_ptr"]
    full = eng.encode_cells(g["pts"], g["meta"], cp).cpu().numpy()
    for c in (0, 4, 7):
        o0, o1 = cp[c], cp[c + 1]
        one = eng.encode_cells(g["pts"][o0:o1], g["meta"][o0:o1], np.array([0, o1 - o0], np.int32)).cpu().numpy()
        assert np.abs(one[0] - full[c]).max() < 1e-6


def test_encode_cells_rejects_bad_input(eng):
    from text2loc_b200.engine import EngineError

    with pytest.raises(EngineError):
        eng.encode_cells(np.zeros((2, 256, 6), np.float32), np.zeros((2, 7), np.float32), np.array([0, 0, 2], np.int32))  # empty cell
    with pytest.raises(EngineError):
        eng.encode_cells(np.zeros((2, 100, 6), np.float32), np.zeros((2, 7), np.float32), np.array([0, 2], np.int32))  # not 256 points


def test_encode_text_golden(eng, golden):
    from oracle import fake_t5

    g = golden("text_small.npz")
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    out = eng.encode_text(feat, n_sent).cpu().numpy()
    err = row_rel_err(out, g["text_emb"])
    print(f"\ntext embedding max row-relative error vs reference: {err:.3e}")
    assert err < EMB_TOL


def test_encode_text_tf32_path_also_within_tolerance(state_dict, golden, monkeypatch):
    """The token layer runs on fp16 operands by default; T2L_TEXT_TF32=1 keeps the tf32 variant for A/B checks.
    Both have an 11-bit significand and must meet the same tolerance against the reference's fp32 modules."""
    from oracle import fake_t5
    from text2loc_b200.engine import Engine

    monkeypatch.setenv("T2L_TEXT_TF32", "1")
    e32 = Engine("cuda:0")
    e32.load_state_dict(state_dict)
    g = golden("text_small.npz")
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    err = row_rel_err(e32.encode_text(feat, n_sent).cpu().numpy(), g["text_emb"])
    print(f"\ntext embedding error of the tf32 token layer: {err:.3e}")
    assert err < EMB_TOL


def test_encode_text_fp32_stream_variant(eng, state_dict, golden, monkeypatch):
    """Default: the token layer's residual stream (x + attn, LayerNorm1, x1 + ffn) is carried as fp16 rows; T2L_TEXT_STREAM32=1
    keeps it in fp32.  Both meet the tolerance against the reference's fp32 modules; the fp32 stream is the closer one."""
    from oracle import fake_t5
    from text2loc_b200.engine import Engine

    g = golden("text_small.npz")
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    err16 = row_rel_err(eng.encode_text(feat, n_sent).cpu().numpy(), g["text_emb"])
    monkeypatch.setenv("T2L_TEXT_STREAM32", "1")
    e32 = Engine("cuda:0")
    e32.load_state_dict(state_dict)
    err32 = row_rel_err(e32.encode_text(feat, n_sent).cpu().numpy(), g["text_emb"])
    print(f"\ntext embedding error: fp16 residual stream {err16:.3e}, fp32 residual stream {err32:.3e}")
    assert err16 < EMB_TOL and err32 < EMB_TOL
    assert err32 < err16 * 1.05


def test_encode_text_feature_magnitudes(eng, state_dict):
    """fp16 operands: features 4x larger / 100x smaller than the synthetic default stay in tolerance (fp16 has the
    range for anything a LayerNorm-ed T5 state can hold); absurd magnitudes saturate at 65504 instead of producing inf."""
    from oracle import restate
    import synth

    base = synth.make_t5_features(5, 8, 6, 12)
    for scale in (4.0, 0.01):
        t5 = base * np.float32(scale)
        got = eng.encode_text(t5, 6).cpu().numpy()
        want = restate.encode_text(state_dict, t5, 6).numpy()
        err = row_rel_err(got, want)
        print(f"\nfeature scale {scale}: text embedding error {err:.3e}")
        assert err < EMB_TOL
    got = eng.encode_text(base * np.float32(1e6), 6).cpu().numpy()
    assert np.isfinite(got).all()


def test_encode_text_shapes(eng, state_dict):
    """Different token counts / sentence counts, chunk boundary in the middle of the batch."""
    from oracle import restate
    import synth

    for nq, S, L in ((3, 6, 9), (5, 4, 17), (2, 1, 1), (40, 6, 12)):
        t5 = synth.make_t5_features(nq + S + L, nq, S, L)
        got = eng.encode_text(t5, S).cpu().numpy()
        want = restate.encode_text(state_dict, t5, S).numpy()
        assert row_rel_err(got, want) < EMB_TOL, (nq, S, L)


def test_encode_text_host_streaming_equals_device(eng):
    """Host (pinned) input is uploaded in chunks on a side stream; same result as a resident tensor."""
    import synth

    t5 = torch.from_numpy(synth.make_t5_features(77, 1000, 6, 12))  # > 2 engine chunks of 455 queries
    a = eng.encode_text(t5.pin_memory(), 6)
    b = eng.encode_text(t5.cuda(), 6)
    c = eng.encode_text(t5, 6)  # pageable host memory also works (copies are then synchronous)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(c, b)


def test_encode_text_fp16_features(eng, state_dict, golden):
    """T5 states delivered as float16 (t2l_encode_text_tokens_f16), from device memory and streamed from pinned host memory:
    same embeddings as the fp32 entry within the tolerance against the REFERENCE (fp32 features), both ways bit-equal."""
    from oracle import fake_t5

    g = golden("text_small.npz")
    feat, n_sent = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    got = eng.encode_text(feat.half().cuda(), n_sent).cpu().numpy()
    err = row_rel_err(got, g["text_emb"])
    print(f"\ntext embedding error with fp16 T5 features: {err:.3e}")
    assert err < EMB_TOL
    import synth

    t5 = torch.from_numpy(synth.make_t5_features(78, 1000, 6, 12)).half()
    a = eng.encode_text(t5.cuda(), 6)
    b = eng.encode_text(t5.pin_memory(), 6)
    assert torch.equal(a, b)
    ref32 = eng.encode_text(t5.float().cuda(), 6)  # the same (fp16-representable) values through the fp32 entry
    assert float((a - ref32).norm(dim=1).max()) < 2e-3  # only the residual operand's width differs... and it is exact here
    assert torch.equal(a, ref32) or float((a - ref32).abs().max()) < 1e-4


# ---- search ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,nq,k", [(3000, 64, 10), (20000, 1000, 10), (257, 130, 5), (100000, 512, 10), (1000, 1, 1), (5000, 300, 12)])
def test_search_matches_fp64_oracle(eng, n, nq, k):
    from oracle import restate
    import synth

    D = synth.make_unit_rows(n, n)
    Q = synth.make_unit_rows(nq + 1, nq)
    eng.db_build(D)
    idx, sc, nfb = eng.search_topk(Q, k)
    oidx, osc = restate.search_topk(D, Q, k)
    assert (idx.cpu().numpy() == oidx).all()
    assert np.abs(sc.cpu().numpy() - osc).max() < 1e-12  # fp64 dots, different summation order only
    print(f"\nN={n} nq={nq} k={k}: exact-rescan fallbacks {int(nfb)} / {nq}")
    assert int(nfb) <= max(1, nq // 20)
    eidx, esc, _ = eng.search_topk(Q, k, exact=True)
    assert (eidx.cpu().numpy() == oidx).all()


def test_search_golden_reference_loop(eng, golden):
    import synth

    g = golden("search_small.npz")
    eng.db_build(synth.make_unit_rows(int(g["d_seed"]), int(g["n"])))
    idx, sc, _ = eng.search_topk(synth.make_unit_rows(int(g["q_seed"]), int(g["nq"])), 10)
    assert (idx.cpu().numpy() == g["idx"]).all()  # the reference's own np.argsort loop
    assert np.abs(sc.cpu().numpy() - g["score"]).max() < 1e-12


def test_search_ties_duplicates_and_clusters(eng):
    """Duplicate rows tie exactly -> index order; a tight cluster forces the margin proof to fail
    and the exact rescan to take over.  Either way the result is the oracle's."""
    from oracle import restate
    import synth

    D = synth.make_unit_rows(5, 4000)
    D[100] = D[7]
    D[3999] = D[7]
    D[2000:2040] = D[1999] + 1e-6 * np.random.default_rng(0).standard_normal((40, 256)).astype(np.float32)  # 41 near-identical rows
    Q = np.concatenate([D[7:8], D[1999:2000], synth.make_unit_rows(6, 30)])
    eng.db_build(D)
    idx, sc, nfb = eng.search_topk(Q, 10)
    oidx, osc = restate.search_topk(D, Q, 10)
    assert (idx.cpu().numpy() == oidx).all()
    assert idx[0, :3].tolist() == [7, 100, 3999]
    assert int(nfb) >= 1  # the cluster query cannot be proven from 16 candidates per split


def test_search_second_pass_and_exhaustive_rescan(eng):
    """Near-collinear rows (what random-weight encoders produce): once more than 16 rows of a split sit within the
    error bound of the k-th score the margin proof fails and the second tensor-core pass collects every row that
    can still matter; 300 exact duplicates (and the tightest queries) overflow its 256-entry buffer and force the
    exhaustive fp64 rescan.  All three routes must return the oracle's answer."""
    from oracle import restate
    import synth

    rng = np.random.default_rng(3)
    base = synth.make_unit_rows(40, 1)[0]
    D = (base[None, :] + 0.004 * rng.standard_normal((6000, 256))).astype(np.float32)  # cosine ~0.998 between any two rows
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    D[1000:1300] = D[999]  # 301 identical rows
    Q = np.concatenate([D[999:1000], (base[None, :] + 0.004 * rng.standard_normal((200, 256))).astype(np.float32)])
    eng.db_build(D)
    idx, sc, nfb = eng.search_topk(Q, 10)
    oidx, osc = restate.search_topk(D, Q, 10)
    assert (idx.cpu().numpy() == oidx).all()
    assert idx[0].tolist() == list(range(999, 1009))  # ties in index order, via the exhaustive rescan
    assert np.abs(sc.cpu().numpy() - osc).max() < 1e-12
    print(f"\nnear-collinear database: {int(nfb)} / {len(Q)} queries took the second pass")
    assert int(nfb) > len(Q) // 8


def test_search_small_db_and_row_offset(eng):
    from oracle import restate
    import synth

    D = synth.make_unit_rows(9, 6)
    Q = synth.make_unit_rows(10, 3)
    eng.db_build(D, row_offset=1000)
    idx, sc, _ = eng.search_topk(Q, 10)
    oidx, osc = restate.search_topk(D, Q, 10)
    got = idx.cpu().numpy()
    assert (got[:, :6] == oidx + 1000).all() and (got[:, 6:] == -1).all()
    assert np.isneginf(sc.cpu().numpy()[:, 6:]).all()


def test_merge_topk_is_shard_count_independent(eng):
    """Row-shard the DB 1/2/4/8 ways on one GPU, merge the per-shard lists: identical result."""
    from oracle import restate
    import synth

    D = synth.make_unit_rows(11, 10000)
    D[9000] = D[10]  # a tie across shards
    Q = np.concatenate([D[10:11], synth.make_unit_rows(12, 99)])
    oidx, osc = restate.search_topk(D, Q, 10)
    for G in (1, 2, 4, 8):
        bounds = np.linspace(0, len(D), G + 1).astype(int)
        idxs, scs = [], []
        for gi in range(G):
            eng.db_build(D[bounds[gi]:bounds[gi + 1]], row_offset=int(bounds[gi]))
            i, s, _ = eng.search_topk(Q, 10)
            idxs.append(i)
            scs.append(s)
        idx, sc = eng.merge_topk(torch.stack(idxs), torch.stack(scs))
        assert (idx.cpu().numpy() == oidx).all(), G
        assert np.abs(sc.cpu().numpy() - osc).max() < 1e-12


# ---- BASELINE.json's full sizes: properties that do not need the oracle to scale ------------------------

def test_search_full_size_properties(eng):
    """configs[2]: 100 000 rows x 32 768 queries.  The fp64 oracle only checks a sample; the rest is covered by
    properties: scores are the exact fp64 dots of the returned rows, lists are sorted by (score desc, row asc),
    the call is idempotent, and 8 row shards merged equal the unsharded search bit for bit."""
    from oracle import restate
    import synth

    N, NQ, K = 100000, 32768, 10
    D = synth.make_unit_rows(101, N)
    Q = synth.make_unit_rows(102, NQ)
    Dt, Qt = torch.from_numpy(D).cuda(), torch.from_numpy(Q).cuda()
    eng.db_build(Dt)
    idx, sc, nfb = eng.search_topk(Qt, K)
    idx2, sc2, _ = eng.search_topk(Qt, K)
    assert torch.equal(idx, idx2) and torch.equal(sc, sc2)  # idempotent
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    # sortedness: score non-increasing, row index increasing inside a tie
    d = sc[:, 1:] - sc[:, :-1]
    assert bool((d <= 0).all())
    assert bool(((d < 0) | (idx[:, 1:] > idx[:, :-1])).all())
    # returned scores are the fp64 dots of the returned rows (any summation order: 1e-12)
    g = Dt[idx.reshape(-1)].double().reshape(NQ, K, 256)
    exact = torch.einsum("qkd,qd->qk", g, Qt.double())
    assert float((exact - sc).abs().max()) < 1e-12
    # oracle on a sample of queries
    pick = np.random.default_rng(0).choice(NQ, 48, replace=False)
    oidx, _ = restate.search_topk(D, Q[pick], K)
    assert (idx[torch.from_numpy(pick).cuda()].cpu().numpy() == oidx).all()
    # 8 shards (BASELINE's 8 x 12 500 rows) merged == unsharded
    bounds = np.linspace(0, N, 9).astype(int)
    idxs, scs = [], []
    for gi in range(8):
        eng.db_build(Dt[bounds[gi]:bounds[gi + 1]], row_offset=int(bounds[gi]))
        i, s, _ = eng.search_topk(Qt, K)
        idxs.append(i)
        scs.append(s)
    midx, msc = eng.merge_topk(torch.stack(idxs), torch.stack(scs))
    assert torch.equal(midx, idx) and torch.equal(msc, sc)
    print(f"\n100000 x 32768: {int(nfb)} queries took the second pass; sorted, idempotent, exact, shard-independent")


def test_encode_cells_full_size_properties(eng, state_dict):
    """3 000 cells x 8 objects (more than one 16 384-object chunk): cells are independent units, so encoding them in
    reverse order must give the same rows bit for bit (chunk boundaries fall elsewhere), rows are unit vectors, and a
    sample agrees with the oracle."""
    from oracle import restate
    import synth

    n_cells = 3000
    pts, meta, ptr = synth.make_packed_cells(77, n_cells, 8)
    fwd = eng.encode_cells(pts, meta, ptr)
    order = np.arange(n_cells)[::-1].copy()
    obj = (order[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    rev = eng.encode_cells(pts[obj], meta[obj], ptr)
    assert torch.equal(rev.flip(0), fwd)
    assert float((fwd.norm(dim=1) - 1).abs().max()) < 1e-5
    sample = [0, 1499, 2047, 2048, 2999]  # includes the cells either side of the chunk boundary
    sobj = np.concatenate([np.arange(c * 8, c * 8 + 8) for c in sample])
    want = restate.encode_cells(state_dict, pts[sobj], meta[sobj], np.arange(0, 8 * len(sample) + 1, 8, dtype=np.int32)).numpy()
    assert row_rel_err(fwd.cpu().numpy()[sample], want) < EMB_TOL


def test_encode_cells_16_objects_per_cell(eng, state_dict):
    """BASELINE configs[3] cell shape: 16 objects per cell (packed), 300 cells = 4 800 objects; every 20th cell against
    the oracle, plus the order-independence property on all of them."""
    from oracle import restate
    import synth

    pts, meta, ptr = synth.make_packed_cells(41, 300, 16)
    got = eng.encode_cells(pts, meta, ptr).cpu().numpy()
    assert np.abs(np.linalg.norm(got, axis=1) - 1).max() < 1e-5
    sample = list(range(0, 300, 20))
    sobj = np.concatenate([np.arange(c * 16, c * 16 + 16) for c in sample])
    want = restate.encode_cells(state_dict, pts[sobj], meta[sobj], np.arange(0, 16 * len(sample) + 1, 16, dtype=np.int32)).numpy()
    err = row_rel_err(got[sample], want)
    print(f"\n16 objects/cell: embeddings vs oracle {err:.3e}")
    assert err < EMB_TOL
    perm = np.arange(299, -1, -1)
    pobj = np.concatenate([np.arange(c * 16, c * 16 + 16) for c in perm])
    rev = eng.encode_cells(pts[pobj], meta[pobj], ptr).cpu().numpy()
    assert (rev[::-1] == got).all()


# ---- streamed database (BASELINE configs[3]) ------------------------------------------------------------------

def test_synthetic_cell_generator_matches_numpy_restatement(eng):
    """t2l_synth_cells is a pure function of (seed, global object index): bit-equal to oracle/synthgen.py, and any
    chunking of the cell range yields the same bytes."""
    from oracle import synthgen

    pts, meta, ptr = eng.synth_cells(9, 100, 7, 16)
    wp, wm, wptr = synthgen.synth_cells(9, 100, 7, 16)
    assert (ptr == wptr).all()
    assert (pts.cpu().numpy() == wp).all() and (meta.cpu().numpy() == wm).all()
    a, am, _ = eng.synth_cells(9, 100, 3, 16)
    b, bm, _ = eng.synth_cells(9, 103, 4, 16)
    assert torch.equal(torch.cat([a, b]), pts) and torch.equal(torch.cat([am, bm]), meta)
    assert int((meta[:, 6] < 256).sum()) > 0  # objects that repeat points are present
    small = int(torch.nonzero(meta[:, 6] < 200)[0])
    assert len(torch.unique(pts[small, :, 0])) < 200


def test_streamed_search_equals_unstreamed_50k_cells(eng, state_dict):
    """configs[3] shape at 1/20 scale: 50 000 cells x 16 objects generated on the device chunk by chunk, encoded and
    scored against 1 024 queries with a running top-k; neither the points (4.9 GB) nor the planes of the whole database
    exist at once.  The result must equal the unstreamed search over the kept embeddings bit for bit, must not depend
    on the chunking, must equal the fp64 oracle on a sample, and sampled cells must match the oracle encoder."""
    from oracle import restate, synthgen
    import synth
    from text2loc_b200 import streaming

    n_cells, per, k = 50000, 16, 10
    Q = eng.encode_text(torch.from_numpy(synth.make_t5_features(31, 1024)).cuda(), 6)
    idx, score, nfb, D = streaming.stream_synthetic(eng, Q, k, seed=5, first_cell=0, n_cells=n_cells, obj_per_cell=per,
                                                    chunk_cells=1024, keep_embeddings=True)
    assert D.shape == (n_cells, 256)
    eng.db_build(D)
    idx2, score2, _ = eng.search_topk(Q, k)
    assert torch.equal(idx, idx2) and torch.equal(score, score2)
    idx3, score3, _ = streaming.stream_synthetic(eng, Q, k, seed=5, first_cell=0, n_cells=n_cells, obj_per_cell=per, chunk_cells=777)
    assert torch.equal(idx, idx3) and torch.equal(score, score3)
    # two "ranks" streaming half of the cells each, merged: the 8-GPU layout of configs[3] on one device
    ia, sa, _ = streaming.stream_synthetic(eng, Q, k, 5, 0, n_cells // 2, per, chunk_cells=1024)
    ib, sb, _ = streaming.stream_synthetic(eng, Q, k, 5, n_cells // 2, n_cells - n_cells // 2, per, chunk_cells=1024)
    im, sm = eng.merge_topk(torch.stack([ia, ib]), torch.stack([sa, sb]))
    assert torch.equal(im, idx) and torch.equal(sm, score)
    # fp64 oracle on a sample of queries (same embeddings into both)
    oidx, oscore = restate.search_topk(D.cpu().numpy(), Q[:48].cpu().numpy(), k)
    assert (idx[:48].cpu().numpy() == oidx).all() and np.abs(score[:48].cpu().numpy() - oscore).max() < 1e-12
    # encoder parity on generated cells
    sample = [0, 1023, 1024, 25000, 49999]
    pts = np.concatenate([synthgen.synth_cells(5, c, 1, per)[0] for c in sample])
    meta = np.concatenate([synthgen.synth_cells(5, c, 1, per)[1] for c in sample])
    want = restate.encode_cells(state_dict, pts, meta, np.arange(0, per * len(sample) + 1, per, dtype=np.int32)).numpy()
    err = row_rel_err(D[sample].cpu().numpy(), want)
    print(f"\nstreamed 50k cells x 16 objects: top-k == unstreamed == merged halves; sampled cells vs oracle {err:.3e}; second-pass queries {int(nfb)}")
    assert err < EMB_TOL


# ---- fine stage (BASELINE configs[4]) -------------------------------------------------

FINE_TOL = 1e-3  # relative to the largest offset component, as for embeddings / similarities


@pytest.fixture(scope="module")
def fine_sd():
    import synth

    return synth.make_fine_state_dict(0)


@pytest.fixture(scope="module")
def fine_eng(fine_sd):
    from text2loc_b200.engine import Engine

    e = Engine("cuda:0")
    e.load_state_dict(fine_sd)
    return e


def test_fine_offsets_golden(fine_eng, golden):
    """t2l_fine_offsets against the reference's own CrossMatch.forward (models/cross_matcher.py:83-129) on 5 cells padded to
    16 objects x 6 hints (tests/golden/fine_small.npz)."""
    from oracle import fake_t5

    g = golden("fine_small.npz")
    feat, n_hints = fake_t5.FakeFrontend(int(g["fake_t5_seed"]))([str(t) for t in g["texts"]])
    got = fine_eng.fine_offsets(g["pts"], g["meta"], g["cell_ptr"], feat, n_hints).cpu().numpy()
    err = np.abs(got - g["offsets"]).max() / np.abs(g["offsets"]).max()
    print(f"\nfine offsets vs the reference's CrossMatch.forward: max error {err:.3e} of the largest component")
    assert got.shape == (5, 2) and err < FINE_TOL
    # the three stages on their own compose to the same result (what the batched run_fine uses)
    obj = fine_eng.fine_encode_objects(g["pts"], g["meta"], g["cell_ptr"])
    assert float((obj.norm(dim=1) - 1).abs().max()) < 1e-5
    hints = fine_eng.fine_encode_hints(feat)
    staged = fine_eng.fine_match(obj, None, hints, None, 16, n_hints).cpu().numpy()
    assert np.abs(staged - got).max() < 1e-6
    # pairs that re-use cells and queries: pair (cell 3, query 1) differs from both diagonal pairs and equals the oracle
    from oracle import restate

    mixed = fine_eng.fine_match(obj, np.array([3, 0], np.int32), hints, np.array([1, 4], np.int32), 16, n_hints).cpu().numpy()
    import synth

    sd = synth.make_fine_state_dict(int(g["weight_seed"]))
    t5 = feat.view(5, n_hints, feat.shape[1], 1024)
    pts, meta = torch.from_numpy(g["pts"]).view(5, 16, 256, 6), torch.from_numpy(g["meta"]).view(5, 16, 7)
    want = restate.fine_offsets(sd, torch.cat([pts[3], pts[0]]), torch.cat([meta[3], meta[0]]), np.array([0, 16, 32]),
                                torch.cat([t5[1], t5[4]]), n_hints).numpy()
    assert np.abs(mixed - want).max() / np.abs(want).max() < FINE_TOL


def test_fine_dropin_crossmatch_and_engine_guards(fine_sd, state_dict, golden):
    from oracle import fake_t5, reference_run
    from oracle.make_golden import fine_case
    from text2loc_b200 import CrossMatch
    from text2loc_b200.engine import Engine, EngineError

    g = golden("fine_small.npz")
    args = reference_run.fine_args()
    model = CrossMatch(["c"] * 22, ["k"] * 8, args, text_frontend=fake_t5.FakeFrontend(int(g["fake_t5_seed"])))
    model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in fine_sd.items()}, strict=False)
    assert model.embed_dim == 128 and model.eval() is model and model.get_device() == model.device
    cells, batches, texts = fine_case()  # the raw objects / point batches / descriptions the golden was made from
    out = model(cells, texts, batches)
    assert out.shape == (5, 2) and out.is_cuda
    assert np.abs(out.cpu().numpy() - g["offsets"]).max() / np.abs(g["offsets"]).max() < FINE_TOL
    # a fine engine refuses the coarse entry points and vice versa; wrong-sized checkpoints are rejected
    with pytest.raises(EngineError, match="fine-stage"):
        model.engine.encode_cells(g["pts"], g["meta"], g["cell_ptr"])
    coarse = Engine("cuda:0")
    coarse.load_state_dict(state_dict)
    with pytest.raises(EngineError, match="coarse model"):
        coarse.fine_encode_hints(torch.zeros(6, 12, 1024))
    bad = dict(fine_sd)
    bad["mlp_offsets.0.weight"] = np.asarray(bad["mlp_offsets.0.weight"])[:, :64]
    with pytest.raises(EngineError, match="size mismatch"):
        Engine("cuda:0").load_state_dict(bad)


def test_run_fine_batched_equals_per_pair_oracle(fine_sd):
    """The drop-in run_fine encodes every distinct retrieved cell once and matches all query x cell pairs in one batch.
    Replaying its (seeded) padding / point sampling in the test, every pair's offsets must equal the oracle's
    CrossMatch restatement on that pair, and the accuracies must equal the reference's calc_sample_accuracies loop."""
    from oracle import fake_t5, reference_run, restate
    import synth
    from text2loc_b200 import CrossMatch, dataio, evaluation

    args = reference_run.fine_args(top_k=[1, 3], threshs=[5, 10, 15])
    ds = synth.SynthCoarseDataset(seed=4, n_cells=12, n_poses=7, n_obj=[3, 16, 20, 1, 8, 5, 16, 2, 9, 30, 4, 6], max_raw=300)
    loader = DataLoader(ds, batch_size=4, collate_fn=dataio.collate_fn, shuffle=False)
    rng = np.random.default_rng(2)
    ids = np.array([c.id for c in ds.all_cells])
    retrievals = np.stack([ids[rng.permutation(12)[:3]] for _ in range(7)])
    frontend = fake_t5.FakeFrontend(0)
    model = CrossMatch([], [], args, text_frontend=frontend)
    model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in fine_sd.items()}, strict=False)
    transform = dataio.FixedPoints(256)
    np.random.seed(77)
    acc, offsets = evaluation.run_fine(model, retrievals, loader, args, transform, return_offsets=True)
    assert offsets.shape == (7, 3, 2)
    # replay: distinct retrieved cells in ascending row order, padded then sampled, exactly as run_fine does
    rows = evaluation.rows_of_ids(ids, retrievals)
    used = np.unique(rows)
    np.random.seed(77)
    objects = [evaluation._padded_objects(ds.all_cells[int(r)], args.pad_size) for r in used]
    points = [dataio.batch_object_points(o, transform) for o in objects]
    pts, meta, _ = dataio.pack_cells(objects, points)
    pts, meta = pts.view(len(used), 16, 256, 6), meta.view(len(used), 16, 7)
    slot = {int(r): i for i, r in enumerate(used)}
    feats, n_hints = frontend([p.text for p in ds.all_poses])
    feats = feats.view(7, n_hints, feats.shape[1], 1024)
    worst = 0.0
    for q in range(7):
        sel = [slot[int(r)] for r in rows[q]]
        want = restate.fine_offsets(fine_sd, pts[sel].reshape(-1, 256, 6), meta[sel].reshape(-1, 7), np.arange(0, 16 * 3 + 1, 16),
                                    feats[q].repeat(3, 1, 1), n_hints).numpy()
        worst = max(worst, float(np.abs(offsets[q] - want).max() / np.abs(want).max()))
    print(f"\nrun_fine: 21 query x cell pairs over {len(used)} distinct cells, worst offset error vs oracle {worst:.3e}")
    assert worst < FINE_TOL
    want_acc = restate.localisation_accuracies(ds.all_poses, ds.all_cells, retrievals, offsets, args.top_k, args.threshs)
    assert all(acc[k][t] == want_acc[k][t] for k in args.top_k for t in args.threshs)


# ---- drop-in API ---------------------------------------------------------------------------------------

def make_model(state_dict, fake_seed=0):
    from oracle import fake_t5, reference_run
    from text2loc_b200 import CellRetrievalNetwork

    args = reference_run.default_args()
    model = CellRetrievalNetwork(["c"] * 22, ["k"] * 8, args, text_frontend=fake_t5.FakeFrontend(fake_seed))
    model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}, strict=False)
    return model.eval(), args


def test_sentence_row_cache_equals_uncached_encode_text(state_dict, fine_sd):
    """With a sentence-caching front end the drop-in models run the token stage once per distinct (sentence, n_tok) and
    assemble batches from cached rows (text_frontend.SentenceRowCache).  Same embeddings as
    pushing every batch through the text head whole -- rows of the token stage do not depend on their neighbours."""
    from oracle import fake_t5, reference_run
    from oracle.stubs import sent_tokenize
    from text2loc_b200 import CellRetrievalNetwork, CrossMatch
    from text2loc_b200.text_frontend import SentenceCacheFrontend

    def frontend():
        return SentenceCacheFrontend(fake_t5.FakeTokenizer(), fake_t5.FakeT5Encoder(0).eval().cuda(), "cuda:0", cap=16, split=sent_tokenize)

    dirs, cols, labs = ["north", "east", "on-top", "west"], ["gray", "red", "dark-green"], ["building", "pole", "traffic light", "vending machine"]
    rng = np.random.default_rng(5)
    texts = [" ".join(f"The pose is {dirs[rng.integers(4)]} of a {cols[rng.integers(3)]} {labs[rng.integers(4)]}." for _ in range(6))
             for _ in range(300)]
    model, _ = make_model(state_dict)
    model._frontend = frontend()
    plain, _ = make_model(state_dict)
    plain._frontend = frontend()
    plain.cache_sentence_rows = False
    worst = 0.0
    for batch in (texts[:64], texts[64:70], texts[:300]):
        got, want = model.encode_text(batch), plain.encode_text(batch)
        assert got.shape == want.shape == (len(batch), 256)
        worst = max(worst, float((got - want).abs().max()))
    n_distinct = len({s for t in texts for s in sent_tokenize(t)})
    print(f"\nsentence row cache: {model._sentence_rows.computed} token-stage rows for {300 * 6 + 70 * 6} sentences "
          f"({n_distinct} distinct), max |cached - uncached| = {worst:.1e}")
    assert worst <= 1e-5  # bit-equal in practice; 1 % of the 1e-3 tolerance is the bound asserted
    assert model._sentence_rows.computed <= 3 * n_distinct  # at most one row per distinct sentence and padding length

    args = reference_run.fine_args(top_k=[1, 3], threshs=[5, 10, 15])
    fine = CrossMatch([], [], args, text_frontend=frontend())
    fine.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in fine_sd.items()}, strict=False)
    got, nh = fine.encode_hints(texts[:40])
    fine.cache_sentence_rows = False
    want, nh2 = fine.encode_hints(texts[:40])
    assert nh == nh2 == 6 and got.shape == want.shape == (240, 128)
    assert float((got - want).abs().max()) <= 1e-5


def test_dropin_eval_epoch_and_run_coarse_golden(state_dict, golden):
    """The reference's own eval_epoch / run_coarse outputs on the same seeded dataset."""
    from oracle.make_golden import e2e_dataset
    from text2loc_b200 import dataio, eval_epoch, run_coarse

    g = golden("eval_e2e.npz")
    model, args = make_model(state_dict, int(g["fake_t5_seed"]))
    args.batch_size = int(g["batch_size"])
    ds = e2e_dataset()
    loader = DataLoader(ds, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    np.random.seed(int(g["np_seed"]))
    acc, acc_close, retr, cell_enc, text_enc = eval_epoch(model, loader, args, return_encodings=True)
    assert cell_enc.dtype == np.float64 and text_enc.dtype == np.float64
    ec, et = row_rel_err(cell_enc, g["cell_enc"]), row_rel_err(text_enc, g["text_enc"])
    print(f"\ne2e embeddings vs reference: cells {ec:.3e}, text {et:.3e}")
    assert ec < EMB_TOL and et < EMB_TOL
    # L1 parity: the engine's top-k on ITS embeddings equals the fp64 oracle on those same embeddings
    from oracle import restate

    oidx, _ = restate.search_topk(cell_enc, text_enc, 10)
    ids = np.array([c.id for c in ds.all_cells])
    got = np.stack([retr[i] for i in range(len(ds))])
    assert got.dtype.kind == "U" and got.shape == (len(ds), 10)
    assert (got == ids[oidx]).all()
    # L3 parity: vs the reference's retrievals; a query may differ only where the reference's own
    # k/k+1 score gap is inside twice the embedding error
    ref = g["retrievals"]
    s_ref = g["text_enc"] @ g["cell_enc"].T
    for q in np.nonzero((got != ref).any(axis=1))[0]:
        srt = np.sort(s_ref[q])[::-1]
        assert np.min(np.abs(np.diff(srt[:11]))) < 2 * (ec + et), f"query {q} differs with a clear score gap"
    np.random.seed(int(g["np_seed"]))
    retrievals, accuracies = run_coarse(model, loader, args, verbose=False)
    assert len(retrievals) == len(ds) and all(len(r) == 10 for r in retrievals)
    assert set(accuracies.keys()) == set(args.top_k) and set(accuracies[1].keys()) == set(args.threshs)
    assert np.allclose([[accuracies[k][t] for t in args.threshs] for k in args.top_k], g["run_coarse_acc"], atol=0.1)
    # bookkeeping: the device rows (t2l_topk_accuracy) equal the reference's per-query loops on the SAME retrievals, exactly
    check_bookkeeping(ds, args, got, acc, acc_close, accuracies)
    if (got == ref).all():  # identical retrievals => identical accuracies, bit for bit
        assert [acc[k] for k in args.top_k] == list(g["acc"]) and [acc_close[k] for k in args.top_k] == list(g["acc_close"])
        assert ([[accuracies[k][t] for t in args.threshs] for k in args.top_k] == g["run_coarse_acc"]).all()


def check_bookkeeping(ds, args, got, acc, acc_close, accuracies):
    from oracle import restate

    cells_dict = {c.id: c for c in ds.all_cells}
    want_acc, want_close, _ = restate.retrieval_accuracies(
        got, np.array([p.cell_id for p in ds.all_poses]), np.array([p.pose_w[0:2] for p in ds.all_poses]), cells_dict,
        ds.all_cells[0].cell_size, args.top_k)
    assert all(acc[k] == want_acc[k] for k in args.top_k), (acc, want_acc)
    assert all(acc_close[k] == want_close[k] for k in args.top_k), (acc_close, want_close)
    want_loc = restate.localisation_accuracies(ds.all_poses, ds.all_cells, got, np.full((len(got), got.shape[1], 2), 0.5), args.top_k, args.threshs)
    assert all(accuracies[k][t] == want_loc[k][t] for k in args.top_k for t in args.threshs), (accuracies, want_loc)


def test_dropin_configs0_against_reference_run_coarse(state_dict, golden):
    """BASELINE configs[0] at its stated size: 1 000 cells x 8 objects x 256 points, 256 queries, seed 1, through the
    drop-in run_coarse / eval_epoch, against what the REFERENCE'S OWN evaluation.coarse.run_coarse returned for the same
    dataset (tests/golden/eval_cfg1.npz, oracle/make_golden.py cfg1).  Reports how many queries' top-10 differ; each of those
    must sit on a reference k/k+1 score gap smaller than twice the measured embedding error (L3 parity)."""
    from oracle import restate
    from oracle.make_golden import cfg1_dataset
    from text2loc_b200 import dataio, eval_epoch, run_coarse

    g = golden("eval_cfg1.npz")
    model, args = make_model(state_dict, int(g["fake_t5_seed"]))
    args.batch_size = int(g["batch_size"])
    ds = cfg1_dataset()
    loader = DataLoader(ds, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    np.random.seed(int(g["np_seed"]))
    retrievals, accuracies = run_coarse(model, loader, args, verbose=False)
    np.random.seed(int(g["np_seed"]))
    acc, acc_close, retr, cell_enc, text_enc = eval_epoch(model, loader, args, return_encodings=True)
    got = np.stack([retr[i] for i in range(len(ds))])
    assert (np.stack(retrievals) == got).all()
    ref_cells, ref_text = g["cell_enc"].astype(np.float64), g["text_enc"].astype(np.float64)
    ec, et = row_rel_err(cell_enc, ref_cells), row_rel_err(text_enc, ref_text)
    assert ec < EMB_TOL and et < EMB_TOL
    # L1: the engine's lists are the fp64 stable order of ITS embeddings
    oidx, _ = restate.search_topk(cell_enc, text_enc, 10)
    ids = np.array([c.id for c in ds.all_cells])
    assert (got == ids[oidx]).all()
    # L3: against the reference's own retrievals
    ref = g["retrievals"]
    differing = np.nonzero((got != ref).any(axis=1))[0]
    s_ref = ref_text @ ref_cells.T
    worst_gap = 0.0
    for q in differing:
        srt = np.sort(s_ref[q])[::-1]
        gap = float(np.min(np.abs(np.diff(srt[:11]))))
        worst_gap = max(worst_gap, gap)
        assert gap < 2 * (ec + et), f"query {q} differs from the reference although its closest top-11 score gap is {gap:.2e}"
    print(f"\nconfigs[0] vs the reference's run_coarse: embeddings cells {ec:.3e} / text {et:.3e}; "
          f"{len(differing)} of {len(ds)} queries differ in their top-10 (largest deciding gap {worst_gap:.2e}, bound {2 * (ec + et):.2e}); "
          f"reference run took {float(g['reference_run_coarse_seconds']):.0f} s on {int(g['reference_cores'])} cores")
    check_bookkeeping(ds, args, got, acc, acc_close, accuracies)
    if len(differing) == 0:
        assert [acc[k] for k in args.top_k] == list(g["acc"]) and [acc_close[k] for k in args.top_k] == list(g["acc_close"])
        assert ([[accuracies[k][t] for t in args.threshs] for k in args.top_k] == g["run_coarse_acc"]).all()


def test_topk_accuracy_kernel_matches_reference_loops(eng):
    """t2l_topk_accuracy on random retrievals (empty slots, several scenes, targets absent from the database) against
    the reference's per-query loops (training/coarse.py:131-150, evaluation/utils.py:31-54)."""
    from oracle import restate
    import synth

    rng = np.random.default_rng(17)
    n_c, n_q, k = 300, 500, 10
    cells = []
    for i in range(n_c):
        scene = f"{rng.integers(3):04d}"
        x0, y0 = rng.uniform(0, 300, 2)
        cells.append(synth.SynthCell(i, scene, [], 30.0, np.array([x0, y0, 0.0, x0 + 30, y0 + 30, 30.0])))
    poses = []
    for _ in range(n_q):
        c = cells[int(rng.integers(n_c))]
        poses.append(synth.SynthPose(c.bbox_w[0:3] + rng.uniform(-20, 50, 3), c.id, c.scene_name, ""))
    idx = np.stack([rng.permutation(n_c)[:k] for _ in range(n_q)]).astype(np.int64)
    ids = np.array([c.id for c in cells])
    top_k, threshs = [1, 3, 5, 10], [5, 10, 15]
    from text2loc_b200 import evaluation

    want_acc, want_close, want_d = restate.retrieval_accuracies(ids[idx], np.array([p.cell_id for p in poses]),
                                                                np.array([p.pose_w[0:2] for p in poses]), {c.id: c for c in cells}, 30.0, top_k)
    hit, close, dists = eng.topk_accuracy(idx, np.array([p.pose_w[0:2] for p in poses]), np.array([c.get_center()[0:2] for c in cells]),
                                          top_k, threshs=[15.0], target_row=evaluation.rows_of_ids(ids, np.array([p.cell_id for p in poses])),
                                          want_dists=True)
    assert (dists.cpu().numpy() == want_d).all()  # same float64 arithmetic as numpy's 2-vector norm
    assert all(hit[:, i].sum().item() / n_q == want_acc[kk] for i, kk in enumerate(top_k))
    assert all(close[:, i, 0].sum().item() / n_q == want_close[kk] for i, kk in enumerate(top_k))
    pos_in = rng.uniform(0, 1, (n_q, k, 2))
    want_loc = restate.localisation_accuracies(poses, cells, ids[idx], pos_in, top_k, threshs)
    got_loc = evaluation.localisation_accuracies(eng, poses, cells, ids[idx], pos_in, top_k, threshs)
    assert all(got_loc[kk][t] == want_loc[kk][t] for kk in top_k for t in threshs)
    assert 0 < want_loc[10][15] < 1  # the case is not degenerate
    # empty slots (-1) are +inf / never hits
    idx2 = idx.copy()
    idx2[:, 5:] = -1
    hit2, close2, d2 = eng.topk_accuracy(idx2, np.array([p.pose_w[0:2] for p in poses]), np.array([c.get_center()[0:2] for c in cells]),
                                         top_k, threshs=[15.0], target_row=np.full(n_q, -1), want_dists=True)
    assert torch.isinf(d2[:, 5:]).all() and (d2[:, :5].cpu().numpy() == want_d[:, :5]).all() and int(hit2.sum()) == 0
    assert (close2[:, 3, 0] == close2[:, 2, 0]).all()


def test_eval_epoch_with_packed_database_cache(state_dict, monkeypatch):
    """model.cache_packed_cells: the database is packed once (vectorised), encoded in one call and
    re-used by the next evaluation; retrievals are then reproducible across calls and remain the fp64 oracle's on the
    engine's own embeddings."""
    from oracle import restate
    from oracle.make_golden import e2e_dataset
    from text2loc_b200 import dataio, eval_epoch

    model, args = make_model(state_dict)
    args.batch_size = 4
    model.cache_packed_cells = True
    ds = e2e_dataset()
    loader = DataLoader(ds, batch_size=4, collate_fn=dataio.collate_fn, shuffle=False)
    calls = []
    real = dataio.pack_cell_database
    monkeypatch.setattr(dataio, "pack_cell_database", lambda *a, **k: calls.append(1) or real(*a, **k))
    np.random.seed(5)
    acc1, _, r1, cells1, text1 = eval_epoch(model, loader, args, return_encodings=True)
    acc2, _, r2 = eval_epoch(model, loader, args)
    assert len(calls) == 1
    got = np.stack([r1[i] for i in range(len(ds))])
    assert (got == np.stack([r2[i] for i in range(len(ds))])).all() and acc1 == acc2
    assert np.abs(np.linalg.norm(cells1, axis=1) - 1).max() < 1e-5
    oidx, _ = restate.search_topk(cells1, text1, 10)
    assert (got == np.array([c.id for c in ds.all_cells])[oidx]).all()


def test_dropin_surface(state_dict):
    from text2loc_b200.engine import EngineError

    model, args = make_model(state_dict)
    assert model.embed_dim == 256 and model.object_size == 28
    assert model.device.type == "cuda" and model.get_device() == model.device
    assert model.to("cuda") is model
    with pytest.raises(Exception, match="Not implemented"):
        model.forward()
    with pytest.raises(EngineError):
        model.to("cpu")
    bad = argparse.Namespace(**{**vars(args), "coarse_embed_dim": 128})
    from text2loc_b200 import CellRetrievalNetwork

    with pytest.raises(EngineError):
        CellRetrievalNetwork([], [], bad)
    # a checkpoint with other dimensions is rejected with a size-mismatch error (load_state_dict raises in the reference too)
    wrong = {k: np.asarray(v) for k, v in state_dict.items()}
    wrong["object_encoder.mlp_merge.0.0.weight"] = wrong["object_encoder.mlp_merge.0.0.weight"][:, :512]
    with pytest.raises(EngineError, match="size mismatch"):
        make_fresh = CellRetrievalNetwork(["c"] * 22, ["k"] * 8, args, text_frontend=lambda d: None)
        make_fresh.load_state_dict({k: torch.as_tensor(v) for k, v in wrong.items()}, strict=False)
    # engine calls leave the caller's current device alone
    assert torch.cuda.current_device() == 0
