"""CPU tests of the host side: the C-ABI library builds, loads and exports every symbol the header
declares (no compute without a GPU), weight folding, input packing, the failure mode without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from text2loc_b200 import _lib

    _lib.build()
    return _lib.load()


def test_library_exports_every_header_symbol(lib):
    from text2loc_b200 import _lib

    header = open(os.path.join(ROOT, "include", "text2loc_b200.h")).read()
    declared = set(re.findall(r"\b(t2l_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None
    assert lib.t2l_version() == 1


def test_built_library_is_blackwell_native(lib):
    """The hot kernels of the built .so carry the sm_100a tensor-core / TMA instructions in their SASS (cuobjdump):
    UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st (tensor memory), UTMALDG / UTMASTG = TMA tensor load / store,
    UBLKCP = cp.async.bulk.  Guards against a build that silently fell back to SIMT code paths."""
    import shutil
    import subprocess
    from collections import Counter, defaultdict

    from text2loc_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    per_kernel = defaultdict(Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP"):
            if op in line:
                per_kernel[cur][op] += 1
    gemms = {k: v for k, v in per_kernel.items() if "umma_gemm_kernel" in k}
    sa = {k: v for k, v in per_kernel.items() if "sa_obj2_kernel" in k}
    assert len(gemms) >= 20 and len(sa) == 3
    for name, ops in gemms.items():  # TMA-fed tcgen05 tiles, accumulators read back from tensor memory
        assert ops["UTCHMMA"] and ops["LDTM"] and ops["UTMALDG"], (name, dict(ops))
    for name, ops in sa.items():  # bulk-copied object blocks, W2 parked in tensor memory (tcgen05.st), TS-mode MMAs
        assert ops["UTCHMMA"] and ops["LDTM"] and ops["STTM"] and ops["UBLKCP"], (name, dict(ops))
    assert any("TopKEpi" in k for k in gemms), "the fused top-k search epilogue is missing"
    assert any("ResidualTmaEpi" in k and v["UTMASTG"] for k, v in gemms.items()), "the TMA-store residual epilogue is missing"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_cuda(lib):
    from text2loc_b200.engine import Engine, EngineError

    h = ctypes.c_void_p()
    assert lib.t2l_create(0, ctypes.byref(h)) != 0
    assert b"no CPU path" in lib.t2l_last_error(None)
    with pytest.raises(EngineError):
        Engine()


def test_bn_folding_matches_torch_modules(state_dict):
    from text2loc_b200 import weights

    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
    pre = "object_encoder.mlp_merge.0"
    lin = torch.nn.Linear(1024, 256)
    bn = torch.nn.BatchNorm1d(256)
    lin.load_state_dict({"weight": sd[pre + ".0.weight"], "bias": sd[pre + ".0.bias"]})
    bn.load_state_dict({k: sd[f"{pre}.1.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")})
    bn.eval()
    x = torch.randn(16, 1024)
    W, b = weights.fold_linear_bn(state_dict, pre)
    with torch.no_grad():
        want = bn(lin(x))
    got = x @ torch.from_numpy(W).T + torch.from_numpy(b)
    assert (got - want).abs().max() < 1e-5


def test_engine_weight_names_cover_what_the_engine_requires(state_dict):
    from text2loc_b200 import weights

    ew = weights.engine_weights(state_dict)
    api = open(os.path.join(ROOT, "text2loc_b200", "csrc", "api.cu")).read()
    required = set(re.findall(r'"((?:sa\d|ga|lin\d|mlp_pointnet|color|pos|num|merge|txt_mlp)\.[a-z0-9]+)"', api))
    for a in ("obj_attn0", "obj_attn1", "txt_intra", "txt_inter"):
        required |= {f"{a}.{p}" for p in ("in_w", "in_b", "out_w", "out_b", "l1_w", "l1_b", "l2_w", "l2_b", "n1_w", "n1_b", "n2_w", "n2_b")}
    assert required <= set(ew), required - set(ew)
    assert ew["sa2.w1x"].shape == (128, 64) and ew["sa2.w1p"].shape == (128, 3) and ew["ga.w1"].shape == (512, 259)
    assert all(v.dtype == np.float32 and v.ndim == 2 for v in ew.values())


def test_pack_cells_layout_and_errors():
    import synth
    from text2loc_b200 import dataio

    cells = synth.make_cell_objects(1, 2, [2, 3], max_raw=100)
    np.random.seed(0)
    batches = [dataio.batch_object_points(o, dataio.FixedPoints(256)) for o in cells]
    pts, meta, ptr = dataio.pack_cells(cells, batches)
    assert pts.shape == (5, 256, 6) and meta.shape == (5, 7) and ptr.tolist() == [0, 2, 5]
    assert torch.equal(pts[2, :, 0:3], batches[1].pos[:256]) and torch.equal(pts[2, :, 3:6], batches[1].x[:256])
    assert np.allclose(meta[3, 3:6].numpy(), cells[1][1].get_center(), atol=1e-6) and meta[3, 6] == len(cells[1][1].xyz)
    bad = dataio.PointsBatch(batches[0].x[:300], batches[0].pos[:300])
    with pytest.raises(ValueError):
        dataio.pack_cells(cells[:1], [bad])


def test_packed_generator_matches_object_generator_statistics():
    import synth

    pts, meta, ptr = synth.make_packed_cells(0, 50, 8)
    assert pts.shape == (400, 256, 6) and ptr[-1] == 400 and pts.dtype == np.float32
    assert 0 <= pts[:, :, 3:].min() and pts[:, :, 3:].max() <= 1 and 30 <= meta[:, 6].min() and meta[:, 6].max() <= 5000
    span = pts[:, :, :3].max(axis=1) - pts[:, :, :3].min(axis=1)
    assert span.max() < 0.31  # objects are smaller than every ball-query radius's diameter: the 32-neighbour cap is hit


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path through the oracle port) prints one JSON line with the
    keys the driver reads; it needs no GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert "workload" in line["config"]


def test_rows_of_ids_maps_strings_to_database_rows():
    from text2loc_b200.evaluation import rows_of_ids

    db = np.array(["0003_00002", "0000_00010", "0003_00000", "0001_00007"], dtype="<U32")
    q = np.array([["0001_00007", "0003_00000"], ["nope", "0003_00002"]])
    assert (rows_of_ids(db, q) == np.array([[3, 2], [-1, 0]])).all()
    assert rows_of_ids(db, np.array([], dtype="<U32")).shape == (0,)


def test_oracle_bookkeeping_restatement_matches_reference_functions():
    """restate.calc_sample_accuracies / localisation_accuracies against the reference's own evaluation/utils.py."""
    from oracle import reference_run, restate
    import synth

    if not reference_run.available():
        pytest.skip("reference not present")
    reference_run.load()
    from evaluation.utils import calc_sample_accuracies as ref_calc

    rng = np.random.default_rng(3)
    cells = [synth.SynthCell(i, f"{i % 2:04d}", [], 30.0, np.array([10.0 * i, 5.0 * i, 0, 10.0 * i + 30, 5.0 * i + 30, 30])) for i in range(12)]
    pose = synth.SynthPose(np.array([40.0, 30.0, 1.0]), cells[4].id, cells[4].scene_name, "")
    top = [cells[i] for i in rng.permutation(12)[:10]]
    pos = rng.uniform(0, 1, (10, 2))
    want = ref_calc(pose, top, pos, [1, 3, 5, 10], [5, 10, 15])
    got = restate.calc_sample_accuracies(pose, top, pos, [1, 3, 5, 10], [5, 10, 15])
    assert got == want


def test_sentence_cache_frontend_reproduces_the_uncached_front_end():
    """Each distinct sentence passes the frozen encoder once; batches are assembled from the cache and equal what the
    uncached front end (the reference's steps, models/language_encoder.py:108-125) produces, pad positions included."""
    from oracle import fake_t5
    from oracle.stubs import sent_tokenize
    from text2loc_b200.text_frontend import SentenceCacheFrontend

    plain = fake_t5.FakeFrontend(0)
    cached = SentenceCacheFrontend(fake_t5.FakeTokenizer(), fake_t5.FakeT5Encoder(0).eval(), "cpu", cap=16, split=sent_tokenize)
    a = ["The pose is north of a gray building. The pose is on-top of a dark-green traffic light.",
         "The pose is east of a red pole. The pose is north of a gray building."]
    b = ["The pose is west of a beige vending machine. The pose is east of a red pole."]
    for batch in (a, b, a + b):
        want, ns = plain(batch)
        got, ns2 = cached(batch)
        assert ns == ns2 == 2 and got.shape == want.shape
        assert torch.equal(got, want)
    assert cached.encoder_calls == 2 and len(cached.cache) == 4  # the third batch was served from the cache alone


def test_sentence_row_cache_computes_each_distinct_sentence_once():
    """SentenceRowCache (token stage of the text head per distinct (sentence, n_tok)): a batch assembled from the cache equals
    the row-by-row computation, repeated sentences are computed once, another padding length is another key, and the table
    is bounded."""
    from text2loc_b200.text_frontend import SentenceRowCache

    calls = []

    def row_of(s, n_tok):
        return torch.full((4,), float(len(s) * 100 + n_tok))

    def compute_for(n_tok):
        def compute(new):
            calls.append(list(new))
            assert len(set(new)) == len(new)
            return torch.stack([row_of(s, n_tok) for s in new])
        return compute

    cache = SentenceRowCache(max_rows=6)
    batch = ["a", "bb", "a", "ccc", "bb", "a"]
    got = cache.rows(batch, 9, compute_for(9))
    assert torch.equal(got, torch.stack([row_of(s, 9) for s in batch]))
    assert calls == [["a", "bb", "ccc"]] and cache.computed == 3
    got = cache.rows(["ccc", "a"], 9, compute_for(9))  # served from the table alone
    assert torch.equal(got, torch.stack([row_of("ccc", 9), row_of("a", 9)])) and len(calls) == 1
    got = cache.rows(["a", "dddd"], 11, compute_for(11))  # same sentence, longer padding: a different row
    assert torch.equal(got, torch.stack([row_of("a", 11), row_of("dddd", 11)])) and calls[-1] == ["a", "dddd"]
    assert len(cache.index) == 5
    got = cache.rows(["e", "ff", "a"], 9, compute_for(9))  # 5 + 2 new > max_rows: the cache starts over with this batch
    assert torch.equal(got, torch.stack([row_of(s, 9) for s in ("e", "ff", "a")]))
    assert calls[-1] == ["e", "ff", "a"] and len(cache.index) == 3 and cache.table.shape == (3, 4)
    cache.clear()  # what load_state_dict does: rows of the previous weights must not survive
    assert cache.index == {} and cache.table is None
    got = cache.rows(["a"], 9, compute_for(9))
    assert torch.equal(got, row_of("a", 9)[None]) and calls[-1] == ["a"]


def test_sentence_cache_frontend_prepare_and_states():
    """prepare() / states() are the two halves of the cached front end's call (the drop-in models use them to run the
    token stage per distinct sentence)."""
    from oracle import fake_t5
    from oracle.stubs import sent_tokenize
    from text2loc_b200.text_frontend import SentenceCacheFrontend

    fe = SentenceCacheFrontend(fake_t5.FakeTokenizer(), fake_t5.FakeT5Encoder(0).eval(), "cpu", cap=16, split=sent_tokenize)
    batch = ["The pose is north of a gray building. The pose is east of a red pole.",
             "The pose is east of a red pole. The pose is on-top of a dark-green traffic light."]
    sentences, n_sent, n_tok = fe.prepare(batch)
    assert n_sent == 2 and len(sentences) == 4 and sentences[1] == sentences[2]
    full, ns = fe(batch)
    assert ns == 2 and full.shape == (4, n_tok, 1024)
    assert torch.equal(fe.states(sentences, n_tok), full)
    assert torch.equal(fe.states([sentences[3]], n_tok)[0], full[3])


def test_pack_cell_database_vectorised_matches_per_object_packing():
    """meta is the per-object reductions of pack_cells; every sampled point is one of the object's raw points; the
    NormalizeScale variant centres each sample and scales it into (-1, 1)."""
    import synth
    from text2loc_b200 import dataio

    objs = synth.make_cell_objects(3, 5, [2, 9, 1, 4, 30], max_raw=400)
    cells = [synth.SynthCell(i, "0000", o, 30.0, np.zeros(6)) for i, o in enumerate(objs)]
    db = dataio.pack_cell_database(cells, rng=np.random.default_rng(0))
    np.random.seed(0)
    pts, meta, ptr = dataio.pack_cells(objs, [dataio.batch_object_points(o, dataio.FixedPoints(256)) for o in objs])
    assert (db.cell_ptr == ptr).all() and db.pts.shape == pts.shape and db.cell_ids == [c.id for c in cells]
    assert np.abs(db.meta.numpy() - meta.numpy()).max() < 1e-6 and (db.meta[:, 6] == meta[:, 6]).all()
    flat = [o for c in objs for o in c]
    for k in (0, 7, 45):
        raw = np.concatenate([flat[k].xyz, flat[k].rgb], axis=1).astype(np.float32)
        d = np.abs(db.pts[k].numpy()[:, None, :] - raw[None, :, :]).max(axis=2).min(axis=1)
        assert d.max() == 0.0
    ns = dataio.pack_cell_database(cells, normalize_scale=True, rng=np.random.default_rng(0))
    p = ns.pts[:, :, 0:3].numpy()
    assert np.abs(p.mean(axis=1)).max() < 1e-5 and abs(np.abs(p).max(axis=(1, 2)).max() - 0.999999) < 1e-5


def test_run_fine_vectorised_packing_feeds_the_engine_the_same_layout(monkeypatch):
    """run_fine's opt-in vectorised packing (model.vectorised_packing) against the per-object path, with the engine replaced
    by a recorder: same shapes and cell_ptr, identical per-object meta rows (mean colour, centre, raw count, padding objects
    included), every sampled point one of its object's raw points."""
    import types

    import synth
    from oracle import fake_t5, reference_run
    from text2loc_b200 import dataio, evaluation

    class Recorder:
        FINE_DIM, device = 128, torch.device("cpu")

        def __init__(self):
            self.calls = []

        def fine_encode_objects(self, pts, meta, cell_ptr):
            self.calls.append((pts.clone(), meta.clone(), np.asarray(cell_ptr).copy()))
            return torch.zeros((pts.shape[0], 128))

        def fine_encode_hints(self, feats):
            return torch.zeros((feats.shape[0], 128))

        def fine_match(self, obj_emb, pair_cell, hints, pair_query, n_obj, n_hints):
            return torch.full((len(pair_cell), 2), 0.5)

    args = reference_run.fine_args(top_k=[1, 3], threshs=[5, 10])
    ds = synth.SynthCoarseDataset(seed=4, n_cells=6, n_poses=4, n_obj=[3, 16, 20, 1, 8, 5], max_raw=200)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, collate_fn=dataio.collate_fn, shuffle=False)
    ids = np.array([c.id for c in ds.all_cells])
    retrievals = np.stack([ids[[0, 2, 4]], ids[[1, 2, 3]], ids[[5, 0, 1]], ids[[2, 4, 5]]])
    monkeypatch.setattr(evaluation, "localisation_accuracies", lambda *a, **k: {1: {5: 0.0, 10: 0.0}, 3: {5: 0.0, 10: 0.0}})
    seen = {}
    for mode in (False, True):
        rec = Recorder()
        frontend = fake_t5.FakeFrontend(0)
        model = types.SimpleNamespace(engine=rec, vectorised_packing=mode, eval=lambda: None,
                                      encode_hints=lambda d, rec=rec, fe=frontend: (rec.fine_encode_hints(fe(d)[0]), fe(d)[1]))
        np.random.seed(11)
        acc, offsets = evaluation.run_fine(model, retrievals, loader, args, dataio.Compose([dataio.FixedPoints(256), dataio.NormalizeScale()]),
                                           return_offsets=True)
        assert offsets.shape == (4, 3, 2) and len(rec.calls) == 1
        seen[mode] = rec.calls[0]
    (pts_a, meta_a, ptr_a), (pts_b, meta_b, ptr_b) = seen[False], seen[True]
    assert pts_a.shape == pts_b.shape == (6 * 16, 256, 6) and (ptr_a == ptr_b).all() and (ptr_a == np.arange(7) * 16).all()
    assert torch.equal(meta_a, meta_b)  # padding objects are drawn before any sampling in both modes
    for pts in (pts_a, pts_b):  # NormalizeScale: centred, inside (-1, 1), the largest coordinate at 0.999999
        assert float(pts[:, :, :3].mean(dim=1).abs().max()) < 1e-4
        assert np.allclose(pts[:, :, :3].abs().amax(dim=(1, 2)).numpy(), 0.999999, atol=1e-6)
    with pytest.raises(ValueError):
        model.vectorised_packing = True
        evaluation.run_fine(model, retrievals, loader, args, lambda d: d)


def test_host_chunk_schedule_covers_every_query_once():
    """engine.host_chunk_schedule: the H2D staging chunks of a host-streamed encode_text partition the batch, never exceed the
    staging buffer, and start / end with short chunks when the batch is long."""
    from text2loc_b200.engine import host_chunk_schedule

    for nq in (0, 1, 5, 37, 221, 222, 223, 4096, 32768):
        for cq in (1, 3, 4, 37, 227, 5000):
            sizes = host_chunk_schedule(nq, cq)
            assert sum(sizes) == nq and all(0 < n <= cq for n in sizes), (nq, cq, sizes)
            if nq >= 6 * cq and cq >= 4:
                assert sizes[0] == cq // 4 and sizes[-1] == cq // 4 and sizes[1] == cq // 2 and sizes[-2] == cq // 2
