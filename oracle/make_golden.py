"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every vector below is an output of the unmodified reference modules imported from
/root/reference under oracle/stubs.py (PyG ops = oracle/pyg_ops.py, T5 = oracle/fake_t5.py),
on seeded synthetic inputs with the seeded synthetic weights of text2loc_b200/synth.py.

  cells_small.npz   CellRetrievalNetwork.encode_objects on ragged cells (1..30 objects, one
                    cell beyond object_size=28, duplicate-point objects, two NormalizeScale'd
                    cells whose ball queries do not hit the 32-neighbour cap)
  text_small.npz    CellRetrievalNetwork.encode_text on 8 six-sentence descriptions
  eval_e2e.npz      training.coarse.eval_epoch(return_encodings=True) + evaluation.coarse.run_coarse
                    on a 24-cell / 40-pose synthetic dataset
  search_small.npz  the training/coarse.py:119-125 loop on random unit rows
  eval_cfg1.npz     evaluation.coarse.run_coarse + eval_epoch(return_encodings=True) on BASELINE configs[0]: 1 000 cells x
                    8 objects, 256 queries, seed 1, with the wall time of the reference's run   (python -m oracle.make_golden cfg1)
  fine_small.npz    CrossMatch.forward (models/cross_matcher.py:83-129, the fine stage = SURVEY.md section 8f row 1) on
                    5 cells padded to 16 objects x 6 hints: offsets [5, 2]   (python -m oracle.make_golden fine)
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

import synth
from text2loc_b200 import dataio

from . import fake_t5, reference_run, restate

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WEIGHT_SEED = 0
FAKE_T5_SEED = 0


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def cells_case():
    """Ragged cells; returns raw objects + per-cell point batches (FixedPoints applied once)."""
    counts = [1, 3, 8, 2, 30, 16, 5, 4]
    cells = synth.make_cell_objects(11, len(counts), counts, max_raw=700)
    # degenerate objects: a single raw point (all 256 samples identical) and a padding-like blob
    rng = np.random.default_rng(99)
    cells[1][0] = synth.SynthObject(0, np.array([[0.3, 0.4, 0.05]]), np.array([[0.2, 0.5, 0.7]]))
    cells[2][3] = synth.SynthObject(3, rng.random((8, 3)) * 0.001, np.zeros((8, 3)))
    np.random.seed(2024)
    fixed = dataio.FixedPoints(256)
    both = dataio.Compose([dataio.FixedPoints(256), dataio.NormalizeScale()])
    batches = [dataio.batch_object_points(objs, both if i >= 6 else fixed) for i, objs in enumerate(cells)]
    return cells, batches


def text_case():
    rng = np.random.default_rng(5)
    out = []
    for _ in range(8):
        out.append(" ".join(
            f"The pose is {synth.DIRECTIONS[rng.integers(5)]} of a {synth.COLOR_WORDS[rng.integers(8)]} {synth.CLASS_WORDS[rng.integers(22)]}."
            for _ in range(6)))
    return out


def e2e_dataset():
    counts = [int(c) for c in np.random.default_rng(8).integers(1, 13, 24)]
    counts[5] = 31
    return synth.SynthCoarseDataset(seed=3, n_cells=24, n_poses=40, n_obj=counts, max_raw=400)


def fine_case():
    """5 top-k cells padded to pad_size = 16 objects, one 6-hint description per cell."""
    cells = synth.make_cell_objects(31, 5, [16] * 5, max_raw=500)
    np.random.seed(7)
    fixed = dataio.FixedPoints(256)
    batches = [dataio.batch_object_points(objs, fixed) for objs in cells]
    rng = np.random.default_rng(6)
    texts = [" ".join(
        f"The pose is {synth.DIRECTIONS[rng.integers(5)]} of a {synth.COLOR_WORDS[rng.integers(8)]} {synth.CLASS_WORDS[rng.integers(22)]}."
        for _ in range(6)) for _ in range(5)]
    return cells, batches, texts


def main_fine():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    sd = synth.make_fine_state_dict(WEIGHT_SEED)
    model = reference_run.build_fine_model(sd, fake_seed=FAKE_T5_SEED)
    cells, batches, texts = fine_case()
    with torch.no_grad():
        offsets = model(cells, texts, batches).numpy()
    pts, meta, cell_ptr = dataio.pack_cells(cells, batches)
    feat, n_sent = fake_t5.FakeFrontend(FAKE_T5_SEED)(texts)
    got = restate.fine_offsets(sd, pts, meta, cell_ptr, feat, n_sent).numpy()
    print("fine: restatement vs reference max abs diff", np.abs(got - offsets).max())
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "fine_small.npz"),
        pts=pts.numpy(), meta=meta.numpy(), cell_ptr=cell_ptr.numpy(), texts=np.array(texts), offsets=offsets,
        t5_digest=digest(feat.numpy()), n_sent=n_sent, weight_seed=WEIGHT_SEED, fake_t5_seed=FAKE_T5_SEED,
    )
    print("fine_small.npz", os.path.getsize(os.path.join(GOLDEN_DIR, "fine_small.npz")))


def cfg1_dataset():
    """BASELINE.json configs[0] / SURVEY.md section 8d config 1: 1 000 cells x 8 objects (30..5000 raw points each),
    256 queries, seed 1."""
    return synth.SynthCoarseDataset(seed=1, n_cells=1000, n_poses=256, n_obj=8, max_raw=5000)


CFG1_NP_SEED = 1
CFG1_BATCH = 8


def main_cfg1():
    """eval_cfg1.npz: the reference's own evaluation.coarse.run_coarse (evaluation/coarse.py:40-84, which calls
    training/coarse.py::eval_epoch) on configs[0], all three loops, timed on this container's cores."""
    import time

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(WEIGHT_SEED)
    model = reference_run.build_model(sd, fake_seed=FAKE_T5_SEED)
    ref = reference_run.load()
    args = reference_run.default_args()
    args.batch_size = CFG1_BATCH
    ds = cfg1_dataset()
    loader = DataLoader(ds, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    np.random.seed(CFG1_NP_SEED)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        retrievals, accuracies = ref["run_coarse"](model, loader, args)
    t_run = time.perf_counter() - t0
    np.random.seed(CFG1_NP_SEED)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        acc, acc_close, retr, cell_enc, text_enc = ref["eval_epoch"](model, loader, args, return_encodings=True)
    assert all((retrievals[i] == retr[i]).all() for i in range(len(ds)))
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "eval_cfg1.npz"),
        retrievals=np.stack(retrievals), acc=np.array([acc[k] for k in args.top_k]),
        acc_close=np.array([acc_close[k] for k in args.top_k]),
        run_coarse_acc=np.array([[accuracies[k][t] for t in args.threshs] for k in args.top_k]),
        cell_enc=cell_enc.astype(np.float32), text_enc=text_enc.astype(np.float32),  # f32 values held in f64 buffers
        top_k=np.array(args.top_k), threshs=np.array(args.threshs), np_seed=CFG1_NP_SEED, batch_size=args.batch_size,
        weight_seed=WEIGHT_SEED, fake_t5_seed=FAKE_T5_SEED, reference_run_coarse_seconds=t_run, reference_cores=os.cpu_count(),
    )
    print(f"cfg1: reference run_coarse took {t_run:.1f} s on {os.cpu_count()} cores; acc {acc} close {acc_close}")
    print("eval_cfg1.npz", os.path.getsize(os.path.join(GOLDEN_DIR, "eval_cfg1.npz")))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(WEIGHT_SEED)
    model = reference_run.build_model(sd, fake_seed=FAKE_T5_SEED)
    ref = reference_run.load()
    args = reference_run.default_args()

    # ---- cells
    cells, batches = cells_case()
    with torch.no_grad():
        cell_emb = model.encode_objects(cells, batches).numpy()
    pts, meta, cell_ptr = dataio.pack_cells(cells, batches)
    out, aux = restate.encode_cells(sd, pts, meta, cell_ptr, return_aux=True)
    print("cells: restatement vs reference max abs diff", np.abs(out.numpy() - cell_emb).max())
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "cells_small.npz"),
        pts=pts.numpy(), meta=meta.numpy(), cell_ptr=cell_ptr.numpy(), cell_emb=cell_emb,
        fps1=aux["fps1"].numpy().astype(np.int16), fps2=aux["fps2"].numpy().astype(np.int16),
        fps3=aux["fps3"].numpy().astype(np.int16),
        nbr_digest=np.array([digest(aux[f"nbr{i}"].numpy().astype(np.int16)) for i in (1, 2, 3)]),
        nbr_count=np.array([(aux[f"nbr{i}"] >= 0).sum().item() for i in (1, 2, 3)]),
        features2=aux["features2"].numpy(), object_emb=aux["object_emb"].numpy(),
        weight_seed=WEIGHT_SEED,
    )

    # ---- text
    texts = text_case()
    with torch.no_grad():
        text_emb = model.encode_text(texts).numpy()
    feat, n_sent = fake_t5.FakeFrontend(FAKE_T5_SEED)(texts)
    out = restate.encode_text(sd, feat, n_sent)
    print("text: restatement vs reference max abs diff", np.abs(out.numpy() - text_emb).max())
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "text_small.npz"),
        texts=np.array(texts), text_emb=text_emb, t5_shape=np.array(feat.shape), t5_digest=digest(feat.numpy()),
        n_sent=n_sent, weight_seed=WEIGHT_SEED, fake_t5_seed=FAKE_T5_SEED,
    )

    # ---- end to end through eval_epoch / run_coarse
    ds = e2e_dataset()
    args.batch_size = 4
    loader = DataLoader(ds, batch_size=args.batch_size, collate_fn=dataio.collate_fn, shuffle=False)
    np.random.seed(123)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        acc, acc_close, retr, cell_enc, text_enc = ref["eval_epoch"](model, loader, args, return_encodings=True)
    np.random.seed(123)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        retrievals, accuracies = ref["run_coarse"](model, loader, args)
    assert all((retrievals[i] == retr[i]).all() for i in range(len(ds)))
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "eval_e2e.npz"),
        retrievals=np.stack([retr[i] for i in range(len(ds))]),
        acc=np.array([acc[k] for k in args.top_k]), acc_close=np.array([acc_close[k] for k in args.top_k]),
        run_coarse_acc=np.array([[accuracies[k][t] for t in args.threshs] for k in args.top_k]),
        cell_enc=cell_enc, text_enc=text_enc, top_k=np.array(args.top_k), threshs=np.array(args.threshs),
        np_seed=123, batch_size=args.batch_size, weight_seed=WEIGHT_SEED, fake_t5_seed=FAKE_T5_SEED,
    )
    print("e2e: acc", acc, "close", acc_close)

    # ---- search loop alone
    D = synth.make_unit_rows(21, 3000)
    Q = synth.make_unit_rows(22, 64)
    idx_ref = restate.search_topk_reference_loop(D.astype(np.float64), Q.astype(np.float64), 10)
    idx, sc = restate.search_topk(D, Q, 10)
    assert (idx == idx_ref).all(), "stable-order oracle differs from the reference loop on tie-free data"
    np.savez_compressed(os.path.join(GOLDEN_DIR, "search_small.npz"), d_seed=21, q_seed=22, n=3000, nq=64, idx=idx_ref, score=sc)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    import sys

    if len(sys.argv) > 1 and sys.argv[1] == "fine":
        main_fine()
    elif len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        main_cfg1()
    else:
        main()
