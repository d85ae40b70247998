#!/usr/bin/env python
"""Headline benchmark: coarse-retrieval queries/sec over an N-cell database (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload)
  N=1   BASELINE configs[1]: 10 000 cells x 4 096 queries, 8 objects/cell x 256 points, d=256, k=10
  N>1   weak scaling towards configs[2]: 12 500 cells and 4 096 queries PER GPU (N=8: 100 000 x 32 768),
        database row-sharded, one all-gather of per-shard top-k.

A "step"  is the query path over one batch: text head on the T5 features ->
(all-gather of query embeddings) -> tensor-core candidate search + fp64 re-rank over the local shard
-> (all-gather + merge of per-shard top-k).  The database is encoded once before the timed region;
that encode is timed on its own (db_encode_cells_per_s) and folded into cold_db_qps.
`value` times the step with inputs resident in HBM; `e2e` times the same call chain through the
public drop-in objects with HOST buffers: pinned-host T5 features copied H2D and the top-k copied
D2H inside the timed region, every step.

--impl reference times the reference's own CPU implementation of the same path (the oracle port
of its Python: torch CPU text head + numpy float64 GEMV/argsort loop) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_TOP = 10
N_SENT, N_TOK = 6, 12
OBJ_PER_CELL = 8
METRIC = "coarse-retrieval queries/sec over N-cell DB; top-k index match"


def workload(n_gpus: int):
    if n_gpus == 1:
        return dict(cells_per_gpu=10000, queries_per_gpu=4096, name="configs[1]: 10k cells x 4k queries, 8 obj/cell x 256 pts, d=256, k=10, 1xB200")
    return dict(cells_per_gpu=12500, queries_per_gpu=4096,
                name=f"weak-scaled configs[2]: {12500 * n_gpus} cells x {4096 * n_gpus} queries, DB row-sharded over {n_gpus}xB200 "
                     f"(12.5k cells + 4k queries per GPU), all-gather of per-shard top-k")


def bench_config(world: int) -> dict:
    """The `config` object of the JSON line -- ONE definition for both arms, so the driver's same-config check compares
    like with like."""
    wl = workload(world)
    return {"workload": wl["name"], "n_cells": wl["cells_per_gpu"] * world, "n_queries": wl["queries_per_gpu"] * world, "k": K_TOP,
            "objects_per_cell": OBJ_PER_CELL, "sentences_x_tokens": [N_SENT, N_TOK],
            "timed_region": "text head + search (+ all-gathers, merge), DB pre-encoded",
            "l2": "inputs larger than L2 (1.2 GB of fp32 T5 features per GPU per step)", "parallelism": f"db-rowshard{world}"}


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present and readable, else the fallback
    B200_PROFILING.md states (6.65 TB/s, 1.59 PFLOP/s burst / ~1.4 sustained)."""
    fb = dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.isfile(p):
        return fb
    try:
        with open(p) as f:
            d = json.load(f)

        def pick(*names):
            for n in names:
                v = d.get(n)
                if isinstance(v, dict):
                    v = v.get("value")
                if isinstance(v, (int, float)) and v > 0:
                    return float(v)
            return None

        hbm = pick("hbm_gbs", "hbm_gb_s", "hbm_GBps", "copy_gbs")
        bf16 = pick("bf16_tflops", "bf16_tflops_burst", "bf16_tf")
        sus = pick("bf16_tflops_sustained", "bf16_sustained_tflops") or bf16
        if hbm and hbm < 100:  # TB/s -> GB/s
            hbm *= 1000.0
        if hbm and bf16:
            return dict(hbm_gbs=hbm, bf16=bf16, bf16_sustained=sus, source="measured")
    except Exception:
        pass
    return fb


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so pinned staging buffers are allocated on the
    NUMA node the GPU's PCIe root hangs off (first touch) and H2D copies do not cross the socket interconnect.
    Best effort: any failure leaves the affinity unchanged."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md), sampled through NVML in-process.

    One sampler per job: rank 0 watches every GPU of the run.  (Round 1 forked `nvidia-smi` five times a second from EVERY rank;
    at 8 ranks those forks and their driver locks sat on the launch path of all eight processes.  NVML is the library
    nvidia-smi reads the same fields from.)"""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, indices, active: bool = True):
        self.indices = list(indices) if not isinstance(indices, int) else [indices]
        self.active = active
        self.rows, self.stop, self.t = [], threading.Event(), None
        self.nvml = None

    def _run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self.nvml = pynvml
        except Exception:
            return
        while not self.stop.is_set():
            for h in handles:
                try:
                    self.rows.append((self.nvml.nvmlDeviceGetClockInfo(h, self.nvml.NVML_CLOCK_SM),
                                      self.nvml.nvmlDeviceGetMaxClockInfo(h, self.nvml.NVML_CLOCK_SM),
                                      int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
                except Exception:
                    pass
            self.stop.wait(0.1)

    def __enter__(self):
        if self.active:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.t:
            self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for n, bit in self.REASONS if any(r[2] & bit for r in self.rows)]
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_min_mhz": float(sm[0]), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "gpus_sampled": len(self.indices), "source": "NVML (in-process, rank 0)"}


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's own path, timed on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_reference_sample(sd, n_db: int, n_queries_total: int, text_q: int, search_q: int, encode_cells: int, seed: int = 0):
    """Bounded sample of the workload through the oracle (reference semantics):
    text head on `text_q` queries, the float64 GEMV + argsort loop for `search_q` queries over an
    n_db-row database, PointNet++/attention encode of `encode_cells` cells.  Returns per-unit times."""
    import torch

    from oracle import restate
    import synth

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t5 = synth.make_t5_features(seed + 100, text_q, N_SENT, N_TOK)
    t0 = time.perf_counter()
    q_emb = restate.encode_text(sd, t5, N_SENT).numpy()
    t_text = (time.perf_counter() - t0) / text_q
    D64 = synth.make_unit_rows(seed + 101, n_db).astype(np.float64)  # cell_encodings is a float64 buffer (training/coarse.py:81)
    Q64 = np.resize(q_emb, (search_q, 256)).astype(np.float64)
    t0 = time.perf_counter()
    restate.search_topk_reference_loop(D64, Q64, K_TOP)
    t_search = (time.perf_counter() - t0) / search_q
    t_cell = None
    if encode_cells:
        pts, meta, ptr = synth.make_packed_cells(seed + 102, encode_cells, OBJ_PER_CELL)
        t0 = time.perf_counter()
        restate.encode_cells(sd, pts, meta, ptr)
        t_cell = (time.perf_counter() - t0) / encode_cells
    return dict(t_text=t_text, t_search=t_search, t_cell=t_cell, threads=threads)


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path on this box's host cores.  /root/reference does not exist on the GPU box, so
    this is the oracle PORT of the reference's Python (oracle/restate.py, pinned to the reference's own modules by
    tests/golden/*.npz): torch CPU text head + the reference's per-query float64 GEMV / argsort loop, all host threads.
    Each step is a bounded SAMPLE of the named workload (the full 4 096-query x 10 000-cell step would take ~15 s, the 8-GPU one
    minutes): `ms_per_step` is the time actually measured per step; `value` = sampled queries / that time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth

    wl = workload(args.gpus)
    n_db, nq = wl["cells_per_gpu"] * args.gpus, wl["queries_per_gpu"] * args.gpus
    sd = synth.make_state_dict(0)
    text_q, search_q = 48, 96
    times, walls = [], []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_reference_sample(sd, n_db, nq, text_q, search_q, 0, seed=step)
        if step >= args.warmup:
            times.append(r)
            walls.append(time.perf_counter() - t0)
    t_text = float(np.mean([r["t_text"] for r in times]))
    t_search = float(np.mean([r["t_search"] for r in times]))
    per_query = t_text + t_search
    value = 1.0 / per_query
    cores = times[0]["threads"]
    sample = (f"per step: text head on {text_q} queries (6 sentences x 12 tokens of T5 features) + float64 GEMV/argsort loop for "
              f"{search_q} queries over the full {n_db}-row database; queries/s = 1 / (t_text + t_search) per query, DB pre-encoded "
              f"as on the engine arm; the oracle port of the reference's Python, {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)), "sampled": True,
        "ms_per_full_step_extrapolated": 1e3 * per_query * nq, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 encoders / f64 scoring", "data": "synthetic", "config": bench_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "ms_text_head_per_query": 1e3 * t_text, "ms_search_per_query": 1e3 * t_search},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_config0_full(sd):
    """configs[0] (1 000 cells x 8 objects x 256 points, 256 queries) END TO END through the port of
    the reference's eval_epoch -- all three loops (query encode, DB encode, float64 search) -- on the host cores."""
    import argparse as ap

    import torch
    from torch.utils.data import DataLoader

    import synth
    from oracle import fake_t5, restate
    from text2loc_b200 import dataio

    torch.set_num_threads(os.cpu_count() or 1)
    ds = synth.SynthCoarseDataset(seed=1, n_cells=1000, n_poses=256, n_obj=8, max_raw=5000)
    loader = DataLoader(ds, batch_size=8, collate_fn=dataio.collate_fn, shuffle=False)
    a = ap.Namespace(top_k=[1, 3, 5, 10], batch_size=8, ranking_loss="pairwise")
    np.random.seed(1)
    t0 = time.perf_counter()
    restate.eval_epoch(sd, loader, a, fake_t5.FakeFrontend(0))
    dt = time.perf_counter() - t0
    return {"seconds": dt, "queries_per_s": 256 / dt, "cells_per_s": 1000 / dt, "cores": os.cpu_count(),
            "what": "oracle port of training/coarse.py::eval_epoch on configs[0]: 256 queries (fake-T5 front end + text head), 1 000 cells x 8 "
                    "objects through PointNet++ / object encoder / attention, float64 GEMV + argsort per query; the reference's OWN run_coarse "
                    "took 171 s on 8 cores in the build container (tests/golden/eval_cfg1.npz)"}


# ---------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------

def run_engine_arm(args):
    import torch
    import torch.distributed as dist

    import synth
    from text2loc_b200 import CellRetrievalNetwork
    from text2loc_b200 import distributed as t2ld

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: the e2e arm streams its T5 features from host memory every step
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = workload(world)
    n_cells_local, nq_local = wl["cells_per_gpu"], wl["queries_per_gpu"]
    n_db, nq = n_cells_local * world, nq_local * world
    row_lo = rank * n_cells_local

    sd = synth.make_state_dict(0)
    ns = argparse.Namespace(coarse_embed_dim=256, pointnet_layers=3, pointnet_variation=0, pointnet_numpoints=256, pointnet_features=2,
                            object_size=28, object_inter_module_num_heads=4, object_inter_module_num_layers=2, inter_module_num_heads=4,
                            inter_module_num_layers=1, intra_module_num_heads=4, intra_module_num_layers=1, class_embed=False,
                            color_embed=False, use_features=["class", "color", "position", "num"], hungging_model="t5-large")

    def no_frontend(descriptions):
        raise RuntimeError("bench.py feeds T5 features directly (encode_text_features)")

    model = CellRetrievalNetwork([], [], ns, text_frontend=no_frontend, device=dev)
    model.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}, strict=False)
    eng = model.engine

    # ---- inputs (synthetic, seeded per rank), host-pinned + device-resident copies
    pts_h, meta_h, ptr = synth.make_packed_cells(1000 + rank, n_cells_local, OBJ_PER_CELL)
    t5_h = torch.from_numpy(synth.make_t5_features(2000 + rank, nq_local, N_SENT, N_TOK)).pin_memory()
    t5_h16 = t5_h.half().pin_memory()  # what the e2e arm ships: the token layer computes on fp16 operand copies either way
    pts_hp, meta_hp = torch.from_numpy(pts_h).pin_memory(), torch.from_numpy(meta_h).pin_memory()
    t5_d = t5_h.to(dev, non_blocking=True)
    pts_d, meta_d = pts_hp.to(dev, non_blocking=True), meta_hp.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---- database encode (before the timed region; timed on its own: one cold run, then three warm ones, median reported)
    enc_ms = []
    for _ in range(4):
        barrier()
        a, b = ev(), ev()
        a.record()
        D_local = model.encode_cells_packed(pts_d, meta_d, ptr)
        b.record()
        torch.cuda.synchronize()
        enc_ms.append(a.elapsed_time(b))
    enc_warm = float(np.median(enc_ms[1:]))
    enc_ms_max = max_over_ranks(enc_warm)
    eng.db_build(D_local, row_offset=row_lo)

    # ---- the step
    def step_resident():
        q_local = model.encode_text_features(t5_d, N_SENT)
        return t2ld.sharded_search(eng, q_local, K_TOP, queries_are_sharded=world > 1)

    def step_e2e():
        q_local = model.encode_text_features(t5_h16, N_SENT)  # pinned host input: the engine streams it H2D in chunks
        idx, score, nfb = t2ld.sharded_search(eng, q_local, K_TOP, queries_are_sharded=world > 1)
        return idx.to("cpu", non_blocking=True), score.to("cpu", non_blocking=True), nfb

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = eng.launch_count
        a, b = ev(), ev()
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps, (eng.launch_count - l0) // steps, out

    # ---- optional (--graph): the resident step as ONE CUDA graph: ~105 kernel launches (+ the collectives at N > 1) captured once
    # after a warm-up and replayed.  Inputs and outputs are static device buffers; the eager step stays the reference the replay
    # is compared with.
    graph, graph_out, graph_err = None, None, None
    if args.graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_resident()
            torch.cuda.current_stream(dev).wait_stream(side)
            barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = step_resident()
            graph.replay()
            torch.cuda.synchronize()
            eager = step_resident()
            torch.cuda.synchronize()
            if not (torch.equal(graph_out[0], eager[0]) and torch.equal(graph_out[1], eager[1])):
                raise RuntimeError("graph replay and eager step disagree")
        except Exception as ex:  # capture is an optimisation: report and fall back to the eager step
            graph, graph_err = None, f"{type(ex).__name__}: {ex}"[:300]
    ok_all = torch.tensor([1 if graph is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    if int(ok_all.item()) == 0:
        graph = None

    def step_graph():
        graph.replay()
        return graph_out

    # ---- two steps in flight (N > 1): consecutive query steps alternate between two engines (own workspace, own copy of the
    # shard) on two streams.  A step still is text head -> all-gather Q -> search -> all-gather top-k -> merge, in order, on
    # its lane; the OTHER lane's text head fills the time this lane waits for the slowest rank at its collectives.
    n_lanes = args.lanes or 1
    lanes = [(model, eng, torch.cuda.current_stream(dev))]
    if n_lanes == 2:
        model_b = CellRetrievalNetwork([], [], ns, text_frontend=no_frontend, device=dev)
        model_b.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}, strict=False)
        model_b.engine.db_build(D_local, row_offset=row_lo)
        lanes = [(model, eng, torch.cuda.Stream(dev)), (model_b, model_b.engine, torch.cuda.Stream(dev))]

    def timed_lanes(steps, warmup):
        cur = torch.cuda.current_stream(dev)

        def run(n):
            out = None
            for _, _, s in lanes:
                s.wait_stream(cur)
            for i in range(n):
                m, e, s = lanes[i % len(lanes)]
                with torch.cuda.stream(s):
                    q_local = m.encode_text_features(t5_d, N_SENT)
                    out = t2ld.sharded_search(e, q_local, K_TOP, queries_are_sharded=world > 1)
            for _, _, s in lanes:
                cur.wait_stream(s)
            return out

        run(warmup)
        barrier()
        l0 = sum(e.launch_count for _, e, _ in lanes)
        a, b = ev(), ev()
        a.record()
        out = run(steps)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps, (sum(e.launch_count for _, e, _ in lanes) - l0) // steps, out

    with ClockSampler(range(world), active=rank == 0) as clocks:
        ms_eager, launches, out = timed(step_resident, args.steps, args.warmup)
        ms_step = ms_eager
        ms_one_lane = ms_eager
        if n_lanes == 2:
            ms_step, launches, out = timed_lanes(args.steps, args.warmup)
        if graph is not None:
            ms_step, _, out = timed(step_graph, args.steps, args.warmup)
        ms_e2e, _, out_e2e = timed(step_e2e, args.steps, args.warmup)
    idx, score, nfb = out

    # stage split (resident): extra passes with CUDA events between the stages and around every collective
    barrier()
    stage = {}
    for _ in range(3):
        e0, e1 = ev(), ev()
        e0.record()
        q_local = model.encode_text_features(t5_d, N_SENT)
        e1.record()
        t2ld.sharded_search(eng, q_local, K_TOP, queries_are_sharded=world > 1, timers=stage)
        stage.setdefault("text_head", []).append(e0.elapsed_time(e1))
    stage_ms = {k: max_over_ranks(float(np.median(v))) for k, v in sorted(stage.items())}
    ms_text, ms_search = stage_ms["text_head"], stage_ms["search"]

    # ---- parity of THIS run's result: the fp64 stable-order oracle over the whole (gathered) database for 64 queries
    if world > 1:
        D_all = t2ld.all_gather_rows(D_local).reshape(-1, 256)
        Q_all = t2ld.all_gather_rows(q_local).reshape(-1, 256)
    else:
        D_all, Q_all = D_local, q_local
    parity = None
    if rank == 0:
        from oracle import restate  # checker only

        sel = np.linspace(0, nq - 1, 64).astype(np.int64)
        oidx, oscore = restate.search_topk(D_all.cpu().numpy(), Q_all[sel].cpu().numpy(), K_TOP)
        parity = bool((idx[sel].cpu().numpy() == oidx).all() and np.abs(score[sel].cpu().numpy() - oscore).max() < 1e-12)
    del D_all

    pk = peaks()
    roof = None
    cpu_base = None
    extra = None
    if rank == 0:
        # ---- roofline of the step's dominant kernel family: the token layer's fp16-operand tcgen05 GEMMs (99 % of the text
        # head's FLOPs), STEP-WEIGHTED: algorithmic FLOPs of the whole text head / its CUDA-event time inside the step, against
        # the SUSTAINED measured bf16 peak (a kernel family timed inside a long step).  The best single GEMM alone is below.
        alg_tf = nq_local * 1.83e9 / (ms_text * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01", "traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("ffn1_f16_bytes_per_launch")
        # DRAM bytes of the whole text head of one step and of one encode chunk, from the committed ncu pass over the same
        # kernels (profiles/r02/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum, 4 096 queries / 16 384 objects)
        traffic_r02 = {}
        tpath = os.path.join(ROOT, "profiles", "r02", "traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic_r02 = json.load(f)
        step_traffic = traffic_r02.get("text_head_step_bytes")
        if step_traffic is not None:
            step_traffic = int(step_traffic * nq_local / 4096)  # linear in the queries of a step (per GPU)
        roof = {"bound": "tensor", "kernel": "umma_gemm_kernel<GemmCfg<256,f16,cta_group::2>,StoreEpiT<...>> x4 per chunk (token layer QKV / out-proj / "
                                             "FFN1 / FFN2) + attention core + LayerNorms = the text head of one step",
                "achieved": alg_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": alg_tf / pk["bf16_sustained"], "traffic": step_traffic,
                "traffic_source": "profiles/r02/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over every launch of the text head of one "
                                  "step, per GPU; algorithmic I/O of the step: %d bytes)" % (nq_local * (N_SENT * N_TOK * 1024 * 4 + 256 * 4)),
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']}); kind::f16 issues at the bf16 rate",
                "ms": ms_text, "algorithmic_flop": nq_local * 1.83e9, "frac_of_burst_peak": alg_tf / pk["bf16"],
                "step_share_text_head": ms_text / ms_step}

        extra = {}
        M, N, K = 32768 // (N_SENT * N_TOK) * (N_SENT * N_TOK), 4096, 1024
        A = torch.randn(M, K, device=dev).half()
        Wt = (torch.randn(N, K, device=dev) / 32).half()
        bias = torch.zeros(N, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def time_alone(fn):
            for _ in range(3):
                fn()
            durs = []
            for _ in range(10):
                flush.zero_()  # L2 flush between timed launches
                a, b = ev(), ev()
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                durs.append(a.elapsed_time(b))
            return float(np.mean(durs))

        dur = time_alone(lambda: eng.debug_linear_f16(A, Wt, bias, act=1, out_half=True))
        achieved = 2.0 * M * N * K / (dur * 1e-3) / 1e12
        cublas = 2.0 * M * N * K / (time_alone(lambda: torch.matmul(A, Wt.T)) * 1e-3) / 1e12
        extra["token_ffn1_alone"] = {"bound": "tensor", "kernel": "umma_gemm_kernel<GemmCfg<256,f16,cta_group::2>,StoreEpiT<half>> FFN1 shape %dx%dx%d alone, "
                                                                  "L2 flushed" % (M, N, K),
                                     "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"], "traffic": traffic,
                                     "ms_per_launch": dur, "cublas_f16_same_shape_tflops": cublas, "frac_of_cublas_f16": achieved / cublas}
        del A, Wt, flush

        # (a) DB encode: compute-bound (377.7 MFLOP + 60.3/8 MFLOP of attention per object vs 7 196 B of I/O);
        #     86 % of the FLOPs are fp16 tcgen05 MMAs, so the f16-rate peak is the denominator; the HBM figure north_star asks for
        #     is reported next to it (the stage is ~53 kFLOP/B: it cannot approach the HBM roof)
        obj_per_s = n_cells_local * OBJ_PER_CELL / (enc_warm * 1e-3)
        flop_per_obj = 377.7e6 + 60.3e6 / OBJ_PER_CELL
        extra["db_encode"] = {"bound": "tensor", "achieved": obj_per_s * flop_per_obj / 1e12, "peak": pk["bf16"], "unit": "TFLOP/s",
                              "frac": obj_per_s * flop_per_obj / 1e12 / pk["bf16"], "frac_of_f16_rate_peak": obj_per_s * flop_per_obj / 1e12 / pk["bf16"],
                              "frac_of_f16_sustained_peak": obj_per_s * flop_per_obj / 1e12 / pk.get("bf16_sustained", pk["bf16"]),
                              "hbm_algorithmic_gbs": obj_per_s * 7196 / 1e9, "hbm_frac": obj_per_s * 7196 / 1e9 / pk["hbm_gbs"],
                              "cells_per_s": n_cells_local / (enc_warm * 1e-3),
                              "traffic_bytes_per_object": traffic_r02.get("encode_chunk_bytes_per_object"),
                              "hbm_measured_traffic_gbs": (obj_per_s * traffic_r02["encode_chunk_bytes_per_object"] / 1e9
                                                           if "encode_chunk_bytes_per_object" in traffic_r02 else None),
                              "note": "algorithmic FLOPs of the whole encode / wall time of encode_cells (FPS, ball query, gathers, attention and all small "
                                      "layers included) against the measured bf16/f16 burst peak"}
        # (b) search at the per-GPU shape of configs[2] (32 768 queries x 12 500 rows) and at 100k rows
        for n_rows in (12500, 100000):
            Dn = torch.from_numpy(synth.make_unit_rows(77, n_rows)).to(dev)
            Qn = torch.from_numpy(synth.make_unit_rows(78, 32768)).to(dev)
            eng.db_build(Dn)
            for _ in range(2):
                eng.search_topk(Qn, K_TOP)
            a, b = ev(), ev()
            a.record()
            for _ in range(5):
                _, _, nfb2 = eng.search_topk(Qn, K_TOP)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            alg = 2.0 * 32768 * n_rows * 256 / (ms * 1e-3) / 1e12
            extra[f"search_32768x{n_rows}"] = {"bound": "tensor", "achieved": alg, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": alg / pk["bf16"],
                                               "ms": ms, "queries_per_s": 32768 / (ms * 1e-3), "second_pass_queries": int(nfb2),
                                               "note": "whole search call: scaling + ONE fp16 tcgen05 pass with fused per-query top-12 x 2 lists + fp64 "
                                                       "re-rank + proof (+ bf16x3 second pass for failed queries); executed = algorithmic FLOPs"}
            del Dn, Qn
        eng.db_build(D_local, row_offset=row_lo)
        # (c) the resident step fed with fp16 T5 states (no conversion kernel)
        t5_d16 = t5_d.half()
        for _ in range(2):
            model.encode_text_features(t5_d16, N_SENT)
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            model.encode_text_features(t5_d16, N_SENT)
        b.record()
        torch.cuda.synchronize()
        extra["text_head_fp16_features"] = {"ms": a.elapsed_time(b) / 3, "queries_per_s_text_head_only": nq_local / (a.elapsed_time(b) / 3 * 1e-3)}
        del t5_d16

        if world == 1 and not args.skip_extras:
            extra["configs3_streamed_small"] = bench_stream(eng, dev, n_cells=20000, nq=4096, obj_per_cell=16, steps=2)[0]
            del model
            extra["configs4_fine_small"] = bench_fine(dev, n_queries=4096, top_cells=5, n_db_cells=2000)

        # ---- CPU baseline beside it (oracle port on the host cores; N=1 only)
        if world == 1:
            r = cpu_reference_sample(sd, n_db, nq, text_q=256, search_q=512, encode_cells=24)
            per_q = r["t_text"] + r["t_search"]
            cpu_base = {"value": 1.0 / per_q, "unit": "queries/s", "cores": r["threads"], "kind": "port",
                        "sample": f"oracle port of the reference path: text head on 256 queries, float64 GEMV+argsort loop for 512 queries over "
                                  f"{n_db} rows, encode of 24 cells; warm-DB queries/s = 1/(t_text+t_search)",
                        "ms_text_head_per_query": 1e3 * r["t_text"], "ms_search_per_query": 1e3 * r["t_search"],
                        "ms_encode_per_cell": 1e3 * r["t_cell"],
                        "cold_db_queries_per_s": nq / (nq * per_q + n_db * r["t_cell"])}
            if not args.skip_extras:
                cpu_base["configs0_end_to_end"] = cpu_config0_full(sd)

    if rank == 0:
        h2d = t5_h16.numel() * 2 * world
        d2h = nq * K_TOP * (8 + 8)
        value = nq / (ms_step * 1e-3)
        cfg = bench_config(world)
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate in the token layer (residual stream carried as f16 rows, sums and LayerNorm statistics in f32), tf32 and 3xtf32 elsewhere; "
                     "search: one f16 candidate pass + f64 re-rank (bf16x3 second pass)", "data": "synthetic", "config": cfg,
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                    "input": "fp16 T5 states in pinned host memory (t2l_encode_text_tokens_f16), streamed H2D in chunks under the compute; "
                             "top-k (i64 row, f64 score) read back D2H every step"},
            "cuda_graph": graph is not None, "cuda_graph_error": graph_err, "ms_per_step_eager": ms_eager,
            "steps_in_flight": n_lanes, "ms_per_step_one_step_in_flight": ms_one_lane,
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roof, "roofline_other_kernels": extra, "cpu_baseline": cpu_base,
            "ms_text_head": ms_text, "ms_search": ms_search, "stage_ms": stage_ms, "search_fallbacks": int(nfb),
            "db_encode_cells_per_s": n_db / (enc_ms_max * 1e-3), "db_encode_ms": enc_ms_max, "db_encode_ms_first": enc_ms[0], "db_encode_ms_runs": enc_ms,
            "cold_db_qps": nq / ((ms_step + enc_ms_max) * 1e-3), "topk_matches_fp64_oracle_sample": parity,
            "reference_arm_note": "--impl reference runs the oracle PORT of the reference's Python (kind: port) on the host cores, sampled",
        }
        print(json.dumps(line), flush=True)
    graph = graph_out = None
    finish(world)


def finish(world: int):
    """Leave without the process-group teardown: with a captured NCCL graph alive destroy_process_group() did not return
    (one N=2 run sat until its timeout); every rank has printed and synchronised by now."""
    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_stream(eng, dev, n_cells: int, nq: int, obj_per_cell: int, steps: int, first_cell: int = 0, seed: int = 5):
    """BASELINE configs[3] shape (16 objects / cell, streamed encode + score with a running top-k) on this rank's cell range."""
    import torch

    import synth
    from text2loc_b200 import streaming

    Q = torch.from_numpy(synth.make_unit_rows(91, nq)).to(dev)
    streaming.stream_synthetic(eng, Q, K_TOP, seed, first_cell, min(n_cells, 2048), obj_per_cell)  # warm-up (arena, planes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        idx, score, nfb = streaming.stream_synthetic(eng, Q, K_TOP, seed, first_cell, n_cells, obj_per_cell)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"n_cells": n_cells, "n_queries": nq, "objects_per_cell": obj_per_cell, "ms": ms, "cells_per_s": n_cells / (ms * 1e-3),
            "queries_per_s": nq / (ms * 1e-3), "second_pass_queries": int(nfb),
            "what": "cells generated on the device chunk by chunk (t2l_synth_cells), encoded, scored, folded into a running top-k "
                    "(t2l_search_topk_accumulate); points and the full embedding matrix never resident"}, idx, score


def bench_fine(dev, n_queries: int, top_cells: int, n_db_cells: int):
    """BASELINE configs[4] shape: every query against its top-`top_cells` retrieved cells through the fine stage (CrossMatch)."""
    import torch

    import synth
    from text2loc_b200.engine import Engine

    eng = Engine(dev)
    eng.load_state_dict(synth.make_fine_state_dict(0))
    pad = 16
    pts, meta, ptr = eng.synth_cells(3, 0, n_db_cells, pad)
    t5 = torch.from_numpy(synth.make_t5_features(4, n_queries, N_SENT, N_TOK)).to(dev)
    rng = np.random.default_rng(0)
    pair_cell = torch.from_numpy(rng.integers(0, n_db_cells, n_queries * top_cells).astype(np.int32)).to(dev)
    pair_query = torch.arange(n_queries, dtype=torch.int32, device=dev).repeat_interleave(top_cells)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def run():
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        obj = eng.fine_encode_objects(pts, meta, ptr)
        e1.record()
        hints = eng.fine_encode_hints(t5)
        e2.record()
        off = eng.fine_match(obj, pair_cell, hints, pair_query, pad, N_SENT)
        e3.record()
        torch.cuda.synchronize()
        return off, e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)

    run()
    off, ms_obj, ms_hint, ms_match = run()
    total = ms_obj + ms_hint + ms_match
    return {"n_queries": n_queries, "top_cells": top_cells, "distinct_cells": n_db_cells, "objects_per_cell": pad,
            "ms_encode_objects": ms_obj, "ms_encode_hints": ms_hint, "ms_match_pairs": ms_match, "queries_per_s": n_queries / (total * 1e-3),
            "pairs_per_s_match_only": n_queries * top_cells / (ms_match * 1e-3), "offsets_finite": bool(torch.isfinite(off).all()),
            "what": "fine stage (CrossMatch) batched over queries: distinct cells' objects encoded once (d = 128), hints once, "
                    "all query x cell pairs through 2 x (cross_objects, cross_hints) decoder layers + offset MLP"}


def run_extra_workload(args):
    """--workload stream: BASELINE configs[3] (1M cells x 4 096 queries, 16 objects / cell, streamed encode + score; every rank
    streams its contiguous range of cells against ALL queries, then one packed all-gather + merge of the running top-k lists).
    --workload fine: configs[4] (32 768 queries x top-5 cells through the fine stage, one GPU)."""
    import torch
    import torch.distributed as dist

    import synth
    from text2loc_b200 import distributed as t2ld
    from text2loc_b200 import streaming
    from text2loc_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "fine":
        if rank == 0:
            nq = args.queries or 32768
            with ClockSampler([local_rank]) as clocks:
                r = bench_fine(dev, n_queries=nq, top_cells=5, n_db_cells=10000)
            print(json.dumps({"metric": "fine-stage queries/sec (CrossMatch offsets for the top-5 retrieved cells of every query)", "value": r["queries_per_s"],
                              "unit": "queries/s", "n_gpus": 1, "higher_is_better": True, "data": "synthetic", "dtype": "f16 token layer + PointNet++, 3xtf32 decoder layers",
                              "config": {"workload": "configs[4]: coarse->fine, top-5 retrieved cells into cross_matcher offset regression, "
                                                     f"{nq} queries, 1xB200"}, "detail": r, "clocks": clocks.summary()}), flush=True)
        finish(world)
        return

    n_cells = args.cells or 1000000 // world
    nq = args.queries or 4096
    per = 16
    eng = Engine(dev)
    eng.load_state_dict(synth.make_state_dict(0))
    Q = torch.from_numpy(synth.make_unit_rows(91, nq)).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(keep=False):
        out = streaming.stream_synthetic(eng, Q, K_TOP, 5, rank * n_cells, n_cells, per, chunk_cells=1024, keep_embeddings=keep)
        idx, score = out[0], out[1]
        if world > 1:
            packed = torch.stack([idx, score.view(torch.int64)])
            idx, score = eng.merge_topk_packed(t2ld.all_gather_rows(packed), nq, K_TOP)
        return idx, score, out

    step()  # warm-up: arena, planes, NCCL
    with ClockSampler(range(world), active=rank == 0) as clocks:
        barrier()
        a, b = ev(), ev()
        a.record()
        for _ in range(max(1, args.steps)):
            idx, score, _ = step()
        b.record()
        barrier()
    ms = a.elapsed_time(b) / max(1, args.steps)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # parity: streamed == unstreamed on this rank's kept embeddings; fp64 oracle on a sample (rank 0's shard)
    idx_l, score_l, nfb, D = streaming.stream_synthetic(eng, Q, K_TOP, 5, rank * n_cells, n_cells, per, chunk_cells=1024, keep_embeddings=True)
    eng.db_build(D, row_offset=rank * n_cells)
    idx_u, score_u, _ = eng.search_topk(Q, K_TOP)
    same = bool(torch.equal(idx_l, idx_u) and torch.equal(score_l, score_u))
    if rank == 0:
        from oracle import restate  # checker only

        oidx, _ = restate.search_topk(D.cpu().numpy(), Q[:16].cpu().numpy(), K_TOP)
        oracle_ok = bool((idx_l[:16].cpu().numpy() == oidx + rank * n_cells).all())
        total_cells = n_cells * world
        print(json.dumps({"metric": "streamed coarse-retrieval queries/sec over an N-cell DB that is never resident (encode + score)", "value": nq / (ms * 1e-3),
                          "unit": "queries/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                          "data": "synthetic", "cells_per_s": total_cells / (ms * 1e-3), "objects_per_s": total_cells * per / (ms * 1e-3),
                          "config": {"workload": f"configs[3]: {total_cells} cells x {nq} queries, {per} obj/cell, streamed encode+score, {world}xB200",
                                     "n_cells": total_cells, "n_queries": nq, "objects_per_cell": per, "chunk_cells": 1024, "k": K_TOP},
                          "streamed_equals_unstreamed_rank_shard": same, "topk_matches_fp64_oracle_sample_rank0_shard": oracle_ok,
                          "second_pass_queries_rank0": int(nfb), "clocks": clocks.summary(),
                          "hbm_note": "per object 6 172 B of generated points are written and read once (never leave the GPU); the stage stays "
                                      "tensor-bound (~53 kFLOP/B), see roofline_other_kernels.db_encode of the default run"}), flush=True)
    finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--graph", action="store_true",
                    help="also capture the resident step in ONE CUDA graph and report its replay time as `value` (measured: no gain at N=1, "
                         "1.7 %% at N=2 -- the step is GPU-bound, launches queue ahead -- so the default times the eager step)")
    ap.add_argument("--lanes", type=int, default=0, choices=[0, 1, 2],
                    help="query steps in flight per rank: 2 = consecutive steps alternate between two engines on two streams, so one "
                         "step's collectives and rank skew hide under the other's text head.  Measured at N = 8: 8.69 against 8.80 ms "
                         "per step (the other lane's long persistent GEMMs delay this lane's collectives as much as they fill its "
                         "waits), so the default stays 1")
    ap.add_argument("--skip-extras", action="store_true", help="leave out the small configs[3] / configs[4] / configs[0]-CPU extras")
    ap.add_argument("--workload", default="coarse", choices=["coarse", "stream", "fine"],
                    help="coarse = the headline step (default); stream = BASELINE configs[3] at full per-GPU size; fine = configs[4]")
    ap.add_argument("--cells", type=int, default=0, help="stream: cells per GPU (default 1 000 000 / gpus)")
    ap.add_argument("--queries", type=int, default=0, help="stream / fine: number of queries (defaults 4 096 / 32 768)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "coarse":
        run_engine_arm(args)
    else:
        run_extra_workload(args)


if __name__ == "__main__":
    main()
