"""GPU parity of the individual kernels through the C ABI: the tcgen05 tf32 GEMM and its fused
epilogues, the fp32 SIMT GEMM, FPS / ball query index sets (bit-exact vs the oracle)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def eng():
    from text2loc_b200.engine import Engine

    return Engine("cuda:0")


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_rna(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 256, 64), (1000, 512, 259), (4096, 1024, 1024), (96, 64, 32), (4224, 64, 32),
                                   (2000, 3072, 1024), (777, 1024, 4096)])
def test_umma_gemm_matches_tf32_emulation(eng, M, N, K):
    """kind::tf32 reads fp32 operands and ignores the low 13 mantissa bits: the result must equal
    an fp64 product of truncated operands up to fp32 accumulation error."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    ld = (K + 3) // 4 * 4
    A = torch.zeros(M, ld, device="cuda")
    W = torch.zeros(N, ld, device="cuda")
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g)
    W[:, :K] = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    C = eng.debug_linear(A[:, :K], W[:, :K], b, act=0, path=1)
    want_trunc = tf32_trunc(A).double() @ tf32_trunc(W).double().T + b.double()
    want_rna = tf32_rna(A).double() @ tf32_rna(W).double().T + b.double()
    e_trunc = (C.double() - want_trunc).abs().max().item()
    e_rna = (C.double() - want_rna).abs().max().item()
    exact = A.double() @ W.double().T + b.double()
    print(f"\n[{M}x{N}x{K}] |C-trunc|={e_trunc:.3e} |C-rna|={e_rna:.3e} |C-fp64|={(C.double() - exact).abs().max().item():.3e}")
    assert min(e_trunc, e_rna) < 2e-5 * max(1.0, K / 256) ** 0.5
    # pre-rounded operands (what the engine feeds): result independent of the hardware's rounding of raw fp32
    Cr = eng.debug_linear(tf32_rna(A)[:, :K], tf32_rna(W)[:, :K], b, act=1, path=1)
    assert (Cr.double() - want_rna.clamp(min=0)).abs().max().item() < 2e-5 * max(1.0, K / 256) ** 0.5


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 64), (1000, 512, 264), (4096, 1024, 1024), (96, 64, 32), (2000, 3072, 1024),
                                   (777, 1024, 4096), (32760, 4096, 1024)])
def test_umma_gemm_f16_operands(eng, M, N, K):
    """kind::f16 with fp16 operands and fp32 accumulation (the token layer): products of fp16 values are exact in
    fp32, so the result equals the fp64 product of the same fp16 operands up to fp32 accumulation error; the
    fp16 output variant is that, rounded to nearest fp16."""
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    b = torch.randn(N, device="cuda", generator=g)
    want = A.double() @ W.double().T + b.double()
    C = eng.debug_linear_f16(A, W, b, act=0)
    err = (C.double() - want).abs().max().item()
    print(f"\n[f16 {M}x{N}x{K}] |C-fp64|={err:.3e}")
    assert err < 2e-5 * max(1.0, K / 256) ** 0.5
    Ch = eng.debug_linear_f16(A, W, b, act=1, out_half=True)
    assert Ch.dtype == torch.float16
    assert (Ch.double() - want.clamp(min=0)).abs().max().item() < 2e-5 * max(1.0, K / 256) ** 0.5 + 2.0 ** -11 * want.abs().max().item()
    # saturation instead of inf
    big = eng.debug_linear_f16(A[:128] * 0 + 250.0, W * 0 + 2.0, None, act=0, out_half=True) if K >= 256 else None
    if big is not None:
        assert torch.isfinite(big).all() and big.max().item() == 65504.0


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 1024, 1024), (1000, 512, 264), (32760, 1024, 1024), (777, 1024, 4096), (129, 256, 128),
                                   (96, 256, 64), (40000, 256, 64)])
def test_umma_gemm_f16_residual_stream(eng, M, N, K):
    """Out-projection / FFN2 on the token layer's fp16 residual stream: fp16(A W^T + bias + R), the sum formed in fp32 and
    rounded once.  The TMA epilogue (residual slabs in, result slabs out, in place in shared memory) and the register-staged
    one must both equal the fp64 result rounded to fp16 up to fp32 accumulation error -- including ragged M (rows beyond M are
    zero-filled on the way in and clipped on the way out) and many tiles per CTA (slab barrier phases)."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    R = (3.0 * torch.randn(M, N, device="cuda", generator=g)).half()
    b = torch.randn(N, device="cuda", generator=g)
    want = A.double() @ W.double().T + b.double() + R.double()
    tol = 2e-5 * max(1.0, K / 256) ** 0.5 + 2.0 ** -11 * want.abs().max().item()
    for reg in (False, True):
        C = eng.debug_linear_f16_residual(A, W, b, R, reg_epilogue=reg)  # (the hook also checks that rows beyond M stay untouched)
        err = (C.double() - want).abs().max().item()
        print(f"\n[f16+residual {M}x{N}x{K} {'registers' if reg else 'tma'}] |C-fp64|={err:.3e} (tol {tol:.3e})")
        assert C.dtype == torch.float16 and err < tol
    # the two epilogues round the same fp32 sums: bit-identical
    assert torch.equal(eng.debug_linear_f16_residual(A, W, b, R, reg_epilogue=False), eng.debug_linear_f16_residual(A, W, b, R, reg_epilogue=True))


@pytest.mark.parametrize("n_seq,S,d,heads", [(2048, 28, 256, 4), (37, 17, 256, 4), (5, 32, 256, 4), (300, 16, 256, 4), (1000, 6, 256, 4),
                                             (64, 9, 128, 4), (1, 28, 256, 4), (500, 6, 128, 4), (333, 16, 128, 4), (7, 3, 128, 4), (1, 1, 128, 4),
                                             (40, 28, 128, 4)])
def test_small_attention_core(eng, n_seq, S, d, heads):
    """nn.MultiheadAttention's core without a mask (cell_retrieval.py:101-103, language_encoder.py:143-145) in fp32: the
    row-per-warp kernel (short sequences) and the sequence-per-warp kernel (64-wide heads, S > 16) against torch in fp64."""
    g = torch.Generator(device="cuda").manual_seed(n_seq * 3 + S)
    qkv = torch.randn(n_seq * S, 3 * d, device="cuda", generator=g)
    got = eng.debug_mha(qkv, n_seq, S, heads)
    q, k, v = (t.reshape(n_seq, S, heads, d // heads).permute(0, 2, 1, 3).double() for t in qkv.split(d, dim=1))
    att = torch.softmax(q @ k.transpose(-1, -2) / (d // heads) ** 0.5, dim=-1)
    want = (att @ v).permute(0, 2, 1, 3).reshape(n_seq * S, d)
    err = (got.double() - want).abs().max().item()
    print(f"\n[mha {n_seq}x{S}x{d}/{heads}] |out-fp64|={err:.3e}")
    assert err < 5e-6


@pytest.mark.parametrize("n_seq,Sq,Sk,d,heads", [(900, 16, 6, 128, 4), (900, 6, 16, 128, 4), (5, 3, 32, 128, 4), (77, 28, 9, 256, 4), (33, 16, 16, 128, 4),
                                                 (3, 40, 7, 128, 4)])
def test_small_cross_attention_core(eng, n_seq, Sq, Sk, d, heads):
    """nn.TransformerDecoderLayer's multihead_attn core (models/cross_matcher.py:113-115): queries from the target rows, keys /
    values from the memory rows, no mask; the packed short-sequence kernel (32-wide heads), the sequence-per-warp kernel and the
    row-per-warp kernel against torch in fp64."""
    g = torch.Generator(device="cuda").manual_seed(n_seq * 5 + Sq * 3 + Sk)
    q = torch.randn(n_seq * Sq, d, device="cuda", generator=g)
    kv = torch.randn(n_seq * Sk, 2 * d, device="cuda", generator=g)
    got = eng.debug_mha_cross(q, kv, n_seq, Sq, Sk, heads)
    hd = d // heads
    qq = q.reshape(n_seq, Sq, heads, hd).permute(0, 2, 1, 3).double()
    kk, vv = (t.reshape(n_seq, Sk, heads, hd).permute(0, 2, 1, 3).double() for t in kv.split(d, dim=1))
    want = (torch.softmax(qq @ kk.transpose(-1, -2) / hd ** 0.5, dim=-1) @ vv).permute(0, 2, 1, 3).reshape(n_seq * Sq, d)
    err = (got.double() - want).abs().max().item()
    print(f"\n[cross mha {n_seq}x{Sq}x{Sk}x{d}/{heads}] |out-fp64|={err:.3e}")
    assert err < 5e-6


def test_intra_cell_attention_without_duplicate_padding_rows(eng):
    """The reference zero-pads every cell to 28 object slots and attends without a mask (cell_retrieval.py:81-103): the
    28 - n padded slots of a cell are identical rows.  The engine carries one of them, weighted by its count as a key.
    Against torch on the full padded [B, 28, 3d] tensor (padded slots = copies of the representative row)."""
    slots, d, heads = 28, 256, 4
    counts = [8, 1, 28, 30, 27, 16, 2, 8, 8, 40, 5] * 20
    g = torch.Generator(device="cuda").manual_seed(5)
    rows = [min(n, slots) + (1 if n < slots else 0) for n in counts]
    row_ptr = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int32)
    cell_ptr = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32)
    qkv = torch.randn(int(row_ptr[-1]), 3 * d, device="cuda", generator=g)
    got = eng.debug_mha_cells(qkv, row_ptr, cell_ptr, slots, heads)
    worst = 0.0
    for b, n in enumerate(counts):
        r0, r = int(row_ptr[b]), rows[b]
        idx = list(range(r0, r0 + r)) + [r0 + r - 1] * (slots - r)  # expand the representative into the padded slots
        full = qkv[idx].double()
        q, k, v = (t.reshape(slots, heads, d // heads).permute(1, 0, 2) for t in full.split(d, dim=1))
        att = torch.softmax(q @ k.transpose(-1, -2) / (d // heads) ** 0.5, dim=-1)
        want = (att @ v).permute(1, 0, 2).reshape(slots, d)[:r]
        worst = max(worst, (got[r0:r0 + r].double() - want).abs().max().item())
    print(f"\n[intra-cell attention, ragged] |out-fp64|={worst:.3e}")
    assert worst < 5e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (4224 * 4, 64, 32), (2112 * 3, 128, 128), (1056, 256, 256), (64, 1024, 512)])
def test_umma_segmax_epilogue(eng, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = tf32_rna(torch.randn(M, K, device="cuda", generator=g))
    W = tf32_rna(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    b = torch.randn(N, device="cuda", generator=g)
    got = eng.debug_linear(A, W, b, act=1, segmax=True, path=1)
    want = (A.double() @ W.double().T + b.double()).clamp(min=0).view(M // 32, 32, N).max(dim=1)[0]
    assert got.shape == (M // 32, N)
    assert (got.double() - want).abs().max().item() < 2e-5
    simt = eng.debug_linear(A, W, b, act=1, segmax=True, path=0)
    assert (simt.double() - want).abs().max().item() < 2e-5


@pytest.mark.parametrize("M,N,K", [(77, 50, 3), (1000, 32, 3), (5, 64, 1), (513, 256, 64), (28 * 9, 768, 256), (300, 256, 1024)])
def test_simt_gemm_fp32(eng, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    got = eng.debug_linear(A, W, b, act=1, path=0)
    want = (A.double() @ W.double().T + b.double()).clamp(min=0)
    assert (got.double() - want).abs().max().item() < 1e-5 * K ** 0.5 * 4


def test_fps_and_ball_query_bit_exact(eng, state_dict, golden):
    """Index sets are discrete: they must equal the oracle's exactly (golden fps*, digest of nbr*)."""
    from oracle import restate

    g = golden("cells_small.npz")
    pts = torch.from_numpy(g["pts"])
    d = eng.encode_objects_debug(pts, g["cell_ptr"]) if eng.has_weights else None
    if d is None:
        eng.load_state_dict(state_dict)
        d = eng.encode_objects_debug(pts, g["cell_ptr"])
    torch.cuda.synchronize()
    for i in (1, 2, 3):
        assert (d[f"fps{i}"].cpu().numpy().astype(np.int16) == g[f"fps{i}"]).all(), f"fps level {i}"
    _, aux = restate.pointnet2_features2(state_dict, pts, g["cell_ptr"], return_aux=True)
    for i in (1, 2, 3):
        want = aux[f"nbr{i}"].numpy()
        cnt = (want >= 0).sum(axis=2)
        assert (d[f"cnt{i}"].cpu().numpy() == cnt).all(), f"neighbour counts level {i}"
        got = d[f"nbr{i}"].cpu().numpy().astype(np.int64)
        mask = want >= 0
        assert (got[mask] == want[mask]).all(), f"neighbour lists level {i}"
    assert int((aux["nbr1"] >= 0).sum()) == int(g["nbr_count"][0])


def test_features2_within_tolerance(eng, state_dict, golden):
    g = golden("cells_small.npz")
    if not eng.has_weights:
        eng.load_state_dict(state_dict)
    d = eng.encode_objects_debug(torch.from_numpy(g["pts"]), g["cell_ptr"])
    got = d["features2"].cpu().numpy()
    want = g["features2"]
    rel = np.abs(got - want).max() / np.abs(want).max()
    print(f"\nfeatures2 max rel-to-max error {rel:.3e}")
    assert rel < 1e-3


def test_fps_ball_query_fma_switch(state_dict, golden, monkeypatch):
    """T2L_DIST_FMA=1 / pyg_ops.DIST_FMA evaluate squared distances with fused multiply-adds (what nvcc's default
    contraction would make of torch-cluster's `dist += tmp * tmp`): engine and oracle must still agree bit for bit."""
    from oracle import pyg_ops, restate
    from text2loc_b200.engine import Engine

    g = golden("cells_small.npz")
    pts = torch.from_numpy(g["pts"])
    monkeypatch.setenv("T2L_DIST_FMA", "1")
    monkeypatch.setattr(pyg_ops, "DIST_FMA", True)
    e = Engine("cuda:0")
    e.load_state_dict(state_dict)
    d = e.encode_objects_debug(pts, g["cell_ptr"])
    _, aux = restate.pointnet2_features2(state_dict, pts, g["cell_ptr"], return_aux=True)
    n_diff_from_default = 0
    for i in (1, 2, 3):
        assert (d[f"fps{i}"].cpu().numpy().astype(np.int64) == aux[f"fps{i}"].numpy()).all(), f"fps level {i} (fma)"
        n_diff_from_default += int((aux[f"fps{i}"].numpy().astype(np.int16) != g[f"fps{i}"]).sum())
        want = aux[f"nbr{i}"].numpy()
        mask = want >= 0
        assert (d[f"nbr{i}"].cpu().numpy().astype(np.int64)[mask] == want[mask]).all(), f"neighbour lists level {i} (fma)"
    print(f"\nFPS picks that differ between the fused and the unfused distance: {n_diff_from_default}")


def test_sa_empty_neighbour_slots(state_dict):
    """Objects whose points are far apart leave most of the 32 neighbour slots empty (down to the centroid alone): the
    fused kernel (sa_obj2.cu) replicates slot 0 there, the oracle masks the slots; the max must agree."""
    from oracle import restate
    import synth
    from text2loc_b200.engine import Engine

    e = Engine("cuda:0")
    e.load_state_dict(state_dict)
    cells = synth.make_cell_objects(11, 3, [2, 3, 1], max_raw=300)
    pts, meta, ptr = synth.pack_cells(cells, 11)
    rng = np.random.default_rng(5)
    pts = pts.copy()
    pts[:, :, 0:3] = rng.uniform(0.0, 3.0, size=pts[:, :, 0:3].shape).astype(np.float32)  # radius 0.2-0.4 balls are nearly empty
    d = e.encode_objects_debug(torch.from_numpy(pts), ptr)
    want, aux = restate.pointnet2_features2(state_dict, torch.from_numpy(pts), ptr, return_aux=True)
    cnt1 = (aux["nbr1"].numpy() >= 0).sum(axis=2)
    assert cnt1.min() < 4 and (d["cnt1"].cpu().numpy() == cnt1).all()
    got = d["features2"].cpu().numpy()
    want = want.numpy()
    rel = np.abs(got - want).max() / np.abs(want).max()
    print(f"\nsparse objects: features2 max rel-to-max error {rel:.3e}, min neighbours {cnt1.min()}")
    assert rel < 1e-3
