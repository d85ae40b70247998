"""Reference checkpoint -> engine weights.

Takes a ``CellRetrievalNetwork.state_dict()`` (key layout: SURVEY.md Appendix B; the same keys
``training/coarse.py:327-332`` saves and ``evaluation/coarse.py:123`` loads with strict=False)
and produces the flat, BatchNorm-folded tensors the C ABI's ``t2l_set_weight`` expects.

Folding (eval mode, models/language_encoder.py:28-31):  y = (Wx + b - mu) / sqrt(var + 1e-5) * g + beta
  ->  W' = diag(g / sqrt(var + eps)) W,   b' = (b - mu) * g / sqrt(var + eps) + beta
computed in float64 and rounded once to float32.

The first Linear of each PointConv MLP is split by input columns ([x_j | pos_j - pos_i],
PyG PointConv.message): ``w1x`` acts on the point features (once per point), ``w1p`` on the
relative position (once per edge).
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5


def _np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def fold_linear_bn(sd, prefix):
    """<prefix>.0 = Linear, <prefix>.1 = BatchNorm1d  ->  (W', b') float32."""
    W = _np(sd[prefix + ".0.weight"]).astype(np.float64)
    b = _np(sd[prefix + ".0.bias"]).astype(np.float64)
    g = _np(sd[prefix + ".1.weight"]).astype(np.float64)
    beta = _np(sd[prefix + ".1.bias"]).astype(np.float64)
    mu = _np(sd[prefix + ".1.running_mean"]).astype(np.float64)
    var = _np(sd[prefix + ".1.running_var"]).astype(np.float64)
    s = g / np.sqrt(var + BN_EPS)
    return (W * s[:, None]).astype(np.float32), ((b - mu) * s + beta).astype(np.float32)


def _attn(out, sd, src, dst):
    f = lambda k: _np(sd[f"{src}.{k}"]).astype(np.float32)
    out[dst + ".in_w"] = f("self_attn.in_proj_weight")
    out[dst + ".in_b"] = f("self_attn.in_proj_bias")
    out[dst + ".out_w"] = f("self_attn.out_proj.weight")
    out[dst + ".out_b"] = f("self_attn.out_proj.bias")
    out[dst + ".l1_w"] = f("linear1.weight")
    out[dst + ".l1_b"] = f("linear1.bias")
    out[dst + ".l2_w"] = f("linear2.weight")
    out[dst + ".l2_b"] = f("linear2.bias")
    out[dst + ".n1_w"] = f("norm1.weight")
    out[dst + ".n1_b"] = f("norm1.bias")
    out[dst + ".n2_w"] = f("norm2.weight")
    out[dst + ".n2_b"] = f("norm2.bias")


def _decoder(out, sd, src, dst):
    """nn.TransformerDecoderLayer keys (models/cross_matcher.py:66-72) -> engine names.  The packed in_proj of the cross
    attention is split by rows: queries come from the target sequence, keys | values from the memory."""
    _attn(out, sd, src, dst)  # self_attn.*, linear1/2, norm1, norm2 (norm2 follows the cross attention)
    f = lambda k: _np(sd[f"{src}.{k}"]).astype(np.float32)
    w, b = f("multihead_attn.in_proj_weight"), f("multihead_attn.in_proj_bias")
    d = w.shape[1]
    out[dst + ".ca_q_w"], out[dst + ".ca_q_b"] = w[:d], b[:d]
    out[dst + ".ca_kv_w"], out[dst + ".ca_kv_b"] = w[d:], b[d:]
    out[dst + ".ca_out_w"], out[dst + ".ca_out_b"] = f("multihead_attn.out_proj.weight"), f("multihead_attn.out_proj.bias")
    out[dst + ".n3_w"], out[dst + ".n3_b"] = f("norm3.weight"), f("norm3.bias")


def is_fine_state_dict(sd: dict) -> bool:
    """CrossMatch.state_dict() (models/cross_matcher.py:39-78) carries the offset MLP; CellRetrievalNetwork's does not."""
    return "mlp_offsets.0.weight" in sd


def engine_weights(sd: dict) -> dict:
    """name -> float32 2-D array (biases and norm vectors as [1, n]).  Accepts the coarse model's state dict
    (CellRetrievalNetwork) or the fine stage's (CrossMatch)."""
    out = {}
    fine = is_fine_state_dict(sd)
    pn = "object_encoder.pointnet"
    for i, c_in in ((1, 3), (2, 64), (3, 128)):
        w1, b1 = fold_linear_bn(sd, f"{pn}.sa{i}.point_conv.local_nn.0")
        w2, b2 = fold_linear_bn(sd, f"{pn}.sa{i}.point_conv.local_nn.1")
        assert w1.shape[1] == c_in + 3
        out[f"sa{i}.w1x"], out[f"sa{i}.w1p"], out[f"sa{i}.b1"] = w1[:, :c_in], w1[:, c_in:], b1
        out[f"sa{i}.w2"], out[f"sa{i}.b2"] = w2, b2
        if i > 1:
            # per-point form of the first Linear: W1 [x_j | pos_j - pos_i] + b = (W1x x_j + W1p (pos_j - o) + b) - W1p (pos_i - o);
            # the position columns come twice because the engine feeds pos_j - o as a tf32 hi | lo pair (csrc/pointnet.cu)
            out[f"sa{i}.w1q"] = np.concatenate([w1[:, :c_in], w1[:, c_in:], w1[:, c_in:]], axis=1)
    out["ga.w1"], out["ga.b1"] = fold_linear_bn(sd, f"{pn}.ga.mlp.0")
    out["ga.w2"], out["ga.b2"] = fold_linear_bn(sd, f"{pn}.ga.mlp.1")
    for n in ("lin1", "lin2"):
        out[n + ".w"], out[n + ".b"] = _np(sd[f"{pn}.{n}.weight"]), _np(sd[f"{pn}.{n}.bias"])
    oe = "object_encoder"
    out["mlp_pointnet.w"], out["mlp_pointnet.b"] = fold_linear_bn(sd, f"{oe}.mlp_pointnet.0")
    for src, dst in (("color_encoder", "color"), ("pos_encoder", "pos"), ("num_encoder", "num")):
        out[dst + ".w1"], out[dst + ".b1"] = fold_linear_bn(sd, f"{oe}.{src}.0")
        out[dst + ".w2"], out[dst + ".b2"] = fold_linear_bn(sd, f"{oe}.{src}.1")
    out["merge.w"], out["merge.b"] = fold_linear_bn(sd, f"{oe}.mlp_merge.0")
    _attn(out, sd, "language_encoder.intra_module.0", "txt_intra")
    out["txt_mlp.w"], out["txt_mlp.b"] = fold_linear_bn(sd, "language_encoder.inter_mlp.0")
    if fine:
        n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("cross_hints."))
        if n_layers != 2 or not any(k.startswith("cross_objects.1.") for k in sd):
            raise ValueError("the B200 fine stage is built for fine_num_decoder_layers = 2 (evaluation/args.py default)")
        for i in range(2):
            _decoder(out, sd, f"cross_objects.{i}", f"cross_objects{i}")
            _decoder(out, sd, f"cross_hints.{i}", f"cross_hints{i}")
        out["offs.w1"], out["offs.b1"] = _np(sd["mlp_offsets.0.weight"]), _np(sd["mlp_offsets.0.bias"])
        out["offs.w2"], out["offs.b2"] = _np(sd["mlp_offsets.2.weight"]), _np(sd["mlp_offsets.2.bias"])
    else:
        _attn(out, sd, "obj_inter_module.0", "obj_attn0")
        _attn(out, sd, "obj_inter_module.1", "obj_attn1")
        _attn(out, sd, "language_encoder.inter_module.0", "txt_inter")
    return {k: np.ascontiguousarray(np.atleast_2d(np.asarray(v, dtype=np.float32))) for k, v in out.items()}


REQUIRED_PREFIXES = ("object_encoder.pointnet.sa1", "object_encoder.mlp_merge", "obj_inter_module.0",
                     "language_encoder.intra_module.0", "language_encoder.inter_mlp.0", "language_encoder.inter_module.0")
