"""state_dict (reference key names) -> kernel-layout device tensors for the C engine.

One-time plumbing: BatchNorm folding, OIHW -> [O, KH*KW*I] (K-major, tap-major / channel-minor,
the order gemm_tc.cu walks K in), 16-bit casts, Q/K/V stacking, and the two column permutations
that let the kernels consume NHWC "cell-major" features where the reference flattens NCHW
"channel-major" ones (seq2seq_highlevel_cma.py:92-100 depth_linear; resnet_encoders.py:58-62
visual_fc).  On a CUDA device every conv / linear weight is packed by ONE launch of the
library's own ``rvb_pack_weight`` kernel (csrc/prep.cu); the torch expressions below are the
layout specification that kernel is tested against (and what the CPU unit tests exercise).

Engine tensor names (``ns`` is "hi" or "lo"):
  {ns}.rgb.stem.{w,b}                     h16 [64,256] (4 K blocks x two 32-wide filter rows), f32 [64]
  {ns}.rgb.l{1-4}.{blk}.{c1,c2,c3,ds}.{w,b}
  {ns}.depth.stem.w f32 [32,49]; {ns}.depth.stem.gn.{w,b}
  {ns}.depth.l{1-4}.{blk}.{c1,c2,c3,ds}.w ; .gn{1,2,3}.{w,b} ; .dsgn.{w,b}
  {ns}.depth.comp.w ; {ns}.depth.comp.gn.{w,b}
  hi.bert.{word,pos,type0,emb_ln.w,emb_ln.b} ; hi.bert.{i}.{qkv,ao,ff1,ff2}.{w,b} ; .ln{1,2}.{w,b}
  hi.{rgb_emb,depth_emb} ; hi.{rgb_kv,depth_kv,rgb_linear,depth_linear}.{w,b}
  hi.vla.{ins_fc,vis_fc,fc_q,fc_kv,fc_o,fc1,fc2}.{w,b} ; hi.vla.ln{0,1,2}.{w,b}
  hi.lstm.{wih,whh,b} ; hi.linear.{w,b}
  lo.{depth_fc,rgb_fc}.{w,b} ; lo.sub_emb ; lo.lstm.{wih,whh,b} ; lo.linear.{w,b} ; lo.stop.{w,b}
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

_STAGES = ((3, 1), (4, 2), (6, 2), (3, 2))


# 16-bit type of the engine the tensors are prepared for (set by the runtime before packing):
# torch.float16 for librobovln_b200.so, torch.bfloat16 for librobovln_b200_bf16.so
H16 = {"dtype": torch.float16}


def set_h16(dtype_name: str) -> None:
    H16["dtype"] = {"fp16": torch.float16, "bf16": torch.bfloat16}[dtype_name]


def _bf(t: torch.Tensor, dev) -> torch.Tensor:
    """fp32 -> the engine's 16-bit operand type (fp16 saturates instead of overflowing)."""
    t = t.detach().to(device=dev, dtype=torch.float32)
    if H16["dtype"] == torch.float16:
        t = t.clamp(-65504.0, 65504.0)
    return t.to(H16["dtype"]).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pack(w: torch.Tensor, dev, bn=None, out: torch.Tensor = None, eps: float = 1e-5):
    """[O,I,KH,KW] / [O,K] fp32 parameter (+ optional eval-mode BatchNorm (gamma, beta, mean, var)) ->
    (16-bit [O, KH*KW*I] K-major weight, fp32 bias or None).  ``out`` (a row slice of a preallocated stacked
    matrix) receives the weight in place."""
    w4 = w.detach()
    if w4.dim() == 2:
        w4 = w4[:, :, None, None]
    O, I, KH, KW = w4.shape
    if torch.device(dev).type != "cuda":
        if bn is not None:
            w4, b = fold_bn(w4, *bn, eps=eps)
        wk = _bf(_conv_kmajor(w4.float()), dev)
        if out is not None:
            out.copy_(wk)
            wk = out
        return wk, (_f32(b, dev) if bn is not None else None)
    from . import _lib

    lib = _lib.load(dtype="fp16" if H16["dtype"] == torch.float16 else "bf16")
    w4 = _f32(w4, dev)
    if out is None:
        out = torch.empty((O, KH * KW * I), dtype=H16["dtype"], device=dev)
    assert out.shape == (O, KH * KW * I) and out.stride(1) == 1
    bias = torch.empty((O,), dtype=torch.float32, device=dev) if bn is not None else None
    g, b, m, v = (_f32(t, dev) for t in bn) if bn is not None else (None, None, None, None)
    p = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(lib.rvb_pack_weight(p(w4), p(g), p(b), p(m), p(v), eps, p(out), p(bias), O, I, KH, KW, out.stride(0),
                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rvb_pack_weight", lib)
    return out, bias


def _conv_kmajor(w: torch.Tensor) -> torch.Tensor:
    """OIHW -> [O, KH*KW*I] with k = (r*KW + s)*I + c."""
    o, i, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(o, kh * kw * i)


def fold_bn(conv_w: torch.Tensor, g, b, mean, var, eps: float = 1e-5):
    """Eval-mode BatchNorm folded into the preceding bias-free conv."""
    scale = g.float() / torch.sqrt(var.float() + eps)
    return conv_w.float() * scale.view(-1, 1, 1, 1), b.float() - mean.float() * scale


def stem_packed_weights(w: torch.Tensor) -> torch.Tensor:
    """[64,3,7,7] (O,C,R,S) -> [64, 4*64]: K block kb holds filter rows 2kb and 2kb+1 as the packed window-mode
    A operand reads the padded row-pair-interleaved image: k = kb*64 + s*8 + (r%2)*4 + c for r < 7, s < 7,
    c < 3; zeros elsewhere (4th channel, 8th pixel of the window, the 8th filter row)."""
    o = w.shape[0]
    wk = torch.zeros((o, 4, 8, 2, 4), dtype=torch.float32, device=w.device)   # [O, kb, s(8), r2, c(4)]
    wr = torch.zeros((o, 8, 8, 4), dtype=torch.float32, device=w.device)      # [O, r(8), s(8), c(4)]
    wr[:, :7, :7, :3] = w.float().permute(0, 2, 3, 1)                          # [O, R, S, C]
    wk[:] = wr.view(o, 4, 2, 8, 4).permute(0, 1, 3, 2, 4)
    return wk.reshape(o, 4 * 64)


def stem_window_weights(w: torch.Tensor) -> torch.Tensor:
    """[64,3,7,7] (O,C,R,S) -> [64, 7*64]: one 64-wide K block per filter row r, laid out as the
    window-mode A operand reads the padded NHW8 image: k = s*8 + c for s < 7, c < 3, zeros
    elsewhere (channels 3..7 and the 8th pixel of the 64-element window)."""
    o = w.shape[0]
    wk = torch.zeros((o, 7, 8, 8), dtype=torch.float32, device=w.device)     # [O, r, s(8), c(8)]
    wk[:, :, :7, :3] = w.float().permute(0, 2, 3, 1)                          # [O, R, S, C]
    return wk.reshape(o, 7 * 64)


def prep_rgb_trunk(sd: Dict[str, torch.Tensor], ns: str, dev) -> Dict[str, torch.Tensor]:
    p = "rgb_encoder.cnn."
    out = {}

    def bn(prefix):
        return (sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"], sd[prefix + ".running_var"])

    w, b = fold_bn(sd[p + "conv1.weight"], *bn(p + "bn1"))
    out[f"{ns}.rgb.stem.w"] = _bf(stem_packed_weights(w), dev)
    out[f"{ns}.rgb.stem.b"] = _f32(b, dev)
    for li, (nb, _s) in enumerate(_STAGES):
        for blk in range(nb):
            q = f"{p}layer{li + 1}.{blk}."
            e = f"{ns}.rgb.l{li + 1}.{blk}."
            for ci in (1, 2, 3):
                out[e + f"c{ci}.w"], out[e + f"c{ci}.b"] = _pack(sd[q + f"conv{ci}.weight"], dev, bn(q + f"bn{ci}"))
            if blk == 0:
                out[e + "ds.w"], out[e + "ds.b"] = _pack(sd[q + "downsample.0.weight"], dev, bn(q + "downsample.1"))
    return out


def prep_depth_trunk(sd: Dict[str, torch.Tensor], ns: str, dev) -> Dict[str, torch.Tensor]:
    p = "depth_encoder.visual_encoder."
    b = p + "backbone."
    out = {}
    out[f"{ns}.depth.stem.w"] = _f32(sd[b + "conv1.0.weight"].reshape(32, 49), dev)
    out[f"{ns}.depth.stem.gn.w"] = _f32(sd[b + "conv1.1.weight"], dev)
    out[f"{ns}.depth.stem.gn.b"] = _f32(sd[b + "conv1.1.bias"], dev)
    for li, (nb, _s) in enumerate(_STAGES):
        for blk in range(nb):
            q = f"{b}layer{li + 1}.{blk}."
            e = f"{ns}.depth.l{li + 1}.{blk}."
            for ci, (cw, gn) in enumerate(((0, 1), (3, 4), (6, 7)), start=1):
                out[e + f"c{ci}.w"] = _pack(sd[q + f"convs.{cw}.weight"], dev)[0]
                out[e + f"gn{ci}.w"] = _f32(sd[q + f"convs.{gn}.weight"], dev)
                out[e + f"gn{ci}.b"] = _f32(sd[q + f"convs.{gn}.bias"], dev)
            if blk == 0:
                out[e + "ds.w"] = _pack(sd[q + "downsample.0.weight"], dev)[0]
                out[e + "dsgn.w"] = _f32(sd[q + "downsample.1.weight"], dev)
                out[e + "dsgn.b"] = _f32(sd[q + "downsample.1.bias"], dev)
    out[f"{ns}.depth.comp.w"] = _pack(sd[p + "compression.0.weight"], dev)[0]
    out[f"{ns}.depth.comp.gn.w"] = _f32(sd[p + "compression.1.weight"], dev)
    out[f"{ns}.depth.comp.gn.b"] = _f32(sd[p + "compression.1.bias"], dev)
    return out


def prep_bert(sd: Dict[str, torch.Tensor], dev) -> Dict[str, torch.Tensor]:
    p = "embedding_layer."
    e = p + "embeddings."
    out = {
        "hi.bert.word": _f32(sd[e + "word_embeddings.weight"], dev),
        "hi.bert.pos": _f32(sd[e + "position_embeddings.weight"], dev),
        "hi.bert.type0": _f32(sd[e + "token_type_embeddings.weight"][0], dev),
        "hi.bert.emb_ln.w": _f32(sd[e + "LayerNorm.weight"], dev),
        "hi.bert.emb_ln.b": _f32(sd[e + "LayerNorm.bias"], dev),
    }
    for i in range(12):
        q = f"{p}encoder.layer.{i}."
        a = q + "attention.self."
        n = f"hi.bert.{i}."
        qkv = torch.empty((2304, 768), dtype=H16["dtype"], device=dev)
        for j, nm in enumerate(("query", "key", "value")):
            _pack(sd[a + nm + ".weight"], dev, out=qkv[768 * j:768 * (j + 1)])
        out[n + "qkv.w"] = qkv
        out[n + "qkv.b"] = _f32(torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], 0), dev)
        out[n + "ao.w"] = _pack(sd[q + "attention.output.dense.weight"], dev)[0]
        out[n + "ao.b"] = _f32(sd[q + "attention.output.dense.bias"], dev)
        out[n + "ln1.w"] = _f32(sd[q + "attention.output.LayerNorm.weight"], dev)
        out[n + "ln1.b"] = _f32(sd[q + "attention.output.LayerNorm.bias"], dev)
        out[n + "ff1.w"] = _pack(sd[q + "intermediate.dense.weight"], dev)[0]
        out[n + "ff1.b"] = _f32(sd[q + "intermediate.dense.bias"], dev)
        out[n + "ff2.w"] = _pack(sd[q + "output.dense.weight"], dev)[0]
        out[n + "ff2.b"] = _f32(sd[q + "output.dense.bias"], dev)
        out[n + "ln2.w"] = _f32(sd[q + "output.LayerNorm.weight"], dev)
        out[n + "ln2.b"] = _f32(sd[q + "output.LayerNorm.bias"], dev)
    return out


def prep_hi_tail(sd: Dict[str, torch.Tensor], dev) -> Dict[str, torch.Tensor]:
    out = {
        "hi.rgb_emb": _f32(sd["rgb_encoder.spatial_embeddings.weight"], dev),
        "hi.depth_emb": _f32(sd["depth_encoder.spatial_embeddings.weight"], dev),
        "hi.rgb_kv.w": _pack(sd["rgb_kv.weight"][:, :, 0], dev)[0],
        "hi.rgb_kv.b": _f32(sd["rgb_kv.bias"], dev),
        "hi.depth_kv.w": _pack(sd["depth_kv.weight"][:, :, 0], dev)[0],
        "hi.depth_kv.b": _f32(sd["depth_kv.bias"], dev),
        "hi.rgb_linear.w": _pack(sd["rgb_linear.2.weight"], dev)[0],
        "hi.rgb_linear.b": _f32(sd["rgb_linear.2.bias"], dev),
        # reference flattens [B,192,16] channel-major (index c*16+cell); tokens are [B,16,192]
        "hi.depth_linear.w": _bf(sd["depth_linear.1.weight"].reshape(128, 192, 16).permute(0, 2, 1).reshape(128, 3072), dev),
        "hi.depth_linear.b": _f32(sd["depth_linear.1.bias"], dev),
    }
    v = "image_cm_encoder."
    a = v + "layers.0.enc_att.attention."
    f = v + "layers.0.pwff."
    pairs = {
        "ins_fc": v + "ins_fc", "vis_fc": v + "vis_fc", "fc_q": a + "fc_q", "fc_o": a + "fc_o",
        "fc1": f + "fc1", "fc2": f + "fc2",
    }
    for n, k in pairs.items():
        out[f"hi.vla.{n}.w"] = _pack(sd[k + ".weight"], dev)[0]
        out[f"hi.vla.{n}.b"] = _f32(sd[k + ".bias"], dev)
    kv = torch.empty((512, 256), dtype=H16["dtype"], device=dev)
    _pack(sd[a + "fc_k.weight"], dev, out=kv[:256])
    _pack(sd[a + "fc_v.weight"], dev, out=kv[256:])
    out["hi.vla.fc_kv.w"] = kv
    # fused block (csrc/vla_block.cu): fc_q folded into the key side.  Per visual cell x (a row of LN0(relu(vis_fc(.)))):
    #   K'_h = Wq_h^T (Wk_h x + bk_h)   [256]   so that   S_h = q_h . k_h = Q0 . K'_h + c_h,   c_h = bq_h . (Wk_h x + bk_h)
    # kvx = x W_all^T + b_all = [K'_0 | K'_1 | K'_2 | K'_3 | c_0..c_3, 0, 0, 0, 0 | v]   (1024 + 8 + 256 = 1288 columns),
    # composed in fp32 and rounded to 16 bits once.
    wq, bq = sd[a + "fc_q.weight"].detach().float(), sd[a + "fc_q.bias"].detach().float()
    wk, bk = sd[a + "fc_k.weight"].detach().float(), sd[a + "fc_k.bias"].detach().float()
    wv, bv = sd[a + "fc_v.weight"].detach().float(), sd[a + "fc_v.bias"].detach().float()
    w_rows, b_rows, c_rows, c_bias = [], [], [], []
    for h in range(4):
        sl = slice(64 * h, 64 * (h + 1))
        w_rows.append(wq[sl].t() @ wk[sl])
        b_rows.append(wq[sl].t() @ bk[sl])
        c_rows.append(wk[sl].t() @ bq[sl])
        c_bias.append((bq[sl] * bk[sl]).sum())
    zeros = torch.zeros((4, 256), dtype=torch.float32, device=wq.device)
    w_all = torch.cat(w_rows + [torch.stack(c_rows), zeros, wv], 0)
    b_all = torch.cat(b_rows + [torch.stack(c_bias), zeros[:, 0], bv], 0)
    out["hi.vla.kvx.w"] = _bf(w_all, dev)
    out["hi.vla.kvx.b"] = _f32(b_all, dev)
    out["hi.vla.fc_kv.b"] = _f32(torch.cat([sd[a + "fc_k.bias"], sd[a + "fc_v.bias"]], 0), dev)
    for n, k in (("ln0", v + "layer_norm"), ("ln1", v + "layers.0.enc_att.layer_norm"), ("ln2", f + "layer_norm")):
        out[f"hi.vla.{n}.w"] = _f32(sd[k + ".weight"], dev)
        out[f"hi.vla.{n}.b"] = _f32(sd[k + ".bias"], dev)
    s = "state_encoder.rnn."
    out["hi.lstm.wih"] = _pack(sd[s + "weight_ih_l0"], dev)[0]
    out["hi.lstm.whh"] = _pack(sd[s + "weight_hh_l0"], dev)[0]
    out["hi.lstm.b"] = _f32(sd[s + "bias_ih_l0"].float() + sd[s + "bias_hh_l0"].float(), dev)
    out["hi.linear.w"] = _f32(sd["linear.weight"], dev)
    out["hi.linear.b"] = _f32(sd["linear.bias"], dev)
    return out


def prep_lo_tail(sd: Dict[str, torch.Tensor], dev) -> Dict[str, torch.Tensor]:
    # visual_fc consumes Flatten([B,128,4,4]) (index c*16+cell); the engine feeds the
    # [B,16,192] depth tokens (128 features + 64 spatial-embedding columns) -> permute to
    # cell-major and put zeros under the embedding columns.
    w = sd["depth_encoder.visual_fc.1.weight"].float().reshape(128, 128, 16).permute(0, 2, 1)   # [o, cell, c]
    wp = torch.zeros((128, 16, 192), dtype=torch.float32, device=w.device)
    wp[:, :, :128] = w
    out = {
        "lo.depth_fc.w": _bf(wp.reshape(128, 3072), dev),
        "lo.depth_fc.b": _f32(sd["depth_encoder.visual_fc.1.bias"], dev),
        "lo.rgb_fc.w": _pack(sd["rgb_encoder.fc.weight"], dev)[0],
        "lo.rgb_fc.b": _f32(sd["rgb_encoder.fc.bias"], dev),
        "lo.sub_emb": _f32(sd["sub_task_embedding.weight"], dev),
    }
    s = "state_encoder.rnn."
    out["lo.lstm.wih"] = _pack(sd[s + "weight_ih_l0"], dev)[0]
    out["lo.lstm.whh"] = _pack(sd[s + "weight_hh_l0"], dev)[0]
    out["lo.lstm.b"] = _f32(sd[s + "bias_ih_l0"].float() + sd[s + "bias_hh_l0"].float(), dev)
    out["lo.linear.w"] = _f32(sd["linear.weight"], dev)
    out["lo.linear.b"] = _f32(sd["linear.bias"], dev)
    out["lo.stop.w"] = _f32(sd["stop_linear.weight"], dev)
    out["lo.stop.b"] = _f32(sd["stop_linear.bias"], dev)
    return out


TRUNK_PREFIXES = ("rgb_encoder.cnn.", "depth_encoder.visual_encoder.")


def trunks_identical(sd_a: Dict[str, torch.Tensor], sd_b: Dict[str, torch.Tensor]) -> bool:
    """True iff every frozen-trunk tensor of the two models is bit-identical (dedup legality,
    SURVEY.md 7.2).  ``cnn.fc`` (lo only, unused) is ignored.  On a CUDA device all pairs are compared by one
    launch of ``rvb_compare_many`` (csrc/prep.cu)."""
    keys_a = [k for k in sd_a if k.startswith(TRUNK_PREFIXES) and not k.startswith("rgb_encoder.cnn.fc.")]
    pairs = []
    for k in keys_a:
        if k not in sd_b:
            return False
        a, b = sd_a[k], sd_b[k]
        if a.shape != b.shape or a.dtype != b.dtype:
            return False
        pairs.append((a.detach(), b.detach()))
    if not pairs:
        return True
    dev = pairs[0][0].device
    on_gpu = dev.type == "cuda" and all(a.device == dev and b.device == dev and a.is_contiguous() and b.is_contiguous()
                                        and (a.numel() * a.element_size()) % 4 == 0 for a, b in pairs)
    if not on_gpu:
        return all(torch.equal(a.to(b.device), b) for a, b in pairs)
    from . import _lib

    lib = _lib.load(dtype="fp16" if H16["dtype"] == torch.float16 else "bf16")
    pairs = [(a, b) for a, b in pairs if a.numel() > 0]
    meta = torch.tensor([[a.data_ptr() for a, _ in pairs], [b.data_ptr() for _, b in pairs],
                         [a.numel() * a.element_size() // 4 for a, _ in pairs]], dtype=torch.int64).to(dev)
    flag = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rvb_compare_many(ctypes.c_void_p(meta[0].data_ptr()), ctypes.c_void_p(meta[1].data_ptr()),
                                        ctypes.c_void_p(meta[2].data_ptr()), len(pairs), ctypes.c_void_p(flag.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "rvb_compare_many", lib)
    return int(flag.item()) == 0
