"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of robo-vln's HCM policy forward pass.

This is the parity oracle for the CUDA path in ``robo-vln_b200/``.  It restates, as plain
functional PyTorch on CPU tensors, the arithmetic of

  * ``Seq2Seq_HighLevel_CMA.forward``  robo_vln_baselines/models/seq2seq_highlevel_cma.py:170-233
  * ``Seq2Seq_LowLevel.forward``       robo_vln_baselines/models/seq2seq_lowlevel.py:116-162
  * ``TorchVisionResNet50.forward``    robo_vln_baselines/models/encoders/resnet_encoders.py:189-237
  * ``VlnResnetDepthEncoder.forward``  robo_vln_baselines/models/encoders/resnet_encoders.py:76-108
  * ``ResNetEncoder.forward`` / ``ResNet`` / ``Bottleneck``
        environments/habitat-lab/habitat_baselines/rl/ddppo/policy/resnet_policy.py:160-189,
        .../resnet.py:65-77,104-137,181-253
  * ``Visual_Ling_Attn.forward`` and its sub-layers
        robo_vln_baselines/models/transformer/transformer.py:38-43,81-126,209-221,262-281
  * ``sinusoid_encoding_table``        robo_vln_baselines/common/utils.py:167-185
  * ``RNNStateEncoder``                environments/habitat-lab/habitat_baselines/rl/models/rnn_state_encoder.py:74-142

and of the third-party modules the reference calls but does not vendor:

  * ``torchvision.models.resnet50`` (v1.5 bottleneck: stride on the 3x3; reference pins
    torchvision==0.2.2.post3 in requirements.txt:9 -- same architecture),
  * ``transformers.BertModel`` (bert-base-uncased architecture, unpinned in
    requirements.txt:14; call site seq2seq_highlevel_cma.py:45,192-195: ``input_ids`` only,
    so token_type 0, positions 0..L-1 and an all-ones attention mask),
  * ``torch.nn.LSTM`` (gate order i,f,g,o).

Pinning: the reference ships no test or golden vector for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``oracle/make_golden.py`` loads the synthetic weights of ``oracle/weights.py`` into the
unmodified reference modules (via ``oracle/ref_loader.py``) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product path never does.

All functions take a ``state_dict`` with the reference's own key names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# RGB trunk: torchvision ResNet-50 v1.5, eval-mode BatchNorm (resnet_encoders.py:143-149)
# --------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=1e-5)


_RGB_STAGES = ((1, 3, 1), (2, 4, 2), (3, 6, 2), (4, 3, 2))   # (layer idx, blocks, first stride)


def rgb_trunk(sd: SD, prefix: str, rgb_nhwc: torch.Tensor) -> torch.Tensor:
    """[B,H,W,3] float 0..255 -> layer4 output [B,2048,H/32,W/32].

    resnet_encoders.py:211-214: permute to NCHW and divide by 255 (no mean/std)."""
    p = prefix + "cnn."
    x = rgb_nhwc.permute(0, 3, 1, 2) / 255.0
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(sd, p + "bn1", x))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nblocks, stride in _RGB_STAGES:
        for b in range(nblocks):
            q = f"{p}layer{li}.{b}."
            s = stride if b == 0 else 1
            idt = x
            o = F.relu(_bn(sd, q + "bn1", F.conv2d(x, sd[q + "conv1.weight"])))
            o = F.relu(_bn(sd, q + "bn2", F.conv2d(o, sd[q + "conv2.weight"], stride=s, padding=1)))
            o = _bn(sd, q + "bn3", F.conv2d(o, sd[q + "conv3.weight"]))
            if (q + "downsample.0.weight") in sd:
                idt = _bn(sd, q + "downsample.1", F.conv2d(x, sd[q + "downsample.0.weight"], stride=s))
            x = F.relu(o + idt)
    return x


def spatial_embedding_channels(weight: torch.Tensor) -> torch.Tensor:
    """resnet_encoders.py:91-102,219-229: ``embedding(arange(16)).view(1,-1,4,4)`` -- a raw
    reinterpretation of the [16,64] table as [64,4,4] (channel c, cell k reads flat[c*16+k])."""
    return weight.reshape(1, -1, 4, 4)


def rgb_encoder_hi(sd: SD, rgb: torch.Tensor) -> torch.Tensor:
    """hi: spatial_output=True -> [B,2112,4,4] (resnet_encoders.py:160-175,216-231)."""
    x = rgb_trunk(sd, "rgb_encoder.", rgb)
    x = F.adaptive_avg_pool2d(x, (4, 4))
    e = spatial_embedding_channels(sd["rgb_encoder.spatial_embeddings.weight"]).expand(x.shape[0], -1, -1, -1)
    return torch.cat([x, e], dim=1)


def rgb_encoder_lo(sd: SD, rgb: torch.Tensor) -> torch.Tensor:
    """lo: global avgpool -> fc 2048->256 -> ReLU (resnet_encoders.py:154-157,235-237)."""
    x = rgb_trunk(sd, "rgb_encoder.", rgb)
    x = torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)
    return F.relu(F.linear(x, sd["rgb_encoder.fc.weight"], sd["rgb_encoder.fc.bias"]))


# --------------------------------------------------------------------------------------
# Depth trunk: DDPPO ResNet-50, base planes 32, GroupNorm(16) (resnet.py, resnet_policy.py)
# --------------------------------------------------------------------------------------
_DEPTH_STAGES = ((1, 3, 1), (2, 4, 2), (3, 6, 2), (4, 3, 2))
NGROUPS = 16


def depth_trunk(sd: SD, prefix: str, depth_nhwc: torch.Tensor) -> torch.Tensor:
    """[B,256,256,1] -> [B,128,4,4].  resnet_policy.py:160-189 (ResizeCenterCropper(256) is
    the identity for 256x256 input, habitat_baselines/common/utils.py:95-105)."""
    p = prefix + "visual_encoder."
    x = depth_nhwc.permute(0, 3, 1, 2)
    x = F.avg_pool2d(x, 2)
    b = p + "backbone."
    x = F.conv2d(x, sd[b + "conv1.0.weight"], None, stride=2, padding=3)
    x = F.relu(F.group_norm(x, NGROUPS, sd[b + "conv1.1.weight"], sd[b + "conv1.1.bias"], eps=1e-5))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nblocks, stride in _DEPTH_STAGES:
        for blk in range(nblocks):
            q = f"{b}layer{li}.{blk}."
            s = stride if blk == 0 else 1
            idt = x
            o = F.conv2d(x, sd[q + "convs.0.weight"])
            o = F.relu(F.group_norm(o, NGROUPS, sd[q + "convs.1.weight"], sd[q + "convs.1.bias"], eps=1e-5))
            o = F.conv2d(o, sd[q + "convs.3.weight"], stride=s, padding=1)
            o = F.relu(F.group_norm(o, NGROUPS, sd[q + "convs.4.weight"], sd[q + "convs.4.bias"], eps=1e-5))
            o = F.conv2d(o, sd[q + "convs.6.weight"])
            o = F.group_norm(o, NGROUPS, sd[q + "convs.7.weight"], sd[q + "convs.7.bias"], eps=1e-5)
            if (q + "downsample.0.weight") in sd:
                idt = F.conv2d(x, sd[q + "downsample.0.weight"], stride=s)
                idt = F.group_norm(idt, NGROUPS, sd[q + "downsample.1.weight"], sd[q + "downsample.1.bias"], eps=1e-5)
            x = F.relu(o + idt)
    x = F.conv2d(x, sd[p + "compression.0.weight"], None, padding=1)
    x = F.relu(F.group_norm(x, 1, sd[p + "compression.1.weight"], sd[p + "compression.1.bias"], eps=1e-5))
    return x


def depth_encoder_hi(sd: SD, depth: torch.Tensor) -> torch.Tensor:
    """hi: [B,192,4,4] (resnet_encoders.py:88-104)."""
    x = depth_trunk(sd, "depth_encoder.", depth)
    e = spatial_embedding_channels(sd["depth_encoder.spatial_embeddings.weight"]).expand(x.shape[0], -1, -1, -1)
    return torch.cat([x, e], dim=1)


def depth_encoder_lo(sd: SD, depth: torch.Tensor) -> torch.Tensor:
    """lo: Flatten -> Linear 2048->128 -> ReLU (resnet_encoders.py:58-62,108)."""
    x = torch.flatten(depth_trunk(sd, "depth_encoder.", depth), 1)
    return F.relu(F.linear(x, sd["depth_encoder.visual_fc.1.weight"], sd["depth_encoder.visual_fc.1.bias"]))


# --------------------------------------------------------------------------------------
# BERT-base encoder (transformers.BertModel, eval, input_ids only)
# --------------------------------------------------------------------------------------
def bert(sd: SD, ids: torch.Tensor, prefix: str = "embedding_layer.", n_layers: int = 12,
         n_heads: int = 12) -> torch.Tensor:
    """ids int64 [R,L] -> last_hidden_state [R,L,768]."""
    R, L = ids.shape
    e = prefix + "embeddings."
    x = sd[e + "word_embeddings.weight"][ids] \
        + sd[e + "token_type_embeddings.weight"][0] \
        + sd[e + "position_embeddings.weight"][:L].unsqueeze(0)
    H = x.shape[-1]
    x = F.layer_norm(x, (H,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps=1e-12)
    dh = H // n_heads
    for i in range(n_layers):
        q_ = f"{prefix}encoder.layer.{i}."
        a = q_ + "attention.self."
        q = F.linear(x, sd[a + "query.weight"], sd[a + "query.bias"]).view(R, L, n_heads, dh).transpose(1, 2)
        k = F.linear(x, sd[a + "key.weight"], sd[a + "key.bias"]).view(R, L, n_heads, dh).transpose(1, 2)
        v = F.linear(x, sd[a + "value.weight"], sd[a + "value.bias"]).view(R, L, n_heads, dh).transpose(1, 2)
        s = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
        ctx = (s @ v).transpose(1, 2).reshape(R, L, H)
        o = q_ + "attention.output."
        y = F.linear(ctx, sd[o + "dense.weight"], sd[o + "dense.bias"])
        x = F.layer_norm(y + x, (H,), sd[o + "LayerNorm.weight"], sd[o + "LayerNorm.bias"], eps=1e-12)
        h = F.gelu(F.linear(x, sd[q_ + "intermediate.dense.weight"], sd[q_ + "intermediate.dense.bias"]))
        y = F.linear(h, sd[q_ + "output.dense.weight"], sd[q_ + "output.dense.bias"])
        x = F.layer_norm(y + x, (H,), sd[q_ + "output.LayerNorm.weight"], sd[q_ + "output.LayerNorm.bias"], eps=1e-12)
    return x


# --------------------------------------------------------------------------------------
# Visual_Ling_Attn (transformer.py:251-281), N=1, d_model=256, h=4, d_ff=1024
# --------------------------------------------------------------------------------------
def sinusoid_table(L: int, d_model: int) -> torch.Tensor:
    """common/utils.py:167-185: PE[p,2i]=sin(p/10000^(2i/d)), PE[p,2i+1]=cos(same)."""
    pos = torch.arange(L, dtype=torch.float32).view(-1, 1)
    dim = torch.arange(d_model // 2, dtype=torch.float32).view(1, -1)
    ang = pos / 10000 ** (2 * dim / d_model)
    out = torch.zeros((L, d_model))
    out[:, ::2] = torch.sin(ang)
    out[:, 1::2] = torch.cos(ang)
    return out


def visual_ling_attn(sd: SD, ins: torch.Tensor, vis: torch.Tensor, prefix: str = "image_cm_encoder.",
                     h: int = 4) -> torch.Tensor:
    """ins [B,L,768] (BERT output), vis [B,16,256] -> [B,L,256]."""
    B, L, _ = ins.shape
    d = sd[prefix + "vis_fc.weight"].shape[0]
    lnw, lnb = sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"]
    V = F.layer_norm(F.relu(F.linear(vis, sd[prefix + "vis_fc.weight"], sd[prefix + "vis_fc.bias"])), (d,), lnw, lnb, 1e-5)
    Q = F.layer_norm(F.relu(F.linear(ins, sd[prefix + "ins_fc.weight"], sd[prefix + "ins_fc.bias"])), (d,), lnw, lnb, 1e-5)
    Q = Q + sinusoid_table(L, d).unsqueeze(0).to(Q.device)   # built on the CPU as the reference does (transformer.py:271-273)
    a = prefix + "layers.0.enc_att.attention."
    dk = d // h
    nk = V.shape[1]
    q = F.linear(Q, sd[a + "fc_q.weight"], sd[a + "fc_q.bias"]).view(B, L, h, dk).permute(0, 2, 1, 3)
    k = F.linear(V, sd[a + "fc_k.weight"], sd[a + "fc_k.bias"]).view(B, nk, h, dk).permute(0, 2, 3, 1)
    v = F.linear(V, sd[a + "fc_v.weight"], sd[a + "fc_v.bias"]).view(B, nk, h, dk).permute(0, 2, 1, 3)
    att = torch.softmax(torch.matmul(q, k) / math.sqrt(dk), -1)
    o = torch.matmul(att, v).permute(0, 2, 1, 3).reshape(B, L, h * dk)
    o = F.linear(o, sd[a + "fc_o.weight"], sd[a + "fc_o.bias"])
    n1 = prefix + "layers.0.enc_att.layer_norm."
    X = F.layer_norm(Q + o, (d,), sd[n1 + "weight"], sd[n1 + "bias"], 1e-5)
    f = prefix + "layers.0.pwff."
    Y = F.linear(F.relu(F.linear(X, sd[f + "fc1.weight"], sd[f + "fc1.bias"])), sd[f + "fc2.weight"], sd[f + "fc2.bias"])
    return F.layer_norm(X + Y, (d,), sd[f + "layer_norm.weight"], sd[f + "layer_norm.bias"], 1e-5)


# --------------------------------------------------------------------------------------
# RNNStateEncoder around nn.LSTM (rnn_state_encoder.py:74-142)
# --------------------------------------------------------------------------------------
def lstm_state_encoder(sd: SD, prefix: str, x: torch.Tensor, hidden: torch.Tensor,
                       masks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [T*N, I] (row = t*N+n), hidden [2,N,H] = cat(h,c), masks [T*N] -> ([T*N,H], [2,N,H]).

    The reference multiplies (h,c) by masks[t] at t=0 and at every later step where ANY env
    has a zero mask (segment starts, rnn_state_encoder.py:104-127); elsewhere masks are all
    ones, so this equals multiplying by masks[t] at exactly those steps.
    Single-step batches (rows == N) multiply by the mask unconditionally (:74-83).
    """
    w_ih, w_hh = sd[prefix + "rnn.weight_ih_l0"], sd[prefix + "rnn.weight_hh_l0"]
    b = sd[prefix + "rnn.bias_ih_l0"] + sd[prefix + "rnn.bias_hh_l0"]
    N, H = hidden.shape[1], hidden.shape[2]
    T = x.shape[0] // N
    x = x.view(T, N, -1)
    m = masks.view(T, N)
    h, c = hidden[0], hidden[1]
    outs = []
    for t in range(T):
        if t == 0 or bool((m[t] == 0.0).any()):
            h = h * m[t].view(N, 1)
            c = c * m[t].view(N, 1)
        g = F.linear(x[t], w_ih) + F.linear(h, w_hh) + b
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 0).view(T * N, H), torch.stack([h, c], 0)


# --------------------------------------------------------------------------------------
# hi / lo forward
# --------------------------------------------------------------------------------------
def hi_forward(sd: SD, rgb: torch.Tensor, depth: torch.Tensor, instruction: torch.Tensor,
               hidden: torch.Tensor, masks: torch.Tensor, return_intermediates: bool = False):
    """seq2seq_highlevel_cma.py:170-233.  Returns (logits [B,4], hidden [2,N,512])."""
    inter = {}
    D = torch.flatten(depth_encoder_hi(sd, depth), 2)            # [B,192,16]
    R = torch.flatten(rgb_encoder_hi(sd, rgb), 2)                # [B,2112,16]
    B = R.shape[0]
    ids = instruction.long().expand(B, instruction.shape[1])
    emb = bert(sd, ids)                                          # [B,L,768]
    Kr = F.conv1d(R, sd["rgb_kv.weight"], sd["rgb_kv.bias"])     # [B,256,16]
    Kd = F.conv1d(D, sd["depth_kv.weight"], sd["depth_kv.bias"])
    Ar = visual_ling_attn(sd, emb, Kr.permute(0, 2, 1))          # [B,L,256]
    Ad = visual_ling_attn(sd, emb, Kd.permute(0, 2, 1))
    ar = Ar.mean(dim=1)                                          # AdaptiveAvgPool1d(1) over tokens
    ad = Ad.mean(dim=1)
    ri = F.relu(F.linear(R.mean(dim=2), sd["rgb_linear.2.weight"], sd["rgb_linear.2.bias"]))
    di = F.relu(F.linear(torch.flatten(D, 1), sd["depth_linear.1.weight"], sd["depth_linear.1.bias"]))
    x = torch.cat((ri, di, ar, ad), dim=1)                       # [B,896]
    y, hid = lstm_state_encoder(sd, "state_encoder.", x, hidden, masks[:, 0])
    logits = F.linear(y, sd["linear.weight"], sd["linear.bias"])
    if return_intermediates:
        inter.update(depth_embedding=D, rgb_embedding=R, bert=emb, rgb_spatial=Kr, depth_spatial=Kd,
                     ins_rgb_att=ar, ins_depth_att=ad, rgb_in=ri, depth_in=di, rnn_in=x, rnn_out=y)
        return logits, hid, inter
    return logits, hid


def lo_forward(sd: SD, rgb: torch.Tensor, depth: torch.Tensor, hidden: torch.Tensor,
               masks: torch.Tensor, sub_goal: torch.Tensor, return_intermediates: bool = False):
    """seq2seq_lowlevel.py:116-162.  Returns (actions [B,2], stop_logit [B,1], hidden)."""
    de = depth_encoder_lo(sd, depth)                             # [B,128]
    re = rgb_encoder_lo(sd, rgb)                                 # [B,256]
    se = sd["sub_task_embedding.weight"][sub_goal.long()]        # [B,32]
    x = torch.cat([de, re, se], dim=1)                           # [B,416]
    y, hid = lstm_state_encoder(sd, "state_encoder.", x, hidden, masks[:, 0])
    act = F.linear(y, sd["linear.weight"], sd["linear.bias"])
    stop = F.linear(y, sd["stop_linear.weight"], sd["stop_linear.bias"])
    if return_intermediates:
        return act, stop, hid, dict(depth_embedding=de, rgb_embedding=re, rnn_in=x, rnn_out=y)
    return act, stop, hid
