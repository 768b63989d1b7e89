"""Differentiable trainable tail of the two HCM models (training path).

In training (`module.train()` with autograd enabled) the frozen encoders -- RGB ResNet-50, depth
ResNet-50 and BERT, 99 % of the FLOPs, none of which receives gradients in the reference
(resnet_encoders.py:35-36,147-148; seq2seq_highlevel_cma.py:192-195) -- run on the sm_100a
engine (`hcm_run_encoders`), and the ~5 M-parameter trainable tail below runs as ordinary
PyTorch ops on the engine's feature buffers so that `loss.backward()` in
hierarchical_trainer.py:506-513,539-555 produces the same gradients the reference would.
Dropout (p = 0.25, five sites in Visual_Ling_Attn) is applied in training mode exactly where the
reference applies it.  Inference never comes here: it is one engine call.

Mirrors: seq2seq_highlevel_cma.py:198-233, seq2seq_lowlevel.py:141-162,
transformer.py:38-43,81-126,209-221,262-281, common/utils.py:167-185,
habitat_baselines/rl/models/rnn_state_encoder.py:74-142.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _spatial(weight: torch.Tensor) -> torch.Tensor:
    """[16,64] embedding table -> [16 cells, 64 channels] as the reference's
    ``.view(1,-1,4,4)`` reads it (channel c of cell k = flat[c*16 + k])."""
    return weight.reshape(64, 16).t()


def sinusoid_table(L: int, d: int, device) -> torch.Tensor:
    pos = torch.arange(L, dtype=torch.float32, device=device).view(-1, 1)
    dim = torch.arange(d // 2, dtype=torch.float32, device=device).view(1, -1)
    ang = pos / 10000 ** (2 * dim / d)
    out = torch.zeros((L, d), device=device)
    out[:, ::2] = torch.sin(ang)
    out[:, 1::2] = torch.cos(ang)
    return out


def ins_projection(m, ins: torch.Tensor) -> torch.Tensor:
    """relu(ins_fc(ins)) -- the part of the query side of Visual_Ling_Attn.forward (transformer.py:266-268) that draws no
    random numbers.  hi_tail computes it ONCE on the un-expanded instruction rows ([1, L, 768] when the trajectory shares one
    instruction) and for both modalities; every call then applies its own dropout mask and LayerNorm to the expanded rows,
    exactly as the reference does on its B-times repeated copy (the weight gradient is the same sum, taken in two stages)."""
    return F.relu(F.linear(ins, m.ins_fc.weight, m.ins_fc.bias))


def visual_ling_attn(m, ins: torch.Tensor, vis: torch.Tensor, p_drop: float, training: bool, h: int = 4, q_lin=None):
    """m = module.image_cm_encoder; ins [B,L,768] (or q_lin = ins_projection(...) [1|B,L,256]), vis [B,16,256] -> [B,L,256]."""
    B = vis.shape[0]
    L = (ins if q_lin is None else q_lin).shape[1]
    d = m.vis_fc.weight.shape[0]
    drop = lambda t: F.dropout(t, p_drop, training)  # noqa: E731
    ln0 = (m.layer_norm.weight, m.layer_norm.bias)
    V = F.layer_norm(drop(F.relu(F.linear(vis, m.vis_fc.weight, m.vis_fc.bias))), (d,), *ln0, 1e-5)
    if q_lin is None:
        q_lin = ins_projection(m, ins)
    Q = F.layer_norm(drop(q_lin.expand(B, -1, -1)), (d,), *ln0, 1e-5)
    Q = Q + sinusoid_table(L, d, vis.device).unsqueeze(0)
    layer = m.layers._modules["0"]
    a = layer.enc_att.attention
    dk = d // h
    nk = V.shape[1]
    q = F.linear(Q, a.fc_q.weight, a.fc_q.bias).view(B, L, h, dk).permute(0, 2, 1, 3)
    k = F.linear(V, a.fc_k.weight, a.fc_k.bias).view(B, nk, h, dk).permute(0, 2, 3, 1)
    v = F.linear(V, a.fc_v.weight, a.fc_v.bias).view(B, nk, h, dk).permute(0, 2, 1, 3)
    att = torch.softmax(torch.matmul(q, k) / math.sqrt(dk), -1)
    o = torch.matmul(att, v).permute(0, 2, 1, 3).reshape(B, L, h * dk)
    o = drop(F.linear(o, a.fc_o.weight, a.fc_o.bias))
    X = F.layer_norm(Q + o, (d,), layer.enc_att.layer_norm.weight, layer.enc_att.layer_norm.bias, 1e-5)
    f = layer.pwff
    Y = F.linear(drop(F.relu(F.linear(X, f.fc1.weight, f.fc1.bias))), f.fc2.weight, f.fc2.bias)
    return F.layer_norm(X + drop(Y), (d,), f.layer_norm.weight, f.layer_norm.bias, 1e-5)


def segment_starts(masks: torch.Tensor, N: int):
    """First steps of the LSTM segments of a [T*N] mask vector: t = 0 and every t where any environment is reset
    (one host read of the masks; CUDA-graph callers compute it before capture and pass it in)."""
    T = masks.numel() // N
    m = masks.reshape(T, N)
    return [0] + ([t + 1 for t, v in enumerate((m[1:] == 0.0).any(dim=1).tolist()) if v] if T > 1 else [])


def lstm_state_encoder(rnn, x: torch.Tensor, hidden: torch.Tensor, masks: torch.Tensor, starts=None):
    """rnn = module.state_encoder.rnn (parameter container); x [T*N, I], hidden [2,N,H], masks [T*N].
    RNNStateEncoder.seq_forward (rnn_state_encoder.py:85-136): the trajectory is cut at t = 0 and at every step
    where any environment is reset, (h, c) are multiplied by the masks of the segment's first step, and each
    segment is ONE fused LSTM call (torch's native/cuDNN kernel, differentiable) instead of a Python loop of
    T cell updates -- 64 steps of forward + backward were 20 of the 42 ms of a DAgger step."""
    N, H = hidden.shape[1], hidden.shape[2]
    T = x.shape[0] // N
    m = masks.view(T, N)
    if starts is None:
        starts = segment_starts(masks, N)
    xs = x.view(T, N, -1)
    weights = [rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0]   # cuDNN re-packs them per call (warning is benign)
    h, c = hidden[0:1], hidden[1:2]
    outs = []
    for i, t0 in enumerate(starts):
        t1 = starts[i + 1] if i + 1 < len(starts) else T
        mk = m[t0].view(1, N, 1)
        # train flag: cuDNN keeps its reserve space for backward only in "training" mode (no dropout here)
        out, h, c = torch._VF.lstm(xs[t0:t1], (h * mk, c * mk), weights, True, 1, 0.0, torch.is_grad_enabled(), False, False)
        outs.append(out)
    y = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
    return y.reshape(T * N, H), torch.cat([h, c], 0)


def hi_tail(mod, rgb_feat, depth_feat, bert, hidden, masks, p_drop: float = 0.25, starts=None):
    """rgb_feat [B,16,2048], depth_feat [B,16,128], bert [1|B,L,768] (all constants, fp32);
    hidden [2,N,512]; masks [B,2] -> (logits [B,4], hidden [2,N,512])."""
    B = rgb_feat.shape[0]
    training = mod.training
    R = torch.cat([rgb_feat, _spatial(mod.rgb_encoder.spatial_embeddings.weight).unsqueeze(0).expand(B, -1, -1)], 2)
    D = torch.cat([depth_feat, _spatial(mod.depth_encoder.spatial_embeddings.weight).unsqueeze(0).expand(B, -1, -1)], 2)
    Kr = F.linear(R, mod.rgb_kv.weight[:, :, 0], mod.rgb_kv.bias)          # Conv1d(k=1) == per-cell linear
    Kd = F.linear(D, mod.depth_kv.weight[:, :, 0], mod.depth_kv.bias)
    q_lin = ins_projection(mod.image_cm_encoder, bert)      # [1|B, L, 256], shared by the two calls
    Ar = visual_ling_attn(mod.image_cm_encoder, None, Kr, p_drop, training, q_lin=q_lin)
    Ad = visual_ling_attn(mod.image_cm_encoder, None, Kd, p_drop, training, q_lin=q_lin)
    lin_r = mod.rgb_linear._modules["2"]
    lin_d = mod.depth_linear._modules["1"]
    ri = F.relu(F.linear(R.mean(dim=1), lin_r.weight, lin_r.bias))
    di = F.relu(F.linear(D.permute(0, 2, 1).reshape(B, -1), lin_d.weight, lin_d.bias))   # channel-major flatten
    x = torch.cat((ri, di, Ar.mean(dim=1), Ad.mean(dim=1)), dim=1)
    y, hid = lstm_state_encoder(mod.state_encoder.rnn, x, hidden, masks[:, 0], starts)
    return F.linear(y, mod.linear.weight, mod.linear.bias), hid


def lo_tail(mod, rgb_gmean, depth_feat, hidden, masks, sub_goal, starts=None):
    """rgb_gmean [B,2048], depth_feat [B,16,128] -> (actions [B,2], stop [B,1], hidden)."""
    B = rgb_gmean.shape[0]
    fc_d = mod.depth_encoder.visual_fc._modules["1"]
    de = F.relu(F.linear(depth_feat.permute(0, 2, 1).reshape(B, -1), fc_d.weight, fc_d.bias))
    re = F.relu(F.linear(rgb_gmean, mod.rgb_encoder.fc.weight, mod.rgb_encoder.fc.bias))
    se = F.embedding(sub_goal.long().view(-1), mod.sub_task_embedding.weight, padding_idx=4)
    x = torch.cat([de, re, se], dim=1)
    y, hid = lstm_state_encoder(mod.state_encoder.rnn, x, hidden, masks[:, 0], starts)
    return (F.linear(y, mod.linear.weight, mod.linear.bias),
            F.linear(y, mod.stop_linear.weight, mod.stop_linear.bias), hid)
