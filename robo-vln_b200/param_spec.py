"""Parameter / buffer layout of the two HCM modules, derived from the architecture.

The trainer saves and loads ``state_dict``s of both models
(robo_vln_baselines/hierarchical_trainer.py:342-363), so key names, shapes and order are part
of the drop-in boundary.  This module enumerates them programmatically (no file from the
reference or from ``oracle/`` is read); ``tests/test_param_spec.py`` checks the result against
the layout dumped from the reference constructors.

Reference constructors mirrored here:
  Seq2Seq_HighLevel_CMA.__init__   robo_vln_baselines/models/seq2seq_highlevel_cma.py:33-141
  Seq2Seq_LowLevel.__init__        robo_vln_baselines/models/seq2seq_lowlevel.py:32-98
  TorchVisionResNet50.__init__     robo_vln_baselines/models/encoders/resnet_encoders.py:121-187
  VlnResnetDepthEncoder.__init__   robo_vln_baselines/models/encoders/resnet_encoders.py:14-73
  ResNetEncoder / ResNet           habitat_baselines/rl/ddppo/policy/resnet_policy.py:80-158, resnet.py:181-242
  Visual_Ling_Attn.__init__        robo_vln_baselines/models/transformer/transformer.py:252-260
  RNNStateEncoder.__init__         habitat_baselines/rl/models/rnn_state_encoder.py:12-41
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

Spec = "OrderedDict[str, Tuple[Tuple[int, ...], torch.dtype, bool]]"  # shape, dtype, is_buffer

F32 = torch.float32
I64 = torch.int64


def _add(spec, key, shape, dtype=F32, buffer=False):
    spec[key] = (tuple(shape), dtype, buffer)


def _linear(spec, p, out_f, in_f, bias=True):
    _add(spec, p + ".weight", (out_f, in_f))
    if bias:
        _add(spec, p + ".bias", (out_f,))


def _norm(spec, p, c):
    _add(spec, p + ".weight", (c,))
    _add(spec, p + ".bias", (c,))


def _bn(spec, p, c):
    _norm(spec, p, c)
    _add(spec, p + ".running_mean", (c,), F32, True)
    _add(spec, p + ".running_var", (c,), F32, True)
    _add(spec, p + ".num_batches_tracked", (), I64, True)


def _bert(spec, p, hidden=768, layers=12, inter=3072, vocab=30522, max_pos=512):
    e = p + "embeddings."
    _add(spec, e + "word_embeddings.weight", (vocab, hidden))
    _add(spec, e + "position_embeddings.weight", (max_pos, hidden))
    _add(spec, e + "token_type_embeddings.weight", (2, hidden))
    _norm(spec, e + "LayerNorm", hidden)
    for i in range(layers):
        q = f"{p}encoder.layer.{i}."
        _linear(spec, q + "attention.self.query", hidden, hidden)
        _linear(spec, q + "attention.self.key", hidden, hidden)
        _linear(spec, q + "attention.self.value", hidden, hidden)
        _linear(spec, q + "attention.output.dense", hidden, hidden)
        _norm(spec, q + "attention.output.LayerNorm", hidden)
        _linear(spec, q + "intermediate.dense", inter, hidden)
        _linear(spec, q + "output.dense", hidden, inter)
        _norm(spec, q + "output.LayerNorm", hidden)
    _linear(spec, p + "pooler.dense", hidden, hidden)


_STAGES = ((3, 1), (4, 2), (6, 2), (3, 2))  # ResNet-50: blocks, first stride


def _depth_trunk(spec, p, base=32):
    b = p + "backbone."
    _add(spec, b + "conv1.0.weight", (base, 1, 7, 7))
    _norm(spec, b + "conv1.1", base)
    cin = base
    for li, (nb, _stride) in enumerate(_STAGES):
        mid = base << li
        cout = mid * 4
        for blk in range(nb):
            q = f"{b}layer{li + 1}.{blk}."
            _add(spec, q + "convs.0.weight", (mid, cin, 1, 1))
            _norm(spec, q + "convs.1", mid)
            _add(spec, q + "convs.3.weight", (mid, mid, 3, 3))
            _norm(spec, q + "convs.4", mid)
            _add(spec, q + "convs.6.weight", (cout, mid, 1, 1))
            _norm(spec, q + "convs.7", cout)
            if blk == 0:
                _add(spec, q + "downsample.0.weight", (cout, cin, 1, 1))
                _norm(spec, q + "downsample.1", cout)
            cin = cout
    _add(spec, p + "compression.0.weight", (128, cin, 3, 3))
    _norm(spec, p + "compression.1", 128)


def _rgb_trunk(spec, p, with_fc: bool):
    _add(spec, p + "conv1.weight", (64, 3, 7, 7))
    _bn(spec, p + "bn1", 64)
    cin = 64
    for li, (nb, _stride) in enumerate(_STAGES):
        mid = 64 << li
        cout = mid * 4
        for blk in range(nb):
            q = f"{p}layer{li + 1}.{blk}."
            _add(spec, q + "conv1.weight", (mid, cin, 1, 1))
            _bn(spec, q + "bn1", mid)
            _add(spec, q + "conv2.weight", (mid, mid, 3, 3))
            _bn(spec, q + "bn2", mid)
            _add(spec, q + "conv3.weight", (cout, mid, 1, 1))
            _bn(spec, q + "bn3", cout)
            if blk == 0:
                _add(spec, q + "downsample.0.weight", (cout, cin, 1, 1))
                _bn(spec, q + "downsample.1", cout)
            cin = cout
    if with_fc:
        _linear(spec, p + "fc", 1000, 2048)   # torchvision's classifier, constructed but unused by lo


def _lstm(spec, p, in_f, hid=512):
    _add(spec, p + "rnn.weight_ih_l0", (4 * hid, in_f))
    _add(spec, p + "rnn.weight_hh_l0", (4 * hid, hid))
    _add(spec, p + "rnn.bias_ih_l0", (4 * hid,))
    _add(spec, p + "rnn.bias_hh_l0", (4 * hid,))


def hi_spec(num_actions: int = 4):
    s = OrderedDict()
    _bert(s, "embedding_layer.")
    _linear(s, "ins_fc", 256, 768)
    _depth_trunk(s, "depth_encoder.visual_encoder.")
    _add(s, "depth_encoder.spatial_embeddings.weight", (16, 64))
    _rgb_trunk(s, "rgb_encoder.cnn.", with_fc=False)
    _add(s, "rgb_encoder.spatial_embeddings.weight", (16, 64))
    _linear(s, "rgb_linear.2", 256, 2112)
    _linear(s, "depth_linear.1", 128, 3072)
    _add(s, "rgb_kv.weight", (256, 2112, 1))
    _add(s, "rgb_kv.bias", (256,))
    _add(s, "depth_kv.weight", (256, 192, 1))
    _add(s, "depth_kv.bias", (256,))
    a = "image_cm_encoder.layers.0.enc_att."
    for n in ("fc_q", "fc_k", "fc_v", "fc_o"):
        _linear(s, a + "attention." + n, 256, 256)
    _norm(s, a + "layer_norm", 256)
    f = "image_cm_encoder.layers.0.pwff."
    _linear(s, f + "fc1", 1024, 256)
    _linear(s, f + "fc2", 256, 1024)
    _norm(s, f + "layer_norm", 256)
    _linear(s, "image_cm_encoder.vis_fc", 256, 256)
    _linear(s, "image_cm_encoder.ins_fc", 256, 768)
    _norm(s, "image_cm_encoder.layer_norm", 256)
    _lstm(s, "state_encoder.", 896)
    _linear(s, "progress_monitor", 1, 512)
    _linear(s, "linear", num_actions, 512)
    return s


def lo_spec(num_actions: int = 2, num_sub_tasks: int = 4):
    s = OrderedDict()
    _depth_trunk(s, "depth_encoder.visual_encoder.")
    _linear(s, "depth_encoder.visual_fc.1", 128, 2048)
    _rgb_trunk(s, "rgb_encoder.cnn.", with_fc=True)
    _linear(s, "rgb_encoder.fc", 256, 2048)
    _add(s, "sub_task_embedding.weight", (num_sub_tasks + 1, 32))
    _lstm(s, "state_encoder.", 416)
    _linear(s, "progress_monitor", 1, 512)
    _linear(s, "linear", num_actions, 512)
    _linear(s, "stop_linear", 1, 512)
    return s


FROZEN_PREFIXES = ("embedding_layer.", "rgb_encoder.cnn.", "depth_encoder.visual_encoder.")


def default_init(key: str, shape, dtype, seed: int = 0) -> torch.Tensor:
    """Deterministic constructor-time initialisation (scale-preserving).  The reference
    constructors download pretrained BERT / ResNet weights, which is not possible offline;
    real use loads a checkpoint with ``load_state_dict`` right after construction
    (hierarchical_trainer.py:342-346)."""
    if dtype == I64:
        return torch.zeros(shape, dtype=I64)
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    leaf = key.rsplit(".", 1)[-1]
    n = len(shape)
    if leaf == "running_mean":
        return torch.zeros(shape)
    if leaf == "running_var":
        return torch.ones(shape)
    if n == 1:
        if leaf == "bias":
            return torch.zeros(shape)
        return torch.ones(shape)
    if "embeddings" in key or "embedding" in key:
        std = 0.02 if "embedding_layer" in key else 1.0
        w = torch.randn(shape, generator=g) * std
        if key == "sub_task_embedding.weight":
            w[4].zero_()          # padding_idx=4 (seq2seq_lowlevel.py:76)
        return w
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    gain = math.sqrt(2.0) if n == 4 else 1.0
    return torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
