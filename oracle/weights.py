"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and inputs for the HCM path.

There are no pretrained weights in this environment (no network), and the reference's
default constructors leave every BatchNorm at identity, which would not exercise BN
folding.  This module fills the reference's exact ``state_dict`` layout
(``oracle/manifest_{hi,lo}.json``: key -> shape, dumped from the reference constructors
``robo_vln_baselines/models/seq2seq_highlevel_cma.py:33-141`` and
``robo_vln_baselines/models/seq2seq_lowlevel.py:32-98``) with seeded values whose scale
keeps activations O(1..10) through all 50+ layers.  Every tensor is generated from its own
``torch.Generator`` seeded by crc32(key) ^ seed, so the values do not depend on iteration
order, on the other model, or on the machine (CPU philox/mt19937 streams are portable).

The frozen trunks (``rgb_encoder.cnn.*`` and ``depth_encoder.visual_encoder.*``) get the
SAME values in hi and lo (as with the pretrained-frozen weights the reference loads),
because the key passed to the generator is the trunk-relative one.
"""
from __future__ import annotations

import json
import math
import os
import zlib

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def load_manifest(which: str) -> dict:
    with open(os.path.join(_HERE, f"manifest_{which}.json")) as f:
        return json.load(f)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _fill(key: str, shape, seed: int) -> torch.Tensor:
    g = _gen(key, seed)
    leaf = key.split(".")[-1]
    n = len(shape)

    def normal(std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    def uniform(lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return normal(0.1)
    if leaf == "running_var":
        return uniform(0.6, 1.4)
    if "spatial_embeddings" in key:
        return normal(1.0)
    if "sub_task_embedding" in key:
        w = normal(1.0)
        w[4].zero_()                       # padding_idx=4 (seq2seq_lowlevel.py:76)
        return w
    if "embeddings.word_embeddings" in key or "embeddings.position_embeddings" in key \
            or "embeddings.token_type_embeddings" in key:
        return normal(0.1)
    if n == 1:
        if leaf == "bias":
            return normal(0.05)
        # 1-D "weight": BatchNorm / GroupNorm / LayerNorm scale.  The last norm of every
        # bottleneck gets a small scale so the residual stream does not blow up.
        if key.endswith("bn3.weight") or key.endswith("convs.7.weight") or key.endswith("downsample.1.weight"):
            return uniform(0.25, 0.55)
        return uniform(0.7, 1.3)
    if n == 4:                              # conv2d OIHW
        fan_in = shape[1] * shape[2] * shape[3]
        return normal(math.sqrt(2.0 / fan_in))
    if n == 3:                              # conv1d (k=1)
        return normal(1.0 / math.sqrt(shape[1] * shape[2]))
    if n == 2:
        if "rnn.weight" in key:
            return normal(1.0 / math.sqrt(shape[1]))
        return normal(1.0 / math.sqrt(shape[1]))
    return normal(0.05)


_TRUNK_PREFIXES = ("rgb_encoder.cnn.", "depth_encoder.visual_encoder.")


def make_state_dict(which: str, seed: int = 0) -> dict:
    """Synthetic state_dict for ``which`` in {"hi", "lo"} with the reference's keys."""
    man = load_manifest(which)
    sd = {}
    for key, (shape, _dtype) in man.items():
        gen_key = key if key.startswith(_TRUNK_PREFIXES) else f"{which}:{key}"
        sd[key] = _fill(gen_key, tuple(shape), seed)
    return sd


def make_inputs(B: int, L: int, N: int, *, rgb_hw: int = 256, seed: int = 1,
                shared_instruction: bool = False, pad_tail: int = 0,
                mask_zero_rows=(0,), hidden_std: float = 0.5) -> dict:
    """Synthetic batch in the layout the reference trainer feeds.

    rows = B = T*N, ordered t-major (row = t*N + n) like ``RNNStateEncoder.seq_forward``.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    rgb = torch.randint(0, 256, (B, rgb_hw, rgb_hw, 3), generator=g).float()
    depth = torch.rand((B, 256, 256, 1), generator=g)
    rows = 1 if shared_instruction else B
    ids = torch.randint(1000, 30522, (rows, L), generator=g)
    ids[:, 0] = 101
    if pad_tail > 0:
        ids[:, L - pad_tail - 1] = 102
        ids[:, L - pad_tail:] = 0
    else:
        ids[:, L - 1] = 102
    masks = torch.ones((B, 2))
    for r in mask_zero_rows:
        masks[r] = 0.0
    hid_hi = torch.randn((2, N, 512), generator=g) * hidden_std
    hid_lo = torch.randn((2, N, 512), generator=g) * hidden_std
    sub_goal = torch.randint(0, 5, (B,), generator=g)
    return {
        "rgb": rgb, "depth": depth, "instruction": ids.float(), "masks": masks,
        "hidden_hi": hid_hi, "hidden_lo": hid_lo, "sub_goal": sub_goal,
        "prev_actions": torch.zeros((B, 2)),
    }
