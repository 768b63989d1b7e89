"""Shared machinery of the two drop-in nn.Modules (parameter tree + runtime binding)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .param_spec import FROZEN_PREFIXES, default_init
from .runtime import HcmRuntime


class _Node(nn.Module):
    """Pure container: holds parameters / buffers under the reference's dotted names."""


class _StateEncoderNode(_Node):
    """Stands in for habitat's RNNStateEncoder: the trainer reads
    ``model.state_encoder.num_recurrent_layers`` to size the hidden state
    (hierarchical_trainer.py:651,657); LSTM packs (h, c) -> 2 * num_layers
    (habitat_baselines/rl/models/rnn_state_encoder.py:43-47)."""

    num_recurrent_layers = 2


def build_param_tree(root: nn.Module, spec, seed: int = 0) -> None:
    for key, (shape, dtype, is_buffer) in spec.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _StateEncoderNode() if (mod is root and p == "state_encoder") else _Node())
            mod = mod._modules[p]
        t = default_init(key, shape, dtype, seed)
        if is_buffer:
            mod.register_buffer(parts[-1], t)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=not key.startswith(FROZEN_PREFIXES)))


class HcmModuleBase(nn.Module):
    """Behaviour common to Seq2Seq_HighLevel_CMA and Seq2Seq_LowLevel."""

    _kind = "hi"

    def __init__(self):
        super().__init__()
        self._rt: Optional[HcmRuntime] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._weights_changed())

    # weights are cached in kernel layout inside the runtime; anything that can change them
    # invalidates the cache
    def _weights_changed(self):
        if self._rt is not None:
            self._rt.mark_dirty()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        rt = self.__dict__.get("_rt")
        if rt is not None:
            first = next(self.parameters())
            if first.device != rt.device:
                self.__dict__["_rt"] = None      # moved to another device: bind a new runtime lazily
            else:
                rt.mark_dirty()
        return out

    def train(self, mode: bool = True):
        # leaving training mode: optimizer steps may have changed the trainable tail, so the
        # engine's packed copies are re-made on the next inference call
        if self.training and not mode:
            self._weights_changed()
        return super().train(mode)

    def notify_weights_updated(self):
        """Call after an optimizer step so the next forward re-packs the trainable weights."""
        self._weights_changed()

    def runtime(self) -> HcmRuntime:
        dev = next(self.parameters()).device
        if self._rt is None or self._rt.device != dev:
            self._rt = HcmRuntime.for_module(self, self._kind, dev)
        return self._rt

    @property
    def output_size(self):
        return 512

    @property
    def is_blind(self):
        return False

    @property
    def num_recurrent_layers(self):
        return self.state_encoder.num_recurrent_layers
