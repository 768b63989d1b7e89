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
        def sig():
            return [(t.data_ptr(), t.dtype, t.device) for t in list(self.parameters()) + list(self.buffers())]

        rt = self.__dict__.get("_rt")
        before = sig() if rt is not None else None
        out = super()._apply(fn, *a, **k)
        if rt is not None:
            first = next(self.parameters())
            if first.device != rt.device:
                self.__dict__["_rt"] = None      # moved to another device: bind a new runtime lazily
            elif sig() != before:
                rt.mark_dirty()                  # storage / dtype changed; a same-device .to() (the trainer calls
                #                                  low_level.to(device2) every update, hierarchical_trainer.py:517) is a no-op
        return out

    # ---- constructor-time weights of the frozen encoders -----------------------------------------
    def _load_ddppo_checkpoint(self, path: str) -> None:
        """DEPTH_ENCODER.ddppo_checkpoint, exactly as VlnResnetDepthEncoder.__init__ loads it
        (resnet_encoders.py:38-52): keys `actor_critic.net.visual_encoder.<name>` -> visual_encoder.<name>, strict."""
        ckpt = torch.load(path, map_location="cpu")
        weights = {}
        for k, v in ckpt["state_dict"].items():
            parts = k.split(".")[2:]
            if not parts or parts[0] != "visual_encoder":
                continue
            weights[".".join(parts[1:])] = v
        del ckpt
        self.depth_encoder.visual_encoder.load_state_dict(weights, strict=True)
        self._weights_changed()

    def _init_frozen_encoders(self, model_config) -> None:
        """The reference constructors fill the frozen encoders from pretrained checkpoints: the DDPPO depth
        ResNet from DEPTH_ENCODER.ddppo_checkpoint, torchvision's ImageNet ResNet-50 and (hi) `bert-base-uncased`.
        The first is a local file and is loaded the same way; the other two are downloads in the reference --
        here they are taken from the local torch-hub / HuggingFace caches when present, never from the network.
        Whatever could not be found stays at its seeded random initialisation and is reported LOUDLY: a trainer
        that starts from such a model (load_from_ckpt=False) would otherwise train on random frozen features."""
        import os
        import warnings

        self.random_frozen_encoders = []
        if model_config is None:
            return            # kernel-level construction (tests, bench): synthetic weights are intended
        ck = getattr(getattr(model_config, "DEPTH_ENCODER", None), "ddppo_checkpoint", "NONE")
        if ck != "NONE":
            self._load_ddppo_checkpoint(ck)     # missing file -> FileNotFoundError, as in the reference
        else:
            self.random_frozen_encoders.append("depth_encoder.visual_encoder (ddppo_checkpoint='NONE')")
        if os.environ.get("ROBOVLN_PRETRAINED", "1") != "0":
            hub = os.path.join(torch.hub.get_dir(), "checkpoints")
            for fn in ("resnet50-0676ba61.pth", "resnet50-19c8e357.pth"):
                fp = os.path.join(hub, fn)
                if os.path.exists(fp):
                    sd = torch.load(fp, map_location="cpu")
                    own = self.rgb_encoder.cnn.state_dict()
                    self.rgb_encoder.cnn.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
                    break
            else:
                self.random_frozen_encoders.append("rgb_encoder.cnn (torchvision resnet50 ImageNet weights not in the torch-hub cache)")
            if self._kind == "hi":
                try:
                    from transformers import BertModel

                    bert = BertModel.from_pretrained("bert-base-uncased", local_files_only=True)
                    own = self.embedding_layer.state_dict()
                    self.embedding_layer.load_state_dict({k: v for k, v in bert.state_dict().items() if k in own}, strict=False)
                except Exception:
                    self.random_frozen_encoders.append("embedding_layer (bert-base-uncased not in the HuggingFace cache)")
        if self.random_frozen_encoders:
            warnings.warn("robovln_b200.%s: frozen encoders left at RANDOM initialisation: %s. Load a checkpoint "
                          "(load_state_dict) before training or evaluating." % (type(self).__name__, "; ".join(self.random_frozen_encoders)),
                          RuntimeWarning, stacklevel=3)
        self._weights_changed()

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
