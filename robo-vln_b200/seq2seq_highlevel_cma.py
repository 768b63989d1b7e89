"""Drop-in for ``robo_vln_baselines.models.seq2seq_highlevel_cma.Seq2Seq_HighLevel_CMA``.

Same constructor signature, ``forward(batch)`` contract, attributes and ``state_dict`` keys as
the reference class (robo_vln_baselines/models/seq2seq_highlevel_cma.py:29-233); the
computation is one call into the sm_100a engine (``hcm_forward_hi``).
"""
from __future__ import annotations

import torch

from .modules import HcmModuleBase, build_param_tree
from .param_spec import hi_spec


def _check_config(model_config):
    """The engine implements the HCM configuration (config/default.py + hierarchical_cma.yaml,
    SURVEY.md Appendix D); refuse anything else instead of silently computing something different."""
    if model_config is None:
        return
    g = lambda node, name, default: getattr(node, name, default) if node is not None else default  # noqa: E731
    want = [
        (g(getattr(model_config, "RGB_ENCODER", None), "cnn_type", "TorchVisionResNet50"), "TorchVisionResNet50", "RGB_ENCODER.cnn_type"),
        (g(getattr(model_config, "DEPTH_ENCODER", None), "cnn_type", "VlnResnetDepthEncoder"), "VlnResnetDepthEncoder", "DEPTH_ENCODER.cnn_type"),
        (g(getattr(model_config, "DEPTH_ENCODER", None), "backbone", "resnet50"), "resnet50", "DEPTH_ENCODER.backbone"),
        (g(getattr(model_config, "DEPTH_ENCODER", None), "output_size", 128), 128, "DEPTH_ENCODER.output_size"),
        (g(getattr(model_config, "RGB_ENCODER", None), "output_size", 256), 256, "RGB_ENCODER.output_size"),
        (g(getattr(model_config, "STATE_ENCODER", None), "hidden_size", 512), 512, "STATE_ENCODER.hidden_size"),
        (g(getattr(model_config, "STATE_ENCODER", None), "rnn_type", "LSTM"), "LSTM", "STATE_ENCODER.rnn_type"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "d_model", 256), 256, "VISUAL_LING_ATTN.d_model"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "h", 4), 4, "VISUAL_LING_ATTN.h"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "d_ff", 1024), 1024, "VISUAL_LING_ATTN.d_ff"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "N", 1), 1, "VISUAL_LING_ATTN.N"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "vis_in_features", 256), 256, "VISUAL_LING_ATTN.vis_in_features"),
        (g(getattr(model_config, "VISUAL_LING_ATTN", None), "ins_in_features", 768), 768, "VISUAL_LING_ATTN.ins_in_features"),
        (g(getattr(model_config, "TRANSFORMER_INSTRUCTION_ENCODER", None), "d_in", 768), 768, "TRANSFORMER_INSTRUCTION_ENCODER.d_in"),
        (g(getattr(model_config, "TRANSFORMER_INSTRUCTION_ENCODER", None), "d_model", 256), 256, "TRANSFORMER_INSTRUCTION_ENCODER.d_model"),
        (g(getattr(model_config, "IMAGE_CROSS_MODAL_ENCODER", None), "d_model", 256), 256, "IMAGE_CROSS_MODAL_ENCODER.d_model"),
        (g(getattr(model_config, "RGB_ENCODER", None), "resnet_output_size", 256), 256, "RGB_ENCODER.resnet_output_size"),
        (g(getattr(model_config, "SEQ2SEQ", None), "use_prev_action", False), False, "SEQ2SEQ.use_prev_action"),
        (g(getattr(model_config, "PROGRESS_MONITOR", None), "use", False), False, "PROGRESS_MONITOR.use"),
        (g(model_config, "ablate_depth", False), False, "ablate_depth"),
        (g(model_config, "ablate_rgb", False), False, "ablate_rgb"),
        (g(model_config, "ablate_instruction", False), False, "ablate_instruction"),
    ]
    for got, exp, name in want:
        if got != exp:
            raise NotImplementedError(f"robovln_b200 implements the HCM configuration only: MODEL.{name}={got!r}, expected {exp!r}")


class Seq2Seq_HighLevel_CMA(HcmModuleBase):
    r"""High-level cross-modal decoder: (RGB, depth, instruction) -> 4 sub-goal logits."""

    _kind = "hi"

    def __init__(self, observation_space=None, num_actions: int = 4, model_config=None, batch_size: int = 1):
        super().__init__()
        _check_config(model_config)
        if num_actions != 4:
            raise NotImplementedError("the HCM high-level head has 4 sub-goal logits")
        self.model_config = model_config
        self.batch_size = batch_size
        vla = getattr(model_config, "VISUAL_LING_ATTN", None) if model_config is not None else None
        self.dropout_p = float(getattr(vla, "dropout", 0.25)) if vla is not None else 0.25
        build_param_tree(self, hi_spec(num_actions))
        self._init_frozen_encoders(model_config)

    def forward(self, batch):
        r"""(observations, rnn_hidden_states, prev_actions, masks) = batch
        -> (logits [B,4], rnn_hidden_states [2,N,512])"""
        observations, rnn_hidden_states, prev_actions, masks = batch
        del batch
        instruction = observations["instruction"]
        rt = self.runtime()
        if "rgb_features" in observations or "depth_features" in observations:
            # pre-computed trunk outputs (resnet_encoders.py:83-84,207-208): rgb_features [B,2048,4,4] (after the
            # adaptive pool), depth_features [B,128,4,4].  BERT (and a trunk whose features are not given) still
            # run on the engine; the tail runs as torch ops on the GPU (autograd-capable, like the training path).
            from . import torch_tail

            dev = rt.device
            have_r, have_d = "rgb_features" in observations, "depth_features" in observations
            n_envs = rnn_hidden_states.shape[1]
            if have_r and have_d:
                B = observations["rgb_features"].shape[0]
                feats = {"bert": rt.encode_bert(instruction, B, n_envs)}
            else:
                feats = rt.encode(observations["rgb"], observations["depth"], instruction, n_envs=n_envs)
            if have_r:
                feats["rgb_feat"] = observations["rgb_features"].to(dev, torch.float32).flatten(2).permute(0, 2, 1)
            if have_d:
                feats["depth_feat"] = observations["depth_features"].to(dev, torch.float32).flatten(2).permute(0, 2, 1)
            with torch.set_grad_enabled(self.training and torch.is_grad_enabled()):
                logits, hidden = torch_tail.hi_tail(self, feats["rgb_feat"], feats["depth_feat"], feats["bert"],
                                                    rnn_hidden_states.to(dev, torch.float32), masks.to(dev, torch.float32),
                                                    self.dropout_p)
            del observations["instruction"]
            return logits, hidden
        if self.training and torch.is_grad_enabled():
            # training step (hierarchical_trainer.py:506-513): frozen encoders on the engine, the
            # trainable tail under autograd (torch_tail.py).  BatchNorm stays in eval mode -- the
            # reference's train()-mode drift of the frozen ResNet statistics is not reproduced.
            from . import torch_tail

            dev = rt.device
            feats = rt.encode(observations["rgb"], observations["depth"], instruction,
                              n_envs=rnn_hidden_states.shape[1])
            logits, hidden = torch_tail.hi_tail(self, feats["rgb_feat"], feats["depth_feat"], feats["bert"],
                                                rnn_hidden_states.to(dev, torch.float32),
                                                masks.to(dev, torch.float32), self.dropout_p)
        else:
            logits, hidden = rt.forward_hi(observations["rgb"], observations["depth"], instruction, masks,
                                           rnn_hidden_states)
        del observations["instruction"]            # the reference mutates the caller's dict (:196)
        return logits, hidden
