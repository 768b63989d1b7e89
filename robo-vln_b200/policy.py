"""HcmPolicy: the hierarchy as one object (hi -> argmax -> lo), the call a rollout loop makes
(robo_vln_baselines/hierarchical_trainer.py:1095-1101), backed by ``hcm_forward_policy``."""
from __future__ import annotations

import torch
import torch.nn as nn

from .seq2seq_highlevel_cma import Seq2Seq_HighLevel_CMA
from .seq2seq_lowlevel import Seq2Seq_LowLevel


class HcmPolicy(nn.Module):
    def __init__(self, high_level: Seq2Seq_HighLevel_CMA = None, low_level: Seq2Seq_LowLevel = None):
        super().__init__()
        self.high_level = high_level if high_level is not None else Seq2Seq_HighLevel_CMA()
        self.low_level = low_level if low_level is not None else Seq2Seq_LowLevel()

    def share_frozen_trunks(self):
        """Copy hi's frozen RGB / depth trunks into lo (what loading the same pretrained
        torchvision / DDPPO checkpoints does in the reference), enabling single-pass trunks."""
        hi_sd = self.high_level.state_dict()
        lo_sd = self.low_level.state_dict()
        with torch.no_grad():
            for k, v in lo_sd.items():
                if k.startswith(("rgb_encoder.cnn.", "depth_encoder.visual_encoder.")) and k in hi_sd:
                    v.copy_(hi_sd[k])
        self.low_level._weights_changed()
        return self

    def _runtime(self):
        rt_hi = self.high_level.runtime()
        rt_lo = self.low_level.runtime()
        if rt_hi is not rt_lo:
            raise RuntimeError("hi and lo are bound to different engines; construct them as a pair on one device")
        return rt_hi

    @torch.no_grad()
    def act(self, observations, hidden_hi, hidden_lo, masks):
        """-> (logits [B,4], actions [B,2], stop_logit [B,1], hidden_hi, hidden_lo, sub_goal [B])"""
        rt = self._runtime()
        return rt.forward_policy(observations["rgb"], observations["depth"], observations["instruction"], masks,
                                 hidden_hi, hidden_lo)

    @torch.no_grad()
    def act_host(self, rgb, depth, instruction, masks, hidden_hi, hidden_lo, out=None):
        """Same step from host (CPU, ideally pinned) float32 buffers; returns host tensors."""
        return self._runtime().forward_policy_host(rgb, depth, instruction, masks, hidden_hi, hidden_lo, out)

    forward = act
