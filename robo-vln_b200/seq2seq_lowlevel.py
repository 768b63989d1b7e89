"""Drop-in for ``robo_vln_baselines.models.seq2seq_lowlevel.Seq2Seq_LowLevel``
(robo_vln_baselines/models/seq2seq_lowlevel.py:21-162): (RGB, depth, sub-goal id) ->
(linear/angular velocity [B,2], stop logit [B,1], hidden state).  One call into the sm_100a
engine (``hcm_forward_lo``); when lo's frozen trunks are bit-identical to the paired hi
model's and the observations are the tensors hi just saw, the trunk features are reused.
"""
from __future__ import annotations

import torch

from .modules import HcmModuleBase, build_param_tree
from .param_spec import lo_spec
from .seq2seq_highlevel_cma import _check_config


class Seq2Seq_LowLevel(HcmModuleBase):
    _kind = "lo"

    def __init__(self, observation_space=None, num_actions: int = 2, num_sub_tasks: int = 4, model_config=None,
                 batch_size: int = 1):
        super().__init__()
        _check_config(model_config)
        if num_actions != 2 or num_sub_tasks != 4:
            raise NotImplementedError("the HCM low-level head is (v, omega) + stop over 4 sub-tasks")
        self.model_config = model_config
        self.batch_size = batch_size
        build_param_tree(self, lo_spec(num_actions, num_sub_tasks))
        self._init_frozen_encoders(model_config)

    def forward(self, batch):
        r"""(observations, rnn_hidden_states, prev_actions, masks, discrete_actions) = batch
        -> (actions [B,2], stop_logit [B,1], rnn_hidden_states [2,N,512])"""
        observations, rnn_hidden_states, prev_actions, masks, discrete_actions = batch
        del batch
        rt = self.runtime()
        if "rgb_features" in observations or "depth_features" in observations:
            # pre-computed trunk outputs (resnet_encoders.py:83-84,207-208): rgb_features [B,2048(,1,1)] (global
            # average pool), depth_features [B,128,4,4]; the tail runs as torch ops on the GPU
            from . import torch_tail

            dev = rt.device
            have_r, have_d = "rgb_features" in observations, "depth_features" in observations
            feats = {}
            if not (have_r and have_d):
                feats = rt.encode(observations["rgb"], observations["depth"], None, n_envs=rnn_hidden_states.shape[1],
                                  use_lo_weights=True)
            if have_r:
                feats["rgb_gmean"] = observations["rgb_features"].to(dev, torch.float32).flatten(1)
            if have_d:
                feats["depth_feat"] = observations["depth_features"].to(dev, torch.float32).flatten(2).permute(0, 2, 1)
            with torch.set_grad_enabled(self.training and torch.is_grad_enabled()):
                return torch_tail.lo_tail(self, feats["rgb_gmean"], feats["depth_feat"], rnn_hidden_states.to(dev, torch.float32),
                                          masks.to(dev, torch.float32), discrete_actions.to(dev))
        if self.training and torch.is_grad_enabled():
            # training step (hierarchical_trainer.py:539-555): see Seq2Seq_HighLevel_CMA.forward
            from . import torch_tail

            dev = rt.device
            feats = rt.encode(observations["rgb"], observations["depth"], None, n_envs=rnn_hidden_states.shape[1],
                              use_lo_weights=True)
            return torch_tail.lo_tail(self, feats["rgb_gmean"], feats["depth_feat"],
                                      rnn_hidden_states.to(dev, torch.float32), masks.to(dev, torch.float32),
                                      discrete_actions.to(dev))
        return rt.forward_lo(observations["rgb"], observations["depth"], masks, rnn_hidden_states, discrete_actions)
