"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* robo-vln reference modules.

This file imports the reference's own model files from ``/root/reference`` behind a
thin layer of stub modules (gym / habitat / yacs are not installed in the build
container).  It exists for two purposes only:

  * ``oracle/make_golden.py`` uses it to generate the committed fixtures under
    ``tests/golden/`` (so the restatement in ``oracle/hcm_oracle.py`` and the CUDA path
    are pinned against the real reference arithmetic), and
  * ``tests/test_oracle_vs_reference.py`` uses it (skipped when ``/root/reference`` is
    absent, i.e. on the GPU box) to re-check the restatement live.

Nothing in the product path (``robo-vln_b200/``) may import this module.

No reference file is edited; the only patches are the ones the reference needs to run
at all on a CPU-only box without network access:

  1. ``BertModel.from_pretrained`` -> ``BertModel(BertConfig())`` (no HF cache here),
     ``torchvision.models.resnet50(pretrained=True)`` -> ``resnet50(weights=None)``;
  2. ``DEPTH_ENCODER.ddppo_checkpoint = "NONE"``;
  3. ``torch.Tensor.get_device`` returns ``tensor.device`` so that
     ``robo_vln_baselines/models/transformer/transformer.py:271-273`` works on CPU;
  4. the RGB observation space is declared 224x224 (a different size raises NameError in
     ``robo_vln_baselines/models/encoders/resnet_encoders.py:130-135``) -- the CNN itself
     is size-agnostic, 256x256 tensors are fed at call time.
"""
from __future__ import annotations

import copy
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ROBOVLN_REFERENCE", "/root/reference")
HL = os.path.join(REF_ROOT, "environments", "habitat-lab")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "robo_vln_baselines", "models"))


class _AttrDict(dict):
    """Minimal yacs.config.CfgNode stand-in (attribute access, no-op freeze)."""

    def __init__(self, init=None, *a, **k):
        super().__init__()
        if isinstance(init, dict):
            for key, val in init.items():
                self[key] = _AttrDict(val) if isinstance(val, dict) and not isinstance(val, _AttrDict) else val

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def defrost(self):
        pass

    def freeze(self):
        pass

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_file(self, *_):
        pass

    def merge_from_list(self, *_):
        pass

    def merge_from_other_cfg(self, *_):
        pass


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load_by_path(name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


_LOADED = None


def load_reference():
    """Return a namespace with the reference classes and the default MODEL config."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    import numpy as np
    import torch
    import torch.nn as nn

    # ---- gym ---------------------------------------------------------------
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

    class Space:  # only used as a type annotation
        pass

    spaces_mod = _mod("gym.spaces", Box=Box, Dict=Dict, Space=Space)
    _mod("gym", spaces=spaces_mod, Space=Space)

    # ---- habitat -----------------------------------------------------------
    import logging

    logger = logging.getLogger("habitat-stub")
    habitat = _mod("habitat", Config=_AttrDict, logger=logger)
    habitat.__path__ = []
    _mod("habitat.tasks").__path__ = []
    _mod("habitat.tasks.nav").__path__ = []
    sensor_names = [
        "EpisodicCompassSensor", "EpisodicGPSSensor", "HeadingSensor", "ImageGoalSensor",
        "IntegratedPointGoalGPSAndCompassSensor", "PointGoalSensor", "ProximitySensor",
    ]
    _mod("habitat.tasks.nav.nav", **{n: type(n, (), {"cls_uuid": n.lower()}) for n in sensor_names})
    _mod("habitat.tasks.nav.object_nav_task", ObjectGoalSensor=type("ObjectGoalSensor", (), {"cls_uuid": "objectgoal"}))
    _mod("habitat.utils").__path__ = []
    _mod("habitat.utils.visualizations").__path__ = []
    _mod("habitat.utils.visualizations.utils", images_to_video=lambda *a, **k: None)

    # ---- yacs --------------------------------------------------------------
    _mod("yacs").__path__ = []
    _mod("yacs.config", CfgNode=_AttrDict)

    # ---- habitat_extensions.config.default ----------------------------------
    _mod("habitat_extensions").__path__ = []
    _mod("habitat_extensions.config").__path__ = []
    _mod("habitat_extensions.config.default", get_extended_config=lambda *a, **k: _AttrDict())

    # ---- habitat_baselines namespace with the REAL model-side files ----------
    for pkg in ["habitat_baselines", "habitat_baselines.common", "habitat_baselines.rl",
                "habitat_baselines.rl.models", "habitat_baselines.rl.ddppo",
                "habitat_baselines.rl.ddppo.policy"]:
        _mod(pkg).__path__ = []

    class _Registry:
        @staticmethod
        def register_policy(x=None, **k):
            return x if x is not None else (lambda y: y)

        register_trainer = register_env = register_policy

    _mod("habitat_baselines.common.baseline_registry", baseline_registry=_Registry())
    _mod("habitat_baselines.common.tensorboard_utils", TensorboardWriter=object)
    _mod("habitat_baselines.rl.ppo", Net=nn.Module, Policy=nn.Module)

    hb = os.path.join(HL, "habitat_baselines")
    _load_by_path("habitat_baselines.common.utils", os.path.join(hb, "common", "utils.py"))
    _load_by_path("habitat_baselines.rl.models.rnn_state_encoder", os.path.join(hb, "rl", "models", "rnn_state_encoder.py"))
    _load_by_path("habitat_baselines.rl.models.simple_cnn", os.path.join(hb, "rl", "models", "simple_cnn.py"))
    resnet = _load_by_path("habitat_baselines.rl.ddppo.policy.resnet", os.path.join(hb, "rl", "ddppo", "policy", "resnet.py"))
    sys.modules["habitat_baselines.rl.ddppo.policy"].resnet = resnet
    _load_by_path("habitat_baselines.rl.ddppo.policy.running_mean_and_var", os.path.join(hb, "rl", "ddppo", "policy", "running_mean_and_var.py"))
    _load_by_path("habitat_baselines.rl.ddppo.policy.resnet_policy", os.path.join(hb, "rl", "ddppo", "policy", "resnet_policy.py"))

    # ---- robo_vln_baselines as a namespace package (its __init__ needs habitat_sim) ----
    rvb = _mod("robo_vln_baselines")
    rvb.__path__ = [os.path.join(REF_ROOT, "robo_vln_baselines")]

    # ---- patches (1) and (3) -------------------------------------------------
    import torchvision.models as tvm
    import transformers

    _orig_resnet50 = tvm.resnet50
    tvm.resnet50 = lambda *a, **k: _orig_resnet50(weights=None)

    class _BertNoDownload(transformers.BertModel):
        @classmethod
        def from_pretrained(cls, *a, **k):
            return transformers.BertModel(transformers.BertConfig())

    transformers.BertModel = _BertNoDownload
    torch.Tensor.get_device = lambda t: t.device

    import importlib

    cfg_default = importlib.import_module("robo_vln_baselines.config.default")
    hi_mod = importlib.import_module("robo_vln_baselines.models.seq2seq_highlevel_cma")
    lo_mod = importlib.import_module("robo_vln_baselines.models.seq2seq_lowlevel")
    tr_mod = importlib.import_module("robo_vln_baselines.models.transformer.transformer")
    hi_mod.BertModel = _BertNoDownload   # the module bound the name at import time

    model_cfg = cfg_default._C.MODEL.clone()
    model_cfg.TORCH_GPU_ID = 0
    model_cfg.DEPTH_ENCODER.ddppo_checkpoint = "NONE"       # patch (2)

    space = Dict({
        "rgb": Box(0, 255, (224, 224, 3), np.uint8),         # patch (4)
        "depth": Box(0.0, 1.0, (256, 256, 1), np.float32),
    })

    ns = types.SimpleNamespace(
        Seq2Seq_HighLevel_CMA=hi_mod.Seq2Seq_HighLevel_CMA,
        Seq2Seq_LowLevel=lo_mod.Seq2Seq_LowLevel,
        Visual_Ling_Attn=tr_mod.Visual_Ling_Attn,
        model_cfg=model_cfg,
        observation_space=space,
        Box=Box, Dict=Dict,
    )
    _LOADED = ns
    return ns


def build_reference_models(seed: int = 0):
    """Instantiate the reference hi / lo modules (eval mode, default-constructed weights)."""
    import torch

    ns = load_reference()
    torch.manual_seed(seed)
    hi = ns.Seq2Seq_HighLevel_CMA(ns.observation_space, 4, ns.model_cfg, 1).eval()
    lo = ns.Seq2Seq_LowLevel(ns.observation_space, 2, 4, ns.model_cfg, 1).eval()
    return hi, lo
