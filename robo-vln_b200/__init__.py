"""robovln_b200 -- B200-native (sm_100a) implementation of robo-vln's HCM policy forward pass.

The directory is called ``robo-vln_b200`` (repo naming); import it as ``robovln_b200`` (the
alias package at the repo root points its ``__path__`` here).
"""
from .seq2seq_highlevel_cma import Seq2Seq_HighLevel_CMA  # noqa: F401
from .seq2seq_lowlevel import Seq2Seq_LowLevel  # noqa: F401
from .policy import HcmPolicy  # noqa: F401
from . import data, losses, obs, optim, trainer  # noqa: F401

__all__ = ["Seq2Seq_HighLevel_CMA", "Seq2Seq_LowLevel", "HcmPolicy"]
