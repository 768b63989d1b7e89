"""Fused losses of the DAgger update (robo_vln_baselines/hierarchical_trainer.py:498-553).

``hi_loss`` / ``lo_loss`` compute exactly what ``_update_agent`` computes with ~25 torch launches per model --
masking included -- as one kernel each that returns the loss value(s) and, for backward, the gradient w.r.t. the model
outputs (``rvb_hi_loss`` / ``rvb_lo_loss``).  CUDA float32 tensors only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _HiLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, oracle):
        if logits.device.type != "cuda" or logits.dtype != torch.float32 or logits.dim() != 2:
            raise RuntimeError("hi_loss: logits must be a float32 CUDA tensor [T, C]")
        logits = logits.contiguous()
        oracle = oracle.reshape(-1).contiguous()
        if oracle.dtype not in (torch.float32, torch.int64) or oracle.numel() != logits.shape[0]:
            raise RuntimeError("hi_loss: oracle must be float32 or int64 with one entry per row")
        lib = _lib.load()
        out = torch.empty((2,), dtype=torch.float32, device=logits.device)
        grad = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            _lib.check(lib.rvb_hi_loss(_p(logits), _p(oracle if oracle.dtype == torch.float32 else None),
                                       _p(oracle if oracle.dtype == torch.int64 else None), logits.shape[0], logits.shape[1], _p(out),
                                       _p(grad), _stream(logits.device)), "rvb_hi_loss", lib)
        ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


class _LoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, actions, stop, corrected, oracle_stop):
        for t in (actions, stop, corrected, oracle_stop):
            if t.device.type != "cuda" or t.dtype != torch.float32:
                raise RuntimeError("lo_loss: float32 CUDA tensors expected")
        actions, corrected = actions.contiguous(), corrected.contiguous()
        stop_f, ostop = stop.reshape(-1).contiguous(), oracle_stop.reshape(-1).contiguous()
        T, A = actions.shape
        if corrected.shape != actions.shape or stop_f.numel() != T or ostop.numel() != T:
            raise RuntimeError("lo_loss: shape mismatch")
        lib = _lib.load()
        out = torch.empty((3,), dtype=torch.float32, device=actions.device)
        d_act, d_stop = torch.empty_like(actions), torch.empty_like(stop_f)
        with torch.cuda.device(actions.device):
            _lib.check(lib.rvb_lo_loss(_p(actions), _p(corrected), _p(stop_f), _p(ostop), T, A, _p(out), _p(d_act), _p(d_stop),
                                       _stream(actions.device)), "rvb_lo_loss", lib)
        ctx.save_for_backward(d_act, d_stop.view(stop.shape))
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_act, g_stop):
        d_act, d_stop = ctx.saved_tensors
        # the kernel's gradients are those of (action loss + stop loss); each output carries its own upstream factor
        return d_act * g_act, d_stop * g_stop, None, None


def hi_loss(logits: torch.Tensor, oracle_action_sensor: torch.Tensor) -> torch.Tensor:
    """= CrossEntropyLoss(ignore_index=-1)(logits.masked_fill_(sensor == 0, 0), sensor - 1), hierarchical_trainer.py:506-511
    (the input logits are not modified)."""
    return _HiLoss.apply(logits, oracle_action_sensor)


def lo_loss(actions: torch.Tensor, stop_logit: torch.Tensor, corrected_actions: torch.Tensor, oracle_stop: torch.Tensor):
    """-> (action loss, stop loss) = (MSELoss()(actions.masked_fill_(corrected == 0, 0), corrected),
    BCEWithLogitsLoss()(stop[oracle_stop != -1], oracle_stop[oracle_stop != -1])), hierarchical_trainer.py:539-553."""
    return _LoLoss.apply(actions, stop_logit, corrected_actions, oracle_stop)
