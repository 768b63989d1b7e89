"""Fused optimizers for the DAgger update.

The reference trains hi with ``torch.optim.AdamW`` and lo with ``torch.optim.Adam`` (L2 weight decay)
(robo_vln_baselines/hierarchical_trainer.py:329-334), each stepping ~40 small tensors with several launches per
tensor.  ``FusedAdamW`` / ``FusedAdam`` are drop-in ``torch.optim.Optimizer`` subclasses (same constructor arguments,
same ``state_dict`` layout: ``step``, ``exp_avg``, ``exp_avg_sq`` -- checkpoints interchange; LR schedulers such as the
reference's CyclicLR act on ``param_groups`` as usual) whose ``step()`` is ONE launch of ``rvb_fused_adam`` per parameter
group, with torch's single-tensor arithmetic.  CUDA float32 parameters only; anything else raises (no fallback).
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    decoupled = False      # Adam: grad += weight_decay * param

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._lists = {}     # per group: (signature, device pointer tables)

    def _tables(self, gi, tensors):
        """Device arrays of pointers for this group's live (param, grad, exp_avg, exp_avg_sq) tensors, rebuilt only when a
        pointer changes (gradients are re-allocated by zero_grad(set_to_none=True))."""
        sig = tuple(t.data_ptr() for quad in tensors for t in quad)
        hit = self._lists.get(gi)
        if hit is not None and hit[0] == sig:
            return hit[1]
        dev = tensors[0][0].device
        chunk = _lib.load().rvb_adam_chunk_elems()
        numel = [q[0].numel() for q in tensors]
        starts = [0]
        for n in numel:
            starts.append(starts[-1] + (n + chunk - 1) // chunk)
        table = torch.tensor([[q[j].data_ptr() for q in tensors] for j in range(4)] + [numel], dtype=torch.int64).to(dev)
        cs = torch.tensor(starts, dtype=torch.int64).to(dev)
        out = (table, cs, len(tensors), starts[-1])
        self._lists[gi] = (sig, out)
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            live = []
            step = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda" or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("robovln_b200.optim: fused Adam handles dense float32 CUDA parameters only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item()) if torch.is_tensor(st["step"]) else int(st["step"])
                if step is None:
                    step = s
                elif s != step:
                    raise RuntimeError("robovln_b200.optim: parameters of one group must share their step count")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("robovln_b200.optim: non-contiguous parameter / gradient")
                live.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            if not live:
                continue
            table, cs, n, chunks = self._tables(gi, live)
            b1, b2 = group["betas"]
            step_size = group["lr"] / (1.0 - b1 ** step)                   # double precision, as torch computes them
            bc2_sqrt = math.sqrt(1.0 - b2 ** step)
            dev = live[0][0].device
            vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
            with torch.cuda.device(dev):
                _lib.check(lib.rvb_fused_adam(vp(table[0]), vp(table[1]), vp(table[2]), vp(table[3]), vp(table[4]), vp(cs), n, chunks,
                                              group["lr"], b1, b2, group["eps"], group["weight_decay"], int(self.decoupled),
                                              step_size, bc2_sqrt, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                           "rvb_fused_adam", lib)
            # the kernel wrote the parameters (and the moments) behind torch's back: bump their version counters so that
            # everything keyed on (data_ptr, _version) -- the engine's packed-tail refresh (runtime._tail_sig_now), autograd's
            # saved-tensor checks -- sees the update, exactly as after an in-place torch op
            torch.autograd.graph.increment_version([t for quad in live for t in (quad[0], quad[2], quad[3])])
        return loss


class FusedAdamW(FusedAdam):
    decoupled = True       # AdamW: param *= 1 - lr * weight_decay

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
