"""The DAgger parameter update of the HCM agent: ``HierarchicalTrainer._update_agent``
(robo_vln_baselines/hierarchical_trainer.py:492-560) behind the same signature.

What the reference does per call -- hi forward, CrossEntropy on the masked logits, backward, AdamW step; lo forward on a
second device, MSE + BCE-with-logits, backward, Adam step -- is kept; how it runs is not:

* the frozen encoders (RGB ResNet-50, depth ResNet-50, BERT: 99 % of the FLOPs, no gradients) run ONCE on the sm_100a
  engine for both models (``HcmRuntime.encode``; lo reuses hi's trunk features when the two share their frozen trunks),
* each loss is one fused kernel (``losses.py``), each optimizer step one fused launch (``optim.py``),
* the trainable tail's forward + loss + backward of each model (a few hundred small launches under autograd) is captured
  once per trajectory shape in a CUDA graph and replayed (``graph=True``, the default): the step is then bound by the GPU
  work instead of by the host launching it.  Dropout stays live under replay (torch's graph-safe Philox offsets).
* no cuda:0 -> cuda:1 shuffle of the observation batch: both models sit on the runtime's device.

``DaggerUpdater.update`` returns exactly what ``_update_agent`` returns and mutates the caller's ``observations`` dict the
same way (``instruction`` deleted by the hi forward, ``vln_oracle_action_sensor`` replaced by its squeezed int64 form).
"""
from __future__ import annotations

import torch

from . import losses, torch_tail


def repackage_hidden(h):
    """robo_vln_baselines/common/utils.py:159-165"""
    if isinstance(h, torch.Tensor):
        return h.detach()
    return tuple(repackage_hidden(v) for v in h)


class _GraphedTail:
    """forward + loss + backward of one model's trainable tail on static buffers, captured in a CUDA graph."""

    def __init__(self, fn, inputs: dict, params, device):
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in inputs.items()}
        self.params = params
        self.fn = fn
        # warm-up on a side stream (lazy cuDNN / autograd initialisation must not land in the capture)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):
                for p in params:
                    p.grad = None
                outs = fn(**self.static)
                outs[0].backward()
        torch.cuda.current_stream(device).wait_stream(side)
        for p in params:
            p.grad = None          # the captured backward allocates the gradients from the graph's pool: static addresses
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outs = fn(**self.static)
            self.outs[0].backward()

    def replay(self, inputs: dict):
        for k, v in inputs.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.outs


class DaggerUpdater:
    """``update(...)`` = ``HierarchicalTrainer._update_agent(...)`` (hierarchical_trainer.py:492-560).

    high_level / low_level: the drop-in modules (``Seq2Seq_HighLevel_CMA`` / ``Seq2Seq_LowLevel`` of this package);
    optimizer_*: any ``torch.optim.Optimizer`` (the reference's AdamW / Adam, or ``optim.FusedAdamW`` / ``optim.FusedAdam``).
    """

    def __init__(self, high_level, low_level, optimizer_high_level, optimizer_low_level, graph: bool = True):
        self.high_level, self.low_level = high_level, low_level
        self.optimizer_high_level, self.optimizer_low_level = optimizer_high_level, optimizer_low_level
        self.graph = bool(graph)
        self._graphs = {}
        self.graph_errors = []     # capture failures fall back to the eager tail for that shape (kept for inspection)

    # ---- the two differentiable tails + losses, as functions of tensors only (capturable) ----
    def _hi_fn(self, starts):
        hi = self.high_level

        def fn(rgb_feat, depth_feat, bert, hidden, masks, sensor):
            logits, hid = torch_tail.hi_tail(hi, rgb_feat, depth_feat, bert, hidden, masks, hi.dropout_p, starts)
            return losses.hi_loss(logits, sensor), hid
        return fn

    def _lo_fn(self, starts):
        lo = self.low_level

        def fn(rgb_gmean, depth_feat, hidden, masks, sub_goal, corrected, oracle_stop):
            act, stop, hid = torch_tail.lo_tail(lo, rgb_gmean, depth_feat, hidden, masks, sub_goal, starts)
            la, ls = losses.lo_loss(act, stop, corrected, oracle_stop)
            return la + ls, hid, la, ls
        return fn

    def _run(self, which, fn_maker, starts, inputs, params, optimizer, device):
        key = (which, tuple(starts)) + tuple((k, tuple(v.shape), v.dtype) for k, v in inputs.items() if torch.is_tensor(v))
        if self.graph and key not in self._graphs:
            try:
                self._graphs[key] = _GraphedTail(fn_maker(starts), inputs, params, device)
            except Exception as exc:       # e.g. an op that cannot be captured on this torch build: eager from now on
                self.graph_errors.append(f"{which}: {type(exc).__name__}: {exc}"[:300])
                self._graphs[key] = None
                torch.cuda.synchronize(device)
        g = self._graphs.get(key) if self.graph else None
        if g is not None:
            outs = g.replay(inputs)
        else:
            optimizer.zero_grad()
            outs = fn_maker(starts)(**inputs)
            outs[0].backward()
        optimizer.step()
        return outs

    def update(self, observations, prev_actions, not_done_masks, corrected_actions, oracle_stop,
               high_recurrent_hidden_states, low_recurrent_hidden_states, detached_state_low):
        hi, lo = self.high_level, self.low_level
        rt = hi.runtime()
        dev = rt.device
        f32 = lambda t: t.to(dev, torch.float32)  # noqa: E731
        high_recurrent_hidden_states = repackage_hidden(high_recurrent_hidden_states)
        low_recurrent_hidden_states = repackage_hidden(low_recurrent_hidden_states)
        n_envs = high_recurrent_hidden_states.shape[1]
        masks = f32(not_done_masks)
        starts = torch_tail.segment_starts(masks[:, 0], n_envs)      # one host read; part of the graph key

        # ---- frozen encoders, once ----
        feats = rt.encode(observations["rgb"], observations["depth"], observations["instruction"], n_envs=n_envs)
        del observations["instruction"]                               # as Seq2Seq_HighLevel_CMA.forward does (:196)

        # ---- hi: CE on the masked logits (:506-513) ----
        sensor = f32(observations["vln_oracle_action_sensor"]).reshape(-1)
        hi_params = [p for p in hi.parameters() if p.requires_grad]
        outs = self._run("hi", self._hi_fn, starts,
                         dict(rgb_feat=feats["rgb_feat"], depth_feat=feats["depth_feat"], bert=feats["bert"],
                              hidden=f32(high_recurrent_hidden_states), masks=masks, sensor=sensor),
                         hi_params, self.optimizer_high_level, dev)
        high_level_loss_data, hi_hidden = outs[0].detach().clone(), outs[1].detach().clone()
        sensor_i = observations["vln_oracle_action_sensor"].squeeze(1).to(dtype=torch.int64)
        observations["vln_oracle_action_sensor"] = sensor_i           # the reference leaves the int64 form in the caller's dict (:510)

        # ---- lo: MSE on the masked actions + BCE-with-logits on the valid stop rows (:516-555) ----
        discrete_actions = (sensor_i.to(dev) - 1).masked_fill(sensor_i.to(dev) == 0, 4).view(-1)
        lo_rt = lo.runtime()
        feats_lo = lo_rt.encode(observations["rgb"], observations["depth"], None, n_envs=n_envs, use_lo_weights=True)
        lo_params = [p for p in lo.parameters() if p.requires_grad]
        outs = self._run("lo", self._lo_fn, starts,
                         dict(rgb_gmean=feats_lo["rgb_gmean"], depth_feat=feats_lo["depth_feat"],
                              hidden=f32(low_recurrent_hidden_states), masks=masks, sub_goal=discrete_actions,
                              corrected=f32(corrected_actions), oracle_stop=f32(oracle_stop)),
                         lo_params, self.optimizer_low_level, dev)
        lo_hidden = outs[1].detach().clone()
        loss = (high_level_loss_data.item(), outs[2].detach().item(), outs[3].detach().item(), 0)
        return loss, hi_hidden, lo_hidden, detached_state_low


def update_agent(high_level, low_level, optimizer_high_level, optimizer_low_level, *args, graph: bool = False):
    """One-shot functional form of ``_update_agent`` (no graph cache across calls unless the caller keeps a DaggerUpdater)."""
    return DaggerUpdater(high_level, low_level, optimizer_high_level, optimizer_low_level, graph=graph).update(*args)
