"""Host-side runtime: owns one C engine (hcm_engine) per hi/lo module pair and device.

PyTorch is used for device memory (weights, workspace, outputs) and for the current CUDA
stream; every computation of the forward pass happens inside librobovln_b200.so.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from . import weight_prep as WP
from ._lib import HcmShape, check

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int64: 2, torch.float16: 3}

# most recent runtime per device still waiting for its other half (hi <-> lo pairing)
_open_runtimes: Dict[Tuple[str, int], "HcmRuntime"] = {}


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class HcmRuntime:
    def __init__(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError(
                "robovln_b200 runs on a CUDA device only (there is no CPU path); move the module with .to('cuda')")
        self.dtype_name = _lib.default_dtype()        # "fp16" (default) or "bf16" via ROBOVLN_DTYPE
        self.lib = _lib.load(dtype=self.dtype_name)
        self.h16 = {"fp16": torch.float16, "bf16": torch.bfloat16}[self.dtype_name]
        self.device = device
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(self.lib.hcm_create(ctypes.byref(self.handle)), "hcm_create")
        self.hi = None          # weakrefs to the nn.Modules
        self.lo = None
        self._tensors: Dict[str, torch.Tensor] = {}   # keeps prepared weights alive
        self._dirty = True
        self._rgb_fmt = 0             # engine-side RGB element type: 0 float32, 1 uint8 (hcm_set_rgb_format)
        self._shares = False
        self._shape_key = None
        self._workspace: Optional[torch.Tensor] = None
        # Trunk-feature reuse (lo after hi on the SAME observations, hierarchical_trainer.py:1096-1100): decided on
        # observation CONTENT -- a 128-bit checksum per frame tensor computed on the device (rvb_checksum) -- never on
        # pointers or version counters.  `trunk_reuse = False` (or ROBOVLN_TRUNK_REUSE=0) always re-runs the trunks.
        self.trunk_reuse = os.environ.get("ROBOVLN_TRUNK_REUSE", "1") != "0"
        self._obs_chk: Optional[torch.Tensor] = None     # int64[4] on the device: content the feature buffers hold
        self._obs_meta = None
        self._feat_lo_weights = False
        self._tail_params = []
        self._tail_sig = None
        # Instruction cache (opt-in: `instruction_cache = True` or ROBOVLN_INSTR_CACHE=1): when the instruction tokens of a
        # call equal those of the previous call, BERT and the query-side projection are skipped (their outputs are still
        # in the engine's buffers).  Rollouts re-send the same instruction at every step of an episode
        # (hierarchical_trainer.py:1193-1196).  bench.py never enables it: the benchmark step always runs BERT.
        self.instruction_cache = os.environ.get("ROBOVLN_INSTR_CACHE", "0") == "1"
        self._bert_key: Optional[torch.Tensor] = None
        self._skip_bert = False

    def __del__(self):
        try:
            if self.handle:
                self.lib.hcm_destroy(self.handle)
        except Exception:
            pass

    # ---- pairing -------------------------------------------------------------------------
    @staticmethod
    def for_module(module, kind: str, device: torch.device) -> "HcmRuntime":
        if device.type != "cuda":
            raise RuntimeError(
                "robovln_b200 runs on a CUDA device only (there is no CPU path); move the module with .to('cuda')")
        key =(device.type, device.index if device.index is not None else torch.cuda.current_device())
        rt = _open_runtimes.get(key)
        if rt is not None and getattr(rt, kind) is None:
            setattr(rt, kind, weakref.ref(module))
            rt._dirty = True
            if rt.hi is not None and rt.lo is not None:
                _open_runtimes.pop(key, None)
            return rt
        rt = HcmRuntime(device)
        setattr(rt, kind, weakref.ref(module))
        _open_runtimes[key] = rt
        return rt

    def mark_dirty(self):
        self._dirty = True

    # ---- weights ---------------------------------------------------------------------------
    def _modules(self):
        hi = self.hi() if self.hi is not None else None
        lo = self.lo() if self.lo is not None else None
        return hi, lo

    def _tail_sig_now(self):
        return tuple((p.data_ptr(), p._version) for p in self._tail_params)

    def _refresh_tails(self):
        """The trainable tail changed (optimizer step, in-place edit) since it was packed: re-pack it INTO the
        engine's existing tensors -- pointers, plans and captured graphs stay valid."""
        hi, lo = self._modules()
        WP.set_h16(self.dtype_name)
        with torch.no_grad(), torch.cuda.device(self.device):
            new = {}
            if hi is not None:
                new.update(WP.prep_hi_tail(hi.state_dict(keep_vars=True), self.device))
            if lo is not None:
                new.update(WP.prep_lo_tail(lo.state_dict(keep_vars=True), self.device))
            for name, t in new.items():
                self._tensors[name].copy_(t)
        self._tail_sig = self._tail_sig_now()
        self._bert_key = None        # the cached query-side projection was made with the old ins_fc / LayerNorm weights

    def sync_weights(self, check_tail: bool = False):
        if not self._dirty:
            if check_tail and self._tail_sig_now() != self._tail_sig:
                self._refresh_tails()
            return
        hi, lo = self._modules()
        dev = self.device
        tensors: Dict[str, torch.Tensor] = {}
        shares = False
        WP.set_h16(self.dtype_name)
        frozen_cache = getattr(self, "_frozen_cache", {})

        def frozen(tag, sd, prefixes, fn):
            """Re-pack a frozen sub-network only if one of its tensors changed (storage or version):
            after an optimizer step only the ~5 M-parameter tail needs new kernel-layout copies."""
            sig = tuple((v.data_ptr(), v._version) for k, v in sd.items() if k.startswith(prefixes))
            hit = frozen_cache.get(tag)
            if hit is not None and hit[0] == sig:
                return hit[1]
            out = fn()
            frozen_cache[tag] = (sig, out)
            return out

        with torch.no_grad():
            sd_hi = hi.state_dict(keep_vars=True) if hi is not None else None
            sd_lo = lo.state_dict(keep_vars=True) if lo is not None else None
            if sd_hi is not None:
                tensors.update(frozen("hi.rgb", sd_hi, ("rgb_encoder.cnn.",), lambda: WP.prep_rgb_trunk(sd_hi, "hi", dev)))
                tensors.update(frozen("hi.depth", sd_hi, ("depth_encoder.visual_encoder.",),
                                      lambda: WP.prep_depth_trunk(sd_hi, "hi", dev)))
                tensors.update(frozen("hi.bert", sd_hi, ("embedding_layer.",), lambda: WP.prep_bert(sd_hi, dev)))
                tensors.update(WP.prep_hi_tail(sd_hi, dev))
            if sd_lo is not None:
                shares = sd_hi is not None and frozen(
                    "shares", {**{"a." + k: v for k, v in sd_hi.items() if k.startswith(WP.TRUNK_PREFIXES)},
                               **{"b." + k: v for k, v in sd_lo.items() if k.startswith(WP.TRUNK_PREFIXES)}},
                    ("a.", "b."), lambda: WP.trunks_identical(sd_hi, sd_lo))
                if not shares:
                    tensors.update(frozen("lo.rgb", sd_lo, ("rgb_encoder.cnn.",), lambda: WP.prep_rgb_trunk(sd_lo, "lo", dev)))
                    tensors.update(frozen("lo.depth", sd_lo, ("depth_encoder.visual_encoder.",),
                                          lambda: WP.prep_depth_trunk(sd_lo, "lo", dev)))
                tensors.update(WP.prep_lo_tail(sd_lo, dev))
        self._frozen_cache = frozen_cache
        with torch.cuda.device(dev):
            torch.cuda.synchronize()
            for name, t in tensors.items():
                shape = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
                check(self.lib.hcm_set_tensor(self.handle, name.encode(), _ptr(t), _DT[t.dtype], t.dim(), shape),
                      f"hcm_set_tensor({name})")
            check(self.lib.hcm_finalize_weights(self.handle, int(hi is not None), int(lo is not None), int(shares)),
                  "hcm_finalize_weights")
        self._tensors = tensors
        self._shares = shares
        self._dirty = False
        self._shape_key = None      # plans capture weight pointers
        self._obs_chk = None
        from .param_spec import FROZEN_PREFIXES

        self._tail_params = [p for m in (hi, lo) if m is not None for n, p in m.named_parameters()
                             if not n.startswith(FROZEN_PREFIXES)]
        self._tail_sig = self._tail_sig_now()

    # ---- planning --------------------------------------------------------------------------
    def ensure_plan(self, B: int, N: int, L: int, instr_rows: int, rgb_hw, depth_hw):
        self.sync_weights(check_tail=True)
        key = (B, N, L, instr_rows, tuple(rgb_hw), tuple(depth_hw))
        if key == self._shape_key:
            return
        shp = HcmShape(B, N, L, instr_rows, rgb_hw[0], rgb_hw[1], depth_hw[0], depth_hw[1])
        with torch.cuda.device(self.device):
            need = self.lib.hcm_workspace_bytes(self.handle, ctypes.byref(shp))
            if need == 0:
                check(-1, "hcm_workspace_bytes")
            if self._workspace is None or self._workspace.numel() < need + 1024:
                self._workspace = None
                self._workspace = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            base = self._workspace.data_ptr()
            aligned = (base + 1023) & ~1023
            torch.cuda.synchronize()
            check(self.lib.hcm_plan(self.handle, ctypes.byref(shp), ctypes.c_void_p(aligned), need), "hcm_plan")
        self._shape_key = key
        self._obs_chk = None
        self._bert_key = None

    def _instruction_hit(self, instr: torch.Tensor) -> bool:
        """Decide whether BERT can be skipped for this call and tell the engine (one small device comparison)."""
        hit = bool(self.instruction_cache and self._bert_key is not None and self._bert_key.shape == instr.shape
                   and self._bert_key.dtype == instr.dtype and self._bert_key.device == instr.device
                   and torch.equal(self._bert_key, instr))
        if hit != self._skip_bert:
            check(self.lib.hcm_set_skip_bert(self.handle, int(hit)), "hcm_set_skip_bert")
            self._skip_bert = hit
        if not hit:
            self._bert_key = instr.clone() if self.instruction_cache else None
        return hit

    # ---- helpers ---------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _prep_obs(self, t: torch.Tensor) -> torch.Tensor:
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def _prep_rgb(self, t: torch.Tensor) -> torch.Tensor:
        """RGB frames: float32 in 0..255 (the reference's batch_obs output) or uint8 exactly as the sensor delivers
        them -- the engine reads either (hcm_set_rgb_format); other dtypes are converted to float32."""
        if t.dtype == torch.uint8:
            if t.device != self.device:
                t = t.to(self.device, non_blocking=True)
            t = t.contiguous()
        else:
            t = self._prep_obs(t)
        fmt = 1 if t.dtype == torch.uint8 else 0
        if fmt != self._rgb_fmt:
            check(self.lib.hcm_set_rgb_format(self.handle, fmt), "hcm_set_rgb_format")
            self._rgb_fmt = fmt
        return t

    def _checksum_obs(self, rgb: torch.Tensor, depth: torch.Tensor):
        """(int64[4] device tensor, meta) -- content checksum of the two frame tensors (asynchronous), or (None, None)
        when reuse is disabled or a tensor is not 16-byte aligned."""
        if not self.trunk_reuse or (rgb.data_ptr() | depth.data_ptr()) & 15:
            return None, None
        out = torch.empty((4,), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            st = self._stream()
            check(self.lib.rvb_checksum(_ptr(rgb), rgb.numel() * rgb.element_size(), ctypes.c_void_p(out.data_ptr()), st),
                  "rvb_checksum")
            check(self.lib.rvb_checksum(_ptr(depth), depth.numel() * depth.element_size(),
                                        ctypes.c_void_p(out.data_ptr() + 16), st), "rvb_checksum")
        return out, (tuple(rgb.shape), rgb.dtype, tuple(depth.shape))

    def _same_obs(self, rgb: torch.Tensor, depth: torch.Tensor) -> bool:
        """Do the engine's trunk-feature buffers hold the features of exactly these frames?  (one 32-byte
        device->host comparison; only asked on the module-API path where hi and lo are called separately)"""
        if self._obs_chk is None:
            return False
        chk, meta = self._checksum_obs(rgb, depth)
        return chk is not None and meta == self._obs_meta and bool(torch.equal(chk, self._obs_chk))

    def launches(self) -> int:
        return int(self.lib.hcm_last_launch_count(self.handle))

    # ---- forward ---------------------------------------------------------------------------
    def forward_hi(self, rgb, depth, instruction, masks, hidden):
        rgb, depth = self._prep_rgb(rgb), self._prep_obs(depth)
        B = rgb.shape[0]
        N = hidden.shape[1]
        instr = instruction
        if instr.dim() != 2:
            raise ValueError("instruction must be [1 or B, L]")
        if instr.shape[0] not in (1, B):
            raise ValueError(f"instruction has {instr.shape[0]} rows; expected 1 or {B}")
        instr = instr.to(self.device)
        i_f32 = instr.float().contiguous() if instr.dtype != torch.int64 else None
        i_i64 = instr.contiguous() if instr.dtype == torch.int64 else None
        masks = masks.to(self.device, torch.float32)
        hidden = hidden.to(self.device, torch.float32).contiguous()
        self.ensure_plan(B, N, instr.shape[1], instr.shape[0], rgb.shape[1:3], depth.shape[1:3])
        self._instruction_hit(instr)
        logits = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        hc_out = torch.empty_like(hidden)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_forward_hi(self.handle, _ptr(rgb), _ptr(depth), _ptr(i_f32), _ptr(i_i64), _ptr(masks),
                                          masks.stride(0), _ptr(hidden), _ptr(logits), _ptr(hc_out), self._stream()),
                  "hcm_forward_hi")
        # lo may follow on the same frames: remember WHAT the feature buffers now hold (content, not pointers)
        self._obs_chk, self._obs_meta = self._checksum_obs(rgb, depth) if (self._shares and self.lo is not None) else (None, None)
        self._feat_lo_weights = False
        self._keep = (rgb, depth, i_f32, i_i64, masks, hidden)   # alive until the stream consumed them
        return logits, hc_out

    def forward_lo(self, rgb, depth, masks, hidden, sub_goal):
        rgb, depth = self._prep_rgb(rgb), self._prep_obs(depth)
        B = rgb.shape[0]
        N = hidden.shape[1]
        masks = masks.to(self.device, torch.float32)
        hidden = hidden.to(self.device, torch.float32).contiguous()
        sub_goal = sub_goal.to(self.device, torch.int64).contiguous().view(-1)
        self.sync_weights(check_tail=True)
        reuse = bool(self._shares and self._shape_key is not None and self._shape_key[0] == B
                     and self._shape_key[1] == N and self._same_obs(rgb, depth))
        if not reuse:
            hi, _ = self._modules()
            if self._shape_key is not None and self._shape_key[0] == B and self._shape_key[1] == N \
                    and self._shape_key[4] == tuple(rgb.shape[1:3]):
                pass  # the current plan already fits (instruction length is irrelevant for lo)
            else:
                self.ensure_plan(B, N, 8 if hi is not None else 1, 1, rgb.shape[1:3], depth.shape[1:3])
        act = torch.empty((B, 2), dtype=torch.float32, device=self.device)
        stop = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        hc_out = torch.empty_like(hidden)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_forward_lo(self.handle, _ptr(rgb), _ptr(depth), _ptr(masks), masks.stride(0),
                                          _ptr(sub_goal), _ptr(hidden), _ptr(act), _ptr(stop), _ptr(hc_out),
                                          int(reuse), self._stream()), "hcm_forward_lo")
        if not reuse:       # the trunks just ran on these frames (with lo's weights)
            self._obs_chk, self._obs_meta = self._checksum_obs(rgb, depth)
            self._feat_lo_weights = True
        self._keep_lo = (rgb, depth, masks, hidden, sub_goal)
        return act, stop, hc_out

    def forward_policy(self, rgb, depth, instruction, masks, hidden_hi, hidden_lo):
        """hi -> argmax -> lo in one engine call (rollout step, hierarchical_trainer.py:1095-1101)."""
        rgb, depth = self._prep_rgb(rgb), self._prep_obs(depth)
        B, N = rgb.shape[0], hidden_hi.shape[1]
        instr = instruction.to(self.device)
        i_f32 = instr.float().contiguous() if instr.dtype != torch.int64 else None
        i_i64 = instr.contiguous() if instr.dtype == torch.int64 else None
        masks = masks.to(self.device, torch.float32)
        hidden_hi = hidden_hi.to(self.device, torch.float32).contiguous()
        hidden_lo = hidden_lo.to(self.device, torch.float32).contiguous()
        self.ensure_plan(B, N, instr.shape[1], instr.shape[0], rgb.shape[1:3], depth.shape[1:3])
        self._instruction_hit(instr)
        logits = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        act = torch.empty((B, 2), dtype=torch.float32, device=self.device)
        stop = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        sub = torch.empty((B,), dtype=torch.int64, device=self.device)
        hc_hi = torch.empty_like(hidden_hi)
        hc_lo = torch.empty_like(hidden_lo)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_forward_policy(self.handle, _ptr(rgb), _ptr(depth), _ptr(i_f32), _ptr(i_i64), _ptr(masks),
                                              masks.stride(0), _ptr(hidden_hi), _ptr(hidden_lo), _ptr(logits), _ptr(act),
                                              _ptr(stop), _ptr(hc_hi), _ptr(hc_lo), _ptr(sub), self._stream()),
                  "hcm_forward_policy")
        self._obs_chk = None      # feature buffers overwritten; content not tracked on this path
        self._keep = (rgb, depth, i_f32, i_i64, masks, hidden_hi, hidden_lo)
        return logits, act, stop, hc_hi, hc_lo, sub

    def encode(self, rgb, depth, instruction=None, n_envs: int = 1, use_lo_weights: bool = False):
        """Frozen encoders only (training path): returns fp32 feature tensors
        {"rgb_feat" [B,16,2048], "rgb_gmean" [B,2048], "depth_feat" [B,16,128], "bert" [1|B,L,768]}."""
        rgb, depth = self._prep_rgb(rgb), self._prep_obs(depth)
        B = rgb.shape[0]
        self.sync_weights()          # may invalidate the current plan (new weights / newly attached half)
        with_bert = instruction is not None
        i_f32 = i_i64 = None
        if with_bert:
            instr = instruction.to(self.device)
            if instr.dim() != 2 or instr.shape[0] not in (1, B):
                raise ValueError("instruction must be [1 or B, L]")
            i_f32 = instr.float().contiguous() if instr.dtype != torch.int64 else None
            i_i64 = instr.contiguous() if instr.dtype == torch.int64 else None
            L, rows = instr.shape[1], instr.shape[0]
        else:
            L, rows = (self._shape_key[2], self._shape_key[3]) if (self._shape_key and self._shape_key[0] == B) else (8, 1)
        N = n_envs if B % max(n_envs, 1) == 0 else 1
        if not (self._shape_key and self._shape_key[0] == B and self._shape_key[4] == tuple(rgb.shape[1:3])
                and (not with_bert or (self._shape_key[2], self._shape_key[3]) == (L, rows))):
            self.ensure_plan(B, N, L, rows, rgb.shape[1:3], depth.shape[1:3])
        self._no_instruction_cache()
        fresh = not ((self._shares or self._feat_lo_weights == bool(use_lo_weights)) and self._same_obs(rgb, depth))
        if fresh or with_bert:
            with torch.cuda.device(self.device):
                check(self.lib.hcm_run_encoders(self.handle, _ptr(rgb), _ptr(depth), _ptr(i_f32), _ptr(i_i64),
                                                int(with_bert), int(use_lo_weights), self._stream()), "hcm_run_encoders")
            self._obs_chk, self._obs_meta = self._checksum_obs(rgb, depth)
            self._feat_lo_weights = bool(use_lo_weights)
            self._keep = (rgb, depth, i_f32, i_i64)
        out = {
            "rgb_feat": self.get_buffer("rgb_tokens", sync=False)[:, :, :2048].float(),
            "rgb_gmean": self.get_buffer("rgb_gmean", sync=False).float(),
            "depth_feat": self.get_buffer("depth_tokens", sync=False)[:, :, :128].float(),
        }
        if with_bert:
            out["bert"] = self.get_buffer("bert", sync=False).float()
        return out

    def _no_instruction_cache(self):
        """Entry points outside the cache protocol (encode / profile / stage runs): BERT always runs and the cached key is dropped."""
        if self._skip_bert:
            check(self.lib.hcm_set_skip_bert(self.handle, 0), "hcm_set_skip_bert")
            self._skip_bert = False
        self._bert_key = None

    def profile_policy(self, rgb, depth, instruction, masks, hidden_hi, hidden_lo):
        """Per-launch device times of one policy step (single stream, CUDA events between
        launches): list of {"name", "ms", "flops"}."""
        import json

        rgb, depth = self._prep_rgb(rgb), self._prep_obs(depth)
        B, N = rgb.shape[0], hidden_hi.shape[1]
        instr = instruction.to(self.device)
        i_f32 = instr.float().contiguous() if instr.dtype != torch.int64 else None
        i_i64 = instr.contiguous() if instr.dtype == torch.int64 else None
        masks = masks.to(self.device, torch.float32)
        hidden_hi = hidden_hi.to(self.device, torch.float32).contiguous()
        hidden_lo = hidden_lo.to(self.device, torch.float32).contiguous()
        self.ensure_plan(B, N, instr.shape[1], instr.shape[0], rgb.shape[1:3], depth.shape[1:3])
        self._no_instruction_cache()
        logits = torch.empty((B, 4), dtype=torch.float32, device=self.device)
        act = torch.empty((B, 2), dtype=torch.float32, device=self.device)
        stop = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        hc_hi = torch.empty_like(hidden_hi)
        hc_lo = torch.empty_like(hidden_lo)
        cap = 1 << 20
        buf = ctypes.create_string_buffer(cap)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_profile_policy(self.handle, _ptr(rgb), _ptr(depth), _ptr(i_f32), _ptr(i_i64), _ptr(masks),
                                              masks.stride(0), _ptr(hidden_hi), _ptr(hidden_lo), _ptr(logits), _ptr(act),
                                              _ptr(stop), _ptr(hc_hi), _ptr(hc_lo), buf, cap, self._stream()),
                  "hcm_profile_policy")
        self._obs_chk = None
        return json.loads(buf.value.decode())

    def forward_policy_host(self, rgb, depth, instruction, masks, hidden_hi, hidden_lo, out=None):
        """Host-buffer entry: all arguments are CPU float32 tensors (pinned for speed); H2D copies,
        the forward and the D2H copies of the results are inside this call."""
        B, N = rgb.shape[0], hidden_hi.shape[1]
        for t in (depth, instruction, masks, hidden_hi, hidden_lo):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("forward_policy_host expects contiguous float32 CPU tensors")
        if rgb.device.type != "cpu" or rgb.dtype not in (torch.float32, torch.uint8) or not rgb.is_contiguous():
            raise ValueError("forward_policy_host expects a contiguous float32 or uint8 CPU tensor for rgb")
        fmt = 1 if rgb.dtype == torch.uint8 else 0
        if fmt != self._rgb_fmt:
            check(self.lib.hcm_set_rgb_format(self.handle, fmt), "hcm_set_rgb_format")
            self._rgb_fmt = fmt
        if masks.dim() != 2 or masks.shape[1] != 2:
            raise ValueError("masks must be [B,2]")
        self.ensure_plan(B, N, instruction.shape[1], instruction.shape[0], rgb.shape[1:3], depth.shape[1:3])
        self._instruction_hit(instruction)        # host tensors: compared on the CPU
        if out is None:
            out = {
                "logits": torch.empty((B, 4), dtype=torch.float32).pin_memory(),
                "actions": torch.empty((B, 2), dtype=torch.float32).pin_memory(),
                "stop": torch.empty((B, 1), dtype=torch.float32).pin_memory(),
                "hidden_hi": torch.empty((2, N, 512), dtype=torch.float32).pin_memory(),
                "hidden_lo": torch.empty((2, N, 512), dtype=torch.float32).pin_memory(),
            }
        with torch.cuda.device(self.device):
            check(self.lib.hcm_forward_policy_host(self.handle, _ptr(rgb), _ptr(depth), _ptr(instruction), _ptr(masks),
                                                   _ptr(hidden_hi), _ptr(hidden_lo), _ptr(out["logits"]),
                                                   _ptr(out["actions"]), _ptr(out["stop"]), _ptr(out["hidden_hi"]),
                                                   _ptr(out["hidden_lo"]), self._stream()), "hcm_forward_policy_host")
        self._obs_chk = None
        return out

    def cross_modal(self, bert: torch.Tensor, rgb_spatial: torch.Tensor, depth_spatial: torch.Tensor) -> torch.Tensor:
        """Visual_Ling_Attn for both modalities + token mean-pool on caller tensors (BASELINE.json configs[2]):
        bert [1|B, L, 768], rgb_spatial / depth_spatial [B, 16, 256] (the rgb_kv / depth_kv outputs, cell-major)
        -> [B, 512] = (ins_rgb_att | ins_depth_att), transformer.py:262-281 + seq2seq_highlevel_cma.py:200-210."""
        B, L, rows = rgb_spatial.shape[0], bert.shape[1], bert.shape[0]
        if rows not in (1, B) or tuple(rgb_spatial.shape[1:]) != (16, 256) or tuple(depth_spatial.shape) != tuple(rgb_spatial.shape):
            raise ValueError("cross_modal: bert [1|B,L,768], rgb/depth spatial [B,16,256]")
        hw = (self._shape_key[4], self._shape_key[5]) if self._shape_key else ((256, 256), (256, 256))
        self.ensure_plan(B, B, L, rows, hw[0], hw[1])
        self._no_instruction_cache()
        b16 = bert.to(self.device, self.h16).contiguous()
        r16 = rgb_spatial.to(self.device, self.h16).contiguous()
        d16 = depth_spatial.to(self.device, self.h16).contiguous()
        out = torch.empty((B, 512), dtype=self.h16, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_run_cross_modal(self.handle, _ptr(b16), _ptr(r16), _ptr(d16), _ptr(out), self._stream()),
                  "hcm_run_cross_modal")
        self._keep = (b16, r16, d16)
        return out

    def encode_bert(self, instruction: torch.Tensor, B: int, n_envs: int = 1) -> torch.Tensor:
        """BERT only (pre-computed visual features path): instruction [1|B, L] -> fp32 [1|B, L, 768]."""
        instr = instruction.to(self.device)
        if instr.dim() != 2 or instr.shape[0] not in (1, B):
            raise ValueError("instruction must be [1 or B, L]")
        self.sync_weights()
        hw = (self._shape_key[4], self._shape_key[5]) if self._shape_key else ((256, 256), (256, 256))
        N = n_envs if B % max(n_envs, 1) == 0 else 1
        self.ensure_plan(B, N, instr.shape[1], instr.shape[0], hw[0], hw[1])
        self._no_instruction_cache()
        i_f32 = instr.float().contiguous() if instr.dtype != torch.int64 else None
        i_i64 = instr.contiguous() if instr.dtype == torch.int64 else None
        with torch.cuda.device(self.device):
            check(self.lib.hcm_run_bert(self.handle, _ptr(i_f32), _ptr(i_i64), self._stream()), "hcm_run_bert")
        self._keep = (i_f32, i_i64)
        return self.get_buffer("bert").float()

    def get_buffer(self, name: str, sync: bool = True) -> torch.Tensor:
        """Copy of an internal stage buffer (parity tests; sync=False: ordered on the current stream only, for callers
        that keep working on that stream)."""
        ptr = ctypes.c_void_p()
        dt = ctypes.c_int()
        nd = ctypes.c_int()
        shape = (ctypes.c_int64 * 8)()
        check(self.lib.hcm_get_buffer(self.handle, name.encode(), ctypes.byref(ptr), ctypes.byref(dt), ctypes.byref(nd),
                                      shape), f"hcm_get_buffer({name})")
        shp = [int(shape[i]) for i in range(nd.value)]
        dtype = {0: torch.float32, 1: torch.bfloat16, 2: torch.int64, 3: torch.float16}[dt.value]
        numel = 1
        for s in shp:
            numel *= s
        out = torch.empty(shp, dtype=dtype, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.hcm_copy_buffer(self.handle, name.encode(), _ptr(out), numel * out.element_size(),
                                           self._stream()), f"hcm_copy_buffer({name})")
            if sync:
                torch.cuda.synchronize()
        return out
