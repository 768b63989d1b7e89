"""The CPU oracle (oracle/hcm_oracle.py) against fixtures produced by the unmodified
reference (oracle/make_golden.py).  fp32 CPU vs fp32 CPU: tolerance 2e-4 abs on O(1..4)
values (different op ordering inside F.* vs nn.Module paths is the only source of drift)."""
import os

import numpy as np
import pytest
import torch

from oracle import hcm_oracle as O
from oracle import weights as W
from oracle.make_golden import CASES, REDUCED, row_stats

TOL = 2e-4


@pytest.fixture(scope="module")
def sds():
    return W.make_state_dict("hi", 0), W.make_state_dict("lo", 0)


def _close(a, b, name, tol=TOL):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    err = float(np.abs(a - b).max())
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert err <= tol, f"{name}: max abs err {err:.3e} > {tol}"


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_golden(case, sds, golden_dir):
    hi_sd, lo_sd = sds
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    kw = CASES[case]
    inp = W.make_inputs(**kw)
    with torch.no_grad():
        logits, hid, it = O.hi_forward(hi_sd, inp["rgb"], inp["depth"], inp["instruction"],
                                       inp["hidden_hi"], inp["masks"], return_intermediates=True)
        act, stop, hid_lo, il = O.lo_forward(lo_sd, inp["rgb"], inp["depth"], inp["hidden_lo"],
                                             inp["masks"], inp["sub_goal"], return_intermediates=True)
    B = inp["rgb"].shape[0]
    if case in REDUCED:      # full-size case: the big intermediates are stored as per-row (mean, mean |x|)
        st = lambda t: row_stats(t.detach().numpy())                                              # noqa: E731
        _close(st(it["depth_embedding"]), gold["hi.depth_embedding.stats"], "hi.depth_embedding.stats")
        _close(st(it["rgb_embedding"]), gold["hi.rgb_embedding.stats"], "hi.rgb_embedding.stats")
        _close(st(it["bert"]), gold["hi.bert.stats"], "hi.bert.stats")
        _close(st(il["depth_embedding"]), gold["lo.depth_embedding.stats"], "lo.depth_embedding.stats")
        _close(st(il["rgb_embedding"]), gold["lo.rgb_embedding.stats"], "lo.rgb_embedding.stats")
    else:
        _close(it["depth_embedding"].view(B, 192, 4, 4), gold["hi.depth_embedding"], "hi.depth_embedding")
        _close(it["rgb_embedding"].view(B, 2112, 4, 4), gold["hi.rgb_embedding"], "hi.rgb_embedding")
        _close(it["bert"], gold["hi.bert"], "hi.bert")
        _close(it["ins_rgb_att"], gold["hi.ins_rgb_att_tokens"].mean(axis=1), "ins_rgb_att")
        _close(it["ins_depth_att"], gold["hi.ins_depth_att_tokens"].mean(axis=1), "ins_depth_att")
        _close(il["depth_embedding"], gold["lo.depth_embedding"], "lo.depth_embedding")
        _close(il["rgb_embedding"], gold["lo.rgb_embedding"], "lo.rgb_embedding")
    _close(it["rnn_in"], gold["hi.rnn_in"], "hi.rnn_in")
    _close(it["rnn_out"], gold["hi.rnn_out"], "hi.rnn_out")
    _close(logits, gold["hi.logits"], "hi.logits")
    _close(hid, gold["hi.hidden"], "hi.hidden", tol=TOL if case not in REDUCED else 2e-3)   # |c| reaches 23 after 64 steps
    _close(il["rnn_in"], gold["lo.rnn_in"], "lo.rnn_in")
    _close(act, gold["lo.actions"], "lo.actions")
    _close(stop, gold["lo.stop"], "lo.stop")
    _close(hid_lo, gold["lo.hidden"], "lo.hidden", tol=TOL if case not in REDUCED else 2e-3)


def test_golden_features_are_nontrivial(golden_dir):
    """Guards against a degenerate fixture (dead ReLUs / all-zero trunk features)."""
    g = np.load(os.path.join(golden_dir, "cfg1_b2_l20.npz"))
    rgb = g["hi.rgb_embedding"][:, :2048]
    dep = g["hi.depth_embedding"][:, :128]
    assert (rgb > 0).mean() > 0.2 and rgb.std() > 0.05
    assert (dep > 0).mean() > 0.2 and dep.std() > 0.05
    assert np.abs(g["hi.logits"][0] - g["hi.logits"][1]).max() > 1e-3


def test_16bit_operand_rounding_floor(sds):
    """How much error 16-bit OPERANDS alone cost on this path, measured with PyTorch's own mixed-precision kernels on the
    CPU (torch.autocast: conv / linear / matmul inputs and outputs rounded to the 16-bit type, fp32 accumulation and
    fp32 normalisation layers -- the same roundings the engine performs), against the fp32 oracle on BASELINE cfg1.

    fp16 (11-bit significand) stays well inside north_star's 1e-2; bf16 (8-bit significand) does NOT: the recurrent state
    is off by ~2.5e-2 with PyTorch's own bf16 kernels.  The 1e-2 "bf16" tolerance is therefore a property of the number
    format on this 100+-layer path, not of an implementation -- which is why the engine's default operand type is fp16
    (same width, same tensor-core rate) and why the bf16 build is held to the bf16 floor measured here
    (tests/test_parity_gpu.py::test_bf16_build_end_to_end)."""
    hi_sd, lo_sd = sds
    inp = W.make_inputs(**CASES["cfg1_b2_l20"])

    def run():
        with torch.no_grad():
            logits, hh = O.hi_forward(hi_sd, inp["rgb"], inp["depth"], inp["instruction"], inp["hidden_hi"], inp["masks"])
            act, stop, hl = O.lo_forward(lo_sd, inp["rgb"], inp["depth"], inp["hidden_lo"], inp["masks"], inp["sub_goal"])
        return {"logits": logits.float(), "hidden_hi": hh.float(), "actions": act.float(), "stop": stop.float(), "hidden_lo": hl.float()}

    ref = run()
    errs = {}
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        with torch.autocast("cpu", dtype=dt):
            got = run()
        errs[name] = {k: float((got[k] - ref[k]).abs().max()) for k in ref}
    print("16-bit operand rounding floor (torch.autocast on the CPU oracle):", errs)
    assert max(errs["fp16"].values()) < 1e-2, errs
    assert max(errs["bf16"]["hidden_hi"], errs["bf16"]["hidden_lo"]) > 1e-2, errs       # inherent to 8-bit significands
    assert max(errs["bf16"].values()) < 6e-2, errs
