"""DaggerUpdater.update (robo-vln_b200/trainer.py) against the reference's _update_agent arithmetic
(robo_vln_baselines/hierarchical_trainer.py:492-560) restated with the stock torch losses and optimizers on the same
drop-in modules, and the CUDA-graph replay path against the eager one."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

T, L = 12, 20


def _policy(seed):
    import robovln_b200 as R

    torch.manual_seed(seed)
    pol = R.HcmPolicy().share_frozen_trunks().to("cuda")
    pol.high_level.train()
    pol.low_level.train()
    pol.high_level.dropout_p = 0.0          # exact comparisons: no dropout noise
    return pol


def _batch(seed, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    masks = torch.ones((T, 2))
    masks[0] = 0.0
    masks[7] = 0.0                           # a reset inside the trajectory: two LSTM segments
    sensor = torch.randint(0, 5, (T, 1), generator=g).float()          # 0 = ignore
    corrected = torch.rand((T, 2), generator=g)
    corrected[3] = 0.0                       # masked action row
    ostop = (torch.rand((T, 1), generator=g) > 0.7).float()
    ostop[5] = -1.0                          # ignored stop row
    return {
        "obs": {"rgb": torch.randint(0, 256, (T, 256, 256, 3), generator=g).float().to(dev),
                "depth": torch.rand((T, 256, 256, 1), generator=g).to(dev),
                "instruction": torch.randint(1000, 30522, (1, L), generator=g).float().to(dev),
                "vln_oracle_action_sensor": sensor.to(dev)},
        "prev": torch.zeros((T, 2), device=dev), "masks": masks.to(dev), "corrected": corrected.to(dev), "ostop": ostop.to(dev),
        "h_hi": (torch.randn((2, 1, 512), generator=g) * 0.1).to(dev), "h_lo": (torch.randn((2, 1, 512), generator=g) * 0.1).to(dev),
    }


def _reference_update(pol, opt_hi, opt_lo, b):
    """_update_agent with the stock criteria, on the drop-in modules' own forward (one device)."""
    obs = dict(b["obs"])
    opt_hi.zero_grad()
    opt_lo.zero_grad()
    out, h_hi = pol.high_level((obs, b["h_hi"].detach(), b["prev"], b["masks"]))
    mask = obs["vln_oracle_action_sensor"] == 0
    out = out.masked_fill(mask, 0)
    tgt = obs["vln_oracle_action_sensor"].squeeze(1).to(torch.int64)
    l_hi = nn.CrossEntropyLoss(ignore_index=-1, reduction="mean")(out, tgt - 1)
    l_hi.backward()
    opt_hi.step()
    disc = (tgt - 1).masked_fill(tgt == 0, 4)
    obs2 = {k: v for k, v in obs.items() if k != "vln_oracle_action_sensor"}
    act, stop, h_lo = pol.low_level((obs2, b["h_lo"].detach(), b["prev"], b["masks"], disc.view(-1)))
    act = act.masked_fill(b["corrected"] == 0, 0)
    l_a = nn.MSELoss()(act, b["corrected"])
    sel = b["ostop"] != -1
    l_s = nn.BCEWithLogitsLoss()(torch.masked_select(stop, sel), torch.masked_select(b["ostop"], sel))
    (l_a + l_s).backward()
    opt_lo.step()
    return (l_hi.item(), l_a.item(), l_s.item(), 0), h_hi.detach(), h_lo.detach()


def _params(pol):
    return {("hi." if m is pol.high_level else "lo.") + n: p.detach().clone() for m in (pol.high_level, pol.low_level)
            for n, p in m.named_parameters() if p.requires_grad}


def _close(a, b, tol, what):
    err = float((a - b).abs().max())
    assert err <= tol * max(1.0, float(b.abs().max())), f"{what}: max abs err {err}"


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graphed"])
def test_update_matches_reference_arithmetic(graph):
    import robovln_b200 as R

    ref, new = _policy(11), _policy(11)
    o_ref = (torch.optim.AdamW([p for p in ref.high_level.parameters() if p.requires_grad], lr=2.5e-4, weight_decay=1e-3),
             torch.optim.Adam([p for p in ref.low_level.parameters() if p.requires_grad], lr=2.5e-4, weight_decay=1e-3))
    o_new = (R.optim.FusedAdamW([p for p in new.high_level.parameters() if p.requires_grad], lr=2.5e-4, weight_decay=1e-3),
             R.optim.FusedAdam([p for p in new.low_level.parameters() if p.requires_grad], lr=2.5e-4, weight_decay=1e-3))
    upd = R.trainer.DaggerUpdater(new.high_level, new.low_level, o_new[0], o_new[1], graph=graph)
    for it in range(3):                      # the second and third calls replay the graphs captured by the first
        if it > 0:
            # every step is compared from IDENTICAL state: Adam's sign-like first steps amplify rounding differences of
            # near-zero gradients, which would otherwise drift the two copies apart over the steps
            new.high_level.load_state_dict(ref.high_level.state_dict())
            new.low_level.load_state_dict(ref.low_level.state_dict())
            o_new[0].load_state_dict(copy.deepcopy(o_ref[0].state_dict()))
            o_new[1].load_state_dict(copy.deepcopy(o_ref[1].state_dict()))
        b = _batch(100 + it)
        want, wh, wl = _reference_update(ref, o_ref[0], o_ref[1], copy.copy(b))
        obs = dict(b["obs"])
        got, gh, gl, dsl = upd.update(obs, b["prev"], b["masks"], b["corrected"], b["ostop"], b["h_hi"], b["h_lo"], "state")
        assert dsl == "state" and got[3] == 0
        assert "instruction" not in obs and obs["vln_oracle_action_sensor"].dtype == torch.int64 and obs["vln_oracle_action_sensor"].dim() == 1
        for k in range(3):
            assert abs(got[k] - want[k]) <= 2e-5 * max(1.0, abs(want[k])), (it, k, got, want)
        _close(gh, wh, 2e-5, "hi hidden")
        _close(gl, wl, 2e-5, "lo hidden")
        pw, pg = _params(ref), _params(new)
        for n in pw:
            # Adam's first steps move every weight by ~lr * sign(grad): where a gradient is ~0 the two implementations
            # may round it to different sides, so the bound on single elements is a fraction of lr, the mean is tight
            _close(pg[n], pw[n], 1.5e-4, f"step {it} {n}")
            assert float((pg[n] - pw[n]).abs().mean()) <= 5e-6, f"step {it} {n}: mean abs err"
    if graph:
        assert not upd.graph_errors, upd.graph_errors
        assert sum(g is not None for g in upd._graphs.values()) == 2      # one graph per model, reused across the three calls


def test_fused_optimizer_step_is_seen_by_the_inference_engine():
    """FusedAdam writes the parameters through raw pointers; the engine re-packs its 16-bit tail copies when the
    (data_ptr, _version) signature of the tail changes -- the optimizer must therefore bump the version counters.
    Inference right after an update (no eval()/train() toggle, no notify call) has to run on the NEW weights."""
    import robovln_b200 as R

    pol = _policy(21)
    opt_hi = R.optim.FusedAdamW([p for p in pol.high_level.parameters() if p.requires_grad], lr=5e-2, weight_decay=0.0)
    opt_lo = R.optim.FusedAdam([p for p in pol.low_level.parameters() if p.requires_grad], lr=5e-2, weight_decay=0.0)
    upd = R.trainer.DaggerUpdater(pol.high_level, pol.low_level, opt_hi, opt_lo, graph=True)
    b = _batch(7)

    def infer():
        with torch.no_grad():
            obs = {k: v for k, v in b["obs"].items() if k != "vln_oracle_action_sensor"}
            return pol.high_level((obs, b["h_hi"], b["prev"], b["masks"]))[0].clone()

    pol.high_level.eval()
    before = infer()
    pol.high_level.train()
    v0 = pol.high_level.linear.weight._version
    upd.update(dict(b["obs"]), b["prev"], b["masks"], b["corrected"], b["ostop"], b["h_hi"], b["h_lo"], None)
    assert pol.high_level.linear.weight._version > v0
    after_engine = infer()                       # train mode + no_grad -> engine inference path, packed tail weights
    assert float((after_engine - before).abs().max()) > 1e-2, "the engine still runs the tail weights of before the update"
    # and it agrees with the torch tail evaluated on the updated parameters (dropout off in _policy)
    rt = pol.high_level.runtime()
    feats = rt.encode(b["obs"]["rgb"], b["obs"]["depth"], b["obs"]["instruction"], n_envs=1)
    from robovln_b200 import torch_tail
    with torch.no_grad():
        want = torch_tail.hi_tail(pol.high_level, feats["rgb_feat"], feats["depth_feat"], feats["bert"], b["h_hi"], b["masks"], 0.0)[0]
    assert float((after_engine - want).abs().max()) < 2e-2
