"""DAgger update tail kernels (csrc/train.cu) against the torch expressions the reference trainer executes
(robo_vln_baselines/hierarchical_trainer.py:329-334 optimizers, :498-553 losses)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T", [1, 7, 64, 300])
@pytest.mark.parametrize("as_int", [False, True])
def test_hi_loss_matches_trainer_expression(T, as_int):
    import robovln_b200 as R

    g = torch.Generator().manual_seed(T)
    logits = (torch.randn((T, 4), generator=g) * 2).cuda().requires_grad_(True)
    sensor = torch.randint(0, 5, (T, 1), generator=g).float().cuda()
    if T == 1:
        sensor[:] = 2.0
    sensor[0] = 3.0                                       # at least one valid row
    # the trainer's sequence (:506-511)
    ref_in = logits.detach().clone().requires_grad_(True)
    out = ref_in * 1.0
    out = out.masked_fill_(sensor == 0, 0)
    tgt = sensor.squeeze(1).to(torch.int64) - 1
    ref = nn.CrossEntropyLoss(ignore_index=-1, reduction="mean")(out, tgt)
    ref.backward()
    loss = R.losses.hi_loss(logits, sensor.squeeze(1).to(torch.int64) if as_int else sensor)
    (loss * 1.0).backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 2e-6 * max(1.0, abs(float(ref.detach())))
    assert float((logits.grad - ref_in.grad).abs().max()) < 1e-6


@pytest.mark.parametrize("T", [1, 9, 64, 300])
def test_lo_loss_matches_trainer_expression(T):
    import robovln_b200 as R

    g = torch.Generator().manual_seed(100 + T)
    act = torch.randn((T, 2), generator=g).cuda().requires_grad_(True)
    stop = torch.randn((T, 1), generator=g).cuda().requires_grad_(True)
    corr = torch.randn((T, 2), generator=g).cuda()
    corr[torch.rand((T, 2), generator=g).cuda() < 0.3] = 0.0
    ostop = (torch.rand((T, 1), generator=g) > 0.7).float().cuda()
    ostop[torch.rand((T, 1), generator=g).cuda() < 0.25] = -1.0
    ostop[0] = 1.0
    a2, s2 = act.detach().clone().requires_grad_(True), stop.detach().clone().requires_grad_(True)
    out = (a2 * 1.0).masked_fill_(corr == 0, 0)                       # :543-547
    la = nn.MSELoss()(out, corr)
    mask = ostop != -1
    ls = nn.BCEWithLogitsLoss()(torch.masked_select(s2, mask), torch.masked_select(ostop, mask))
    (la + ls).backward()
    fa, fs = R.losses.lo_loss(act, stop, corr, ostop)
    (fa + fs).backward()
    assert abs(float(fa) - float(la)) < 2e-6 * max(1.0, float(la)) and abs(float(fs) - float(ls)) < 2e-6 * max(1.0, float(ls))
    assert float((act.grad - a2.grad).abs().max()) < 1e-6 and float((stop.grad - s2.grad).abs().max()) < 1e-6


@pytest.mark.parametrize("decoupled", [True, False])
def test_fused_adam_matches_torch(decoupled):
    import robovln_b200 as R

    g = torch.Generator().manual_seed(5)
    shapes = [(1,), (7,), (4096,), (4097,), (256, 768), (2048, 896), (3, 5, 7)]
    ref_p = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    my_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    frozen_ref = torch.randn(10, generator=g).cuda().requires_grad_(True)          # never gets a gradient: skipped
    frozen_my = frozen_ref.detach().clone().requires_grad_(True)
    kw = dict(lr=1e-4, weight_decay=1e-3)
    ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref_p + [frozen_ref], **kw)
    mine = (R.optim.FusedAdamW if decoupled else R.optim.FusedAdam)(my_p + [frozen_my], **kw)
    for it in range(6):
        if it == 3:                                   # an LR scheduler acts on param_groups
            for o in (ref, mine):
                o.param_groups[0]["lr"] = 3e-3
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 3))
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        mine.step()
        if it == 1:
            mine.zero_grad(set_to_none=True)          # gradients get new storage: pointer tables are rebuilt
    torch.cuda.synchronize()
    for a, b in zip(ref_p, my_p):
        # a few ulp after 6 steps: same formula, but fused multiply-adds are contracted differently than in torch's kernels
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(a.abs().max())), tuple(a.shape)
        sa, sb = ref.state[a], mine.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 6
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=2e-6, atol=1e-12)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=2e-6, atol=1e-20)
    assert torch.equal(frozen_ref, frozen_my) and len(mine.state[frozen_my]) == 0
    # the state dict is interchangeable with torch's
    ref2 = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref_p + [frozen_ref], **kw)
    ref2.load_state_dict(mine.state_dict())
    with pytest.raises(RuntimeError, match="float32 CUDA"):
        cpu = torch.zeros(3, requires_grad=True)
        cpu.grad = torch.zeros(3)
        R.optim.FusedAdam([cpu]).step()
