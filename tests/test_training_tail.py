"""Training path: the trainable tail (robo-vln_b200/torch_tail.py) against the oracle -- values and
gradients -- on CPU, and on the GPU the full train()-mode forward/backward through the modules
(frozen encoders on the engine, tail under autograd) with the losses of
hierarchical_trainer.py:498-553 (BASELINE.json configs[4], dropout disabled for parity)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import hcm_oracle as O
from oracle import weights as W


def _losses_hi(logits, targets):
    return F.cross_entropy(logits, targets, ignore_index=-1)


def _losses_lo(act, stop, gt_act, gt_stop):
    return F.mse_loss(act, gt_act) + F.binary_cross_entropy_with_logits(stop, gt_stop)


@pytest.fixture(scope="module")
def cpu_case():
    sd_hi, sd_lo = W.make_state_dict("hi", 0), W.make_state_dict("lo", 0)
    inp = W.make_inputs(B=3, L=10, N=1, rgb_hw=256, seed=21, mask_zero_rows=(0, 2))
    with torch.no_grad():
        logits, hid, it = O.hi_forward(sd_hi, inp["rgb"], inp["depth"], inp["instruction"], inp["hidden_hi"], inp["masks"],
                                       return_intermediates=True)
        rgb_trunk = O.rgb_trunk(sd_lo, "rgb_encoder.", inp["rgb"])
    feats = {
        "rgb_feat": it["rgb_embedding"][:, :2048].permute(0, 2, 1).contiguous(),
        "depth_feat": it["depth_embedding"][:, :128].permute(0, 2, 1).contiguous(),
        "bert": it["bert"], "rgb_gmean": rgb_trunk.mean(dim=(2, 3)),
    }
    return sd_hi, sd_lo, inp, feats, logits, hid


def test_torch_tail_matches_oracle_values_and_grads(cpu_case):
    import robovln_b200 as R
    from robovln_b200 import torch_tail

    sd_hi, sd_lo, inp, feats, ref_logits, ref_hid = cpu_case
    hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1)
    hi.load_state_dict(sd_hi)
    hi.eval()      # dropout off; torch_tail itself is device agnostic
    logits, hid = torch_tail.hi_tail(hi, feats["rgb_feat"], feats["depth_feat"], feats["bert"], inp["hidden_hi"],
                                     inp["masks"], 0.25)
    assert float((logits - ref_logits).abs().max()) < 1e-4
    assert float((hid - ref_hid).abs().max()) < 1e-4
    targets = torch.tensor([1, -1, 3])
    _losses_hi(logits, targets).backward()
    # oracle gradients: the same loss through the restated forward with the tail weights requiring grad
    keys = ["linear.weight", "state_encoder.rnn.weight_hh_l0", "image_cm_encoder.vis_fc.weight",
            "image_cm_encoder.layers.0.enc_att.attention.fc_q.weight", "rgb_kv.weight", "depth_linear.1.weight",
            "rgb_encoder.spatial_embeddings.weight"]
    sd = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd_hi.items()}
    o_logits, _ = O.hi_forward(sd, inp["rgb"], inp["depth"], inp["instruction"], inp["hidden_hi"], inp["masks"])
    _losses_hi(o_logits, targets).backward()
    params = dict(hi.named_parameters())
    for k in keys:
        g, r = params[k].grad, sd[k].grad
        assert g is not None and r is not None, k
        assert float((g - r).abs().max()) <= 1e-4 + 1e-3 * float(r.abs().max()), k
    assert params["rgb_encoder.cnn.conv1.weight"].grad is None        # frozen trunk

    lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    lo.load_state_dict(sd_lo)
    lo.eval()
    act, stop, hl = torch_tail.lo_tail(lo, feats["rgb_gmean"], feats["depth_feat"], inp["hidden_lo"], inp["masks"],
                                       inp["sub_goal"])
    with torch.no_grad():
        r_act, r_stop, r_hl = O.lo_forward(sd_lo, inp["rgb"], inp["depth"], inp["hidden_lo"], inp["masks"], inp["sub_goal"])
    assert float((act - r_act).abs().max()) < 1e-4 and float((stop - r_stop).abs().max()) < 1e-4
    assert float((hl - r_hl).abs().max()) < 1e-4


@pytest.mark.gpu
def test_train_mode_forward_backward_on_gpu(cpu_case):
    import robovln_b200 as R

    sd_hi, sd_lo, inp, feats, ref_logits, ref_hid = cpu_case
    hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1)
    lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    hi.load_state_dict(sd_hi)
    lo.load_state_dict(sd_lo)
    hi.cuda().train()
    lo.cuda().train()
    hi.dropout_p = 0.0                      # parity run: dropout disabled
    dev = "cuda"
    obs = {"rgb": inp["rgb"].to(dev), "depth": inp["depth"].to(dev), "instruction": inp["instruction"].to(dev)}
    logits, hid = hi((obs, inp["hidden_hi"].to(dev), None, inp["masks"].to(dev)))
    assert logits.requires_grad and "instruction" not in obs
    assert float((logits.detach().cpu() - ref_logits).abs().max()) < 1e-2
    assert float((hid.detach().cpu() - ref_hid).abs().max()) < 1e-2
    targets = torch.tensor([1, -1, 3], device=dev)
    _losses_hi(logits, targets).backward()
    sd = {k: (v.clone().requires_grad_(True) if k in ("linear.weight", "image_cm_encoder.vis_fc.weight", "rgb_kv.weight") else v)
          for k, v in sd_hi.items()}
    o_logits, _ = O.hi_forward(sd, inp["rgb"], inp["depth"], inp["instruction"], inp["hidden_hi"], inp["masks"])
    _losses_hi(o_logits, targets.cpu()).backward()
    params = dict(hi.named_parameters())
    for k in ("linear.weight", "image_cm_encoder.vis_fc.weight", "rgb_kv.weight"):
        g, r = params[k].grad.cpu(), sd[k].grad
        # the gradient logic is pinned exactly on CPU above; here fp16 encoder features (errors ~5e-3)
        # perturb the tail's activations, hence its gradients: 15 % of the largest entry
        assert float((g - r).abs().max()) <= 0.15 * float(r.abs().max()) + 1e-5, k
    assert params["embedding_layer.encoder.layer.0.output.dense.weight"].grad is None
    assert params["rgb_encoder.cnn.layer1.0.conv1.weight"].grad is None

    act, stop, hl = lo((obs, inp["hidden_lo"].to(dev), None, inp["masks"].to(dev), inp["sub_goal"].to(dev)))
    loss = _losses_lo(act, stop, torch.zeros_like(act), torch.ones_like(stop))
    loss.backward()
    with torch.no_grad():
        r_act, r_stop, _ = O.lo_forward(sd_lo, inp["rgb"], inp["depth"], inp["hidden_lo"], inp["masks"], inp["sub_goal"])
    assert float((act.detach().cpu() - r_act).abs().max()) < 1e-2
    assert dict(lo.named_parameters())["stop_linear.weight"].grad is not None
    # an optimizer step changes the weights the ENGINE sees on the next eval-mode call
    opt = torch.optim.SGD([p for p in hi.parameters() if p.requires_grad and p.grad is not None], lr=0.5)
    opt.step()
    hi.notify_weights_updated()
    hi.eval()
    obs["instruction"] = inp["instruction"].to(dev)
    with torch.no_grad():
        logits2, _ = hi((obs, inp["hidden_hi"].to(dev), None, inp["masks"].to(dev)))
    assert float((logits2 - logits.detach()).abs().max()) > 1e-3


def test_segment_starts_and_precomputed_segments():
    """trainer.DaggerUpdater computes the LSTM segment boundaries once on the host (they key its CUDA graphs) and hands
    them to the tail: same boundaries as RNNStateEncoder.seq_forward (rnn_state_encoder.py:95-110), same results."""
    from robovln_b200 import torch_tail, trainer

    m = torch.ones(10)
    m[0] = 0.0
    m[4] = 0.0
    m[9] = 0.0
    assert torch_tail.segment_starts(m, 1) == [0, 4, 9]
    assert torch_tail.segment_starts(torch.ones(6), 1) == [0]
    m2 = torch.ones(8)          # T = 4, N = 2: a reset of either environment cuts the sequence
    m2[5] = 0.0                 # t = 2, env 1
    assert torch_tail.segment_starts(m2, 2) == [0, 2]
    g = torch.Generator().manual_seed(3)
    rnn = torch.nn.LSTM(6, 5)
    x = torch.randn((10, 6), generator=g)
    h = torch.randn((2, 1, 5), generator=g)
    y0, h0 = torch_tail.lstm_state_encoder(rnn, x, h, m)
    y1, h1 = torch_tail.lstm_state_encoder(rnn, x, h, m, starts=[0, 4, 9])
    assert torch.equal(y0, y1) and torch.equal(h0, h1)
    hh = (torch.ones(2, requires_grad=True) * 2, (torch.ones(1, requires_grad=True) * 3,))
    out = trainer.repackage_hidden(hh)
    assert not out[0].requires_grad and not out[1][0].requires_grad


def test_dagger_updater_mirrors_update_agent_signature(golden_dir):
    """trainer.DaggerUpdater.update takes the arguments of the reference's _update_agent in the same order and returns
    the same 4-tuple (fixture: oracle/make_golden_api.py, read with ast from the unmodified reference source)."""
    import inspect
    import json
    import os

    from robovln_b200 import trainer

    api = json.load(open(os.path.join(golden_dir, "update_agent_api.json")))
    params = [p for p in inspect.signature(trainer.DaggerUpdater.update).parameters if p != "self"]
    assert params == api["args"]
    assert api["returns"] == ["loss", "high_recurrent_hidden_states", "low_recurrent_hidden_states", "detached_state_low"]
    assert api["loss_tuple_len"] == 4
    src = inspect.getsource(trainer.DaggerUpdater.update)
    assert "return loss, hi_hidden, lo_hidden, detached_state_low" in src
