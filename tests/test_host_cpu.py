"""CPU-side checks (no GPU): the C-ABI library loads and exports every header symbol, the
parameter layout equals the reference's, weight preparation is numerically right, the module
API mirrors the reference's, and the N>1 sharding logic works over gloo with world_size 2."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as G

    G.build()
    from robovln_b200 import _lib

    lib = _lib.load(build_if_missing=False)
    header = open(os.path.join(ROOT, "include", "robovln_b200.h")).read()
    declared = set(re.findall(r"\b((?:hcm|rvb)_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature"
    assert b"sm_100a" in lib.hcm_version()


def test_no_gpu_means_loud_failure():
    import robovln_b200 as R

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1)
    obs = {"rgb": torch.zeros(1, 256, 256, 3), "depth": torch.zeros(1, 256, 256, 1), "instruction": torch.zeros(1, 4)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        hi((obs, torch.zeros(2, 1, 512), None, torch.ones(1, 2)))


@pytest.mark.parametrize("which", ["hi", "lo"])
def test_param_layout_matches_reference_manifest(which):
    from robovln_b200.param_spec import hi_spec, lo_spec

    spec = hi_spec() if which == "hi" else lo_spec()
    man = json.load(open(os.path.join(ROOT, "oracle", f"manifest_{which}.json")))
    assert list(spec.keys()) == list(man.keys())
    for k, (shape, dtype, _buf) in spec.items():
        assert list(shape) == man[k][0], k
        assert str(dtype) == man[k][1], k


def test_module_api_mirrors_reference():
    import robovln_b200 as R
    from oracle import weights as W

    hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1)
    lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    assert hi.state_encoder.num_recurrent_layers == 2 and lo.state_encoder.num_recurrent_layers == 2
    assert hi.num_recurrent_layers == 2 and hi.output_size == 512 and not hi.is_blind
    man_hi = json.load(open(os.path.join(ROOT, "oracle", "manifest_hi.json")))
    assert list(hi.state_dict().keys()) == list(man_hi.keys())
    res = hi.load_state_dict(W.make_state_dict("hi", 0), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    lo.load_state_dict(W.make_state_dict("lo", 0), strict=True)
    # frozen trunks carry no gradient, the trainable tail does (resnet_encoders.py:35-36,147-148)
    assert not hi.rgb_encoder.cnn.conv1.weight.requires_grad
    assert not hi.depth_encoder.visual_encoder.backbone.conv1._modules["0"].weight.requires_grad
    assert hi.image_cm_encoder.vis_fc.weight.requires_grad and lo.stop_linear.weight.requires_grad
    assert sum(p.numel() for p in hi.parameters()) == sum(
        int(torch.tensor(v[0]).prod()) if v[0] else 1 for k, v in man_hi.items()
        if "running_" not in k and "num_batches" not in k)
    hi.train(); hi.eval()
    with pytest.raises(NotImplementedError):
        from types import SimpleNamespace as NS
        R.Seq2Seq_HighLevel_CMA(None, 4, NS(STATE_ENCODER=NS(rnn_type="GRU", hidden_size=512)), 1)


def test_weight_prep_numerics():
    from oracle import hcm_oracle as O
    from oracle import weights as W
    from robovln_b200 import weight_prep as WP

    sd = W.make_state_dict("hi", 0)
    # BN folding == eval-mode BatchNorm after the conv (layer2.0 conv2: 3x3 stride 2)
    q = "rgb_encoder.cnn.layer2.0."
    x = torch.randn(2, 128, 16, 16)
    ref = F.batch_norm(F.conv2d(x, sd[q + "conv2.weight"], stride=2, padding=1), sd[q + "bn2.running_mean"],
                       sd[q + "bn2.running_var"], sd[q + "bn2.weight"], sd[q + "bn2.bias"], False, 0.0, 1e-5)
    w, b = WP.fold_bn(sd[q + "conv2.weight"], sd[q + "bn2.weight"], sd[q + "bn2.bias"], sd[q + "bn2.running_mean"],
                      sd[q + "bn2.running_var"])
    got = F.conv2d(x, w, b, stride=2, padding=1)
    assert float((got - ref).abs().max()) < 1e-4
    # K-major layout: k = (r*KW + s)*Cin + c
    wk = WP._conv_kmajor(w)
    assert wk.shape == (128, 9 * 128)
    assert torch.equal(wk[5, (1 * 3 + 2) * 128 + 7], w[5, 7, 1, 2])
    # packed tensors: names, dtypes, shapes
    t = {}
    t.update(WP.prep_rgb_trunk(sd, "hi", "cpu"))
    t.update(WP.prep_depth_trunk(sd, "hi", "cpu"))
    t.update(WP.prep_bert(sd, "cpu"))
    t.update(WP.prep_hi_tail(sd, "cpu"))
    assert t["hi.rgb.stem.w"].shape == (64, 256) and t["hi.rgb.stem.w"].dtype == torch.float16   # default build
    sw = t["hi.rgb.stem.w"].float().view(64, 4, 8, 2, 4).permute(0, 1, 3, 2, 4).reshape(64, 8, 8, 4)   # [o, r (8th = 0), s, c]
    assert torch.all(sw[:, :, 7] == 0) and torch.all(sw[:, :, :, 3:] == 0) and torch.all(sw[:, 7] == 0)
    wf, _ = WP.fold_bn(sd["rgb_encoder.cnn.conv1.weight"], sd["rgb_encoder.cnn.bn1.weight"], sd["rgb_encoder.cnn.bn1.bias"],
                       sd["rgb_encoder.cnn.bn1.running_mean"], sd["rgb_encoder.cnn.bn1.running_var"])
    assert float((sw[9, 2, 5, 1] - wf[9, 1, 2, 5]).abs()) < 2e-3   # [o, c, r, s] -> k = (r//2)*64 + s*8 + (r%2)*4 + c
    assert t["hi.rgb.l4.0.c2.w"].shape == (512, 9 * 512) and t["hi.depth.comp.w"].shape == (128, 9 * 1024)
    assert t["hi.bert.3.qkv.w"].shape == (2304, 768) and t["hi.vla.fc_kv.w"].shape == (512, 256)
    assert t["hi.lstm.b"].dtype == torch.float32
    # depth_linear column permutation: cell-major tokens x permuted weight == reference flatten order
    D = torch.randn(3, 192, 16)                                  # [B, C, cell] as the reference holds it
    ref = F.linear(torch.flatten(D, 1), sd["depth_linear.1.weight"])
    tokens = D.permute(0, 2, 1).reshape(3, 16 * 192)             # [B, cell, C] as the engine holds it
    got = F.linear(tokens, t["hi.depth_linear.w"].float())
    assert float((got - ref).abs().max()) < 0.05                 # bf16 weights
    # lo visual_fc: zero columns under the spatial-embedding slots
    sdl = W.make_state_dict("lo", 0)
    tl = WP.prep_lo_tail(sdl, "cpu")
    wl = tl["lo.depth_fc.w"].float().view(128, 16, 192)
    assert torch.all(wl[:, :, 128:] == 0)
    feat = torch.randn(3, 128, 16)
    ref = F.linear(torch.flatten(feat, 1), sdl["depth_encoder.visual_fc.1.weight"])
    tok = torch.zeros(3, 16, 192); tok[:, :, :128] = feat.permute(0, 2, 1); tok[:, :, 128:] = 123.0
    got = F.linear(tok.reshape(3, -1), tl["lo.depth_fc.w"].float())
    assert float((got - ref).abs().max()) < 0.05
    # trunk identity detection (dedup legality)
    assert WP.trunks_identical(sd, sdl)
    sdl2 = dict(sdl); sdl2["rgb_encoder.cnn.layer1.0.conv1.weight"] = sdl["rgb_encoder.cnn.layer1.0.conv1.weight"] + 1e-3
    assert not WP.trunks_identical(sd, sdl2)
    # the spatial-embedding reinterpretation the kernels implement (flat[c*16 + cell])
    e = sd["rgb_encoder.spatial_embeddings.weight"]
    assert torch.equal(O.spatial_embedding_channels(e)[0, 5, 2, 3], e.flatten()[5 * 16 + 2 * 4 + 3])


def test_shard_range_partition():
    from robovln_b200.sharding import shard_range

    for rows in (1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import robovln_b200
from robovln_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for rows in (8, 7):
    full = torch.arange(rows * 7, dtype=torch.float32).view(rows, 7)
    lo, hi = sharding.shard_range(rows, rank, world)
    local = sharding.pack_outputs(full[lo:hi, :4], full[lo:hi, 4:6], full[lo:hi, 6:7])
    out = sharding.all_gather_outputs(local, rows)
    assert torch.equal(out, full), (rank, rows, out)
    l, a, s = sharding.unpack_outputs(out)
    assert l.shape == (rows, 4) and a.shape == (rows, 2) and s.shape == (rows, 1)
dist.destroy_process_group()
print("ok", rank)
"""


def test_output_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_constructor_loads_ddppo_checkpoint_and_reports_random_encoders(tmp_path):
    """Constructor parity (resnet_encoders.py:38-52, :151; seq2seq_highlevel_cma.py:45): DEPTH_ENCODER.ddppo_checkpoint is
    loaded with the reference's key remap; pretrained BERT / ImageNet weights that are not available offline are
    reported loudly instead of silently training on random frozen encoders."""
    import warnings
    from types import SimpleNamespace as NS

    import robovln_b200 as R

    donor = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    src = donor.depth_encoder.visual_encoder.state_dict()
    g = torch.Generator().manual_seed(11)
    ck = {"state_dict": {"actor_critic.net.visual_encoder." + k: torch.randn(v.shape, generator=g) for k, v in src.items()}}
    ck["state_dict"]["actor_critic.net.prev_action_embedding.weight"] = torch.zeros(5, 32)     # ignored, as in the reference
    path = str(tmp_path / "gibson-2plus-resnet50.pth")
    torch.save(ck, path)
    cfg = NS(DEPTH_ENCODER=NS(cnn_type="VlnResnetDepthEncoder", backbone="resnet50", output_size=128, ddppo_checkpoint=path))
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        hi = R.Seq2Seq_HighLevel_CMA(None, 4, cfg, 1)
    for k, v in hi.depth_encoder.visual_encoder.state_dict().items():
        assert torch.equal(v, ck["state_dict"]["actor_critic.net.visual_encoder." + k]), k
    msgs = [str(w.message) for w in rec if issubclass(w.category, RuntimeWarning)]
    assert any("RANDOM" in m and "rgb_encoder.cnn" in m and "embedding_layer" in m for m in msgs), msgs
    assert not any("depth_encoder" in m for m in msgs)
    assert hi.random_frozen_encoders and all("depth" not in s for s in hi.random_frozen_encoders)
    # no checkpoint configured -> the depth trunk is reported too; a missing file raises like torch.load does
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        R.Seq2Seq_LowLevel(None, 2, 4, NS(DEPTH_ENCODER=NS(ddppo_checkpoint="NONE")), 1)
    assert any("depth_encoder" in str(w.message) for w in rec)
    with pytest.raises(FileNotFoundError):
        R.Seq2Seq_LowLevel(None, 2, 4, NS(DEPTH_ENCODER=NS(ddppo_checkpoint=str(tmp_path / "missing.pth"))), 1)
    # dims that are hard-wired into the kernels are validated instead of silently ignored
    with pytest.raises(NotImplementedError, match="ins_in_features"):
        R.Seq2Seq_HighLevel_CMA(None, 4, NS(VISUAL_LING_ATTN=NS(ins_in_features=512)), 1)
    with pytest.raises(NotImplementedError, match="d_in"):
        R.Seq2Seq_HighLevel_CMA(None, 4, NS(TRANSFORMER_INSTRUCTION_ENCODER=NS(d_in=512)), 1)


def test_same_device_to_does_not_invalidate_the_runtime():
    """hierarchical_trainer.py:517 calls low_level.to(device2) on every update: a move that changes nothing must not
    mark the engine's packed weights dirty (it used to force a re-pack + re-plan per training step)."""
    import robovln_b200 as R

    class FakeRt:
        device = torch.device("cpu")
        dirty = 0

        def mark_dirty(self):
            self.dirty += 1

    lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    rt = FakeRt()
    lo.__dict__["_rt"] = rt
    lo.to("cpu")
    lo.to(torch.device("cpu"))
    lo.float()
    assert rt.dirty == 0
    lo.double()            # storage really changed
    assert rt.dirty == 1
