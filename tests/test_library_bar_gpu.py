"""SURVEY.md 8(d) "library bar": the same forward as plain PyTorch ops on the SAME B200 (the oracle restatement
moved to the GPU: cuDNN convolutions, cuBLAS GEMMs, torch LayerNorm/GroupNorm/LSTM; fp32 with TF32 and fp16
autocast), timed next to the sm_100a engine.  The engine must not be slower than the library path; the
measured numbers are written to gpurun_out/library_bar.json for profiles/."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_engine_beats_torch_eager_on_the_same_gpu():
    import robovln_b200 as R
    from oracle import hcm_oracle as O
    from oracle import weights as W

    B, L = 64, 80
    dev = "cuda"
    sd_hi, sd_lo = W.make_state_dict("hi", 0), W.make_state_dict("lo", 0)
    inp = W.make_inputs(B=B, L=L, N=B, rgb_hw=256, seed=1, mask_zero_rows=(0,))
    g_hi = {k: v.to(dev) for k, v in sd_hi.items()}
    g_lo = {k: v.to(dev) for k, v in sd_lo.items()}
    gin = {k: v.to(dev) for k, v in inp.items() if isinstance(v, torch.Tensor)}

    def eager():
        with torch.no_grad():
            logits, _ = O.hi_forward(g_hi, gin["rgb"], gin["depth"], gin["instruction"], gin["hidden_hi"], gin["masks"])
            O.lo_forward(g_lo, gin["rgb"], gin["depth"], gin["hidden_lo"], gin["masks"], logits.argmax(1))

    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = False    # autotuning the 100+ conv shapes costs minutes of test time for ~no gain here
    ms_tf32 = _time(eager)

    def eager_fp16():
        with torch.autocast("cuda", dtype=torch.float16):
            eager()

    ms_fp16 = _time(eager_fp16)

    hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1)
    lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
    hi.load_state_dict(sd_hi)
    lo.load_state_dict(sd_lo)
    pol = R.HcmPolicy(hi, lo).to(dev).eval()
    obs = {k: gin[k] for k in ("rgb", "depth", "instruction")}
    ms_engine = _time(lambda: pol.act(obs, gin["hidden_hi"], gin["hidden_lo"], gin["masks"]), warm=3, reps=20)
    out = {"batch": B, "seq_len": L, "torch_eager_tf32_ms": ms_tf32, "torch_eager_fp16_autocast_ms": ms_fp16,
           "engine_ms": ms_engine, "note": "torch eager executes both trunks twice (hi and lo), as the reference does; "
           "the engine shares the frozen trunks.  Same GPU, same weights, same inputs, CUDA-event timing."}
    print(json.dumps(out))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(os.path.join("gpurun_out", "library_bar.json"), "w"), indent=1)
    except OSError:
        pass
    assert ms_engine < ms_fp16 and ms_engine < ms_tf32


def test_engine_beats_the_tuned_library_arm():
    """The stronger bar: channels_last fp16 torchvision ResNet-50 + HF BERT (SDPA) + cuDNN LSTM with
    cudnn.benchmark, trunks once, CUDA-graph replay (tools/library_bar.py) vs the engine's policy step."""
    import robovln_b200 as R
    from tools import library_bar

    B, L = 64, 80
    lib = library_bar.measure(B=B, L=L, steps=10, warmup=5)
    pol = R.HcmPolicy().share_frozen_trunks().to("cuda").eval()
    g = torch.Generator().manual_seed(2)
    obs = {"rgb": torch.randint(0, 256, (B, 256, 256, 3), generator=g).float().cuda(), "depth": torch.rand((B, 256, 256, 1), generator=g).cuda(),
           "instruction": torch.randint(1000, 30522, (B, L), generator=g).float().cuda()}
    h = torch.zeros((2, B, 512), device="cuda")
    masks = torch.ones((B, 2), device="cuda")
    ms_engine = _time(lambda: pol.act(obs, h, h.clone(), masks), warm=3, reps=20)
    out = {"library": lib, "engine_ms": ms_engine}
    print(json.dumps(out))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(os.path.join("gpurun_out", "library_bar_tuned.json"), "w"), indent=1)
    except OSError:
        pass
    assert ms_engine < lib["ms_per_step"], out
