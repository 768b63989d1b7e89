"""tcgen05 implicit-GEMM kernel (csrc/gemm_tc.cu) against an fp32 torch reference evaluated on
the same 16-bit-rounded operands, and against the CUDA-core validation kernel, for both builds
of the library (fp16 default, bf16).

Tolerance: 16-bit outputs are held to the output rounding (2e-3 fp16 / 1e-2 bf16 of the
tensor's max magnitude); fp32 outputs to 2e-3 (accumulation-order differences over K <= 9216)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

F32_TOL = 2e-3
DTYPES = ["fp16", "bf16"]


def _mk(shape, scale=1.0, seed=0, dtype="fp16"):
    from tests.gpu_util import H16

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randn(shape, generator=g, device="cuda") * scale
    return (t if dtype == "f32" else t.to(H16[dtype])).contiguous()


# (name, NB, H, W, Cin, Cout, K, stride, pad)
CONV_CASES = [
    ("plain_tail_m300", 1, 1, 300, 64, 64, 1, 1, 0),
    ("plain_k32_n32_depth_l1", 2, 32, 32, 32, 32, 1, 1, 0),
    ("plain_layer1_c64_256", 2, 64, 64, 64, 256, 1, 1, 0),
    ("plain_many_tiles", 8, 64, 64, 64, 256, 1, 1, 0),           # 256 x 1..2 tiles > 148 CTAs: persistent loop
    ("plain_k416_tail", 1, 1, 64, 416, 2048, 1, 1, 0),           # LSTM-lo input projection (K tail via OOB fill)
    ("plain_k2112", 1, 1, 1024, 2112, 256, 1, 1, 0),             # rgb_kv
    ("plain_bert_qkv", 1, 1, 1600, 768, 2304, 1, 1, 0),
    ("plain_bert_ff2", 1, 1, 1600, 3072, 768, 1, 1, 0),
    ("c3x3_64x64_c64", 2, 64, 64, 64, 64, 3, 1, 1),              # RGB layer1 conv2: th=2
    ("c3x3_32x32_c128", 3, 32, 32, 128, 128, 3, 1, 1),           # th=4
    ("c3x3_16x16_c256", 3, 16, 16, 256, 256, 3, 1, 1),           # th=8
    ("c3x3_8x8_c512", 5, 8, 8, 512, 512, 3, 1, 1),               # two images per tile, 72 k-blocks, odd NB
    ("c3x3_4x4_c1024_comp", 3, 4, 4, 1024, 128, 3, 1, 1),        # eight images per tile, NB < nb
    ("c3x3_32x32_c32_depth", 2, 32, 32, 32, 32, 3, 1, 1),        # Cin=32: half of every A box is OOB
    ("c3x3_s2_64_to_32", 2, 64, 64, 128, 128, 3, 2, 1),          # elementStrides = 2
    ("c3x3_s2_16_to_8", 3, 16, 16, 512, 512, 3, 2, 1),
    ("c3x3_s2_8_to_4", 3, 8, 8, 256, 256, 3, 2, 1),
    ("c1x1_s2_ds_64_to_32", 2, 64, 64, 256, 512, 1, 2, 0),       # downsample branch
    ("c1x1_s2_ds_8_to_4", 3, 8, 8, 512, 1024, 1, 2, 0),
    ("c3x3_56x56_c64_rgb224", 2, 56, 56, 64, 64, 3, 1, 1),       # 224x224 geometry: 112-row tiles
    ("c3x3_s2_56_to_28", 2, 56, 56, 128, 128, 3, 2, 1),
    ("c3x3_14x14_c256", 3, 14, 14, 256, 256, 3, 1, 1),           # th=7 -> 98-row tiles
    ("c3x3_7x7_c512", 3, 7, 7, 512, 512, 3, 1, 1),               # two 49-pixel images per tile
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_matches_torch(case, dtype):
    from tests.gpu_util import OUT_TOL, conv_gemm, conv_ref, rel_err

    name, NB, H, W, Cin, Cout, K, stride, pad = case
    x = _mk((NB, H, W, Cin), 1.0, 1, dtype)
    w = _mk((Cout, K * K * Cin), (2.0 / (K * K * Cin)) ** 0.5, 2, dtype)
    bias = _mk((Cout,), 0.5, 3, "f32")
    ref = conv_ref(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, act=1)
    out = conv_gemm(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, act=1)
    assert out.shape == ref.shape
    e = rel_err(out, ref)
    assert e < OUT_TOL[dtype], f"{name}: tcgen05 vs torch rel err {e:.3e}"
    simt = conv_gemm(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, act=1, impl=1)
    e2 = rel_err(simt, ref)
    assert e2 < OUT_TOL[dtype], f"{name}: simt vs torch rel err {e2:.3e}"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_forced_tile_widths(bn, dtype):
    from tests.gpu_util import conv_gemm, conv_ref, rel_err

    x = _mk((1, 1, 1000, 512), 1.0, 4, dtype)
    w = _mk((768, 512), 512 ** -0.5, 5, dtype)
    ref = conv_ref(x, w)
    out = conv_gemm(x, w, force_bn=bn, out_f32=True)
    e = rel_err(out, ref)
    assert e < F32_TOL, f"BN={bn}: rel err {e:.3e}"


PAIR = 256 | (1 << 16)   # force_bn encoding of the CTA-pair (cta_group::2) kernel, 256 x 256 tiles


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [
    ("pair_plain_even", 1, 1, 1024, 768, 768, 1, 1, 0),
    ("pair_plain_odd_tiles_tail", 1, 1, 1600 + 37, 512, 512, 1, 1, 0),      # 13 M tiles (odd) + ragged tail
    ("pair_plain_n_tail", 1, 1, 512, 1024, 384, 1, 1, 0),                   # N not a multiple of 256
    ("pair_c3x3_16x16", 4, 16, 16, 256, 256, 3, 1, 1),
    ("pair_c3x3_s2_8x8_two_imgs", 5, 16, 16, 512, 512, 3, 2, 1),             # 2 images per 128-row tile, odd tile count
    ("pair_many_units", 1, 1, 128 * 60, 512, 1024, 1, 1, 0),                # 30 x 4 = 120 pair tiles > 74 clusters
], ids=lambda c: c[0])
def test_cta_pair_kernel(case, dtype):
    """cta_group::2: two CTAs of a cluster run one M=256 UMMA, each staging half of B."""
    from tests.gpu_util import OUT_TOL, conv_gemm, conv_ref, rel_err

    name, NB, H, W, Cin, Cout, K, stride, pad = case
    x = _mk((NB, H, W, Cin), 1.0, 31, dtype)
    w = _mk((Cout, K * K * Cin), (2.0 / (K * K * Cin)) ** 0.5, 32, dtype)
    bias = _mk((Cout,), 0.5, 33, "f32")
    M = NB * ((H + 2 * pad - K) // stride + 1) * ((W + 2 * pad - K) // stride + 1)
    res = _mk((M, Cout), 1.0, 34, dtype)
    ref = conv_ref(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, res=res, act=1)
    out = conv_gemm(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, res=res, act=1, force_bn=PAIR)
    assert rel_err(out, ref) < OUT_TOL[dtype], name
    one = conv_gemm(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, res=res, act=1, force_bn=128, out_f32=True)
    two = conv_gemm(x, w, KH=K, KW=K, stride=stride, pad=pad, bias=bias, res=res, act=1, force_bn=PAIR, out_f32=True)
    assert rel_err(two, one) < 1e-5, name          # same products, same fp32 accumulation order per K block


@pytest.mark.parametrize("dtype", DTYPES)
def test_epilogue_variants(dtype):
    from tests.gpu_util import OUT_TOL, conv_gemm, conv_ref, rel_err

    M, K, N = 640, 256, 256
    x = _mk((1, 1, M, K), 1.0, 6, dtype)
    w = _mk((N, K), K ** -0.5, 7, dtype)
    bias = _mk((N,), 0.3, 8, "f32")
    res = _mk((M, N), 1.0, 9, dtype)
    # bias + residual, fp32 out (pre-LayerNorm tensors)
    ref = conv_ref(x, w, bias=bias, res=res)
    out = conv_gemm(x, w, bias=bias, res=res, out_f32=True)
    assert rel_err(out, ref) < F32_TOL
    # residual broadcast over row blocks (cross-modal fc_o: res row = m % res_rows)
    res2 = _mk((160, N), 1.0, 10, dtype)
    ref = conv_ref(x, w, bias=bias, res=res2, res_rows=160)
    out = conv_gemm(x, w, bias=bias, res=res2, res_rows=160, out_f32=True)
    assert rel_err(out, ref) < F32_TOL
    # GELU(erf), 16-bit out
    ref = conv_ref(x, w, bias=bias, act=2)
    out = conv_gemm(x, w, bias=bias, act=2)
    assert rel_err(out, ref) < OUT_TOL[dtype]
    # no bias, no activation (depth convs feeding GroupNorm)
    ref = conv_ref(x, w)
    out = conv_gemm(x, w)
    assert rel_err(out, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw", [(256, 256), (224, 224), (128, 160)])
def test_rgb_stem_window_conv(hw, dtype):
    """7x7 stride-2 pad-3 conv (3 -> 64) + bias + ReLU on the tensor cores through overlapping-window
    TMA loads of the zero-padded NHW8 image, against torch's conv2d on the /255 image."""
    from robovln_b200.weight_prep import stem_window_weights
    from tests.gpu_util import H16, OUT_TOL, rel_err, rgb_stem

    H, W = hw
    NB = 3
    g = torch.Generator(device="cuda")
    g.manual_seed(21)
    rgb = torch.randint(0, 256, (NB, H, W, 3), generator=g, device="cuda").float()
    w = torch.randn((64, 3, 7, 7), generator=g, device="cuda") * (2.0 / 147) ** 0.5
    bias = torch.randn((64,), generator=g, device="cuda") * 0.2
    w_win = stem_window_weights(w).to(H16[dtype]).contiguous()
    out, padded = rgb_stem(rgb, w_win, bias, dtype)
    # pre-pass: interior = rgb/255 in channels 0..2, everything else zero
    img = (rgb / 255.0).to(H16[dtype]).float()
    assert torch.equal(padded[:, 3:H + 3, 3:W + 3, :3].float(), img)
    assert float(padded[..., 3:].float().abs().max()) == 0.0
    assert float(padded[:, :3].float().abs().max()) == 0.0 and float(padded[:, :, W + 3:].float().abs().max()) == 0.0
    w_q = w_win.float().view(64, 7, 8, 8)[:, :, :7, :3].permute(0, 3, 1, 2).contiguous()      # rounded weights, OIHW
    ref = torch.relu(torch.nn.functional.conv2d(img.permute(0, 3, 1, 2), w_q, bias, stride=2, padding=3))
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 64)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < OUT_TOL[dtype]
    simt, _ = rgb_stem(rgb, w_win, bias, dtype, impl=1)
    assert rel_err(simt, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("hw", [(256, 256), (224, 224), (128, 160)])
def test_rgb_stem_packed_window_conv(hw, dtype):
    """The stem the engine runs: row-pair-interleaved padded image, overlapping-window TMA boxes that put TWO
    filter rows in every 64-wide K block (4 K blocks instead of 7), against torch's conv2d on the /255 image."""
    from robovln_b200.weight_prep import stem_packed_weights
    from tests.gpu_util import H16, OUT_TOL, rel_err, rgb_stem_packed

    H, W = hw
    NB = 3
    g = torch.Generator(device="cuda")
    g.manual_seed(22)
    rgb = torch.randint(0, 256, (NB, H, W, 3), generator=g, device="cuda").float()
    w = torch.randn((64, 3, 7, 7), generator=g, device="cuda") * (2.0 / 147) ** 0.5
    bias = torch.randn((64,), generator=g, device="cuda") * 0.2
    w_pk = stem_packed_weights(w).to(H16[dtype]).contiguous()
    out, padded = rgb_stem_packed(rgb, w_pk, bias, dtype)
    img = (rgb / 255.0).to(H16[dtype]).float()
    assert torch.equal(padded[:, 3:H + 3, 3:W + 3, :3].float(), img)
    assert float(padded[..., 3:].float().abs().max()) == 0.0
    assert float(padded[:, :3].float().abs().max()) == 0.0 and float(padded[:, :, W + 3:].float().abs().max()) == 0.0
    w_q = w_pk.float().view(64, 4, 8, 2, 4).permute(0, 1, 3, 2, 4).reshape(64, 8, 8, 4)[:, :7, :7, :3]
    w_q = w_q.permute(0, 3, 1, 2).contiguous()                                                   # rounded weights, OIHW
    ref = torch.relu(torch.nn.functional.conv2d(img.permute(0, 3, 1, 2), w_q, bias, stride=2, padding=3))
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 64)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < OUT_TOL[dtype]
    simt, _ = rgb_stem_packed(rgb, w_pk, bias, dtype, impl=1)
    assert rel_err(simt, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,N,res_rows,act,pe_rows", [(5120, 768, 768, 0, 0, 0), (1000, 3072, 768, 0, 0, 0),
                                                        (640, 768, 256, -1, 1, 80), (2048, 256, 256, 1024, 0, 0),
                                                        (300, 1024, 256, 0, 0, 0), (128 * 60, 256, 512, 0, 1, 0)])
def test_gemm_layernorm_epilogue(M, K, N, res_rows, act, pe_rows, dtype):
    """LayerNorm folded into the GEMM store (row statistics exchanged between the N/256 CTAs of a cluster over
    DSMEM) against torch: LN(act(A W^T + bias + res)) * gamma + beta (+ pe).  res_rows -1 = no residual."""
    import ctypes

    from tests.gpu_util import H16, OUT_TOL, P, check, lib, rel_err, stream

    a = _mk((M, K), 1.0, 31, dtype)
    w = _mk((N, K), K ** -0.5, 32, dtype)
    bias = _mk((N,), 0.3, 33, "f32")
    gamma = _mk((N,), 0.2, 34, "f32") + 1.0
    beta = _mk((N,), 0.2, 35, "f32")
    res = None if res_rows < 0 else _mk((res_rows if res_rows > 0 else M, N), 1.0, 36, dtype)
    pe = _mk((pe_rows, N), 0.5, 37, "f32") if pe_rows > 0 else None
    out = torch.zeros((M, N), dtype=H16[dtype], device="cuda")
    for eps in (1e-12, 1e-5):
        check(lib(dtype).rvb_gemm_ln(P(a), M, K, P(w), N, P(bias), P(res), max(res_rows, 0), act, P(gamma), P(beta),
                                     ctypes.c_float(eps), P(pe), pe_rows, P(out), stream()), "rvb_gemm_ln", dtype)
        torch.cuda.synchronize()
        y = a.float() @ w.float().t() + bias
        if res is not None:
            y = y + res.float()[torch.arange(M, device="cuda") % res.shape[0]]
        if act == 1:
            y = torch.relu(y)
        ref = torch.nn.functional.layer_norm(y, (N,), gamma, beta, eps)
        if pe is not None:
            ref = ref + pe[torch.arange(M, device="cuda") % pe_rows]
        assert rel_err(out, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("NB,HW,Cin,Cout,K,stride,res,relu", [
    (6, 8, 256, 128, 1, 1, False, True),      # layer3 conv1: 64-pixel samples, 8 channels per group, BN = 128
    (6, 8, 128, 128, 3, 1, False, True),      # layer3 conv2 (3x3)
    (5, 8, 128, 512, 1, 1, True, True),       # layer3 conv3 + residual: 32 channels per group
    (4, 16, 128, 128, 3, 2, False, True),     # layer3 block 0 conv2, stride 2: 16x16 -> 8x8
    (19, 4, 512, 256, 1, 1, False, True),     # layer4 conv1: 16-pixel samples, 16 channels per group, ragged last tile
    (16, 4, 256, 256, 3, 1, False, False),    # layer4 conv2 (3x3), no ReLU
    (9, 4, 256, 1024, 1, 1, True, True),      # layer4 conv3 + residual: 64 channels per group
])
def test_conv_groupnorm_epilogue(NB, HW, Cin, Cout, K, stride, res, relu, dtype):
    """GroupNorm(16) folded into the conv GEMM's store (depth trunk layers 3-4) against torch."""
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, rel_err, stream

    x = _mk((NB, HW, HW, Cin), 1.0, 41, dtype)
    w = _mk((Cout, K * K * Cin), (K * K * Cin) ** -0.5, 42, dtype)
    gamma = _mk((Cout,), 0.3, 43, "f32") + 1.0
    beta = _mk((Cout,), 0.3, 44, "f32")
    Ho = (HW + 2 * (K // 2) - K) // stride + 1
    r = _mk((NB * Ho * Ho, Cout), 1.0, 45, dtype) if res else None
    out = torch.zeros((NB * Ho * Ho, Cout), dtype=H16[dtype], device="cuda")
    outs = []
    for _ in range(2):
        check(lib(dtype).rvb_conv_gemm_gn(P(x), NB, HW, HW, Cin, P(w), Cout, K, stride, K // 2, P(gamma), P(beta), 16, int(relu),
                                          P(r), P(out), stream()), "rvb_conv_gemm_gn", dtype)
        torch.cuda.synchronize()
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])          # fixed-order statistics: bit-reproducible
    w4 = w.float().view(Cout, K, K, Cin).permute(0, 3, 1, 2).contiguous()
    y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w4, None, stride=stride, padding=K // 2)
    y = torch.nn.functional.group_norm(y, 16, gamma, beta, 1e-5).permute(0, 2, 3, 1).reshape(-1, Cout)
    if r is not None:
        y = y + r.float()
    if relu:
        y = torch.relu(y)
    assert rel_err(out, y) < OUT_TOL[dtype]


def test_output_column_slice_and_pitch():
    """GEMM epilogues write straight into column slices of the LSTM input (ldc > N)."""
    from tests.gpu_util import OUT_TOL, conv_gemm, conv_ref, rel_err

    M, K, N, LDC = 64, 2112, 256, 896
    x = _mk((1, 1, M, K), 1.0, 11)
    w = _mk((N, K), K ** -0.5, 12)
    buf = torch.full((M, LDC), 7.0, dtype=torch.float16, device="cuda")
    view = buf[:, 384:]
    conv_gemm(x, w, act=1, ldc=LDC, out=view)
    ref = conv_ref(x, w, act=1)
    assert rel_err(buf[:, 384:640], ref) < OUT_TOL["fp16"]
    assert torch.all(buf[:, :384] == 7.0) and torch.all(buf[:, 640:] == 7.0)


def test_fp16_epilogue_saturates_instead_of_overflowing():
    from tests.gpu_util import conv_gemm

    x = _mk((1, 1, 128, 64), 100.0, 15)
    w = _mk((64, 64), 100.0, 16)
    out = conv_gemm(x, w)
    assert torch.isfinite(out.float()).all() and float(out.float().abs().max()) == 65504.0


def test_linearity_full_size():
    """Size-independent property at BASELINE shapes (RGB layer1, batch 64): conv(2x) = 2 conv(x)
    and agreement of two tile widths, without needing a CPU reference at this size."""
    from tests.gpu_util import conv_gemm, rel_err

    x = _mk((64, 64, 64, 64), 1.0, 13)
    w = _mk((64, 9 * 64), (2.0 / 576) ** 0.5, 14)
    y1 = conv_gemm(x, w, KH=3, KW=3, pad=1, out_f32=True)
    y2 = conv_gemm((x.float() * 2).to(x.dtype), w, KH=3, KW=3, pad=1, out_f32=True)
    assert rel_err(y2, 2 * y1) < 1e-5          # scaling by 2 is exact in 16-bit floats / fp32
    y3 = conv_gemm(x, w, KH=3, KW=3, pad=1, out_f32=True, force_bn=128)
    assert rel_err(y3, y1) < 1e-5
