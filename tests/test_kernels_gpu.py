"""Non-GEMM kernels against fp32 torch references on the same (16-bit-rounded) inputs, for both
builds of the library (fp16 default / bf16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DTYPES = ["fp16", "bf16"]


def _mk(shape, scale=1.0, seed=0, dtype="fp16"):
    from tests.gpu_util import H16

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.randn(shape, generator=g, device="cuda") * scale
    return (t if dtype == "f32" else t.to(H16[dtype])).contiguous()


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C,G,HW,NB", [(32, 16, 4096, 2), (64, 16, 1024, 3), (128, 16, 1024, 2), (256, 16, 256, 2),
                                         (512, 16, 64, 3), (1024, 16, 16, 5), (128, 1, 16, 4)])
def test_groupnorm(C, G, HW, NB, dtype):
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, stream

    x = (_mk((NB, HW, C), 2.0, 1, "f32") + 0.5).to(H16[dtype])
    gamma = _mk((C,), 0.3, 2, "f32") + 1.0
    beta = _mk((C,), 0.2, 3, "f32")
    res = _mk((NB, HW, C), 1.0, 4, dtype)
    stats = torch.empty((NB, G, 2), dtype=torch.float32, device="cuda")
    out = torch.empty((NB, HW, C), dtype=H16[dtype], device="cuda")
    check(lib(dtype).rvb_groupnorm(P(x), P(stats), P(gamma), P(beta), NB, HW, C, G, 1, P(res), P(out), C, stream()),
          "rvb_groupnorm", dtype)
    torch.cuda.synchronize()
    xr = x.float().permute(0, 2, 1)  # [NB, C, HW]
    ref = torch.relu(F.group_norm(xr, G, gamma, beta, 1e-5) + res.float().permute(0, 2, 1)).permute(0, 2, 1)
    assert _rel(out, ref) < OUT_TOL[dtype]
    # fused statistics + apply kernel (stats = NULL): what the depth trunk runs; must agree bit for bit
    # from run to run (fixed-order reductions) and with the two-kernel path up to fp32 summation order
    out2 = torch.empty_like(out)
    out3 = torch.empty_like(out)
    for o in (out2, out3):
        check(lib(dtype).rvb_groupnorm(P(x), None, P(gamma), P(beta), NB, HW, C, G, 1, P(res), P(o), C, stream()),
              "rvb_groupnorm(fused)", dtype)
    torch.cuda.synchronize()
    assert _rel(out2, ref) < OUT_TOL[dtype]
    assert torch.equal(out2, out3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("D,M", [(768, 1600), (256, 333)])
def test_layernorm(D, M, dtype):
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, stream

    x = _mk((M, D), 1.5, 5, "f32") + 0.3
    g = _mk((D,), 0.2, 6, "f32") + 1.0
    b = _mk((D,), 0.2, 7, "f32")
    pe = _mk((80, D), 1.0, 8, "f32")
    out = torch.empty((M, D), dtype=H16[dtype], device="cuda")
    check(lib(dtype).rvb_layernorm(P(x), M, D, P(g), P(b), 1e-5, P(pe), 80, P(out), stream()), "rvb_layernorm", dtype)
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (D,), g, b, 1e-5) + pe[torch.arange(M, device="cuda") % 80]
    assert _rel(out, ref) < OUT_TOL[dtype]
    check(lib(dtype).rvb_layernorm(P(x), M, D, P(g), P(b), 1e-12, None, 0, P(out), stream()), "rvb_layernorm", dtype)
    torch.cuda.synchronize()
    assert _rel(out, F.layer_norm(x, (D,), g, b, 1e-12)) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("impl", ["mma_sync", "tcgen05"])
@pytest.mark.parametrize("L,R", [(20, 3), (80, 4), (12, 2), (128, 2), (200, 1), (1, 2), (80, 64)])
def test_bert_attention(L, R, impl, dtype):
    """Both self-attention kernels: warp-level mma.sync (the engine's default) and the tcgen05 / TMEM / TMA kernel
    (L <= 128; V as an MN-major operand) against torch."""
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, stream

    if impl == "tcgen05" and L > 128:
        pytest.skip("the tcgen05 kernel holds one <= 128-key tile")
    heads = 12
    qkv = _mk((R * L, 3 * heads * 64), 1.0, 9, dtype)
    ctx = torch.empty((R * L, heads * 64), dtype=H16[dtype], device="cuda")
    fn = lib(dtype).rvb_bert_attention if impl == "mma_sync" else lib(dtype).rvb_bert_attention_tc
    check(fn(P(qkv), P(ctx), R, L, heads, stream()), "rvb_bert_attention[%s]" % impl, dtype)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(R, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    ref = (s @ v).permute(0, 2, 1, 3).reshape(R * L, heads * 64)
    # probabilities are rounded to 16 bits before P.V: 1.5x the output-rounding tolerance
    assert _rel(ctx, ref) < 1.5 * OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,L,shared", [(3, 20, False), (4, 80, False), (5, 12, True)])
def test_vla_attention(B, L, shared, dtype):
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, stream

    qrows = L if shared else B * L
    q = _mk((qrows, 256), 1.0, 10, dtype)
    kv = _mk((B * 16, 512), 1.0, 11, dtype)
    ctx = torch.empty((B * L, 256), dtype=H16[dtype], device="cuda")
    check(lib(dtype).rvb_vla_attention(P(q), P(kv), P(ctx), B, L, qrows, stream()), "rvb_vla_attention", dtype)
    torch.cuda.synchronize()
    qf = q.float().view(1 if shared else B, L, 4, 64).expand(B, L, 4, 64).permute(0, 2, 1, 3)
    kf = kv.float()[:, :256].view(B, 16, 4, 64).permute(0, 2, 3, 1)
    vf = kv.float()[:, 256:].view(B, 16, 4, 64).permute(0, 2, 1, 3)
    ref = (torch.softmax(qf @ kf / 8.0, -1) @ vf).permute(0, 2, 1, 3).reshape(B * L, 256)
    assert _rel(ctx, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T,N,zero_rows", [(1, 1, ()), (1, 64, (3, 17)), (5, 1, (0, 3)), (7, 3, (0, 1, 2, 10)),
                                            (64, 1, (0,))])
def test_lstm(T, N, zero_rows, dtype):
    from tests.gpu_util import P, check, lib, stream

    H = 512
    gx = _mk((T * N, 4 * H), 1.0, 12, "f32")
    whh = _mk((4 * H, H), H ** -0.5, 13, dtype)
    hc = _mk((2, N, H), 0.5, 14, "f32")
    masks = torch.ones((T * N, 2), device="cuda")
    for r in zero_rows:
        masks[r] = 0
    hc_out = torch.empty_like(hc)
    scratch = torch.empty((2, N, H), device="cuda")
    y = torch.empty((T * N, H), device="cuda")
    check(lib(dtype).rvb_lstm(P(gx), P(whh), P(masks), 2, P(hc), P(hc_out), P(scratch), P(y), T, N, stream()),
          "rvb_lstm", dtype)
    torch.cuda.synchronize()
    h, c = hc[0].clone(), hc[1].clone()
    m = masks[:, 0].view(T, N)
    outs = []
    W = whh.float()
    for t in range(T):
        if t == 0 or bool((m[t] == 0).any()):
            h = h * m[t].view(N, 1)
            c = c * m[t].view(N, 1)
        g = gx.view(T, N, -1)[t] + h @ W.t()
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    ref_y = torch.stack(outs).view(T * N, H)
    assert float((y - ref_y).abs().max()) < 2e-4
    assert float((hc_out[0] - h).abs().max()) < 2e-4
    assert float((hc_out[1] - c).abs().max()) < 2e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_and_stems(dtype):
    from tests.gpu_util import H16, OUT_TOL, P, check, lib, stream

    L = lib(dtype)
    x = _mk((2, 64, 64, 32), 1.0, 15, dtype)
    out = torch.empty((2, 32, 32, 32), dtype=H16[dtype], device="cuda")
    check(L.rvb_maxpool3x3s2(P(x), P(out), 2, 64, 64, 32, stream()), "rvb_maxpool3x3s2", dtype)
    torch.cuda.synchronize()
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(out.float(), ref)

    # RGB stem im2col: equals unfold of the /255 image (k = r*21 + s*3 + c)
    g = torch.Generator(device="cuda")
    g.manual_seed(16)
    rgb = torch.randint(0, 256, (2, 64, 96, 3), generator=g, device="cuda").float()
    Ho, Wo = 32, 48
    col = torch.full((2 * Ho * Wo, 160), 9.0, dtype=H16[dtype], device="cuda")
    check(L.rvb_rgb_stem_im2col(P(rgb), P(col), 2, 64, 96, 160, stream()), "rvb_rgb_stem_im2col", dtype)
    torch.cuda.synchronize()
    img = (rgb / 255.0).permute(0, 3, 1, 2)
    unf = F.unfold(img, 7, padding=3, stride=2)               # [2, 3*49, Ho*Wo] with index c*49 + r*7 + s
    unf = unf.view(2, 3, 7, 7, Ho * Wo).permute(0, 4, 2, 3, 1).reshape(2 * Ho * Wo, 147)
    assert torch.equal(col[:, :147].float(), unf.to(H16[dtype]).float())
    assert torch.all(col[:, 147:] == 0)

    # depth stem: avg_pool2d(2) -> conv7x7 s2 p3 (1 -> 32)
    depth = torch.rand((2, 256, 256, 1), generator=g, device="cuda")
    w = _mk((32, 49), 0.2, 17, "f32")
    o = torch.empty((2, 64, 64, 32), dtype=H16[dtype], device="cuda")
    check(L.rvb_depth_stem(P(depth), P(w), P(o), 2, 256, 256, stream()), "rvb_depth_stem", dtype)
    torch.cuda.synchronize()
    ref = F.conv2d(F.avg_pool2d(depth.permute(0, 3, 1, 2), 2), w.view(32, 1, 7, 7), None, 2, 3).permute(0, 2, 3, 1)
    assert _rel(o, ref) < OUT_TOL[dtype]
