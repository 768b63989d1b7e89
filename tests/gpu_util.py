"""Helpers for the -m gpu tests: ctypes calls into librobovln_b200*.so with torch tensors."""
import ctypes

import torch

import robovln_b200  # noqa: F401  (alias package; makes robovln_b200._lib importable)
from robovln_b200 import _lib

H16 = {"fp16": torch.float16, "bf16": torch.bfloat16}
# relative-to-max tolerance of a 16-bit OUTPUT tensor (half an ulp at the top of the range, x2)
OUT_TOL = {"fp16": 2e-3, "bf16": 1e-2}


def lib(dtype="fp16"):
    return _lib.load(dtype=dtype)


def check(rc, what, dtype="fp16"):
    _lib.check(rc, what, lib(dtype))


def P(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_name(t: torch.Tensor) -> str:
    return "fp16" if t.dtype == torch.float16 else "bf16"


def conv_gemm(x, w, *, KH=1, KW=1, stride=1, pad=0, bias=None, res=None, res_rows=0, act=0, out_f32=False,
              force_bn=0, impl=0, ldc=None, out=None, sync=True):
    """x: [NB,H,W,Cin] h16 (contiguous), w: [Cout, KH*KW*Cin] h16 -> out [M, Cout]."""
    dt = dtype_name(x)
    NB, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    M = NB * Ho * Wo
    if ldc is None:
        ldc = Cout
    if out is None:
        out = torch.zeros((M, ldc), dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    rc = lib(dt).rvb_conv_gemm(P(x), NB, H, W, Cin, Cin, P(w), Cout, KH, KW, stride, pad, P(bias), P(res),
                               0 if res is None else res.shape[-1], res_rows, act, P(out), ldc, int(out_f32), force_bn,
                               impl, 0, 0, stream())
    check(rc, "rvb_conv_gemm", dt)
    if sync:
        torch.cuda.synchronize()
    return out


def rgb_stem(rgb, w_win, bias, dtype="fp16", impl=0):
    """RGB stem as the engine runs it: pad/convert pre-pass, then the window-mode 7x7 s2 conv.
    rgb [NB,H,W,3] f32, w_win [64, 448] h16 (weight_prep.stem_window_weights) -> [NB*Ho*Wo, 64]."""
    NB, H, W, _ = rgb.shape
    Hp, Wp = H + 6, W + 6
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    padded = torch.full((NB, Hp, Wp, 8), 3.0, dtype=H16[dtype], device=rgb.device)
    check(lib(dtype).rvb_rgb_pad_convert(P(rgb), P(padded), NB, H, W, Wp, stream()), "rvb_rgb_pad_convert", dtype)
    out = torch.zeros((NB * Ho * Wo, 64), dtype=H16[dtype], device=rgb.device)
    rc = lib(dtype).rvb_conv_gemm(P(padded), NB, Hp, Wo, 64, 8, P(w_win), 64, 7, 1, 2, 0, P(bias), None, 0, 0, 1,
                                  P(out), 64, 0, 0, impl, 1, Wp * 8, stream())
    check(rc, "rvb_conv_gemm(window)", dtype)
    torch.cuda.synchronize()
    return out, padded


def rgb_stem_packed(rgb, w_pk, bias, dtype="fp16", impl=0):
    """RGB stem as the engine runs it: row-pair-interleaved pad/convert pre-pass, then the packed window-mode conv (two
    filter rows per 64-wide K block).  rgb [NB,H,W,3] f32, w_pk [64, 256] h16
    (weight_prep.stem_packed_weights) -> [NB*Ho*Wo, 64]."""
    NB, H, W, _ = rgb.shape
    Hp, Wp = H + 6, W + 6
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    inter = torch.full((NB, Hp // 2, Wp, 2, 4), 3.0, dtype=H16[dtype], device=rgb.device)
    check(lib(dtype).rvb_rgb_pad_convert4(P(rgb), P(inter), NB, H, W, Wp, stream()), "rvb_rgb_pad_convert4", dtype)
    out = torch.zeros((NB * Ho * Wo, 64), dtype=H16[dtype], device=rgb.device)
    rc = lib(dtype).rvb_conv_gemm(P(inter), NB, Hp // 2, Wo, 64, 8, P(w_pk), 64, 4, 1, 2, 0, P(bias), None, 0, 0, 1,
                                  P(out), 64, 0, 0, impl, 2, Wp * 8, stream())
    check(rc, "rvb_conv_gemm(packed window)", dtype)
    torch.cuda.synchronize()
    padded = inter.permute(0, 1, 3, 2, 4).reshape(NB, Hp, Wp, 4)      # back to plain NHW4 for inspection
    return out, padded


def conv_ref(x, w, *, KH=1, KW=1, stride=1, pad=0, bias=None, res=None, res_rows=0, act=0):
    """fp32 torch reference on the same 16-bit-rounded operands."""
    NB, H, W, Cin = x.shape
    Cout = w.shape[0]
    if KH == 1 and KW == 1 and stride == 1 and pad == 0:
        y = x.float().reshape(-1, Cin) @ w.float().t()          # plain GEMM (avoids slow cuDNN paths for H=1 images)
    else:
        w4 = w.float().view(Cout, KH, KW, Cin).permute(0, 3, 1, 2).contiguous()
        y = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w4, None, stride=stride, padding=pad)
        y = y.permute(0, 2, 3, 1).reshape(-1, Cout)
    if bias is not None:
        y = y + bias.float()
    if res is not None:
        r = res.float().view(-1, Cout)
        if res_rows:
            idx = torch.arange(y.shape[0], device=y.device) % res_rows
            r = r[idx]
        y = y + r
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    return y


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
