"""Time a few epilogue-bound GEMM shapes of the RGB trunk (CUDA events, min of 7), for experiments
with ROBOVLN_EPI_DEBUG / ROBOVLN_RES_TMA etc.  Prints one line per case."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.gpu_util import conv_gemm

def mk(shape, scale, seed, dt=torch.float16):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dt).contiguous()

cases = []
def add(name, M, K, N, res, act=1, KH=1, img=None):
    if img is None:
        x = mk((1, 1, M, K), 1.0, 1)
    else:
        x = mk(img, 1.0, 1)
    w = mk((N, KH * KH * K), (KH * KH * K) ** -0.5, 2); b = mk((N,), 0.1, 3, torch.float32)
    r = mk((M, N), 1.0, 4) if res else None
    o = torch.zeros((M, N), dtype=torch.float16, device="cuda")
    byt = x.numel() * 2 + M * N * 2 * (2 if res else 1)
    cases.append((name, lambda: conv_gemm(x, w, KH=KH, KW=KH, pad=KH // 2, bias=b, res=r, act=act, out=o), byt, 2.0 * M * N * K * KH * KH))

add("l1c3 K=64 N=256 +res", 262144, 64, 256, True)
add("l1ds K=64 N=256", 262144, 64, 256, False)
add("l1c1 K=256 N=64", 262144, 256, 64, False)
add("l1c2 3x3 K=576 N=64", 262144, 64, 64, False, KH=3, img=(64, 64, 64, 64))
add("l2c3 K=128 N=512 +res", 65536, 128, 512, True)
add("l3c3 K=256 N=1024 +res", 16384, 256, 1024, True)
add("l3c1 K=1024 N=256", 16384, 1024, 256, False)
add("l4c3 K=512 N=2048 +res", 4096, 512, 2048, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn, byt, fl in cases:
    fn(); fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"{name:26s} {t*1e3:7.1f} us  {byt/t/1e6:7.0f} GB/s  {fl/t/1e9:7.1f} TF/s", flush=True)
