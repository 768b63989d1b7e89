import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.gpu_util import conv_gemm
PAIR = 256 | (1 << 16)
def mk(shape, scale, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).half().contiguous()
for (M, K, N) in [(1024, 768, 768), (5120, 768, 2304), (5120, 3072, 768)]:
    x = mk((1, 1, M, K), 1.0, 1); w = mk((N, K), K ** -0.5, 2)
    for fb, nm in ((128, "bn128"), (256, "bn256"), (PAIR, "pair")):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            conv_gemm(x, w, force_bn=fb)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"M={M} K={K} N={N} {nm} rep{rep}: {dt*1e3:.3f} ms", flush=True)
