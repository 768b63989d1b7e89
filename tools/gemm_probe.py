import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.gpu_util import conv_gemm
PAIR = 256 | (1 << 16)
def mk(shape, scale, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).half().contiguous()
# clocks up before the first measurement
_w = torch.randn((8192, 8192), device="cuda").half()
for _ in range(30): torch.matmul(_w, _w)
torch.cuda.synchronize()
shapes = [(8192, 8192, 8192), (5120, 768, 2304), (5120, 768, 3072), (5120, 3072, 768), (5120, 768, 768), (37888, 768, 2304)]
for (M, K, N) in shapes:
    x = mk((1, 1, M, K), 1.0, 1); w = mk((N, K), K ** -0.5, 2)
    out = torch.zeros((M, N), dtype=torch.float16, device="cuda")
    for fb, nm in ((128, "bn128"), (256, "bn256"), (PAIR, "pair")):
        for act in (0, 2):
            for _ in range(2): conv_gemm(x, w, force_bn=fb, out=out, act=act)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # NB launches enqueued back to back per measurement: the host side of the test entry point (it encodes four
            # tensor maps per call) stays ahead of the GPU, so the interval holds kernel time, not launch latency
            NB = 100
            ts = []
            for _ in range(5):
                e0.record()
                for _ in range(NB): conv_gemm(x, w, force_bn=fb, out=out, act=act, sync=False)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / NB)
            t = min(ts)
            print(f"M={M} K={K} N={N} {nm:6s} act={act}: {t*1e3:8.1f} us  {2*M*N*K/t/1e9:7.1f} TF/s", flush=True)
    a = x.view(M, K); b = w.t().contiguous()
    for _ in range(2): torch.matmul(a, b)
    ts = []
    c = torch.empty((M, N), dtype=torch.float16, device="cuda")
    for _ in range(5):
        e0.record()
        for _ in range(NB): torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / NB)
    print(f"M={M} K={K} N={N} cublas      : {min(ts)*1e3:8.1f} us  {2*M*N*K/min(ts)/1e9:7.1f} TF/s", flush=True)
