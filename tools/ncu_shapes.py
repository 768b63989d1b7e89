"""Run a few of the trunk's GEMM shapes once each inside a cudaProfilerStart/Stop window (for ncu --set full)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.gpu_util import conv_gemm, rgb_stem_packed

def mk(shape, scale, seed, dt=torch.float16):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dt).contiguous()

cases = []
x = mk((64, 64, 64, 64), 1.0, 1); w = mk((64, 576), 576 ** -0.5, 2); b = mk((64,), 0.1, 3, torch.float32)
cases.append(("l1c2 conv3x3 N=64", lambda: conv_gemm(x, w, KH=3, KW=3, pad=1, bias=b, act=1)))
x2 = mk((1, 1, 262144, 64), 1.0, 4); w2 = mk((256, 64), 0.125, 5); r2 = mk((262144, 256), 1.0, 6); b2 = mk((256,), 0.1, 7, torch.float32)
o2 = torch.zeros((262144, 256), dtype=torch.float16, device="cuda")
cases.append(("l1c3 plain K=64 N=256 +res", lambda: conv_gemm(x2, w2, bias=b2, res=r2, act=1, out=o2)))
rgb = torch.randint(0, 256, (64, 256, 256, 3), device="cuda").float(); ww = mk((64, 256), 0.08, 8); b3 = mk((64,), 0.1, 9, torch.float32)
cases.append(("stem packed window", lambda: rgb_stem_packed(rgb, ww, b3)))
x4 = mk((1, 1, 5120, 768), 1.0, 10); w4 = mk((3072, 768), 768 ** -0.5, 11); b4 = mk((3072,), 0.1, 12, torch.float32)
o4 = torch.zeros((5120, 3072), dtype=torch.float16, device="cuda")
cases.append(("ffn1 gelu", lambda: conv_gemm(x4, w4, bias=b4, act=2, out=o4)))
x5 = mk((64, 16, 16, 256), 1.0, 13); w5 = mk((256, 2304), 2304 ** -0.5, 14)
cases.append(("l3c2 conv3x3 N=256", lambda: conv_gemm(x5, w5, KH=3, KW=3, pad=1, act=1)))
x6 = mk((1, 1, 65536, 128), 1.0, 15); w6 = mk((512, 128), 128 ** -0.5, 16); r6 = mk((65536, 512), 1.0, 17)
o6 = torch.zeros((65536, 512), dtype=torch.float16, device="cuda")
cases.append(("l2c3 plain K=128 N=512 +res", lambda: conv_gemm(x6, w6, res=r6, act=1, out=o6)))
for name, fn in cases:
    fn(); fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for name, fn in cases:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1)*1e3:.1f} us", flush=True)
torch.cuda.cudart().cudaProfilerStop()
