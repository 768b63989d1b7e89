"""Phase timeline of gemm_tc_kernel on the BERT-shaped problems of the step.  Needs a library built with the stamps compiled in
(ROBOVLN_BUILD_STAMPS=1 python robo-vln_b200/build.py) and ROBOVLN_GEMM_TIMES=<file>:
   run:      ROBOVLN_GEMM_TIMES=gpurun_out/gemm_times.csv python tools/gemm_timeline.py run
   analyse:  python tools/gemm_timeline.py show gpurun_out/gemm_times.csv"""
import os, sys
import numpy as np

NAMES = ["kernel start", "setup done (barriers, TMEM)", "first TMA issued", "tile 0 loads issued", "first operands landed", "tile 0 MMAs issued",
         "tile 0 accumulator complete", "tile 0 epilogue done", "all epilogues done"]


def run():
    import torch
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
    from tests.gpu_util import conv_gemm
    def mk(shape, scale, seed, dt=torch.float16):
        g = torch.Generator(device="cuda"); g.manual_seed(seed)
        return (torch.randn(shape, generator=g, device="cuda") * scale).to(dt).contiguous()
    M = 5120
    x768 = mk((1, 1, M, 768), 1.0, 1); x3072 = mk((1, 1, M, 3072), 1.0, 2)
    res = mk((M, 768), 1.0, 3)
    cases = [("qkv", x768, mk((2304, 768), 768 ** -0.5, 4), dict(bias=mk((2304,), 0.1, 5, torch.float32))),
             ("ao", x768, mk((768, 768), 768 ** -0.5, 6), dict(bias=mk((768,), 0.1, 7, torch.float32), res=res, out_f32=True)),
             ("ff1", x768, mk((3072, 768), 768 ** -0.5, 8), dict(bias=mk((3072,), 0.1, 9, torch.float32), act=2)),
             ("ff2", x3072, mk((768, 3072), 3072 ** -0.5, 10), dict(bias=mk((768,), 0.1, 11, torch.float32), res=res, out_f32=True))]
    if len(sys.argv) > 2 and sys.argv[2] == "conv":
        xc = mk((64, 64, 64, 64), 1.0, 31)
        cases = [("l1conv2", xc, mk((64, 576), 576 ** -0.5, 32), dict(KH=3, KW=3, stride=1, pad=1, bias=mk((64,), 0.1, 33, torch.float32), act=1)),
                 ("l1conv1", mk((1, 1, 262144, 256), 1.0, 34), mk((64, 256), 256 ** -0.5, 35), dict(bias=mk((64,), 0.1, 36, torch.float32), act=1))]
    for name, x, w, kw in cases:
        for _ in range(3):
            conv_gemm(x, w, **kw)
    print("done")


def show(path):
    blocks, cur = [], None
    for l in open(path):
        if l.startswith("#"):
            cur = [l.strip(), []]; blocks.append(cur)
        elif l.strip():
            cur[1].append([int(v) for v in l.strip().split(",")][1:])
    for hdr, rows in blocks[2::3]:          # third (warm) launch of every case
        t = np.array(rows, dtype=np.int64)
        print(hdr)
        for i, n in enumerate(NAMES):        # SM clocks are per SM: every CTA is measured against its own start
            ok = t[:, i] != 0
            col = (t[:, i] - t[:, 0])[ok]
            if len(col): print(f"   {n:34s} min {col.min():7d}  median {int(np.median(col)):7d}  max {col.max():7d}")


if __name__ == "__main__":
    run() if sys.argv[1] == "run" else show(sys.argv[2])
