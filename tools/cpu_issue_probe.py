"""How long does the host take to ISSUE one policy step (no synchronisation) vs the device time?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import robovln_b200 as R
B, L = 64, 80
dev = torch.device("cuda", 0)
policy = R.HcmPolicy().share_frozen_trunks().to(dev).eval()
g = torch.Generator().manual_seed(1)
obs = {"rgb": torch.randint(0, 256, (B, 256, 256, 3), generator=g).float().to(dev),
       "depth": torch.rand((B, 256, 256, 1), generator=g).to(dev),
       "instruction": torch.randint(1000, 30522, (B, L), generator=g).float().to(dev)}
masks = torch.ones((B, 2), device=dev); hh = torch.zeros((2, B, 512), device=dev); hl = torch.zeros((2, B, 512), device=dev)
for _ in range(5): policy.act(obs, hh, hl, masks)
torch.cuda.synchronize()
K = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K): policy.act(obs, hh, hl, masks)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {1e3*(t1-t0)/K:.3f} ms/step, device {e0.elapsed_time(e1)/K:.3f} ms/step, wall incl. drain {1e3*(t2-t0)/K:.3f} ms/step")
