"""One policy step (BASELINE cfg2 shapes) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Env: STEP_BATCH (default 64), STEP_L (80), STEP_MS=0|1."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robovln_b200 as R

B = int(os.environ.get("STEP_BATCH", "64")); L = int(os.environ.get("STEP_L", "80"))
dev = torch.device("cuda", 0)
policy = R.HcmPolicy().share_frozen_trunks().to(dev).eval()
g = torch.Generator().manual_seed(1)
obs = {"rgb": torch.randint(0, 256, (B, 256, 256, 3), generator=g).float().to(dev),
       "depth": torch.rand((B, 256, 256, 1), generator=g).to(dev),
       "instruction": torch.randint(1000, 30522, (B, L), generator=g).float().to(dev)}
masks = torch.ones((B, 2), device=dev); hh = torch.zeros((2, B, 512), device=dev); hl = torch.zeros((2, B, 512), device=dev)
for _ in range(3):
    policy.act(obs, hh, hl, masks)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = policy.act(obs, hh, hl, masks)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(out[0].abs().sum()))
