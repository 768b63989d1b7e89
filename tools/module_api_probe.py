"""The reference trainer's rollout calls (hierarchical_trainer.py:1096-1100) through the drop-in modules:
out, h = hi(batch); pred = out.argmax(1); act, stop, h2 = lo(batch + pred).  Device time per step at B=64."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import robovln_b200 as R
B, L = 64, 80
dev = torch.device("cuda", 0)
pol = R.HcmPolicy().share_frozen_trunks().to(dev).eval(); hi, lo = pol.high_level, pol.low_level
g = torch.Generator().manual_seed(1)
rgb = torch.randint(0, 256, (B, 256, 256, 3), generator=g).float().to(dev); depth = torch.rand((B, 256, 256, 1), generator=g).to(dev)
ids = torch.randint(1000, 30522, (B, L), generator=g).float().to(dev)
masks = torch.ones((B, 2), device=dev); hh = torch.zeros((2, B, 512), device=dev); hl = torch.zeros((2, B, 512), device=dev)
def step():
    obs = {"rgb": rgb, "depth": depth, "instruction": ids}
    with torch.no_grad():
        out, h1 = hi((obs, hh, None, masks)); pred = out.argmax(dim=1)
        act, stop, h2 = lo((obs, hl, None, masks, pred))
    return out, act, stop
ref = [t.clone() for t in step()]
for _ in range(4): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): o = step()
e1.record(); torch.cuda.synchronize()
print(f"module API (hi -> argmax -> lo): {e0.elapsed_time(e1)/30:.3f} ms/step; replays equal first call: {all(torch.equal(a, b) for a, b in zip(o, ref))}")
