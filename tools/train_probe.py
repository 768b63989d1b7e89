"""Where does a cfg5 training step spend its time? (cuda-synchronised wall clock per phase)"""
import os, sys, time, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import robovln_b200 as R
dev = torch.device("cuda", 0); T, L = 64, 80
policy = R.HcmPolicy().share_frozen_trunks().to(dev); hi, lo = policy.high_level, policy.low_level; hi.train(); lo.train()
g = torch.Generator().manual_seed(5)
rgb = torch.randint(0, 256, (T, 256, 256, 3), generator=g).float().to(dev); depth = torch.rand((T, 256, 256, 1), generator=g).to(dev)
ids = torch.randint(1000, 30522, (1, L), generator=g).float().to(dev); masks = torch.ones((T, 2), device=dev); masks[0] = 0
tgt = torch.randint(0, 4, (T,), generator=g).to(dev); sub = torch.randint(0, 5, (T,), generator=g).to(dev)
params = [p for m in (hi, lo) for p in m.parameters() if p.requires_grad]; opt = torch.optim.AdamW(params, lr=1e-4)
def t(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    obs = {"rgb": rgb, "depth": depth, "instruction": ids}
    t0 = t(); logits, _ = hi((obs, torch.zeros((2, 1, 512), device=dev), None, masks)); t1 = t()
    F.cross_entropy(logits, tgt).backward(); t2 = t()
    act, stop, _ = lo((obs, torch.zeros((2, 1, 512), device=dev), None, masks, sub)); t3 = t()
    (act.sum() + stop.sum()).backward(); t4 = t()
    opt.step(); opt.zero_grad(); t5 = t()
    print(f"hi fwd {1e3*(t1-t0):.2f}  hi bwd {1e3*(t2-t1):.2f}  lo fwd {1e3*(t3-t2):.2f}  lo bwd {1e3*(t4-t3):.2f}  opt {1e3*(t5-t4):.2f} ms")
