"""Where a cfg5 DAgger update (trainer.DaggerUpdater, CUDA-graph tails) spends its time: wall clock with a device
synchronize after every section."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import robovln_b200 as R
from robovln_b200 import torch_tail

dev = torch.device("cuda")
T, L = 64, 80
policy = R.HcmPolicy().share_frozen_trunks().to(dev)
hi, lo = policy.high_level, policy.low_level
hi.train(); lo.train()
g = torch.Generator().manual_seed(5)
rgb = torch.randint(0, 256, (T, 256, 256, 3), generator=g).float().to(dev)
depth = torch.rand((T, 256, 256, 1), generator=g).to(dev)
ids = torch.randint(1000, 30522, (1, L), generator=g).float().to(dev)
masks = torch.ones((T, 2), device=dev); masks[0] = 0.0
tgt_act = torch.rand((T, 2), generator=g).to(dev)
tgt_stop = (torch.rand((T, 1), generator=g) > 0.9).float().to(dev)
sensor = (torch.randint(0, 4, (T,), generator=g) + 1).float().view(T, 1).to(dev)
opt_hi = R.optim.FusedAdamW([p for p in hi.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3)
opt_lo = R.optim.FusedAdam([p for p in lo.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-3)
h0 = torch.zeros((2, 1, 512), device=dev)
prev = torch.zeros((T, 2), device=dev)
upd = R.trainer.DaggerUpdater(hi, lo, opt_hi, opt_lo, graph=True)
def call():
    obs = {"rgb": rgb, "depth": depth, "instruction": ids, "vln_oracle_action_sensor": sensor}
    return upd.update(obs, prev, masks, tgt_act, tgt_stop, h0, h0, None)
for _ in range(5): call()
rt = hi.runtime()
acc = {}
def sec(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3
    return r
graphs = {k[0]: v for k, v in upd._graphs.items()}
n = 10
for _ in range(n):
    feats = sec("encode (engine + 4 buffer copies)", lambda: rt.encode(rgb, depth, ids, n_envs=1))
    sec("hi graph replay (incl. input copies)", lambda: graphs["hi"].replay(dict(rgb_feat=feats["rgb_feat"], depth_feat=feats["depth_feat"], bert=feats["bert"], hidden=h0, masks=masks, sensor=sensor.view(-1))))
    sec("hi optimizer", lambda: opt_hi.step())
    f2 = sec("encode lo (trunks reused)", lambda: lo.runtime().encode(rgb, depth, None, n_envs=1, use_lo_weights=True))
    sec("lo graph replay", lambda: graphs["lo"].replay(dict(rgb_gmean=f2["rgb_gmean"], depth_feat=f2["depth_feat"], hidden=h0, masks=masks, sub_goal=(sensor.view(-1).long() - 1), corrected=tgt_act, oracle_stop=tgt_stop)))
    sec("lo optimizer", lambda: opt_lo.step())
print(json.dumps({k: round(v / n, 3) for k, v in acc.items()}), "total", round(sum(acc.values()) / n, 3))
