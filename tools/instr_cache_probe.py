"""Rollout-style use of the instruction cache: the same 64 instructions at every step, new frames
every step.  Prints ms/step with the cache off and on.  NOT part of bench.py's metric (the benchmark step always runs BERT)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import robovln_b200 as R
B, L = 64, 80
dev = torch.device("cuda", 0)
policy = R.HcmPolicy().share_frozen_trunks().to(dev).eval()
rt = policy._runtime()
g = torch.Generator().manual_seed(1)
frames = [{"rgb": torch.randint(0, 256, (B, 256, 256, 3), generator=g, dtype=torch.uint8).to(dev), "depth": torch.rand((B, 256, 256, 1), generator=g).to(dev)} for _ in range(3)]
ids = torch.randint(1000, 30522, (B, L), generator=g).float().to(dev)
masks = torch.ones((B, 2), device=dev); hh = torch.zeros((2, B, 512), device=dev); hl = torch.zeros((2, B, 512), device=dev)
def run(n):
    for i in range(n):
        f = frames[i % 3]
        out = policy.act({"rgb": f["rgb"], "depth": f["depth"], "instruction": ids}, hh, hl, masks)
    return out
res = {}
for cache in (False, True):
    rt.instruction_cache = cache
    run(6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = run(30); e1.record(); torch.cuda.synchronize()
    res["cache_on" if cache else "cache_off"] = {"ms_per_step": e0.elapsed_time(e1) / 30, "launches": int(rt.launches()), "logit0": float(out[0][0, 0])}
res["note"] = "64 environments, uint8 frames changing every step, the same 64 instructions at every step (an episode); the cache skips BERT + the query-side projection"
print(json.dumps(res))
