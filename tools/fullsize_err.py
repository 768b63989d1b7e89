import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import robovln_b200 as R
from oracle import weights as W
from oracle.make_golden import CASES
hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1); lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
hi.load_state_dict(W.make_state_dict("hi", 0)); lo.load_state_dict(W.make_state_dict("lo", 0)); hi.cuda().eval(); lo.cuda().eval()
gold = np.load("/root/repo/tests/golden/cfg2_b64_l80.npz")
inp = W.make_inputs(**CASES["cfg2_b64_l80"]); dev = "cuda"
obs = {"rgb": inp["rgb"].to(dev), "depth": inp["depth"].to(dev), "instruction": inp["instruction"].to(dev)}
with torch.no_grad():
    logits, hh = hi((obs, inp["hidden_hi"].to(dev), inp["prev_actions"].to(dev), inp["masks"].to(dev)))
    rt = hi.runtime(); rin = rt.get_buffer("hi_rnn_in").float().cpu().numpy(); rout = rt.get_buffer("hi_rnn_out").float().cpu().numpy()
    act, stop, hl = lo((obs, inp["hidden_lo"].to(dev), inp["prev_actions"].to(dev), inp["masks"].to(dev), inp["sub_goal"].to(dev)))
def e(a, b): return float(np.abs(a - b).max())
print("rnn_in abs err", e(rin, gold["hi.rnn_in"]), "max", float(np.abs(gold["hi.rnn_in"]).max()))
err = np.abs(rout - gold["hi.rnn_out"]).max(axis=1); print("rnn_out err per step:", np.round(err, 4).tolist())
print("logits", e(logits.cpu().numpy(), gold["hi.logits"]), "act", e(act.cpu().numpy(), gold["lo.actions"]), "stop", e(stop.cpu().numpy(), gold["lo.stop"]))
print("h hi", e(hh[0].cpu().numpy(), gold["hi.hidden"][0]), "c hi", e(hh[1].cpu().numpy(), gold["hi.hidden"][1]), "h lo", e(hl[0].cpu().numpy(), gold["lo.hidden"][0]), "c lo", e(hl[1].cpu().numpy(), gold["lo.hidden"][1]))
