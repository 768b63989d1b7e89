"""Diagnostic (GPU): per-stage errors of the CUDA path vs the reference fixtures."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import robovln_b200 as R
from oracle import weights as W
from oracle.make_golden import CASES

hi = R.Seq2Seq_HighLevel_CMA(None, 4, None, 1); lo = R.Seq2Seq_LowLevel(None, 2, 4, None, 1)
hi.load_state_dict(W.make_state_dict("hi", 0)); lo.load_state_dict(W.make_state_dict("lo", 0))
hi.cuda().eval(); lo.cuda().eval()
def rep(name, got, ref):
    got = got.float().cpu().numpy(); 
    d = np.abs(got - ref)
    print(f"  {name:28s} max|ref|={np.abs(ref).max():7.3f} rms_ref={np.sqrt((ref**2).mean()):7.4f} max_err={d.max():.4e} rms_err={np.sqrt((d**2).mean()):.4e} rel2max={d.max()/np.abs(ref).max():.3e}")
for case, kw in CASES.items():
    print(case)
    gold = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    inp = W.make_inputs(**kw); dev = "cuda"
    obs = {"rgb": inp["rgb"].to(dev), "depth": inp["depth"].to(dev), "instruction": inp["instruction"].to(dev)}
    with torch.no_grad():
        logits, hid_hi = hi((obs, inp["hidden_hi"].to(dev), None, inp["masks"].to(dev)))
        rt = hi.runtime()
        m = {k: rt.get_buffer(k) for k in ("rgb_tokens", "depth_tokens", "bert", "vla_tokens", "hi_rnn_in", "hi_rnn_out")}
        act, stop, hid_lo = lo((obs, inp["hidden_lo"].to(dev), None, inp["masks"].to(dev), inp["sub_goal"].to(dev)))
        m["lo_rnn_in"] = rt.get_buffer("lo_rnn_in")
    B = inp["rgb"].shape[0]
    rep("rgb_embedding", m["rgb_tokens"].permute(0, 2, 1).reshape(B, 2112, 4, 4)[:, :2048], gold["hi.rgb_embedding"][:, :2048])
    rep("depth_embedding", m["depth_tokens"].permute(0, 2, 1).reshape(B, 192, 4, 4)[:, :128], gold["hi.depth_embedding"][:, :128])
    bert = m["bert"]; bert = bert.expand(B, -1, -1) if bert.shape[0] == 1 else bert
    rep("bert", bert, gold["hi.bert"])
    rep("ins_rgb_att_tokens", m["vla_tokens"][0], gold["hi.ins_rgb_att_tokens"])
    rep("ins_depth_att_tokens", m["vla_tokens"][1], gold["hi.ins_depth_att_tokens"])
    x = m["hi_rnn_in"]; g = gold["hi.rnn_in"]
    for nm, a, b in (("rgb_in", 0, 256), ("depth_in", 256, 384), ("ins_rgb_att", 384, 640), ("ins_depth_att", 640, 896)):
        rep("hi.rnn_in." + nm, x[:, a:b], g[:, a:b])
    x = m["lo_rnn_in"]; g = gold["lo.rnn_in"]
    for nm, a, b in (("depth", 0, 128), ("rgb", 128, 384), ("sub", 384, 416)):
        rep("lo.rnn_in." + nm, x[:, a:b], g[:, a:b])
    rep("hi.rnn_out", m["hi_rnn_out"], gold["hi.rnn_out"])
    rep("hi.logits", logits, gold["hi.logits"]); rep("hi.hidden", hid_hi, gold["hi.hidden"])
    rep("lo.actions", act, gold["lo.actions"]); rep("lo.stop", stop, gold["lo.stop"]); rep("lo.hidden", hid_lo, gold["lo.hidden"])
