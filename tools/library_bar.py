"""The "library" arm of SURVEY.md 2a / 8(d): the same policy step built from stock library kernels on the same B200,
as favourably as PyTorch allows -- channels_last fp16 torchvision ResNet-50 and a GroupNorm ResNet-50 (cuDNN), HuggingFace
BertModel with SDPA attention (cuBLAS + flash attention), cuDNN LSTM, cudnn.benchmark = True, frozen trunks run ONCE
for hi and lo (as the engine does), the whole step replayed from a CUDA graph.  Random weights: only time is compared.

Used by bench.py (`library_baseline` key of the N=1 line) and tests/test_library_bar_gpu.py.  Independent of oracle/.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class GNBottleneck(nn.Module):
    def __init__(self, cin, mid, stride, groups=16):
        super().__init__()
        cout = mid * 4
        self.c1, self.g1 = nn.Conv2d(cin, mid, 1, bias=False), nn.GroupNorm(groups, mid)
        self.c2, self.g2 = nn.Conv2d(mid, mid, 3, stride, 1, bias=False), nn.GroupNorm(groups, mid)
        self.c3, self.g3 = nn.Conv2d(mid, cout, 1, bias=False), nn.GroupNorm(groups, cout)
        self.ds = None
        if stride != 1 or cin != cout:
            self.ds = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.GroupNorm(groups, cout))

    def forward(self, x):
        y = F.relu(self.g1(self.c1(x)))
        y = F.relu(self.g2(self.c2(y)))
        y = self.g3(self.c3(y))
        return F.relu(y + (x if self.ds is None else self.ds(x)))


class DepthResNet50GN(nn.Module):
    """DDPPO ResNet-50, base planes 32, GroupNorm(16) + compression conv (habitat_baselines/rl/ddppo/policy/resnet.py)."""

    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(1, 32, 7, 2, 3, bias=False), nn.GroupNorm(16, 32), nn.ReLU(True), nn.MaxPool2d(3, 2, 1))
        layers, cin = [], 32
        for li, nb in enumerate((3, 4, 6, 3)):
            mid = 32 << li
            for b in range(nb):
                layers.append(GNBottleneck(cin, mid, 2 if (b == 0 and li > 0) else 1))
                cin = mid * 4
        self.layers = nn.Sequential(*layers)
        self.comp = nn.Sequential(nn.Conv2d(cin, 128, 3, 1, 1, bias=False), nn.GroupNorm(1, 128), nn.ReLU(True))

    def forward(self, depth):          # [B,1,256,256]
        return self.comp(self.layers(self.stem(F.avg_pool2d(depth, 2))))


class CrossModal(nn.Module):
    def __init__(self):
        super().__init__()
        self.ins_fc, self.vis_fc = nn.Linear(768, 256), nn.Linear(256, 256)
        self.ln0, self.ln1, self.ln2 = nn.LayerNorm(256), nn.LayerNorm(256), nn.LayerNorm(256)
        self.q, self.k, self.v, self.o = (nn.Linear(256, 256) for _ in range(4))
        self.fc1, self.fc2 = nn.Linear(256, 1024), nn.Linear(1024, 256)

    def forward(self, bert, kv, pe):   # bert [B,L,768], kv [B,16,256]
        B, L, _ = bert.shape
        q0 = self.ln0(F.relu(self.ins_fc(bert))) + pe
        v_ = self.ln0(F.relu(self.vis_fc(kv)))
        q = self.q(q0).view(B, L, 4, 64).transpose(1, 2)
        k = self.k(v_).view(B, 16, 4, 64).transpose(1, 2)
        v = self.v(v_).view(B, 16, 4, 64).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, 256)
        x = self.ln1(q0 + self.o(o))
        return self.ln2(x + self.fc2(F.relu(self.fc1(x)))).mean(1)


class LibraryPolicy(nn.Module):
    def __init__(self):
        super().__init__()
        import torchvision
        from transformers import BertConfig, BertModel

        self.rgb = torchvision.models.resnet50(weights=None)
        self.rgb.fc = nn.Identity()
        self.rgb.avgpool = nn.Identity()
        self.depth = DepthResNet50GN()
        try:
            self.bert = BertModel(BertConfig(), add_pooling_layer=False, attn_implementation="sdpa")
        except TypeError:
            self.bert = BertModel(BertConfig(), add_pooling_layer=False)
        self.rgb_emb, self.depth_emb = nn.Parameter(torch.randn(1, 64, 4, 4)), nn.Parameter(torch.randn(1, 64, 4, 4))
        self.rgb_kv, self.depth_kv = nn.Conv1d(2112, 256, 1), nn.Conv1d(192, 256, 1)
        self.rgb_linear, self.depth_linear = nn.Linear(2112, 256), nn.Linear(3072, 128)
        self.cm = CrossModal()
        self.lstm_hi, self.head_hi = nn.LSTM(896, 512), nn.Linear(512, 4)
        self.lo_depth_fc, self.lo_rgb_fc = nn.Linear(2048, 128), nn.Linear(2048, 256)
        self.sub_emb = nn.Embedding(5, 32, padding_idx=4)
        self.lstm_lo, self.head_lo, self.stop_lo = nn.LSTM(416, 512), nn.Linear(512, 2), nn.Linear(512, 1)

    def rgb_features(self, x):   # torchvision forward without avgpool / fc
        m = self.rgb
        x = m.maxpool(m.relu(m.bn1(m.conv1(x))))
        return m.layer4(m.layer3(m.layer2(m.layer1(x))))

    def forward(self, rgb, depth, ids, masks, h_hi, h_lo, pe):
        B = rgb.shape[0]
        x = (rgb.permute(0, 3, 1, 2) / 255.0).contiguous(memory_format=torch.channels_last)
        r4 = self.rgb_features(x)                                        # [B,2048,8,8]
        d4 = self.depth(depth.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last))   # [B,128,4,4]
        bert = self.bert(input_ids=ids).last_hidden_state
        r = torch.cat([F.adaptive_avg_pool2d(r4, 4), self.rgb_emb.expand(B, -1, -1, -1)], 1).flatten(2)    # [B,2112,16]
        d = torch.cat([d4, self.depth_emb.expand(B, -1, -1, -1)], 1).flatten(2)                             # [B,192,16]
        ar = self.cm(bert, self.rgb_kv(r).transpose(1, 2), pe)
        ad = self.cm(bert, self.depth_kv(d).transpose(1, 2), pe)
        xi = torch.cat([F.relu(self.rgb_linear(r.mean(2))), F.relu(self.depth_linear(d.flatten(1))), ar, ad], 1)
        m = masks[:, :1].view(1, B, 1)
        y, (hh, ch) = self.lstm_hi(xi.unsqueeze(0), (h_hi[:1] * m, h_hi[1:] * m))
        logits = self.head_hi(y[0])
        sub = logits.argmax(1)
        xl = torch.cat([F.relu(self.lo_depth_fc(d4.flatten(1))), F.relu(self.lo_rgb_fc(r4.mean((2, 3)))), self.sub_emb(sub)], 1)
        y2, (hl, cl) = self.lstm_lo(xl.unsqueeze(0), (h_lo[:1] * m, h_lo[1:] * m))
        return logits, self.head_lo(y2[0]), self.stop_lo(y2[0]), hh, hl


def measure(B: int = 64, L: int = 80, steps: int = 20, warmup: int = 5, device="cuda"):
    """-> dict(ms_per_step eager / graphed, obs_per_s) of the library arm on `device`."""
    torch.backends.cudnn.benchmark = True
    dev = torch.device(device)
    g = torch.Generator().manual_seed(0)
    pol = LibraryPolicy().to(dev).half().eval().to(memory_format=torch.channels_last)
    rgb = torch.randint(0, 256, (B, 256, 256, 3), generator=g).to(dev).half()
    depth = torch.rand((B, 256, 256, 1), generator=g).to(dev).half()
    ids = torch.randint(1000, 30522, (B, L), generator=g).to(dev)
    masks = torch.ones((B, 2), device=dev, dtype=torch.half)
    h_hi = torch.zeros((2, B, 512), device=dev, dtype=torch.half)
    h_lo = torch.zeros((2, B, 512), device=dev, dtype=torch.half)
    pos = torch.arange(L, device=dev).float().unsqueeze(1)
    div = torch.pow(10000.0, torch.arange(0, 256, 2, device=dev).float() / 256)
    pe = torch.zeros((L, 256), device=dev)
    pe[:, 0::2], pe[:, 1::2] = torch.sin(pos / div), torch.cos(pos / div)
    pe = pe.half()

    def step():
        with torch.no_grad():
            return pol(rgb, depth, ids, masks, h_hi, h_lo, pe)

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(max(warmup, 3)):
        out = step()
    eager_ms = timed(step, steps)
    res = {"batch": B, "seq_len": L, "eager_ms_per_step": eager_ms, "dtype": "fp16 channels_last, cudnn.benchmark, SDPA, trunks once",
           "outputs_finite": bool(all(torch.isfinite(o.float()).all() for o in out))}
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        res["graphed_ms_per_step"] = timed(graph.replay, steps)
    except Exception as exc:       # report the eager number if capture is not possible
        res["graph_error"] = str(exc)[:200]
    best = min(res.get("graphed_ms_per_step", math.inf), eager_ms)
    res["ms_per_step"] = best
    res["value"] = B / (best * 1e-3)
    res["unit"] = "obs/s"
    return res


if __name__ == "__main__":
    import json

    print(json.dumps(measure()))
