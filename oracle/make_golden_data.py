"""Golden vectors for the tbptt data path -- TEST INFRASTRUCTURE, not product code.

The reference's `collate_fn`, `_block_shuffle`, `IWTrajectoryDataset.__next__/_load_next` ordering logic
(robo_vln_baselines/hierarchical_trainer.py:66-274) and `split_batch_tbptt` (robo_vln_baselines/common/utils.py:120-144)
cannot be imported as modules here (their files import habitat / lmdb / tensorflow at the top).  This script reads the
UNMODIFIED source files under /root/reference, extracts exactly those definitions with `ast`, executes them in a
namespace that provides only torch / numpy / random / defaultdict, feeds them seeded synthetic trajectories, and writes
what they return to tests/golden/data_path.npz.  Nothing from the reference is copied into the repo: only the outputs.

    python oracle/make_golden_data.py           (needs /root/reference; run in the build container, not on the GPU box)
"""
from __future__ import annotations

import ast
import os
import random
import sys
from collections import defaultdict

import numpy as np
import torch

REF = "/root/reference/robo_vln_baselines"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def extract(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "random": random, "defaultdict": defaultdict}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = ast.get_source_segment(src, node)
            exec(compile(code, path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def synthetic_episode(g: torch.Generator, T: int, L: int, hw: int = 8):
    """One trajectory in the layout IWTrajectoryDataset.__next__ hands to collate_fn (small frames: the functions are
    shape-agnostic).  RGB holds integral 0..255 values as float32 (what batch_obs_data_collect stores)."""
    obs = {
        "rgb": torch.randint(0, 256, (T, hw, hw, 3), generator=g).float(),
        "depth": torch.rand((T, hw, hw, 1), generator=g),
        "vln_oracle_action_sensor": torch.randint(0, 5, (T, 1), generator=g).float(),
        "instruction": torch.randint(1, 30000, (1, L), generator=g).float(),
    }
    prev_actions = torch.rand((T, 1, 2), generator=g, dtype=torch.float64)
    oracle_actions = torch.rand((T, 1, 2), generator=g, dtype=torch.float64)
    oracle_stop = (torch.rand((T, 1), generator=g) > 0.8).float()
    return obs, prev_actions, oracle_actions, oracle_stop


def make_batch(seed, lens, ilens):
    g = torch.Generator().manual_seed(seed)
    return [synthetic_episode(g, T, L) for T, L in zip(lens, ilens)]


INGEST_VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "walk", "past", "the", "sofa", "and", "stop", "at", "door", "turn",
                "left", "right", "##s", "##ing", "kitchen", ",", "."]
INGEST_TEXTS = ["Walk past the sofa and stop at the door.", "Turn left, walking past the kitchen doors.", "unknownword turns right"]


def ingest_observation(i: int, text: str):
    """One simulator observation as habitat hands it to transform_obs (uint8 RGB frame, float depth, instruction dict)."""
    g = np.random.default_rng(100 + i)
    return {"rgb": g.integers(0, 256, (8, 8, 3), dtype=np.uint8), "depth": g.random((8, 8, 1), dtype=np.float32),
            "instruction": {"text": text, "tokens": [int(x) for x in g.integers(1, 50, 6)]}, "progress": np.float32(0.25 * i)}


CASES = {"b1": ([7], [5]), "b3_ragged": ([5, 9, 3], [6, 4, 8]), "b2_equal": ([4, 4], [3, 3])}


def main():
    tr = extract(os.path.join(REF, "hierarchical_trainer.py"), ["ObservationsDict", "collate_fn", "_block_shuffle"])
    ut = extract(os.path.join(REF, "common", "utils.py"), ["split_batch_tbptt"])
    out = {}
    for name, (lens, ilens) in CASES.items():
        batch = make_batch(sum(map(ord, name)), lens, ilens)
        obs, pa, nd, ca, os_ = tr["collate_fn"](batch)
        for k, v in obs.items():
            out[f"{name}.obs.{k}"] = v.numpy()
        out[f"{name}.prev_actions"] = pa.numpy()
        out[f"{name}.not_done"] = nd.numpy()
        out[f"{name}.corrected"] = ca.numpy()
        out[f"{name}.oracle_stop"] = os_.numpy()
        # tbptt split exactly as train() calls it (hierarchical_trainer.py:741-750): per-sensor [B*T, ...] tensors, split_dim 0
        for steps in (2, 4, 100):
            chunks = ut["split_batch_tbptt"](obs, pa, nd, ca, os_, steps, 0)
            out[f"{name}.tbptt{steps}.n"] = np.array(len(chunks))
            for i, (o, a, m, c, s) in enumerate(chunks):
                for k, v in o.items():
                    out[f"{name}.tbptt{steps}.{i}.obs.{k}"] = v.numpy()
                out[f"{name}.tbptt{steps}.{i}.prev_actions"] = a.numpy()
                out[f"{name}.tbptt{steps}.{i}.not_done"] = m.numpy()
                out[f"{name}.tbptt{steps}.{i}.corrected"] = c.numpy()
                out[f"{name}.tbptt{steps}.{i}.oracle_stop"] = s.numpy()
    # ordering logic: _block_shuffle under a fixed Python RNG seed
    for n, bs in ((10, 3), (7, 1), (100, 16)):
        random.seed(1234 + n)
        out[f"block_shuffle.{n}.{bs}"] = np.array(tr["_block_shuffle"](list(range(n)), bs))
    # observation ingest (common/utils.py:18-118) with a tiny WordPiece vocabulary at the path the reference hard-codes
    import tempfile

    from tokenizers import BertWordPieceTokenizer

    ut2 = extract(os.path.join(REF, "common", "utils.py"), ["get_bert_tokens", "_to_tensor", "batch_obs_data_collect", "batch_obs", "transform_obs"])
    ut2["BertWordPieceTokenizer"] = BertWordPieceTokenizer
    ut2["List"], ut2["Dict"], ut2["Optional"] = __import__("typing").List, __import__("typing").Dict, __import__("typing").Optional
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "vocab_files"))
        with open(os.path.join(tmp, "vocab_files", "bert-base-uncased-vocab.txt"), "w") as fh:
            fh.write("\n".join(INGEST_VOCAB) + "\n")
        os.chdir(tmp)
        try:
            for i, text in enumerate(INGEST_TEXTS):
                obs = ingest_observation(i, text)
                got = ut2["transform_obs"](obs, "instruction", is_bert=True)
                out[f"ingest.{i}.instruction"] = np.array(got["instruction"], dtype=np.int64)
                out[f"ingest.{i}.glove_tokens"] = np.array(got["glove_tokens"], dtype=np.int64)
                b = ut2["batch_obs"](got)
                for k, v in b.items():
                    out[f"ingest.{i}.batch.{k}"] = v.numpy()
            steps = [ut2["transform_obs"](ingest_observation(i, INGEST_TEXTS[0]), "instruction", is_bert=True) for i in range(3)]
            for k, v in ut2["batch_obs_data_collect"](steps).items():
                out[f"ingest.collect.{k}"] = v.numpy()
        finally:
            os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "data_path.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    sys.exit(main())
