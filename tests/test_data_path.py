"""tbptt data path (robo-vln_b200/data.py) against golden outputs of the UNMODIFIED reference functions
(`collate_fn`, `_block_shuffle`: robo_vln_baselines/hierarchical_trainer.py:66-161; `split_batch_tbptt`:
robo_vln_baselines/common/utils.py:120-144), produced by oracle/make_golden_data.py -> tests/golden/data_path.npz.
Index / byte work: the bar is bit-exact."""
import os
import random

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "data_path.npz"))


def _eq(t: torch.Tensor, ref: np.ndarray, what: str):
    assert tuple(t.shape) == tuple(ref.shape), f"{what}: shape {tuple(t.shape)} vs {ref.shape}"
    assert str(t.dtype).replace("torch.", "") == str(ref.dtype), f"{what}: dtype {t.dtype} vs {ref.dtype}"
    assert np.array_equal(t.numpy(), ref), what


@pytest.mark.parametrize("case", ["b1", "b3_ragged", "b2_equal"])
def test_collate_and_tbptt_split_match_reference(case, gold):
    from oracle.make_golden_data import CASES, make_batch
    from robovln_b200 import data as D

    lens, ilens = CASES[case]
    batch = make_batch(sum(map(ord, case)), lens, ilens)
    obs, pa, nd, ca, os_ = D.collate_fn(batch)
    assert isinstance(obs, D.ObservationsDict)
    assert set(obs.keys()) == {k.split(".obs.")[1] for k in gold.files if k.startswith(case + ".obs.")}
    for k, v in obs.items():
        _eq(v, gold[f"{case}.obs.{k}"], f"obs[{k}]")
    _eq(pa, gold[f"{case}.prev_actions"], "prev_actions")
    _eq(nd, gold[f"{case}.not_done"], "not_done_masks")
    _eq(ca, gold[f"{case}.corrected"], "corrected_actions")
    _eq(os_, gold[f"{case}.oracle_stop"], "oracle_stop")
    for steps in (2, 4, 100):
        chunks = D.split_batch_tbptt(obs, pa, nd, ca, os_, steps, 0)
        assert len(chunks) == int(gold[f"{case}.tbptt{steps}.n"])
        for i, (o, a, m, c, s) in enumerate(chunks):
            for k, v in o.items():
                _eq(v, gold[f"{case}.tbptt{steps}.{i}.obs.{k}"], f"tbptt{steps}[{i}].obs[{k}]")
            _eq(a, gold[f"{case}.tbptt{steps}.{i}.prev_actions"], "prev_actions")
            _eq(m, gold[f"{case}.tbptt{steps}.{i}.not_done"], "not_done")
            _eq(c, gold[f"{case}.tbptt{steps}.{i}.corrected"], "corrected")
            _eq(s, gold[f"{case}.tbptt{steps}.{i}.oracle_stop"], "oracle_stop")
            # views, not copies
            if i == 0:
                assert o["rgb"].data_ptr() == obs["rgb"].data_ptr()


def test_collate_keeps_uint8_frames(gold):
    """RGB frames stored / collated as uint8 are the same numbers as the reference's float32 frames."""
    from oracle.make_golden_data import CASES, make_batch
    from robovln_b200 import data as D

    lens, ilens = CASES["b3_ragged"]
    batch = make_batch(sum(map(ord, "b3_ragged")), lens, ilens)
    batch8 = [({**o, "rgb": o["rgb"].to(torch.uint8)}, a, b, c) for (o, a, b, c) in batch]
    obs, *_ = D.collate_fn(batch8)
    assert obs["rgb"].dtype == torch.uint8
    assert np.array_equal(obs["rgb"].float().numpy(), gold["b3_ragged.obs.rgb"])


def test_block_shuffle_matches_reference(gold):
    from robovln_b200 import data as D

    for n, bs in ((10, 3), (7, 1), (100, 16)):
        random.seed(1234 + n)
        assert D._block_shuffle(list(range(n)), bs) == gold[f"block_shuffle.{n}.{bs}"].tolist()


def test_store_dataset_and_prefetch_round_trip(tmp_path):
    """TrajectoryStore write -> IWTrajectoryDataset (reference ordering + per-episode post-processing,
    hierarchical_trainer.py:229-255) -> PrefetchLoader on the CPU: every episode comes back exactly once with its
    frames bit-identical, oracle actions clamped to 4 after the stop step, oracle_stop set from the stop step on."""
    from robovln_b200 import data as D

    g = np.random.default_rng(5)
    store = D.TrajectoryStore(str(tmp_path / "traj"), "w")
    eps = []
    for e in range(11):
        T, L = int(g.integers(3, 9)), int(g.integers(4, 10))
        obs = {
            "rgb": g.integers(0, 256, (T, 8, 8, 3)).astype(np.float32),
            "depth": g.random((T, 8, 8, 1)).astype(np.float32),
            "vln_oracle_action_sensor": g.integers(1, 4, (T, 1)).astype(np.float32),
            "instruction": np.tile(g.integers(1, 30000, (1, L)).astype(np.float32), (T, 1)),
            "glove_tokens": np.zeros((T, 3), np.float32),
        }
        obs["rgb"][:, 0, 0, 0] = e                      # tag
        pa, oa = g.random((T, 1, 2)), g.random((T, 1, 2))
        stop = [0] * (T - 1) + [int(g.integers(2, T + 1))]
        store.append(obs, pa, oa, stop)
        eps.append((obs, pa, oa, stop))
    store.close()
    assert sorted(os.listdir(tmp_path / "traj"))[-1] == "index.json"
    ds = D.IWTrajectoryDataset(str(tmp_path / "traj"), use_iw=True, inflection_weight_coef=3.2, batch_size=2, is_bert=True)
    assert ds.length == 11 and ds.inflec_weights.tolist() == pytest.approx([1.0, 3.2])
    random.seed(0)
    seen = []
    for obs, pa, oa, ostop in ds:
        e = int(obs["rgb"][0, 0, 0, 0])
        seen.append(e)
        ref_obs, ref_pa, ref_oa, ref_stop = eps[e]
        assert obs["rgb"].dtype == torch.uint8 and np.array_equal(obs["rgb"].numpy(), ref_obs["rgb"].astype(np.uint8))
        assert np.array_equal(obs["depth"].numpy(), ref_obs["depth"])
        assert tuple(obs["instruction"].shape) == (1, ref_obs["instruction"].shape[1])
        val = ref_stop[-1] - 1
        exp = ref_obs["vln_oracle_action_sensor"].copy()
        exp[val:] = 4
        assert np.array_equal(obs["vln_oracle_action_sensor"].numpy(), exp)
        es = np.zeros_like(exp)
        es[val:] = 1
        assert np.array_equal(ostop.numpy(), es)
        assert np.array_equal(pa.numpy(), ref_pa) and np.array_equal(oa.numpy(), ref_oa)
    assert sorted(seen) == list(range(11))
    # batches of 2 through the prefetcher (CPU device: same code path minus the CUDA stream)
    random.seed(0)
    n_rows = 0
    for obs, pa, nd, ca, ostop in D.PrefetchLoader(ds, 2, torch.device("cpu"), drop_last=False):
        B = obs["instruction"].shape[0]
        assert obs["rgb"].shape[0] == pa.shape[0] == nd.shape[0] == ostop.shape[0] and obs["rgb"].shape[0] % B == 0
        T = obs["rgb"].shape[0] // B
        assert (nd.view(B, T, 2)[:, 0] == 0).all() and (nd.view(B, T, 2)[:, 1:] == 1).all()
        n_rows += B
    assert n_rows == 11
    with pytest.raises(ImportError, match="lmdb"):
        D.IWTrajectoryDataset(str(tmp_path / "not_a_store"), use_iw=False)


def test_observation_ingest_matches_reference(gold, tmp_path):
    """robo-vln_b200/obs.py against the UNMODIFIED common/utils.py functions (transform_obs -> batch_obs /
    batch_obs_data_collect) on the same simulator-style observations and the same tiny WordPiece vocabulary; the
    tokenizer is built once and the token ids of a repeated instruction come from the cache."""
    from oracle.make_golden_data import INGEST_TEXTS, INGEST_VOCAB, ingest_observation
    from robovln_b200 import obs as OB

    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(INGEST_VOCAB) + "\n")
    before = dict(OB.stats)
    for i, text in enumerate(INGEST_TEXTS):
        o = OB.transform_obs(ingest_observation(i, text), "instruction", is_bert=True, vocab_file=str(vocab))
        assert o["instruction"] == gold[f"ingest.{i}.instruction"].tolist()
        assert list(o["glove_tokens"]) == gold[f"ingest.{i}.glove_tokens"].tolist()
        b = OB.batch_obs(o)
        assert set(b.keys()) == {k.split(".batch.")[1] for k in gold.files if k.startswith(f"ingest.{i}.batch.")}
        for k, v in b.items():
            _eq(v, gold[f"ingest.{i}.batch.{k}"], f"batch_obs[{k}]")
        b8 = OB.batch_obs(o, keep_uint8=True)                  # the sensor's uint8 frame is kept; same numbers
        assert b8["rgb"].dtype == torch.uint8 and np.array_equal(b8["rgb"].float().numpy(), gold[f"ingest.{i}.batch.rgb"])
    steps = [OB.transform_obs(ingest_observation(i, INGEST_TEXTS[0]), "instruction", is_bert=True, vocab_file=str(vocab)) for i in range(3)]
    for k, v in OB.batch_obs_data_collect(steps).items():
        _eq(v, gold[f"ingest.collect.{k}"], f"batch_obs_data_collect[{k}]")
    assert OB.stats["tokenizer_builds"] - before["tokenizer_builds"] <= 1          # once per vocabulary, not once per step
    assert OB.stats["token_cache_hits"] - before["token_cache_hits"] >= 3           # the repeated instruction
    # non-BERT branch (GloVe token ids pass through)
    o = OB.transform_obs(ingest_observation(0, INGEST_TEXTS[0]), "instruction", is_bert=False)
    assert o["instruction"] == ingest_observation(0, INGEST_TEXTS[0])["instruction"]["tokens"] and "glove_tokens" not in o
