"""tbptt data path of the DAgger trainer, feeding BASELINE.json configs[4].

Mirrors, with the same names, argument meaning and results:
  collate_fn            robo_vln_baselines/hierarchical_trainer.py:66-154
  _block_shuffle        robo_vln_baselines/hierarchical_trainer.py:157-161
  IWTrajectoryDataset   robo_vln_baselines/hierarchical_trainer.py:164-274   (ordering + per-episode post-processing)
  split_batch_tbptt     robo_vln_baselines/common/utils.py:120-144

What is different is how the bytes move:
  * collate_fn writes every trajectory ONCE into a preallocated (optionally pinned) [B, T_max, ...] buffer per sensor
    instead of pad -> stack -> transpose -> contiguous (four full copies of the frames), and keeps the element type it
    is given: uint8 RGB frames stay uint8 all the way to the engine, which normalises them on the GPU
    (hcm_set_rgb_format) -- a quarter of the bytes of the float32 frames the reference stores and uploads.
  * TrajectoryStore is a flat, memory-mappable on-disk format (uint8 RGB, float32 or float16 depth) standing in for the
    reference's LMDB of msgpack'd float32 arrays (hierarchical_trainer.py:466-475; ~1.5 TB, README.md:213); lmdb and
    msgpack_numpy are not part of this image, so IWTrajectoryDataset reads the reference's LMDB only where those
    modules exist.
  * PrefetchLoader collates the next batch on a worker thread into pinned buffers and uploads it on its own CUDA
    stream while the current tbptt chunks train.
"""
from __future__ import annotations

import json
import os
import queue
import random
import threading
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


class ObservationsDict(dict):
    """dict of batched sensor tensors; DataLoader(pin_memory=True) calls pin_memory() on it
    (hierarchical_trainer.py:58-63)."""

    def pin_memory(self):
        for k, v in self.items():
            if not v.is_pinned():
                self[k] = v.pin_memory()
        return self


def _alloc(shape, dtype, fill, pin: bool) -> torch.Tensor:
    t = torch.empty(shape, dtype=dtype, pin_memory=pin)
    t.fill_(fill)
    return t


def collate_fn(batch, pin: bool = False):
    """Each sample: (obs, prev_actions, oracle_actions, oracle_stop) with obs[sensor] [T_i, ...] and
    obs['instruction'] [1, L_i].  Returns, exactly as the reference:
      observations  ObservationsDict: sensor -> [B*T_max, ...] (episode-major), 'instruction' -> [B, L_max]
      prev_actions  [B*T_max, 2]      not_done_masks [B*T_max, 2] (0 on every episode's first step)
      corrected_actions [B*T_max, 2]  oracle_stop [B*T_max, 1] (padding = -1)
    Padding is zeros (observations, actions) / -1 (oracle_stop), hierarchical_trainer.py:108-126."""
    B = len(batch)
    obs0 = batch[0][0]
    t_max = max(int(s[1].size(0)) for s in batch)
    l_max = max(int(s[0]["instruction"].size(1)) for s in batch)
    observations = ObservationsDict()
    for sensor in obs0:
        first = obs0[sensor]
        if sensor == "instruction":
            out = _alloc((B, l_max), first.dtype, 0, pin)
            for b, s in enumerate(batch):
                ins = s[0][sensor]
                out[b, : ins.size(1)] = ins[0]
            observations[sensor] = out
            continue
        out = _alloc((B, t_max) + tuple(first.shape[1:]), first.dtype, 0, pin)
        for b, s in enumerate(batch):
            t = s[0][sensor]
            out[b, : t.size(0)] = t
        observations[sensor] = out.view((B * t_max,) + tuple(first.shape[1:]))

    def padded(idx, fill):
        first = batch[0][idx]
        out = _alloc((B, t_max) + tuple(first.shape[1:]), first.dtype, fill, pin)
        for b, s in enumerate(batch):
            out[b, : s[idx].size(0)] = s[idx]
        return out

    prev_actions = padded(1, 0)
    corrected = padded(2, 0)
    oracle_stop = padded(3, -1)
    # torch.ones_like(corrected [T, B, ...], dtype=float) with step 0 zeroed, then transposed to [B, T, ...]
    not_done = _alloc(tuple(corrected.shape), torch.float32, 1.0, pin)
    not_done[:, 0] = 0
    return (observations, prev_actions.view(-1, 2), not_done.view(-1, 2), corrected.view(-1, 2), oracle_stop.view(-1, 1))


def split_batch_tbptt(batch, prev_actions, not_done_masks, corrected_actions, oracle_stop_batch, tbptt_steps, split_dim):
    """Chunks of `tbptt_steps` along `split_dim` (views, no copies); the instruction is shared by every chunk.
    -> list of (observations, prev_actions, not_done_masks, corrected_actions, oracle_stop) (common/utils.py:120-144)."""
    pieces = {k: (v if k == "instruction" else v.split(tbptt_steps, dim=split_dim)) for k, v in batch.items()}
    out = []
    for i, (pa, ca, nd, os_) in enumerate(zip(prev_actions.split(tbptt_steps, dim=split_dim),
                                               corrected_actions.split(tbptt_steps, dim=split_dim),
                                               not_done_masks.split(tbptt_steps, dim=split_dim),
                                               oracle_stop_batch.split(tbptt_steps, dim=split_dim))):
        obs = {k: (v if k == "instruction" else v[i]) for k, v in pieces.items()}
        out.append((obs, pa, nd, ca, os_))
    return out


def _block_shuffle(lst, block_size):
    blocks = [lst[i: i + block_size] for i in range(0, len(lst), block_size)]
    random.shuffle(blocks)
    return [ele for block in blocks for ele in block]


# -----------------------------------------------------------------------------------------------
# on-disk trajectory store
# -----------------------------------------------------------------------------------------------
class TrajectoryStore:
    """Directory of episodes: `index.json` + one raw little-endian file per (episode, field), read through np.memmap.

    An episode is what the reference packs into one LMDB value (hierarchical_trainer.py:452-470):
    [obs dict of [T, ...] arrays, prev_actions [T,1,2], oracle_actions [T,1,2], stop_step list]."""

    VERSION = 1

    def __init__(self, root: str, mode: str = "r"):
        self.root = root
        self.mode = mode
        self.index_path = os.path.join(root, "index.json")
        if mode == "w":
            os.makedirs(root, exist_ok=True)
            self.index = {"version": self.VERSION, "episodes": []}
        else:
            with open(self.index_path) as fh:
                self.index = json.load(fh)
            if self.index.get("version") != self.VERSION:
                raise ValueError(f"{root}: unsupported trajectory store version {self.index.get('version')}")

    def __len__(self):
        return len(self.index["episodes"])

    def append(self, obs: Dict[str, np.ndarray], prev_actions: np.ndarray, oracle_actions: np.ndarray, stop_step: Sequence,
               rgb_uint8: bool = True, depth_dtype: str = "float32") -> int:
        """Add one episode.  RGB frames holding integral values in 0..255 are stored as uint8 (lossless); depth as
        float32 (default) or float16."""
        if self.mode != "w":
            raise RuntimeError("store opened read-only")
        eid = len(self.index["episodes"])
        fields = {}

        def put(name, arr):
            arr = np.ascontiguousarray(arr)
            fn = f"{eid:08d}.{name}.bin"
            arr.tofile(os.path.join(self.root, fn))
            fields[name] = {"file": fn, "dtype": str(arr.dtype), "shape": list(arr.shape)}

        for k, v in obs.items():
            v = np.asarray(v)
            if k == "rgb" and rgb_uint8 and v.dtype != np.uint8:
                r = np.rint(v)
                if not (np.array_equal(r, v) and r.min() >= 0 and r.max() <= 255):
                    raise ValueError("rgb frames are not integral 0..255 values; pass rgb_uint8=False")
                v = r.astype(np.uint8)
            if k == "depth" and v.dtype != np.dtype(depth_dtype):
                v = v.astype(depth_dtype)
            put("obs." + k, v)
        put("prev_actions", np.asarray(prev_actions))
        put("oracle_actions", np.asarray(oracle_actions))
        put("stop_step", np.asarray(stop_step, dtype=np.int64))
        self.index["episodes"].append({"length": int(np.asarray(prev_actions).shape[0]), "fields": fields})
        return eid

    def close(self):
        if self.mode == "w":
            with open(self.index_path, "w") as fh:
                json.dump(self.index, fh)

    def length_of(self, eid: int) -> int:
        return self.index["episodes"][eid]["length"]

    def read(self, eid: int):
        """-> [obs dict, prev_actions, oracle_actions, stop_step] of read-only memory-mapped arrays."""
        ep = self.index["episodes"][eid]
        arrs = {}
        for name, f in ep["fields"].items():
            shape = tuple(f["shape"])
            path = os.path.join(self.root, f["file"])
            if int(np.prod(shape)) == 0:
                arrs[name] = np.zeros(shape, dtype=f["dtype"])
            else:
                arrs[name] = np.memmap(path, dtype=f["dtype"], mode="r", shape=shape)
        obs = {k[4:]: v for k, v in arrs.items() if k.startswith("obs.")}
        return [obs, arrs["prev_actions"], arrs["oracle_actions"], list(arrs["stop_step"])]


def _is_native_store(path: str) -> bool:
    return os.path.exists(os.path.join(path, "index.json"))


class IWTrajectoryDataset(torch.utils.data.IterableDataset):
    """Same constructor and iteration order as the reference class (length-sorted blocks of `batch_size` inside preloads
    of batch_size*100 episodes, block-shuffled with Python's `random`), reading either a TrajectoryStore directory or
    -- where lmdb + msgpack_numpy are installed -- the reference's LMDB."""

    def __init__(self, lmdb_features_dir, use_iw, inflection_weight_coef=1.0, lmdb_map_size=1e9, batch_size=1, is_bert=False):
        super().__init__()
        self.lmdb_features_dir = lmdb_features_dir
        self.lmdb_map_size = lmdb_map_size
        self.preload_size = batch_size * 100
        self._preload: List = []
        self.batch_size = batch_size
        self.is_bert = is_bert
        self.inflec_weights = torch.tensor([1.0, inflection_weight_coef]) if use_iw else torch.tensor([1.0, 1.0])
        self._store: Optional[TrajectoryStore] = None
        if _is_native_store(lmdb_features_dir):
            self._store = TrajectoryStore(lmdb_features_dir, "r")
            self.length = len(self._store)
        else:
            lmdb = self._lmdb()
            with lmdb.open(lmdb_features_dir, map_size=int(lmdb_map_size), readonly=True, lock=False) as env:
                self.length = env.stat()["entries"]

    @staticmethod
    def _lmdb():
        try:
            import lmdb  # noqa: F401
            import msgpack_numpy  # noqa: F401
        except ImportError as exc:      # loud: there is no silent fallback for a format we cannot read
            raise ImportError("reading the reference's LMDB trajectory buffer needs the `lmdb` and `msgpack_numpy` packages; "
                              "convert it to a TrajectoryStore or install them") from exc
        return lmdb

    def _read_many(self, ids: List[int]):
        if self._store is not None:
            return [self._store.read(i) for i in ids]
        import msgpack_numpy

        lmdb = self._lmdb()
        out = []
        with lmdb.open(self.lmdb_features_dir, map_size=int(self.lmdb_map_size), readonly=True, lock=False) as env, \
                env.begin(buffers=True) as txn:
            for i in ids:
                out.append(msgpack_numpy.unpackb(txn.get(str(i).encode()), raw=False))
        return out

    def _load_next(self):
        if len(self._preload) == 0:
            if len(self.load_ordering) == 0:
                raise StopIteration
            ids = []
            for _ in range(self.preload_size):
                if len(self.load_ordering) == 0:
                    break
                ids.append(self.load_ordering.pop())
            new_preload = self._read_many(ids)
            lengths = [len(ep[0]) for ep in new_preload]       # NB: the reference sorts on len(obs dict), kept as is
            sort_priority = list(range(len(lengths)))
            random.shuffle(sort_priority)
            sorted_ordering = list(range(len(lengths)))
            sorted_ordering.sort(key=lambda k: (lengths[k], sort_priority[k]))
            for idx in _block_shuffle(sorted_ordering, self.batch_size):
                self._preload.append(new_preload[idx])
        return self._preload.pop()

    def __next__(self):
        obs, prev_actions, oracle_actions, stop_step = self._load_next()
        obs = dict(obs)
        discrete = np.array(obs["vln_oracle_action_sensor"])              # copy (the store is read-only)
        val = int(stop_step[-1]) - 1
        discrete[val:] = 4
        obs["vln_oracle_action_sensor"] = discrete
        oracle_stop = np.zeros_like(discrete)
        oracle_stop[val:] = 1
        if self.is_bert:
            obs["instruction"] = np.expand_dims(np.asarray(obs["instruction"][0]), axis=0)
        else:
            obs["instruction"] = np.expand_dims(np.asarray(obs["glove_tokens"][0]), axis=0)
            del obs["glove_tokens"]
        obs = {k: torch.from_numpy(np.asarray(v)) for k, v in obs.items()}
        return (obs, torch.from_numpy(np.asarray(prev_actions)), torch.from_numpy(np.asarray(oracle_actions)),
                torch.from_numpy(oracle_stop))

    def __iter__(self):
        worker_info = torch.utils.data.get_worker_info()
        if worker_info is None:
            start, end = 0, self.length
        else:
            per_worker = int(np.ceil(self.length / worker_info.num_workers))
            start = per_worker * worker_info.id
            end = min(start + per_worker, self.length)
        self._preload = []
        self.load_ordering = list(reversed(_block_shuffle(list(range(start, end)), self.preload_size)))
        return self


# -----------------------------------------------------------------------------------------------
# pinned, asynchronous feeding of the GPU
# -----------------------------------------------------------------------------------------------
class PrefetchLoader:
    """Iterates (observations, prev_actions, not_done_masks, corrected_actions, oracle_stop) batches ON THE DEVICE.

    A worker thread pulls `batch_size` episodes from `dataset`, collates them into pinned host buffers (collate_fn,
    one copy per trajectory) and uploads them on a private CUDA stream; the consumer's stream waits on the upload's event
    only when it takes the batch, so disk reads, collation and H2D copies overlap the training of the previous batch
    (the reference does the collation in DataLoader workers and the upload synchronously, hierarchical_trainer.py:674-677)."""

    def __init__(self, dataset: Iterable, batch_size: int, device: torch.device, depth: int = 2, drop_last: bool = True):
        self.dataset, self.batch_size, self.device, self.depth, self.drop_last = dataset, batch_size, torch.device(device), depth, drop_last
        self.cuda = self.device.type == "cuda"

    def _produce(self, q: "queue.Queue"):
        try:
            stream = torch.cuda.Stream(self.device) if self.cuda else None
            it = iter(self.dataset)
            while True:
                samples = []
                try:
                    for _ in range(self.batch_size):
                        samples.append(next(it))
                except StopIteration:
                    pass
                if not samples or (len(samples) < self.batch_size and self.drop_last):
                    break
                host = collate_fn(samples, pin=self.cuda)
                if not self.cuda:
                    q.put((host, None, host))
                    continue
                with torch.cuda.stream(stream):
                    obs = ObservationsDict({k: v.to(self.device, non_blocking=True) for k, v in host[0].items()})
                    rest = tuple(t.to(self.device, non_blocking=True) for t in host[1:])
                    ev = torch.cuda.Event()
                    ev.record(stream)
                q.put(((obs,) + rest, ev, host))      # `host` keeps the pinned buffers alive until the copy is consumed
            q.put(None)
        except BaseException as exc:  # surface worker failures in the consumer
            q.put(exc)

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        th = threading.Thread(target=self._produce, args=(q,), daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            dev_batch, ev, _host = item
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            yield dev_batch
        th.join()
