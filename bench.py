#!/usr/bin/env python
"""Benchmark of the HCM policy step (hi -> argmax -> lo) -- BASELINE.json metric
"policy-forward obs/sec (batch=64, seq=80, 256x256 RGB-D)".

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU

One "step" = one policy forward over one batch of synthetic observations: 64 environments per
GPU (rollout-shaped: N=64, T=1), 256x256 RGB + 256x256 depth + an 80-token instruction per
environment (64 distinct instructions), through hi AND lo.  With N GPUs every rank runs its own
64 environments (weak scaling; N=8 is BASELINE.json configs[3], 512 environments) and the only
collective is the all-gather of the packed [B,7] outputs, inside the timed region.

Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_OBS = 25.29          # algorithmic, SURVEY.md 8(d) / BASELINE.md section 3 (L=80, trunks shared, BERT per row)
METRIC = "policy_forward_obs_per_sec"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "bf16_tflops_burst": float(d["bf16_tflops"]), "hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
        except Exception:
            pass
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_obs_per_sec(sample_rows: int, L: int, steps: int, warmup: int):
    """The reference algorithm (oracle port, fp32, torch CPU) on a bounded sample of the workload:
    `sample_rows` observations of the same shapes, hi + lo (each running its own trunks, as the
    reference executes them)."""
    import torch

    from oracle import hcm_oracle as O
    from oracle import weights as W

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd_hi, sd_lo = W.make_state_dict("hi", 0), W.make_state_dict("lo", 0)
    inp = W.make_inputs(B=sample_rows, L=L, N=sample_rows, rgb_hw=256, seed=1, mask_zero_rows=(0,))

    def step():
        with torch.no_grad():
            logits, _ = O.hi_forward(sd_hi, inp["rgb"], inp["depth"], inp["instruction"], inp["hidden_hi"], inp["masks"])
            O.lo_forward(sd_lo, inp["rgb"], inp["depth"], inp["hidden_lo"], inp["masks"], logits.argmax(1))

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return sample_rows * steps / total, total / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_sample_rows
    v, ms, cores = cpu_oracle_obs_per_sec(rows, args.seq_len, args.steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "obs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: HCM policy forward (hi+lo), 256x256 RGB-D, L=%d, rollout-shaped" % args.seq_len,
                   "per_gpu_batch": args.batch, "seq_len": args.seq_len,
                   "note": "CPU arm: each step is a %d-observation sample of the batch" % rows},
        "cpu_baseline": {"value": v, "unit": "obs/s", "cores": cores, "kind": "port",
                         "sample": "%d observations/step x %d steps, oracle/hcm_oracle.py (fp32 torch CPU restatement "
                                   "of the reference modules; the Python reference itself cannot travel to the box)" % (rows, args.steps)},
        "e2e": {"value": v, "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    import robovln_b200 as R
    from robovln_b200 import sharding

    B, L = args.batch, args.seq_len
    N = B                                   # rollout-shaped: one step of B environments
    policy = R.HcmPolicy().share_frozen_trunks().to(dev).eval()
    rt = policy._runtime()

    g = torch.Generator(device="cpu")
    g.manual_seed(1 + rank)
    n_sets = 3                              # rotating input sets: 3 x 67 MB > 126 MB L2

    def make_host_set():
        ids = torch.randint(1000, 30522, (B, L), generator=g).float()
        ids[:, 0] = 101
        ids[:, -1] = 102
        masks = torch.ones((B, 2))
        masks[::7] = 0.0
        return {
            "rgb": torch.randint(0, 256, (B, 256, 256, 3), generator=g).float().pin_memory(),
            "depth": torch.rand((B, 256, 256, 1), generator=g).pin_memory(),
            "instruction": ids.pin_memory(), "masks": masks.pin_memory(),
            "hidden_hi": (torch.randn((2, N, 512), generator=g) * 0.1).pin_memory(),
            "hidden_lo": (torch.randn((2, N, 512), generator=g) * 0.1).pin_memory(),
        }

    host_sets = [make_host_set() for _ in range(n_sets)]
    dev_sets = [{k: v.to(dev) for k, v in hs.items()} for hs in host_sets]
    global_rows = B * world

    def step_device(i):
        d = dev_sets[i % n_sets]
        obs = {"rgb": d["rgb"], "depth": d["depth"], "instruction": d["instruction"]}
        logits, act, stop, hh, hl, sub = policy.act(obs, d["hidden_hi"], d["hidden_lo"], d["masks"])
        packed = sharding.pack_outputs(logits, act, stop)
        if world > 1:
            packed = sharding.all_gather_outputs(packed, global_rows)
        return packed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------
    for i in range(max(args.warmup, 3)):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step_device(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = rt.launches() * args.steps + (args.steps if world > 1 else 0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = global_rows / (ms_step * 1e-3)
    finite = bool(torch.isfinite(out).all().item())

    # ---- end to end through the host-buffer API (H2D + forward + D2H per step) --------------
    host_out = None
    for i in range(3):
        host_out = policy.act_host(**{k: host_sets[i % n_sets][k] for k in ("rgb", "depth", "instruction", "masks", "hidden_hi", "hidden_lo")}, out=host_out)
    barrier()
    e0.record()
    for i in range(args.steps):
        hs = host_sets[i % n_sets]
        host_out = policy.act_host(hs["rgb"], hs["depth"], hs["instruction"], hs["masks"], hs["hidden_hi"], hs["hidden_lo"], out=host_out)
        _ = float(host_out["logits"][0, 0])          # the step's result is read on the host
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = global_rows / (ms_e2e / args.steps * 1e-3)
    # same call with the RGB frames as uint8 (observation ingest): a quarter of the RGB bytes
    e2e_u8 = None
    try:
        u8_sets = [hs["rgb"].to(torch.uint8).pin_memory() for hs in host_sets]
        for i in range(3):
            hs = host_sets[i % n_sets]
            host_out = policy.act_host(u8_sets[i % n_sets], hs["depth"], hs["instruction"], hs["masks"], hs["hidden_hi"], hs["hidden_lo"], out=host_out)
        barrier()
        e0.record()
        for i in range(args.steps):
            hs = host_sets[i % n_sets]
            host_out = policy.act_host(u8_sets[i % n_sets], hs["depth"], hs["instruction"], hs["masks"], hs["hidden_hi"], hs["hidden_lo"], out=host_out)
            _ = float(host_out["logits"][0, 0])
        e1.record()
        barrier()
        ms_u8 = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_u8], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_u8 = float(t.item())
        e2e_u8 = {"value": global_rows / (ms_u8 / args.steps * 1e-3), "unit": "obs/s", "ms_per_step": ms_u8 / args.steps,
                  "h2d_bytes_per_step": B * (256 * 256 * 3 + 256 * 256 * 4 + L * 4 + 2 * 4) + 2 * (2 * N * 512 * 4)}
    except Exception as exc:
        e2e_u8 = {"error": str(exc)[:200]}
    h2d = B * (256 * 256 * 3 * 4 + 256 * 256 * 4 + L * 4 + 2 * 4) + 2 * (2 * N * 512 * 4)
    d2h = B * (4 + 2 + 1) * 4 + 2 * (2 * N * 512 * 4)

    # ---- per-launch profile of one step -> roofline of the dominant kernel ------------------
    peaks = _peaks()
    d = dev_sets[0]
    rt.profile_policy(d["rgb"], d["depth"], d["instruction"], d["masks"], d["hidden_hi"], d["hidden_lo"])      # warm
    ops = rt.profile_policy(d["rgb"], d["depth"], d["instruction"], d["masks"], d["hidden_hi"], d["hidden_lo"])
    gemm_ms = sum(o["ms"] for o in ops if o["flops"] > 0)
    gemm_fl = sum(o["flops"] for o in ops if o["flops"] > 0)
    n_gemm = sum(1 for o in ops if o["flops"] > 0)
    all_ms = sum(o["ms"] for o in ops)
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tp):          # dram__bytes_read.sum + dram__bytes_write.sum per gemm_tc launch, from the committed ncu pass
        try:
            traffic = float(json.load(open(tp))["dram_bytes_per_launch"])
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM), %d launches/step" % n_gemm,
        "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
        "peak_source": "%s, sustained bf16 (kernel timed inside the step)" % peaks["source"], "traffic": traffic,
        "traffic_note": "average DRAM bytes per gemm_tc launch (ncu, profiles/r01_gemm_traffic.json); 'achieved' is FLOP/s "
                        "(tensor-bound kernel), average FLOPs per launch = flops_per_step / launches",
        "flops_per_step": gemm_fl, "ms_per_step_in_kernel": gemm_ms, "kernel_share_of_step": gemm_ms / all_ms if all_ms else None,
        "step_frac_of_tensor_peak": value / world * GFLOP_PER_OBS * 1e9 / (peaks["bf16_tflops"] * 1e12),
        "single_stream_step_ms": all_ms,
    }
    if args.profile_out and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({"ops": ops, "roofline": roofline, "ms_per_step": ms_step}, open(args.profile_out, "w"), indent=1)

    # ---- secondary shape: one 64-step trajectory, N=1, one shared instruction ----------
    traj = None
    if rank == 0 and world == 1:
        try:
            d0 = dev_sets[0]
            tobs = {"rgb": d0["rgb"], "depth": d0["depth"], "instruction": d0["instruction"][:1].contiguous()}
            th = torch.zeros((2, 1, 512), device=dev)
            for _ in range(3):
                policy.act(dict(tobs), th, th.clone(), d0["masks"])
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                policy.act(dict(tobs), th, th.clone(), d0["masks"])
            e1.record()
            torch.cuda.synchronize()
            tms = e0.elapsed_time(e1) / 10
            traj = {"workload": "trajectory-shaped: T=%d steps of ONE environment (N=1), one shared instruction (BERT runs once), "
                                "LSTMs serial in T" % B, "ms_per_step": tms, "value": B / (tms * 1e-3), "unit": "obs/s"}
        except Exception as exc:      # never lose the contract line over the secondary number
            traj = {"error": str(exc)[:200]}

    # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        v, ms, cores = cpu_oracle_obs_per_sec(args.cpu_sample_rows, L, 3, 1)
        cpu = {"value": v, "unit": "obs/s", "cores": cores, "kind": "port",
               "sample": "%d observations/step x 3 steps (+1 warm-up) of the same shapes through oracle/hcm_oracle.py "
                         "(fp32 torch CPU restatement pinned to the reference's outputs)" % args.cpu_sample_rows}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "obs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": rt.dtype_name, "data": "synthetic",
            "config": {
                "workload": "cfg2: HCM policy forward (hi -> argmax -> lo), batch=64/GPU rollout-shaped (N=64,T=1), "
                            "256x256 RGB + 256x256 depth, 64 distinct 80-token instructions, random-init weights",
                "per_gpu_batch": B, "global_batch": global_rows, "seq_len": L, "parallelism": "dp%d" % world,
                "l2": "3 rotating input sets (201 MB) and a >2 GB per-step working set vs 126 MB L2",
                "outputs_finite": finite, "trajectory_shaped": traj,
            },
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "obs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "ms_per_step": ms_e2e / args.steps, "uint8_rgb_frames": e2e_u8},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cross_modal(args):
    """BASELINE.json configs[2]: cross-modal attention block only, rollout-shaped batch of 128 environments."""
    import torch

    import robovln_b200 as R

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, L = 128, args.seq_len
    policy = R.HcmPolicy().share_frozen_trunks().to(dev).eval()
    rt = policy._runtime()
    g = torch.Generator().manual_seed(3)
    sets = [(torch.randn((B, L, 768), generator=g).to(dev, rt.h16), torch.randn((B, 16, 256), generator=g).to(dev, rt.h16),
             torch.randn((B, 16, 256), generator=g).to(dev, rt.h16)) for _ in range(3)]
    for i in range(max(args.warmup, 3)):
        out = rt.cross_modal(*sets[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = rt.cross_modal(*sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = B * 0.246e9 * 1.0      # 0.288 GFLOP/obs, 0.246 with the query side shared by both modalities
    peaks = _peaks()
    print(json.dumps({
        "metric": "cross_modal_obs_per_sec", "value": B / (ms * 1e-3), "unit": "obs/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": rt.dtype_name, "data": "synthetic",
        "config": {"workload": "cfg3: Visual_Ling_Attn x2 + token mean-pool only, 128 environments, L=%d, 16 visual cells" % L,
                   "launches": int(rt.launches())},
        "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
                     "note": "31.5 GFLOP per call in %d launches: launch/latency bound (27 us at peak)" % int(rt.launches())},
        "outputs_finite": bool(torch.isfinite(out.float()).all().item()),
    }), flush=True)


def run_train(args):
    """BASELINE.json configs[4]: DAgger inner loop on a 64-step trajectory (N=1, T=64, one shared instruction):
    hi forward + CE + backward, lo forward + MSE + BCE-with-logits + backward (hierarchical_trainer.py:498-553),
    frozen encoders on the engine, trainable tail under autograd, dropout as the reference (p=0.25)."""
    import torch
    import torch.nn.functional as F

    import robovln_b200 as R

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    T, L = args.batch, args.seq_len
    policy = R.HcmPolicy().share_frozen_trunks().to(dev)
    hi, lo = policy.high_level, policy.low_level
    hi.train()
    lo.train()
    g = torch.Generator().manual_seed(5)
    rgb = torch.randint(0, 256, (T, 256, 256, 3), generator=g).float().to(dev)
    depth = torch.rand((T, 256, 256, 1), generator=g).to(dev)
    ids = torch.randint(1000, 30522, (1, L), generator=g).float().to(dev)
    masks = torch.ones((T, 2), device=dev)
    masks[0] = 0.0
    tgt_hi = torch.randint(0, 4, (T,), generator=g).to(dev)
    tgt_act = torch.rand((T, 2), generator=g).to(dev)
    tgt_stop = (torch.rand((T, 1), generator=g) > 0.9).float().to(dev)
    sub = torch.randint(0, 5, (T,), generator=g).to(dev)
    params = [p for m in (hi, lo) for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        obs = {"rgb": rgb, "depth": depth, "instruction": ids}
        logits, _ = hi((obs, torch.zeros((2, 1, 512), device=dev), None, masks))
        F.cross_entropy(logits, tgt_hi, ignore_index=-1).backward()
        act, stop, _ = lo((obs, torch.zeros((2, 1, 512), device=dev), None, masks, sub))
        (F.mse_loss(act, tgt_act) + F.binary_cross_entropy_with_logits(stop, tgt_stop)).backward()
        opt.step()       # the engine runs only the FROZEN encoders in train mode: no re-pack of its weights per step
        return logits

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    obs_s = T / (ms * 1e-3)
    print(json.dumps({
        "metric": "dagger_step_tokens_and_pixels_per_sec", "value": obs_s * (L + 2 * 256 * 256), "unit": "tokens+pixels/s",
        "obs_per_sec": obs_s, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "dtype": hi.runtime().dtype_name, "data": "synthetic",
        "config": {"workload": "cfg5: DAgger inner loop, trajectory T=%d (N=1), shared %d-token instruction, hi fwd+CE+bwd, "
                               "lo fwd+MSE+BCE+bwd, AdamW step; frozen encoders on the engine, trainable tail in torch autograd" % (T, L)},
        "outputs_finite": bool(torch.isfinite(out).all().item()),
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)      # ~0.8 s per timed region: enough nvidia-smi clock samples
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="environments per GPU")
    ap.add_argument("--seq-len", type=int, default=80)
    ap.add_argument("--cpu-sample-rows", type=int, default=4)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    ap.add_argument("--workload", default="policy", choices=["policy", "cross_modal", "train"],
                    help="policy = BASELINE.json metric (default, the contract line); cross_modal = configs[2]; train = configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cross_modal":
        run_cross_modal(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
