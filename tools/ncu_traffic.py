"""Aggregate an `ncu --csv` launch list of one policy step (tools/ncu_step.py) into the `roofline.traffic` evidence:
per tensor-core kernel family launches/step, DRAM bytes per launch, time share, time-weighted tensor-pipe utilisation.
    python tools/ncu_traffic.py profiles/r02_ncu_launches.csv > profiles/r02_gemm_traffic.json"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
i0 = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[i0]
recs = {}
for r in rows[i0 + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    k = int(d["ID"])
    rec = recs.setdefault(k, {"name": d["Kernel Name"]})
    try:
        rec[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        pass
fam = lambda n: "gemm_tc_kernel" if "gemm_tc_kernel" in n else ("vla_block" if "vla_" in n else None)
tot_t = sum(r.get("gpu__time_duration.sum", 0.0) for r in recs.values())
out = {"how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active "
              "--clock-control none --cache-control none over one eager policy step (tools/ncu_step.py, cfg2 shapes); serialised single-kernel replays: compare shares",
       "all_kernels_ncu_time_us": tot_t / 1e3, "launches_per_step": len(recs)}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for r in recs.values():
    f = fam(r["name"])
    if f is None:
        continue
    a = agg[f]
    t = r.get("gpu__time_duration.sum", 0.0)
    a["launches"] += 1
    a["time_ns"] += t
    a["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
    a["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
    a["tensor_x_time"] += r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
tc = {"launches": 0, "time_ns": 0.0, "dram": 0.0, "tensor_x_time": 0.0}
for f, a in agg.items():
    out[f] = {"launches_per_step": int(a["launches"]), "ncu_time_us_per_step": a["time_ns"] / 1e3,
              "dram_bytes_read_per_step": a["dram_read"], "dram_bytes_write_per_step": a["dram_write"],
              "dram_bytes_per_launch": (a["dram_read"] + a["dram_write"]) / max(a["launches"], 1),
              "tensor_pipe_pct_time_weighted": a["tensor_x_time"] / max(a["time_ns"], 1.0),
              "kernel_share_of_step": a["time_ns"] / max(tot_t, 1.0)}
    tc["launches"] += a["launches"]; tc["time_ns"] += a["time_ns"]; tc["dram"] += a["dram_read"] + a["dram_write"]; tc["tensor_x_time"] += a["tensor_x_time"]
out["kernel"] = "tcgen05 kernels (gemm_tc_kernel + fused cross-modal block)"
out["tensor_core_launches_per_step"] = int(tc["launches"])
out["dram_bytes_per_launch"] = tc["dram"] / max(tc["launches"], 1)
out["kernel_share_of_step"] = tc["time_ns"] / max(tot_t, 1.0)
out["tensor_pipe_pct_time_weighted"] = tc["tensor_x_time"] / max(tc["time_ns"], 1.0)
print(json.dumps(out, indent=1))
