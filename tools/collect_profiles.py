"""Copy the evidence a `tools/gpu_r2_final.sh` / `gpu_r2_27.sh` / `gpu_r2_scale.sh` run left under gpurun_out/ into profiles/
(the tracked, judged set).  JSON lines are extracted from the logs; ncu CSVs and sanitizer logs are copied as they are."""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, S, P = (os.path.join(ROOT, "gpurun_out", "final"), os.path.join(ROOT, "gpurun_out", "scale"), os.path.join(ROOT, "profiles"))


def last_json(path):
    line = [l for l in open(path) if l.startswith("{")][-1]
    return json.loads(line)


def put(name, obj):
    json.dump(obj, open(os.path.join(P, name), "w"), indent=1)
    print("wrote", name)


jobs = [("bench_default.log", "r02_bench_line.json"), ("bench.log", "r02_bench_line_20steps.json"), ("reference.log", "r02_reference_arm.json"),
        ("bench_bf16.log", "r02_bench_bf16.json"), ("cfg3.log", "r02_cfg3_cross_modal.json"), ("cfg3_single.log", "r02_cfg3_cross_modal_single_cta.json"),
        ("cfg5.log", "r02_cfg5_train_step.json"), ("icache.log", "r02_instruction_cache.json")]
for src, dst in jobs:
    p = os.path.join(F, src)
    if os.path.exists(p):
        put(dst, last_json(p))
for n, dst in (("bench8.log", "r02_bench_8gpu_strong.json"), ("bench4.log", "r02_bench_4gpu_strong.json"), ("bench2.log", "r02_bench_2gpu_strong.json"),
               ("bench2_weak.log", "r02_bench_2gpu_weak.json")):
    p = os.path.join(S, n)
    if os.path.exists(p):
        put(dst, last_json(p))
for src, dst in (("r02_bench_ops.json", "r02_bench_ops.json"), ("clocks.csv", "r02_bench_clocks.csv"), ("r02_cfg3_ncu.csv", "r02_cfg3_ncu.csv"),
                 ("r02_ncu_launches.csv", "r02_ncu_launches.csv"), ("probe.log", "r02_gemm_probe.txt"), ("sanit_mem.log", "r02_sanitizer_memcheck_kernels.log"),
                 ("sanit_mem_policy.log", "r02_sanitizer_memcheck_policy.log")):
    p = os.path.join(F, src)
    if os.path.exists(p):
        shutil.copyfile(p, os.path.join(P, dst))
        print("copied", dst)
tr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(P, "r02_ncu_launches.csv")], capture_output=True, text=True)
if tr.returncode == 0 and tr.stdout.strip().startswith("{"):
    open(os.path.join(P, "r02_gemm_traffic.json"), "w").write(tr.stdout)
    print("wrote r02_gemm_traffic.json")
lb = os.path.join(ROOT, "gpurun_out", "library_bar_tuned.json")
if os.path.exists(lb):
    shutil.copyfile(lb, os.path.join(P, "r02_library_bar.json"))
    print("copied r02_library_bar.json")
