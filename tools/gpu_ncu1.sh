#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 ROBOVLN_MULTISTREAM=0
M1=gpu__time_duration.sum
M2=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread
timeout 600 ncu --profile-from-start off --metrics $M1 --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --profile-from-start off --metrics $M2 --clock-control none -k regex:gemm_tc --csv --log-file gpurun_out/gemm_metrics.csv python tools/ncu_step.py > gpurun_out/ncu2.log 2>&1
# full capture of 3 launches of the top kernel (BERT FFN1 GEMM = launches of gemm_tc<128> with the most flops): skip into BERT
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tc -s 110 -c 6 -o gpurun_out/prof_gemm python tools/ncu_step.py > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out
