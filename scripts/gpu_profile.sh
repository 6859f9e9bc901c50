#!/bin/bash
# ncu launch list + full capture of the dominant GEMM kernel (run on the GPU box via gpurun)
mkdir -p gpurun_out
B="python bench.py --chains 65536 --mcmc_per_flow_steps 2 --steps 1 --warmup 3 --no_e2e --no_cpu_baseline"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/prof_gemm_tc -f $B > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
