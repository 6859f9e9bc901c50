#!/bin/bash
# ncu launch list + full capture of the dominant GEMM kernel (run on the GPU box via gpurun; 1 GPU)
mkdir -p gpurun_out
B="python bench.py --chains 65536 --mcmc_per_flow_steps 2 --steps 1 --warmup 3 --warmup_unit iteration --no_e2e --no_cpu_baseline"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
# persistent CTA-pair kernel: a forward layer, a backward-data layer (bf16 cross terms) inside one FM update
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2p_kernel -s 12 -c 4 -o gpurun_out/prof_gemm_tc2p -f $B > gpurun_out/ncu_full.log 2>&1
echo "full capture (persistent) rc=$?"
if [ "$1" == "all" ]; then
  # one-tile CTA-pair kernel: the split-K weight-gradient GEMMs
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 2 -c 2 -o gpurun_out/prof_gemm_tc2 -f $B > gpurun_out/ncu_full2.log 2>&1
  echo "full capture (one-tile) rc=$?"
fi
ls -la gpurun_out/*.ncu-rep
