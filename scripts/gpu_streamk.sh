#!/bin/bash
# stream-K validation on the GPU box: numerics first (hard timeouts: a flag bug would spin), then timings
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "streamk" > gpurun_out/pytest_streamk.log 2>&1
echo "streamk tests rc=$?"; tail -5 gpurun_out/pytest_streamk.log
timeout -s KILL 300 python scripts/streamk_bench.py > gpurun_out/streamk_bench.txt 2>&1
echo "bench rc=$?"; cat gpurun_out/streamk_bench.txt
for sk in 0 1; do
  MFM_STREAMK=$sk timeout -s KILL 300 python bench.py --chains 8192 --no_e2e --no_cpu_baseline --warmup 3 > gpurun_out/bench_8k_sk$sk.json 2> gpurun_out/bench_8k_sk$sk.err
  echo "bench 8k streamk=$sk rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_8k_sk$sk.json')); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'])"
done
