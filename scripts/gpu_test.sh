#!/bin/bash
# usage: scripts/gpu_test.sh [pytest args]   (run on the GPU box via gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -40
