#!/bin/bash
# Runs on the GPU box under gpurun: GPU parity tests, smoke, a short bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 180 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
