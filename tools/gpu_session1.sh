#!/bin/bash
# first GPU session: smoke, parity tests, first bench lines
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
