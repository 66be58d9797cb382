#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 50 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload batch_f32 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_batch.json 2> gpurun_out/bench_2gpu_batch.err; echo "bench 2gpu batch exit $?"
tail -c 600 gpurun_out/bench_2gpu_batch.json; tail -3 gpurun_out/bench_2gpu_batch.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 10 --warmup 2 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "ref 2gpu exit $?"; cut -c1-200 gpurun_out/bench_2gpu_ref.json
