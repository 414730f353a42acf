#!/bin/bash
# N-GPU job: NCCL parity test (2 ranks) + bench at N ranks.  usage: bash scripts/jobs/multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -q -k two_rank > gpurun_out/r2_nccl_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_nccl_test.log)
tail -5 gpurun_out/r2_nccl_test.log
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2_bench_n$N.err)
tail -5 gpurun_out/r2_bench_n$N.err
head -c 3000 gpurun_out/r2_bench_n$N.json
