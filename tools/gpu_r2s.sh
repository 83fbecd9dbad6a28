#!/bin/bash
# round 2, GPU call S (2 GPUs): the GPU tests that need two devices, the bench under torchrun at N = 2
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_classify.py -m gpu -x -q -k "peer_memory or two_processes_nccl or multi or sharded" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -4 gpurun_out/r2s_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2s_bench_n2.json; tail -3 gpurun_out/r2s_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref_n2.json 2> gpurun_out/r2s_bench_ref_n2.err; echo "reference rc=$?"; head -c 300 gpurun_out/r2s_bench_ref_n2.json
