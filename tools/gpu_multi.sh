#!/bin/bash
# run with: gpurun --gpus 2 -- 'bash tools/gpu_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tools/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1
timeout 300 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_samebox.log 2>&1
timeout 300 $TR --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_${N}gpu_b16.log 2>&1
timeout 300 $TR --master-port 29536 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu.log 2>&1
grep -h MULTI_GPU_OK gpurun_out/multi_check_$N.log; tail -c 400 gpurun_out/bench_${N}gpu.log
