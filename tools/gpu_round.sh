#!/bin/bash
# One GPU-box session: tests, bench, launch list, one full ncu capture.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/env.txt 2>&1
timeout 600 python tools/umma_diag.py > gpurun_out/umma_diag.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not umma" > gpurun_out/t_kernels.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "umma" > gpurun_out/t_umma.log 2>&1
timeout 900 python -m pytest tests/test_gpu_modules.py -q -m gpu -s > gpurun_out/t_modules.log 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48.log 2>&1
tail -3 gpurun_out/*.log
