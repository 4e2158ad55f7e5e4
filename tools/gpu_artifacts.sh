#!/bin/bash
# Round artifacts (1 GPU): tests, default bench (+CPU baseline), side workloads, ncu launch list of the bench command,
# DRAM traffic of the hot kernels, full captures of the hot kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
NCU="ncu --clock-control none"
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/traffic_vocoder.csv python tools/profile_kernels.py forward > gpurun_out/p_forward.log 2>&1
python tools/ncu_traffic.py hierspeechpp_vocoder_sn+dec_B1x10s gpurun_out/traffic_vocoder.csv 3 > gpurun_out/traffic.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/traffic_sr48.csv python tools/profile_kernels.py sr48 16 > gpurun_out/p_sr48.log 2>&1
python tools/ncu_traffic.py speechsr48_B16x10s gpurun_out/traffic_sr48.csv 2 >> gpurun_out/traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --workload speechsr48 --batch 16 --no-cpu-baseline > gpurun_out/bench_sr48.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 16 --no-cpu-baseline > gpurun_out/bench_voc_b16.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --workload speechsr48 --batch 64 --no-cpu-baseline > gpurun_out/bench_sr48_b64.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --batch 32 --seconds 30 --no-cpu-baseline > gpurun_out/bench_voc_b32x30s.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench.log 2>&1
$NCU --set full --import-source on -k regex:act1d -s 2 -c 1 -o gpurun_out/prof3_act_sat -f python tools/profile_kernels.py act 16 32 480000 1 > gpurun_out/p3_act.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof3_umma_c32_b16 -f python tools/profile_kernels.py umma 16 32 480000 7 3 > gpurun_out/p3_umma1.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof3_umma_c128_b16 -f python tools/profile_kernels.py umma 16 128 10000 11 5 > gpurun_out/p3_umma2.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof3_umma_c256_b16 -f python tools/profile_kernels.py umma 16 256 2000 11 5 > gpurun_out/p3_umma4.log 2>&1
$NCU --set full --import-source on -k regex:conv_umma -s 2 -c 1 -o gpurun_out/prof3_umma_c128_b1 -f python tools/profile_kernels.py umma 1 128 10000 11 5 > gpurun_out/p3_umma3.log 2>&1
tail -c 600 gpurun_out/bench.log; echo; tail -c 300 gpurun_out/bench_ref.log; cat gpurun_out/traffic.log | head -40
