#!/bin/bash
# gpurun_out/r02_* (scratch) -> profiles/r02_* (tracked): bench lines, ncu summaries, launch list, sanitizer, parity, timelines.
G=gpurun_out/r02; P=profiles/r02
lastjson() { grep '^{' "$1" | tail -1; }
lastjson ${G}_bench.log > ${P}_bench.json
lastjson ${G}_bench_ref.log > ${P}_bench_ref.json
for w in sr48_b16 voc_b16 sr48_b64 chain24 synth sr24_3s tts; do lastjson ${G}_bench_$w.log > ${P}_bench_$w.json; done
cp ${G}_parity.log ${P}_parity.log
cp ${G}_env.txt ${P}_env.txt
cp ${G}_microbench_mha.txt ${P}_microbench_mha.txt
[ -f ${G}_microbench_convT.txt ] && cp ${G}_microbench_convT.txt ${P}_microbench_convT.txt
cp ${G}_timeline_vocoder_b1.txt ${P}_timeline_vocoder_b1.txt
cp ${G}_timeline_vocoder_b1_serial.txt ${P}_timeline_vocoder_b1_serial.txt
cp ${G}_timeline_synth.txt ${P}_timeline_synth.txt
python tools/launch_summary.py ${G}_launches_bench.csv > ${P}_launches_bench_b1x10s.txt
for k in act_sat actmma_sat umma_c32_b16 umma_c128_b16 umma_c256_b16 umma_c16_b1 mha; do
  python tools/ncu_summary.py ${G}_prof_$k.ncu-rep > ${P}_ncu_$k.txt 2>/dev/null
done
{ for t in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $t python tools/sanitize_target.py"; grep -E "sanitize target ok|ERROR SUMMARY|RACECHECK SUMMARY" ${G}_sanitize_$t.log; done; } > ${P}_compute_sanitizer.txt
cp ${G}_ncu_traffic.json profiles/ncu_traffic.json
python tools/sass_summary.py > profiles/sass_summary.txt 2>/dev/null
ls -la profiles | wc -l
