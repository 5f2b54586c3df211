#!/bin/bash
# Round 2, GPU call 15: compute-sanitizer over small runs of every kernel family; final captures (ncu) of K4 and K1
mkdir -p gpurun_out
timeout -k 10 300 python profiles/tools/sanitize_small.py > gpurun_out/r2b15_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r2b15_plain.log
for tool in memcheck synccheck racecheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/tools/sanitize_small.py > gpurun_out/r2b15_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok$" gpurun_out/r2b15_$tool.log | tail -4
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k4_kernel -s 1 -c 1 -f -o gpurun_out/r2b15_k4_n20_final python profiles/tools/k4_variants.py ncu 20:16384 > gpurun_out/r2b15_ncu_k4.log 2>&1; echo "ncu k4 rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k4_kernel -s 1 -c 1 -f -o gpurun_out/r2b15_k4_n100_final python profiles/tools/k4_variants.py ncu 100:16384 > gpurun_out/r2b15_ncu_k4_n100.log 2>&1; echo "ncu k4 n100 rc=$?"
