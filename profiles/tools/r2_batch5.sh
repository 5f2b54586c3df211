#!/bin/bash
# Round 2, GPU call 5: GPU suite after the lu_static_ok fix; ncu --set full of K4 (n = 20, 16,384 ropes, prefetch + vote) and of
# K1 with the pattern-restricted factors (262,144 pendulums); launch list of the default bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b5_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2b5_pytest_all.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_kernel -s 1 -c 1 -o gpurun_out/r2b5_k4_n20 python profiles/tools/k4_variants.py ncu 20:16384 > gpurun_out/r2b5_ncu_k4.log 2>&1; echo "ncu k4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 3 -c 1 -o gpurun_out/r2b5_k1_pendulum python bench.py --systems 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gather > gpurun_out/r2b5_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b5_ncu_launches.log 2>&1; echo "launch list rc=$?"
