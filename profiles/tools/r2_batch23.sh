#!/bin/bash
# Round 2, GPU call 23: ncu --set full of the Robertson (two 192-thread CTAs per SM) and transistor kernels
mkdir -p gpurun_out
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 3 -c 1 -f -o gpurun_out/r2b23_k1_robertson python bench.py --workload robertson_4M --systems 1048576 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gather > gpurun_out/r2b23_ncu_rob.log 2>&1; echo "ncu rob rc=$?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 3 -c 1 -f -o gpurun_out/r2b23_k1_transistor python bench.py --workload transistor_256k --systems 65536 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-gather > gpurun_out/r2b23_ncu_tra.log 2>&1; echo "ncu tra rc=$?"
