#!/bin/bash
# Round 2, GPU call 2: K4 geometry sweep (ropes per warp x resident warps per SM) on the chain workloads
mkdir -p gpurun_out
run() {  # workload minb rpw
  BRDAE_K4_MINB=$2 BRDAE_K4_RPW=$3 timeout 300 python bench.py --workload $1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b2_$1_m$2_r$3.json 2>gpurun_out/r2b2_err.log
  python -c "
import json; d=json.loads(open('gpurun_out/r2b2_$1_m$2_r$3.json').read()); print('$1 minb=$2 rpw=$3', round(d['value']), 'ropes/s', round(d['ms_per_step'],1), 'ms', d['config']['launch'])"
}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "npendulum" > gpurun_out/r2b2_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -2 gpurun_out/r2b2_pytest_chain.log
for cfg in "8 32" "8 16" "8 8" "16 16" "16 8" "16 4" "32 8" "32 4" "32 2"; do run npendulum_n20_16k $cfg; done
for cfg in "8 16" "16 8" "32 4"; do run npendulum_n100_16k $cfg; done
for cfg in "16 8" "32 4"; do run npendulum_n50_16k $cfg; done
