#!/bin/bash
# Round 2, GPU call 1: the new chain kernel K4 on hardware (parity tests, the three chain workloads, a first ncu capture),
# then the whole GPU suite and the default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "npendulum or dense_output or host_buffer" > gpurun_out/r2b1_pytest_chain.log 2>&1; echo "pytest chain rc=$?"; tail -5 gpurun_out/r2b1_pytest_chain.log
for wl in npendulum_n20_16k npendulum_n50_16k npendulum_n100_16k; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b1_$wl.json 2> gpurun_out/r2b1_$wl.err; echo "$wl rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2b1_$wl.json').read()); print('$wl', round(d['value']), 'ropes/s', round(d['ms_per_step'],1), 'ms', d['status_histogram'], d['mean_accepted_steps'])"
done
for reps in 2 4; do
  BRDAE_NEWTON_REPS=$reps timeout 300 python bench.py --workload npendulum_n20_16k --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b1_n20_reps$reps.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2b1_n20_reps$reps.json').read()); print('n20 reps=$reps', round(d['value']), 'ropes/s', round(d['ms_per_step'],1), 'ms')"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k4_kernel -c 1 -o gpurun_out/r2b1_k4_n20 python bench.py --workload npendulum_n20_16k --systems 4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b1_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b1_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 gpurun_out/r2b1_pytest_all.log
timeout 600 python bench.py > gpurun_out/r2b1_bench_default.json 2> gpurun_out/r2b1_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b1_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['cpu_baseline']['value'], d['clocks'])"
