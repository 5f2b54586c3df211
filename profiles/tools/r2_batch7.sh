#!/bin/bash
# Round 2, GPU call 7: staged K4 with the new default geometry: small-batch geometry check, ncu --set full (n = 20, 16,384 ropes),
# the three chain workloads through bench.py (CPU baselines included), whole GPU suite.
mkdir -p gpurun_out
timeout -k 10 300 python profiles/tools/k4_variants.py staged 20:4096,20:8192,20:2048 "BRDAE_K4_RPW=8;BRDAE_K4_RPW=16;BRDAE_K4_RPW=32;BRDAE_K4_VOTE=1" > gpurun_out/r2b7_k4_small.jsonl 2>gpurun_out/r2b7_k4_small.err; cat gpurun_out/r2b7_k4_small.jsonl
timeout -k 10 300 python profiles/tools/k4_variants.py staged 20:65536 "BRDAE_K4_VOTE=1" > gpurun_out/r2b7_k4_64k.jsonl 2>gpurun_out/r2b7_k4_64k.err; cat gpurun_out/r2b7_k4_64k.jsonl
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k4_kernel -s 1 -c 1 -f -o gpurun_out/r2b7_k4_n20_staged python profiles/tools/k4_variants.py ncu 20:16384 > gpurun_out/r2b7_ncu_k4.log 2>&1; echo "ncu k4 rc=$?"
for wl in npendulum_n20_16k npendulum_n50_16k npendulum_n100_16k; do
  timeout -k 10 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2b7_$wl.json 2> gpurun_out/r2b7_$wl.err; echo "$wl rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2b7_$wl.json').read()); print('$wl', round(d['value']), round(d['ms_per_step'],1), 'ms frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'C', d['cpu_baseline_native'].get('value'), d['config']['launch'])"
done
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b7_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2b7_pytest_all.log
