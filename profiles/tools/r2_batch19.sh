#!/bin/bash
# Round 2, GPU call 19: final tree -- GPU suite, smoke, default bench, Robertson and chain lines
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b19_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/r2b19_pytest_all.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b19_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2b19_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r2b19_bench_default.json 2> gpurun_out/r2b19_bench_default.err; echo "bench rc=$?"
for wl in robertson_4M npendulum_n20_16k; do
  timeout -k 10 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2b19_$wl.json 2> gpurun_out/r2b19_$wl.err; echo "$wl rc=$?"
done
for f in bench_default robertson_4M npendulum_n20_16k; do python -c "
import json; d=json.loads(open('gpurun_out/r2b19_$f.json').read()); r=d['roofline']; print('$f', round(d['value']), round(d['ms_per_step'],2), r['bound'], round(r['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1), d['clocks']['reasons'])"; done
