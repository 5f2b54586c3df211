#!/bin/bash
# Round 2, last GPU call: GPU suite, smoke and the default bench line on the final tree
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b26_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -2 gpurun_out/r2b26_pytest_all.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b26_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2b26_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r2b26_bench_default.json 2> gpurun_out/r2b26_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b26_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1), d['cpu_baseline']['kind'], d['clocks'])"
