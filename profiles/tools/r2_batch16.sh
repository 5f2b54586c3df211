#!/bin/bash
# Round 2, GPU call 16: whole GPU suite on the final tree, smoke(), the default bench line
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b16_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2b16_pytest_all.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b16_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2b16_smoke.log
BRDAE_ROWS_TIMING=1 timeout -k 10 900 python bench.py > gpurun_out/r2b16_bench_default.json 2> gpurun_out/r2b16_bench_default.err; echo "bench rc=$?"
tail -2 gpurun_out/r2b16_bench_default.err | cut -c1-400
python -c "
import json; d=json.loads(open('gpurun_out/r2b16_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('d2h_copy_probe_gbs_all_ranks'), d['cpu_baseline']['value'], d['clocks'])"
