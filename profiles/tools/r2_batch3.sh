#!/bin/bash
# Round 2, GPU call 3: whole GPU suite with the new parity tests, the reference's self-consistency band on this box,
# default bench line + reference arm, every single-GPU workload.
mkdir -p gpurun_out
nproc
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2b3_pytest_all.log 2>&1; echo "pytest all rc=$?"; grep -E "transistor 512|passed|failed" gpurun_out/r2b3_pytest_all.log | tail -5
timeout 900 python profiles/tools/ref_selfconsistency.py --systems 512 --with-c-oracle --out gpurun_out/r2b3_ref_selfconsistency.json > gpurun_out/r2b3_selfc.log 2>&1; echo "selfc rc=$?"; tail -4 gpurun_out/r2b3_selfc.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b3_bench_reference.json 2> gpurun_out/r2b3_bench_reference.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/r2b3_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2b3_bench_default.json 2> gpurun_out/r2b3_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b3_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['cpu_baseline'], d['cpu_baseline_native'], d['value_with_gather'], d['clocks'])"
for wl in robertson_4M transistor_256k npendulum_n20_16k npendulum_n50_16k npendulum_n100_16k pendulum3_1M_tend10; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/r2b3_$wl.json 2> gpurun_out/r2b3_$wl.err; echo "$wl rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2b3_$wl.json').read()); print('$wl', round(d['value']), round(d['ms_per_step'],1), 'ms frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'C', d['cpu_baseline_native'].get('value'))"
done
