#!/bin/bash
# Round 2, final single-GPU call: whole GPU suite, the reference arm and the default bench line, every single-GPU workload.
tag=${1:-r2final}
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/${tag}_pytest_all.log
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref arm rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout -k 10 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"
for wl in pendulum3_1M_tend2 robertson_4M transistor_256k pendulum3_1M_tend10 npendulum_n20_16k npendulum_n50_16k npendulum_n100_16k; do
  if [ $wl != pendulum3_1M_tend2 ]; then
    timeout -k 10 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/${tag}_$wl.json 2> gpurun_out/${tag}_$wl.err; echo "$wl rc=$?"
    f=gpurun_out/${tag}_$wl.json
  else f=gpurun_out/${tag}_bench_default.json; fi
  python -c "
import json; d=json.loads(open('$f').read()); r=d['roofline']; print('$wl', round(d['value']), round(d['ms_per_step'],1), 'ms', r['bound'], round(r['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],2), d['cpu_baseline']['kind'], 'C', round(d['cpu_baseline_native'].get('value',0)), d['clocks']['reasons'])"
done
