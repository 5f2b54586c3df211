#!/bin/bash
# Round 2, GPU call 17: the default bench line twice (FP64 probe moved behind the timed regions)
mkdir -p gpurun_out
for k in 1 2; do
  timeout -k 10 900 python bench.py --no-cpu-baseline > gpurun_out/r2b17_bench_$k.json 2> gpurun_out/r2b17_bench_$k.err; echo "bench $k rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2b17_bench_$k.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['roofline']['peak'],2), 'gather', round(d['value_with_gather']['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
