#!/bin/bash
# final bench lines of round 1 for every single-GPU workload (device-timed value, e2e through the host-buffer
# C ABI entry, CPU comparators); the default line first, then the reference arm on the same box
mkdir -p gpurun_out
python bench.py > gpurun_out/final_default.log 2> gpurun_out/final_default.err; echo "default rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.log 2> gpurun_out/final_reference.err; echo "reference rc=$?"
for wc in pendulum3_1M_tend10:2048 robertson_4M:4096 transistor_256k:1024; do   # workload:systems of the numpy-port sample
  w=${wc%%:*}; c=${wc##*:}
  python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample $c > gpurun_out/final_$w.log 2> gpurun_out/final_$w.err; echo "$w rc=$?"
done
for w in npendulum_n20_16k npendulum_n50_16k; do   # the numpy port needs ~1.8 s per rope: 8 ropes per host core
  python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample 128 > gpurun_out/final_$w.log 2> gpurun_out/final_$w.err; echo "$w rc=$?"
done
python bench.py --workload npendulum_n100_16k --systems 4096 --steps 2 --warmup 3 --cpu-sample 64 > gpurun_out/final_npendulum_n100_4k.log 2> gpurun_out/final_npendulum_n100_4k.err; echo "n100 rc=$?"
for f in gpurun_out/final_*.log; do
  python - "$f" <<'EOF'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        r = d.get('roofline') or {}
        cb = d.get('cpu_baseline') or {}
        cn = d.get('cpu_baseline_native') or {}
        e = d.get('e2e') or {}
        print(sys.argv[1], d['config'].get('workload'), 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2),
              'frac', r.get('frac'), 'e2e', e.get('value'), 'cpu', cb.get('value'), cb.get('cores'), 'cpu_c', cn.get('value'),
              d.get('clocks'))
EOF
done
