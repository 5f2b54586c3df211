#!/bin/bash
# ncu --set full captures of the Robertson and transistor K1 kernels (262,144 / 65,536 systems), each after the same
# command exited 0 without ncu
mkdir -p gpurun_out
for wc in robertson_4M:262144 transistor_256k:65536; do
  w=${wc%%:*}; b=${wc##*:}
  python bench.py --workload $w --systems $b --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$w.log 2>&1 || { echo "plain $w failed"; tail -5 gpurun_out/plain_$w.log; continue; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_kernel -s 3 -c 1 -f -o gpurun_out/prof_k1_$w \
    python bench.py --workload $w --systems $b --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
ls -la gpurun_out/prof_k1_*
