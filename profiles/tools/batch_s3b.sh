#!/bin/bash
# GPU tests (incl. the plugin path) and the n-pendulum bench lines with e2e and the CPU comparator
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s3b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_s3b.log
for w in npendulum_n20_16k npendulum_n50_16k; do
  timeout 240 python bench.py --workload $w --steps 2 --warmup 3 --cpu-sample 128 > gpurun_out/final_$w.log 2> gpurun_out/final_$w.err; echo "$w rc=$?"
done
timeout 200 python bench.py --workload npendulum_n100_16k --systems 4096 --steps 2 --warmup 3 --cpu-sample 64 > gpurun_out/final_npendulum_n100_4k.log 2> gpurun_out/final_npendulum_n100_4k.err; echo "n100 rc=$?"
