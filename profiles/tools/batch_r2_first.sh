#!/bin/bash
# First GPU call of round 2: confirm on hardware what round 1 could only check through the CPU harness (the GPU tests marked
# `first_hw_run` in tests/test_gpu_parity.py show up as XPASS when they pass), then the default bench line and the reference arm.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu_r2_first.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r2_first.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_r2_first.log 2> gpurun_out/bench_r2_first.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_first.log').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['cpu_baseline']['value'], d['clocks'])"
