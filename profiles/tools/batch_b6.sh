#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host" > gpurun_out/pytest_gpu_b6.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_b6.log
python bench.py --no-cpu-baseline > gpurun_out/bench8.log 2> gpurun_out/bench8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench8.log').read().strip().splitlines()[-1]); print('value', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'e2e', round(d['e2e']['value']))"
