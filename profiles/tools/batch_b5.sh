#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b5.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_b5.log
bash profiles/tools/run_variants_wl.sh transistor_256k --steps 3 --warmup 3 -- default tra128x2
bash profiles/tools/run_variants_wl.sh pendulum3_1M_tend2 --steps 3 --warmup 3 -- default
