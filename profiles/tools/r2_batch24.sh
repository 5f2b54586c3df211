#!/bin/bash
# Round 2, GPU call 24: the host-buffer entries under CUDA_LAUNCH_BLOCKING=1 (the launch blocks: the pipelined copy-in must not be used)
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "host or rows or integration or invalid_y0" > gpurun_out/r2b24_launch_blocking.log 2>&1; echo "launch-blocking rc=$?"; tail -2 gpurun_out/r2b24_launch_blocking.log
