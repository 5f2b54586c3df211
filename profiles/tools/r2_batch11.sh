#!/bin/bash
# Round 2, GPU call 11: the band kernel as a parameterised body (chain regression: K2 analytic + FD) and the first user band problem
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "band or k2 or fd_jacobian_npendulum or fd_npendulum" > gpurun_out/r2b11_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "rod fixture|passed|failed|Error" gpurun_out/r2b11_pytest.log | tail -8
