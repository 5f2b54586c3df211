#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python profiles/tools/e2e_probe3.py > gpurun_out/r2b22_e2e.log 2>&1; grep -E "^==|brdae rows" gpurun_out/r2b22_e2e.log | cut -c1-330 | awk 'NR%1==0' | tail -14
