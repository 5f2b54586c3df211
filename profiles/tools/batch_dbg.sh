#!/bin/bash
for v in "$@"; do
  if [ "$v" != "default" ]; then export BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$v.so; else unset BRDAE_LIB; fi
  echo "== $v"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bit_identical_to_c_oracle" 2>&1 | tail -4
done
