#!/bin/bash
# Round 2, GPU call 4: whole GPU suite on the current tree (K1 pattern-sparse factors, report buffers, K4 prefetch + vote),
# default bench with the phase times of the host entry, K4 variants (prefetch target / distance, Newton vote, geometry).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2b4_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/r2b4_pytest_all.log
BRDAE_ROWS_TIMING=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b4_bench_default.json 2> gpurun_out/r2b4_bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b4_bench_default.err | cut -c1-600
python -c "
import json; d=json.loads(open('gpurun_out/r2b4_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['clocks'])"
S="BRDAE_K4_VOTE=0;BRDAE_K4_VOTE=1;BRDAE_K4_VOTE=4;BRDAE_K4_VOTE=8;BRDAE_K4_VOTE=1 BRDAE_NEWTON_REPS=1;BRDAE_K4_VOTE=1 BRDAE_K4_RPW=32;BRDAE_K4_VOTE=1 BRDAE_K4_RPW=8"
timeout 600 python profiles/tools/k4_variants.py default 20:16384 "$S" > gpurun_out/r2b4_k4_default.jsonl 2>gpurun_out/r2b4_k4_default.err; cat gpurun_out/r2b4_k4_default.jsonl
for v in pf0 pf2 pfd1 pfd3; do
  BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_$v.so timeout 300 python profiles/tools/k4_variants.py $v 20:16384 "BRDAE_K4_VOTE=0;BRDAE_K4_VOTE=1" > gpurun_out/r2b4_k4_$v.jsonl 2>gpurun_out/r2b4_k4_$v.err; cat gpurun_out/r2b4_k4_$v.jsonl
done
timeout 600 python profiles/tools/k4_variants.py default 50:16384,100:16384 "BRDAE_K4_VOTE=0;BRDAE_K4_VOTE=1" > gpurun_out/r2b4_k4_big.jsonl 2>gpurun_out/r2b4_k4_big.err; cat gpurun_out/r2b4_k4_big.jsonl
BRDAE_LIB=$PWD/dae_scipy_b200/libbrdae_pf0.so timeout 600 python profiles/tools/k4_variants.py pf0 50:16384,100:16384 "BRDAE_K4_VOTE=1" > gpurun_out/r2b4_k4_big_pf0.jsonl 2>gpurun_out/r2b4_k4_big_pf0.err; cat gpurun_out/r2b4_k4_big_pf0.jsonl
