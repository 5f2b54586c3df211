#!/bin/bash
# Round 2, multi-GPU call: N = $1 GPUs of one box.  The two-rank NCCL test, weak scaling of the default workload (with the timed
# gather and the end-to-end leg), strong scaling of the 4M-system Robertson ensemble (BASELINE configs[2]) at every N <= $1.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2mg_topo_n$N.txt 2>&1
run() {  # tag nproc args...
  tag=$1; np=$2; shift 2
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/r2mg_$tag.json 2> gpurun_out/r2mg_$tag.err; echo "$tag rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r2mg_$tag.json').read().strip().splitlines()[-1]); print('$tag', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value']), 'gather', d['value_with_gather'] and round(d['value_with_gather']['value']), d['config'].get('host_placement'), d['status_histogram'])"
}
timeout -k 10 300 python -m pytest tests -m gpu -q -x -k "nccl" > gpurun_out/r2mg_pytest_nccl_n$N.log 2>&1; echo "pytest nccl rc=$?"; tail -2 gpurun_out/r2mg_pytest_nccl_n$N.log
for np in 2 4 8; do
  [ $np -le $N ] || continue
  run weak_n$np $np --steps 5 --warmup 3
  run strong_rob_n$np $np --workload robertson_4M --scaling strong --steps 5 --warmup 3
done
