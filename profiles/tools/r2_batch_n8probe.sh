#!/bin/bash
# Round 2: the default line on 8 GPUs of one box with the D2H copy-rate probe (the ceiling of the end-to-end rate)
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n8_weak.json 2> gpurun_out/r2n8_weak.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2n8_weak.json').read().strip().splitlines()[-1]); e=d['e2e']; print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(e['value']), 'host writes', round(e['host_write_gbs_all_ranks'],1), 'GB/s; plain D2H probe', e.get('d2h_copy_probe_gbs_all_ranks'), 'bound', e.get('d2h_bound_on_e2e'), 'gather', round(d['value_with_gather']['value']))"
