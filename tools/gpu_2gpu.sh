#!/bin/bash
OUT=gpurun_out/${1:-g2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/test_peer_gather.py 2>&1 | tail -5 | tee $OUT/peer_gather_check.txt
for g in peer nccl; do
RBA_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu_$g.json 2> $OUT/bench_2gpu_$g.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_2gpu_$g.json'));print('$g', d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['parallelism'])" || tail -5 $OUT/bench_2gpu_$g.err
done
