#!/bin/bash
# N GPUs of one box: PeerGather check, bench (peer and nccl transports), sharded evaluation
N=${2:-8}
OUT=gpurun_out/${1:-g8}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/test_peer_gather.py 2>&1 | grep PeerGather | tee $OUT/peer_gather_check.txt
for g in peer nccl; do
RBA_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${N}gpu_$g.json 2> $OUT/bench_${N}gpu_$g.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_${N}gpu_$g.json'));print('$g', d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks'])" || tail -5 $OUT/bench_${N}gpu_$g.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/run_evaluate_sharded.py --images $((N*32)) --out $OUT/evaluate_sharded_${N}gpu.json 2>&1 | tail -1 | cut -c1-400
