#!/bin/bash
# closing 2-GPU check of the round: the driver's own launch line (torchrun, default transport) + the world-size-2 GPU test
OUT=gpurun_out/g2s; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_2gpu.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['parallelism'],d['clocks'])" || tail -5 $OUT/bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_2gpu_reference.json 2> $OUT/bench_2gpu_reference.err; echo "ref rc=$?"; head -c 300 $OUT/bench_2gpu_reference.json
