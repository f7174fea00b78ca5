#!/bin/bash
# closing 8-GPU bench of the round: the driver's own launch line (default transport), then 1 GPU on the same box for the efficiency
OUT=gpurun_out/g8t; mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${1:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_${N}gpu.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['parallelism'],d['clocks'])" || tail -5 $OUT/bench_${N}gpu.err
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_1gpu_same_box.json 2> $OUT/bench_1gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_1gpu_same_box.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks'])"
