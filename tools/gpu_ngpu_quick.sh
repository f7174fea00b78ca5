#!/bin/bash
# bench.py on N GPUs (the driver's launch line, no baselines): prints the line's per-rank diagnostics
N=${1:-2}; OUT=gpurun_out/gq$N; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench_${N}gpu.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks']);print(json.dumps(d.get('per_rank')))" || tail -5 $OUT/bench_${N}gpu.err
