#!/bin/bash
# bench.py on N GPUs of one box (weak headline + strong block), plus --scaling strong as the headline.  usage: bash tools/gpu_multi.sh N tag
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "weak rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --scaling strong --no-extras \
   > gpurun_out/${TAG}_bench_strong_${N}gpu.json 2> gpurun_out/${TAG}_bench_strong_${N}gpu.err; echo "strong rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_strong_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_${N}gpu.json", "gpurun_out/${TAG}_bench_strong_${N}gpu.json"):
    try:
        d = json.load(open(f))
        print(f, d["scaling"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "strong-block", d.get("strong", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "parse failed", e)
PY
