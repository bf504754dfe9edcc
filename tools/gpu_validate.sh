#!/bin/bash
# Full validation on one B200: GPU tests, smoke, both bench arms.  usage: bash tools/gpu_validate.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/${TAG}_gpu_tests.txt
tail -3 gpurun_out/${TAG}_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["whole_forward_frac_of_peak"])
    print(d["roofline"]["ms_per_class"])
    print("strong", d.get("strong",{}).get("value"), "workloads", {k:v.get("value") for k,v in d.get("workloads",{}).items()})
    print("hbm", {k:(round(v["gbs"]),round(v["frac"],3)) for k,v in d["roofline"].get("hbm_stages",{}).items() if isinstance(v,dict)})
    print("topk", json.dumps(d.get("topk"))[:1500])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "ref rc=$?"; head -c 600 gpurun_out/${TAG}_bench_reference.json
