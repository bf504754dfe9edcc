#!/bin/bash
# Quick GPU loop for kernel work on one B200: the kernel-level GPU tests, the per-kernel timings at 64 and 229 copies and a short
# headline bench.  usage: gpurun -- bash tools/gpu_quick.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_dsp.py tests/test_gpu_reduce.py -x -q 2>&1 | tail -4
timeout 300 python tools/kernel_bench.py 64 2>&1 | tail -24
timeout 300 python tools/kernel_bench.py 229 2>&1 | head -16
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_class'], d['roofline']['whole_forward_frac_of_peak'])"
