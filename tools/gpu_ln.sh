timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4
timeout 300 python tools/kernel_bench.py 229 2>&1 | grep "qkv\|fc1"
timeout 300 python tools/kernel_bench.py 64 2>&1 | grep "qkv\|fc1"
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02p3_bench.json 2> gpurun_out/r02p3_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r02p3_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02p3_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_class'], d['roofline']['whole_forward_frac_of_peak'], d['roofline']['gemm_class'])"
