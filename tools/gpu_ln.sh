timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_reduce.py -x -q 2>&1 | tail -4
for d in 0 1; do echo "== debug $d"; B200X_RLN_DEBUG=$d timeout 300 python tools/kernel_bench.py 229 2>&1 | grep "tail\|resid\|layernorm"; done
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_fullsize_golden.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r02p_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02p_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_class'], d['roofline']['whole_forward_frac_of_peak'])"
