timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k "fullrow" 2>&1 | tail -8
timeout 300 python tools/kernel_bench.py 229 2>&1 | grep "resid\|tail\|full-row\|layernorm"
