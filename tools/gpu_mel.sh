timeout 900 python -m pytest tests/test_gpu_dsp.py tests/test_gpu_engine.py tests/test_gpu_mel_variant.py -x -q 2>&1 | tail -4
timeout 300 python tools/kernel_bench.py 64 2>&1 | tail -9
