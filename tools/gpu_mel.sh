timeout 900 python -m pytest tests/test_gpu_mel_variant.py -x -q 2>&1 | tail -30
