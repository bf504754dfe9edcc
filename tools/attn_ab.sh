set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -5
for nq in 1 5 6; do KB_ATTN_VARIANTS=1 KB_ATTN_NQ=$nq KB_ATTN_LIST=256 timeout 120 python tools/kernel_bench.py 64 2>&1 | tail -2; done
for nq in 1 5 6; do KB_ATTN_VARIANTS=1 KB_ATTN_NQ=$nq KB_ATTN_LIST=256 timeout 120 python tools/kernel_bench.py 229 2>&1 | tail -2; done
