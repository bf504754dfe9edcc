for v in 256 1 2 3 4 8 15 16; do echo "variant $v"; KB_ATTN_VARIANTS=1 KB_ATTN_NQ=5 KB_ATTN_LIST=$v timeout 120 python tools/kernel_bench.py 64 2>&1 | tail -1; done
