#!/bin/bash
# (gpurun copies back at most 64 MiB: the two ncu captures are capped at 10 / 6 kernels)
# Round validation on one B200: GPU parity tests, smoke, both bench arms, ncu launch list and full captures.
# usage (under gpurun): bash tools/gpu_validate.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu_$tag.txt
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?" | tee -a $out/summary_$tag.txt
python __graft_entry__.py --smoke > $out/smoke_$tag.log 2>&1; echo "smoke exit $?" | tee -a $out/summary_$tag.txt
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; echo "ref exit $?" | tee -a $out/summary_$tag.txt
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench exit $?" | tee -a $out/summary_$tag.txt
python tools/kernel_bench.py 64 > $out/kb_$tag.txt 2>&1; echo "kb exit $?" | tee -a $out/summary_$tag.txt
KB_ATTN_VARIANTS=1 python tools/kernel_bench.py 64 > $out/kbv_$tag.txt 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launches_$tag.log 2>&1
echo "ncu launches exit $?" | tee -a $out/summary_$tag.txt
KB_ITERS=1 KB_WARMUP=1 python tools/kernel_bench.py 64 > $out/plain_kb_$tag.log 2>&1 &&
KB_ITERS=1 KB_WARMUP=1 ncu --set full --clock-control none -k regex:'gemm|attention|layernorm|istft|mel' -c 10 \
    -o $out/prof_$tag -f python tools/kernel_bench.py 64 > $out/ncu_full_$tag.log 2>&1
echo "ncu full exit $?" | tee -a $out/summary_$tag.txt
# DRAM traffic of the dominant kernels at the bench's own launch shape (228-copy chunk): 2 launches each, skipping the warm-up steps
ncu --set full --clock-control none -k regex:'attention|gemm2' -s 30 -c 6 \
    -o $out/prof_bench_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_bench_$tag.log 2>&1
echo "ncu bench exit $?" | tee -a $out/summary_$tag.txt
tail -3 $out/pytest_gpu_$tag.log; cat $out/bench_$tag.json; cat $out/kb_$tag.txt $out/kbv_$tag.txt
