#!/bin/bash
# ncu evidence of the bench command on one B200: launch list of `bench.py --steps 1`, one --set full capture of the transformer
# kernels inside the bench (229-copy launches) and one of the DSP kernels (kernel_bench, 64 copies).  usage: bash tools/ncu_bench.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches_bench_steps1.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
# the second layer of the first forward after the warm-up passes: skip the DSP + tokenizer + first-layer launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attention_kernel|gemm2_|layernorm_kernel' -s 24 -c 9 \
   -o gpurun_out/${TAG}_ncu_bench229 -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
timeout 300 python tools/kernel_bench.py 64 > gpurun_out/${TAG}_kernel_bench64.txt 2>&1 || exit 1
KB_ITERS=1 KB_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mel_db|istft_masked|stft_kernel|mel_resize|mel_stats' -c 12 \
   -o gpurun_out/${TAG}_ncu_dsp -f python tools/kernel_bench.py 64 > gpurun_out/${TAG}_ncu_dsp.log 2>&1; echo "dsp rc=$?"
ls -la gpurun_out/${TAG}_*.ncu-rep
