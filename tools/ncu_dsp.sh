#!/bin/bash
# ncu --set full capture of the DSP kernels (kernel_bench, 64 copies) + plain kernel_bench timing.  usage: bash tools/ncu_dsp.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py 64 > gpurun_out/${TAG}_kernel_bench64.txt 2>&1; tail -22 gpurun_out/${TAG}_kernel_bench64.txt
KB_ITERS=1 KB_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mel_db|istft_masked|stft_kernel|mel_resize|mel_stats' -c 16 \
   -o gpurun_out/${TAG}_ncu_dsp -f python tools/kernel_bench.py 64 > gpurun_out/${TAG}_ncu_dsp.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}_ncu_dsp.ncu-rep
