timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_dsp.py tests/test_gpu_mel_variant.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02u2_bench.json 2> gpurun_out/r02u2_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r02u2_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02u2_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_class'], d['roofline']['whole_forward_frac_of_peak'], d['roofline']['frac'])"
