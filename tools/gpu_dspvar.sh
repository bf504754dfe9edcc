for v in "" _DB200X_MEL_MIN_CTAS4DB200X_MEL_NO_STAGE _DB200X_MEL_MIN_CTAS5DB200X_MEL_NO_STAGE; do
echo "== lib$v"
B200X_LIB_PATH=/root/repo/audio-deepfake-explainability_b200/libb200xai$v.so timeout 300 python tools/kernel_bench.py 64 2>&1 | grep "mel_db"
done
