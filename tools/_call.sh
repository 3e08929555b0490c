timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "fft or transforms" 2>&1 | tail -5
python tools/fft_sizes.py 2>&1 | tail -6
for v in "B2_FFT_LAG=2 B2_FFT_RING=6" "B2_FFT_LAG=4 B2_FFT_RING=8" "B2_FFT_GC=16 B2_FFT_LAG=2 B2_FFT_RING=5"; do
  env $v FFT_SIZES=4096,4224 python tools/fft_sizes.py 2>&1 | tail -3
done
