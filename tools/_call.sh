timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "fft or transforms" 2>&1 | tail -3
FFT_SIZES=4096,4224,4416,2240,2048 python tools/fft_sizes.py 2>&1 | tail -6
