timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "fft or transforms" 2>&1 | tail -3
FFT_SIZES=4224,4416,4352,2176 python tools/fft_sizes.py 2>&1 | tail -5
