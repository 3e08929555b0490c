timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fft_z" -p no:cacheprovider 2>&1 | tail -2
python tools/fft_sizes.py 2>&1 | grep Nz
