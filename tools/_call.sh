mkdir -p gpurun_out
for g in 1 2 4 6; do echo "== group $g"; B2_FFT_GROUP=$g python tools/fft_sizes.py 2>&1 | grep Nz; done | tee gpurun_out/r02_fft_group.txt
echo "== cufft"; B2_FFT_IMPL=cufft python tools/fft_sizes.py 2>&1 | grep Nz | tee -a gpurun_out/r02_fft_group.txt
