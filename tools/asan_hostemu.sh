#!/bin/bash
# Kernel SOURCE of fbpic_b200/csrc/b2_ext_kernels.cuh under AddressSanitizer (a CPU stand-in for compute-sanitizer in
# the GPU-less build container): tests/hostemu is rebuilt with -fsanitize=address and tests/test_hostemu_ext.py runs with
# the ASan runtime preloaded, so that every out-of-bounds access of a kernel on the exact-size NumPy buffers aborts.
#   bash tools/asan_hostemu.sh
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
SO=tests/hostemu/libemu_ext.so
[ -f $SO ] && cp $SO /tmp/libemu_ext.bak
g++ -O1 -g -ffp-contract=off -fsanitize=address -fno-omit-frame-pointer -Db2ext=b2ext_hostemu -shared -fPIC \
    -Wl,-Bsymbolic -o $SO tests/hostemu/emu_ext.cpp
touch $SO
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$ASAN python -m pytest tests/test_hostemu_ext.py -x -q -p no:cacheprovider || RC=$?
# ... and the host flows that launch these kernels on the arrays of whole simulations (the fake device then allocates
# its "device" buffers without slack: B2_FAKE_PAD=0)
B2_FAKE_PAD=0 ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$ASAN python -m pytest tests/test_host_flow.py -x -q -p no:cacheprovider \
    -k "pml_flow or cross_flow or lab_frame or ionization_kernels or compton_momentum or antenna_flow or ext_kernel or bunch_injection_plane" || RC=$?
if [ -f /tmp/libemu_ext.bak ]; then cp /tmp/libemu_ext.bak $SO; else rm -f $SO; fi
touch $SO
exit ${RC:-0}
