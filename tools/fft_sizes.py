"""Time the strided z-FFT (cuFFT Z2Z, [Nz][Nr] layout) for candidate local grid lengths."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fbpic_b200 import _lib
from fbpic_b200._lib import DeviceArray, call

ctx = _lib.context()
Nr = 256
ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
call.b2_event_create(ctypes.byref(ev0)); call.b2_event_create(ctypes.byref(ev1))
print('impl', os.environ.get('B2_FFT_IMPL', 'own (two-pass) where planned'), {k: v for k, v in os.environ.items() if k.startswith('B2_FFT_')})
for Nz in [int(v) for v in os.environ.get('FFT_SIZES', '4096,4224,4416,2240,2048').split(',')]:
    a = DeviceArray.zeros((Nz, Nr), np.complex128)
    b = DeviceArray.zeros((Nz, Nr), np.complex128)
    for _ in range(3):
        call.b2_fft_z(ctx.handle, a.ptr, b.ptr, Nz, Nr, 0, None)
    call.b2_event_record(ev0, ctx.stream)
    for _ in range(20):
        call.b2_fft_z(ctx.handle, a.ptr, b.ptr, Nz, Nr, 0, None)
    call.b2_event_record(ev1, ctx.stream)
    ms = ctypes.c_float(0.)
    call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms))
    t1 = ms.value / 20 * 1e3
    # the batched call of the PIC step: 12 independent arrays (E, B of two modes)
    from fbpic_b200._lib import ptr_array
    A = [DeviceArray.zeros((Nz, Nr), np.complex128) for _ in range(12)]
    B = [DeviceArray.zeros((Nz, Nr), np.complex128) for _ in range(12)]
    pa, pb = ptr_array(A), ptr_array(B)
    for _ in range(3):
        call.b2_fft_z_multi(ctx.handle, 12, pa, pb, Nz, Nr, 2, None)
    call.b2_event_record(ev0, ctx.stream)
    for _ in range(10):
        call.b2_fft_z_multi(ctx.handle, 12, pa, pb, Nz, Nr, 2, None)
    call.b2_event_record(ev1, ctx.stream)
    call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms))
    t12 = ms.value / 120 * 1e3
    print('Nz=%d  single call %.1f us/FFT (%.0f GB/s)   batch of 12: %.1f us/FFT (%.0f GB/s)' % (
        Nz, t1, 2 * Nz * Nr * 16 / (t1 * 1e-6) / 1e9, t12, 2 * Nz * Nr * 16 / (t12 * 1e-6) / 1e9))
    del A, B
