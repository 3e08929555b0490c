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
for Nz in (4096, 4222, 4224, 4256, 4290, 4312, 4320, 4352, 4400, 4410, 4480):
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
    print('Nz=%d  n_guard=%d  %.1f us/FFT  %.0f GB/s' % (Nz, (Nz - 4096) // 2, ms.value / 20 * 1e3,
                                                       2 * Nz * Nr * 16 / (ms.value / 20 * 1e-3) / 1e9))
