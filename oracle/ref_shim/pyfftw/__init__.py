"""Stand-in for `pyfftw`, used ONLY to import the unmodified FBPIC reference in
the build container (neither MKL nor pyfftw is installed there).  It exposes
the three calls the reference makes (fbpic/fields/spectral_transform/
fourier.py:98-101,132-134,166-168) on top of scipy.fft.  Test infrastructure."""
import scipy.fft as _sf


class FFTW(object):
    def __init__(self, a, b, axes=(0,), direction='FFTW_FORWARD', threads=1):
        self.a, self.b = a, b
        self.axis = axes[0]
        self.forward = (direction == 'FFTW_FORWARD')
        self.threads = threads

    def update_arrays(self, new_input_array, new_output_array):
        self.a, self.b = new_input_array, new_output_array

    def __call__(self):
        f = _sf.fft if self.forward else _sf.ifft
        self.b[...] = f(self.a, axis=self.axis, workers=self.threads)
