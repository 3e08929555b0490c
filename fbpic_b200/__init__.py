"""
fbpic_b200 -- B200-native (sm_100a) implementation of FBPIC's per-step PIC hot loop
behind the reference's own operator surface (Simulation.step / Particles / Fields /
BoundaryCommunicator).  Host code is Python; every per-step operation runs in
hand-written CUDA (plus cuFFT for the z-FFT) through the C ABI of
include/fbpic_b200.h.  There is no CPU fallback.
"""
__version__ = '0.1.0'

from .main import Simulation, GpuMemoryManager    # noqa: F401
from .particles import Particles                  # noqa: F401
from .fields import Fields, BinomialSmoother      # noqa: F401
from .boundaries import BoundaryCommunicator      # noqa: F401
from ._lib import DeviceArray, cuda_available, B200Error   # noqa: F401


def set_random_seed(random_seed):
    """Fix the seeds of the Monte-Carlo parts (plasma loading, bunches, ionization, Compton scattering): rank r uses
    random_seed + r (fbpic/utils/random_seed.py).  The device generators of the elementary processes take their seeds
    from `np.random` when `make_ionizable` / `activate_compton` are called, so call this first."""
    import os
    import numpy as np
    np.random.seed(random_seed + int(os.environ.get('RANK', '0')))
