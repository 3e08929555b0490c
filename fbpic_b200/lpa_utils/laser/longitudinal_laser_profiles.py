"""Import path of `fbpic.lpa_utils.laser.longitudinal_laser_profiles`; the classes live in the package itself."""
from . import (LaserLongitudinalProfile, GaussianChirpedLongitudinalProfile,                      # noqa: F401
               CustomSpectrumLongitudinalProfile)
