"""Import path of `fbpic.lpa_utils.laser.transverse_laser_profiles`; the classes live in the package itself."""
from . import (LaserTransverseProfile, GaussianTransverseProfile, LaguerreGaussTransverseProfile,  # noqa: F401
               DonutLikeLaguerreGaussTransverseProfile, FlattenedGaussianTransverseProfile)
