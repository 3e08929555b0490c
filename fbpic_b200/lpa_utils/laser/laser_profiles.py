"""Import path of the reference's profile module (`fbpic.lpa_utils.laser.laser_profiles`); the classes live in the
package itself."""
from . import (LaserProfile, SummedLaserProfile, ParaxialApproximationLaser, GaussianLaser,      # noqa: F401
               LaguerreGaussLaser, DonutLikeLaguerreGaussLaser, FlattenedGaussianLaser, FewCycleLaser, FromLasyFileLaser)
