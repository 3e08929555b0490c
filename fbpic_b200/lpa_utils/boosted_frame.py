"""`BoostConverter`: lab frame -> boosted frame conversion of simulation parameters
(fbpic/lpa_utils/boosted_frame.py:12-318).  Host-side set-up arithmetic only."""
import numpy as np
from scipy.constants import c


class BoostConverter(object):

    def __init__(self, gamma0):
        self.gamma0 = gamma0
        self.beta0 = np.sqrt(1 - 1. / gamma0**2)

    # every converter maps a list of lab-frame values to the list of boosted-frame values
    def _scaled(self, values, factor):
        return [v * factor for v in values]

    def static_length(self, lab_frame_vars):
        """L' = L / gamma0 (boosted_frame.py:30-50)"""
        return [length / self.gamma0 for length in lab_frame_vars]

    def copropag_length(self, lab_frame_vars, beta_object=1.):
        """L' = L / [gamma0 (1 - beta_object beta0)] (boosted_frame.py:52-79)"""
        return self._scaled(lab_frame_vars, 1. / (self.gamma0 * (1. - self.beta0 * beta_object)))

    def static_density(self, lab_frame_vars):
        """n' = n gamma0 (boosted_frame.py:81-101)"""
        return self._scaled(lab_frame_vars, self.gamma0)

    def copropag_density(self, lab_frame_vars, beta_object=1.):
        """n' = n gamma0 (1 - beta_object beta0) (boosted_frame.py:103-130)"""
        return self._scaled(lab_frame_vars, self.gamma0 * (1. - self.beta0 * beta_object))

    def velocity(self, lab_frame_vars):
        """v' = (v - c beta0) / (1 - beta0 v / c) (boosted_frame.py:132-152)"""
        return [(v - c * self.beta0) / (1 - v * self.beta0 / c) for v in lab_frame_vars]

    def longitudinal_momentum(self, lab_frame_vars):
        """u_z' = gamma0 (u_z - sqrt(1 + u_z^2) beta0), no transverse motion (boosted_frame.py:154-179)"""
        return [self.gamma0 * (uz - np.sqrt(1 + uz**2) * self.beta0) for uz in lab_frame_vars]

    def gamma(self, lab_frame_vars):
        """gamma' = gamma0 (gamma - beta0 sqrt(gamma^2 - 1)) (boosted_frame.py:181-207)"""
        return [self.gamma0 * (g - self.beta0 * np.sqrt(g**2 - 1)) for g in lab_frame_vars]

    def wavenumber(self, lab_frame_vars):
        """k' = k / (gamma0 (1 + beta0)) (boosted_frame.py:209-229)"""
        return [k / (self.gamma0 * (1 + self.beta0)) for k in lab_frame_vars]

    def boost_particle_arrays(self, x, y, z, ux, uy, uz, inv_gamma):
        """Lorentz-transform a particle distribution and propagate it ballistically to the boosted-frame
        time t' = 0 (boosted_frame.py:231-279)."""
        uz_frame = self.gamma0 * self.beta0
        t_b = -uz_frame * z / c
        gamma_lab = np.sqrt(1. + (ux * ux + uy * uy + uz * uz))
        nux, nuy = ux.copy(), uy.copy()
        nuz = self.gamma0 * uz - uz_frame * gamma_lab
        gamma_b = np.sqrt(1. + (nux**2 + nuy**2 + nuz**2))
        nx = x - t_b * nux * c / gamma_b
        ny = y - t_b * nuy * c / gamma_b
        nz = self.gamma0 * z - t_b * nuz * c / gamma_b
        return nx, ny, nz, nux, nuy, nuz, 1. / gamma_b

    def interaction_time(self, L_interact, l_window, v_window):
        """Time for the moving window to slide across the plasma, in the boosted frame
        (boosted_frame.py:281-318)."""
        L_i, = self.static_length([L_interact])
        l_w, = self.copropag_length([l_window])
        v_w, = self.velocity([v_window])
        return (L_i + l_w) / (v_w + self.beta0 * c)
