"""
ADK field ionization (fbpic/particles/elementary_process/ionization/{ionizer,numba_methods,cuda_methods,
inline_functions,read_atomic_data}.py; Chen et al., JCP 236 (2013), eq. 2), fully relativistic: the rate is evaluated
with the field amplitude in the rest frame of the ion and the proper time of one cycle, so it also holds in a boosted
frame.

As in the reference, ions of all charge states of an element live in ONE `Particles` object with a per-particle
`ionization_level` (here `Ionizer.levels`, an 8-byte array that follows the particles through the cell sort and the
particle exchange with the plumbing of the tracked ids) and deposit with the weight `w_times_level`; the freed
electrons are appended to the target species.

What differs (B200-first, not a port): the reference counts the new electrons per batch of 10 ions, brings the counts
to the host for a cumulative sum, reallocates and fills the electron arrays in a second batched pass
(ionizer.py:186-330).  Here `b2_ionize` appends (ion index, former level) of every ionization event to a short device
list with one atomic counter; the host reads back only that list (a few entries per cycle), sorts it -- which also
makes the result independent of the order of the atomics -- and `b2_permute`, used as a gather, copies the 8 state
arrays of those ions to the end of the electron arrays.  Random numbers: a counter-based generator in the kernel keyed
by (seed, cycle, ion index); `Ionizer.host_draws = True` draws them with `np.random.rand` on the host instead, as the
reference's CPU path does (ionizer.py:221).
"""
import ctypes
import json
import os
import numpy as np
from scipy.constants import c, e, m_e, physical_constants
from scipy.special import gamma as gamma_function

from . import _lib
from ._lib import DeviceArray, call, ptr_array

_ENERGIES = None


def get_ionization_energies(element):
    """Ionization energies (J) of all charge states of `element` (atomic symbol), or None if it is not tabulated
    (read_atomic_data.py:14-91; NIST data, fbpic_b200/ionization_energies.json)."""
    global _ENERGIES
    if _ENERGIES is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ionization_energies.json')) as f:
            _ENERGIES = json.load(f)['ionization_energies_eV']
    if element not in _ENERGIES:
        return None
    return e * np.array(_ENERGIES[element])


class Ionizer(object):
    """Ionization data of one ionizable species (ionizer.py:41-183) and the per-cycle ionization step."""
    host_draws = False

    def __init__(self, element, ionizable_species, target_species, level_start, level_max=None):
        from .particles import LevelCarrier, Particles
        self.level_start, self.level_max = level_start, level_max
        self.initialize_ADK_parameters(element, ionizable_species.dt)
        self.levels = LevelCarrier(level_start, ionizable_species.Ntot)
        self.w_times_level = np.asarray(ionizable_species.w) * self.levels.id
        if type(target_species) is dict:
            for level in range(self.level_start, self.level_max):
                if level not in target_species:
                    raise ValueError('When passing a dictionary for `target_species`, its keys should be\nthe '
                                     'integers corresponding to the ionizable levels.\n (i.e. the integers from %d to '
                                     '%d for %s with level_start=%d.)' % (self.level_start, self.level_max, element,
                                                                         self.level_start))
                assert isinstance(target_species[level], Particles)
            self.target_species = [target_species[level] for level in range(self.level_start, self.level_max)]
            self.store_electrons_per_level = True
        elif isinstance(target_species, Particles):
            self.target_species = [target_species]
            self.store_electrons_per_level = False
        else:
            raise ValueError("Unexpected type for target_species: %s\nPlease pass a `Particles` object, or a dictionary"
                             % type(target_species))
        for species in self.target_species:
            assert species.q == -e and species.m == m_e
        self.seed = int(np.random.randint(0, 2**31 - 1))      # reproducible with np.random.seed
        self.n_calls = 0
        self._events = self._count = self._tables = None

    @property
    def ionization_level(self):
        """per-particle charge state (host array between `step()` calls)"""
        a = self.levels.id
        return a.get() if isinstance(a, DeviceArray) else a

    def initialize_ADK_parameters(self, element, dt):
        """Per-level tables of the ADK probability per cycle (ionizer.py:137-183):
        W dt = prefactor E^power exp(exp_prefactor / E), with the effective quantum numbers n* = Z sqrt(U_H / U)."""
        Uion = get_ionization_energies(element)
        if Uion is None:
            raise ValueError("Unknown ionizable element %s.\n" % element
                             + "Please use atomic symbol (e.g. 'He') not full name (e.g. Helium)")
        self.element = element
        if self.level_max is None:
            self.level_max = len(Uion)
        else:
            assert type(self.level_max) is int, "level_max must be integer"
            if self.level_max > len(Uion):
                raise ValueError("Chosen level_max for {}".format(element) + " cannot exceed {}".format(len(Uion)))
        alpha = physical_constants['fine-structure constant'][0]
        r_e = physical_constants['classical electron radius'][0]
        wa = alpha**3 * c / r_e                   # atomic unit of frequency
        Ea = m_e * c**2 / e * alpha**4 / r_e      # atomic unit of electric field
        UH = get_ionization_energies('H')[0]
        n_eff = (np.arange(len(Uion)) + 1) * np.sqrt(UH / Uion)
        l_eff = n_eff[0] - 1
        C2 = 2**(2 * n_eff) / (n_eff * gamma_function(n_eff + l_eff + 1) * gamma_function(n_eff - l_eff))
        self.adk_power = -(2 * n_eff - 1)
        self.adk_prefactor = dt * wa * C2 * (Uion / (2 * UH)) * (2 * (Uion / UH)**(3. / 2) * Ea)**(2 * n_eff - 1)
        self.adk_exp_prefactor = -2. / 3 * (Uion / UH)**(3. / 2) * Ea

    # ---- residency ----
    def send_to_gpu(self, species):
        """(the level array itself travels with `species.uint_carriers()`)"""
        self.w_times_level = DeviceArray(species._capacity, np.float64).view((species.Ntot,))
        self.update_weights(species)

    def receive_from_gpu(self, species):
        """(kept current on the device after every ionization, sort and exchange)"""
        self.w_times_level = self.w_times_level.get()

    def update_weights(self, species):
        """w_times_level = w * level on the device (after an ionization, a sort or an exchange)"""
        n = species.Ntot
        if not isinstance(self.w_times_level, DeviceArray) or self.w_times_level.capacity < n:
            self.w_times_level = DeviceArray(max(species._capacity, n), np.float64)
        self.w_times_level = self.w_times_level.view((n,))
        call.b2_w_times_level(_lib.context().handle, n, species.w.ptr, self.levels.id.ptr, self.w_times_level.ptr, None)

    # ---- the per-cycle step ----
    def handle_ionization(self, ion):
        """One ADK draw per ion; every event frees one electron that starts with the position and momentum of its ion
        (ionizer.py:186-330)."""
        n = ion.Ntot
        if n == 0:
            return
        ion._need_gpu()
        ctx = _lib.context().handle
        if self._count is None:
            self._count = DeviceArray(1, np.int64)
            self._tables = [DeviceArray.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
                            for a in (self.adk_prefactor, self.adk_power, self.adk_exp_prefactor)]
        if self._events is None or self._events.size < 2 * n:
            self._events = DeviceArray(2 * ion._capacity_for(n), np.int64)
        draws = DeviceArray.from_numpy(np.random.rand(n)) if self.host_draws else None
        self.n_calls += 1
        found = ctypes.c_int64(0)
        call.b2_ionize(ctx, n, self.levels.id.ptr, self.level_max, self._tables[0].ptr, self._tables[1].ptr,
                       self._tables[2].ptr, ion.ux.ptr, ion.uy.ptr, ion.uz.ptr, ion.Ex.ptr, ion.Ey.ptr, ion.Ez.ptr,
                       ion.Bx.ptr, ion.By.ptr, ion.Bz.ptr, draws.ptr if draws is not None else None,
                       (self.seed * 1000003 + self.n_calls) & (2**64 - 1), self._events.size // 2, self._events.ptr,
                       self._count.ptr, ctypes.byref(found), None)
        k = int(found.value)
        if k == 0:
            return
        self.update_weights(ion)
        events = self._events.view((k, 2)).get()
        events = events[np.argsort(events[:, 0], kind='stable')]          # ion order, whatever the atomics did
        for i_level, elec in enumerate(self.target_species):
            if self.store_electrons_per_level:
                idx = events[events[:, 1] == self.level_start + i_level, 0]
            else:
                idx = events[:, 0]
            if len(idx) == 0:
                continue
            elec._need_gpu()
            d_idx = DeviceArray.from_numpy(np.ascontiguousarray(idx, dtype=np.int64))
            old = elec.Ntot
            elec.grow_device_arrays(old + len(idx))
            from .particles import FLOAT_ATTRS
            call.b2_permute(ctx, len(idx), d_idx.ptr, len(FLOAT_ATTRS), ptr_array([getattr(ion, a) for a in FLOAT_ATTRS]),
                            ptr_array([getattr(elec, a).ptr + 8 * old for a in FLOAT_ATTRS]), None)
