"""Device time per PIC cycle of the solver variants around the hot loop at the C2 grid size (Nz=4096, Nr=256, Nm=2,
16.8 M particles), next to the default periodic cycle: radial PML, cross-deposition, laser antenna, external field,
open z + moving window.  One JSON line per variant (CUDA events on the context stream around K cycles, data resident
in HBM).  Not part of bench.py's contract: a tool for deciding what to optimise next."""
import ctypes
import json
import math
import os
import sys
import numpy as np
from scipy.constants import c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbpic_b200 import Simulation, _lib                                       # noqa: E402
from fbpic_b200._lib import call                                              # noqa: E402
from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser        # noqa: E402
from fbpic_b200.lpa_utils.external_fields import ExternalField               # noqa: E402

Nz, Nr, Nm, dz, dr = 4096, 256, 2, 0.05e-6, 0.4e-6
N_ORDER = 32
zmax, rmax = Nz * dz, Nr * dr


def undulator(F, x, y, z, t, amplitude, length_scale):
    return F + amplitude * math.cos(2 * math.pi * z / length_scale)


def build(variant):
    np.random.seed(0)
    kw = dict(p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2, p_nt=4, n_e=4.e24, sort_period=4)
    if variant == 'default':
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, **kw)
    elif variant == 'pml':
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, boundaries={'z': 'periodic', 'r': 'open'}, **kw)
    elif variant == 'cross_deposition':
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, current_correction='cross-deposition', **kw)
    elif variant == 'external_field':
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, **kw)
        sim.external_fields = [ExternalField(undulator, 'By', 1., 1.e-5)]
    elif variant in ('open_window', 'antenna'):
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, n_order=N_ORDER, boundaries={'z': 'open', 'r': 'reflective'}, **kw)
        sim.set_moving_window(v=c)
        if variant == 'antenna':
            add_laser_pulse(sim, GaussianLaser(1., 20.e-6, 16.e-15, -10.e-6, zf=0.5 * zmax), method='antenna',
                            z0_antenna=0.1 * zmax)
    elif variant in ('field_diag', 'lab_diag'):
        # diagnostics overhead: a field output every 10 cycles resp. 10 lab-frame snapshots (fields + particles)
        # fed every cycle; files go to a scratch directory
        import tempfile
        from fbpic_b200.openpmd_diag import (FieldDiagnostic, BackTransformedFieldDiagnostic,
                                             BackTransformedParticleDiagnostic)
        out = tempfile.mkdtemp()
        if variant == 'field_diag':
            sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, **kw)
            sim.diags = [FieldDiagnostic(10, sim.fld, comm=sim.comm, fieldtypes=['E', 'B', 'rho'], write_dir=out)]
        else:
            # (the run itself is not boosted: the lab-frame diagnostics only slice it, which is what is timed)
            sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, **kw)
            sim.ptcl[0].track(sim.comm)
            args = (0., 5. * zmax, 0., 0.2 * zmax / c, 10, 5., 20, sim.fld)
            sim.diags = [BackTransformedFieldDiagnostic(*args, comm=sim.comm, write_dir=out),
                         BackTransformedParticleDiagnostic(*args, species={'electrons': sim.ptcl[0]}, comm=sim.comm,
                                                           write_dir=out)]
    elif variant == 'ionization':
        # a second, ionizable species (nitrogen, same sampling) that feeds the electrons: the ADK pass, the
        # per-particle-charge push and the unfused route of that species (the electrons keep the fused kernels)
        from scipy.constants import m_p
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, dz / c, **kw)
        ions = sim.add_new_species(q=0, m=14. * m_p, n=4.e24, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=2,
                                   p_nr=2, p_nt=4)
        ions.make_ionizable('N', target_species=sim.ptcl[0], level_start=1)
    else:
        raise ValueError(variant)
    sp = sim.ptcl[0]
    k0 = 2 * np.pi / zmax * 8
    sp.uz = 0.05 * np.sin(k0 * sp.z) * np.exp(-(sp.x**2 + sp.y**2) / (30.e-6)**2)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.uz**2)
    return sim


def main():
    steps = int(os.environ.get('B2_VARIANT_STEPS', '40'))
    ctx = _lib.context()
    ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
    call.b2_event_create(ctypes.byref(ev0))
    call.b2_event_create(ctypes.byref(ev1))
    for variant in ('default', 'pml', 'cross_deposition', 'external_field', 'open_window', 'antenna', 'field_diag',
                    'lab_diag', 'ionization'):
        try:
            sim = build(variant)
            n = sum(s.Ntot for s in sim.ptcl)
            sim.step(10, keep_on_gpu=True)
            call.b2_device_sync()
            l0 = _lib.load().b2_launch_count()
            call.b2_event_record(ev0, ctx.stream)
            sim.step(steps, keep_on_gpu=True)
            call.b2_event_record(ev1, ctx.stream)
            ms = ctypes.c_float(0.)
            call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms))
            print(json.dumps({'variant': variant, 'ms_per_step': ms.value / steps, 'particles': n,
                              'particle_updates_per_s': n * steps / (ms.value * 1e-3),
                              'launches_per_step': (_lib.load().b2_launch_count() - l0) / steps,
                              'Nz_local': sim.fld.interp[0].Nz, 'Nr_local': sim.fld.interp[0].Nr}), flush=True)
            del sim
        except Exception as exc:                                              # keep going: one line per variant
            print(json.dumps({'variant': variant, 'error': repr(exc)[:300]}), flush=True)


if __name__ == '__main__':
    main()
