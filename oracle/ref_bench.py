"""oracle/ref_bench.py -- times the UNMODIFIED reference (FBPIC numba CPU path, installed by oracle/make_ref.sh
into oracle/_ref) on a bench.py workload.  MEASUREMENT INFRASTRUCTURE (the `--impl reference` arm and nothing
else): run as a subprocess of bench.py so that the numba / OpenBLAS thread environment (NUMBA_THREADING_LAYER=omp,
OPENBLAS_NUM_THREADS=1, NUMBA_NUM_THREADS=<threads>) is set before anything is imported.

    python oracle/ref_bench.py '<json: {cfg, Nz, steps, warmup, threads, laser}>'   -> one JSON line on stdout

The simulation is built through the reference's own public API (fbpic.main.Simulation(use_cuda=False), electrons from
its own evenly-spaced loader with np.random.seed(0), z periodic) with the same analytic laser pulse written to the
mode-1 interpolation grid as bench.py writes for the B200 arm (the formula is passed in as arrays, see bench.py:laser_fields)."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    job = json.loads(sys.argv[1])
    threads = int(job['threads'])
    os.environ['NUMBA_THREADING_LAYER'] = 'omp'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['OMP_NUM_THREADS'] = str(threads)
    os.environ['NUMBA_NUM_THREADS'] = str(threads)
    os.environ['FBPIC_DISABLE_CACHING'] = os.environ.get('FBPIC_DISABLE_CACHING', '0')
    sys.path.insert(0, os.path.join(HERE, 'ref_shim'))      # pyfftw -> scipy.fft (neither MKL nor pyfftw in the image)
    sys.path.insert(0, os.path.join(HERE, '_ref'))
    import numpy as np
    from scipy.constants import c
    t_imp = time.perf_counter()
    from fbpic.main import Simulation
    import fbpic
    assert os.path.realpath(fbpic.__file__).startswith(os.path.realpath(os.path.join(HERE, '_ref'))), fbpic.__file__
    cfg = job['cfg']
    Nz = int(job['Nz'])
    np.random.seed(0)
    zmax = Nz * cfg['dz']
    dt = cfg['dz'] / c
    p_nz, p_nr, p_nt = cfg['ppc']
    sim = Simulation(Nz, zmax, cfg['Nr'], cfg['rmax'], cfg['Nm'], dt, p_zmin=0., p_zmax=zmax, p_rmin=0.,
                     p_rmax=cfg['rmax'], p_nz=p_nz, p_nr=p_nr, p_nt=p_nt, n_e=cfg['n_e'], n_order=-1,
                     boundaries={'z': 'periodic', 'r': 'reflective'}, use_cuda=False, verbose_level=0)
    if job.get('laser', True):
        sys.path.insert(0, os.path.dirname(HERE))
        from bench import laser_fields
        g1 = sim.fld.interp[1]
        Er1, Et1, Br1, Bt1 = laser_fields(g1.z, g1.r, z0=0.5 * zmax)
        g1.Er[:, :], g1.Et[:, :], g1.Br[:, :], g1.Bt[:, :] = Er1, Et1, Br1, Bt1
    Ntot = int(sum(s.Ntot for s in sim.ptcl))
    t0 = time.perf_counter()
    sim.step(max(int(job['warmup']), 1), show_progress=False)       # includes the numba JIT
    t_warm = time.perf_counter() - t0
    steps = int(job['steps'])
    budget = float(job.get('budget_s', 0.))
    if budget > 0.:                       # bounded run: one probe step decides how many steps fit
        t0 = time.perf_counter()
        sim.step(1, show_progress=False)
        t1 = time.perf_counter() - t0
        steps = max(1, min(steps, int(budget / max(t1, 1e-9))))
    t0 = time.perf_counter()
    sim.step(steps, show_progress=False)
    dt_run = time.perf_counter() - t0
    import numba
    print(json.dumps({'value': Ntot * steps / dt_run, 'ms_per_step': dt_run / steps * 1e3, 'Ntot': Ntot,
                      'steps': steps, 'threads': threads, 'warmup_s': t_warm, 'import_s': t0 - t_imp,
                      'numba': numba.__version__, 'threading_layer': numba.threading_layer(),
                      'fbpic': fbpic.__version__, 'Nz': Nz}))


if __name__ == '__main__':
    main()
