"""
bench.py -- PIC hot-loop throughput of fbpic_b200 on B200 (and of the CPU oracle).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): Nz=4096 per GPU, Nr=256, Nm=2, linear
shapes, uniform electrons 2x2x4 per cell (16.8 M macro-particles per GPU) filling the box,
a0=4 / w0=5um / 16fs Gaussian laser pulse initialised analytically on the grid, z periodic
(Nz stays 4096: isolates the hot loop).  N>1: weak scaling over z-slabs (n_order=32, NCCL guard-cell
exchange + particle migration), C3 as named: 4096 physical cells + 2x64 guard cells = 4224 local cells per GPU
(`--compact-slab`: the guards inside a local box of 4096 cells); the particle count in `value` is the real one.
The timed region holds the K steps and nothing else; the per-kernel-family device times (`kernels`, `roofline`) come
from a second pass of K steps with CUDA events around every launch (`ms_per_step_instrumented`).
One "step" = one full PIC cycle (Simulation.step(1)); metric = particle-updates/s =
(sum over ranks of macro-particles) * K / (max over ranks of the device time of K steps).
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
# host-thread hygiene for the CPU legs, before NumPy/OpenBLAS load (oversubscription costs 25x)
os.environ.setdefault('OMP_WAIT_POLICY', 'passive')
os.environ.setdefault('OPENBLAS_NUM_THREADS', str(min(os.cpu_count() or 1, 32)))
import subprocess
import sys
import threading
import time

import numpy as np
from scipy.constants import c, e, m_e

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: Nz (per GPU), Nr, Nm, (p_nz, p_nr, p_nt), dz [m], rmax [m], n_e
    'C2': dict(Nz=4096, Nr=256, Nm=2, ppc=(2, 2, 4), dz=0.05e-6, rmax=20.e-6 * 256 / 50, n_e=4.e24),
    # C2 with cubic particle shapes (the 4 x 4 stencil kernels)
    'C2c': dict(Nz=4096, Nr=256, Nm=2, ppc=(2, 2, 4), dz=0.05e-6, rmax=20.e-6 * 256 / 50, n_e=4.e24, shape='cubic'),
    'C1': dict(Nz=256, Nr=64, Nm=2, ppc=(2, 2, 4), dz=0.2e-6, rmax=20.e-6, n_e=2.e24),
    'C4': dict(Nz=2048, Nr=512, Nm=4, ppc=(2, 2, 16), dz=0.05e-6, rmax=40.e-6, n_e=4.e24),
    # C4 with the minimum azimuthal sampling main.py:817-821 allows for the default of p_nt (transform-dominated)
    'C4t': dict(Nz=2048, Nr=512, Nm=4, ppc=(2, 2, 4), dz=0.05e-6, rmax=40.e-6, n_e=4.e24),
    # C2 as the script runs it: open z (damping + injection cells: 4416 local cells) and a moving window at c
    'C2w': dict(Nz=4096, Nr=256, Nm=2, ppc=(2, 2, 4), dz=0.05e-6, rmax=20.e-6 * 256 / 50, n_e=4.e24, window=True),
    # C5: boosted frame, Galilean PSATD, gamma = 10, electrons + ions, 8 per cell in theta; Nz per GPU (named on 4)
    'C5': dict(Nz=2048, Nr=256, Nm=2, ppc=(2, 2, 8), dz=0.05e-6 * 10, rmax=20.e-6 * 256 / 50, n_e=4.e24 * 10,
               gamma_boost=10., ions=True),
    'tiny': dict(Nz=128, Nr=32, Nm=2, ppc=(2, 2, 4), dz=0.1e-6, rmax=10.e-6, n_e=4.e24),
}


def laser_fields(z, r, a0=4., w0=5.e-6, ctau=16.e-15 * c, z0=None, lambda0=0.8e-6):
    """Linearly (x) polarised Gaussian pulse at focus, as mode-1 amplitudes
    (Er1, Et1, Br1, Bt1) on the (z, r) mesh: Ex = 2 Re[Er1 e^{-i theta}] cos(theta) ..."""
    k0 = 2 * np.pi / lambda0
    E0 = a0 * m_e * c**2 * k0 / e
    if z0 is None:
        z0 = 0.5 * (z[0] + z[-1])
    prof = E0 * np.exp(-(z[:, None] - z0)**2 / ctau**2) * np.exp(-r[None, :]**2 / w0**2) \
        * np.cos(k0 * (z[:, None] - z0))
    Er1 = 0.5 * prof + 0.j
    Et1 = -0.5j * prof
    Br1 = 0.5j * prof / c
    Bt1 = 0.5 * prof / c + 0.j
    return Er1, Et1, Br1, Bt1


def build_b200_sim(cfg, n_gpus, fused=True, seed=0, sort_period=1, full_slab=True):
    from fbpic_b200 import Simulation
    np.random.seed(seed + int(os.environ.get('RANK', '0')))
    dt = cfg['dz'] / c
    p_nz, p_nr, p_nt = cfg['ppc']
    n_order = -1 if n_gpus == 1 else 32
    n_guard = None
    nz_phys = cfg['Nz']
    if n_gpus > 1:
        # z-slabs: every rank works on a LOCAL periodic box = physical cells + 2*n_guard guard cells and FFTs
        # that length.  Guard width >= stencil reach of n_order=32 (boundary_communicator.py:243-250).
        # Default (`full_slab`): cfg['Nz'] PHYSICAL cells per GPU as BASELINE names C3 / C5, local length
        # cfg['Nz'] + 2*n_guard (4096 + 2*64 = 4224 = 64*66: the two-pass z-FFT of b2_fft.cu runs that length at
        # 17 us per array, cuFFT needs 36 us -- profiles/r02_fft_group.txt).
        # --compact-slab: the local box keeps the single-GPU length (cfg['Nz'] - 2*n_guard physical cells).
        from fbpic_b200.host_tables import stencil_reach
        n_guard = stencil_reach(cfg['Nz'] * n_gpus, cfg['dz'], cfg['dz'], n_order, None, False) + 1
        n_guard = (n_guard + 7) // 8 * 8
        if full_slab:
            # widen the guard (steps of 8 cells) until the local length has a two-pass FFT plan with small radices
            from fbpic_b200 import _lib as _l
            for extra in range(0, 129, 8):
                if _l.load().b2_fft_has_plan(cfg['Nz'] + 2 * (n_guard + extra), 13):
                    n_guard += extra
                    break
        else:
            nz_phys = cfg['Nz'] - 2 * n_guard
    Nz_g = nz_phys * n_gpus
    zmax = Nz_g * cfg['dz']
    kw = {}
    uz_m = 0.
    if cfg.get('gamma_boost'):
        gb = cfg['gamma_boost']
        kw.update(v_comoving=-c * np.sqrt(1. - 1. / gb**2), use_galilean=True, initialize_ions=cfg.get('ions', False))
        uz_m = -np.sqrt(gb**2 - 1.)
        if n_gpus == 1:
            n_order = 32
    window = bool(cfg.get('window')) and n_gpus == 1
    bz = 'open' if window else 'periodic'
    sim = Simulation(Nz_g, zmax, cfg['Nr'], cfg['rmax'], cfg['Nm'], dt, p_zmin=0., p_zmax=zmax,
                     p_rmin=0., p_rmax=cfg['rmax'], p_nz=p_nz, p_nr=p_nr, p_nt=p_nt, n_e=cfg['n_e'],
                     n_order=n_order, n_guard=n_guard, boundaries={'z': bz, 'r': 'reflective'},
                     fused=fused, sort_period=sort_period, particle_shape=cfg.get('shape', 'linear'), **kw)
    if uz_m != 0.:
        for sp in sim.ptcl:                 # the plasma flows backwards at gamma in the boosted frame (main.py:909-936)
            sp.uz = np.full(sp.Ntot, uz_m)
            sp.inv_gamma = np.full(sp.Ntot, 1. / np.sqrt(1. + uz_m**2))
    if window:
        sim.set_moving_window(v=c)
    g1 = sim.fld.interp[1]
    lam = 0.8e-6 * (cfg['dz'] / 0.05e-6)
    Er1, Et1, Br1, Bt1 = laser_fields(g1.z, g1.r, lambda0=lam,
                                      z0=0.5 * zmax if n_gpus == 1 else 0.5 * nz_phys * cfg['dz'])
    g1.Er[:, :], g1.Et[:, :], g1.Br[:, :], g1.Bt[:, :] = Er1, Et1, Br1, Bt1
    return sim


def build_oracle_sim(cfg, Nz, nthreads, seed=0):
    from oracle import oracle as orc
    np.random.seed(seed)
    zmax = Nz * cfg['dz']
    dt = cfg['dz'] / c
    p_nz, p_nr, p_nt = cfg['ppc']
    sim = orc.OracleSim(Nz, zmax, cfg['Nr'], cfg['rmax'], cfg['Nm'], dt, nthreads=nthreads)
    # particles exactly as Simulation.add_new_species would create them (last two r cells empty)
    Npz, Npr = Nz * p_nz, cfg['Nr'] * p_nr
    # (the port's own statement of the lattice of continuous_injection.py:203-275: cold, uniform density)
    ddz, ddr, dth = zmax / Npz, cfg['rmax'] / Npr, 2 * np.pi / p_nt
    zp, rp, tp = np.meshgrid(ddz * (np.arange(Npz) + 0.5), ddr * (np.arange(Npr) + 0.5), dth * np.arange(p_nt),
                             copy=True, indexing='ij')
    tp += (2 * np.pi * np.random.rand(Npz, Npr))[:, :, None]
    r, th, z = rp.ravel(), tp.ravel(), zp.ravel()
    x, y, w = r * np.cos(th), r * np.sin(th), cfg['n_e'] * r * dth * ddr * ddz
    Ntot = z.size
    ux, uy, uz, ig = np.zeros(Ntot), np.zeros(Ntot), np.zeros(Ntot), np.ones(Ntot)
    sim.add_species(-e, m_e, x, y, z, ux, uy, uz, ig, w)
    zz = (0.5 + np.arange(Nz)) * cfg['dz']
    rr = (0.5 + np.arange(cfg['Nr'])) * (cfg['rmax'] / cfg['Nr'])
    Er1, Et1, Br1, Bt1 = laser_fields(zz, rr, z0=0.5 * zmax)
    g1 = sim.interp[1]
    g1['Er'][:, :], g1['Et'][:, :], g1['Br'][:, :], g1['Bt'][:, :] = Er1, Et1, Br1, Bt1
    return sim, Ntot


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 5 ms (the same counters
    nvidia-smi prints as clocks.sm / clocks_event_reasons.*); falls back to polling nvidia-smi."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index=0, period=0.005):
        super().__init__(daemon=True)
        self.gpu, self.period, self._halt = gpu_index, period, threading.Event()
        self.sm, self.reasons, self.sm_max, self.source = [], set(), None, 'nvml'
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[gpu_index]) if vis and vis.split(',')[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.masks = {'hw_slowdown': pynvml.nvmlClocksThrottleReasonHwSlowdown,
                          'hw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                          'sw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                          'sw_power_cap': pynvml.nvmlClocksThrottleReasonSwPowerCap}
        except Exception:
            self.nv, self.source, self.period = None, 'nvidia-smi', 0.05

    def _sample(self):
        if self.nv is not None:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.reasons.update(n for n, m in self.masks.items() if r & m)
            return
        out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        if out:
            f = [v.strip() for v in out.split(',')]
            if f[0].replace('.', '').isdigit():
                self.sm.append(float(f[0]))
            if f[1].replace('.', '').isdigit():
                self.sm_max = max(self.sm_max or 0., float(f[1]))
            names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
            self.reasons.update(n for n, v in zip(names, f[3:7]) if v.lower().startswith('active'))

    def run(self):
        while not self._halt.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return dict(sm_mhz=(float(np.median(self.sm)) if self.sm else None),
                    sm_min_mhz=(min(self.sm) if self.sm else None), sm_max_mhz=self.sm_max,
                    reasons=sorted(self.reasons), samples=len(self.sm), source=self.source)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI
    function), before any host buffer is allocated: page-locked staging buffers and the H2D/D2H copies of the
    e2e leg then stay on the GPU's socket.  Returns a description (None if nothing was changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[device_index]) if vis and vis.split(',')[device_index].isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = '/sys/bus/pci/devices/%s/local_cpulist' % bus.lower()[-12:]
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return 'cpus %s (NUMA node of GPU %s)' % (txt, bus)
    except Exception:
        pass
    return None


def csrc_hash():
    """sha256 over the CUDA sources of the library (ties an ncu traffic figure to the build it was captured from)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'fbpic_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return None


def profile_table():
    from fbpic_b200 import _lib
    lib = _lib.load()
    out = {}
    for s in range(lib.b2_profile_slots()):
        ms, n = ctypes.c_double(0.), ctypes.c_uint64(0)
        _lib.call.b2_profile_read(s, ctypes.byref(ms), ctypes.byref(n))
        if n.value:
            out[lib.b2_profile_name(s).decode()] = dict(ms=ms.value, launches=int(n.value))
    return out


def time_oracle(cfg, Nz, steps, warmup, nthreads):
    sim, Ntot = build_oracle_sim(cfg, Nz, nthreads)
    sim.step(max(warmup, 1))
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    return Ntot * steps / dt, dt / steps * 1e3, Ntot


def cpu_model():
    try:
        for ln in open('/proc/cpuinfo'):
            if ln.startswith('model name'):
                return ln.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown CPU'


def time_reference(cfg, Nz, steps, warmup, threads, budget_s=0.):
    """FBPIC's own CPU path (oracle/_ref, installed by oracle/make_ref.sh) in a subprocess with the thread
    environment of SURVEY 8d; returns the dict printed by oracle/ref_bench.py."""
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'fbpic', 'main.py')):
        raise RuntimeError('oracle/_ref not installed (bash oracle/make_ref.sh)')
    job = {'cfg': {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}, 'Nz': Nz, 'steps': steps,
           'warmup': warmup, 'threads': threads, 'budget_s': budget_s}
    env = dict(os.environ)
    for k in ('OMP_WAIT_POLICY', 'OPENBLAS_NUM_THREADS'):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_bench.py'), json.dumps(job)],
                       capture_output=True, text=True, env=env, timeout=float(os.environ.get('REF_TIMEOUT_S', 900.)))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    if r.returncode != 0 or not lines:
        raise RuntimeError('reference run failed: ' + (r.stderr or r.stdout)[-300:])
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2', choices=sorted(CONFIGS))
    ap.add_argument('--preroll', type=int, default=30, help='untimed setup steps that disorder the plasma')
    ap.add_argument('--no-fused', action='store_true')
    ap.add_argument('--sort-period', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--compact-slab', action='store_true',
                    help='N>1: local box of Nz cells INCLUDING the guards (Nz - 2*n_guard physical cells per GPU) '
                         'instead of Nz physical cells per GPU plus guards')
    ap.add_argument('--no-parity', action='store_true', help='N>1: skip the untimed N-rank vs 1-rank parity check')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    n_gpus = max(args.gpus, world)
    metric = 'particle-updates/s (PIC hot loop: gather+push+deposit+sort+spectral solve)'
    workload = '%s periodic plasma + laser: Nz=%d/GPU Nr=%d Nm=%d ppc=%dx%dx%d %s' % (
        (args.config, cfg['Nz'], cfg['Nr'], cfg['Nm']) + cfg['ppc'] + (cfg.get('shape', 'linear'),))
    ncores = os.cpu_count() or 1
    nthreads = int(os.environ.get('ORACLE_NUM_THREADS', min(ncores, 64)))

    # ------------------------------------------------------------------ reference arm
    if args.impl == 'reference':
        if rank != 0:
            return
        # FBPIC's own numba CPU path (unmodified, oracle/_ref) on THIS workload at full size, all host threads
        # this process may use; the C/OpenMP oracle port is timed next to it as a second, labelled number.
        threads = int(os.environ.get('REF_NUM_THREADS', len(os.sched_getaffinity(0))))
        warm = max(1, min(args.warmup, 5))
        try:
            ref = time_reference(cfg, cfg['Nz'], args.steps, warm, threads,
                                 budget_s=float(os.environ.get('REF_BUDGET_S', 200.)))
            kind, val, ms, steps, Ntot = 'reference', ref['value'], ref['ms_per_step'], ref['steps'], ref['Ntot']
            sample = 'the whole workload (Nz=%d, %d particles), %d steps after %d warm-up steps: unmodified FBPIC %s ' \
                     'Simulation(use_cuda=False).step(), numba %s threading layer %s, NUMBA_NUM_THREADS=%d of %d ' \
                     'logical cores (%s), OPENBLAS_NUM_THREADS=1, FFT = scipy.fft through the pyfftw shim' % (
                         cfg['Nz'], Ntot, steps, warm, ref['fbpic'], ref['numba'], ref['threading_layer'],
                         threads, ncores, cpu_model())
            extra = {'reference_detail': ref}
        except Exception as exc:            # oracle/_ref missing or numba unusable on this box: the port, labelled
            from oracle import oracle as orc
            orc.build()
            Nz_s, steps = min(cfg['Nz'], 1024), max(1, min(args.steps, 10))
            val, ms, Ntot = time_oracle(cfg, Nz_s, steps, min(args.warmup, 2), min(threads, 32))
            kind = 'port'
            sample = 'z-slab Nz=%d of the workload (%d particles), %d steps, oracle port (C+OpenMP particle ' \
                     'kernels, scipy.fft, OpenBLAS dgemm), %d threads; FBPIC itself failed: %s' % (
                         Nz_s, Ntot, steps, min(threads, 32), repr(exc)[:200])
            extra = {}
        line = {
            'impl': 'reference', 'metric': metric, 'value': val, 'unit': 'particle-updates/s',
            'n_gpus': 0, 'steps': steps, 'warmup': warm, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': {'workload': workload, 'sample': sample},
            'cpu_baseline': {'value': val, 'unit': 'particle-updates/s', 'cores': threads, 'kind': kind,
                             'sample': sample, 'cpu': cpu_model()},
            'e2e': {'value': val, 'unit': 'particle-updates/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}
        line.update(extra)
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    # NCCL's own report of the communicator (rank count, transport; NCCL_DEBUG=INFO prints to stdout) must reach
    # stderr while stdout stays the one JSON line: file descriptor 1 is pointed at stderr for the whole run and the
    # JSON line is written to the saved descriptor at the end.
    json_fd = 1
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() not in ('INFO', 'TRACE'):
            os.environ['NCCL_DEBUG'] = os.environ.get('B2_NCCL_DEBUG', 'INFO')
        os.environ.setdefault('NCCL_DEBUG_SUBSYS', 'INIT')
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
    from fbpic_b200 import _lib
    from fbpic_b200._lib import call
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('gloo')
    ctx = _lib.context()
    affinity = bind_to_gpu_numa(ctx.device)
    sim = build_b200_sim(cfg, n_gpus, fused=not args.no_fused, sort_period=args.sort_period,
                         full_slab=not args.compact_slab)
    nz_local = sim.fld.interp[0].Nz
    if n_gpus > 1:
        workload += ' | z-slabs: %d physical + 2x%d guard = %d local cells per GPU, n_order=32' % (
            nz_local - 2 * sim.comm.n_guard, sim.comm.n_guard, nz_local)
    Ntot_local = sum(s.Ntot for s in sim.ptcl)


    def barrier():
        call.b2_device_sync()
        if dist is not None:
            dist.barrier()

    # setup (untimed): disorder the plasma, then W warm-up steps
    sim.step(max(args.preroll, 1), keep_on_gpu=True)
    sim.step(max(args.warmup, 3), keep_on_gpu=True)

    ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
    call.b2_event_create(ctypes.byref(ev0))
    call.b2_event_create(ctypes.byref(ev1))
    sampler = ClockSampler(ctx.device) if rank == 0 else None
    launches0 = _lib.load().b2_launch_count()
    flops0 = _lib.load().b2_dht_flops()
    barrier()
    if sampler:
        sampler.start()
    # the timed region: K steps, nothing else on the stream (the per-launch events of the kernel table cost
    # ~2 us of stream time per launch: they are recorded in a second pass of K steps below)
    call.b2_event_record(ev0, ctx.stream)
    sim.step(args.steps, keep_on_gpu=True)
    call.b2_event_record(ev1, ctx.stream)
    host_enqueue_ms = sim.last_enqueue_s * 1e3 / args.steps
    barrier()
    ms = ctypes.c_float(0.)
    call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms))
    clocks = sampler.stop() if sampler else None
    launches = _lib.load().b2_launch_count() - launches0
    dht_flops = _lib.load().b2_dht_flops() - flops0
    # second pass, same K steps, with CUDA events around every launch: per-kernel-family device time (roofline).
    # The second-stream overlap of the post-exchange z-FFTs is off in this pass: events around a launch that shares
    # the device with another stream would time the sharing, not the kernel.
    overlap_eb = bool(sim.fld.side_allowed()) and (n_gpus > 1 or bool(cfg.get('window')))
    sim.fld.join_side()
    sim.fld._side_ok = False
    call.b2_profile_reset()
    call.b2_profile_enable(1)
    call.b2_event_record(ev0, ctx.stream)
    sim.step(args.steps, keep_on_gpu=True)
    call.b2_event_record(ev1, ctx.stream)
    barrier()
    call.b2_profile_enable(0)
    sim.fld._side_ok = None          # (re-evaluated at the next use)
    ms_prof = ctypes.c_float(0.)
    call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms_prof))
    prof = profile_table()
    t_ms, n_tot = ms.value, float(Ntot_local)
    if dist is not None:
        import torch
        t = torch.tensor([t_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t[0])
        nt = torch.tensor([n_tot], dtype=torch.float64)
        dist.all_reduce(nt, op=dist.ReduceOp.SUM)
        n_tot = float(nt[0])
    value = n_tot * args.steps / (t_ms * 1e-3)

    # ---- e2e: the user-facing call with HOST buffers: Simulation.step(K) copies the whole state
    #      host->device at entry and device->host at exit (main.py:402-403, 579-581) ----
    e2e = None
    if not args.no_e2e:
        try:
            sim.receive_data_from_gpu()
            k_e2e = args.steps
            sim.step(1)                          # untimed: the first round trip allocates the page-locked buffers
            barrier()
            t0 = time.perf_counter()
            sim.step(k_e2e)                      # H2D of all state, K steps, D2H of all state
            call.b2_device_sync()
            t_e2e = time.perf_counter() - t0
            if dist is not None:
                import torch
                t = torch.tensor([t_e2e], dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_e2e = float(t[0])
            e2e = {'value': n_tot * k_e2e / t_e2e, 'unit': 'particle-updates/s', 'wall_s': t_e2e,
                   'split_s': {k: round(v, 4) for k, v in sim.last_step_timing.items()},
                   'h2d_bytes_per_step': sim.last_step_bytes['h2d'] / k_e2e,
                   'd2h_bytes_per_step': sim.last_step_bytes['d2h'] / k_e2e,
                   'note': 'Simulation.step(%d) from/to host NumPy arrays: particle + field state H2D at entry and D2H '
                           'at exit, as the reference API does (bytes counted from the arrays copied; the gathered '
                           'fields Ex..Bz of the particles stay in registers in the fused step and are not copied); '
                           'per-step bytes = total/%d' % (k_e2e, k_e2e)}
        except Exception as exc:      # the device-timed line above must survive a failure of this leg
            e2e = {'value': None, 'unit': 'particle-updates/s', 'error': repr(exc)[:300]}

    # ---- N > 1: the sharded loop against the same problem on one GPU (untimed; small periodic box, all ranks)
    mgpu_parity = None
    if dist is not None and not args.no_parity:
        try:
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            from mgpu_parity import periodic_case
            cases = []
            for correct, tol in ((False, 1e-9), (True, 5e-4)):
                ok, err, counts = periodic_case(dist, correct, tol, nsteps=24, nzr=96, verbose=False)
                cases.append({'correct_currents': correct, 'ok': ok, 'max_err_rel_to_field_max': err, 'tol': tol,
                              'particles_per_rank': counts})
            mgpu_parity = {'ranks': world, 'ok': all(cse['ok'] for cse in cases), 'cases': cases,
                           'what': 'E, B, J, rho of %d z-slabs (96 physical cells each, NCCL halo + particle migration, '
                                   '24 steps) vs the same box on one GPU (tools/mgpu_parity.py)' % world}
        except Exception as exc:
            mgpu_parity = {'ranks': world, 'ok': False, 'error': repr(exc)[:300]}

    if rank != 0:
        return
    # ---- roofline of the dominant kernel family (device time from CUDA events inside the timed region)
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get('hbm_gbs', 6650.)
    peak_src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback'
    cells = nz_local * cfg['Nr'] * 16
    Nm = cfg['Nm']
    exchange_path = (n_gpus > 1) or bool(cfg.get('window'))      # guard exchange / damping between iFFT and FFT
    # algorithmic bytes PER STEP of each kernel family, whatever the number of launches:
    alg = {
        # J: + (4+64+64-64) B/particle on the steps where the SoA permutation rides along;
        # rho: the second position push is fused in (reads 64, writes 24 B/particle)
        'deposit_J': (64 + 68 / max(args.sort_period, 1)) * Ntot_local + 3 * Nm * cells,
        'deposit_rho': 88 * Ntot_local + Nm * cells,
        'gather_push': 112 * Ntot_local + 6 * Nm * cells,
        'sort': (4 + 12 + 4) * Ntot_local / max(args.sort_period, 1),
        # z-FFTs: J (3) + rho (1) forward and E, B (6) inverse per mode; the exchange path adds the
        # iFFT -> exchange -> FFT round trips of J (3 + 3) and of E, B (6 more forward)
        'fft': (22 if exchange_path else 10) * Nm * 2 * cells,
        # fused correct + push (+ push_rho): reads 11, writes 8 arrays, coefficients 2.5 arrays-worth per mode;
        # as separate correct / push / push_rho launches: 27 arrays
        'spectral': Nm * ((27 if n_gpus > 1 else 19) * cells + 5 * cells // 2),
    }
    top = max(prof.items(), key=lambda kv: kv[1]['ms'])[0] if prof else None
    roofline = None
    shares = {k: v['ms'] / t_ms for k, v in prof.items()}
    if top == 'dht':
        # flops counted by the library for the launches of the timed region (4*Nz*Nr^2 per transformed array)
        ach = dht_flops / (prof['dht']['ms'] * 1e-3) / 1e12
        roofline = {'kernel': 'k_dht_tma (Hankel GEMM, TMA-fed fp64 DMMA)', 'bound': 'tensor', 'achieved': ach,
                    'peak': 37.1, 'unit': 'TFLOP/s', 'frac': ach / 37.1, 'traffic': None,
                    'peak_source': 'measured DMMA.8x8x4 issue peak on B200 (profiles/r01_microbench.txt, '
                                   'profiles/r02_microbench_dmma_operands.txt); MEASURED_PEAKS.json has no fp64 entry'}
    elif top in alg:
        ach = alg[top] * args.steps / (prof[top]['ms'] * 1e-3) / 1e9
        roofline = {'kernel': top, 'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': ach / hbm_peak, 'traffic': None, 'peak_source': peak_src}
    # DRAM traffic per launch of the dominant kernel: from the ncu --set full capture of THIS build
    # (profiles/r02_traffic.json records the hash of csrc/ it was taken from; null if the sources changed since)
    if roofline is not None:
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        except Exception:
            traffic = {}
        ent = (traffic.get('kernels') or {}).get('%s:%s' % (args.config, top))
        if ent and n_gpus == 1 and traffic.get('csrc_sha256') == csrc_hash():
            roofline['traffic'] = ent['dram_bytes_per_launch']
            roofline['traffic_source'] = ent['source']
        else:
            roofline['traffic_source'] = 'no ncu --set full capture of this kernel for this build and config'
    kernels = {}
    for k, v in prof.items():
        d = dict(ms_per_step=v['ms'] / args.steps, launches_per_step=v['launches'] / args.steps,
                 share=shares[k])
        if k in alg:
            d['GBps'] = alg[k] * args.steps / (v['ms'] * 1e-3) / 1e9
            d['hbm_frac'] = d['GBps'] / hbm_peak
        if k == 'dht':
            d['TFLOPs'] = dht_flops / (v['ms'] * 1e-3) / 1e12
            d['tensor_frac'] = d['TFLOPs'] / 37.1
        kernels[k] = d

    cpu_baseline = None
    if not args.no_cpu_baseline and n_gpus == 1:
        # bounded sample on this box's host cores: FBPIC's own numba path (oracle/_ref) on a z-slab of the
        # workload; the C/OpenMP oracle port only if the reference cannot run here (labelled)
        threads = len(os.sched_getaffinity(0))
        Nz_s = min(cfg['Nz'], 1024)
        try:
            ref = time_reference(cfg, Nz_s, 4, 1, threads, budget_s=25.)
            cpu_baseline = {'value': ref['value'], 'unit': 'particle-updates/s', 'cores': threads,
                            'kind': 'reference', 'cpu': cpu_model(),
                            'sample': 'z-slab Nz=%d of the workload (%d particles), %d steps after 1 warm-up: unmodified '
                                      'FBPIC %s numba CPU path, %d threads (`--impl reference` times the whole workload)'
                                      % (Nz_s, ref['Ntot'], ref['steps'], ref['fbpic'], threads)}
        except Exception as exc:
            from oracle import oracle as orc
            orc.build()
            Nz_s = min(cfg['Nz'], 512)
            val, ms_cpu, n_cpu = time_oracle(cfg, Nz_s, 2, 1, min(threads, 32))
            cpu_baseline = {'value': val, 'unit': 'particle-updates/s', 'cores': min(threads, 32), 'kind': 'port',
                            'sample': 'z-slab Nz=%d of the workload (%d particles), 2 steps after 1 warm-up, oracle '
                                      'port; FBPIC itself failed: %s' % (Nz_s, n_cpu, repr(exc)[:160])}
    out = {
        'metric': metric, 'value': value, 'unit': 'particle-updates/s', 'n_gpus': n_gpus,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': t_ms / args.steps,
        'pic_steps_per_s': args.steps / (t_ms * 1e-3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload, 'particles_total': n_tot, 'host_affinity': affinity, 'fused': not args.no_fused,
                   'n_order': -1 if n_gpus == 1 else 32, 'n_guard': sim.comm.n_guard, 'preroll_steps': args.preroll, 'sort_period': args.sort_period,
                   'l2': 'inputs larger than L2 (particle state %.1f GB per GPU)' % (Ntot_local * 64 / 1e9)},
        'gpu_launches': int(launches), 'host_enqueue_ms_per_step': round(host_enqueue_ms, 4),
        'ms_per_step_instrumented': round(ms_prof.value / args.steps, 4), 'overlap_eb_ffts': overlap_eb, 'clocks': clocks, 'e2e': e2e, 'roofline': roofline,
        'cpu_baseline': cpu_baseline, 'kernels': kernels,
    }
    if mgpu_parity is not None:
        out['mgpu_parity'] = mgpu_parity
    if json_fd == 1:
        print(json.dumps(out))
    else:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + '\n').encode())


if __name__ == '__main__':
    main()
