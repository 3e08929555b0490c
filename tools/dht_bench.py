"""Times the batched Hankel launches of one PIC step (rho, J, E+B job lists) at the C2 and C4 grid sizes,
for the TMA-fed kernel and for the LDG-staged one (B2_DHT_IMPL=legacy), CUDA events on the context stream.
    python tools/dht_bench.py            -> one JSON line per (impl, shape, job list)"""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one():
    import numpy as np
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray, DhtJob, call
    ctx = _lib.context()
    impl = os.environ.get('B2_DHT_IMPL', 'tma')
    rng = np.random.default_rng(0)
    data = os.environ.get('DHT_BENCH_DATA', 'random')
    if data == 'zeros':                 # same launches on all-zero operands: separates data-dependent power / clock effects
        class _Z(object):
            def normal(self, size): return np.zeros(size)
        rng = _Z()
    from bench import ClockSampler
    for name, Nz, Nr, Nm in (('C2', 4096, 256, 2), ('C4', 2048, 512, 4), ('C1', 256, 64, 2)):
        mats = [DeviceArray.from_numpy(rng.normal(size=(Nr, Nr))) for _ in range(3 * Nm)]
        arrs = [DeviceArray.from_numpy(rng.normal(size=(Nz, Nr)) + 1j * rng.normal(size=(Nz, Nr)))
                for _ in range(6 * Nm)]
        outs = [DeviceArray((Nz, Nr), np.complex128) for _ in range(6 * Nm)]
        lists = {'rho': [], 'J': [], 'EB': [], 'J+rho': []}
        for m in range(Nm):
            M0, Mp, Mm = mats[3 * m].ptr, mats[3 * m + 1].ptr, mats[3 * m + 2].ptr
            a, o = arrs[6 * m:6 * m + 6], outs[6 * m:6 * m + 6]
            lists['rho'].append(DhtJob(a[0].ptr, None, o[0].ptr, None, M0, None, None, _lib.DHT_SCALAR))
            lists['J'] += [DhtJob(a[0].ptr, None, o[0].ptr, None, M0, None, None, _lib.DHT_SCALAR),
                           DhtJob(a[1].ptr, a[2].ptr, o[1].ptr, o[2].ptr, Mp, Mm, None, _lib.DHT_RT_TO_PM)]
            for f in (0, 3):
                lists['EB'] += [DhtJob(a[f].ptr, None, o[f].ptr, None, M0, None, None, _lib.DHT_SCALAR),
                                DhtJob(a[f + 1].ptr, a[f + 2].ptr, o[f + 1].ptr, o[f + 2].ptr, Mp, Mm, None,
                                       _lib.DHT_PM_TO_RT)]
        lists['J+rho'] = lists['J'] + lists['rho']
        ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
        call.b2_event_create(ctypes.byref(ev0))
        call.b2_event_create(ctypes.byref(ev1))
        step_ms, step_flop = 0., 0.
        for lname, jobs in lists.items():
            chunks = [jobs[i:i + 16] for i in range(0, len(jobs), 16)]     # <= 16 jobs per launch
            def go():
                for ch in chunks:
                    arr = (DhtJob * len(ch))(*ch)
                    call.b2_dht_batch(ctx.handle, len(ch), arr, Nz, Nr, None)
            for _ in range(3):
                go()
            reps = int(os.environ.get('DHT_BENCH_REPS', '100'))
            sampler = ClockSampler(ctx.device)
            sampler.start()
            call.b2_event_record(ev0, ctx.stream)
            for _ in range(reps):
                go()
            call.b2_event_record(ev1, ctx.stream)
            ms = ctypes.c_float(0.)
            call.b2_event_elapsed_ms(ev0, ev1, ctypes.byref(ms))
            clk = sampler.stop()
            nprod = sum(1 if j.kind == _lib.DHT_SCALAR else 2 for j in jobs)
            flop = nprod * 4. * Nz * Nr * Nr
            t = ms.value / reps
            if lname in ('J+rho', 'EB'):          # the two Hankel launches of a fused PIC step
                step_ms += t
                step_flop += flop
            print(json.dumps({'impl': impl, 'shape': name, 'Nz': Nz, 'Nr': Nr, 'jobs': lname, 'products': nprod,
                              'ms': round(t, 4), 'TFLOPs': round(flop / (t * 1e-3) / 1e12, 2), 'data': data,
                              'sm_mhz': clk['sm_mhz'], 'sm_min_mhz': clk['sm_min_mhz'], 'reasons': clk['reasons']}), flush=True)
        print(json.dumps({'impl': impl, 'shape': name, 'jobs': 'whole step', 'ms': round(step_ms, 4),
                          'TFLOPs': round(step_flop / (step_ms * 1e-3) / 1e12, 2)}), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--one':
        run_one()
    else:
        for impl in ('tma', 'legacy'):
            env = dict(os.environ, B2_DHT_IMPL=impl)
            subprocess.run([sys.executable, os.path.abspath(__file__), '--one'], env=env, timeout=600)
