"""Console output of a run (the role of fbpic/utils/printing.py): the set-up banner of `Simulation(verbose_level=...)`
and the progress line of `step(show_progress=True)`.

The progress line never synchronises the device: it reports the HOST time per cycle, i.e. the pace at which cycles are
enqueued, which equals the device pace once the launch queue is full; it is refreshed at most twice a second.  The
total printed at the end of `step()` is taken after the data came back from the device, so it is a true wall time."""
import sys
import time

from . import _lib


class ProgressBar(object):
    """printing.py:22-137: `|#####     | 120/400, 0:00:03 left, 3 ms/step`"""

    def __init__(self, N, n_avg=20, Nbars=35, char=u'|'):
        self.N, self.n_avg, self.Nbars, self.bar_char = N, n_avg, Nbars, char
        self.start_time = self.prev_time = self.last_print = time.time()
        self.avg_time_per_step, self.time_per_step, self.total_duration, self.eta, self.i_step = 0., 0., 0., None, 0

    def time(self, i_step):
        now = time.time()
        self.i_step = i_step
        self.total_duration = now - self.start_time
        self.time_per_step = now - self.prev_time
        if i_step <= 2:                    # the first cycles carry the uploads and one-off table set-up
            self.avg_time_per_step = self.time_per_step
        else:
            self.avg_time_per_step += (self.time_per_step - self.avg_time_per_step) / self.n_avg
        self.eta = None if i_step < self.n_avg else self.avg_time_per_step * (self.N - i_step)
        self.prev_time = now

    def print_progress(self):
        now = time.time()
        if now - self.last_print < 0.5 and self.i_step + 1 < self.N:
            return
        self.last_print = now
        i = self.i_step
        nbars = int((i + 1) * 1. / self.N * self.Nbars)
        line = '\r|' + nbars * self.bar_char + (self.Nbars - nbars) * ' ' + '| %d/%d' % (i + 1, self.N)
        if self.eta is None:
            line += ', calc. ETA...'
        else:
            m, s = divmod(self.eta, 60)
            h, m = divmod(m, 60)
            line += ', %d:%02d:%02d left' % (h, m, s)
        line += ', %.1f ms/step' % (self.avg_time_per_step * 1.e3)
        sys.stdout.write(line + '\033[K')
        sys.stdout.flush()

    def print_summary(self):
        total = time.time() - self.start_time
        m, s = divmod(total, 60)
        h, m = divmod(m, 60)
        print('\nTotal time taken (with data transfers): %d:%02d:%02d' % (h, m, s))
        print('Average time per iteration (with data transfers): %.1f ms\n' % (total / max(self.N, 1) * 1.e3))


def print_simulation_setup(sim, verbose_level=1):
    """Banner printed by rank 0 (printing.py:139-260): 1 = one line on the hardware, 2 = solver and domain details."""
    if verbose_level <= 0 or sim.comm.rank != 0:
        return
    from .diags import __version__
    lines = ['', 'fbpic_b200 (%s)' % __version__, '']
    n = sim.comm.size
    where = 'Running on %d B200 GPU%s (one process per GPU, NCCL)' % (n, 's' if n > 1 else '')
    if verbose_level == 1:
        lines.append(where)
    else:
        comm, g0 = sim.comm, sim.fld.interp[0]
        lines.append('Library: %s' % _lib.load().b2_version().decode())
        lines += [where,
                  'Grid (local, with guard / damp cells): Nz = %d, Nr = %d, %d azimuthal modes' % (g0.Nz, g0.Nr, sim.fld.Nm),
                  'PSATD stencil order: %s' % ('infinite' if sim.fld.n_order == -1 else '%d' % sim.fld.n_order),
                  'Current correction: %s' % sim.fld.current_correction,
                  'Particle shape: %s' % sim.particle_shape,
                  'Longitudinal boundaries: %s' % comm.boundaries['z'],
                  'Transverse boundaries: %s' % comm.boundaries['r'],
                  'Guard region size: %d cells' % comm.n_guard,
                  'Damping region size: %d cells' % comm.nz_damp,
                  'Injection region size: %d cells' % comm.n_inject,
                  'Particle exchange period: every %d step' % comm.exchange_period,
                  'Fused kernels: %s (sort period %d)' % ('yes' if sim.fused else 'no', sim.sort_period)]
        if getattr(sim, 'boost', None) is not None:
            lines += ['Boosted frame: Yes', 'Boosted frame gamma: %d' % sim.boost.gamma0,
                      'Galilean frame: %s' % ('Yes' if sim.use_galilean else 'No')]
        else:
            lines.append('Boosted frame: False')
    print('\n'.join(lines) + '\n')
