import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    """A GPU test that hangs would take the whole box with it: every `-m gpu` test gets a time limit when the
    pytest-timeout plugin is there (the slowest one takes well under a minute on a B200)."""
    if not config.pluginmanager.hasplugin('timeout'):
        return
    import pytest
    for item in items:
        if item.get_closest_marker('gpu') is not None and item.get_closest_marker('timeout') is None:
            item.add_marker(pytest.mark.timeout(900, method='thread'))     # a hang inside a CUDA call ignores signals


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def group_scale(golden, prefix, group, Nm):
    """max |F| over all components and modes of one vector field (E, B, J) or rho:
    components that vanish by symmetry hold only rounding noise, so parity is
    measured against the magnitude of the whole field."""
    names = ['rho'] if group == 'rho' else [group + 'r', group + 't', group + 'z']
    return max(np.abs(golden['%s%s_m%d' % (prefix, n, m)]).max() for n in names for m in range(Nm))


def assert_close(a, b, rel, what='', scale=None):
    """|a-b| <= rel * (max|a| + max|b|): the tolerance form of the reference's
    own CPU/GPU parity test (tests/test_cpu_gpu_deposition.py:96-98)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if scale is None:
        scale = np.abs(a).max() + np.abs(b).max()
    err = np.abs(a - b).max() if a.size else 0.
    assert err <= rel * scale + 1e-300, '%s: max err %.3e > %.1e * %.3e' % (what, err, rel, scale)


# Test files are collected in alphabetical order and the driver runs `pytest -x`: the hardware-proven GPU tests
# (test_gpu_kernels / multi / plasma_wave / step) therefore come first, the tests of the SURVEY 8f widening
# (test_gpu_w0 .. w9c, ordered from kernel-level goldens to whole-script and analytic acceptance runs, then the
# full-size property tests test_gpu_x_*) after them; the external-field tests, whose kernels are compiled at run time
# by NVRTC and loaded through the CUDA library API, come last (test_gpu_y_external, test_gpu_y2_ionization_laser).
