"""CPU tests of the Python -> CUDA C translation behind `ExternalField`: the generated source compiles under
NVRTC for sm_100a (inside libfbpic_b200.so, no GPU needed for that step) and -- compiled for the host by the
test -- evaluates to the same numbers as the Python function."""
import ctypes
import math
import numpy as np
import pytest

from conftest import assert_close
import fake_device


def f_undulator(F, x, y, z, t, amplitude, length_scale):
    return F + amplitude * math.cos(2 * np.pi * z / length_scale)


def f_branches(field, xx, yy, zz, time, amp, L):
    """local variables, renamed arguments, if / elif / else, conditional expression, power, modulo"""
    r2 = xx**2 + yy**2
    phase = (zz - 3.e8 * time) / L
    if zz < 0:
        return field
    elif r2 > L**2:
        env = 0.
    else:
        env = math.exp(-r2 / L**2) ** 1.5
    env *= 2.
    s = 1. if phase % 2. < 1. else -1.
    return field * 0.5 + amp * env * s * abs(math.sin(phase)) + math.atan2(yy, xx) * 1.e-3 * amp


K_CAPTURED = 2.5


def f_captured(F, x, y, z, t, amplitude, length_scale):
    return amplitude * K_CAPTURED * math.tanh(x / length_scale) - np.sqrt(y * y + 1.e-12) / length_scale


f_lambda = lambda F, x, y, z, t, a, L: F + a * math.hypot(x, y) / L      # noqa: E731


@pytest.mark.parametrize('func', [f_undulator, f_branches, f_captured, f_lambda])
def test_translation_matches_python(func):
    from fbpic_b200.lpa_utils.external_fields import python_to_cuda
    from fbpic_b200 import _lib
    body = python_to_cuda(func)
    # 1. NVRTC accepts it for sm_100a
    h, nb = ctypes.c_void_p(), ctypes.c_size_t()
    lib = _lib.load()
    rc = lib.b2_external_field_compile(body.encode(), ctypes.byref(h))
    assert rc == 0, lib.b2_error_string()
    lib.b2_external_field_cubin_size(h, ctypes.byref(nb))
    assert nb.value > 1000
    lib.b2_external_field_free(h)
    # 2. same numbers as the Python function (host build of the same body)
    rng = np.random.default_rng(8)
    n = 500
    F, x, y = rng.normal(size=n), rng.normal(size=n) * 4.e-6, rng.normal(size=n) * 4.e-6
    z = rng.uniform(-5.e-6, 20.e-6, n)
    t, amp, L = 7.e-15, 3.7, 5.e-6
    want = np.array([func(F[i], x[i], y[i], z[i], t, amp, L) for i in range(n)])
    fake = fake_device.FakeLib()
    hh = ctypes.c_void_p()
    fake.b2_external_field_compile(body.encode(), ctypes.byref(hh))
    got = F.copy()
    fake.b2_external_field_apply(None, hh, n, got.ctypes.data, x.ctypes.data, y.ctypes.data, z.ctypes.data, t, amp, L,
                                 1., 0., None)
    assert_close(got, want, 1e-14, func.__name__)


def test_untranslatable_function_is_rejected():
    from fbpic_b200.lpa_utils.external_fields import ExternalField, TranslationError

    def uses_a_loop(F, x, y, z, t, amplitude, length_scale):
        for k in range(3):
            F = F + k
        return F

    with pytest.raises(TranslationError):
        ExternalField(uses_a_loop, 'Ex', 1., 1.)
    with pytest.raises(ValueError):
        ExternalField(f_undulator, 'Er', 1., 1.)


def test_nvrtc_reports_bad_expression():
    from fbpic_b200 import _lib
    h = ctypes.c_void_p()
    lib = _lib.load()
    assert lib.b2_external_field_compile(b'    F_[i_] = F + no_such_symbol;', ctypes.byref(h)) != 0
    assert b'no_such_symbol' in lib.b2_error_string()


def test_boosted_frame_amplitudes():
    """external_fields.py:149-181"""
    from fbpic_b200.lpa_utils.external_fields import ExternalField
    from scipy.constants import c
    g = 5.
    b = math.sqrt(1 - 1 / g**2)
    e = ExternalField(f_undulator, 'By', 2., 1.e-2, gamma_boost=g)
    (f1, a1), (f2, a2) = e.fieldtypes_and_amplitudes
    assert (f1, f2) == ('By', 'Ex') and abs(a1 - g * 2.) < 1e-14 and abs(a2 + g * b * c * 2.) < 1e-6
    e = ExternalField(f_undulator, 'Ez', 2., 1.e-2, gamma_boost=g)
    assert e.fieldtypes_and_amplitudes == (('Ez', 2.),)


def test_module_attribute_constants_are_resolved_on_the_captured_object():
    """`const.e` of `import scipy.constants as const` is the elementary charge, `math.e` Euler's number: an attribute
    is evaluated on the module the function captured, never matched by its bare name."""
    import math
    import scipy.constants as const
    from fbpic_b200.lpa_utils.external_fields import python_to_cuda

    def field(F, x, y, z, t, amplitude, length_scale):
        return F + amplitude * const.e * math.cos(2 * math.pi * z / length_scale) + math.e

    src = python_to_cuda(field)
    assert '1.602176634e-19' in src and '2.718281828459045' in src
