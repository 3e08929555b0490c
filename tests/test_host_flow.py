"""CPU tests of the HOST-SIDE logic of the operator surface.  `Simulation.step` and friends are run with
tests/fake_device.py standing in for libfbpic_b200.so (device memory in host RAM, every C-ABI entry point
executed by the oracle / NumPy / the host-compiled kernel source of tests/hostemu) and compared with the
golden outputs of the unmodified reference.  What this checks: call order, sort-state bookkeeping, PML /
cross-deposition / antenna flows, particle exchange and injection on one rank.  What it does NOT check:
the CUDA kernels (the `-m gpu` tests do, through the real library).  The test bodies are the GPU tests
themselves, imported from their modules."""
import gc
import pytest

import fake_device
import test_gpu_step
import test_gpu_widen
import test_gpu_laser
import test_gpu_external
import test_gpu_bunch
import test_gpu_scripts


@pytest.fixture
def fake(monkeypatch):
    from fbpic_b200 import _lib
    f = fake_device.install(monkeypatch)
    yield f
    gc.collect()                    # device arrays of the test are released through the fake
    _lib.call.__dict__.clear()


@pytest.mark.parametrize('fused', [False, True, 3])
@pytest.mark.parametrize('tag', test_gpu_step.TAGS)
def test_step_flow(fake, tag, fused):
    test_gpu_step.test_step_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
def test_moving_window_flow(fake, fused):
    test_gpu_step.test_moving_window_vs_reference_golden(fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['periodic', 'open', 'galilean', 'window'])
def test_pml_flow(fake, tag, fused):
    test_gpu_widen.test_pml_step_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['std', 'galilean'])
def test_cross_deposition_flow(fake, tag, fused):
    test_gpu_widen.test_cross_deposition_step_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('tag', ['gauss', 'lg_pml', 'boost'])
def test_add_laser_direct_flow(fake, tag):
    test_gpu_laser.test_add_laser_direct_vs_reference_golden(tag)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'moving', 'boost'])
def test_laser_antenna_flow(fake, tag, fused):
    test_gpu_laser.test_laser_antenna_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'boost'])
def test_external_fields_flow(fake, tag, fused):
    test_gpu_external.test_external_fields_step_vs_reference_golden(tag, fused)


def test_external_field_string_flow(fake):
    test_gpu_external.test_external_field_string_expression()


@pytest.mark.parametrize('tag', ['uniform', 'gaussian', 'gaussian_boost'])
def test_bunch_space_charge_flow(fake, tag):
    test_gpu_bunch.test_bunch_space_charge_vs_reference_golden(tag)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lwfa', 'boosted'])
def test_example_script_flow(fake, tag, fused):
    test_gpu_scripts.test_example_script_vs_reference_golden(tag, fused)
