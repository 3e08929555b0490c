"""CPU tests of the HOST-SIDE logic of the operator surface.  `Simulation.step` and friends are run with
tests/fake_device.py standing in for libfbpic_b200.so (device memory in host RAM, every C-ABI entry point
executed by the oracle / NumPy / the host-compiled kernel source of tests/hostemu) and compared with the
golden outputs of the unmodified reference.  What this checks: call order, sort-state bookkeeping, PML /
cross-deposition / antenna flows, particle exchange and injection on one rank.  What it does NOT check:
the CUDA kernels (the `-m gpu` tests do, through the real library).  The test bodies are the GPU tests
themselves, imported from their modules."""
import gc
import os
import subprocess
import sys
import pytest

from conftest import ROOT

import fake_device

# flows that take minutes on the CPU stand-in: run once with B2_SLOW_FLOWS=1 (results recorded in DESIGN.md section 9)
slow_flow = pytest.mark.skipif(os.environ.get('B2_SLOW_FLOWS') != '1',
                               reason='minutes on the CPU stand-in; set B2_SLOW_FLOWS=1')
import test_gpu_step
import test_gpu_w0_ext_kernels
import test_gpu_w1_pml_cross
import test_gpu_w2_laser
import test_gpu_y_external
import test_gpu_w3_bunch
import test_gpu_w4_scripts
import test_gpu_w6_acceptance
import test_gpu_w8_diags
import test_gpu_w9_step_options
import test_gpu_w9b_ionization
import test_gpu_y2_ionization_laser
import test_gpu_w9c_compton


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
    test_gpu_w1_pml_cross.test_pml_step_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['std', 'galilean'])
def test_cross_deposition_flow(fake, tag, fused):
    test_gpu_w1_pml_cross.test_cross_deposition_step_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('tag', ['gauss', 'lg_pml', 'boost'])
def test_add_laser_direct_flow(fake, tag):
    test_gpu_w2_laser.test_add_laser_direct_vs_reference_golden(tag)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'moving', 'boost'])
def test_laser_antenna_flow(fake, tag, fused):
    test_gpu_w2_laser.test_laser_antenna_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'boost'])
def test_external_fields_flow(fake, tag, fused):
    test_gpu_y_external.test_external_fields_step_vs_reference_golden(tag, fused)


def test_external_field_string_flow(fake):
    test_gpu_y_external.test_external_field_string_expression()


@pytest.mark.parametrize('tag', ['uniform', 'gaussian', 'gaussian_boost'])
def test_bunch_space_charge_flow(fake, tag):
    test_gpu_w3_bunch.test_bunch_space_charge_vs_reference_golden(tag)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lwfa', 'boosted'])
def test_example_script_flow(fake, tag, fused):
    test_gpu_w4_scripts.test_example_script_vs_reference_golden(tag, fused)


def test_two_rank_flow_gloo():
    """The 2-rank parity worker of tests/test_gpu_multi.py (periodic slabs with and without the current
    correction, open z + moving window + injection + migration) on the CPU: fake device + gloo."""
    env = dict(os.environ, OMP_NUM_THREADS='2', ORACLE_NUM_THREADS='2', MGPU_NZ_PER_RANK='64', MGPU_STEPS='10',
               MGPU_WINDOW_STEPS='24')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29647',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py'), '--fake-device']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'MGPU_PARITY_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize('Nm', [2, 4])
def test_two_rank_pml_antenna_flow_gloo(Nm):
    """2 ranks, open z + radial PML + laser antenna + moving window with injection vs the single domain
    (Nm = 4: the E/B + PML guard exchange exceeds one staging launch and is chunked)."""
    env = dict(os.environ, OMP_NUM_THREADS='2', ORACLE_NUM_THREADS='2', MGPU_NZ_PER_RANK='64',
               MGPU_WINDOW_STEPS='24', MGPU_EXTRA='1', MGPU_NM=str(Nm))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(29649 + Nm),
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py'), '--fake-device']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'MGPU_EXTRA_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


# ---- the reference's physics acceptance tests (restated in test_gpu_w6_acceptance.py), host flow on the CPU
@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_uniform_rho_flow(fake, shape):
    test_gpu_w6_acceptance.test_uniform_electron_plasma(shape)
    test_gpu_w6_acceptance.test_neutral_plasma_shifted(shape)


@slow_flow
def test_cherenkov_instability_flow(fake):
    test_gpu_w6_acceptance.test_cherenkov_instability()


@pytest.mark.parametrize('case', [pytest.param('labframe_with_preexisting_plasma', marks=slow_flow),
                                  'boosted_with_preexisting_plasma',
                                  pytest.param('labframe_without_preexisting_plasma', marks=slow_flow)])
def test_continuous_injection_flow(fake, case):
    getattr(test_gpu_w6_acceptance, 'test_' + case)()


@pytest.mark.parametrize('variant', ['periodic', 'moving_window', 'galilean'])
def test_laser_propagation_flow(fake, variant):
    """mode 1 (Gaussian beam) of each variant; the other modes run in the GPU suite"""
    getattr(test_gpu_w6_acceptance, 'test_laser_' + variant)(1)


@pytest.mark.parametrize('fused', [False, True])
def test_diagnostics_flow(fake, fused, tmp_path):
    test_gpu_w8_diags.test_diagnostics_hold_the_state_of_their_iteration(fused, tmp_path)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('window', [False, True])
def test_restart_flow(fake, fused, window, tmp_path):
    test_gpu_w8_diags.test_restart_from_checkpoint_continues_the_run(fused, window, tmp_path)


def test_ext_kernel_tests_flow(fake):
    """the device-side kernel tests of b2_ext.cu, run here against the host-compiled kernel source"""
    k = test_gpu_w0_ext_kernels
    for comoving in (False, True):
        k.test_push_eb_pml(comoving, (9, 130))
        k.test_correct_currents_cross(comoving)
    k.test_damp_pml((33, 70), 33)
    k.test_correct_divE()
    k.test_antenna_helpers()
    k.test_push_p_after_plane()
    k.test_extract_slice(3)
    k.test_select_crossing()
    test_gpu_y_external.test_external_field_jit()


def test_two_rank_ionization_flow_gloo():
    """Ionization levels migrate with the ions between 2 ranks; electrons = events over all ranks."""
    env = dict(os.environ, OMP_NUM_THREADS='2', ORACLE_NUM_THREADS='2', MGPU_NZ_PER_RANK='64', MGPU_EXTRA='4')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29655',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py'), '--fake-device']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'MGPU_IONIZATION_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_rank_restart_flow_gloo():
    """Per-rank checkpoints of a 2-rank run, restart, same continuation (rank by rank)."""
    env = dict(os.environ, OMP_NUM_THREADS='2', ORACLE_NUM_THREADS='2', MGPU_NZ_PER_RANK='64', MGPU_EXTRA='3')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29657',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py'), '--fake-device']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'MGPU_RESTART_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize('fused', [False, True])
def test_transfer_accounting(fake, fused):
    """Bytes copied by a step() call: the fused step neither uploads nor reads back the 6 gathered-field arrays
    of the particles (they stay in registers); the unfused one moves them and returns the gathered fields.  Of the
    grids only E and B of the interpolation grid are uploaded (the sources and every spectral array are recomputed
    before their first use); everything is read back."""
    import numpy as np
    from scipy.constants import c
    from fbpic_b200 import Simulation
    np.random.seed(0)
    zmax, rmax = 16.e-6, 8.e-6
    sim = Simulation(32, zmax, 12, rmax, 2, zmax / 32 / c, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4, n_e=1.e24, fused=fused)
    sp = sim.ptcl[0]
    sp.uz = 0.1 * np.sin(2 * np.pi * sp.z / zmax)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.uz**2)
    sim.step(2)
    grids = sum(getattr(g, k).nbytes for g in sim.fld.interp for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')) \
        + sum(getattr(g, k).nbytes for g in sim.fld.spect for k in ('Ep', 'Em', 'Ez', 'Bp', 'Bm', 'Bz', 'Jp', 'Jm', 'Jz', 'rho_prev', 'rho_next'))
    n_arrays = 8 if fused else 14
    assert sim.last_step_bytes['d2h'] == n_arrays * 8 * sp.Ntot + grids
    eb = sum(getattr(g, k).nbytes for g in sim.fld.interp for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'))
    assert sim.last_step_bytes['h2d'] >= n_arrays * 8 * sp.Ntot + eb               # + the one-off table uploads
    assert sim.last_step_bytes['h2d'] < (n_arrays + 1) * 8 * sp.Ntot + eb + 4 * grids
    assert len(sp.Ez) == sp.Ntot and (np.any(sp.Ez != 0) != fused)


def test_two_rank_diagnostics_flow_gloo():
    """Diagnostics of a 2-rank run (gathered over the ranks) equal those of the single-domain run."""
    env = dict(os.environ, OMP_NUM_THREADS='2', ORACLE_NUM_THREADS='2', MGPU_NZ_PER_RANK='64', MGPU_EXTRA='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29659',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py'), '--fake-device']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'MGPU_DIAG_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', sorted(test_gpu_w9_step_options.OPTIONS))
def test_step_options_flow(fake, tag, fused):
    test_gpu_w9_step_options.test_step_options_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'pml', 'boost'])
def test_mirror_flow(fake, tag, fused):
    test_gpu_w2_laser.test_mirror_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
def test_tracking_flow(fake, fused):
    test_gpu_w8_diags.test_tracked_ids_follow_the_particles(fused)


def test_fullsize_property_tests_flow(fake, monkeypatch):
    """The property tests of the full-size GPU suite, on a small grid (their logic, not their size)."""
    import test_gpu_x_fullsize_properties as t
    monkeypatch.setattr(t, 'NZ', 48)
    monkeypatch.setattr(t, 'NR', 24)
    t.test_sort_contract_full_size()
    for shape in ('linear', 'cubic'):
        t.test_deposition_conserves_charge_and_is_linear(shape)
        t.test_gather_reproduces_uniform_fields(shape)
    t.test_push_invariants()
    t.test_transform_round_trips()
    for fused in (False, True):
        t.test_cold_plasma_is_a_fixed_point_of_the_cycle(fused)


@pytest.mark.parametrize('fused', [False, True])
def test_species_mix_flow(fake, fused):
    test_gpu_w9_step_options.test_species_mix_vs_reference_golden(fused)


def test_kernel_level_operator_flow(fake):
    """The operator-level GPU tests of test_gpu_kernels.py / the Nm = 4 step (kernel results are the oracle's own here:
    what this run checks is the host side of the operator calls -- argument order, array swaps, sort-state flags)."""
    import test_gpu_kernels as K
    for shape in ('linear', 'cubic'):
        for Nm in (1, 2, 3):
            K.test_kernels_vs_reference_golden(shape, Nm)
        K.test_deposit_gather_vs_oracle_large(shape, 2)
        K.test_fused_deposition_paths_vs_oracle(shape, 2)
        test_gpu_step.test_step_Nm4_vs_oracle(shape)
    K.test_transforms_vs_numpy(48, 20)


def test_config_shape_tests_flow(fake, monkeypatch):
    """The BASELINE-shape parity tests on reduced sizes (their host logic; the oracle checks itself here)."""
    import test_gpu_x_config_shapes as t
    monkeypatch.setattr(t, 'SCALE', 0.125)
    for fused in (False, True):
        t.test_c2_shape(fused)
        t.test_c5_shape(fused)
    t.test_c4_shape(True)
    t.test_transforms_at_config_sizes(256, 256)
    t.test_mode3_transforms_512()


def test_bench_flow(fake, monkeypatch, capsys):
    """bench.py end to end on its `tiny` configuration (host flow: set-up, timed region bracketing, e2e leg with the
    byte accounting, JSON line with every key of the contract)."""
    import json
    import bench
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--config', 'tiny', '--steps', '3', '--warmup', '3', '--preroll', '2',
                                      '--no-cpu-baseline'])
    bench.main()
    line = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith('{')][-1]
    d = json.loads(line)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'gpu_launches', 'clocks', 'e2e', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['steps'] == 3 and d['n_gpus'] == 1 and d['dtype'] == 'f64' and d['config']['workload'].startswith('tiny')
    assert d['e2e']['value'] > 0 and 'error' not in d['e2e']
    n = d['config']['particles_total']
    assert 8 * 8 * n / 3 < d['e2e']['h2d_bytes_per_step'] < 14 * 8 * n / 3       # 8 particle arrays + the grids


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'boost'])
def test_bunch_injection_plane_flow(fake, tag, fused):
    test_gpu_w3_bunch.test_bunch_injection_plane_vs_reference_golden(tag, fused)


@pytest.mark.parametrize('fused', [False, True])
def test_diagnostic_trees_flow(fake, fused, tmp_path):
    test_gpu_w8_diags.test_diagnostic_trees_vs_reference_golden(fused, tmp_path)


@pytest.mark.parametrize('fused', [False, True])
def test_lab_frame_snapshots_flow(fake, fused, tmp_path):
    test_gpu_w8_diags.test_lab_frame_snapshots_vs_reference_golden(fused, tmp_path)


@pytest.fixture
def shim_h5py():
    """`import h5py` resolves to the API stand-in of oracle/ref_shim for one test: the diagnostics then take their
    h5py branch (fbpic_b200/openpmd_store.py) and write `.h5` containers."""
    path = os.path.join(ROOT, 'oracle', 'ref_shim')
    sys.modules.pop('h5py', None)
    sys.path.insert(0, path)
    yield
    sys.path.remove(path)
    sys.modules.pop('h5py', None)


def test_npz_archives_convert_to_hdf5(fake, tmp_path):
    """A diagnostics directory written without h5py (`.npz` archives) converts to the `.h5` files the h5py branch
    would have written: same tree, compared with the reference's."""
    import diag_cases
    from conftest import load_golden
    from fbpic_b200 import openpmd_store
    test_gpu_w8_diags.test_diagnostic_trees_vs_reference_golden(True, tmp_path)
    path = os.path.join(ROOT, 'oracle', 'ref_shim')
    sys.modules.pop('h5py', None)
    sys.path.insert(0, path)
    try:
        d = str(tmp_path / 'all' / 'hdf5')
        for name in sorted(os.listdir(d)):
            openpmd_store.npz_to_hdf5(os.path.join(d, name))
            os.remove(os.path.join(d, name))
        assert sorted(os.listdir(d)) == ['data00000000.h5', 'data00000004.h5']
        g = load_golden('diags_tree')
        ref, got = diag_cases.golden_files(g, 'all'), diag_cases.written_files(str(tmp_path / 'all'))
        for name in ref:
            diag_cases.compare_trees(got[name], ref[name], 1e-9, 'converted ' + name)
    finally:
        sys.path.remove(path)
        sys.modules.pop('h5py', None)


def test_diagnostics_through_the_h5py_api(fake, shim_h5py, tmp_path):
    """Same comparisons with the reference's trees, the files written and read back through the h5py calls."""
    import h5py
    assert h5py.__version__.endswith('shim')
    test_gpu_w8_diags.test_diagnostic_trees_vs_reference_golden(True, tmp_path / 'a')
    assert sorted(os.listdir(str(tmp_path / 'a' / 'all' / 'hdf5'))) == ['data00000000.h5', 'data00000004.h5']
    test_gpu_w8_diags.test_lab_frame_snapshots_vs_reference_golden(True, tmp_path / 'b')
    test_gpu_w8_diags.test_restart_from_checkpoint_continues_the_run(True, False, tmp_path / 'c')


@pytest.mark.parametrize('script', ['lwfa.py', 'boosted_frame.py', 'ionization_injection.py'])
def test_example_scripts_flow(fake, script, tmp_path, monkeypatch, capsys):
    """The two example scripts (the reference's documented input scripts, full size) run a few cycles end to end --
    laser set-up, bunch with space charge, antenna, moving window, boosted-frame and lab-frame diagnostics -- and
    leave readable diagnostics."""
    import runpy
    import numpy as np
    from fbpic_b200.diags import read_diag, list_iterations
    out = str(tmp_path / 'diags')
    monkeypatch.setattr(sys, 'argv', [script, '--steps', '3', '--out', out])
    np.random.seed(0)
    runpy.run_path(os.path.join(ROOT, 'examples', script), run_name='__main__')
    assert list_iterations(out) == [0]
    d = read_diag(out, 0)
    assert d['fields/E/r'].ndim == 3 and 'particles/electrons/position/x' in d
    if script == 'ionization_injection.py':
        assert 'fields/rho_electrons' in d and 'particles/electrons from N/weighting' in d
    if script == 'boosted_frame.py':
        assert list_iterations(out + '_lab') == list(range(11))
        lab = read_diag(out + '_lab', 0)
        assert lab['fields/E/z'].shape[:2] == (3, 75) and 'particles/bunch/momentum/z' in lab


def test_smoke_entry_flow(fake, capsys):
    """`__graft_entry__.smoke()` (the driver's round-end check) runs its comparison with the oracle"""
    import __graft_entry__ as entry
    entry.smoke()
    assert 'smoke OK' in capsys.readouterr().out


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_cpu_gpu_deposition_flow(fake, shape, fused, tmp_path):
    test_gpu_w8_diags.test_cpu_gpu_deposition_as_written(shape, fused, tmp_path)


@slow_flow
def test_boosted_particle_output_flow(fake, tmp_path):
    """the reference's tests/test_boosted_particle_output.py (500 cycles, 3000 tracked particles)"""
    test_gpu_w6_acceptance.test_boosted_output(tmp_path)


@slow_flow
def test_beam_focusing_flow(fake, tmp_path):
    """the reference's tests/test_beam_focusing.py (2 x 101 cycles, 40000 particles, Nr = 200)"""
    test_gpu_w6_acceptance.test_beam_focusing(tmp_path)


def test_bunch_gaussian_as_written_flow(fake, tmp_path):
    test_gpu_w3_bunch.test_bunch_gaussian_as_written(tmp_path)


@slow_flow
@pytest.mark.parametrize('z_boundary,use_galilean', [('periodic', False), ('open', True)])
def test_pml_laser_as_written_flow(fake, z_boundary, use_galilean, tmp_path):
    """the reference's tests/test_pml.py (2 x 601 cycles, fields only, restart in the middle)"""
    test_gpu_w1_pml_cross.test_pml_laser_as_written(z_boundary, use_galilean, tmp_path)


@slow_flow
@pytest.mark.parametrize('case', ['labframe', 'labframe_moving', 'boostedframe'])
def test_antenna_as_written_flow(fake, case):
    """the reference's tests/test_laser_antenna.py (420 cycles, fields + antenna particles)"""
    test_gpu_w2_laser.test_antenna_as_written(case)


@slow_flow
@pytest.mark.parametrize('gamma_boost', [None, 10])
def test_external_fields_as_written_flow(fake, gamma_boost):
    """the reference's tests/test_external_fields.py (400 calls of step(1))"""
    test_gpu_y_external.test_external_fields_as_written(gamma_boost)


def test_fewcycle_laser_as_written_flow(fake):
    """the reference's tests/test_fewcycle_laser.py"""
    test_gpu_w2_laser.test_fewcycle_laser_as_written()


@slow_flow
def test_flattenedgauss_laser_as_written_flow(fake):
    """the reference's tests/test_flattenedgauss_laser.py (Nz = 1600, Nr = 600)"""
    test_gpu_w2_laser.test_flattenedgauss_laser_as_written()


@slow_flow
@pytest.mark.parametrize('case', ['custom', 'gaussian', 'flattened_chirped', 'donut_chirped'])
def test_parax_approx_laser_as_written_flow(fake, case):
    """the reference's tests/test_parax_approx_laser.py (Nz = 800, Nr = 300, Nm = 3)"""
    test_gpu_w2_laser.test_parax_approx_laser_as_written(case)


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_charge_cylinder_as_written_flow(fake, shape):
    """the reference's tests/test_charge_cylinder.py"""
    test_gpu_w3_bunch.test_charge_cylinder_as_written(shape)


def test_bunch_from_openpmd_series_flow(fake, tmp_path):
    test_gpu_w3_bunch.test_bunch_from_openpmd_series(tmp_path)


def test_console_output_flow(fake, capsys):
    """`verbose_level` banner and `show_progress` line (fbpic/utils/printing.py), silent by default"""
    import numpy as np
    from scipy.constants import c
    from fbpic_b200 import Simulation
    zmax, rmax = 16.e-6, 8.e-6
    kw = dict(p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=1, p_nr=1, p_nt=4, n_e=1.e24)
    np.random.seed(0)
    Simulation(32, zmax, 12, rmax, 2, zmax / 32 / c, **kw).step(2)
    assert capsys.readouterr().out == ''
    sim = Simulation(32, zmax, 12, rmax, 2, zmax / 32 / c, verbose_level=2, n_order=8, gamma_boost=3., n_guard=12,
                     n_damp={'z': 12, 'r': 6}, boundaries={'z': 'open', 'r': 'open'}, **kw)
    out = capsys.readouterr().out
    for text in ('fbpic_b200', 'PSATD stencil order: 8', 'Transverse boundaries: open', 'Boosted frame gamma: 3',
                 'Guard region size'):
        assert text in out, (text, out)
    sim.step(3, show_progress=True)
    out = capsys.readouterr().out
    assert '3/3' in out and 'ms/step' in out and 'Total time taken' in out


@slow_flow
@pytest.mark.parametrize('frame', ['labframe', 'boostedframe'])
def test_ionization_as_written_flow(fake, frame, tmp_path):
    """the reference's tests/test_ionization.py (N5+ fraction after a laser pulse, Chen et al. 2013)"""
    getattr(test_gpu_y2_ionization_laser, 'test_ionization_' + frame)(tmp_path)


def test_ionization_restart_flow(fake, tmp_path):
    test_gpu_w9b_ionization.test_restart_keeps_the_ionization_levels(tmp_path)


def test_ionization_kernels_and_plumbing_flow(fake):
    test_gpu_w9b_ionization.test_ionize_kernel_vs_reference_probabilities()
    test_gpu_w9b_ionization.test_push_p_ioniz_and_weights()
    test_gpu_w9b_ionization.test_grow_device_arrays_beyond_capacity()
    test_gpu_w9b_ionization.test_ionization_events_free_one_electron_each()
    test_gpu_w9b_ionization.test_ionizable_species_through_window_sort_and_exchange()


@slow_flow
@pytest.mark.parametrize('gamma_boost', [1., 10.])
def test_compton_as_written_flow(fake, gamma_boost):
    """the reference's tests/test_compton.py (300000 electrons, 101 cycles of push + scattering)"""
    test_gpu_w9c_compton.test_compton_as_written(gamma_boost)


@pytest.mark.parametrize('gamma_boost', [1., 10.])
def test_compton_momentum_conservation_flow(fake, gamma_boost):
    test_gpu_w9c_compton.test_compton_momentum_conservation(gamma_boost)


def test_input_script_diagnostic_flow(fake, tmp_path, monkeypatch):
    """InputScriptDiagnostic: the text of the running script and the extra attributes land in the file attributes"""
    import numpy as np
    from scipy.constants import c
    from fbpic_b200 import Simulation, set_random_seed
    from fbpic_b200.openpmd_diag import InputScriptDiagnostic, FieldDiagnostic
    from fbpic_b200.openpmd_store import read_tree, existing_file
    script = tmp_path / 'my_run.py'
    script.write_text('# the input deck\nNz = 32\n')
    monkeypatch.setattr(sys, 'argv', ['python', str(script)])
    set_random_seed(4)
    first = np.random.rand()
    set_random_seed(4)
    assert np.random.rand() == first
    zmax, rmax = 16.e-6, 8.e-6
    sim = Simulation(32, zmax, 12, rmax, 2, zmax / 32 / c, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=1, p_nr=1,
                     p_nt=4, n_e=1.e24)
    out = str(tmp_path / 'diags')
    sim.diags = [FieldDiagnostic(2, sim.fld, comm=sim.comm, fieldtypes=['E'], write_dir=out),
                 InputScriptDiagnostic(2, sim.comm, param_dict={'runLabel': 'scan 7', 'a0': 2.5, 'box': [1, 2]},
                                       write_dir=out)]
    sim.step(3)
    tree = read_tree(existing_file(os.path.join(out, 'hdf5', 'data00000002')))
    assert bytes(tree['/@inputScript']).decode() == script.read_text()
    assert bytes(tree['/@runLabel']).decode() == 'scan 7' and float(tree['/@a0']) == 2.5
    assert list(tree['/@box']) == [1, 2] and '/data/2/fields/E/r' in tree


def test_lasy_file_laser_through_the_antenna_flow(fake, tmp_path):
    """A laser read from a (synthetic) lasy file is emitted by the antenna: after 60 cycles the pulse is on the grid
    with the polarisation of the file."""
    import numpy as np
    from scipy.constants import c
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, FromLasyFileLaser
    from test_lpa_utils_host import _write_lasy_like
    nt, nr, dt_, dr_ = 80, 40, 1.e-15, 1.e-6
    tt, rr = np.meshgrid(dt_ * np.arange(nt), dr_ * np.arange(nr), indexing='ij')
    env = 2.e12 * np.exp(-(tt - 30.e-15)**2 / (8.e-15)**2 - rr**2 / (10.e-6)**2)[None].astype(np.complex128)
    path = str(tmp_path / 'lasy_laser_00000.npz')
    _write_lasy_like(path, env, 'thetaMode', np.array([dt_, dr_]), np.array([0., 0.]), 2 * np.pi * c / 0.8e-6,
                     np.array([1. + 0.j, 0.j]))
    Nz, Nr, Nm, zmax, rmax = 160, 24, 2, 16.e-6, 24.e-6
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, zmax / Nz / c, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6},
                     boundaries={'z': 'open', 'r': 'reflective'})
    add_laser_pulse(sim, FromLasyFileLaser(path), method='antenna', z0_antenna=2.e-6)
    sim.step(120)
    g1 = sim.fld.interp[1]
    Er, Et = np.abs(g1.Er).max(), np.abs(g1.Et).max()
    assert 0.2e12 < 2 * Er < 2.4e12 and abs(Er - Et) < 1e-3 * Er          # x-polarised: |Er| = |Et| in mode 1
    assert np.abs(sim.fld.interp[0].Er).max() < 1e-4 * Er
    iz = np.unravel_index(np.abs(g1.Er).argmax(), g1.Er.shape)[0]
    assert 4.e-6 < g1.z[iz] < 14.e-6                                       # the pulse left the antenna towards +z
