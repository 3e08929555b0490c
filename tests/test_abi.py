"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that include/fbpic_b200.h declares; the ctypes table covers the same set; and the
product fails loudly (no CPU fallback) when no GPU is present."""
import os
import re
import ctypes
import numpy as np
import pytest

from conftest import ROOT
from fbpic_b200 import _lib, build


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'fbpic_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(b2_[A-Za-z0-9_]+)\s*\(', txt)))


def test_library_builds_and_exports_header_symbols():
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = header_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, 'declared in the header but not exported: %s' % missing


def test_ctypes_table_matches_header():
    assert sorted(_lib.EXPORTED) == header_symbols()


def test_spectral_struct_layout():
    # 23 pointers + 2 doubles, no padding (mirrors b2_spectral_mode)
    assert ctypes.sizeof(_lib.SpectralMode) == 23 * 8 + 16


@pytest.mark.skipif(_lib.cuda_available(), reason='needs a box without GPU')
def test_no_cpu_fallback():
    from fbpic_b200 import Simulation, B200Error
    from scipy.constants import c
    sim = Simulation(16, 16.e-6, 8, 8.e-6, 2, 1.e-6 / c, p_zmin=0, p_zmax=16.e-6, p_rmin=0, p_rmax=8.e-6,
                     p_nz=1, p_nr=1, p_nt=4, n_e=1.e24)
    assert sim.ptcl[0].Ntot > 0
    with pytest.raises(B200Error):
        sim.step(1)
    with pytest.raises(B200Error):
        sim.ptcl[0].push_x(1.e-16)
