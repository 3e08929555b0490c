"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that include/fbpic_b200.h declares; the ctypes table covers the same set; and the
product fails loudly (no CPU fallback) when no GPU is present."""
import os
import re
import ctypes
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


def header_prototypes():
    """name -> number of parameters, parsed from the declarations of include/fbpic_b200.h"""
    txt = open(os.path.join(ROOT, 'include', 'fbpic_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(b2_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;', txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ('', 'void') else len(args.split(','))
    return out


def test_ctypes_argument_counts_match_header():
    """Every ctypes signature has as many arguments as the C prototype it binds, and pointer / integer /
    floating-point arguments sit at the same positions."""
    protos = header_prototypes()
    assert sorted(protos) == sorted(_lib.EXPORTED)
    txt = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'fbpic_b200.h')).read(), flags=re.S)
    for name, argtypes in _lib._SIGNATURES.items():
        assert len(argtypes) == protos[name], (name, len(argtypes), protos[name])
        m = re.search(r'\b%s\s*\(([^;{]*?)\)\s*;' % name, txt, flags=re.S)
        params = [p.strip() for p in m.group(1).split(',')] if protos[name] else []
        for i, (p, t) in enumerate(zip(params, argtypes)):
            is_ptr = '*' in p
            is_dbl = (not is_ptr) and p.startswith('double')
            if is_ptr:
                ok = t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, 'contents') or 'LP_' in t.__name__
            elif is_dbl:
                ok = t is ctypes.c_double
            elif p.startswith('int64_t'):
                ok = t is ctypes.c_int64
            elif p.startswith('uint64_t'):
                ok = t is ctypes.c_uint64
            elif p.startswith('size_t'):
                ok = t is ctypes.c_size_t
            else:
                ok = t is ctypes.c_int
            assert ok, '%s: argument %d (%s) is bound as %s' % (name, i, p, t.__name__)


def test_product_never_reaches_the_checkers():
    """The oracle (oracle/), the fake device and the kernel-source emulation (tests/) are test infrastructure:
    no module of the product imports or loads them, and bench.py touches the oracle only in its cpu_baseline /
    --impl reference legs."""
    import ast
    pkg = os.path.join(ROOT, 'fbpic_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith('.py'):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom) and node.level == 0:
                    names = [node.module or '']
                for n in names:
                    top = n.split('.')[0]
                    assert top not in ('oracle', 'tests', 'fake_device', 'conftest'), (f, n)
            src = open(os.path.join(dirpath, f)).read()
            assert 'liboracle' not in src and 'libemu_ext' not in src, f
    bench = open(os.path.join(ROOT, 'bench.py')).read()
    uses = [m.start() for m in re.finditer(r'from oracle import|import oracle', bench)]
    assert uses, 'bench.py is expected to time the oracle as the CPU baseline'
    for pos in uses:
        ctx = bench[max(0, pos - 1500):pos]
        assert ('def time_oracle' in ctx) or ('def build_oracle_sim' in ctx) or ("args.impl == 'reference'" in ctx) \
            or ('cpu_baseline = None' in ctx), \
            'bench.py imports the oracle outside the cpu_baseline / reference legs'


def test_launch_limits_match_the_sources():
    """The per-launch limits the host layer chunks its work by are those compiled into the library."""
    import re
    from fbpic_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'fbpic_b200.h')).read()
    dht = open(os.path.join(ROOT, 'fbpic_b200', 'csrc', 'b2_dht.cu')).read()
    assert int(re.search(r'#define\s+B2_MAX_ARRAYS\s+(\d+)', header).group(1)) == _lib.MAX_ARRAYS
    assert int(re.search(r'#define\s+DHT_MAX_JOBS\s+(\d+)', dht).group(1)) == _lib.MAX_DHT_JOBS
