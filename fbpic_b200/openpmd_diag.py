"""Import path of the reference's diagnostics package (`from fbpic.openpmd_diag import FieldDiagnostic, ...`).
The classes are those of `fbpic_b200.diags`: same arguments and hooks, but `.npz` archives keyed by the openPMD record
paths instead of openPMD/HDF5 files (h5py is not available to this build; see fbpic_b200/diags.py)."""
from .diags import (FieldDiagnostic, ParticleDiagnostic, set_periodic_checkpoint,       # noqa: F401
                    restart_from_checkpoint)


def __getattr__(name):
    if name in ('BackTransformedFieldDiagnostic', 'BackTransformedParticleDiagnostic',
                'BoostedFieldDiagnostic', 'BoostedParticleDiagnostic', 'ParticleChargeDensityDiagnostic'):
        raise NotImplementedError('%s is not built (back-transformed / derived diagnostics are outside the hot path)'
                                  % name)
    raise AttributeError(name)
