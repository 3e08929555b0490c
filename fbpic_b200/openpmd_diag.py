"""Import path of the reference's diagnostics package (`from fbpic.openpmd_diag import FieldDiagnostic, ...`).
The classes are those of `fbpic_b200.diags`: same arguments, hooks and openPMD tree; HDF5 files when h5py is installed,
`.npz` archives of the same tree otherwise (fbpic_b200/openpmd_store.py)."""
from .diags import (FieldDiagnostic, ParticleDiagnostic, ParticleChargeDensityDiagnostic,         # noqa: F401
                    InputScriptDiagnostic,
                    BackTransformedFieldDiagnostic, BoostedFieldDiagnostic, BackTransformedParticleDiagnostic,
                    BoostedParticleDiagnostic, set_periodic_checkpoint, restart_from_checkpoint)
