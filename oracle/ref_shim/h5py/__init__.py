"""Import-only stand-in for `h5py` (not installed in the build container): the reference's laser package
imports it at module level for `FromLasyFileLaser` (fbpic/lpa_utils/laser/laser_profiles.py:10), which the
fixtures never use.  Test infrastructure."""


def __getattr__(name):
    raise ImportError('h5py is not available in this container (stub in oracle/ref_shim)')
