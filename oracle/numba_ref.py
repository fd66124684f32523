"""Imports the UNMODIFIED reference hop loop from /root/reference (build container
only -- the path does not exist on the GPU box).  Test infrastructure.

`fenics` is stubbed because the reference imports it at module level
(kmc_dopant_networks.py:28) although the hop loop never touches it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KMC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "kmc_dopant_networks.py"))


def load():
    """Return (module, seed_fn).  seed_fn(s) seeds numba's internal generator, which
    is what np.random.* inside the @jit loop draws from."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.modules.setdefault("fenics", types.ModuleType("fenics"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import kmc_dopant_networks as ref  # noqa: E402
    import numpy as np
    from numba import njit

    @njit
    def seed_fn(s):
        np.random.seed(s)

    return ref, seed_fn
