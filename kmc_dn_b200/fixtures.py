"""Reader for the reference's `.kmc` files (pickled attribute dicts written by
`kmc_dn.saveSelf`, reference kmc_dopant_networks.py:943-963).

The reference loads them with a bare `pickle.load` (:965-978).  The files are
third-party data, so this reader only admits the numpy globals that the 400
fixtures under thesis_indrek/tests/ actually contain; anything else raises.
"""
import io
import pickle

import numpy as np

_ALLOWED = {
    ("numpy", "ndarray"),
    ("numpy", "dtype"),
    ("numpy.core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "scalar"),
}


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) not in _ALLOWED:
            raise pickle.UnpicklingError(f"refusing global {module}.{name} in .kmc file")
        # import the module as the file names it; only if this numpy does not have it, try the other spelling of the
        # same package (numpy.core <-> numpy._core: the fixtures were written by a numpy that had `numpy.core`)
        try:
            mod = __import__(module, fromlist=[name])
        except ImportError:
            other = (module.replace("numpy._core", "numpy.core") if module.startswith("numpy._core")
                     else module.replace("numpy.core", "numpy._core"))
            mod = __import__(other, fromlist=[name])
        return getattr(mod, name)


def load_kmc(path_or_bytes):
    """Return the attribute dict stored in a `.kmc` file."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        data = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            data = f.read()
    d = _RestrictedUnpickler(io.BytesIO(data)).load()
    if not isinstance(d, dict):
        raise pickle.UnpicklingError(".kmc payload is not a dict")
    return d


def save_kmc(path, attrs):
    """Write an attribute dict in the reference's `.kmc` format (plain pickle of
    list/tuple/int/float/ndarray values, reference :957-963)."""
    d = {k: v for k, v in attrs.items() if isinstance(v, (list, tuple, int, float, np.ndarray))}
    with open(path, "wb") as f:
        pickle.dump(d, f)
