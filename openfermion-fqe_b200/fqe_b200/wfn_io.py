"""Wavefunction files interchangeable with the reference (Wavefunction.save / read,
/root/reference/src/fqe/wavefunction.py:726-765).

The reference pickles  [symmetry_map, conserved, conserve_spin, conserve_number, norb,
[key, FqeData], ...]  with its own FqeData objects inside.  This module

* READS such a file without the reference package: a restricted unpickler maps every
  ``fqe.*`` class to an attribute bag and only lets numpy's array reconstruction through
  (nothing else is importable from a wavefunction file), then takes ``.coeff`` of every sector;
* WRITES the same list layout with sector entries that, when un-pickled where the reference is
  installed, call ``fqe.fqe_data.FqeData(nalpha, nbeta, norb)`` and set its ``coeff`` - so a
  file written here loads with the reference's ``Wavefunction.read`` and vice versa.

Host-only (numpy); the device side is in wavefunction.py.
"""
import io
import pickle
import sys
import types
from typing import Dict, Tuple

import numpy

_NUMPY_OK = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "complex128"), ("numpy", "float64"),
    ("numpy", "int64"), ("numpy", "int32"), ("numpy", "uint64"), ("numpy", "bool_"),
}
_BUILTINS_OK = {"complex", "set", "frozenset", "tuple", "list", "dict", "int", "float", "bool",
                "slice", "range", "bytearray", "bytes"}


class _Bag:
    """stands in for any class of the reference package: keeps the pickled attributes"""

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            state = {**(state[0] or {}), **state[1]}
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "fqe" or module.startswith("fqe."):
            return type(name, (_Bag,), {"__module__": module})
        if (module, name) in _NUMPY_OK:
            return super().find_class(module, name)
        if module == "builtins" and name in _BUILTINS_OK:
            return super().find_class(module, name)
        if module == "collections" and name == "OrderedDict":
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"wavefunction file refers to {module}.{name}: refused")


def load(fh) -> dict:
    """-> {'conserved': {...}, 'conserve_spin': bool, 'conserve_number': bool, 'norb': int,
    'sectors': {(nele, m_s): complex128 ndarray}} from a file in either layout."""
    data = _RestrictedUnpickler(fh).load()
    if not isinstance(data, list) or len(data) < 2:
        raise ValueError("not a wavefunction file: top-level object is not a list")
    if len(data) >= 5 and isinstance(data[0], dict) and isinstance(data[1], dict) and \
            isinstance(data[2], bool) and isinstance(data[3], bool):
        # reference layout (wavefunction.py:756-765)
        out = {"conserved": dict(data[1]), "conserve_spin": data[2], "conserve_number": data[3],
               "norb": int(data[4]), "sectors": {}}
        entries = data[5:]
    elif isinstance(data[0], dict) and isinstance(data[1], (int, numpy.integer)):
        # layout of fqe_b200 round 1: [conserved, norb, [key, ndarray]...]
        out = {"conserved": dict(data[0]), "conserve_spin": True, "conserve_number": True,
               "norb": int(data[1]), "sectors": {}}
        entries = data[2:]
    else:
        raise ValueError("unknown wavefunction file layout")
    for entry in entries:
        key, sec = entry[0], entry[1]
        coeff = sec if isinstance(sec, numpy.ndarray) else getattr(sec, "coeff", None)
        if not isinstance(coeff, numpy.ndarray) or coeff.ndim != 2:
            raise ValueError(f"sector {key}: no coefficient matrix in the file")
        out["sectors"][(int(key[0]), int(key[1]))] = numpy.ascontiguousarray(
            coeff, dtype=numpy.complex128)
    return out


class _SectorEntry:
    """pickles as  fqe.fqe_data.FqeData(nalpha, nbeta, norb)  + {'coeff': array}"""

    def __init__(self, nalpha: int, nbeta: int, norb: int, coeff: numpy.ndarray):
        self.args = (int(nalpha), int(nbeta), int(norb))
        self.coeff = numpy.ascontiguousarray(coeff, dtype=numpy.complex128)


def dump(fh, conserved: Dict[str, int], norb: int,
         sectors: Dict[Tuple[int, int], Tuple[int, int, numpy.ndarray]],
         conserve_spin: bool = True, conserve_number: bool = True) -> None:
    """Write the reference's list layout; ``sectors[(nele, m_s)] = (nalpha, nbeta, coeff)``."""
    # pickle records a class by module and name and checks that the name resolves: offer a
    # placeholder ``fqe.fqe_data.FqeData`` for the duration of the dump if the reference is absent
    placeholders = {}
    try:
        import fqe.fqe_data as _ref   # the real package, when it is installed
        target = _ref.FqeData
    except Exception:
        for modname in ("fqe", "fqe.fqe_data"):
            if modname not in sys.modules:
                placeholders[modname] = types.ModuleType(modname)
                sys.modules[modname] = placeholders[modname]
        mod = sys.modules["fqe.fqe_data"]
        if not hasattr(mod, "FqeData"):
            mod.FqeData = type("FqeData", (), {"__module__": "fqe.fqe_data"})
            placeholders.setdefault("fqe.fqe_data:FqeData", mod)
        target = mod.FqeData

    class _Pickler(pickle.Pickler):
        def reducer_override(self, obj):
            if isinstance(obj, _SectorEntry):
                return target, obj.args, {"coeff": obj.coeff}
            return NotImplemented

    try:
        data = [{}, dict(conserved), bool(conserve_spin), bool(conserve_number), int(norb)]
        for key, (nalpha, nbeta, coeff) in sectors.items():
            data.append([(int(key[0]), int(key[1])), _SectorEntry(nalpha, nbeta, norb, coeff)])
        buf = io.BytesIO()
        _Pickler(buf, protocol=4).dump(data)
        fh.write(buf.getvalue())
    finally:
        if "fqe.fqe_data:FqeData" in placeholders:
            try:
                delattr(placeholders["fqe.fqe_data:FqeData"], "FqeData")
            except AttributeError:
                pass
        for modname in ("fqe.fqe_data", "fqe"):
            if modname in placeholders:
                sys.modules.pop(modname, None)
