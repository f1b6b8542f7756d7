"""Generate tests/golden/augm_offsets.npz by RUNNING the reference's own delay iterators.

The iterators (reference src/augm_iterators/*.py) are pure NumPy, so unlike the GPy-backed parts
they can be executed in the build container.  They are loaded by file path (the package's
__init__ imports GPy, which is absent).  Run once in the build container:
    python tests/golden/make_iterator_golden.py
/root/reference does not exist on the GPU box; tests only read the committed .npz.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/src/augm_iterators"


def _load():
    pkg = types.ModuleType("refaug")
    pkg.__path__ = [REF]
    sys.modules["refaug"] = pkg
    mods = {}
    for name in ("abstract_augm_iterator", "backward_augm_iterator", "even_augm_iterator"):
        spec = importlib.util.spec_from_file_location("refaug." + name, os.path.join(REF, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refaug." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["backward_augm_iterator"].BackwardAugmentation, mods["even_augm_iterator"].EvenAugmentation


if __name__ == "__main__":
    Backward, Even = _load()
    out = {}
    for n in range(0, 4):
        for dim in range(1, 5):
            out["backward_n%d_d%d" % (n, dim)] = np.array(list(Backward(n, dim))).reshape(-1, dim)
            out["even_n%d_d%d" % (n, dim)] = np.array(list(Even(n, dim))).reshape(-1, dim)
            assert len(out["backward_n%d_d%d" % (n, dim)]) == Backward(n, dim).new_entries_count()
            assert len(out["even_n%d_d%d" % (n, dim)]) == Even(n, dim).new_entries_count()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "augm_offsets.npz"), **out)
    print("wrote", len(out), "tables")
