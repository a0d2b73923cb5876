"""Import the upstream reference (read-only at /root/reference) in THIS container only.

The reference's ``__init__`` chain imports plotting / profiling modules that are not
installed here and that the QMF hot path never touches; they are stubbed so the
package imports.  This file is dev tooling used to *generate* golden fixtures under
``tests/golden/`` (see tools/make_golden.py).  Nothing in ``tests/ -m gpu``,
``bench.py`` or ``__graft_entry__`` imports it: /root/reference does not exist on
the GPU box.
"""
import sys
import types

REF_ROOT = "/root/reference"


def import_reference():
    for name in [
        "skimage", "skimage.metrics", "skimage.io", "pyinstrument", "seaborn",
        "matplotlib", "matplotlib.pyplot", "opt_einsum",
    ]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: float("nan")
    sys.modules["skimage.io"].imread = None
    sys.modules["pyinstrument"].Profiler = object

    class _Stub:
        pass

    sys.modules["matplotlib.pyplot"].Figure = _Stub
    sys.modules["matplotlib.pyplot"].Axes = _Stub
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import lrf  # noqa: E402

    return lrf
