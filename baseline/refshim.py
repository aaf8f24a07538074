"""Import the unmodified reference package from `baseline/_ref`.

`plspy/__init__.py:17-18` pulls in its io / visualize sub-packages, which need nibabel, matplotlib, seaborn and
nilearn; none of them is installed in this image and none is on the resampling path, so they are stubbed with
`MagicMock` before the import (the same shim `tests/golden/make_golden.py` uses).  No reference code is changed."""
import os
import sys
from unittest.mock import MagicMock

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_STUBS = ["nibabel", "nibabel.nifti1", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.patches",
          "nilearn", "nilearn.plotting", "seaborn"]


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "plspy", "core", "bootstrap_permutation.py"))


def import_reference():
    """Returns the reference's `plspy` module (installed copy under baseline/_ref)."""
    if not reference_available():
        raise ImportError(f"{REF_DIR}/plspy not found: run `python -m baseline.install_ref` in the build container "
                          "(needs /root/reference)")
    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import plspy
    if not os.path.abspath(plspy.__file__).startswith(REF_DIR):
        raise ImportError(f"`import plspy` resolved to {plspy.__file__}, not to the copy under {REF_DIR}")
    return plspy
