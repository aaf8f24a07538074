"""plspy_b200 -- B200-native resampling engine behind plspy's `PLS(...)` API.

    import plspy_b200 as plspy
    result = plspy.PLS(X, groups_sizes, num_conditions, num_perm=5000, num_boot=5000, pls_method="mct")

The permutation test, the bootstrap test and the split-half loops of `plspy.core` run as hand-written
sm_100a CUDA kernels behind a C ABI (include/plsb200.h); everything else mirrors the reference's
Python interface.  There is no CPU fallback: importing the engine without the built library, or
running it without a CUDA device, raises.
"""
__version__ = "0.1.0"

from . import exceptions  # noqa: F401


def __getattr__(name):
    # heavy imports (torch, the CUDA library) happen on first use of the public API
    if name in ("PLS", "methods"):
        from . import pls
        return getattr(pls, name)
    if name in ("install", "uninstall"):
        from . import plugin
        return getattr(plugin, name)
    if name in ("io", "plugin", "bootstrap_permutation", "class_functions", "pls", "pls_classes", "resample",
                "split_half_resampling", "engine", "dist", "build", "_lib", "nifti", "device_analysis"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
