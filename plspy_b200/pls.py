"""`PLS(...)` driver with the reference's signature and argument validation (plspy/core/pls.py:21-93)."""
from . import pls_classes

methods = {
    "mct": pls_classes._MeanCentreTaskPLS,
    "rb": pls_classes._RegularBehaviourPLS,
    "cst": pls_classes._ContrastTaskPLS,
    "csb": pls_classes._ContrastBehaviourPLS,
    "mb": pls_classes._MultiblockPLS,
    "cmb": pls_classes._ContrastMultiblockPLS,
}


def PLS(*args, **kwargs):
    """PLS(X, groups_sizes, num_conditions, Y=None, cond_order=None, num_perm=..., num_boot=..., mctype=...,
    contrasts=..., bscan=..., num_split=..., lv=..., CI=0.95, pls_method="mct")

    Drop-in for `plspy.PLS`: same arguments, same result attributes; the permutation, bootstrap and
    split-half loops run on a B200 through libplsb200.  Extra optional keywords: `perm_indices`,
    `boot_indices` (pre-generated resampling index matrices), `engine` (an `Engine` already holding X),
    `precision` ("fp64" = exact mode, default; "tf32x3" = fast mode for the bootstrap moment GEMM),
    `rotate_method` (2 = derived, the reference's behaviour and the default; 1 = Procrustes; 0 = per-permutation SVD,
    mct only -- see `_ResampleTestPLS._permutation_test`), `analysis` ("host" = LAPACK on the cross-block matrix as
    in the reference, default; "device" = through the Gram matrix on the GPU, sign of each latent
    variable fixed by convention -- see device_analysis.py).
    """
    pls_method = kwargs.pop("pls_method", "mct")
    kwargs["pls_alg"] = pls_method
    if "num_split" in kwargs:
        if kwargs["num_split"] < 0 or not isinstance(kwargs["num_split"], int):
            raise ValueError("Invalid number of splits provided. Value must be a positive integer.")
        if "CI" in kwargs:
            if kwargs["CI"] is None or kwargs["CI"] < 0 or kwargs["CI"] > 1:
                raise ValueError("CI should be within 0 and 1.")
        if "lv" in kwargs:
            if kwargs["lv"] <= 0 or not isinstance(kwargs["lv"], int):
                raise ValueError("lv must be a positive integer greater than 0.")
    if "num_boot" in kwargs:
        if kwargs["num_boot"] < 0 or not isinstance(kwargs["num_boot"], int):
            raise ValueError("Invalid number of bootstraps provided. Value must be a positive integer.")
    if "num_perm" in kwargs:
        if kwargs["num_perm"] < 0 or not isinstance(kwargs["num_perm"], int):
            raise ValueError("Invalid number of permutations provided. Value must be a positive integer.")
    return pls_classes.PLSBase._create(pls_method, *args, **kwargs)
