"""Split-half reproducibility tests (plspy/core/split_half_resampling.py:23-401, :404-861) -- GPU path.

Placeholder until the K x K Jacobi eigensolver kernel lands: the functions keep the reference's
signatures and fail loudly rather than falling back to the CPU."""
from . import exceptions


def split_half_test_train(pls_alg, matrix, Y, cond_order, num_split, mctype=None, contrasts=None, bscan=None,
                          Xbscan=None, Ybscan=None, engine=None, draws=None):
    raise exceptions.NotImplementedError("split_half_test_train is not yet available on the B200 path")


def split_half(pls_alg, matrix, Y, cond_order, num_split, mctype=None, contrasts=None, bscan=None, Xbscan=None,
               Ybscan=None, lv=1, CI=0.95, engine=None, draws=None):
    raise exceptions.NotImplementedError("split_half is not yet available on the B200 path")
