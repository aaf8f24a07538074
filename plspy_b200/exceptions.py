"""Exception types of the drop-in, same names as plspy/core/exceptions.py:4-45."""


class Error(Exception):
    """Base class of the package's exceptions."""


class InputMatrixDimensionMismatchError(Error):
    """Input dimensions do not match the declared groups/conditions."""


class ImproperShapeError(Error):
    """A matrix has the wrong number of dimensions."""


class ConditionMatrixMalformedError(Error):
    """The condition matrix is not of shape (n,)."""


class NotImplementedError(Error):  # noqa: A001  (the reference shadows the builtin the same way)
    """The requested method/variant is not available."""


class MissingParameterError(Error):
    """A required parameter was not passed."""


class OutOfRangeError(Error):
    """An out-of-range index was referenced."""
