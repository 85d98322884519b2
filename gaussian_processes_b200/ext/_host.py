"""Argument validation shared by the numpy-signature shims.  The reference's Cython
declares every array as ``np.ndarray[float64, mode='c', ndim=k]`` and raises
``ValueError`` on a mismatch before the body runs; so do these."""
import numpy as np


def carray(a, ndim, name):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % a.dtype)
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    if not a.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return a


def out_array(a, shape, name="out"):
    carray(a, len(shape), name)
    if tuple(a.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, a.shape, tuple(shape)))
    if not a.flags.writeable:
        raise ValueError("buffer source array is read-only")
    return a
