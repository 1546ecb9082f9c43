"""Composite n-dimensional rotation (interface of 2_Manifold_Rotations/Generate_Nd_Rot.py:4-45)."""
import numpy as np

from . import _capi


def genNdRotations(n, theta):
    """R = G_0 G_1 ... G_{T-1}, T = n(n-1)/2 Givens rotations in lexicographic plane order.

    Same argument conventions and errors as the reference: `theta` is a list of T angles or an
    ndarray whose first dimension is T ([T] or [T, 1]); a wrong count raises ValueError.
    """
    total = int(n * (n - 1.) / 2.)
    if isinstance(theta, list):
        if len(theta) != total:
            raise ValueError('%s angles required' % (total))
    if isinstance(theta, np.ndarray):
        if theta.shape[0] != total:
            raise ValueError('%sx1 array of angles required' % (total))
    flat = np.array([float(np.asarray(theta[r]).reshape(-1)[0]) for r in range(total)], dtype=np.float64)
    return _capi.nd_rotation(int(n), flat)
