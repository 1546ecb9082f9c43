"""TEST INFRASTRUCTURE: the input side of the reference's consumer, 3_Manifold_Binning/Manifold_Binning.py:88-139, restated so
that the files our stages write can be loaded the way that script loads them (paths, dtypes, shapes, the v1/v2 lists, the
manifold normalisation and the U_init slice).  Nothing in the product imports this file."""
import os

import numpy as np


def normalize(_d, copy=True):
    """Manifold_Binning.py:119-128."""
    d = _d if not copy else np.copy(_d)
    d -= np.min(d, axis=0)
    d /= np.ptp(d, axis=0)
    return ((d - np.min(d)) / (np.ptp(d))) * 2 - 1


def load_like_manifold_binning(parDir, PD, PCA=False):
    """Returns dict(U0, U_init, rotMatrix, rotEigs, dim, v1_list, v2_list) or raises like the consumer would."""
    maniPath = os.path.join(parDir, '1_Embedding/DM/Data_Manifolds/PD_%s_vec.npy' % PD)                        # :88
    U0 = np.load(maniPath)
    rotMatrix = np.load(os.path.join(parDir, '2_Manifold_Rotations/Data_Rotations/PD%s/PD%s_RotMatrices.npy' % (PD, PD)))       # :106-111
    rotEigs = np.load(os.path.join(parDir, '2_Manifold_Rotations/Data_Rotations/PD%s/PD%s_RotEigenfunctions.npy' % (PD, PD)))
    rotParams = np.load(os.path.join(parDir, '2_Manifold_Rotations/Data_Rotations/PD%s/PD%s_RotParameters.npy' % (PD, PD)))
    dim = rotParams[0]
    v1_list, v2_list = [], []
    for i in range(len(rotEigs) - 1):                                                                        # :114-118
        if rotEigs[i] != 0:
            v1_list.append(i + 1)
            v2_list.append(int(rotEigs[i] + 1))
    U = normalize(U0)
    U_init = U[:, 0:dim] if PCA else U[:, 1:dim + 1]                                                         # :131-134
    assert U0.dtype == np.float64 and U0.ndim == 2 and U0.shape[1] >= dim + (0 if PCA else 1)
    assert U_init.shape == (U0.shape[0], dim) and np.all(U_init.min(axis=0) == -1.0) and np.all(U_init.max(axis=0) == 1.0)
    assert rotMatrix.dtype == np.float64 and rotMatrix.ndim == 2 and rotMatrix.shape[1] == dim * (dim - 1) // 2
    assert rotEigs.shape == (dim - 1,) and rotParams.dtype.kind == 'i'
    assert all(1 <= v <= dim for v in v1_list + v2_list)
    return dict(U0=U0, U_init=U_init, rotMatrix=rotMatrix, rotEigs=rotEigs, dim=int(dim), v1_list=v1_list, v2_list=v2_list)
