"""PCA embedding per PD (interface of 1_Embedding/PCA/PCA.py).

`op(pyDir, PD)` loads `0_Data_Inputs/_Pristine_2D/PD_%s.mrcs` (PCA.py:35,44-45), subtracts the mean image
(PCA.py:53-55; no per-image standardisation on this path), and writes `Data_Manifolds/PD_%s_vec.npy` -- the
scores `Y.T.dot(W)` of the leading `dim` = 20 principal components (PCA.py:65-68), `[N, dim]` -- and
`PD_%s_val.npy` = s^2.  The reference saves all N squared singular values; this implementation computes
the leading `dim` (+ `extra`) on the GPU: the tensor-core Gram kernel gives Y^T Y and the Lanczos solver its
leading eigenpairs (v, s^2); scores = v s.  Column signs follow the library's convention (largest entry
positive); the reference's are LAPACK's.
"""
import os
import sys

import numpy as np

from . import _capi, mrc


def embed(stack, dim=20, tol=0.0, ctx=None):
    """stack [N, box, box] or [N, P] float32 -> (eig_vals[dim] = s^2, U[N, dim] = scores, lanczos steps)."""
    ctx = ctx or _capi.default_context()
    stack = np.asarray(stack)
    N = stack.shape[0]
    flat = np.ascontiguousarray(stack.reshape(N, -1), dtype=np.float32)
    if dim > N - 1:
        raise ValueError('PCA.embed: dim=%d but the centred stack of %d images has rank <= %d' % (dim, N, N - 1))
    return ctx.pca_embed(flat, k=dim, tol=tol)


def op(pyDir, PD, dim=20, dataDir=None, ctx=None):
    parDir1 = os.path.abspath(os.path.join(pyDir, os.pardir))
    parDir2 = os.path.abspath(os.path.join(parDir1, os.pardir))
    if dataDir is None:
        dataDir = os.path.join(parDir2, '0_Data_Inputs/_Pristine_2D')
    outDir = os.path.join(pyDir, 'Data_Manifolds')
    if not os.path.exists(outDir):
        os.makedirs(outDir)
    print('PD:', PD)
    init_stack = mrc.mmap(os.path.join(dataDir, 'PD_%s.mrcs' % PD))
    print('Computing SVD...')
    eig_vals, U, _ = embed(init_stack.data, dim=dim, ctx=ctx)
    print('SVD complete')
    np.save(os.path.join(outDir, 'PD_%s_vec.npy' % PD), U)
    np.save(os.path.join(outDir, 'PD_%s_val.npy' % PD), eig_vals)
    init_stack.close()


if __name__ == '__main__':
    path1 = os.path.splitext(sys.argv[0])[0]
    path2, tail = os.path.split(path1)
    op(path2, sys.argv[1])
