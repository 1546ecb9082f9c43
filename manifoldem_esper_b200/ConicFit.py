"""Least-squares conic fits of a 2-D point cloud (interface of 3_Manifold_Binning/ConicFit.py:205-283).

Host-side glue called by Rotation_Histograms.findParabolas.  The reference solves the normal
equations with sympy's `Matrix.rref`; for floating-point entries that routine is a fraction-free
Gauss-Jordan elimination in 53-bit arithmetic (largest-magnitude pivot, rows normalised last), which
`_rref_solve` restates in numpy float64 so no symbolic package is needed.
"""
import numpy as np


def _rref_solve(A_aug):
    """Last column of the reduced row-echelon form of [AtA | Atb] (fraction-free Gauss-Jordan)."""
    M = np.array(A_aug, dtype=np.float64)
    rows, cols = M.shape
    piv_row = 0
    pivots = []
    for piv_col in range(cols):
        if piv_row >= rows:
            break
        col_abs = np.abs(M[piv_row:, piv_col])
        off = int(np.argmax(col_abs))                 # numeric column: partial pivoting, first maximum
        if col_abs[off] == 0:
            continue
        if off != 0:
            r = piv_row + off
            M[[piv_row, r]] = M[[r, piv_row]]
        pivots.append(piv_col)
        pv = M[piv_row, piv_col]
        prow = M[piv_row].copy()
        for r in range(rows):
            if r == piv_row:
                continue
            val = M[r, piv_col]
            if val == 0:
                continue
            M[r] = pv * M[r] - val * prow
        piv_row += 1
    for i, j in enumerate(pivots):
        pv = M[i, j]
        M[i, j + 1:] = M[i, j + 1:] / pv
        M[i, j] = 1.0
    return M[:, -1]


def fit2(U_rot_Nd1, v1, v2):
    """General conic x^2 + Bxy + Cy^2 + Dx + Ey + F = 0 -> [cF1, cF0, Theta, R2] (ConicFit.py:205-264)."""
    x = U_rot_Nd1[:, v1]
    y = U_rot_Nd1[:, v2]
    A = np.column_stack((x * y, y * y, x, y, np.ones(np.shape(x)[0])))
    b = -1. * (x * x)
    AtA = np.matmul(A.T, A)
    Atb = np.matmul(A.T, b)
    cF = _rref_solve(np.column_stack((AtA, Atb)))
    cF1 = np.array([1., cF[0], cF[1], cF[2], cF[3], cF[4]])
    tss = np.sum((b - np.mean(b)) ** 2)
    ssr = np.sum((b - np.matmul(A, cF)) ** 2)
    R2 = 1 - (ssr / tss)
    # rotation angle that removes the cross term (half-angle identities), then rotated coefficients
    adj = 1. - cF[1]
    hyp = np.sqrt(adj ** 2 + cF[0] ** 2)
    cos2Th = adj / hyp
    cosTh = np.sqrt((1 + cos2Th) / 2.)
    Theta = np.arccos(cosTh) if cF1[1] > 0 else -1. * np.arccos(cosTh)
    ct, st = np.cos(Theta), np.sin(Theta)
    cF0 = np.zeros(6)
    cF0[0] = (cF1[0] * ct ** 2) + (cF1[1] * st * ct) + (cF1[2] * st ** 2)
    cF0[1] = 2. * (cF1[2] - cF1[0]) * st * ct + cF1[1] * (ct ** 2 - st ** 2)
    cF0[2] = (cF1[0] * st ** 2) - (cF1[1] * st * ct) + (cF1[2] * ct ** 2)
    cF0[3] = cF1[3] * ct + cF1[4] * st
    cF0[4] = -1. * cF1[3] * st + cF1[4] * ct
    cF0[5] = cF1[5]
    return np.array([cF1, cF0, Theta, R2], dtype='object')


def fit3(U_rot_Nd, v1, v2):
    """Parabola y = a x^2 + b x + c -> [cF, R2] (ConicFit.py:267-283)."""
    x = U_rot_Nd[:, v1]
    b = U_rot_Nd[:, v2]
    A = np.column_stack((x ** 2, x, np.ones(np.shape(x)[0])))
    AtA = np.matmul(A.T, A)
    Atb = np.matmul(A.T, b)
    cF = _rref_solve(np.column_stack((AtA, Atb)))
    tss = np.sum((b - np.mean(b)) ** 2)
    ssr = np.sum((b - np.matmul(A, cF)) ** 2)
    R2 = 1 - (ssr / tss)
    return np.array([cF, R2], dtype='object')
