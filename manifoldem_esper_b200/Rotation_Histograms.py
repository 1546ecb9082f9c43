"""Optimal Nd eigen-rotations per PD by 2-D histogram sweeps (interface of
2_Manifold_Rotations/Rotation_Histograms.py).

Device side: manifold normalisation (:100-112) and the conic fits of the parabola detection (:153-212) for
every column pair of any number of PDs (esper_manifold_prepare, k7_conic.cu); the whole theta search
(:283-365) for any number of PDs (K5, esper_rot_sweep).  Host side (this file): the partner selection
from the R^2 table, the comboList / Rij bookkeeping (:217-281) and the files.  `op(pyDir, PD)` reads and
writes the same files as the reference; `prepare_many` + `sweep_many` (or `run_many`) are the batched entries
the launcher uses.  `prepare` / `findParabolas` are the host-only (numpy) statement of the same preparation.
"""
import os
import sys

import numpy as np

from . import ConicFit, Generate_Nd_Rot, _capi

THETA_START, THETA_STOP, THETA_STEP = -40, 40, .5     # Rotation_Histograms.py:290
BINS, HIST_LO, HIST_HI = 51, -1.5, 1.5                 # :308-310


def normalize(_d, to_sum=False, copy=True):
    """Rotation_Histograms.py:100-106."""
    d = _d if not copy else np.copy(_d)
    d -= np.min(d, axis=0)
    d /= (np.sum(d, axis=0) if to_sum else np.ptp(d, axis=0))
    d = 2. * (d - np.min(d)) / np.ptp(d) - 1
    return d


def findParabolas(dim, U_init, itIdx, CTF=True):
    """Best parabolic partner of every eigenfunction (:153-212, figures omitted)."""
    theta_total = int(dim * (dim - 1) / 2)
    R2_best_vals = np.zeros(dim - 1, dtype=float)
    R2_best_psi = np.zeros(dim - 1, dtype=int)
    for v1 in range(dim - 1):
        R2_best = 0
        for v2 in range(v1 + 1, dim):
            if CTF is False:
                R2 = ConicFit.fit3(U_init, v1, v2)[1]
            else:
                R2 = ConicFit.fit2(U_init, v1, v2)[3]
            if R2 > R2_best and v2 != R2_best_psi[0]:
                R2_best = R2
                R2_best_vals[v1] = R2
                R2_best_psi[v1] = v2
    if (R2_best_vals[0] < 0.2) and (itIdx == 1):          # rare pre-rotation retry (:202-209)
        thetas = np.zeros(shape=(theta_total, 1), dtype=float)
        thetas[0] = 45 * np.pi / 180
        R = Generate_Nd_Rot.genNdRotations(dim, thetas)
        U_init = np.matmul(R, U_init.T).T
        return findParabolas(dim, U_init, itIdx + 1, CTF)
    return R2_best_psi, U_init, itIdx


def plan(dim, R2_best_psi):
    """comboList, saveEigs, comboFinal, Rij_final of Rotation_Histograms.py:217-281."""
    comboList, saveEigs, seen = [], np.zeros(dim - 1), []
    for i in range(dim - 1):
        if i in seen:
            continue
        partner = i + 1 if R2_best_psi[i] < i else int(R2_best_psi[i])
        comboList.append((i, partner))
        saveEigs[i] = partner
        seen += [i, R2_best_psi[i]]
    comboFinal = []
    c1, c2 = comboList[0]
    for c in range(1, dim - 1):
        if c1 in comboList[c] or c2 in comboList[c]:
            continue
        comboFinal = [(c1, comboList[c][0]), (c1, comboList[c][1]), (c2, comboList[c][0]), (c2, comboList[c][1])]
        break
    planes = [(a, b) for a in range(dim - 1) for b in range(a + 1, dim)]
    Rij_final = []
    for r, (a, b) in enumerate(planes):
        for q in range(4):
            pa, pb = comboFinal[q]           # IndexError when no mixed pair exists, as in the reference (:266)
            if (a, b) == (pa, pb) or (b, a) == (pa, pb):
                Rij_final.append(r)
    return comboList, saveEigs, comboFinal, Rij_final


def prepare(U0, dim=6, PCA=False, CTF=True):
    """Everything up to the sweep for one PD -> dict(U_init, comboList, saveEigs, Rij_final, itIdx)."""
    U = normalize(U0)
    U_init = U[:, 0:dim] if PCA else U[:, 1:dim + 1]
    R2_best_psi, U_init, itIdx = findParabolas(dim, U_init, 1, CTF)
    comboList, saveEigs, comboFinal, Rij_final = plan(dim, R2_best_psi)
    return dict(U_init=np.ascontiguousarray(U_init), comboList=comboList, saveEigs=saveEigs,
                comboFinal=comboFinal, Rij_final=Rij_final, itIdx=itIdx)


def select_partners(dim, R2_pairs):
    """Partner selection of findParabolas (:159-177) given R^2 of every pair v1 < v2 in lexicographic order."""
    R2_best_vals = np.zeros(dim - 1, dtype=float)
    R2_best_psi = np.zeros(dim - 1, dtype=int)
    q = 0
    for v1 in range(dim - 1):
        R2_best = 0
        for v2 in range(v1 + 1, dim):
            R2 = R2_pairs[q]
            q += 1
            if R2 > R2_best and v2 != R2_best_psi[0]:
                R2_best = R2
                R2_best_vals[v1] = R2
                R2_best_psi[v1] = v2
    return R2_best_psi, R2_best_vals


def prepare_many(U0s, dim=6, PCA=False, CTF=True, ctx=None):
    """`prepare` for a batch of equally shaped manifolds with the normalisation and all dim(dim-1)/2 conic fits
    per PD on the GPU (one esper_manifold_prepare call).  Returns the list of prep dicts; `preps[0]['resident']`
    is the token `sweep_many` compares with the ctx to know that the device still holds exactly these U_init."""
    ctx = ctx or _capi.default_context()
    U0 = np.stack([np.asarray(u, dtype=np.float64) for u in U0s])
    kind = 2 if CTF else 3
    Ui, R2 = ctx.manifold_prepare(U0, dim, first=0 if PCA else 1, kind=kind)
    n_pd = U0.shape[0]
    psi, itIdx = [None] * n_pd, [1] * n_pd
    retry = []
    for i in range(n_pd):
        psi[i], vals = select_partners(dim, R2[i])
        if vals[0] < 0.2:                                   # rare pre-rotation retry (:202-209)
            retry.append(i)
    if retry:
        thetas = np.zeros(shape=(int(dim * (dim - 1) / 2), 1), dtype=float)
        thetas[0] = 45 * np.pi / 180
        R = Generate_Nd_Rot.genNdRotations(dim, thetas)
        for i in retry:
            Ui[i] = np.matmul(R, Ui[i].T).T
            itIdx[i] = 2
        R2r = ctx.conic_fit_batch(np.ascontiguousarray(Ui[retry]), kind=kind)
        for j, i in enumerate(retry):
            psi[i], _ = select_partners(dim, R2r[j])
    token = None if retry else ctx.rot_token           # a retry changed U_init on the host: upload again for the sweep
    preps = []
    for i in range(n_pd):
        comboList, saveEigs, comboFinal, Rij_final = plan(dim, psi[i])
        preps.append(dict(U_init=Ui[i], comboList=comboList, saveEigs=saveEigs, comboFinal=comboFinal,
                          Rij_final=Rij_final, itIdx=itIdx[i], R2=R2[i], resident=token))
    return preps


def run_many(U0s, dim=6, PCA=False, CTF=True, ctx=None):
    """Normalise, fit, plan and sweep a batch of manifolds -> (preps, list of thetas_all)."""
    ctx = ctx or _capi.default_context()
    preps = prepare_many(U0s, dim=dim, PCA=PCA, CTF=CTF, ctx=ctx)
    return preps, sweep_many(preps, dim, ctx=ctx)


def theta_grid():
    return np.arange(THETA_START, THETA_STOP, THETA_STEP) * np.pi / 180


def sweep_many(preps, dim, ctx=None, want_counts=False):
    """Run the theta search (:283-365) for a list of `prepare` results on the GPU in one call.

    Returns a list of thetas_all arrays ([n_subspaces, T], what the reference saves as
    PD%s_RotMatrices.npy) and, optionally, the non-zero-bin counts [n_pd, max_sub, 4, 160].
    """
    ctx = ctx or _capi.default_context()
    n_pd = len(preps)
    max_sub = max(len(p['comboList']) for p in preps)
    n_rij = len(preps[0]['Rij_final'])
    if any(len(p['Rij_final']) != n_rij for p in preps):
        raise ValueError('all PDs of one batch must have the same number of Rij operators')
    m = preps[0]['U_init'].shape[0]
    tok = getattr(ctx, 'rot_token', None)                          # device copy left by prepare_many, still current?
    resident = tok is not None and all(p.get('resident') is tok for p in preps)
    U = None if resident else np.stack([p['U_init'] for p in preps]).astype(np.float64)
    combo = np.zeros((n_pd, max_sub, 2), dtype=np.int32)
    n_sub = np.zeros(n_pd, dtype=np.int32)
    rij = np.zeros((n_pd, max(n_rij, 1)), dtype=np.int32)
    for i, p in enumerate(preps):
        if p['U_init'].shape != (m, dim):
            raise ValueError('all PDs of one batch must share the manifold shape [m, dim]')
        n_sub[i] = len(p['comboList'])
        combo[i, :n_sub[i]] = np.array(p['comboList'], dtype=np.int32)
        rij[i, :n_rij] = p['Rij_final']
    res = ctx.rot_sweep(U, combo, n_sub, rij[:, :n_rij], theta_grid(), BINS, HIST_LO, HIST_HI, want_counts=want_counts,
                        shape=(n_pd, m, dim))
    thetas, counts = res if want_counts else (res, None)
    out = []
    for i, p in enumerate(preps):
        th = thetas[i, :n_sub[i]].copy()
        if p['itIdx'] > 1:
            th[:, 0] += 45 * np.pi / 180            # :358-359
        out.append(th)
    return (out, counts) if want_counts else out


def op(pyDir, PD, dim=6, PCA=False, CTF=True, maniPath=None, ctx=None):
    parDir = os.path.abspath(os.path.join(pyDir, os.pardir))
    outDir = os.path.join(pyDir, 'Data_Rotations')
    if not os.path.exists(outDir):
        os.mkdir(outDir)
    pdDir = os.path.join(outDir, 'PD%s' % PD)
    if not os.path.exists(pdDir):
        os.mkdir(pdDir)
    if maniPath is None:
        if PCA is True:
            maniPath = os.path.join(parDir, '1_Embedding/PCA/Data_Manifold/PD_%s_vec.npy' % PD)
        else:
            maniPath = os.path.join(parDir, '1_Embedding/DM/Data_Manifolds/PD_%s_vec.npy' % PD)
    U0 = np.load(maniPath)
    print('Manifold Info:', np.shape(U0))
    prep = prepare_many([U0], dim=dim, PCA=PCA, CTF=CTF, ctx=ctx)[0]
    print('Parabolic subspaces:', prep['comboList'])
    print('Rij:', prep['comboFinal'])
    print('Rij indices:', prep['Rij_final'])
    thetas_all = sweep_many([prep], dim, ctx=ctx)[0]
    np.save(os.path.join(pdDir, 'PD%s_RotMatrices.npy' % PD), thetas_all)
    np.save(os.path.join(pdDir, 'PD%s_RotEigenfunctions.npy' % PD), prep['saveEigs'])
    np.save(os.path.join(pdDir, 'PD%s_RotParameters' % PD), np.array([dim]))
    print('')


if __name__ == '__main__':
    path1 = os.path.splitext(sys.argv[0])[0]
    path2, tail = os.path.split(path1)
    op(path2, sys.argv[1])
