#!/usr/bin/env python
"""Generate tests/golden/* by EXECUTING THE REFERENCE's own code (authoring container only).

The reference (/root/reference, read-only, pure Python) is copied to a scratch directory,
given the three documented text patches and run with non-recording
stubs for the plotting modules and an `mrcfile` stand-in.  Nothing from the reference is
copied into this repository: only the numeric outputs of running it on seeded inputs.

    python tools/gen_goldens.py [--ref /root/reference] [--jobs 8] [--skip-rot]

Outputs (all under tests/golden/):
  nd_rot.npz            genNdRotations(n, theta) for n in {2,3,6,8}            (P7)
  stage123_small.npz    stack -> Distances.op -> GaussianBandwidth.op -> DiffusionMaps.op
  stage123_pristine.npz same on a noise-free stack (small distances)            (P2)
  stage1_dupes.npz      Distances.op on a stack with exact duplicates           (P3)
  stage23_medium.npz    N=400 distance matrix -> DiffusionMaps.op
  rot_inputs.npz        first 9 columns of 12 shipped Data_Manifolds_126 fixtures
  rot_ref_dim{6,8}.npz  Rotation_Histograms.op outputs for all 126 fixtures     (P5)
  rot_counts_dim8.npz   per-evaluation non-zero-bin counts for 3 PDs (histogramdd spy)
  stage1_p2.npz         Distances.op on the config-1 stand-in at full box: pristine N=400, 250x250 (--only-big)
  stage1_c2_rows.npz    Distances.op at BASELINE config 2 (N=2000, 250x250, SNR 0.1; ~20 min): every 25th row of D (--only-big)
  stage23_c2.npz        DiffusionMaps.op at BASELINE config 2 on the float64-Gram distance matrix of the same stack:
                        logSumWij, popt, eps, R2, val[:16], vec[:, :16] under np.random.seed (--only-big)
  tau_snr.npz           0_Data_Inputs/Pristine_AddTauSNR: AddNoise.find_SNR per state and Generate_TauSNR.op
                        (tau = 10, SNR = 0.1 as shipped) on a 6-state pristine stack under np.random.seed
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, 'tests', 'golden')
ROT_INPUT_PDS = [1, 2, 3, 17, 32, 48, 63, 64, 80, 97, 111, 126]
COUNT_PDS = [1, 63, 126]


class _Null(types.ModuleType):
    """Module/object stub: every attribute, call and item is the stub itself; records nothing."""

    def __init__(self, name='null'):
        super().__init__(name)
        self.__path__ = []

    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self

    def __iter__(self):
        return iter(())


def install_stubs():
    null = _Null('matplotlib')
    for name in ['matplotlib', 'matplotlib.figure', 'matplotlib.pyplot', 'matplotlib.cm',
                 'matplotlib.gridspec', 'matplotlib.ticker', 'pylab', 'mpl_toolkits', 'imageio',
                 'mpl_toolkits.mplot3d']:
        sys.modules[name] = null
    from manifoldem_esper_b200 import mrc
    m = types.ModuleType('mrcfile')
    m.mmap = mrc.mmap
    m.new_mmap = mrc.new_mmap
    sys.modules['mrcfile'] = m


def patched_copy(ref, dst):
    for d in ['1_Embedding/DM', '1_Embedding/PCA', '2_Manifold_Rotations', '3_Manifold_Binning']:
        shutil.copytree(os.path.join(ref, d), os.path.join(dst, d),
                        ignore=shutil.ignore_patterns('*.npy', '*.mp4'))
    os.makedirs(os.path.join(dst, '0_Data_Inputs'), exist_ok=True)

    def sub(path, pairs):
        p = os.path.join(dst, path)
        s = open(p).read()
        for a, b in pairs:
            assert a in s, (path, a)
            s = s.replace(a, b)
        open(p, 'w').write(s)

    # (1) ground-truth colour indices are not shipped; (2) `normed` kwarg removed from numpy;
    # (3) select the non-CTF branch / the PCA fixtures.  Arithmetic is untouched.
    sub('1_Embedding/DM/DiffusionMaps.py', [('groundTruth = True', 'groundTruth = False')])
    sub('1_Embedding/DM/Distances.py', [
        ('CTF = True #if using', 'CTF = False #if using'),
        ("'0_Data_Inputs/CTF5k15k_SNRpt1_ELS_2D'", "'0_Data_Inputs/Synth'")])
    # (4) CTF branch: a second copy of Distances.py with the shipped default `CTF = True` left as is (only the data
    # directory is redirected); Read_Alignment opens the star file with mode 'rU', which Python >= 3.11 rejects
    shutil.copy(os.path.join(ref, '1_Embedding/DM/Distances.py'), os.path.join(dst, '1_Embedding/DM/Distances_ctf.py'))
    sub('1_Embedding/DM/Distances_ctf.py', [("'0_Data_Inputs/CTF5k15k_SNRpt1_ELS_2D'", "'0_Data_Inputs/SynthCTF'")])
    sub('1_Embedding/DM/Read_Alignment.py', [("open(starfile, 'rU')", "open(starfile, 'r')")])
    sub('2_Manifold_Rotations/Rotation_Histograms.py', [
        ('groundTruth = True', 'groundTruth = False'),
        ('PCA = False', 'PCA = True'),
        (', normed=False', '')])


def import_from(path, name):
    import importlib.util
    d = os.path.dirname(path)
    if d not in sys.path:
        sys.path.insert(0, d)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def run_quiet(fn, *a):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def stage123(root, mods, stack, pd, seed):
    """stack (float32 [N,b,b]) -> reference outputs dict."""
    from manifoldem_esper_b200 import mrc
    ddir = os.path.join(root, '0_Data_Inputs', 'Synth')
    os.makedirs(ddir, exist_ok=True)
    mrc.write_stack(os.path.join(ddir, 'PD%s_SNR_tau5_stack.mrcs' % pd), stack)
    dm = os.path.join(root, '1_Embedding', 'DM')
    run_quiet(mods['Distances'].op, dm, pd)
    D = np.load(os.path.join(dm, 'Data_Distances', 'PD_%s_dist.npy' % pd))
    out = dict(stack=stack, D=D)
    if seed is None:
        return out
    logEps = np.arange(-100, 100.2, 0.2)
    np.random.seed(seed)
    a0 = 1 * (np.random.rand(4, 1) - .5)
    popt, logSum, resnorm, R2 = mods['GaussianBandwidth'].op(D, logEps, a0)
    np.random.seed(seed)
    run_quiet(mods['DiffusionMaps'].op, dm, pd)
    val = np.load(os.path.join(dm, 'Data_Manifolds', 'PD_%s_val.npy' % pd))
    vec = np.load(os.path.join(dm, 'Data_Manifolds', 'PD_%s_vec.npy' % pd))
    out.update(seed=seed, a0=a0, popt=popt, logSumWij=logSum, R2=R2,
               eps=np.exp(-(popt[1] / popt[0])), val=val[:32], vec=vec[:, :32])
    return out


def ctf_golden(root, stack, params, pd):
    """Reference Distances.op with CTF = True on (stack, star file) -> dict(stack, params, D, wiener)."""
    import warnings
    from manifoldem_esper_b200 import mrc, Read_Alignment
    ddir = os.path.join(root, '0_Data_Inputs', 'SynthCTF')
    os.makedirs(ddir, exist_ok=True)
    name = 'Hsp2D_5k15k_PD_%s' % pd
    mrc.write_stack(os.path.join(ddir, name + '.mrcs'), stack)
    Read_Alignment.write_star(os.path.join(ddir, name + '.star'), params[:, 2], volt=params[:, 3], Cs=params[:, 1],
                              ampc=params[:, 4], pix=params[:, 0], name=name)
    dm = os.path.join(root, '1_Embedding', 'DM')
    mod = import_from(os.path.join(dm, 'Distances_ctf.py'), 'Distances_ctf')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        run_quiet(mod.op, dm, pd)
    D = np.load(os.path.join(dm, 'Data_Distances', 'PD_%s_dist.npy' % pd))
    w = np.array(mrc.mmap(os.path.join(ddir, name + '_filtered.mrcs')).data)
    return dict(stack=stack, params=params, D=D, wiener=w)


def pca_golden(root, stack, pd):
    """Reference 1_Embedding/PCA/PCA.py op() on a stack -> dict(stack, val [N], vec [N, 20])."""
    from manifoldem_esper_b200 import mrc
    ddir = os.path.join(root, '0_Data_Inputs', '_Pristine_2D')
    os.makedirs(ddir, exist_ok=True)
    mrc.write_stack(os.path.join(ddir, 'PD_%s.mrcs' % pd), stack)
    pdir = os.path.join(root, '1_Embedding', 'PCA')
    mod = import_from(os.path.join(pdir, 'PCA.py'), 'PCA')
    run_quiet(mod.op, pdir, pd)
    return dict(stack=stack,
                val=np.load(os.path.join(pdir, 'Data_Manifolds', 'PD_%s_val.npy' % pd)),
                vec=np.load(os.path.join(pdir, 'Data_Manifolds', 'PD_%s_vec.npy' % pd)))


def rot_worker(args):
    root, ref, pd, dim, spy = args
    install_stubs()
    rh_dir = os.path.join(root, '2_Manifold_Rotations')
    src = open(os.path.join(rh_dir, 'Rotation_Histograms.py')).read().replace('dim = 6 #', 'dim = %d #' % dim)
    path = os.path.join(rh_dir, 'RH_dim%d.py' % dim)
    if not os.path.exists(path):
        tmp = path + '.%d' % os.getpid()
        open(tmp, 'w').write(src)
        os.replace(tmp, path)
    mod = import_from(path, 'RH_dim%d' % dim)
    pds = '%03d' % pd
    mdir = os.path.join(root, '1_Embedding', 'PCA', 'Data_Manifold')
    os.makedirs(mdir, exist_ok=True)
    dst = os.path.join(mdir, 'PD_%s_vec.npy' % pds)
    if not os.path.exists(dst):
        shutil.copy(os.path.join(ref, '1_Embedding/PCA/Data_Manifolds_126/PD%s_SS2_SNRpt1_tau5_vec.npy' % pds), dst)
    counts = []
    if spy:
        real = np.histogramdd

        def spy_hdd(*a, **k):
            H, e = real(*a, **k)
            counts.append(int(np.count_nonzero(H)))
            return H, e
        np.histogramdd = spy_hdd
    try:
        run_quiet(mod.op, rh_dir, pds)
    finally:
        if spy:
            np.histogramdd = real
    o = os.path.join(rh_dir, 'Data_Rotations', 'PD%s' % pds)
    return (pd, dim, np.load(os.path.join(o, 'PD%s_RotMatrices.npy' % pds)),
            np.load(os.path.join(o, 'PD%s_RotEigenfunctions.npy' % pds)),
            np.load(os.path.join(o, 'PD%s_RotParameters.npy' % pds)), np.array(counts, dtype=np.int32))


def tau_snr_golden(ref):
    """Run the reference's Generate_TauSNR.op / AddNoise from a scratch copy of 0_Data_Inputs (it writes next to itself)."""
    from manifoldem_esper_b200 import mrc, synth
    root = tempfile.mkdtemp(prefix='esper_prep_')
    shutil.copytree(os.path.join(ref, '0_Data_Inputs/Pristine_AddTauSNR'), os.path.join(root, 'Pristine_AddTauSNR'))
    os.makedirs(os.path.join(root, 'Pristine_2D'))
    pristine = synth.pristine_states(box=24, n_side=3, pd=2)[:6]
    mrc.write_stack(os.path.join(root, 'Pristine_2D', 'PD_001.mrcs'), pristine)
    d = os.path.join(root, 'Pristine_AddTauSNR')
    AN = import_from(os.path.join(d, 'AddNoise.py'), 'AddNoise')
    GT = import_from(os.path.join(d, 'Generate_TauSNR.py'), 'Generate_TauSNR')
    params = np.array([AN.find_SNR(img, 0.1) for img in pristine])            # float32 (sig_mean, noise_std)
    seed = 20201
    np.random.seed(seed)
    run_quiet(GT.op, d, '001')
    mm = mrc.mmap(os.path.join(d, "Modified_Stacks", "PD001_SNR_tau10_stack.mrcs"))
    out = np.array(mm.data)
    mm.close()
    shutil.rmtree(root, ignore_errors=True)
    return dict(pristine=pristine, params=params, seed=np.array(seed), tau=np.array(10), snr=np.array(0.1), stack=np.array(out))


def big_goldens(root, parts):
    """Reference runs at the BASELINE sizes.  The stacks are regenerated from synth's seeds by the tests, only outputs are stored."""
    from manifoldem_esper_b200 import synth
    from oracle import esper_oracle as O
    dm = os.path.join(root, '1_Embedding', 'DM')
    mods = {n: import_from(os.path.join(dm, n + '.py'), n) for n in ['GaussianBandwidth', 'Distances', 'DiffusionMaps']}
    if 'p2' in parts:
        prist = synth.synthetic_pd(pd=4, box=250, n_side=20, tau=1, snr=None)         # N=400, the config-1 stand-in
        o = stage123(root, mods, prist, '004', 5)
        o.pop('stack')
        o['synth'] = np.array([4, 250, 20, 1, -1])                                    # pd, box, n_side, tau, snr (-1 = none)
        np.savez_compressed(os.path.join(GOLD, 'stage1_p2.npz'), **o)
        print('stage1_p2 done', flush=True)
    if 'c2dm' in parts:
        stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)          # N=2000, config 2
        D = O.distances_gram_f64(O.standardize_stack(stack))
        os.makedirs(os.path.join(dm, 'Data_Distances'), exist_ok=True)
        np.save(os.path.join(dm, 'Data_Distances', 'PD_001_dist.npy'), D)
        logEps = np.arange(-100, 100.2, 0.2)
        seed = 17
        np.random.seed(seed)
        a0 = 1 * (np.random.rand(4, 1) - .5)
        popt, logSum, resnorm, R2 = mods['GaussianBandwidth'].op(D, logEps, a0)
        np.random.seed(seed)
        run_quiet(mods['DiffusionMaps'].op, dm, '001')
        val = np.load(os.path.join(dm, 'Data_Manifolds', 'PD_001_val.npy'))
        vec = np.load(os.path.join(dm, 'Data_Manifolds', 'PD_001_vec.npy'))
        np.savez_compressed(os.path.join(GOLD, 'stage23_c2.npz'), seed=seed, a0=a0, popt=popt, logSumWij=logSum, R2=R2,
                            eps=np.exp(-(popt[1] / popt[0])), val=val[:16], vec=vec[:, :16], D_rows=D[::100],
                            synth=np.array([1, 250, 20, 5, 0.1]))
        print('stage23_c2 done', flush=True)
    if 'c2dist' in parts:
        stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
        o = stage123(root, mods, stack, '011', None)
        np.savez_compressed(os.path.join(GOLD, 'stage1_c2_rows.npz'), D_rows=o['D'][::25], step=np.array(25),
                            synth=np.array([1, 250, 20, 5, 0.1]))
        print('stage1_c2_rows done', flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--jobs', type=int, default=8)
    ap.add_argument('--skip-rot', action='store_true')
    ap.add_argument('--only-rot', action='store_true')
    ap.add_argument('--only-pca', action='store_true')
    ap.add_argument('--only-ctf', action='store_true')
    ap.add_argument('--only-prep', action='store_true')
    ap.add_argument('--only-big', action='store_true', help='the BASELINE-size goldens only (about half an hour of CPU)')
    ap.add_argument('--big-part', default='p2,c2dm,c2dist', help='which BASELINE-size goldens: p2, c2dm, c2dist')
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    os.environ['PYTHONDONTWRITEBYTECODE'] = '1'
    sys.dont_write_bytecode = True
    root = tempfile.mkdtemp(prefix='esper_ref_')
    patched_copy(args.ref, root)
    install_stubs()
    from manifoldem_esper_b200 import synth

    if args.only_big:
        big_goldens(root, args.big_part.split(','))
        shutil.rmtree(root, ignore_errors=True)
        return
    if args.only_prep or not (args.only_rot or args.only_pca or args.only_ctf):
        np.savez_compressed(os.path.join(GOLD, 'tau_snr.npz'), **tau_snr_golden(args.ref))
        print('data prep done')
    if args.only_prep:
        shutil.rmtree(root, ignore_errors=True)
        return
    if args.only_ctf or not (args.only_rot or args.only_pca):
        st, pa = synth.ctf_stack(pd=3, box=32, n_side=6, tau=2, snr=0.1)              # N=72, even box
        np.savez_compressed(os.path.join(GOLD, 'ctf_small.npz'), **ctf_golden(root, st, pa, '003'))
        st, pa = synth.ctf_stack(pd=5, box=25, n_side=4, tau=2, snr=0.5)              # N=32, odd box
        np.savez_compressed(os.path.join(GOLD, 'ctf_odd.npz'), **ctf_golden(root, st, pa, '005'))
        st, pa = synth.ctf_stack(pd=7, box=24, n_side=5, tau=1, snr=None)             # N=25, noise-free (cancellation)
        np.savez_compressed(os.path.join(GOLD, 'ctf_pristine.npz'), **ctf_golden(root, st, pa, '007'))
        print('ctf done')
    if args.only_ctf:
        shutil.rmtree(root, ignore_errors=True)
        return
    if args.only_pca or not args.only_rot:
        prist = synth.synthetic_pd(pd=5, box=32, n_side=8, tau=1, snr=None)           # N=64, raw pixel values
        np.savez_compressed(os.path.join(GOLD, 'pca_pristine.npz'), **pca_golden(root, prist, '005'))
        noisy = synth.synthetic_pd(pd=3, box=32, n_side=6, tau=3, snr=0.1)            # N=108
        np.savez_compressed(os.path.join(GOLD, 'pca_small.npz'), **pca_golden(root, noisy, '003'))
        print('pca done')
    if args.only_pca:
        shutil.rmtree(root, ignore_errors=True)
        return

    if not args.only_rot:
        dm = os.path.join(root, '1_Embedding', 'DM')
        mods = {n: import_from(os.path.join(dm, n + '.py'), n)
                for n in ['GaussianBandwidth', 'Distances', 'DiffusionMaps']}
        nd = import_from(os.path.join(root, '2_Manifold_Rotations', 'Generate_Nd_Rot.py'), 'Generate_Nd_Rot')

        rng = np.random.default_rng(42)
        nd_out = {}
        for n in (2, 3, 6, 8):
            T = n * (n - 1) // 2
            th = rng.uniform(-np.pi, np.pi, size=(5, T))
            nd_out['theta_%d' % n] = th
            nd_out['R_%d' % n] = np.stack([nd.genNdRotations(n, th[i].reshape(T, 1)) for i in range(5)])
            assert np.array_equal(nd.genNdRotations(n, list(th[0])), nd_out['R_%d' % n][0])
        np.savez_compressed(os.path.join(GOLD, 'nd_rot.npz'), **nd_out)
        print('nd_rot done')

        small = synth.synthetic_pd(pd=3, box=32, n_side=6, tau=3, snr=0.1)           # N=108
        np.savez_compressed(os.path.join(GOLD, 'stage123_small.npz'), **stage123(root, mods, small, '003', 7))
        print('stage123_small done')
        prist = synth.synthetic_pd(pd=5, box=32, n_side=8, tau=1, snr=None)           # N=64
        np.savez_compressed(os.path.join(GOLD, 'stage123_pristine.npz'), **stage123(root, mods, prist, '005', 11))
        print('stage123_pristine done')
        dupes = synth.synthetic_pd(pd=9, box=24, n_side=5, tau=2, snr=None)           # N=50, pairs equal
        np.savez_compressed(os.path.join(GOLD, 'stage1_dupes.npz'), **stage123(root, mods, dupes, '009', None))
        print('stage1_dupes done')
        med = synth.synthetic_pd(pd=2, box=24, n_side=10, tau=4, snr=1.0)             # N=400
        o = stage123(root, mods, med, '002', 3)
        o.pop('stack')
        np.savez_compressed(os.path.join(GOLD, 'stage23_medium.npz'), **o)
        print('stage23_medium done')

    if not args.skip_rot:
        fx = {}
        for pd in ROT_INPUT_PDS:
            fx['PD%03d' % pd] = np.load(os.path.join(
                args.ref, '1_Embedding/PCA/Data_Manifolds_126/PD%03d_SS2_SNRpt1_tau5_vec.npy' % pd))[:, :9]
        np.savez_compressed(os.path.join(GOLD, 'rot_inputs.npz'), **fx)
        import multiprocessing as mp
        with mp.get_context('fork').Pool(args.jobs) as pool:
            for dim in (6, 8):
                jobs = [(root, args.ref, pd, dim, dim == 8 and pd in COUNT_PDS) for pd in range(1, 127)]
                res = pool.map(rot_worker, jobs, chunksize=2)
                out, cnt = {}, {}
                for pd, _, rm, re_, rp, counts in res:
                    out['RotMatrices_%03d' % pd] = rm
                    out['RotEigenfunctions_%03d' % pd] = re_
                    assert int(rp[0]) == dim
                    if len(counts):
                        cnt['PD%03d' % pd] = counts
                np.savez_compressed(os.path.join(GOLD, 'rot_ref_dim%d.npz' % dim), **out)
                if cnt:
                    np.savez_compressed(os.path.join(GOLD, 'rot_counts_dim%d.npz' % dim), **cnt)
                print('rot dim', dim, 'done')
    shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main()
