"""ctypes binding of libesper_b200.so (include/esper_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present, the
calls raise.  Everything numeric in this package goes through this module.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libesper_b200.so')

ESPER_OK = 0
E_ARG, E_CUDA, E_NOMEM, E_NOCONV, E_STATE, E_ARCH, E_NOFIT = -1, -2, -3, -4, -5, -6, -7
DIST_STANDARDIZE, DIST_CENTER, DIST_REFINE = 1, 2, 4
DIST_DEFAULT = DIST_STANDARDIZE | DIST_CENTER | DIST_REFINE

_lib = None
_lib_lock = threading.Lock()

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/esper_b200.h declares
SIGNATURES = {
    'esper_abi_version': (C.c_int, []),
    'esper_create': (C.c_void_p, [C.c_int, C.c_void_p]),
    'esper_destroy': (None, [C.c_void_p]),
    'esper_last_error': (C.c_char_p, [C.c_void_p]),
    'esper_sync': (C.c_int, [C.c_void_p]),
    'esper_stream': (C.c_void_p, [C.c_void_p]),
    'esper_set_profiling': (C.c_int, [C.c_void_p, C.c_int]),
    'esper_get_kernel_ms': (C.c_int, [C.c_void_p, C.c_char_p, _f64p, _i32p]),
    'esper_reset_profiling': (C.c_int, [C.c_void_p]),
    'esper_launch_count': (C.c_longlong, [C.c_void_p]),
    'esper_stat': (C.c_longlong, [C.c_void_p, C.c_char_p]),
    'esper_dist_rms': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    'esper_upload_images': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    'esper_pd_embed_fused': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_dist_config': (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    'esper_upload_dist': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'esper_eps_sweep': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    'esper_markov': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    'esper_dmap_config': (C.c_int, [C.c_void_p, C.c_int]),
    'esper_gram_config': (C.c_int, [C.c_void_p, C.c_int]),
    'esper_gram_info': (C.c_int, [C.c_void_p, C.c_void_p]),
    'esper_tanh_fit': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_sweep_config': (C.c_int, [C.c_void_p, C.c_int]),
    'esper_sweep_info': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_dmap_embed': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int,
                                   C.c_void_p, C.c_void_p, _i32p]),
    'esper_image_stats': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'esper_ctf_dist': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]),
    'esper_pca_embed': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                  C.c_void_p, C.c_void_p, _i32p]),
    'esper_dist_rms_rows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_void_p]),
    'esper_moment_layout': (C.c_int, [C.c_void_p]),
    'esper_moment_stats_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'esper_moment_accum_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    'esper_eps_sweep_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    'esper_degree_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    'esper_markov_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    'esper_symv_rows_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'esper_lanczos_start_d': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'esper_lanczos_ortho_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_tridiag_eigen_all': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'esper_tridiag_ritz': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, _i32p, C.c_void_p, C.c_void_p]),
    'esper_ritz_vectors_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_nd_rotation': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'esper_hist_edges': (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_void_p]),
    'esper_rot_hist_eval': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_void_p, C.c_void_p]),
    'esper_rot_sweep': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                  C.c_void_p, C.c_void_p]),
    'esper_manifold_prepare': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    'esper_comm_alloc': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'esper_comm_open': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    'esper_comm_close': (C.c_int, [C.c_void_p]),
    'esper_symv_allgather_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    'esper_comm_wait_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    'esper_comm_status': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint)]),
    'esper_lanczos_fused_steps_d': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    'esper_tau_snr_stack': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_ulonglong,
                                      C.c_void_p, C.c_void_p]),
    'esper_download_images': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    'esper_mrc_info': (C.c_int, [C.c_char_p, _i32p, _i32p, C.POINTER(C.c_longlong)]),
    'esper_upload_mrc': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, _i32p]),
    'esper_bin_accumulate': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'esper_conic_fit_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}


def load_library(path: str | None = None):
    """dlopen libesper_b200.so and declare every prototype.  Raises OSError if it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise OSError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                          '(or `make -C manifoldem_esper_b200/csrc`); there is no CPU fallback' % p)
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.esper_abi_version() != 1:
            raise OSError('libesper_b200.so ABI version mismatch')
        if path is None:
            _lib = lib
        return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class EsperError(RuntimeError):
    pass


class Context:
    """One esper_ctx: device buffers, a stream and the stage calls.  Not thread-safe (one per thread)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load_library()
        self._h = self._lib.esper_create(int(device), C.c_void_p(stream) if stream else None)
        if not self._h:
            raise EsperError(self._lib.esper_last_error(None).decode())
        self.device = int(device)
        self.rot_token = None          # identifies the batch of normalised manifolds resident on the device

    def close(self):
        if getattr(self, '_h', None):
            self._lib.esper_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc == ESPER_OK:
            return
        msg = self._lib.esper_last_error(self._h).decode()
        if rc in (E_ARG, E_STATE):
            raise ValueError(msg)
        raise EsperError('esper error %d: %s' % (rc, msg))

    # -- plumbing ---------------------------------------------------------------------------
    def sync(self):
        self._check(self._lib.esper_sync(self._h))

    @property
    def stream_ptr(self):
        """cudaStream_t (as an integer) every call of this ctx is enqueued on."""
        return int(self._lib.esper_stream(self._h) or 0)

    def set_profiling(self, on=True):
        self._check(self._lib.esper_set_profiling(self._h, int(on)))       # 0 off, 1 every launch, 2 the Gram launches only

    def reset_profiling(self):
        self._check(self._lib.esper_reset_profiling(self._h))

    def kernel_ms(self, name):
        ms, n = C.c_double(0), C.c_int(0)
        self._check(self._lib.esper_get_kernel_ms(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return int(self._lib.esper_launch_count(self._h))

    def stat(self, name):
        return int(self._lib.esper_stat(self._h, name.encode()))

    # -- stage 1 ----------------------------------------------------------------------------
    def dist_config(self, split_k=0, refine_tau=0.05):
        self._check(self._lib.esper_dist_config(self._h, int(split_k), float(refine_tau)))

    def upload_images(self, imgs):
        """Asynchronous H2D of a [n, P] float32 stack; `imgs` must stay alive until the next sync."""
        imgs = _check_stack(imgs)
        self._keep = imgs
        self._check(self._lib.esper_upload_images(self._h, _ptr(imgs), imgs.shape[0], imgs.shape[1]))
        return imgs.shape

    def dist_rms(self, imgs=None, n=None, P=None, flags=DIST_DEFAULT, download=True, out=None):
        """Distances.py:87-104 (non-CTF).  imgs=None -> resident stack of shape (n, P)."""
        if imgs is not None:
            imgs = _check_stack(imgs)
            n, P = imgs.shape
        if n is None or P is None:
            raise ValueError('dist_rms: n and P are required when imgs is None')
        D = None
        if download:
            D = out if out is not None else np.empty((n, n), dtype=np.float64)
            if D.shape != (n, n) or D.dtype != np.float64 or not D.flags.c_contiguous:
                raise ValueError('dist_rms: out must be a C-contiguous float64 [n, n] array')
        self._check(self._lib.esper_dist_rms(self._h, _ptr(imgs), int(n), int(P), int(flags), _ptr(D)))
        return D

    def gram_config(self, mode=1):
        """0 = 3xTF32 operand planes, 1 = 3xFP16 (default; used for z-scored stacks of more than 128 images)."""
        self._check(self._lib.esper_gram_config(self._h, int(mode)))

    def gram_info(self):
        """Operand format the last dist_rms used: 0 = TF32, 1 = FP16."""
        v = C.c_int(0)
        self._check(self._lib.esper_gram_info(self._h, C.byref(v)))
        return v.value

    def upload_dist(self, D):
        D = _check_square(D)
        self._keep = D
        self._check(self._lib.esper_upload_dist(self._h, _ptr(D), D.shape[0]))
        self.sync()

    # -- stage 2 ----------------------------------------------------------------------------
    def eps_sweep(self, D, coef, thr, m=None):
        """GaussianBandwidth.py:36-44.  D=None -> resident distance matrix of order m."""
        if D is not None:
            D = _check_square(D)
            m = D.shape[0]
        if m is None:
            raise ValueError('eps_sweep: m is required when D is None')
        coef = _as(coef, np.float64).reshape(-1)
        out = np.empty(coef.shape[0], dtype=np.float64)
        self._check(self._lib.esper_eps_sweep(self._h, _ptr(D), int(m), _ptr(coef), coef.shape[0], float(thr), _ptr(out)))
        return out

    def sweep_config(self, mode=1):
        """0 = general sweep kernel, 1 = moment sweep (default) with automatic return to the general kernel."""
        self._check(self._lib.esper_sweep_config(self._h, int(mode)))

    def sweep_info(self):
        """(path, bins) of the last eps_sweep: path 0 = general kernel, 1 = moment sweep."""
        path, bins = C.c_int(0), C.c_int(0)
        self._check(self._lib.esper_sweep_info(self._h, C.byref(path), C.byref(bins)))
        return path.value, bins.value

    # -- stage 3 ----------------------------------------------------------------------------
    def markov(self, D, eps, m=None, want_matrix=True):
        if D is not None:
            D = _check_square(D)
            m = D.shape[0]
        deg = np.empty(m, dtype=np.float64)
        Ms = np.empty((m, m), dtype=np.float64) if want_matrix else None
        self._check(self._lib.esper_markov(self._h, _ptr(D), int(m), float(eps), _ptr(deg), _ptr(Ms)))
        return deg, Ms

    def dmap_config(self, mode=0):
        """0 = latency (Ms resident on chip), 1 = throughput (two solves share every SM), 2 = throughput without the
        speculative batch (opt-in)."""
        self._check(self._lib.esper_dmap_config(self._h, int(mode)))

    def dmap_embed(self, D, eps, k=16, tol=0.0, max_iter=0, m=None):
        """DiffusionMaps.py:109-169 at a given eps -> (val[k] = lambda^2, vec[m, k], lanczos steps)."""
        if D is not None:
            D = _check_square(D)
            m = D.shape[0]
        if m is None:
            raise ValueError('dmap_embed: m is required when D is None')
        k = int(k)
        val = np.empty(max(k, 1), dtype=np.float64)
        vec = np.empty((m, max(k, 1)), dtype=np.float64)
        it = C.c_int(0)
        self._check(self._lib.esper_dmap_embed(self._h, _ptr(D), int(m), float(eps), k, float(tol), int(max_iter),
                                               _ptr(val), _ptr(vec), C.byref(it)))
        return val, vec, it.value

    def pd_embed_fused(self, imgs, logEps, a0, k=16, n=None, P=None, flags=DIST_DEFAULT, tol=0.0, max_iter=0, want_D=False, out=None):
        """Stages 1-3 in ONE library call (esper_pd_embed_fused): stack -> distances -> eps sweep -> tanh fit from a0 ->
        Markov matrix -> leading k eigenpairs.  imgs=None -> the resident stack of shape (n, P).
        Returns dict(logSumWij, popt, resnorm, R2, eps, slope, val, vec, iters[, D]).  When the fit from a0 fails the reference's
        `resnorm > 100` test, the retry loop of GaussianBandwidth.py:47-57 continues here (new random starts, esper_tanh_fit)
        and the embedding finishes on the resident matrix."""
        if imgs is not None:
            imgs = _check_stack(imgs)
            n, P = imgs.shape
        if n is None or P is None:
            raise ValueError('pd_embed_fused: n and P are required when imgs is None')
        logEps = _as(logEps, np.float64).reshape(-1)
        coef = 1. / (2. * np.exp(logEps))
        a0 = _as(a0, np.float64).reshape(-1)
        if a0.shape[0] != 4:
            raise ValueError('pd_embed_fused: a0 must have four entries')
        k = int(k)
        D = None
        if want_D or out is not None:
            D = out if out is not None else np.empty((n, n), dtype=np.float64)
            if D.shape != (n, n) or D.dtype != np.float64 or not D.flags.c_contiguous:
                raise ValueError('pd_embed_fused: out must be a C-contiguous float64 [n, n] array')
        L = np.empty(logEps.shape[0], dtype=np.float64)
        fit = np.zeros(8, dtype=np.float64)
        val = np.empty(max(k, 1), dtype=np.float64)
        vec = np.empty((n, max(k, 1)), dtype=np.float64)
        it = C.c_int(0)
        rc = self._lib.esper_pd_embed_fused(self._h, _ptr(imgs), int(n), int(P), int(flags), _ptr(logEps), _ptr(coef), logEps.shape[0],
                                            _ptr(a0), k, float(tol), int(max_iter), _ptr(D), _ptr(L), _ptr(fit), _ptr(val), _ptr(vec), C.byref(it))
        if rc == E_NOFIT:
            from . import GaussianBandwidth
            popt, resnorm, R2 = GaussianBandwidth.fit(logEps, L, 1 * (np.random.rand(4, 1) - .5), native=True)
            eps = float(np.exp(-(popt[1] / popt[0])))
            val, vec, iters = self.dmap_embed(None, eps, k=k, tol=tol, max_iter=max_iter, m=n)
            res = dict(logSumWij=L, popt=np.asarray(popt), resnorm=resnorm, R2=R2, eps=eps, slope=popt[0] * popt[2], val=val, vec=vec, iters=iters)
        else:
            self._check(rc)
            res = dict(logSumWij=L, popt=fit[:4].copy(), resnorm=float(fit[4]), R2=float(fit[5]), eps=float(fit[6]), slope=float(fit[7]),
                       val=val, vec=vec, iters=it.value)
        if D is not None:
            res['D'] = D
        return res

    def image_stats(self, imgs):
        """float32 (mean, std) per image, bit-identical to numpy's float32 `img.mean()`, `img.std()`."""
        imgs = _check_stack(imgs)
        n, P = imgs.shape
        mean = np.empty(n, dtype=np.float32)
        std = np.empty(n, dtype=np.float32)
        self._check(self._lib.esper_image_stats(self._h, _ptr(imgs), n, P, _ptr(mean), _ptr(std)))
        return mean, std

    def ctf_dist(self, stack, params, flags=DIST_DEFAULT, wiener=True, download=True):
        """Distances.py:107-145 (CTF branch): stack [n, box, box] float32, params [n, 5] float64 =
        (pixel size, Cs, defocus, voltage, amplitude contrast) -> (D [n, n] or None, wiener [n, box, box] float32 or None)."""
        stack = np.ascontiguousarray(stack, dtype=np.float32)
        if stack.ndim != 3 or stack.shape[1] != stack.shape[2]:
            raise ValueError('ctf_dist: stack must be [n, box, box], got %r' % (stack.shape,))
        n, box = stack.shape[0], stack.shape[1]
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.shape != (n, 5):
            raise ValueError('ctf_dist: params must be [%d, 5], got %r' % (n, params.shape))
        D = np.empty((n, n), dtype=np.float64) if download else None
        W = np.empty((n, box, box), dtype=np.float32) if wiener else None
        self._check(self._lib.esper_ctf_dist(self._h, _ptr(stack), n, box, _ptr(params), int(flags), _ptr(D), _ptr(W)))
        return D, W

    def pca_embed(self, imgs, k=20, tol=0.0, max_iter=0, n=None, P=None):
        """PCA.py:49-71 -> (val[k] = s^2 descending, U[n, k] = scores v s, lanczos steps)."""
        if imgs is not None:
            imgs = _check_stack(imgs)
            n, P = imgs.shape
        if n is None or P is None:
            raise ValueError('pca_embed: n and P are required when imgs is None')
        k = int(k)
        val = np.empty(max(k, 1), dtype=np.float64)
        U = np.empty((n, max(k, 1)), dtype=np.float64)
        it = C.c_int(0)
        self._check(self._lib.esper_pca_embed(self._h, _ptr(imgs), int(n), int(P), k, float(tol), int(max_iter),
                                              _ptr(val), _ptr(U), C.byref(it)))
        return val, U, it.value

    # -- oversized-PD row-block path (device pointers are plain ints, e.g. torch.Tensor.data_ptr()) ----
    def dist_rms_rows(self, imgs, n, P, row0, nrows, flags=DIST_DEFAULT, download=False):
        if imgs is not None:
            imgs = _check_stack(imgs)
            n, P = imgs.shape
        D = np.empty((nrows, n), dtype=np.float64) if download else None
        self._check(self._lib.esper_dist_rms_rows(self._h, _ptr(imgs), int(n), int(P), int(flags), int(row0), int(nrows), _ptr(D)))
        return D

    def moment_stats_rows(self, nrows, n):
        """[min non-zero D^2, max D^2, zeros, valid elements] of the resident row block (moment sweep, step 1)."""
        out = np.empty(4, dtype=np.float64)
        self._check(self._lib.esper_moment_stats_rows(self._h, int(nrows), int(n), _ptr(out)))
        return out

    def moment_accum_rows(self, nrows, n, coef, thr, plan):
        """Moments and boundary sums [MB, MV] of the resident row block against `plan` (moment_sweep.plan)."""
        from . import moment_sweep
        coef = _as(coef, np.float64).reshape(-1)
        plan = _as(plan, np.float64).reshape(-1)
        if plan.shape[0] != moment_sweep.PLAN_LEN:
            raise ValueError('moment_accum_rows: plan must have %d entries' % moment_sweep.PLAN_LEN)
        out = np.empty((moment_sweep.MB, moment_sweep.MV), dtype=np.float64)
        self._check(self._lib.esper_moment_accum_rows(self._h, int(nrows), int(n), _ptr(coef), coef.shape[0], float(thr), _ptr(plan), _ptr(out)))
        return out

    def eps_sweep_rows(self, nrows, n, coef, thr):
        coef = _as(coef, np.float64).reshape(-1)
        out = np.empty(coef.shape[0], dtype=np.float64)
        self._check(self._lib.esper_eps_sweep_rows(self._h, int(nrows), int(n), _ptr(coef), coef.shape[0], float(thr), _ptr(out)))
        return out

    def degree_rows(self, nrows, n, eps):
        out = np.empty(nrows, dtype=np.float64)
        self._check(self._lib.esper_degree_rows(self._h, int(nrows), int(n), float(eps), _ptr(out)))
        return out

    def markov_rows(self, nrows, n, eps, degree_all):
        degree_all = _as(degree_all, np.float64).reshape(-1)
        if degree_all.shape[0] != n:
            raise ValueError('markov_rows: degree_all must have n entries')
        self._check(self._lib.esper_markov_rows(self._h, int(nrows), int(n), float(eps), _ptr(degree_all)))

    def symv_rows_d(self, nrows, n, q_ptr, w_ptr):
        self._check(self._lib.esper_symv_rows_d(self._h, int(nrows), int(n), C.c_void_p(q_ptr), C.c_void_p(w_ptr)))

    # -- fused SYMV + all-gather over peer memory (row-block path) ---------------------------
    def comm_alloc(self, n_pad, world):
        """Allocate this rank's gather buffer; returns its 64-byte CUDA IPC handle (bytes) for the peers."""
        h = (C.c_char * 64)()
        self._check(self._lib.esper_comm_alloc(self._h, int(n_pad), int(world), h))
        return bytes(h.raw)

    def comm_open(self, handles, rank, world):
        """handles: list of `world` 64-byte IPC handles in rank order."""
        blob = b''.join(handles)
        if len(blob) != 64 * int(world):
            raise ValueError('comm_open: need one 64-byte handle per rank')
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self._lib.esper_comm_open(self._h, buf, int(rank), int(world)))

    def comm_close(self):
        self._check(self._lib.esper_comm_close(self._h))

    def symv_allgather_d(self, nrows, n, row0, q_ptr, step):
        self._check(self._lib.esper_symv_allgather_d(self._h, int(nrows), int(n), int(row0), C.c_void_p(q_ptr), int(step)))

    def comm_wait_d(self, step, n):
        """Enqueue the unpack of elements [0, n) of `step`; returns the device address of the dense gathered vector."""
        p = C.c_void_p(0)
        self._check(self._lib.esper_comm_wait_d(self._h, int(step), int(n), C.byref(p)))
        return int(p.value)

    def comm_status(self):
        st = C.c_uint(0)
        self._check(self._lib.esper_comm_status(self._h, C.byref(st)))
        return int(st.value)

    def lanczos_fused_steps_d(self, nrows, n, row0, V_ptr, ab_ptr, j0, n_steps, step0):
        """n_steps Lanczos steps (fused SYMV + packet gather + orthogonalisation) enqueued by one host call."""
        self._check(self._lib.esper_lanczos_fused_steps_d(self._h, int(nrows), int(n), int(row0), C.c_void_p(V_ptr), C.c_void_p(ab_ptr),
                                                          int(j0), int(n_steps), int(step0)))

    def lanczos_start_d(self, n, degree_all, V_ptr):
        degree_all = _as(degree_all, np.float64).reshape(-1)
        self._check(self._lib.esper_lanczos_start_d(self._h, int(n), _ptr(degree_all), C.c_void_p(V_ptr)))

    def lanczos_ortho_d(self, n, j, V_ptr, w_ptr, ab_ptr):
        self._check(self._lib.esper_lanczos_ortho_d(self._h, int(n), int(j), C.c_void_p(V_ptr), C.c_void_p(w_ptr), C.c_void_p(ab_ptr)))

    def ritz_vectors_d(self, n, steps, k, V_ptr, S):
        S = _as(S, np.float64)
        vec = np.empty((n, k), dtype=np.float64)
        self._check(self._lib.esper_ritz_vectors_d(self._h, int(n), int(steps), int(k), C.c_void_p(V_ptr), _ptr(S), _ptr(vec)))
        return vec

    # -- stage 4 ----------------------------------------------------------------------------
    def rot_hist_eval(self, Uinit, R, pd_of_eval, v1, v2, bins=51, lo=-1.5, hi=1.5, want_hist=False):
        Uinit = _as(Uinit, np.float64)
        if Uinit.ndim == 2:
            Uinit = Uinit[None]
        n_pd, m, dim = Uinit.shape
        R = _as(R, np.float64).reshape(-1, dim, dim)
        ne = R.shape[0]
        pd_of_eval, v1, v2 = (_as(a, np.int32).reshape(-1) for a in (pd_of_eval, v1, v2))
        if not (pd_of_eval.shape[0] == v1.shape[0] == v2.shape[0] == ne):
            raise ValueError('rot_hist_eval: R, pd_of_eval, v1, v2 must have one entry per evaluation')
        nz = np.empty(ne, dtype=np.int32)
        H = np.empty((ne, bins, bins), dtype=np.int32) if want_hist else None
        self.rot_token = None
        self._check(self._lib.esper_rot_hist_eval(self._h, _ptr(Uinit), n_pd, m, dim, _ptr(R), _ptr(pd_of_eval), _ptr(v1),
                                                  _ptr(v2), ne, int(bins), float(lo), float(hi), _ptr(nz), _ptr(H)))
        return (nz, H) if want_hist else nz

    # -- data preparation / formats ---------------------------------------------
    def tau_snr_stack(self, pristine, tau, snr, seed=0, download=True):
        """Generate_TauSNR.py:56-64 + AddNoise.py:17-86 on the device: float32 [ss, box, box] pristine states ->
        (stack float32 [ss*tau, box, box] or None, params float64 [ss, 2] = (signal mean, noise std)).
        snr=None/0 -> tau exact copies.  The stack stays resident for dist_rms(None, n=ss*tau, P=box*box)."""
        pristine = np.ascontiguousarray(pristine, dtype=np.float32)
        if pristine.ndim != 3 or pristine.shape[1] != pristine.shape[2]:
            raise ValueError('tau_snr_stack: pristine must be [ss, box, box]')
        ss, box, _ = pristine.shape
        tau = int(tau)
        params = np.empty((ss, 2), dtype=np.float64)
        stack = np.empty((ss * max(tau, 0), box, box), dtype=np.float32) if download else None
        self._check(self._lib.esper_tau_snr_stack(self._h, _ptr(pristine), ss, box, tau, float(snr or 0.0),
                                                  C.c_ulonglong(int(seed) & 0xFFFFFFFFFFFFFFFF), _ptr(params), _ptr(stack)))
        return stack, params

    def download_images(self, n, P):
        out = np.empty((int(n), int(P)), dtype=np.float32)
        self._check(self._lib.esper_download_images(self._h, _ptr(out), int(n), int(P)))
        return out

    def upload_mrc(self, path, first=0, count=0):
        """Stream a mode-2 MRC stack from disk into the resident image buffer -> (images, ny, nx)."""
        dims = (C.c_int32 * 3)()
        self._check(self._lib.esper_upload_mrc(self._h, os.fsencode(path), int(first), int(count), dims))
        return int(dims[0]), int(dims[1]), int(dims[2])

    def bin_accumulate(self, imgs, order, offsets, n=None, P=None):
        """Manifold_Binning.py:940-974: float64 sums [n_bins, P] of the images order[offsets[b]:offsets[b+1]], in order."""
        if imgs is not None:
            imgs = _check_stack(imgs)
            n, P = imgs.shape
        if n is None or P is None:
            raise ValueError('bin_accumulate: n and P are required when imgs is None')
        order = _as(order, np.int32).reshape(-1)
        offsets = _as(offsets, np.int32).reshape(-1)
        if offsets.shape[0] < 2 or offsets[-1] != order.shape[0]:
            raise ValueError('bin_accumulate: offsets must be [n_bins + 1] with offsets[-1] == len(order)')
        sums = np.empty((offsets.shape[0] - 1, int(P)), dtype=np.float64)
        self._check(self._lib.esper_bin_accumulate(self._h, _ptr(imgs), int(n), int(P), _ptr(order), _ptr(offsets),
                                                   offsets.shape[0] - 1, _ptr(sums)))
        return sums

    def manifold_prepare(self, U0, dim, first, kind=2, want_coef=False):
        """Rotation_Histograms.py:100-112 (normalize + slice) and the conic fits of findParabolas (:153-212) for a
        batch of manifolds U0 [n_pd, m, ncol] -> (Uinit [n_pd, m, dim], R2 [n_pd, pairs] or None[, coef]).  The
        normalised manifolds stay resident for rot_sweep(None, ..., shape=...) / conic_fit_batch(None, ...)."""
        U0 = _as(U0, np.float64)
        if U0.ndim == 2:
            U0 = U0[None]
        n_pd, m, ncol = U0.shape
        dim = int(dim)
        n_pairs = dim * (dim - 1) // 2
        Ui = np.empty((n_pd, m, dim), dtype=np.float64)
        R2 = np.empty((n_pd, n_pairs), dtype=np.float64) if kind else None
        coef = np.empty((n_pd, n_pairs, 5), dtype=np.float64) if (kind and want_coef) else None
        self.rot_token = None
        self._check(self._lib.esper_manifold_prepare(self._h, _ptr(U0), n_pd, m, ncol, int(first), dim, int(kind),
                                                     _ptr(Ui), _ptr(R2), _ptr(coef)))
        self.rot_token = object()
        return (Ui, R2, coef) if want_coef else (Ui, R2)

    def conic_fit_batch(self, Uinit, kind=2, shape=None, want_coef=False):
        """ConicFit.fit2 (kind 2) / fit3 (kind 3) for every column pair of every manifold -> R2 [n_pd, pairs]."""
        if Uinit is not None:
            Uinit = _as(Uinit, np.float64)
            if Uinit.ndim == 2:
                Uinit = Uinit[None]
            shape = Uinit.shape
            self.rot_token = None
        n_pd, m, dim = (int(x) for x in shape)
        n_pairs = dim * (dim - 1) // 2
        R2 = np.empty((n_pd, n_pairs), dtype=np.float64)
        coef = np.empty((n_pd, n_pairs, 5), dtype=np.float64) if want_coef else None
        self._check(self._lib.esper_conic_fit_batch(self._h, _ptr(Uinit), n_pd, m, dim, int(kind), _ptr(R2), _ptr(coef)))
        return (R2, coef) if want_coef else R2

    def rot_sweep(self, Uinit, combo, n_sub, rij, theta_grid, bins=51, lo=-1.5, hi=1.5, want_counts=False, shape=None):
        """Uinit=None with shape=(n_pd, m, dim): the manifolds left resident by manifold_prepare."""
        if Uinit is not None:
            Uinit = _as(Uinit, np.float64)
            shape = Uinit.shape
            self.rot_token = None
        n_pd, m, dim = (int(x) for x in shape)
        combo = _as(combo, np.int32)
        max_sub = combo.shape[1]
        n_sub = _as(n_sub, np.int32).reshape(-1)
        rij = _as(rij, np.int32).reshape(n_pd, -1)
        n_rij = rij.shape[1]
        theta_grid = _as(theta_grid, np.float64).reshape(-1)
        T = dim * (dim - 1) // 2
        if combo.shape != (n_pd, max_sub, 2) or n_sub.shape[0] != n_pd:
            raise ValueError('rot_sweep: combo must be [n_pd, max_sub, 2] and n_sub [n_pd]')
        out = np.empty((n_pd, max_sub, T), dtype=np.float64)
        cnt = np.empty((n_pd, max_sub, max(n_rij, 1), theta_grid.shape[0]), dtype=np.int32) if want_counts else None
        self._check(self._lib.esper_rot_sweep(self._h, _ptr(Uinit), n_pd, m, dim, _ptr(combo), _ptr(n_sub), max_sub,
                                              _ptr(rij), n_rij, _ptr(theta_grid), theta_grid.shape[0], int(bins),
                                              float(lo), float(hi), _ptr(out), _ptr(cnt)))
        return (out, cnt) if want_counts else out


def _check_stack(imgs):
    imgs = np.asarray(imgs)
    if imgs.ndim == 3:
        imgs = imgs.reshape(imgs.shape[0], -1)
    if imgs.ndim != 2 or imgs.shape[0] < 1 or imgs.shape[1] < 1:
        raise ValueError('image stack must be [n, N, N] or [n, P] with n >= 1')
    return _as(imgs, np.float32)


def _check_square(D):
    D = np.asarray(D)
    if D.ndim != 2 or D.shape[0] != D.shape[1] or D.shape[0] < 1:
        raise ValueError('distance matrix must be square [m, m] with m >= 1')
    return _as(D, np.float64)


def tridiag_ritz(alpha, beta, k, tol=0.0):
    """Host only: (converged, theta[k-1], S[steps, k-1]) of T = tridiag(beta, alpha, beta) (esper_tridiag_ritz)."""
    lib = load_library()
    alpha, beta = _as(alpha, np.float64).reshape(-1), _as(beta, np.float64).reshape(-1)
    steps = alpha.shape[0]
    if beta.shape[0] != steps or k < 2:
        raise ValueError('tridiag_ritz: alpha and beta must have one entry per step and k >= 2')
    theta = np.zeros(k - 1); S = np.zeros((steps, k - 1)); conv = C.c_int(0)
    rc = lib.esper_tridiag_ritz(_ptr(alpha), _ptr(beta), steps, int(k), float(tol), C.byref(conv), _ptr(theta), _ptr(S))
    if rc != ESPER_OK:
        raise ValueError('tridiag_ritz: bad argument')
    return bool(conv.value), theta, S


def tanh_fit(x, y, a0):
    """Host only: esper_tanh_fit, one curve_fit(fun, x, y, p0=a0) of GaussianBandwidth.py:47-57 ->
    (popt[4], resnorm, R_squared, info, nfev)."""
    lib = load_library()
    x, y = _as(x, np.float64).reshape(-1), _as(y, np.float64).reshape(-1)
    a0 = _as(a0, np.float64).reshape(-1)
    if x.shape != y.shape or a0.shape[0] != 4:
        raise ValueError('tanh_fit: x and y must have the same length and a0 four entries')
    popt = np.empty(4)
    rn, r2, info, nfev = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
    rc = lib.esper_tanh_fit(_ptr(x), _ptr(y), x.shape[0], _ptr(a0), _ptr(popt), C.byref(rn), C.byref(r2), C.byref(info), C.byref(nfev))
    if rc != ESPER_OK:
        raise ValueError('tanh_fit: array must not contain infs or NaNs (and needs at least 4 points)')
    return popt, rn.value, r2.value, info.value, nfev.value


def tridiag_eigen_all(alpha, beta):
    """Host only: (theta[n] descending, Z[n, n]) = all eigenpairs of T = tridiag(beta[:-1], alpha, beta[:-1]) (esper_tridiag_eigen_all)."""
    lib = load_library()
    alpha, beta = _as(alpha, np.float64).reshape(-1), _as(beta, np.float64).reshape(-1)
    n = alpha.shape[0]
    if beta.shape[0] < n:
        beta = np.concatenate([beta, np.zeros(n - beta.shape[0])])
    theta, Z = np.empty(n), np.empty((n, n))
    rc = lib.esper_tridiag_eigen_all(_ptr(alpha), _ptr(beta), n, _ptr(theta), _ptr(Z))
    if rc != ESPER_OK:
        raise EsperError('tridiag_eigen_all: QL iteration did not converge')
    return theta, Z


def nd_rotation(n, theta):
    """Host-only: esper_nd_rotation (Generate_Nd_Rot.py:4-45)."""
    lib = load_library()
    theta = _as(theta, np.float64).reshape(-1)
    R = np.empty((n, n), dtype=np.float64)
    rc = lib.esper_nd_rotation(int(n), _ptr(theta), theta.shape[0], _ptr(R))
    if rc != ESPER_OK:
        raise ValueError('%s angles required' % (int(n * (n - 1) / 2)))
    return R


def mrc_info(path):
    """(nx, ny, nz), mode, data offset of an MRC2014 file (host only; what mrcfile.mmap reports)."""
    lib = load_library()
    dims = (C.c_int32 * 3)()
    mode = C.c_int32(0)
    off = C.c_longlong(0)
    rc = lib.esper_mrc_info(os.fsencode(path), dims, C.byref(mode), C.byref(off))
    if rc != ESPER_OK:
        raise ValueError(lib.esper_last_error(None).decode())
    return (int(dims[0]), int(dims[1]), int(dims[2])), int(mode.value), int(off.value)


def hist_edges(bins, lo, hi):
    lib = load_library()
    e = np.empty(bins + 1, dtype=np.float64)
    if lib.esper_hist_edges(int(bins), float(lo), float(hi), _ptr(e)) != ESPER_OK:
        raise ValueError('bad histogram specification')
    return e


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context per device for the module-level reference-style functions."""
    key = (os.getpid(), threading.get_ident(), device)
    ctx = _default_ctx.get(key)
    if ctx is None:
        ctx = _default_ctx[key] = Context(device)
    return ctx
