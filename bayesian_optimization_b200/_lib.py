"""ctypes binding of libb200bo.so (include/b200bo.h).  There is NO CPU fallback: if the shared library
is missing, or no sm_100a device is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200bo.so")

# ids, kept in sync with include/b200bo.h (tests/test_abi.py parses the header and compares)
CORR_RBF, CORR_MATERN12, CORR_MATERN32, CORR_MATERN52, CORR_ABSEXP, CORR_CUBIC, CORR_GENEXP, CORR_MATERN_NU = range(8)
MODE_NOISELESS, MODE_NOISY, MODE_NOISE_ESTIM = range(3)
TREND_CONSTANT, TREND_LINEAR, TREND_QUADRATIC = 0, 1, 2
FIT_OK, FIT_NOT_SPD, FIT_REJECTED = range(3)
ACQ_EI, ACQ_PI, ACQ_UCB, ACQ_MGFI = range(4)
HOST, DEVICE = 0, 1
PREC_FP64, PREC_FAST = 0, 1
STATE_L, STATE_LINV, STATE_GAMMA, STATE_YT, STATE_FT, STATE_RHO, STATE_BETA, STATE_G, STATE_R = range(9)
N_TIMINGS = 12
N_BAND_INFO = 32
BAND_INFO_KEYS = (
    "dy_model", "du_model", "ds_abs_1", "ds_rel_1", "ds_abs_3", "ds_rel_3", "ds_deterministic", "linv_row2_max", "linv_2norm",
    "linv_fro", "linv_row1_max", "gamma_2norm", "f_2norm", "x_sqnorm_max", "sd_r", "dy_cal_1", "ds_cal_1", "dy_cal_3",
    "ds_cal_3", "cal_err_y_1", "cal_err_s_1", "cal_err_y_3", "cal_err_s_3", "dy_used", "ds_used", "ds_abs_used", "ds_rel_used",
    "du_used", "band_err_y", "band_err_s", "band_ratio", "widen")

E_ARG, E_CUDA, E_STATE, E_NODEVICE = -1, -2, -3, -4


class B200BOError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200bo error {code}: {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/b200bo.h declares
SIGNATURES = {
    "b200bo_last_error": (C.c_char_p, []),
    "b200bo_version": (C.c_int, []),
    "b200bo_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "b200bo_destroy": (C.c_int, [C.c_void_p]),
    "b200bo_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200bo_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "b200bo_set_keep_R": (C.c_int, [C.c_void_p, C.c_int]),
    "b200bo_set_fast_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "b200bo_set_fast_products": (C.c_int, [C.c_void_p, C.c_int]),
    "b200bo_set_replay": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "b200bo_set_chol_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "b200bo_debug_oz_syrk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200bo_gradient": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200bo_acq_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double,
                                  C.c_void_p, C.c_void_p]),
    "b200bo_debug_fused_time": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200bo_set_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "b200bo_factor": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_int, C.c_void_p, _dp, _dp, _dp, _ip]),
    "b200bo_factor_restricted": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                           C.c_void_p, _dp, _ip]),
    "b200bo_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _dp, _dp, _dp, _ip]),
    "b200bo_llf_grad_restricted": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200bo_llf_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200bo_get_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "b200bo_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "b200bo_acq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double,
                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200bo_acq_from_moments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                          C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "b200bo_debug_fast_rt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "b200bo_best_pairs_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "b200bo_get_band_info": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200bo_debug_fast_check": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]),
    "b200bo_get_timings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200bo_get_fit_timings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
}

_lib = None


def load_library():
    """dlopen libb200bo.so and declare every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m bayesian_optimization_b200.build` "
            "(there is no CPU fallback for the CUDA path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise B200BOError(rc, load_library().b200bo_last_error().decode())


def _f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a) -> int:
    """host numpy array or (torch) device tensor -> raw address"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())


class Engine:
    """One engine = one b200bo handle = one fitted GP on one GPU."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        _check(self._lib.b200bo_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.N = self.D = 0
        self._torch_stream = None  # address of the torch stream the handle was last bound to (device-pointer calls)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200bo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int):
        _check(self._lib.b200bo_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))
        self._torch_stream = int(cuda_stream_ptr)

    def _follow_torch_stream(self):
        """Device-pointer calls read / write tensors that the caller's CURRENT torch stream produced or will consume:
        run the handle on that stream, so the kernels are ordered with the caller's work (the handle's own stream is
        non-blocking and would race with it)."""
        import torch

        s = int(torch.cuda.current_stream(self.device).cuda_stream)
        if s != self._torch_stream:
            self.set_stream(s)

    def set_precision(self, prec: int):
        _check(self._lib.b200bo_set_precision(self._h, int(prec)))

    def set_fast_kernel(self, generation: int):
        _check(self._lib.b200bo_set_fast_kernel(self._h, int(generation)))

    def set_replay(self, budget_mb: int, max_chunks: int = -1):
        _check(self._lib.b200bo_set_replay(self._h, int(budget_mb), int(max_chunks)))

    def set_fast_products(self, products: int):
        _check(self._lib.b200bo_set_fast_products(self._h, int(products)))

    def set_keep_R(self, keep: bool):
        _check(self._lib.b200bo_set_keep_R(self._h, int(bool(keep))))

    # ---- fit ---------------------------------------------------------------------------------------
    def set_train(self, X: np.ndarray, y: np.ndarray):
        X = _f64(X)
        y = _f64(y).ravel()
        if X.ndim != 2 or y.shape[0] != X.shape[0]:
            raise ValueError("X must be (N, D) and y (N,)")
        self.N, self.D = X.shape
        _check(self._lib.b200bo_set_train(self._h, X.ctypes.data, y.ctypes.data, self.N, self.D))

    def factor(self, corr: int, theta, mode: int, par_last: float = 0.0, noise_var: float = 0.0,
               trend: int = TREND_CONSTANT, beta: Optional[Sequence[float]] = None):
        """-> (llf, sigma2, noise_var, status)"""
        theta = _f64(theta).ravel()
        b = None if beta is None else _f64(beta).ravel()
        llf, s2, nv, st = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        _check(self._lib.b200bo_factor(
            self._h, int(corr), theta.ctypes.data, int(theta.size), int(mode), float(par_last), float(noise_var),
            int(trend), None if b is None else b.ctypes.data, C.byref(llf), C.byref(s2), C.byref(nv), C.byref(st)))
        return llf.value, s2.value, nv.value, st.value

    def factor_restricted(self, corr: int, theta, sigma2: float, noise_var: float = 0.0, trend: int = TREND_CONSTANT,
                          beta: Optional[Sequence[float]] = None):
        """log_likelihood_restricted at (theta, sigma2, noise_var) -> (llf, status)"""
        theta = _f64(theta).ravel()
        b = None if beta is None else _f64(beta).ravel()
        llf, st = C.c_double(), C.c_int()
        _check(self._lib.b200bo_factor_restricted(
            self._h, int(corr), theta.ctypes.data, int(theta.size), float(sigma2), float(noise_var), int(trend),
            None if b is None else b.ctypes.data, C.byref(llf), C.byref(st)))
        return llf.value, st.value

    def append(self, X_new, y_all):
        """m new training points at the parameters of the last factor(): -> (llf, sigma2, noise_var, status)"""
        X_new = _f64(X_new)
        y_all = _f64(y_all).ravel()
        if X_new.ndim != 2 or X_new.shape[1] != self.D or y_all.size != self.N + X_new.shape[0]:
            raise ValueError("X_new must be (m, D) and y_all (N + m,)")
        llf, s2, nv, st = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        _check(self._lib.b200bo_append(self._h, X_new.ctypes.data, int(X_new.shape[0]), y_all.ctypes.data, C.byref(llf),
                                       C.byref(s2), C.byref(nv), C.byref(st)))
        self.N += int(X_new.shape[0])
        return llf.value, s2.value, nv.value, st.value

    def llf_grad_restricted(self, n_par: int) -> np.ndarray:
        g = np.empty(n_par)
        _check(self._lib.b200bo_llf_grad_restricted(self._h, g.ctypes.data, int(n_par)))
        return g

    def llf_grad(self, n_par: int) -> np.ndarray:
        g = np.empty(n_par)
        _check(self._lib.b200bo_llf_grad(self._h, g.ctypes.data, int(n_par)))
        return g

    def state(self, what: int, p: int = 1) -> np.ndarray:
        """p: trend basis size of the last factor() (1 for the constant trend)"""
        N = self.N
        shape = {STATE_L: (N, N), STATE_LINV: (N, N), STATE_R: (N, N), STATE_BETA: (p,), STATE_G: (p, p) if p > 1 else (1,),
                 STATE_FT: (N, p) if p > 1 else (N,)}.get(what, (N,))
        out = np.empty(shape)
        _check(self._lib.b200bo_get_state(self._h, int(what), out.ctypes.data, out.size))
        return out

    # ---- predict / acquisition -------------------------------------------------------------------
    def predict(self, Xc: np.ndarray, eval_mse: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        Xc = _f64(Xc)
        if Xc.ndim != 2 or Xc.shape[1] != self.D:
            raise ValueError("Xc must be (M, D)")
        M = Xc.shape[0]
        yhat = np.empty(M)
        mse = np.empty(M) if eval_mse else None
        _check(self._lib.b200bo_predict(self._h, Xc.ctypes.data, M, HOST, int(eval_mse), yhat.ctypes.data,
                                        None if mse is None else mse.ctypes.data))
        return yhat, mse

    def predict_device(self, Xc, yhat, mse=None):
        """torch CUDA float64 tensors in/out (Xc (M,D) contiguous)."""
        if Xc.ndim != 2 or Xc.shape[1] != self.D:
            raise ValueError("Xc must be (M, D)")
        M = int(Xc.shape[0])
        self._follow_torch_stream()
        _check(self._lib.b200bo_predict(self._h, _ptr(Xc), M, DEVICE, int(mse is not None), _ptr(yhat), _ptr(mse)))

    def acq(self, Xc, acq_id: int, minimize: bool, plugin: float, params, return_values: bool = False,
            device_vals=None):
        """Xc: host ndarray (M,D) or torch CUDA tensor.  -> (best_val (q,), best_idx (q,), vals (q,M) | None)"""
        params = _f64(np.atleast_1d(params if params is not None else [0.0])).ravel()
        q = int(params.size)
        on_dev = not isinstance(Xc, np.ndarray)
        if not on_dev:
            Xc = _f64(Xc)
        if Xc.ndim != 2 or Xc.shape[1] != self.D:
            raise ValueError("Xc must be (M, D)")
        M = int(Xc.shape[0])
        best_val = np.empty(q)
        best_idx = np.empty(q, dtype=np.int64)
        vals = None
        if on_dev:
            vals = device_vals
            self._follow_torch_stream()
        elif return_values:
            vals = np.empty((q, M))
        _check(self._lib.b200bo_acq(self._h, _ptr(Xc), M, DEVICE if on_dev else HOST, int(acq_id), int(bool(minimize)),
                                    float(plugin), params.ctypes.data, q, _ptr(vals), best_val.ctypes.data,
                                    best_idx.ctypes.data))
        return best_val, best_idx, vals

    def best_pairs_device(self, out, index_offset: int, rank: int, world: int):
        """after acq(): this rank's (value bits, global index) pairs into row ``rank`` of the (world, 2q) int64 CUDA
        tensor ``out`` (other rows zeroed), ordered on the handle's stream -- the payload of the one all-reduce"""
        _check(self._lib.b200bo_best_pairs_device(self._h, int(index_offset), int(rank), int(world), C.c_void_p(_ptr(out))))

    def acq_from_moments(self, yhat, mse, acq_id: int, minimize: bool, plugin: float, params,
                         return_values: bool = True):
        params = _f64(np.atleast_1d(params if params is not None else [0.0])).ravel()
        q = int(params.size)
        yhat = _f64(yhat).ravel()
        mse = _f64(mse).ravel()
        M = yhat.size
        best_val = np.empty(q)
        best_idx = np.empty(q, dtype=np.int64)
        vals = np.empty((q, M)) if return_values else None
        _check(self._lib.b200bo_acq_from_moments(self._h, yhat.ctypes.data, mse.ctypes.data, M, HOST, int(acq_id),
                                                 int(bool(minimize)), float(plugin), params.ctypes.data, q,
                                                 _ptr(vals), best_val.ctypes.data, best_idx.ctypes.data))
        return best_val, best_idx, vals

    def debug_fast_rt(self, Xc: np.ndarray):
        """test hook: (rt (M,N) float32, yhat, sum rt^2, Ft^T rt) exactly as the tensor-core kernel produced them"""
        Xc = _f64(Xc)
        M = Xc.shape[0]
        rt = np.empty((M, self.N), dtype=np.float32)
        yh, ss, df = np.empty(M), np.empty(M), np.empty(M)
        _check(self._lib.b200bo_debug_fast_rt(self._h, Xc.ctypes.data, M, rt.ctypes.data, yh.ctypes.data,
                                              ss.ctypes.data, df.ctypes.data))
        return rt, yh, ss, df

    def gradient(self, Xc: np.ndarray):
        """(yhat (M,), mse (M,), y_dx (M, D), mse_dx (M, D)) -- gpr.py:537-576 for every row of Xc"""
        Xc = _f64(Xc)
        if Xc.ndim != 2 or Xc.shape[1] != self.D:
            raise ValueError("Xc must be (M, D)")
        M, D = Xc.shape
        yh, ms = np.empty(M), np.empty(M)
        ydx, mdx = np.empty((M, D)), np.empty((M, D))
        _check(self._lib.b200bo_gradient(self._h, Xc.ctypes.data, M, yh.ctypes.data, ms.ctypes.data, ydx.ctypes.data,
                                         mdx.ctypes.data))
        return yh, ms, ydx, mdx

    def acq_grad(self, Xc: np.ndarray, acq_id: int, minimize: bool, plugin: float, param: float):
        """(value (M,), dx (M, D)) of one acquisition function -- its return_dx=True path for every row of Xc"""
        Xc = _f64(Xc)
        if Xc.ndim != 2 or Xc.shape[1] != self.D:
            raise ValueError("Xc must be (M, D)")
        M, D = Xc.shape
        val, dx = np.empty(M), np.empty((M, D))
        _check(self._lib.b200bo_acq_grad(self._h, Xc.ctypes.data, M, int(acq_id), int(bool(minimize)), float(plugin),
                                         float(param), val.ctypes.data, dx.ctypes.data))
        return val, dx

    def set_chol_tc(self, digits: int, min_rows: int = 0):
        """Cholesky trailing updates on tcgen05 int8 digit planes (7 or 8), 0 = fp64 DMMA"""
        _check(self._lib.b200bo_set_chol_tc(self._h, int(digits), int(min_rows)))

    def debug_oz_syrk(self, P: np.ndarray, Cm: np.ndarray, digits: int = 8, reps: int = 1):
        """C - P P^T on the lower tiles through the tensor-core (digits 7 / 8) or the DMMA (0) kernel; (C', best ms)"""
        P = _f64(P)
        out = np.array(Cm, dtype=np.float64, order="C", copy=True)
        rows = P.shape[0]
        if P.shape != (rows, 64) or out.shape != (rows, rows):
            raise ValueError("P must be (rows, 64) and C (rows, rows)")
        ms = C.c_double(0.0)
        _check(self._lib.b200bo_debug_oz_syrk(self._h, P.ctypes.data, rows, out.ctypes.data, int(digits), int(reps), C.byref(ms)))
        return out, ms.value

    def debug_fused_time(self, Xc: np.ndarray, products: int = 1, reps: int = 3) -> float:
        """average device ms of the fused tensor-core kernel alone (developer hook)"""
        Xc = _f64(Xc)
        out = C.c_double(0.0)
        _check(self._lib.b200bo_debug_fused_time(self._h, Xc.ctypes.data, Xc.shape[0], int(products), int(reps), C.byref(out)))
        return out.value

    def band_info(self) -> dict:
        """half-widths of the arg-max band of the tensor-core path: a-priori model, calibration, last pass (b200bo.h)"""
        t = np.zeros(N_BAND_INFO)
        _check(self._lib.b200bo_get_band_info(self._h, t.ctypes.data, N_BAND_INFO))
        return dict(zip(BAND_INFO_KEYS, t.tolist()))

    def fast_check(self, stride: int = 100, max_samples: int = 1 << 20) -> dict:
        """float64 check of every stride-th candidate of the last tensor-core call (developer / bench hook)"""
        t = np.zeros(8)
        _check(self._lib.b200bo_debug_fast_check(self._h, int(stride), int(max_samples), t.ctypes.data, 8))
        return dict(zip(("checked", "max_err_yhat", "max_err_mse", "max_ratio_to_allowed", "dy", "ds_cal", "ds_model_at_ss1",
                         "products"), t.tolist()))

    def timings(self) -> np.ndarray:
        t = np.zeros(N_TIMINGS)
        _check(self._lib.b200bo_get_timings(self._h, t.ctypes.data, N_TIMINGS))
        return t

    def fit_timings(self) -> np.ndarray:
        t = np.zeros(N_TIMINGS)
        _check(self._lib.b200bo_get_fit_timings(self._h, t.ctypes.data, N_TIMINGS))
        return t
