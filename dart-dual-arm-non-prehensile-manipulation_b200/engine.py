"""Handle wrapper over the C ABI: one batched NMPC engine per (method, device)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DartCfg, check


_BYTE = C.c_char


def _p(a):
    """Address of a numpy array's data.  ``a.ctypes.data`` builds a helper object per call (2.6 us each, nine arguments per
    solve: a fifth of the host API's time at the headline batch); the buffer protocol gives the same address in 0.4 us."""
    if a is None:
        return None
    try:
        return C.addressof(_BYTE.from_buffer(a))
    except (TypeError, ValueError):          # read-only or empty array
        return a.ctypes.data


def _f64(a, shape, name):
    if a is None:
        return None
    if type(a) is np.ndarray and a.dtype == np.float64 and a.shape == shape and a.flags.c_contiguous:
        return a                             # the usual case: no conversion, no copy
    a = np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64)
    if a.shape != shape:
        raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
    return a


class _Out(dict):
    """Result dict of ``NMPCEngine.solve`` (u0, J, w, status, iters) that remembers the addresses of its arrays."""
    __slots__ = ("ptrs",)


class NMPCEngine:
    """Batched solve of B independent tray-tilt NLPs on one GPU.

    Host entry (``solve``) mirrors the data the reference passes to ``self.solver(x0=, p=, ...)``;
    device entry (``solve_device``) takes torch CUDA tensors and is asynchronous on the current stream.
    """

    def __init__(self, cfg: DartCfg, device: int = 0):
        self._lib = _lib.lib()
        self.cfg = cfg
        self.device = int(device)
        self._h = C.c_void_p()
        check(self._lib.dart_create(C.byref(self._h), C.byref(cfg), self.device), "dart_create")
        self.nx = self._lib.dart_nx(self._h)
        self.nref = self._lib.dart_nref(self._h)
        self.naux = self._lib.dart_naux(self._h)
        self.nw = self._lib.dart_nw(self._h)
        self.N = cfg.N
        self._nvtx_name = "dart_solve/" + {0: "pmpc", 1: "rmpc", 2: "lmpc"}.get(int(cfg.method), "nmpc")

    def set_mu_init(self, mu_init):
        """Initial barrier parameter of the following solves (0 = the strategy's default: 0.1 monotone, 0.01 predictor-corrector);
        see dart_set_mu_init."""
        check(self._lib.dart_set_mu_init(self._h, float(mu_init)), "dart_set_mu_init")

    def set_barrier_strategy(self, strategy):
        """'auto' (default: per method, the strategy measured faster -- predictor-corrector for PMPC / LMPC and for RMPC
        calls without a warm plan, monotone for warm-started RMPC), 'mehrotra' (predictor-corrector wherever the kernel has it) or 'monotone' (IPOPT's default schedule for every
        method); see dart_set_barrier_strategy."""
        code = {"monotone": 0, "mehrotra": 1, "auto": 2}[strategy]
        check(self._lib.dart_set_barrier_strategy(self._h, code), "dart_set_barrier_strategy")

    @property
    def ndual(self):
        return self._lib.dart_ndual(self._h)

    def set_dual_state(self, dual):
        """Register a zero-initialised float64 CUDA tensor [B, ndual] as the dual warm-start state (None unregisters);
        see dart_set_dual_state.  The tensor must outlive the registration."""
        if dual is None:
            check(self._lib.dart_set_dual_state(self._h, None, 0), "dart_set_dual_state")
            self._dual = None
            return
        import torch
        if not (dual.is_cuda and dual.dtype == torch.float64 and dual.is_contiguous() and dual.dim() == 2 and dual.shape[1] == self.ndual):
            raise ValueError(f"dual state must be a contiguous float64 CUDA tensor [B, {self.ndual}]")
        check(self._lib.dart_set_dual_state(self._h, C.c_void_p(dual.data_ptr()), int(dual.shape[0])), "dart_set_dual_state")
        self._dual = dual

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dart_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host arrays
    def make_out(self, B, want_w=True):
        """Reusable result arrays for ``solve(..., out=...)`` (the numpy ``out=`` idiom): a control loop that calls ``solve``
        every cycle saves four allocations and five address look-ups per call (8 of the 95 us of a headline batch).  Each
        ``solve`` overwrites them."""
        o = _Out(u0=np.empty((B, 2)), J=np.empty(B), w=np.empty((B, self.nw)) if want_w else None,
                 status=np.empty(B, dtype=np.int32), iters=np.empty(B, dtype=np.int32))
        o.ptrs = (B, _p(o["w"]), _p(o["u0"]), _p(o["J"]), _p(o["status"]), _p(o["iters"]))
        return o

    def solve(self, x0, ref, aux=None, warm_w=None, want_w=True, out=None):
        if not (type(x0) is np.ndarray and x0.ndim == 2):
            x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        B = x0.shape[0]
        x0 = _f64(x0, (B, self.nx), "x0")
        ref = _f64(ref, (B, self.nref), "ref")
        aux = _f64(aux, (B, self.naux), "aux")
        warm_w = _f64(warm_w, (B, self.nw), "warm_w")
        if out is None:
            out = self.make_out(B, want_w)
        Bo, pw, pu, pJ, ps, pi = out.ptrs
        if Bo != B:
            raise ValueError(f"out= was made for {Bo} instances, this call has {B}")
        check(self._lib.dart_solve_host(self._h, B, _p(x0), _p(ref), _p(aux), _p(warm_w), pw, pu, pJ, ps, pi), "dart_solve_host")
        return out

    # ------------------------------------------------------------------ torch CUDA tensors
    def solve_device(self, x0, ref, aux=None, warm_w=None, w_out=None, u0_out=None, J_out=None, status=None, iters=None):
        import torch
        B = x0.shape[0]

        def ptr(t, shape, dtype, name):
            if t is None:
                return None
            if (not t.is_cuda) or t.device.index != self.device or t.dtype != dtype or not t.is_contiguous() \
                    or tuple(t.shape) != shape:
                raise ValueError(f"{name}: need contiguous {dtype} CUDA:{self.device} tensor of shape {shape}")
            return C.c_void_p(t.data_ptr())

        f64, i32 = torch.float64, torch.int32
        dev = torch.device("cuda", self.device)
        if u0_out is None:
            u0_out = torch.empty((B, 2), dtype=f64, device=dev)
        if J_out is None:
            J_out = torch.empty((B,), dtype=f64, device=dev)
        if status is None:
            status = torch.empty((B,), dtype=i32, device=dev)
        if iters is None:
            iters = torch.empty((B,), dtype=i32, device=dev)
        # dart_solve launches on the CURRENT device and rejects a mismatch with the handle's (DART_ERR_ARG): run under the
        # handle's device and restore the caller's afterwards
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.nvtx.range_push(self._nvtx_name)          # NVTX range around the solve launch (tracing)
            try:
                return self._solve_device(B, ptr, x0, ref, aux, warm_w, w_out, u0_out, J_out, status, iters, stream, f64, i32)
            finally:
                torch.cuda.nvtx.range_pop()

    def _solve_device(self, B, ptr, x0, ref, aux, warm_w, w_out, u0_out, J_out, status, iters, stream, f64, i32):
        check(self._lib.dart_solve(self._h, B, ptr(x0, (B, self.nx), f64, "x0"), ptr(ref, (B, self.nref), f64, "ref"),
                                   ptr(aux, (B, self.naux), f64, "aux"), ptr(warm_w, (B, self.nw), f64, "warm_w"),
                                   ptr(w_out, (B, self.nw), f64, "w_out"), ptr(u0_out, (B, 2), f64, "u0_out"),
                                   ptr(J_out, (B,), f64, "J_out"), ptr(status, (B,), i32, "status"),
                                   ptr(iters, (B,), i32, "iters"), stream), "dart_solve")
        return dict(u0=u0_out, J=J_out, w=w_out, status=status, iters=iters)

    def set_result_rows(self, rows):
        """rows: CUDA float64 tensor [B,4] (or None); later solves also write [u0x, u0y, J, status] rows into it."""
        if rows is not None and (rows.dim() != 2 or rows.shape[1] != 4 or not rows.is_contiguous() or str(rows.dtype) != "torch.float64"):
            raise ValueError("result rows: need a contiguous float64 CUDA tensor [capacity, 4]")
        self._rows = rows
        check(self._lib.dart_set_result_rows(self._h, None if rows is None else C.c_void_p(rows.data_ptr()),
                                             0 if rows is None else int(rows.shape[0])), "dart_set_result_rows")

    @property
    def launch_count(self):
        return int(self._lib.dart_launch_count(self._h))

    def last_launch_config(self):
        v = [C.c_int32() for _ in range(4)]
        check(self._lib.dart_last_launch_config(self._h, *[C.byref(x) for x in v]), "dart_last_launch_config")
        return dict(lanes=v[0].value, block_threads=v[1].value, grid=v[2].value, smem_bytes=v[3].value)


def tilt_to_quat_device(u, quat=None):
    """Device epilogue: tilt command [B,2] -> tray quaternion wxyz [B,4] (PMPC/main.py:107-116)."""
    import torch
    B = u.shape[0]
    if quat is None:
        quat = torch.empty((B, 4), dtype=torch.float64, device=u.device)
    stream = C.c_void_p(torch.cuda.current_stream(u.device).cuda_stream)
    check(_lib.lib().dart_tilt_to_quat(B, C.c_void_p(u.data_ptr()), C.c_void_p(quat.data_ptr()), stream), "dart_tilt_to_quat")
    return quat


def measure_fp64_tflops(device=0):
    """FP64 FMA-pipe peak of the device (TFLOP/s) from the library's DFMA microbenchmark."""
    v = C.c_double()
    check(_lib.lib().dart_measure_fp64_tflops(int(device), C.byref(v)), "dart_measure_fp64_tflops")
    return v.value
