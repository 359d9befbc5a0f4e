"""Engine: one C-ABI handle (= one B200 + one CUDA stream) as a Python object.  Pure marshalling -- every
method is one call into libgamc_b200.so; arrays are NumPy float64, matrices column-major (Fortran order),
the layout of Julia's `Matrix{Float64}`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=[order + "_CONTIGUOUS", "ALIGNED"])


class Engine:
    def __init__(self, device: int = 0):
        self.lib = L.load()
        self.h = C.c_void_p()
        st = self.lib.gamc_create(C.byref(self.h), int(device))
        if st != L.GAMC_OK:
            raise L.CudaError(f"gamc_create(device={device}) failed with status {st}: no usable sm_100 CUDA device "
                              "(the product has no CPU fallback)")
        self.device = int(device)
        self.p = None
        self.N = None

    # -- lifecycle
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.gamc_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        L.check(st, self.h)

    def version(self):
        return self.lib.gamc_version().decode()

    def sync(self):
        self._ck(self.lib.gamc_sync(self.h))

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self.lib.gamc_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # -- target
    def target_logistic(self, X, y, lam):
        X = _f64(X, "F")
        y = _f64(y, "C").ravel()
        N, p = X.shape
        assert y.shape[0] == N
        self._ck(self.lib.gamc_target_logistic(self.h, L.dptr(X), N, L.dptr(y), N, p, float(lam)))
        self.N, self.p = N, p

    def target_logistic_device(self, dX_ptr: int, ldX: int, dy_ptr: int, N: int, p: int, lam: float):
        self._ck(self.lib.gamc_target_logistic_device(self.h, C.c_void_p(dX_ptr), ldX, C.c_void_p(dy_ptr), N, p, float(lam)))
        self.N, self.p = N, p

    def target_logistic_synth(self, seed, row0, nrows, p, theta_true, lam):
        tt = _f64(theta_true, "C").ravel()
        assert tt.shape[0] == p
        self._ck(self.lib.gamc_target_logistic_synth(self.h, int(seed), int(row0), int(nrows), int(p), L.dptr(tt), float(lam)))
        self.N, self.p = int(nrows), int(p)

    def target_callback(self, p, fn):
        """Generic target: `fn(theta, level) -> (logtarget, grad | None, G | None)` evaluated on the host for
        levels 0 (log-target), 1 (+ gradient), 2 (+ metric tensor, already transformed).  The wrapper is kept alive here."""
        p = int(p)

        def _cb(theta_ptr, p_, level, lt_ptr, grad_ptr, G_ptr, _user):
            try:
                theta = np.ctypeslib.as_array(theta_ptr, shape=(p_,)).copy()
                lt, grad, G = fn(theta, int(level))
                lt_ptr[0] = float(lt)
                if level >= 1:
                    np.ctypeslib.as_array(grad_ptr, shape=(p_,))[:] = np.asarray(grad, dtype=np.float64).ravel()
                if level >= 2:
                    np.ctypeslib.as_array(G_ptr, shape=(p_ * p_,))[:] = np.asarray(G, dtype=np.float64).ravel(order="F")
                return 0
            except Exception:   # noqa: BLE001 -- an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        self._target_cb = L.TARGET_FN(_cb)
        self._ck(self.lib.gamc_target_callback(self.h, p, C.cast(self._target_cb, C.c_void_p), None))
        self.N, self.p = 0, p

    def target_device_callback(self, p, enqueue):
        """Generic target evaluated ON THE DEVICE: `enqueue(d_theta_ptr, level, d_out_ptr, stream_ptr)` must only
        enqueue work on the CUDA stream `stream_ptr` that reads p doubles at `d_theta_ptr` and writes [logtarget, grad(p), pad,
        G(p x p column-major)] at `d_out_ptr` (G at offset (p + 2) & ~1); no host synchronisation per evaluation."""
        p = int(p)

        def _cb(d_theta, p_, level, d_out, stream, _user):
            try:
                enqueue(int(d_theta), int(level), int(d_out), int(stream or 0))
                return 0
            except Exception:   # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        self._target_dcb = L.TARGET_ENQUEUE_FN(_cb)
        self._ck(self.lib.gamc_target_device_callback(self.h, p, C.cast(self._target_dcb, C.c_void_p), None))
        self.N, self.p = 0, p

    def read_rows(self, r0, n):
        X = np.empty((n, self.p), dtype=np.float64, order="F")
        y = np.empty(n, dtype=np.float64)
        self._ck(self.lib.gamc_target_read_rows(self.h, int(r0), int(n), L.dptr(X), L.dptr(y)))
        return X, y

    # -- evaluation
    def eval_logtarget(self, theta):
        th = _f64(theta, "C").ravel()
        out = C.c_double()
        self._ck(self.lib.gamc_eval_logtarget(self.h, L.dptr(th), C.byref(out)))
        return out.value

    def eval_upto_tensor(self, theta):
        th = _f64(theta, "C").ravel()
        lt = C.c_double()
        g = np.empty(self.p)
        G = np.empty((self.p, self.p), order="F")
        self._ck(self.lib.gamc_eval_upto_tensor(self.h, L.dptr(th), C.byref(lt), L.dptr(g), L.dptr(G)))
        return lt.value, g, G

    def eval_partials(self, theta, want_grad=True, want_tensor=True):
        th = _f64(theta, "C").ravel()
        ll = C.c_double()
        g = np.empty(self.p) if (want_grad or want_tensor) else None
        G = np.empty((self.p, self.p), order="F") if want_tensor else None
        self._ck(self.lib.gamc_eval_partials(self.h, L.dptr(th), C.byref(ll), L.dptr(g), L.dptr(G)))
        return ll.value, g, G

    # -- communicator
    @staticmethod
    def comm_unique_id():
        lib = L.load()
        buf = np.zeros(128, dtype=np.uint8)
        L.check(lib.gamc_comm_unique_id(L.u8ptr(buf)))
        return buf

    def comm_init(self, uid, nranks, rank):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._ck(self.lib.gamc_comm_init(self.h, L.u8ptr(uid), int(nranks), int(rank)))

    def comm_peer_path(self):
        """Bit mask of the reductions that run over NVLink peer memory instead of NCCL: 1 = the scalar of AM steps
        (inside k_am_post), 2 = the gradient/metric payload of SMMALA steps (inside k_land; known after sampler_init)."""
        out = C.c_int32(0)
        self._ck(self.lib.gamc_comm_peer_path(self.h, C.byref(out)))
        return int(out.value)

    def set_metric_route(self, route="auto", slices=0, smax=0):
        """Route of the logistic metric for targets set up after this call: "auto", "f64" (FP64 DMMA) or "i8" (exact int8 digit
        products on tcgen05; `slices` digit planes, products with a + b <= smax)."""
        code = {"auto": 0, "f64": 1, "i8": 2}[route] if isinstance(route, str) else int(route)
        self._ck(self.lib.gamc_set_metric_route(self.h, code, int(slices), int(smax)))

    def metric_route(self):
        """("f64" | "i8", slices, smax) of the current target."""
        r, k, s = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.gamc_get_metric_route(self.h, C.byref(r), C.byref(k), C.byref(s)))
        return ("i8" if r.value == 2 else "f64"), int(k.value), int(s.value)

    def comm_set_timeout(self, seconds):
        self._ck(self.lib.gamc_comm_set_timeout(self.h, float(seconds)))

    # -- sampler
    def set_schedule_callback(self, fn):
        """`GAMC.update!::Function`: fn(i, tot, u) -> bool (True = SMMALA step); None restores the configured schedule."""
        if fn is None:
            self._sched_cb = None
            self._ck(self.lib.gamc_set_schedule_callback(self.h, None, None))
            return

        def _cb(i, tot, u, _user):
            try:
                return 1 if fn(int(i), int(tot), float(u)) else 0
            except Exception:   # noqa: BLE001
                import traceback
                traceback.print_exc()
                return 1

        self._sched_cb = L.SCHEDULE_FN(_cb)
        self._ck(self.lib.gamc_set_schedule_callback(self.h, C.cast(self._sched_cb, C.c_void_p), None))

    def set_state(self, st: L.StateScalars, theta, lastmean, secondlastmean, oldinvtensor=None):
        th, m1, m2 = (_f64(a, "C").ravel() for a in (theta, lastmean, secondlastmean))
        Cm = None if oldinvtensor is None else _f64(oldinvtensor, "F")
        self._ck(self.lib.gamc_set_state(self.h, C.byref(st), L.dptr(th), L.dptr(m1), L.dptr(m2), L.dptr(Cm)))

    def monitor(self):
        """(pstate.loglikelihood, pstate.logprior) of the current point."""
        ll, lp = C.c_double(), C.c_double()
        self._ck(self.lib.gamc_get_monitor(self.h, C.byref(ll), C.byref(lp)))
        return ll.value, lp.value

    def chain_get_logtarget(self, first, n):
        out = np.empty(n, dtype=np.float64)
        self._ck(self.lib.gamc_chain_get_logtarget(self.h, int(first), int(n), L.dptr(out)))
        return out

    def tune_log(self):
        """Records of the burn-in report, one row of 16 doubles per tuning event (see gamc_tune_log_get)."""
        n = C.c_int64()
        self._ck(self.lib.gamc_tune_log_get(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 16), dtype=np.float64)
        self._ck(self.lib.gamc_tune_log_get(self.h, L.dptr(out), n.value, C.byref(n)))
        return out[: n.value]

    def probe_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.lib.gamc_probe_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def sampler_init(self, cfg: L.Config, theta0):
        th = _f64(theta0, "C").ravel()
        assert th.shape[0] == self.p
        self._cfg = cfg
        self._ck(self.lib.gamc_sampler_init(self.h, C.byref(cfg), L.dptr(th)))

    def tape_set(self, u_sched, u_mix, z, u_acc):
        """z is (n, p) C-order == p x n column-major."""
        us, um, ua = _f64(u_sched, "C"), _f64(u_mix, "C"), _f64(u_acc, "C")
        zz = _f64(z, "C")
        n = us.shape[0]
        assert zz.shape == (n, self.p)
        self._ck(self.lib.gamc_tape_set(self.h, n, L.dptr(us), L.dptr(um), L.dptr(zz), L.dptr(ua)))

    def tape_generate(self, n, seed):
        self._ck(self.lib.gamc_tape_generate(self.h, int(n), int(seed)))

    def tape_get(self, first, n):
        """(u_sched, u_mix, z, u_acc) of tape slots [first, first+n); z is (n, p)."""
        us, um, ua = np.empty(n), np.empty(n), np.empty(n)
        z = np.empty((n, self.p))
        self._ck(self.lib.gamc_tape_get(self.h, int(first), int(n), L.dptr(us), L.dptr(um), L.dptr(z), L.dptr(ua)))
        return us, um, z, ua

    def run(self, n):
        self._ck(self.lib.gamc_run(self.h, int(n)))

    def chain_get(self, first, n):
        value = np.empty((self.p, n), dtype=np.float64, order="F")
        accept = np.empty(n, dtype=np.uint8)
        self._ck(self.lib.gamc_chain_get(self.h, int(first), int(n), L.dptr(value), L.u8ptr(accept)))
        return value, accept.astype(bool)

    def iterate(self, u_sched, u_mix, z, u_acc):
        us, um, ua = _f64(u_sched, "C"), _f64(u_mix, "C"), _f64(u_acc, "C")
        zz = _f64(z, "C")
        n = us.shape[0]
        value = np.empty((self.p, n), dtype=np.float64, order="F")
        accept = np.empty(n, dtype=np.uint8)
        nsaved = C.c_int64()
        self._ck(self.lib.gamc_iterate(self.h, n, L.dptr(us), L.dptr(um), L.dptr(zz), L.dptr(ua), L.dptr(value),
                                       L.u8ptr(accept), C.byref(nsaved)))
        k = nsaved.value
        return value[:, :k], accept[:k].astype(bool)

    def peek_kinds(self, n):
        out = np.empty(n, dtype=np.uint8)
        self._ck(self.lib.gamc_peek_kinds(self.h, int(n), L.u8ptr(out)))
        return out

    def state(self) -> L.StateScalars:
        st = L.StateScalars()
        self._ck(self.lib.gamc_get_state(self.h, C.byref(st)))
        return st

    def array(self, array_id):
        is_mat = array_id in (L.ARR_TENSORLOGTARGET, L.ARR_OLDINVTENSOR, L.ARR_CHOLINVTENSOR, L.ARR_TENSORLOGLIKELIHOOD,
                              L.ARR_TENSORLOGPRIOR)
        out = np.empty(self.p * self.p if is_mat else self.p, dtype=np.float64)
        self._ck(self.lib.gamc_get_array(self.h, int(array_id), L.dptr(out)))
        return out.reshape((self.p, self.p), order="F") if is_mat else out

    # -- batched chains (small p)
    def batched_target_mvt(self, Sinv, nu, logconst):
        S = _f64(Sinv, "C")
        n = S.shape[0]
        self._ck(self.lib.gamc_batched_target_mvt(self.h, n, float(nu), L.dptr(S), float(logconst)))
        self.bp = n

    def batched_target_rv(self, t, obs, sigma, nplanets=1):
        t, obs, sigma = (_f64(a, "C").ravel() for a in (t, obs, sigma))
        assert t.shape == obs.shape == sigma.shape
        self._ck(self.lib.gamc_batched_target_rv(self.h, L.dptr(t), L.dptr(obs), L.dptr(sigma), t.shape[0], int(nplanets)))
        self.bp = 5 * int(nplanets) + 1

    def batched_target_logistic(self, X, y, lam):
        X = _f64(X, "F")
        y = _f64(y, "C").ravel()
        N, p = X.shape
        assert y.shape[0] == N
        self._ck(self.lib.gamc_batched_target_logistic(self.h, L.dptr(X), N, L.dptr(y), N, p, float(lam)))
        self.bp = p

    def batched_set_chain_offset(self, offset):
        self._ck(self.lib.gamc_batched_set_chain_offset(self.h, int(offset)))

    def batched_init(self, cfg: L.Config, theta0, transform=0, softabs_alpha=1000.0):
        """theta0: (nchains, p) array (row c = initial value of chain c)."""
        th = _f64(theta0, "C")
        assert th.ndim == 2 and th.shape[1] == self.bp
        self.bn = th.shape[0]
        self.bcap = max(0, int(cfg.nsteps) - int(cfg.burnin))
        self._ck(self.lib.gamc_batched_init(self.h, C.byref(cfg), self.bn, L.dptr(th), int(transform), float(softabs_alpha)))

    def batched_run(self, n, seed=0, tape=None):
        """tape: None (device Philox) or dict with u_sched/u_mix/u_acc of shape (nchains, n) and z of shape (nchains, n, p)."""
        if tape is None:
            self._ck(self.lib.gamc_batched_run(self.h, int(n), int(seed), None, None, None, None))
        else:
            us, um, ua, z = (_f64(tape[k], "C") for k in ("u_sched", "u_mix", "u_acc", "z"))
            assert us.shape == (self.bn, n) and z.shape == (self.bn, n, self.bp)
            self._ck(self.lib.gamc_batched_run(self.h, int(n), int(seed), L.dptr(us), L.dptr(um), L.dptr(z), L.dptr(ua)))

    def batched_chain_get(self, first=0, nchains=None):
        nchains = self.bn - first if nchains is None else nchains
        value = np.empty((nchains, self.bcap, self.bp), dtype=np.float64)
        accept = np.empty((nchains, self.bcap), dtype=np.uint8)
        self._ck(self.lib.gamc_batched_chain_get(self.h, int(first), int(nchains), L.dptr(value), L.u8ptr(accept)))
        return value, accept.astype(bool)

    def batched_state(self, chain):
        st = L.StateScalars()
        th = np.empty(self.bp)
        self._ck(self.lib.gamc_batched_get_state(self.h, int(chain), C.byref(st), L.dptr(th)))
        return st, th

    # -- batched chains, large dimension (structured t target)
    def bigd_target_mvt(self, tdiag, toff, nu):
        td, te = _f64(tdiag, "C").ravel(), _f64(toff, "C").ravel()
        n = td.shape[0]
        assert te.shape[0] == max(n - 1, 0)
        self._ck(self.lib.gamc_bigd_target_mvt(self.h, n, float(nu), L.dptr(td), L.dptr(te) if n > 1 else None))
        self.bp = n

    def bigd_set_chain_offset(self, offset):
        self._ck(self.lib.gamc_bigd_set_chain_offset(self.h, int(offset)))

    def bigd_init(self, cfg: L.Config, theta0, transform=1, softabs_alpha=1000.0, max_factors=0):
        th = _f64(theta0, "C")
        assert th.ndim == 2 and th.shape[1] == self.bp
        self.bn = th.shape[0]
        self.bcap = max(0, int(cfg.nsteps) - int(cfg.burnin))
        self._ck(self.lib.gamc_bigd_init(self.h, C.byref(cfg), self.bn, L.dptr(th), int(transform), float(softabs_alpha), int(max_factors)))

    def bigd_run(self, n, seed=0, tape=None):
        if tape is None:
            self._ck(self.lib.gamc_bigd_run(self.h, int(n), int(seed), None, None, None, None))
        else:
            us, um, ua, z = (_f64(tape[k], "C") for k in ("u_sched", "u_mix", "u_acc", "z"))
            assert us.shape == (self.bn, n) and z.shape == (self.bn, n, self.bp)
            self._ck(self.lib.gamc_bigd_run(self.h, int(n), int(seed), L.dptr(us), L.dptr(um), L.dptr(z), L.dptr(ua)))

    def bigd_chain_get(self, first=0, nchains=None):
        nchains = self.bn - first if nchains is None else nchains
        value = np.empty((nchains, self.bcap, self.bp), dtype=np.float64)
        accept = np.empty((nchains, self.bcap), dtype=np.uint8)
        self._ck(self.lib.gamc_bigd_chain_get(self.h, int(first), int(nchains), L.dptr(value), L.u8ptr(accept)))
        return value, accept.astype(bool)

    def bigd_state(self, chain):
        st = L.StateScalars()
        th = np.empty(self.bp)
        self._ck(self.lib.gamc_bigd_get_state(self.h, int(chain), C.byref(st), L.dptr(th)))
        return st, th

    # -- measurement
    def profile_enable(self, on=True):
        self._ck(self.lib.gamc_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self.lib.gamc_profile_reset(self.h))

    def profile_get(self):
        out = {}
        for fam, name in enumerate(L.KERNEL_FAMILIES):
            n = C.c_int64()
            ms = C.c_double()
            self._ck(self.lib.gamc_profile_get(self.h, fam, C.byref(n), C.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.lib.gamc_launch_count(self.h, C.byref(n)))
        return n.value
