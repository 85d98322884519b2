"""GP over a *user-defined* ``Kernel`` subclass.

The reference's ``GP`` accepts any object with the ``Kernel`` interface
(gp/kernels/base.py:7-121): it only ever calls ``K(x1, x2)``, ``K.jacobian``, ``K.hessian``,
``K.params`` and ``K.set_param`` (gp/gp.py:211-238,264,271,276,524,548,572,656).  The two
built-in kernels have fused CUDA functors; for any other kernel the element formulas live in
the user's Python methods, so the kernel matrices are evaluated there -- on the host, exactly
where gp.py evaluates them -- uploaded once, and every O(N^3) / O(N^2) step after that
(Cholesky, solves, inverse, traces, quadratic forms, posterior products) runs through the
same device entry points as the built-in path.  Nothing here computes linear algebra on the
CPU.
"""
import numpy as np

from . import device as D
from ._lib import call
from .engine import Engine, DTYPE


class HostKernelEngine(Engine):
    """Device state of (user kernel, x, y, s).  Same interface as ``Engine``."""

    def __init__(self, kernel, s, x, y):
        self.kernel = kernel
        self.kind = None
        self.hx = np.array(x, dtype=DTYPE)
        self.hy = np.array(y, dtype=DTYPE)
        self.kparams = [float(v) for v in kernel.params]
        self.n_p = len(self.kparams)
        self.s = float(s)
        self.n = int(self.hx.size)
        self.npad = D.roundup(self.n)
        self.T = self.npad // D.NB
        self.dx = D.to_device(self.hx)
        self.dy = D.to_device(self.hy, pad_to=self.npad)
        self.finite = bool(np.isfinite(self.hx).all() and np.isfinite(self.hy).all())
        self._c = {}
        self._ws = None
        self._done = 0

    def rebind(self, kparams, s):
        self.kparams = [float(v) for v in kparams]
        self.s = float(s)
        self.reset()
        return self

    # ------------------------------------------------------------------ uploads
    def _up(self, a, rows, cols, identity_pad=False):
        return D.mat_to_device(np.asarray(a, dtype=DTYPE), rows, cols, identity_pad=identity_pad)

    # Kernel matrices on the device, zero padded to [rows, cols].  A kernel with generated CUDA
    # functors (kernels/symbolic.py: ``device_slices``) builds them there; any other Kernel subclass
    # evaluates them in its own Python methods and they are uploaded.
    def _dev(self, xh):
        return self.dx if xh is self.hx else D.to_device(xh)

    def _mK(self, x1, x2, rows, cols):
        if hasattr(self.kernel, "device_slices"):
            return self.kernel.device_slices(self._dev(x1), x1.size, self._dev(x2), x2.size, rows, cols, 1)[0]
        return self._up(self.kernel(x1, x2), rows, cols)

    def _mJ(self, x1, x2, rows, cols):
        n_p = self.n_p
        if hasattr(self.kernel, "device_slices"):
            J = self.kernel.device_slices(self._dev(x1), x1.size, self._dev(x2), x2.size, rows, cols,
                                          ((1 << n_p) - 1) << 1)
            return [J[i] for i in range(n_p)]
        J = np.asarray(self.kernel.jacobian(x1, x2), dtype=DTYPE)
        return [self._up(J[i], rows, cols) for i in range(n_p)]

    def _mH(self, x1, x2, rows, cols):
        """{(i, j): d2K_ij, j >= i} (the Hessian is symmetric in the parameter pair)."""
        n_p = self.n_p
        pairs = [(i, j) for i in range(n_p) for j in range(i, n_p)]
        if hasattr(self.kernel, "device_slices"):
            mask = sum(1 << (1 + n_p + i * n_p + j) for i, j in pairs)
            H = self.kernel.device_slices(self._dev(x1), x1.size, self._dev(x2), x2.size, rows, cols, mask)
            return {pr: H[q] for q, pr in enumerate(pairs)}          # slice order = ascending bit order
        H = np.asarray(self.kernel.hessian(x1, x2), dtype=DTYPE)
        return {(i, j): self._up(H[i, j], rows, cols) for i, j in pairs}

    def Kxx(self):
        if "K" not in self._c:
            n = self.npad
            if hasattr(self.kernel, "device_slices"):
                K = self.kernel.device_slices(self.dx, self.n, self.dx, self.n, n, n, 1, s2=self.s ** 2,
                                              add_diag=True, pad_identity=True)[0]
                import torch
                finite = bool(torch.isfinite(K).all().item())
            else:
                Kh = np.array(self.kernel(self.hx, self.hx), dtype=DTYPE)          # gp.py:264
                Kh[np.diag_indices(self.n)] += self.s ** 2                         # gp.py:265 (index diagonal)
                finite = bool(np.isfinite(Kh).all())
                K = self._up(Kh, n, n, identity_pad=True) if finite else None
            if not finite:
                raise ValueError("array must not contain infs or NaNs")            # scipy check_finite, gp.py:294
            self._c["K"] = K
        return self._c["K"]

    def _jac(self):
        """dK_i(x, x), i < n_p, zero padded, on the device (gp.py:271)."""
        if "J" not in self._c:
            self._c["J"] = self._mJ(self.hx, self.hx, self.npad, self.npad)
        return self._c["J"]

    # ------------------------------------------------------------------ staged chain
    def _run(self, stages):
        stages |= self.ST_FACTOR
        if stages & self.ST_GRAD:
            stages |= self.ST_LAUUM
        if stages & self.ST_LAUUM:
            stages |= self.ST_TRTRI
        need = stages & ~self._done
        if not need:
            return
        self._workspace()
        n, st, c = self.npad, D.stream_ptr(), self._c
        L, W, V, Ki, a = self._mL, self._mW, self._mV, self._mKi, self._valpha
        if need & self.ST_FACTOR:
            L.copy_(self.Kxx())
            info, out3 = D.izeros(1), D.empty(3)
            z, flags = D.empty(n), D.izeros(2 * self.T + 2)
            call("gpb_potrf", D.ptr(L), n, n, 0, 1, D.ptr(W), n, 0, D.ptr(V), n, 0, D.ptr(info), st)
            call("gpb_potrs", D.ptr(L), D.ptr(W), n, n, n, 0, 0, 1, D.ptr(self.dy), 0, D.ptr(z), D.ptr(a), n,
                 D.ptr(flags), st)
            call("gpb_loglh", D.ptr(L), self.n, n, D.ptr(self.dy), D.ptr(a), D.ptr(info), D.ptr(out3), st)
            h = D.to_host(out3)
            c.update(L=L, W=W, V=V, alpha=a, info=int(info.item()), loglh3=(float(h[0]), float(h[1]), float(h[2])))
        if need & self.ST_TRTRI:
            call("gpb_trtri", D.ptr(L), n, n, 0, 1, D.ptr(W), n, 0, D.ptr(V), n, 0, D.ptr(Ki), n, 0, st)
            c["trtri"] = True
        if need & self.ST_LAUUM:
            call("gpb_lauum", D.ptr(V), n, n, 0, 1, D.ptr(Ki), n, 0, st)
            c["Ki"] = Ki
        if need & self.ST_GRAD:
            # gp_c.pyx:41-49 as O(N^2) reductions over the uploaded Jacobian slices
            raw = np.zeros(16)
            vals = []
            for Ji in self._jac():
                vals.append((self._quad(a, Ji, a), self._trace(Ki, Ji)))
            diag = self.gemv(Ki, self.n, 1, D.ones(1), lda=n + 1)          # Ki[r, r]
            tr, aa = self.dot(diag, D.ones(self.n), self.n), self.dot(a, a, self.n)
            for i, (q, t) in enumerate(vals):
                raw[i], raw[6 + i] = float(q.item()), float(t.item())
            raw[12], raw[13] = float(tr.item()), float(aa.item())
            c["grad_raw"] = raw
        self._done |= need

    def grad_terms(self):
        if self.n_p > 6:
            raise NotImplementedError("kernels with more than 6 parameters")
        return Engine.grad_terms(self)

    # ------------------------------------------------------------------ small device helpers
    def gemv(self, A, rows, cols, x, y=None, alpha=1.0, beta=0.0, lda=None):
        if y is None:
            y = D.empty(rows)
        call("gpb_gemv", D.ptr(A), rows, cols, lda if lda is not None else A.stride(-2), D.ptr(x), D.ptr(y),
             alpha, beta, D.stream_ptr())
        return y

    def _trace(self, A, B):
        """tr(A B) over the leading n x n block."""
        out = D.empty(1)
        call("gpb_trace_prod", D.ptr(A), self.npad, D.ptr(B), self.npad, self.n, D.ptr(self._partial()),
             D.ptr(out), D.stream_ptr())
        return out

    def _quad(self, u, M, v):
        out = D.empty(1)
        call("gpb_quadform", D.ptr(u), D.ptr(M), self.npad, D.ptr(v), self.n, D.ptr(self._partial()),
             D.ptr(out), D.stream_ptr())
        return out

    # ------------------------------------------------------------------ second derivatives
    def d2_terms(self):
        """G, Q, TP, TH of ``Engine.d2_terms`` from uploaded Jacobian / Hessian slices
        (gp_c.pyx:70-111); the N^3 products Ki dK_i run on the DMMA GEMM."""
        if "d2" in self._c:
            return self._c["d2"]
        n, npad, n_p, s = self.n, self.npad, self.n_p, self.s
        nth = n_p + 1
        Ki, a = self.Ki(), self.alpha()
        J = self._jac()
        H = self._mH(self.hx, self.hx, npad, npad)                              # gp.py:276
        B = [self.gemv(J[i], n, n, a, D.zeros(npad)) for i in range(n_p)]       # b_i = dK_i alpha
        C = [self.gemv(Ki, n, n, B[i], D.zeros(npad)) for i in range(n_p)]      # c_i = Ki b_i
        Kia = self.gemv(Ki, n, n, a, D.zeros(npad))
        P = []
        for i in range(n_p):
            Pi = D.empty(npad, npad)
            self.gemm(Ki, J[i], Pi, npad, npad, npad)                            # J symmetric: NT form is Ki dK_i
            P.append(Pi)
        G, Q, TP, TH = (np.zeros((nth, nth)) for _ in range(4))
        dev = {}
        for i in range(n_p):
            for j in range(n_p):
                dev["G", i, j] = self.dot(B[j], C[i], n)
                dev["TP", i, j] = self._trace(P[j], P[i])
                if j >= i:
                    dev["Q", i, j] = self._quad(a, H[i, j], a)
                    dev["TH", i, j] = self._trace(Ki, H[i, j])
            dev["Gs", i] = self.dot(a, C[i], n)
            dev["sG", i] = self.dot(B[i], Kia, n)
            dev["TPs", i] = self._trace(Ki, P[i])
        dev["akia"] = self.dot(a, Kia, n)
        dev["trKK"] = self._trace(Ki, Ki)
        host = {k: float(v.item()) for k, v in dev.items()}
        _, _, tr, aa = self._traces()
        for i in range(n_p):
            for j in range(n_p):
                G[i, j] = host["G", i, j]
                TP[i, j] = host["TP", i, j]
                Q[i, j] = host["Q", min(i, j), max(i, j)]
                TH[i, j] = host["TH", min(i, j), max(i, j)]
            G[i, n_p] = 2 * s * host["Gs", i]
            G[n_p, i] = 2 * s * host["sG", i]
            TP[i, n_p] = TP[n_p, i] = 2 * s * host["TPs", i]
        G[n_p, n_p] = 4 * s * s * host["akia"]
        TP[n_p, n_p] = 4 * s * s * host["trKK"]
        Q[n_p, n_p] = 2.0 * aa
        TH[n_p, n_p] = 2.0 * tr
        self._c["d2"] = (G, Q, TP, TH)
        return self._c["d2"]

    def _traces(self):
        self.grad_terms()
        h = self._c["grad_raw"]
        return None, None, float(h[12]), float(h[13])

    # ------------------------------------------------------------------ posterior
    def mean(self, xo):
        a = self.alpha()
        xo = np.ascontiguousarray(xo, dtype=DTYPE).reshape(-1)
        m = int(xo.size)
        if m == 0:
            return np.empty(0, dtype=DTYPE)
        Kxox = self._mK(xo, self.hx, m, self.npad)                               # gp.py:572
        return D.to_host(self.gemv(Kxox, m, self.n, a)).copy()                   # gp.py:597

    def cov(self, xo, host=True):
        W, _ = self.inv_factor()
        xo = np.ascontiguousarray(xo, dtype=DTYPE).reshape(-1)
        m = int(xo.size)
        if m == 0:
            return np.empty((0, 0), dtype=DTYPE)
        mp = D.roundup(m)
        Kxox = self._mK(xo, self.hx, mp, self.npad)
        Z = D.empty(mp, self.npad)
        self.gemm(Kxox, W, Z, mp, self.npad, self.npad, b_tri=1)
        Cm = self._mK(xo, xo, mp, mp)                                            # gp.py:524
        self.gemm(Z, Z, Cm, mp, mp, self.npad, alpha=-1.0, beta=1.0, lower_only=1, Ct=Cm)
        if not host:
            return Cm[:m, :m]
        return D.download_2d(Cm, m, m)

    def var(self, xo):
        return np.diag(self.cov(xo)).copy()

    def cov_rows(self, xo, lo, hi, host=True):
        return np.ascontiguousarray(self.cov(xo)[int(lo):int(hi)])

    def dm(self, xo):
        """gp_c.dm_dtheta (gp_c.pyx:114-131) with device mat-vecs on uploaded slices."""
        n, npad, n_p = self.n, self.npad, self.n_p
        nth = n_p + 1
        Ki, a = self.Ki(), self.alpha()
        xo = np.ascontiguousarray(xo, dtype=DTYPE).reshape(-1)
        m = int(xo.size)
        out = np.zeros((nth, m))
        if m == 0:
            return out
        J = self._jac()
        Kxox = self._mK(xo, self.hx, m, npad)
        Jxo = self._mJ(xo, self.hx, m, npad)                                     # gp.py:656
        res = D.zeros(nth, m)
        for i in range(n_p):
            b = self.gemv(J[i], n, n, a, D.zeros(npad))                          # dK_i alpha
            c = self.gemv(Ki, n, n, b, D.zeros(npad))                            # Ki dK_i alpha
            self.gemv(Jxo[i], m, n, a, res[i])                                   # dK_i(xo, x) alpha
            self.gemv(Kxox, m, n, c, res[i], alpha=-1.0, beta=1.0)
        kia = self.gemv(Ki, n, n, a, D.zeros(npad))
        self.gemv(Kxox, m, n, kia, res[n_p], alpha=-2.0 * self.s)
        return D.to_host(res).copy()

    def solve_residual(self):
        a = self.alpha()
        r = self.gemv(self.Kxx(), self.n, self.n, a) - self.dy[:self.n]
        return float(r.abs().max().item() / self.dy[:self.n].abs().max().item())
