"""Device-resident state of one GP and every computation of the hot path, expressed as
calls into libgpb200.so on torch-owned HBM buffers.

Nothing here computes on the CPU: torch allocates, ctypes launches, and host values
appear only when a scalar / small vector (or a matrix the user asked for) is read back.
Matrices are padded to a multiple of 128 with an identity pad (DESIGN.md "Data layout"),
so the factor of the padded matrix is [[L, 0], [0, I]] and every pad contribution to
log-determinants, traces and quadratic forms is exactly zero.
"""
import ctypes

import numpy as np
import torch

from . import _lib, device as D
from ._lib import call, darr, iarr, parr

GAUSSIAN, PERIODIC = 0, 1
N_KP = {GAUSSIAN: 2, PERIODIC: 3}
DTYPE = np.float64
MIN = float(np.log(np.exp2(DTYPE(np.finfo(DTYPE).minexp + 4))))   # gp.py:17


def jac_slices(kind):
    return list(range(1, 1 + N_KP[kind]))


def hess_slice(kind, i, j):
    n_p = N_KP[kind]
    return 1 + n_p + i * n_p + j


class Engine(object):
    """One (kernel, theta, s, x, y) on the device; results are cached until dropped."""

    # stages of gpb_gp_stages (include/gpb200.h)
    ST_FACTOR, ST_TRTRI, ST_LAUUM, ST_GRAD = 1, 2, 4, 8

    def __init__(self, kind, kparams, s, x, y):
        self.kind = int(kind)
        self.n_p = N_KP[self.kind]
        self.kparams = [float(v) for v in kparams]
        self.s = float(s)
        self.n = int(np.asarray(x).size)
        self.npad = D.roundup(self.n)
        self.T = self.npad // D.NB
        self.dx = D.to_device(x)
        self.dy = D.to_device(y, pad_to=self.npad)
        self.finite = bool(np.isfinite(np.asarray(x)).all() and np.isfinite(np.asarray(y)).all())
        self._c = {}
        self._ws = None
        self._done = 0

    def rebind(self, kparams, s):
        """New hyperparameters on the same observations: results are dropped, x / y and the
        workspace stay on the device (what an optimiser iteration needs; the reference rebuilds
        everything, gp.py:231-240)."""
        self.kparams = [float(v) for v in kparams]
        self.s = float(s)
        self._c = {k: v for k, v in self._c.items() if k in ("partial",)}
        self._done = 0
        return self

    def reset(self):
        """Forget every result (the workspace stays allocated)."""
        self._c = {}
        self._done = 0

    # ------------------------------------------------------------------ staged chain
    def _workspace(self):
        """The batch-1 evaluator workspace; L / W / V / Ki / alpha are views into it."""
        if self._ws is None:
            off = (ctypes.c_int64 * 16)()
            call("gpb_eval_layout", self.n, off, 16)
            ws = torch.empty(int(off[14]), dtype=torch.uint8, device=D.require_cuda())
            nn = self.npad * self.npad * 8

            def mat(i):
                return ws[int(off[i]):int(off[i]) + nn].view(D.F64).view(self.npad, self.npad)
            self._ws = ws
            self._ws_ptr, self._ws_bytes = ws.data_ptr(), ws.numel()
            self._mL, self._mW, self._mV, self._mKi = mat(0), mat(1), mat(2), mat(3)
            self._valpha = ws[int(off[5]):int(off[5]) + self.npad * 8].view(D.F64)
            self._hout = np.empty(24)
            self._hout_ptr = self._hout.ctypes.data
        return self._ws

    def _run(self, stages):
        """Enqueue the missing stages of the chain in ONE library call with ONE read-back."""
        stages |= self.ST_FACTOR
        if stages & self.ST_GRAD:
            stages |= self.ST_LAUUM
        if stages & self.ST_LAUUM:
            stages |= self.ST_TRTRI
        need = stages & ~self._done
        if not need:
            return
        self._workspace()
        call("gpb_gp_stages", self.kind, darr(self.kparams + [self.s]), D.ptr(self.dx), D.ptr(self.dy),
             self.n, need, self._ws_ptr, self._ws_bytes, self._hout_ptr, D.stream_ptr())
        h = self._hout
        c = self._c
        need |= int(h[20])          # a one-block GP (N <= 128) completes every stage in its two launches
        if need & self.ST_FACTOR:
            c.update(L=self._mL, W=self._mW, V=self._mV, alpha=self._valpha, info=int(h[19]),
                     loglh3=(float(h[0]), float(h[1]), float(h[2])))
        if need & self.ST_TRTRI:
            c["trtri"] = True
        if need & self.ST_LAUUM:
            c["Ki"] = self._mKi
        if need & self.ST_GRAD:
            c["grad_raw"] = h[3:19].copy()
        self._done |= need

    def _ensure(self, stages):
        """Run the chain up to ``stages`` and raise what the reference raises: ValueError for
        non-finite data (scipy check_finite, gp.py:294), LinAlgError when Kxx is not PD."""
        if not self.finite:
            raise ValueError("array must not contain infs or NaNs")
        if self._c.get("info", 0) == 0:         # a failed factorisation is not inverted
            self._run(stages)
        info = self._c["info"]
        if info != 0:
            raise np.linalg.LinAlgError(
                "%d-th leading minor of the array is not positive definite" % info)

    # ------------------------------------------------------------------ helpers
    def _theta(self):
        return darr(self.kparams)

    def _partial(self):
        if "partial" not in self._c:
            self._c["partial"] = D.empty(int(_lib.lib.gpb_grad_partial_doubles(self.n)) + 1024)
        return self._c["partial"]

    def build(self, x1, n1, x2, n2, rows, cols, mask, add_diag=False, pad_identity=False, out=None):
        """Selected slices of the kernel between device vectors x1, x2 -> [nsel, rows, cols]."""
        nsel = bin(mask).count("1")
        if out is None:
            out = D.empty(nsel, rows, cols)
        call("gpb_kernel_build", self.kind, self._theta(), self.s, D.ptr(x1), n1, D.ptr(x2), n2, rows,
             cols, mask, D.ptr(out), cols, rows * cols, int(add_diag), int(pad_identity), D.stream_ptr())
        return out

    def gemm(self, A, B, C, M, N, K, alpha=1.0, beta=0.0, a_tri=0, b_tri=0, lower_only=0, Ct=None):
        call("gpb_gemm_nt", D.ptr(A), A.stride(-2), D.ptr(B), B.stride(-2), D.ptr(C), C.stride(-2),
             D.ptr(Ct), Ct.stride(-2) if Ct is not None else 0, M, N, K, alpha, beta, a_tri, b_tri,
             lower_only, D.stream_ptr())
        return C

    def gemv(self, A, rows, cols, x, y=None, alpha=1.0, beta=0.0):
        if y is None:
            y = D.empty(rows)
        call("gpb_gemv", D.ptr(A), rows, cols, A.stride(-2) if A.dim() > 1 else cols, D.ptr(x), D.ptr(y),
             alpha, beta, D.stream_ptr())
        return y

    def dot(self, u, v, n):
        """u . v via a 1-row GEMV (device scalar tensor)."""
        out = D.empty(1)
        call("gpb_gemv", D.ptr(u), 1, n, n, D.ptr(v), D.ptr(out), 1.0, 0.0, D.stream_ptr())
        return out

    # ------------------------------------------------------------------ Kxx and its factor
    def Kxx(self):
        """Kxx + s^2 I (index diagonal, gp.py:265) on the device, [npad, npad] identity-padded."""
        if "K" not in self._c:
            self._c["K"] = self.build(self.dx, self.n, self.dx, self.n, self.npad, self.npad, 1,
                                      add_diag=True, pad_identity=True)[0]
        return self._c["K"]

    def factor(self):
        """Blocked Cholesky (gp.py:294) + solves + log_lh in one enqueue.  Returns LAPACK-style
        info (0 = ok)."""
        self._run(self.ST_FACTOR)
        return self._c["info"]

    def require_pd(self):
        self._ensure(self.ST_FACTOR)

    def Lxx_host(self):
        self.require_pd()
        L = D.empty(self.n, self.n)          # tril(L): the never-written upper triangle is not read
        call("gpb_tril_copy", D.ptr(L), self.n, D.ptr(self._c["L"]), self.npad, self.n, D.stream_ptr())
        return D.download_2d(L, self.n, self.n)

    def alpha(self):
        """K^-1 y by forward/backward substitution (cho_solve, gp.py:332-334); [npad], pad = 0."""
        self.require_pd()
        return self._c["alpha"]

    def solve(self, rhs):
        """K^-1 rhs for another right-hand side on the cached factor (cho_solve, gp.py:332-334);
        device vector [npad]."""
        self.require_pd()
        n = self.npad
        b = D.to_device(rhs, pad_to=n)
        z, a = D.empty(n), D.empty(n)
        flags = D.izeros(2 * self.T + 2)
        call("gpb_potrs", D.ptr(self._c["L"]), D.ptr(self._c["W"]), n, n, n, 0, 0, 1, D.ptr(b), 0,
             D.ptr(z), D.ptr(a), n, D.ptr(flags), D.stream_ptr())
        return a

    def inv_factor(self):
        """W = L^-1 (lower) and V = L^-T (upper), completed from potrf's diagonal blocks."""
        self._ensure(self.ST_FACTOR | self.ST_TRTRI)
        return self._c["W"], self._c["V"]

    def Ki(self):
        """inv(L)^T inv(L) (gp.py:311-312), full symmetric [npad, npad]."""
        self._ensure(self.ST_FACTOR | self.ST_TRTRI | self.ST_LAUUM)
        return self._c["Ki"]

    # ------------------------------------------------------------------ likelihood
    def loglh3(self):
        """(log_lh, logdet, y.alpha) -- gp_c.log_lh (gp_c.pyx:17-31) with logdet from the Cholesky."""
        self.require_pd()
        return self._c["loglh3"]

    def slice_reduce(self, slices):
        """For each slice S: (alpha^T S alpha, sum(Ki o S)); plus tr(Ki) and alpha.alpha."""
        Ki, a = self.Ki(), self.alpha()
        t0, t1 = [], []
        tr = aa = None
        for lo in range(0, max(len(slices), 1), 6):
            chunk = list(slices[lo:lo + 6])
            out = D.empty(16)
            call("gpb_slice_reduce", self.kind, self._theta(), D.ptr(self.dx), self.n, D.ptr(Ki), self.npad,
                 D.ptr(a), len(chunk), iarr(chunk + [0]), D.ptr(self._partial()), D.ptr(out), D.stream_ptr())
            h = D.to_host(out)
            t0 += list(h[:len(chunk)])
            t1 += list(h[6:6 + len(chunk)])
            tr, aa = float(h[12]), float(h[13])
        return np.array(t0), np.array(t1), tr, aa

    def grad_terms(self):
        """t0[i] = y^T Ki dK_i Ki y, t1[i] = tr(Ki dK_i) for the kernel params and s (gp_c.pyx:41-49).
        Cold, this is the whole chain (factor, solves, inverse, fused reductions) in one call."""
        if "grad_terms" not in self._c:
            self._ensure(self.ST_FACTOR | self.ST_TRTRI | self.ST_LAUUM | self.ST_GRAD)
            h = self._c["grad_raw"]
            n_p = self.n_p
            t0 = np.append(h[:n_p], 2.0 * self.s * h[13])        # dK_s = 2 s I  (gp_c.pyx:45)
            t1 = np.append(h[6:6 + n_p], 2.0 * self.s * h[12])
            self._c["grad_terms"] = (t0, t1)
        return self._c["grad_terms"]

    # ------------------------------------------------------------------ second derivatives
    def d2_terms(self):
        """Everything gp_c.d2lh_dtheta2 (gp_c.pyx:70-111) needs, as [n_theta, n_theta] arrays:
        G[i,j] = b_j . Ki b_i (b_i = dK_i alpha), Q[i,j] = alpha^T d2K_ij alpha,
        TP[i,j] = tr(Ki dK_j Ki dK_i), TH[i,j] = tr(Ki d2K_ij)."""
        if "d2" in self._c:
            return self._c["d2"]
        n, npad, n_p, s = self.n, self.npad, self.n_p, self.s
        nth = n_p + 1
        Ki, a = self.Ki(), self.alpha()
        st = D.stream_ptr()
        # b_i = dK_i alpha (kernel tiles regenerated on the fly); b_s = 2 s alpha
        B = D.zeros(nth, npad)
        sl = jac_slices(self.kind)
        call("gpb_kernel_matvec", self.kind, self._theta(), D.ptr(self.dx), n, D.ptr(self.dx), n, n_p,
             iarr(sl), iarr(range(n_p)), darr([1.0] * n_p), parr([D.ptr(a)] * n_p), n_p,
             parr([B[i].data_ptr() for i in range(n_p)]), st)
        # c_i = Ki b_i
        C = D.zeros(nth, npad)
        for i in range(n_p):
            self.gemv(Ki, n, n, B[i], C[i])
        Kia = self.gemv(Ki, n, n, a)                       # Ki alpha, for the s row
        # every scalar of this method lands in ONE device block: G | traces | dots -> one read-back at the end
        # (was 2 n_p + 3 blocking reads; at the reference's own N <= 50 those were the cost of the call)
        ng, nt_ = nth * nth, nth * nth + 1
        blk = D.zeros(ng + nt_ + 2 * n_p + 1)
        Gd = blk[:ng].view(nth, nth)
        tp = blk[ng:ng + nt_]
        dots = blk[ng + nt_:]

        def dot_into(u, v, slot):
            call("gpb_gemv", D.ptr(u), 1, n, n, D.ptr(v), dots[slot:].data_ptr(), 1.0, 0.0, st)
        for i in range(n_p):
            # column over j of b_j . c_i  (the s row/column is assembled from Kia below)
            call("gpb_gemv", D.ptr(B), n_p, n, npad, D.ptr(C[i]), Gd[i].data_ptr(), 1.0, 0.0, st)
        for i in range(n_p):
            dot_into(a, C[i], i)                           # alpha . c_i
            dot_into(B[i], Kia, n_p + i)                   # b_i . Ki alpha
        dot_into(a, Kia, 2 * n_p)
        # P_i = Ki dK_i (dense products on the DMMA GEMM), then traces of products.  The h-slice of
        # both kernels is proportional to K itself (dK_h = (2/h) K, gaussian_c.pyx:60-69,
        # periodic_c.pyx:65), so P_h = (2/h) (I - s^2 Ki) needs no product: its traces follow from
        # tr Ki, tr Ki Ki, tr P_j and tr Ki P_j.
        ch = 2.0 / self.kparams[0]
        dense = list(range(1, n_p))                        # parameter indices that need a real product
        J = self.build(self.dx, n, self.dx, n, npad, npad, sum(1 << sl[i] for i in dense))
        P = D.empty(len(dense), npad, npad)
        for q, i in enumerate(dense):
            self.gemm(Ki, J[q], P[q], npad, npad, npad)
        part = self._partial()

        def trace_prod(A, Bm, slot):
            call("gpb_trace_prod", D.ptr(A), npad, D.ptr(Bm), npad, n, D.ptr(part), tp[slot:].data_ptr(), st)
        for qi, i in enumerate(dense):
            for qj, j in enumerate(dense):
                trace_prod(P[qj], P[qi], i * nth + j)
            trace_prod(Ki, P[qi], i * nth + n_p)
        trace_prod(Ki, Ki, nth * nth - 1)
        _, t1g = self.grad_terms()                         # t1g[j] = tr(Ki dK_j) = tr P_j
        # Hessian slices: alpha^T H alpha and sum(Ki o H)
        pairs = [(i, j) for i in range(n_p) for j in range(i, n_p)]
        q0, q1, tr, aa = self.slice_reduce([hess_slice(self.kind, i, j) for i, j in pairs])
        blkh = D.to_host(blk).copy()
        Gh = blkh[:ng].reshape(nth, nth)
        tph = blkh[ng:ng + nt_]
        bs_ci, bj_kia, a_kia = blkh[ng + nt_:ng + nt_ + n_p], blkh[ng + nt_ + n_p:ng + nt_ + 2 * n_p], blkh[ng + nt_ + 2 * n_p]
        s2, trKi = s * s, tr
        trKK = tph[nth * nth - 1]
        tph[0 * nth + 0] = ch * ch * (n - 2.0 * s2 * trKi + s2 * s2 * trKK)          # tr(P_h P_h)
        tph[0 * nth + n_p] = ch * (trKi - s2 * trKK)                                 # tr(Ki P_h)
        for j in dense:
            v = ch * (t1g[j] - s2 * tph[j * nth + n_p])                              # tr(P_h P_j)
            tph[0 * nth + j] = tph[j * nth + 0] = v
        G = np.zeros((nth, nth))
        Q = np.zeros((nth, nth))
        TP = np.zeros((nth, nth))
        TH = np.zeros((nth, nth))
        for i in range(n_p):
            for j in range(n_p):
                G[i, j] = Gh[i, j]
                TP[i, j] = tph[i * nth + j]
            G[i, n_p] = 2 * s * float(bs_ci[i])                 # b_s . c_i
            G[n_p, i] = 2 * s * float(bj_kia[i])                # b_i . c_s
            TP[i, n_p] = TP[n_p, i] = 2 * s * tph[i * nth + n_p]
        G[n_p, n_p] = 4 * s * s * float(a_kia)
        TP[n_p, n_p] = 4 * s * s * tph[nth * nth - 1]
        for (i, j), v0, v1 in zip(pairs, q0, q1):
            Q[i, j] = Q[j, i] = v0
            TH[i, j] = TH[j, i] = v1
        Q[n_p, n_p] = 2.0 * aa                                  # d2k = 2 I (gp_c.pyx:99-100)
        TH[n_p, n_p] = 2.0 * tr
        self._c["d2"] = (G, Q, TP, TH)
        return self._c["d2"]

    # ------------------------------------------------------------------ posterior
    def mean(self, xo):
        """K(xo, x) alpha without materialising K(xo, x) (gp.py:597): upload, one fused launch and
        the download in a single library call."""
        a = self.alpha()
        xo = np.ascontiguousarray(xo, dtype=DTYPE).reshape(-1)
        m = int(xo.size)
        out = np.empty(m, dtype=DTYPE)
        if m:
            # up to 2048 points the library reads xo / writes the means in page-locked host memory and
            # needs no device scratch
            scratch = D.empty(2 * D.roundup(m, 32)) if m > 2048 else None
            call("gpb_post_mean_host", self.kind, self._theta(), xo.ctypes.data, m, D.ptr(self.dx), self.n,
                 D.ptr(a), D.ptr(scratch), out.ctypes.data, D.stream_ptr())
        return out

    def _rows_out(self, C, rows, cols, blocks, host=True):
        """Download row blocks of the device matrix ``C`` as they are produced.  ``blocks`` yields
        (r0, r1) after enqueuing the work that completes rows [r0, r1) on the compute stream; each
        block's DMA runs on the copy stream behind an event, overlapping the next block's GEMMs."""
        if not host:                 # device-resident result (timing / chaining): just run the blocks
            for _ in blocks:
                pass
            return C[:rows, :cols]
        out, pinned = D.host_array(rows, cols)
        cs = D.copy_stream() if pinned else None
        for r0, r1 in blocks:
            r1 = min(r1, rows)
            if r1 <= r0:
                continue
            if pinned:
                cs.wait_stream(torch.cuda.current_stream())
                D.download_2d(C, r1 - r0, cols, out=out, pinned=True, r0=r0, stream=cs, sync=False)
            else:
                D.download_2d(C, r1 - r0, cols, out=out, pinned=False, r0=r0)
        if pinned:
            cs.synchronize()
            torch.cuda.current_stream().wait_stream(cs)      # C may be freed / reused after this
        return out

    COV_HOST_MAX_M = 512      # up to here cov() is one gpb_post_cov_host call (result <= 2 MB)

    @staticmethod
    def _panel(mp):
        """Row-panel height for pipelined result downloads: ~8 panels, multiples of 128, >= 1024."""
        return max(1024, D.roundup(mp // 8))

    def cov(self, xo, host=True):
        """K(xo,xo) - K(xo,x) K^-1 K(x,xo) as Kxoxo - Z Z^T with Z = K(xo,x) L^-T (gp.py:599-625).
        Only tiles on or below the diagonal are computed (mirrored store).  Row panels run bottom-up
        so that each panel is final -- its upper part was mirrored by the panels below it -- when its
        own GEMM ends, and its download overlaps the next panel's GEMM."""
        W, _ = self.inv_factor()
        m = int(xo.size)
        if m == 0:
            return np.empty((0, 0), dtype=DTYPE)
        mp = D.roundup(m)
        if host and m <= self.COV_HOST_MAX_M:
            # small result: the four launches and both copies in one library call
            xo = np.ascontiguousarray(xo, dtype=DTYPE).reshape(-1)
            out = np.empty((m, m), dtype=DTYPE)
            nscr = int(_lib.lib.gpb_post_cov_scratch_doubles(m, self.n))      # 0: the one-launch path
            scratch = D.empty(nscr) if nscr else None
            call("gpb_post_cov_host", self.kind, self._theta(), xo.ctypes.data, m, D.ptr(self.dx), self.n,
                 D.ptr(W), self.npad, D.ptr(scratch), out.ctypes.data, m, D.stream_ptr())
            return out
        dxo = D.to_device(xo)
        Kxox = self.build(dxo, m, self.dx, self.n, mp, self.npad, 1)[0]
        Z = D.empty(mp, self.npad)
        self.gemm(Kxox, W, Z, mp, self.npad, self.npad, b_tri=1)
        del Kxox
        C = self.build(dxo, m, dxo, m, mp, mp, 1)[0]
        P = self._panel(mp)

        def blocks():
            for r1 in range(mp, 0, -P):
                r0 = max(0, r1 - P)
                if r0 > 0:      # C[r0:r1, :r0] -= Z[r0:r1] Z[:r0]^T, mirrored into C[:r0, r0:r1]
                    self.gemm(Z[r0:r1], Z[:r0], C[r0:r1, :r0], r1 - r0, r0, self.npad, alpha=-1.0, beta=1.0,
                              Ct=C[:r0, r0:r1])
                Cd = C[r0:r1, r0:r1]
                self.gemm(Z[r0:r1], Z[r0:r1], Cd, r1 - r0, r1 - r0, self.npad, alpha=-1.0, beta=1.0,
                          lower_only=1, Ct=Cd)
                yield r0, r1
        return self._rows_out(C, m, m, blocks(), host)

    def var(self, xo):
        """diag(cov(xo)) = k(x*, x*) - |K(x*, x) L^-T|^2 row by row: N^2 M flop instead of
        N^2 M + N M^2, and M doubles back to the host instead of M^2 (additive API)."""
        W, _ = self.inv_factor()
        m = int(xo.size)
        if m == 0:
            return np.empty(0, dtype=DTYPE)
        out = D.empty(m)
        dxo = D.to_device(xo)
        step = max(D.NB, (1 << 28) // self.npad // D.NB * D.NB)      # <= 2 GiB of Z at a time
        for lo in range(0, m, step):
            hi = min(m, lo + step)
            mp = D.roundup(hi - lo)
            Kxox = self.build(dxo[lo:hi], hi - lo, self.dx, self.n, mp, self.npad, 1)[0]
            Z = D.empty(mp, self.npad)
            self.gemm(Kxox, W, Z, mp, self.npad, self.npad, b_tri=1)
            call("gpb_post_var", self.kind, self._theta(), D.ptr(Z), self.npad, hi - lo, self.n,
                 out[lo:].data_ptr(), D.stream_ptr())
        return D.to_host(out).copy()

    def cov_rows(self, xo, lo, hi, host=True):
        """Rows lo:hi of cov(xo) -- the unit of test-point sharding: needs no peer
        data.  U = K(xo_r, x) Ki, then K(xo_r, xo) - U K(xo, x)^T; 2 N^2 m_r + 2 N M m_r flop.
        The whole range (one shard) takes the symmetric path of ``cov``."""
        m, mb = int(xo.size), int(hi - lo)
        if mb <= 0 or m == 0:
            return np.empty((max(mb, 0), m), dtype=DTYPE)
        if lo == 0 and hi == m:
            return self.cov(xo, host)
        Ki = self.Ki()
        mp, mbp = D.roundup(m), D.roundup(mb)
        dxo = D.to_device(xo)
        dxr = dxo[lo:hi]
        Kr = self.build(dxr, mb, self.dx, self.n, mbp, self.npad, 1)[0]
        U = D.empty(mbp, self.npad)
        self.gemm(Kr, Ki, U, mbp, self.npad, self.npad)           # Ki symmetric: NT form is K(xo_r,x) Ki
        del Kr
        Kall = self.build(dxo, m, self.dx, self.n, mp, self.npad, 1)[0]
        C = self.build(dxr, mb, dxo, m, mbp, mp, 1)[0]
        P = self._panel(mbp)

        # The shard's own diagonal block C[:, lo:hi] is symmetric: when it is tile-aligned, compute its
        # lower tiles only (mirrored), bottom-up like ``cov``, and the columns left / right of it densely.
        sym = (lo % D.NB == 0) and (mb == mbp) and (mbp >= 2 * D.NB)

        def blocks():
            if not sym:
                for r0 in range(0, mbp, P):
                    r1 = min(mbp, r0 + P)
                    self.gemm(U[r0:r1], Kall, C[r0:r1], r1 - r0, mp, self.npad, alpha=-1.0, beta=1.0)
                    yield r0, r1
                return
            Kd = Kall[lo:lo + mbp]
            for r1 in range(mbp, 0, -P):
                r0 = max(0, r1 - P)
                Ub = U[r0:r1]
                if lo > 0:
                    self.gemm(Ub, Kall[:lo], C[r0:r1, :lo], r1 - r0, lo, self.npad, alpha=-1.0, beta=1.0)
                if lo + mbp < mp:
                    self.gemm(Ub, Kall[lo + mbp:], C[r0:r1, lo + mbp:], r1 - r0, mp - lo - mbp, self.npad,
                              alpha=-1.0, beta=1.0)
                if r0 > 0:
                    self.gemm(Ub, Kd[:r0], C[r0:r1, lo:lo + r0], r1 - r0, r0, self.npad, alpha=-1.0, beta=1.0,
                              Ct=C[:r0, lo + r0:lo + r1])
                Cd = C[r0:r1, lo + r0:lo + r1]
                self.gemm(Ub, Kd[r0:r1], Cd, r1 - r0, r1 - r0, self.npad, alpha=-1.0, beta=1.0, lower_only=1, Ct=Cd)
                yield r0, r1
        return self._rows_out(C, mb, m, blocks(), host)

    # ------------------------------------------------------------------ test points sharded over GPUs
    def z_block(self, dxo_block, mb, bs):
        """Z_b = K(xo_b, x) L^-T for one block of test points, zero rows beyond ``mb``: [bs, npad] on the
        device -- N^2 bs flop (W lower triangular).  The unit a rank contributes to the all-gather."""
        W, _ = self.inv_factor()
        K = self.build(dxo_block, mb, self.dx, self.n, bs, self.npad, 1)[0]
        Z = D.empty(bs, self.npad)
        self.gemm(K, W, Z, bs, self.npad, self.npad, b_tri=1)
        return Z

    def cov_panel(self, dxo, m, b, bs, zget):
        """Block row b of the LOWER triangle of cov(xo): C[lo:hi, 0:hi] = K(xo_b, xo[:hi]) - Z_b Z[:hi]^T with
        lo = b bs (``zget(c)`` -> Z_c on this device).  Blocks left of the diagonal are dense products, the
        diagonal block computes its lower tiles and mirrors them, nothing right of it is touched: summed over
        all block rows that is N M^2 / 2 flop + the diagonal blocks -- half the row-shard form.
        Returns the device tensor [bs, (b + 1) bs] (valid part [:hi - lo, :hi])."""
        lo = b * bs
        hi = min(m, lo + bs)
        width = (b + 1) * bs
        C = self.build(dxo[lo:hi], hi - lo, dxo[:hi], hi, bs, width, 1)[0]
        Zb = zget(b)
        for c in range(b):
            self.gemm(Zb, zget(c), C[:, c * bs:(c + 1) * bs], bs, bs, self.npad, alpha=-1.0, beta=1.0)
        Cd = C[:, b * bs:(b + 1) * bs]
        self.gemm(Zb, Zb, Cd, bs, bs, self.npad, alpha=-1.0, beta=1.0, lower_only=1, Ct=Cd)
        return C

    def solve_residual(self):
        """max |Kxx alpha - y| / max |y| computed on the device with regenerated kernel tiles
        (size-independent check of the factorisation + solves)."""
        a = self.alpha()
        out = D.empty(self.n)
        call("gpb_kernel_matvec", self.kind, self._theta(), D.ptr(self.dx), self.n, D.ptr(self.dx), self.n, 1,
             iarr([0]), iarr([0]), darr([1.0]), parr([D.ptr(a)]), 1, parr([D.ptr(out)]), D.stream_ptr())
        r = out + (self.s ** 2) * a[:self.n] - self.dy[:self.n]
        return float(r.abs().max().item() / self.dy[:self.n].abs().max().item())

    def dm(self, xo):
        """gp_c.dm_dtheta (gp_c.pyx:114-131) with mat-vecs only:
        dm[i] = dK_i(xo,x) alpha - K(xo,x) Ki (dK_i alpha);  s row: -K(xo,x) Ki (2 s alpha)."""
        n, npad, n_p = self.n, self.npad, self.n_p
        nth = n_p + 1
        Ki, a = self.Ki(), self.alpha()
        st = D.stream_ptr()
        m = int(xo.size)
        sl = jac_slices(self.kind)
        B = D.zeros(n_p, npad)
        call("gpb_kernel_matvec", self.kind, self._theta(), D.ptr(self.dx), n, D.ptr(self.dx), n, n_p,
             iarr(sl), iarr(range(n_p)), darr([1.0] * n_p), parr([D.ptr(a)] * n_p), n_p,
             parr([B[i].data_ptr() for i in range(n_p)]), st)
        C = D.zeros(nth, npad)
        for i in range(n_p):
            self.gemv(Ki, n, n, B[i], C[i])
        self.gemv(Ki, n, n, a, C[n_p])
        out = D.zeros(nth, max(m, 1))
        if m:
            dxo = D.to_device(xo)
            slices = sl + [0] * nth
            outidx = list(range(n_p)) + list(range(nth))
            coef = [1.0] * n_p + [-1.0] * n_p + [-2.0 * self.s]
            vecs = [D.ptr(a)] * n_p + [C[i].data_ptr() for i in range(nth)]
            call("gpb_kernel_matvec", self.kind, self._theta(), D.ptr(dxo), m, D.ptr(self.dx), n, len(slices),
                 iarr(slices), iarr(outidx), darr(coef), parr(vecs), nth,
                 parr([out[i].data_ptr() for i in range(nth)]), st)
        return D.to_host(out[:, :m]).copy()


# ---------------------------------------------------------------------------
# batched evaluation: log_lh + dloglh_dtheta for many hyperparameter candidates
# ---------------------------------------------------------------------------
class BatchEvaluator(object):
    """Fixed (x, y) on the device; evaluates chunks of theta candidates through the
    fused C-ABI evaluator ``gpb_gp_eval`` (one workspace, reused)."""

    def __init__(self, kind, x, y, max_batch=None, workspace_gb=24.0):
        self.kind = int(kind)
        self.nth = N_KP[self.kind] + 1
        self.n = int(np.asarray(x).size)
        dev = D.require_cuda()
        # pinned staging: host inputs cross PCIe asynchronously on the compute stream
        self._hx = torch.empty(self.n, dtype=D.F64).pin_memory()
        self._hy = torch.empty(self.n, dtype=D.F64).pin_memory()
        self._hres = None
        self.dx = torch.empty(self.n, dtype=D.F64, device=dev)
        self.dy = torch.empty(self.n, dtype=D.F64, device=dev)
        self.set_data(x, y)
        per = int(_lib.lib.gpb_eval_workspace_bytes(self.n, 1, 1))
        cap = max(1, int(workspace_gb * 2 ** 30) // per)
        self.max_batch = int(min(cap, max_batch or cap, 65535 // max(1, D.roundup(self.n) // D.NB)))
        self._ws = None
        self._ws_batch = 0

    def set_data(self, x, y):
        """(Re)upload the observations: numpy -> pinned staging -> HBM."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        if x.size != self.n or y.size != self.n:
            raise ValueError("x and y must keep their length (%d)" % self.n)
        self._hx.numpy()[:] = x
        self._hy.numpy()[:] = y
        self.dx.copy_(self._hx, non_blocking=True)
        self.dy.copy_(self._hy, non_blocking=True)

    def _workspace(self, batch, want_grad):
        if self._ws is None or batch > self._ws_batch:
            nbytes = int(_lib.lib.gpb_eval_workspace_bytes(self.n, batch, 1))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=D.require_cuda())
            self._ws_batch = batch
        return self._ws

    def eval_device(self, thetas, want_grad=True):
        """thetas: host [B, n_theta].  Returns the device tensor [B, 8]:
        log_lh, dloglh[0..3], logdet, y.alpha, info."""
        thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        B = thetas.shape[0]
        res = D.empty(B, 8)
        for lo in range(0, B, self.max_batch):
            hi = min(B, lo + self.max_batch)
            ws = self._workspace(hi - lo, want_grad)
            chunk = np.ascontiguousarray(thetas[lo:hi])
            call("gpb_gp_eval", self.kind, chunk.ctypes.data_as(_lib.dp), hi - lo, D.ptr(self.dx),
                 D.ptr(self.dy), self.n, int(want_grad), D.ptr(ws), ws.numel(), res[lo:].data_ptr(),
                 D.stream_ptr())
        return res

    def eval(self, thetas, want_grad=True):
        """(log_lh[B], dloglh[B, n_theta], info[B]) as numpy arrays."""
        res = self.eval_device(thetas, want_grad)
        if self._hres is None or self._hres.shape[0] < res.shape[0]:
            self._hres = torch.empty(res.shape[0], 8, dtype=D.F64).pin_memory()
        h = self._hres[:res.shape[0]]
        h.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        r = h.numpy()
        return r[:, 0].copy(), r[:, 1:1 + self.nth].copy(), r[:, 7].astype(np.int64)
