"""Likelihood / derivative reductions with the reference's Cython signatures
(reference: gp/ext/gp_c.pyx:17,34,52,70,114): numpy arrays in, outputs written in place.

These shims take *host* arrays (that is the reference contract), stage them into
identity/zero padded device buffers and run the same CUDA kernels the ``GP`` object uses
on resident data (``GP`` itself never goes through here -- it keeps everything on the
device and regenerates kernel tiles instead of reading ``Kj`` / ``Kh``).
"""
import numpy as np

from .. import _lib, device as D
from .._lib import call
from ._host import carray, out_array

__all__ = ["log_lh", "dloglh_dtheta", "dlh_dtheta", "d2lh_dtheta2", "dm_dtheta"]

MIN = float(np.log(np.exp2(np.float64(np.finfo(np.float64).minexp + 4))))


def _gemm(A, B, C, M, N, K, alpha=1.0, beta=0.0):
    call("gpb_gemm_nt", D.ptr(A), A.stride(-2), D.ptr(B), B.stride(-2), D.ptr(C), C.stride(-2), None, 0,
         M, N, K, alpha, beta, 0, 0, 0, D.stream_ptr())
    return C


def _gemv(A, rows, cols, lda, x, y=None, alpha=1.0, beta=0.0):
    if y is None:
        y = D.empty(rows)
    call("gpb_gemv", D.ptr(A), rows, cols, lda, D.ptr(x), D.ptr(y), alpha, beta, D.stream_ptr())
    return y


class _Scratch(object):
    def __init__(self, n):
        self.part = D.empty(int(_lib.lib.gpb_grad_partial_doubles(n)) + 1024)
        self.out = D.empty(64)
        self.k = 0

    def slot(self):
        s = self.out[self.k:]
        self.k += 1
        return s

    def trace_prod(self, A, B, ld, n):
        s = self.slot()
        call("gpb_trace_prod", D.ptr(A), ld, D.ptr(B), ld, n, D.ptr(self.part), s.data_ptr(), D.stream_ptr())
        return s

    def quadform(self, u, M, ld, v, n):
        s = self.slot()
        call("gpb_quadform", D.ptr(u), D.ptr(M), ld, D.ptr(v), n, D.ptr(self.part), s.data_ptr(), D.stream_ptr())
        return s

    def values(self):
        return D.to_host(self.out[:self.k])


def log_lh(y, K, Kiy):
    """gp_c.log_lh (gp_c.pyx:17-31).  log|K| comes from a Cholesky of K instead of the
    reference's LU ``slogdet``; a K that is not positive definite returns ``-inf`` (for
    such K the reference returns -inf too whenever det K <= 0)."""
    carray(y, 1, "y"); carray(K, 2, "K"); carray(Kiy, 1, "Kiy")
    n = y.size
    npad = D.roundup(n)
    L = D.mat_to_device(K, npad, npad, identity_pad=True)
    W = D.empty(npad, npad)
    info = D.izeros(1)
    call("gpb_potrf", D.ptr(L), npad, npad, 0, 1, D.ptr(W), npad, 0, None, 0, 0, D.ptr(info), D.stream_ptr())
    out = D.empty(3)
    dy, dKiy = D.to_device(y), D.to_device(Kiy)     # named: must outlive the launch
    call("gpb_loglh", D.ptr(L), n, npad, D.ptr(dy), D.ptr(dKiy), D.ptr(info), D.ptr(out), D.stream_ptr())
    return float(D.to_host(out)[0])


def _brackets(y, Ki, Kj, Kiy, s):
    """(y^T Ki dK_i Kiy, tr(Ki dK_i)) for i < n_p and the noise row (gp_c.pyx:41-49, 59-67)."""
    carray(y, 1, "y"); carray(Ki, 2, "Ki"); carray(Kj, 3, "Kj"); carray(Kiy, 1, "Kiy")
    n_p, n = Kj.shape[0], Kj.shape[1]
    dKi, dy, dKiy = D.mat_to_device(Ki, n, n), D.to_device(y), D.to_device(Kiy)
    sc = _Scratch(n)
    # y^T (Ki dK) Kiy = (Ki^T y)^T dK Kiy: one transposed mat-vec, then quadratic forms
    dKiT = dKi.t().contiguous()
    KiT_y = _gemv(dKiT, n, n, n, dy)
    keep = [dKiT]
    for i in range(n_p):
        dKj = D.mat_to_device(Kj[i], n, n)
        keep.append(dKj)
        sc.quadform(KiT_y, dKj, n, dKiy, n)
        sc.trace_prod(dKi, dKj, n, n)
    eye = _eye(n)
    sc.trace_prod(dKi, eye, n, n)
    dot = D.empty(1)
    call("gpb_gemv", D.ptr(KiT_y), 1, n, n, D.ptr(dKiy), D.ptr(dot), 1.0, 0.0, D.stream_ptr())
    v = sc.values()
    t0 = np.append(v[0:2 * n_p:2], 2.0 * s * float(dot.item()))
    t1 = np.append(v[1:2 * n_p:2], 2.0 * s * v[2 * n_p])
    return t0, t1


def _eye(n):
    import torch
    return torch.eye(n, dtype=D.F64, device=D.require_cuda())


def dloglh_dtheta(y, Ki, Kj, Kiy, s, dloglh):
    """gp_c.dloglh_dtheta (gp_c.pyx:34-49)."""
    out_array(dloglh, (Kj.shape[0] + 1,), "dloglh")
    t0, t1 = _brackets(y, Ki, Kj, Kiy, float(s))
    dloglh[:] = 0.5 * t0 + -0.5 * t1


def dlh_dtheta(y, Ki, Kj, Kiy, s, lh, dlh):
    """gp_c.dlh_dtheta (gp_c.pyx:52-67)."""
    out_array(dlh, (Kj.shape[0] + 1,), "dlh")
    t0, t1 = _brackets(y, Ki, Kj, Kiy, float(s))
    dlh[:] = 0.5 * float(lh) * (t0 - t1)


def d2lh_dtheta2(y, Ki, Kj, Kh, Kiy, s, lh, dlh, d2lh):
    """gp_c.d2lh_dtheta2 (gp_c.pyx:70-111) on the DMMA GEMM: dKi_j = -Ki dK_j Ki and the
    per-(i, j) traces / quadratic forms, from the caller's dense Ki, Kj, Kh."""
    carray(y, 1, "y"); carray(Ki, 2, "Ki"); carray(Kj, 3, "Kj"); carray(Kh, 4, "Kh")
    carray(Kiy, 1, "Kiy"); carray(dlh, 1, "dlh")
    n_p, n = Kj.shape[0], Kj.shape[1]
    nth = n_p + 1
    out_array(d2lh, (nth, nth), "d2lh")
    s, lh = float(s), float(lh)
    npad = D.roundup(n)
    dKi = D.mat_to_device(Ki, npad, npad)
    dKiT = dKi.t().contiguous()
    dy, dKiy = D.to_device(y, npad), D.to_device(Kiy, npad)
    eye2s = _eye(npad) * (2 * s)
    eye2s[n:, n:] = 0
    dK = [D.mat_to_device(Kj[i], npad, npad) for i in range(n_p)] + [eye2s]
    # dKi_j = -Ki (dK_j Ki): two NT products (B operand = transposed right factor)
    dKi_l = []
    for j in range(nth):
        T = _gemm(dK[j], dKiT, D.empty(npad, npad), npad, npad, npad)             # dK_j Ki
        dKi_l.append(_gemm(dKi, T.t().contiguous(), D.empty(npad, npad), npad, npad, npad, alpha=-1.0))
    sc = _Scratch(n)
    res = np.empty((nth, nth))
    KiT_y = _gemv(dKiT, npad, npad, npad, dy)                                      # (y^T Ki)^T
    eye = _eye(npad)
    for i in range(nth):
        dKiT_i = dK[i].t().contiguous()
        sc.quadform(KiT_y, dK[i], npad, dKiy, n)                                   # y^T Ki dK_i Kiy
        sc.trace_prod(dKi, dK[i], npad, n)                                         # tr(Ki dK_i)
        for j in range(nth):
            G = _gemm(dKi_l[j], dKiT_i, D.empty(npad, npad), npad, npad, npad)     # dKi_j dK_i
            if i < n_p and j < n_p:
                d2k = D.mat_to_device(Kh[i, j], npad, npad)
            elif i == n_p and j == n_p:
                d2k = _eye(npad) * 2.0
                d2k[n:, n:] = 0
            else:
                d2k = D.zeros(npad, npad)
            sc2 = _Scratch(n)
            sc2.quadform(dy, G, npad, dKiy, n)                                     # t1a
            sc2.quadform(dKiy, d2k, npad, dKiy, n)                                 # t1b
            w = _gemv(dKi_l[j], npad, npad, npad, dy)                              # dKi_j y
            sc2.quadform(dKiy, dK[i], npad, w, n)                                  # t1c
            sc2.trace_prod(G, eye, npad, n)                                        # tr(dKi_j dK_i)
            sc2.trace_prod(dKi, d2k, npad, n)                                      # tr(Ki d2k)
            v = sc2.values()
            res[i, j] = v[0] + v[1] + v[2] - (v[3] + v[4])
    v = sc.values()
    for i in range(nth):
        r_i = v[2 * i] - v[2 * i + 1]
        for j in range(nth):
            d2lh[i, j] = 0.5 * (dlh[j] * r_i + lh * res[i, j])


def dm_dtheta(y, Ki, Kj, Kjxo, Kxox, s, dm):
    """gp_c.dm_dtheta (gp_c.pyx:114-131) with mat-vecs:
    dm[i] = dKxox_i (Ki y) - Kxox (Ki (dK_i (Ki y)))."""
    carray(y, 1, "y"); carray(Ki, 2, "Ki"); carray(Kj, 3, "Kj"); carray(Kjxo, 3, "Kjxo"); carray(Kxox, 2, "Kxox")
    n_p, n, m = Kj.shape[0], Kj.shape[1], Kjxo.shape[1]
    out_array(dm, (n_p + 1, m), "dm")
    s = float(s)
    dKi, dy = D.mat_to_device(Ki, n, n), D.to_device(y)
    dKxox = D.mat_to_device(Kxox, m, n)
    Kiy = _gemv(dKi, n, n, n, dy)
    out = D.zeros(n_p + 1, max(m, 1))
    for i in range(n_p + 1):
        if i < n_p:
            b = _gemv(D.mat_to_device(Kj[i], n, n), n, n, n, Kiy)
            _gemv(D.mat_to_device(Kjxo[i], m, n), m, n, n, Kiy, out[i])
            c = _gemv(dKi, n, n, n, b)
            _gemv(dKxox, m, n, n, c, out[i], alpha=-1.0, beta=1.0)
        else:
            c = _gemv(dKi, n, n, n, Kiy)
            _gemv(dKxox, m, n, n, c, out[i], alpha=-2.0 * s, beta=0.0)
    dm[:] = D.to_host(out[:, :m])
