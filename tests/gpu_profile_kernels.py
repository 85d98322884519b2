"""Isolated launches of the hot kernels for `ncu --set full` (one GPU, few launches so the
report stays small).  Modes:
  gemm : four DMMA GEMM shapes of the N=4096 factorisation, called through the C ABI
         (trailing SYRK K=512 beta=1 lower-only | lauum-shaped upper x upper | TRSM-shaped |
          skinny column update K=384)
  eval : one scalar log_lh + dloglh_dtheta at N=4096 (use -k to pick diag / trsv / grad / build)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gaussian_processes_b200 as gpb  # noqa: E402
from gaussian_processes_b200 import _lib, device as D  # noqa: E402
from gaussian_processes_b200._lib import call  # noqa: E402
from conftest import synth_xy  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "gemm"
n = 4096

if mode == "gemm":
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    U = torch.triu(A).contiguous()
    C = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    Wd = torch.tril(torch.randn(128, 128, dtype=torch.float64, device="cuda", generator=g)).contiguous()
    st = D.stream_ptr()

    def gemm(Ap, lda, Bp, ldb, Cp, ldc, Ctp, ldct, M, N, K, alpha, beta, a_tri, b_tri, lower):
        call("gpb_gemm_nt", Ap, lda, Bp, ldb, Cp, ldc, Ctp, ldct, M, N, K, alpha, beta, a_tri, b_tri, lower, st)
    for rep in range(2):
        # 1. trailing update after the first 512-wide panel: C[512:,512:] -= P P^T, P = A[512:, 0:512]
        P = A[512:, :512]
        Ct = C[512:, 512:]
        gemm(P.data_ptr(), n, P.data_ptr(), n, Ct.data_ptr(), n, None, 0, n - 512, n - 512, 512, -1.0, 1.0, 0, 0, 1)
        # 2. lauum: Ki = U U^T, lower tiles + mirror
        gemm(U.data_ptr(), n, U.data_ptr(), n, C.data_ptr(), n, C.data_ptr(), n, n, n, n, 1.0, 0.0, 2, 2, 1)
        # 3. TRSM by inverted diagonal block, in place
        Pn = A[128:, :128]
        gemm(Pn.data_ptr(), n, Wd.data_ptr(), 128, Pn.data_ptr(), n, None, 0, n - 128, 128, 128, 1.0, 0.0, 0, 1, 0)
        # 4. skinny column update inside a panel (K = 384)
        Pc = A[384:, :384]
        Cc = C[384:, 384:512]
        gemm(Pc.data_ptr(), n, Pc.data_ptr(), n, Cc.data_ptr(), n, None, 0, n - 384, 128, 384, -1.0, 1.0, 0, 0, 0)
    torch.cuda.synchronize()
    print("gemm shapes done")
else:
    x, y = synth_xy(n, 0)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    for rep in range(2):
        gp.set_param("w", 0.5 + 0.01 * rep)
        gp.log_lh
        gp.dloglh_dtheta
    torch.cuda.synchronize()
    print("eval done", gp.log_lh)
