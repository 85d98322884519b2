"""Device plumbing: torch owns HBM allocations and streams; every computation is a
call into libgpb200.so with raw device pointers."""
import numpy as np
import torch

from . import _lib

NB = 128          # GPB_BLOCK: padding / blocking unit of device matrices
F64 = torch.float64


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.GpbError("gaussian_processes_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def roundup(n, m=NB):
    return (int(n) + m - 1) // m * m


def empty(*shape):
    return torch.empty(*shape, dtype=F64, device=require_cuda())


def zeros(*shape):
    return torch.zeros(*shape, dtype=F64, device=require_cuda())


def izeros(*shape):
    return torch.zeros(*shape, dtype=torch.int32, device=require_cuda())


def to_device(a, pad_to=None):
    """1-D float64 host array -> device tensor, optionally zero-padded."""
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    dev = require_cuda()
    if pad_to is None or pad_to == a.size:
        return torch.from_numpy(a.copy()).to(dev)
    out = torch.zeros(pad_to, dtype=F64, device=dev)
    out[:a.size] = torch.from_numpy(a.copy()).to(dev)
    return out


def mat_to_device(a, rows, cols, identity_pad=False):
    """2-D host array -> zero (or identity) padded [rows, cols] device tensor."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = torch.zeros(rows, cols, dtype=F64, device=require_cuda())
    out[:a.shape[0], :a.shape[1]] = torch.from_numpy(a).to(out.device)
    if identity_pad:
        k = min(rows, cols)
        idx = torch.arange(min(a.shape), k, device=out.device)
        out[idx, idx] = 1.0
    return out


def to_host(t):
    return t.detach().cpu().numpy()


def ptr(t):
    return t.data_ptr() if t is not None else None
