"""Device plumbing: torch owns HBM allocations and streams; every computation is a
call into libgpb200.so with raw device pointers."""
import numpy as np
import torch

from . import _lib

NB = 128          # GPB_BLOCK: padding / blocking unit of device matrices
F64 = torch.float64


_cuda_ok = False


def require_cuda():
    """The current CUDA device (raises when there is none: no CPU fallback).  Availability is
    probed once -- torch.cuda.is_available() costs ~4 us per call, which is real money for a
    50-point GP whose whole evaluation is ~20 kernel launches."""
    global _cuda_ok
    if not _cuda_ok:
        if not torch.cuda.is_available():
            raise _lib.GpbError("gaussian_processes_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _cuda_ok = True
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw handle, ~0.3 us)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def roundup(n, m=NB):
    return (int(n) + m - 1) // m * m


def empty(*shape):
    return torch.empty(*shape, dtype=F64, device=require_cuda())


def zeros(*shape):
    return torch.zeros(*shape, dtype=F64, device=require_cuda())


def ones(*shape):
    return torch.ones(*shape, dtype=F64, device=require_cuda())


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


# ---------------------------------------------------------------------------
# matrix-valued results: page-locked result arrays from a recycling pool
# ---------------------------------------------------------------------------
class _PinnedPool(object):
    """Page-locked host buffers handed out as numpy arrays.

    The array returned to the user is a view of a pooled buffer; a ``weakref.finalize`` on the
    buffer's root ndarray (the ``.base`` every user-side view collapses to) gives the buffer
    back once the last view is gone, so a result is never overwritten while it is reachable
    and ``cudaHostAlloc`` (~0.3 s/GB) is paid once per size class, not per call."""

    GRAIN = 1 << 21            # size classes: multiples of 2 MiB
    KEEP_BYTES = 6 << 30       # idle buffers kept for reuse

    def __init__(self):
        self._free = {}
        self._idle = 0

    def _give(self, buf):
        n = buf.numel()
        if self._idle + n > self.KEEP_BYTES:
            return                                   # dropped: torch frees the pinned block
        self._free.setdefault(n, []).append(buf)
        self._idle += n

    def take(self, nbytes):
        import weakref
        n = max(self.GRAIN, (int(nbytes) + self.GRAIN - 1) // self.GRAIN * self.GRAIN)
        lst = self._free.get(n)
        if lst:
            buf = lst.pop()
            self._idle -= n
        else:
            buf = torch.empty(n, dtype=torch.uint8).pin_memory()
        root = buf.numpy()
        weakref.finalize(root, self._give, buf)
        return root, buf.data_ptr()


_pool = _PinnedPool()
_copy_streams = {}
PINNED_MIN_BYTES = 1 << 20
PINNED_MAX_BYTES = 4 << 30      # larger results take the staged pageable path (do not lock that much RAM)


def copy_stream():
    """Side stream for device->host result traffic that overlaps compute."""
    dev = torch.cuda.current_device()
    if dev not in _copy_streams:
        _copy_streams[dev] = torch.cuda.Stream(device=dev)
    return _copy_streams[dev]


def host_array(rows, cols):
    """Fresh [rows, cols] float64 result array, page-locked when large.  -> (array, pinned)"""
    nbytes = int(rows) * int(cols) * 8
    if nbytes < PINNED_MIN_BYTES or nbytes > PINNED_MAX_BYTES:
        return np.empty((int(rows), int(cols)), dtype=np.float64), False
    root, _ = _pool.take(nbytes)
    return root[:nbytes].view(np.float64).reshape(int(rows), int(cols)), True


def download_2d(t, rows, cols, out=None, pinned=False, r0=0, stream=None, sync=True):
    """Rows [r0, r0+rows) x cols of the 2-D device tensor ``t`` -> the same rows of the host
    array ``out`` (``pinned`` says whether it came page-locked from ``host_array``), or a new array."""
    require_cuda()
    if out is None:
        out, pinned = host_array(r0 + rows, cols)
    dst = out[r0:r0 + rows]
    st = stream if stream is not None else torch.cuda.current_stream()
    if rows and cols:
        _lib.call("gpb_download_2d", dst.ctypes.data, dst.strides[0] // 8, t.data_ptr() + r0 * t.stride(0) * 8,
                  t.stride(0), int(rows), int(cols), int(bool(pinned)), st.cuda_stream)
        if sync and pinned:
            st.synchronize()
    return out


def ptr(t):
    return t.data_ptr() if t is not None else None
