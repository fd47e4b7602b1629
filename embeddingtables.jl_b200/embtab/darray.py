"""DeviceArray: a column-major (Julia-layout) N-d array resident in HBM.

The reference keeps tables, indices, outputs and cotangents in Julia `Array`s (column-major).
Here they live in device memory allocated through PyTorch (device memory / streams are the only
things torch is used for); shapes are Julia shapes, memory order is Julia's, so pointers and
leading dimensions go straight into the C ABI.  2-d arrays may be row-slice views of a parent
(leading dimension > rows) -- what `view(dst, rows, :)` is in the reference's
PreallocationStrategy (src/lookup.jl:336-340) and its pullback (src/utils.jl:50-63).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
         np.dtype(np.float16): torch.float16}
_NP2ELT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64,
           np.dtype(np.int32): _lib.I32, np.dtype(np.int64): _lib.I64,
           np.dtype(np.float16): _lib.F16}
try:   # numpy has no bfloat16 of its own; ml_dtypes provides one (round-to-nearest-even conversions)
    import ml_dtypes
    bfloat16 = np.dtype(ml_dtypes.bfloat16)
    _NP2T[bfloat16] = torch.bfloat16
    _NP2ELT[bfloat16] = _lib.BF16
except ImportError:   # bf16 tables then need the C ABI directly
    bfloat16 = None


def _np_to_torch(flat: np.ndarray) -> torch.Tensor:
    """torch.from_numpy, also for bfloat16 (which torch's numpy bridge does not know)"""
    if bfloat16 is not None and flat.dtype == bfloat16:
        return torch.from_numpy(flat.view(np.int16)).view(torch.bfloat16)
    return torch.from_numpy(flat)


def _torch_to_np(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(bfloat16)
    return t.numpy()


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.EmbTabError("no CUDA device: embtab computes on a B200 only (no CPU fallback)")


class DeviceArray:
    __slots__ = ("buf", "shape", "offset", "ld", "dtype")

    def __init__(self, buf: torch.Tensor, shape, offset=0, ld=None, dtype=None):
        self.buf = buf                      # 1-d torch tensor owning (or sharing) the memory
        self.shape = tuple(int(s) for s in shape)
        self.offset = int(offset)           # elements from buf[0]
        self.ld = int(ld) if ld is not None else (self.shape[0] if self.shape else 1)
        self.dtype = np.dtype(dtype) if dtype is not None else np.dtype(
            {v: k for k, v in _NP2T.items()}[buf.dtype])

    # ---- construction ---------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype=np.float32, device=None) -> "DeviceArray":
        require_cuda()
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(shape)) if shape else 1
        buf = torch.empty(n, dtype=_NP2T[np.dtype(dtype)], device=device or "cuda")
        return DeviceArray(buf, shape, dtype=dtype)

    @staticmethod
    def zeros(shape, dtype=np.float32, device=None) -> "DeviceArray":
        a = DeviceArray.empty(shape, dtype, device)
        a.buf.zero_()
        return a

    @staticmethod
    def from_numpy(a, dtype=None, device=None) -> "DeviceArray":
        """Upload a host array given in Julia shape (any memory order)."""
        require_cuda()
        a = np.asarray(a, dtype=dtype)
        flat = np.ascontiguousarray(a.reshape(-1, order="F"))
        buf = _np_to_torch(flat).to(device or "cuda")
        return DeviceArray(buf, a.shape, dtype=a.dtype)

    @staticmethod
    def from_pointer(ptr: int, shape, dtype=np.float32, ld=None) -> "DeviceArray":
        """Wrap raw device memory (e.g. an etb_malloc allocation or a peer's IPC-mapped buffer).
        No ownership: the caller keeps the allocation alive."""
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        ld_ = int(ld) if ld is not None else shape[0]
        n = (shape[-1] - 1) * ld_ + shape[0] if len(shape) == 2 else int(np.prod(shape))

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (max(n, 1),), "typestr": np.dtype(dtype).str,
                                        "data": (int(ptr), False), "version": 3}
        buf = torch.as_tensor(raw, device="cuda")
        return DeviceArray(buf, shape, 0, ld_, dtype)

    @staticmethod
    def mapped(host: np.ndarray) -> "DeviceArray":
        """Device view of a PAGE-LOCKED host array (see pinned_empty).  Pinned memory is device-addressable
        under unified addressing, so a kernel can store its result into it / load its operand from it over
        PCIe without a staging copy.  No ownership: keep `host` alive."""
        assert host.flags.f_contiguous or host.ndim <= 1, "Julia (column-major) layout expected"
        return DeviceArray.from_pointer(host.ctypes.data, host.shape, host.dtype)

    def similar(self, dtype=None, shape=None) -> "DeviceArray":
        """Julia `similar(example(A), T, dims)` (reference src/lookup.jl:20-22)."""
        return DeviceArray.empty(self.shape if shape is None else shape, dtype or self.dtype,
                                 self.buf.device)

    # ---- properties -----------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def elt(self):
        return _NP2ELT[self.dtype]

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.offset * self.itemsize

    @property
    def is_dense(self):
        return self.ndim != 2 or self.ld == self.shape[0] or self.shape[1] <= 1

    def size(self, d=None):
        return self.shape if d is None else self.shape[d - 1]  # 1-based like Julia

    # ---- views ----------------------------------------------------------------------
    def rows(self, lo: int, hi: int) -> "DeviceArray":
        """view(A, lo+1:hi, :) for a matrix (0-based half-open here)."""
        assert self.ndim == 2 and 0 <= lo <= hi <= self.shape[0]
        return DeviceArray(self.buf, (hi - lo, self.shape[1]), self.offset + lo, self.ld, self.dtype)

    def cols(self, lo: int, hi: int) -> "DeviceArray":
        """view(A, :, lo+1:hi) for a matrix, or A[lo+1:hi] for a vector."""
        if self.ndim == 1:
            return DeviceArray(self.buf, (hi - lo,), self.offset + lo, None, self.dtype)
        assert self.ndim == 2
        return DeviceArray(self.buf, (self.shape[0], hi - lo), self.offset + lo * self.ld, self.ld, self.dtype)

    def lastdim(self, t: int) -> "DeviceArray":
        """view(A, :, ..., :, t+1): the ColumnWrap element (reference src/lookup.jl:195-208)."""
        assert self.ndim >= 2 and self.is_dense
        inner = int(np.prod(self.shape[:-1]))
        return DeviceArray(self.buf, self.shape[:-1], self.offset + t * inner, None, self.dtype)

    # ---- host transfer --------------------------------------------------------------
    def numpy(self) -> np.ndarray:
        """Download as a Fortran-ordered numpy array of the Julia shape."""
        if self.ndim == 2 and not self.is_dense:
            span = (self.shape[1] - 1) * self.ld + self.shape[0]
            flat = _torch_to_np(self.buf[self.offset:self.offset + span].cpu())
            full = np.lib.stride_tricks.as_strided(
                flat, shape=self.shape, strides=(self.itemsize, self.ld * self.itemsize))
            return np.asfortranarray(full)
        n = int(np.prod(self.shape)) if self.shape else 1
        flat = _torch_to_np(self.buf[self.offset:self.offset + n].cpu())
        return flat.reshape(self.shape, order="F")

    def copy_from(self, a) -> "DeviceArray":
        a = np.asarray(a, dtype=self.dtype)
        assert a.shape == self.shape, (a.shape, self.shape)
        if self.ndim == 2 and not self.is_dense:
            for j in range(self.shape[1]):
                s = self.offset + j * self.ld
                self.buf[s:s + self.shape[0]].copy_(_np_to_torch(np.ascontiguousarray(a[:, j])))
            return self
        n = int(np.prod(self.shape)) if self.shape else 1
        self.buf[self.offset:self.offset + n].copy_(
            _np_to_torch(np.ascontiguousarray(a.reshape(-1, order="F"))))
        return self

    def upload(self, host: np.ndarray) -> "DeviceArray":
        """Asynchronous host->device copy on the current stream.  `host` holds this array's
        elements in column-major memory order (e.g. a pinned buffer from `pinned_empty`, or a row-slice
        view of one); the copy is truly asynchronous only from pinned memory.  Row-slice views on either
        side go as one strided (2-D) copy, without staging."""
        assert host.dtype == self.dtype and host.size == int(np.prod(self.shape))
        hp = _host_pitch(host)
        if self.is_dense and hp is None:
            _lib.check(_lib.lib().etb_memcpy_h2d(self.ptr, host.ctypes.data, host.nbytes, current_stream_ptr()))
        else:
            assert self.ndim == 2 and host.shape == self.shape
            w = self.shape[0] * self.itemsize
            _lib.check(_lib.lib().etb_memcpy2d_h2d(self.ptr, self.ld * self.itemsize, host.ctypes.data, hp or w, w,
                                                   self.shape[1], current_stream_ptr()))
        return self

    def download(self, host: np.ndarray) -> np.ndarray:
        """Asynchronous device->host copy on the current stream (synchronise before reading)."""
        assert host.dtype == self.dtype and host.size == int(np.prod(self.shape))
        hp = _host_pitch(host)
        if self.is_dense and hp is None:
            _lib.check(_lib.lib().etb_memcpy_d2h(host.ctypes.data, self.ptr, host.nbytes, current_stream_ptr()))
        else:
            assert self.ndim == 2 and host.shape == self.shape
            w = self.shape[0] * self.itemsize
            _lib.check(_lib.lib().etb_memcpy2d_d2h(host.ctypes.data, hp or w, self.ptr, self.ld * self.itemsize, w,
                                                   self.shape[1], current_stream_ptr()))
        return host

    def fill(self, v):
        if self.is_dense:
            n = int(np.prod(self.shape)) if self.shape else 1
            self.buf[self.offset:self.offset + n].fill_(v)
        else:
            for j in range(self.shape[1]):
                s = self.offset + j * self.ld
                self.buf[s:s + self.shape[0]].fill_(v)
        return self

    def copy(self) -> "DeviceArray":
        out = DeviceArray.empty(self.shape, self.dtype, self.buf.device)
        if self.is_dense:
            n = int(np.prod(self.shape)) if self.shape else 1
            out.buf.copy_(self.buf[self.offset:self.offset + n])
        else:
            for j in range(self.shape[1]):
                s = self.offset + j * self.ld
                out.buf[j * self.shape[0]:(j + 1) * self.shape[0]].copy_(self.buf[s:s + self.shape[0]])
        return out

    def __repr__(self):
        return f"DeviceArray{self.shape}<{self.dtype}, ld={self.ld}>"


def _host_pitch(host: np.ndarray):
    """None for one contiguous block; the column pitch in bytes for a row-slice view of a column-major matrix."""
    if host.flags.f_contiguous or host.flags.c_contiguous:
        return None
    assert host.ndim == 2 and host.strides[0] == host.itemsize and host.strides[1] >= host.shape[0] * host.itemsize, \
        "host array must be contiguous or a row-slice view of a column-major matrix"
    return host.strides[1]


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A page-locked host array (Fortran order, Julia shape) for asynchronous upload/download."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    t = torch.empty(int(np.prod(shape)), dtype=_NP2T[np.dtype(dtype)], pin_memory=True)
    a = _torch_to_np(t).reshape(shape, order="F")
    _PINNED_KEEPALIVE[a.ctypes.data] = t
    return a


_PINNED_KEEPALIVE = {}


def as_device(a, dtype=None) -> DeviceArray:
    """Host arrays (Julia shape) are uploaded; DeviceArrays pass through."""
    if isinstance(a, DeviceArray):
        return a
    return DeviceArray.from_numpy(a, dtype=dtype)


def as_device_indices(I) -> DeviceArray:
    if isinstance(I, DeviceArray):
        if I.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            raise TypeError(f"indices must be int32/int64, got {I.dtype}")
        return I
    I = np.asarray(I)
    if I.dtype != np.int32:
        I = I.astype(np.int64, copy=False)
    return DeviceArray.from_numpy(I)
