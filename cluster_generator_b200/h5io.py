"""
Host-side handle over the native HDF5 subset of ``libcgb200`` (``cgb_h5_*`` in
``include/cgb200.h``): what ``h5py.File`` is to the reference's particle I/O
(``particles.py:239-363, 453-552``).  Only the pieces that path uses: groups, contiguous
numeric datasets, scalar / 1-D numeric attributes and string attributes.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_char_p, c_int, c_int64, c_size_t, c_void_p

import numpy as np
import torch

from . import _native as N

_CODES = {
    np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3,
    np.dtype("uint32"): 4, np.dtype("int64"): 5, np.dtype("uint64"): 6,
}
_DTYPES = {v: k for k, v in _CODES.items()}
H5_STR = 7


def _b(s: str) -> bytes:
    return s.encode("utf-8")


class H5Writer:
    """``with H5Writer(path, overwrite) as f: f.create_group(...); f.write(...)``"""

    def __init__(self, path, overwrite: bool = False):
        self._h = c_void_p()
        N.check(N.lib().cgb_h5_create(_b(str(path)), int(bool(overwrite)), byref(self._h)))

    def create_group(self, path: str) -> None:
        N.check(N.lib().cgb_h5_create_group(self._h, _b(path)))

    def write(self, path: str, data) -> None:
        """Host array (numpy, rank 0..4) -> dataset."""
        arr = np.asarray(data)
        if arr.ndim:  # (ascontiguousarray would turn a scalar into a 1-d array)
            arr = np.ascontiguousarray(arr)
        if arr.dtype not in _CODES:
            raise ValueError(f"dtype {arr.dtype} is outside the HDF5 subset written here")
        dims = (c_int64 * 4)(*arr.shape)
        N.check(N.lib().cgb_h5_write_dataset(self._h, _b(path), _CODES[arr.dtype], arr.ndim, dims,
                                             c_void_p(arr.ctypes.data)))

    def write_device(self, path: str, t: torch.Tensor, out_dtype="float64", scale: float = 1.0) -> None:
        """CUDA tensor -> dataset, without a host copy of the whole field.  ``t`` is ``(n,)``
        or an ``(n, k)`` view of ``(k, n)`` planes (the layout the samplers produce); other
        layouts are made plane-major first."""
        out_dtype = np.dtype(out_dtype)
        if t.dtype not in (torch.float32, torch.float64):
            raise ValueError("write_device takes float32/float64 tensors")
        if t.ndim == 1:
            planes, n, k = t.contiguous(), t.shape[0], 1
        elif t.ndim == 2:
            n, k = t.shape
            planes = t.t()
            if not planes.is_contiguous():
                planes = planes.contiguous()
        else:
            raise ValueError("write_device takes (n,) or (n, k) tensors")
        ld = planes.stride(0) if k > 1 else n
        with torch.cuda.device(t.device):  # staging buffers and the pack kernel live on t's device
            N.check(N.lib().cgb_h5_write_dataset_device(
                self._h, _b(path), _CODES[out_dtype], c_void_p(planes.data_ptr()),
                1 if t.dtype == torch.float32 else 2, n, k, ld, float(scale),
                c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))

    def set_attr(self, obj: str, name: str, value) -> None:
        if isinstance(value, str):
            N.check(N.lib().cgb_h5_set_attr(self._h, _b(obj), _b(name), H5_STR, -1, c_char_p(_b(value))))
            return
        if isinstance(value, (bool, int)):  # Python scalars as h5py stores them
            arr = np.asarray(value, dtype=np.int64)
        elif isinstance(value, float):
            arr = np.asarray(value, dtype=np.float64)
        else:
            arr = np.asarray(value)
            if arr.ndim:
                arr = np.ascontiguousarray(arr)
        if arr.dtype not in _CODES or arr.ndim > 1:
            raise ValueError(f"attribute {name!r}: dtype {arr.dtype} / rank {arr.ndim} not supported")
        N.check(N.lib().cgb_h5_set_attr(self._h, _b(obj), _b(name), _CODES[arr.dtype],
                                        -1 if arr.ndim == 0 else arr.shape[0], c_void_p(arr.ctypes.data)))

    def close(self) -> None:
        if self._h:
            h, self._h = self._h, c_void_p()
            N.check(N.lib().cgb_h5_close(h))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None:
            self.close()
        elif self._h:  # keep the original exception
            h, self._h = self._h, c_void_p()
            N.lib().cgb_h5_close(h)


class H5Reader:
    def __init__(self, path):
        self._h = c_void_p()
        N.check(N.lib().cgb_h5_open(_b(str(path)), byref(self._h)))

    def keys(self, group: str = "/") -> list:
        n = c_int()
        N.check(N.lib().cgb_h5_num_children(self._h, _b(group), byref(n)))
        buf = ctypes.create_string_buffer(1024)
        out = []
        for i in range(n.value):
            N.check(N.lib().cgb_h5_child_name(self._h, _b(group), i, buf, 1024, None))
            out.append(buf.value.decode("utf-8"))
        return out

    def is_group(self, group: str, name: str) -> bool:
        n = c_int()
        N.check(N.lib().cgb_h5_num_children(self._h, _b(group), byref(n)))
        buf = ctypes.create_string_buffer(1024)
        flag = c_int()
        for i in range(n.value):
            N.check(N.lib().cgb_h5_child_name(self._h, _b(group), i, buf, 1024, byref(flag)))
            if buf.value.decode("utf-8") == name:
                return bool(flag.value)
        raise KeyError(name)

    def info(self, path: str):
        """(numpy dtype, shape, absolute file offset of the data or -1)"""
        dt, rank, off = c_int(), c_int(), c_int64()
        dims = (c_int64 * 4)()
        N.check(N.lib().cgb_h5_dataset_info(self._h, _b(path), byref(dt), byref(rank), dims, byref(off)))
        return _DTYPES[dt.value], tuple(dims[i] for i in range(rank.value)), off.value

    def read(self, path: str) -> np.ndarray:
        dtype, shape, _ = self.info(path)
        out = np.empty(shape, dtype=dtype)
        N.check(N.lib().cgb_h5_read_dataset(self._h, _b(path), c_void_p(out.ctypes.data), out.nbytes))
        return out

    def read_device(self, path: str, device, dtype=torch.float64, scale: float = 1.0) -> torch.Tensor:
        """f32/f64 dataset -> CUDA tensor of ``dtype`` (values times ``scale``) without a whole-field
        host copy; an ``(n, k)`` dataset comes back as an ``(n, k)`` view of ``(k, n)`` planes."""
        _, shape, _ = self.info(path)
        if len(shape) not in (1, 2):
            raise ValueError("read_device takes datasets of rank 1 or 2")
        n, k = shape[0], (shape[1] if len(shape) == 2 else 1)
        device = torch.device(device)
        out = torch.empty((k, n) if len(shape) == 2 else (n,), dtype=dtype, device=device)
        with torch.cuda.device(device):
            N.check(N.lib().cgb_h5_read_dataset_device(
                self._h, _b(path), c_void_p(out.data_ptr()), 1 if dtype == torch.float32 else 2, n, float(scale),
                c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        return out.T if len(shape) == 2 else out

    def attrs(self, obj: str) -> dict:
        n = c_int()
        N.check(N.lib().cgb_h5_num_attrs(self._h, _b(obj), byref(n)))
        buf = ctypes.create_string_buffer(1024)
        out = {}
        for i in range(n.value):
            N.check(N.lib().cgb_h5_attr_name(self._h, _b(obj), i, buf, 1024))
            name = buf.value
            dt, cnt, nb = c_int(), c_int64(), c_size_t()
            N.check(N.lib().cgb_h5_attr_info(self._h, _b(obj), name, byref(dt), byref(cnt), byref(nb)))
            if dt.value == 0:
                continue  # type outside the subset (e.g. unyt's pickled unit registry)
            raw = ctypes.create_string_buffer(max(nb.value, 1))
            N.check(N.lib().cgb_h5_read_attr(self._h, _b(obj), name, raw, nb.value))
            if dt.value == H5_STR:
                out[name.decode("utf-8")] = raw.raw[:nb.value].decode("utf-8")
            else:
                arr = np.frombuffer(raw.raw[:nb.value], dtype=_DTYPES[dt.value]).copy()
                out[name.decode("utf-8")] = arr[0] if cnt.value < 0 else arr
        return out

    def close(self) -> None:
        if self._h:
            h, self._h = self._h, c_void_p()
            N.check(N.lib().cgb_h5_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
