"""numpy arrays whose last byte is followed by an inaccessible page: any read or write past the end of the buffer faults.
The CPU analogue of compute-sanitizer memcheck for kernels run on the host emulator (overruns past the end only)."""
import ctypes
import mmap

import numpy as np

_libc = ctypes.CDLL(None, use_errno=True)
_libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
_keep = []


def guarded(array: np.ndarray, align: int = 16) -> np.ndarray:
    """copy `array` so that it ends (up to `align` padding in front, never behind) right before a PROT_NONE page"""
    a = np.ascontiguousarray(array)
    page = mmap.PAGESIZE
    nbytes = max(a.nbytes, 1)
    data_pages = (nbytes + page - 1) // page
    m = mmap.mmap(-1, (data_pages + 1) * page)
    base = ctypes.addressof(ctypes.c_char.from_buffer(m))
    end = base + data_pages * page
    if _libc.mprotect(end, page, 0) != 0:
        raise OSError(ctypes.get_errno(), "mprotect failed")
    start = end - a.nbytes
    if start % align:
        raise ValueError(f"buffer of {a.nbytes} bytes cannot end at a page boundary and start {align}-byte aligned")
    out = np.frombuffer(m, dtype=a.dtype, count=a.size, offset=start - base).reshape(a.shape)
    out[...] = a
    _keep.append(m)
    return out
