"""ctypes binding of oracle/logmel_oracle.c (TEST INFRASTRUCTURE; see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblogmel_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def _lib():
    if not os.path.isfile(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.oracle_logmel_num_frames.restype = ctypes.c_long
    lib.oracle_logmel_num_frames.argtypes = [ctypes.c_long]
    lib.oracle_logmel_f32.restype = ctypes.c_int
    lib.oracle_logmel_f32.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
    return lib


def logmel(wave, window, fb, log_offset=1e-8, n_threads=None):
    lib = _lib()
    x = np.ascontiguousarray(wave, dtype=np.float32)
    w = np.ascontiguousarray(window, dtype=np.float32)
    f = np.ascontiguousarray(fb, dtype=np.float32)
    T = lib.oracle_logmel_num_frames(x.shape[0])
    out = np.empty((T, 256), dtype=np.float32)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    rc = lib.oracle_logmel_f32(x.ctypes.data, x.shape[0], w.ctypes.data, f.ctypes.data, log_offset, out.ctypes.data,
                               int(n_threads))
    assert rc == 0
    return out
