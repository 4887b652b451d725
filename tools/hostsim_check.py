"""Run the lane-serial host build of the log-mel building blocks against the reference goldens."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import logmel_oracle as lo

def build():
    so = os.path.join(ROOT, "oracle", "_build", "libhostsim.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "logmel_hostsim.cpp")])
    lib = ctypes.CDLL(so)
    lib.hostsim_logmel.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p]
    return lib

def run(lib, x, win, fb):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty((1 + x.shape[0] // 256, 256), np.float32)
    rc = lib.hostsim_logmel(x.ctypes.data, x.shape[0], win.ctypes.data, fb.ctypes.data, 1e-8, out.ctypes.data)
    assert rc == 0
    return out

if __name__ == "__main__":
    g = np.load(os.path.join(ROOT, 'tests/golden/logmel.npz')); fbz = np.load(os.path.join(ROOT, 'tests/golden/mel_fb.npz'))
    fb = np.zeros((1025, 256), np.float32); off = 0
    for m, (s, l) in enumerate(zip(fbz['start'], fbz['length'])):
        fb[s:s + l, m] = fbz['weights'][off:off + l]; off += l
    win = np.ascontiguousarray(fbz['window'])
    lib = build()
    for k in g.files:
        if not k.startswith('pcm_'): continue
        name = k[4:]; x = g[k].astype(np.float32) / 32768.0; ref = g['feat_' + name]
        out = run(lib, x, win, fb)
        print(name, lo.close_logmel(out, ref), lo.close_logmel(out, ref, fft_noise=256.0))
