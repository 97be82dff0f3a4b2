"""ctypes access to fade_b200/csrc/emu/libfadeemu.so (host lock-step emulation of one kernel group).

Test infrastructure: lets the CPU suite check the shared kernel core against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "fade_b200", "csrc", "emu")
EMU_SO = os.path.join(EMU_DIR, "libfadeemu.so")
OPS_CAP = 10

CODE = {"A": 0, "C": 1, "T": 2, "G": 3, "N": 4}


class AlnOut(C.Structure):
    _fields_ = [("score", C.c_int32), ("end_query", C.c_int32), ("end_ref", C.c_int32),
                ("beg_query", C.c_int32), ("beg_ref", C.c_int32), ("n_ops", C.c_int32),
                ("flags", C.c_uint32), ("read", C.c_int32), ("ops", C.c_uint32 * OPS_CAP)]


_lib = None


def build_emu(force=False):
    srcs = [os.path.join(EMU_DIR, "emu.cu"), os.path.join(EMU_DIR, "..", "sw_core.cuh")]
    if force or not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                               "-shared", "-o", EMU_SO, srcs[0]])
    return EMU_SO


def lib():
    global _lib
    if _lib is None:
        build_emu()
        L = C.CDLL(EMU_SO)
        L.fadeemu_align_pair.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.POINTER(AlnOut), C.POINTER(AlnOut)]
        L.fadeemu_align_pair.restype = C.c_int
        L.fadeemu_prmt.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.fadeemu_prmt.restype = C.c_uint32
        assert L.fadeemu_sizeof_alnout() == C.sizeof(AlnOut)
        _lib = L
    return _lib


def codes(s: str) -> np.ndarray:
    return np.array([CODE.get(c.upper(), 5) for c in s], dtype=np.uint8)


def align_pair(R, qa, ta, qb, tb, clips=(0, 0, 0, 0), scoring=(10, 2, 2, -3), extra_blocks=0, min_length=5,
               tagged=1):
    a, b = AlnOut(), AlnOut()
    ca, cta, cb, ctb = codes(qa), codes(ta), codes(qb), codes(tb)
    cl = np.array(clips, dtype=np.uint32)
    rc = lib().fadeemu_align_pair(R, ca.ctypes.data, len(ca), cta.ctypes.data, len(cta), cb.ctypes.data, len(cb),
                                  ctb.ctypes.data, len(ctb), *scoring, extra_blocks, tagged, min_length, cl.ctypes.data,
                                  C.byref(a), C.byref(b))
    if rc:
        raise RuntimeError(f"fadeemu_align_pair rc={rc}")
    return a, b
