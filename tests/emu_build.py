"""Builds tests/emu/libqremu.so (g++), the CPU emulation of the product's work-item functions."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "qr_emu.cpp")
LIB = os.path.join(HERE, "emu", "libqremu.so")
CSRC = os.path.join(os.path.dirname(HERE), "qam-reconciliation_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    return LIB


def load():
    L = C.CDLL(build())
    L.emu_last_error.restype = C.c_char_p
    return L
