"""CPU oracle for the reverse-reconciliation hot path -- TEST INFRASTRUCTURE ONLY.

`oracle.port` wraps oracle/qr_oracle.c (a C restatement of the reference's
algorithm, built with gcc by `oracle.build()`); `oracle.ref` runs the compiled,
unmodified reference (oracle/_ref, built by oracle/build_ref.py) in a
subprocess-safe way.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
package never does (tests/test_abi_and_host.py::test_product_never_touches_the_oracle enforces it).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libqroracle.so")
SRC = os.path.join(HERE, "qr_oracle.c")


def build(force: bool = False) -> str:
    """Compile oracle/qr_oracle.c -> oracle/libqroracle.so (plain gcc -O2, no -march:
    the reference is built with CPython's default -O2 and no FMA contraction)."""
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-o", LIB, SRC, "-lm"])
    return LIB
