"""TEST INFRASTRUCTURE ONLY: builds / loads tests/emul/libjiminy_b200_emul.so (the product's C ABI
and device code compiled for the host, warp lanes emulated by threads -- see jb_emul_shim.h) so
that CPU-only tests can run the kernel source against the oracle.  Never imported by jiminy_b200/."""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)
from jiminy_b200.core import Api  # noqa: E402

_LIB = os.path.join(_HERE, "libjiminy_b200_emul.so")
_SRCS = [os.path.join(_HERE, f) for f in ("jb_emul.cpp", "jb_emul_shim.h")] + \
        [os.path.join(_ROOT, "jiminy_b200", "csrc", f) for f in
         ("jb_capi.cu", "jb_kernel.cuh", "jb_device.cuh", "jb_constraints.cuh", "jb_constraints_quadruped.cuh", "jb_constraints_blocks.cuh", "jb_constraints_bodies.cuh", "jb_plan.cpp", "jb_plan.h")] + \
        [os.path.join(_ROOT, "include", "jiminy_b200.h")]


def build(force=False, fma=False):
    """`fma=True`: multiply-adds contracted like nvcc does for the device (needs an x86 host with FMA): the rounding
    of the GPU build, to check that the parity tolerances of the `-m gpu` suite hold before GPU time is spent."""
    lib = _LIB.replace(".so", "_fma.so") if fma else _LIB
    stale = not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in _SRCS)
    if force or stale:
        fp = ["-mfma", "-ffp-contract=fast"] if fma else ["-ffp-contract=off"]
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-DJB_HOST_EMUL=1"] + fp +
                       ["-I", _HERE, "-x", "c++", os.path.join(_HERE, "jb_emul.cpp"),
                        os.path.join(_ROOT, "jiminy_b200", "csrc", "jb_plan.cpp"), "-o", lib], check=True)
    return lib


_api = {}


def host_has_fma() -> bool:
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def emul_api(fma: bool = False) -> Api:
    if fma not in _api:
        _api[fma] = Api(C.CDLL(build(fma=fma)))
    return _api[fma]
