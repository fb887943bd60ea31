"""Builds libzigflac_b200.so (sm_100a, in-tree) and the `flac` CLI with nvcc.

    python zig-flac_b200/build.py [--force]

The shared library is the C-ABI drop-in boundary (include/zigflac_b200.h).  It links cudart
statically and has no torch dependency.  nvcc cross-compiles sm_100a without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzigflac_b200.so")
CLI = os.path.join(HERE, "flac")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    cu = [os.path.join(CSRC, f) for f in ("zf_capi.cu",)]
    cpp = [os.path.join(CSRC, f) for f in ("zf_host.cpp", "zf_driver.cpp")]
    c = [os.path.join(CSRC, "zf_synth.c")]
    hdr = [os.path.join(CSRC, f) for f in ("zf_kernel.cuh", "zf_kernel_indep.cuh", "zf_kernel_full.cuh", "zf_kernel_v3.cuh", "zf_dev.h")] + [
        os.path.join(HERE, "..", "include", "zigflac_b200.h")]
    if force or _stale(LIB, cu + cpp + c + hdr + [os.path.abspath(__file__)]):
        synth_o = os.path.join(CSRC, "zf_synth.o")
        # -ffp-contract=off: the generator must round identically wherever it is built
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-c", "-o", synth_o] + c, check=True)
        cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo"] + ARCH + [
            "-Xcompiler", "-fPIC,-pthread", "-shared", "-cudart", "static",
            "-o", LIB] + cu + cpp + [synth_o, "-lm"]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.run(cmd, check=True)
    if force or _stale(CLI, [os.path.join(CSRC, "zf_cli.cpp"), LIB]):
        subprocess.run(["g++", "-O2", "-o", CLI, os.path.join(CSRC, "zf_cli.cpp"), "-L" + HERE,
                        "-lzigflac_b200", "-Wl,-rpath,$ORIGIN", "-pthread"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
