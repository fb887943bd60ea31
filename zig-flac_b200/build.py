"""Builds the C-ABI boundary with nvcc for sm_100a, in-tree:

    libzigflac_b200.so   shared library (ctypes tests, the `flac` CLI, bench.py)
    libzigflac_b200.a    static library -- what BASELINE.json's north_star names: a Zig (or C) host links it from
                         build.zig with addObjectFile + cudart_static
    flac                 the CLI with the reference's argv / exit-code contract (src/cli.zig)

    python zig-flac_b200/build.py [--force] [-v]

Both libraries hold the same objects and link cudart statically; there is no torch dependency.  nvcc cross-compiles
sm_100a without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libzigflac_b200.so")
STATIC = os.path.join(HERE, "libzigflac_b200.a")
CLI = os.path.join(HERE, "flac")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

CU = ["zf_capi.cu", "zf_decode.cu"]
CPP = ["zf_host.cpp", "zf_driver.cpp"]
C_SRC = ["zf_synth.c"]


def _headers():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hdr + [os.path.join(HERE, "..", "include", "zigflac_b200.h"), os.path.abspath(__file__)]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers()
    objs = []
    for f in CU + CPP + C_SRC:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, os.path.splitext(f)[0] + ".o")
        objs.append(obj)
        if not (force or _stale(obj, [src] + hdr)):
            continue
        if f.endswith(".c"):
            # -ffp-contract=off: the generator must round identically wherever it is built
            cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-c", "-o", obj, src]
        else:
            # -fmad=false: the LPC extension's double arithmetic is one IEEE rounding per operation (zf_kernel_lpc.cuh); nothing
            # else in the kernels is floating point
            cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo", "-fmad=false"] + ARCH + ["-Xcompiler", "-fPIC,-pthread", "-c", "-o", obj, src]
            if verbose and f.endswith(".cu"):
                cmd[1:1] = ["-Xptxas", "-v"]
        subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        subprocess.run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-Xcompiler", "-pthread", "-o", LIB] + objs + ["-lm"],
                       check=True)
    if force or _stale(STATIC, objs):
        if os.path.exists(STATIC):
            os.remove(STATIC)
        subprocess.run(["ar", "rcs", STATIC] + objs, check=True)
    if force or _stale(CLI, [os.path.join(CSRC, "zf_cli.cpp"), LIB]):
        subprocess.run(["g++", "-O2", "-o", CLI, os.path.join(CSRC, "zf_cli.cpp"), "-L" + HERE,
                        "-lzigflac_b200", "-Wl,-rpath,$ORIGIN", "-pthread"], check=True)
    return LIB


def link_static_check(out_path):
    """Links a C++ program against libzigflac_b200.a the way a Zig build would (static cudart from the toolkit), as a check
    that the archive is self-contained.  Returns the path of the binary."""
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    subprocess.run(["g++", "-O2", "-o", out_path, os.path.join(CSRC, "zf_cli.cpp"), STATIC, "-L" + cuda_lib,
                    "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lm"], check=True)
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
