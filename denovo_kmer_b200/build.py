"""In-tree build of libdkb.so for sm_100a (explicit nvcc; the .so travels with gpurun)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(_HERE, "csrc", f) for f in ("dkb_api.cu", "dkb_host.cpp")]
DEPS = SRC + [os.path.join(_HERE, "csrc", f) for f in
              ("dkb_device.cuh", "dkb_scan.cuh", "dkb_build.cuh", "dkb_pack.cuh")] + [
    os.path.join(_HERE, "..", "include", "dkb.h")]
OUT = os.path.join(_HERE, "libdkb.so")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
              "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++", "-Xcompiler", "-pthread"]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(
            os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
