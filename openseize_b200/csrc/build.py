"""Compile the sm_100a kernels into openseize_b200/_lib/libosz_b200.so.

nvcc cross-compiles without a GPU.  The library is built in-tree (git-ignored,
but it travels to the GPU box with the gpurun snapshot).
"""

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libosz_b200.so")
SOURCES = ["runtime.cu", "fir.cu", "sos.cu", "tf.cu", "upfirdn.cu", "sosdec.cu", "spectra.cu",
           "spectra_generic.cu", "spectra_mixed.cu", "protools.cu"]
HEADERS = ["common.cuh", "sos_core.cuh", "sos_tile.cuh", "ufd_mma.cuh", "fft_core.cuh", "fft_core_body.inc", os.path.join("..", "..", "include", "osz_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr", "-diag-suppress", "20208",
]


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(HERE, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Build the shared library if any source changed.  Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("openseize_b200: CUDA build failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                 "-lcudart"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
