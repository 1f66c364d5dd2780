"""Build libjudi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjudi_b200.so")
SOURCES = ["capi.cu", "model.cu", "sparse.cu", "kernels.cu", "tma3d.cu", "tti3d.cu", "varc3d.cu", "flux3d.cu", "persist2d.cu", "propagate.cu", "hostops.cu", "comm.cu", "isic3d.cu"]
HEADERS = ["common.cuh", "tma_util.cuh", "propagate.h", os.path.join("..", "..", "include", "judi_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr", "-cudart", "static",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False):
    if not force and not stale():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose or ptxas_info:
            sys.stdout.write(out.decode())
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = ["nvcc", "-shared", "-cudart", "static", "-o", LIB] + objs + ["-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(LIB)
