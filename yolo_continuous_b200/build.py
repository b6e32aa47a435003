"""Build csrc/*.cu into yolo_continuous_b200/csrc/libyc_b200.so for sm_100a (in-tree, so the
shared object travels to the GPU box with the repo snapshot)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libyc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Wno-deprecated-gpu-targets",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


EXTRA = os.environ.get("YC_NVCC_FLAGS", "").split()   # experiments only (e.g. -DT2_BK=64)


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "yc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        cmd = [NVCC] + FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        ok = ok and p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-cudart", "static", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
