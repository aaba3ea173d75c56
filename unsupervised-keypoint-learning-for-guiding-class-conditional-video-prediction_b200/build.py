"""In-tree nvcc build of libkp_b200.so (sm_100a only; the .so travels to the GPU box with the snapshot)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkp_b200.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the B200 path has no fallback")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_trace_variant():
    """Debug variant with -DKP_TRACE (per-tile timeline stamps, KP_TAPCONV_TRACE=1) as libkp_b200_trace.so; selected with
    the environment variable KP_B200_LIB.  Not part of the product build."""
    out = os.path.join(HERE, "build", "libkp_b200_trace.so")
    objs = []
    for src in _sources():
        obj = os.path.join(HERE, "build", "trace_" + os.path.basename(src)[:-3] + ".o")
        subprocess.check_call([_nvcc(), *NVCC_FLAGS, "-DKP_TRACE", "-c", src, "-o", obj])
        objs.append(obj)
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs, "-lcudart"])
    return out


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link them into libkp_b200.so. Returns the library path."""
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s\n" % src)
    if failed:
        raise RuntimeError("nvcc build failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    if "--trace" in sys.argv:
        os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
        print(build_trace_variant())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
