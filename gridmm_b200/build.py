"""Compile the CUDA sources of gridmm_b200 into ONE in-tree shared library (C ABI, no torch/pybind types).

    python -m gridmm_b200.build            # builds gridmm_b200/lib/libgridmm_b200.so for sm_100a

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libgridmm_b200.so"
SOURCES = ["host_util.cu", "grid.cu", "pool.cu", "gemm_tc.cu", "gemm_ln.cu", "attn.cu", "attn_tc.cu", "rowops.cu", "heads.cu", "optim.cu", "train.cu", "graph.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177",
]


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; gridmm_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(dig):
    stamp = os.path.join(LIBDIR, "build.sha256")
    return os.path.exists(lib_path()) and os.path.exists(stamp) and open(stamp).read().strip() == dig


def build(force=False, verbose=False):
    """Build the library if the sources changed.  Returns the path of the .so.  Safe to call from several processes at once
    (one rank per GPU): builders serialise on a file lock, compile into a per-process object directory and publish the library
    with an atomic rename."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    dig = _digest()
    if not force and _up_to_date(dig):
        return lib_path()
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(dig):      # another process built it while this one waited
                return lib_path()
            return _build_locked(dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig, verbose):
    stamp = os.path.join(LIBDIR, "build.sha256")
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, "obj", "pid%d" % os.getpid())
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out:
            sys.stderr.write(out)
        objs.append(obj)
    tmp = lib_path() + ".tmp%d" % os.getpid()
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    os.replace(tmp, lib_path())
    with open(stamp, "w") as f:
        f.write(dig)
    shutil.rmtree(objdir, ignore_errors=True)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
