"""In-tree build of libhawq_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
OUT = os.path.join(_HERE, "libhawq_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def up_to_date():
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in _deps())


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile every .cu under csrc/ into hawq_b200/libhawq_b200.so.  Returns the path.
    Safe when several processes call it at once (one rank per GPU under torchrun): an exclusive file lock serialises them, the
    winner compiles into a temporary file and renames it into place, the others find the library up to date."""
    if not force and up_to_date():
        return OUT
    import fcntl
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():               # another process built it while we waited
                return OUT
            tmp = "%s.tmp.%d" % (OUT, os.getpid())
            cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-I", INCLUDE, "-o", tmp] + sources()
            if verbose:
                print(" ".join(cmd).replace(tmp, OUT))
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout)
            os.replace(tmp, OUT)
            if verbose and r.stdout:
                print(r.stdout)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT
