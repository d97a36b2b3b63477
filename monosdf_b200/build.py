"""Builds libmonosdf_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

`python -m monosdf_b200.build` or `__graft_entry__.build()`.  The shared library only links libcudart; its
C ABI is declared in include/monosdf_b200.h and bound with ctypes in monosdf_b200/_lib.py.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmonosdf_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-DMSDF_BUILDING"]
# per-file extra flags: the sampler's bit-exact contract forbids FMA contraction (see include/msdf_detmath.h)
EXTRA = {"sampler.cu": ["-fmad=false"]}


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stamp(src, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for fn in sorted(os.listdir(root)):
            if fn.endswith((".cuh", ".h")) or os.path.join(root, fn) == src:
                with open(os.path.join(root, fn), "rb") as f:
                    h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs, rebuilt = [], False
    procs = []
    for fn in sorted(os.listdir(CSRC)):
        if not fn.endswith(".cu"):
            continue
        src = os.path.join(CSRC, fn)
        obj = os.path.join(OBJ_DIR, fn[:-3] + ".o")
        flags = ARCH + COMMON + EXTRA.get(fn, [])
        if verbose:
            flags = flags + ["-Xptxas", "-v"]
        stamp_file = obj + ".stamp"
        stamp = _stamp(src, flags)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        procs.append((fn, stamp_file, stamp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for fn, stamp_file, stamp, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % fn)
        if verbose or "warning" in out:
            sys.stderr.write("[%s]\n%s" % (fn, out))
        with open(stamp_file, "w") as f:
            f.write(stamp)
        rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
