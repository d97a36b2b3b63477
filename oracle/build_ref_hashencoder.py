"""Compiles the REFERENCE's own hash-grid kernels (/root/reference/code/hashencoder/src/{hashencoder.cu,bindings.cpp},
sources used where they lie, untouched) into oracle/_ref/_hash_encoder_ref*.so -- a checker for the -m gpu tests:
tests/test_gpu_hashgrid.py compares msdf_hash_encode_* with these kernels on the same inputs.

TEST INFRASTRUCTURE ONLY (only tests/ may import the result).  Container-only recipe: /root/reference does not exist on
the GPU box, the built module travels there with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
The reference builds the same two files with torch.utils.cpp_extension.load (hashencoder/backend.py:10-24); the only
deviations are flags: -std=c++17 instead of c++14 (torch 2.11's headers need it) and an explicit sm_100 target.

  python oracle/build_ref_hashencoder.py [reference_root]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "_hash_encoder_ref"


def build(ref_root="/root/reference"):
    src_dir = os.path.join(ref_root, "code", "hashencoder", "src")
    if not os.path.isdir(src_dir):
        return None                       # GPU box: nothing to build, the prebuilt module (if any) is used
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, NAME + sysconfig.get_config_var("EXT_SUFFIX"))
    srcs = [os.path.join(src_dir, "hashencoder.cu"), os.path.join(src_dir, "bindings.cpp")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs + [__file__]):
        return out
    inc = []
    for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
        inc += ["-I", p]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-O3", "-std=c++17", "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H", abi] + inc
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s) + ".o")
        if s.endswith(".cu"):
            cmd = [nvcc, "-c", s, "-o", o, "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC",
                   "-allow-unsupported-compiler", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                   "-U__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr"] + common
        else:
            cmd = ["g++", "-c", s, "-o", o, "-fPIC"] + common
        subprocess.check_call(cmd)
        objs.append(o)
    lib_dir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["g++", "-shared", "-o", out] + objs + ["-L" + lib_dir, "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda",
                                                                 "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                                                                 "-Wl,-rpath," + lib_dir])
    for o in objs:
        os.remove(o)
    return out


def load():
    """The built module, or None when it has not been built (import torch first: it needs libtorch loaded)."""
    import glob
    import importlib.util
    import torch  # noqa: F401
    hits = glob.glob(os.path.join(OUT_DIR, NAME + "*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location(NAME, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
