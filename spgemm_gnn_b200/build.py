"""In-tree build of libmaxk_b200.so with nvcc for sm_100a (no torch headers involved).

The shared object sits next to this file; it is git-ignored but travels to the GPU box.
cudart is linked statically so that the library does not depend on which libcudart the
host process (torch) happens to have loaded -- the boundary is a C ABI on raw pointers.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libmaxk_b200.so")
SOURCES = ["api.cu", "topk.cu", "topk_tile.cu", "cbsr.cu", "partition.cu", "spgemm_fwd.cu", "sspmm_bwd.cu", "bank.cu", "banked.cu", "layernorm.cu", "peer.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP = SO + ".srchash"


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to
    the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [
        os.path.join(ROOT, "include", "maxk_b200.h"), os.path.abspath(__file__)]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(SO) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False, defs=(), out: str = None) -> str:
    """`defs` / `out` build an experimental variant (extra -D flags) next to the product library;
    `MAXK_LIB=<path>` makes `_lib` load it (tuning runs only)."""
    so = out or SO
    if not defs and not out and not force and not _stale():
        return SO
    nvcc = nvcc_path()
    objdir = os.path.join(HERE, "build" if not out else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    common = [
        nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-ccbin", "/usr/bin/g++",
    ]
    common += list(defs)
    if verbose:
        common += ["-Xptxas", "-v"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        log = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{log}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, *ARCH, "-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++",
                           "-o", so] + objs)
    if so == SO and not defs:
        with open(STAMP, "w") as f:
            f.write(_source_hash())
    return so


BINDING_SRC = os.path.join(HERE, "binding", "maxk_bindings.cpp")


def binding_path() -> str:
    import sysconfig
    return os.path.join(HERE, "maxk_kernels_ext" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_binding(force: bool = False) -> str:
    """The compiled `maxk_kernels` front end (pybind11 / ATen, binding/maxk_bindings.cpp) over
    libmaxk_b200.so -- what the reference's setup.py:25-31 builds, minus the kernels, which stay behind
    the C ABI.  One g++ call with torch's own include / library paths (binding/setup.py is the
    setuptools form of the same build); in-tree, next to the library, so that it travels with it."""
    import hashlib
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    build()
    out = binding_path()
    stamp = out + ".srchash"
    h = hashlib.sha256()
    for f in (BINDING_SRC, os.path.join(ROOT, "include", "maxk_b200.h")):
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(torch.__version__.encode())
    digest = h.hexdigest()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return out
    inc = ce.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"], os.path.join(ROOT, "include")]
    libs = ce.library_paths(device_type="cuda")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=maxk_kernels_ext",
            "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
           + [f"-I{i}" for i in inc] + [BINDING_SRC, "-o", out] + [f"-L{l}" for l in libs]
           + [f"-L{HERE}", "-lmaxk_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
              "-ltorch_python", "-Wl,-rpath," + HERE] + [f"-Wl,-rpath,{l}" for l in libs])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode())
        raise RuntimeError("building the maxk_kernels binding failed")
    with open(stamp, "w") as f:
        f.write(digest)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defs=defs, out=out))
    if "--binding" in sys.argv:
        print(build_binding(force="--force" in sys.argv))
