"""Builds annchor_b200/libannb.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libannb.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "-shared",
    "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: build a variant library (e.g. -DANNB_VARIANT=...) next to the default one;
    select it at run time with ANNB_LIBRARY=<path>."""
    global SO
    if out is not None:
        return _build_variant(list(extra_flags), out, verbose)
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        deps = [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + \
            glob.glob(os.path.join(HERE, "..", "include", "*.h"))
        if not force and os.path.exists(obj) and all(os.path.getmtime(obj) > os.path.getmtime(p) for p in deps):
            continue
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart",
                           "static", "-o", SO] + objs)
    return SO


def _build_variant(extra, out, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    bdir = os.path.join(HERE, "build", os.path.basename(out).replace(".", "_"))
    os.makedirs(bdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared",)] + extra
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc] + flags + ["-c", src, "-o", obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT)))
    for src, p in procs:
        o = p.communicate()[0].decode()
        if p.returncode != 0:
            sys.stderr.write(o)
            raise RuntimeError("nvcc failed on %s" % src)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                           "-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
