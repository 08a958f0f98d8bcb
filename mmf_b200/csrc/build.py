"""Builds libmmfb200.so (sm_100a only) in-tree with nvcc.

    python -m mmf_b200.csrc.build [--force] [--verbose]
    python -m mmf_b200.csrc.build --variant NAME -DFLAG=1 ...   # libmmfb200_NAME.so with extra defines, for A/B runs
                                                                # (select it at run time with MMFB_LIB=<path>)

The built library lives next to the sources (git-ignored, but it travels to the GPU box with
the gpurun snapshot).  There is exactly one target architecture: compute_100a / sm_100a.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["capi.cu", "gemm.cu", "attention.cu", "rowops.cu", "optim.cu", "ce.cu"]
HEADERS = ["common.cuh", "mmfb_internal.h", "../../include/mmfb200.h"]
LIB = os.path.join(HERE, "libmmfb200.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def _digest():
    h = hashlib.sha256()
    for f in _sources() + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, variant=None, defines=()):
    """Compile every .cu into objects (parallel) and link the shared library. Returns its path.
    `variant` + `defines`: an experimental build next to the product library (never loaded unless MMFB_LIB names it)."""
    if variant:
        return _build_variant(variant, list(defines), verbose)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose:
            sys.stderr.write(out)
        else:
            # always surface spills: they are performance bugs
            for line in out.splitlines():
                if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                    sys.stderr.write("[%s] %s\n" % (src, line))
    if failed:
        raise RuntimeError("building libmmfb200.so failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("linking libmmfb200.so failed:\n" + r.stdout)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


def _build_variant(name, defines, verbose):
    nvcc = _nvcc()
    out = os.path.join(HERE, "libmmfb200_%s.so" % name)
    objs, procs = [], []
    for src in _sources():
        obj = os.path.join(HERE, src.replace(".cu", ".%s.o" % name))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + defines + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s (%s):\n%s" % (src, name, o))
        for line in o.splitlines():
            if verbose or ("spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line):
                sys.stderr.write("[%s %s] %s\n" % (name, src, line))
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                             "-cudart", "static"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("linking %s failed:\n%s" % (out, r.stdout))
    for o in objs:
        os.remove(o)
    return out


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant,
                 defines=[a for a in sys.argv[1:] if a.startswith("-D")])
    print(path)
