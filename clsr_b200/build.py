"""Build libclsr_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

Every source is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into one shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "build")
LIB = os.path.join(HERE, "libclsr_b200.so")
SOURCES = ["engine.cu", "shard.cu", "crc32c.cu", "evalmetrics.cu", "batchbuild.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "clsr_b200.h"))
    return hs


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _extra():
    cmd = os.environ.get("CLSR_NVCC_EXTRA", "").split()   # developer knobs (-DCLSR_TC_UN1=8 ...)
    if os.environ.get("CLSR_NCCL", "1") != "0" and os.path.exists("/usr/include/nccl.h"):
        cmd += ["-DCLSR_WITH_NCCL"]
    return cmd


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    extra = _extra()
    tag = os.path.join(OBJ, "flags.txt")
    flag_sig = " ".join(FLAGS + extra)
    if not os.path.exists(tag) or open(tag).read() != flag_sig:
        force = True

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_t):
            return o, None
        cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", s, "-o", o]
        return o, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(compile_one, _sources()))
    for o, r in res:
        if r is not None and r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building %s" % o)
        if r is not None and verbose:
            print(r.stderr)
    with open(tag, "w") as f:
        f.write(flag_sig)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [o for o, _ in res] + ["-o", LIB],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libclsr_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
