"""Build libb2dglue.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

Every ``csrc/*.cu`` is compiled to its own object (in parallel, rebuilt only when the source or a header
changed) and linked into one shared library.  A/B builds for profiling pass extra ``-D`` macros and another
output path; the shipped library is always built without any.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2dglue.so")
OBJ_DIR = os.path.join(CSRC, "_build")
SOURCES = ["lib.cu", "proposal.cu", "nms.cu", "roi_align.cu", "roi_align_rows.cu", "roi_align_bwd_rows.cu",
           "codecs.cu", "targets.cu", "uncertainty.cu", "detections.cu", "bev.cu", "head_tail.cu", "eval.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]
CC_FLAGS = ARCH_FLAGS + ["-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
LINK_FLAGS = ["--shared", "-cudart", "shared"] + ARCH_FLAGS


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "b2d_glue.h"), os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, [os.path.join(CSRC, s) for s in SOURCES] + _headers())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compile + link.  ``defines`` (e.g. ["B2D_AB_NOSTORE=1"]) and ``out`` are for A/B builds only."""
    out = out or LIB
    ab = bool(defines) or out != LIB
    if not force and not ab and not needs_build():
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_dir = OBJ_DIR + ("_ab" if ab else "")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = _headers()
    dflags = [f"-D{d}" for d in defines]

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src[:-3] + ".o")
        if not (force or ab) and not _stale(o, [s] + hdrs):
            return o, ""
        cmd = [nvcc] + CC_FLAGS + dflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}{res.stderr}")
        return o, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    res = subprocess.run([nvcc] + LINK_FLAGS + [o for o, _ in results] + ["-o", out], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libb2dglue.so")
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
