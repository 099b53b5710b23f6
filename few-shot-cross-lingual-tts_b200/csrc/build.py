"""Build libfs2b200.so in-tree with nvcc for sm_100a (no torch headers: the ABI is plain C).

Usage: python build.py [--force] [--verbose]
The .so stays next to the sources so that it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfs2b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(HERE, "..", "..", "include", "fs2b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(HERE, "build", src[:-3] + ".o")
    srcp = os.path.join(HERE, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), _deps_mtime()):
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", srcp, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    if force:
        for f in os.listdir(os.path.join(HERE, "build")):
            os.remove(os.path.join(HERE, "build", f))
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if verbose and log:
        print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        cmd = [NVCC, "-shared", "-o", SO, *objs, "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
