"""Build libhands_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python -m hands_b200._build            # or: from hands_b200._build import build; build()
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libhands_b200.so")
SOURCES = ["hb_api.cu", "mano_kernels.cu", "mano_tc.cu", "pcl_kernels.cu", "pcl_setup.cu", "loss_kernels.cu", "silhouette.cu"]
PER_FILE_FLAGS = {"pcl_setup.cu": ["-fmad=false"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libhands_b200.so cannot be built")
    return exe


STAMP = os.path.join(HERE, "lib", "build.sha256")
HEADER = os.path.join(HERE, "..", "include", "hands_b200.h")


def source_digest():
    """sha256 over every file the library is built from plus the flags.  Stored beside the .so at build time; a
    mismatch means the binary is stale (mtimes do not survive the copy to a GPU box, content does)."""
    import hashlib

    h = hashlib.sha256()
    for path in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [HEADER]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(repr((NVCC_FLAGS, sorted(PER_FILE_FLAGS.items()))).encode())
    return h.hexdigest()


def header_version():
    import re

    with open(HEADER) as fh:
        return int(re.search(r"#define\s+HB_VERSION\s+(\d+)", fh.read()).group(1))


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_digest()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(HERE, "lib", src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + PER_FILE_FLAGS.get(src, ["-fmad=true"]) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(HERE, "lib", "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(source_digest())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
