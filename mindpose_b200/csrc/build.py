"""Build libposecodec.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mindpose_b200.csrc.build [--force] [--verbose]

The shared library lands next to the sources (mindpose_b200/csrc/libposecodec.so)
so that it travels with the repo snapshot; it is git-ignored.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libposecodec.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")
SOURCES = [
    "lib.cu",
    "topdown_decode.cu",
    "topdown_encode.cu",
    "warp_affine.cu",
    "rescale.cu",
    "bottomup_decode.cu",
    "bottomup_encode.cu",
    "grouping.cu",
    "refine_missing.cu",
    "oks_nms.cu",
    "peer_gather.cu",
]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "posecodec.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact parity with the reference's op order: no fused multiply-add
    # contraction anywhere (the kernels are HBM-bound, not FMA-bound)
    "--fmad=false",
    "-Xcompiler", "-fPIC,-O2",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libposecodec needs the CUDA 12.9 toolkit")
    return exe


def _extra_defines():
    """Experiment switches (``PC_NVCC_DEFINES="-DPC_SOME_SWITCH=1"``): part of the
    fingerprint, so a library built with them is rebuilt by the next plain build()."""
    extra = os.environ.get("PC_NVCC_DEFINES", "").split()
    bad = [d for d in extra if not d.startswith("-D")]
    if bad:
        raise RuntimeError(f"PC_NVCC_DEFINES may only hold -D switches, got {bad}")
    return extra


def _fingerprint() -> str:
    h = hashlib.sha256()
    h.update(" ".join(_extra_defines()).encode())
    for rel in SOURCES + HEADERS + ["build.py"]:
        with open(os.path.join(HERE, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant: build an experiment library next to the shipped one
    (libposecodec_<variant>.so, objects under build/<variant>/; select it at run time with
    POSECODEC_LIB) without touching libposecodec.so or its stamp."""
    if variant:
        return _build_into(os.path.join(HERE, f"libposecodec_{variant}.so"),
                           os.path.join(HERE, "build", variant), verbose)
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB
    _build_into(LIB, os.path.join(HERE, "build"), verbose)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


def _build_into(lib_path: str, obj_dir: str, verbose: bool) -> str:
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *_extra_defines(), "-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src} ==\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(obj_dir, "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise RuntimeError(f"nvcc failed; see {os.path.join(obj_dir, 'nvcc.log')}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *objs,
            "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise RuntimeError(f"link of {os.path.basename(lib_path)} failed")
    return lib_path


if __name__ == "__main__":
    variant = ""
    if "--variant" in sys.argv:
        variant = sys.argv[sys.argv.index("--variant") + 1]
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant)
    print(path)
