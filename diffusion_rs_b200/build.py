"""Build libfluxb200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

The shared library is the product; there is no CPU fallback. `build()` is what `__graft_entry__.build()` calls.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libfluxb200.so"
OBJ_DIR = PKG / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _flags() -> list[str]:
    """NVCC_FLAGS plus opt-in debug instrumentation (FLUXB200_GEMM_TRACE=1: clock64 trace of the GEMM's MMA thread)."""
    extra = ["-DFB_GEMM_TRACE=1"] if os.environ.get("FLUXB200_GEMM_TRACE") == "1" else []
    if os.environ.get("FLUXB200_FULL_WAIT_CLUSTER") == "1":  # A/B build: round-1 cluster-scope wait in the GEMM main loop
        extra.append("-DFB_FULL_WAIT_CLUSTER=1")
    return NVCC_FLAGS + extra


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) +
                    list((PKG.parent / "include").glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, obj: Path, log: Path) -> tuple[Path, int, str]:
    cmd = [nvcc, *_flags(), "-I", str(PKG.parent / "include"), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(r.stdout + r.stderr)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = OBJ_DIR / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    srcs = _sources()
    objs = [OBJ_DIR / (s.stem + ".o") for s in srcs]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        futs = [ex.submit(_compile_one, nvcc, s, o, OBJ_DIR / (s.stem + ".log")) for s, o in zip(srcs, objs)]
        for f in futs:
            src, rc, out = f.result()
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
            if verbose:
                sys.stderr.write(f"== {src.name}\n{out}\n")
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
