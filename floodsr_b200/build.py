"""Build `libfloodsr_b200.so` in-tree with nvcc for sm_100a (`python -m floodsr_b200.build`)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libfloodsr_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: floodsr_b200 needs the CUDA toolkit to build (no CPU fallback)")
    return cand


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "floodsr_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def _compile_one(args):
    cmd, src = args
    proc = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    return src, proc.returncode, proc.stdout + proc.stderr, cmd


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every source to an object (in parallel, only the stale ones) and link the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = find_nvcc()
    obj_dir = PKG.parent / "build" / "obj"
    obj_dir.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "floodsr_b200.h"]
    h_time = max(h.stat().st_mtime for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, h_time):
            cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "-shared"], "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append((cmd, src))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for src, rc, out, cmd in pool.map(_compile_one, jobs):
            if verbose or rc != 0:
                sys.stderr.write(out)
            if rc != 0:
                raise RuntimeError(f"nvcc failed ({rc}): {' '.join(cmd)}")
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", *[str(o) for o in objs], "-lcuda", "-o", str(LIB)]
    proc = subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc link failed ({proc.returncode}): {' '.join(cmd)}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
