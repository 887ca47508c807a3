"""In-tree build of libvinum_b200.so (the C-ABI library, include/vinum_b200.h).

nvcc cross-compiles for sm_100a without a GPU.  The shared object is written next
to this file (vinum_b200/_C/libvinum_b200.so): it is git-ignored but travels to the
GPU box with the repo snapshot.

    python vinum_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT_DIR = HERE / "_C"
LIB_PATH = OUT_DIR / "libvinum_b200.so"

# (source, object stem, extra defines): vk_agg_fast_inst.cu is compiled once per slice of
# the fused-aggregate kernel variants so that they build in parallel
SOURCES = [
    ("vk_agg_fast_inst.cu", f"vk_agg_fast_inst{part}", [f"-DVK_FAST_PART={part}"]) for part in range(15)
] + [
    ("vk_hashagg.cu", "vk_hashagg", []),
    ("vk_sort.cu", "vk_sort", []),
    ("vk_filter.cu", "vk_filter", []),
    ("vk_runtime.cu", "vk_runtime", []),
    ("vk_arith.cu", "vk_arith", []),
    ("vk_ingest.cu", "vk_ingest", []),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # NumPy parity: every float result is one correctly rounded IEEE op (no FMA contraction)
    "-fmad=false",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: vinum_b200 needs the CUDA toolkit to build its sm_100a kernels")


def _newest_source_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "vinum_b200.h",
                                                                   Path(__file__)]
    return max(f.stat().st_mtime for f in files)


def needs_build() -> bool:
    return not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < _newest_source_mtime()


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> Path:
    """`variant` / `defines` build a tuning variant (libvinum_b200_<variant>.so, selected at
    run time with VINUM_B200_LIB) next to the product library."""
    lib_path = OUT_DIR / f"libvinum_b200_{variant}.so" if variant else LIB_PATH
    if not variant and not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR / ("obj_" + variant if variant else "obj")
    obj_dir.mkdir(exist_ok=True)

    hdr_mtime = max(f.stat().st_mtime for f in list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "vinum_b200.h",
                                                                            Path(__file__)])

    extra_defines = list(defines)

    def compile_one(item) -> Path:
        src, stem, defines = item
        obj = obj_dir / (stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime >= max(hdr_mtime, (CSRC / src).stat().st_mtime):
            return obj  # up to date
        cmd = [nvcc, *NVCC_FLAGS, *defines, *extra_defines, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, flush=True)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = lib_path.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    variant, defs = "", []
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            variant = a.split("=", 1)[1]
        elif a.startswith("-D"):
            defs.append(a)
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant, defines=defs)
    print(p)
