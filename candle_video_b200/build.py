"""In-tree build of the sm_100a shared library and the standalone bring-up tools.

    python -m candle_video_b200.build            # library + tools (skips up-to-date targets)
    python -m candle_video_b200.build --force

nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU build box; the resulting
`candle_video_b200/lib/libltxv_b200.so` travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libltxv_b200.so"

NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I", str(ROOT / "include"),
          "-I", str(CSRC)]

LIB_SOURCES = [
    "gemm_tcgen05.cu",
    "attention_tcgen05.cu",
    "glue.cu",
    "vae_glue.cu",
    "dit.cu",
    "vae.cu",
    "vae_encoder.cu",
    "conv3d_op.cu",
    "pipeline.cu",
    "comm.cu",
    "weights.cc",
    "ffi.cu",
    "model_common.cu",
    "tensormap.cc",
    "profile.cc",
    "options.cc",
]

TOOLS = {
    "gemm_test": (["tools/gemm_test.cu"], ["gemm_tcgen05.cu", "tensormap.cc", "profile.cc", "options.cc"]),
    "attn_test": (["tools/attn_test.cu"], ["attention_tcgen05.cu", "tensormap.cc", "profile.cc", "options.cc"]),
    "attn_prof": (["tools/attn_prof.cu"], ["attention_tcgen05.cu", "tensormap.cc", "profile.cc", "options.cc"]),
    "attn_bench": (["tools/attn_bench.cu"], ["attention_tcgen05.cu", "tensormap.cc", "profile.cc", "options.cc"]),
}


def _newer(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def _headers() -> list[Path]:
    return list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {cmd[-1]}")
    if r.stderr.strip() and os.environ.get("LTXV_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def _compile(src: Path, obj: Path, force: bool) -> Path:
    if force or _newer(obj, [src] + _headers()):
        obj.parent.mkdir(parents=True, exist_ok=True)
        _run([NVCC, *ARCH, *COMMON, "-c", str(src), "-o", str(obj)])
    return obj


def build_library(force: bool = False) -> Path:
    srcs = [CSRC / s for s in LIB_SOURCES if (CSRC / s).exists()]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, OBJDIR / (s.name + ".o"), force), srcs))
    if force or _newer(LIB, objs):
        LIBDIR.mkdir(parents=True, exist_ok=True)
        _run([NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"])
    return LIB


def build_tools(force: bool = False) -> list[Path]:
    outs = []
    for name, (mains, deps) in TOOLS.items():
        if not all((ROOT / m).exists() for m in mains) or not all((CSRC / d).exists() for d in deps):
            continue
        out = ROOT / "tools" / "bin" / name
        srcs = [ROOT / m for m in mains] + [CSRC / d for d in deps]
        if force or _newer(out, srcs + _headers()):
            out.parent.mkdir(parents=True, exist_ok=True)
            _run([NVCC, *ARCH, *COMMON, "-o", str(out), *map(str, srcs), "-cudart", "static"])
        outs.append(out)
    return outs


def main() -> None:
    force = "--force" in sys.argv
    lib = build_library(force)
    tools = build_tools(force)
    print(f"built {lib}")
    for t in tools:
        print(f"built {t}")


if __name__ == "__main__":
    main()
