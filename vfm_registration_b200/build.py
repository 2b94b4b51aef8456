"""Builds vfm_registration_b200/libvfmreg_b200.so in-tree with nvcc for sm_100a (no other target, no JIT cache).

    python -m vfm_registration_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# VFM_BUILD_SUFFIX=epi8 builds a second library (libvfmreg_b200_epi8.so, objects in _obj_epi8) next to the product one, for
# A/B runs of a compile-time knob: VFMREG_LIB selects it at load time
_SUFFIX = os.environ.get("VFM_BUILD_SUFFIX", "")
OBJ = os.path.join(HERE, "_obj" + ("_" + _SUFFIX if _SUFFIX else ""))
LIB = os.path.join(HERE, "libvfmreg_b200" + ("_" + _SUFFIX if _SUFFIX else "") + ".so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
# per-file extra flags; ransac.cu spells every fused multiply-add explicitly (bit parity with the C oracle)
EXTRA = {"ransac.cu": ["-fmad=false"], "project.cu": ["-fmad=false"], "voxel.cu": ["-fmad=false"], "nnscore.cu": ["-fmad=false"], "teaser.cu": ["-fmad=false"]}
# tuning knobs: VFM_NVCC_DEFS="-DVFM_TBN=128 -DVFM_EPI_WARPS=4" python -m vfm_registration_b200.build --force
EXTRA_ALL = os.environ.get("VFM_NVCC_DEFS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvfmreg_b200.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_all(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "vfmreg_b200.h"))
    headers.append(os.path.abspath(__file__))
    env = dict(os.environ)
    env.pop("CC", None)   # the image's CC wrapper lacks libgomp specs; nvcc must use /usr/bin/g++
    env.pop("CXX", None)

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        if not force and not _stale(obj, [os.path.join(CSRC, src)] + headers):
            return obj, ""
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj, "-ccbin", "/usr/bin/g++"] + ARCH + COMMON + EXTRA.get(src, []) + EXTRA_ALL
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if verbose and log:
        print(log)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, "-ccbin", "/usr/bin/g++"] + ARCH + objs + ["-lcuda", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
