"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo).

  libldpc_b200.so   -- sm_100a CUDA kernels + the extern "C" boundary (include/ldpc_b200.h)
  libmyldpc_b200.so -- the drop-in C++ `Coder` class (include/MyLdpc.h) over that boundary
"""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess

PKG = pathlib.Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libldpc_b200.so"
CODER_LIB = PKG / "libmyldpc_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # never contract a*b+c: the posterior must round after every add
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math",
]

OBJ = ROOT / "build" / "obj"   # git-ignored; objects are an implementation detail of this script

# One translation unit per kernel family so nvcc builds them in parallel (the compiled quasi-cyclic profiles: one per
# 802.16e rate).  (source, object name, extra defines)
QC_RATES = [("34B", "QcProfile34B"), ("34A", "QcProfile34A"), ("23B", "QcProfile23B"),
            ("23A", "QcProfile23A"), ("12", "QcProfile12"), ("56", "QcProfile56")]
UNITS = [("ldpc_b200.cu", "ldpc_b200", []), ("ldpc_tables.cpp", "ldpc_tables", []),
         ("k_group.cu", "k_group", []), ("k_qcg.cu", "k_qcg", []), ("k_sp.cu", "k_sp", []),
         ("k_tdmp.cu", "k_tdmp", []), ("k_misc.cu", "k_misc", []), ("k_qcw.cu", "k_qcw", []), ("k_qcm.cu", "k_qcm", []), ("k_spq.cu", "k_spq", [])]
UNITS += [("k_qc.cu", "k_qc_" + tag, ["-DLDPC_QC_RATE=" + prof, "-DLDPC_QC_RATE_FN=qc_profiles_" + tag]) for tag, prof in QC_RATES]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: pathlib.Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(pathlib.Path(s).stat().st_mtime > t for s in sources)


def _headers():
    return sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "ldpc_b200.h"]


def build_cuda(force: bool = False, verbose: bool = False, ptxas_info: bool = False, only=None) -> pathlib.Path:
    """nvcc -c every unit that is older than its source or any header (in parallel), then link."""
    from concurrent.futures import ThreadPoolExecutor

    hdrs = _headers()
    if not (force or only or ptxas_info) and not _stale(LIB, [CSRC / src for src, _, _ in UNITS] + hdrs):
        return LIB  # the shipped library is current (the objects need not exist: build/ does not travel to the GPU box)
    OBJ.mkdir(parents=True, exist_ok=True)
    jobs = []
    for src, name, defs in UNITS:
        obj = OBJ / (name + ".o")
        if only and name not in only and obj.exists():
            continue
        if force or _stale(obj, [CSRC / src] + hdrs):
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), *defs, "-c", str(CSRC / src), "-o", str(obj)]
            if ptxas_info:
                cmd += ["-Xptxas", "-v"]
            jobs.append((name, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, cmd, r

    if jobs:
        workers = max(1, min(len(jobs), (os.cpu_count() or 4)))
        with ThreadPoolExecutor(workers) as ex:
            for name, cmd, r in ex.map(run, jobs):
                if verbose or r.returncode or ptxas_info:
                    print(" ".join(cmd))
                    print(r.stdout, r.stderr)
                if r.returncode:
                    raise RuntimeError("nvcc failed for " + name)
    objs = [OBJ / (name + ".o") for _, name, _ in UNITS]
    if jobs or force or _stale(LIB, objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd))
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("link failed for libldpc_b200.so")
    return LIB


def build_coder(force: bool = False, verbose: bool = False) -> pathlib.Path:
    src = CSRC / "mycoder.cpp"
    if not src.exists():
        return CODER_LIB
    deps = [src, ROOT / "include" / "MyLdpc.h", ROOT / "include" / "MyLdpc_c.h", ROOT / "include" / "ldpc_b200.h"]
    if force or _stale(CODER_LIB, deps) or _stale(CODER_LIB, [LIB]):
        cxx = shutil.which("g++") or "g++"
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-I", str(ROOT / "include"), str(src),
               "-o", str(CODER_LIB), "-L", str(PKG), "-lldpc_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd))
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("g++ failed for libmyldpc_b200.so")
    return CODER_LIB


BIN = PKG / "bin"
REFERENCE_TEST = pathlib.Path("/root/reference/Test.cpp")


def build_harness(force: bool = False, verbose: bool = False) -> None:
    """tools/mytest.cpp (our Test.cpp-shaped CLI) and, when the reference tree is mounted, the
    reference's OWN Test.cpp compiled unmodified against include/MyLdpc.h -- the drop-in proof.
    Binaries land in myldpccppapi_b200/bin/ (git-ignored, shipped to the GPU box)."""
    if not CODER_LIB.exists():
        return
    BIN.mkdir(exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    jobs = [(ROOT / "tools" / "mytest.cpp", BIN / "mytest", [])]
    if REFERENCE_TEST.exists():
        # `#include "MyLdpc.h"` / "cl.hpp" resolve next to the including file first, so the reference's
        # Test.cpp is compiled through a symlink in a scratch directory (nothing is copied): its quoted
        # includes then fall through to -I include, i.e. to OUR MyLdpc.h and the empty cl.hpp stub.
        import tempfile
        scratch = pathlib.Path(tempfile.mkdtemp(prefix="myldpc_dropin_"))
        link = scratch / "Test.cpp"
        link.symlink_to(REFERENCE_TEST)
        jobs.append((link, BIN / "MyTest_reference_harness", []))
    for src, out, extra in jobs:
        if not (force or _stale(out, [src.resolve(), ROOT / "include" / "MyLdpc.h", CODER_LIB])):
            continue
        cmd = [cxx, "-std=c++17", "-O2", "-w", "-I", str(ROOT / "include"), str(src), "-o", str(out),
               "-L", str(PKG), "-lmyldpc_b200", "-lldpc_b200", "-Wl,-rpath,$ORIGIN/..", *extra]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd))
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("g++ failed for " + out.name)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force=force, verbose=verbose)
    build_coder(force=force, verbose=verbose)
    build_harness(force=force, verbose=verbose)


if __name__ == "__main__":
    import sys
    build_cuda(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    build_coder(force="--force" in sys.argv, verbose=True)
    build_harness(force="--force" in sys.argv, verbose=True)
