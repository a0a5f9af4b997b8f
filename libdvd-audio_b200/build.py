"""Build recipe for the engine: nvcc for the sm_100a kernels + C ABI
(lib/libdvdagpu.so), gcc for the C host library (lib/libdvd-audio.so) and the
API dumper linked against it (lib/b200_dump).  Everything is built in-tree so
the binaries travel to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
INCLUDE = os.path.join(ROOT, "include")

ENGINE_LIB = os.path.join(LIBDIR, "libdvdagpu.so")
HOST_LIB = os.path.join(LIBDIR, "libdvd-audio.so")
DUMP_BIN = os.path.join(LIBDIR, "b200_dump")
WAV_BIN = os.path.join(LIBDIR, "dvda2wav")

CU_FILES = ["scan.cu", "demux.cu", "mlp_index.cu", "mlp_decode.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", INCLUDE,
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_engine(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("common.cuh", "kernels.cuh", "mlp_common.cuh")] + [os.path.join(INCLUDE, "dvdagpu.h")]
    objs, jobs = [], []
    for cu in CU_FILES:
        src = os.path.join(CSRC, cu)
        obj = os.path.join(OBJDIR, cu.replace(".cu", ".o"))
        if force or _newer(obj, [src] + headers):
            flags = list(NVCC_FLAGS) + os.environ.get("DVDA_NVCC_EXTRA", "").split()
            if ptxas_info:
                flags += ["-Xptxas", "-v"]
            jobs.append([_nvcc()] + flags + ["-c", src, "-o", obj])
        objs.append(obj)
    if jobs:
        # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
    if force or _newer(ENGINE_LIB, objs):
        _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", ENGINE_LIB] + objs, verbose)
    return ENGINE_LIB


def build_host(force=False, verbose=False):
    build_engine(force=force, verbose=verbose)
    src = os.path.join(HOST, "dvd-audio.c")
    deps = [src, os.path.join(INCLUDE, "dvd-audio.h"), os.path.join(INCLUDE, "dvdagpu.h"), ENGINE_LIB]
    if force or _newer(HOST_LIB, deps):
        _run(["gcc", "-O2", "-g", "-Wall", "-Wextra", "-std=c11", "-fPIC", "-shared", "-I", INCLUDE,
              "-o", HOST_LIB, src, "-L", LIBDIR, "-ldvdagpu", "-lpthread", "-Wl,-rpath,$ORIGIN"], verbose)
    dump_src = os.path.join(ROOT, "oracle", "api_dump.c")
    if force or _newer(DUMP_BIN, [dump_src, HOST_LIB]):
        # the same dumper source the reference build uses (oracle/_ref/ref_dump)
        _run(["gcc", "-O2", "-g", "-Wall", "-I", INCLUDE, "-o", DUMP_BIN, dump_src,
              "-L", LIBDIR, "-ldvd-audio", "-Wl,-rpath,$ORIGIN"], verbose)
    wav_src = os.path.join(ROOT, "tools", "dvda2wav.c")
    if force or _newer(WAV_BIN, [wav_src, HOST_LIB]):
        # our equivalent of the reference's extraction tool, on the GPU-backed library
        _run(["gcc", "-O2", "-g", "-Wall", "-I", INCLUDE, "-o", WAV_BIN, wav_src,
              "-L", LIBDIR, "-ldvd-audio", "-Wl,-rpath,$ORIGIN"], verbose)
    return HOST_LIB


def build_all(force=False, verbose=False):
    build_host(force=force, verbose=verbose)
    return ENGINE_LIB, HOST_LIB, DUMP_BIN


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv, verbose=True)
