"""Builds libcafe_gpu.so (CUDA kernels + C ABI + host problem builders) in-tree for sm_100a.

Every source is compiled to an object (relocatable device code) that is cached by modification time, then nvcc links
the shared library; the generated whole-body routines live in their own translation unit because ptxas needs minutes for them."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libcafe_gpu.so")
HOST_LIB = os.path.join(HERE, "libcafe_host.so")   # the problem builders / settings readers alone, g++ only: what the CPU baseline arm loads

# source -> dependencies besides itself (paths relative to csrc/)
ABI = ["../../include/cafe_gpu.h", "../../include/cafe_deck.h"]
SOURCES = {
    "solver.cu": ["bwd2.cuh", "device_types.cuh", "launchers.h"] + ABI,
    "knot_kernels.cu": ["knot_kernels.cuh", "launchers.h", "device_types.cuh", "model_hkd.cuh", "model_srb.cuh", "model_wb.cuh", "gen/hkd_gen.h", "gen/srb_gen.h"] + ABI,
    "wb_gen_wrappers.cu": ["gen/wb_gen.h", "wb_pieces.h"],
    "wb_leg_kernels.cu": ["gen/wb_leg_gen.h", "gen/wb_gen.h", "wb_leg_tables.h", "device_types.cuh", "launchers.h"] + ABI,
    "wb_coop.cu": ["wb_coop.cuh", "wb_leg_tables.h", "device_types.cuh", "model_hkd.cuh", "gen/hkd_gen.h", "launchers.h"] + ABI,
    "host/abi_host.cpp": ["host/problem_builders.h", "host/quad_reference.h"] + ABI,
    "host/hkd_problem.cpp": ["host/problem_builders.h", "host/quad_reference.h", "host/info_reader.h", "gen/hkd_gen.h"] + ABI,
    "host/mhpc_problem.cpp": ["host/problem_builders.h", "host/quad_reference.h", "host/info_reader.h"] + ABI,
    "host/quad_reference.cpp": ["host/quad_reference.h"],
}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-rdc=true", "-Xcompiler", "-fPIC",
    "--fmad=true", "-Xptxas", "-v", "-Xcudafe", "--diag_suppress=177",
]


# per-source extra flags: the generated whole-body routines are capped at 128 registers so that the per-(problem,knot)
# kernels calling them reach 4 CTAs of 128 threads per SM (their local-memory traffic needs the latency hiding)
_RC = os.environ.get("CAFE_KNOT_MAXRREG", "128")   # dev switch for occupancy experiments
_MINB = str(max(1, 65536 // (128 * int(_RC))))
_LRC = os.environ.get("CAFE_LEG_MAXRREG", "255")    # register cap of the leg-parallel straight-line kernels
EXTRA = {"solver.cu": os.environ.get("CAFE_SOLVER_DEFS", "").split(), "wb_gen_wrappers.cu": ["-maxrregcount=" + _RC], "wb_leg_kernels.cu": ["-maxrregcount=" + _LRC, "-DCAFE_LEG_MINB=" + str(max(1, 65536 // (128 * int(_LRC))))], "knot_kernels.cu": ["-maxrregcount=" + _RC, "-DCAFE_KNOT_MINB=" + _MINB] + os.environ.get("CAFE_KNOT_DEFS", "").split()}


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _stale(target, deps):
    t = _mtime(target)
    return t == 0.0 or any(_mtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    objs, logs, rebuilt = [], [], False
    jobs = []
    for src, deps in SOURCES.items():
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(obj)
        all_deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, d) for d in deps]
        if force or _stale(obj, all_deps):
            jobs.append((src, [nvcc] + NVCC_FLAGS + EXTRA.get(src, []) + ["-c", "-o", obj, os.path.join(CSRC, src)]))
    # the translation units are independent: compile them side by side (ptxas needs minutes for the generated routines)
    procs = [(src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for src, cmd in jobs]
    failed = None
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        logs.append(" ".join(cmd) + "\n" + out)
        if pr.returncode != 0 and failed is None:
            failed = src
            sys.stderr.write(logs[-1])
        rebuilt = True
    if failed is not None:
        raise RuntimeError("nvcc failed on " + failed)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError("nvcc link failed")
    # host-only library: the same four host sources compiled by g++ (no CUDA runtime in the process that loads it)
    host_src = [os.path.join(CSRC, s) for s in SOURCES if s.startswith("host/")]
    host_deps = host_src + [os.path.join(CSRC, d) for s in SOURCES if s.startswith("host/") for d in SOURCES[s]]
    if force or _stale(HOST_LIB, host_deps):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", HOST_LIB] + host_src + ["-lstdc++", "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError("host library build failed")
    if logs:
        with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
            f.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
