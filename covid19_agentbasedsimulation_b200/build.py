"""Builds libcovidabs.so in-tree with nvcc for sm_100a (B200).

    python -m covid19_agentbasedsimulation_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcovidabs.so")
SOURCES = [os.path.join(CSRC, "abs_capi.cu"), os.path.join(CSRC, "graph_capi.cu"), os.path.join(CSRC, "world_builder.cu")]
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "covidabs.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # exact reproducibility of the draw transforms and of the wealth arithmetic: no FMA contraction
              "-fmad=false",
              "-shared", "-Xcompiler", "-fPIC"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n{}\n{}".format(" ".join(cmd), res.stdout))
    if verbose:
        print(res.stdout)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
