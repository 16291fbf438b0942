"""TEST INFRASTRUCTURE: compiles csrc/*.cu as host C++ against tests/emu/cpu_emu.h so the
CPU suite can execute the SIMT kernels' source (see cpu_emu.h).  Kernels that use inline PTX
(tcgen05/TMA) are excluded -- they only run on the GPU box."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "equivarianttransformermpnn4quantumcomputations_b200", "csrc")
LIB = os.path.join(HERE, "libeqv2_emu.so")
EXCLUDE = {"gemm_tc.cu"}


def build(force=False):
    srcs = [s for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if os.path.basename(s) not in EXCLUDE]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cpu_emu.h"), os.path.join(HERE, "cpu_emu.cc")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) < os.path.getmtime(LIB) for d in deps):
        return LIB
    objs = []
    procs = []
    for s in srcs + [os.path.join(HERE, "cpu_emu.cc")]:
        o = os.path.join(HERE, os.path.basename(s) + ".emu.o")
        objs.append(o)
        cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-DEQV2_CPU_EMU", "-I", HERE, "-x", "c++", "-c", s, "-o", o,
               "-Wno-unknown-pragmas", "-pthread"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emu compile failed on {s}:\n{out}")
    r = subprocess.run(["g++", "-shared", "-o", LIB, *objs, "-pthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
