"""Builds libporeseq_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m poreseq_b200.build [--force]

-fmad=false keeps the FP64 recurrence bit-identical to the reference's x86-64 build (which emits no
FMA); FP32 kernels that want FMA ask for it explicitly with intrinsics.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libporeseq_b200.so")
SOURCES = ["ps_host.cu", "ps_drivers.cu", "ps_viterbi.cu", "ps_sw.cu", "ps_pack.cu", "ps_comm.cu", "ps_lockstep.cu", "ps_swhost.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-O3",
              "-Xptxas", "-v", "-ldl"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "poreseq_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: experiment builds (kernel A/Bs, ablations) beside the shipped library; loaded through
    PORESEQ_B200_LIB (poreseqcpp.lib)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = LIB if out is None else os.path.join(HERE, out)
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log" if out is None else out + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building %s" % LIB)
    if verbose:
        sys.stderr.write(log)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, out=outs[0] if outs else None))
