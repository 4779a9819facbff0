"""Build libstcgan_b200.so (sm_100a) in-tree with nvcc.  No GPU is needed to build.

    python shadow-removal-istd_b200/build.py [--force] [--verbose]

Output: shadow-removal-istd_b200/stcgan_b200/libstcgan_b200.so  (git-ignored, travels with gpurun snapshots)
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "stcgan_b200", "libstcgan_b200.so")
STAMP = OUT + ".stamp"
SOURCES = ["api.cu", "tapconv_ffma.cu", "tapconv_tc.cu", "bn_act.cu", "misc.cu", "thin_col2im.cu", "augment.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_ptx.cuh"), os.path.join(HERE, "..", "include", "stcgan_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas=-v"]


def _digest():
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose:
            sys.stderr.write(f"--- {s} ---\n{out}\n")
        else:
            spills = [l for l in out.splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
            if spills:
                sys.stderr.write(f"--- {s}: register spills ---\n" + "\n".join(spills) + "\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
