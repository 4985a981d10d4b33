"""Build camkifu_b200/libcamkifu_b200.so (hand-written sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

The shared library travels to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcamkifu_b200.so")
SOURCES = ["ckb_api.cu", "geometry.cu", "warp.cu", "mog2.cu", "kmeans.cu", "kmeans_cluster.cu", "zones.cu", "jpeg_ingest.cu", "cnn_pack.cu", "cnn_simt.cu", "cnn_tc.cu", "cnn_tc_front.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "camkifu_b200.h"))
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None, bdir: str = None) -> str:
    """extra_flags / out / bdir: instrumented variants for tools/ (e.g. -DKC_TIMING); the product is the default call."""
    if out is None and not force and up_to_date():
        return OUT
    out = out or OUT
    objs = []
    bdir = bdir or os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on " + src)
    # the arch on the link line too: nvcc otherwise adds a device-link stub for its default architecture (sm_52)
    subprocess.run([nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart", "-ldl"], check=True)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
