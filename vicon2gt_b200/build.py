"""In-tree native builds (no JIT cache): the CUDA solver library for sm_100a and the host
simulator library.  The oracle (test infrastructure) has its own recipe in oracle/Makefile.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
HOST = os.path.join(PKG_DIR, "host")

_LIBS = {
    "cuda": os.path.join(PKG_DIR, "libvicon2gt_b200.so"),
    "sim": os.path.join(HOST, "libv2gt_sim.so"),
    "oracle": os.path.join(ROOT, "oracle", "liboracle.so"),
}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def lib_path(name):
    # V2GT_LIB_CUDA points the loader at an alternative build of the CUDA library (kernel tuning experiments)
    if name == "cuda" and os.environ.get("V2GT_LIB_CUDA"):
        return os.environ["V2GT_LIB_CUDA"]
    return _LIBS[name]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def _find(exe, fallbacks):
    p = shutil.which(exe)
    if p:
        return p
    for f in fallbacks:
        if os.path.exists(f):
            return f
    raise RuntimeError(f"{exe} not found")


def cuda_sources():
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdrs.append(os.path.join(ROOT, "include", "vicon2gt_b200.h"))
    return srcs, hdrs


def build_cuda(force=False, verbose=False, extra=(), out=None, objdir_name="_obj"):
    """nvcc -> vicon2gt_b200/libvicon2gt_b200.so (sm_100a only; cross-compiles without a GPU)."""
    srcs, hdrs = cuda_sources()
    out = out or _LIBS["cuda"]
    if not force and not _newer(out, srcs + hdrs):
        return out
    nvcc = _find("nvcc", ["/usr/local/cuda/bin/nvcc"])
    nccl_inc = []
    for cand in ("/usr/include/nccl.h",):
        if os.path.exists(cand):
            nccl_inc = ["-DV2GT_HAVE_NCCL=1"]
    # one nvcc per translation unit, in parallel, then one link step (objects live in _obj/, git-ignored)
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(PKG_DIR, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    common = [nvcc] + NVCC_FLAGS + list(extra) + nccl_inc + ["-I", os.path.join(ROOT, "include")]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            log = _run(common + ["-c", src, "-o", obj])
            if verbose:
                print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return out


def build_sim(force=False):
    src = os.path.join(HOST, "sim.cpp")
    deps = [src, os.path.join(HOST, "v2gt_sim.h"), os.path.join(CSRC, "v2gt_math.cuh"), os.path.join(CSRC, "v2gt_spline.cuh")]
    out = _LIBS["sim"]
    if not force and not _newer(out, deps):
        return out
    gxx = _find("g++", ["/usr/bin/g++"])
    _run([gxx, "-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-shared", "-o", out, src])
    return out


def build_oracle(force=False):
    """Test infrastructure: compiles oracle/'s C++ restatement (building the checker is not using it)."""
    odir = os.path.join(ROOT, "oracle")
    if force and os.path.exists(_LIBS["oracle"]):
        os.remove(_LIBS["oracle"])
    _run(["make", "-s", "-C", odir])
    return _LIBS["oracle"]


def build_ref(force=False):
    """Test infrastructure: oracle/_ref/libv2gt_ref.so from the reference's own sources where they lie under
    /root/reference (oracle/Makefile.ref, against the interface stand-ins of oracle/shim/).  Returns None where the
    reference is not present (the GPU box): the prebuilt library and tests/golden/ref_pin.npz travel instead."""
    odir = os.path.join(ROOT, "oracle")
    ref = os.environ.get("V2GT_REFERENCE_ROOT", "/root/reference")
    out = os.path.join(odir, "_ref", "libv2gt_ref.so")
    if not os.path.isdir(os.path.join(ref, "src", "cpi")):
        return out if os.path.exists(out) else None
    if force and os.path.exists(out):
        os.remove(out)
    _run(["make", "-s", "-C", odir, "-f", "Makefile.ref", "REF=" + ref])
    return out


def build_examples(force=False):
    """g++ -> examples/run_simulation, run_monte_carlo and run_chain (host C++ over the C ABI: include/vicon2gt_b200.hpp); needs
    the two libraries.  Returns the path of run_simulation."""
    gxx = _find("g++", ["/usr/bin/g++"])
    outs = []
    for name in ("run_simulation", "run_monte_carlo", "run_chain"):
        src = os.path.join(ROOT, "examples", name + ".cpp")
        out = os.path.join(ROOT, "examples", name)
        deps = [src, os.path.join(ROOT, "include", "vicon2gt_b200.hpp"), os.path.join(ROOT, "include", "vicon2gt_b200.h"), _LIBS["cuda"], _LIBS["sim"]]
        outs.append(out)
        if not force and not _newer(out, deps):
            continue
        _run([gxx, "-O2", "-std=c++14", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", out, src,
              "-L", PKG_DIR, "-L", HOST, "-lvicon2gt_b200", "-lv2gt_sim",
              "-Wl,-rpath," + PKG_DIR, "-Wl,-rpath," + HOST, "-Wl,-rpath,$ORIGIN/../vicon2gt_b200", "-Wl,-rpath,$ORIGIN/../vicon2gt_b200/host",
              "-Wl,-rpath,/usr/local/cuda/lib64"])
    return outs[0]


def build_all(force=False, verbose=False):
    build_sim(force)
    build_cuda(force, verbose)
    build_oracle(force)
    build_examples(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", ", ".join(_LIBS.values()))
