"""Build libcbinfer_sm100.so in-tree with nvcc (sm_100a only).

Replaces the reference's pycbinfer/build.sh:1-10 (three libraries for compute_52/61) with one
library for Blackwell.  Run as ``python -m cbinfer_b200.build`` or call :func:`build`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcbinfer_sm100.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                  if f.endswith((".cu", ".cuh")) and f != "compat_shim.cu") + \
        [os.path.join(os.path.dirname(HERE), "include", "cbinfer_b200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/cb_abi.cu -> cbinfer_b200/libcbinfer_sm100.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcbinfer_sm100.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB + ".tmp", os.path.join(CSRC, "cb_abi.cu")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout[-4000:])
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(proc.stdout)
    return LIB


COMPAT_DIR = os.path.join(HERE, "compat")


def compat_libs():
    """the three libraries the UNMODIFIED reference dlopens (pycbinfer/conv2d_cg.py:44,50,
    conv2d_fg.py:31), built from csrc/compat_shim.cu on top of libcbinfer_sm100.so"""
    import platform
    m = platform.machine()
    return {"cbconv2d_cg_backend": (os.path.join(COMPAT_DIR, "cbconv2d_cg_backend_%s.so" % m), []),
            "cbconv2d_cg_half_backend": (os.path.join(COMPAT_DIR, "cbconv2d_cg_half_backend_%s.so" % m),
                                         ["-DCB_COMPAT_HALF"]),
            "cbconv2d_fg_backend": (os.path.join(COMPAT_DIR, "cbconv2d_fg_backend_%s.so" % m), ["-DCB_COMPAT_FG"])}


def build_compat(force=False):
    """Compile the reference-symbol shim libraries into cbinfer_b200/compat/.  Returns {name: path}."""
    build()
    src = os.path.join(CSRC, "compat_shim.cu")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(COMPAT_DIR, exist_ok=True)
    out = {}
    for name, (path, defs) in compat_libs().items():
        out[name] = path
        deps = [src, os.path.join(os.path.dirname(HERE), "include", "cbinfer_b200.h")]
        if not force and os.path.exists(path) and all(os.path.getmtime(path) >= os.path.getmtime(d) for d in deps):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "--shared",
               "-Xcompiler", "-fPIC"] + defs + ["-o", path + ".tmp", src]
        if "-DCB_COMPAT_FG" not in defs:
            cmd += ["-L" + HERE, "-lcbinfer_sm100", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/..",
                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (compat shim):\n" + proc.stdout[-4000:])
        os.replace(path + ".tmp", path)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--compat" in sys.argv:
        print(build_compat(force="--force" in sys.argv))
