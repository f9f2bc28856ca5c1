"""In-tree nvcc build of the CUDA extension (sm_100a only).  The .so stays in fplll_b200/lib/ (git-ignored,
travels to the GPU box with the snapshot)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # the reference build never contracts a*b+c (configure.ac:25: -O3, no -march): bit parity needs the same
              "--fmad=false", "-Xcompiler", "-fPIC", "-shared"]

TARGETS = {
    "libb200gso.so": ["gso_api.cu"],
    "libb200enum.so": ["enum_api.cu"],
    "libb200bkz.so": ["gso_api.cu", "enum_api.cu", "bkz_api.cu"],
    "libb200hh.so": ["hh_api.cu"],
}


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f)
                                                                  for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(p) > os.path.getmtime(out) for p in deps)


def build_all(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    built = []
    for name, srcs in TARGETS.items():
        paths = [os.path.join(CSRC, s) for s in srcs]
        if not all(os.path.exists(p) for p in paths):
            continue
        out = os.path.join(LIBDIR, name)
        if force or _stale(out, paths):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + paths
            if name == "libb200enum.so":
                cmd += ["-lnccl"] if os.environ.get("B200_WITH_NCCL") else []
            subprocess.check_call(cmd)
        built.append(out)
    return built


if __name__ == "__main__":
    import sys
    print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))
