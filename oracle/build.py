"""Builds the CPU oracle (TEST INFRASTRUCTURE): liboracle.so from the C restatement, and — only where
/root/reference exists (this container, never the GPU box) — oracle/_ref/ from the reference's own sources."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = ["gso_oracle.c", "hh_oracle.c", "enum_oracle.c"]


def build_oracle(force=False):
    out = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, s) for s in SRCS if os.path.exists(os.path.join(HERE, s))]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    # -ffp-contract=off: the reference build has no FMA contraction (configure.ac:25, no -march)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", out] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return out


def build_ref(jobs=8):
    if not os.path.isdir("/root/reference/fplll"):
        return None  # GPU box: use the prebuilt oracle/_ref that travelled with the snapshot
    need = ["libfplll.so", "fplll", "latticegen", "ref_probe", "strategies/default.json"]
    probe_src = os.path.join(HERE, "ref_probe.cpp")
    fresh = all(os.path.exists(os.path.join(HERE, "_ref", f)) for f in need) and \
        os.path.getmtime(os.path.join(HERE, "_ref", "ref_probe")) >= os.path.getmtime(probe_src)
    if not fresh:
        subprocess.check_call(["make", "-f", os.path.join(HERE, "Makefile.ref"), "-j%d" % jobs, "all"],
                              cwd=HERE, stdout=subprocess.DEVNULL)
    # the reference's own test programs (run under the MatGSO shim by tests/test_shim_gpu.py)
    if not all(os.path.exists(os.path.join(HERE, "_ref", t)) for t in ("test_lll", "test_gso", "test_bkz")):
        subprocess.check_call(["make", "-f", os.path.join(HERE, "Makefile.ref"), "-j%d" % jobs,
                               os.path.join(HERE, "_ref", "test_lll"), os.path.join(HERE, "_ref", "test_gso"),
                               os.path.join(HERE, "_ref", "test_bkz")], cwd=HERE, stdout=subprocess.DEVNULL)
    return os.path.join(HERE, "_ref")


if __name__ == "__main__":
    build_oracle(force=True)
    if "--ref" in sys.argv:
        build_ref()
