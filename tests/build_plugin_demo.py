"""Builds tests/_build/plugin_demo (reference library + fplll_extenum_adapter + libb200enum) where the reference
headers exist (development container only).  The binary travels to the GPU box with the snapshot."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build():
    if not os.path.isdir("/root/reference/fplll") or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfplll.so")):
        return None
    out = os.path.join(ROOT, "tests", "_build", "plugin_demo")
    srcs = [os.path.join(ROOT, "tests", "plugin_demo.cpp"),
            os.path.join(ROOT, "fplll_b200", "csrc", "fplll_extenum_adapter.cpp")]
    lib = os.path.join(ROOT, "fplll_b200", "lib", "libb200enum.so")
    deps = srcs + [lib, os.path.join(ROOT, "include", "b200enum.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(p) for p in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    sh = os.path.join(ROOT, "oracle", "shim")
    cmd = ["g++", "-O2", "-std=c++11", "-pthread", "-I%s/include" % sh, "-I%s/cfg/fplll" % sh, "-I%s/cfg/fplll/enum" % sh,
           "-I%s/cfg" % sh, "-I/root/reference", "-I/root/reference/fplll"] + srcs + [
           "-o", out, "-L" + os.path.join(ROOT, "fplll_b200", "lib"), "-lb200enum",
           "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lfplll", "-l:libmpfr.so.6", "-l:libgmp.so.10",
           "-Wl,-rpath,$ORIGIN/../../fplll_b200/lib:$ORIGIN/../../oracle/_ref"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build())
