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


def build_shim():
    """tests/_build/libb200fplll.so (the MatGSO forwarding shim, fplll_b200/csrc/fplll_matgso_shim.cpp) and
    tests/_build/shim_demo: both need the reference headers, so they are built here and travel to the GPU box."""
    if not os.path.isdir("/root/reference/fplll") or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfplll.so")):
        return None
    bdir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(bdir, exist_ok=True)
    sh = os.path.join(ROOT, "oracle", "shim")
    inc = ["-I%s/include" % sh, "-I%s/cfg/fplll" % sh, "-I%s/cfg/fplll/enum" % sh, "-I%s/cfg" % sh, "-I/root/reference",
           "-I/root/reference/fplll"]
    lib = os.path.join(bdir, "libb200fplll.so")
    src = os.path.join(ROOT, "fplll_b200", "csrc", "fplll_matgso_shim.cpp")
    deps = [src, os.path.join(ROOT, "include", "b200gso.h"), os.path.join(ROOT, "include", "b200bkz.h")]
    if not (os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(p) for p in deps)):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-fPIC", "-shared", "-pthread"] + inc + [src, "-o", lib,
                              "-L" + os.path.join(ROOT, "fplll_b200", "lib"), "-lb200bkz", "-ldl",
                              "-Wl,-rpath,$ORIGIN/../../fplll_b200/lib"])
    exe = os.path.join(bdir, "shim_demo")
    dsrc = os.path.join(ROOT, "tests", "shim_demo.cpp")
    if not (os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(dsrc)):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-pthread"] + inc + [dsrc, "-o", exe,
                              "-L" + os.path.join(ROOT, "oracle", "_ref"), "-lfplll", "-l:libmpfr.so.6", "-l:libgmp.so.10",
                              "-Wl,-rpath,$ORIGIN/../../oracle/_ref"])
    return lib, exe


if __name__ == "__main__":
    print(build())
    print(build_shim())
