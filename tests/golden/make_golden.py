"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile.ref).  Run in the build container only:  python tests/golden/make_golden.py
Inputs come from the reference's own latticegen (default seed), never from a re-implementation."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import helpers as H  # noqa: E402

TMP = "/tmp/lat"
os.makedirs(TMP, exist_ok=True)


def save_state(path, rec, extra=None):
    keys = ["d", "n", "n_known_rows", "n_known_cols", "n_source_rows", "flags", "row_expo", "gso_valid_cols",
            "init_row_size", "bf", "gf", "mu", "r", "b"]
    out = {k: rec[k] for k in keys if k in rec}
    out.update(extra or {})
    np.savez_compressed(path, **out)


def gen(args, name):
    p = os.path.join(TMP, name)
    if not os.path.exists(p):
        open(p, "w").write(O.latticegen(args))
    return p


def bkz_fixtures():
    # --- BKZ fixtures: q-ary dim-60 (latticegen q 60 30 12 b), reference wrapper-LLL, then reference BKZ ----------
    q60 = gen(["q", 60, 30, 12, "b"], "q60.txt")
    q60l = os.path.join(TMP, "q60_lll.txt")
    O.run_ref("load %s\nlll 0.99 0.51 wrapper default 0\nsave %s\n" % (q60, q60l))
    b_in = np.array(O.read_matrix(q60l), dtype=np.int64)
    pack = {"b_in": b_in}
    # (block, flags, max_loops, strategies): BKZ_NO_LLL = 2, BKZ_MAX_LOOPS = 4
    for tag, bs, fl, ml, strat in [("bkz20_none", 20, 2, 0, "none"), ("bkz30_default", 30, 2 | 4, 2, "default"),
                                   ("bkz40_default", 40, 2 | 4, 2, "default")]:
        outp = os.path.join(TMP, "q60_%s.txt" % tag)
        o = O.run_ref("load %s\nbkz %d %d %d %s enumlib 1\nsave %s\n" % (q60l, bs, fl, ml, strat, outp), timeout=600)
        st = int(o.split("bkz status=")[1].split()[0])
        pack[tag + "_status"] = np.int32(st)
        pack[tag + "_b"] = np.array(O.read_matrix(outp), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "bkz_q60.npz"), **pack)


def hlll_fixtures():
    # --- HLLL fixtures: the reference's HLLLReduction<long,double>::hlll (ROW_EXPO | OP_FORCE_LONG, defaults
    # delta 0.99, eta 0.51, theta 0.001, c 0.1) on latticegen inputs that fit int64 -------------------------------
    pack = {}
    for tag, args in [("u40", ["u", 40, 16]), ("r60", ["r", 60, 55]), ("q40", ["q", 40, 20, 12, "b"]),
                      ("u100", ["u", 100, 20]), ("q80", ["q", 80, 40, 16, "b"])]:
        inp = gen(args, "hlll_%s.txt" % tag)
        outp = os.path.join(TMP, "hlll_%s_out.txt" % tag)
        o = O.run_ref("load %s\ntolong\nhlll_long 0.99 0.51 0.001 0.1\nsave_long %s\n" % (inp, outp), timeout=600)
        pack[tag + "_in"] = np.array(O.read_matrix(inp), dtype=np.int64)
        pack[tag + "_out"] = np.array(O.read_matrix(outp), dtype=np.int64)
        pack[tag + "_status"] = np.int32(int(o.split("hlll_long status=")[1].split()[0]))
    np.savez_compressed(os.path.join(HERE, "hlll_long.npz"), **pack)


def svp_kat_fixture():
    # --- the reference's SVP known-answer pair (tests/test_svp.cpp:373-374): input basis and a shortest vector -------
    import re
    src = "/root/reference/tests/lattices/"
    rows = re.findall(r"\[([^\[\]]*)\]", open(src + "example_svp_in").read())
    b = np.array([[int(x) for x in r.split()] for r in rows if r.strip()], dtype=np.int64)
    sv = np.array([int(x) for x in re.findall(r"-?\d+", open(src + "example_svp_out").read())], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "example_svp.npz"), b_in=b, sv=sv)


def main():
    if '--only-svp' in sys.argv:
        return svp_kat_fixture()
    if '--only-bkz' in sys.argv:
        return bkz_fixtures()
    if '--only-hlll' in sys.argv:
        return hlll_fixtures()
    # --- config #1 input: latticegen u 40 40 (md5 a6fe01e1...) -------------------------------
    u40 = np.array(O.read_matrix(gen(["u", 40, 40], "u40.txt")), dtype=np.int64)
    s = O.RefSession(u40)
    s.cmd("update_gso")
    s.dump_state()
    _, recs = s.run()
    save_state(os.path.join(HERE, "u40_update_gso.npz"), recs[0])

    # op-sequence trace on u40 (state machine: row ops, move_row, partial updates), dumped after every op group
    rng = np.random.default_rng(20260923)
    ops = H.random_op_script(rng, 40, 60)
    s = O.RefSession(u40)
    marks = []
    for k, op in enumerate(ops):
        for line in H.ops_to_ref_script([op]):
            s.lines.append(line)
        if op[0] in ("move_row", "update_gso") or (op[0] in ("row_op_end", "update_rows_to") and k % 3 == 0):
            s.dump_state()
            marks.append(k)
    _, recs = s.run()
    assert len(recs) == len(marks)
    import json
    pack = {"ops_json": np.frombuffer(json.dumps(ops).encode(), dtype=np.uint8), "marks": np.array(marks),
            "b0": u40}
    for t, r in enumerate(recs):
        for key in ["n_known_rows", "n_known_cols", "n_source_rows"]:
            pack["s%d_%s" % (t, key)] = np.int32(r[key])
        for key in ["row_expo", "gso_valid_cols", "init_row_size", "bf", "gf", "mu", "r", "b"]:
            pack["s%d_%s" % (t, key)] = r[key]
    np.savez_compressed(os.path.join(HERE, "u40_ops_trace.npz"), **pack)

    # reference LLL on u40 through the <long,double> code path (MatGSO<long,double>+LLLReduction, GSO_ROW_EXPO)
    mat = os.path.join(TMP, "u40.txt")
    out = O.run_ref("load %s\ntolong\nlll_long 0.99 0.51\ngso l 2\nsave %s/u40_lll_long.txt\n" % (mat, TMP))
    st = dict(tok.split("=") for tok in out.split("lll_long")[1].split() if "=" in tok)
    red = np.array(O.read_matrix(TMP + "/u40_lll_long.txt"), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "u40_lll_long.npz"), b_in=u40, b_out=red, status=np.int32(st["status"]),
                        n_swaps=np.int32(st["swaps"]))

    # --- config #2/#5 input after the host wrapper LLL (bkz.cpp:873-875): int64 regime dim-200 -------------
    r200 = gen(["r", 200, 2000], "r200.txt")
    red_path = os.path.join(TMP, "r200_lll.txt")
    if not os.path.exists(red_path):
        O.run_ref("load %s\nlll 0.99 0.51 wrapper default 0\nsave %s\n" % (r200, red_path), timeout=900)
    b200 = np.array(O.read_matrix(red_path), dtype=np.int64)
    s = O.RefSession(b200)
    s.cmd("update_gso")
    s.dump_state()
    _, recs = s.run()
    rec = recs[0]
    # keep the fixture small: b + the derived state in compact form (mu/r lower triangles as float64)
    tl = np.tril_indices(200)
    np.savez_compressed(os.path.join(HERE, "r200_lll_update_gso.npz"), b=b200, row_expo=rec["row_expo"],
                        gso_valid_cols=rec["gso_valid_cols"], init_row_size=rec["init_row_size"],
                        n_known_rows=np.int32(rec["n_known_rows"]), n_known_cols=np.int32(rec["n_known_cols"]),
                        n_source_rows=np.int32(rec["n_source_rows"]),
                        mu_tril=rec["mu"][tl], r_tril=rec["r"][tl], gf_tril=rec["gf"][tl])
    # --- enumeration fixtures (reference's internal enumerator through the plugin-API capture hook) ----------
    import json
    strat = json.load(open(os.path.join(O.REF_DIR, "strategies", "default.json")))
    s60 = [e for e in strat if e["block_size"] == 60][0]
    pr60 = s60["pruning_parameters"][15]   # [gh_factor = 1.05, coefficients, probability ~ 0.41]
    prfile = os.path.join(TMP, "prune60.txt")
    open(prfile, "w").write(" ".join(repr(float(c)) for c in pr60[1]))
    out = os.path.join(TMP, "enum_fixture.bin")
    if os.path.exists(out):
        os.remove(out)
    # radius as BKZ sets it (bkz.cpp:310-323): gh_factor * GH^2 of the block, expressed as a multiple of r(first,first)
    import math
    Rm = np.zeros((200, 200))
    Rm[tl] = rec["r"][tl]
    rii = np.array([Rm[i, i] * 2.0 ** (2 * int(rec["row_expo"][i])) for i in range(200)])

    def gh_ratio(first, last):
        n = last - first
        gh2 = math.exp((2.0 / n) * math.lgamma(n / 2 + 1)) / math.pi * math.exp(np.sum(np.log(rii[first:last])) / n)
        return gh2 / rii[first]
    f100, f140 = float(1.05 * gh_ratio(100, 160)), float(1.05 * gh_ratio(140, 200))
    script = ("load %s\ntolong\ngso l 2\nupdate_gso\n"
              "enum 0 30 0.99 - %s capture\n"
              "enum 100 160 %r %s %s capture\n"
              "enum 140 200 %r %s %s capture\n") % (red_path, out, f100, prfile, out, f140, prfile, out)
    print(O.run_ref(script, timeout=600))
    recs = O.read_enum_records(out)
    for name, r in zip(["enum_r200_b30_unpruned", "enum_r200_b60_pruned_100", "enum_r200_b60_pruned_140"], recs):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mut=r["mut"], rdiag=r["rdiag"], pruning=r["pruning"],
                            maxdist=np.float64(r["maxdist"]), normexp=np.int64(r["normexp"]),
                            found=np.int32(r["found"]), best=np.float64(r["best"]), sol=r["sol"], nodes=r["nodes"],
                            gh_factor=np.float64(pr60[0]))
    # Leech lattice known-answer test (tests/test_enum.cpp:55-100): 196560 minimal vectors of squared norm 32
    leech_in = "/root/reference/tests/lattices/example_list_cvp_in_lattice"
    lo = os.path.join(TMP, "leech_lll.txt")
    O.run_ref("load %s\nlll 0.99 0.51 wrapper default 0\nsave %s\n" % (leech_in, lo))
    np.savez_compressed(os.path.join(HERE, "leech_lll.npz"), b=np.array(O.read_matrix(lo), dtype=np.int64))
    bkz_fixtures()
    hlll_fixtures()
    svp_kat_fixture()
    print("golden fixtures written to", HERE)


def config_size_fixtures():
    """BASELINE.json configs #3-#5 at their stated sizes (python tests/golden/make_golden.py --configs; ~25 min of CPU):
      #3 HLLL on latticegen q 400 200 30 b: the reference's HLLLReduction<long,double> on the full 400 x 400 basis
         (status + output) and on the 120-row sub-lattice rows [140, 260) (a case a test can afford on every run);
      #4 BKZ-40 on the dim-180 Goldstein-Mayer basis (latticegen q 180 1 1800 p), wrapper-LLL first (bkz.cpp:869-876),
         no pruning => deterministic => the output basis is the known answer;
      #5 BKZ-60 with strategies/default.json, one tour, on the LLL-reduced r200 basis: status, r(0,0) and slope."""
    import time
    pack = {}
    q400 = gen(["q", 400, 200, 30, "b"], "q400.txt")
    b400 = np.array(O.read_matrix(q400), dtype=np.int64)
    sub = os.path.join(TMP, "q400_rows140_260.txt")
    O.write_matrix(sub, b400[140:260])
    for tag, inp in (("q400sub", sub), ("q400", q400)):
        outp = os.path.join(TMP, "hlll_%s_out.txt" % tag)
        t0 = time.time()
        o = O.run_ref("load %s\ntolong\nhlll_long 0.99 0.51 0.001 0.1\nsave_long %s\n" % (inp, outp), timeout=7200)
        pack[tag + "_hlll_sec"] = np.float64(time.time() - t0)
        pack[tag + "_in"] = np.array(O.read_matrix(inp), dtype=np.int64)
        pack[tag + "_out"] = np.array(O.read_matrix(outp), dtype=np.int64)
        pack[tag + "_status"] = np.int32(int(o.split("hlll_long status=")[1].split()[0]))
        print(tag, "hlll status", pack[tag + "_status"], "sec", pack[tag + "_hlll_sec"], flush=True)
    del pack["q400_in"], pack["q400_out"]  # the full run ends in status 10 after minutes: keep its status and time only
    np.savez_compressed(os.path.join(HERE, "hlll_q400.npz"), **pack)
    pack = {}
    gm = gen(["q", 180, 1, 1800, "p"], "gm180.txt")
    gml = os.path.join(TMP, "gm180_lll.txt")
    O.run_ref("load %s\nlll 0.99 0.51 wrapper default 0\nsave %s\n" % (gm, gml), timeout=3600)
    pack["b_in"] = np.array(O.read_matrix(gml), dtype=np.int64)
    outp = os.path.join(TMP, "gm180_bkz40.txt")
    t0 = time.time()
    o = O.run_ref("load %s\nbkz 40 %d 1 none enumlib 1\nsave %s\n" % (gml, 2 | 4, outp), timeout=7200)
    pack["bkz40_none_sec"] = np.float64(time.time() - t0)
    pack["bkz40_none_status"] = np.int32(int(o.split("bkz status=")[1].split()[0]))
    pack["bkz40_none_b"] = np.array(O.read_matrix(outp), dtype=np.int64)
    print("gm180 bkz40 status", pack["bkz40_none_status"], "sec", pack["bkz40_none_sec"], flush=True)
    np.savez_compressed(os.path.join(HERE, "bkz_gm180.npz"), **pack)
    g = np.load(os.path.join(HERE, "r200_lll_update_gso.npz"))
    mat = os.path.join(TMP, "r200_red.txt")
    O.write_matrix(mat, g["b"])
    outp = os.path.join(TMP, "r200_bkz60.txt")
    o = O.run_ref("load %s\nbkz 60 %d 1 default enumlib 1\nsave %s\n" % (mat, 2 | 4, outp), timeout=7200)
    b = np.array(O.read_matrix(outp), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "bkz60_r200_ref.npz"), status=np.int32(int(o.split("bkz status=")[1].split()[0])),
                        b_out=b)
    print(o, flush=True)


def enum_dual_subsol_fixtures():
    """Dual SVP enumeration and sub-solutions: block [140,170) of the LLL-reduced r200 basis, unpruned,
    through the capture hook (records what the plugin API hands an external enumerator, then the reference's own
    enumerator runs).  Standalone: python tests/golden/make_golden.py --dual"""
    g = np.load(os.path.join(HERE, "r200_lll_update_gso.npz"))
    os.makedirs(TMP, exist_ok=True)
    mat = os.path.join(TMP, "r200_red.txt")
    O.write_matrix(mat, g["b"])
    out = os.path.join(TMP, "enum_dual_fixture.bin")
    if os.path.exists(out):
        os.remove(out)
    modes = ["capture", "capture_dual", "capture_subsols", "capture_dual_subsols"]
    script = "load %s\ntolong\ngso l 2\nupdate_gso\n" % mat + "".join(
        "enum 140 170 0.99 - %s %s\n" % (out, m) for m in modes)
    print(O.run_ref(script, timeout=600))
    recs = O.read_enum_records(out)
    arrs = dict(mut=recs[0]["mut"], rdiag=recs[0]["rdiag"], pruning=recs[0]["pruning"])
    for m, r in zip(["primal", "dual", "subsols", "dual_subsols"], recs):
        assert np.array_equal(r["mut"], recs[0]["mut"]) and np.array_equal(r["rdiag"], recs[0]["rdiag"])
        arrs[m + "_maxdist"] = np.float64(r["maxdist"])
        arrs[m + "_best"] = np.float64(r["best"])
        arrs[m + "_normexp"] = np.int64(r["normexp"])
        arrs[m + "_sol"] = r["sol"]
        arrs[m + "_nodes"] = r["nodes"]
        if r["subsols"]:
            arrs[m + "_subdist"] = r["subdist"]
            arrs[m + "_subsol"] = r["subsol"]
    np.savez_compressed(os.path.join(HERE, "enum_r200_b30_dual_subsols.npz"), **arrs)


def config2_mpz_reference():
    """BASELINE config #2 as written (LLL<mpz_t,double> on latticegen r 200 2000): the reference's own outcome — status,
    GSO dump and basis of tests/shim_demo without the shim — as hashes in lll_r200_mpz_ref.json (33 s on one core)."""
    import hashlib
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(HERE))
    inp, out = os.path.join(TMP, "r200.txt"), os.path.join(TMP, "r200_out.bin")
    open(inp, "w").write(O.latticegen(["r", 200, 2000]))
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(root, "oracle", "_ref"))
    env.pop("LD_PRELOAD", None)
    p = subprocess.run([os.path.join(root, "tests", "_build", "shim_demo"), inp, "mpz", out], capture_output=True,
                       text=True, env=env, timeout=900)
    tok = dict(t.split("=") for t in p.stdout.split() if "=" in t)
    sha = lambda f: hashlib.sha256(open(f, "rb").read()).hexdigest()
    rec = dict(what="BASELINE config #2: reference LLLReduction<Z_NR<mpz_t>, FP_NR<double>> on latticegen r 200 2000",
               made_by="tests/golden/make_golden.py::config2_mpz_reference",
               input_md5=hashlib.md5(open(inp, "rb").read()).hexdigest(), status=int(tok["status"]),
               gso_ok=int(tok["gso_ok"]), out_bin_sha256=sha(out), out_basis_sha256=sha(out + ".basis"))
    json.dump(rec, open(os.path.join(HERE, "lll_r200_mpz_ref.json"), "w"), indent=2)
    print(rec)


if __name__ == "__main__":
    if "--config2" in sys.argv:
        config2_mpz_reference()
    elif "--dual" in sys.argv:
        enum_dual_subsol_fixtures()
    elif "--configs" in sys.argv:
        config_size_fixtures()
    else:
        main()
