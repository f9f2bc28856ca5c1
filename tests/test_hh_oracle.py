"""CPU tests (no GPU): the C restatement of MatHouseholder<long,double> (oracle/hh_oracle.c) against the UNMODIFIED
reference driven call by call through oracle/_ref/ref_probe (bit-exact), plus the reference's own test_householder
relation (tests/test_gso.cpp:101-152): mu(i,j) = R(i,j)/R(j,j) and r(i,j) = R(i,j)*R(j,j) against the GSO."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O


def _run_both(b, ops, flags=5):
    tmp = O.tempfile.mkdtemp(prefix="hh_")
    mat, dump = os.path.join(tmp, "in.txt"), os.path.join(tmp, "d.bin")
    O.write_matrix(mat, b)
    script = ["load " + mat, "tolong", "hh %d" % flags]
    m = O.OracleHouseholder(b, flags)
    states = []
    for op in ops:
        if op[0] == "dump":
            script.append("hh_dump " + dump)
            states.append(m.state())
        else:
            script.append("hh_" + op[0] + " " + " ".join(str(int(x)) for x in op[1:]))
            getattr(m, op[0])(*op[1:])
    O.run_ref("\n".join(script) + "\n")
    return states, O.read_hh_dumps(dump)


def _assert_same(s, r, rows_R, what):
    nk = r["n_known_rows"]
    assert s["n_known_rows"] == nk and s["n_known_cols"] == r["n_known_cols"], what
    for k in ["row_expo", "expo_norm_square_b", "b"]:
        assert np.array_equal(s[k], r[k]), what + " " + k
    for k in ["sigma", "norm_square_b", "bf"]:
        assert H.eq_f64(s[k][:rows_R], r[k][:rows_R]), what + " " + k
    assert H.eq_f64(s["R"][:rows_R], r["R"][:rows_R]), what + " R"
    assert H.eq_f64(s["V"][:nk], r["V"][:nk]), what + " V"


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,d,n,bits", [(1, 10, 10, 20), (2, 24, 30, 12), (3, 40, 40, 30)])
def test_hlll_like_call_sequence_live_vs_reference(seed, d, n, bits):
    """refresh_R_bf / update_R / size_reduce (with integer + float row operations) / swap / recover_R in the order
    HLLLReduction::hlll issues them (hlll.cpp:49-171) on an UNREDUCED random basis."""
    rng = np.random.default_rng(seed)
    b = rng.integers(-(1 << bits), 1 << bits, size=(d, n), dtype=np.int64)
    ops = [("refresh_R_bf", 0), ("update_R_last", 0), ("refresh_R_bf", 1)]
    k, kmax = 1, 1
    for step in range(3 * d):
        ops += [("update_R", k, 0), ("size_reduce", k, k, 0), ("refresh_R_bf", k), ("update_R", k, 0), ("dump",)]
        if step % 3 != 2 or k == 1:   # "Lovasz holds": finish the row and go up
            ops += [("update_R_last", k)]
            k += 1
            if k >= d:
                break
            ops += [("refresh_R_bf", k)] if k > kmax else [("refresh_R", k)]
            kmax = max(kmax, k)
        else:                           # "Lovasz fails": swap down and recover
            ops += [("swap", k - 1, k)]
            k -= 1
            ops += [("recover_R", k), ("dump",), ("set_updated_R_false",)]
    ops.append(("dump",))
    states, recs = _run_both(b, ops)
    assert len(states) == len(recs) > 5
    for t, (s, r) in enumerate(zip(states, recs)):
        _assert_same(s, r, min(d, kmax + 1), "seed %d dump %d" % (seed, t))


def test_householder_relation_with_gso():
    """tests/test_gso.cpp:101-152: R(i,j)/R(j,j) = mu(i,j), R(i,j)*R(j,j) = r(i,j), R(i,i) > 0."""
    b = H.gold("bkz_q60.npz")["b_in"]
    d = b.shape[0]
    hh = O.OracleHouseholder(b, 0)
    for i in range(d):
        hh.refresh_R_bf(i)
        hh.update_R(i)
    R = hh.state()["R"]
    g = O.OracleGSO(b, 0)
    assert g.update_gso()
    s = g.state()
    for i in range(d):
        assert R[i, i] > 0
        for j in range(i):
            assert abs(R[i, j] / R[j, j] - s["mu"][i, j]) < 1e-9 * max(1.0, abs(s["mu"][i, j]))
            assert abs(R[i, j] * R[j, j] - s["r"][i, j]) < 1e-9 * max(1.0, abs(s["r"][i, j]))


HLLL_TAGS = ["u40", "r60", "q40", "u100", "q80"]


@pytest.mark.parametrize("tag", HLLL_TAGS)
def test_oracle_hlll_equals_reference_golden(tag):
    """ohh_hlll (the restated HLLLReduction::hlll, hlll.cpp:25-171) ends on the basis the reference's own
    HLLLReduction<long,double> produced (tests/golden/hlll_long.npz, made by make_golden.py --only-hlll)."""
    g = H.gold("hlll_long.npz")
    m = O.OracleHouseholder(g[tag + "_in"])
    st = m.hlll(0.99, 0.51, 0.001, 0.1)
    assert st == int(g[tag + "_status"]) == 0
    assert np.array_equal(m.state()["b"], g[tag + "_out"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,d,bits", [(5, 30, 25), (6, 50, 12), (7, 70, 40)])
def test_oracle_hlll_live_vs_reference(seed, d, bits):
    """same, on seeded random bases, against the reference run live (status and output basis)."""
    rng = np.random.default_rng(seed)
    b = rng.integers(-(1 << bits), 1 << bits, size=(d, d), dtype=np.int64)
    tmp = O.tempfile.mkdtemp(prefix="hlll_")
    pin, pout = os.path.join(tmp, "in.txt"), os.path.join(tmp, "out.txt")
    O.write_matrix(pin, b)
    o = O.run_ref("load %s\ntolong\nhlll_long 0.99 0.51 0.001 0.1\nsave_long %s\n" % (pin, pout))
    rst = int(o.split("hlll_long status=")[1].split()[0])
    m = O.OracleHouseholder(b)
    assert m.hlll() == rst
    assert np.array_equal(m.state()["b"], np.array(O.read_matrix(pout), dtype=np.int64))
