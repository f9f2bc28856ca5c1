"""GPU parity tests of the device Householder state (include/b200hh.h) against the oracle restatement, which is itself
pinned bit-exactly to the reference (tests/test_hh_oracle.py).  Same HLLL-like call order as hlll.cpp:49-171."""
import numpy as np
import pytest

import helpers as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _hlll_like_ops(d):
    ops = [("refresh_R_bf", 0), ("update_R_last", 0), ("refresh_R_bf", 1)]
    k, kmax = 1, 1
    for step in range(3 * d):
        ops += [("update_R", k, 0), ("size_reduce", k, k, 0), ("refresh_R_bf", k), ("update_R", k, 0), ("check",)]
        if step % 3 != 2 or k == 1:
            ops += [("update_R_last", k)]
            k += 1
            if k >= d:
                break
            ops += [("refresh_R_bf", k)] if k > kmax else [("refresh_R", k)]
            kmax = max(kmax, k)
        else:
            ops += [("swap", k - 1, k)]
            k -= 1
            ops += [("recover_R", k), ("check",), ("set_updated_R_false",)]
    ops.append(("check",))
    return ops, kmax


@pytest.mark.parametrize("seed,d,n,bits,batch", [(1, 10, 10, 20, 1), (2, 24, 30, 12, 3), (3, 40, 40, 30, 2),
                                                 (4, 70, 75, 10, 2)])
def test_hlll_like_sequence_bit_exact(seed, d, n, bits, batch):
    from fplll_b200.householder import MatHouseholder
    rng = np.random.default_rng(seed)
    b = rng.integers(-(1 << bits), 1 << bits, size=(batch, d, n), dtype=np.int64)
    ops, kmax = _hlll_like_ops(d)
    md = MatHouseholder(b, 5)
    mos = [O.OracleHouseholder(b[l], 5) for l in range(batch)]
    checks = 0
    for op in ops:
        if op[0] == "check":
            st = md.state()
            for l, mo in enumerate(mos):
                s = mo.state()
                nk = s["n_known_rows"]
                what = "seed %d lattice %d check %d" % (seed, l, checks)
                assert st["n_known_rows"][l] == nk and st["n_known_cols"][l] == s["n_known_cols"], what
                assert np.array_equal(st["b"][l], s["b"]), what + " b"
                rows = min(d, kmax + 1)
                for k in ["row_expo", "expo_norm_square_b"]:
                    assert np.array_equal(st[k][l][:rows], s[k][:rows]), what + " " + k
                for k in ["sigma", "norm_square_b"]:
                    assert H.eq_f64(st[k][l][:nk], s[k][:nk]), what + " " + k
                assert H.eq_f64(st["bf"][l][:rows], s["bf"][:rows]), what + " bf"
                assert H.eq_f64(st["R"][l][:rows], s["R"][:rows]), what + " R"
                assert H.eq_f64(st["V"][l][:nk], s["V"][:nk]), what + " V"
            checks += 1
        else:
            r = getattr(md, op[0])(*op[1:])
            for l, mo in enumerate(mos):
                ro = getattr(mo, op[0])(*op[1:])
                if op[0] == "size_reduce":
                    assert bool(r[l]) == bool(ro)
    assert checks > 5


def test_full_qr_matches_gso_relation():
    """tests/test_gso.cpp:101-152 on the device R."""
    from fplll_b200.householder import MatHouseholder
    b = H.gold("bkz_q60.npz")["b_in"]
    d = b.shape[0]
    md = MatHouseholder(b, 0, keep_history=False)
    for i in range(d):
        md.refresh_R_bf(i)
        md.update_R(i)
    R = md.state()["R"][0]
    g = O.OracleGSO(b, 0)
    assert g.update_gso()
    s = g.state()
    for i in range(d):
        assert R[i, i] > 0
        for j in range(i):
            assert abs(R[i, j] / R[j, j] - s["mu"][i, j]) < 1e-9 * max(1.0, abs(s["mu"][i, j]))
            assert abs(R[i, j] * R[j, j] - s["r"][i, j]) < 1e-9 * max(1.0, abs(s["r"][i, j]))


@pytest.mark.parametrize("tag", ["u40", "r60", "q40", "u100", "q80"])
def test_device_hlll_equals_reference_golden(tag):
    """b200hh_hlll (HLLLReduction::hlll, hlll.cpp:25-171, whole loop on the device) ends on the basis the reference's
    HLLLReduction<long,double> produced; a batch of copies plus one differently-scaled lattice checks lattices of a
    batch do not interact."""
    from fplll_b200.householder import hlll_reduction
    g = H.gold("hlll_long.npz")
    b_in = g[tag + "_in"]
    out, st = hlll_reduction(b_in)
    assert st == int(g[tag + "_status"]) == 0
    assert np.array_equal(out, g[tag + "_out"])
    batch = np.stack([b_in, 3 * b_in, b_in])
    outs, sts = hlll_reduction(batch)
    assert list(sts) == [0, 0, 0]
    assert np.array_equal(outs[0], g[tag + "_out"]) and np.array_equal(outs[2], g[tag + "_out"])
    mo = O.OracleHouseholder(3 * b_in)
    assert mo.hlll() == 0 and np.array_equal(outs[1], mo.state()["b"])


@pytest.mark.parametrize("seed,d,n,bits,batch", [(11, 30, 30, 25, 4), (12, 50, 55, 12, 3), (13, 70, 70, 40, 2),
                                                 (14, 1, 5, 10, 2), (15, 2, 2, 30, 2)])
def test_device_hlll_equals_oracle_random(seed, d, n, bits, batch):
    from fplll_b200.householder import hlll_reduction
    rng = np.random.default_rng(seed)
    b = rng.integers(-(1 << bits), 1 << bits, size=(batch, d, n), dtype=np.int64)
    outs, sts = hlll_reduction(b)
    for l in range(batch):
        mo = O.OracleHouseholder(b[l])
        assert mo.hlll() == sts[l], "lattice %d status" % l
        assert np.array_equal(outs[l], mo.state()["b"]), "lattice %d basis" % l


def test_device_hlll_needs_history():
    from fplll_b200.householder import MatHouseholder
    from fplll_b200 import B200Error
    m = MatHouseholder(H.gold("hlll_long.npz")["u40_in"], 5, keep_history=False)
    with pytest.raises(B200Error):
        m.hlll()


def test_hlll_config3_sublattice_of_q400(fb=None):
    """BASELINE config #3 shape (latticegen q 400 200 30 b, n = 400 columns): the whole device hlll() loop on the
    120-row sub-lattice rows [140, 260) ends on the basis of the reference's HLLLReduction<long,double>
    (tests/golden/hlll_q400.npz; the reference's own run on all 400 rows ends in status 10 after 260 s of CPU —
    fp64 is not enough there, make_golden.py --configs)."""
    import fplll_b200
    z = H.gold("hlll_q400.npz")
    out, st = fplll_b200.hlll_reduction(z["q400sub_in"].copy())
    assert st == int(z["q400sub_status"]) == 0
    assert np.array_equal(out, z["q400sub_out"])


@pytest.mark.parametrize("d,n,B,bits", [(40, 40, 70, 20), (33, 50, 97, 12), (64, 64, 64, 30)])
def test_update_R_lane_per_lattice_equals_warp_per_lattice_and_oracle(monkeypatch, d, n, B, bits):
    """hk_update_R_x32 (32 lattices per warp, operands transposed by the TMA unit; taken for batches without R_history)
    against hk_update_R (one warp per lattice, B200_HH_X32=0) on the same QR sweep, bit for bit, and against the oracle
    on a sample of lattices.  Batches that are not a multiple of 32, n not a multiple of 16, d < n."""
    from fplll_b200.householder import MatHouseholder
    rng = np.random.default_rng(1000 + d)
    b = rng.integers(-(1 << bits), 1 << bits, size=(B, d, n), dtype=np.int64)
    states = {}
    for x32 in ("1", "0"):
        monkeypatch.setenv("B200_HH_X32", x32)
        m = MatHouseholder(b, 5, keep_history=False)
        for i in range(d):
            m.refresh_R_bf(i)
            m.update_R(i, last_j=(i % 3 != 1))   # both forms of the call: with and without update_R_last
            if i % 3 == 1:
                m.update_R_last(i)
        states[x32] = m.state()
        m.close()
    a, c = states["1"], states["0"]
    for k in ("R", "V", "sigma"):
        assert H.eq_f64(a[k], c[k]), k
    for l in (0, 31, 32, B - 1):
        mo = O.OracleHouseholder(b[l], 5)
        for i in range(d):
            mo.refresh_R_bf(i)
            mo.update_R(i)
        s = mo.state()
        assert H.eq_f64(a["R"][l], s["R"]), "R lattice %d" % l
        assert H.eq_f64(a["V"][l], s["V"]), "V lattice %d" % l
