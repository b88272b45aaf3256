"""Known-answer vectors that the reference's own PyTAT test-suite holds for this path,
restated against this repository's ``TAT`` module.  They pin the integer rules (block order and
offsets, merged-edge segment order, fermi signs of transpose / merge) and the contract / svd
corner cases WITHOUT needing oracle/_ref, so they run on the CPU checker (-m "not gpu") and, in
test_tat_gpu.py, on the CUDA path.

Sources (reference file:line):
  PyTAT/tests/test_create_symmetry_tensor.py:5-29, 32-75   block layout, rank-0, zero-size
  PyTAT/tests/test_split_and_merge.py:32-40, 66-105         merged edge order, merge sign (n_odd & 2)
  PyTAT/tests/test_transpose.py:88-140                      transposition parity
  PyTAT/tests/test_contract.py:5-37, 71-115, 176-214        contract integers, fermi == bose, corners
  PyTAT/tests/test_edge_operator.py:5-46                    fused edge_operator == split.merge.transpose
  PyTAT/tests/test_svd.py:94-129                            truncation counts
  PyTAT/tests/test_scalar.py:41-60                          range_ accumulates first + k*step by repeated addition
"""
import itertools

import numpy as np

import tnsp_b200.TAT as TAT


def _st(t):
    return np.asarray(t.storage, dtype=np.float64).reshape(-1)


def test_z2_block_layout_offsets():
    a = TAT.BoseZ2.D.Tensor(["Left", "Right", "Up"], [[(True, 3), (False, 1)], [(True, 1), (False, 2)], [(True, 2), (False, 3)]]).range_()
    assert a.names == ["Left", "Right", "Up"] and a.rank == 3
    assert a.storage.size == 1 * 2 * 3 + 1 * 1 * 2 + 3 * 2 * 2 + 3 * 1 * 3
    assert a.blocks[[("Left", True), ("Right", False), ("Up", True)]].shape == (3, 2, 2)
    assert a.blocks[[("Left", False), ("Right", True), ("Up", True)]].shape == (1, 1, 2)
    # block order (T,T,F) (T,F,T) (F,T,T) (F,F,F): offsets 0, 9, 9+12, 9+12+2
    assert a[{"Left": (True, 2), "Right": (False, 0), "Up": (True, 1)}] == 9 + 9
    assert a[{"Left": 2, "Right": 1, "Up": 1}] == 9 + 9
    assert a[{"Left": (False, 0), "Right": (False, 1), "Up": (False, 2)}] == 9 + 12 + 2 + 5
    assert a[{"Left": 3, "Right": 2, "Up": 4}] == 9 + 12 + 2 + 5


def test_rank0_zero_size_zero_block():
    a = TAT.BoseU1.D.Tensor([], []).range_(2333)
    assert a.rank == 0 and _st(a).tolist() == [2333.0] and a[{}] == 2333
    b = TAT.BoseU1.D.Tensor(["Left", "Right", "Up"], [[(0, 0)], [(-1, 1), (0, 2), (1, 3)], [(-1, 2), (0, 3), (1, 1)]]).zero_()
    assert b.storage.size == 0
    assert b.blocks[[("Left", 0), ("Right", +1), ("Up", -1)]].shape == (0, 3, 2)
    c = TAT.BoseU1.D.Tensor(["Left", "Right", "Up"], [[], [(-1, 1), (0, 2), (1, 3)], [(-1, 2), (0, 3), (1, 1)]]).zero_()
    assert c.storage.size == 0


def test_merged_edge_first_appearance_order():
    a = TAT.BoseU1.D.Tensor(["i", "j"], [[-1, 0, +1], [-1, 0, +1]]).range_()
    d = TAT.BoseU1.D.Tensor(["m"], [[(-2, 1), (-1, 2), (0, 3), (+1, 2), (+2, 1)]]).range_()
    b = a.merge_edge({"m": ["i", "j"]})
    assert b.edge_by_name("m") == d.edge_by_name("m")
    assert (d - b).norm_max() == 0
    c = b.split_edge({"m": [("i", [-1, 0, +1]), ("j", [-1, 0, +1])]})
    assert (c - a).norm_max() == 0


def test_merge_split_round_trip_high_rank():
    edge = [(-1, 2), (0, 2), (+1, 2)]
    for mod, parities in ((TAT.BoseU1, [False]), (TAT.FermiU1, [False, True])):
        a = mod.D.Tensor(list("12345"), [edge] * 5).range_()
        for i in range(5):
            for j in range(i, 5):
                for p in parities:
                    names = a.names[i:j]
                    b = a.merge_edge({"m": names}, p)
                    c = b.split_edge({"m": [(n, edge) for n in names]}, p)
                    assert (c - a).norm_max() == 0


def test_fermi_merge_sign_is_count_and_2():
    edge = [(-1, 1), (0, 1), (+1, 1)]
    a_u1 = TAT.BoseU1.D.Tensor(list("12345"), [edge] * 5).range_()
    a_f = TAT.FermiU1.D.Tensor(list("12345"), [edge] * 5).range_()
    for i in range(5):
        for j in range(i, 5):
            names = a_u1.names[i:j]
            b_u1 = _st(a_u1.merge_edge({"m": names}))
            assert np.array_equal(b_u1, _st(a_f.merge_edge({"m": names}, False)))
            b_f = _st(a_f.merge_edge({"m": names}, True))
            for s in itertools.product([-1, 0, 1], repeat=5):
                if sum(s) != 0:
                    continue
                item = a_u1[{str(k + 1): (s[k], 0) for k in range(5)}]
                assert item in b_u1
                odd = sum(s[x] != 0 for x in range(i, j))
                assert ((-item) if (odd & 2) else item) in b_f


def test_transpose_parity_sign():
    edge = TAT.FermiZ2.Edge([(False, 2), (True, 2)])
    a = TAT.FermiZ2.D.Tensor(list("ijklmn"), [edge] * 6).range_()
    b = a.transpose(["l", "j", "i", "n", "k", "m"])
    for idx in itertools.product(range(4), repeat=6):
        p = [x >= 2 for x in idx]
        if p[0] ^ p[1] ^ p[2] ^ p[3] ^ p[4] ^ p[5]:
            continue
        pi, pj, pk, pl, pm, pn = p
        sign = (pl and (pi ^ pj ^ pk)) ^ (pj and pi) ^ (pn and (pk ^ pm))
        pos = dict(zip("ijklmn", idx))
        assert b[pos] == (-a[pos] if sign else a[pos])


def test_contract_integers():
    T = TAT.No.D.Tensor
    a = T(["A", "B"], [2, 2]).range_()
    b = T(["C", "D"], [2, 2]).range_()
    for pair, names, want in ((("A", "C"), ["B", "D"], [4, 6, 6, 10]), (("A", "D"), ["B", "C"], [2, 6, 3, 11]),
                              (("B", "C"), ["A", "D"], [2, 3, 6, 11]), (("B", "D"), ["A", "C"], [1, 3, 3, 13])):
        c = a.contract(b, {pair})
        assert c.names == names and _st(c).tolist() == want
    x = T(["A", "B", "C", "D"], [1, 2, 3, 4]).range_()
    y = T(["E", "F", "G", "H"], [3, 1, 2, 4]).range_()
    z = x.contract(y, {("B", "G"), ("D", "H")})
    assert z.names == ["A", "C", "E", "F"]
    assert _st(z).tolist() == [316, 796, 1276, 428, 1164, 1900, 540, 1532, 2524]


def test_contract_fermi_equals_bose_for_this_layout():
    e1, e2 = [(-1, 2), (0, 2), (+1, 2)], [(+1, 2), (0, 2), (-1, 2)]
    F, U = TAT.FermiU1.D.Tensor, TAT.BoseU1.D.Tensor
    fa = F(list("abcd"), [(e1, True), e2, (e1, True), (e2, True)]).range_()
    fb = F(list("efgh"), [e1, e2, (e1, True), e2]).range_()
    fc = fa.contract(fb, {("d", "e"), ("c", "f")})
    fd = fb.contract(fa, {("e", "d"), ("f", "c")})
    assert (fc - fd).norm_max() == 0
    ua = U(list("abcd"), [e1, e2, e1, e2]).range_()
    ub = U(list("efgh"), [e1, e2, e1, e2]).range_()
    uc = ua.contract(ub, {("d", "e"), ("c", "f")})
    assert (uc - ub.contract(ua, {("e", "d"), ("f", "c")})).norm_max() == 0
    assert np.array_equal(_st(fc), _st(uc))


def test_contract_with_merge_and_reverse_signs():
    T = TAT.FermiU1.D.Tensor
    e1, e2 = ([(-1, 2), (0, 2), (+1, 2)], False), ([(+1, 2), (0, 2), (-1, 2)], True)
    a = T(list("abcd"), [e1, e2, e1, e2]).range_()
    b = T(list("efgh"), [e1, e2, e1, e2]).range_()
    c = a.contract(b, {("a", "f"), ("b", "g"), ("c", "h")})
    cm = a.merge_edge({"m": ["b", "a"]}, False).contract(b.merge_edge({"m": ["g", "f"]}, True), {("m", "m"), ("c", "h")})
    assert (c - cm).norm_max() == 0
    cr = a.reverse_edge({"b", "a"}, False).contract(b.reverse_edge({"g", "f"}, True), {("a", "f"), ("b", "g"), ("c", "h")})
    assert (c - cr).norm_max() == 0
    z = TAT.FermiZ2.D.Tensor
    ez = [(False, 2), (True, 2)]
    p = z(["i", "j"], [(ez, False), (ez, True)]).range_()
    q = z(["i", "j"], [(ez, False), (ez, True)]).range_().transpose(["j", "i"])
    r = p.contract(q, {("j", "i")})
    rr = p.reverse_edge({"j"}, False).contract(q.reverse_edge({"i"}, True), {("j", "i")})
    assert (r - rr).norm_max() == 0


def test_contract_corner_cases():
    N, Z = TAT.No.D.Tensor, TAT.BoseZ2.D.Tensor
    c = N(["A", "B"], [2, 0]).range_().contract(N(["C", "D"], [0, 2]).range_(), {("B", "C")})
    assert c.storage.size == 4 and c.norm_max() == 0
    c = Z(["A", "B"], [[(False, 2)], [(False, 0)]]).range_().contract(Z(["C", "D"], [[(False, 0)], [(False, 2)]]).range_(), {("B", "C")})
    assert c.storage.size == 4 and c.norm_max() == 0
    c = Z(["A", "B"], [[(True, 2)], [(False, 0)]]).range_().contract(Z(["C", "D"], [[(False, 0)], [(False, 2)]]).range_(), {("B", "C")})
    assert c.storage.size == 0
    c = Z(["A", "B"], [[(False, 2)], [(False, 0)]]).range_().contract(Z(["C", "D"], [[(False, 0)], [(True, 2)]]).range_(), {("B", "C")})
    assert c.storage.size == 0
    c = Z(["A", "B"], [[(False, 2)], [(True, 0)]]).range_().contract(Z(["C", "D"], [[(True, 0)], [(False, 2)]]).range_(), {("B", "C")})
    assert c.storage.size == 4 and c.norm_max() == 0


def test_contract_fuse():
    N = TAT.No.D.Tensor
    a = N(["A", "B", "C"], [3, 4, 5]).range_()
    b = N(["A", "B", "D"], [3, 4, 7]).range_()
    c = a.contract(b, {("B", "B")}, {"A"})
    for i in range(3):
        hat = N(["A"], [3]).zero_()
        hat.storage[i] = 1
        a0, b0, c0 = (t.contract(hat, {("A", "A")}) for t in (a, b, c))
        assert (a0.contract(b0, {("B", "B")}) - c0).norm_max() == 0


def test_edge_operator_equals_split_merge_transpose():
    a = TAT.No.D.Tensor(["A", "B"], [8, 8]).range_().edge_rename({"A": "C"})
    sp = {"C": [("D", 4), ("E", 2)], "B": [("F", 2), ("G", 4)]}
    mg = {"I": ["D", "F"], "J": ["G", "E"]}
    fused = a.edge_operator(sp, {"D", "F"}, mg, ["J", "I"])
    steps = a.split_edge(sp).merge_edge(mg).transpose(["J", "I"])
    assert (fused - steps).norm_max() == 0
    for mod in (TAT.BoseU1, TAT.FermiU1):
        t = mod.D.Tensor(["Left", "Right", "Up", "Down"], [[(-1, 3), (0, 1), (1, 2)], [(-1, 1), (0, 4), (1, 2)], [(-1, 2), (0, 3), (1, 1)],
                                                           [(-1, 1), (0, 3), (1, 2)]]).range_().edge_rename({"Right": "Right1"})
        sp = {"Down": [("Down1", [(0, 1), (1, 2)]), ("Down2", [(-1, 1), (0, 1)])]}
        d = t.split_edge(sp).transpose(["Down1", "Right1", "Up", "Left", "Down2"]).merge_edge({"Left": ["Left", "Down2"]})
        total = t.edge_operator(sp, set(), {"Left": ["Left", "Down2"]}, ["Down1", "Right1", "Up", "Left"])
        assert (total - d).norm_max() == 0


def _unitary(t, name, fermi):
    d = t.conjugate().edge_rename({name: name + "'"})
    pairs = {(n, n) for n in t.names if n != name}
    m = t.contract(d, pairs)
    ident = m.same_shape().identity_({(name, name + "'")}) if not fermi else None
    if ident is not None:
        assert (m - ident).norm_max() < 1e-6


def test_svd_cut_counts():
    a = TAT.No.D.Tensor(["A", "B"], [2, 2]).zero_()
    a[{"A": 0, "B": 0}] = 1
    u, s, v = a.svd({"B"}, "E", "F", "U", "V", 8)
    _unitary(u, "E", False)
    _unitary(v, "F", False)
    b = v.contract(s, {("F", "V")}).contract(u, {("U", "E")})
    assert (a - b.transpose(a.names)).norm_max() < 1e-6
    assert s.storage.size == 1
    f = TAT.FermiU1.D.Tensor(["A", "B"], [[(0, 1), (+1, 1)], [(-1, 1), (0, 1)]]).range_(0, 1)
    u, s, v = f.svd({"B"}, "E", "F", "U", "V", 8)
    b = v.contract(s, {("F", "V")}).contract(u, {("U", "E")})
    assert (f - b.transpose(f.names)).norm_max() < 1e-6
    assert s.storage.size == 1
    g = TAT.FermiU1.D.Tensor(["A", "B", "C", "D"], [([(-1, 1), (0, 1), (-2, 1)], True), ([(0, 1), (1, 2)], False), ([(0, 2), (1, 2)], False),
                                                    ([(-2, 2), (-1, 1), (0, 2)], True)]).range_()
    u, s, v = g.svd({"B", "D"}, "E", "F", "U", "V", 8)
    b = v.contract(s, {("F", "V")}).contract(u, {("U", "E")})
    assert (g - b.transpose(g.names)).norm_max() < 1e-6 * g.norm_max()


def test_range_accumulates_like_the_reference():
    b = TAT.No.D.Tensor(["Left", "Right"], [3, 4]).range_(0, 0.1)
    want = np.zeros(12)
    for i in range(1, 12):
        want[i] = want[i - 1] + 0.1
    assert np.array_equal(_st(b), want)


def test_rng_seed_reproducible_and_in_range():
    TAT.random.seed(233)
    x = [TAT.random.uniform_int(0, 100)() for _ in range(8)]
    TAT.random.seed(233)
    assert x == [TAT.random.uniform_int(0, 100)() for _ in range(8)]
    assert all(0 <= v <= 100 for v in x)
    r = TAT.random.uniform_real(0, 100)
    assert all(0 <= r() <= 100 for _ in range(1000))
