"""Boundary-MPS environments of one sampled configuration (or of a lock-step batch of them).

Behavioural mirror of the reference's ``SingleLayerAuxiliaries``
(tetragono/tetragono/auxiliaries/single_layer_auxiliaries.py):

* boundary MPS in four directions, each row/column compressed by QR sweep + truncated SVD sweep
  (``_two_line_to_one_line`` :462-498) -- this is where the batched QR / Jacobi-SVD kernels run;
* 3-row and 4-row "inline" strips (:254-428);
* ``replace`` (amplitude with <= 2x2 sites replaced, :500-708) with the reference's
  cache-state dependent route selection, and ``hole`` (:710-812).

The implementation is direction-generic (one strip builder parameterised by a direction table
instead of one method per direction) and every tensor may carry a leading chain axis, so one call
evaluates all Monte-Carlo chains of a batch.  Tensor names follow the reference exactly
(L1,R1,..,U3,D3, L0/R0/U0/D0 for holes) because they are part of the drop-in contract.
"""
from __future__ import annotations

from .. import lazy


def safe_contract(t1, t2, pairs):
    """contract on those pairs whose names exist on both sides (utility.py:340-350)"""
    n1, n2 = t1.names, t2.names
    return t1.contract(t2, {(a, b) for a, b in pairs if a in n1 and b in n2})


def safe_rename(t, name_map):
    names = t.names
    return t.edge_rename({k: v for k, v in name_map.items() if k in names})


# direction table of the inline strips:
#   F/B       forward / backward bond names along the sweep
#   po/pi     perpendicular bond: name on the piece already absorbed / name on the piece being absorbed
#   first     index (1 or last) of the boundary that is attached as the "tail"
_DIRS = {
    "l2r": dict(F="R", B="L", po="D", pi="U", first_is_low=True),
    "r2l": dict(F="L", B="R", po="U", pi="D", first_is_low=False),
    "u2d": dict(F="D", B="U", po="R", pi="L", first_is_low=True),
    "d2u": dict(F="U", B="D", po="L", pi="R", first_is_low=False),
}


def two_line_to_one_line(udlr, line_1, line_2, cut, normalize):
    """Absorb one lattice row/column (line_2) into a boundary MPS (line_1) and compress it to bond
    dimension ``cut``: contract site-wise, QR sweep forward, truncated-SVD sweep backward."""
    up, down, left, right = udlr
    left1, left2, right1, right2 = left + "1", left + "2", right + "1", right + "2"
    n = len(line_1)
    if n != len(line_2):
        raise ValueError("Different Length in Two Line to One Line")
    dl = [
        safe_contract(safe_rename(line_1[i], {left: left1, right: right1}), safe_rename(line_2[i], {left: left2, right: right2}),
                      {(down, up)}) for i in range(n)
    ]
    for i in range(n - 1):
        q, r = dl[i].qr("r", {name for name in (right1, right2) if name in dl[i].names}, right, left)
        dl[i] = q
        dl[i + 1] = safe_contract(dl[i + 1], r, {(left1, right1), (left2, right2)})
    for i in range(n - 1, 0, -1):
        u, s, v = dl[i].svd({left}, right, left, left, right, cut)
        if normalize:
            s /= s.norm_2()
        dl[i] = v
        dl[i - 1] = safe_contract(safe_contract(dl[i - 1], u, {(right, left)}), s, {(right, left)})
    return dl


def _strip_body(tail, *pieces_and_dir):
    """tail (x) site(s) (x) closing boundary site, each absorbed along the sweep direction"""
    *pieces, d, first_index, normalize = pieces_and_dir
    F, B, po, pi = d["F"], d["B"], d["po"], d["pi"]
    step = 1 if d["first_is_low"] else -1
    result = tail
    idx = first_index
    for piece in pieces:
        idx += step
        name = F + str(idx)
        result = safe_contract(result, safe_rename(piece, {F: name}), {(name, B), (po, pi)})
    if normalize:
        result /= result.norm_2()
    return result


def _strip_tail(strip, boundary_site, d, index):
    name = d["F"] + str(index)
    return safe_rename(safe_contract(strip, boundary_site, {(name, d["B"])}), {d["F"]: name})


def _getitem(seq, i):
    return seq[i]


def _zip(*args):
    return args


class SingleLayerAuxiliaries:
    def __init__(self, L1, L2, cut_dimension, normalize, Tensor):
        self.L1, self.L2 = L1, L2
        self.cut_dimension = cut_dimension
        self.normalize = normalize
        self.Tensor = Tensor
        one = Tensor(1)
        self._one = lazy.Root(one)
        self._one_l1 = lazy.Root([one] * L1)
        self._one_l2 = lazy.Root([one] * L2)
        self._lattice = [[lazy.Root() for _ in range(L2)] for _ in range(L1)]
        self._build_graph()

    # -- graph construction ----------------------------------------------------------------------
    def _build_graph(self):
        L1, L2, cut, norm = self.L1, self.L2, self.cut_dimension, self.normalize
        N = lazy.Node
        lat = self._lattice
        self._zip_row = [N(_zip, *lat[l1]) for l1 in range(L1)]
        self._zip_column = [N(_zip, *(lat[l1][l2] for l1 in range(L1))) for l2 in range(L2)]
        # boundary MPS (single_layer_auxiliaries.py:430-456)
        self._up_to_down = {-1: self._one_l2}
        self._down_to_up = {L1: self._one_l2}
        self._left_to_right = {-1: self._one_l1}
        self._right_to_left = {L2: self._one_l1}
        for l1 in range(L1):
            self._up_to_down[l1] = N(two_line_to_one_line, "UDLR", self._up_to_down[l1 - 1], self._zip_row[l1], cut, norm)
        for l1 in reversed(range(L1)):
            self._down_to_up[l1] = N(two_line_to_one_line, "DULR", self._down_to_up[l1 + 1], self._zip_row[l1], cut, norm)
        for l2 in range(L2):
            self._left_to_right[l2] = N(two_line_to_one_line, "LRUD", self._left_to_right[l2 - 1], self._zip_column[l2], cut, norm)
        for l2 in reversed(range(L2)):
            self._right_to_left[l2] = N(two_line_to_one_line, "RLUD", self._right_to_left[l2 + 1], self._zip_column[l2], cut, norm)
        self._up_to_down_site = {(l1, l2): N(_getitem, self._up_to_down[l1], l2) for l1 in range(-1, L1) for l2 in range(L2)}
        self._down_to_up_site = {(l1, l2): N(_getitem, self._down_to_up[l1], l2) for l1 in range(L1 + 1) for l2 in range(L2)}
        self._left_to_right_site = {(l1, l2): N(_getitem, self._left_to_right[l2], l1) for l2 in range(-1, L2) for l1 in range(L1)}
        self._right_to_left_site = {(l1, l2): N(_getitem, self._right_to_left[l2], l1) for l2 in range(L2 + 1) for l1 in range(L1)}

        # inline strips; *_tailed = strip with the next boundary site already attached
        d = _DIRS
        s = self
        s._inline_left_to_right, s._inline_left_to_right_tailed = {}, {}
        s._inline_right_to_left, s._inline_right_to_left_tailed = {}, {}
        s._inline_up_to_down, s._inline_up_to_down_tailed = {}, {}
        s._inline_down_to_up, s._inline_down_to_up_tailed = {}, {}
        s._4_inline_left_to_right, s._4_inline_left_to_right_tailed = {}, {}
        s._4_inline_right_to_left, s._4_inline_right_to_left_tailed = {}, {}
        for l1 in range(L1):
            for l2 in range(-1, L2):
                if l2 == -1:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._inline_left_to_right_tailed[l1, l2 - 1], lat[l1][l2], s._down_to_up_site[l1 + 1, l2],
                              d["l2r"], 1, norm)
                s._inline_left_to_right[l1, l2] = strip
                s._inline_left_to_right_tailed[l1, l2] = strip if l2 == L2 - 1 else N(
                    _strip_tail, strip, s._up_to_down_site[l1 - 1, l2 + 1], d["l2r"], 1)
            for l2 in reversed(range(L2 + 1)):
                if l2 == L2:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._inline_right_to_left_tailed[l1, l2 + 1], lat[l1][l2], s._up_to_down_site[l1 - 1, l2],
                              d["r2l"], 3, norm)
                s._inline_right_to_left[l1, l2] = strip
                s._inline_right_to_left_tailed[l1, l2] = strip if l2 == 0 else N(
                    _strip_tail, strip, s._down_to_up_site[l1 + 1, l2 - 1], d["r2l"], 3)
        for l2 in range(L2):
            for l1 in range(-1, L1):
                if l1 == -1:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._inline_up_to_down_tailed[l1 - 1, l2], lat[l1][l2], s._right_to_left_site[l1, l2 + 1],
                              d["u2d"], 1, norm)
                s._inline_up_to_down[l1, l2] = strip
                # NB the reference returns _inline_left_to_right[l1, l2] for the last row here
                # (single_layer_auxiliaries.py:338-340); that entry is never reached by replace/hole.
                s._inline_up_to_down_tailed[l1, l2] = strip if l1 == L1 - 1 else N(
                    _strip_tail, strip, s._left_to_right_site[l1 + 1, l2 - 1], d["u2d"], 1)
            for l1 in reversed(range(L1 + 1)):
                if l1 == L1:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._inline_down_to_up_tailed[l1 + 1, l2], lat[l1][l2], s._left_to_right_site[l1, l2 - 1],
                              d["d2u"], 3, norm)
                s._inline_down_to_up[l1, l2] = strip
                s._inline_down_to_up_tailed[l1, l2] = strip if l1 == 0 else N(
                    _strip_tail, strip, s._right_to_left_site[l1 - 1, l2 + 1], d["d2u"], 3)
        for l1 in range(L1 - 1):
            for l2 in range(-1, L2):
                if l2 == -1:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._4_inline_left_to_right_tailed[l1, l2 - 1], lat[l1][l2], lat[l1 + 1][l2],
                              s._down_to_up_site[l1 + 2, l2], d["l2r"], 1, norm)
                s._4_inline_left_to_right[l1, l2] = strip
                s._4_inline_left_to_right_tailed[l1, l2] = strip if l2 == L2 - 1 else N(
                    _strip_tail, strip, s._up_to_down_site[l1 - 1, l2 + 1], d["l2r"], 1)
            for l2 in reversed(range(L2 + 1)):
                if l2 == L2:
                    strip = self._one
                else:
                    strip = N(_strip_body, s._4_inline_right_to_left_tailed[l1, l2 + 1], lat[l1 + 1][l2], lat[l1][l2],
                              s._up_to_down_site[l1 - 1, l2], d["r2l"], 4, norm)
                s._4_inline_right_to_left[l1, l2] = strip
                s._4_inline_right_to_left_tailed[l1, l2] = strip if l2 == 0 else N(
                    _strip_tail, strip, s._down_to_up_site[l1 + 2, l2 - 1], d["r2l"], 4)

    _GRAPH_FIELDS = ("_zip_row", "_zip_column", "_up_to_down", "_down_to_up", "_left_to_right", "_right_to_left", "_up_to_down_site",
                     "_down_to_up_site", "_left_to_right_site", "_right_to_left_site", "_inline_left_to_right",
                     "_inline_left_to_right_tailed", "_inline_right_to_left", "_inline_right_to_left_tailed", "_inline_up_to_down",
                     "_inline_up_to_down_tailed", "_inline_down_to_up", "_inline_down_to_up_tailed", "_4_inline_left_to_right",
                     "_4_inline_left_to_right_tailed", "_4_inline_right_to_left", "_4_inline_right_to_left_tailed")

    def copy(self, cp=None):
        """Clone with warm caches (single_layer_auxiliaries.py:33-104)."""
        result = self.__new__(type(self))
        if cp is None:
            cp = lazy.Copy()
        result.L1, result.L2 = self.L1, self.L2
        result.cut_dimension, result.normalize, result.Tensor = self.cut_dimension, self.normalize, self.Tensor
        result._one, result._one_l1, result._one_l2 = cp(self._one), cp(self._one_l1), cp(self._one_l2)
        result._lattice = [[cp(root) for root in row] for row in self._lattice]
        # dict / list insertion order is a valid topological order of the graph
        for f in self._GRAPH_FIELDS:
            v = getattr(self, f)
            if isinstance(v, dict):
                setattr(result, f, {k: cp(n) for k, n in v.items()})
            else:
                setattr(result, f, [cp(n) for n in v])
        return result

    def __setitem__(self, l1l2, tensor):
        l1, l2 = l1l2
        self._lattice[l1][l2].reset(tensor)

    def __getitem__(self, l1l2):
        l1, l2 = l1l2
        return self._lattice[l1][l2]()

    # -- amplitude with replaced sites -------------------------------------------------------------
    def _amplitude(self, hint):
        direction, line = ("H", 0) if hint is None else hint
        if direction == "H":
            return self._inline_right_to_left[line, 0]()
        if direction == "V":
            return self._inline_down_to_up[0, line]()
        raise ValueError("Unrecognized hint")

    def _close_row(self, left, pieces, right, width):
        """left strip (tailed) x sites x right strip (tailed) for a `width`-row strip"""
        result = left
        for i, piece in enumerate(pieces):
            name = "R" + str(i + 2)
            result = safe_contract(result, safe_rename(piece, {"R": name}), {("D", "U"), (name, "L")})
        return safe_contract(result, right, {("R" + str(i), "L" + str(i)) for i in range(1, width + 3)} | {("D", "U")})

    def _single_h(self, l1, l2, t):
        left = self._inline_left_to_right_tailed[l1, l2 - 1]()
        right = self._inline_right_to_left_tailed[l1, l2 + 1]()
        result = safe_contract(left, t, {("R2", "L"), ("D", "U")})
        return safe_contract(result, right, {("R1", "L1"), ("R", "L2"), ("D", "U"), ("R3", "L3")})

    def _single_v(self, l1, l2, t):
        up = self._inline_up_to_down_tailed[l1 - 1, l2]()
        down = self._inline_down_to_up_tailed[l1 + 1, l2]()
        result = safe_contract(up, t, {("D2", "U"), ("R", "L")})
        return safe_contract(result, down, {("D1", "U1"), ("D", "U2"), ("R", "L"), ("D3", "U3")})

    def _pair_h(self, l1, l2, t0, t1):
        """3-row strip, two neighbouring columns l2, l2+1 replaced"""
        result = safe_contract(self._inline_left_to_right_tailed[l1, l2 - 1](), safe_rename(t0, {"R": "R2"}), {("D", "U"), ("R2", "L")})
        result = safe_contract(result, safe_rename(self._down_to_up_site[l1 + 1, l2](), {"R": "R3"}), {("D", "U"), ("R3", "L")})
        result = safe_contract(result, safe_rename(self._up_to_down_site[l1 - 1, l2 + 1](), {"R": "R1"}), {("R1", "L")})
        result = safe_contract(result, safe_rename(t1, {"R": "R2"}), {("D", "U"), ("R2", "L")})
        return safe_contract(result, self._inline_right_to_left_tailed[l1, l2 + 2](), {("R1", "L1"), ("R2", "L2"), ("D", "U"), ("R3", "L3")})

    def _pair_v(self, l1, l2, t0, t1):
        result = safe_contract(self._inline_up_to_down_tailed[l1 - 1, l2](), safe_rename(t0, {"D": "D2"}), {("R", "L"), ("D2", "U")})
        result = safe_contract(result, safe_rename(self._right_to_left_site[l1, l2 + 1](), {"D": "D3"}), {("R", "L"), ("D3", "U")})
        result = safe_contract(result, safe_rename(self._left_to_right_site[l1 + 1, l2 - 1](), {"D": "D1"}), {("D1", "U")})
        result = safe_contract(result, safe_rename(t1, {"D": "D2"}), {("R", "L"), ("D2", "U")})
        return safe_contract(result, self._inline_down_to_up_tailed[l1 + 2, l2](), {("D1", "U1"), ("D2", "U2"), ("R", "L"), ("D3", "U3")})

    def _block_4(self, l1, l2, t00, t10, t01, t11):
        """4-row strip over rows l1, l1+1 with columns l2, l2+1 given explicitly"""
        result = safe_contract(self._4_inline_left_to_right_tailed[l1, l2 - 1](), safe_rename(t00, {"R": "R2"}), {("D", "U"), ("R2", "L")})
        result = safe_contract(result, safe_rename(t10, {"R": "R3"}), {("D", "U"), ("R3", "L")})
        result = safe_contract(result, safe_rename(self._down_to_up_site[l1 + 2, l2](), {"R": "R4"}), {("D", "U"), ("R4", "L")})
        result = safe_contract(result, safe_rename(self._up_to_down_site[l1 - 1, l2 + 1](), {"R": "R1"}), {("R1", "L")})
        result = safe_contract(result, safe_rename(t01, {"R": "R2"}), {("D", "U"), ("R2", "L")})
        result = safe_contract(result, safe_rename(t11, {"R": "R3"}), {("D", "U"), ("R3", "L")})
        return safe_contract(result, self._4_inline_right_to_left_tailed[l1, l2 + 2](),
                             {("R1", "L1"), ("R2", "L2"), ("R3", "L3"), ("D", "U"), ("R4", "L4")})

    def replace(self, replacement, *, hint=None):
        """<s'|psi> with the site tensors in ``replacement`` substituted.  Which of the mathematically
        equivalent (but differently truncated) routes is taken depends on which environments are
        already cached, exactly as in the reference (:517-705)."""
        if len(replacement) == 0:
            return self._amplitude(hint)
        ks = list(replacement.keys())
        minl1, maxl1 = min(k[0] for k in ks), max(k[0] for k in ks)
        minl2, maxl2 = min(k[1] for k in ks), max(k[1] for k in ks)
        s = self
        if maxl1 == minl1 and maxl2 == minl2:
            l1, l2 = minl1, minl2
            t = replacement[l1, l2]
            if s._inline_left_to_right_tailed[l1, l2 - 1] and s._inline_right_to_left_tailed[l1, l2 + 1]:
                return s._single_h(l1, l2, t)
            if s._inline_up_to_down_tailed[l1 - 1, l2] and s._inline_down_to_up_tailed[l1 + 1, l2]:
                return s._single_v(l1, l2, t)
            if l1 != s.L1 - 1 and s._4_inline_left_to_right_tailed[l1, l2 - 1] and s._4_inline_right_to_left_tailed[l1, l2 + 1]:
                return s._close_row(s._4_inline_left_to_right_tailed[l1, l2 - 1](), [t, s._lattice[l1 + 1][l2]()],
                                    s._4_inline_right_to_left_tailed[l1, l2 + 1](), 2)
            if l1 != 0 and s._4_inline_left_to_right_tailed[l1 - 1, l2 - 1] and s._4_inline_right_to_left_tailed[l1 - 1, l2 + 1]:
                return s._close_row(s._4_inline_left_to_right_tailed[l1 - 1, l2 - 1](), [s._lattice[l1 - 1][l2](), t],
                                    s._4_inline_right_to_left_tailed[l1 - 1, l2 + 1](), 2)
            if hint is None or hint == "H":
                return s._single_h(l1, l2, t)
            if hint == "V":
                return s._single_v(l1, l2, t)
            raise ValueError("Unrecognized hint")
        if hint is not None:
            raise ValueError("Unrecognized hint")
        if maxl1 == minl1 and maxl2 - minl2 == 1:
            l1, l2 = minl1, minl2
            t0, t1 = replacement[l1, l2], replacement[l1, l2 + 1]
            if (s._inline_left_to_right_tailed[l1, l2 - 1] and s._down_to_up_site[l1 + 1, l2] and s._inline_right_to_left_tailed[l1, l2 + 2]
                    and s._up_to_down_site[l1 - 1, l2 + 1]):
                return s._pair_h(l1, l2, t0, t1)
            if l1 != s.L1 - 1 and (s._4_inline_left_to_right_tailed[l1, l2 - 1] and s._down_to_up_site[l1 + 2, l2]
                                   and s._4_inline_right_to_left_tailed[l1, l2 + 2] and s._up_to_down_site[l1 - 1, l2 + 1]):
                return s._block_4(l1, l2, t0, s._lattice[l1 + 1][l2](), t1, s._lattice[l1 + 1][l2 + 1]())
            if l1 != 0 and (s._4_inline_left_to_right_tailed[l1 - 1, l2 - 1] and s._down_to_up_site[l1 + 1, l2]
                            and s._4_inline_right_to_left_tailed[l1 - 1, l2 + 2] and s._up_to_down_site[l1 - 2, l2 + 1]):
                return s._block_4(l1 - 1, l2, s._lattice[l1 - 1][l2](), t0, s._lattice[l1 - 1][l2 + 1](), t1)
            return s._pair_h(l1, l2, t0, t1)
        if maxl1 - minl1 == 1 and maxl2 == minl2:
            l1, l2 = minl1, minl2
            t0, t1 = replacement[l1, l2], replacement[l1 + 1, l2]
            if (s._inline_up_to_down_tailed[l1 - 1, l2] and s._right_to_left_site[l1, l2 + 1] and s._inline_down_to_up_tailed[l1 + 2, l2]
                    and s._left_to_right_site[l1 + 1, l2 - 1]):
                return s._pair_v(l1, l2, t0, t1)
            if s._4_inline_left_to_right_tailed[l1, l2 - 1] and s._4_inline_right_to_left_tailed[l1, l2 + 1]:
                return s._close_row(s._4_inline_left_to_right_tailed[l1, l2 - 1](), [t0, t1], s._4_inline_right_to_left_tailed[l1, l2 + 1](), 2)
            return s._pair_v(l1, l2, t0, t1)
        if maxl1 - minl1 == 1 and maxl2 - minl2 == 1:
            t = [[s._lattice[minl1][minl2](), s._lattice[minl1][maxl2]()], [s._lattice[maxl1][minl2](), s._lattice[maxl1][maxl2]()]]
            for (l1, l2), tensor in replacement.items():
                t[l1 - minl1][l2 - minl2] = tensor
            return s._block_4(minl1, minl2, t[0][0], t[1][0], t[0][1], t[1][1])
        return None  # not implemented replace style (as the reference)

    # -- environments ----------------------------------------------------------------------------
    def hole(self, position, *, hint=None):
        s = self
        if len(position) == 0:
            return self._amplitude(hint)
        if len(position) == 1:
            l1, l2 = position[0]
            if hint is None or hint == "H":
                big = safe_contract(s._inline_left_to_right_tailed[l1, l2 - 1](), s._inline_right_to_left_tailed[l1, l2 + 1](),
                                    {("R1", "L1"), ("R3", "L3")})
                return safe_rename(big, {"R2": "R0", "L2": "L0", "U": "U0", "D": "D0"})
            if hint == "V":
                big = safe_contract(s._inline_up_to_down_tailed[l1 - 1, l2](), s._inline_down_to_up_tailed[l1 + 1, l2](),
                                    {("U1", "D1"), ("U3", "D3")})
                return safe_rename(big, {"U2": "U0", "D2": "D0", "R": "R0", "L": "L0"})
            raise ValueError("Unrecognized hint")
        if len(position) == 2:
            if hint is not None:
                raise ValueError("Unrecognized hint")
            p0, p1 = position
            if p0[0] == p1[0] and abs(p0[1] - p1[1]) == 1:
                a, b = ("0", "1") if p0[1] < p1[1] else ("1", "0")   # tag of the left / right site
                pl, pr = (p0, p1) if p0[1] < p1[1] else (p1, p0)
                r = safe_contract(s._inline_left_to_right_tailed[pl[0], pl[1] - 1](),
                                  safe_rename(s._down_to_up_site[pl[0] + 1, pl[1]](), {"R": "R3"}), {("R3", "L")})
                r = safe_rename(r, {"D": "D" + a, "U": "U" + a})
                r = safe_rename(safe_contract(r, safe_rename(s._up_to_down_site[pr[0] - 1, pr[1]](), {"R": "R1"}), {("R1", "L")}), {"D": "D" + b})
                r = safe_contract(r, s._inline_right_to_left_tailed[pr[0], pr[1] + 1](), {("R1", "L1"), ("R3", "L3")})
                return safe_rename(r, {"U": "U" + b, "R2": "R" + a, "L2": "L" + b})
            if p0[1] == p1[1] and abs(p0[0] - p1[0]) == 1:
                a, b = ("0", "1") if p0[0] < p1[0] else ("1", "0")   # tag of the upper / lower site
                pu, pd = (p0, p1) if p0[0] < p1[0] else (p1, p0)
                r = safe_contract(s._inline_up_to_down_tailed[pu[0] - 1, pu[1]](),
                                  safe_rename(s._right_to_left_site[pu[0], pu[1] + 1](), {"D": "D3"}), {("D3", "U")})
                r = safe_rename(r, {"R": "R" + a, "L": "L" + a})
                r = safe_rename(safe_contract(r, safe_rename(s._left_to_right_site[pd[0], pd[1] - 1](), {"D": "D1"}), {("D1", "U")}), {"R": "R" + b})
                r = safe_contract(r, s._inline_down_to_up_tailed[pd[0] + 1, pd[1]](), {("D1", "U1"), ("D3", "U3")})
                return safe_rename(r, {"L": "L" + b, "D2": "D" + a, "U2": "U" + b})
        raise NotImplementedError("Unsupported auxilary hole style")
