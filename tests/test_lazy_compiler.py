"""CPU tests of the lazy-expression compiler (animal_vision_b200/lazy.py): the register program it emits is run by a
NumPy interpreter of the K7 instruction set and compared with direct NumPy evaluation of the same expression."""
import numpy as np

from animal_vision_b200 import lazy as L

NAMES = {v: k for k, v in L.OPS.items()}


class _FakeTensor:
    def __init__(self, arr):
        self.arr = arr

    def data_ptr(self):
        return id(self.arr)


def _interp(ins, n_regs, sources, npx, n_dst_ch):
    r = np.zeros((n_regs, npx), np.float32)
    out = {}
    for w0, w1 in ins:
        op, dst, a, b = NAMES[int(w0) & 255], (int(w0) >> 8) & 255, (int(w0) >> 16) & 255, int(w0) >> 24
        imm = int(w1)
        if op == "LOAD":
            v = sources[a].tensor.arr[:, b]
        elif op == "CONST":
            v = np.full(npx, np.uint32(imm).view(np.float32))
        elif op == "STORE":
            out[(b, imm)] = r[a].copy()
            continue
        elif op == "SELECT":
            v = np.where(r[a] != 0, r[b], r[imm & 255])
        elif op.endswith("I") and op[:-1] in ("ADD", "SUB", "RSUB", "MUL", "DIV", "RDIV", "MIN", "MAX", "POW", "GT", "GE", "LT", "LE"):
            x, m = r[a], np.float32(np.uint32(imm).view(np.float32))
            with np.errstate(all="ignore"):
                v = {"ADDI": lambda: x + m, "SUBI": lambda: x - m, "RSUBI": lambda: m - x, "MULI": lambda: x * m, "DIVI": lambda: x / m,
                     "RDIVI": lambda: m / x, "MINI": lambda: np.minimum(x, m), "MAXI": lambda: np.maximum(x, m), "POWI": lambda: np.power(x, m),
                     "GTI": lambda: (x > m).astype(np.float32), "GEI": lambda: (x >= m).astype(np.float32),
                     "LTI": lambda: (x < m).astype(np.float32), "LEI": lambda: (x <= m).astype(np.float32)}[op]()
        else:
            x, y = r[a], r[b]
            with np.errstate(all="ignore"):
                v = {"ADD": lambda: x + y, "SUB": lambda: x - y, "MUL": lambda: x * y, "DIV": lambda: x / y, "MIN": lambda: np.minimum(x, y),
                     "MAX": lambda: np.maximum(x, y), "POW": lambda: np.power(x, y), "ATAN2": lambda: np.arctan2(x, y),
                     "GT": lambda: (x > y).astype(np.float32), "GE": lambda: (x >= y).astype(np.float32),
                     "LT": lambda: (x < y).astype(np.float32), "LE": lambda: (x <= y).astype(np.float32), "NEG": lambda: -x,
                     "ABS": lambda: np.abs(x), "SQRT": lambda: np.sqrt(x), "EXP": lambda: np.exp(x), "SIN": lambda: np.sin(x),
                     "COS": lambda: np.cos(x), "FLOOR": lambda: np.floor(x), "MOV": lambda: x}[op]()
        r[dst] = v.astype(np.float32)
    return out


def test_program_equals_numpy_and_registers_are_reused():
    rng = np.random.default_rng(0)
    npx = 257
    img = rng.random((npx, 3), dtype=np.float32)
    src = L.Source(_FakeTensor(img), L.SRC_PLANE, 0, 3)
    x, y, z = (L.E("LOAD", src=src, ch=c) for c in range(3))
    e0 = L.clip((x * 0.5 + y) ** 2 - L.sqrt(x) / (y + 1e-8), 0.0, 1.0)
    e1 = L.where(x > y, e0, x * y) + L.luma([x, y, z])
    chain = x
    for k in range(40):                                  # a long dependent chain must not grow the register file
        chain = L.minimum(chain * 1.01 + (y if k % 2 else z), 2.0)
    ins, n_regs, sources = L._compile([(e0, 0, 0), (e1, 0, 1), (chain, 1, 0), (x, 1, 1)])
    assert len(sources) == 1 and n_regs <= 8
    got = _interp(ins, n_regs, sources, npx, 2)
    X, Y, Z = img[:, 0], img[:, 1], img[:, 2]
    f = np.float32
    r0 = np.clip((X * f(0.5) + Y) ** 2 - np.sqrt(X) / (Y + f(1e-8)), 0, 1)
    r1 = np.where(X > Y, r0, X * Y) + (f(0.2126) * X + f(0.7152) * Y + f(0.0722) * Z)
    c = X
    for k in range(40):
        c = np.minimum(c * f(1.01) + (Y if k % 2 else Z), f(2.0))
    assert np.array_equal(got[(0, 0)], r0) and np.array_equal(got[(0, 1)], r1)
    assert np.array_equal(got[(1, 0)], c) and np.array_equal(got[(1, 1)], X)
    names = [NAMES[int(w0) & 255] for w0, _ in ins]
    assert "CONST" not in names and "MULI" in names and "MINI" in names          # constants ride in the instruction word


def test_constants_on_either_side_and_comparisons():
    img = np.random.default_rng(1).random((129, 2), dtype=np.float32) * 2 - 1
    src = L.Source(_FakeTensor(img), L.SRC_PLANE, 0, 2)
    x, y = (L.E("LOAD", src=src, ch=c) for c in range(2))
    exprs = [1.0 - x, 2.0 / (y + 3.0), L.where(0.25 < x, x - 0.5, 0.5 - y), L.maximum(0.1, x) ** 1.3, L.where(x >= 0.0, 1.0, L.where(0.0 > y, 2.0, 3.0)),
             L._e(0.75) + L._e(0.5) * x]
    ins, n_regs, sources = L._compile([(e, 0, i) for i, e in enumerate(exprs)])
    got = _interp(ins, n_regs, sources, 129, len(exprs))
    X, Y, f = img[:, 0], img[:, 1], np.float32
    ref = [f(1.0) - X, f(2.0) / (Y + f(3.0)), np.where(f(0.25) < X, X - f(0.5), f(0.5) - Y), np.power(np.maximum(f(0.1), X), f(1.3)),
           np.where(X >= 0, f(1.0), np.where(f(0.0) > Y, f(2.0), f(3.0))), f(0.75) + f(0.5) * X]
    for i, r in enumerate(ref):
        assert np.array_equal(got[(0, i)], r.astype(np.float32)), i


def test_limits_raise():
    import pytest
    from animal_vision_b200._abi import AvbError
    src = L.Source(_FakeTensor(np.zeros((4, 1), np.float32)), L.SRC_PLANE, 0, 1)
    x = L.E("LOAD", src=src, ch=0)

    def stage(k):
        vals = [L.exp(x * float(i)) for i in range(k)]   # all k values are stored first and summed afterwards: all alive at once
        tot = vals[0]
        for v in vals[1:]:
            tot = tot + v
        return [(v, 0, 0) for v in vals] + [(tot, 0, 1)]
    ins, n_regs, _ = L._compile(stage(20))
    assert 20 <= n_regs <= L.MAX_REGS
    with pytest.raises(AvbError):
        L._compile(stage(60))
