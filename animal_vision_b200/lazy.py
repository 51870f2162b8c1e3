"""Lazy element-wise expressions over device planes, compiled into K7 register programs (csrc/k7_pointwise.cu).

The UV species of the reference are written as long runs of NumPy element-wise statements; here the same statements
build an expression DAG (`E` nodes) on the host and `Lazy.run` executes every output of a stage in ONE kernel launch.
Each node is one IEEE float32 operation, evaluated in the order the species code wrote it (no folding, no
re-association), which is what NumPy does with float32 arrays and Python-float scalars (NEP 50: the scalar adopts
float32).  Reduction results (min / max / percentiles) enter as per-frame DEVICE scalars, so a stage never waits
for the host.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import List, Sequence

import numpy as np

from ._abi import AvbError, check

OPS = {name: i for i, name in enumerate(
    ["NOP", "LOAD", "CONST", "MOV", "ADD", "SUB", "MUL", "DIV", "MIN", "MAX", "POW", "ATAN2", "GT", "GE", "LT", "LE", "NEG",
     "ABS", "SQRT", "EXP", "SIN", "COS", "FLOOR", "SRGB_DEC", "SRGB_ENC", "QUANT", "SELECT", "STORE",
     "ADDI", "SUBI", "RSUBI", "MULI", "DIVI", "RDIVI", "MINI", "MAXI", "POWI", "GTI", "GEI", "LTI", "LEI"])}   # enum in include/avb200.h
# (op, constant on the right) / (op, constant on the left) -> immediate form; comparisons flip when the constant is on the left
_IMM_RIGHT = {"ADD": "ADDI", "SUB": "SUBI", "MUL": "MULI", "DIV": "DIVI", "MIN": "MINI", "MAX": "MAXI", "POW": "POWI",
              "GT": "GTI", "GE": "GEI", "LT": "LTI", "LE": "LEI"}
_IMM_LEFT = {"ADD": "ADDI", "SUB": "RSUBI", "MUL": "MULI", "DIV": "RDIVI", "MIN": "MINI", "MAX": "MAXI",
             "GT": "LTI", "GE": "LEI", "LT": "GTI", "LE": "GEI"}
SRC_PLANE, SRC_ROW, SRC_COL, SRC_FRAME = 0, 1, 2, 3
MAX_SRC, MAX_DST, MAX_REGS, MAX_INS = 24, 4, 48, 2048


class VmSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("frame_stride", C.c_int64), ("pix_stride", C.c_int32), ("kind", C.c_int32)]


class VmDst(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("frame_stride", C.c_int64), ("row_stride", C.c_int64), ("pix_stride", C.c_int32), ("kind", C.c_int32)]


class Source:
    """A device buffer an expression can load from."""
    __slots__ = ("tensor", "kind", "frame_stride", "pix_stride")

    def __init__(self, tensor, kind, frame_stride, pix_stride):
        self.tensor, self.kind, self.frame_stride, self.pix_stride = tensor, kind, int(frame_stride), int(pix_stride)

    def key(self):
        return (self.tensor.data_ptr(), self.kind, self.frame_stride, self.pix_stride)


def _e(v) -> "E":
    return v if isinstance(v, E) else E("CONST", imm=np.float32(v))


class E:
    """One float32 value per pixel."""
    __slots__ = ("op", "args", "imm", "src", "ch")

    def __init__(self, op, args=(), imm=np.float32(0), src=None, ch=0):
        self.op, self.args, self.imm, self.src, self.ch = op, tuple(args), imm, src, ch

    def __add__(self, o): return E("ADD", (self, _e(o)))
    def __radd__(self, o): return E("ADD", (_e(o), self))
    def __sub__(self, o): return E("SUB", (self, _e(o)))
    def __rsub__(self, o): return E("SUB", (_e(o), self))
    def __mul__(self, o): return E("MUL", (self, _e(o)))
    def __rmul__(self, o): return E("MUL", (_e(o), self))
    def __truediv__(self, o): return E("DIV", (self, _e(o)))
    def __rtruediv__(self, o): return E("DIV", (_e(o), self))
    def __neg__(self): return E("NEG", (self,))
    def __gt__(self, o): return E("GT", (self, _e(o)))
    def __ge__(self, o): return E("GE", (self, _e(o)))
    def __lt__(self, o): return E("LT", (self, _e(o)))
    def __le__(self, o): return E("LE", (self, _e(o)))

    def __pow__(self, o):
        if not isinstance(o, E) and float(o) == 2.0:
            return E("MUL", (self, self))                 # numpy: x ** 2 is np.square
        return E("POW", (self, _e(o)))


def minimum(a, b): return E("MIN", (_e(a), _e(b)))
def maximum(a, b): return E("MAX", (_e(a), _e(b)))


def clip(x, lo, hi):
    """np.clip(x, lo, hi); either bound may be None."""
    x = _e(x)
    if lo is not None:
        x = maximum(x, lo)
    if hi is not None:
        x = minimum(x, hi)
    return x


def sqrt(x): return E("SQRT", (_e(x),))
def exp(x): return E("EXP", (_e(x),))
def sin(x): return E("SIN", (_e(x),))
def cos(x): return E("COS", (_e(x),))
def absolute(x): return E("ABS", (_e(x),))
def floor(x): return E("FLOOR", (_e(x),))
def arctan2(y, x): return E("ATAN2", (_e(y), _e(x)))
def where(c, a, b): return E("SELECT", (_e(c), _e(a), _e(b)))
def srgb_to_linear(x): return E("SRGB_DEC", (_e(x),))          # uv_helpers.py:33-37
def linear_to_srgb(x): return E("SRGB_ENC", (_e(x),))          # uv_helpers.py:40-44
def quantize(x): return E("QUANT", (_e(x),))                   # uv_helpers.py:26-30, integer dtypes


def luma(rgb: Sequence[E]) -> E:
    """0.2126 R + 0.7152 G + 0.0722 B in NumPy's left-to-right order."""
    return 0.2126 * rgb[0] + 0.7152 * rgb[1] + 0.0722 * rgb[2]


def _imm_form(nd):
    """(immediate op, register operand, constant) when one operand of a binary node is a constant, else None.
    IEEE add / mul / min / max are commutative, so `c + x` may issue as `x + c` with identical bits."""
    if len(nd.args) != 2:
        return None
    a, b = nd.args
    if b.op == "CONST" and nd.op in _IMM_RIGHT and a.op != "CONST":
        return _IMM_RIGHT[nd.op], a, b.imm
    if a.op == "CONST" and nd.op in _IMM_LEFT and b.op != "CONST":
        return _IMM_LEFT[nd.op], b, a.imm
    return None


def _compile(stores):
    """stores: [(E, dst_index, channel)] -> (instruction array uint32[n,2], n_regs, sources)."""
    roots = {id(r) for r, _, _ in stores}
    order, state = [], {}
    for root, _, _ in stores:                      # iterative post-order over the DAG (shared nodes once)
        stack = [(root, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0:
                if id(node) in state:
                    continue
                state[id(node)] = 1
            imm = _imm_form(node)
            kids = (imm[1],) if imm else node.args      # a constant folded into an immediate is never materialised
            if i < len(kids):
                stack.append((node, i + 1))
                child = kids[i]
                if id(child) not in state:
                    stack.append((child, 0))
            else:
                order.append(node)
    pos = {id(nd): k for k, nd in enumerate(order)}
    last = {id(nd): -1 for nd in order}
    for k, nd in enumerate(order):
        imm = _imm_form(nd)
        for a in ((imm[1],) if imm else nd.args):
            last[id(a)] = max(last[id(a)], k)
    store_at = {}
    for root, d, ch in stores:
        store_at.setdefault(pos[id(root)], []).append((d, ch))
    sources, src_index = [], {}
    free, n_regs, reg = [], 0, {}
    ins = []

    def emit(op, dst=0, a=0, b=0, imm=0):
        ins.append((OPS[op] | (dst << 8) | (a << 16) | (b << 24), int(imm) & 0xffffffff))

    def fbits(v):
        return int(np.float32(v).view(np.uint32))

    for k, nd in enumerate(order):
        imm = _imm_form(nd)
        operands = (imm[1],) if imm else nd.args
        args = [reg[id(a)] for a in operands]
        # operands whose last consumer this is: their registers are free again -- also for this very result, since an
        # instruction reads its operands before it writes
        for aid in {id(x) for x in operands}:
            if last[aid] == k:
                free.append(reg[aid])
        if free:
            r = free.pop()
        else:
            r = n_regs
            n_regs += 1
        reg[id(nd)] = r
        if nd.op == "LOAD":
            key = nd.src.key()
            if key not in src_index:
                src_index[key] = len(sources)
                sources.append(nd.src)
            emit("LOAD", r, src_index[key], nd.ch)
        elif nd.op == "CONST":
            emit("CONST", r, imm=fbits(nd.imm))
        elif imm:
            emit(imm[0], r, args[0], imm=fbits(imm[2]))
        elif nd.op == "SELECT":
            emit("SELECT", r, args[0], args[1], args[2])
        elif len(args) == 1:
            emit(nd.op, r, args[0])
        else:
            emit(nd.op, r, args[0], args[1])
        for d, ch in store_at.get(k, ()):
            emit("STORE", 0, r, d, ch)
        if last[id(nd)] == -1:                      # nothing reads it later (a stored root, or dead code)
            free.append(r)
    if n_regs > MAX_REGS:
        raise AvbError(f"element-wise stage needs {n_regs} registers (limit {MAX_REGS}): split the stage")
    if len(ins) > MAX_INS:
        raise AvbError(f"element-wise stage has {len(ins)} instructions (limit {MAX_INS}): split the stage")
    if len(sources) > MAX_SRC:
        raise AvbError(f"element-wise stage reads {len(sources)} buffers (limit {MAX_SRC}): split the stage")
    return np.array(ins, dtype=np.uint32), max(n_regs, 1), sources


class Lazy:
    """Binds expressions to a batch geometry [n, H, W] on one engine."""

    def __init__(self, eng, n: int, H: int, W: int):
        self.eng, self.t, self.n, self.H, self.W = eng, eng.torch, int(n), int(H), int(W)
        self.launches = 0

    # ---- leaves
    def plane(self, tensor, ch: int = 0) -> E:
        """Channel `ch` of a float32 CUDA tensor [n or 1, H, W, C] (a single frame broadcasts over the batch)."""
        t = self.t
        if not (tensor.is_cuda and tensor.dtype == t.float32 and tensor.dim() == 4 and tensor.is_contiguous()):
            raise AvbError("plane: expected a contiguous CUDA float32 tensor [n,H,W,C]")
        nn, H, W, Cn = tensor.shape
        if (H, W) != (self.H, self.W) or nn not in (1, self.n):
            raise AvbError(f"plane: shape {tuple(tensor.shape)} does not match the batch [{self.n},{self.H},{self.W}]")
        fs = 0 if (nn == 1 and self.n > 1) else H * W * Cn
        return E("LOAD", src=Source(tensor, SRC_PLANE, fs, Cn), ch=ch)

    def channels(self, tensor) -> List[E]:
        return [self.plane(tensor, c) for c in range(tensor.shape[3])]

    def _table(self, arr: np.ndarray, kind: int, length: int) -> E:
        a = np.ascontiguousarray(arr, np.float32).reshape(-1)
        if a.size != length:
            raise AvbError(f"table of {a.size} values, expected {length}")
        dev = self.eng.cached(("lazy_tab", kind, hashlib.sha1(a.tobytes()).hexdigest()), lambda: self.eng._dev(a))
        return E("LOAD", src=Source(dev, kind, 0, 1), ch=0)

    def row(self, arr) -> E:
        """A value per image row ((H,1) arrays of the reference)."""
        return self._table(arr, SRC_ROW, self.H)

    def col(self, arr) -> E:
        return self._table(arr, SRC_COL, self.W)

    def table(self, arr) -> E:
        """A pixel-independent (H,W) plane computed on the host with the reference's own NumPy expression (radial
        masks, seams, attention spots); uploaded once per content."""
        a = np.ascontiguousarray(arr, np.float32)
        if a.shape != (self.H, self.W):
            raise AvbError(f"table shape {a.shape}, expected {(self.H, self.W)}")
        dev = self.eng.cached(("lazy_tab2", hashlib.sha1(a.tobytes()).hexdigest()), lambda: self.eng._dev(a.reshape(1, self.H, self.W, 1)))
        return E("LOAD", src=Source(dev, SRC_PLANE, 0, 1), ch=0)

    def keyed(self, key, build, kind: str = "table") -> E:
        """A pixel-independent table that is a pure function of `key` (species parameters + geometry): `build()` runs on the
        host only the first time the key is seen on this device -- a video stream never recomputes its masks.
        kind: "table" (H,W), "row" (H) or "col" (W)."""
        full = ("lazy_keyed", kind, self.H, self.W, key)
        dev = self.eng._cache.get(full)
        if dev is None:
            a = np.ascontiguousarray(build(), np.float32)
            shape = {"table": (self.H, self.W), "row": (self.H,), "col": (self.W,)}[kind]
            if a.size != int(np.prod(shape)):
                raise AvbError(f"keyed {kind} has {a.size} values, expected {shape}")
            dev = self.eng._cache[full] = self.eng._dev(a.reshape((1, self.H, self.W, 1)) if kind == "table" else a.reshape(-1))
        if kind == "table":
            return E("LOAD", src=Source(dev, SRC_PLANE, 0, 1), ch=0)
        return E("LOAD", src=Source(dev, SRC_ROW if kind == "row" else SRC_COL, 0, 1), ch=0)

    def scalar(self, tensor, index: int = 0) -> E:
        """A per-frame device scalar: tensor is float32 [n, k] (contiguous), value tensor[frame, index]."""
        t = self.t
        if not (tensor.is_cuda and tensor.dtype == t.float32 and tensor.is_contiguous() and tensor.shape[0] == self.n):
            raise AvbError("scalar: expected a contiguous CUDA float32 tensor [n, k]")
        k = int(tensor.numel() // self.n)
        return E("LOAD", src=Source(tensor, SRC_FRAME, k, 1), ch=int(index))

    # ---- execution
    def run(self, outputs):
        """outputs: [(expression list, destination tensor)]; destination float32 [n,H,W,len] contiguous or uint8
        [n,H,W,len] (strided rows / frames allowed).  One launch for all of them."""
        t = self.t
        if len(outputs) > MAX_DST:
            raise AvbError("too many destinations in one stage")
        stores, dsts = [], (VmDst * len(outputs))()
        for d, (exprs, out) in enumerate(outputs):
            if tuple(out.shape) != (self.n, self.H, self.W, len(exprs)) or not out.is_cuda:
                raise AvbError(f"destination shape {tuple(out.shape)} != {(self.n, self.H, self.W, len(exprs))}")
            if out.dtype == t.float32:
                kind = 0
            elif out.dtype == t.uint8:
                kind = 1
            else:
                raise AvbError("destination must be float32 or uint8")
            fs, rs, ps, cs = out.stride()
            if cs != 1:
                raise AvbError("destination channels must be contiguous")
            dsts[d] = VmDst(out.data_ptr(), fs, rs, ps, kind)
            for ch, e in enumerate(exprs):
                stores.append((_e(e), d, ch))
        ins, n_regs, sources = _compile(stores)
        prog = self.eng.cached(("lazy_prog", hashlib.sha1(ins.tobytes()).hexdigest()), lambda: self.eng._dev(ins.view(np.int32)))
        srcs = (VmSrc * max(len(sources), 1))()
        for i, s in enumerate(sources):
            srcs[i] = VmSrc(s.tensor.data_ptr(), s.frame_stride, s.pix_stride, s.kind)
        with t.cuda.device(self.eng.device):
            rc = self.eng.lib.avb_vm_run(prog.data_ptr(), int(ins.shape[0]), int(n_regs), self.n, self.H, self.W,
                                         C.cast(srcs, C.c_void_p), len(sources), C.cast(dsts, C.c_void_p), len(outputs),
                                         self.eng.stream_ptr())
        check(rc, "avb_vm_run")
        self.eng.launches += 1
        self.launches += 1
        self._keep = sources           # the source tensors stay referenced until the next stage is enqueued

    def eval(self, exprs: Sequence[E]):
        out = self.t.empty((self.n, self.H, self.W, len(exprs)), dtype=self.t.float32, device=self.eng.device)
        self.run([(list(exprs), out)])
        return out
