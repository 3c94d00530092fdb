"""Minimal pure-Python reader for R's RDX3 (.rda / save()) serialisation format.

Test infrastructure only: used by tools/make_golden.py to turn the reference's
bundled datasets (data/*.rda under the reference checkout) into the portable
fixtures committed under tests/golden/.  No R is available in this image.

Format notes (XDR / big-endian, serialisation version 3):
  file  = xz | gzip | bzip2 stream wrapping  b"RDX3\n" b"X\n" int32 version,
          int32 writer-version, int32 min-reader-version, int32 nchar, native
          encoding name, then ONE item (a pairlist tagged with object names).
  item  = int32 flags; type = flags & 0xFF, is_obj = bit 8, has_attr = bit 9,
          has_tag = bit 10; payload by type; attributes follow vector payloads.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from typing import Any

import numpy as np

NA_INT = -2147483648


class RObj:
    """An R value with attributes. `value` is a numpy array / list / str / dict."""

    __slots__ = ("type", "value", "attr")

    def __init__(self, type_: str, value: Any, attr: dict | None = None):
        self.type = type_
        self.value = value
        self.attr = attr or {}

    # convenience ---------------------------------------------------------
    def names(self):
        n = self.attr.get("names")
        return None if n is None else list(n.value)

    def __getitem__(self, key):
        if isinstance(key, str):
            if self.type == "S4":
                return self.attr[key]
            if self.type == "pairlist":
                return self.value[key]
            nm = self.names()
            return self.value[nm.index(key)]
        return self.value[key]

    def keys(self):
        if self.type == "S4":
            return list(self.attr.keys())
        if self.type == "pairlist":
            return list(self.value.keys())
        return self.names()

    def __repr__(self):
        v = self.value
        if isinstance(v, np.ndarray):
            d = f"array{v.shape}"
        elif isinstance(v, list):
            d = f"list[{len(v)}]"
        else:
            d = repr(v)[:40]
        return f"RObj({self.type}, {d}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list[Any] = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def attributes(self) -> dict:
        a = self.item()
        if a is None:
            return {}
        assert a.type == "pairlist", a
        return a.value

    def item(self):  # noqa: C901 - a type switch
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:  # NILVALUE
            return None
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 253:  # global env
            return RObj("env", "R_GlobalEnv")
        if t == 242:  # empty env
            return RObj("env", "R_EmptyEnv")
        if t == 241:  # base env
            return RObj("env", "R_BaseEnv")
        if t == 251 or t == 252:  # missing / unbound
            return RObj("special", t)
        if t in (249, 250, 247):  # package / namespace / persistent refs
            s = self.item()
            o = RObj("nsref", s)
            self.refs.append(o)
            return o
        if t == 1:  # SYMSXP
            ch = self.item()
            o = RObj("symbol", ch)
            self.refs.append(o)
            return o
        if t == 4:  # ENVSXP
            o = RObj("env", {})
            self.refs.append(o)
            self.i32()  # locked
            encl, frame, hashtab, attr = self.item(), self.item(), self.item(), self.item()
            o.value = {"enclos": encl, "frame": frame, "hashtab": hashtab}
            if attr is not None:
                o.attr = attr.value
            return o
        if t in (2, 6, 5, 17):  # LISTSXP, LANGSXP, PROMSXP, DOTSXP
            out: dict = {}
            attr = {}
            k = 0
            # iterative over cdr to avoid deep recursion
            while True:
                if has_attr:
                    attr.update(self.attributes())
                tag = None
                if has_tag:
                    tg = self.item()
                    tag = tg.value if isinstance(tg, RObj) else tg
                car = self.item()
                out[tag if tag is not None else f"_{k}"] = car
                k += 1
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6, 5, 17):
                    # dotted tail: rewind and read as a generic item
                    self.p -= 4
                    out[f"_cdr{k}"] = self.item()
                    break
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return RObj("pairlist", out, attr)
        if t == 3:  # CLOSXP
            attr = self.attributes() if has_attr else {}
            env, formals, body = self.item(), self.item(), self.item()
            return RObj("closure", (env, formals, body), attr)
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            return self.raw(n).decode("utf-8", "replace")
        if t == 10 or t == 13:  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            o = RObj("logical" if t == 10 else "integer", v)
        elif t == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            o = RObj("double", v)
        elif t == 15:  # CPLXSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">c16", count=n, offset=self.p).astype(np.complex128)
            self.p += 16 * n
            o = RObj("complex", v)
        elif t == 16:  # STRSXP
            n = self.length()
            o = RObj("character", [self.item() for _ in range(n)])
        elif t == 19 or t == 20:  # VECSXP / EXPRSXP
            n = self.length()
            o = RObj("list", [self.item() for _ in range(n)])
        elif t == 24:  # RAWSXP
            n = self.length()
            o = RObj("raw", self.raw(n))
        elif t == 25:  # S4SXP
            o = RObj("S4", None)
        elif t == 8 or t == 7:  # BUILTINSXP / SPECIALSXP
            n = self.i32()
            return RObj("builtin", self.raw(n).decode())
        elif t == 21:  # BCODESXP - not needed; bail out loudly
            raise NotImplementedError("bytecode objects are not supported")
        elif t == 22:  # EXTPTRSXP
            o = RObj("extptr", None)
            self.refs.append(o)
            prot, tag = self.item(), self.item()
            o.value = (prot, tag)
        elif t == 238:  # ALTREP
            info, state, attr = self.item(), self.item(), self.item()
            cls = list(info.value.values())[0].value  # class symbol -> CHARSXP string
            o = self._altrep(cls, state)
            if attr is not None:
                o.attr.update(attr.value)
            return o
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.p}")
        if has_attr:
            o.attr.update(self.attributes())
        return o

    @staticmethod
    def _altrep(cls: str, state) -> RObj:
        if cls == "compact_intseq":
            n, start, step = (float(x) for x in state.value)
            return RObj("integer", (start + step * np.arange(int(n))).astype(np.int32))
        if cls == "compact_realseq":
            n, start, step = (float(x) for x in state.value)
            return RObj("double", start + step * np.arange(int(n), dtype=np.float64))
        if cls.startswith("wrap_"):
            inner = state.value[0] if state.type == "list" else list(state.value.values())[0]
            return RObj(inner.type, inner.value, dict(inner.attr))
        if cls == "deferred_string":
            # state = pairlist(arg, scipen) or already-expanded STRSXP
            if state.type == "character":
                return state
            arg = list(state.value.values())[0]
            vals = arg.value
            return RObj("character", [_fmt_num(v) for v in vals])
        raise NotImplementedError(f"ALTREP class {cls}")


def _fmt_num(v) -> str:
    if isinstance(v, (np.integer, int)):
        return str(int(v))
    return repr(float(v)) if float(v) != int(v) else str(int(v))


def _decompress(raw: bytes) -> bytes:
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


def read_rda(path: str) -> dict[str, Any]:
    """Return {object name: RObj} for every object stored by save() in `path`."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    if buf[:5] != b"RDX3\n":
        raise ValueError(f"{path}: not an RDX3 stream ({buf[:5]!r})")
    r = _Reader(buf)
    r.p = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version, _w, _m = r.i32(), r.i32(), r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)  # native encoding
    top = r.item()
    return dict(top.value)


def data_frame(o: RObj) -> dict[str, Any]:
    """VECSXP with names -> {column: numpy array or list of str}."""
    out = {}
    for nm, col in zip(o.names(), o.value):
        out[nm] = col.value if col is not None else None
    return out
