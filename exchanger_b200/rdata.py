"""Minimal reader for R ``save()`` files (RDX2, XDR serialisation).

Only what the fixtures ``RLdata500.rda`` / ``RLdata10000.rda`` need
(reference: /root/reference/data/*.rda, documented in R/data.R:1-60):
pairlists, symbols, character/integer/real/logical/string/list vectors,
attributes and back-references.  It exists because neither R nor pyreadr is
available here; it is a test-data loader, not part of the sampler.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from typing import Any

NA_INT = -2147483648


class RObject:
    """A decoded R value: python payload plus its attribute dict."""

    __slots__ = ("value", "attrs")

    def __init__(self, value: Any, attrs: dict | None = None):
        self.value = value
        self.attrs = attrs or {}

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"RObject({type(self.value).__name__}, attrs={list(self.attrs)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.refs: list[Any] = []

    def _int(self) -> int:
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def _double(self) -> float:
        v = struct.unpack_from(">d", self.buf, self.pos)[0]
        self.pos += 8
        return v

    def _bytes(self, n: int) -> bytes:
        v = self.buf[self.pos:self.pos + n]
        self.pos += n
        return v

    def _length(self) -> int:
        n = self._int()
        if n == -1:  # long vector
            hi, lo = self._int(), self._int()
            n = (hi << 32) + lo
        return n

    def item(self) -> Any:
        flags = self._int()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:      # NILVALUE_SXP
            return None
        if typ == 253:      # GLOBALENV_SXP
            return "<globalenv>"
        if typ == 255:      # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self._int()
            return self.refs[idx - 1]
        if typ == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:        # LISTSXP (pairlist): returned as list of (tag, value)
            out = []
            attrs = None
            while True:
                if has_attr:
                    attrs = self._attrs()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self._int()
                typ = flags & 0xFF
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
                if typ == 254:
                    break
                if typ != 2:
                    raise ValueError(f"malformed pairlist (cdr type {typ})")
            return out
        if typ == 9:        # CHARSXP
            n = self._int()
            if n == -1:
                return None  # NA_character_
            return self._bytes(n).decode("utf-8", errors="replace")
        if typ in (10, 13):  # LGLSXP / INTSXP
            n = self._length()
            vals = list(struct.unpack_from(f">{n}i", self.buf, self.pos))
            self.pos += 4 * n
            obj = RObject(vals)
        elif typ == 14:     # REALSXP
            n = self._length()
            vals = list(struct.unpack_from(f">{n}d", self.buf, self.pos))
            self.pos += 8 * n
            obj = RObject(vals)
        elif typ == 16:     # STRSXP
            n = self._length()
            obj = RObject([self.item() for _ in range(n)])
        elif typ == 19:     # VECSXP
            n = self._length()
            obj = RObject([self.item() for _ in range(n)])
        else:
            raise NotImplementedError(f"R SEXP type {typ} not supported")
        if has_attr:
            obj.attrs = self._attrs()
        return obj

    def _attrs(self) -> dict:
        pl = self.item()
        return {tag: val for tag, val in (pl or [])}


def read_rda(path: str) -> dict[str, Any]:
    """Return ``{name: RObject}`` for every object stored in an .rda file."""
    raw = open(path, "rb").read()
    if raw[:2] == b"BZ":
        raw = bz2.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 file")
    if raw[5:7] != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    rd = _Reader(raw)
    rd.pos = 7
    version = rd._int()
    rd._int()  # writer version
    rd._int()  # min reader version
    if version != 2:
        raise ValueError(f"serialisation version {version} not supported")
    top = rd.item()
    return {tag: val for tag, val in top}


def data_frame_columns(df: RObject) -> dict[str, list]:
    """Columns of a data.frame RObject as python lists.

    Factor columns are expanded to their character levels (``None`` for NA);
    integer/real columns keep numbers with ``None`` for NA.
    """
    names = df.attrs["names"].value
    out: dict[str, list] = {}
    for name, col in zip(names, df.value):
        vals = col.value
        if "levels" in col.attrs:
            lev = col.attrs["levels"].value
            out[name] = [None if v == NA_INT else lev[v - 1] for v in vals]
        elif vals and isinstance(vals[0], float):
            out[name] = [None if v != v else v for v in vals]
        elif vals and isinstance(vals[0], int):
            out[name] = [None if v == NA_INT else v for v in vals]
        else:
            out[name] = list(vals)
    return out
