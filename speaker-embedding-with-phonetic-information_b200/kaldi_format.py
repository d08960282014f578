"""Kaldi on-disk formats touched by the extraction path (Python side).

Host-side utility used by the fixture generator, the tests and bench.py to
write inputs the C++ host (csrc/kaldi_io.cc, csrc/nnet3_model.cc) reads, and
to read back what it writes.  It is an independent implementation of the same
formats, so a write here / read there round trip checks both.

Formats follow upstream Kaldi (not shipped in the reference, see SURVEY.md
section 0 and 8a-1/8a-7); the reference pins only where they are used:
  * feature rspecifier / vector wspecifier:
    egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:79,93
  * nnet3 raw model `final.raw`: extract_xvectors_new.sh:46,59
"""
from __future__ import annotations

import io
import os
import re
import struct
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import BinaryIO, Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

# --------------------------------------------------------------------------
# basic tokens / scalars
# --------------------------------------------------------------------------


def write_token(f: BinaryIO, binary: bool, tok: str) -> None:
    f.write(tok.encode() + b" ")


def write_int32(f: BinaryIO, binary: bool, v: int) -> None:
    if binary:
        f.write(b"\x04" + struct.pack("<i", v))
    else:
        f.write(f"{v} ".encode())


def write_float(f: BinaryIO, binary: bool, v: float) -> None:
    if binary:
        f.write(b"\x04" + struct.pack("<f", v))
    else:
        f.write(f"{float(np.float32(v))!r} ".encode())


def write_double(f: BinaryIO, binary: bool, v: float) -> None:
    if binary:
        f.write(b"\x08" + struct.pack("<d", v))
    else:
        f.write(f"{float(v)!r} ".encode())


def write_bool(f: BinaryIO, binary: bool, v: bool) -> None:
    f.write(b"T" if v else b"F")
    if not binary:
        f.write(b" ")


def write_vector(f: BinaryIO, binary: bool, v: np.ndarray) -> None:
    """Kaldi Vector<float|double>::Write."""
    v = np.ascontiguousarray(v)
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float32)
    if binary:
        f.write(b"FV " if v.dtype == np.float32 else b"DV ")
        write_int32(f, True, v.shape[0])
        f.write(v.tobytes())
    else:
        f.write(b" [ " + _fmt_row(v).encode() + b" ]\n")


def write_matrix(f: BinaryIO, binary: bool, m: np.ndarray) -> None:
    """Kaldi Matrix<float|double>::Write (uncompressed)."""
    m = np.ascontiguousarray(m)
    if m.dtype not in (np.float32, np.float64):
        m = m.astype(np.float32)
    assert m.ndim == 2
    if binary:
        f.write(b"FM " if m.dtype == np.float32 else b"DM ")
        write_int32(f, True, m.shape[0])
        write_int32(f, True, m.shape[1])
        f.write(m.tobytes())
    else:
        if m.shape[0] == 0:
            f.write(b" [ ]\n")
            return
        f.write(b" [")
        for r in range(m.shape[0]):
            f.write(b"\n  " + _fmt_row(m[r]).encode())
        f.write(b" ]\n")


def _fmt_row(row: np.ndarray) -> str:
    # 9 (float32) / 17 (float64) significant digits round-trip exactly
    f = "%.9g" if row.dtype == np.float32 else "%.17g"
    return " ".join([f % x for x in row.tolist()])


# --------------------------------------------------------------------------
# reading
# --------------------------------------------------------------------------


class _Reader:
    """Byte reader with one-byte peek over a bytes object."""

    def __init__(self, data: bytes, pos: int = 0):
        self.d = data
        self.p = pos

    def peek(self, n: int = 1) -> bytes:
        return self.d[self.p:self.p + n]

    def read(self, n: int) -> bytes:
        b = self.d[self.p:self.p + n]
        if len(b) != n:
            raise EOFError("unexpected end of Kaldi stream")
        self.p += n
        return b

    def eof(self) -> bool:
        return self.p >= len(self.d)

    def skip_ws(self) -> None:
        while self.p < len(self.d) and self.d[self.p:self.p + 1] in b" \t\r\n":
            self.p += 1

    def token(self) -> str:
        self.skip_ws()
        s = self.p
        while self.p < len(self.d) and self.d[self.p:self.p + 1] not in b" \t\r\n":
            self.p += 1
        tok = self.d[s:self.p].decode()
        if self.p < len(self.d):
            self.p += 1  # consume exactly one whitespace, as Kaldi ReadToken does
        return tok

    def peek_token(self) -> str:
        save = self.p
        t = self.token()
        self.p = save
        return t


def read_int32(r: _Reader, binary: bool) -> int:
    if binary:
        n = r.read(1)[0]
        if n != 4:
            raise ValueError(f"expected int32 size byte 4, got {n}")
        return struct.unpack("<i", r.read(4))[0]
    return int(r.token())


def read_float(r: _Reader, binary: bool) -> float:
    if binary:
        n = r.read(1)[0]
        if n == 4:
            return struct.unpack("<f", r.read(4))[0]
        if n == 8:
            return struct.unpack("<d", r.read(8))[0]
        raise ValueError(f"bad float size byte {n}")
    return float(r.token())


def read_bool(r: _Reader, binary: bool) -> bool:
    if not binary:
        r.skip_ws()
    c = r.read(1)
    if c not in b"TF":
        raise ValueError(f"bad bool {c!r}")
    if not binary and r.peek() in (b" ", b"\n"):
        r.read(1)
    return c == b"T"


def read_vector_or_matrix(r: _Reader, binary: bool) -> np.ndarray:
    """Reads FV/DV/FM/DM (binary) or `[ ... ]` (text; 1-row text matrices come
    back as vectors -- callers reshape by field)."""
    if binary:
        tag = r.token()
        if tag in ("FV", "DV"):
            dt = np.float32 if tag == "FV" else np.float64
            n = read_int32(r, True)
            return np.frombuffer(r.read(n * dt().itemsize), dtype=dt).copy()
        if tag in ("FM", "DM"):
            dt = np.float32 if tag == "FM" else np.float64
            rows = read_int32(r, True)
            cols = read_int32(r, True)
            a = np.frombuffer(r.read(rows * cols * dt().itemsize), dtype=dt)
            return a.reshape(rows, cols).copy()
        if tag in ("CM", "CM2", "CM3"):
            return _read_compressed(r, tag)
        raise ValueError(f"unknown binary matrix tag {tag!r}")
    r.skip_ws()
    if r.read(1) != b"[":
        raise ValueError("expected '[' in text matrix")
    end = r.d.index(b"]", r.p)
    body = r.d[r.p:end].decode()
    r.p = end + 1
    rows = [ln.split() for ln in body.split("\n") if ln.strip()]
    if len(rows) == 0:
        return np.zeros((0,), dtype=np.float32)
    if len(rows) == 1:
        return np.array(rows[0], dtype=np.float32)
    return np.array(rows, dtype=np.float32)


def _read_compressed(r: _Reader, tag: str) -> np.ndarray:
    """Kaldi CompressedMatrix (formats 1-3), as found in on-disk feats.scp
    archives written by steps/make_mfcc.sh."""
    min_value, rng = struct.unpack("<ff", r.read(8))
    rows, cols = struct.unpack("<ii", r.read(8))
    if tag == "CM":
        hdr = np.frombuffer(r.read(cols * 8), dtype="<u2").reshape(cols, 4).astype(np.float32)
        p = min_value + rng * (1.0 / 65535.0) * hdr  # p0,p25,p75,p100 per column
        b = np.frombuffer(r.read(rows * cols), dtype=np.uint8).reshape(cols, rows).astype(np.float32)
        out = np.empty((cols, rows), dtype=np.float32)
        for c in range(cols):
            v = b[c]
            p0, p25, p75, p100 = p[c]
            out[c] = np.where(
                v <= 64, p0 + (p25 - p0) * v * (1 / 64.0),
                np.where(v <= 192, p25 + (p75 - p25) * (v - 64) * (1 / 128.0),
                         p75 + (p100 - p75) * (v - 192) * (1 / 63.0)))
        return np.ascontiguousarray(out.T)
    if tag == "CM2":
        a = np.frombuffer(r.read(rows * cols * 2), dtype="<u2").reshape(rows, cols)
        return (min_value + rng * (1.0 / 65535.0) * a.astype(np.float32)).astype(np.float32)
    a = np.frombuffer(r.read(rows * cols), dtype=np.uint8).reshape(rows, cols)
    return (min_value + rng * (1.0 / 255.0) * a.astype(np.float32)).astype(np.float32)


# --------------------------------------------------------------------------
# tables (ark / scp)
# --------------------------------------------------------------------------


def write_ark(path_or_file, items, binary: bool = True, scp_path: Optional[str] = None,
              ark_name_in_scp: Optional[str] = None) -> None:
    """Write (key, ndarray) pairs as a Kaldi archive; 1-D arrays are vectors,
    2-D arrays are matrices.  With scp_path also writes `key ark:offset`."""
    own = isinstance(path_or_file, (str, os.PathLike))
    f = open(path_or_file, "wb") if own else path_or_file
    scp = open(scp_path, "w") if scp_path else None
    name = ark_name_in_scp or (str(path_or_file) if own else "-")
    try:
        for key, a in items:
            f.write(key.encode() + b" ")
            off = f.tell()
            if binary:
                f.write(b"\0B")
            (write_vector if np.ndim(a) == 1 else write_matrix)(f, binary, np.asarray(a))
            if scp:
                scp.write(f"{key} {name}:{off}\n")
    finally:
        if own:
            f.close()
        if scp:
            scp.close()


def read_ark(path_or_bytes) -> Iterator[Tuple[str, np.ndarray]]:
    data = path_or_bytes if isinstance(path_or_bytes, bytes) else open(path_or_bytes, "rb").read()
    r = _Reader(data)
    while True:
        r.skip_ws()
        if r.eof():
            return
        key = r.token()
        binary = r.peek(2) == b"\0B"
        if binary:
            r.read(2)
        yield key, read_vector_or_matrix(r, binary)


def read_scp(path: str) -> Iterator[Tuple[str, np.ndarray]]:
    for ln in open(path):
        ln = ln.strip()
        if not ln:
            continue
        key, rx = ln.split(None, 1)
        m = re.match(r"^(.*):(\d+)$", rx)
        fn, off = (m.group(1), int(m.group(2))) if m else (rx, 0)
        data = open(fn, "rb").read()
        r = _Reader(data, off)
        binary = r.peek(2) == b"\0B"
        if binary:
            r.read(2)
        yield key, read_vector_or_matrix(r, binary)


# --------------------------------------------------------------------------
# nnet3 raw model
# --------------------------------------------------------------------------


@dataclass
class Component:
    type: str                      # e.g. "NaturalGradientAffineComponent"
    fields: "OrderedDict[str, object]" = field(default_factory=OrderedDict)  # token (no <>) -> value


@dataclass
class Nnet3Model:
    config_lines: List[str] = field(default_factory=list)       # node lines, as `nnet3-copy` would print them
    components: "OrderedDict[str, Component]" = field(default_factory=OrderedDict)


# Field typing used by the *writer* (the reader is token driven and generic).
_INT_FIELDS = {"Dim", "BlockDim", "InputDim", "OutputDim", "InputPeriod", "OutputPeriod", "LeftContext",
               "RightContext", "NumLogCountFeatures", "RankIn", "RankOut", "UpdatePeriod"}
_BOOL_FIELDS = {"IsGradient", "TestMode", "IncludeVarinance", "OutputStddevs"}
_DOUBLE_FIELDS = {"Count", "OderivCount", "NumDimsSelfRepaired", "NumDimsProcessed"}


def write_nnet3(path: str, model: Nnet3Model, binary: bool = True) -> None:
    """`Nnet::Write` layout: `<Nnet3>` + node config lines (always text) + blank
    line + `<NumComponents>` + per-component blobs + `</Nnet3>`."""
    with open(path, "wb") as f:
        if binary:
            f.write(b"\0B")
        write_token(f, binary, "<Nnet3>")
        f.write(b"\n")
        for ln in model.config_lines:
            f.write(ln.encode() + b"\n")
        f.write(b"\n")
        write_token(f, binary, "<NumComponents>")
        write_int32(f, binary, len(model.components))
        if not binary:
            f.write(b"\n")
        for name, comp in model.components.items():
            write_token(f, binary, "<ComponentName>")
            write_token(f, binary, name)
            write_token(f, binary, f"<{comp.type}>")
            for tok, val in comp.fields.items():
                write_token(f, binary, f"<{tok}>")
                if isinstance(val, np.ndarray):
                    (write_vector if val.ndim == 1 else write_matrix)(f, binary, val)
                elif tok in _BOOL_FIELDS or isinstance(val, (bool, np.bool_)):
                    write_bool(f, binary, bool(val))
                elif tok in _INT_FIELDS:
                    write_int32(f, binary, int(val))
                elif tok in _DOUBLE_FIELDS:
                    write_double(f, binary, float(val))
                else:
                    write_float(f, binary, float(val))
            write_token(f, binary, f"</{comp.type}>")
            if not binary:
                f.write(b"\n")
        write_token(f, binary, "</Nnet3>")


def read_nnet3(path_or_bytes: Union[str, bytes]) -> Nnet3Model:
    """Token-driven tolerant reader (SURVEY Appendix D `load_model`)."""
    data = path_or_bytes if isinstance(path_or_bytes, bytes) else open(path_or_bytes, "rb").read()
    r = _Reader(data)
    binary = r.peek(2) == b"\0B"
    if binary:
        r.read(2)
    if r.token() != "<Nnet3>":
        raise ValueError("expected <Nnet3>")
    model = Nnet3Model()
    # config section: text lines up to an empty line
    nl = r.d.index(b"\n", r.p)
    r.p = nl + 1
    while True:
        nl = r.d.index(b"\n", r.p)
        ln = r.d[r.p:nl].decode().strip()
        r.p = nl + 1
        if ln == "":
            break
        model.config_lines.append(ln)
    if r.token() != "<NumComponents>":
        raise ValueError("expected <NumComponents>")
    n = read_int32(r, binary)
    for _ in range(n):
        if r.token() != "<ComponentName>":
            raise ValueError("expected <ComponentName>")
        name = r.token()
        ctype = r.token()
        assert ctype.startswith("<") and ctype.endswith(">"), ctype
        ctype = ctype[1:-1]
        comp = Component(ctype)
        while True:
            tok = r.token()
            if tok == f"</{ctype}>":
                break
            if not (tok.startswith("<") and tok.endswith(">")):
                raise ValueError(f"component {name}: expected a <Token>, got {tok!r}")
            comp.fields[tok[1:-1]] = _read_generic_value(r, binary)
        model.components[name] = comp
    if r.token() != "</Nnet3>":
        raise ValueError("expected </Nnet3>")
    return model


def _read_generic_value(r: _Reader, binary: bool):
    if binary:
        c = r.peek(1)
        if c == b"\x04":
            # int32 and float share the size byte; keep raw bytes, decide by caller
            raw = r.read(5)[1:]
            return _Scalar4(raw)
        if c == b"\x08":
            return struct.unpack("<d", r.read(9)[1:])[0]
        if c in b"TF" and r.peek(3)[1:2] in (b"<", b" ") and r.peek(3) not in (b"FV ", b"FM "):
            return r.read(1) == b"T"
        if r.peek(3) in (b"FV ", b"DV ", b"FM ", b"DM "):
            return read_vector_or_matrix(r, True)
        if c in b"TF":
            return r.read(1) == b"T"
        raise ValueError(f"cannot skip unknown binary value starting {r.peek(4)!r}")
    r.skip_ws()
    if r.peek(1) == b"[":
        return read_vector_or_matrix(r, False)
    t = r.token()
    if t in ("T", "F"):
        return t == "T"
    return _ScalarText(t)


class _Scalar4:
    """4-byte scalar of unknown type (int32 or float32) from a binary model."""

    def __init__(self, raw: bytes):
        self.raw = raw

    def __int__(self):
        return struct.unpack("<i", self.raw)[0]

    def __float__(self):
        return struct.unpack("<f", self.raw)[0]


class _ScalarText:
    def __init__(self, s: str):
        self.s = s

    def __int__(self):
        return int(float(self.s))

    def __float__(self):
        return float(self.s)
