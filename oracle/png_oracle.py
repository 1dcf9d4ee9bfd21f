"""CPU oracle for the GPU-side PNG (zlib / deflate) encoder of the pseudo-label writer.

TEST INFRASTRUCTURE ONLY — nothing under ``diga_b200/`` imports this module.

The reference writes its pseudo-labels with ``colorize_mask(label).save(path)`` (``pseudolabel_generator.py:45-49,
:100-105``): Pillow's PNG plugin on top of zlib, neither of which is part of ``/root/reference``
(``requirements.txt``: ``Pillow==8.1.0``; zlib is whatever the Python build links).  What the reference pins is the
*decoded* file: mode 'P', palette = the Cityscapes trainId colours, pixel index = trainId.  ``csrc/png.cu`` emits a
particular, much simpler deflate stream (Up filter, one block with a STATIC Huffman table tuned on label maps, literal +
distance-1 matches); this module restates exactly that token stream from the published formats (RFC 1950, RFC 1951, PNG
1.2 §6 / §9) in numpy, so the GPU bytes can be checked bit for bit.  The table itself is data: the code LENGTHS and the
block-header bits in ``tests/golden/png_table.json`` (written by ``tools/make_png_table.py`` together with the constants
the kernels index); the codes are derived from the lengths here, independently, by RFC 1951 §3.2.2, and pins itself against the standard decoders: ``zlib.decompress`` must return the
filtered scanlines, and Pillow must open the framed file to the same pixels / mode / palette as the file Pillow itself
writes for ``colorize_mask`` (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

LEN_BASE = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195,
                     227, 258], dtype=np.int64)                      # RFC 1951 §3.2.5, length codes 257..285
LEN_EXTRA = np.array([0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0], dtype=np.int64)


def filtered_scanlines(label: np.ndarray) -> np.ndarray:
    """PNG 1.2 §6.3 filter type 2 (Up) on every scanline: ``[H, W+1]`` uint8, column 0 = the filter byte."""
    label = np.ascontiguousarray(label, dtype=np.uint8)
    prev = np.zeros_like(label)
    prev[1:] = label[:-1]
    out = np.empty((label.shape[0], label.shape[1] + 1), dtype=np.uint8)
    out[:, 0] = 2
    out[:, 1:] = label - prev                                         # uint8 wrap-around = mod 256
    return out


def _table():
    global _TABLE
    if _TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "png_table.json")) as f:
            t = json.load(f)
        lengths = np.array(t["lit_lengths"], dtype=np.int64)
        assert lengths.size == 286 and t["dist_lengths"] == [1]
        # canonical Huffman codes from the lengths alone (RFC 1951 §3.2.2): symbols ordered by (length, symbol) take
        # consecutive code values; the value is shifted left whenever the length grows
        codes = np.zeros(286, dtype=np.int64)
        code, prev = 0, None
        for sym in sorted(np.flatnonzero(lengths), key=lambda i: (lengths[i], i)):
            if prev is not None:
                code = (code + 1) << int(lengths[sym] - lengths[prev])
            codes[sym] = code
            prev = sym
        _TABLE = {"len": lengths, "pat": _rev(codes, lengths), "header": np.array([int(c) for c in t["header_bits"]], dtype=np.uint8)}
    return _TABLE


_TABLE = None


def _rev(x: np.ndarray, n: np.ndarray) -> np.ndarray:
    """Huffman codes are packed starting from their most significant bit (RFC 1951 §3.1.1)."""
    r = np.zeros_like(x)
    for j in range(16):
        r |= np.where(j < n, ((x >> j) & 1) << np.maximum(n - 1 - j, 0), 0)
    return r


def _literal(v: np.ndarray):
    t = _table()
    return t["pat"][v], t["len"][v]


def _match(length: np.ndarray):
    """Length code + extra bits (LSB first) + the single distance code (distance 1): one 0 bit."""
    t = _table()
    idx = np.searchsorted(LEN_BASE, length, side="right") - 1
    hb = t["len"][257 + idx]
    pat = t["pat"][257 + idx] | ((length - LEN_BASE[idx]) << hb)
    return pat, hb + LEN_EXTRA[idx] + 1


def token_stream(label: np.ndarray):
    """The encoder's tokens for one image as ``(pattern, nbits)`` arrays in stream order: per scanline, per run of equal
    filtered bytes: literal, ``(len-1) // 258`` matches of 258, then one match of the remainder (>= 3) or 1-2 literals."""
    f = filtered_scanlines(label)
    h, rowlen = f.shape
    flat = f.reshape(-1).astype(np.int64)
    start = np.ones(flat.shape, dtype=bool)
    start[1:] = flat[1:] != flat[:-1]
    start[::rowlen] = True                                            # a run never crosses a scanline
    s = np.flatnonzero(start)
    e = np.append(s[1:], flat.size)
    v = flat[s]
    m = e - s - 1
    nfull, rem = m // 258, m % 258
    count = 1 + nfull + np.where(rem >= 3, 1, rem)
    first = np.cumsum(count) - count
    run = np.repeat(np.arange(s.size), count)
    k = np.arange(int(count.sum())) - first[run]
    lit_p, lit_n = _literal(v)
    rem_p, rem_n = _match(np.maximum(rem, 3))
    full_p, full_n = _match(np.array([258]))
    is_lit = (k == 0) | ((k > nfull[run]) & (rem[run] < 3))
    is_full = (k >= 1) & (k <= nfull[run])
    pat = np.where(is_lit, lit_p[run], np.where(is_full, full_p[0], rem_p[run]))
    nb = np.where(is_lit, lit_n[run], np.where(is_full, full_n[0], rem_n[run]))
    return pat.astype(np.int64), nb.astype(np.int64)


def deflate_stream(label: np.ndarray) -> bytes:
    """zlib stream (RFC 1950) of the filtered scanlines exactly as ``csrc/png.cu`` lays it out: 78 01, the block header
    (BFINAL=1, BTYPE=10 and the static code table), the tokens, the end-of-block code, zero padding, Adler-32."""
    t = _table()
    pat, nb = token_stream(label)
    hdr = t["header"]
    first = 16 + hdr.size
    pos = first + np.cumsum(nb) - nb
    eob_at = first + int(nb.sum())
    eob_n = int(t["len"][256])
    nbytes = (eob_at + eob_n + 7) // 8
    bits = np.zeros(nbytes * 8, dtype=np.uint8)
    bits[16:first] = hdr
    for j in range(int(nb.max()) if nb.size else 0):
        sel = nb > j
        bits[pos[sel] + j] = (pat[sel] >> j) & 1
    for j in range(eob_n):
        bits[eob_at + j] = (int(t["pat"][256]) >> j) & 1
    body = np.packbits(bits, bitorder="little")
    body[0], body[1] = 0x78, 0x01
    adler = zlib.adler32(filtered_scanlines(label).tobytes())
    return body.tobytes() + adler.to_bytes(4, "big")
