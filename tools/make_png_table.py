#!/usr/bin/env python
"""Generates the static Huffman table of the GPU PNG encoder (csrc/png.cu).

The encoder emits ONE deflate block of type 2 ("dynamic Huffman", RFC 1951 §3.2.7) whose code table is not computed per
image but fixed here once, tuned on label maps: after PNG's Up filter a 19-class segmentation map only contains the bytes
0, +-1..+-18, 255-k and the filter byte, and nearly every match is a distance-1 run.  The table is data, written to

    diga_b200/csrc/png_table.inc    constants the kernels index (bit-reversed codes, lengths, the block header bits)
    tests/golden/png_table.json     the same code LENGTHS + header bits, from which the oracle (oracle/png_oracle.py) derives
                                    the codes with its own restatement of RFC 1951 §3.2.2

Deterministic (seeded); re-run only to re-tune:  python tools/make_png_table.py
"""
import heapq
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
MAX_LIT_BITS = 12          # bounds the worst-case stream (diga_png_deflate_capacity)
CLEN_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


# ------------------------------------------------------------------------------------------------ model label maps
def model_maps():
    rng = np.random.default_rng(20260101)
    maps = []
    for block in (64, 32, 16):                                   # piecewise-constant maps, 10 % ignore
        coarse = rng.integers(0, 19, (512 // block, 1024 // block)).astype(np.uint8)
        coarse[rng.random(coarse.shape) < 0.1] = 255
        maps.append(np.kron(coarse, np.ones((block, block), np.uint8)))
    # smooth blobs: arg-max of bilinearly up-sampled random score maps (what the synthetic bench produces)
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7)
    for low in ((33, 65), (65, 129)):
        z = F.interpolate(3 * torch.randn((1, 19, *low), generator=g), (256, 512), mode="bilinear", align_corners=True)
        lab = z.argmax(1)[0].numpy().astype(np.uint8)
        maps.append(lab)
        kept = lab.copy()                                         # a consensus-selected map: disagreeing blobs -> 255
        z2 = F.interpolate(3 * torch.randn((1, 19, *low), generator=g), (256, 512), mode="bilinear", align_corners=True)
        kept[(z2.argmax(1)[0].numpy() % 3) == 0] = 255
        maps.append(kept)
    return maps


def symbol_histogram(lab):
    """Counts of the literal/length symbols (0..285) the encoder emits for one map: per scanline and run of equal filtered
    bytes a literal, (len-1)//258 matches of 258, then one match of the remainder (>= 3) or 1-2 literals."""
    prev = np.zeros_like(lab)
    prev[1:] = lab[:-1]
    f = np.concatenate([np.full((lab.shape[0], 1), 2, np.uint8), lab - prev], axis=1)
    rowlen = f.shape[1]
    flat = f.reshape(-1).astype(np.int64)
    start = np.ones(flat.shape, bool)
    start[1:] = flat[1:] != flat[:-1]
    start[::rowlen] = True
    s = np.flatnonzero(start)
    e = np.append(s[1:], flat.size)
    v, m = flat[s], e - s - 1
    nfull, rem = m // 258, m % 258
    hist = np.zeros(286, np.float64)
    np.add.at(hist, v, 1 + np.where(rem < 3, rem, 0))
    hist[285] += nfull.sum()
    big = rem >= 3
    idx = np.searchsorted(np.array(LEN_BASE), rem[big], side="right") - 1
    np.add.at(hist, 257 + idx, 1)
    hist[256] += 1
    return hist


# ------------------------------------------------------------------------------------------------ Huffman
def huffman_lengths(weights):
    heap = [(w, i, (i,)) for i, w in enumerate(weights)]
    heapq.heapify(heap)
    depth = [0] * len(weights)
    tick = len(weights)
    while len(heap) > 1:
        w1, _, s1 = heapq.heappop(heap)
        w2, _, s2 = heapq.heappop(heap)
        for i in s1 + s2:
            depth[i] += 1
        heapq.heappush(heap, (w1 + w2, tick, s1 + s2))
        tick += 1
    return depth


def limited_lengths(weights, max_bits):
    """Huffman code lengths over ALL symbols (every symbol keeps a code, so the code is complete) with the depth bounded by
    raising the floor weight until the tree fits."""
    w = np.asarray(weights, np.float64)
    total = w.sum()
    floor = total * 2.0 ** -(max_bits + 4)
    while True:
        lens = huffman_lengths(np.maximum(w, floor).tolist())
        if max(lens) <= max_bits:
            return lens
        floor *= 1.5


def canonical_codes(lengths):
    """RFC 1951 §3.2.2."""
    max_bits = max(lengths)
    bl_count = [0] * (max_bits + 1)
    for n in lengths:
        if n:
            bl_count[n] += 1
    code, next_code = 0, [0] * (max_bits + 2)
    for bits in range(1, max_bits + 1):
        code = (code + bl_count[bits - 1]) << 1
        next_code[bits] = code
    codes = [0] * len(lengths)
    for i, n in enumerate(lengths):
        if n:
            codes[i] = next_code[n]
            next_code[n] += 1
    return codes


def rev(code, n):
    return int(format(code, f"0{n}b")[::-1], 2) if n else 0


# ------------------------------------------------------------------------------------------------ block header
def rle_code_lengths(seq):
    """Code-length alphabet (RFC 1951 §3.2.7): 0-15 literal lengths, 16 = repeat previous 3-6 (2 extra bits), 17 = zeros
    3-10 (3 bits), 18 = zeros 11-138 (7 bits).  Greedy."""
    out, i = [], 0
    while i < len(seq):
        v, j = seq[i], i
        while j < len(seq) and seq[j] == v:
            j += 1
        run = j - i
        if v == 0:
            while run >= 11:
                k = min(run, 138)
                out.append((18, k - 11, 7))
                run -= k
            if run >= 3:
                out.append((17, run - 3, 3))
                run = 0
            out += [(0, 0, 0)] * run
        else:
            out.append((v, 0, 0))
            run -= 1
            while run >= 3:
                k = min(run, 6)
                out.append((16, k - 3, 2))
                run -= k
            out += [(v, 0, 0)] * run
        i = j
    return out


class Bits:
    def __init__(self):
        self.bits = []

    def put(self, value, n):                  # LSB first
        self.bits += [(value >> k) & 1 for k in range(n)]

    def put_code(self, code, n):              # Huffman codes: MSB first
        self.put(rev(code, n), n)


def block_header(lit_lengths, dist_lengths):
    syms = rle_code_lengths(list(lit_lengths) + list(dist_lengths))
    freq = [0] * 19
    for s, _, _ in syms:
        freq[s] += 1
    used = [i for i in range(19) if freq[i]]
    cl = [0] * 19
    sub = limited_lengths([freq[i] for i in used], 7) if len(used) > 1 else [1]
    for i, n in zip(used, sub):
        cl[i] = n
    cl_codes = canonical_codes(cl)
    hclen = max(k for k in range(19) if cl[CLEN_ORDER[k]]) + 1
    hclen = max(hclen, 4)
    b = Bits()
    b.put(1, 1)                               # BFINAL
    b.put(2, 2)                               # BTYPE = 10
    b.put(len(lit_lengths) - 257, 5)          # HLIT
    b.put(len(dist_lengths) - 1, 5)           # HDIST
    b.put(hclen - 4, 4)                       # HCLEN
    for k in range(hclen):
        b.put(cl[CLEN_ORDER[k]], 3)
    for s, extra, nextra in syms:
        b.put_code(cl_codes[s], cl[s])
        b.put(extra, nextra)
    return b.bits


def c_array(name, ctype, values, per_line=16, space="__device__ const"):
    """Lane-divergent look-ups (one literal value per lane) go through L1 as plain global loads; __constant__ memory would
    serialise them per distinct address."""
    rows = [", ".join(str(v) for v in values[i:i + per_line]) for i in range(0, len(values), per_line)]
    return f"{space} {ctype} {name}[{len(values)}] = {{\n    " + ",\n    ".join(rows) + "};\n"


def main():
    hist = np.zeros(286)
    for lab in model_maps():
        h = symbol_histogram(lab)
        hist += h / h.sum()
    lit_lengths = limited_lengths(hist, MAX_LIT_BITS)
    assert abs(sum(2.0 ** -n for n in lit_lengths) - 1.0) < 1e-12, "literal/length code must be complete"
    dist_lengths = [1]                        # a single distance code (distance 1): one bit, RFC 1951 §3.2.7
    codes = canonical_codes(lit_lengths)
    header = block_header(lit_lengths, dist_lengths)
    words = [0] * ((len(header) + 31) // 32)
    for i, bit in enumerate(header):
        words[i // 32] |= bit << (i % 32)
    with open(os.path.join(ROOT, "tests", "golden", "png_table.json"), "w") as f:
        json.dump({"lit_lengths": lit_lengths, "dist_lengths": dist_lengths, "header_bits": "".join(map(str, header))}, f)
        f.write("\n")
    inc = ["// GENERATED by tools/make_png_table.py — static Huffman table of the PNG encoder (do not edit).\n",
           f"constexpr int kPngHeaderBits = {len(header)};   // BFINAL, BTYPE=10, HLIT, HDIST, HCLEN, code-length codes, code lengths\n",
           f"constexpr int kPngMaxLitBits = {max(lit_lengths[:256])};\n",
           c_array("kPngHeaderWords", "uint32_t", [f"0x{w:08x}u" for w in words], 6, "__constant__"),
           c_array("kPngLitPat", "uint16_t", [rev(codes[v], lit_lengths[v]) for v in range(256)]),
           c_array("kPngLitLen", "uint8_t", lit_lengths[:256], 32),
           c_array("kPngLenPat", "uint16_t", [rev(codes[257 + i], lit_lengths[257 + i]) for i in range(29)]),
           c_array("kPngLenLen", "uint8_t", lit_lengths[257:286], 32),
           f"constexpr uint32_t kPngEobPat = {rev(codes[256], lit_lengths[256])}u;\n",
           f"constexpr int kPngEobLen = {lit_lengths[256]};\n"]
    with open(os.path.join(ROOT, "diga_b200", "csrc", "png_table.inc"), "w") as f:
        f.writelines(inc)
    used = {v: lit_lengths[v] for v in (0, 1, 2, 18, 19, 237, 255, 100)}
    print(f"header {len(header)} bits; literal bits {used}; len258 {lit_lengths[285]} bits; eob {lit_lengths[256]} bits; "
          f"max literal {max(lit_lengths[:256])}")


if __name__ == "__main__":
    main()
