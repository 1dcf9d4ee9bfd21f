// f3, writer half — the zlib stream of a pseudo-label PNG, produced on the GPU.
// Replaces the host-side `Image.save` of pseudolabel_generator.py:45-49,100-105 (PIL + zlib, 17-43 ms per 2048x1024 map
// on one host core — two orders of magnitude more than every GPU kernel of config 5 together).  The label map stays
// in HBM; what crosses PCIe is the finished IDAT payload (10-60 KB instead of 2 MB) and the host only frames it
// (signature, IHDR, PLTE, IDAT + CRC, IEND).  Decoded pixels, mode ('P') and palette equal the reference's files; the
// bytes differ (as they do between zlib versions).
//
// Stream layout (RFC 1950 / 1951 / PNG 1.2):
//   zlib header 78 01 | ONE deflate block (BFINAL=1, BTYPE=10) with a STATIC code table | pad to byte | Adler-32 (big endian)
//   scanline y = filter byte 2 (Up) followed by label[y][x] - label[y-1][x] (mod 256; row -1 is zero).
// A segmentation map filtered that way is zero except along horizontal class edges, so the encoder only needs runs:
// every run of equal bytes becomes   literal(v)   then   match(len <= 258, distance 1)   tokens.
// The Huffman table is not computed per image: a filtered 19-class map only holds the bytes 0, +-1..+-18, 255-k and the
// filter byte, and nearly all matches are 258-byte runs, so ONE table tuned on label maps (tools/make_png_table.py ->
// png_table.inc: 2 bits for a zero, 4 bits for a 258-byte match instead of 8 and 13 with the fixed code of BTYPE=01) is
// sent as the block's "dynamic" header (350 constant bits).  Every byte value keeps a code (<= 12 bits), so any uint8 map
// encodes.  Files come out at 0.4x (blocky maps) to 1.1x (noise-like maps) of Pillow's.
//
// Three launches per batch, no host sync:
//   rows<false> : one warp per scanline — filter into shared memory, find the run starts (lane-segment scan + warp prefix
//                 sum into a shared list), one lane per run computes its token bits  -> row bit count, Adler partials
//   layout      : one CTA per image — exclusive scan of the row bit counts, Adler-32 combine, zero the output span,
//                 zlib header, trailer, byte length
//   rows<true>  : same walk, tokens OR-ed into the bit stream at their global bit offset (atomicOr on 32-bit words)
#include <errno.h>
#include <string.h>

#include "common.cuh"

namespace diga {

struct PngRowStat {
  uint32_t bits;      // token bits of the scanline
  uint32_t sum;       // sum of its filtered bytes
  uint64_t wsum_off;  // rows<false>: sum of i * byte[i];  after layout: bit offset of the scanline's first token
};

__constant__ uint16_t kLenBase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                      31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};

#include "png_table.inc"

// The look-ups are lane-divergent (one byte value per lane), so every CTA first packs the table into shared memory:
// literal entry = pattern | bits << 16;  length entry = pattern | code bits << 16 | total bits (code + extra + distance) << 24.
struct PngLut {
  uint32_t lit[256];
  uint32_t len[29];
};

__device__ __forceinline__ void png_lut_fill(PngLut& L) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) L.lit[i] = (uint32_t)kPngLitPat[i] | ((uint32_t)kPngLitLen[i] << 16);
  for (int i = threadIdx.x; i < 29; i += blockDim.x)
    L.len[i] = (uint32_t)kPngLenPat[i] | ((uint32_t)kPngLenLen[i] << 16) | ((uint32_t)(kPngLenLen[i] + kLenExtra[i] + 1) << 24);
  __syncthreads();
}

// match of `len` (3..258) at distance 1: length code + extra bits (LSB first) + the single distance code (one 0 bit)
__device__ __forceinline__ int match_code(int len) {
  int lo = 0, hi = 28;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int)kLenBase[mid] <= len) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int match_bits_of(const PngLut& L, int idx) { return (int)(L.len[idx] >> 24); }
__device__ __forceinline__ uint32_t match_pattern(const PngLut& L, int len, int idx) {
  const uint32_t e = L.len[idx];
  return (e & 0xffffu) | ((uint32_t)(len - kLenBase[idx]) << ((e >> 16) & 0xffu));
}

__device__ __forceinline__ void put_bits(uint32_t* out, uint64_t pos, uint32_t pattern, int nbits) {
  const uint64_t word = pos >> 5;
  const int sh = (int)(pos & 31);
  const uint64_t p = (uint64_t)pattern << sh;
  atomicOr(out + word, (uint32_t)p);
  if (sh + nbits > 32) atomicOr(out + word + 1, (uint32_t)(p >> 32));
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t& total) {
  const int lane = threadIdx.x & 31;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

template <bool WRITE>
__global__ void png_rows_kernel(const uint8_t* __restrict__ labels, int H, int W, PngRowStat* __restrict__ stats,
                                uint8_t* __restrict__ out, int64_t capacity, int smem_per_warp) {
  extern __shared__ __align__(16) unsigned char png_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int y = blockIdx.x * warps + warp;
  const int64_t img = blockIdx.y;
  __shared__ PngLut lut;
  png_lut_fill(lut);
  if (y >= H) return;
  const int rowlen = W + 1;
  uint8_t* f = png_smem + (size_t)warp * smem_per_warp;                            // filtered scanline
  uint16_t* starts = reinterpret_cast<uint16_t*>(f + ((rowlen + 3) & ~3));          // run start positions
  const uint8_t* cur = labels + (img * H + y) * (int64_t)W;
  const uint8_t* prv = cur - W;

  // ---- filter (Up) into shared memory, Adler partials --------------------------------------------------------
  uint32_t s1 = 0;
  uint64_t s2 = 0;
  if (lane == 0) {
    f[0] = 2;
    s1 = 2;
  }
  if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(labels) & 3) == 0)) {
    for (int x = lane * 4; x < W; x += 128) {
      const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(cur + x));
      const uint32_t b = y > 0 ? __ldg(reinterpret_cast<const uint32_t*>(prv + x)) : 0u;
      const uint32_t d = __vsub4(a, b);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t v = (d >> (8 * k)) & 0xffu;
        f[1 + x + k] = (uint8_t)v;
        s1 += v;
        s2 += (uint64_t)(1 + x + k) * v;
      }
    }
  } else {
    for (int x = lane; x < W; x += 32) {
      const uint32_t v = (uint32_t)(uint8_t)(__ldg(cur + x) - (y > 0 ? __ldg(prv + x) : (uint8_t)0));
      f[1 + x] = (uint8_t)v;
      s1 += v;
      s2 += (uint64_t)(1 + x) * v;
    }
  }
  __syncwarp();

  // ---- run starts: every lane scans its own segment, a warp prefix sum places them in order -----------------------
  const int seg = (rowlen + 31) >> 5;
  const int a0 = lane * seg, a1 = min(a0 + seg, rowlen);
  uint32_t mine = 0;
  {
    int pv = a0 > 0 && a0 < rowlen ? f[a0 - 1] : -1;
    for (int i = a0; i < a1; ++i) {
      const int v = f[i];
      mine += (v != pv);
      pv = v;
    }
  }
  uint32_t nruns;
  uint32_t at = warp_excl_scan(mine, nruns);
  {
    int pv = a0 > 0 && a0 < rowlen ? f[a0 - 1] : -1;
    for (int i = a0; i < a1; ++i) {
      const int v = f[i];
      if (v != pv) starts[at++] = (uint16_t)i;
      pv = v;
    }
  }
  __syncwarp();

  // ---- one lane per run: literal + distance-1 matches ------------------------------------------------------------
  uint32_t* out32 = WRITE ? reinterpret_cast<uint32_t*>(out + img * capacity) : nullptr;
  const uint64_t row_off = WRITE ? stats[img * H + y].wsum_off : 0;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nruns; base += 32) {
    const uint32_t r = base + lane;
    uint32_t bits = 0, v = 0;
    int nfull = 0, rem = 0, ridx = 0;
    if (r < nruns) {
      const int s = starts[r], e = r + 1 < nruns ? (int)starts[r + 1] : rowlen;
      v = f[s];
      const int m = e - s - 1;                 // bytes after the literal
      nfull = m / 258;
      rem = m - nfull * 258;
      const int lb = (int)(lut.lit[v] >> 16);
      bits = lb + nfull * match_bits_of(lut, 28);
      if (rem >= 3) {
        ridx = match_code(rem);
        bits += match_bits_of(lut, ridx);
      } else {
        bits += rem * lb;
      }
    }
    uint32_t total;
    const uint32_t excl = warp_excl_scan(bits, total);
    if (WRITE && r < nruns) {
      uint64_t pos = row_off + carry + excl;
      const int lb = (int)(lut.lit[v] >> 16);
      const uint32_t lp = lut.lit[v] & 0xffffu;
      put_bits(out32, pos, lp, lb);
      pos += lb;
      for (int k = 0; k < nfull; ++k) {
        put_bits(out32, pos, lut.len[28] & 0xffffu, match_bits_of(lut, 28));   // code 285 (len 258), no extra bits, distance code
        pos += match_bits_of(lut, 28);
      }
      if (rem >= 3) {
        put_bits(out32, pos, match_pattern(lut, rem, ridx), match_bits_of(lut, ridx));
      } else {
        for (int k = 0; k < rem; ++k) {
          put_bits(out32, pos, lp, lb);
          pos += lb;
        }
      }
    }
    carry += total;
  }

  if (!WRITE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      PngRowStat st;
      st.bits = carry;
      st.sum = s1;
      st.wsum_off = s2;
      stats[img * H + y] = st;
    }
  }
}

constexpr uint32_t kAdlerMod = 65521u;
constexpr int kLayoutBlock = 1024;

__global__ void __launch_bounds__(kLayoutBlock)
png_layout_kernel(PngRowStat* __restrict__ stats, int H, int W, uint8_t* __restrict__ out, int64_t capacity,
                  int64_t* __restrict__ lengths) {
  __shared__ uint64_t warp_tot[kLayoutBlock / 32];
  __shared__ uint64_t red[3][kLayoutBlock / 32];
  __shared__ uint64_t carry_s;
  const int64_t img = blockIdx.x;
  PngRowStat* st = stats + img * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t rowlen = (uint64_t)W + 1;
  if (threadIdx.x == 0) carry_s = 16 + kPngHeaderBits;     // zlib header (2 bytes) + block header with the code table
  __syncthreads();
  uint64_t sum_d = 0, sum_g = 0;                            // both kept < kAdlerMod * small
  for (int y0 = 0; y0 < H; y0 += kLayoutBlock) {
    const int y = y0 + threadIdx.x;
    uint64_t bits = 0;
    if (y < H) {
      const PngRowStat s = st[y];
      bits = s.bits;
      sum_d = (sum_d + s.sum) % kAdlerMod;
      const uint64_t g = ((((uint64_t)y * rowlen) % kAdlerMod) * (s.sum % kAdlerMod) + s.wsum_off % kAdlerMod) % kAdlerMod;
      sum_g = (sum_g + g) % kAdlerMod;
    }
    uint64_t inc = bits;                                    // block-wide exclusive scan, 64-bit
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint64_t before = carry_s;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (y < H) st[y].wsum_off = before + inc - bits;
    __syncthreads();
    if (threadIdx.x == kLayoutBlock - 1) carry_s = before + inc;
    __syncthreads();
  }
  // Adler-32 over the H*(W+1) filtered bytes:  A = 1 + sum d,  B = n + n * sum d - sum g * d_g   (mod 65521)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum_d += __shfl_xor_sync(0xffffffffu, sum_d, o);
    sum_g += __shfl_xor_sync(0xffffffffu, sum_g, o);
  }
  if (lane == 0) {
    red[0][warp] = sum_d;
    red[1][warp] = sum_g;
  }
  __syncthreads();
  const uint64_t end_bits = carry_s + kPngEobLen;           // + end-of-block code
  const uint64_t deflate_end = (end_bits + 7) >> 3;          // byte offset of the Adler-32 trailer
  const uint64_t total = deflate_end + 4;
  uint32_t* out32 = reinterpret_cast<uint32_t*>(out + img * capacity);
  const uint64_t words = (total + 3) >> 2;
  for (uint64_t i = threadIdx.x; i < words; i += kLayoutBlock) out32[i] = 0u;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t d = 0, g = 0;
    for (int w = 0; w < kLayoutBlock / 32; ++w) {
      d += red[0][w];
      g += red[1][w];
    }
    d %= kAdlerMod;
    g %= kAdlerMod;
    const uint64_t n = ((uint64_t)H * rowlen) % kAdlerMod;
    const uint32_t A = (uint32_t)((1 + d) % kAdlerMod);
    const uint32_t B = (uint32_t)((n + n * d % kAdlerMod + kAdlerMod - g) % kAdlerMod);
    uint8_t* o = out + img * capacity;
    out32[0] = 0x00000178u;                                 // 78 01
    for (int k = 0; k < (kPngHeaderBits + 31) / 32; ++k) {   // block header (constant bits) from bit 16
      const int nb = min(32, kPngHeaderBits - 32 * k);
      put_bits(out32, 16 + 32 * (uint64_t)k, kPngHeaderWords[k] & 0xffffu, min(nb, 16));
      if (nb > 16) put_bits(out32, 16 + 32 * (uint64_t)k + 16, kPngHeaderWords[k] >> 16, nb - 16);
    }
    put_bits(out32, carry_s, kPngEobPat, kPngEobLen);
    o[deflate_end + 0] = (uint8_t)(B >> 8);
    o[deflate_end + 1] = (uint8_t)(B & 0xff);
    o[deflate_end + 2] = (uint8_t)(A >> 8);
    o[deflate_end + 3] = (uint8_t)(A & 0xff);
    lengths[img] = (int64_t)total;
  }
}

}  // namespace diga

// ---------------------------------------------------------------------------------------------
// host side: CRC-32 (PNG 1.2 annex D polynomial, slice-by-8) and the file framing
// ---------------------------------------------------------------------------------------------
namespace {

struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static const CrcTables tab;                                // thread-safe initialisation (C++11 magic static)
  crc = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = tab.t[7][lo & 0xff] ^ tab.t[6][(lo >> 8) & 0xff] ^ tab.t[5][(lo >> 16) & 0xff] ^ tab.t[4][lo >> 24] ^
          tab.t[3][hi & 0xff] ^ tab.t[2][(hi >> 8) & 0xff] ^ tab.t[1][(hi >> 16) & 0xff] ^ tab.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = tab.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return ~crc;
}

void put_be32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24);
  p[1] = (uint8_t)(v >> 16);
  p[2] = (uint8_t)(v >> 8);
  p[3] = (uint8_t)v;
}

// `known_crc`: the chunk's CRC-32 when it was already computed (the IDAT chunk: on the GPU, diga_png_crc)
bool write_chunk(FILE* f, const char tag[4], const uint8_t* data, size_t n, const uint32_t* known_crc = nullptr) {
  uint8_t head[8], tail[4];
  put_be32(head, (uint32_t)n);
  memcpy(head + 4, tag, 4);
  uint32_t crc;
  if (known_crc) {
    crc = *known_crc;
  } else {
    crc = crc32_update(0, head + 4, 4);
    if (n) crc = crc32_update(crc, data, n);
  }
  put_be32(tail, crc);
  return fwrite(head, 1, 8, f) == 8 && (n == 0 || fwrite(data, 1, n, f) == n) && fwrite(tail, 1, 4, f) == 4;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// CRC-32 of the IDAT chunk ("IDAT" + zlib stream) on the GPU, so that the host only writes bytes.  CRC is linear over GF(2):
// with raw(M) = M(x) * x^32 mod P (register starts at 0, no final inversion) the register after a message is
//     crc32(M) = raw(M) ^ (0xffffffff * x^(8|M|) mod P) ^ 0xffffffff        and       raw(A || B) = raw(A) * x^(8|B|) ^ raw(B),
// and leading zero bytes do not change raw().  One CTA per image: the message is cut, FROM ITS END, into 256 ranges of R bytes
// (the first ranges may start before the message: virtual zeros), every thread runs the byte-wise table recurrence over its
// range, and a shared-memory tree combines the 256 remainders — every right operand of level k has length R * 2^k, so one
// multiplier x^(8 R 2^k) per level serves all pairs (multiplication mod P: zlib's multmodp, reflected bit order).
// ------------------------------------------------------------------------------------------------
namespace diga {

constexpr uint32_t kCrcPoly = 0xedb88320u;
__constant__ uint32_t kCrcX2n[32] = {   // x^(2^n) mod P, n = 0..31
    0x40000000u, 0x20000000u, 0x08000000u, 0x00800000u, 0x00008000u, 0xedb88320u, 0xb1e6b092u, 0xa06a2517u,
    0xed627daeu, 0x88d14467u, 0xd7bbfe6au, 0xec447f11u, 0x8e7ea170u, 0x6427800eu, 0x4d47bae0u, 0x09fe548fu,
    0x83852d0fu, 0x30362f1au, 0x7b5a9cc3u, 0x31fec169u, 0x9fec022au, 0x6c8dedc4u, 0x15d6874du, 0x5fde7a4eu,
    0xbad90e37u, 0x2e4e5eefu, 0x4eaba214u, 0xa8a472c0u, 0x429a969eu, 0x148d302au, 0xc40ba6d0u, 0xc4e22c3cu};

__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (uint32_t m = 0x80000000u;; m >>= 1) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    if (m == 1) break;
    b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
  }
  return p;
}
// x^(8 * nbytes) mod P
__device__ __forceinline__ uint32_t crc_x8n(uint64_t nbytes) {
  uint32_t p = 0x80000000u;
  for (unsigned k = 3; nbytes; nbytes >>= 1, ++k)
    if (nbytes & 1) p = crc_multmodp(kCrcX2n[k & 31], p);
  return p;
}

constexpr int kCrcBlock = 256;

__global__ void __launch_bounds__(kCrcBlock)
png_crc_kernel(const uint8_t* __restrict__ payload, int64_t capacity, const int64_t* __restrict__ lengths, uint32_t* __restrict__ crc_out) {
  __shared__ uint32_t table[256];
  __shared__ uint32_t part[kCrcBlock];
  const int t = threadIdx.x;
  {
    uint32_t c = (uint32_t)t;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
    table[t] = c;
  }
  __syncthreads();
  const int64_t img = blockIdx.x;
  const uint8_t* p = payload + img * capacity;
  const int64_t len = lengths[img];
  const int64_t total = (len < 0 ? 0 : len > capacity ? capacity : len) + 4;       // "IDAT" + stream (never past the image's slot)
  int64_t R = (total + kCrcBlock - 1) / kCrcBlock;
  R = (R + 3) & ~(int64_t)3;
  if (R < 4) R = 4;
  const int64_t lo = total - (int64_t)(kCrcBlock - t) * R, hi = lo + R;
  uint32_t c = 0;
  for (int64_t j = lo < 0 ? 0 : lo; j < hi; ++j) {
    const uint32_t byte = j < 4 ? (uint32_t)("IDAT"[j]) : (uint32_t)__ldg(p + (j - 4));
    c = table[(c ^ byte) & 0xffu] ^ (c >> 8);
  }
  part[t] = c;
  uint32_t op = crc_x8n((uint64_t)R);
  for (int n = kCrcBlock / 2; n >= 1; n >>= 1) {
    __syncthreads();
    uint32_t v = 0;
    if (t < n) v = crc_multmodp(op, part[2 * t]) ^ part[2 * t + 1];
    __syncthreads();
    if (t < n) part[t] = v;
    op = crc_multmodp(op, op);
  }
  if (t == 0) crc_out[img] = part[0] ^ crc_multmodp(crc_x8n((uint64_t)total), 0xffffffffu) ^ 0xffffffffu;
}

}  // namespace diga

extern "C" int diga_png_crc(const uint8_t* payload, int64_t n, int64_t capacity, const int64_t* lengths, uint32_t* crc_out,
                            diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(payload && lengths && crc_out, DIGA_ERR_INVALID, "png_crc: null pointer");
  DIGA_REQUIRE(n >= 0 && n <= 65535 && capacity >= 1, DIGA_ERR_INVALID, "png_crc: bad sizes");
  DIGA_REQUIRE(aligned(lengths, 8) && aligned(crc_out, 4), DIGA_ERR_MISALIGNED, "png_crc: misaligned pointer");
  if (n == 0) return DIGA_OK;
  png_crc_kernel<<<(unsigned)n, kCrcBlock, 0, (cudaStream_t)stream>>>(payload, capacity, lengths, crc_out);
  DIGA_CHECK_LAUNCH("png_crc_kernel");
  return DIGA_OK;
}

static int png_write_file_impl(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                               const uint8_t* palette_host, int64_t palette_bytes, const uint32_t* idat_crc);

extern "C" int diga_png_write_file(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                                   const uint8_t* palette_host, int64_t palette_bytes) {
  return png_write_file_impl(path, payload_host, length, H, W, palette_host, palette_bytes, nullptr);
}

extern "C" int diga_png_write_file_crc(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                                       const uint8_t* palette_host, int64_t palette_bytes, uint32_t idat_crc) {
  return png_write_file_impl(path, payload_host, length, H, W, palette_host, palette_bytes, &idat_crc);
}

static int png_write_file_impl(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                               const uint8_t* palette_host, int64_t palette_bytes, const uint32_t* idat_crc) {
  using namespace diga;
  DIGA_REQUIRE(path && payload_host && palette_host, DIGA_ERR_INVALID, "png_write_file: null pointer");
  DIGA_REQUIRE(length > 6 && length < (int64_t(1) << 31) && H >= 1 && W >= 1 && H < (int64_t(1) << 31) && W < (int64_t(1) << 31),
               DIGA_ERR_INVALID, "png_write_file: bad sizes");
  DIGA_REQUIRE(palette_bytes >= 3 && palette_bytes <= 768 && palette_bytes % 3 == 0, DIGA_ERR_INVALID,
               "png_write_file: palette of %lld bytes", (long long)palette_bytes);
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error("png_write_file: cannot open %s: %s", path, strerror(errno));
    return DIGA_ERR_IO;
  }
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  uint8_t ihdr[13];
  put_be32(ihdr, (uint32_t)W);
  put_be32(ihdr + 4, (uint32_t)H);
  ihdr[8] = 8;   // bit depth
  ihdr[9] = 3;   // colour type: palette
  ihdr[10] = ihdr[11] = ihdr[12] = 0;
  bool ok = fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr, 13) &&
            write_chunk(f, "PLTE", palette_host, (size_t)palette_bytes) &&
            write_chunk(f, "IDAT", payload_host, (size_t)length, idat_crc) && write_chunk(f, "IEND", nullptr, 0);
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    set_error("png_write_file: short write to %s: %s", path, strerror(errno));
    return DIGA_ERR_IO;
  }
  return DIGA_OK;
}

extern "C" int64_t diga_png_deflate_capacity(int64_t H, int64_t W) {
  if (H < 1 || W < 1) return 0;
  // worst case: every byte a literal of the longest code; + headers, end of block, trailer, rounded up to 16 bytes
  const int64_t bits = H * (W + 1) * diga::kPngMaxLitBits + 16 + diga::kPngHeaderBits + diga::kPngEobLen;
  return (((bits + 7) / 8 + 4 + 8) + 15) / 16 * 16;
}

extern "C" int64_t diga_png_deflate_scratch_bytes(int64_t n, int64_t H) {
  return n < 0 || H < 0 ? 0 : n * H * (int64_t)sizeof(diga::PngRowStat);
}

extern "C" int diga_png_deflate(const uint8_t* labels, int64_t n, int64_t H, int64_t W, uint8_t* out, int64_t capacity,
                                void* scratch, int64_t* lengths, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(labels && out && scratch && lengths, DIGA_ERR_INVALID, "png_deflate: null pointer");
  DIGA_REQUIRE(n >= 0 && H >= 1 && W >= 1 && W <= 65534 && H <= (1 << 24) && n <= 65535, DIGA_ERR_INVALID,
               "png_deflate: bad sizes n=%lld H=%lld W=%lld", (long long)n, (long long)H, (long long)W);
  DIGA_REQUIRE(capacity >= diga_png_deflate_capacity(H, W) && capacity % 4 == 0, DIGA_ERR_INVALID,
               "png_deflate: capacity %lld below diga_png_deflate_capacity", (long long)capacity);
  DIGA_REQUIRE(aligned(out, 4) && aligned(scratch, 8) && aligned(lengths, 8), DIGA_ERR_MISALIGNED, "png_deflate: misaligned pointer");
  if (n == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PngRowStat* stats = static_cast<PngRowStat*>(scratch);
  const int rowlen = (int)W + 1;
  const int per_warp = (((rowlen + 3) & ~3) + 2 * (rowlen + 1) + 15) & ~15;
  int warps = 8;
  while (warps > 1 && (size_t)warps * per_warp > 200 * 1024) warps >>= 1;
  const size_t smem = (size_t)warps * per_warp;
  DIGA_REQUIRE(smem <= 200 * 1024, DIGA_ERR_INVALID, "png_deflate: W=%lld needs %zu bytes of shared memory", (long long)W, smem);
  static size_t configured_dev[64] = {0};            // the attribute is per device
  size_t& configured = configured_dev[device_slot()];
  if (smem > 48 * 1024 && configured < smem) {
    if (cudaFuncSetAttribute(png_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(png_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("png_deflate: cannot reserve %zu bytes of shared memory", smem);
      return DIGA_ERR_CUDA;
    }
    configured = smem;
  }
  const dim3 grid((unsigned)((H + warps - 1) / warps), (unsigned)n);
  png_rows_kernel<false><<<grid, warps * 32, smem, st>>>(labels, (int)H, (int)W, stats, nullptr, capacity, per_warp);
  DIGA_CHECK_LAUNCH("png_rows_kernel<count>");
  png_layout_kernel<<<(unsigned)n, kLayoutBlock, 0, st>>>(stats, (int)H, (int)W, out, capacity, lengths);
  DIGA_CHECK_LAUNCH("png_layout_kernel");
  png_rows_kernel<true><<<grid, warps * 32, smem, st>>>(labels, (int)H, (int)W, stats, out, capacity, per_warp);
  DIGA_CHECK_LAUNCH("png_rows_kernel<write>");
  return DIGA_OK;
}
