// PNG decode on the device ("decode -> crop -> resize -> ..."): the two stages of
// LoadImageFromFile's cv2.imdecode (mmdet/datasets/pipelines/loading.py:58-69) for 8-bit non-interlaced PNG files, the
// on-disk format of data/gaze360/*_rawframes (tools/gaze360_img_reorganize.py:108,137 writes them with cv2.imwrite):
//   1. inflate (RFC 1950 / 1951) of the concatenated IDAT payload into the filtered scanlines,
//   2. scanline reconstruction (PNG spec section 9: None / Sub / Up / Average / Paeth) + the conversion
//      cv2.imdecode(IMREAD_COLOR) applies: gray -> replicated, palette -> RGB, alpha dropped, RGB -> BGR.
// One WARP per image.  Inflate is a serial bit stream: lane 0 decodes symbols (primary lookup tables in shared memory,
// canonical bit-by-bit decoding for the rare codes longer than the table index), literals are stored by lane 0, every
// LZ77 match is copied by the whole warp.  Reconstruction runs as a 32-row wavefront: lane k owns row 32*band + k and
// trails lane k-1 by one byte, so the byte above arrives by one __shfl_up per step and Paeth / Average rows cost the same
// as Sub rows.
//
// The file compiles twice: by nvcc as device code, and by g++ with MCG_PNG_HOST_SIM for the CPU tests
// (tests/test_png.py): there a "warp" is emulated lane by lane in lock step for the wavefront and is a single lane for
// inflate, so the table construction and the bit reader are exercised without a GPU.
#pragma once
#include <stdint.h>

#ifdef MCG_PNG_HOST_SIM
#define MCG_PNG_FN static inline
#define MCG_PNG_RARE static
#define MCG_PNG_LANES 1
#else
#define MCG_PNG_FN __device__ __forceinline__
#define MCG_PNG_RARE __device__ __noinline__   // rare paths stay out of the hot loop's code
#define MCG_PNG_LANES 32
#endif

namespace mcg {
namespace png {

enum Status : int32_t {
  ST_OK = 0,
  ST_BAD_ZLIB_HEADER = 1,
  ST_BAD_BLOCK_TYPE = 2,
  ST_BAD_STORED_LEN = 3,
  ST_BAD_CODE_LENGTHS = 4,
  ST_BAD_SYMBOL = 5,
  ST_BAD_DISTANCE = 6,
  ST_OUTPUT_OVERFLOW = 7,
  ST_INPUT_EXHAUSTED = 8,
  ST_OUTPUT_SHORT = 9,
  ST_BAD_FILTER = 10,
  ST_BAD_JOB = 11,
  ST_BAD_CHECKSUM = 12,
};

constexpr int kLitBits = 10;   // primary table index widths: a literal/length code has <= 15 bits, codes longer than
constexpr int kDistBits = 8;   // the index (probability <= 2^-10 / 2^-8 per symbol) take the canonical slow path
constexpr int kMaxLit = 288, kMaxDist = 32, kMaxBits = 15;

// per-image decoding tables (shared memory on the device)
struct alignas(1024) Tables {
  // first, and 1 KB aligned: the fast loop forms their addresses as (window & 1023) | base, one LOP3 on the lookup chain
  uint8_t lit_len[1 << kLitBits]; // the literal fast loop's view of `lit`: code length, or 0x80 for anything that is not
  uint8_t lit_sym[1 << kLitBits]; // a literal with a code of <= kLitBits bits; and the literal itself
  uint16_t lit[1 << kLitBits];    // (symbol << 4) | code length ; 0 = not in the table
  uint16_t dist[1 << kDistBits];
  uint16_t lit_sorted[kMaxLit];   // symbols ordered by (length, symbol): canonical decoding
  uint16_t dist_sorted[kMaxDist];
  uint16_t lit_count[kMaxBits + 1];
  uint16_t dist_count[kMaxBits + 1];
  uint16_t code[kMaxLit];         // scratch: canonical code of every symbol
  uint8_t lens[kMaxLit + kMaxDist];
};

#ifdef MCG_PNG_HOST_SIM
static long long g_slow_symbols = 0;   // symbols that took the canonical path (tests assert that it is exercised)
MCG_PNG_FN void warp_sync() {}
MCG_PNG_FN int bcast(int v) { return v; }
MCG_PNG_FN long long bcastll(long long v) { return v; }
MCG_PNG_FN unsigned long long warp_sum(unsigned long long v) { return v; }
MCG_PNG_FN uint8_t load_cg(const uint8_t* p) { return *p; }
MCG_PNG_FN uint32_t load_word(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
MCG_PNG_FN uint32_t load_word_early(const uint8_t* p) { return load_word(p); }
MCG_PNG_FN uint32_t shr_wrap(uint32_t v, uint32_t n) { return v >> (n & 31u); }
struct FastTabs {
  const uint8_t* len;
  const uint8_t* sym;
};
MCG_PNG_FN FastTabs fast_tabs(const Tables& T) { return FastTabs{T.lit_len, T.lit_sym}; }
MCG_PNG_FN uint32_t fast_len(const FastTabs& f, uint32_t win) { return f.len[win & ((1u << kLitBits) - 1u)]; }
MCG_PNG_FN uint32_t fast_sym(const FastTabs& f, uint32_t win) { return f.sym[win & ((1u << kLitBits) - 1u)]; }
MCG_PNG_FN void prefetch(const uint8_t*) {}
MCG_PNG_FN uint32_t brev32(uint32_t v) {
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
#else
MCG_PNG_FN void warp_sync() { __syncwarp(); }
MCG_PNG_FN int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }
MCG_PNG_FN long long bcastll(long long v) { return __shfl_sync(0xffffffffu, v, 0); }
MCG_PNG_FN unsigned long long warp_sum(unsigned long long v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// bytes another lane of this warp has just written: read at L2, never from a stale L1 line
MCG_PNG_FN uint8_t load_cg(const uint8_t* p) { return __ldcg(p); }
MCG_PNG_FN uint32_t load_word(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
// the same load pinned where it is written (the fast loop issues the next pass's word early; left to itself the
// compiler sinks the load to its use and puts the L1 latency in front of every pass)
MCG_PNG_FN uint32_t load_word_early(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// v >> (n mod 32) in one instruction (the 0x80 of a non-literal shifts by 0)
MCG_PNG_FN uint32_t shr_wrap(uint32_t v, uint32_t n) { return __funnelshift_r(v, 0u, n); }
struct FastTabs {
  uint32_t len, sym;   // shared-space addresses, multiples of 1024
};
MCG_PNG_FN FastTabs fast_tabs(const Tables& T) {
  return FastTabs{static_cast<uint32_t>(__cvta_generic_to_shared(T.lit_len)), static_cast<uint32_t>(__cvta_generic_to_shared(T.lit_sym))};
}
MCG_PNG_FN uint32_t fast_len(const FastTabs& f, uint32_t win) {
  uint32_t a, v;
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(a) : "r"(win), "r"((1u << kLitBits) - 1u), "r"(f.len));   // (win & 1023) | base
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
MCG_PNG_FN uint32_t fast_sym(const FastTabs& f, uint32_t win) {
  uint32_t a, v;
  asm("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(a) : "r"(win), "r"((1u << kLitBits) - 1u), "r"(f.sym));
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
MCG_PNG_FN void prefetch(const uint8_t* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
MCG_PNG_FN uint32_t brev32(uint32_t v) { return __brev(v); }
#endif

// LSB-first bit reader over a byte stream (lane 0 only).  `p` is the next byte to fetch; it is 4-byte aligned whenever
// at least 4 bytes remain (br_align), so the steady-state refill is one 32-bit load.  Reads past the end deliver zeros
// and are detected by br_consumed() > n afterwards.
struct BitReader {
  const uint8_t* in;
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf;
  int cnt;
};
// bytes until `p` is aligned (needs cnt <= 32 on entry)
MCG_PNG_FN void br_align(BitReader& b) {
  while ((reinterpret_cast<uintptr_t>(b.p) & 3) != 0 && b.p < b.end) {
    b.buf |= static_cast<uint64_t>(*b.p++) << b.cnt;
    b.cnt += 8;
  }
}
MCG_PNG_FN void br_init(BitReader& b, const uint8_t* in, long long n) {
  b.in = in;
  b.p = in;
  b.end = in + n;
  b.buf = 0;
  b.cnt = 0;
  br_align(b);
}
// the last bytes of the stream (and zeros behind it)
MCG_PNG_RARE void br_refill_tail(BitReader& b) {
  while (b.cnt < 32) {
    const uint32_t v = b.p < b.end ? *b.p : 0u;
    ++b.p;
    b.buf |= static_cast<uint64_t>(v) << b.cnt;
    b.cnt += 8;
  }
}
// at least 32 valid bits afterwards
MCG_PNG_FN void br_refill(BitReader& b) {
  if (b.cnt < 32) {
    if (b.p + 4 <= b.end) {
      b.buf |= static_cast<uint64_t>(load_word(b.p)) << b.cnt;
      b.cnt += 32;
      if ((reinterpret_cast<uintptr_t>(b.p) & 127) == 0) prefetch(b.p + 256);
      b.p += 4;
    } else {
      br_refill_tail(b);
    }
  }
}
MCG_PNG_FN uint32_t br_peek(const BitReader& b, int nbits) { return static_cast<uint32_t>(b.buf) & ((1u << nbits) - 1u); }
MCG_PNG_FN void br_skip(BitReader& b, int nbits) {
  b.buf >>= nbits;
  b.cnt -= nbits;
}
MCG_PNG_FN uint32_t br_get(BitReader& b, int nbits) {  // nbits <= 16, caller keeps cnt >= nbits
  const uint32_t v = br_peek(b, nbits);
  br_skip(b, nbits);
  return v;
}
MCG_PNG_FN long long br_consumed(const BitReader& b) { return static_cast<long long>(b.p - b.in) - (b.cnt >> 3); }
// continue at byte offset `at` (behind a stored block)
MCG_PNG_FN void br_seek(BitReader& b, long long at) {
  b.p = b.in + at;
  b.buf = 0;
  b.cnt = 0;
  br_align(b);
}

// Canonical Huffman tables for `n` symbols with code lengths lens[0..n) (RFC 1951 3.2.2).  Called by every lane.
// Returns false for an over-subscribed set of lengths.  Incomplete sets are accepted (a stream that then uses a
// missing code fails with ST_BAD_SYMBOL), as zlib does for a single distance code.
MCG_PNG_FN bool build_table(const uint8_t* lens, int n, uint16_t* tab, int tab_bits, uint16_t* sorted, uint16_t* count,
                            uint16_t* code, int lane) {
  warp_sync();
  int ok = 1;
  if (lane == 0) {
    for (int l = 0; l <= kMaxBits; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) count[lens[s]]++;
    int left = 1;
    for (int l = 1; l <= kMaxBits; ++l) {
      left <<= 1;
      left -= count[l];
      if (left < 0) ok = 0;
    }
    if (ok) {
      uint16_t next[kMaxBits + 2], offs[kMaxBits + 2];
      uint32_t c = 0;
      offs[1] = 0;
      for (int l = 1; l <= kMaxBits; ++l) {
        c = (c + (l > 1 ? count[l - 1] : 0)) << 1;
        next[l] = static_cast<uint16_t>(c);
        offs[l + 1] = offs[l] + count[l];
      }
      for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l) {
          code[s] = next[l]++;
          sorted[offs[l]++] = static_cast<uint16_t>(s);
        }
      }
    }
  }
  ok = bcast(ok);
  if (!ok) return false;
  for (int i = lane; i < (1 << tab_bits); i += MCG_PNG_LANES) tab[i] = 0;
  warp_sync();
  // a code of length l <= tab_bits owns every index whose low l bits are the bit-reversed code
  for (int s = lane; s < n; s += MCG_PNG_LANES) {
    const int l = lens[s];
    if (l && l <= tab_bits) {
      const uint32_t r = brev32(code[s]) >> (32 - l);
      const uint16_t e = static_cast<uint16_t>((s << 4) | l);
      for (uint32_t i = r; i < (1u << tab_bits); i += (1u << l)) tab[i] = e;
    }
  }
  warp_sync();
  return true;
}

// a code longer than the primary table's index: bit by bit over the canonical code (zlib's puff.c formulation)
MCG_PNG_RARE int decode_symbol_slow(BitReader& b, const uint16_t* sorted, const uint16_t* count) {
#ifdef MCG_PNG_HOST_SIM
  ++g_slow_symbols;
#endif
  int code = 0, first = 0, index = 0;
  uint32_t bits = static_cast<uint32_t>(b.buf);
  for (int l = 1; l <= kMaxBits; ++l) {
    code |= static_cast<int>(bits & 1u);
    bits >>= 1;
    const int c = count[l];
    if (code - c < first) {
      br_skip(b, l);
      return sorted[index + (code - first)];
    }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}
// one symbol (lane 0): primary lookup, else the canonical path
MCG_PNG_FN int decode_symbol(BitReader& b, const uint16_t* tab, int tab_bits, const uint16_t* sorted, const uint16_t* count) {
  const uint32_t e = tab[br_peek(b, tab_bits)];
  if (e) {
    br_skip(b, e & 15);
    return static_cast<int>(e >> 4);
  }
  return decode_symbol_slow(b, sorted, count);
}

// the whole warp copies an LZ77 match: out[pos + i] = out[pos - dist + i], bytes repeat with period dist when dist < len
MCG_PNG_FN void copy_match(uint8_t* out, long long pos, int len, int dist, int lane) {
  warp_sync();
  const uint8_t* src = out + pos - dist;
  if (dist == 1) {
    const uint8_t v = load_cg(src);
    for (int i = lane; i < len; i += MCG_PNG_LANES) out[pos + i] = v;
  } else if (dist >= len) {
    for (int i = lane; i < len; i += MCG_PNG_LANES) out[pos + i] = load_cg(src + i);
  } else {
    for (int i = lane; i < len; i += MCG_PNG_LANES) out[pos + i] = load_cg(src + i % dist);
  }
  warp_sync();
}

// Adler-32 (RFC 1950) of out[0..n) by the whole warp: a = 1 + sum d_i, b = n + sum (n - i) d_i, both mod 65521.
// Lane l takes bytes l, l + 32, ...; 64-bit sums hold any image (terms <= 255 n, n / 32 of them per lane).
MCG_PNG_FN uint32_t adler32_warp(const uint8_t* out, long long n, int lane) {
  warp_sync();
  unsigned long long sa = 0, sb = 0;
  for (long long i = lane; i < n; i += MCG_PNG_LANES) {
    const unsigned long long d = load_cg(out + i);
    sa += d;
    sb += d * static_cast<unsigned long long>(n - i);
    if ((i & 0xfffffLL) < MCG_PNG_LANES) sb %= 65521ull;   // every 2^20 bytes: keeps sb far from 2^64 for any n
  }
  sa = warp_sum(sa);
  sb = warp_sum(sb % 65521ull);
  const uint32_t ra = static_cast<uint32_t>((1ull + sa) % 65521ull);
  const uint32_t rb = static_cast<uint32_t>((static_cast<unsigned long long>(n % 65521) + sb) % 65521ull);
  return (rb << 16) | ra;
}

// zlib stream `in[0..n)` -> out[0..cap).  Every lane of the warp calls it with the same arguments; returns the status
// (uniform) and the number of bytes produced.  The Adler-32 trailer is verified against the produced bytes (the one
// integrity check of the whole path when the host skips the chunk CRCs).
MCG_PNG_FN int inflate_warp(const uint8_t* in, long long n, uint8_t* out, long long cap, Tables& T, int lane,
                            long long* produced) {
#ifndef MCG_PNG_HOST_SIM
  // the arguments come from an indexed kernel-parameter table: without this the compiler re-reads them from the
  // constant bank inside the symbol loop instead of keeping them in registers
  asm volatile("" : "+l"(in), "+l"(n), "+l"(out), "+l"(cap));
  __builtin_assume(__isGlobal(in));
  __builtin_assume(__isGlobal(out));
#endif
  BitReader br;
  br_init(br, in, n);
  long long pos = 0;
  int st = ST_OK;
  if (lane == 0) {
    br_refill(br);
    const uint32_t cmf = br_get(br, 8), flg = br_get(br, 8);
    if ((cmf & 15) != 8 || (cmf >> 4) > 7 || (flg & 32) || ((cmf << 8) | flg) % 31 != 0) st = ST_BAD_ZLIB_HEADER;
  }
  st = bcast(st);
  int final_block = 0;
  while (st == ST_OK && !final_block) {
    int btype = 0, hlit = 0, hdist = 0;
    if (lane == 0) {
      br_refill(br);
      final_block = static_cast<int>(br_get(br, 1));
      btype = static_cast<int>(br_get(br, 2));
      if (btype == 3) st = ST_BAD_BLOCK_TYPE;
    }
    st = bcast(st);
    if (st != ST_OK) break;
    final_block = bcast(final_block);
    btype = bcast(btype);
    if (btype == 0) {
      // stored block: LEN / NLEN at the next byte boundary, then LEN raw bytes
      long long src = 0;
      int len = 0;
      if (lane == 0) {
        br_skip(br, br.cnt & 7);
        br_refill(br);
        len = static_cast<int>(br_get(br, 16));
        const int nlen = static_cast<int>(br_get(br, 16));
        src = br_consumed(br);
        if ((len ^ nlen) != 0xffff) st = ST_BAD_STORED_LEN;
        else if (src + len > n) st = ST_INPUT_EXHAUSTED;
        else if (pos + len > cap) st = ST_OUTPUT_OVERFLOW;
      }
      st = bcast(st);
      if (st != ST_OK) break;
      len = bcast(len);
      src = bcastll(src);
      warp_sync();
      for (int i = lane; i < len; i += MCG_PNG_LANES) out[pos + i] = in[src + i];
      warp_sync();
      pos += len;
      if (lane == 0) br_seek(br, src + len);  // restart the bit reader behind the raw bytes
      continue;
    }
    if (btype == 1) {
      // fixed code (RFC 1951 3.2.6)
      for (int s = lane; s < kMaxLit; s += MCG_PNG_LANES) T.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
      for (int s = lane; s < 30; s += MCG_PNG_LANES) T.lens[kMaxLit + s] = 5;
      hlit = kMaxLit;
      hdist = 30;
    } else {
      // dynamic code: the code lengths themselves are Huffman coded (3.2.7); lane 0 reads them
      if (lane == 0) {
        br_refill(br);
        hlit = static_cast<int>(br_get(br, 5)) + 257;
        hdist = static_cast<int>(br_get(br, 5)) + 1;
        const int hclen = static_cast<int>(br_get(br, 4)) + 4;
        if (hlit > 286 || hdist > 30) st = ST_BAD_CODE_LENGTHS;
        uint8_t cl[19];
        for (int i = 0; i < 19; ++i) cl[i] = 0;
        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        for (int i = 0; i < hclen; ++i) {
          br_refill(br);
          cl[order[i]] = static_cast<uint8_t>(br_get(br, 3));
        }
        // code-length code: 19 symbols, <= 7 bits; a direct 128-entry table in the (not yet used) distance table
        uint16_t* ct = T.dist;
        {
          int count[8], left = 1;
          for (int l = 0; l < 8; ++l) count[l] = 0;
          for (int s = 0; s < 19; ++s) count[cl[s]]++;
          for (int l = 1; l < 8; ++l) {
            left <<= 1;
            left -= count[l];
            if (left < 0) st = ST_BAD_CODE_LENGTHS;
          }
          uint32_t next[9], c = 0;
          for (int l = 1; l < 8; ++l) {
            c = (c + (l > 1 ? count[l - 1] : 0)) << 1;
            next[l] = c;
          }
          for (int i = 0; i < 128; ++i) ct[i] = 0;
          if (st == ST_OK)
            for (int s = 0; s < 19; ++s) {
              const int l = cl[s];
              if (!l) continue;
              const uint32_t r = brev32(next[l]++) >> (32 - l);
              for (uint32_t i = r; i < 128; i += (1u << l)) ct[i] = static_cast<uint16_t>((s << 4) | l);
            }
        }
        int i = 0;
        const int total = hlit + hdist;
        while (st == ST_OK && i < total) {
          br_refill(br);
          const uint32_t e = ct[br_peek(br, 7)];
          if (!e) {
            st = ST_BAD_CODE_LENGTHS;
            break;
          }
          br_skip(br, e & 15);
          const int sym = static_cast<int>(e >> 4);
          if (sym < 16) {
            T.lens[i++] = static_cast<uint8_t>(sym);
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (i == 0) {
                st = ST_BAD_CODE_LENGTHS;
                break;
              }
              val = T.lens[i - 1];
              rep = 3 + static_cast<int>(br_get(br, 2));
            } else if (sym == 17) {
              rep = 3 + static_cast<int>(br_get(br, 3));
            } else {
              rep = 11 + static_cast<int>(br_get(br, 7));
            }
            if (i + rep > total) {
              st = ST_BAD_CODE_LENGTHS;
              break;
            }
            while (rep--) T.lens[i++] = static_cast<uint8_t>(val);
          }
        }
        if (st == ST_OK && T.lens[256] == 0) st = ST_BAD_CODE_LENGTHS;  // no end-of-block code
        if (st == ST_OK) {
          // the tables take the literal/length lengths at lens[0..hlit) and the distance lengths at lens[288..)
          for (int d = hdist - 1; d >= 0; --d) T.lens[kMaxLit + d] = T.lens[hlit + d];
          for (int s = hlit; s < kMaxLit; ++s) T.lens[s] = 0;
        }
      }
      st = bcast(st);
      if (st != ST_OK) break;
      hlit = kMaxLit;
      hdist = bcast(hdist);
    }
    warp_sync();
    if (!build_table(T.lens, hlit, T.lit, kLitBits, T.lit_sorted, T.lit_count, T.code, lane) ||
        !build_table(T.lens + kMaxLit, hdist, T.dist, kDistBits, T.dist_sorted, T.dist_count, T.code, lane)) {
      st = ST_BAD_CODE_LENGTHS;
      break;
    }
    for (int i = lane; i < (1 << kLitBits); i += MCG_PNG_LANES) {
      const uint32_t e = T.lit[i];
      T.lit_len[i] = (e - 1u < 0xfffu) ? static_cast<uint8_t>(e & 15u) : static_cast<uint8_t>(0x80);
      T.lit_sym[i] = static_cast<uint8_t>(e >> 4);
    }
    warp_sync();
    // symbols: lane 0 decodes and stores literals until it meets a match, which the warp copies together
    for (;;) {
      int ev = 0, len = 0, dist = 0;  // ev: 0 = match, 1 = end of block, 2 = error (st set)
      if (lane == 0) {
        // The literal path is the hot loop (a photograph is ~1 literal per output byte): refill test, one table
        // lookup, one 64-bit shift, one byte store.  Its dependent chain (mask -> LDS -> length -> shift) is what
        // bounds the decoder, so everything else (bounds checks, pointer increments) is kept off that chain.
        uint8_t* op = out + pos;
        uint8_t* const oend = out + cap;
        // the reader's state in registers for the loop; `br` itself is memory (the rare paths are calls that take it)
        uint64_t buf = br.buf;
        int cnt = br.cnt;
        const uint8_t* ip = br.p;
        const uint8_t* const iend = br.end;
        const FastTabs ft = fast_tabs(T);
        for (;;) {
          // Fast loop over runs of literals, three per pass.  One warp per scheduler cannot hide any latency, so the
          // loop is bound by its instruction count (~4 cycles each) as much as by the lookup chain:
          //  * the pass count is fixed up front from what is left of the stream and of the output (a pass reads at most
          //    one word and writes at most three bytes), so there is one counter and no pointer compare per pass;
          //  * the refill is branch-free (the next word is merged under a predicate; it was loaded a refill earlier);
          //  * after the refill the low 32 bits of the buffer hold the 30 bits three primary-table codes can take, so
          //    the three lookups shift a 32-bit window and the 64-bit buffer is shifted once per pass;
          //  * lit_len / lit_sym give "length, or 0x80 = not a short literal" and the literal as bytes: nothing is
          //    consumed before the test.
          // Long codes, matches, end of block and the last bytes drop to the general step below (one symbol) and come back.
          {
            long long trips = (iend - ip) / 4 - 3;
            const long long out_trips = (oend - op) / 3;
            trips = trips < out_trips ? trips : out_trips;
            int left = trips > (1 << 30) ? (1 << 30) : static_cast<int>(trips);
            if (left > 0) {
              uint32_t wa = load_word_early(ip);          // the next two words of the stream
              uint32_t wb = load_word_early(ip + 4);
              for (; left > 0; --left) {
                const bool need = cnt < 32;
                buf |= need ? static_cast<uint64_t>(wa) << cnt : 0ull;   // cnt <= 31 when it counts
                ip += need ? 4 : 0;
                cnt += need ? 32 : 0;
                wa = need ? wb : wa;
                wb = load_word_early(ip + 4);                           // (the same word again when nothing was consumed)
                // three lookups without a branch between them (a branch costs a single warp ~13 cycles even when it
                // falls through): a symbol that is not a short literal has length code 0x80, shifts the window by 0 and
                // switches the rest of the pass off
                const uint32_t win = static_cast<uint32_t>(buf);
                const uint32_t l1 = fast_len(ft, win), s1 = fast_sym(ft, win);
                const uint32_t w2 = shr_wrap(win, l1);
                const uint32_t l2 = fast_len(ft, w2), s2 = fast_sym(ft, w2);
                const uint32_t w3 = shr_wrap(w2, l2);
                const uint32_t l3 = fast_len(ft, w3), s3 = fast_sym(ft, w3);
                const bool ok1 = (l1 & 0x80u) == 0;
                const bool ok2 = ok1 && (l2 & 0x80u) == 0;
                const bool ok3 = ok2 && (l3 & 0x80u) == 0;
                if (ok1) op[0] = static_cast<uint8_t>(s1);
                if (ok2) op[1] = static_cast<uint8_t>(s2);
                if (ok3) op[2] = static_cast<uint8_t>(s3);
                const int used = static_cast<int>((ok1 ? l1 : 0u) + (ok2 ? l2 : 0u) + (ok3 ? l3 : 0u));
                const int nout = (ok1 ? 1 : 0) + (ok2 ? 1 : 0) + (ok3 ? 1 : 0);
                const bool stop = !ok3;
                buf >>= used;
                cnt -= used;
                op += nout;
                if (stop) break;
              }
            }
          }
          if (cnt < 32) {
            if (ip + 4 <= iend) {
              buf |= static_cast<uint64_t>(load_word(ip)) << cnt;
              cnt += 32;
              ip += 4;
            } else {
              br.buf = buf, br.cnt = cnt, br.p = ip;
              br_refill_tail(br);
              buf = br.buf, cnt = br.cnt, ip = br.p;
            }
          }
          int sym;
          const uint32_t e = T.lit[static_cast<uint32_t>(buf) & ((1u << kLitBits) - 1u)];
          if (e) {
            const int l = static_cast<int>(e & 15u);
            buf >>= l;
            cnt -= l;
            sym = static_cast<int>(e >> 4);
          } else {
            br.buf = buf, br.cnt = cnt, br.p = ip;
            sym = decode_symbol_slow(br, T.lit_sorted, T.lit_count);
            buf = br.buf, cnt = br.cnt, ip = br.p;
            if (sym < 0) {
              st = ST_BAD_SYMBOL;
              ev = 2;
              break;
            }
          }
          if (sym < 256) {
            if (op >= oend) {
              st = ST_OUTPUT_OVERFLOW;
              ev = 2;
              break;
            }
            *op++ = static_cast<uint8_t>(sym);
            continue;
          }
          br.buf = buf, br.cnt = cnt, br.p = ip;        // everything below is per match / per block: the generic reader
          if (sym == 256) {
            ev = 1;
            break;
          }
          sym -= 257;
          if (sym >= 29) {
            st = ST_BAD_SYMBOL;
            ev = 2;
            break;
          }
          // length: codes 257..264 -> 3..10, then groups of four codes per extra bit, 285 -> 258
          if (sym < 8) {
            len = 3 + sym;
          } else if (sym == 28) {
            len = 258;
          } else {
            const int eb = (sym >> 2) - 1;
            len = 3 + ((4 + (sym & 3)) << eb) + static_cast<int>(br_get(br, eb));
          }
          br_refill(br);
          const int ds = decode_symbol(br, T.dist, kDistBits, T.dist_sorted, T.dist_count);
          if (ds < 0 || ds >= 30) {
            st = ST_BAD_SYMBOL;
            ev = 2;
            break;
          }
          if (ds < 4) {
            dist = 1 + ds;
          } else {
            const int eb = (ds >> 1) - 1;
            dist = 1 + ((2 + (ds & 1)) << eb) + static_cast<int>(br_get(br, eb));
          }
          if (dist > op - out) {
            st = ST_BAD_DISTANCE;
            ev = 2;
          } else if (len > oend - op) {
            st = ST_OUTPUT_OVERFLOW;
            ev = 2;
          }
          break;
        }
        pos = op - out;
        if (ev != 2 && br_consumed(br) > n) {
          st = ST_INPUT_EXHAUSTED;
          ev = 2;
        }
      }
      ev = bcast(ev);
      if (ev != 0) break;
      len = bcast(len);
      dist = bcast(dist);
      pos = bcastll(pos);
      copy_match(out, pos, len, dist, lane);
      pos += len;
    }
    st = bcast(st);
    pos = bcastll(pos);
  }
  st = bcast(st);
  pos = bcastll(pos);
  *produced = pos;
  if (st != ST_OK) return st;
  int want = 0;
  if (lane == 0) {
    br_skip(br, br.cnt & 7);              // the trailer starts at the next byte boundary, most significant byte first
    br_refill(br);
    uint32_t w = 0;
    for (int k = 0; k < 4; ++k) w = (w << 8) | br_get(br, 8);
    want = static_cast<int>(w);
    if (br_consumed(br) > n) st = ST_INPUT_EXHAUSTED;
  }
  st = bcast(st);
  if (st != ST_OK) return st;
  want = bcast(want);
  return adler32_warp(out, pos, lane) == static_cast<uint32_t>(want) ? ST_OK : ST_BAD_CHECKSUM;
}

// ---------------------------------------------------------------------------------------------------------------
// Scanline reconstruction, one lane per row of a 32-row band.
// ---------------------------------------------------------------------------------------------------------------
struct LaneState {
  uint32_t hist_a;   // the last four reconstructed bytes of this row (newest in the low byte)
  uint32_t hist_b;   // the last four bytes of the row above
  uint32_t last;     // this row's newest reconstructed byte (what the lane below reads next step)
  int ch, px;        // channel / pixel of the next byte
};

MCG_PNG_FN uint32_t paeth(uint32_t a, uint32_t b, uint32_t c) {
  const int p = static_cast<int>(a) + static_cast<int>(b) - static_cast<int>(c);
  int pa = p - static_cast<int>(a), pb = p - static_cast<int>(b), pc = p - static_cast<int>(c);
  pa = pa < 0 ? -pa : pa;
  pb = pb < 0 ? -pb : pb;
  pc = pc < 0 ? -pc : pc;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// channels of a scanline pixel for a colour type (bit depth 8); 0 = not a PNG colour type
MCG_PNG_FN int channels_of(int color_type) {
  return color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 3 ? 1 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0;
}

// one byte of one row: reconstruct (PNG spec 9.2) and write what cv2.imdecode(IMREAD_COLOR) would hold for it
MCG_PNG_FN void unfilter_byte(LaneState& s, int ft, uint32_t raw, uint32_t up, int bpp, int color_type,
                              const uint8_t* palette, uint8_t* dst_row) {
  const int sh = 8 * (bpp - 1);
  const uint32_t a = (s.hist_a >> sh) & 255u, c = (s.hist_b >> sh) & 255u, b = up;
  const uint32_t pred = ft == 0 ? 0u : ft == 1 ? a : ft == 2 ? b : ft == 3 ? ((a + b) >> 1) : paeth(a, b, c);
  const uint32_t x = (raw + pred) & 255u;
  s.hist_a = (s.hist_a << 8) | x;
  s.hist_b = (s.hist_b << 8) | b;
  s.last = x;
  uint8_t* d = dst_row + 3 * s.px;
  const uint8_t v = static_cast<uint8_t>(x);
  if (color_type == 2 || color_type == 6) {
    if (s.ch < 3) d[2 - s.ch] = v;          // RGB(A) -> BGR, alpha dropped
  } else if (color_type == 3) {
    d[0] = palette[3 * x + 2];
    d[1] = palette[3 * x + 1];
    d[2] = palette[3 * x];
  } else if (s.ch == 0) {                   // gray (+ alpha): replicated
    d[0] = v;
    d[1] = v;
    d[2] = v;
  }
  if (++s.ch == bpp) {
    s.ch = 0;
    ++s.px;
  }
}

}  // namespace png
}  // namespace mcg
