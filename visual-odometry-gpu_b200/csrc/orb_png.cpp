// orb_png.cpp -- see orb_png.h.  Host code only (the decode threads of the ingest stage run it).
#include "orb_png.h"

#include <cstdlib>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace orbpng {
namespace {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t le64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // x86-64 / aarch64 little endian

// ---- checksums -------------------------------------------------------------------------------
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
      for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
const CrcTables& crc_tables() { static const CrcTables T; return T; }

}  // namespace

uint32_t crc32(const uint8_t* p, size_t n, uint32_t crc) {
  const CrcTables& T = crc_tables();
  uint32_t c = ~crc;
  while (n >= 8) {   // slicing-by-8
    const uint64_t v = le64(p) ^ c;
    c = T.t[7][v & 0xff] ^ T.t[6][(v >> 8) & 0xff] ^ T.t[5][(v >> 16) & 0xff] ^ T.t[4][(v >> 24) & 0xff] ^
        T.t[3][(v >> 32) & 0xff] ^ T.t[2][(v >> 40) & 0xff] ^ T.t[1][(v >> 48) & 0xff] ^ T.t[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

uint32_t adler32(const uint8_t* p, size_t n, uint32_t adler) {
  uint32_t a = adler & 0xffff, b = adler >> 16;
  while (n) {
    // 5552 is the largest block for which b cannot overflow 32 bits before the modulo
    size_t blk = n < 5552 ? n : 5552;
    n -= blk;
    // per 16-byte group: a += sum(x_i), b += 16 * a_before + sum((16 - i) * x_i): both sums vectorise
    while (blk >= 16) {
      uint32_t s = 0, ws = 0;
      for (int i = 0; i < 16; i++) { s += p[i]; ws += (uint32_t)(16 - i) * p[i]; }
      b += 16 * a + ws;
      a += s;
      p += 16; blk -= 16;
    }
    while (blk--) { a += *p++; b += a; }
    a %= 65521u; b %= 65521u;
  }
  return (b << 16) | a;
}

namespace {

// ---- inflate ---------------------------------------------------------------------------------
// Decode tables: u32 entries  [31:16] literal / base value / subtable offset   [15:12] kind   [11:8] extra bits (or
// subtable index bits)   [7:0] code bits to consume (0 = unused code -> corrupt stream).
enum : uint32_t { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_LINK = 3, K_DIST = 4 };
constexpr int LIT_BITS = 11, DIST_BITS = 8;
constexpr size_t IN_PAD = 64;
constexpr int LIT_CAP = (1 << LIT_BITS) + 288 * 16, DIST_CAP = (1 << DIST_BITS) + 32 * 128;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t bits) { return (value << 16) | (kind << 12) | (extra << 8) | bits; }

inline uint32_t symbol_entry(int sym, bool dist, int bits) {
  if (dist) return sym < 30 ? entry(DIST_BASE[sym], K_DIST, DIST_EXTRA[sym], bits) : 0;
  if (sym < 256) return entry(sym, K_LIT, 0, bits);
  if (sym == 256) return entry(0, K_EOB, 0, bits);
  return sym < 286 ? entry(LEN_BASE[sym - 257], K_LEN, LEN_EXTRA[sym - 257], bits) : 0;
}

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1); code >>= 1; }
  return r;
}

// canonical Huffman code lengths -> two-level lookup table indexed by the (LSB-first) bit buffer
const char* build_table(const uint8_t* lens, int n, bool dist, int root, uint32_t* table, int cap) {
  int count[16] = {0};
  for (int i = 0; i < n; i++) count[lens[i]]++;
  count[0] = 0;
  int left = 1, used = 0;
  for (int l = 1; l <= 15; l++) {
    left = (left << 1) - count[l];
    if (left < 0) return "inflate: over-subscribed Huffman code";
    used += count[l];
  }
  // incomplete codes: zlib accepts only the single-code (or empty) distance tree
  if (left > 0 && !(dist && used <= 1)) return "inflate: incomplete Huffman code";
  uint32_t next[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next[l] = code; }
  const int nroot = 1 << root;
  for (int i = 0; i < nroot; i++) table[i] = 0;
  uint8_t sub_max[1 << LIT_BITS];
  bool any_long = false;
  uint32_t codes[320];
  for (int s = 0; s < n; s++) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t r = reverse_bits(next[l]++, l);
    codes[s] = r;
    if (l <= root) {
      const uint32_t e = symbol_entry(s, dist, l);
      for (int i = r; i < nroot; i += 1 << l) table[i] = e;
    } else {
      if (!any_long) { memset(sub_max, 0, sizeof(sub_max)); any_long = true; }
      uint8_t& m = sub_max[r & (nroot - 1)];
      if (l > m) m = (uint8_t)l;
    }
  }
  if (any_long) {
    int top = nroot;
    for (int s = 0; s < n; s++) {
      const int l = lens[s];
      if (l <= root) continue;
      const uint32_t r = codes[s], pre = r & (nroot - 1);
      const int sub_bits = sub_max[pre] - root;
      if ((table[pre] >> 12 & 15) != K_LINK || (table[pre] & 0xff) == 0) {
        if (top + (1 << sub_bits) > cap) return "inflate: decode table overflow";
        table[pre] = entry(top, K_LINK, sub_bits, root);
        for (int i = 0; i < (1 << sub_bits); i++) table[top + i] = 0;
        top += 1 << sub_bits;
      }
      const uint32_t base = table[pre] >> 16;
      const uint32_t e = symbol_entry(s, dist, l - root);
      for (int i = r >> root; i < (1 << sub_bits); i += 1 << (l - root)) table[base + i] = e;
    }
  }
  return nullptr;
}

struct FixedTables {
  uint32_t lit[LIT_CAP], dist[DIST_CAP];
  FixedTables() {
    uint8_t l[288];
    for (int i = 0; i < 144; i++) l[i] = 8;
    for (int i = 144; i < 256; i++) l[i] = 9;
    for (int i = 256; i < 280; i++) l[i] = 7;
    for (int i = 280; i < 288; i++) l[i] = 8;
    build_table(l, 288, false, LIT_BITS, lit, LIT_CAP);
    uint8_t d[32];
    for (int i = 0; i < 32; i++) d[i] = 5;
    build_table(d, 32, true, DIST_BITS, dist, DIST_CAP);
  }
};

// `in` must be readable for IN_PAD bytes past in + n_in (the callers pad with zeros: the bit reader loads 8 bytes at a
// time and a corrupt block header can run ~25 bytes past the end before the next bounds check).  One raw deflate stream.
const char* inflate_raw(const uint8_t* in, size_t n_in, uint8_t* out, size_t out_cap, size_t* produced, size_t* consumed) {
  const uint8_t* const in_begin = in;
  const uint8_t* const in_limit = in + n_in + 8;   // the refill may run this far into the padding; beyond it = truncated
  uint8_t* const out_begin = out;
  uint8_t* const out_end = out + out_cap;
  uint64_t bitbuf = 0;
  int nbits = 0;
  uint32_t lit_dyn[LIT_CAP], dist_dyn[DIST_CAP];
#define ORB_REFILL()                                   \
  do {                                                 \
    bitbuf |= le64(in) << nbits;                       \
    in += (63 - nbits) >> 3;                           \
    nbits |= 56;                                       \
  } while (0)
#define ORB_TAKE(n) (tmp = (uint32_t)(bitbuf & ((1ull << (n)) - 1)), bitbuf >>= (n), nbits -= (n), tmp)
  uint32_t tmp;
  int last;
  do {
    if (in > in_limit) return "inflate: truncated stream";
    ORB_REFILL();
    last = ORB_TAKE(1);
    const int type = ORB_TAKE(2);
    const uint32_t *lit, *dist;
    if (type == 0) {
      // stored: drop to a byte boundary, give whole unread bytes back to the input
      const int drop = nbits & 7;
      bitbuf >>= drop; nbits -= drop;
      in -= nbits >> 3;
      bitbuf = 0; nbits = 0;
      if (in + 4 > in_begin + n_in) return "inflate: truncated stored block";
      const uint32_t len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
      if ((len ^ 0xffff) != nlen) return "inflate: stored block length mismatch";
      in += 4;
      if (in + len > in_begin + n_in) return "inflate: truncated stored block";
      if (out + len > out_end) return "inflate: output overflow";
      memcpy(out, in, len);
      in += len; out += len;
      continue;
    } else if (type == 1) {
      static const FixedTables F;
      lit = F.lit; dist = F.dist;
    } else if (type == 2) {
      const int hlit = ORB_TAKE(5) + 257, hdist = ORB_TAKE(5) + 1, hclen = ORB_TAKE(4) + 4;
      if (hlit > 286 || hdist > 30) return "inflate: too many length or distance symbols";
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (int i = 0; i < hclen; i++) {
        if (nbits < 3) ORB_REFILL();
        cl[order[i]] = (uint8_t)ORB_TAKE(3);
      }
      uint32_t cl_table[128 + 19 * 2];
      // a 7-bit root covers every code-length code (max 7 bits): no subtables
      {
        int count[8] = {0};
        for (int i = 0; i < 19; i++) count[cl[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l <= 7; l++) { left = (left << 1) - count[l]; if (left < 0) return "inflate: bad code-length code"; }
        if (left > 0) return "inflate: incomplete code-length code";
        uint32_t next[8], code = 0;
        for (int l = 1; l <= 7; l++) { code = (code + count[l - 1]) << 1; next[l] = code; }
        for (int i = 0; i < 128; i++) cl_table[i] = 0;
        for (int s = 0; s < 19; s++) {
          const int l = cl[s];
          if (!l) continue;
          const uint32_t r = reverse_bits(next[l]++, l);
          for (int i = r; i < 128; i += 1 << l) cl_table[i] = (s << 8) | l;
        }
      }
      uint8_t lens[320];
      int n = 0;
      while (n < hlit + hdist) {
        if (in > in_limit) return "inflate: truncated stream";
        ORB_REFILL();
        const uint32_t e = cl_table[bitbuf & 127];
        if (!(e & 0xff)) return "inflate: bad code-length symbol";
        bitbuf >>= (e & 0xff); nbits -= (e & 0xff);
        const int sym = e >> 8;
        if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
          if (n == 0) return "inflate: repeat with no previous length";
          val = lens[n - 1];
          rep = 3 + ORB_TAKE(2);
        } else if (sym == 17) rep = 3 + ORB_TAKE(3);
        else rep = 11 + ORB_TAKE(7);
        if (n + rep > hlit + hdist) return "inflate: repeat runs past the code lengths";
        while (rep--) lens[n++] = (uint8_t)val;
      }
      if (lens[256] == 0) return "inflate: no end-of-block code";
      const char* e1 = build_table(lens, hlit, false, LIT_BITS, lit_dyn, LIT_CAP);
      if (e1) return e1;
      const char* e2 = build_table(lens + hlit, hdist, true, DIST_BITS, dist_dyn, DIST_CAP);
      if (e2) return e2;
      lit = lit_dyn; dist = dist_dyn;
    } else {
      return "inflate: reserved block type";
    }

    for (;;) {
      if (in > in_limit) return "inflate: truncated stream";
      ORB_REFILL();
      uint32_t e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
      // up to three literals per refill: 3 * 15 < 56 bits
      if ((e & 0xff) && ((e >> 12) & 15) == K_LIT && out_end - out >= 4) {
        bitbuf >>= (e & 0xff); nbits -= (e & 0xff);
        *out++ = (uint8_t)(e >> 16);
        e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
        if (((e >> 12) & 15) == K_LIT && (e & 0xff)) {
          bitbuf >>= (e & 0xff); nbits -= (e & 0xff);
          *out++ = (uint8_t)(e >> 16);
          e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
          if (((e >> 12) & 15) == K_LIT && (e & 0xff)) {
            bitbuf >>= (e & 0xff); nbits -= (e & 0xff);
            *out++ = (uint8_t)(e >> 16);
            continue;
          }
        }
        // fewer than 30 bits used so far; what follows needs at most 48 more: refill to stay safe
        ORB_REFILL();
      }
      if (((e >> 12) & 15) == K_LINK) {
        bitbuf >>= LIT_BITS; nbits -= LIT_BITS;
        e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & 15)) - 1))];
      }
      const int cb = e & 0xff;
      if (!cb) return "inflate: invalid literal/length code";
      bitbuf >>= cb; nbits -= cb;
      const uint32_t kind = (e >> 12) & 15;
      if (kind == K_LIT) {
        if (out >= out_end) return "inflate: output overflow";
        *out++ = (uint8_t)(e >> 16);
        continue;
      }
      if (kind == K_EOB) break;
      const int xl = (e >> 8) & 15;
      const uint32_t length = (e >> 16) + ORB_TAKE(xl);
      uint32_t d = dist[bitbuf & ((1u << DIST_BITS) - 1)];
      if (((d >> 12) & 15) == K_LINK) {
        bitbuf >>= DIST_BITS; nbits -= DIST_BITS;
        d = dist[(d >> 16) + (bitbuf & ((1u << ((d >> 8) & 15)) - 1))];
      }
      const int db = d & 0xff;
      if (!db) return "inflate: invalid distance code";
      bitbuf >>= db; nbits -= db;
      const int xd = (d >> 8) & 15;
      const uint32_t distance = (d >> 16) + ORB_TAKE(xd);
      if (distance > (size_t)(out - out_begin)) return "inflate: distance too far back";
      if (length > (size_t)(out_end - out)) return "inflate: output overflow";
      const uint8_t* src = out - distance;
      if (distance >= 8 && (size_t)(out_end - out) >= length + 8) {
        // 8 bytes at a time, may write up to 7 bytes past the match (inside the output buffer)
        uint8_t* o = out;
        const uint8_t* const oe = out + length;
        do { memcpy(o, src, 8); o += 8; src += 8; } while (o < oe);
      } else if (distance == 1) {
        memset(out, *src, length);
      } else {
        for (uint32_t i = 0; i < length; i++) out[i] = src[i];
      }
      out += length;
    }
  } while (!last);
#undef ORB_REFILL
#undef ORB_TAKE
  // whole bytes still in the bit buffer were not consumed
  in -= nbits >> 3;
  if (in > in_begin + n_in) return "inflate: truncated stream";
  *produced = out - out_begin;
  *consumed = in - in_begin;
  return nullptr;
}

}  // namespace

Scratch::~Scratch() { free(buf); }
uint8_t* Scratch::need(size_t n) {
  if (n > cap) {
    free(buf);
    cap = n + n / 4 + 4096;
    buf = (uint8_t*)malloc(cap);
    if (!buf) cap = 0;
  }
  return buf;
}

const char* inflate_zlib(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out, size_t* produced, bool verify_adler) {
  if (n_in < 6) return "zlib: stream too short";
  const int cmf = in[0], flg = in[1];
  if ((cmf & 15) != 8 || (cmf >> 4) > 7) return "zlib: unknown compression method";
  if (((cmf << 8) | flg) % 31) return "zlib: header check failed";
  if (flg & 0x20) return "zlib: preset dictionary";
  // the bit reader needs 8 readable bytes past the end: decode from a padded copy
  Scratch pad;
  uint8_t* p = pad.need(n_in + IN_PAD);
  if (!p) return "out of memory";
  memcpy(p, in + 2, n_in - 2);
  memset(p + n_in - 2, 0, IN_PAD);
  size_t got = 0, used = 0;
  const char* e = inflate_raw(p, n_in - 2, out, n_out, &got, &used);
  if (e) return e;
  if (used + 4 > n_in - 2) return "zlib: missing checksum";
  if (verify_adler && be32(p + used) != adler32(out, got)) return "zlib: incorrect data check";
  *produced = got;
  return nullptr;
}

// ---- PNG -------------------------------------------------------------------------------------
namespace {
const uint8_t PNG_SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// reconstructs one scanline in place; prev = reconstructed previous line or nullptr for the first
const char* unfilter_row(int ft, uint8_t* cur, const uint8_t* prev, size_t rowbytes, int bpp) {
  switch (ft) {
    case 0: break;
    case 1: {
      size_t i = bpp;
#if defined(__SSE2__)
      if (bpp == 1) {
        // byte prefix sum, 16 at a time: log-step shifts inside the vector, then add the running total of the row
        __m128i carry = _mm_set1_epi8((char)cur[0]);
        for (; i + 16 <= rowbytes; i += 16) {
          __m128i x = _mm_loadu_si128((const __m128i*)(cur + i));
          x = _mm_add_epi8(x, _mm_slli_si128(x, 1));
          x = _mm_add_epi8(x, _mm_slli_si128(x, 2));
          x = _mm_add_epi8(x, _mm_slli_si128(x, 4));
          x = _mm_add_epi8(x, _mm_slli_si128(x, 8));
          x = _mm_add_epi8(x, carry);
          _mm_storeu_si128((__m128i*)(cur + i), x);
          carry = _mm_set1_epi8((char)cur[i + 15]);
        }
      }
#endif
      for (; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
      break;
    }
    case 2:
      if (prev) for (size_t i = 0; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + prev[i]);
      break;
    case 3:
      if (prev) {
        for (size_t i = 0; i < (size_t)bpp && i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + (prev[i] >> 1));
        for (size_t i = bpp; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + ((cur[i - bpp] + prev[i]) >> 1));
      } else {
        for (size_t i = bpp; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + (cur[i - bpp] >> 1));
      }
      break;
    case 4:
      if (prev) {
        for (size_t i = 0; i < (size_t)bpp && i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + prev[i]);
        if (bpp == 1) {
          int a = cur[0], c = prev[0];
          for (size_t i = 1; i < rowbytes; i++) {
            const int b = prev[i];
            a = (cur[i] + paeth(a, b, c)) & 0xff;
            cur[i] = (uint8_t)a;
            c = b;
          }
        } else {
          for (size_t i = bpp; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + paeth(cur[i - bpp], prev[i], prev[i - bpp]));
        }
      } else {
        for (size_t i = bpp; i < rowbytes; i++) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
      }
      break;
    default: return "png: unknown filter type";
  }
  return nullptr;
}

const char* parse_ihdr(const uint8_t* data, Info* info) {
  Info& I = *info;
  I.width = (int)be32(data); I.height = (int)be32(data + 4);
  I.bit_depth = data[8]; I.color_type = data[9];
  if (I.width <= 0 || I.height <= 0) return "png: bad image size";
  if (data[10] != 0 || data[11] != 0) return "png: unknown compression or filter method";
  if (data[12] != 0) return "png: interlaced files are not supported";
  if (I.bit_depth != 8 && I.bit_depth != 16) return "png: only 8 and 16 bit depths are supported";
  switch (I.color_type) {
    case 0: I.channels = 1; break;
    case 2: I.channels = 3; break;
    case 4: I.channels = 2; break;
    case 6: I.channels = 4; break;
    default: return "png: palette images are not supported";
  }
  return nullptr;
}

struct Parsed {
  Info info;
  size_t idat_total = 0;
};

// walks the chunk list; when `idat` is given, concatenates the IDAT payloads into it
const char* parse(const uint8_t* f, size_t n, Parsed* P, uint8_t* idat, bool check_crc, size_t skip = 0, size_t idat_cap = ~(size_t)0,
                  ChunkCrc* collect = nullptr, int collect_cap = 0, int* n_collected = nullptr) {
  if (n < 8 + 25 || memcmp(f, PNG_SIG, 8)) return "png: bad signature";
  size_t pos = 8;
  bool have_ihdr = false, have_iend = false;
  size_t total = 0;
  while (pos + 12 <= n) {
    const uint32_t len = be32(f + pos);
    const uint8_t* type = f + pos + 4;
    if (len > 0x7fffffffu || pos + 12 + (size_t)len > n) return "png: truncated chunk";
    const uint8_t* data = f + pos + 8;
    const bool is_ihdr = !memcmp(type, "IHDR", 4), is_idat = !memcmp(type, "IDAT", 4), is_iend = !memcmp(type, "IEND", 4);
    if (!have_ihdr && !is_ihdr) return "png: IHDR is not the first chunk";
    // with `collect`, the IDAT checksums are left to the caller (the device checks them): only IHDR is verified here
    if (check_crc && (is_ihdr || (is_idat && !collect)) && crc32(type, (size_t)len + 4) != be32(data + len)) return "png: chunk CRC mismatch";
    if (is_idat && collect) {
      if (*n_collected >= collect_cap) return "png: too many IDAT chunks for the device checksum table";
      collect[(*n_collected)++] = ChunkCrc{(uint32_t)total, len, be32(data + len)};
    }
    if (is_ihdr) {
      if (len != 13) return "png: bad IHDR";
      const char* e = parse_ihdr(data, &P->info);
      if (e) return e;
      have_ihdr = true;
    } else if (is_idat) {
      if (idat) {
        // the first `skip` bytes of the concatenated payload are dropped (zlib header for the device path)
        const size_t drop = total < skip ? (skip - total < len ? skip - total : len) : 0;
        if (total + len - skip > idat_cap && total + len > skip) return "png: compressed data larger than the frame slot";
        if (len > drop) memcpy(idat + (total + drop - skip), data + drop, len - drop);
      }
      total += len;
    } else if (is_iend) {
      have_iend = true;
      break;
    } else if (!(type[0] & 0x20)) {
      if (memcmp(type, "PLTE", 4)) return "png: unknown critical chunk";
    }
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr) return "png: no IHDR";
  if (!have_iend) return "png: no IEND";
  if (!total) return "png: no image data";
  P->idat_total = total;
  return nullptr;
}
}  // namespace

const char* read_info(const uint8_t* file, size_t n, Info* info) {
  // signature + IHDR only (33 bytes), so that a file head is enough
  if (n < 33 || memcmp(file, PNG_SIG, 8)) return "png: bad signature";
  if (be32(file + 8) != 13 || memcmp(file + 12, "IHDR", 4)) return "png: IHDR is not the first chunk";
  if (crc32(file + 12, 17) != be32(file + 29)) return "png: chunk CRC mismatch";
  Info& I = *info;
  return parse_ihdr(file + 16, &I);
}

const char* extract_deflate(const uint8_t* file, size_t n, Info* info, uint8_t* dst, size_t dst_cap, size_t* deflate_bytes,
                            ChunkCrc* crcs, int crc_cap, int* n_crcs) {
  Parsed P;
  const char* e = parse(file, n, &P, nullptr, false);
  if (e) return e;
  *info = P.info;
  if (P.idat_total < 2 + 4 + 1) return "zlib: stream too short";
  if (P.idat_total + 16 > dst_cap) return "png: compressed data larger than the frame slot";
  int nc = 0;
  if (crcs) {
    // many tiny IDAT chunks (never seen from libpng, which writes 8 KB chunks): check them here after all
    e = parse(file, n, &P, dst, true, 0, dst_cap, crcs, crc_cap, &nc);
    if (e && !strcmp(e, "png: too many IDAT chunks for the device checksum table")) { crcs = nullptr; nc = 0; }
    else if (e) return e;
  }
  if (!crcs && (e = parse(file, n, &P, dst, true, 0, dst_cap))) return e;
  if (n_crcs) *n_crcs = nc;
  // zlib framing (RFC 1950) in front of the deflate stream
  if ((dst[0] & 15) != 8 || (dst[0] >> 4) > 7) return "zlib: unknown compression method";
  if (((dst[0] << 8) | dst[1]) % 31) return "zlib: header check failed";
  if (dst[1] & 0x20) return "zlib: preset dictionary";
  memset(dst + P.idat_total, 0, 16);
  *deflate_bytes = P.idat_total - 2 - 4;      // without the zlib header and the Adler-32 trailer
  return nullptr;
}

const char* decode_gray8(const uint8_t* file, size_t n, uint8_t* dst, size_t pitch, int expect_w, int expect_h, Scratch* scratch) {
  Parsed P;
  const char* e = parse(file, n, &P, nullptr, false);
  if (e) return e;
  const Info& I = P.info;
  if (I.width != expect_w || I.height != expect_h) return "png: image size differs from the expected frame size";
  if (pitch < (size_t)I.width) return "png: destination pitch too small";
  const int bpp = I.channels * (I.bit_depth / 8);
  const size_t rowbytes = (size_t)I.width * bpp, raw_bytes = (rowbytes + 1) * (size_t)I.height;
  const size_t idat_cap = (P.idat_total + IN_PAD + 63) & ~(size_t)63;
  Scratch local;
  if (!scratch) scratch = &local;
  uint8_t* base = scratch->need(idat_cap + raw_bytes + 64);
  if (!base) return "out of memory";
  uint8_t* idat = base;
  uint8_t* raw = base + idat_cap;
  if ((e = parse(file, n, &P, idat, true))) return e;
  memset(idat + P.idat_total, 0, IN_PAD);
  // zlib framing (RFC 1950) around the deflate stream
  if (P.idat_total < 6) return "zlib: stream too short";
  if ((idat[0] & 15) != 8 || (idat[0] >> 4) > 7) return "zlib: unknown compression method";
  if (((idat[0] << 8) | idat[1]) % 31) return "zlib: header check failed";
  if (idat[1] & 0x20) return "zlib: preset dictionary";
  size_t got = 0, used = 0;
  if ((e = inflate_raw(idat + 2, P.idat_total - 2, raw, raw_bytes, &got, &used))) return e;
  if (got != raw_bytes) return "png: not enough image data";
  if (used + 4 > P.idat_total - 2) return "zlib: missing checksum";
  if (be32(idat + 2 + used) != adler32(raw, got)) return "zlib: incorrect data check";

  if (bpp == 1) {
    // 8-bit gray: reconstruct straight into the destination rows
    for (int y = 0; y < I.height; y++) {
      const uint8_t* line = raw + (size_t)y * (rowbytes + 1);
      uint8_t* row = dst + (size_t)y * pitch;
      memcpy(row, line + 1, rowbytes);
      if ((e = unfilter_row(line[0], row, y ? row - pitch : nullptr, rowbytes, 1))) return e;
    }
    return nullptr;
  }
  // other layouts: reconstruct in place in the scratch lines, then reduce to 8-bit gray
  const int step = I.bit_depth / 8;   // 16-bit samples are big endian: the high byte is the 8-bit value (strip)
  for (int y = 0; y < I.height; y++) {
    uint8_t* line = raw + (size_t)y * (rowbytes + 1);
    if ((e = unfilter_row(line[0], line + 1, y ? line - rowbytes : nullptr, rowbytes, bpp))) return e;
    const uint8_t* s = line + 1;
    uint8_t* row = dst + (size_t)y * pitch;
    if (I.color_type == 0 || I.color_type == 4) {
      for (int x = 0; x < I.width; x++) row[x] = s[(size_t)x * bpp];
    } else {
      // RGB -> gray with the 15-bit coefficients libpng derives from (0.299, 0.587): 9797, 19234, 3737; truncating for
      // 8-bit samples, rounding for 16-bit ones, which are then stripped to their high byte (pinned against cv2 4.13)
      for (int x = 0; x < I.width; x++) {
        const uint8_t* q = s + (size_t)x * bpp;
        if (step == 1) {
          const int r = q[0], g = q[1], b = q[2];
          row[x] = (r == g && g == b) ? (uint8_t)r : (uint8_t)((9797 * r + 19234 * g + 3737 * b) >> 15);
        } else {
          const uint32_t r = (q[0] << 8) | q[1], g = (q[2] << 8) | q[3], b = (q[4] << 8) | q[5];
          const uint32_t v = (r == g && g == b) ? r : ((9797u * r + 19234u * g + 3737u * b + 16384u) >> 15);
          row[x] = (uint8_t)(v >> 8);
        }
      }
    }
  }
  return nullptr;
}

}  // namespace orbpng
