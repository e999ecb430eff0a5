// DEFLATE (RFC 1951) decoder for the PNG reader of the predict path's input side (robosat/tiles.py:150-159,181 reads every tile
// through PIL -> libpng -> zlib inflate). Decoding a 512x512 RGB tile is 85 % inflate, and with the hosts' CPU quota the tile
// decoder bounds `rs predict` over a PNG directory from two GPUs up (profiles/r2_cfg4.md), so the hot loop is written for
// throughput: a 64-bit bit buffer refilled with one unaligned load, 11-bit primary decode tables with sub-tables for longer
// codes, up to three literals per refill, 8-byte match copies. Host code; no device work. rsb_png.cpp falls back to zlib for any
// stream this decoder rejects, and the zlib wrapper's Adler-32 is verified here, so a wrong decode cannot pass silently.
#include "rsb_inflate.h"

#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace rsb {
namespace {

constexpr int kLitBits = 11;   // primary litlen table: 2048 entries
constexpr int kDistBits = 8;   // primary distance table
constexpr int kPreBits = 7;    // code-length code: codes are at most 7 bits, no sub-tables
constexpr int kLitTableSize = (1 << kLitBits) + 288 * 16;
constexpr int kDistTableSize = (1 << kDistBits) + 32 * 128;

// table entry (32 bits): low byte = bits to consume for this entry; flags; payload in the high half
constexpr uint32_t kLiteral = 0x8000;   // payload = the byte
constexpr uint32_t kEndOfBlock = 0x4000;
constexpr uint32_t kSubtable = 0x2000;  // payload = start of the sub-table, bits 8..11 = its index width
constexpr uint32_t kInvalid = 0x1000;   // unused code space / reserved symbol
// length / distance entries: payload = base value, bits 8..11 = number of extra bits

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Kind { kLitLen, kDist, kPre };

inline uint32_t symbol_entry(Kind kind, int sym) {
    if (kind == kPre) return uint32_t(sym) << 16;
    if (kind == kDist) return sym < 30 ? (uint32_t(kDistBase[sym]) << 16) | (uint32_t(kDistExtra[sym]) << 8) : kInvalid;
    if (sym < 256) return (uint32_t(sym) << 16) | kLiteral;
    if (sym == 256) return kEndOfBlock;
    if (sym < 286) return (uint32_t(kLenBase[sym - 257]) << 16) | (uint32_t(kLenExtra[sym - 257]) << 8);
    return kInvalid;
}

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) {
        r = (r << 1) | (code & 1);
        code >>= 1;
    }
    return r;
}

// Canonical Huffman code -> decode table indexed by the next `tablebits` stream bits (LSB first). Codes longer than tablebits
// go through a sub-table per primary prefix, as wide as the longest code of that prefix. Returns false for an over-subscribed
// code; an incomplete code leaves kInvalid entries (decoding one fails), which also covers the one-distance-code streams.
bool build_table(const uint8_t* lens, int nsyms, Kind kind, int tablebits, uint32_t* table, int table_capacity) {
    int count[16] = {0};
    for (int i = 0; i < nsyms; ++i) ++count[lens[i]];
    count[0] = 0;
    int left = 1;
    for (int len = 1; len <= 15; ++len) {
        left = (left << 1) - count[len];
        if (left < 0) return false;
    }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int len = 1; len <= 15; ++len) {
        next_code[len] = code;
        code = (code + uint32_t(count[len])) << 1;
    }
    const int primary = 1 << tablebits;
    for (int i = 0; i < primary; ++i) table[i] = kInvalid | 1;
    // pass 1: codewords; width of every sub-table = longest code sharing the primary prefix
    uint32_t rev[288];
    uint8_t subbits[1 << kLitBits];
    bool any_long = false;
    for (int s = 0; s < nsyms; ++s) {
        const int len = lens[s];
        if (!len) continue;
        rev[s] = reverse_bits(next_code[len]++, len);
        if (len > tablebits) {
            if (!any_long) {
                memset(subbits, 0, size_t(primary));
                any_long = true;
            }
            const uint32_t prefix = rev[s] & uint32_t(primary - 1);
            if (len - tablebits > subbits[prefix]) subbits[prefix] = uint8_t(len - tablebits);
        }
    }
    int used = primary;
    if (any_long) {
        for (int p = 0; p < primary; ++p) {
            if (!subbits[p]) continue;
            const int size = 1 << subbits[p];
            if (used + size > table_capacity) return false;
            table[p] = (uint32_t(used) << 16) | kSubtable | (uint32_t(subbits[p]) << 8) | uint32_t(tablebits);
            for (int i = 0; i < size; ++i) table[used + i] = kInvalid | 1;
            used += size;
        }
    }
    // pass 2: fill
    for (int s = 0; s < nsyms; ++s) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t e = symbol_entry(kind, s);
        if (len <= tablebits) {
            const uint32_t entry = e | uint32_t(len);
            for (uint32_t i = rev[s]; i < uint32_t(primary); i += 1u << len) table[i] = entry;
        } else {
            const uint32_t prefix = rev[s] & uint32_t(primary - 1);
            const uint32_t sub = table[prefix];
            const int width = int((sub >> 8) & 0xf);
            const uint32_t start = sub >> 16;
            const int rest = len - tablebits;
            const uint32_t entry = e | uint32_t(rest);
            for (uint32_t i = rev[s] >> tablebits; i < (1u << width); i += 1u << rest) table[start + i] = entry;
        }
    }
    return true;
}

inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64 / aarch64 as deployed); checked in rsb_inflate_raw
}

struct Tables {
    uint32_t lit[kLitTableSize];
    uint32_t dist[kDistTableSize];
    uint32_t pre[1 << kPreBits];
};

}  // namespace

// Branch-free refill (bits past `bitcnt` may already hold the next bytes: OR-ing the same bytes again is idempotent).
// Needs 8 readable bytes at `in`; afterwards 56 <= bitcnt <= 63.
#define RSB_REFILL()                                   \
    do {                                               \
        bitbuf |= load64(in) << bitcnt;                \
        in += (63 - bitcnt) >> 3;                      \
        bitcnt |= 56;                                  \
    } while (0)
#define RSB_DROP(n)            \
    do {                       \
        bitbuf >>= (n);        \
        bitcnt -= unsigned(n); \
    } while (0)

int rsb_inflate_raw(const uint8_t* in_begin, size_t in_len, uint8_t* out_begin, size_t out_len, size_t* consumed) {
    const uint16_t endian_probe = 1;
    if (*reinterpret_cast<const uint8_t*>(&endian_probe) != 1) return -1;
    Tables tb;  // 45 KB on the stack (a thread_local here costs a __tls_get_addr round trip per use inside a shared library: +50 %)
    const uint8_t* in = in_begin;
    const uint8_t* const in_end = in_begin + in_len;   // real end; kInflateInPad readable bytes follow
    uint8_t* out = out_begin;
    uint8_t* const out_end = out_begin + out_len;      // real end; kInflateOutPad writable bytes follow
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;
    bool last = false;
    while (!last) {
        if (in > in_end + 8) return -2;
        RSB_REFILL();
        last = (bitbuf & 1) != 0;
        const unsigned type = unsigned(bitbuf >> 1) & 3;
        RSB_DROP(3);
        if (type == 0) {
            // stored: back to the byte boundary, hand the unread whole bytes back to the input
            RSB_DROP(bitcnt & 7);
            in -= bitcnt >> 3;
            bitbuf = 0;
            bitcnt = 0;
            if (in + 4 > in_end) return -2;
            const unsigned len = unsigned(in[0]) | (unsigned(in[1]) << 8), nlen = unsigned(in[2]) | (unsigned(in[3]) << 8);
            in += 4;
            if ((len ^ nlen) != 0xffff) return -3;
            if (in + len > in_end || out + len > out_end) return -2;
            memcpy(out, in, len);
            in += len;
            out += len;
            continue;
        }
        if (type == 3) return -3;
        if (type == 1) {
            uint8_t lens[288 + 32];
            for (int i = 0; i < 144; ++i) lens[i] = 8;
            for (int i = 144; i < 256; ++i) lens[i] = 9;
            for (int i = 256; i < 280; ++i) lens[i] = 7;
            for (int i = 280; i < 288; ++i) lens[i] = 8;
            for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
            if (!build_table(lens, 288, kLitLen, kLitBits, tb.lit, kLitTableSize)) return -3;
            if (!build_table(lens + 288, 32, kDist, kDistBits, tb.dist, kDistTableSize)) return -3;
        } else {
            const unsigned hlit = unsigned(bitbuf & 31) + 257, hdist = unsigned((bitbuf >> 5) & 31) + 1, hclen = unsigned((bitbuf >> 10) & 15) + 4;
            RSB_DROP(14);
            if (hlit > 286 || hdist > 30) return -3;
            uint8_t pre_lens[19] = {0};
            for (unsigned i = 0; i < hclen; ++i) {
                if (bitcnt < 3) RSB_REFILL();
                pre_lens[kPreOrder[i]] = uint8_t(bitbuf & 7);
                RSB_DROP(3);
            }
            if (!build_table(pre_lens, 19, kPre, kPreBits, tb.pre, 1 << kPreBits)) return -3;
            uint8_t lens[286 + 30 + 138];
            unsigned n = 0;
            while (n < hlit + hdist) {
                if (in > in_end + 8) return -2;
                RSB_REFILL();
                const uint32_t e = tb.pre[bitbuf & ((1u << kPreBits) - 1)];
                if (e & kInvalid) return -3;
                RSB_DROP(e & 0xff);
                const unsigned sym = e >> 16;
                if (sym < 16) {
                    lens[n++] = uint8_t(sym);
                } else if (sym == 16) {
                    if (n == 0) return -3;
                    const unsigned rep = 3 + unsigned(bitbuf & 3);
                    RSB_DROP(2);
                    memset(lens + n, lens[n - 1], rep);
                    n += rep;
                } else if (sym == 17) {
                    const unsigned rep = 3 + unsigned(bitbuf & 7);
                    RSB_DROP(3);
                    memset(lens + n, 0, rep);
                    n += rep;
                } else {
                    const unsigned rep = 11 + unsigned(bitbuf & 127);
                    RSB_DROP(7);
                    memset(lens + n, 0, rep);
                    n += rep;
                }
            }
            if (n != hlit + hdist) return -3;  // a repeat ran past the end
            if (lens[256] == 0) return -3;     // no end-of-block code
            if (!build_table(lens, int(hlit), kLitLen, kLitBits, tb.lit, kLitTableSize)) return -3;
            if (!build_table(lens + hlit, int(hdist), kDist, kDistBits, tb.dist, kDistTableSize)) return -3;
        }

        // ---- the block's symbols. Per iteration: one refill (>= 56 bits), up to two literals and one more code (3 x 15 bits),
        // length extra bits (5), a second refill, distance code + extra (15 + 13). Bounds are checked once per iteration; what a
        // corrupt stream can write past out_end before the check (3 literals + one 258-byte match + 7) fits the output padding.
        const uint32_t* const lit = tb.lit;
        const uint32_t* const dist_table = tb.dist;
        for (;;) {
            if (in > in_end + 8 || out > out_end) return -2;
            RSB_REFILL();
            uint32_t e = lit[bitbuf & ((1u << kLitBits) - 1)];
            if (e & kSubtable) {
                RSB_DROP(kLitBits);
                e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & 0xf)) - 1))];
            }
            RSB_DROP(e & 0xff);
            if (e & kLiteral) {
                *out++ = uint8_t(e >> 16);
                e = lit[bitbuf & ((1u << kLitBits) - 1)];
                if (e & kSubtable) {
                    RSB_DROP(kLitBits);
                    e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & 0xf)) - 1))];
                }
                RSB_DROP(e & 0xff);
                if (e & kLiteral) {
                    *out++ = uint8_t(e >> 16);
                    e = lit[bitbuf & ((1u << kLitBits) - 1)];
                    if (e & kSubtable) {
                        RSB_DROP(kLitBits);
                        e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & 0xf)) - 1))];
                    }
                    RSB_DROP(e & 0xff);
                    if (e & kLiteral) {
                        *out++ = uint8_t(e >> 16);
                        continue;
                    }
                }
            }
            if (e & (kEndOfBlock | kInvalid)) {
                if (e & kInvalid) return -3;
                break;
            }
            const unsigned lbits = (e >> 8) & 0xf;
            const unsigned length = (e >> 16) + unsigned(bitbuf & ((1u << lbits) - 1));
            RSB_DROP(lbits);
            RSB_REFILL();
            uint32_t d = dist_table[bitbuf & ((1u << kDistBits) - 1)];
            if (d & kSubtable) {
                RSB_DROP(kDistBits);
                d = dist_table[(d >> 16) + (bitbuf & ((1u << ((d >> 8) & 0xf)) - 1))];
            }
            if (d & kInvalid) return -3;
            RSB_DROP(d & 0xff);
            const unsigned dbits = (d >> 8) & 0xf;
            const size_t distance = (d >> 16) + size_t(bitbuf & ((1u << dbits) - 1));
            RSB_DROP(dbits);
            if (distance > size_t(out - out_begin) || out + length > out_end) return -2;
            const uint8_t* src = out - distance;
            uint8_t* dst = out;
            out += length;
            if (distance >= 8) {
                do {
                    memcpy(dst, src, 8);  // may write up to 7 bytes past the match: inside the buffer or its padding
                    dst += 8;
                    src += 8;
                } while (dst < out);
            } else if (distance == 1) {
                memset(dst, *src, length);
            } else {
                do {
                    *dst++ = *src++;
                } while (dst < out);
            }
        }
    }
    // bytes really consumed: whole bytes still in the bit buffer were not used
    in -= bitcnt >> 3;
    if (in > in_end) return -2;
    if (out != out_end) return -4;
    if (consumed) *consumed = size_t(in - in_begin);
    return 0;
}

#undef RSB_REFILL
#undef RSB_DROP

#if defined(__x86_64__)
// 32 bytes per step: s1 through psadbw, s2 = 32 * (s1 before the block) + sum of (32 - i) * byte_i through pmaddubsw / pmaddwd
__attribute__((target("avx2"))) static inline uint32_t hsum(__m256i x) {
    __m128i lo = _mm_add_epi32(_mm256_castsi256_si128(x), _mm256_extracti128_si256(x, 1));
    lo = _mm_add_epi32(lo, _mm_shuffle_epi32(lo, 0x4e));
    lo = _mm_add_epi32(lo, _mm_shuffle_epi32(lo, 0xb1));
    return uint32_t(_mm_cvtsi128_si32(lo));
}

__attribute__((target("avx2"))) static void adler32_avx2(const uint8_t*& p, size_t blocks, uint32_t& a, uint32_t& b) {
    const __m256i weights = _mm256_setr_epi8(32, 31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3,
                                             2, 1);
    const __m256i ones = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
    __m256i s1 = zero, s1_before = zero, s2 = zero;
    for (size_t j = 0; j < blocks; ++j) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
        p += 32;
        s1_before = _mm256_add_epi32(s1_before, s1);
        s1 = _mm256_add_epi32(s1, _mm256_sad_epu8(v, zero));
        s2 = _mm256_add_epi32(s2, _mm256_madd_epi16(_mm256_maddubs_epi16(v, weights), ones));
    }
    b += uint32_t(blocks) * 32u * a + 32u * hsum(s1_before) + hsum(s2);
    a += hsum(s1);
}
#endif

uint32_t rsb_adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
#endif
    while (n) {
        // 5552 = the most bytes whose sums cannot overflow 32 bits before the modulo
        size_t chunk = n < 5552 ? n : 5552;
        n -= chunk;
#if defined(__x86_64__)
        if (have_avx2 && chunk >= 32) {
            const size_t blocks = chunk / 32;  // <= 173: the vector lanes stay far below 2^32
            adler32_avx2(p, blocks, a, b);
            chunk -= blocks * 32;
        }
#endif
        while (chunk >= 8) {
            a += p[0]; b += a;
            a += p[1]; b += a;
            a += p[2]; b += a;
            a += p[3]; b += a;
            a += p[4]; b += a;
            a += p[5]; b += a;
            a += p[6]; b += a;
            a += p[7]; b += a;
            p += 8;
            chunk -= 8;
        }
        while (chunk--) {
            a += *p++;
            b += a;
        }
        a %= 65521;
        b %= 65521;
    }
    return (b << 16) | a;
}

int rsb_inflate_zlib_padded(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    if (in_len < 6) return -2;
    const unsigned cmf = in[0], flg = in[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return -3;  // deflate, <= 32 K window, no preset dictionary
    size_t used = 0;
    const int rc = rsb_inflate_raw(in + 2, in_len - 2 - 4, out, out_len, &used);
    if (rc) return rc;
    const uint8_t* t = in + 2 + used;
    if (t + 4 > in + in_len) return -2;
    const uint32_t want = (uint32_t(t[0]) << 24) | (uint32_t(t[1]) << 16) | (uint32_t(t[2]) << 8) | uint32_t(t[3]);
    return rsb_adler32(out, out_len) == want ? 0 : -5;
}

}  // namespace rsb
