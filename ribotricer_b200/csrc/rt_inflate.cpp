// rt_inflate.cpp -- raw DEFLATE (RFC 1951) decoder and CRC-32 for the BGZF blocks of a BAM file
// (SURVEY.md 8(f) "next #2": the inflate under pysam's AlignmentFile, bam.py:65-71, is where a BAM load spends its time).
//
// zlib 1.3's inflate runs at ~0.2 GB/s per core on Ribo-seq BGZF blocks, which bounds rt_bam_load on any number of cores.
// This decoder is built for whole-buffer decoding of small independent streams (a BGZF block holds <= 64 KiB):
//   * a 64-bit bit buffer refilled with one unaligned 8-byte load, without a branch, while >= 32 input bytes remain;
//   * one 2,048-entry table lookup per literal / length symbol (11 bits; longer codes go through a subtable) and one
//     256-entry lookup per distance symbol; an entry carries the number of bits to drop, the base value and where the
//     extra bits sit, so a symbol costs a load, two shifts and an add;
//   * up to three literals per refill; matches are copied in 8-byte words (run fill for distance 1);
//   * a careful loop with per-symbol bounds checks for the last bytes of the input / output.
// Anything unusual (an incomplete or over-subscribed code, a bad stored-block length, a distance before the start of
// the output, an output that is not exactly the announced size) returns false; rt_bam.cpp then hands the block to zlib,
// and every block is checked against the CRC-32 of its BGZF footer either way.
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "ribotricer_b200.h"

namespace {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitCap = (1 << kLitBits) + 1024, kDistCap = (1 << kDistBits) + 512;
constexpr uint32_t F_LIT = 1u << 31;   // litlen table: a literal, value in bits 16..23
constexpr uint32_t F_EXC = 1u << 15;   // not a plain symbol: subtable pointer (F_SUB), invalid (F_BAD), else end of block
constexpr uint32_t F_SUB = 1u << 14;
constexpr uint32_t F_BAD = 1u << 13;
// plain entry: [0..5] bits to drop (codeword + extra bits), [8..11] codeword bits, [16..30] base value
// subtable pointer: [0..5] bits to drop (the table bits), [8..11] index bits of the subtable, [16..30] its first entry

constexpr uint32_t F_LIT2 = 1u << 30;  // pair table only: two literals in one entry
// pair table, literal entry: [0..5] bits to drop (both codewords), [8..15] first literal, [16..23] second literal

struct Tables {
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    uint32_t pair[1 << kLitBits];       // the first kLitBits-bit lookup of the fast loop: `lit` with literal pairs folded in
};

// Literal-heavy streams (base and quality strings) are bound by the lookup -> shift -> lookup chain of one symbol after
// the other; when the codewords of two literals fit into one table index, the pair table yields both from one lookup.
void build_pairs(Tables& t) {
    for (uint32_t idx = 0; idx < (1u << kLitBits); ++idx) {
        const uint32_t e1 = t.lit[idx];
        if (!(e1 & F_LIT)) {
            t.pair[idx] = e1;
            continue;
        }
        const uint32_t l1 = e1 & 63u, e2 = t.lit[idx >> l1], l2 = e2 & 63u;
        const uint32_t one = F_LIT | (((e1 >> 16) & 0xffu) << 8) | l1;
        t.pair[idx] = ((e2 & F_LIT) && l1 + l2 <= (uint32_t)kLitBits) ? (one + l2) | F_LIT2 | (((e2 >> 16) & 0xffu) << 16) : one;
    }
}

inline uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline void st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
inline void st16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }
inline uint32_t rev16(uint32_t x) {
    x = ((x & 0x5555u) << 1) | ((x >> 1) & 0x5555u);
    x = ((x & 0x3333u) << 2) | ((x >> 2) & 0x3333u);
    x = ((x & 0x0f0fu) << 4) | ((x >> 4) & 0x0f0fu);
    return ((x & 0x00ffu) << 8) | ((x >> 8) & 0x00ffu);
}
inline uint32_t rev_bits(uint32_t code, int len) { return rev16(code) >> (16 - len); }

// per-symbol payload and number of extra bits (low byte)
uint32_t g_lit_info[288], g_dist_info[32];
std::once_flag g_info_once;
void init_info() {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (int s = 0; s < 256; ++s) g_lit_info[s] = F_LIT | ((uint32_t)s << 16);
    g_lit_info[256] = F_EXC;
    for (int s = 0; s < 29; ++s) g_lit_info[257 + s] = ((uint32_t)len_base[s] << 16) | len_extra[s];
    g_lit_info[286] = g_lit_info[287] = F_EXC | F_BAD;
    for (int s = 0; s < 30; ++s) g_dist_info[s] = ((uint32_t)dist_base[s] << 16) | dist_extra[s];
    g_dist_info[30] = g_dist_info[31] = F_EXC | F_BAD;
}

inline uint32_t make_entry(uint32_t info, int code_bits) {
    const uint32_t extra = info & 0xffu;
    return (info & ~0xffu) | ((uint32_t)code_bits << 8) | ((uint32_t)code_bits + extra);
}

// Canonical Huffman decode table.  false = over-subscribed, incomplete (unless `allow_single`: at most one code, as zlib
// accepts for the distance code) or too large for `cap`.
bool build_table(uint32_t* tab, int tab_bits, int cap, const uint8_t* lens, int n_sym, const uint32_t* info, bool allow_single) {
    int count[16] = {0};
    for (int s = 0; s < n_sym; ++s) count[lens[s]]++;
    const int n_codes = n_sym - count[0];
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
    }
    const int tab_size = 1 << tab_bits;
    if (left > 0) {
        if (!allow_single || n_codes > 1 || (n_codes == 1 && count[1] != 1)) return false;
        for (int i = 0; i < tab_size; ++i) tab[i] = F_EXC | F_BAD | 1u;
        if (n_codes == 1)
            for (int s = 0; s < n_sym; ++s)
                if (lens[s] == 1)
                    for (int i = 0; i < tab_size; i += 2) tab[i] = make_entry(info[s], 1);
        return true;
    }
    // symbols in canonical order
    uint16_t offs[17], sorted[288];
    offs[1] = 0;
    for (int l = 1; l <= 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n_sym; ++s)
        if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
    uint32_t code = 0;
    int i = 0, len = 1;
    while (len <= 15 && count[len] == 0) ++len;
    int in_len = 0;
    // short codes: replicate over the table
    for (; len <= tab_bits && len <= 15; ++len, code <<= 1, in_len = 0)
        for (in_len = 0; in_len < count[len]; ++in_len, ++i, ++code) {
            const uint32_t e = make_entry(info[sorted[i]], len);
            for (int j = (int)rev_bits(code, len); j < tab_size; j += 1 << len) tab[j] = e;
        }
    if (i == n_codes) return true;
    // long codes: codes with the same first tab_bits bits are neighbours in canonical order and share a subtable whose
    // width is that of the longest of them
    int next_free = tab_size;
    const int first_i = i, first_len = len;
    const uint32_t first_code = code;
    uint8_t max_len[1 << kLitBits];
    memset(max_len, 0, (size_t)tab_size);
    for (; len <= 15; ++len, code <<= 1)
        for (int c = 0; c < count[len]; ++c, ++i, ++code) max_len[rev_bits(code >> (len - tab_bits), tab_bits)] = (uint8_t)len;
    i = first_i;
    code = first_code;
    int cur_prefix = -1, sub_start = 0, sub_bits = 0;
    for (len = first_len; len <= 15; ++len, code <<= 1)
        for (int c = 0; c < count[len]; ++c, ++i, ++code) {
            const int prefix = (int)rev_bits(code >> (len - tab_bits), tab_bits);
            if (prefix != cur_prefix) {
                cur_prefix = prefix;
                sub_bits = max_len[prefix] - tab_bits;
                sub_start = next_free;
                next_free += 1 << sub_bits;
                if (next_free > cap) return false;
                tab[prefix] = F_EXC | F_SUB | ((uint32_t)sub_start << 16) | ((uint32_t)sub_bits << 8) | (uint32_t)tab_bits;
            }
            const int l2 = len - tab_bits;
            const uint32_t e = make_entry(info[sorted[i]], l2);
            for (int j = (int)rev_bits(code & ((1u << l2) - 1), l2); j < (1 << sub_bits); j += 1 << l2) tab[sub_start + j] = e;
        }
    return true;
}

Tables g_fixed;
bool g_fixed_ok = false;
std::once_flag g_fixed_once;
void init_fixed() {
    uint8_t lens[288 + 32];
    int s = 0;
    for (; s < 144; ++s) lens[s] = 8;
    for (; s < 256; ++s) lens[s] = 9;
    for (; s < 280; ++s) lens[s] = 7;
    for (; s < 288; ++s) lens[s] = 8;
    for (int d = 0; d < 32; ++d) lens[288 + d] = 5;
    g_fixed_ok = build_table(g_fixed.lit, kLitBits, kLitCap, lens, 288, g_lit_info, false) &&
                 build_table(g_fixed.dist, kDistBits, kDistCap, lens + 288, 32, g_dist_info, false);
    if (g_fixed_ok) build_pairs(g_fixed);
}

#define RT_MASK(n) ((1ull << (n)) - 1)

__attribute__((always_inline)) inline bool inflate_body(const uint8_t* in, size_t in_n, uint8_t* const out0, size_t out_n) {
    const uint8_t* ip = in;
    const uint8_t* const in_end = in + in_n;
    uint8_t* out = out0;
    uint8_t* const out_end = out0 + out_n;
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;
    Tables dyn;
    auto refill_slow = [&]() {
        while (bitcnt <= 55 && ip < in_end) {
            bitbuf |= (uint64_t)*ip++ << bitcnt;
            bitcnt += 8;
        }
    };
#define RT_NEED(n)                        \
    do {                                  \
        refill_slow();                    \
        if (bitcnt < (unsigned)(n)) return false; \
    } while (0)
#define RT_DROP(n)            \
    do {                      \
        bitbuf >>= (n);       \
        bitcnt -= (unsigned)(n); \
    } while (0)

    for (;;) {
        RT_NEED(3);
        const unsigned final_block = (unsigned)bitbuf & 1u, type = ((unsigned)bitbuf >> 1) & 3u;
        RT_DROP(3);
        const Tables* T = nullptr;
        if (type == 0) {                                   // stored
            RT_DROP(bitcnt & 7u);
            ip -= bitcnt >> 3;
            bitbuf = 0;
            bitcnt = 0;
            if (in_end - ip < 4) return false;
            const unsigned len = ip[0] | ((unsigned)ip[1] << 8), nlen = ip[2] | ((unsigned)ip[3] << 8);
            if (len != (~nlen & 0xffffu)) return false;
            ip += 4;
            if ((size_t)(in_end - ip) < len || (size_t)(out_end - out) < len) return false;
            memcpy(out, ip, len);
            ip += len;
            out += len;
            if (final_block) break;
            continue;
        } else if (type == 1) {
            std::call_once(g_fixed_once, init_fixed);
            if (!g_fixed_ok) return false;
            T = &g_fixed;
        } else if (type == 2) {
            RT_NEED(14);
            const int hlit = (int)(bitbuf & 31) + 257, hdist = (int)((bitbuf >> 5) & 31) + 1, hclen = (int)((bitbuf >> 10) & 15) + 4;
            RT_DROP(14);
            if (hlit > 286 || hdist > 30) return false;   // as zlib: "too many length or distance symbols"
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pre_lens[19] = {0};
            for (int k = 0; k < hclen; ++k) {
                RT_NEED(3);
                pre_lens[order[k]] = (uint8_t)(bitbuf & 7);
                RT_DROP(3);
            }
            // precode: 7-bit direct table of (symbol << 8 | length); 0 = invalid
            uint16_t pre[1 << kPreBits];
            {
                int count[8] = {0};
                for (int s = 0; s < 19; ++s) count[pre_lens[s]]++;
                int left = 1;
                for (int l = 1; l <= 7; ++l) {
                    left = (left << 1) - count[l];
                    if (left < 0) return false;
                }
                if (left != 0) return false;              // zlib rejects an incomplete code-length code
                uint32_t code = 0;
                for (int l = 1; l <= 7; ++l, code <<= 1)
                    for (int s = 0; s < 19; ++s)
                        if (pre_lens[s] == l) {
                            for (int j = (int)rev_bits(code, l); j < (1 << kPreBits); j += 1 << l) pre[j] = (uint16_t)((s << 8) | l);
                            ++code;
                        }
            }
            uint8_t lens[286 + 30 + 138];
            const int total = hlit + hdist;
            for (int k = 0; k < total;) {
                refill_slow();
                const unsigned pe = pre[bitbuf & RT_MASK(kPreBits)], sym = pe >> 8, pl = pe & 0xffu;
                if (pl + (sym < 16 ? 0u : sym == 16 ? 2u : sym == 17 ? 3u : 7u) > bitcnt) return false;
                RT_DROP(pl);
                if (sym < 16) {
                    lens[k++] = (uint8_t)sym;
                } else if (sym == 16) {
                    if (k == 0) return false;
                    const int rep = 3 + (int)(bitbuf & 3);
                    RT_DROP(2);
                    memset(lens + k, lens[k - 1], (size_t)rep);
                    k += rep;
                } else if (sym == 17) {
                    const int rep = 3 + (int)(bitbuf & 7);
                    RT_DROP(3);
                    memset(lens + k, 0, (size_t)rep);
                    k += rep;
                } else {
                    const int rep = 11 + (int)(bitbuf & 127);
                    RT_DROP(7);
                    memset(lens + k, 0, (size_t)rep);
                    k += rep;
                }
                if (k > total) return false;
            }
            if (lens[256] == 0) return false;              // no end-of-block code
            if (!build_table(dyn.lit, kLitBits, kLitCap, lens, hlit, g_lit_info, false)) return false;
            if (!build_table(dyn.dist, kDistBits, kDistCap, lens + hlit, hdist, g_dist_info, true)) return false;
            build_pairs(dyn);
            T = &dyn;
        } else {
            return false;
        }
        const uint32_t* const lit = T->lit;
        const uint32_t* const dist = T->dist;
        const uint32_t* const pair = T->pair;
        bool block_done = false;

        // ---- fast loop: no bounds checks inside
        while (in_end - ip >= 32 && out_end - out >= 288) {
            bitbuf |= ld64(ip) << bitcnt;
            ip += (63 - bitcnt) >> 3;
            bitcnt |= 56;
            uint32_t e = pair[bitbuf & RT_MASK(kLitBits)];
            if (e & F_LIT) {                               // one or two literals per lookup, three lookups per refill
                RT_DROP(e & 63u);
                st16(out, (uint16_t)(e >> 8));
                out += (e >> 30) - 1;
                e = pair[bitbuf & RT_MASK(kLitBits)];
                if (e & F_LIT) {
                    RT_DROP(e & 63u);
                    st16(out, (uint16_t)(e >> 8));
                    out += (e >> 30) - 1;
                    e = pair[bitbuf & RT_MASK(kLitBits)];
                    if (e & F_LIT) {
                        RT_DROP(e & 63u);
                        st16(out, (uint16_t)(e >> 8));
                        out += (e >> 30) - 1;
                        continue;
                    }
                }
            }
            if (e & F_EXC) {
                if (!(e & F_SUB)) {
                    if (e & F_BAD) return false;
                    RT_DROP(e & 63u);
                    block_done = true;
                    break;
                }
                RT_DROP(e & 63u);
                e = lit[(e >> 16) + (bitbuf & RT_MASK((e >> 8) & 15u))];
                if (e & F_LIT) {
                    RT_DROP(e & 63u);
                    *out++ = (uint8_t)(e >> 16);
                    continue;
                }
                if (e & F_EXC) {
                    if (e & F_BAD) return false;
                    RT_DROP(e & 63u);
                    block_done = true;
                    break;
                }
            }
            uint64_t saved = bitbuf;
            RT_DROP(e & 63u);
            const unsigned len = ((e >> 16) & 0x1ffu) + (unsigned)((saved & RT_MASK(e & 63u)) >> ((e >> 8) & 15u));
            if (bitcnt < 28) {                             // a distance takes up to 15 + 13 bits
                bitbuf |= ld64(ip) << bitcnt;
                ip += (63 - bitcnt) >> 3;
                bitcnt |= 56;
            }
            e = dist[bitbuf & RT_MASK(kDistBits)];
            if (e & F_EXC) {
                if (!(e & F_SUB)) return false;
                RT_DROP(e & 63u);
                e = dist[(e >> 16) + (bitbuf & RT_MASK((e >> 8) & 15u))];
                if (e & F_EXC) return false;
            }
            saved = bitbuf;
            RT_DROP(e & 63u);
            const size_t d = (e >> 16) + (size_t)((saved & RT_MASK(e & 63u)) >> ((e >> 8) & 15u));
            if (d > (size_t)(out - out0)) return false;
            const uint8_t* s = out - d;
            uint8_t* o = out;
            out += len;
            if (d >= 8) {
                st64(o, ld64(s));
                st64(o + 8, ld64(s + 8));
                if (len > 16) {
                    o += 16; s += 16;
                    do {
                        st64(o, ld64(s));
                        o += 8; s += 8;
                    } while (o < out);
                }
            } else if (d == 1) {
                const uint64_t v = 0x0101010101010101ull * *s;
                do {
                    st64(o, v);
                    o += 8;
                } while (o < out);
            } else {
                do *o++ = *s++; while (o < out);
            }
        }

        // ---- careful loop: the last bytes of the input or of the output
        while (!block_done) {
            refill_slow();
            uint32_t e = lit[bitbuf & RT_MASK(kLitBits)];
            if ((e & F_EXC) && (e & F_SUB)) {
                if ((e & 63u) > bitcnt) return false;
                RT_DROP(e & 63u);
                e = lit[(e >> 16) + (bitbuf & RT_MASK((e >> 8) & 15u))];
            }
            if ((e & 63u) > bitcnt) return false;
            if (e & F_EXC) {
                if (e & (F_BAD | F_SUB)) return false;
                RT_DROP(e & 63u);
                break;
            }
            uint64_t saved = bitbuf;
            RT_DROP(e & 63u);
            if (e & F_LIT) {
                if (out == out_end) return false;
                *out++ = (uint8_t)(e >> 16);
                continue;
            }
            const size_t len = ((e >> 16) & 0x1ffu) + (size_t)((saved & RT_MASK(e & 63u)) >> ((e >> 8) & 15u));
            refill_slow();
            e = dist[bitbuf & RT_MASK(kDistBits)];
            if ((e & F_EXC) && (e & F_SUB)) {
                if ((e & 63u) > bitcnt) return false;
                RT_DROP(e & 63u);
                e = dist[(e >> 16) + (bitbuf & RT_MASK((e >> 8) & 15u))];
            }
            if (e & F_EXC) return false;
            if ((e & 63u) > bitcnt) return false;
            saved = bitbuf;
            RT_DROP(e & 63u);
            const size_t d = (e >> 16) + (size_t)((saved & RT_MASK(e & 63u)) >> ((e >> 8) & 15u));
            if (d > (size_t)(out - out0) || len > (size_t)(out_end - out)) return false;
            const uint8_t* s = out - d;
            for (size_t k = 0; k < len; ++k) out[k] = s[k];
            out += len;
        }
        if (final_block) break;
    }
    return out == out_end;
#undef RT_NEED
#undef RT_DROP
}

// The same body compiled twice: with BMI2 (shrx / bzhi take the variable shifts and masks off the symbol chain) and plain.
bool inflate_plain(const uint8_t* in, size_t in_n, uint8_t* out, size_t out_n) { return inflate_body(in, in_n, out, out_n); }
#if defined(__x86_64__)
__attribute__((target("bmi2"))) bool inflate_bmi2(const uint8_t* in, size_t in_n, uint8_t* out, size_t out_n) {
    return inflate_body(in, in_n, out, out_n);
}
#endif
bool inflate_fast(const uint8_t* in, size_t in_n, uint8_t* out, size_t out_n) {
    std::call_once(g_info_once, init_info);
#if defined(__x86_64__)
    static const bool bmi2 = __builtin_cpu_supports("bmi2");
    if (bmi2) return inflate_bmi2(in, in_n, out, out_n);
#endif
    return inflate_plain(in, in_n, out, out_n);
}

// CRC-32 (the gzip polynomial) by carry-less multiplication: four 128-bit lanes folded over 64-byte strides, then
// reduced (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009).
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t* buf, size_t len, uint32_t crc) {
    // len >= 64 and a multiple of 16; crc is the running register (already inverted)
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
        y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
        y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

}  // namespace

// C++ entry points used by rt_bam.cpp
bool rt_inflate_fast(const uint8_t* in, size_t in_n, uint8_t* out, size_t out_n) { return inflate_fast(in, in_n, out, out_n); }

uint32_t rt_crc32_fast(const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) {
        const size_t body = n & ~(size_t)15;
        const uint32_t reg = crc32_clmul(p, body, 0xffffffffu);
        return (uint32_t)crc32(~reg, p + body, (uInt)(n - body));
    }
#endif
    uint32_t c = 0;
    while (n) {
        const size_t step = n > (1u << 30) ? (1u << 30) : n;
        c = (uint32_t)crc32(c, p, (uInt)step);
        p += step;
        n -= step;
    }
    return c;
}

extern "C" {

int rt_inflate_raw(const uint8_t* src, int64_t n_src, uint8_t* dst, int64_t n_dst) {
    if (n_src < 0 || n_dst < 0 || (!src && n_src) || (!dst && n_dst)) return RT_EINVAL;
    return inflate_fast(src, (size_t)n_src, dst, (size_t)n_dst) ? RT_OK : RT_EINVAL;
}

uint32_t rt_crc32(const uint8_t* p, int64_t n) { return (p && n > 0) ? rt_crc32_fast(p, (size_t)n) : 0u; }

}  // extern "C"
