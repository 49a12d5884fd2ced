// rt_kernels.cuh -- sm_100a kernels of the detect-orfs scoring path.
//
//   K1 bin_psites_kernel      bam.py:71-137 + detect_orfs.py:54-83
//   K3 score_orfs_kernel      detect_orfs.py:134-203,274-299 + statistics.py:48-115
//                             + common.py:164-180 (gather fused in: K2)
//   K4 gather_profiles_kernel detect_orfs.py:134-203 for the reported ORFs (:322)
//
// All "file:line" citations are relative to the reference tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ribotricer_b200.h"

namespace rt {

constexpr unsigned kFull = 0xffffffffu;

// ---- device-resident index encoding ------------------------------------------------------
// exon entry : (slot_offset << 24) | len      slot_offset relative to d_cov (strand plane
//              included), len in [1, 2^24); slot_offset == kZeroOff means "reads as zeros"
//              (unknown contig / strand, or the part of an exon outside the padded contig).
// ORF desc   : exon_begin (40 bits) | n_entries (23 bits) << 40 | reverse << 63
constexpr int kLenBits = 24;
constexpr uint64_t kLenMask = (1ull << kLenBits) - 1;
constexpr uint64_t kZeroOff = (1ull << 40) - 1;
constexpr uint64_t kBeginMask = (1ull << 40) - 1;
constexpr int kMaxEntriesPerOrf = (1 << 23) - 1;

// ---- K3 tiling ------------------------------------------------------------------------------
constexpr int kScoreWarps = 8;                 // warps per CTA
constexpr int kTileCodons = 256;               // codons per warp tile (8 rounds of 32 lanes)
constexpr int kTileNt = 3 * kTileCodons;       // window starts per tile
constexpr int kBufNt = kTileNt + 8;            // + 2 halo + zero pad, keeps 16 B multiples
constexpr int kFetchBatch = 4;                 // ORFs claimed per atomic

struct FrameAcc {
    int K = 0;        // kept codons (statistics.py:72 negated)
    int na = 0;       // codons (a,0,0): unit vector (1, 0)
    int nb = 0;       // codons (0,b,0): unit vector (-1/2, +sqrt3/2)
    int nc = 0;       // codons (0,0,c): unit vector (-1/2, -sqrt3/2)
    int ng = 0;       // other non-uniform codons, summed in fp64 below
    double sre = 0.0; // sum of A / sqrt(A^2 + 3 B^2),          A = 2a - b - c
    double sim = 0.0; // sum of B / sqrt(A^2 + 3 B^2) (x sqrt3), B = b - c
};

// One codon (a,b,c) of one frame: statistics.py:72-90 in closed form.  The reference
// normalises the triplet by |a + b w + c w^2| and SciPy's coherence then only sees the unit
// vector u = (a + b w + c w^2) / |.| of every non-uniform kept codon;
// 2 Re = A, 2 Im = sqrt3 * B, 4 |.|^2 = A^2 + 3 B^2.
__device__ __forceinline__ void accumulate_codon(int a, int b, int c, bool complete, FrameAcc& f) {
    if (complete && (a | b | c) != 0) {
        f.K++;
        const int nz = (a != 0) + (b != 0) + (c != 0);
        if (nz == 1) {
            if (a != 0) f.na++;
            else if (b != 0) f.nb++;
            else f.nc++;
        } else if (a != b || b != c) {
            const double dA = (double)(2ll * a - b - c);
            const double dB = (double)((long long)b - c);
            const double r = rsqrt(fma(dA, dA, 3.0 * dB * dB));
            f.sre = fma(dA, r, f.sre);
            f.sim = fma(dB, r, f.sim);
            f.ng++;
        }
    }
}

__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ long long warp_min_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long t = __shfl_xor_sync(kFull, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ int ld_cov(const int32_t* p) { return __ldg(p); }

// Cursor over the concatenated exons of one ORF in PROFILE order (ascending genomic
// positions for '+', descending for '-': detect_orfs.py:176-187,201-202).  Warp-uniform.
struct ExonCursor {
    const uint64_t* entries;  // first entry of this ORF
    int n;                    // entries of this ORF
    bool rev;
    int next = 0;             // entries consumed
    int rem = 0;              // values left in the current entry
    long long pos = 0;        // slot of the next value (relative to cov)
    bool zero = false;
    uint64_t cached = 0;      // lane-private: entry (next & ~31) + lane

    __device__ __forceinline__ bool advance(int lane) {
        if (next == n) return false;
        if ((next & 31) == 0) {
            const int idx = next + lane;
            if (idx < n) cached = __ldg(entries + (rev ? n - 1 - idx : idx));
        }
        const uint64_t ent = __shfl_sync(kFull, cached, next & 31);
        rem = (int)(ent & kLenMask);
        const uint64_t off = ent >> kLenBits;
        zero = off == kZeroOff;
        pos = rev ? (long long)off + rem - 1 : (long long)off;
        ++next;
        return true;
    }
};

struct ScoreArgs {
    const int32_t* cov;
    const uint64_t* orf_desc;   // indexed by absolute ORF id
    const uint64_t* exon_entries;
    long long orf_lo, orf_hi;
    unsigned long long* work_counter;
    rt_score_params prm;
    rt_score_out out;           // element k <-> ORF orf_lo + k
};

__global__ void __launch_bounds__(kScoreWarps * 32)
score_orfs_kernel(const ScoreArgs args) {
    __shared__ __align__(16) int32_t s_buf[kScoreWarps][kBufNt];
    const int lane = threadIdx.x & 31;
    int32_t* buf = s_buf[threadIdx.x >> 5];
    const double kSqrt3 = 1.7320508075688772;

    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(args.work_counter, (unsigned long long)kFetchBatch);
        base = __shfl_sync(kFull, base, 0);
        if ((long long)base + args.orf_lo >= args.orf_hi) break;

        for (int bi = 0; bi < kFetchBatch; ++bi) {
            const long long k_out = (long long)base + bi;
            const long long orf = args.orf_lo + k_out;
            if (orf >= args.orf_hi) break;

            const uint64_t desc = __ldg(args.orf_desc + orf);
            ExonCursor cur;
            cur.entries = args.exon_entries + (desc & kBeginMask);
            cur.n = (int)((desc >> 40) & kMaxEntriesPerOrf);
            cur.rev = (desc >> 63) != 0;

            FrameAcc f0, f1, f2;
            long long count = 0;
            long long min_codon = 0x7fffffffffffffffll;
            long long total = 0;   // profile length so far
            int fill = 0;
            bool last = false;

            while (!last) {
                // ---- K2: stage the next profile tile into shared memory ----
                while (fill < kTileNt + 2) {
                    if (cur.rem == 0 && !cur.advance(lane)) { last = true; break; }
                    const int take = min(cur.rem, kTileNt + 2 - fill);
                    if (cur.zero) {
                        for (int k = lane; k < take; k += 32) buf[fill + k] = 0;
                    } else if (cur.rev) {
                        const int32_t* src = args.cov + cur.pos;
                        for (int k = lane; k < take; k += 32) buf[fill + k] = ld_cov(src - k);
                    } else {
                        const int32_t* src = args.cov + cur.pos;
                        for (int k = lane; k < take; k += 32) buf[fill + k] = ld_cov(src + k);
                    }
                    fill += take;
                    total += take;
                    cur.rem -= take;
                    cur.pos += cur.rev ? -take : take;
                }
                const int nvals = fill;
                if (last && lane < 6) buf[nvals + lane] = 0;   // nvals + 5 < kBufNt
                __syncwarp();

                // ---- K3: one lane per codon; frames are fixed per register set ----
                const int ncod = last ? (nvals + 2) / 3 : kTileCodons;
                for (int c = lane; c < ncod; c += 32) {
                    const int32_t* p = buf + 3 * c;
                    const int v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3], v4 = p[4];
                    const long long cs = (long long)v0 + v1 + v2;     // common.py:177-179
                    count += cs;                                      // detect_orfs.py:278
                    min_codon = cs < min_codon ? cs : min_codon;
                    accumulate_codon(v0, v1, v2, 3 * c + 2 < nvals, f0);   // statistics.py:71
                    accumulate_codon(v1, v2, v3, 3 * c + 3 < nvals, f1);
                    accumulate_codon(v2, v3, v4, 3 * c + 4 < nvals, f2);
                }
                __syncwarp();
                if (!last) {   // carry the 2-value halo to the front of the next tile
                    int t = 0;
                    if (lane < 2) t = buf[kTileNt + lane];
                    __syncwarp();
                    if (lane < 2) buf[lane] = t;
                    fill = 2;
                    total -= 0;
                    __syncwarp();
                }
            }

            // ---- warp reductions ----
            FrameAcc* fr[3] = {&f0, &f1, &f2};
            int K[3], M[3];
            double re[3], im[3];
            const bool any_general = __any_sync(kFull, (f0.ng | f1.ng | f2.ng) != 0);
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                K[f] = __reduce_add_sync(kFull, fr[f]->K);
                const int na = __reduce_add_sync(kFull, fr[f]->na);
                const int nb = __reduce_add_sync(kFull, fr[f]->nb);
                const int nc = __reduce_add_sync(kFull, fr[f]->nc);
                const int ng = __reduce_add_sync(kFull, fr[f]->ng);
                double sre = 0.0, sim = 0.0;
                if (any_general) {
                    sre = warp_sum_f64(fr[f]->sre);
                    sim = warp_sum_f64(fr[f]->sim);
                }
                M[f] = na + nb + nc + ng;
                re[f] = sre + ((double)na - 0.5 * ((double)nb + (double)nc));
                im[f] = kSqrt3 * (sim + 0.5 * ((double)nb - (double)nc));
            }
            count = warp_sum_i64(count);
            min_codon = warp_min_i64(min_codon);

            if (lane == 0) {
                // statistics.py:64-66,92-115: running maximum with the K==0 reset quirk
                double coh = 0.0;
                int valid = -1;
                double s3[3];
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (K[f] == 0) {
                        s3[f] = __longlong_as_double(0x7ff8000000000000ll);
                        coh = 0.0;
                        valid = 0;
                        continue;
                    }
                    // Cxy(1/3) = |sum u|^2 / (K * M); 0/0 -> NaN never wins (all codons uniform)
                    const double s = (re[f] * re[f] + im[f] * im[f]) / ((double)K[f] * (double)M[f]);
                    s3[f] = s;
                    if (s > coh) { coh = s; valid = K[f]; }
                    if (valid == -1) valid = K[f];
                }
                const double score = sqrt(coh);
                const long long L = total;
                const long long n_codons = L / 3 > 1 ? L / 3 : 1;            // detect_orfs.py:281
                const double ratio = (double)valid / (double)n_codons;      // :285
                const double density = (double)count / (double)n_codons;    // :287
                const bool ok = score >= args.prm.phase_score_cutoff &&
                                (double)valid >= args.prm.min_valid_codons &&
                                (L == 0 || (double)min_codon >= args.prm.min_reads_per_codon) &&
                                ratio >= args.prm.min_valid_codons_ratio &&
                                density >= args.prm.min_density_over_orf;   // :289-299
                args.out.score[k_out] = score;
                args.out.valid[k_out] = valid;
                args.out.count[k_out] = count;
                args.out.length[k_out] = (int32_t)L;
                if (args.out.min_codon)
                    args.out.min_codon[k_out] =
                        L == 0 ? 0 : (min_codon > 0x7fffffffll ? 0x7fffffff : (int32_t)min_codon);
                if (args.out.status) args.out.status[k_out] = ok ? 1 : 0;
                if (args.out.frame_K) {
                    args.out.frame_K[3 * k_out + 0] = K[0];
                    args.out.frame_K[3 * k_out + 1] = K[1];
                    args.out.frame_K[3 * k_out + 2] = K[2];
                }
                if (args.out.frame_s) {
                    args.out.frame_s[3 * k_out + 0] = s3[0];
                    args.out.frame_s[3 * k_out + 1] = s3[1];
                    args.out.frame_s[3 * k_out + 2] = s3[2];
                }
            }
            __syncwarp();
        }
    }
}

// ---- K4 -------------------------------------------------------------------------------------
struct GatherArgs {
    const int32_t* cov;
    const uint64_t* orf_desc;
    const uint64_t* exon_entries;
    const int64_t* orf_ids;
    const int64_t* out_ptr;
    int32_t* out;
    long long n_sel;
};

__global__ void __launch_bounds__(256) gather_profiles_kernel(const GatherArgs args) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < args.n_sel; i += nwarps) {
        const uint64_t desc = __ldg(args.orf_desc + args.orf_ids[i]);
        ExonCursor cur;
        cur.entries = args.exon_entries + (desc & kBeginMask);
        cur.n = (int)((desc >> 40) & kMaxEntriesPerOrf);
        cur.rev = (desc >> 63) != 0;
        int32_t* dst = args.out + args.out_ptr[i];
        while (cur.advance(lane)) {
            const int take = cur.rem;
            if (cur.zero) {
                for (int k = lane; k < take; k += 32) dst[k] = 0;
            } else if (cur.rev) {
                const int32_t* src = args.cov + cur.pos;
                for (int k = lane; k < take; k += 32) dst[k] = ld_cov(src - k);
            } else {
                const int32_t* src = args.cov + cur.pos;
                for (int k = lane; k < take; k += 32) dst[k] = ld_cov(src + k);
            }
            dst += take;
            cur.rem = 0;
        }
    }
}

// ---- K1 -------------------------------------------------------------------------------------
struct BinArgs {
    int32_t* cov;
    const int32_t* ref_id;
    const int32_t* first;
    const int32_t* last;
    const uint16_t* mlen;
    const uint16_t* flag;
    const uint8_t* mapq;
    const uint8_t* nh;
    long long n;
    int protocol;
    int weight;                    // +1 bin, -1 un-bin
    const int32_t* len_table;      // RT_LEN_TABLE
    const long long* contig_base;  // n_contig
    const long long* contig_len;   // n_contig
    int n_contig;
    int pad;
    long long plane;
    unsigned long long* stats;       // RT_N_STATS
    unsigned long long* len_counts;  // RT_LEN_TABLE
};

constexpr int kBinThreads = 256;
constexpr int kBinReadsPerThread = 8;
constexpr int kLenHist = 512;   // read lengths below this are histogrammed in shared memory

// Category of one read after the cascade of bam.py:77-91 and the validity test of :133.
// 0..5 map to RT_ST_QCFAIL..RT_ST_MULTI (offset by 1), 6 = valid, 7 = bad ref, 8 = ignored.
__device__ __forceinline__ int classify_read(unsigned fl, unsigned mapq, unsigned nh) {
    if (fl & 0x200) return RT_ST_QCFAIL;      // bam.py:77
    if (fl & 0x400) return RT_ST_DUPLICATE;   // bam.py:80
    if (fl & 0x100) return RT_ST_SECONDARY;   // bam.py:83
    if (fl & 0x4) return RT_ST_UNMAPPED;      // bam.py:86
    // is_read_uniq_mapping, common.py:33-69 (None is falsy -> counted as multi, bam.py:89)
    bool uniq;
    if (nh != 0) uniq = nh == 1;              // common.py:54-56
    else uniq = mapq == 255;                  // common.py:59-69: every other branch is falsy
    return uniq ? RT_ST_VALID : RT_ST_MULTI;
}

__global__ void __launch_bounds__(kBinThreads) bin_psites_kernel(const BinArgs a) {
    __shared__ unsigned int s_stats[RT_N_STATS];
    __shared__ unsigned int s_len[kLenHist];
    for (int i = threadIdx.x; i < kLenHist; i += kBinThreads) s_len[i] = 0;
    if (threadIdx.x < RT_N_STATS) s_stats[threadIdx.x] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long block_base = (long long)blockIdx.x * (kBinThreads * kBinReadsPerThread);
#pragma unroll 2
    for (int it = 0; it < kBinReadsPerThread; ++it) {
        const long long i = block_base + (long long)it * kBinThreads + threadIdx.x;
        const bool in = i < a.n;
        int cat = -1;       // -1: no read
        int len = -1;       // >= 0: counts in read_length_counts
        if (in) {
            const unsigned fl = a.flag[i];
            cat = classify_read(fl, a.mapq[i], a.nh[i]);
            if (cat == RT_ST_VALID) {
                const int l = a.mlen[i];                        // bam.py:99
                const int mode = __ldg(a.len_table + l);
                const bool rev = (fl & 0x10) != 0;              // bam.py:94
                int strand;
                long long pos;
                if (a.protocol == RT_PROTOCOL_FORWARD) {        // bam.py:105-117
                    strand = rev ? 1 : 0;
                    pos = rev ? a.last[i] : a.first[i];
                } else {                                        // bam.py:118-131
                    strand = rev ? 0 : 1;
                    pos = rev ? a.first[i] : a.last[i];
                }
                const int c = a.ref_id[i];
                if (mode == RT_LEN_FILTERED || a.protocol > RT_PROTOCOL_REVERSE) {
                    cat = -2;                                   // bam.py:101 / no protocol branch
                } else if (c < 0 || c >= a.n_contig) {
                    cat = RT_ST_BADREF;                         // chrom is None, bam.py:133
                } else {
                    len = l;                                    // bam.py:136
                    if (mode >= 0) {                            // detect_orfs.py:74
                        const long long p = pos + 1 + (strand == 0 ? mode : -mode);   // bam.py:135, detect_orfs.py:78-81
                        if (p < 1 - a.pad || p > a.contig_len[c] + a.pad) {
                            atomicAdd(&s_stats[RT_ST_OOB], 1u);
                        } else {
                            atomicAdd(a.cov + ((long long)strand * a.plane + a.contig_base[c] + a.pad + p), a.weight);   // detect_orfs.py:82
                        }
                    }
                }
            }
        }
        // warp-aggregated counters (bam.py:61,73-91,137)
        const unsigned m_in = __ballot_sync(kFull, cat != -1);
        if (m_in == 0) continue;
#pragma unroll
        for (int k = RT_ST_QCFAIL; k <= RT_ST_VALID; ++k) {
            const unsigned m = __ballot_sync(kFull, cat == k);
            if (lane == 0 && m) atomicAdd(&s_stats[k], (unsigned)__popc(m));
        }
        const unsigned m_bad = __ballot_sync(kFull, cat == RT_ST_BADREF);
        if (lane == 0) {
            atomicAdd(&s_stats[RT_ST_TOTAL], (unsigned)__popc(m_in));
            if (m_bad) atomicAdd(&s_stats[RT_ST_BADREF], (unsigned)__popc(m_bad));
        }
        // read_length_counts (bam.py:136): one leader per distinct length in the warp
        unsigned todo = __ballot_sync(kFull, len >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int l = __shfl_sync(kFull, len, leader);
            const unsigned same = __ballot_sync(kFull, len == l);
            if (lane == leader) {
                if (l < kLenHist) atomicAdd(&s_len[l], (unsigned)__popc(same));
                else atomicAdd(a.len_counts + l, (unsigned long long)((long long)a.weight * __popc(same)));
            }
            todo &= ~same;
        }
    }
    __syncthreads();
    // RT_ST_VALID in s_stats counts reads that passed the cascade; badref/ignored were re-labelled above
    // two's-complement wrap makes weight = -1 subtract
    if (threadIdx.x < RT_N_STATS && s_stats[threadIdx.x])
        atomicAdd(a.stats + threadIdx.x, (unsigned long long)((long long)a.weight * s_stats[threadIdx.x]));
    for (int i = threadIdx.x; i < kLenHist; i += kBinThreads)
        if (s_len[i]) atomicAdd(a.len_counts + i, (unsigned long long)((long long)a.weight * s_len[i]));
}

}  // namespace rt
