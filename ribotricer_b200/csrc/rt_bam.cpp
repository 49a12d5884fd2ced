// rt_bam.cpp -- native BAM/BGZF decode straight to the read columns of the C ABI
// (SURVEY.md 8(f) "next #2"; replaces the two pysam passes of bam.py:65-71 on the host).
//
// Per alignment record it produces what split_bam reads through pysam:
//   ref_id  = reference_id                       flag = flag        mapq = mapping_quality
//   first / last / mlen = get_reference_positions()[0] / [-1] / len()   (bam.py:95-99): reference
//             positions of M, = and X operations only (D and N advance the reference, S I H P do not)
//   nh      = the NH aux tag as is_read_uniq_mapping sees it (common.py:53-56: `dict(get_tags())["NH"] == 1`):
//             0 = absent, 1 = present and equal to 1, anything else = present and different from 1
//             (2..254 the value itself; 255 for values <= 0, > 254 or of a non-numeric type).  When the tag
//             occurs twice the LAST one counts, as in dict().
//   pos / ref_end = reference_start / reference_end of infer_protocol.py:84-85 (ref_end = -1 for None:
//             unmapped flag or no CIGAR; else pos + reference length of the CIGAR, at least 1 as in htslib)
// BGZF blocks are inflated in parallel (zlib raw inflate), records are cut sequentially and
// decoded in parallel.  No GPU is involved.
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "ribotricer_b200.h"

struct rt_bam {
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::vector<int32_t> ref_id, first, last, pos, ref_end;
    std::vector<uint16_t> mlen, flag;
    std::vector<uint8_t> mapq, nh;
    bool sorted = false;
};

namespace {

thread_local std::string g_bam_error;

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }

struct Block { size_t in_off, in_len; size_t out_off; uint32_t out_len; };

// One BGZF member: gzip header with a 'BC' extra subfield holding the block size (SAM spec 4.1).
bool parse_block(const uint8_t* base, size_t size, size_t off, Block& b, size_t& next) {
    if (off + 18 > size) return false;
    const uint8_t* p = base + off;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
    const uint16_t xlen = rd16(p + 10);
    if (off + 12 + xlen > size) return false;
    int bsize = -1;
    for (size_t x = 12; x + 4 <= 12u + xlen;) {
        const uint16_t slen = rd16(p + x + 2);
        if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) bsize = rd16(p + x + 4);
        x += 4 + slen;
    }
    if (bsize < 0) return false;
    const size_t total = (size_t)bsize + 1;
    if (off + total > size || total < 12u + xlen + 8) return false;
    b.in_off = off + 12 + xlen;
    b.in_len = total - (12 + xlen) - 8;
    b.out_len = rd32(p + total - 4);
    next = off + total;
    return true;
}

bool inflate_block(const uint8_t* src, size_t n, uint8_t* dst, uint32_t out_len) {
    if (out_len == 0) return true;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(src);
    zs.avail_in = (uInt)n;
    zs.next_out = dst;
    zs.avail_out = out_len;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    return rc == Z_STREAM_END && zs.total_out == out_len;
}

void parallel_for(int n_threads, size_t n, const std::function<void(size_t, size_t)>& fn) {
    n_threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, n));
    if (n_threads == 1) { fn(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const size_t lo = std::min(n, (size_t)t * per), hi = std::min(n, lo + per);
        if (lo < hi) th.emplace_back(fn, lo, hi);
    }
    for (auto& x : th) x.join();
}

// Decode one alignment record (after its block_size field) into slot i of the columns.
// Returns false when the fixed fields, the CIGAR or an aux field run past the record (untrusted input).
bool decode_record(const uint8_t* r, uint32_t len, rt_bam& out, size_t i) {
    const int32_t ref = rdi32(r), pos = rdi32(r + 4);
    const uint32_t l_name = r[8];
    const uint32_t n_cigar = rd16(r + 12);
    const int32_t l_seq = rdi32(r + 16);
    if (l_seq < 0) return false;
    const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (fixed > len) return false;
    out.ref_id[i] = ref;
    out.mapq[i] = r[9];
    const uint16_t flag = rd16(r + 14);
    out.flag[i] = flag;
    const uint8_t* cig = r + 32 + l_name;
    int64_t cur = pos, first = 0, last = 0, n = 0;
    for (uint32_t k = 0; k < n_cigar; ++k) {
        const uint32_t v = rd32(cig + 4 * k), op = v & 15u, l = v >> 4;
        if (op == 0 || op == 7 || op == 8) {          // M = X
            if (n == 0) first = cur;
            last = cur + l - 1;
            n += l;
            cur += l;
        } else if (op == 2 || op == 3) {              // D N
            cur += l;
        }
    }
    out.first[i] = (int32_t)first;
    out.last[i] = (int32_t)last;
    out.mlen[i] = (uint16_t)std::min<int64_t>(n, 65535);
    out.pos[i] = pos;
    const int64_t rlen = cur - pos;
    out.ref_end[i] = ((flag & 4) || n_cigar == 0) ? -1 : (int32_t)(pos + (rlen ? rlen : 1));
    // aux fields: the NH tag as `dict(read.get_tags())["NH"] == 1` sees it
    const uint8_t* end = r + len;
    const uint8_t* a = r + fixed;
    int nh = 0;                                           // 0 absent, 1 equal to 1, else present and not 1
    while (a + 3 <= end) {
        const uint8_t t0 = a[0], t1 = a[1], ty = a[2];
        a += 3;
        size_t adv = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': adv = 1; break;
            case 's': case 'S': adv = 2; break;
            case 'i': case 'I': case 'f': adv = 4; break;
            case 'Z': case 'H': {
                const uint8_t* z = (const uint8_t*)memchr(a, 0, (size_t)(end - a));
                if (!z) return false;
                adv = (size_t)(z - a) + 1;
                break;
            }
            case 'B': {
                if (a + 5 > end) return false;
                const uint8_t sub = a[0];
                const uint32_t cnt = rd32(a + 1);
                const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                adv = 5 + (size_t)cnt * w;
                break;
            }
            default: return false;                    // unknown aux type
        }
        if (adv > (size_t)(end - a)) return false;
        if (t0 == 'N' && t1 == 'H') {
            long long val = 0;
            bool numeric = true;
            switch (ty) {
                case 'c': val = (int8_t)a[0]; break;
                case 'C': val = a[0]; break;
                case 's': val = (int16_t)rd16(a); break;
                case 'S': val = rd16(a); break;
                case 'i': val = rdi32(a); break;
                case 'I': val = rd32(a); break;
                case 'f': { float f; memcpy(&f, a, 4); val = f == 1.0f ? 1 : 0; break; }   // 1.0 == 1 in Python
                default: numeric = false;
            }
            nh = !numeric ? 255 : val == 1 ? 1 : (val >= 2 && val <= 254) ? (int)val : 255;
        }
        a += adv;
    }
    out.nh[i] = (uint8_t)nh;
    return true;
}

}  // namespace

extern "C" {

const char* rt_bam_last_error(void) { return g_bam_error.c_str(); }

int rt_bam_load(const char* path, int n_threads, rt_bam** out) {
    if (!path || !out) return RT_EINVAL;
    *out = nullptr;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { g_bam_error = std::string("cannot open ") + path; return RT_EINVAL; }
    struct stat st;
    fstat(fd, &st);
    const size_t size = (size_t)st.st_size;
    const uint8_t* base = size ? (const uint8_t*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (size && base == MAP_FAILED) { g_bam_error = "mmap failed"; return RT_EINVAL; }
    auto fail = [&](const std::string& msg) {
        if (size) munmap((void*)base, size);
        g_bam_error = msg;
        return RT_EINVAL;
    };
    rt_bam* bam = new rt_bam();
    std::vector<uint8_t> buf;          // uncompressed bytes not yet consumed + the current chunk
    size_t file_off = 0;
    bool header_done = false;
    const size_t kChunkOut = 256u << 20;
    std::vector<Block> blocks;
    std::vector<uint64_t> rec_off;
    while (file_off < size) {
        // 1. a chunk of BGZF blocks
        blocks.clear();
        size_t out_total = 0;
        while (file_off < size && out_total < kChunkOut) {
            Block b;
            size_t next;
            if (!parse_block(base, size, file_off, b, next)) { delete bam; return fail("not a BGZF/BAM file (bad block header)"); }
            b.out_off = out_total;
            out_total += b.out_len;
            blocks.push_back(b);
            file_off = next;
        }
        const size_t keep = buf.size();
        buf.resize(keep + out_total);
        std::atomic<bool> ok{true};
        parallel_for(n_threads, blocks.size(), [&](size_t lo, size_t hi) {
            for (size_t k = lo; k < hi; ++k)
                if (!inflate_block(base + blocks[k].in_off, blocks[k].in_len, buf.data() + keep + blocks[k].out_off,
                                   blocks[k].out_len))
                    ok = false;
        });
        if (!ok) { delete bam; return fail("BGZF inflate failed"); }
        size_t p = 0;
        const size_t n = buf.size();
        // 2. header (once): magic, text, references
        if (!header_done) {
            if (n < 12) continue;   // need more data
            if (memcmp(buf.data(), "BAM\1", 4) != 0) { delete bam; return fail("not a BAM file (bad magic)"); }
            const uint32_t l_text = rd32(buf.data() + 4);
            if (n < 12 + (size_t)l_text) continue;
            const std::string text((const char*)buf.data() + 8, l_text);
            const size_t hd = text.find("@HD");
            if (hd != std::string::npos) {
                const size_t eol = text.find('\n', hd);
                bam->sorted = text.substr(hd, eol - hd).find("SO:coordinate") != std::string::npos;
            }
            size_t q = 8 + l_text;
            const uint32_t n_ref = rd32(buf.data() + q);
            q += 4;
            bool complete = true;
            std::vector<std::string> names;
            std::vector<int64_t> lens;
            for (uint32_t r = 0; r < n_ref; ++r) {
                if (q + 4 > n) { complete = false; break; }
                const uint32_t l_name = rd32(buf.data() + q);
                if (q + 4 + l_name + 4 > n) { complete = false; break; }
                names.emplace_back((const char*)buf.data() + q + 4, l_name ? l_name - 1 : 0);
                lens.push_back(rd32(buf.data() + q + 4 + l_name));
                q += 8 + l_name;
            }
            if (!complete) continue;
            bam->ref_names.swap(names);
            bam->ref_lens.swap(lens);
            header_done = true;
            p = q;
        }
        // 3. cut complete records, decode them in parallel
        rec_off.clear();
        while (p + 4 <= n) {
            const uint32_t bs = rd32(buf.data() + p);
            if (bs < 32) { delete bam; return fail("corrupt BAM record"); }
            if (p + 4 + bs > n) break;
            rec_off.push_back(p);
            p += 4 + (size_t)bs;
        }
        const size_t base_i = bam->ref_id.size(), m = rec_off.size();
        bam->ref_id.resize(base_i + m); bam->first.resize(base_i + m); bam->last.resize(base_i + m);
        bam->mlen.resize(base_i + m); bam->flag.resize(base_i + m); bam->mapq.resize(base_i + m); bam->nh.resize(base_i + m);
        bam->pos.resize(base_i + m); bam->ref_end.resize(base_i + m);
        std::atomic<bool> rec_ok{true};
        parallel_for(n_threads, m, [&](size_t lo, size_t hi) {
            for (size_t k = lo; k < hi; ++k)
                if (!decode_record(buf.data() + rec_off[k] + 4, rd32(buf.data() + rec_off[k]), *bam, base_i + k))
                    rec_ok = false;
        });
        if (!rec_ok) { delete bam; return fail("corrupt BAM record"); }
        buf.erase(buf.begin(), buf.begin() + (ptrdiff_t)p);
    }
    if (size) munmap((void*)base, size);
    if (!header_done) { delete bam; g_bam_error = "truncated BAM header"; return RT_EINVAL; }
    if (!buf.empty()) { delete bam; g_bam_error = "truncated BAM record at end of file"; return RT_EINVAL; }
    *out = bam;
    return RT_OK;
}

void rt_bam_free(rt_bam* b) { delete b; }
int64_t rt_bam_n_reads(const rt_bam* b) { return b ? (int64_t)b->ref_id.size() : 0; }
int rt_bam_n_ref(const rt_bam* b) { return b ? (int)b->ref_names.size() : 0; }
const char* rt_bam_ref_name(const rt_bam* b, int i) {
    return b && i >= 0 && i < (int)b->ref_names.size() ? b->ref_names[i].c_str() : nullptr;
}
int64_t rt_bam_ref_len(const rt_bam* b, int i) { return b && i >= 0 && i < (int)b->ref_lens.size() ? b->ref_lens[i] : -1; }
int rt_bam_sorted(const rt_bam* b) { return b && b->sorted ? 1 : 0; }

int rt_bam_copy(const rt_bam* b, int32_t* ref_id, int32_t* first, int32_t* last, uint16_t* mlen, uint16_t* flag,
                uint8_t* mapq, uint8_t* nh) {
    if (!b) return RT_EINVAL;
    if (ref_id) std::copy(b->ref_id.begin(), b->ref_id.end(), ref_id);
    if (first) std::copy(b->first.begin(), b->first.end(), first);
    if (last) std::copy(b->last.begin(), b->last.end(), last);
    if (mlen) std::copy(b->mlen.begin(), b->mlen.end(), mlen);
    if (flag) std::copy(b->flag.begin(), b->flag.end(), flag);
    if (mapq) std::copy(b->mapq.begin(), b->mapq.end(), mapq);
    if (nh) std::copy(b->nh.begin(), b->nh.end(), nh);
    return RT_OK;
}

int rt_bam_copy_span(const rt_bam* b, int32_t* pos, int32_t* ref_end) {
    if (!b) return RT_EINVAL;
    if (pos) std::copy(b->pos.begin(), b->pos.end(), pos);
    if (ref_end) std::copy(b->ref_end.begin(), b->ref_end.end(), ref_end);
    return RT_OK;
}

int rt_bam_pack(const rt_bam* b, uint8_t* meta, int64_t run_cap, int64_t* run_start, int32_t* run_ref, int64_t* n_runs) {
    if (!b) return RT_EINVAL;
    return rt_pack_read_meta((int64_t)b->ref_id.size(), b->ref_id.data(), b->flag.data(), b->mapq.data(), b->nh.data(),
                             meta, run_cap, run_start, run_ref, n_runs);
}

int rt_bam_stream(const rt_bam* b, int n_threads, int64_t cap_blocks, uint32_t* records, int32_t* hdr, int64_t* n_blocks) {
    if (!b) return RT_EINVAL;
    return rt_stream_pack((int64_t)b->ref_id.size(), b->ref_id.data(), b->first.data(), b->last.data(), b->mlen.data(), b->flag.data(),
                          b->mapq.data(), b->nh.data(), n_threads, cap_blocks, records, hdr, n_blocks);
}

}  // extern "C"
