"""Throughput of the native BAM/BGZF decoder (rt_bam_load: threaded inflate + record walk) and of the hand-over as a
record stream (rt_bam_stream), on a synthetic coordinate-sorted BAM with Illumina-like names, random bases and
four-valued qualities (so that the BGZF blocks compress like a real Ribo-seq library: 35-40 B/read).  CPU only.

usage: bam_decode_time.py [n_reads] [bam_path]     (an existing bam_path is reused)"""
import ctypes as C
import os, struct, sys, tempfile, time, zlib
from multiprocessing import Pool
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bam_writer as W


def synth_stream(n, refs, seed=1):
    """The uncompressed BAM byte stream of n single-M-CIGAR reads, built column-wise with numpy."""
    rng = np.random.default_rng(seed)
    ref = np.sort(rng.integers(0, len(refs), n)).astype(np.int32)
    pos = np.zeros(n, np.int32)
    for r in range(len(refs)):
        m = ref == r
        pos[m] = np.sort(rng.integers(0, refs[r][1] - 100, int(m.sum())))
    length = rng.integers(26, 33, n).astype(np.int32)
    flag = np.where(rng.random(n) < 0.5, 16, 0).astype(np.uint16)
    name_w = 20                                                   # "SRR1234567.%09d\0"
    size = 36 + name_w + 4 + (length + 1) // 2 + length + 4       # block_size field + body (NH:C:1 = 4 bytes)
    off = np.concatenate([[0], np.cumsum(size)]).astype(np.int64)
    out = np.zeros(int(off[-1]), np.uint8)
    serial = rng.permutation(n)
    digits = np.array([(serial // 10 ** k) % 10 for k in range(8, -1, -1)], np.uint8).T + ord("0")
    for L in range(26, 33):
        idx = np.flatnonzero(length == L)
        k = len(idx)
        rec = np.zeros((k, int(36 + name_w + 4 + (L + 1) // 2 + L + 4)), np.uint8)
        fixed = np.zeros(k, dtype=[("bs", "<u4"), ("ref", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                                   ("n_cig", "<u2"), ("flag", "<u2"), ("l_seq", "<i4"), ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
        fixed["bs"] = rec.shape[1] - 4; fixed["ref"] = ref[idx]; fixed["pos"] = pos[idx]; fixed["l_name"] = name_w
        fixed["mapq"] = 255; fixed["bin"] = 4680; fixed["n_cig"] = 1; fixed["flag"] = flag[idx]; fixed["l_seq"] = L
        fixed["nref"] = -1; fixed["npos"] = -1
        rec[:, :36] = fixed.view(np.uint8).reshape(k, 36)
        rec[:, 36:46] = np.frombuffer(b"SRR1234567", np.uint8); rec[:, 46] = ord("."); rec[:, 47:56] = digits[idx]
        c = 36 + name_w
        rec[:, c:c + 4] = np.frombuffer(struct.pack("<I", (L << 4) | 0), np.uint8)
        c += 4
        nb = (L + 1) // 2
        rec[:, c:c + nb] = (1 << rng.integers(0, 4, (k, nb))) << 4 | (1 << rng.integers(0, 4, (k, nb)))
        c += nb
        rec[:, c:c + L] = np.array([2, 14, 27, 37], np.uint8)[rng.choice(4, (k, L), p=[0.03, 0.07, 0.2, 0.7])]
        c += L
        rec[:, c:c + 4] = np.frombuffer(b"NHC\x01", np.uint8)
        out[(off[idx][:, None] + np.arange(rec.shape[1])[None, :]).ravel()] = rec.ravel()
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(("@SQ\tSN:%s\tLN:%d\n" % (nm, ln)).encode() for nm, ln in refs)
    head = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for nm, ln in refs:
        nb = nm.encode() + b"\0"
        head += struct.pack("<I", len(nb)) + nb + struct.pack("<I", ln)
    return head + out.tobytes()


def write_bgzf(path, stream, payload=0xFF00):
    with Pool(os.cpu_count()) as pool, open(path, "wb") as fh:
        for blk in pool.imap(W.bgzf_block, (stream[i:i + payload] for i in range(0, len(stream), payload)), chunksize=64):
            fh.write(blk)
        fh.write(W.BGZF_EOF)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(tempfile.mkdtemp(prefix="rt_bam_"), "lib.bam")
    if not os.path.exists(path):
        t0 = time.perf_counter()
        stream = synth_stream(n, [("chr%d" % i, 50_000_000) for i in range(1, 5)])
        write_bgzf(path, stream)
        print(f"wrote {n} reads, {len(stream) / 1e6:.0f} MB of BAM records as {os.path.getsize(path) / 1e6:.1f} MB of BGZF "
              f"in {time.perf_counter() - t0:.1f} s")
    size = os.path.getsize(path)
    from ribotricer_b200 import _lib
    lib = _lib.load()
    for thr in (1, 2, 4, 8, 0):
        h = C.c_void_p()
        t0 = time.perf_counter()
        assert lib.rt_bam_load(path.encode(), thr, C.byref(h)) == 0, lib.rt_bam_last_error()
        t1 = time.perf_counter()
        nb = C.c_int64(0)
        assert lib.rt_bam_stream(h, thr, 0, None, None, C.byref(nb)) == 0
        rec = np.empty(nb.value * _lib.RT_STREAM_BLOCK, np.uint32); hdr = np.empty(nb.value * 4, np.int32)
        assert lib.rt_bam_stream(h, thr, nb.value, rec.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p), C.byref(nb)) == 0
        t2 = time.perf_counter()
        nr = lib.rt_bam_n_reads(h)
        lib.rt_bam_free(h)
        print(f"threads {thr or os.cpu_count()}: rt_bam_load {1e3 * (t1 - t0):.0f} ms = {nr / (t1 - t0) / 1e6:.1f} M reads/s, "
              f"{size / (t1 - t0) / 1e6:.0f} MB/s of BGZF; rt_bam_stream {1e3 * (t2 - t1):.0f} ms ({nb.value} blocks)")
