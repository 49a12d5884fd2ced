"""ctypes binding of oracle/rt_oracle.c (dense-array CPU restatement).

TEST INFRASTRUCTURE ONLY (checker; see rt_oracle.c's header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librt_oracle.so")
ST_NAMES = ("total", "qcfail", "duplicate", "secondary", "unmapped", "multi", "valid", "oob", "badref")
LEN_TABLE = 65536
LEN_UNUSED, LEN_FILTERED = -2147483648, -2147483647

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "rt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "librt_oracle.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_phasescore.restype = C.c_double
        _lib.orc_select.restype = C.c_double
        _lib.orc_profile.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_len_table(psite_offsets: dict | None, read_lengths=None) -> np.ndarray:
    """len -> offset (any sign), LEN_UNUSED (kept, not merged) or LEN_FILTERED."""
    t = np.full(LEN_TABLE, LEN_UNUSED, dtype=np.int32)
    if read_lengths is not None:
        t[:] = LEN_FILTERED
        for length in read_lengths:
            t[int(length)] = LEN_UNUSED
    for length, off in (psite_offsets or {}).items():
        if t[int(length)] != LEN_FILTERED:
            t[int(length)] = int(off)
    return t


def genome_layout(contig_len, pad: int):
    """Shared layout rule (include/ribotricer_b200.h): 32-element aligned strides."""
    contig_len = np.asarray(contig_len, dtype=np.int64)
    stride = (contig_len + 2 * pad + 1 + 31) // 32 * 32
    base = np.zeros(len(contig_len), dtype=np.int64)
    if len(contig_len):
        base[1:] = np.cumsum(stride)[:-1]
    plane = int(stride.sum()) if len(contig_len) else 32
    return base, plane


def phasescore(cov):
    a = np.ascontiguousarray(cov, dtype=np.int32)
    v = C.c_int32()
    s = lib().orc_phasescore(_p(a), C.c_int64(len(a)), C.byref(v))
    return s, v.value


def frame_spectra(cov):
    a = np.ascontiguousarray(cov, dtype=np.int32)
    K = np.zeros(3, np.int32)
    s = np.zeros(3, np.float64)
    lib().orc_frame_spectra(_p(a), C.c_int64(len(a)), _p(K), _p(s))
    return K, s


def bin_reads(cols, protocol: int, len_table, contig_base, contig_len, pad, plane, cov=None):
    n = len(cols["ref_id"])
    if cov is None:
        cov = np.zeros(2 * plane, dtype=np.int32)
    stats = np.zeros(len(ST_NAMES), dtype=np.int64)
    len_counts = np.zeros(LEN_TABLE, dtype=np.int64)
    c = {k: np.ascontiguousarray(cols[k], dtype=dt) for k, dt in (
        ("ref_id", np.int32), ("first", np.int32), ("last", np.int32), ("mlen", np.uint16),
        ("flag", np.uint16), ("mapq", np.uint8), ("nh", np.uint8))}
    cb = np.ascontiguousarray(contig_base, np.int64)
    cl = np.ascontiguousarray(contig_len, np.int64)
    lt = np.ascontiguousarray(len_table, np.int32)
    lib().orc_bin_reads(C.c_int64(n), _p(c["ref_id"]), _p(c["first"]), _p(c["last"]), _p(c["mlen"]),
                        _p(c["flag"]), _p(c["mapq"]), _p(c["nh"]), C.c_int(protocol), _p(lt),
                        C.c_int(len(cl)), _p(cb), _p(cl), C.c_int(pad), C.c_int64(plane), _p(cov),
                        _p(stats), _p(len_counts))
    return cov, dict(zip(ST_NAMES, stats.tolist())), len_counts


def score(index, cov, contig_base, contig_len, pad, plane, params, lo=0, hi=None, nthreads=0,
          diagnostics=True):
    """index: dict with exon_ptr(i64), exon_start(i32), exon_end(i32), orf_contig(i32), orf_strand(u8)."""
    hi = len(index["orf_contig"]) if hi is None else hi
    n = hi - lo
    out = dict(score=np.zeros(n, np.float64), valid=np.zeros(n, np.int32), count=np.zeros(n, np.int64),
               length=np.zeros(n, np.int32), min_codon=np.zeros(n, np.int32), status=np.zeros(n, np.uint8))
    K3 = np.zeros((n, 3), np.int32) if diagnostics else None
    s3 = np.zeros((n, 3), np.float64) if diagnostics else None
    prm = np.ascontiguousarray(params, np.float64)
    cb = np.ascontiguousarray(contig_base, np.int64)
    cl = np.ascontiguousarray(contig_len, np.int64)
    lib().orc_score(C.c_int64(lo), C.c_int64(hi), _p(index["exon_ptr"]), _p(index["exon_start"]),
                    _p(index["exon_end"]), _p(index["orf_contig"]), _p(index["orf_strand"]), _p(cb), _p(cl),
                    C.c_int(pad), C.c_int64(plane), _p(cov), _p(prm), _p(out["score"]), _p(out["valid"]),
                    _p(out["count"]), _p(out["length"]), _p(out["min_codon"]), _p(out["status"]),
                    _p(K3), _p(s3), C.c_int(nthreads))
    if diagnostics:
        out["frame_K"], out["frame_s"] = K3, s3
    return out


def gather_profiles(index, orf_ids, cov, contig_base, contig_len, pad, plane):
    orf_ids = np.ascontiguousarray(orf_ids, np.int64)
    ep = index["exon_ptr"]
    lens = np.zeros(len(orf_ids), np.int64)
    exlen = (index["exon_end"].astype(np.int64) - index["exon_start"] + 1)
    cs = np.concatenate([[0], np.cumsum(exlen)])
    lens = cs[ep[orf_ids + 1]] - cs[ep[orf_ids]]
    out_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = np.zeros(int(out_ptr[-1]), np.int32)
    cb = np.ascontiguousarray(contig_base, np.int64)
    cl = np.ascontiguousarray(contig_len, np.int64)
    lib().orc_gather_profiles(C.c_int64(len(orf_ids)), _p(orf_ids), _p(out_ptr), _p(ep),
                              _p(index["exon_start"]), _p(index["exon_end"]), _p(index["orf_contig"]),
                              _p(index["orf_strand"]), _p(cb), _p(cl), C.c_int(pad), C.c_int64(plane),
                              _p(cov), _p(out))
    return out_ptr, out


def tie_mask(frame_K, frame_s, tol=1e-12):
    """Vectorised H1 classifier (same rule as oracle_py.is_frame_tie): True where the running
    maximum of statistics.py:109 is decided between frames whose scores agree within ``tol``
    while their K differ -- the reference's valid_codons is then SciPy rounding noise."""
    n = len(frame_K)
    best = np.zeros(n)
    valid = np.full(n, -1, np.int64)
    tie = np.zeros(n, bool)
    for f in range(3):
        k = frame_K[:, f].astype(np.int64)
        s = frame_s[:, f]
        zero = k == 0
        ok = ~zero & ~np.isnan(s)
        with np.errstate(invalid="ignore"):
            win = ok & (s > best + tol)
            close = ok & ~win & (np.abs(s - best) <= tol)
        amb = close & (valid != -1) & (valid != k)
        tie = np.where(zero | win, False, tie | amb)
        best = np.where(zero, 0.0, np.where(win | close, np.fmax(best, np.where(ok, s, best)), best))
        valid = np.where(zero, 0, np.where(win, k, valid))
        valid = np.where(~zero & (valid == -1), k, valid)
    return tie
