"""Where the end-to-end time goes: the host-buffer entry points timed one by one on C2 (wall clock around synchronous
calls), the streamed rt_bin_reads_host at several pipeline widths, and rt_stream_pack alone at several thread counts."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine, ScoreParams
cfg = synth.config("C2"); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
eng.set_index(**idx.as_dict()); eng.set_layout("compact")
d = synth.make_reads(cfg, idx, device="cuda")
h = {k: v.cpu().pin_memory() for k, v in d.items()}
del d; torch.cuda.empty_cache()
hn = {k: v.numpy() for k, v in h.items()}
stream = eng.stream_reads(hn, pinned=True)
cov = eng.new_coverage(); hout = eng.new_host_score_columns(idx.n_orf)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("cpus", os.cpu_count())
print("stream_host ms", t(lambda: (eng.clear_coverage(cov), eng.bin_stream_host(cov, stream, "forward"))))
for pipes in (4, 8, 12, 16):
    os.environ["RT_PACK_PIPES"] = str(pipes)
    print(f"columns sorted_hint, {pipes} pipes ms", t(lambda: (eng.clear_coverage(cov), eng.bin_reads_host(cov, h, "forward", sorted_hint=True))))
os.environ.pop("RT_PACK_PIPES")
print("columns plain ms", t(lambda: (eng.clear_coverage(cov), eng.bin_reads_host(cov, h, "forward", sorted_hint=False))))
print("score_host ms", t(lambda: eng.score_host(cov, 0, idx.n_orf, ScoreParams(), out=hout)))
lib = eng.lib
n = len(hn["ref_id"])
cols = [hn[k] if hn[k].dtype.itemsize != 2 else hn[k].view(np.uint16) for k in ("ref_id", "first", "last", "mlen", "flag", "mapq", "nh")]
ptrs = [c.ctypes.data_as(C.c_void_p) for c in cols]
rec = np.empty(stream["n_blocks"] * 256, np.uint32); hdr = np.empty(stream["n_blocks"] * 4, np.int32); nb = C.c_int64(0)
rec[:] = 0
for thr in (4, 8, 16):
    t0 = time.perf_counter(); lib.rt_stream_pack(n, *ptrs, thr, 0, None, None, C.byref(nb)); t1 = time.perf_counter()
    lib.rt_stream_pack(n, *ptrs, thr, nb.value, rec.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p), C.byref(nb)); t2 = time.perf_counter()
    print(f"rt_stream_pack {thr} threads: count pass {1e3 * (t1 - t0):.1f} ms, count + write passes {1e3 * (t2 - t1):.1f} ms")
a = np.empty(1 << 28, np.uint8); b = np.empty(1 << 28, np.uint8); a[:] = 1; b[:] = 2
t0 = time.perf_counter(); np.copyto(b, a); print(f"one-thread memcpy {2 * (1 << 28) / (time.perf_counter() - t0) / 1e9:.1f} GB/s (read + write)")
