"""Host text stages on a C2-like index at 10 % scale (250 k ORFs), CPU only: rt_index_load (file read + chunked parse) and
rt_tsv_write (every row reported, 96 M profile values, 10 % of them non-zero) into /dev/null and into a file."""
import ctypes as C, time, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ribotricer_b200 import synth, _lib
from ribotricer_b200.index import load_native_index
cfg = synth.config("C2", 0.1, 0.1)
idx = synth.make_index(cfg)
path = "/tmp/idx_c2_01.tsv"
if not os.path.exists(path): idx.write_tsv(path)
print("index", idx.n_orf, os.path.getsize(path)/1e6, "MB")
for i in range(3):
    t0=time.perf_counter(); nat = load_native_index(path); t=time.perf_counter()-t0
    print(f"rt_index_load {t*1e3:.0f} ms = {os.path.getsize(path)/t/1e6:.0f} MB/s")
lib = _lib.load()
n = idx.n_orf
d = idx.as_dict()
ptr = d["exon_ptr"]; L = np.zeros(n, np.int64)
ex_len = (d["exon_end"] - d["exon_start"] + 1).astype(np.int64)
L = np.add.reduceat(ex_len, ptr[:-1])
rng = np.random.default_rng(0)
prof_ptr = np.concatenate([[0], np.cumsum(L)]).astype(np.int64)
prof = (rng.random(int(prof_ptr[-1])) < 0.1).astype(np.int32) * rng.integers(1, 40, int(prof_ptr[-1]), dtype=np.int32)
sel = np.arange(n, dtype=np.int64)
score = rng.random(n); valid = rng.integers(0, 100, n, dtype=np.int32); count = rng.integers(0, 10000, n).astype(np.int64)
length = L.astype(np.int32); status = np.ones(n, np.uint8)
p = lambda a: a.ctypes.data_as(C.c_void_p)
for out in ("/dev/null", "/tmp/out.tsv"):
  for i in range(2):
    h = C.c_void_p(); assert lib.rt_tsv_open(out.encode(), 1, C.byref(h)) == 0
    t0=time.perf_counter()
    assert lib.rt_tsv_write(h, nat.handle, n, p(sel), 0, p(score), p(valid), p(count), p(length), p(status), p(prof_ptr), p(prof)) == 0
    lib.rt_tsv_close(h); t=time.perf_counter()-t0
    sz = os.path.getsize("/tmp/out.tsv") if out != "/dev/null" else 0
    print(f"rt_tsv_write -> {out}: {t*1e3:.0f} ms, {prof_ptr[-1]/1e6:.0f} M profile values, {sz/1e6:.0f} MB")
