"""K1 on the resident record stream of C2, CUDA events: clear + rt_bin_stream against rt_bin_stream_fresh (zones)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = synth.config(name); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
eng.set_index(**idx.as_dict()); eng.set_layout("compact")
d = synth.make_reads(cfg, idx, device="cuda")
h = {k: v.cpu().numpy() for k, v in d.items()}
del d; torch.cuda.empty_cache()
ds = eng.upload_stream(eng.stream_reads(h))
cov = eng.new_coverage(); st, lc = eng.new_bin_accumulators()
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
lib = os.environ.get("RT_LIB_PATH", "default").split("/")[-1]
t_red = timed(lambda: (eng.clear_coverage(cov), eng.bin_stream_device(cov, ds, "forward", st, lc)))
want = cov.clone()
print(lib, "clear + rt_bin_stream ms", t_red)
if os.environ.get("RT_NO_FRESH") != "1":
    t_fresh = timed(lambda: eng.bin_stream_device(cov, ds, "forward", st, lc, fresh=True))
    print(lib, "rt_bin_stream_fresh ms", t_fresh, "identical coverage:", bool(torch.equal(cov, want)))
    import ctypes
    n_spill = torch.zeros(1, dtype=torch.int64)
