import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine
cfg = synth.config("C2"); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
eng.set_index(**idx.as_dict()); eng.set_layout("compact")
d = synth.make_reads(cfg, idx, device="cuda")
h = {k: v.cpu().numpy() for k, v in d.items()}
del d; torch.cuda.empty_cache()
ds = eng.upload_stream(eng.stream_reads(h))
cov = eng.new_coverage(); st, lc = eng.new_bin_accumulators()
for _ in range(3): eng.bin_stream_device(cov, ds, "forward", st, lc)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.bin_stream_device(cov, ds, "forward", st, lc)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("RT_LIB_PATH", "default").split("/")[-1], "K1 ms", e0.elapsed_time(e1) / 10, "nonzero slots", int((cov != 0).sum()), "sum", int(cov.sum()) // 13)
