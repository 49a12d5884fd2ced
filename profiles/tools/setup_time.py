import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine
cfg = synth.config("C2"); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
t0 = time.perf_counter(); eng.set_index(**idx.as_dict()); t1 = time.perf_counter()
eng.set_layout("compact"); t2 = time.perf_counter()
cov = eng.new_coverage(); out = eng.new_score_columns(idx.n_orf)
eng.score_device(cov, out); torch.cuda.synchronize(); t3 = time.perf_counter()
hout = eng.new_host_score_columns(idx.n_orf); eng.score_host(cov, out=hout); t4 = time.perf_counter()
print(f"set_index {t1-t0:.2f} s, set_layout {t2-t1:.2f} s, first score (plan) {t3-t2:.2f} s, first score_host (4 part plans) {t4-t3:.2f} s")
