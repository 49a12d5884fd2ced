import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine, ScoreParams
cfg = synth.config("C2"); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
eng.set_index(**idx.as_dict()); eng.set_layout("compact")
d = synth.make_reads(cfg, idx, device="cuda")
h = {k: v.cpu().pin_memory() for k, v in d.items()}
packed = eng.pack_reads(h, pinned=True)
del d; torch.cuda.empty_cache()
cov = eng.new_coverage(); hout = eng.new_host_score_columns(idx.n_orf)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("packed_host ms", t(lambda: (eng.clear_coverage(cov), eng.bin_reads_packed_host(cov, packed, "forward"))))
print("columns sorted_hint ms", t(lambda: (eng.clear_coverage(cov), eng.bin_reads_host(cov, h, "forward", sorted_hint=True))))
print("columns plain ms", t(lambda: (eng.clear_coverage(cov), eng.bin_reads_host(cov, h, "forward", sorted_hint=False))))
print("score_host ms", t(lambda: eng.score_host(cov, 0, idx.n_orf, ScoreParams(), out=hout)))
