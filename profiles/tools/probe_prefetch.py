import os, sys, time, subprocess
# each setting in its own process: the prefetch distance is read once per process
for pf in ("0", "512", "2048", "8192"):
    env = dict(os.environ, RT_PACK_PREFETCH=pf)
    out = subprocess.run([sys.executable, "-c", """
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine
cfg = synth.config("C2", 0.5); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len); eng.set_length_table(synth.TRUE_OFFSETS, None)
eng.set_index(**idx.as_dict()); eng.set_layout("compact")
d = synth.make_reads(cfg, idx, device="cuda")
h = {k: v.cpu().pin_memory() for k, v in d.items()}
del d; torch.cuda.empty_cache()
cov = eng.new_coverage()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for pipes in (8, 16):
    os.environ["RT_PACK_PIPES"] = str(pipes)
    print("prefetch", os.environ["RT_PACK_PREFETCH"], "pipes", pipes, "bin_reads_host ms (50 M reads)", round(t(lambda: (eng.clear_coverage(cov), eng.bin_reads_host(cov, h, "forward", sorted_hint=True))), 2))
"""], env=env, capture_output=True, text=True)
    print(out.stdout.strip() or out.stderr[-500:])
