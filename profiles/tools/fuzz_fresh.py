"""Differential soak of rt_bin_stream_fresh against clear + rt_bin_stream: random length tables (negative offsets, unused and
filtered lengths), protocols, read subsets (dense piles, sparse libraries), with and without spliced / long reads."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from ribotricer_b200 import synth
from ribotricer_b200.engine import Engine
from test_gpu_parity import _stream_library
cfg = synth.config("tiny"); idx = synth.make_index(cfg)
eng = Engine(0); eng.set_genome(idx.contig_names, idx.contig_len, 64)
eng.set_length_table(synth.TRUE_OFFSETS, None); eng.set_index(**idx.as_dict()); eng.set_layout("compact")
rng = np.random.default_rng(11)
bad = 0
for trial in range(60):
    n = int(rng.choice([300, 5_000, 80_000, 400_000]))
    if trial % 2:
        reads = _stream_library(idx, n, trial)
    else:
        reads = synth.reads_to_numpy(synth.make_reads(cfg, idx, n_reads=n, seed_offset=trial))
        if trial % 4 == 0:      # piles: many reads on few positions
            keep = np.sort(rng.choice(len(reads["first"]), size=max(50, n // 50), replace=False))
            reads = {k: np.repeat(v[keep], 50)[:n] for k, v in reads.items()}
    offs = {int(l): int(rng.integers(-40, 60)) for l in rng.choice(np.arange(20, 40), size=int(rng.integers(1, 10)), replace=False)}
    rl = None if trial % 3 else [int(x) for x in rng.choice(np.arange(20, 40), size=8, replace=False)]
    eng.set_length_table(offs, rl)
    stream = eng.upload_stream(eng.stream_reads(reads))
    for protocol in ("forward", "reverse"):
        want = eng.new_coverage(); st, lc = eng.new_bin_accumulators()
        eng.bin_stream_device(want, stream, protocol, st, lc)
        got = torch.full_like(want, 99); st2, lc2 = eng.new_bin_accumulators()
        eng.bin_stream_device(got, stream, protocol, st2, lc2, fresh=True)
        ok = bool(torch.equal(got, want) and torch.equal(st, st2) and torch.equal(lc, lc2))
        bad += not ok
        if not ok: print("MISMATCH trial", trial, protocol, n, offs, rl, int((got != want).sum()))
print("trials", 60 * 2, "mismatches", bad, "binned in last", int(want.sum()))
