"""N > 1 host logic on CPU: world_size-2 gloo job.  Each rank scores its ORF shard (with the
oracle standing in for the GPU -- this test is about sharding, ordering and the control-plane
gather, not about kernels) and rank 0 must end up with exactly the single-process result."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from oracle import c_oracle as CO
from ribotricer_b200 import multi_gpu, synth

dist.init_process_group("gloo")
rank, size = dist.get_rank(), dist.get_world_size()
cfg = synth.config("tiny")
idx = synth.make_index(cfg)
reads = synth.reads_to_numpy(synth.make_reads(cfg, idx, n_reads=40_000))
pad = 64
base, plane = CO.genome_layout(idx.contig_len, pad)
cov, _, _ = CO.bin_reads(reads, 0, CO.make_len_table(synth.TRUE_OFFSETS), base, idx.contig_len, pad, plane)
bounds = multi_gpu.shard_bounds(idx.orf_len, np.diff(idx.exon_ptr), size)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
local = CO.score(idx.as_dict(), cov, base, idx.contig_len, pad, plane, [0.428571428571, 5, 0, 0, 0.0], lo, hi,
                 diagnostics=False)
# TSV parts in rank order
prefix = sys.argv[2]
with open(f"{prefix}_translating_ORFs.tsv.part{rank}", "w") as fh:
    if rank == 0:
        fh.write("header\n")
    for k in range(hi - lo):
        fh.write(f"{lo + k}\t{local['count'][k]}\n")
full = multi_gpu.gather_columns(local, dist)
dist.barrier()
if rank == 0:
    ref = CO.score(idx.as_dict(), cov, base, idx.contig_len, pad, plane, [0.428571428571, 5, 0, 0, 0.0],
                   diagnostics=False)
    for k in ref:
        assert np.array_equal(full[k], ref[k], equal_nan=True), k
    multi_gpu.join_parts(prefix, size)
    rows = open(f"{prefix}_translating_ORFs.tsv").read().split("\n")
    assert rows[0] == "header" and [int(r.split("\t")[0]) for r in rows[1:] if r] == list(range(idx.n_orf))
    print("GLOO_OK", bounds.tolist())
dist.destroy_process_group()
'''


def test_shard_bounds_properties():
    from ribotricer_b200 import multi_gpu

    rng = np.random.default_rng(0)
    L = (3 * np.maximum(20, rng.lognormal(np.log(110), 1.0, 50_000))).astype(np.int64)
    E = rng.geometric(0.25, len(L))
    for n in (1, 2, 3, 4, 8):
        b = multi_gpu.shard_bounds(L, E, n)
        assert b[0] == 0 and b[-1] == len(L) and (np.diff(b) >= 0).all()
        cost = 4 * L + 8 * E + 42
        per = np.array([cost[b[i]:b[i + 1]].sum() for i in range(n)])
        assert per.max() <= cost.sum() / n + cost.max()          # balanced within one ORF
    b = multi_gpu.shard_bounds([60, 60], [1, 1], 8)              # more shards than ORFs
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b) >= 0).all()
    assert multi_gpu.shard_bounds([], [], 4).tolist() == [0, 0, 0, 0, 0]


def test_two_rank_gloo_job(tmp_path, built):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29631", str(script), ROOT, str(tmp_path / "job")],
        capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GLOO_OK" in out.stdout
