"""N > 1 host logic on CPU: world_size-2 gloo job.  Each rank scores its genomic block of the index from its
slice of the reads (with the oracle standing in for the GPU -- this test is about the shard plan, the read
slices, ordering and the control-plane gather, not about kernels) and rank 0 must end up with exactly the
single-process result."""
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
# each rank: its genomic block of the index (shard_plan), scored from ITS SLICE of the coordinate-sorted reads only
plan = multi_gpu.shard_plan(idx.exon_ptr, idx.exon_start, idx.exon_end, idx.orf_contig, size)
sh = plan[rank]
slices = multi_gpu.read_slices(reads["ref_id"], reads["first"], reads["last"], sh.spans, max(synth.TRUE_OFFSETS.values()))
mine = {k: np.concatenate([v[a:b] for a, b in slices]) for k, v in reads.items()}
assert len(mine["ref_id"]) < len(reads["ref_id"])
cov_mine, _, _ = CO.bin_reads(mine, 0, CO.make_len_table(synth.TRUE_OFFSETS), base, idx.contig_len, pad, plane)
sub = multi_gpu.sub_index(idx.as_dict(), sh.rows)
local = CO.score(sub, cov_mine, base, idx.contig_len, pad, plane, [0.428571428571, 5, 0, 0, 0.0], diagnostics=False)
local["rows"] = sh.rows
# TSV parts: one per run of consecutive rows, joined in row order
prefix = sys.argv[2]
if rank == 0:
    open(f"{prefix}_translating_ORFs.tsv.header", "w").write("header\n")
for g_lo, g_hi, r_lo in sh.runs:
    with open(f"{prefix}_translating_ORFs.tsv.rows{g_lo:012d}", "w") as fh:
        for k in range(g_hi - g_lo):
            fh.write(f"{g_lo + k}\t{local['count'][r_lo + k]}\n")
full = multi_gpu.gather_columns(local, dist)
dist.barrier()
if rank == 0:
    ref = CO.score(idx.as_dict(), cov, base, idx.contig_len, pad, plane, [0.428571428571, 5, 0, 0, 0.0],
                   diagnostics=False)
    order = np.argsort(full["rows"], kind="stable")
    assert np.array_equal(full["rows"][order], np.arange(idx.n_orf))
    for k in ref:
        assert np.array_equal(full[k][order], ref[k], equal_nan=True), k
    multi_gpu.join_runs(prefix, plan)
    rows = open(f"{prefix}_translating_ORFs.tsv").read().split("\n")
    assert rows[0] == "header" and [int(r.split("\t")[0]) for r in rows[1:] if r] == list(range(idx.n_orf))
    assert [int(r.split("\t")[1]) for r in rows[1:] if r] == ref["count"].tolist()
    print("GLOO_OK", [len(p.rows) for p in plan], [len(p.runs) for p in plan])
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


def test_shard_plan_and_read_slices():
    """Blocks along the genome: a partition of the rows, balanced in bytes, few runs each for a genome-ordered
    index; the read slices of a block hold every read that can reach it (checked by brute force), also for spliced
    reads whose two ends lie far apart; an index in random order falls back to contiguous row ranges."""
    from ribotricer_b200 import multi_gpu, synth

    cfg = synth.config("tiny")
    idx = synth.make_index(cfg)
    reads = synth.reads_to_numpy(synth.make_reads(cfg, idx, n_reads=60_000))
    rng = np.random.default_rng(5)
    spliced = rng.choice(len(reads["first"]), 500, replace=False)       # long introns: last far to the right of first
    reads["last"][spliced] += rng.integers(1000, 60_000, 500).astype(np.int32)
    cost = 4 * idx.orf_len + 8 * np.diff(idx.exon_ptr) + 42
    for n in (1, 2, 3, 8):
        plan = multi_gpu.shard_plan(idx.exon_ptr, idx.exon_start, idx.exon_end, idx.orf_contig, n)
        assert np.array_equal(np.sort(np.concatenate([s.rows for s in plan])), np.arange(idx.n_orf))
        per = np.array([cost[s.rows].sum() for s in plan])
        assert per.max() <= cost.sum() / n + cost.max()
        assert sum(len(s.runs) for s in plan) <= 16 * n
        for s in plan:
            assert sum(hi - lo for lo, hi, _ in s.runs) == len(s.rows)
            assert all(np.array_equal(s.rows[r:r + hi - lo], np.arange(lo, hi)) for lo, hi, r in s.runs)
            sub = multi_gpu.sub_index(idx.as_dict(), s.rows)
            assert np.array_equal(np.diff(sub["exon_ptr"]), np.diff(idx.exon_ptr)[s.rows])
            # brute force: reads that can put a P-site (5' end +- 13) on an exon position of this block
            off = 13
            slices = multi_gpu.read_slices(reads["ref_id"], reads["first"], reads["last"], s.spans, off)
            taken = np.zeros(len(reads["first"]), bool)
            for a, b in slices:
                taken[a:b] = True
            for c, lo, hi in s.spans:
                m = reads["ref_id"] == c
                p_first, p_last = reads["first"].astype(np.int64) + 1, reads["last"].astype(np.int64) + 1   # 1-based 5' ends
                near = m & (((p_first >= lo - off) & (p_first <= hi + off)) | ((p_last >= lo - off) & (p_last <= hi + off)))
                assert taken[near].all()
        if n == 8:
            assert sum(t.sum() for t in [taken]) < len(taken)       # a block does not pull the whole library
    perm = np.random.default_rng(3).permutation(idx.n_orf)          # an index in random order: contiguous row ranges
    shuffled = multi_gpu.sub_index(idx.as_dict(), perm)
    plan = multi_gpu.shard_plan(shuffled["exon_ptr"], shuffled["exon_start"], shuffled["exon_end"], shuffled["orf_contig"], 4,
                                max_runs_per_shard=8)
    assert all(len(s.runs) <= 1 for s in plan) and sum(len(s.rows) for s in plan) == idx.n_orf


SHARDED_WORKER = r'''
import os, sys
ROOT, out_dir = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch.distributed as dist
from helpers import alignments_to_reads, load_golden
from oracle import oracle_py as O
from oracle_engine import OracleEngine
from ribotricer_b200 import detect_orfs as D, multi_gpu, statistics as S
from ribotricer_b200.bam import ReadColumns

dist.init_process_group("gloo")
rank, size = dist.get_rank(), dist.get_world_size()
eng = OracleEngine(int(os.environ["LOCAL_RANK"]))           # the device calls of the product answered by the oracle
eng._aux = OracleEngine(eng.device_index)
D._ENGINE = eng
S.phasescore = lambda values, engine=None: O.phasescore_scipy(list(values))
case = load_golden("metagene_case.json.gz")["case"]
names = [c[0] for c in case["contigs"]]
cols = alignments_to_reads(case, names)
order = np.lexsort((cols["first"], cols["ref_id"]))         # a coordinate-sorted library: ranks take slices of it
reads = ReadColumns(names, np.array([c[1] for c in case["contigs"]], np.int64), {k: v[order] for k, v in cols.items()}, True)
idx_path = os.path.join(out_dir, f"index_{rank}.tsv")
open(idx_path, "w").write("\n".join(case["index"]) + "\n")
for tag, args in (("default", (None, None, None)), ("given", ("forward", None, {int(k): v for k, v in case["psite_offsets"].items()}))):
    prefix = os.path.join(out_dir, tag, "job")
    shard = multi_gpu.detect_orfs_sharded(reads, idx_path, prefix, args[0], args[1], args[2], 0.428571428571, 5, 0, 0, 0.0, True,
                                          meta_min_reads=case["meta_min_reads"])
    assert 0 < len(shard.rows) < len(case["index"]) - 1      # a real split
    dist.barrier()
    if rank == 0:
        single = os.path.join(out_dir, tag, "single")
        D.detect_orfs(reads, idx_path, single, args[0], args[1], args[2], 0.428571428571, 5, 0, 0, 0.0, True,
                      meta_min_reads=case["meta_min_reads"])
        for suffix in ("_translating_ORFs.tsv", "_pos.wig", "_neg.wig", "_bam_summary.txt"):
            a, b = open(prefix + suffix, "rb").read(), open(single + suffix, "rb").read()
            assert a == b and len(a) > 0, (tag, suffix)
        left = [f for f in os.listdir(os.path.dirname(prefix)) if ".rows" in f or f.endswith(".header")]
        assert not left, left
    dist.barrier()
if rank == 0:
    print("SHARDED_OK")
dist.destroy_process_group()
'''


def test_detect_orfs_sharded_two_rank_gloo_job(tmp_path, built):
    """The product's multi-rank entry point, multi_gpu.detect_orfs_sharded, as a world-size-2 gloo job on the CPU (the
    device calls answered by the oracle, tests/oracle_engine.py): rank 0 infers protocol and offsets and broadcasts them,
    every rank scores its genomic block from its slice of the sorted library and writes its row runs, rank 0 joins them.
    TSV, WIG and summary must equal the single-process detect_orfs() files byte for byte, with the default flags and
    with given offsets.  (GPU twin: test_gpu_api.py::test_two_rank_job_is_byte_identical, which needs two GPUs.)"""
    script = tmp_path / "sharded_worker.py"
    script.write_text(SHARDED_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29641", str(script), ROOT, str(tmp_path)],
        capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "SHARDED_OK" in out.stdout
