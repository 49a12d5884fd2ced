"""Run under torchrun on N GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank scores its genomic block of one synthetic library with detect_orfs_sharded() from its slice of the
reads; rank 0 then
runs the single-GPU detect_orfs() and requires byte-identical TSV / WIG / summary files, with and
without P-site offset inference (the offsets are inferred on rank 0 and broadcast).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

from ribotricer_b200 import multi_gpu, synth  # noqa: E402
from ribotricer_b200.bam import ReadColumns, save_read_columns  # noqa: E402
from ribotricer_b200.detect_orfs import detect_orfs  # noqa: E402


def main():
    rank, size, local = multi_gpu.world()
    tmp = os.environ.get("MGPU_TMP") or os.path.join(tempfile.gettempdir(), "rt_mgpu_check")
    os.makedirs(tmp, exist_ok=True)
    cfg = synth.config("tiny")
    idx = synth.make_index(cfg)
    index_path = os.path.join(tmp, "index.tsv")
    reads_path = os.path.join(tmp, "reads.npz")
    if rank == 0:
        idx.write_tsv(index_path)
        cols = synth.reads_to_numpy(synth.make_reads(cfg, idx, n_reads=400_000))
        save_read_columns(reads_path, ReadColumns(idx.contig_names, idx.contig_len, cols, True))
    # rendezvous before anybody reads the files
    import torch
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
    for tag, lengths, offsets, meta in (("given", [28, 29, 30], {28: 12, 29: 12, 30: 13}, 10 ** 9),
                                        ("inferred", None, None, 20000)):
        prefix = os.path.join(tmp, f"sharded_{tag}")
        shard = multi_gpu.detect_orfs_sharded(reads_path, index_path, prefix, "forward", lengths, offsets,
                                              0.428571428571, 5, 0, 0, 0.0, True, meta_min_reads=meta)
        print(f"rank {rank}: {len(shard.rows)} of {idx.n_orf} ORFs in {len(shard.runs)} runs, spans {shard.spans} ({tag})",
              flush=True)
        dist.barrier()
        if rank == 0:
            single = os.path.join(tmp, f"single_{tag}")
            detect_orfs(reads_path, index_path, single, "forward", lengths, offsets, 0.428571428571, 5, 0, 0, 0.0,
                        True, meta_min_reads=meta)
            for suffix in ("_translating_ORFs.tsv", "_pos.wig", "_neg.wig", "_bam_summary.txt"):
                a, b = open(prefix + suffix).read(), open(single + suffix).read()
                assert a == b, f"{tag}: {suffix} differs between {size} ranks and 1 rank"
            rows = open(prefix + "_translating_ORFs.tsv").read().count("\n") - 1
            assert rows == idx.n_orf
            if tag == "inferred":
                print(open(prefix + "_psite_offsets.txt").read())
            print(f"MULTI_GPU_OK {tag}: {size} ranks, {rows} rows identical to the single-GPU run", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
