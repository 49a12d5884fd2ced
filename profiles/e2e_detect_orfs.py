#!/usr/bin/env python
"""Wall-clock of the whole detect_orfs() call (host stages included) on BASELINE configs[0]
(yeast R64 scale: 100 k candidate ORFs, 10 M reads), with the stage split.

    python profiles/e2e_detect_orfs.py [scale] [config]
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from ribotricer_b200 import detect_orfs as D  # noqa: E402
from ribotricer_b200 import synth  # noqa: E402
from ribotricer_b200.bam import ReadColumns, load_reads, save_read_columns, split_bam  # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    name = sys.argv[2] if len(sys.argv) > 2 else "C1"
    cfg = synth.config(name, scale)
    idx = synth.make_index(cfg)
    tmp = tempfile.mkdtemp(prefix="rt_e2e_")
    index_path, reads_path = os.path.join(tmp, "index.tsv"), os.path.join(tmp, "reads.npz")
    idx.write_tsv(index_path)
    eng = D.get_engine(0)
    cols = synth.reads_to_numpy(synth.make_reads(cfg, idx, device=eng.device))
    save_read_columns(reads_path, ReadColumns(idx.contig_names, idx.contig_len, cols, True))
    out = {"config": f"{name} x{scale}: {idx.n_orf} ORFs ({int(idx.orf_len.sum())} nt), {len(cols['ref_id'])} reads"}
    offsets = {k: v for k, v in synth.TRUE_OFFSETS.items()}
    lengths = sorted(offsets)
    runs = (("offsets given, translating rows only", dict(read_lengths=lengths, psite_offsets=offsets, report_all=False)),
            ("offsets given, --report_all", dict(read_lengths=lengths, psite_offsets=offsets, report_all=True)),
            ("default flags (protocol, lengths and offsets inferred)", dict(read_lengths=None, psite_offsets=None, report_all=False)))
    for tag, kw in runs[:int(os.environ.get("E2E_RUNS", "3"))]:        # E2E_RUNS=1: a short run on a large configuration
        prefix = os.path.join(tmp, "run", "lib")
        t0 = time.perf_counter()
        try:
            D.detect_orfs(reads_path, index_path, prefix, "forward" if kw["psite_offsets"] else None, kw["read_lengths"],
                          kw["psite_offsets"], 0.428571428571, 5, 0, 0, 0.0, kw["report_all"])
        except SystemExit as exc:   # e.g. "no periodic read length found" (metagene.py:304-307)
            out[tag] = {"seconds": round(time.perf_counter() - t0, 3), "sys.exit": str(exc)}
            continue
        dt = time.perf_counter() - t0
        rows = sum(1 for _ in open(prefix + "_translating_ORFs.tsv")) - 1
        out[tag] = {"seconds": round(dt, 3), "rows": rows, "tsv_MB": round(os.path.getsize(prefix + "_translating_ORFs.tsv") / 1e6, 1)}
    # stage split of the second run
    t = {}
    t0 = time.perf_counter(); reads = load_reads(reads_path); t["load read columns (.npz)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); D._INDEX_CACHE.clear(); ix = D.load_index(index_path); t["native index load"] = time.perf_counter() - t0
    t0 = time.perf_counter(); al, _ = split_bam(reads, "forward", os.path.join(tmp, "s"), lengths, engine=eng); t["split_bam (H2D + K1 count pass)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); merged = D.merge_read_lengths(al, offsets); eng.torch.cuda.synchronize(); t["merge_read_lengths (K1)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); D.export_wig(merged, os.path.join(tmp, "w")); t["export_wig"] = time.perf_counter() - t0
    t0 = time.perf_counter(); D.export_orf_coverages(index_path, merged, os.path.join(tmp, "x"), report_all=True); t["export_orf_coverages (set_index + score + K4 + native TSV)"] = time.perf_counter() - t0
    t0 = time.perf_counter(); D.export_orf_coverages(index_path, merged, os.path.join(tmp, "y"), report_all=True); t["export_orf_coverages again (index resident)"] = time.perf_counter() - t0
    out["stages_s"] = {k: round(v, 3) for k, v in t.items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
