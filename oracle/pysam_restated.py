"""Pure-Python restatement of the slice of pysam that ribotricer's ``split_bam`` calls.

TEST INFRASTRUCTURE ONLY (see oracle/ref_import.py): pysam (``pysam>=0.19.1`` in the reference's
requirements.txt:7 / pyproject.toml:46) is a third-party dependency that is absent from this
image and not installable offline.  ``split_bam`` (bam.py:33-153) and ``is_read_uniq_mapping``
(common.py:33-69) use exactly this surface:

    pysam.AlignmentFile(path, "rb")   .count(until_eof=True)  .fetch(until_eof=True)  .close()
    AlignedSegment: is_qcfail, is_duplicate, is_secondary, is_unmapped, is_reverse, flag,
                    mapping_quality, reference_name, get_tags(), get_reference_positions()
and ``infer_protocol`` (infer_protocol.py:34-124) additionally ``reference_start`` / ``reference_end``.

Every one of them is defined by the SAM/BAM specification (SAMv1 section 4.2) plus pysam's
documented behaviour: ``get_reference_positions()`` lists the reference positions of the aligned
bases, i.e. of CIGAR operations M, = and X (insertions and soft clips have no reference position
and are left out because ``full_length`` defaults to False; D and N advance the reference without
yielding positions).  With this module installed as ``pysam`` the UNMODIFIED reference
``split_bam`` runs here on real BAM bytes, which is how tests/golden/split_bam_case.json.gz was made.
``reference_end`` follows pysam's property (libcalignedsegment.pyx): ``None`` when the read carries the unmapped
flag or has no CIGAR, otherwise htslib's ``bam_endpos``: ``pos`` + the reference length of the CIGAR (operations
M, D, N, = and X), counting at least 1.
It shares no code with the product's decoder (ribotricer_b200/csrc/rt_bam.cpp).
"""
from __future__ import annotations

import gzip
import struct

_CIGAR_CONSUMES_REF_WITH_POS = {0, 7, 8}     # M, =, X
_CIGAR_SKIPS_REF = {2, 3}                    # D, N


class AlignedSegment:
    def __init__(self, header, ref_id, pos, mapq, flag, cigar, tags):
        self._header = header
        self.reference_id = ref_id
        self.reference_start = pos
        self.mapping_quality = mapq
        self.flag = flag
        self.cigartuples = cigar
        self._tags = tags

    is_qcfail = property(lambda self: bool(self.flag & 0x200))
    is_duplicate = property(lambda self: bool(self.flag & 0x400))
    is_secondary = property(lambda self: bool(self.flag & 0x100))
    is_unmapped = property(lambda self: bool(self.flag & 0x4))
    is_reverse = property(lambda self: bool(self.flag & 0x10))

    @property
    def reference_end(self):
        if self.flag & 0x4 or not self.cigartuples:
            return None
        rlen = sum(n for op, n in self.cigartuples if op in (0, 2, 3, 7, 8))
        return self.reference_start + (rlen or 1)

    @property
    def reference_name(self):
        return self._header[self.reference_id][0] if 0 <= self.reference_id < len(self._header) else None

    def get_tags(self):
        return list(self._tags)

    def get_reference_positions(self, full_length=False):
        assert not full_length
        out, pos = [], self.reference_start
        for op, n in self.cigartuples:
            if op in _CIGAR_CONSUMES_REF_WITH_POS:
                out.extend(range(pos, pos + n))
                pos += n
            elif op in _CIGAR_SKIPS_REF:
                pos += n
        return out


def _parse_tags(buf: bytes):
    tags, i = [], 0
    scalar = {"A": ("c", 1), "c": ("b", 1), "C": ("B", 1), "s": ("h", 2), "S": ("H", 2), "i": ("i", 4), "I": ("I", 4),
              "f": ("f", 4)}
    while i < len(buf):
        tag, typ = buf[i:i + 2].decode(), chr(buf[i + 2])
        i += 3
        if typ in scalar:
            fmt, size = scalar[typ]
            val = struct.unpack_from("<" + fmt, buf, i)[0]
            if typ == "A":
                val = val.decode()
            i += size
        elif typ in "ZH":
            end = buf.index(b"\0", i)
            val = buf[i:end].decode()
            i = end + 1
        elif typ == "B":
            sub = chr(buf[i])
            n = struct.unpack_from("<I", buf, i + 1)[0]
            fmt, size = scalar[sub]
            val = list(struct.unpack_from("<%d%s" % (n, fmt), buf, i + 5))
            i += 5 + n * size
        else:
            raise ValueError(f"unknown aux type {typ!r}")
        tags.append((tag, val))
    return tags


class AlignmentFile:
    def __init__(self, path, mode="rb"):
        assert mode == "rb"
        with gzip.open(path, "rb") as fh:       # BGZF = concatenated gzip members
            data = fh.read()
        assert data[:4] == b"BAM\1", "not a BAM file"
        l_text = struct.unpack_from("<I", data, 4)[0]
        at = 8 + l_text
        n_ref = struct.unpack_from("<I", data, at)[0]
        at += 4
        self.references_and_lengths = []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<I", data, at)[0]
            name = data[at + 4:at + 4 + l_name - 1].decode()
            l_ref = struct.unpack_from("<I", data, at + 4 + l_name)[0]
            self.references_and_lengths.append((name, l_ref))
            at += 8 + l_name
        self._data, self._first = data, at

    references = property(lambda self: tuple(n for n, _ in self.references_and_lengths))
    lengths = property(lambda self: tuple(l for _, l in self.references_and_lengths))

    def fetch(self, until_eof=False):
        assert until_eof
        data, at = self._data, self._first
        while at < len(data):
            block_size = struct.unpack_from("<I", data, at)[0]
            ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, at + 4)
            p = at + 36 + l_name
            cigar = [(v & 0xF, v >> 4) for v in struct.unpack_from("<%dI" % n_cig, data, p)]
            p += 4 * n_cig + (l_seq + 1) // 2 + l_seq
            tags = _parse_tags(data[p:at + 4 + block_size])
            yield AlignedSegment(self.references_and_lengths, ref_id, pos, mapq, flag, cigar, tags)
            at += 4 + block_size

    def count(self, until_eof=False):
        return sum(1 for _ in self.fetch(until_eof=until_eof))

    def close(self):
        pass
