"""Minimal BAM/BGZF writer for tests (SAM spec v1 section 4): lets the native decoder be
exercised in an image that has neither pysam nor samtools."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def bgzf_block(payload: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    data = comp.compress(payload) + comp.flush()
    bsize = 12 + 6 + len(data) + 8 - 1
    return (b"\x1f\x8b\x08\x04" + struct.pack("<IBBH", 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, bsize)
            + data + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def aux_field(tag: str, typ: str, value) -> bytes:
    b = tag.encode() + typ[0].encode()
    if typ in "cCsSiI":
        return b + struct.pack("<" + {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ], value)
    if typ == "A":
        return b + value.encode()
    if typ == "f":
        return b + struct.pack("<f", value)
    if typ == "Z":
        return b + value.encode() + b"\0"
    if typ[0] == "B":      # typ = "B" + subtype
        sub = typ[1]
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
        return tag.encode() + b"B" + sub.encode() + struct.pack("<I", len(value)) + struct.pack("<%d%s" % (len(value), fmt), *value)
    raise ValueError(typ)


def record(ref_id, pos, mapq, flag, cigar, name=b"r", l_seq=None, aux=b"") -> bytes:
    """cigar: list of (op_char, length)."""
    if l_seq is None:
        l_seq = sum(l for op, l in cigar if op in "MIS=X")
    name = name + b"\0"
    cig = b"".join(struct.pack("<I", (l << 4) | CIGAR_OPS.index(op)) for op, l in cigar)
    seq = bytes((l_seq + 1) // 2)
    qual = b"\xff" * l_seq
    body = (struct.pack("<iiBBHHHiiii", ref_id, pos, len(name), mapq, 4680, len(cigar), flag, l_seq, -1, -1, 0)
            + name + cig + seq + qual + aux)
    return struct.pack("<I", len(body)) + body


def write_bam(path, refs, records, sorted_header=True, block_payload=0xFF00, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, eof=True):
    """refs: [(name, length)]; records: list of bytes from record()."""
    text = ("@HD\tVN:1.6\tSO:%s\n" % ("coordinate" if sorted_header else "unsorted")).encode()
    text += b"".join(("@SQ\tSN:%s\tLN:%d\n" % (n, l)).encode() for n, l in refs)
    head = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for n, l in refs:
        nb = n.encode() + b"\0"
        head += struct.pack("<I", len(nb)) + nb + struct.pack("<I", l)
    stream = head + b"".join(records)
    with open(path, "wb") as fh:
        for i in range(0, len(stream), block_payload):     # records straddle block boundaries on purpose
            fh.write(bgzf_block(stream[i:i + block_payload], level, strategy))
        if eof:
            fh.write(BGZF_EOF)
