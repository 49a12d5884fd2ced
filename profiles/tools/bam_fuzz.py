"""Fuzz of rt_bam_load: damaged copies of a small BAM (bytes of the uncompressed stream overwritten under valid CRCs, bits of
the compressed file flipped, either one cut short) at batch sizes from 1 B to 1 MiB; every file must load or be refused.

    python profiles/tools/bam_fuzz.py [seed] [seconds]
    # under AddressSanitizer + UBSan: build the three host sources with -fsanitize=address,undefined into a .so, then
    # ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) RT_FUZZ_LIB=that.so python ...
"""
import ctypes as C, os, sys, time, struct, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bam_writer as W
from test_host import _bam_stream_bytes
lib = C.CDLL(os.environ.get('RT_FUZZ_LIB', os.path.join(ROOT, 'ribotricer_b200', 'libribotricer_b200.so')))
lib.rt_bam_last_error.restype = C.c_char_p
lib.rt_bam_n_reads.restype = C.c_int64
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
refs = [("c%d" % i, 100000) for i in range(5)]
recs = []
for i in range(600):
    cigar = [("M", int(rng.integers(20, 40)))]
    if rng.random() < 0.3: cigar += [("N", 100), ("M", 10)]
    aux = W.aux_field("NH", str(rng.choice(list("cCsSiI"))), 1) + (W.aux_field("MD", "Z", "30") if rng.random() < 0.5 else b"") + (W.aux_field("ZB", "Bs", [1, 2, 3]) if rng.random() < 0.3 else b"")
    recs.append(W.record(int(rng.integers(0, 5)), int(rng.integers(0, 90000)), 255, int(rng.choice([0, 16])), cigar, name=b"q%d" % i, aux=aux))
stream = _bam_stream_bytes(W, refs, recs)
def bgzf(data, payload):
    return b"".join(W.bgzf_block(data[i:i + payload]) for i in range(0, len(data), payload)) + W.BGZF_EOF
good = bgzf(stream, 3000)
path = "/tmp/fuzz.bam"
n_ok = n_err = 0
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 30
while time.time() < t_end:
    mode = int(rng.integers(0, 4))
    if mode == 0:      # corrupt the uncompressed stream (valid CRCs): exercises the header / record / aux parsers
        b = bytearray(stream)
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.integers(0, len(b)))
            b[k] = int(rng.integers(0, 256))
        data = bgzf(bytes(b), int(rng.choice([500, 3000, 65280])))
    elif mode == 1:    # corrupt the compressed file
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        data = bytes(b)
    elif mode == 2:    # truncate
        data = good[:int(rng.integers(0, len(good)))]
    else:              # truncate the uncompressed stream
        data = bgzf(stream[:int(rng.integers(0, len(stream)))], 3000)
    open(path, "wb").write(data)
    os.environ["RT_BAM_BATCH_BYTES"] = str(int(rng.choice([1, 100, 5000, 1 << 20])))
    h = C.c_void_p()
    rc = lib.rt_bam_load(path.encode(), int(rng.choice([1, 3])), C.byref(h))
    if rc == 0:
        n_ok += 1; lib.rt_bam_free(h)
    else:
        n_err += 1
print("loaded", n_ok, "refused", n_err)
