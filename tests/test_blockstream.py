"""Raw codec streams: Hadoop's BlockCompressorStream / BlockDecompressorStream framing that
Lz4Codec / ZstdCodec put around the per-block natives (Lz4Codec.java:95-104,128-138; ZstdCodec.java:103-112,136-146).
CPU: the framing code of the library (4mc_b200/csrc/blockstream.h) bound to the oracle's codecs -- writer rules,
reader rules, byte layouts.  GPU: the C-ABI calls, cross-decoded both ways against that CPU build."""
import ctypes as C
import os
import random

import pytest

from conftest import build_native, gen_logtext

MIB = 1 << 20
MAX_LZ4 = 4 * MIB - (4 * MIB // 255 + 16)          # bufferSize - compressionOverhead = 4 177 840
MAX_ZSTD = 4 * MIB - (4 * MIB >> 8)                # 4 177 920


@pytest.fixture(scope="module")
def bs(oracle):
    B = build_native("bs_emul", ["tests/native/bs_emul.cpp"], deps=["4mc_b200/csrc/blockstream.h"])
    B.bs_emul_max_input.restype = C.c_uint
    B.bs_emul_bound.restype = C.c_size_t
    B.bs_emul_bound.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
    B.bs_emul_compress.restype = C.c_longlong
    B.bs_emul_compress.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    B.bs_emul_decompress.restype = C.c_longlong
    B.bs_emul_decompress.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    B.bs_emul_plan.restype = C.c_longlong
    B.bs_emul_plan.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_uint), C.c_size_t, C.POINTER(C.c_int)]
    B.bs_emul_predict.restype = C.c_longlong
    B.bs_emul_predict.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint), C.c_size_t, C.POINTER(C.c_size_t)]
    lz4_c = C.cast(oracle.fmo_lz4_compress, C.c_void_p)
    lz4_d = C.cast(oracle.fmo_lz4_decompress_safe, C.c_void_p)
    zstd_d = C.cast(oracle.fmo_zstd_decompress, C.c_void_p)

    class Api:
        lib = B

        @staticmethod
        def compress(data, write_size=0, level=1, kind=0, cfn=None):
            cap = B.bs_emul_bound(kind, len(data), write_size)
            out = C.create_string_buffer(cap)
            r = B.bs_emul_compress(kind, cfn or lz4_c, level, bytes(data), len(data), write_size, out, cap)
            assert r > 0, r
            return out.raw[:r]

        @staticmethod
        def decompress(stream, cap, kind=0):
            out = C.create_string_buffer(max(cap, 1))
            r = B.bs_emul_decompress(kind, zstd_d if kind else lz4_d, bytes(stream), len(stream), out, cap)
            return r, out.raw[:max(r, 0)]

        @staticmethod
        def plan(n, write_size, kind=0):
            raws = (C.c_uint * 4096)()
            tz = C.c_int()
            k = B.bs_emul_plan(kind, n, write_size, raws, 4096, C.byref(tz))
            return list(raws[:k]), bool(tz.value)

        @staticmethod
        def predict(stream, kind=0):
            us = (C.c_uint * 4096)()
            total = C.c_size_t()
            k = B.bs_emul_predict(kind, bytes(stream), len(stream), us, 4096, C.byref(total))
            return (None, 0) if k < 0 else (list(us[:k]), total.value)
    return Api


def walk(stream, max_input):
    """[(rawLen, [cLen, ...]), ...] by the reader's rules, chunk sizes taken as the reference writer makes them."""
    out, ip = [], 0
    while len(stream) - ip >= 4:
        raw = int.from_bytes(stream[ip:ip + 4], "big"); ip += 4
        if raw == 0:
            out.append((0, []))
            break
        chunks, got = [], 0
        while got < raw:
            c = int.from_bytes(stream[ip:ip + 4], "big"); ip += 4 + c
            chunks.append(c)
            got += min(max_input, raw - got)
        out.append((raw, chunks))
    assert ip == len(stream)
    return out


# ---- CPU: writer / reader rules ------------------------------------------------------------------

def test_constants(bs):
    assert bs.lib.bs_emul_max_input(0) == MAX_LZ4 == 4177840
    assert bs.lib.bs_emul_max_input(1) == MAX_ZSTD == 4177920


def test_writer_block_plan(bs):
    assert bs.plan(0, 0) == ([], True)                                  # finish() on an untouched compressor: rawLen 0
    assert bs.plan(1, 0) == ([1], False)
    assert bs.plan(MAX_LZ4, 0) == ([MAX_LZ4], False)                    # fits the buffer: one block, one chunk
    assert bs.plan(MAX_LZ4 + 1, 0) == ([MAX_LZ4 + 1], True)             # large write: written at once, then rawLen 0 at close
    assert bs.plan(10 * MIB, 0) == ([10 * MIB], True)
    per = MAX_LZ4 // 65536 * 65536                                      # 64 KiB writes: a block is cut before the write that overflows
    assert bs.plan(10 * MIB, 65536) == ([per, per, 10 * MIB - 2 * per], False)
    assert bs.plan(3 * MAX_LZ4, MAX_LZ4) == ([MAX_LZ4] * 3, False)
    assert bs.plan(MAX_LZ4 + 5 * MIB + 7, 5 * MIB) == ([5 * MIB, MAX_LZ4 + 7], True)   # both writes large; the 2nd is the rest
    assert bs.plan(6 * MIB, MIB) == ([3 * MIB, 3 * MIB], False)         # 3 MiB buffered + 1 MiB > MAX -> cut


def test_byte_layouts(bs):
    assert bs.compress(b"") == bytes(4)
    assert bs.compress(b"A") == bytes.fromhex("00000001" "00000002" "1041")      # LZ4: token 0x10, one literal
    s = bs.compress(b"A" * 100, write_size=40)                           # small writes join one block
    assert s[:4] == (100).to_bytes(4, "big") and len(walk(s, MAX_LZ4)) == 1
    data = gen_logtext_cached(MAX_LZ4 + 1)
    s = bs.compress(data)
    w = walk(s, MAX_LZ4)
    assert [r for r, _ in w] == [MAX_LZ4 + 1, 0] and len(w[0][1]) == 2
    assert w[0][1][1] == 2                                               # the 1-byte tail: token + literal
    assert bs.decompress(s, len(data)) == (len(data), data)


_text = {}


def gen_logtext_cached(n):
    import importlib
    if "t" not in _text:
        _text["t"] = gen_logtext(importlib.import_module("4mc_b200"), 12 * MIB + 4096)
    return _text["t"][:n]


@pytest.mark.parametrize("n,ws", [(0, 0), (1, 0), (70000, 0), (70000, 1000), (MAX_LZ4, 0), (MAX_LZ4 + 1, 0),
                                  (9 * MIB + 321, 0), (9 * MIB + 321, 65536), (9 * MIB + 321, 5 * MIB), (12 * MIB, MAX_LZ4)])
def test_cpu_round_trip_and_prediction(bs, n, ws):
    data = gen_logtext_cached(n)
    s = bs.compress(data, ws)
    assert len(s) <= bs.lib.bs_emul_bound(0, n, ws)
    assert bs.decompress(s, n) == (n, data)
    raws, tz = bs.plan(n, ws)
    w = walk(s, MAX_LZ4)
    assert [r for r, _ in w] == raws + ([0] if tz else [])
    us, total = bs.predict(s)
    assert total == n and sum(us) == n and len(us) == sum(len(c) for _, c in w)


def test_bound_holds_for_any_write_size(bs):
    data = gen_logtext_cached(9 * MIB + 321)
    rnd = random.Random(11).randbytes(5 * MIB)                     # incompressible: every chunk grows
    for ws in (1, 7, 4096, MAX_LZ4 // 2, MAX_LZ4 // 2 + 1, MAX_LZ4 - 1, MAX_LZ4, MAX_LZ4 + 1, 6 * MIB):
        raws, tz = bs.plan(len(data), ws)
        assert sum(raws) == len(data) and len(raws) <= 2 * (len(data) // MAX_LZ4) + 3
        assert all(r >= min(MAX_LZ4 // 2, len(data)) for r in raws[:-1])
    for ws in (0, 1000, MAX_LZ4, 3 * MIB):
        assert len(bs.compress(rnd, ws)) <= bs.lib.bs_emul_bound(0, len(rnd), ws) < len(rnd) + len(rnd) // 100 + 4096


def test_reader_rules(bs):
    data = gen_logtext_cached(200000)
    s = bs.compress(data, 50000)
    assert bs.decompress(s + bytes(4), len(data)) == (len(data), data)          # trailing rawLen 0
    assert bs.decompress(s + bytes(4) + b"junk", len(data)) == (len(data), data)  # nothing is read past it
    assert bs.decompress(s + b"\x00\x00", len(data)) == (len(data), data)       # rawReadInt fails at a boundary: EOF
    assert bs.decompress(s[:-10], len(data))[0] == -2                           # cut inside a chunk: EOFException
    assert bs.decompress(s[:6], len(data))[0] == -2                             # cut inside a chunk length
    bad = bytearray(s); bad[8] = 0xff                                           # first token: huge literal run
    assert bs.decompress(bytes(bad), len(data))[0] == -4
    assert bs.decompress(s, len(data) - 1)[0] == -3                             # output too small
    big = (5 * MIB).to_bytes(4, "big")
    assert bs.decompress((1).to_bytes(4, "big") + big + bytes(10), 100)[0] == -4   # chunk beyond the 4 MiB buffer
    # a foreign writer: one block, chunks of another size -- the reader does not care
    a, b = bs.compress(data[:70000])[4:], bs.compress(data[70000:])[4:]
    foreign = (200000).to_bytes(4, "big") + a + b
    assert bs.decompress(foreign, 200000) == (200000, data)
    us, total = bs.predict(foreign)                                              # the framing alone cannot tell: sizes differ
    assert us == [200000] or us is None


def test_zstd_streams_from_the_reference_codec(bs, ref):
    """kind 1: chunks are zstd frames; compressed by the reference's ZSTD_compress, read by the oracle's decoder."""
    ref.ZSTD_compress.restype = C.c_size_t
    cfn = C.cast(ref.ZSTD_compress, C.c_void_p)
    for n, ws in ((0, 0), (70000, 0), (MAX_ZSTD + 1, 0), (9 * MIB, 65536)):
        data = gen_logtext_cached(n)
        for level in (1, 4):
            s = bs.compress(data, ws, level=level, kind=1, cfn=cfn)
            assert bs.decompress(s, n, kind=1) == (n, data)
    s = bs.compress(gen_logtext_cached(MAX_ZSTD + 1), kind=1, cfn=cfn)
    assert [r for r, _ in walk(s, MAX_ZSTD)] == [MAX_ZSTD + 1, 0]


def random_datum_records(count, seed):
    """The input of the reference's own codec test (TestFourMcCodec.codecTest, java/hadoop-4mc/src/test/java/com/fing/
    compression/fourmc/TestFourMcCodec.java): `count` key / value pairs of Hadoop RandomDatum records, each a BE32
    length of 10 + 10^(3 * U[0,1)) followed by that many random bytes (same shape; Java's PRNG is not reproduced)."""
    rng = random.Random(seed)
    parts = []
    for _ in range(2 * count):
        n = 10 + int(10.0 ** (rng.random() * 3.0))
        parts.append(n.to_bytes(4, "big") + rng.randbytes(n))
    return b"".join(parts)


def test_reference_codec_test_shape_cpu(bs, ref):
    """TestFourMcCodec.testZstdCodec: 100 000 random records through ZstdCodec, handed to the stream in ONE write
    (DataOutputStream over BufferedOutputStream passes a large array straight through) -> one block of
    ceil(n / MAX_INPUT_SIZE) chunks and a closing zero, read back record for record."""
    ref.ZSTD_compress.restype = C.c_size_t
    cfn = C.cast(ref.ZSTD_compress, C.c_void_p)
    data = random_datum_records(100000, 20240917)
    s = bs.compress(data, 0, kind=1, cfn=cfn)
    w = walk(s, MAX_ZSTD)
    assert [r for r, _ in w] == [len(data), 0] and len(w[0][1]) == -(-len(data) // MAX_ZSTD)
    assert len(data) / len(s) < 1.1                               # "Compression ratio should be very small, almost 1"
    assert bs.decompress(s, len(data), kind=1) == (len(data), data)


# ---- GPU: the C-ABI calls ------------------------------------------------------------------------

def _gpu_api(pkg):
    L = pkg.lib()
    L.fourmc_blockstream_bound.restype = C.c_size_t
    L.fourmc_blockstream_bound.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
    L.fourmc_blockstream_compress_host.restype = C.c_longlong
    L.fourmc_blockstream_compress_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t]
    L.fourmc_blockstream_decompress_host.restype = C.c_longlong
    L.fourmc_blockstream_decompress_host.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    return L


def gpu_compress(ctx, pkg, data, zstd=0, level=1, ws=0, serial=False):
    L = _gpu_api(pkg)
    cap = L.fourmc_blockstream_bound(zstd, len(data), ws)
    out = C.create_string_buffer(cap)
    if serial:
        os.environ["FOURMC_BS_SERIAL"] = "1"          # chunk by chunk through the per-block call, like the Java stream
    try:
        r = L.fourmc_blockstream_compress_host(ctx.handle, zstd, level, bytes(data), len(data), ws, out, cap)
    finally:
        os.environ.pop("FOURMC_BS_SERIAL", None)
    assert r > 0, (r, ctx.last_error())
    return out.raw[:r]


def gpu_decompress(ctx, pkg, stream, cap, zstd=0, serial=False):
    L = _gpu_api(pkg)
    out = C.create_string_buffer(max(cap, 1))
    if serial:
        os.environ["FOURMC_BS_SERIAL"] = "1"
    try:
        r = L.fourmc_blockstream_decompress_host(ctx.handle, zstd, bytes(stream), len(stream), out, cap)
    finally:
        os.environ.pop("FOURMC_BS_SERIAL", None)
    return r, out.raw[:max(r, 0)]


@pytest.mark.gpu
@pytest.mark.parametrize("zstd", [0, 1])
def test_gpu_streams_round_trip_and_cross_decode(ctx, pkg, bs, zstd):
    rnd = random.Random(5).randbytes(MAX_LZ4 + 100000)                 # incompressible: chunks larger than their input
    cases = [(b"", 0), (b"A", 0), (gen_logtext_cached(70000), 1000), (gen_logtext_cached(9 * MIB + 321), 0),
             (gen_logtext_cached(9 * MIB + 321), 65536), (gen_logtext_cached(12 * MIB), MAX_LZ4), (rnd, 0), (rnd + rnd, 65536), (gen_logtext_cached(5 * MIB) + rnd[:MIB] + bytes(3 * MIB), MIB)]
    mx = MAX_ZSTD if zstd else MAX_LZ4
    for data, ws in cases:
        for level in ((1, 3) if len(data) < 6 * MIB else (1,)):
            s = gpu_compress(ctx, pkg, data, zstd, level, ws)
            raws, tz = bs.plan(len(data), ws, kind=zstd)
            assert [r for r, _ in walk(s, mx)] == raws + ([0] if tz else [])
            assert bs.decompress(s, len(data), kind=zstd) == (len(data), data)          # CPU reader (oracle codec)
            assert gpu_decompress(ctx, pkg, s, len(data), zstd) == (len(data), data)      # batch
            assert gpu_decompress(ctx, pkg, s, len(data), zstd, serial=True) == (len(data), data)
            if level == 1:                                                               # the writer's batch path vs chunk by chunk
                s2 = gpu_compress(ctx, pkg, data, zstd, level, ws, serial=True)
                assert [r for r, _ in walk(s2, mx)] == raws + ([0] if tz else [])
                assert bs.decompress(s2, len(data), kind=zstd) == (len(data), data)
    assert gpu_compress(ctx, pkg, b"", zstd) == bytes(4)


@pytest.mark.gpu
def test_gpu_reference_codec_test_shape(ctx, pkg, bs):
    """TestFourMcCodec.testZstdCodec on the GPU path (and its LZ4 twin): see test_reference_codec_test_shape_cpu."""
    data = random_datum_records(100000, 20240917)
    for zstd, mx in ((1, MAX_ZSTD), (0, MAX_LZ4)):
        s = gpu_compress(ctx, pkg, data, zstd)
        w = walk(s, mx)
        assert [r for r, _ in w] == [len(data), 0] and len(w[0][1]) == -(-len(data) // mx)
        assert gpu_decompress(ctx, pkg, s, len(data), zstd) == (len(data), data)
        assert bs.decompress(s, len(data), kind=zstd) == (len(data), data)      # the oracle's decoders agree


@pytest.mark.gpu
def test_gpu_reads_cpu_made_streams_and_foreign_chunking(ctx, pkg, bs):
    data = gen_logtext_cached(9 * MIB + 321)
    for ws in (0, 65536, 5 * MIB):
        s = bs.compress(data, ws)                                       # oracle LZ4 chunks
        assert gpu_decompress(ctx, pkg, s, len(data)) == (len(data), data)
    # chunks of another size than the reference writer's: the prediction fails, the chunk-by-chunk reader decides
    a, b = bs.compress(data[:70000])[4:], bs.compress(data[70000:200000])[4:]
    foreign = (200000).to_bytes(4, "big") + a + b
    assert gpu_decompress(ctx, pkg, foreign, 200000) == (200000, data[:200000])
    # errors as the CPU reader reports them
    s = bs.compress(data[:200000], 50000)
    for bad in (s[:-10], s[:6]):
        assert gpu_decompress(ctx, pkg, bad, 200000)[0] == bs.decompress(bad, 200000)[0] == -2
    bad = bytearray(s); bad[8] = 0xff
    assert gpu_decompress(ctx, pkg, bytes(bad), 200000)[0] == -4
    assert gpu_decompress(ctx, pkg, s, 199999)[0] == -3
