"""4mz / zstd frame encoding (SURVEY.md rows a8 writer, a10).

Compressed bytes need not match the reference (BASELINE.json north_star); the bar is that the
reference's ZSTD_decompress / `4mc -z -d` restores the input exactly.

CPU half: the product's encoder source (4mc_b200/csrc/zstd_encode.h) is written as CTA phases;
tests/native/zenc_emul.cpp runs the phases with a loop over thread ids, and every frame it makes is
decoded by the reference (oracle/_ref, where built) and by the product's own frame decoder built for
the host.  GPU half: the kernels through the C-ABI, the CLI and the reference CLI."""
import ctypes as C
import os
import random
import subprocess

import pytest

from conftest import ROOT, golden_bytes, gen_logtext, build_native

MIB = 1024 * 1024
CLI = os.path.join(ROOT, "4mc_b200", "host", "4mc")


@pytest.fixture(scope="module")
def zenc():
    Z = build_native("zenc_emul", ["tests/native/zenc_emul.cpp"],
                     deps=["4mc_b200/csrc/zstd_encode.h", "4mc_b200/csrc/zstd_decode.h"])
    Z.zenc_emul_compress.restype = C.c_longlong
    Z.zenc_emul_compress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong, C.c_int]
    return Z


@pytest.fixture(scope="module")
def zdec():
    Z = build_native("zstd_shim", ["tests/native/zstd_shim.cpp"], deps=["4mc_b200/csrc/zstd_decode.h"])
    Z.zstd_shim_decompress.restype = C.c_longlong
    Z.zstd_shim_decompress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong]

    def dec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = Z.zstd_shim_decompress(out, cap, src, len(src))
        return int(r), out.raw[:max(r, 0)]
    return dec


def ref_decoder():
    p = os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    R.ZSTD_decompress.restype = C.c_size_t
    R.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    R.ZSTD_isError.restype = C.c_uint
    R.ZSTD_isError.argtypes = [C.c_size_t]

    def dec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = R.ZSTD_decompress(out, cap, src, len(src))
        if R.ZSTD_isError(r):
            return -1, b""
        return int(r), out.raw[:r]
    return dec


def sample_inputs(pkg):
    rng = random.Random(7)
    text = gen_logtext(pkg, 600000)
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    return {
        "empty": b"", "one": b"A", "tiny": b"abc" * 5, "zeros_64k": bytes(65536), "zeros_300k": bytes(300000),
        "text_300": text[:300], "text_5k": text[:5000], "text_64k": text[:65536], "text_64k+1": text[:65537],
        "text_600k": text,
        "random_70k": golden_bytes("random_70000.bin"),
        "skew_all_bytes": bytes(min(255, int(rng.expovariate(0.05))) for _ in range(150000)),      # symbols above 128: FSE-coded tree
        "two_symbols": bytes(rng.choice(b"ab") for _ in range(100000)),
        "high_symbols": bytes(rng.choice(b"\xf0\xf1\xf2\xf3\x80\x81 abcdefgh") for _ in range(100000)),
        "words": b" ".join(rng.choice(words) for _ in range(40000)),
        "ramp": bytes((i * 7 + (i >> 8)) & 255 for i in range(200000)),
        "long_runs": b"".join(bytes([rng.randrange(256)]) * rng.randint(1, 5000) for _ in range(200)),
        "one_literal_byte": b"x" * 300 + bytes(100000) + b"x" * 300,
    }


def test_emulated_encoder_frames_decode(zenc, zdec, pkg, ora):
    ref = ref_decoder()
    for name, data in sample_inputs(pkg).items():
        for mm in (4, 5, 4 | (1 << 4), 5 | (2 << 4)):           # bits 4..: the stand-in matcher (recent / first / first + repeat offset)
            cap = len(data) + len(data) // 64 + 1024
            out = C.create_string_buffer(cap)
            c = zenc.zenc_emul_compress(out, cap, data, len(data), mm)
            assert c > 0, name
            frame = out.raw[:c]
            assert c <= len(data) + 12 + 3 * ((len(data) + 65535) // 65536), name     # never worse than raw blocks
            assert zdec(frame, len(data)) == (len(data), data), name
            assert ora.zstd_decompress(frame, len(data)) == (len(data), data), name      # the strict oracle
            if ref:
                assert ref(frame, len(data)) == (len(data), data), name
    # the encoder must actually compress: log text well under half, zeros to almost nothing
    text = gen_logtext(pkg, 600000)
    out = C.create_string_buffer(len(text) + 1024)
    assert zenc.zenc_emul_compress(out, len(out), text, len(text), 5) < len(text) * 0.40
    assert zenc.zenc_emul_compress(out, len(out), bytes(300000), 300000, 5) < 200


def test_fse_table_description_round_trip(zenc):
    rng = random.Random(11)
    sizes = []
    for _ in range(3000):
        max_sym = rng.randint(1, 52)
        log = rng.randint(5, 9)
        shape = rng.random()
        count = [0] * 64
        for s in range(max_sym + 1):
            if shape < 0.3:
                count[s] = rng.randint(0, 3)
            elif shape < 0.6:
                count[s] = int(rng.expovariate(1 / 200.0)) if rng.random() < 0.7 else 0
            else:
                count[s] = rng.choice((0, 0, 1, 1, 2, 5000, 30))
        count[max_sym] = max(count[max_sym], 1)
        count[rng.randrange(max_sym)] += 1                     # at least two symbols present
        if sum(1 for c in count if c) > (1 << log):
            continue
        arr = (C.c_uint32 * 64)(*count)
        r = zenc.zenc_emul_ncount_roundtrip(arr, max_sym, log)
        assert r > 0, (count, max_sym, log, r)
        sizes.append(r)
    assert len(sizes) > 2000 and max(sizes) < 100


def test_huffman_lengths_are_complete_and_limited(zenc):
    rng = random.Random(5)
    for trial in range(400):
        n = rng.randint(2, 256)
        syms = rng.sample(range(256), n)
        count = [0] * 256
        kind = trial % 4
        for i, s in enumerate(syms):
            if kind == 0:
                count[s] = rng.randint(1, 1000)
            elif kind == 1:
                count[s] = 1 << min(i, 15)                     # Fibonacci-like depth: forces the length limit
            elif kind == 2:
                count[s] = 1
            else:
                count[s] = max(1, int(rng.expovariate(1 / 50.0)))
        arr = (C.c_uint32 * 256)(*count)
        nb = (C.c_uint8 * 256)()
        log = zenc.zenc_emul_huffman_check(arr, nb)
        assert 1 <= log <= 11, (trial, log)


# ---- GPU: through the C-ABI ----------------------------------------------------------------------

def walk_4mz(stream):
    assert stream[:12] == bytes.fromhex("344d5a00 00000001 289a1c9a")
    pos, out = 12, []
    while True:
        u, c = int.from_bytes(stream[pos:pos + 4], "big"), int.from_bytes(stream[pos + 4:pos + 8], "big")
        if u == 0:
            return out
        out.append((u, c, stream[pos + 12:pos + 12 + c]))
        pos += 12 + c


@pytest.mark.gpu
def test_gpu_4mz_streams_decode_everywhere(ctx, pkg, ora, zdec, ref_cli, tmp_path):
    """GPU-written 4mz: decoded by the reference CLI (`4mc -z -d`), by the oracle-independent host build
    of the frame decoder block by block, and by the GPU reader."""
    inputs = sample_inputs(pkg)
    inputs["text_9m"] = gen_logtext(pkg, 9 * MIB + 777)
    inputs["mixed_12m"] = gen_logtext(pkg, 4 * MIB) + os.urandom(4 * MIB) + bytes(4 * MIB - 5)     # compressed, stored, tiny
    for name, data in inputs.items():
        stream = ctx.compress_4mz(data)
        blocks = walk_4mz(stream)
        assert sum(u for u, _, _ in blocks) == len(data), name
        got = b""
        for u, c, payload in blocks:
            assert c <= u, name                                   # stored fallback (native/4mc.c:469-485)
            if c == u:
                got += payload
            else:
                r, out = zdec(payload, u)
                assert r == u, name
                got += out
        assert got == data, name
        assert ctx.decompress_4mz(stream) == data, name
        assert ora.decompress_4mz(stream, len(data)) == (len(data), data), name          # the strict oracle
        src, out = tmp_path / "s.4mz", tmp_path / "s.out"
        src.write_bytes(stream)
        subprocess.run([ref_cli, "-f", "-q", "-q", "-z", "-d", str(src), str(out)], check=True)
        assert out.read_bytes() == data, name
    stream = ctx.compress_4mz(inputs["mixed_12m"])
    kinds = [c == u for u, c, _ in walk_4mz(stream)]
    assert kinds == [False, True, False]
    assert len(ctx.compress_4mz(inputs["text_9m"])) < len(inputs["text_9m"]) * 0.42


@pytest.mark.gpu
def test_gpu_4mz_empty_and_footer_match_reference_layout(ctx):
    assert ctx.compress_4mz(b"") == golden_bytes("empty.4mz")
    assert ctx.compress_4mz(b"A") == golden_bytes("A.4mz")        # stored block: byte-identical to the reference


@pytest.mark.gpu
def test_gpu_zstd_compress_per_block(ctx, pkg, zdec):
    """fourmc_zstd_compress = ZSTD_compress on one block (native/jniZstdCompressor.c:93)."""
    ref = ref_decoder()
    text = gen_logtext(pkg, 4 * MIB)
    for data in (b"", b"A", text[:100], text[:70000], text, os.urandom(200000), bytes(4 * MIB)):
        frame = ctx.zstd_compress(data)
        assert frame is not None and len(frame) <= int(pkg.lib().fourmc_zstd_compress_bound(len(data)))
        assert zdec(frame, len(data)) == (len(data), data)
        if ref:
            assert ref(frame, len(data)) == (len(data), data)
        assert pkg.ZstdDecompressor(ctx).decompress_bytes_direct(frame) == data
    assert ctx.zstd_compress(os.urandom(100000), capacity=1000) is None      # dstSize_tooSmall
    assert len(pkg.ZstdCompressor(ctx).compress_bytes_direct(text)) < len(text) * 0.42


@pytest.mark.gpu
def test_gpu_4mz_device_calls(ctx, pkg):
    """Device-resident writer (whole stream and span + index) against the device-resident reader."""
    import torch
    n = 21 * MIB + 12345
    pages = (n + 4095) // 4096
    d_in = torch.empty(pages * 4096, dtype=torch.uint8, device="cuda")
    ctx.gen_device(d_in.data_ptr(), pages)
    ctx.sync()
    cap = int(pkg.lib().fourmc_4mc_bound(n))
    d_out = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    d_size = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_lens = torch.zeros(6, dtype=torch.int32, device="cuda")
    ctx.compress_4mz_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, d_size.data_ptr(), d_lens.data_ptr())
    ctx.sync()
    size = int(d_size.item())
    assert 0 < size < n * 0.45
    lens = d_lens.tolist()
    assert sum(lens) + 12 + 12 + 20 + 4 * 6 == size
    d_back = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx.decompress_4mz_device(d_out.data_ptr(), size, d_back.data_ptr(), n, d_res.data_ptr())
    ctx.sync()
    assert d_res.tolist() == [n, -1]
    assert torch.equal(d_back[:n], d_in[:n])
    # span + index: what a sharded writer does
    d_span = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    d_ssz = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_l2 = torch.zeros(6, dtype=torch.int32, device="cuda")
    ctx.compress_4mz_span_device(d_in.data_ptr(), n, d_span.data_ptr(), cap, d_ssz.data_ptr(), d_l2.data_ptr())
    d_hdr = torch.zeros(16, dtype=torch.uint8, device="cuda")
    d_tail = torch.zeros(12 + 20 + 4 * 6 + 16, dtype=torch.uint8, device="cuda")
    ctx.build_index_4mz_device(d_l2.data_ptr(), 6, d_hdr.data_ptr(), d_tail.data_ptr())
    ctx.sync()
    span = int(d_ssz.item())
    whole = bytes(d_hdr[:12].cpu().numpy()) + bytes(d_span[:span].cpu().numpy()) + bytes(d_tail[:12 + 20 + 24].cpu().numpy())
    host = bytes(d_in[:n].cpu().numpy())
    assert ctx.decompress_4mz(whole) == host
    # small groups (several batches appended through the device-side carry) decode to the same bytes
    os.environ["FOURMC_ZGROUP"] = "2"
    try:
        d_out.zero_()
        ctx.compress_4mz_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, d_size.data_ptr(), d_lens.data_ptr())
        ctx.sync()
        size2 = int(d_size.item())
        assert ctx.decompress_4mz(bytes(d_out[:size2].cpu().numpy())) == host
        assert ctx.decompress_4mz(ctx.compress_4mz(host)) == host
        assert ctx.decompress_4mz(ctx.compress_4mz(host, 3)) == host      # the chain links are rebuilt per group
    finally:
        del os.environ["FOURMC_ZGROUP"]


@pytest.mark.gpu
def test_cli_writes_4mz_the_reference_reads(pkg, ref_cli, tmp_path):
    data = gen_logtext(pkg, 5 * MIB + 99, first_page=17)
    src, ours, back, back2 = tmp_path / "in.txt", tmp_path / "in.4mz", tmp_path / "back", tmp_path / "back2"
    src.write_bytes(data)
    assert subprocess.run([CLI, "-f", "-q", "-z", "-1", str(src), str(ours)]).returncode == 0
    subprocess.run([ref_cli, "-f", "-q", "-q", "-z", "-d", str(ours), str(back)], check=True)
    assert back.read_bytes() == data
    assert subprocess.run([CLI, "-f", "-q", "-z", "-d", str(ours), str(back2)]).returncode == 0
    assert back2.read_bytes() == data
