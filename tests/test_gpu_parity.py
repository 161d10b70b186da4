"""Parity tests proper: the CUDA path, called through the C-ABI of lib4mcgpu.so, against the oracle,
the committed golden vectors (outputs of the reference) and -- where oracle/_ref travelled -- the
reference CLI itself.  Bit-exact throughout (integer / byte work)."""
import ctypes as C
import random
import subprocess

import pytest

from conftest import golden_bytes, golden_json, gen_logtext

pytestmark = pytest.mark.gpu
MIB = 1024 * 1024


# ---------------------------------------------------------------- XXH32

def test_xxh32_golden(ctx, lcg):
    for v in golden_json("xxh32.json"):
        data = bytes.fromhex(v["hex"]) if "hex" in v else lcg[v["lcg_off"]:v["lcg_off"] + v["lcg_len"]]
        assert ctx.xxh32(data, v["seed"]) == v["xxh32"], v


def test_xxh32_vs_oracle_large(ctx, ora, pkg):
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    for n in (4 * 1024 * 1024, 4 * 1024 * 1024 - 1, 2035746, 2048, 2049, 2063, 2064, 2065, 4095):
        for off in (0, 1, 7, 12):
            assert ctx.xxh32(text[off:off + n - off]) == ora.xxh32(text[off:off + n - off])


def test_xxh32_batch_device(ctx, ora, pkg):
    import torch
    data = gen_logtext(pkg, 1 << 20)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    rng = random.Random(9)
    items = [(rng.randrange(0, 1 << 19), rng.choice([0, 1, 15, 16, 17, 100, 4096, 70001, 1 << 19])) for _ in range(300)]
    off = torch.tensor([o for o, _ in items], dtype=torch.int64).cuda()
    ln = torch.tensor([l for _, l in items], dtype=torch.int32).cuda()
    out = torch.zeros(len(items), dtype=torch.int32).cuda()
    torch.cuda.synchronize()
    ctx.xxh32_batch_device(len(items), t.data_ptr(), off.data_ptr(), ln.data_ptr(), out.data_ptr(), seed=0,
                           stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = [x & 0xFFFFFFFF for x in out.cpu().tolist()]
    assert got == [ora.xxh32(data[o:o + l]) for o, l in items]


# ---------------------------------------------------------------- LZ4 decode

def test_lz4_decode_golden(ctx, ora):
    """Return value (incl. the exact negative code) and bytes equal LZ4_decompress_safe's."""
    for v in golden_json("lz4_decode.json"):
        src = bytes.fromhex(v["hex"])
        r, out = ctx.lz4_decompress_safe(src, v["cap"])
        assert r == v["ret"], (v["hex"][:40], v["cap"], r, v["ret"])
        assert ora.xxh32(out) == v["out_xxh32"]


def test_lz4_decode_vs_oracle_blocks(ctx, ora, pkg):
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    rng = random.Random(4)
    cases = [text, text[:4194303], text[:65537], bytes(4 * 1024 * 1024), (b"abcdefg" * 600000)[:4 * 1024 * 1024],
             bytes(rng.getrandbits(8) for _ in range(100000)) + text[:200000],
             b"".join(bytes([rng.getrandbits(8)]) * rng.choice([1, 2, 3, 40, 300, 5000]) for _ in range(3000))]
    for src in cases:
        comp = ora.lz4_compress(src)
        for cap in (len(src), 4 * 1024 * 1024):
            if cap < len(src):
                continue
            r, out = ctx.lz4_decompress_safe(comp, cap)
            assert r == len(src) and out == src
        # capacity one short: the reference fails, with a specific code
        if len(src) > 64:
            assert ctx.lz4_decompress_safe(comp, len(src) - 1)[0] == ora.lz4_decompress(comp, len(src) - 1)[0]


def test_lz4_decode_mutations_vs_oracle(ctx, ora, pkg):
    rng = random.Random(8)
    text = gen_logtext(pkg, 100000)
    comp = ora.lz4_compress(text)
    for k in range(150):
        m = bytearray(comp)
        i = rng.randrange(len(m))
        if k % 3 == 0:
            m[i] = rng.getrandbits(8)
        elif k % 3 == 1:
            m = m[:i + 1]
        else:
            m[i:i] = bytes([rng.choice([0, 255, 0xF0, 0x0F])])
        m = bytes(m)
        cap = rng.choice([len(text), len(text) + 50, 4 << 20])
        r, out = ctx.lz4_decompress_safe(m, cap)
        er, eout = ora.lz4_decompress(m, cap)
        assert r == er and out == eout


# ---------------------------------------------------------------- container decode

def test_container_decode_golden(ctx):
    text = golden_bytes("logtext_128k.bin")
    for lvl in (1, 2, 3, 4):                      # reference-compressed at all four LZ4 levels
        assert ctx.decompress_4mc(golden_bytes(f"logtext_128k.l{lvl}.4mc")) == text
    assert ctx.decompress_4mc(golden_bytes("empty.4mc")) == b""
    assert ctx.decompress_4mc(golden_bytes("A.4mc")) == b"A"
    assert ctx.decompress_4mc(golden_bytes("zeros_4m1.4mc")) == bytes(4 * 1024 * 1024 + 1)
    assert ctx.decompress_4mc(golden_bytes("random_70000.4mc")) == golden_bytes("random_70000.bin")
    assert ctx.decompress_4mc(golden_bytes("two_streams.4mc")) == b"A" + text


@pytest.mark.parametrize("zstd", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_writer_and_reader_share_one_stream(ctx, ora, ref_cli, pkg, tmp_path, world, zstd):
    """SURVEY.md 8e on the GPU, the ranks played by separate contexts in one process: contiguous block ranges, every
    rank's span from fourmc_*_compress_span_device, the gathered block lengths give the span offsets and the footer
    (fourmc_*_build_index_device), the spans are placed -- and the result is ONE stream: the oracle and the reference CLI
    decode it to the input, and every rank decodes its own block range of it through the footer index
    (fourmc_*_decompress_range_device).  bench.py --gpus N runs the same steps with NCCL between processes."""
    import subprocess
    import torch
    n = 11 * 4194304 + 33333                                       # a ragged last block
    data = gen_logtext(pkg, n, seed=5)
    nb = (n + 4194304 - 1) // 4194304
    ranks = [pkg.Context(0) for _ in range(world)]
    try:
        d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        spans, lens_all, sizes = [], [], []
        for r, c in enumerate(ranks):
            lo, hi = pkg.shard_blocks(nb, world, r)
            off, cnt = lo * 4194304, min(n, hi * 4194304) - lo * 4194304
            cap = cnt + 12 * (hi - lo) + 64
            d_span = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
            d_l = torch.zeros(max(hi - lo, 1), dtype=torch.int32, device="cuda")
            f = c.compress_4mz_span_device if zstd else c.compress_span_device
            f(d_in.data_ptr() + off, cnt, d_span.data_ptr(), cap, d_sz.data_ptr(), d_l.data_ptr())
            c.sync()
            spans.append(d_span[:int(d_sz.item())].clone()); sizes.append(int(d_sz.item())); lens_all.append(d_l[:hi - lo])
        lens = torch.cat(lens_all)
        assert [int(x.sum()) for x in lens_all] == sizes            # a span is exactly its block records
        offs = pkg.span_base_offsets(sizes)
        total = 12 + sum(sizes) + 12 + 20 + 4 * nb
        d_stream = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        for r in range(world):
            d_stream[offs[r]:offs[r] + sizes[r]].copy_(spans[r])
        build = ranks[0].build_index_4mz_device if zstd else ranks[0].build_index_device
        build(lens.data_ptr(), nb, d_stream.data_ptr(), d_stream.data_ptr() + 12 + sum(sizes))
        ranks[0].sync()
        stream = bytes(d_stream[:total].cpu().numpy())
        dec = ora.decompress_4mz if zstd else ora.decompress_4mc
        assert dec(stream, n + 16) == (n, data)
        p, q = tmp_path / ("s.4mz" if zstd else "s.4mc"), tmp_path / "s.out"
        p.write_bytes(stream)
        subprocess.run([ref_cli, "-f", "-q", "-q", "-d"] + (["-z"] if zstd else []) + [str(p), str(q)], check=True)
        assert q.read_bytes() == data
        # the readers: every rank its own range of the ONE stream
        for r, c in enumerate(ranks):
            lo, hi = pkg.shard_blocks(nb, world, r)
            want = data[lo * 4194304:min(n, hi * 4194304)]
            d_out = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
            res = torch.zeros(2, dtype=torch.int64, device="cuda")
            c.decompress_range_device(d_stream.data_ptr(), total, lo, hi - lo, d_out.data_ptr(), len(want), res.data_ptr(), zstd=zstd)
            c.sync()
            assert res.cpu().tolist() == [len(want), -1]
            assert bytes(d_out[:len(want)].cpu().numpy()) == want
            # a rank that holds only ITS part of the stream (header, own span, end mark + footer): a range reader
            # looks at its own blocks only
            d_part = torch.full((total + 64,), 0xAB, dtype=torch.uint8, device="cuda")
            d_part[:12].copy_(d_stream[:12])
            d_part[offs[r]:offs[r] + sizes[r]].copy_(spans[r])
            d_part[12 + sum(sizes):total].copy_(d_stream[12 + sum(sizes):total])
            d_out.zero_()
            c.decompress_range_device(d_part.data_ptr(), total, lo, hi - lo, d_out.data_ptr(), len(want), res.data_ptr(), zstd=zstd)
            c.sync()
            assert res.cpu().tolist() == [len(want), -1]
            assert bytes(d_out[:len(want)].cpu().numpy()) == want
    finally:
        for c in ranks:
            c.close()


def test_compressed_bytes_are_reproducible(ctx, pkg):
    """The match finders' tables are filled by racing stores; ties are settled by position (lowest / highest),
    so two runs -- and two contexts -- give the same bytes (a storage format wants reproducible artefacts)."""
    data = gen_logtext(pkg, 3 * 4194304 + 4321, seed=77)
    other = pkg.Context(0)
    try:
        for level in (1, 2, 3, 4):
            a = ctx.compress_4mc(data, level=level)
            assert ctx.compress_4mc(data, level=level) == a and other.compress_4mc(data, level=level) == a, level
        for level in (1, 3):
            z = ctx.compress_4mz(data, level=level)
            assert ctx.compress_4mz(data, level=level) == z and other.compress_4mz(data, level=level) == z, level
        # the device-resident calls favour speed unless asked (fourmc_ctx_set_reproducible)
        import torch
        d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        cap = pkg.lib().fourmc_4mc_bound(len(data))
        outs = []
        for c in (ctx, other):
            c.set_reproducible(1)
            d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            d_sz = torch.zeros(1, dtype=torch.int64, device="cuda")
            c.compress_device(d_in.data_ptr(), len(data), d_out.data_ptr(), cap, d_sz.data_ptr())
            c.sync()
            outs.append(bytes(d_out[:int(d_sz.item())].cpu().numpy()))
            c.set_reproducible(-1)
        assert outs[0] == outs[1] == ctx.compress_4mc(data, level=1)
    finally:
        ctx.set_reproducible(-1)
        other.close()


def test_short_block_and_empty_stream(ctx, ora, pkg):
    """A block that decodes to fewer bytes than its header announces moves the following blocks down
    (native/4mc.c:661-666); a stream that decodes to nothing ends the loop over streams (:909-913).  Host call,
    device call and the split reader against the oracle (itself checked against the reference CLI)."""
    import torch
    from conftest import short_block_stream
    stream, expect = short_block_stream(ora)
    assert ora.decompress_4mc(stream, len(expect) + 4096) == (len(expect), expect)
    assert ctx.decompress_4mc(stream) == expect
    d_in = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
    d_out = torch.zeros(len(expect) + 4096, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx.decompress_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), d_out.numel(), res.data_ptr())
    torch.cuda.synchronize()
    assert res.cpu().tolist() == [len(expect), -1]
    assert bytes(d_out[:len(expect)].cpu().numpy()) == expect
    assert ctx.decompress_4mc(golden_bytes("empty.4mc") + golden_bytes("A.4mc")) == b""
    assert ctx.decompress_4mc(golden_bytes("A.4mc") + golden_bytes("empty.4mc") + golden_bytes("A.4mc")) == b"A"


def test_container_decode_errors_match_reference_exit_codes(ctx, ora):
    good = golden_bytes("logtext_128k.l1.4mc")
    n = 128 * 1024
    muts = []
    for pos in (100, 9, 7, len(good) - 1, len(good) - 30, 14, 20, 5000):
        b = bytearray(good); b[pos] ^= 0x10; muts.append(bytes(b))
    muts += [good[:20], good[:30000], good[:-5], good.replace(b"4MC\0", b"4MZ\0", 1), good + good[:7]]
    for m in muts:
        assert ctx.decompress_4mc_rc(m, n + 16) == ora.decompress_4mc(m, n + 16)[0]


# ---------------------------------------------------------------- compress

SIZES = [0, 1, 5, 12, 13, 14, 131, 132, 133, 4096, 65535, 65536, 65537, 65549, 131072 + 5, 200000,
         4194304 - 1, 4194304, 4194304 + 1, 4194304 + 65536 + 7, 3 * 4194304 + 12345]


@pytest.mark.parametrize("n", SIZES)
def test_compress_4mc_roundtrip_text(ctx, ora, pkg, n):
    data = gen_logtext(pkg, n)
    stream = ctx.compress_4mc(data)
    r, out = ora.decompress_4mc(stream, n)            # oracle = reference reader semantics
    assert r == n and out == data
    assert ctx.decompress_4mc(stream) == data         # and our own reader


def test_compress_4mc_special_inputs(ctx, ora):
    rng = random.Random(2)
    n = 4 * 1024 * 1024 + 70001
    rnd = rng.randbytes(n)
    cases = {"random": rnd, "zeros": bytes(n), "period7": (b"abcdefg" * (n // 7 + 1))[:n],
             "mixed": rnd[:100000] + bytes(300000) + rnd[:50000] * 3 + b"xyz" * 100000}
    for name, data in cases.items():
        stream = ctx.compress_4mc(data)
        r, out = ora.decompress_4mc(stream, len(data))
        assert r == len(data) and out == data, name
        assert ctx.decompress_4mc(stream) == data, name
    # incompressible blocks are stored: file = input + 12 header + 12/block + 12 EOS + footer
    s = ctx.compress_4mc(rnd)
    assert len(s) == n + 12 + 2 * 12 + 12 + 20 + 2 * 4
    # empty and 1-byte files are byte-identical to the reference's
    assert ctx.compress_4mc(b"") == golden_bytes("empty.4mc")
    assert ctx.compress_4mc(b"A") == golden_bytes("A.4mc")
    assert ctx.compress_4mc(golden_bytes("random_70000.bin")) == golden_bytes("random_70000.4mc")


def test_compress_ratio_not_worse_than_reference(ctx, pkg):
    data = gen_logtext(pkg, 8 * 1024 * 1024)
    assert len(data) / len(ctx.compress_4mc(data)) > 2.0      # reference: 2.02 on this input


def test_lz4_compress_block_api(ctx, ora, pkg):
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    for n in (0, 1, 13, 100, 65536, 1000000, 4 * 1024 * 1024):
        c = ctx.lz4_compress(text[:n])
        assert c is not None and len(c) <= pkg.Lz4Compressor.compress_bound(n)
        assert ora.lz4_decompress(c, n) == (n, text[:n])
    rnd = random.Random(1).randbytes(100000)
    c = ctx.lz4_compress(rnd)                               # bound-sized destination: always fits
    assert ora.lz4_decompress(c, len(rnd)) == (len(rnd), rnd)
    assert ctx.lz4_compress(rnd, capacity=len(rnd) - 1) is None      # native/4mc.c:301 -> stored


def test_reference_cli_decodes_gpu_output(ctx, ref_cli, pkg, tmp_path):
    data = gen_logtext(pkg, 9 * 1024 * 1024 + 4321) + bytes(100000) + random.Random(3).randbytes(4 * 1024 * 1024 + 5)
    p, q = tmp_path / "g.4mc", tmp_path / "g.out"
    p.write_bytes(ctx.compress_4mc(data))
    subprocess.run([ref_cli, "-f", "-q", "-q", "-d", str(p), str(q)], check=True)
    assert q.read_bytes() == data


def test_gpu_decodes_reference_cli_output(ctx, ref_cli, pkg, tmp_path):
    data = gen_logtext(pkg, 9 * 1024 * 1024 + 4321) + bytes(100000)
    p = tmp_path / "r.bin"
    p.write_bytes(data)
    for lvl in (1, 2, 3, 4):
        q = tmp_path / f"r{lvl}.4mc"
        subprocess.run([ref_cli, "-f", "-q", "-q", f"-{lvl}", str(p), str(q)], check=True)
        assert ctx.decompress_4mc(q.read_bytes()) == data


# ---------------------------------------------------------------- device-resident pipeline

def _device_roundtrip(ctx, pkg, n_bytes, seed=0x4D43):
    import torch
    st = None                              # the context's own stream
    pages = (n_bytes + 4095) // 4096
    src = torch.empty(pages * 4096, dtype=torch.uint8, device="cuda")
    cap = pkg.lib().fourmc_4mc_bound(n_bytes)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.zeros(n_bytes + 16, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()               # torch's fills run on another stream than the context's
    ctx.gen_device(src.data_ptr(), pages, seed=seed, stream=st)
    ctx.compress_device(src.data_ptr(), n_bytes, comp.data_ptr(), cap, size.data_ptr(), stream=st)
    torch.cuda.synchronize()
    csz = int(size.item())
    ctx.decompress_device(comp.data_ptr(), csz, out.data_ptr(), n_bytes, res.data_ptr(), stream=st)
    torch.cuda.synchronize()
    return src, comp[:csz], out, res.cpu().tolist()


@pytest.mark.parametrize("n", [4096, 4 * 1024 * 1024 + 4096, 64 * 1024 * 1024, 1024 * 1024 * 1024 + 8192])
def test_device_roundtrip(ctx, ora, pkg, n):
    import torch
    src, comp, out, res = _device_roundtrip(ctx, pkg, n)
    assert res == [n, -1]
    assert torch.equal(out[:n], src[:n])
    if n >= 4 * 1024 * 1024:
        assert n / comp.numel() > 1.9
    if n <= 64 * 1024 * 1024:
        # the generator on the device equals the host generator; the stream decodes with the oracle
        host = gen_logtext(pkg, n)
        assert bytes(src[:n].cpu().numpy().tobytes()) == host
        r, o = ora.decompress_4mc(comp.cpu().numpy().tobytes(), n)
        assert r == n and o == host


def test_device_decode_detects_corruption(ctx, pkg):
    import torch
    n = 16 * 1024 * 1024
    src, comp, out, res = _device_roundtrip(ctx, pkg, n)
    st = None
    for pos, expect_block in ((5 * 1024 * 1024, None), (30, 0)):
        bad = comp.clone()
        bad[pos] ^= 0x40
        r = torch.zeros(2, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.decompress_device(bad.data_ptr(), bad.numel(), out.data_ptr(), n, r.data_ptr(), stream=st)
        torch.cuda.synchronize()
        code, blk = r.cpu().tolist()
        assert code == pkg.E_CONTENT and blk >= 0
        if expect_block is not None:
            assert blk == expect_block


def test_levels_2_to_4_chain_parse(ctx, ora, pkg, ref_cli, tmp_path):
    """Levels 2..4 (SURVEY rows a11 / a12, and the upper 4mz levels of a10): the hash-chain parse.  Valid
    streams (oracle, reference CLI), ratio above the Fast level and growing with the level."""
    import os
    import subprocess
    text = gen_logtext(pkg, 9 * MIB + 4321, first_page=5)
    mixed = text[:3 * MIB] + bytes(300000) + os.urandom(MIB) + (b"abcdefg" * 200000)[:MIB] + text[:100]
    sizes = {}
    for level in (1, 2, 3, 4):
        for name, data in (("text", text), ("mixed", mixed)):
            s = ctx.compress_4mc(data, level)
            assert ora.decompress_4mc(s, len(data)) == (len(data), data), (level, name)
            assert ctx.decompress_4mc(s) == data
            z = ctx.compress_4mz(data, level)
            assert ctx.decompress_4mz(z) == data, (level, name)
            if name == "text":
                sizes[level] = (len(s), len(z))
            if level == 3:
                for blob, flag in ((s, []), (z, ["-z"])):
                    src, out = tmp_path / "lv.bin", tmp_path / "lv.out"
                    src.write_bytes(blob)
                    subprocess.run([ref_cli, "-f", "-q", "-q"] + flag + ["-d", str(src), str(out)], check=True)
                    assert out.read_bytes() == data
    for k in (0, 1):
        assert sizes[4][k] < sizes[3][k] < sizes[2][k] < sizes[1][k], sizes
    assert len(text) / sizes[2][0] > 2.45 and len(text) / sizes[3][0] > 2.6 and len(text) / sizes[4][0] > 2.65, sizes
    # row a12's bar: High / Ultra compress at least as well as the reference's HC 4 / HC 8 (native/4mc.c:248-251)
    import ctypes as C
    rp = os.path.join(os.path.dirname(ref_cli), "libref4mc.so")
    R = C.CDLL(rp)
    R.LZ4_compress_HC.restype = C.c_int
    R.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    dst = C.create_string_buffer(4 * MIB + 4 * MIB // 255 + 64)
    for level, hc in ((3, 4), (4, 8)):
        theirs = sum(12 + R.LZ4_compress_HC(text[o:o + 4 * MIB], dst, len(text[o:o + 4 * MIB]), len(dst), hc)
                     for o in range(0, len(text), 4 * MIB))
        assert sizes[level][0] <= theirs + 64, (level, len(text) / sizes[level][0], len(text) / theirs)
    # the chain links are exact, so the bytes are the same on every run and however the blocks are grouped
    whole = ctx.compress_4mc(text, 3)
    os.environ["FOURMC_CHAIN_GROUP"] = "1"
    try:
        assert ctx.compress_4mc(text, 3) == whole
    finally:
        del os.environ["FOURMC_CHAIN_GROUP"]
    # per-block calls: LZ4_compressMC / LZ4_compressHC2 / ZSTD_compress(level) of the JNI natives
    blk = text[:4 * MIB]
    for level in (2, 3, 4):
        c = ctx.lz4_compress(blk, level)
        assert ora.lz4_decompress(c, len(blk)) == (len(blk), blk)
        f = ctx.zstd_compress(blk, level)
        assert ctx.zstd_decompress(f, len(blk)) == (len(blk), blk)


@pytest.mark.parametrize("n", [1, 12, 13, 4095, 65535, 65537, 98305, 131071, 4 * MIB - 1, 4 * MIB + 1, 5 * MIB + 77])
def test_levels_2_to_4_ragged_sizes(ctx, ora, pkg, n):
    """The chain kernels at sizes around their own boundaries (the 32-position steps and 512-position pieces of the link
    kernel, 64 KiB of warm-up, the 32 KiB regions, the block end rules): every stream decodes to the input, on text, on a
    period shorter than a step (every step holds equal hashes) and on zeros (every match hits the length cap)."""
    text = gen_logtext(pkg, n, first_page=3)
    for data in (text, (b"abcdefg" * (n // 7 + 1))[:n], bytes(n)):
        for level in (2, 3, 4):
            s = ctx.compress_4mc(data, level)
            assert ora.decompress_4mc(s, len(data)) == (len(data), data), (n, level)
        z = ctx.compress_4mz(data, 3)
        assert ctx.decompress_4mz(z) == data, n


def test_generators_host_equals_device_and_round_trip(ctx, ora, pkg, ref_cli, tmp_path):
    """The JSON (configs[2]) and silesia-like mix (configs[3]) inputs: the device generator is bit-identical to
    the host one, and both containers restore them -- the mix holds incompressible blocks (stored path)."""
    import torch
    for kind, seed, pages in ((1, 0x4D5A, 3 * 1024 + 7), (2, 0x5148, 9 * 1024)):
        n = pages * 4096
        host = C.create_string_buffer(n)
        assert pkg.lib().fourmc_gen_host(kind, seed, 5, pages, host) == 0
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        ctx.gen_device(dev.data_ptr(), pages, seed=seed, first_page=5, kind=kind)
        ctx.sync()
        data = host.raw
        assert bytes(dev.cpu().numpy()) == data, kind
        s = ctx.compress_4mc(data)
        assert ora.decompress_4mc(s, n) == (n, data) and ctx.decompress_4mc(s) == data
        z = ctx.compress_4mz(data)
        assert ctx.decompress_4mz(z) == data
        src, out = tmp_path / "g.4mz", tmp_path / "g.out"
        src.write_bytes(z)
        subprocess.run([ref_cli, "-f", "-q", "-q", "-z", "-d", str(src), str(out)], check=True)
        assert out.read_bytes() == data
        if kind == 1:
            assert n / len(z) > 2.5 and n / len(s) > 1.9           # reference: ZSTD 1 3.35, LZ4 2.12 on this input
        else:
            stored = [int.from_bytes(s[o + 4:o + 8], "big") == int.from_bytes(s[o:o + 4], "big") for o in ctx.read_index(s)]
            assert any(stored) and not all(stored)
    assert pkg.lib().fourmc_gen_host(3, 1, 0, 1, C.create_string_buffer(4096)) != 0


def test_distinct_contexts_are_thread_safe(pkg, ora):
    """include/fourmc.h: a context is single-threaded, distinct contexts are independent (the Java objects
    call the natives from many task threads at once)."""
    import threading
    datas = [gen_logtext(pkg, 6 * MIB + 1000 * i, first_page=100 * i) for i in range(4)]
    errors = []

    def work(i):
        try:
            c = pkg.Context(0)
            for _ in range(3):
                s = c.compress_4mc(datas[i])
                assert c.decompress_4mc(s) == datas[i]
                z = c.compress_4mz(datas[i], 1 + i % 2)
                assert c.decompress_4mz(z) == datas[i]
                blk = datas[i][:300000]
                assert ora.lz4_decompress(c.lz4_compress(blk), len(blk)) == (len(blk), blk)
            c.close()
        except Exception as e:      # noqa: BLE001 -- reported below
            errors.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_randomized_round_trips_all_levels_both_containers(ctx, ora, pkg):
    """Seeded sweep over sizes, inputs and levels: whatever the GPU writes, the oracle (4mc) and the reference's
    ZSTD_decompress (4mz, where oracle/_ref is present) restore it, and so does the GPU reader."""
    import os
    rng = random.Random(20261017)
    R = None
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref4mc.so")
    if os.path.exists(p):
        R = C.CDLL(p)
        R.ZSTD_decompress.restype = C.c_size_t
        R.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]

    def make(kind, n):
        if kind == "zeros":
            return bytes(n)
        if kind == "random":
            return rng.randbytes(n)
        if kind == "period":
            unit = rng.randbytes(rng.randint(1, 300))
            return (unit * (n // len(unit) + 1))[:n]
        if kind == "twosym":
            return bytes(rng.choice(b"ab") for _ in range(min(n, 300000))) * (n // 300000 + 1)
        k = {"text": 0, "json": 1, "mix": 2}[kind]
        pages = (n + 4095) // 4096
        buf = C.create_string_buffer(max(pages, 1) * 4096)
        assert pkg.lib().fourmc_gen_host(k, rng.getrandbits(32), rng.randrange(1 << 20), pages, buf) == 0
        return buf.raw[:n]

    for case in range(28):
        kind = rng.choice(["text", "text", "json", "mix", "mix", "zeros", "random", "period", "twosym"])
        n = rng.choice([0, 1, rng.randrange(2, 70000), rng.randrange(70000, 5 * MIB), rng.randrange(5 * MIB, 13 * MIB), 4 * MIB, 8 * MIB + 1])
        if kind == "twosym":
            n = min(n, 2 * MIB)
        data = make(kind, n)[:n]
        level = rng.randint(1, 4)
        s = ctx.compress_4mc(data, level)
        assert ora.decompress_4mc(s, len(data)) == (len(data), data), (case, kind, n, level)
        assert ctx.decompress_4mc(s) == data, (case, kind, n, level)
        z = ctx.compress_4mz(data, level)
        assert ctx.decompress_4mz(z) == data, (case, kind, n, level)
        if R is not None:
            pos, got = 12, b""
            while True:
                u, c = int.from_bytes(z[pos:pos + 4], "big"), int.from_bytes(z[pos + 4:pos + 8], "big")
                if u == 0:
                    break
                payload = z[pos + 12:pos + 12 + c]
                if c == u:
                    got += payload
                else:
                    back = C.create_string_buffer(u)
                    assert R.ZSTD_decompress(back, u, payload, c) == u, (case, kind, n, level)
                    got += back.raw
                pos += 12 + c
            assert got == data, (case, kind, n, level)


@pytest.mark.parametrize("codec", ["4mc", "4mz"])
def test_sampled_blocks_of_a_large_run_decode_with_the_reference_cli(ctx, ora, pkg, ref_cli, tmp_path, codec):
    """SURVEY.md 8(d) verification at bench size: 16 GiB compressed on the device (4096 blocks, the bench's step),
    64 randomly sampled block records re-assembled into a valid stream (header / EOS / footer regenerated from
    the sampled lengths -- blocks are independent) -> reference `4mc -d` -> compared with the regenerated
    originals; the records' checksums against the CPU XXH32; the whole run decoded again on the device."""
    import numpy as np
    import torch
    nb, blk = 4096, 4 * 1024 * 1024
    n = nb * blk
    zst = codec == "4mz"
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    cap = pkg.lib().fourmc_4mc_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    lens = torch.zeros(nb, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.gen_device(src.data_ptr(), n // 4096)
    (ctx.compress_4mz_device if zst else ctx.compress_device)(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(),
                                                              d_block_lens=lens.data_ptr())
    ctx.sync()
    torch.cuda.synchronize()
    hl = lens.cpu().numpy().astype(np.int64)
    offs = 12 + np.concatenate(([0], np.cumsum(hl)[:-1]))
    assert int(size.item()) == 12 + int(hl.sum()) + 12 + 20 + 4 * nb
    picks = sorted(random.Random(0x5A).sample(range(nb), 64))
    records = [bytes(comp[int(offs[b]):int(offs[b] + hl[b])].cpu().numpy().tobytes()) for b in picks]
    for b, rec in zip(picks, records):
        u, c, ck = (int.from_bytes(rec[i:i + 4], "big") for i in (0, 4, 8))
        assert u == blk and c == len(rec) - 12 and c < u
        assert ck == ora.xxh32(rec[12:])
    # header + EOS + footer for the sampled blocks, built by the library from their lengths
    sl = torch.tensor([len(r) for r in records], dtype=torch.int32, device="cuda")
    head = torch.zeros(12, dtype=torch.uint8, device="cuda")
    tail = torch.zeros(12 + 20 + 4 * len(picks), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    (ctx.build_index_4mz_device if zst else ctx.build_index_device)(sl.data_ptr(), len(picks), head.data_ptr(), tail.data_ptr())
    ctx.sync()
    stream = head.cpu().numpy().tobytes() + b"".join(records) + tail.cpu().numpy().tobytes()
    want = b"".join(gen_logtext(pkg, blk, first_page=b * 1024) for b in picks)
    p, q = tmp_path / "s.bin", tmp_path / "s.out"
    p.write_bytes(stream)
    subprocess.run([ref_cli, "-f", "-q", "-q"] + (["-z"] if zst else []) + ["-d", str(p), str(q)], check=True)
    assert q.read_bytes() == want
    assert (ora.decompress_4mz if zst else ora.decompress_4mc)(stream, len(want)) == (len(want), want)
    # and the whole 16 GiB back on the device
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    (ctx.decompress_4mz_device if zst else ctx.decompress_device)(comp.data_ptr(), int(size.item()), out.data_ptr(), n, res.data_ptr())
    ctx.sync()
    assert res.cpu().tolist() == [n, -1]
    assert torch.equal(out, src)
