"""CPU checks of product logic that is host-compilable: the D1 parser (lz4_parse.h, HD code shared
with the kernel), the generator, the encode algorithm's sequential emulation, and the ABI surface."""
import ctypes as C
import os
import random
import re
import sys

import pytest

from conftest import ROOT, build_native, golden_bytes, golden_json, gen_logtext


@pytest.fixture(scope="module")
def shim():
    L = build_native("parse_shim", ["tests/native/parse_shim.cpp"])
    L.parse_shim.restype = C.c_int
    L.parse_shim.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    return L


def test_parser_matches_reference_golden(shim):
    """The product's D1 state machine returns exactly what LZ4_decompress_safe returned."""
    for v in golden_json("lz4_decode.json"):
        src = bytes.fromhex(v["hex"])
        for step in (0, 1, 7, 2048):            # the parser is resumable at any sequence boundary
            r = shim.parse_shim(src, len(src), v["cap"], None, None, None, step)
            assert r == v["ret"], (v["hex"][:40], v["cap"], step, r, v["ret"])


def test_parser_matches_oracle_fuzz(shim, ora, pkg):
    rng = random.Random(3)
    text = gen_logtext(pkg, 200000)
    for it in range(30):
        n = rng.choice([13, 64, 200, 5000, 70000, 200000])
        comp = ora.lz4_compress(text[:n])
        nt = C.c_int()
        assert shim.parse_shim(comp, len(comp), n, C.byref(nt), None, None, 0) == n and nt.value >= 1
        for k in range(60):
            m = bytearray(comp)
            i = rng.randrange(len(m))
            if k % 3 == 0:
                m[i] = rng.getrandbits(8)
            elif k % 3 == 1:
                m = m[:i]
            else:
                m[i:i] = bytes([rng.choice([0, 255, 0xF0, 0x0F])])
            m = bytes(m)
            for cap in (n, n + 50, 4 << 20):
                assert shim.parse_shim(m, len(m), cap, None, None, None, [0, 3, 100][k % 3]) == ora.lz4_decompress(m, cap)[0]


def test_generator_deterministic(pkg):
    a = gen_logtext(pkg, 3 * 4096 + 100)
    b = gen_logtext(pkg, 4096, first_page=2)
    assert a[2 * 4096:3 * 4096] == b
    assert a[:11] == b"1700000002 " or a[:4] == b"1700"
    assert a == golden_bytes("logtext_128k.bin")[:len(a)]


@pytest.fixture(scope="module")
def emul():
    L = build_native("enc_emul", ["tests/native/enc_emul.cpp"])
    L.enc_emul_block.restype = C.c_int
    L.enc_emul_block.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.enc_emul_block_chain.restype = C.c_int
    L.enc_emul_block_chain.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
    return L


@pytest.mark.parametrize("n", [0, 1, 12, 13, 14, 131, 132, 133, 65535, 65536, 65537, 65549, 200000, 4194304 - 1, 4194304])
def test_encode_algorithm_roundtrip(emul, ora, pkg, n):
    """Sequential emulation of lz4_region_kernel + block write: output decodes to the input."""
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    for name, src in (("text", text[:n]), ("zeros", bytes(n)), ("period", (b"abcdefg" * (n // 7 + 1))[:n])):
        dst = C.create_string_buffer(n + n // 255 + 128)
        c = emul.enc_emul_block(src, n, dst, 5)
        assert ora.lz4_decompress(dst.raw[:c], n) == (n, src), (name, n)


@pytest.mark.parametrize("n", [0, 13, 133, 65537, 200000, 4194304])
def test_chain_parse_algorithm_roundtrip(emul, ora, pkg, n):
    """Levels 2..4: sequential emulation of lz4_chain_kernel + the chain variant of
    lz4_region_kernel (exact links over a sliding 64 KiB history, a search with chain swap at every position,
    cost-optimal choices per slice, two walks); the output decodes to the input and the ratio grows with the depth."""
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    kinds = (("text", text[:n]), ("zeros", bytes(n)), ("period", (b"abcdefg" * (n // 7 + 1))[:n]))
    for name, src in kinds:
        sizes = []
        for depth in (4, 32, 128):
            dst = C.create_string_buffer(n + n // 255 + 128)
            c = emul.enc_emul_block_chain(src, n, dst, 4, depth)
            assert ora.lz4_decompress(dst.raw[:c], n) == (n, src), (name, n, depth)
            sizes.append(c)
        if name == "text" and n == 4194304:
            fast = emul.enc_emul_block(src, n, C.create_string_buffer(n + n // 255 + 128), 5)
            assert sizes[2] < sizes[1] < sizes[0] < fast
            assert n / sizes[0] > 2.5 and n / sizes[1] > 2.66 and n / sizes[2] > 2.69


def test_chain_links_chunked_equal_sequential(emul, pkg):
    """lz4_chain_kernel's scheme (chunks with 64 KiB of warm-up, steps of 32 positions, read-back + lane-to-lane repair
    of equal hashes inside a step) gives exactly the sequential links, whatever the chunk size."""
    import random
    emul.enc_emul_chain_links_check.restype = C.c_longlong
    emul.enc_emul_chain_links_check.argtypes = [C.c_char_p, C.c_int, C.c_int]
    text = gen_logtext(pkg, 1 << 20)
    rng = random.Random(3)
    cases = [text, text[:300001], (b"abcdefg" * 60000)[:400003], bytes(200000), rng.randbytes(150000),
             b"".join(bytes([rng.getrandbits(8)]) * rng.choice([1, 2, 3, 33, 70]) for _ in range(20000)), b"abc", b""]
    for data in cases:
        for chunk in (1 << 16, 1 << 18, 1 << 20):
            assert emul.enc_emul_chain_links_check(data, len(data), chunk) == 0, (len(data), chunk)


def test_chain_parse_reaches_the_reference_ratios(emul, ref, pkg):
    """Row a12's bar: levels 3 / 4 (32 / 128 candidates) compress the bench inputs at least as well as the
    reference's LZ4_compress_HC at its levels 4 / 8 (native/4mc.c:248-251 -> native/lz4/lz4hc.c), level 2 better
    than LZ4_compressMC; the JSON input within 0.1 %."""
    ref.LZ4_compress_HC.restype = C.c_int
    ref.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    n = 4 * 1024 * 1024
    for kind, seed, slack in ((0, 0x4D43, 1.0), (1, 0x4D43, 0.999), (2, 0x4D43, 1.0)):
        buf = C.create_string_buffer(n)
        assert pkg.lib().fourmc_gen_host(kind, seed, 0, n // 4096, buf) == 0
        src = buf.raw
        dst = C.create_string_buffer(n + n // 255 + 128)
        for depth, hc in ((32, 4), (128, 8)):
            mine = emul.enc_emul_block_chain(src, n, dst, 4, depth)
            theirs = ref.LZ4_compress_HC(src, dst, n, len(dst), hc)
            assert mine * slack <= theirs, (kind, depth, n / mine, n / theirs)


def test_encode_algorithm_ratio(emul, pkg):
    text = gen_logtext(pkg, 4 * 1024 * 1024)
    dst = C.create_string_buffer(len(text) + len(text) // 255 + 128)
    c = emul.enc_emul_block(text, len(text), dst, 5)
    assert len(text) / c > 2.0          # the reference's LZ4_compress_default gives 2.02 on this input


def test_abi_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "fourmc.h")).read()
    declared = sorted(set(re.findall(r"\b(fourmc_[a-z0-9_]+)\s*\(", hdr)) - {"fourmc_ctx"})
    assert len(declared) >= 20
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device(pkg):
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.FourMcError) as e:
        pkg.Context(0)
    assert e.value.code == pkg.E_CUDA
    assert pkg.lib().fourmc_lz4_compress_bound(4 * 1024 * 1024) == 4210768   # lz4.h:212


def test_entry_points_reject_a_missing_context(pkg):
    """Every call that needs the device takes a context; NULL is an argument error, never a CPU path."""
    L = pkg.lib()
    buf = C.create_string_buffer(64)
    assert L.fourmc_blockstream_compress_host(None, 0, 1, buf, 64, 0, buf, 64) == pkg.E_ARG
    assert L.fourmc_blockstream_decompress_host(None, 0, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_4mc_compress_host(None, 1, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_4mc_decompress_host(None, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_4mz_compress_host(None, 1, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_4mz_decompress_host(None, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_zstd_compress(None, 1, buf, 64, buf, 64) == pkg.E_ARG
    assert L.fourmc_zstd_decompress(None, buf, 64, buf, 64) == pkg.E_ARG
    # sizes that need no device
    assert L.fourmc_blockstream_bound(0, 0, 0) >= 4 and L.fourmc_blockstream_bound(1, 1 << 30, 65536) > 1 << 30
    assert L.fourmc_4mc_bound(0) == 44                                        # header + EOS + footer of an empty file


def test_shard_blocks(pkg):
    for nb in (0, 1, 7, 8, 9, 16384, 4096 + 3):
        for g in (1, 2, 4, 8):
            ranges = [pkg.shard_blocks(nb, g, r) for r in range(g)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


@pytest.fixture(scope="module")
def d1():
    L = build_native("d1_emul", ["tests/native/d1_emul.cpp"])
    for f in (L.d1_emul, L.d1_serial):
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int64), C.c_int, C.c_int]
    return L


def _d1_both(d1, comp, cap, d):
    nw, nc = (len(comp) + 16) // 32 + 2, (len(comp) + 16) // 128 + 2
    out = []
    for f in (d1.d1_emul, d1.d1_serial):
        m = (C.c_uint32 * nw)(); c = (C.c_int64 * nc)()
        r = f(comp, len(comp), cap, d, m, c, nw, nc)
        out.append((r, list(m), list(c)) if r >= 0 else (r, None, None))     # tables only matter for valid blocks
    return out


def test_warp_parallel_parse_equals_serial_walk(d1, ora, pkg):
    """Sequential emulation of lz4_parse_kernel's bulk phase (per-position decode, exit DP, hop chain,
    per-lane re-walk, hand-over to the state machine): same verdict, token bits and chunk positions."""
    rng = random.Random(12)
    text = gen_logtext(pkg, 600000)
    samples = [text, text[:70000], text[:5000], text[:100], bytes(300000), (b"abcdefg" * 90000)[:500000],
               rng.randbytes(5000) + text[:100000] + bytes(5000) + rng.randbytes(300) * 50,
               b"".join(bytes([rng.getrandbits(8)]) * rng.choice([1, 2, 3, 40, 300, 5000, 70000]) for _ in range(400))]
    for src in samples:
        comp = ora.lz4_compress(src)
        for d in (0, 1, 7, 12, 15):
            for cap in (len(src), len(src) + 100, 4 << 20, max(0, len(src) - 1), len(src) // 2):
                a, b = _d1_both(d1, comp, cap, d)
                assert a == b, (len(src), d, cap, a[0], b[0])
        # corrupt streams: same error value
        for k in range(120):
            m = bytearray(comp)
            i = rng.randrange(len(m))
            if k % 4 == 0:
                m[i] = rng.getrandbits(8)
            elif k % 4 == 1:
                m = m[:i + 1]
            elif k % 4 == 2:
                m[i:i] = bytes([rng.choice([0, 255, 0xF0, 0x0F])])
            else:
                j = rng.randrange(len(m) - 1); m[j] = 0xFF; m[j + 1] = 0xFF      # far offsets / long lengths
            m = bytes(m)
            a, b = _d1_both(d1, m, rng.choice([len(src), len(src) + 64, 4 << 20]), rng.choice([0, 3, 12]))
            assert a == b, (len(src), k, a[0], b[0])


def test_d2_word_assembly_rules(ora):
    """The word stage of lz4_copy_kernel (which bytes travel as aligned words of the span, who stores a word two
    sequences share, what is patched in with byte stores), restated lane by lane in tools/d2_words_emul.py, on the
    sequences of a reference-written block and on adversarial batches; the GPU suite checks the kernel itself."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import d2_words_emul as E
    rnd = random.Random(7)
    stream = golden_bytes("logtext_128k.l1.4mc")
    text = golden_bytes("logtext_128k.bin")
    usize, csize = int.from_bytes(stream[12:16], "big"), int.from_bytes(stream[16:20], "big")
    block = stream[24:24 + csize]
    assert ora.lz4_decompress(block, usize) == (usize, text[:usize])
    seqs = E.parse_lz4_block(block)
    assert sum(s[1] + s[2] for s in seqs) == usize
    for s0 in range(0, len(seqs), 32):
        grp = [(lit, ml, off, op, text[op:op + lit]) for _, lit, ml, off, op in seqs[s0:s0 + 32]]
        if sum(g[0] + g[1] for g in grp) > E.SPAN:
            continue
        for dstbase in (0, 5):
            got, op0, total = E.run_batch(grp, text, dstbase, rnd)
            assert got == text[op0:op0 + total], (s0, dstbase)
    for _ in range(1500):
        grp, out = E.adversarial_case(rnd)
        if sum(g[0] + g[1] for g in grp) > E.SPAN:
            continue
        got, op0, total = E.run_batch(grp, out, rnd.randrange(16), rnd)
        assert got == out[op0:op0 + total]
