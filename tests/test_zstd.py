"""4mz / zstd frame decoding (SURVEY.md rows a8 reader, a9).

CPU half: the product's frame decoder (4mc_b200/csrc/zstd_decode.h) compiled for the host by
tests/native/zstd_shim.cpp, against vectors produced by the reference's own ZSTD_compress /
ZSTD_decompress and its CLI (tests/golden/make_golden.py), and -- where oracle/_ref exists -- against
ZSTD_decompress live on mutated frames.  GPU half: the same vectors through the C-ABI.
Bit-exact: decoded bytes, decoded sizes and accept / reject must equal the reference's."""
import ctypes as C
import os
import random
import subprocess

import pytest

from conftest import ROOT, golden_bytes, golden_json, gen_logtext, build_native

MIB = 1024 * 1024


@pytest.fixture(scope="module")
def zshim():
    Z = build_native("zstd_shim", ["tests/native/zstd_shim.cpp"], deps=["4mc_b200/csrc/zstd_decode.h"])
    Z.zstd_shim_decompress.restype = C.c_longlong
    Z.zstd_shim_decompress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong]

    def dec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = Z.zstd_shim_decompress(out, cap, src, len(src))
        return int(r), out.raw[:max(r, 0)]
    return dec


def walk_4mz(stream):
    """(usize, csize, payload) of every block of one 4mz stream."""
    assert stream[:12] == bytes.fromhex("344d5a00 00000001 289a1c9a")
    pos, out = 12, []
    while True:
        u, c = int.from_bytes(stream[pos:pos + 4], "big"), int.from_bytes(stream[pos + 4:pos + 8], "big")
        if u == 0:
            return out
        out.append((u, c, stream[pos + 12:pos + 12 + c]))
        pos += 12 + c


def expected_of(pkg, name):
    if name.startswith("logtext_128k"):
        return golden_bytes("logtext_128k.bin")
    if name.startswith("logtext_1280k"):
        return gen_logtext(pkg, 1280 * 1024, first_page=64)
    return {"empty.4mz": b"", "A.4mz": b"A", "zeros_4m1.4mz": bytes(4 * MIB + 1), "random_70000.4mz": golden_bytes("random_70000.bin")}[name]


FILES_4MZ = ["empty.4mz", "A.4mz", "zeros_4m1.4mz", "random_70000.4mz", "logtext_128k.z1.4mz", "logtext_128k.z2.4mz",
             "logtext_128k.z3.4mz", "logtext_128k.z4.4mz", "logtext_1280k.z1.4mz", "logtext_1280k.z2.4mz"]


# ---- CPU: the decoder source against the reference's vectors ------------------------------------

def test_host_build_matches_reference_vectors(zshim, ora):
    n_ok = n_bad = 0
    for z in golden_json("zstd_decode.json"):
        src = bytes.fromhex(z["hex"])
        for cap, ret, xxh in z["runs"]:
            r, out = zshim(src, cap)
            if ret < 0:
                assert r < 0, (z["hex"][:64], cap, ret, r)
                n_bad += 1
            else:
                assert r == ret and ora.xxh32(out) == xxh, (z["hex"][:64], cap, ret, r)
                n_ok += 1
    assert n_ok > 400 and n_bad > 1000


@pytest.mark.parametrize("name", FILES_4MZ)
def test_host_build_decodes_reference_4mz_blocks(zshim, ora, pkg, name):
    want = expected_of(pkg, name)
    got = b""
    for u, c, payload in walk_4mz(golden_bytes(name)):
        if c == u:
            got += payload                                       # stored block (native/4mc.c:797-803)
        else:
            r, out = zshim(payload, u)
            assert r == u
            got += out
    assert got == want


def test_host_build_agrees_with_reference_on_mutated_frames(zshim, ref, pkg):
    """Accept / reject and the decoded bytes on damaged frames: live against ZSTD_decompress."""
    ref.ZSTD_compress.restype = C.c_size_t
    ref.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    ref.ZSTD_decompress.restype = C.c_size_t
    ref.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    ref.ZSTD_isError.restype = C.c_uint
    ref.ZSTD_isError.argtypes = [C.c_size_t]
    ref.ZSTD_compressBound.restype = C.c_size_t
    ref.ZSTD_compressBound.argtypes = [C.c_size_t]
    rng = random.Random(4)
    text = gen_logtext(pkg, 300000)
    skew = bytes(min(255, int(rng.expovariate(0.05))) for _ in range(60000))
    both_ok = 0
    for src in (text[:70000], text[:3000], text[:200000], skew, skew[:900]):
        for lvl in (1, 3, 19):
            cap = ref.ZSTD_compressBound(len(src))
            cb = C.create_string_buffer(cap)
            csz = ref.ZSTD_compress(cb, cap, src, len(src), lvl)
            comp = cb.raw[:csz]
            assert zshim(comp, len(src)) == (len(src), src)
            for _ in range(250):
                m = bytearray(comp)
                for _k in range(rng.choice((1, 1, 1, 2, 3))):
                    m[rng.randrange(len(m))] = rng.getrandbits(8)
                if rng.random() < 0.05:
                    m = m[:rng.randrange(1, len(m))]
                m = bytes(m)
                o2 = C.create_string_buffer(len(src) + 64)
                b = ref.ZSTD_decompress(o2, len(src), m, len(m))
                a, out = zshim(m, len(src))
                if ref.ZSTD_isError(b):
                    assert a < 0
                else:
                    assert a == b and out == o2.raw[:b]
                    both_ok += 1
    assert both_ok > 500


# ---- GPU: through the C-ABI ----------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", FILES_4MZ)
def test_gpu_decodes_reference_4mz(ctx, pkg, name):
    assert ctx.decompress_4mz(golden_bytes(name)) == expected_of(pkg, name)
    assert pkg.FourMzCodec(ctx).decompress(golden_bytes(name)) == expected_of(pkg, name)


@pytest.mark.gpu
def test_gpu_zstd_decompress_matches_reference_vectors(ctx, ora):
    n_ok = n_bad = 0
    for z in golden_json("zstd_decode.json"):
        src = bytes.fromhex(z["hex"])
        for cap, ret, xxh in z["runs"]:
            r, out = ctx.zstd_decompress(src, cap)
            if ret < 0:
                assert r < 0, (z["hex"][:64], cap, ret, r)
                n_bad += 1
            else:
                assert r == ret and ora.xxh32(out) == xxh, (z["hex"][:64], cap, ret, r)
                n_ok += 1
    assert n_ok > 400 and n_bad > 1000


@pytest.mark.gpu
def test_gpu_4mz_container_errors(ctx, pkg):
    good = golden_bytes("logtext_128k.z1.4mz")
    n = 128 * 1024
    assert ctx.decompress_4mz_rc(good, n) == n
    assert ctx.decompress_4mz_rc(good, n - 1) == pkg.E_OUTPUT
    assert ctx.decompress_4mz_rc(golden_bytes("logtext_128k.l1.4mc"), n) == pkg.E_CONTENT    # "not a 4mc file", native/4mc.c:888
    bad = bytearray(good); bad[100] ^= 0x40                      # payload byte: block checksum mismatch (native/4mc.c:790)
    assert ctx.decompress_4mz_rc(bytes(bad), n) == pkg.E_CONTENT
    assert ctx.decompress_4mz_rc(good[:len(good) // 2], n) == pkg.E_INPUT
    two = golden_bytes("A.4mz") + good                           # concatenated streams (native/4mc.c:908-912)
    assert ctx.decompress_4mz(two) == b"A" + golden_bytes("logtext_128k.bin")
    assert pkg.FourMzCodec(ctx).decompress(pkg.FourMzCodec(ctx).compress(b"abc" * 1000)) == b"abc" * 1000


def assemble_4mz(ora, blocks):
    """One 4mz stream from (usize, csize, payload) records."""
    body, deltas, off, prev = b"", [], 12, 0
    for u, c, payload in blocks:
        body += u.to_bytes(4, "big") + c.to_bytes(4, "big") + ora.xxh32(payload).to_bytes(4, "big") + payload
        deltas.append(off - prev)
        prev = off
        off += 12 + c
    n = len(blocks)
    foot = (20 + 4 * n).to_bytes(4, "big") + (1).to_bytes(4, "big") + b"".join(d.to_bytes(4, "big") for d in deltas)
    foot += (20 + 4 * n).to_bytes(4, "big") + bytes.fromhex("344d5a00")
    return bytes.fromhex("344d5a00 00000001 289a1c9a") + body + bytes(12) + foot + ora.xxh32(foot).to_bytes(4, "big")


@pytest.mark.gpu
def test_gpu_4mz_device_call_many_blocks(ctx, pkg, ora):
    """The device-resident call takes ONE stream (found through its footer index): a multi-block
    stream is assembled from the blocks the reference CLI wrote into the committed fixtures."""
    import torch
    one = walk_4mz(golden_bytes("logtext_1280k.z1.4mz"))
    zeros = walk_4mz(golden_bytes("zeros_4m1.4mz"))
    rnd = walk_4mz(golden_bytes("random_70000.4mz"))
    blocks = one * 3 + zeros[:1] + rnd + one * 2
    stream = assemble_4mz(ora, blocks)
    text = gen_logtext(pkg, 1280 * 1024, first_page=64)
    want = text * 3 + bytes(4 * MIB) + golden_bytes("random_70000.bin") + text * 2
    assert ctx.decompress_4mz(stream) == want
    d_in = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
    d_out = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.decompress_4mz_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), len(want), d_res.data_ptr())
    ctx.sync()
    assert d_res.tolist() == [len(want), -1]
    assert bytes(d_out[:len(want)].cpu().numpy()) == want


@pytest.mark.gpu
def test_gpu_decodes_live_reference_frames(ctx, pkg, ora, ref):
    """Full 4 MiB blocks compressed by the reference's ZSTD_compress here and now, at the four 4mz levels
    (1 / 3 / 6 / 12: native/4mc.c:415-425): 32 sub-blocks per frame with treeless literals and repeat-mode
    tables, decoded by the batch path and by the per-block call."""
    ref.ZSTD_compress.restype = C.c_size_t
    ref.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    data = gen_logtext(pkg, 2 * 4 * MIB + 123457, first_page=999)
    for lvl in (1, 3, 6, 12):
        blocks = []
        for o in range(0, len(data), 4 * MIB):
            src = data[o:o + 4 * MIB]
            cb = C.create_string_buffer(len(src))
            c = ref.ZSTD_compress(cb, len(src) - 1, src, len(src), lvl)
            assert 0 < c < len(src)
            blocks.append((len(src), c, cb.raw[:c]))
        assert ctx.decompress_4mz(assemble_4mz(ora, blocks)) == data, lvl
        r, out = ctx.zstd_decompress(blocks[0][2], 4 * MIB)
        assert r == 4 * MIB and out == data[:4 * MIB], lvl
    # a frame of more than 32 blocks through the per-block call: the tables in force ("repeat" modes, treeless
    # literals) are carried from one chunk of 32 blocks to the next inside the lane-parallel kernel
    big = gen_logtext(pkg, 7 * MIB + 99, first_page=4242)
    for lvl in (1, 3):
        cb = C.create_string_buffer(len(big))
        c = ref.ZSTD_compress(cb, len(big), big, len(big), lvl)
        assert 0 < c < 8 * MIB
        assert ctx.zstd_decompress(cb.raw[:c], len(big)) == (len(big), big), lvl
    # more sequences in one block than the fast path holds (two-symbol noise: ~6 bytes per sequence): the
    # frame is handed to the serial decoder, same bytes
    rng = random.Random(12)
    dense = bytes(rng.choice(b"ab") for _ in range(2 * MIB))
    cb = C.create_string_buffer(len(dense))
    c = ref.ZSTD_compress(cb, len(dense), dense, len(dense), 1)
    assert ctx.zstd_decompress(cb.raw[:c], len(dense)) == (len(dense), dense)
    assert ctx.decompress_4mz(assemble_4mz(ora, [(len(dense), c, cb.raw[:c])])) == dense


@pytest.mark.gpu
def test_cli_decodes_reference_4mz(pkg, tmp_path):
    cli = os.path.join(ROOT, "4mc_b200", "host", "4mc")
    src = os.path.join(ROOT, "tests", "golden", "logtext_1280k.z2.4mz")
    out = tmp_path / "o.bin"
    subprocess.run([cli, "-f", "-q", "-z", "-d", src, str(out)], check=True)
    assert out.read_bytes() == gen_logtext(pkg, 1280 * 1024, first_page=64)
