"""The oracle (oracle/fourmc_oracle.c) against the reference's own outputs: committed golden vectors
(tests/golden, harvested from the reference build) and, where oracle/_ref exists, the reference live."""
import ctypes as C
import random

import pytest

from conftest import golden_bytes, golden_json, gen_logtext


def test_xxh32_golden(ora, lcg):
    for v in golden_json("xxh32.json"):
        data = bytes.fromhex(v["hex"]) if "hex" in v else lcg[v["lcg_off"]:v["lcg_off"] + v["lcg_len"]]
        assert ora.xxh32(data, v["seed"]) == v["xxh32"], v


def test_xxh32_survey_vectors(ora, lcg):
    # SURVEY.md 8c, measured on the reference build
    assert ora.xxh32(b"") == 0x02CC5D05
    assert ora.xxh32(b"abc") == 0x32D153FF
    assert ora.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F
    assert ora.xxh32(b"a", 1) == 0xF514706F
    assert ora.xxh32(lcg) == 0x0152B317
    assert ora.xxh32(lcg[:(1 << 20) - 3]) == 0x8E213116
    assert ora.xxh32(bytes.fromhex("344d430000000001")) == 0xA4B73443
    assert ora.xxh32(bytes.fromhex("344d5a0000000001")) == 0x289A1C9A


def test_lz4_decode_golden(ora):
    for v in golden_json("lz4_decode.json"):
        r, out = ora.lz4_decompress(bytes.fromhex(v["hex"]), v["cap"])
        assert r == v["ret"], (v["hex"][:40], v["cap"], r, v["ret"])
        assert ora.xxh32(out) == v["out_xxh32"]


def test_container_golden_decode(ora):
    text = golden_bytes("logtext_128k.bin")
    for lvl in (1, 2, 3, 4):
        r, out = ora.decompress_4mc(golden_bytes(f"logtext_128k.l{lvl}.4mc"), len(text))
        assert r == len(text) and out == text
    assert ora.decompress_4mc(golden_bytes("empty.4mc"), 0)[0] == 0
    assert ora.decompress_4mc(golden_bytes("A.4mc"), 1) == (1, b"A")
    n = 4 * 1024 * 1024 + 1
    r, out = ora.decompress_4mc(golden_bytes("zeros_4m1.4mc"), n)
    assert r == n and out == bytes(n)
    rnd = golden_bytes("random_70000.bin")
    assert ora.decompress_4mc(golden_bytes("random_70000.4mc"), len(rnd)) == (len(rnd), rnd)
    r, out = ora.decompress_4mc(golden_bytes("two_streams.4mc"), 1 + len(text))
    assert out == b"A" + text


def test_short_block_and_empty_stream_follow_the_reference(ora, ref_cli, tmp_path):
    """Two corners of the serial reader: a block that decodes to fewer bytes than it announces is
    written as decoded (native/4mc.c:661-666), and a stream that decodes to nothing ends the loop over
    concatenated streams (:909-913).  The oracle against the reference CLI itself."""
    import subprocess
    from conftest import short_block_stream
    stream, expect = short_block_stream(ora)
    cases = [(stream, expect), (golden_bytes("empty.4mc") + golden_bytes("A.4mc"), b""),
             (golden_bytes("A.4mc") + golden_bytes("empty.4mc") + golden_bytes("A.4mc"), b"A")]
    for i, (data, want) in enumerate(cases):
        assert ora.decompress_4mc(data, len(want) + 4096) == (len(want), want)
        src, dst = tmp_path / f"c{i}.4mc", tmp_path / f"c{i}.out"
        src.write_bytes(data)
        assert subprocess.run([ref_cli, "-d", "-f", "-q", str(src), str(dst)]).returncode == 0
        assert dst.read_bytes() == want


def test_container_golden_layout(ora):
    # SURVEY.md 8c byte layouts
    assert golden_bytes("empty.4mc").hex() == ("344d430000000001a4b73443" "000000000000000000000000"
                                               "000000140000000100000014344d4300849b8d65")
    a = golden_bytes("A.4mc")
    assert a[12:24].hex() == "000000010000000110659a4d" and a[24:25] == b"A"
    assert a[-24:].hex() == "00000018000000010000000c00000018344d43004a23827e"
    # writer restatement reproduces the stored-only and empty files byte for byte
    assert ora.compress_4mc(b"") == golden_bytes("empty.4mc")
    assert ora.compress_4mc(b"A") == a
    rnd = golden_bytes("random_70000.bin")
    assert ora.compress_4mc(rnd) == golden_bytes("random_70000.4mc")


def test_container_errors(ora):
    good = golden_bytes("logtext_128k.l1.4mc")
    n = 128 * 1024
    bad = bytearray(good); bad[100] ^= 1
    assert ora.decompress_4mc(bad, n)[0] == -4          # invalid block checksum -> exit 4
    bad = bytearray(good); bad[9] ^= 1
    assert ora.decompress_4mc(bad, n)[0] == -4          # header checksum
    bad = bytearray(good); bad[7] = 2
    assert ora.decompress_4mc(bad, n)[0] == -4          # version
    bad = bytearray(good); bad[-1] ^= 1
    assert ora.decompress_4mc(bad, n)[0] == -4          # footer checksum
    assert ora.decompress_4mc(good[:20], n)[0] == -2    # truncated inside a block header -> exit 2
    assert ora.decompress_4mc(golden_bytes("logtext_128k.l1.4mc").replace(b"4MC\0", b"4MZ\0", 1), n)[0] == -4


def test_read_index(oracle, ora):
    f = golden_bytes("zeros_4m1.4mc")
    offs = (C.c_int64 * 8)()
    nb = oracle.fmo_4mc_read_index(f, len(f), 0x344D4300, offs, 8)
    assert nb == 2 and offs[0] == 12 and offs[1] == 12 + 0x4057       # second delta 0x4057
    assert oracle.fmo_4mc_read_index(golden_bytes("empty.4mc"), 44, 0x344D4300, offs, 8) == 0
    bad = bytearray(f); bad[-1] ^= 1
    assert oracle.fmo_4mc_read_index(bytes(bad), len(bad), 0x344D4300, offs, 8) == -4


def test_oracle_compressor_roundtrip(ora, pkg):
    rng = random.Random(5)
    text = gen_logtext(pkg, 300000)
    for n in (0, 1, 12, 13, 64, 1000, 65536, 300000):
        for src in (text[:n], bytes(n), bytes(rng.getrandbits(8) for _ in range(min(n, 4000)))):
            c = ora.lz4_compress(src)
            assert ora.lz4_decompress(c, len(src)) == (len(src), src)


# ---- live against the reference, when it was built here ----

def test_live_lz4_decode_fuzz(ora, ref):
    rng = random.Random(11)

    def ref_dec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = ref.LZ4_decompress_safe(src, out, len(src), cap)
        return r, out.raw[:max(r, 0)]

    words = [("w%d" % rng.randrange(10 ** rng.randrange(1, 6))).encode() for _ in range(300)]
    for it in range(40):
        n = rng.choice([0, 1, 5, 12, 13, 20, 63, 64, 65, 100, 300, 5000])
        src = bytearray()
        while len(src) < n:
            src += rng.choice(words) + b" "
        src = bytes(src[:n])
        cb = C.create_string_buffer(n + n // 255 + 32)
        c = ref.LZ4_compress_default(src, cb, n, len(cb))
        comp = cb.raw[:c]
        for k in range(80):
            m = bytearray(comp)
            if not m:
                break
            kind = rng.randrange(4)
            if kind == 0:
                m[rng.randrange(len(m))] = rng.getrandbits(8)
            elif kind == 1:
                m = m[:rng.randrange(len(m) + 1)]
            elif kind == 2:
                m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
            else:
                i = rng.randrange(len(m)); m[i:i] = bytes([rng.choice([0, 255, 0xF0, 0x0F, 0xFF])])
            m = bytes(m)
            for cap in (n, n + 37, 4 << 20):
                assert ora.lz4_decompress(m, cap) == ref_dec(m, cap)


def test_live_cli_decodes_oracle_stream(ora, ref_cli, pkg, tmp_path):
    import subprocess
    data = gen_logtext(pkg, 5 * 1024 * 1024 + 123)
    p = tmp_path / "o.4mc"
    p.write_bytes(ora.compress_4mc(data))
    out = tmp_path / "o.bin"
    subprocess.run([ref_cli, "-f", "-q", "-q", "-d", str(p), str(out)], check=True)
    assert out.read_bytes() == data
