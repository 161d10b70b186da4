"""The Zstandard oracle (oracle/zstd_oracle.c: a strict frame decoder and the 4mz reader) pinned to the
reference: the .4mz files the reference CLI wrote (tests/golden, harvested by make_golden.py), the reference's
own ZSTD_decompress verdicts on 1767 valid and mutated frames (tests/golden/zstd_decode.json) and, where
oracle/_ref exists, frames compressed live by the reference at the four 4mz levels."""
import ctypes as C
import random

import pytest

from conftest import golden_bytes, golden_json, gen_logtext

MIB = 1 << 20


@pytest.fixture(scope="module")
def zora(ora):
    class Z:
        frame = staticmethod(ora.zstd_decompress)
        stream = staticmethod(ora.decompress_4mz)
    return Z


def test_reference_4mz_files(zora, pkg):
    text = golden_bytes("logtext_128k.bin")
    for lvl in (1, 2, 3, 4):
        assert zora.stream(golden_bytes(f"logtext_128k.z{lvl}.4mz"), len(text)) == (len(text), text)
    big = gen_logtext(pkg, 1280 * 1024, first_page=64)
    for lvl in (1, 2):
        assert zora.stream(golden_bytes(f"logtext_1280k.z{lvl}.4mz"), len(big)) == (len(big), big)
    assert zora.stream(golden_bytes("empty.4mz"), 0)[0] == 0
    assert zora.stream(golden_bytes("A.4mz"), 1) == (1, b"A")
    n = 4 * MIB + 1
    assert zora.stream(golden_bytes("zeros_4m1.4mz"), n) == (n, bytes(n))
    rnd = golden_bytes("random_70000.bin")
    assert zora.stream(golden_bytes("random_70000.4mz"), len(rnd)) == (len(rnd), rnd)


def test_4mz_stream_errors(zora):
    good = golden_bytes("logtext_128k.z1.4mz")
    n = 128 * 1024
    for at in (9, 100, len(good) - 1):              # header checksum, block payload, footer checksum
        bad = bytearray(good); bad[at] ^= 1
        assert zora.stream(bad, n)[0] < 0, at
    assert zora.stream(good[:20], n)[0] < 0         # cut inside a block header
    assert zora.stream(good.replace(b"4MZ\0", b"4MC\0", 1), n)[0] < 0
    assert zora.stream(good, n - 1)[0] < 0          # output too small


def test_reference_decode_verdicts(zora, ora):
    """Every frame the reference decodes is decoded to the same bytes, or (for the handful of malformed frames the
    reference tolerates) rejected; nothing the reference rejects is accepted."""
    same = strict = 0
    for z in golden_json("zstd_decode.json"):
        src = bytes.fromhex(z["hex"])
        for cap, ret, xxh in z["runs"]:
            r, out = zora.frame(src, cap)
            if ret < 0:
                assert r < 0, (z["hex"][:64], cap, ret, r)
            elif r >= 0:
                assert r == ret and ora.xxh32(out) == xxh, (z["hex"][:64], cap, ret, r)
                same += 1
            else:
                strict += 1
    assert same > 500 and strict <= 4


def test_live_reference_frames(zora, ref, pkg):
    ref.ZSTD_compress.restype = C.c_size_t
    ref.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    rng = random.Random(0x4D5A)
    text = gen_logtext(pkg, 2 * MIB)
    cases = [b"", b"x", bytes(300000), text, text[:70001], bytes(rng.randrange(256) for _ in range(5000)),
             bytes(rng.choice(b"ab") for _ in range(200000)), (text[:1000] * 900)]
    for data in cases:
        for lvl in (1, 3, 6, 12):
            cap = len(data) + len(data) // 128 + 1024
            out = C.create_string_buffer(cap)
            c = ref.ZSTD_compress(out, cap, data, len(data), lvl)
            assert c < cap
            assert zora.frame(out.raw[:c], len(data)) == (len(data), data), (len(data), lvl)


def test_mutated_frames_against_the_live_reference(zora, ref, pkg):
    """Seeded mutations (bit flips, byte substitutions, truncations, insertions) of frames the reference wrote:
    the strict oracle never accepts what the reference rejects, and where both accept the bytes agree."""
    ref.ZSTD_compress.restype = C.c_size_t
    ref.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    ref.ZSTD_decompress.restype = C.c_size_t
    ref.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    ref.ZSTD_isError.restype = C.c_uint
    ref.ZSTD_isError.argtypes = [C.c_size_t]
    rng = random.Random(0x0FAC)
    text = gen_logtext(pkg, 300000)
    frames = []
    noisy = b"".join(rng.randbytes(rng.randrange(20, 400)) + text[k * 50:k * 50 + rng.randrange(8, 60)] * rng.randrange(1, 4)
                     for k in range(300))                       # mostly raw literals: a flipped literal byte still decodes
    for data in (text[:3000], text[:70000], text, bytes(rng.choice(b"abcd") for _ in range(50000)), b"q" * 40000, noisy,
                 noisy[:5000]):
        for lvl in (1, 3, 6):
            cap = len(data) + 1024
            out = C.create_string_buffer(cap)
            c = ref.ZSTD_compress(out, cap, data, len(data), lvl)
            assert c < cap
            frames.append((out.raw[:c], len(data)))
    both = strict_only = rejected = 0
    for frame, n in frames:
        for _ in range(120):
            m = bytearray(frame)
            kind = rng.randrange(4)
            at = rng.randrange(len(m))
            if kind == 0:
                m[at] ^= 1 << rng.randrange(8)
            elif kind == 1:
                m[at] = rng.randrange(256)
            elif kind == 2:
                del m[at:]
            else:
                m[at:at] = bytes([rng.randrange(256)])
            m = bytes(m)
            cap = n + 64
            back = C.create_string_buffer(cap)
            r = ref.ZSTD_decompress(back, cap, m, len(m))
            ro, oo = zora.frame(m, cap)
            if ref.ZSTD_isError(r):
                assert ro < 0, (m[:32].hex(), kind, at)
                rejected += 1
            elif ro >= 0:
                assert ro == r and oo == back.raw[:r], (m[:32].hex(), kind, at)
                both += 1
            else:
                strict_only += 1
    print('both', both, 'strict_only', strict_only, 'rejected', rejected)
    assert both > 40 and rejected > 500, (both, strict_only, rejected)
    assert strict_only <= both // 2          # strictness may reject a tolerated frame, but not as a rule
