#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference built by oracle/Makefile (oracle/_ref/).

Run in the build container (needs /root/reference through oracle/_ref):   python tests/golden/make_golden.py
Everything written here is an OUTPUT OF THE REFERENCE ITSELF (its CLI `4mc` and its
LZ4_compress_default / LZ4_decompress_safe / XXH32 symbols); the tests compare the oracle
restatement and the CUDA path against these files, so they run where /root/reference does not exist.
"""
import ctypes as C
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "4mc")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")
sys.path.insert(0, ROOT)
import importlib
pkg = importlib.import_module("4mc_b200")


def gen_logtext(nbytes, seed=0x4D43, first_page=0):
    pages = (nbytes + 4095) // 4096
    buf = C.create_string_buffer(pages * 4096)
    assert pkg.lib().fourmc_gen_host(0, seed, first_page, pages, buf) == 0
    return buf.raw[:nbytes]


def cli(args, inp, out):
    subprocess.run([REF_CLI, "-f", "-q", "-q"] + args + [inp, out], check=True)


def main():
    R = C.CDLL(REF_LIB)
    R.XXH32.restype = C.c_uint32
    R.XXH32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    R.LZ4_compress_default.restype = C.c_int
    R.LZ4_decompress_safe.restype = C.c_int
    rng = random.Random(20261017)
    tmp = os.path.join(HERE, "_tmp.bin")

    def w(name, data):
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)

    # ---- container fixtures from the reference CLI
    text = gen_logtext(128 * 1024)
    w("logtext_128k.bin", text)
    for lvl in (1, 2, 3, 4):
        cli([f"-{lvl}"], os.path.join(HERE, "logtext_128k.bin"), os.path.join(HERE, f"logtext_128k.l{lvl}.4mc"))
    for name, data in (("empty", b""), ("A", b"A"), ("zeros_4m1", bytes(4 * 1024 * 1024 + 1))):
        w("_tmp.bin", data)
        cli(["-1"], tmp, os.path.join(HERE, f"{name}.4mc"))
    rnd = bytes(rng.getrandbits(8) for _ in range(70000))
    w("random_70000.bin", rnd)
    cli(["-1"], os.path.join(HERE, "random_70000.bin"), os.path.join(HERE, "random_70000.4mc"))
    # two concatenated streams (native/4mc.c:908-912)
    w("two_streams.4mc", open(os.path.join(HERE, "A.4mc"), "rb").read() + open(os.path.join(HERE, "logtext_128k.l1.4mc"), "rb").read())
    os.remove(tmp)

    # ---- XXH32 known answers
    xs = []
    lcg = bytearray(1 << 20)
    s = 12345
    for i in range(len(lcg)):
        s = (s * 1103515245 + 12345) & 0xFFFFFFFF
        lcg[i] = (s >> 16) & 0xFF
    cases = [b"", b"a", b"abc", b"Nobody inspects the spammish repetition", b"0123456789abcdef", b"0123456789abcdefX"]
    for c in cases:
        for seed in (0, 1, 0x9E3779B1):
            xs.append({"hex": c.hex(), "seed": seed, "xxh32": R.XXH32(c, len(c), seed)})
    for ln in list(range(0, 70)) + [255, 256, 257, 4095, 4096, 4097, 65535, 65536, 65537, (1 << 20) - 3, 1 << 20]:
        for off in (0, 1, 2, 3, 5, 17):
            if off + ln > len(lcg):
                continue
            b = bytes(lcg[off:off + ln])
            xs.append({"lcg_off": off, "lcg_len": ln, "seed": 0, "xxh32": R.XXH32(b, ln, 0)})
    json.dump(xs, open(os.path.join(HERE, "xxh32.json"), "w"), indent=0)

    # ---- LZ4 decode known answers: valid blocks from LZ4_compress_default + adversarial + mutated
    def ref_dec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = R.LZ4_decompress_safe(src, out, len(src), cap)
        o = out.raw[:max(r, 0)]
        return r, R.XXH32(o, len(o), 0)

    def ref_comp(src):
        cb = C.create_string_buffer(len(src) + len(src) // 255 + 32)
        c = R.LZ4_compress_default(src, cb, len(src), len(cb))
        return cb.raw[:c]

    ks = []
    adv = [("1041 0000 C0" , b"0123456789ab", 17), ("1041 0900 C0", b"0123456789ab", 17), ("1041 0100 40", b"0123", 9),
           ("1041 0100 C0", b"0123456789ab", 64), ("1041 0100 C0", b"0123456789ab", 16), ("00", b"", 0), ("00", b"", 5), ("", b"", 5)]
    for hx, tail, cap in adv:
        src = bytes.fromhex(hx) + tail
        r, x = ref_dec(src, cap)
        ks.append({"hex": src.hex(), "cap": cap, "ret": r, "out_xxh32": x})
    samples = [text[:n] for n in (0, 1, 5, 12, 13, 14, 20, 63, 64, 65, 100, 300, 5000, 70000)] + [rnd[:3000], bytes(5000), text[1000:1300] * 40]
    for sm in samples:
        comp = ref_comp(sm)
        n = len(sm)
        for cap in sorted({n, n + 1, n + 100, max(n - 1, 0), 4 << 20}):
            r, x = ref_dec(comp, cap)
            ks.append({"hex": comp.hex(), "cap": cap, "ret": r, "out_xxh32": x})
        if n > 6000 or not comp:
            continue
        for _ in range(60):
            m = bytearray(comp)
            kind = rng.randrange(4)
            if kind == 0:
                m[rng.randrange(len(m))] = rng.getrandbits(8)
            elif kind == 1:
                m = m[:rng.randrange(len(m) + 1)]
            elif kind == 2:
                m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
            else:
                i = rng.randrange(len(m))
                m[i:i] = bytes([rng.choice([0, 255, 0xF0, 0x0F, 0xFF])])
            m = bytes(m)
            cap = rng.choice([n, n + 37, 4 << 20])
            r, x = ref_dec(m, cap)
            ks.append({"hex": m.hex(), "cap": cap, "ret": r, "out_xxh32": x})
    json.dump(ks, open(os.path.join(HERE, "lz4_decode.json"), "w"), indent=0)
    print("golden:", len(xs), "xxh32 vectors,", len(ks), "lz4 decode vectors")

    # ---- 4mz containers from the reference CLI (-z: zstd levels 1 / 3 / 6 / 12, native/4mc.c:389-553)
    for lvl in (1, 2, 3, 4):
        cli(["-z", f"-{lvl}"], os.path.join(HERE, "logtext_128k.bin"), os.path.join(HERE, f"logtext_128k.z{lvl}.4mz"))
    for name, data in (("empty", b""), ("A", b"A"), ("zeros_4m1", bytes(4 * 1024 * 1024 + 1))):
        w("_tmp.bin", data)
        cli(["-z", "-1"], tmp, os.path.join(HERE, f"{name}.4mz"))
    cli(["-z", "-1"], os.path.join(HERE, "random_70000.bin"), os.path.join(HERE, "random_70000.4mz"))
    # 1.25 MiB of generator text (regenerated by the tests, not stored): ten 128 KiB zstd blocks per
    # frame, so treeless literals and repeat-mode sequence tables occur
    w("_tmp.bin", gen_logtext(1280 * 1024, first_page=64))
    for lvl in (1, 2):
        cli(["-z", f"-{lvl}"], tmp, os.path.join(HERE, f"logtext_1280k.z{lvl}.4mz"))
    os.remove(tmp)

    # ---- zstd frame known answers: ZSTD_compress at several levels, decoded by ZSTD_decompress with
    # several capacities, plus mutated frames (accept / reject and the decoded bytes must match)
    R.ZSTD_compress.restype = C.c_size_t
    R.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    R.ZSTD_decompress.restype = C.c_size_t
    R.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    R.ZSTD_isError.restype = C.c_uint
    R.ZSTD_isError.argtypes = [C.c_size_t]
    R.ZSTD_compressBound.restype = C.c_size_t
    R.ZSTD_compressBound.argtypes = [C.c_size_t]

    def zcomp(src, lvl):
        cap = R.ZSTD_compressBound(len(src))
        out = C.create_string_buffer(cap)
        n = R.ZSTD_compress(out, cap, src, len(src), lvl)
        return out.raw[:n]

    def zdec(src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = R.ZSTD_decompress(out, cap, src, len(src))
        if R.ZSTD_isError(r):
            return int(r) - (1 << 64), 0
        return int(r), R.XXH32(out.raw[:r], r, 0)

    skew = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(40000))     # many symbols, Huffman weights via FSE
    zsamples = [text[:n] for n in (0, 1, 2, 9, 50, 300, 1000, 3000, 20000)] + \
               [rnd[:2000], bytes(3000), bytes(200000), b"ab" * 5000, text[1000:1300] * 40, skew[:20000], skew[:700],
                text[:15000] + rnd[:5000] + text[40000:50000]]
    zs = []
    for sm in zsamples:
        for lvl in (1, 3, 6, 12, 19):
            comp = zcomp(sm, lvl)
            n = len(sm)
            zs.append({"hex": comp.hex(), "runs": [[cap, *zdec(comp, cap)] for cap in sorted({n, n + 1, max(n - 1, 0), n + (1 << 20)})]})
            if len(comp) > 6000 or lvl not in (1, 3, 19):
                continue
            for _ in range(40):
                m = bytearray(comp)
                for _k in range(rng.choice((1, 1, 2))):
                    kind = rng.randrange(4)
                    if kind == 0:
                        m[rng.randrange(len(m))] = rng.getrandbits(8)
                    elif kind == 1 and len(m) > 1:
                        m = m[:rng.randrange(1, len(m) + 1)]
                    elif kind == 2:
                        m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
                    else:
                        i = rng.randrange(len(m))
                        m[i:i] = bytes([rng.getrandbits(8)])
                m = bytes(m)
                cap = rng.choice([n, n + 37, n + (1 << 20)])
                zs.append({"hex": m.hex(), "runs": [[cap, *zdec(m, cap)]]})
    # two frames back to back and a skippable frame in front (zstd_decompress.c:989-1100)
    two = zcomp(text[:5000], 3) + zcomp(text[5000:9000], 1)
    skip = bytes.fromhex("502a4d18") + (7).to_bytes(4, "little") + b"skipped" + zcomp(text[:777], 1)
    for comp, n in ((two, 9000), (skip, 777)):
        zs.append({"hex": comp.hex(), "runs": [[cap, *zdec(comp, cap)] for cap in (n, n - 1, n + 100)]})
    json.dump(zs, open(os.path.join(HERE, "zstd_decode.json"), "w"), indent=0)
    runs = [r for z in zs for r in z["runs"]]
    print("golden:", len(zs), "zstd frames,", len(runs), "decode runs,", sum(1 for r in runs if r[1] < 0), "of them rejected")


if __name__ == "__main__":
    main()
