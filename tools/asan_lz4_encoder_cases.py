"""ASan / UBSan over the LZ4 encoder emulation (tests/native/enc_emul.cpp: the Fast parse and the chain parse of levels
2..4 exactly as lz4_chain_kernel / lz4_region_kernel run them) -- see tools/asan_host.sh."""
import ctypes as C
import importlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("4mc_b200")
E = C.CDLL("/tmp/fourmc_asan/enc_emul_asan.so")
E.enc_emul_block.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
E.enc_emul_block_chain.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
E.enc_emul_chain_links_check.restype = C.c_longlong
E.enc_emul_chain_links_check.argtypes = [C.c_char_p, C.c_int, C.c_int]
O = C.CDLL("/tmp/fourmc_asan/liboracle_asan.so")
O.fmo_lz4_decompress_safe.restype = C.c_int
O.fmo_lz4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
rng = random.Random(5)
n = 600000
buf = C.create_string_buffer((n + 4095) // 4096 * 4096)
assert pkg.lib().fourmc_gen_host(0, 0x4D43, 0, (n + 4095) // 4096, buf) == 0
text = buf.raw[:n]
cases = [b"", b"A", text[:13], text[:4095], text[:65537], text[:98305], text, bytes(200000), rng.randbytes(70000),
         (b"abcdefg" * 30000)[:200003], bytes(rng.choice(b"ab") for _ in range(100000)),
         b"".join(bytes([rng.randrange(256)]) * rng.randint(1, 3000) for _ in range(150))]
runs = 0
for data in cases:
    m = len(data)
    src = (C.c_char * max(m, 1)).from_buffer_copy(data or b"\0")          # exact-size input: reads past its end are caught
    for depth in (0, 4, 32, 128):
        cap = m + m // 255 + 128
        dst = C.create_string_buffer(cap)
        c = E.enc_emul_block(C.cast(src, C.c_char_p), m, dst, 5) if depth == 0 else E.enc_emul_block_chain(C.cast(src, C.c_char_p), m, dst, 4, depth)
        back = C.create_string_buffer(m + 1)
        assert O.fmo_lz4_decompress_safe(dst.raw[:c], back, c, m) == m and back.raw[:m] == data, (m, depth)
        runs += 1
    for chunk in (1 << 16, 1 << 18):
        assert E.enc_emul_chain_links_check(C.cast(src, C.c_char_p), m, chunk) == 0
print("ran", runs, "LZ4 encoder emulations clean")
