"""ASan / UBSan over the zstd encoder source the kernels run (zstd_encode.h through tests/native/zenc_emul.cpp, one
loop iteration per thread id) and over the block-stream framing (blockstream.h) -- see tools/asan_host.sh."""
import ctypes as C
import importlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("4mc_b200")
Z = C.CDLL("/tmp/fourmc_asan/zenc_emul_asan.so")
Z.zenc_emul_compress.restype = C.c_longlong
Z.zenc_emul_compress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong, C.c_int]
D = C.CDLL("/tmp/fourmc_asan/zstd_shim_asan.so")
D.zstd_shim_decompress.restype = C.c_longlong
D.zstd_shim_decompress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong]
rng = random.Random(21)
n = 700000
buf = C.create_string_buffer((n + 4095) // 4096 * 4096)
assert pkg.lib().fourmc_gen_host(0, 0x4D43, 0, (n + 4095) // 4096, buf) == 0
text = buf.raw[:n]
assert pkg.lib().fourmc_gen_host(1, 0x4D5A, 0, (n + 4095) // 4096, buf) == 0
js = buf.raw[:n]
cases = [b"", b"A", text[:300], text[:65536], text[:65537], text, js, bytes(300000), rng.randbytes(70000),
         bytes(rng.choice(b"ab") for _ in range(100000)), bytes(min(255, int(rng.expovariate(0.05))) for _ in range(150000)),
         b"".join(bytes([rng.randrange(256)]) * rng.randint(1, 5000) for _ in range(200))]
runs = 0
for data in cases:
    for mm in (4, 5, 4 | (1 << 4), 5 | (2 << 4)):
        cap = len(data) + len(data) // 64 + 1024
        out = C.create_string_buffer(cap)                      # exact capacity
        c = Z.zenc_emul_compress(out, cap, data, len(data), mm)
        assert c > 0
        back = C.create_string_buffer(len(data) + 64)
        src = (C.c_char * c).from_buffer_copy(out.raw[:c])     # exact-size frame: reads past its end are caught
        assert D.zstd_shim_decompress(back, len(data), C.cast(src, C.c_char_p), c) == len(data)
        assert back.raw[:len(data)] == data
        runs += 1
print("ran", runs, "encoder emulations clean")
