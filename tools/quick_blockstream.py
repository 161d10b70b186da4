"""Raw codec streams (Hadoop block-stream framing, fourmc_blockstream_*): host-buffer timings (development aid).
usage: python tools/quick_blockstream.py [GiB=1] [zstd=0]"""
import ctypes as C
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    zstd = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n = int(gib * (1 << 30)) // 4096 * 4096
    L = pkg.lib()
    L.fourmc_blockstream_bound.restype = C.c_size_t
    L.fourmc_blockstream_bound.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
    L.fourmc_blockstream_compress_host.restype = C.c_longlong
    L.fourmc_blockstream_compress_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    L.fourmc_blockstream_decompress_host.restype = C.c_longlong
    L.fourmc_blockstream_decompress_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    ctx = pkg.Context(0)
    src = C.create_string_buffer(n)
    assert L.fourmc_gen_host(0, 0x4D43, 0, n // 4096, src) == 0
    cap = L.fourmc_blockstream_bound(zstd, n, 65536)
    comp = C.create_string_buffer(cap)
    out = C.create_string_buffer(n)
    for it in range(3):
        t0 = time.perf_counter()
        c = L.fourmc_blockstream_compress_host(ctx.handle, zstd, 1, src, n, 65536, comp, cap)
        t1 = time.perf_counter()
        assert c > 0, (c, ctx.last_error())
        d = L.fourmc_blockstream_decompress_host(ctx.handle, zstd, comp, c, out, n)
        t2 = time.perf_counter()
        assert d == n, (d, ctx.last_error())
        os.environ["FOURMC_BS_SERIAL"] = "1"
        t3 = time.perf_counter()
        d = L.fourmc_blockstream_decompress_host(ctx.handle, zstd, comp, c, out, n)
        t4 = time.perf_counter()
        del os.environ["FOURMC_BS_SERIAL"]
        assert d == n
        print(f"{'zstd' if zstd else 'lz4'} {n / 2**30:.2f} GiB, 64 KiB writes: ratio {n / c:.3f}  compress (one batch) {n / (t1 - t0) / 1e9:.2f} GB/s  "
              f"decompress (one batch) {n / (t2 - t1) / 1e9:.2f} GB/s  decompress (chunk by chunk) {n / (t4 - t3) / 1e9:.2f} GB/s")
    assert out.raw == src.raw
    print("equal: True")


if __name__ == "__main__":
    main()
