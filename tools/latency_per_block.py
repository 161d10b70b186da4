"""Latency of the per-block calls (what the JNI natives do: one 4 MiB block per call, host buffers)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conftest
pkg = importlib.import_module("4mc_b200")
ctx = pkg.Context(0)
blk = conftest.gen_logtext(pkg, 4 << 20)


def timeit(fn, n=6):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


c1 = ctx.lz4_compress(blk, 1); z1 = ctx.zstd_compress(blk, 1)
print(f"per 4 MiB block, host buffers (ms): LZ4 compress L1 {timeit(lambda: ctx.lz4_compress(blk, 1)):.2f}  L3 {timeit(lambda: ctx.lz4_compress(blk, 3)):.2f}  "
      f"LZ4 decompress {timeit(lambda: ctx.lz4_decompress_safe(c1, 4 << 20)):.2f}  ZSTD compress L1 {timeit(lambda: ctx.zstd_compress(blk, 1)):.2f}  "
      f"ZSTD decompress {timeit(lambda: ctx.zstd_decompress(z1, 4 << 20)):.2f}  XXH32 {timeit(lambda: ctx.xxh32(blk)):.2f}")
