"""Small workload for compute-sanitizer: the chain kernels of levels 2..4 (lz4_chain_kernel, lz4_region_kernel<., true>), both
containers.  usage: python tools/sanitize_chain.py [tiny]"""
import importlib, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conftest
pkg = importlib.import_module("4mc_b200")
ctx = pkg.Context(0)
tiny = len(sys.argv) > 1
text = conftest.gen_logtext(pkg, (150 * 1024 + 77) if tiny else (4 * 1024 * 1024 + 300 * 1024 + 777))
data = text + bytes(70000) + random.Random(1).randbytes(20000 if tiny else 200000) + b"ab" * (5000 if tiny else 50000)
for level in ((3,) if tiny else (2, 3)):
    s = ctx.compress_4mc(data, level); assert ctx.decompress_4mc(s) == data
    z = ctx.compress_4mz(data, level); assert ctx.decompress_4mz(z) == data
    for n in (0, 1, 13, 4096, 65537):
        assert ctx.decompress_4mc(ctx.compress_4mc(data[:n], level)) == data[:n]
if not tiny:
    c = ctx.lz4_compress(text[:4 * 1024 * 1024], 4)            # one bare block through the per-block call, depth 128
    assert ctx.lz4_decompress_safe(c, 4 * 1024 * 1024)[1] == text[:4 * 1024 * 1024]
print("sanitize chain workload ok")
