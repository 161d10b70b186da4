"""Small workload for compute-sanitizer: every kernel, small sizes."""
import importlib, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conftest
pkg = importlib.import_module("4mc_b200")
ctx = pkg.Context(0)
data = conftest.gen_logtext(pkg, 5 * 1024 * 1024 + 777) + bytes(70000) + random.Random(1).randbytes(200000)
s = ctx.compress_4mc(data)
assert ctx.decompress_4mc(s) == data
for n in (0, 1, 13, 4096, 65537):
    s2 = ctx.compress_4mc(data[:n]); assert ctx.decompress_4mc(s2) == data[:n]
for name in ("logtext_128k.l1.4mc", "logtext_128k.l3.4mc", "zeros_4m1.4mc", "random_70000.4mc", "two_streams.4mc"):
    ctx.decompress_4mc(conftest.golden_bytes(name))
for v in conftest.golden_json("lz4_decode.json")[::25]:
    ctx.lz4_decompress_safe(bytes.fromhex(v["hex"]), v["cap"])
assert ctx.xxh32(data[:100001]) >= 0
c = ctx.lz4_compress(data[:300000]); r, o = ctx.lz4_decompress_safe(c, 300000); assert o == data[:300000]
z = ctx.compress_4mz(data)
assert ctx.decompress_4mz(z) == data
for n in (0, 1, 13, 4096, 65537, 300000):
    z2 = ctx.compress_4mz(data[:n]); assert ctx.decompress_4mz(z2) == data[:n]
f = ctx.zstd_compress(data[:200000]); r, o = ctx.zstd_decompress(f, 200000); assert o == data[:200000]
ix = pkg.FourMcBlockIndex(ctx.read_index(s))
assert b"".join(ctx.read_split_lines(s, a, ln) for a, ln in ix.plan_splits(len(s), 1 << 20)) == data
print("sanitize workload ok")
