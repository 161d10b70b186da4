"""Tiny workload for the slow sanitizer tools (racecheck / synccheck): every kernel once, a few hundred KiB."""
import importlib, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conftest
pkg = importlib.import_module("4mc_b200")
ctx = pkg.Context(0)
data = conftest.gen_logtext(pkg, 300000) + bytes(20000) + random.Random(1).randbytes(30000) + b"ab" * 5000
for level in (1, 3):
    s = ctx.compress_4mc(data, level); assert ctx.decompress_4mc(s) == data
    z = ctx.compress_4mz(data, level); assert ctx.decompress_4mz(z) == data
assert ctx.decompress_4mz(conftest.golden_bytes("logtext_128k.z3.4mz")) == conftest.golden_bytes("logtext_128k.bin")
assert ctx.decompress_4mc(conftest.golden_bytes("logtext_128k.l3.4mc")) == conftest.golden_bytes("logtext_128k.bin")
ix = pkg.FourMcBlockIndex(ctx.read_index(s))
assert b"".join(ctx.read_split_lines(s, a, ln) for a, ln in ix.plan_splits(len(s), 1 << 20)) == data
print("sanitize tiny workload ok")
