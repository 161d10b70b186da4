"""Small workload for compute-sanitizer: the chain parse (levels 2..4), both containers."""
import importlib, os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conftest
pkg = importlib.import_module("4mc_b200")
ctx = pkg.Context(0)
data = conftest.gen_logtext(pkg, 5 * 1024 * 1024 + 777) + bytes(70000) + random.Random(1).randbytes(200000) + b"ab" * 50000
for level in (2, 4):
    s = ctx.compress_4mc(data, level); assert ctx.decompress_4mc(s) == data
    z = ctx.compress_4mz(data, level); assert ctx.decompress_4mz(z) == data
    for n in (0, 1, 13, 4096, 65537):
        assert ctx.decompress_4mc(ctx.compress_4mc(data[:n], level)) == data[:n]
print("sanitize chain workload ok")
