"""Device-resident compress timing and ratio per level, both containers (development aid).
Usage: python tools/quick_levels.py [gib]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(gib * (1 << 30)) // 4096 * 4096
ctx = pkg.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
src = torch.empty(n, dtype=torch.uint8, device="cuda")
ctx.gen_device(src.data_ptr(), n // 4096, stream=st)
cap = pkg.lib().fourmc_4mc_bound(n)
comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
size = torch.zeros(1, dtype=torch.int64, device="cuda")
out = torch.zeros(n, dtype=torch.uint8, device="cuda")
res = torch.zeros(2, dtype=torch.int64, device="cuda")
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for name, cfn, dfn in (("4mc", ctx.compress_device, ctx.decompress_device), ("4mz", ctx.compress_4mz_device, ctx.decompress_4mz_device)):
    for level in (1, 2, 3, 4):
        ts = []
        for it in range(2):
            torch.cuda.synchronize(); e[0].record()
            cfn(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(), level=level, stream=st)
            e[1].record(); torch.cuda.synchronize(); ts.append(e[0].elapsed_time(e[1]))
        csz = int(size.item())
        torch.cuda.synchronize(); e[0].record()
        dfn(comp.data_ptr(), csz, out.data_ptr(), n, res.data_ptr(), stream=st)
        e[1].record(); torch.cuda.synchronize(); td = e[0].elapsed_time(e[1])
        print(f"{name} level {level}: ratio {n / csz:.3f}  compress {n / min(ts) / 1e6:.1f} GB/s  decompress {n / td / 1e6:.1f} GB/s  ok {bool(torch.equal(out, src))}", flush=True)
