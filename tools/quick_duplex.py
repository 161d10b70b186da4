"""Host-buffer writer and reader calls side by side (two contexts, two host threads): do both PCIe directions carry
payload at once?  Development aid.  usage: python tools/quick_duplex.py [GiB=4] [rounds=4]"""
import importlib, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")
lib = pkg.lib()
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(gib * (1 << 30)) // (4 << 20) * (4 << 20)
a, b = pkg.Context(0), pkg.Context(0)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
a.gen_device(src.data_ptr(), n // 4096); a.sync()
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.copy_(src)
cap = lib.fourmc_4mc_bound(n)
h_c1 = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
h_c2 = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
c = a._check(lib.fourmc_4mc_compress_host(a.handle, 1, h_in.data_ptr(), n, h_c1.data_ptr(), cap))
assert b._check(lib.fourmc_4mc_decompress_host(b.handle, h_c1.data_ptr(), c, h_out.data_ptr(), n)) == n


def wr(k):
    for _ in range(k):
        a._check(lib.fourmc_4mc_compress_host(a.handle, 1, h_in.data_ptr(), n, h_c2.data_ptr(), cap))


def rd(k):
    for _ in range(k):
        b._check(lib.fourmc_4mc_decompress_host(b.handle, h_c1.data_ptr(), c, h_out.data_ptr(), n))


def timed(fns):
    ts = [threading.Thread(target=f, args=(rounds,)) for f in fns]
    t = time.perf_counter()
    for x in ts: x.start()
    for x in ts: x.join()
    return time.perf_counter() - t


tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FOURMC_"))
tw, tr = timed([wr]), timed([rd])
tb = timed([wr, rd])
print(f"[{tag}] {gib:g} GiB x {rounds}: writer alone {n * rounds / tw / 1e9:.1f} GB/s, reader alone {n * rounds / tr / 1e9:.1f} GB/s, "
      f"serial round trip {n * rounds / (tw + tr) / 1e9:.1f} GB/s, side by side {n * rounds / tb / 1e9:.1f} GB/s of round trips "
      f"({tb / rounds * 1e3:.0f} ms per pair; alone {tw / rounds * 1e3:.0f} + {tr / rounds * 1e3:.0f})", flush=True)
