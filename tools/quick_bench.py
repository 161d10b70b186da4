"""Quick device-resident timing of the compress / decompress pipelines (development aid)."""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = int(gib * (1 << 30)) // 4096 * 4096
    ctx = pkg.Context(0)
    stream = torch.cuda.Stream()          # a real (non-default) stream: handle 0 would mean "context stream"
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    t0 = time.time()
    ctx.gen_device(src.data_ptr(), n // 4096, stream=st)
    torch.cuda.synchronize()
    print(f"gen {n / 2**30:.2f} GiB in {time.time() - t0:.3f}s")
    cap = pkg.lib().fourmc_4mc_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(iters):
        e[0].record()
        ctx.compress_device(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(), stream=st)
        e[1].record()
        torch.cuda.synchronize()
        csz = int(size.item())
        e[1].record()
        ctx.decompress_device(comp.data_ptr(), csz, out.data_ptr(), n, res.data_ptr(), stream=st)
        e[2].record()
        torch.cuda.synchronize()
        tc = e[0].elapsed_time(e[1]) if False else None
        print(f"iter {it}: csize {csz} ratio {n / csz:.3f} result {res.cpu().tolist()}")
    # separate clean timings
    def comp_fn():
        ctx.compress_device(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(), stream=st)

    def dec_fn():
        ctx.decompress_device(comp.data_ptr(), int(size.item()), out.data_ptr(), n, res.data_ptr(), stream=st)

    for name, fn in (("compress", comp_fn), ("decompress", dec_fn)):
        ts = []
        for it in range(iters):
            torch.cuda.synchronize()
            e[0].record(); fn(); e[1].record()
            torch.cuda.synchronize()
            ts.append(e[0].elapsed_time(e[1]))
        best = min(ts)
        print(f"{name}: best {best:.2f} ms  -> {n / best / 1e6:.1f} GB/s (uncompressed)   all: {[round(t, 2) for t in ts]}")
    print("equal:", bool(torch.equal(out, src)), "result", res.cpu().tolist())


if __name__ == "__main__":
    main()
