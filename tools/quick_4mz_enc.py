"""Device-resident 4mz timing (development aid): GPU compress of N GiB of log text, GPU decompress,
byte compare.  Usage: python tools/quick_4mz_enc.py [gib] [iters] [decode 0|1]"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    decode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    n = int(gib * (1 << 30)) // 4096 * 4096
    ctx = pkg.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.gen_device(src.data_ptr(), n // 4096, stream=st)
    cap = pkg.lib().fourmc_4mc_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for it in range(iters):
        torch.cuda.synchronize()
        e[0].record()
        ctx.compress_4mz_device(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(), stream=st)
        e[1].record()
        torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1]))
    csz = int(size.item())
    best = min(ts)
    print(f"4mz compress: {n / 2**30:.2f} GiB -> {csz} ratio {n / csz:.3f}; best {best:.2f} ms -> {n / best / 1e6:.1f} GB/s   all: {[round(t, 2) for t in ts]}")
    if decode:
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        res = torch.zeros(2, dtype=torch.int64, device="cuda")
        ts = []
        for it in range(max(1, iters - 1)):
            torch.cuda.synchronize()
            e[0].record()
            ctx.decompress_4mz_device(comp.data_ptr(), csz, out.data_ptr(), n, res.data_ptr(), stream=st)
            e[1].record()
            torch.cuda.synchronize()
            ts.append(e[0].elapsed_time(e[1]))
        best = min(ts)
        print(f"4mz decompress: best {best:.2f} ms -> {n / best / 1e6:.2f} GB/s  result {res.tolist()}  equal {bool(torch.equal(out, src))}")


if __name__ == "__main__":
    main()
