"""Device-resident timing of one codec leg with the per-kernel breakdown (development aid).

  python tools/quick_decode.py [GiB=4] [iters=3] [codec=4mc|4mz] [kind=0 log-text|1 JSON|2 mix] [level=1]

Environment switches of the library (FOURMC_D2_GATHER, FOURMC_D2_WARPS, FOURMC_DEC_MODE ...) are read by the
library itself, so A/B runs are separate processes.  Prints one line per leg and the kernels' event times.
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("4mc_b200")


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    codec = sys.argv[3] if len(sys.argv) > 3 else "4mc"
    kind = int(sys.argv[4]) if len(sys.argv) > 4 else (1 if codec == "4mz" else 0)
    level = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    n = int(gib * (1 << 30)) // 4096 * 4096
    ctx = pkg.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.gen_device(src.data_ptr(), n // 4096, seed=(0x4D43, 0x4D5A, 0x5148)[kind], kind=kind, stream=st)
    cap = pkg.lib().fourmc_4mc_bound(n)
    comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    cfn = ctx.compress_4mz_device if codec == "4mz" else ctx.compress_device
    dfn = ctx.decompress_4mz_device if codec == "4mz" else ctx.decompress_device
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FOURMC_"))

    def leg(name, fn):
        ts = []
        for it in range(iters + 1):
            if it == 1:
                ctx.timing_enable(True)
            torch.cuda.synchronize()
            e[0].record(stream)
            fn()
            e[1].record(stream)
            torch.cuda.synchronize()
            if it:
                ts.append(e[0].elapsed_time(e[1]))
        kt = ctx.timing_collect()
        ctx.timing_enable(False)
        best = min(ts)
        print(f"[{tag}] {codec} L{level} kind{kind} {gib:g} GiB {name}: best {best:.2f} ms = {n / best / 1e6:.1f} GB/s   all {[round(t, 2) for t in ts]}")
        for k, (cnt, ms) in sorted(kt.items(), key=lambda kv: -kv[1][1]):
            print(f"      {k:32s} x{cnt // iters:<3d} {ms / iters:9.3f} ms per leg")

    leg("compress", lambda: cfn(src.data_ptr(), n, comp.data_ptr(), cap, size.data_ptr(), stream=st, level=level))
    csz = int(size.item())
    leg("decompress", lambda: dfn(comp.data_ptr(), csz, out.data_ptr(), n, res.data_ptr(), stream=st))
    ok = bool(torch.equal(out, src)) and res.cpu().tolist() == [n, -1]
    print(f"[{tag}] ratio {n / csz:.3f}  round trip {'ok' if ok else 'FAILED ' + str(res.cpu().tolist())}")
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
