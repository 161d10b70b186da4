#!/usr/bin/env python
"""Device-resident 4mz decode timing: builds a 4mz stream of N 4 MiB blocks with the reference's
ZSTD_compress (oracle/_ref/libref4mc.so, level 1), decodes it on the GPU, checks the bytes and
prints GB/s.  Usage: python tools/quick_4mz.py [n_blocks] [iters]"""
import ctypes as C
import importlib
import os
import sys
import struct
from concurrent.futures import ThreadPoolExecutor

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("4mc_b200")
MIB = 1 << 20


def main():
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref4mc.so"))
    R.ZSTD_compress.restype = C.c_size_t
    R.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
    R.XXH32.restype = C.c_uint32
    R.XXH32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    buf = C.create_string_buffer(nb * 4 * MIB)
    assert pkg.lib().fourmc_gen_host(0, 0x4D43, 0, nb * 1024, buf) == 0
    data = buf.raw

    def one(i):
        src = data[i * 4 * MIB:(i + 1) * 4 * MIB]
        out = C.create_string_buffer(4 * MIB)
        c = R.ZSTD_compress(out, 4 * MIB - 1, src, len(src), 1)
        payload = out.raw[:c]
        return struct.pack(">III", len(src), c, R.XXH32(payload, c, 0)) + payload

    with ThreadPoolExecutor(16) as ex:
        recs = list(ex.map(one, range(nb)))
    stream = bytes.fromhex("344d5a00 00000001 289a1c9a") + b"".join(recs) + bytes(12)
    deltas, prev = [], 0
    off = 12
    for r in recs:
        deltas.append(off - prev)
        prev = off
        off += len(r)
    foot = struct.pack(">II", 20 + 4 * nb, 1) + b"".join(struct.pack(">I", d) for d in deltas) + struct.pack(">II", 20 + 4 * nb, 0x344D5A00)
    stream += foot + struct.pack(">I", R.XXH32(foot, len(foot), 0))
    print(f"4mz: {nb} blocks, {len(stream) / MIB:.1f} MiB, ratio {len(data) / len(stream):.3f}", flush=True)

    ctx = pkg.Context(0)
    st = torch.cuda.Stream()
    d_in = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
    d_out = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for it in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.decompress_4mz_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), len(data), d_res.data_ptr(), stream=st.cuda_stream)
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"iter {it}: {ms:.1f} ms  {len(data) / ms / 1e6:.2f} GB/s  result {d_res.tolist()}", flush=True)
    ok = bytes(d_out[:len(data)].cpu().numpy()) == data
    print("bytes identical:", ok)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
