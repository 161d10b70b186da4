"""file -> file through the CLIs, disk (tmpfs) + PCIe included, for honesty.
Times this repo's CLI (4mc_b200/host/4mc, GPU) and the reference CLI (oracle/_ref/4mc, one core) on the same
synthetic log-text file, both directions, and cross-checks: each CLI decodes the other's file.
usage: python tools/cli_file_timing.py [GiB=2] [-z]"""
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("4mc_b200")
import ctypes as C  # noqa: E402


def run(cmd):
    t = time.perf_counter()
    subprocess.run(cmd, check=True)
    return time.perf_counter() - t


def main():
    args = [a for a in sys.argv[1:] if a != "-z"]
    z = ["-z"] if "-z" in sys.argv else []
    gib = float(args[0]) if args else 2.0
    n = int(gib * (1 << 30)) // 4096 * 4096
    d = "/dev/shm/fourmc_t2"
    os.makedirs(d, exist_ok=True)
    src = os.path.join(d, "in.bin")
    buf = C.create_string_buffer(n)
    assert pkg.lib().fourmc_gen_host(0, 0x4D43, 0, n // 4096, buf) == 0
    t = time.perf_counter()
    with open(src, "wb") as f:
        f.write(buf.raw)
    t_w = time.perf_counter() - t
    t = time.perf_counter()
    with open(src, "rb") as f:
        while f.read(1 << 26):
            pass
    t_r = time.perf_counter() - t
    del buf
    ours, ref = os.path.join(ROOT, "4mc_b200", "host", "4mc"), os.path.join(ROOT, "oracle", "_ref", "4mc")
    ext = ".4mz" if z else ".4mc"
    res = {"bytes": n, "codec": "4mz" if z else "4mc", "storage": "tmpfs",
           "storage_one_thread_GBps": {"write": n / t_w / 1e9, "read": n / t_r / 1e9}}
    for name, exe in (("gpu_cli", ours), ("reference_cli_1_core", ref)):
        comp, back = os.path.join(d, name + ext), os.path.join(d, name + ".out")
        tc = min(run([exe, "-f", "-q", "-q"] + z + ["-1", src, comp]) for _ in range(2))
        td = min(run([exe, "-f", "-q", "-q"] + z + ["-d", comp, back]) for _ in range(2))
        assert subprocess.run(["cmp", "-s", src, back]).returncode == 0
        res[name] = {"compress_GBps": n / tc / 1e9, "decompress_GBps": n / td / 1e9, "file_bytes": os.path.getsize(comp)}
    # each CLI reads the other's file
    for a, b in (("gpu_cli", ref), ("reference_cli_1_core", ours)):
        back = os.path.join(d, "cross.out")
        run([b, "-f", "-q", "-q"] + z + ["-d", os.path.join(d, a + ext), back])
        assert subprocess.run(["cmp", "-s", src, back]).returncode == 0
    res["cross_decode"] = "ok"
    print(json.dumps(res))
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))


if __name__ == "__main__":
    main()
