#!/usr/bin/env python
"""bench.py -- 4mc block hot path on B200: compress + decompress throughput (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path, configs[1] (4mc Fast, log text)
  python bench.py --config 2|3|4 ...                       configs[2] 4mz Fast on JSON, configs[3] 4mc High on the mix,
                                                           configs[4] per-split reads of one .4mc through the InputFormat rules
  python bench.py --impl reference ...                     the reference's own CPU functions, all host cores

A "step" is one pass of the hot path over one batch of the synthetic input (16 GiB of the 64 GiB workload): the
batch is compressed to ONE complete .4mc / .4mz stream (block encode, XXH32, block headers, footer index) and that
stream is decompressed again (index parse, XXH32 verify, block decode), everything resident in HBM.
`value` = uncompressed bytes of the batch / step time.

Multi-GPU (one process per GPU, torchrun): STRONG scaling of that one step.  The batch's blocks are split into
contiguous ranges, one per rank; every rank compresses its range into a span; the block lengths and
span sizes are all-gathered (the only collectives), which gives every rank the footer index and every span's offset;
the spans travel over NVLink (NCCL send / recv straight into place) to rank 0, where the single stream -- header, all
spans in order, end mark, footer index -- stands assembled in HBM when the step ends (and is decoded as a whole and
compared with the input after the timed region); meanwhile every rank decodes ITS block range of that stream, found
through the shared footer index -- the bytes of its range are the ones it wrote, so it reads its local copy.
`weak` in the line is the former figure (every rank round-trips a private stream of its own data).

`e2e` is the same round trip through the host-buffer C-ABI calls (fourmc_4mc_compress_host / fourmc_4mc_decompress_host)
with pinned host buffers: PCIe copies inside the timed region.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
BLOCK = 4 * 1024 * 1024
GIB = 1 << 30

CONFIGS = {
    1: dict(workload="configs[1]: 4mc Fast (LZ4) compress+decompress 64 GiB synthetic log-text, 4 MiB blocks",
            metric="4mc_fast_lz4_compress_plus_decompress_uncompressed_GBps", codec="4mc", level=1, kind=0, seed=0x4D43,
            total_gib=64.0, batch_gib=16.0, cpu_gib=2.0, cpu1_gib=0.25, data="log-text"),
    2: dict(workload="configs[2]: 4mz Fast (ZSTD level 1) compress+decompress 64 GiB synthetic JSON, 4 MiB blocks",
            metric="4mz_fast_zstd_compress_plus_decompress_uncompressed_GBps", codec="4mz", level=1, kind=1, seed=0x4D5A,
            total_gib=64.0, batch_gib=16.0, cpu_gib=2.0, cpu1_gib=0.25, data="JSON"),
    3: dict(workload="configs[3]: 4mc High (LZ4 HC level 4) compress 16 GiB silesia-like mix, decompress, 4 MiB blocks",
            metric="4mc_high_lz4hc_compress_plus_decompress_uncompressed_GBps", codec="4mc", level=3, kind=2, seed=0x5148,
            total_gib=16.0, batch_gib=4.0, cpu_gib=0.5, cpu1_gib=0.0625, data="silesia-like mix"),
    4: dict(workload="configs[4]: FourMcCodec read path: 10 000 splits of a 40 GiB .4mc (log text) planned by the InputFormat rules, "
                     "per-split decode + line records",
            metric="4mc_split_read_uncompressed_GBps", codec="4mc", level=1, kind=0, seed=0x4D43,
            total_gib=40.0, batch_gib=4.0, cpu_gib=1.0, cpu1_gib=0.25, data="log-text"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4], help="BASELINE.json configs[k]; default 1")
    ap.add_argument("--codec", default=None, choices=["4mc", "4mz"], help="shorthand: 4mc = --config 1, 4mz = --config 2")
    ap.add_argument("--total-gib", type=float, default=None, help="the workload: synthetic input over all GPUs")
    ap.add_argument("--batch-gib", type=float, default=None, help="bytes per step over all GPUs")
    ap.add_argument("--e2e-gib", type=float, default=8.0, help="bytes per end-to-end step over all GPUs")
    ap.add_argument("--cpu-gib", type=float, default=None, help="bytes per CPU step (reference arm / cpu_baseline)")
    ap.add_argument("--split-threads", type=int, default=2, help="configs[4]: reader threads (contexts) per GPU")
    ap.add_argument("--splits-per-call", type=int, default=64, help="configs[4]: splits handed to one fourmc_read_splits_lines_host call "
                    "(1 = the single-split call)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    a = ap.parse_args()
    if not a.config:
        a.config = 2 if a.codec == "4mz" else 1
    cfg = dict(CONFIGS[a.config])
    if a.total_gib is not None:
        cfg["total_gib"] = a.total_gib
    if a.batch_gib is not None:
        cfg["batch_gib"] = a.batch_gib
    if a.cpu_gib is not None:
        cfg["cpu_gib"] = a.cpu_gib
    cfg["batch_gib"] = min(cfg["batch_gib"], cfg["total_gib"])
    a.cfg = cfg
    return a


def config_dict(cfg):
    """Identical in both arms: what the workload is, not how an arm runs it."""
    return {"workload": cfg["workload"], "total_gib": cfg["total_gib"], "batch_gib_per_step": cfg["batch_gib"], "block_mib": 4,
            "generator": f"fourmc_gen kind {cfg['kind']} ({cfg['data']}), seed {cfg['seed']:#x}",
            "l2": "inputs (GiBs per step) far larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own functions (oracle/_ref, built from its sources)
# ------------------------------------------------------------------------------------------------

class CpuReference:
    """Per block, exactly the calls of native/4mc.c:301-311 (compress + XXH32) and :637-661 (XXH32 + decompress) -- the
    function the config's level selects (:243-253, :415-425) -- spread over host threads (the reference itself is
    single-threaded; Hadoop runs one task per core).  Loads the reference build and the host-only input generator; never
    this repo's CUDA library."""

    def __init__(self, cfg, threads=None):
        ref = os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")
        self.cfg = cfg
        self.cores = threads or os.cpu_count() or 1
        self.pool = ThreadPoolExecutor(self.cores)
        level, zst = cfg["level"], cfg["codec"] == "4mz"
        if os.path.exists(ref):
            self.kind = "reference"
            L = C.CDLL(ref)
            L.XXH32.restype = C.c_uint32
            L.XXH32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
            self.xxh = L.XXH32
            if zst:
                L.ZSTD_compress.restype = C.c_size_t
                L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
                L.ZSTD_decompress.restype = C.c_size_t
                L.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
                zl = {1: 1, 2: 3, 3: 6, 4: 12}[level]                                      # native/4mc.c:415-425
                self.compress = lambda src, dst, n, cap: L.ZSTD_compress(dst, cap, src, n, zl)
                self.decompress = lambda src, dst, c, cap: L.ZSTD_decompress(dst, cap, src, c)
                self.fn = f"ZSTD_compress level {zl}+XXH32 / XXH32+ZSTD_decompress"
            else:
                for f in (L.LZ4_compress_default, L.LZ4_compressMC_limitedOutput, L.LZ4_decompress_safe):
                    f.restype = C.c_int
                    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
                L.LZ4_compress_HC.restype = C.c_int
                L.LZ4_compress_HC.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
                if level <= 1:
                    self.compress, name = L.LZ4_compress_default, "LZ4_compress_default"
                elif level == 2:
                    self.compress, name = L.LZ4_compressMC_limitedOutput, "LZ4_compressMC"
                else:
                    hl = 4 if level == 3 else 8                                            # native/4mc.c:248-252
                    self.compress, name = (lambda s, d, n, cap: L.LZ4_compress_HC(s, d, n, cap, hl)), f"LZ4_compress_HC level {hl}"
                self.decompress = L.LZ4_decompress_safe
                self.fn = f"{name}+XXH32 / XXH32+LZ4_decompress_safe"
        else:                                   # the oracle port (plain C restatement): Fast LZ4 only
            if zst or level > 1:
                raise SystemExit("this CPU arm needs oracle/_ref/libref4mc.so (make -C oracle ref)")
            self.kind = "port"
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "oracle"], check=True)
            L = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
            self.compress, self.decompress, self.xxh = L.fmo_lz4_compress, L.fmo_lz4_decompress_safe, L.fmo_xxh32
            for f in (self.compress, self.decompress):
                f.restype = C.c_int
                f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            self.xxh.restype = C.c_uint32
            self.xxh.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
            self.fn = "fmo_lz4_compress+XXH32 / XXH32+fmo_lz4_decompress_safe"

    def make_input(self, nbytes):
        host = os.path.join(ROOT, "4mc_b200", "host")
        if not os.path.exists(os.path.join(host, "libfourmcgen.so")):
            subprocess.run(["make", "-C", host, "-s", "libfourmcgen.so"], check=True)
        G = C.CDLL(os.path.join(host, "libfourmcgen.so"))
        G.fourmcgen_pages.restype = C.c_int
        G.fourmcgen_pages.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        buf = (C.c_char * nbytes)()
        base = C.addressof(buf)
        pages = nbytes // 4096
        per = max(1, pages // (max(self.cores, os.cpu_count() or 1) * 4))
        pool = self.pool if self.cores > 1 else ThreadPoolExecutor(os.cpu_count() or 1)

        def work(p0):
            G.fourmcgen_pages(self.cfg["kind"], self.cfg["seed"], p0, min(per, pages - p0), base + p0 * 4096)
        list(pool.map(work, range(0, pages, per)))
        return buf

    def step(self, src, nbytes, comp, out):
        """One round trip of nbytes; returns (t_compress, t_decompress, stored bytes)."""
        nb = nbytes // BLOCK
        sb, cb, ob = C.addressof(src), C.addressof(comp), C.addressof(out)
        slot = BLOCK + BLOCK // 255 + 64
        csz = [0] * nb
        cks = [0] * nb

        def comp_block(i):
            c = self.compress(sb + i * BLOCK, cb + i * slot, BLOCK, BLOCK - 1)        # native/4mc.c:301 / :467
            if c <= 0 or c >= BLOCK:                                                   # stored (:318-329 / :469-485)
                C.memmove(cb + i * slot, sb + i * BLOCK, BLOCK)
                c = BLOCK
            csz[i] = c
            cks[i] = self.xxh(cb + i * slot, c, 0)                                     # :311 / :323

        def dec_block(i):
            if self.xxh(cb + i * slot, csz[i], 0) != cks[i]:                           # :637 / :645
                raise RuntimeError("checksum")
            if csz[i] == BLOCK:
                C.memmove(ob + i * BLOCK, cb + i * slot, BLOCK)                        # :635-642
            elif self.decompress(cb + i * slot, ob + i * BLOCK, csz[i], BLOCK) != BLOCK:  # :661 / :810
                raise RuntimeError("decode")

        t0 = time.perf_counter()
        list(self.pool.map(comp_block, range(nb)))
        t1 = time.perf_counter()
        list(self.pool.map(dec_block, range(nb)))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, sum(csz)


def run_cpu(cfg, gib, steps, warmup, threads=None):
    ref = CpuReference(cfg, threads)
    nbytes = max(BLOCK, int(gib * GIB) // BLOCK * BLOCK)
    src = ref.make_input(nbytes)
    slot = BLOCK + BLOCK // 255 + 64
    comp = (C.c_char * ((nbytes // BLOCK) * slot))()
    out = (C.c_char * nbytes)()
    for _ in range(warmup):
        ref.step(src, nbytes, comp, out)
    tc = td = 0.0
    csum = 0
    for _ in range(steps):
        a, b, csum = ref.step(src, nbytes, comp, out)
        tc += a
        td += b
    assert bytes(out[:4096]) == bytes(src[:4096]) and bytes(out[-4096:]) == bytes(src[-4096:])
    total = nbytes * steps
    return {
        "value": total / (tc + td) / 1e9, "unit": "GB/s", "cores": ref.cores, "kind": ref.kind,
        "sample": f"{nbytes / GIB:.3f} GiB {cfg['data']} per step x {steps} steps, {ref.cores} thread(s), in memory ({ref.fn} per 4 MiB block)",
        "compress_GBps": total / tc / 1e9, "decompress_GBps": total / td / 1e9, "ratio": nbytes / csum,
        "ms_per_step": (tc + td) / steps * 1e3,
    }


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    cpu = run_cpu(cfg, cfg["cpu_gib"], args.steps, args.warmup)
    if args.config == 4:                            # a split read is the reader's half: XXH32 + LZ4_decompress_safe per block
        cpu["value"] = cpu["decompress_GBps"]
        cpu["ms_per_step"] = cfg["cpu_gib"] * GIB / (cpu["value"] * 1e9) * 1e3
        cpu["sample"] += " -- decode half only"
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": cpu["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(cfg),
        "cpu_baseline": {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "detail": {k: cpu[k] for k in ("compress_GBps", "decompress_GBps", "ratio")},
        "e2e": {"value": cpu["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons while the timed region runs (B200_PROFILING.md), sampled every 200 ms
    through NVML in a thread (an `nvidia-smi -lms 100` child was seen to stall a step by ~50 ms now and then);
    falls back to the nvidia-smi loop when the NVML binding is missing."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []                      # (time, sm_mhz, sm_max_mhz, [reason names])
        self.p = None
        self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            masks = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                     (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                     (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                     (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)

            def loop():
                while not self._stop.is_set():
                    try:
                        r = get_reasons(h)
                        self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(mx),
                                          [n for m, n in masks if r & m]))
                    except Exception:      # noqa: BLE001 -- a failed sample is just missing
                        pass
                    self._stop.wait(0.2)
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            return
        except Exception:                   # noqa: BLE001 -- no NVML binding: the nvidia-smi loop below
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "500"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]),
                                  [self.NAMES[i] for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")]))
            except (ValueError, IndexError):
                pass

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = [r[1] for r in rows]
        mx = [r[2] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted({n for r in rows for n in r[3]}), "samples": len(rows)}

    def stop(self):
        self._stop.set()
        if self.p:
            self.p.terminate()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Dist:
    """torch.distributed plumbing of one bench process (NCCL, one rank per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        # stdout carries the one JSON line only: everything libraries print there (NCCL's version banner under
        # NCCL_DEBUG, for one) goes to stderr; the line itself is written to the saved descriptor at the end
        sys.stdout.flush()
        self.json_fd = os.dup(1)
        os.dup2(2, 1)
        if self.world > 1:
            # the span exchange is pairwise send / recv of hundreds of MB: give a pair more than NCCL's default two channels
            os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def emit(self, line):
        if self.rank == 0:
            sys.stdout.flush()
            os.write(self.json_fd, (json.dumps(line) + "\n").encode())
        if self.world > 1:
            self.dist.destroy_process_group()


def check_stream_layout(stream_bytes_head, tail_bytes, total_len, nb, magic):
    """Container fields of an assembled stream: header, end mark, footer sizes / magic, the
    deltas adding up to the end mark's position.  (The bytes themselves are checked by decoding them.)"""
    be = lambda b, o: int.from_bytes(b[o:o + 4], "big")          # noqa: E731
    assert be(stream_bytes_head, 0) == magic and be(stream_bytes_head, 4) == 1
    fsize = 20 + 4 * nb
    assert len(tail_bytes) == 12 + fsize and tail_bytes[:12] == bytes(12)
    foot = tail_bytes[12:]
    assert be(foot, 0) == fsize and be(foot, 4) == 1 and be(foot, fsize - 12) == fsize and be(foot, fsize - 8) == magic
    pos = 0
    for i in range(nb):
        pos += be(foot, 8 + 4 * i)
    assert nb == 0 or 12 <= pos < total_len - 12 - fsize


def main_ours(args):
    cfg = args.cfg
    if args.config == 4:
        return main_splits(args)
    D = Dist()
    torch, dist, rank, world = D.torch, D.dist, D.rank, D.world
    pkg = importlib.import_module("4mc_b200")
    ctx = pkg.Context(D.local)                      # raises (no CPU fallback) if the device is unusable
    zst, level = cfg["codec"] == "4mz", cfg["level"]
    lib = pkg.lib()
    compress_device = ctx.compress_4mz_device if zst else ctx.compress_device
    compress_span_device = ctx.compress_4mz_span_device if zst else ctx.compress_span_device
    decompress_device = ctx.decompress_4mz_device if zst else ctx.decompress_device
    build_index_device = ctx.build_index_4mz_device if zst else ctx.build_index_device
    compress_host = lib.fourmc_4mz_compress_host if zst else lib.fourmc_4mc_compress_host
    decompress_host = lib.fourmc_4mz_decompress_host if zst else lib.fourmc_4mc_decompress_host
    magic = 0x344D5A00 if zst else 0x344D4300
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream

    # ---- the workload and this rank's share of it: contiguous block ranges of every step's batch
    batch_g = max(BLOCK * world, int(cfg["batch_gib"] * GIB) // BLOCK * BLOCK)
    total_g = max(batch_g, int(cfg["total_gib"] * GIB) // batch_g * batch_g)
    n_batches = total_g // batch_g
    nb_g = batch_g // BLOCK
    lo, hi = pkg.shard_blocks(nb_g, world, rank)
    per = -(-nb_g // world)
    my_nb = hi - lo
    my_bytes = my_nb * BLOCK
    src = torch.empty(n_batches * my_bytes, dtype=torch.uint8, device="cuda")
    for b in range(n_batches):                      # global page index = position in the 64 GiB workload
        ctx.gen_device(src.data_ptr() + b * my_bytes, my_bytes // 4096, seed=cfg["seed"], first_page=(b * nb_g + lo) * (BLOCK // 4096),
                       kind=cfg["kind"], stream=st)
    cap = lib.fourmc_4mc_bound(batch_g)
    comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")          # the ONE stream of a step, whole, on every rank
    out = torch.empty(my_bytes, dtype=torch.uint8, device="cuda")
    size = torch.zeros(1, dtype=torch.int64, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    lens = torch.zeros(per, dtype=torch.int32, device="cuda")
    if world > 1:
        span_cap = my_bytes + 12 * my_nb + 64
        span = torch.empty(span_cap, dtype=torch.uint8, device="cuda")
        all_lens = torch.zeros(per * world, dtype=torch.int32, device="cuda")
        all_sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        keep = torch.cat([torch.arange(r * per, r * per + (pkg.shard_blocks(nb_g, world, r)[1] - pkg.shard_blocks(nb_g, world, r)[0]))
                          for r in range(world)]).to("cuda")
        packed_lens = torch.zeros(nb_g, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    last = {}

    def step(i, record=None):
        b = i % n_batches
        s_ptr = src.data_ptr() + b * my_bytes
        if record:
            record[0].record(stream)
        if world == 1:
            compress_device(s_ptr, batch_g, comp.data_ptr(), cap, size.data_ptr(), d_block_lens=lens.data_ptr(), level=level, stream=st)
            if record:
                record[1].record(stream)
            csz = int(size.item())                # the stream length the reader needs (one 8-byte D2H)
            decompress_device(comp.data_ptr(), csz, out.data_ptr(), batch_g, res.data_ptr(), stream=st)
        else:
            # writer: my block range -> a span; block lengths and span sizes to everybody (the only collectives)
            compress_span_device(s_ptr, my_bytes, span.data_ptr(), span_cap, size.data_ptr(), d_block_lens=lens.data_ptr(),
                                 level=level, stream=st)
            dist.all_gather_into_tensor(all_lens, lens)
            dist.all_gather_into_tensor(all_sizes, size)
            sizes = all_sizes.cpu().tolist()      # span sizes are needed on the host: they size the transfers
            offs = pkg.span_base_offsets(sizes)
            spans_end = 12 + sum(sizes)
            csz = spans_end + 12 + 20 + 4 * nb_g
            # every rank: its own span at its place in the stream, header + end mark + footer index from the gathered lengths
            comp[offs[rank]:offs[rank] + sizes[rank]].copy_(span[:sizes[rank]], non_blocking=True)
            torch.index_select(all_lens, 0, keep, out=packed_lens)
            build_index_device(packed_lens.data_ptr(), nb_g, comp.data_ptr(), comp.data_ptr() + spans_end, stream=st)
            if record:
                record[1].record(stream)
            # the ONE stream is assembled on rank 0: the other ranks' spans travel there over NVLink (NCCL send / recv
            # straight into place) while ...
            if rank == 0:
                ops = [dist.P2POp(dist.irecv, comp[offs[p]:offs[p] + sizes[p]], p) for p in range(1, world) if sizes[p]]
            else:
                ops = [dist.P2POp(dist.isend, span[:sizes[rank]], 0)] if sizes[rank] else []
            reqs = dist.batch_isend_irecv(ops) if ops else []
            # ... every rank reads ITS block range of that stream, located through the shared footer index; the bytes of
            # its range are the ones it has just written, so it reads its local copy of them (no collective, no transfer)
            ctx.decompress_range_device(comp.data_ptr(), csz, lo, my_nb, out.data_ptr(), my_bytes, res.data_ptr(), stream=st, zstd=zst)
            for r in reqs:
                r.wait()                          # the step ends when the assembled stream stands on rank 0
        if record:
            record[2].record(stream)
        last.update(b=b, csz=csz)
        return b, csz

    for i in range(args.warmup):
        step(i)
    D.barrier()
    sampler = ClockSampler(D.local)
    time.sleep(1.0)                              # nvidia-smi's start-up queries every GPU of the box: let it pass ...
    step(args.warmup)                            # ... under one more untimed step (same batch the first timed step takes)
    D.barrier()
    ctx.timing_enable(True)
    launches0 = ctx.kernel_launches()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    csizes = []
    D.barrier()
    t_wall0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        _, csz = step(args.warmup + i, evs[i])
        csizes.append(csz)
    e1.record(stream)
    D.barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.kernel_launches() - launches0
    elapsed_ms = D.max(e0.elapsed_time(e1))
    ktimes = ctx.timing_collect()
    ctx.timing_enable(False)
    clocks = sampler.window(t_wall0, t_wall1)
    t_c = D.max(sum(e[0].elapsed_time(e[1]) for e in evs))
    t_d = D.max(sum(e[1].elapsed_time(e[2]) for e in evs))

    # verification outside the timed region: every rank's decoded range equals its input range; and the stream that was
    # assembled on rank 0 is decoded there as a whole and compared with the regenerated input of the whole batch
    r = res.cpu().tolist()
    ok = r == [my_bytes, -1] and bool(torch.equal(out, src[last["b"] * my_bytes:(last["b"] + 1) * my_bytes]))
    if rank == 0:
        csz = last["csz"]
        tail_len = 12 + 20 + 4 * nb_g
        check_stream_layout(bytes(comp[:12].cpu().numpy()), bytes(comp[csz - tail_len:csz].cpu().numpy()), csz, nb_g, magic)
        if world > 1:
            whole = torch.empty(batch_g, dtype=torch.uint8, device="cuda")
            want = torch.empty(batch_g, dtype=torch.uint8, device="cuda")
            ctx.gen_device(want.data_ptr(), batch_g // 4096, seed=cfg["seed"], first_page=last["b"] * nb_g * (BLOCK // 4096),
                           kind=cfg["kind"], stream=st)
            res2 = torch.zeros(2, dtype=torch.int64, device="cuda")
            decompress_device(comp.data_ptr(), csz, whole.data_ptr(), batch_g, res2.data_ptr(), stream=st)
            torch.cuda.synchronize()
            ok = ok and res2.cpu().tolist() == [batch_g, -1] and bool(torch.equal(whole, want))
            del whole, want
    if D.sum(0.0 if ok else 1.0) > 0:
        raise SystemExit(f"round trip FAILED on some rank (rank {rank}: result {r})")

    value = batch_g * args.steps / (elapsed_ms / 1e3) / 1e9
    mean_c = sum(csizes) / len(csizes)
    ratio = batch_g / mean_c

    # ---- roofline: algorithmic bytes (compress = read u + write 12+c; decompress = read 12+c + write u)
    # of THIS rank's share over measured times: the dominant kernel's launches, and each leg as a whole
    peak, peak_src = hbm_peak()
    my_algo = my_bytes + 12 * my_nb + mean_c * my_nb / nb_g
    main_stream = {k: v for k, v in ktimes.items() if k != "xxh_verify_kernel"}      # runs on a side stream beside the parse
    dom = max(main_stream.items(), key=lambda kv: kv[1][1]) if main_stream else None
    roofline = None
    if dom:
        name, (cnt, ms) = dom
        per_step = max(1, round(cnt / args.steps))   # launches of this kernel per step (the 4mz writer works in groups of blocks)
        algo = my_algo / per_step
        ach = algo / (ms / cnt / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(name)
            if t:
                traffic = t["dram_bytes_per_uncompressed_byte"] * my_bytes / per_step
        legs = {}
        for leg, t_leg in (("compress", t_c), ("decompress", t_d)):
            a = my_algo / (t_leg / args.steps / 1e3) / 1e9
            legs[leg] = {"ms_per_step": t_leg / args.steps, "achieved": a, "frac": a / peak}
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo,
                    "avg_launch_ms": ms / cnt, "legs": legs,
                    "step": {"achieved": 2 * my_algo / ((t_c + t_d) / args.steps / 1e3) / 1e9,
                             "frac": 2 * my_algo / ((t_c + t_d) / args.steps / 1e3) / 1e9 / peak}}

    # ---- weak scaling beside it: every rank round-trips a private stream of its own resident data
    weak = None
    if world > 1 and not args.no_weak:
        wbytes = min(len(src), int(cfg["batch_gib"] * GIB)) // BLOCK * BLOCK
        wcap = lib.fourmc_4mc_bound(wbytes)
        wcomp = comp if wcap <= comp.numel() else torch.empty(wcap, dtype=torch.uint8, device="cuda")
        wout = torch.empty(wbytes, dtype=torch.uint8, device="cuda")

        def wstep():
            compress_device(src.data_ptr(), wbytes, wcomp.data_ptr(), wcap, size.data_ptr(), level=level, stream=st)
            decompress_device(wcomp.data_ptr(), int(size.item()), wout.data_ptr(), wbytes, res.data_ptr(), stream=st)
        wstep()
        D.barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(stream)
        wn = min(args.steps, 4)
        for _ in range(wn):
            wstep()
        w1.record(stream)
        D.barrier()
        wms = D.max(w0.elapsed_time(w1))
        assert res.cpu().tolist() == [wbytes, -1]
        weak = {"value": world * wbytes * wn / (wms / 1e3) / 1e9, "unit": "GB/s", "batch_gib_per_gpu": wbytes / GIB, "steps": wn}
        del wout

    # ---- end to end through the host-buffer C-ABI (pinned host memory, PCIe inside the timed region)
    e2e = None
    if not args.no_e2e:
        en_g = min(batch_g, int(args.e2e_gib * GIB)) // (BLOCK * world) * (BLOCK * world)
        en = en_g // world                        # this rank's share of the end-to-end batch
        h_in = torch.empty(en, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(src[:en])
        ecap = lib.fourmc_4mc_bound(en)
        h_comp = torch.empty(ecap, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(en, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        e2e_t = [0.0, 0.0]

        def e2e_step():
            t_a = time.perf_counter()
            c = ctx._check(compress_host(ctx.handle, level, h_in.data_ptr(), en, h_comp.data_ptr(), ecap))
            t_b = time.perf_counter()
            d = ctx._check(decompress_host(ctx.handle, h_comp.data_ptr(), c, h_out.data_ptr(), en))
            e2e_t[0] += t_b - t_a
            e2e_t[1] += time.perf_counter() - t_b
            assert d == en
            return c
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        D.barrier()
        e2e_t[0] = e2e_t[1] = 0.0
        t0 = time.perf_counter()
        ec = 0
        for _ in range(args.steps):
            ec = e2e_step()
        torch.cuda.synchronize()
        dt = D.max(time.perf_counter() - t0)
        assert bool(torch.equal(h_out[:1 << 20], h_in[:1 << 20])) and bool(torch.equal(h_out[-(1 << 20):], h_in[-(1 << 20):]))
        # the link itself, same pinned buffers: one direction alone, then both at once
        pn = min(en, 1 << 30)
        d_a = torch.empty(pn, dtype=torch.uint8, device="cuda")
        d_b = torch.empty(pn, dtype=torch.uint8, device="cuda")
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

        def link(h2d, d2h, reps=4):
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s_a):
                        d_a.copy_(h_in[:pn], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s_b):
                        h_out[:pn].copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            return (int(h2d) + int(d2h)) * reps * pn / (time.perf_counter() - t) / 1e9
        link(True, True, 1)
        pcie = {"h2d_GBps": link(True, False), "d2h_GBps": link(False, True), "both_GBps": link(True, True)}
        del d_a, d_b
        ec_all = D.sum(float(ec))
        e2e = {"value": en_g * args.steps / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(en_g + ec_all),
               "d2h_bytes_per_step": int(ec_all + en_g), "bytes_per_step": en_g, "ms_per_step": dt / args.steps * 1e3,
               "compress_GBps": en_g * args.steps / D.max(e2e_t[0]) / 1e9, "decompress_GBps": en_g * args.steps / D.max(e2e_t[1]) / 1e9,
               "mode": "writer call, then reader call, per step; every rank its share of the batch", "pcie_rank0": pcie}
        del h_in, h_comp, h_out

    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        c = run_cpu(cfg, cfg["cpu_gib"], 2, 1)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["compress_GBps"], cpu["decompress_GBps"], cpu["ratio"] = c["compress_GBps"], c["decompress_GBps"], c["ratio"]
        c1 = run_cpu(cfg, cfg["cpu1_gib"], 1, 0, threads=1)      # the reference as shipped: one core
        cpu["one_core"] = {k: c1[k] for k in ("value", "unit", "cores", "sample", "compress_GBps", "decompress_GBps")}
    sampler.stop()

    conf = config_dict(cfg)
    line = {
        "metric": cfg["metric"], "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": conf,
        "detail": {"compress_GBps": batch_g * args.steps / (t_c / 1e3) / 1e9, "decompress_GBps": batch_g * args.steps / (t_d / 1e3) / 1e9,
                   "ratio": ratio, "level": level, "verified_round_trip": True, "single_stream_bytes": last["csz"],
                   "blocks_per_step": nb_g, "blocks_per_step_per_gpu": my_nb, "resident_input_gib_per_gpu": len(src) / GIB,
                   "parallelism": f"one stream, contiguous block ranges x{world}" + ("" if world == 1 else
                                  "; lengths + span sizes all-gathered, spans gathered to rank 0 by NCCL send/recv while every rank decodes its block range"),
                   "step_ms_rank0": [[round(e[0].elapsed_time(e[1]), 2), round(e[1].elapsed_time(e[2]), 2)] for e in evs],
                   "kernel_ms_rank0": {k: {"launches": v[0], "total_ms": round(v[1], 3)} for k, v in sorted(ktimes.items())}},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "weak": weak, "gpu_launches": launches, "clocks": clocks,
    }
    D.emit(line)


def main_splits(args):
    """configs[4]: one .4mc file (log text, written by this repo's writer) in host memory; the splits Hadoop would
    plan for it (FourMcInputFormat.getSplits over byte ranges of file/10000 bytes) are dealt round-robin to the ranks
     and read through fourmc_read_splits_lines_host, `--splits-per-call` at a time (a node runs many map
    tasks over one file: their splits are handed over together and decoded as one batch) -- per split: index lookup,
    its blocks decoded on the GPU, line boundaries found there, the records copied back.  --splits-per-call 1 is the
    single-split call (fourmc_read_split_lines_host).  A step = `batch_gib` worth of splits, over `--split-threads`
    reader threads (one context each)."""
    cfg = args.cfg
    D = Dist()
    torch, rank, world = D.torch, D.rank, D.world
    pkg = importlib.import_module("4mc_b200")
    import psutil
    lib = pkg.lib()
    ctx = pkg.Context(D.local)
    # the file: as much of the 40 GiB as this host can hold next to the other ranks' copies
    avail = psutil.virtual_memory().available
    file_target = int(min(cfg["total_gib"] * GIB, avail * 0.5 / world))
    n_in = max(64 * BLOCK, int(file_target * 2.0) // BLOCK * BLOCK)        # log text compresses about 2:1
    slice_bytes = min(n_in, 4 * GIB)
    d_src = torch.empty(slice_bytes, dtype=torch.uint8, device="cuda")
    h_file = torch.empty(int(n_in * 0.62) + (64 << 20), dtype=torch.uint8)   # pageable, like a file read into memory (log text: 2:1)
    # written as ONE stream: spans per slice, footer from all lengths (the sharded writer's path, on one rank)
    lens_all, pos = [], 12
    d_span = torch.empty(slice_bytes + 12 * (slice_bytes // BLOCK) + 64, dtype=torch.uint8, device="cuda")
    d_size = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_lens = torch.zeros(slice_bytes // BLOCK, dtype=torch.int32, device="cuda")
    for off in range(0, n_in, slice_bytes):
        n = min(slice_bytes, n_in - off)
        ctx.gen_device(d_src.data_ptr(), n // 4096, seed=cfg["seed"], first_page=off // 4096, kind=cfg["kind"])
        ctx.compress_span_device(d_src.data_ptr(), n, d_span.data_ptr(), d_span.numel(), d_size.data_ptr(), d_block_lens=d_lens.data_ptr())
        torch.cuda.synchronize()
        s = int(d_size.item())
        assert pos + s + 64 + 4 * (n_in // BLOCK) < h_file.numel(), "the synthetic text compressed worse than 1.6:1"
        h_file[pos:pos + s].copy_(d_span[:s])
        pos += s
        lens_all.append(d_lens[:n // BLOCK].clone())
    lens_t = torch.cat(lens_all)
    nb = lens_t.numel()
    d_hdr = torch.zeros(16, dtype=torch.uint8, device="cuda")
    d_tail = torch.zeros(12 + 20 + 4 * nb + 16, dtype=torch.uint8, device="cuda")
    ctx.build_index_device(lens_t.data_ptr(), nb, d_hdr.data_ptr(), d_tail.data_ptr())
    torch.cuda.synchronize()
    h_file[:12].copy_(d_hdr[:12])
    tail_len = 12 + 20 + 4 * nb
    h_file[pos:pos + tail_len].copy_(d_tail[:tail_len])
    file_size = pos + tail_len
    del d_src, d_span
    fptr = h_file.data_ptr()
    offsets = (C.c_int64 * nb)()
    assert lib.fourmc_read_index_host(ctx.handle, fptr, file_size, offsets, nb) == nb
    n_splits_target = max(world, int(10000 * file_size / (40 * GIB)))
    split_size = max(1, file_size // n_splits_target)
    ns = lib.fourmc_plan_splits(offsets, nb, file_size, split_size, None, None, 0)
    st_arr, ln_arr = (C.c_int64 * ns)(), (C.c_int64 * ns)()
    lib.fourmc_plan_splits(offsets, nb, file_size, split_size, st_arr, ln_arr, ns)
    mine = pkg.shard_splits(ns, world, rank)
    per_step = max(1, int(len(mine) * min(1.0, cfg["batch_gib"] * GIB / n_in)))
    T = max(1, args.split_threads)
    B = max(1, args.splits_per_call)
    ctxs = [ctx] + [pkg.Context(D.local) for _ in range(T - 1)]
    out_cap = int(B * (split_size * 8 + 3 * BLOCK))
    bufs = [torch.empty(out_cap, dtype=torch.uint8, pin_memory=True) for _ in range(T)]
    pool = ThreadPoolExecutor(T)

    def read_some(t, idxs):
        got = 0
        for j0 in range(0, len(idxs), B):
            part = idxs[j0:j0 + B]
            if B == 1:
                r = lib.fourmc_read_split_lines_host(ctxs[t].handle, fptr, file_size, st_arr[part[0]], ln_arr[part[0]], bufs[t].data_ptr(), out_cap)
            else:
                a = (C.c_int64 * len(part))(*[st_arr[i] for i in part])
                b = (C.c_int64 * len(part))(*[ln_arr[i] for i in part])
                o = (C.c_int64 * (len(part) + 1))()
                r = lib.fourmc_read_splits_lines_host(ctxs[t].handle, fptr, file_size, len(part), a, b, bufs[t].data_ptr(), out_cap, o)
            if r < 0:
                raise RuntimeError(f"splits {part[:3]}..: error {r}")
            got += r
        return got

    cursor = [0]

    def step():
        idxs = [mine[(cursor[0] + j) % len(mine)] for j in range(per_step)]
        cursor[0] += per_step
        chunk = -(-len(idxs) // T)
        return sum(pool.map(lambda t: read_some(t, idxs[t * chunk:(t + 1) * chunk]), range(T)))

    for _ in range(min(args.warmup, 2)):
        step()
    D.barrier()
    launches0 = sum(c.kernel_launches() for c in ctxs)
    sampler = ClockSampler(D.local)
    t0w = time.perf_counter()
    got = 0
    for _ in range(args.steps):
        got += step()
    torch.cuda.synchronize()
    dt = D.max(time.perf_counter() - t0w)
    t1w = time.perf_counter()
    launches = sum(c.kernel_launches() for c in ctxs) - launches0
    total_got = D.sum(float(got))
    splits_done = D.sum(float(per_step * args.steps))
    # every line exactly once: all of this rank's splits together return the bytes of... checked on a sample
    # against the plain decode of the covering blocks in tests/test_splits_gpu.py; here: totals are plausible
    assert got > 0
    clocks = sampler.window(t0w, t1w)
    sampler.stop()
    peak, peak_src = hbm_peak()
    value = total_got / dt / 1e9
    comp_read = splits_done * split_size
    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        c = run_cpu(dict(cfg, level=1), cfg["cpu_gib"], 1, 1)
        cpu = {"value": c["decompress_GBps"], "unit": "GB/s", "cores": c["cores"], "kind": c["kind"],
               "sample": c["sample"] + " -- the decode half only (XXH32 + LZ4_decompress_safe), which is what a split read runs"}
    line = {
        "metric": cfg["metric"], "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": config_dict(cfg),
        "detail": {"file_gib": file_size / GIB, "uncompressed_gib": n_in / GIB, "blocks": nb, "splits_planned": ns,
                   "split_bytes": split_size, "splits_per_step_per_gpu": per_step, "threads_per_gpu": T, "splits_per_call": B,
                   "splits_per_s": splits_done / dt, "host_copy": "the file is replicated in every rank's host memory (pageable)",
                   "note": "file scaled to the host memory of this box when 40 GiB x ranks does not fit"},
        "roofline": {"bound": "hbm", "kernel": "split decode (parse + copy kernels over the blocks of one call)",
                     "achieved": (total_got + comp_read) / dt / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": (total_got + comp_read) / dt / 1e9 / peak, "traffic": None, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": int(comp_read / args.steps), "d2h_bytes_per_step": int(total_got / args.steps),
                "mode": "this path IS end to end: host file in, host records out, per split"},
        "gpu_launches": launches, "clocks": clocks,
    }
    for c in ctxs[1:]:
        c.close()
    D.emit(line)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
