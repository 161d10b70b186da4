"""4mc_b200 -- host-side mirror of the reference's codec interface over lib4mcgpu.so.

The product is the C-ABI library (include/fourmc.h, 4mc_b200/csrc/*): hand-written sm_100a CUDA
for the per-4 MiB-block LZ4 compress / decompress, XXH32 and block-index path of fingltd/4mc.
This module is the thin Python binding used by the tests and bench.py; it adds no arithmetic of
its own and never falls back to a CPU codec -- if the library or a CUDA device is missing, calls
raise FourMcError.

Names follow the reference (java/hadoop-4mc/src/main/java/com/fing/compression/fourmc):
  Lz4Compressor.compress_bytes_direct / Lz4Decompressor.decompress_bytes_direct / xxhash32
      <- Lz4Compressor.java:314-321, Lz4Decompressor.java:287-290 (JNI natives)
  FourMcCodec.compress / decompress (whole .4mc streams)
      <- FourMcCodec.java:82-168, native/4mc.c:220 fourMCcompressFilename, :896 fourMcDecompressFileName
The package name starts with a digit, so import it with importlib.import_module("4mc_b200").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FOURMC_LIB") or os.path.join(_HERE, "lib4mcgpu.so")      # FOURMC_LIB: a development build to A/B against
CSRC = os.path.join(_HERE, "csrc")

BLOCKSIZE = 4 * 1024 * 1024
OK, E_GENERIC, E_INPUT, E_OUTPUT, E_CONTENT = 0, -1, -2, -3, -4
E_CUDA, E_ARG, E_UNSUPPORTED = -10, -11, -12
BLOCK_OK, BLOCK_CHECKSUM, BLOCK_CORRUPT, BLOCK_TOOLARGE = 0, 1, 2, 3


class FourMcError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"fourmc error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compiles csrc/ into lib4mcgpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "fourmc.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.isfile(s))
    if stale:
        subprocess.run(["make", "-C", CSRC, "-s"], check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Loads lib4mcgpu.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
    sig = {
        "fourmc_ctx_create": (i32, [C.POINTER(vp), i32]),
        "fourmc_ctx_destroy": (None, [vp]),
        "fourmc_last_error": (C.c_char_p, [vp]),
        "fourmc_kernel_launches": (u64, [vp]),
        "fourmc_sync": (i32, [vp, vp]),
        "fourmc_timing_enable": (i32, [vp, i32]),
        "fourmc_ctx_set_reproducible": (i32, [vp, i32]),
        "fourmc_timing_collect": (C.c_longlong, [vp, C.c_char_p, sz]),
        "fourmc_lz4_compress_bound": (i32, [i32]),
        "fourmc_lz4_compress": (i32, [vp, i32, vp, i32, vp, i32]),
        "fourmc_lz4_decompress_safe": (i32, [vp, vp, i32, vp, i32]),
        "fourmc_xxh32": (u32, [vp, vp, sz, u32, C.POINTER(i32)]),
        "fourmc_4mc_bound": (sz, [sz]),
        "fourmc_4mc_compress_host": (C.c_longlong, [vp, i32, vp, sz, vp, sz]),
        "fourmc_4mc_decompress_host": (C.c_longlong, [vp, vp, sz, vp, sz]),
        "fourmc_4mc_decoded_size_host": (C.c_longlong, [vp, sz]),
        "fourmc_4mc_compress_device": (i32, [vp, vp, i32, vp, sz, vp, sz, vp, vp]),
        "fourmc_4mc_compress_span_device": (i32, [vp, vp, i32, vp, sz, vp, sz, vp, vp]),
        "fourmc_4mc_build_index_device": (i32, [vp, vp, vp, u32, vp, vp]),
        "fourmc_4mc_decompress_device": (i32, [vp, vp, vp, sz, vp, sz, vp]),
        "fourmc_4mc_decompress_range_device": (i32, [vp, vp, vp, sz, u32, u32, vp, sz, vp]),
        "fourmc_4mz_decompress_range_device": (i32, [vp, vp, vp, sz, u32, u32, vp, sz, vp]),
        "fourmc_compress_fd": (C.c_longlong, [vp, i32, i32, i32, i32, C.POINTER(u64)]),
        "fourmc_decompress_fd": (C.c_longlong, [vp, i32, i32, i32, C.POINTER(u64)]),
        "fourmc_lz4_decompress_batch_device": (i32, [vp, vp, u32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
        "fourmc_xxh32_batch_device": (i32, [vp, vp, u32, vp, vp, vp, u32, vp]),
        "fourmc_4mz_decompress_host": (C.c_longlong, [vp, vp, sz, vp, sz]),
        "fourmc_4mz_decoded_size_host": (C.c_longlong, [vp, sz]),
        "fourmc_4mz_decompress_device": (i32, [vp, vp, vp, sz, vp, sz, vp]),
        "fourmc_zstd_decompress": (C.c_longlong, [vp, vp, sz, vp, sz]),
        "fourmc_zstd_compress": (C.c_longlong, [vp, i32, vp, sz, vp, sz]),
        "fourmc_zstd_compress_bound": (sz, [sz]),
        "fourmc_4mz_compress_host": (C.c_longlong, [vp, i32, vp, sz, vp, sz]),
        "fourmc_4mz_compress_device": (i32, [vp, vp, i32, vp, sz, vp, sz, vp, vp]),
        "fourmc_4mz_compress_span_device": (i32, [vp, vp, i32, vp, sz, vp, sz, vp, vp]),
        "fourmc_4mz_build_index_device": (i32, [vp, vp, vp, u32, vp, vp]),
        "fourmc_read_index_host": (C.c_longlong, [vp, vp, sz, vp, sz]),
        "fourmc_index_find_next_position": (C.c_int64, [vp, i32, C.c_int64]),
        "fourmc_index_find_belonging_block": (C.c_int64, [vp, i32, C.c_int64]),
        "fourmc_index_align_slice_start": (C.c_int64, [vp, i32, C.c_int64, C.c_int64]),
        "fourmc_index_align_slice_end": (C.c_int64, [vp, i32, C.c_int64, C.c_int64]),
        "fourmc_plan_splits": (i32, [vp, i32, C.c_int64, C.c_int64, vp, vp, i32]),
        "fourmc_read_split_lines_host": (C.c_longlong, [vp, vp, sz, C.c_int64, C.c_int64, vp, sz]),
        "fourmc_read_splits_lines_host": (C.c_longlong, [vp, vp, sz, i32, vp, vp, vp, sz, vp]),
        "fourmc_blockstream_bound": (sz, [i32, sz, sz]),
        "fourmc_blockstream_compress_host": (C.c_longlong, [vp, i32, i32, vp, sz, sz, vp, sz]),
        "fourmc_blockstream_decompress_host": (C.c_longlong, [vp, i32, vp, sz, vp, sz]),
        "fourmc_gen_device": (i32, [vp, vp, i32, u64, u64, u64, vp]),
        "fourmc_gen_host": (i32, [i32, u64, u64, u64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _buf(b):
    """ctypes view of a bytes-like object without copying (read-only objects are copied)."""
    if isinstance(b, (bytes, bytearray)):
        return (C.c_char * len(b)).from_buffer_copy(b) if isinstance(b, bytes) else (C.c_char * len(b)).from_buffer(b)
    mv = memoryview(b).cast("B")
    return (C.c_char * len(mv)).from_buffer(mv) if not mv.readonly else (C.c_char * len(mv)).from_buffer_copy(mv)


class Context:
    """One fourmc_ctx: a CUDA device, a stream and grow-only workspaces (include/fourmc.h)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        rc = lib().fourmc_ctx_create(C.byref(self._h), device)
        if rc != OK:
            self._h = C.c_void_p()
            raise FourMcError(rc, "fourmc_ctx_create failed (no CUDA device?)")

    def close(self):
        if self._h:
            lib().fourmc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except TypeError:       # interpreter shutdown: the module globals are already gone, the process frees the rest
            pass

    @property
    def handle(self):
        return self._h

    def _check(self, rc: int):
        if rc < 0:
            raise FourMcError(rc, lib().fourmc_last_error(self._h).decode())
        return rc

    def last_error(self) -> str:
        return lib().fourmc_last_error(self._h).decode()

    def kernel_launches(self) -> int:
        return int(lib().fourmc_kernel_launches(self._h))

    def sync(self, stream=None):
        self._check(lib().fourmc_sync(self._h, stream))

    def set_reproducible(self, mode: int):
        """-1: per entry point (host-facing calls yes, device-resident calls no), 0 / 1: never / always."""
        self._check(lib().fourmc_ctx_set_reproducible(self._h, mode))

    def timing_enable(self, on: bool = True):
        self._check(lib().fourmc_timing_enable(self._h, 1 if on else 0))

    def timing_collect(self) -> dict:
        """{kernel name: (launches, total ms)} since the last collect; synchronises the device."""
        buf = C.create_string_buffer(1 << 16)
        n = self._check(lib().fourmc_timing_collect(self._h, buf, len(buf)))
        out = {}
        for line in buf.raw[:n].decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    # ---- per block, host buffers (what the JNI natives and the CLI loop call) ----
    def xxh32(self, data, seed: int = 0) -> int:
        st = C.c_int(0)
        b = _buf(data)
        h = lib().fourmc_xxh32(self._h, b, len(b), seed & 0xFFFFFFFF, C.byref(st))
        self._check(st.value)
        return int(h)

    def lz4_compress(self, data, level: int = 1, capacity: int | None = None) -> bytes | None:
        """LZ4_compress_default semantics: None when the block does not fit in `capacity`."""
        n = len(data)
        cap = lib().fourmc_lz4_compress_bound(n) if capacity is None else capacity
        out = C.create_string_buffer(max(cap, 1))
        rc = lib().fourmc_lz4_compress(self._h, level, _buf(data), n, out, cap)
        if rc < 0:
            self._check(rc)
        return None if rc == 0 else out.raw[:rc]

    def lz4_decompress_safe(self, data, capacity: int):
        """Returns (return value of LZ4_decompress_safe, decoded bytes)."""
        out = C.create_string_buffer(max(capacity, 1))
        rc = lib().fourmc_lz4_decompress_safe(self._h, _buf(data), len(data), out, capacity)
        if rc in (E_CUDA, E_ARG) and self.last_error():
            raise FourMcError(rc, self.last_error())
        return rc, out.raw[:max(rc, 0)]

    # ---- whole streams, host buffers ----
    def compress_4mc(self, data, level: int = 1) -> bytes:
        n = len(data)
        cap = lib().fourmc_4mc_bound(n)
        out = C.create_string_buffer(cap)
        rc = lib().fourmc_4mc_compress_host(self._h, level, _buf(data), n, out, cap)
        self._check(rc)
        return out.raw[:rc]

    def decompress_4mc(self, stream) -> bytes:
        b = _buf(stream)
        size = lib().fourmc_4mc_decoded_size_host(b, len(b))
        cap = max(int(size), 0)
        out = C.create_string_buffer(max(cap, 1))
        rc = lib().fourmc_4mc_decompress_host(self._h, b, len(b), out, cap)
        self._check(rc)
        return out.raw[:rc]

    def decompress_4mc_rc(self, stream, capacity: int) -> int:
        """Only the status / decoded size (negative FOURMC_E_* on failure), no exception."""
        b = _buf(stream)
        cap = capacity
        out = C.create_string_buffer(max(cap, 1))
        return int(lib().fourmc_4mc_decompress_host(self._h, b, len(b), out, cap))

    # ---- 4mz (zstd blocks) ----
    def zstd_compress(self, data, level: int = 1, capacity: int | None = None):
        """ZSTD_compress on one block: the frame, or None when it does not fit in `capacity`
        (the reference returns dstSize_tooSmall there and 4mc stores the block, native/4mc.c:469)."""
        n = len(data)
        cap = int(lib().fourmc_zstd_compress_bound(n)) if capacity is None else capacity
        out = C.create_string_buffer(max(cap, 1))
        rc = int(lib().fourmc_zstd_compress(self._h, level, _buf(data), n, out, cap))
        if rc == -70:
            return None
        self._check(rc)
        return out.raw[:rc]

    def compress_4mz(self, data, level: int = 1) -> bytes:
        n = len(data)
        cap = lib().fourmc_4mc_bound(n)
        out = C.create_string_buffer(cap)
        rc = lib().fourmc_4mz_compress_host(self._h, level, _buf(data), n, out, cap)
        self._check(rc)
        return out.raw[:rc]

    def compress_4mz_device(self, d_in: int, n: int, d_out: int, out_capacity: int, d_out_size: int,
                            d_block_lens: int | None = None, level: int = 1, stream=None):
        self._check(lib().fourmc_4mz_compress_device(self._h, stream, level, d_in, n, d_out, out_capacity,
                                                     d_out_size, d_block_lens))

    def compress_4mz_span_device(self, d_in: int, n: int, d_span: int, span_capacity: int, d_span_size: int,
                                 d_block_lens: int | None = None, level: int = 1, stream=None):
        self._check(lib().fourmc_4mz_compress_span_device(self._h, stream, level, d_in, n, d_span, span_capacity,
                                                          d_span_size, d_block_lens))

    def build_index_4mz_device(self, d_block_lens: int, n_blocks: int, d_header: int | None, d_tail: int, stream=None):
        self._check(lib().fourmc_4mz_build_index_device(self._h, stream, d_block_lens, n_blocks, d_header, d_tail))

    def zstd_decompress(self, data, capacity: int):
        """Returns (decoded size or negative zstd-style error, decoded bytes): ZSTD_decompress on one block."""
        out = C.create_string_buffer(max(capacity, 1))
        rc = int(lib().fourmc_zstd_decompress(self._h, _buf(data), len(data), out, capacity))
        if rc in (E_CUDA, E_ARG) and self.last_error():
            raise FourMcError(rc, self.last_error())
        return rc, out.raw[:max(rc, 0)]

    def decompress_4mz(self, stream) -> bytes:
        b = _buf(stream)
        cap = max(int(lib().fourmc_4mz_decoded_size_host(b, len(b))), 0)
        out = C.create_string_buffer(max(cap, 1))
        rc = lib().fourmc_4mz_decompress_host(self._h, b, len(b), out, cap)
        self._check(rc)
        return out.raw[:rc]

    def decompress_4mz_rc(self, stream, capacity: int) -> int:
        b = _buf(stream)
        out = C.create_string_buffer(max(capacity, 1))
        return int(lib().fourmc_4mz_decompress_host(self._h, b, len(b), out, capacity))

    def decompress_4mz_device(self, d_in: int, n: int, d_out: int, out_capacity: int, d_result: int, stream=None):
        self._check(lib().fourmc_4mz_decompress_device(self._h, stream, d_in, n, d_out, out_capacity, d_result))

    # ---- raw codec streams: Lz4Codec / ZstdCodec over Hadoop's BlockCompressorStream framing ----
    def compress_blockstream(self, data, zstd: bool = False, level: int = 1, write_size: int = 0) -> bytes:
        """What `Lz4Codec.createOutputStream` (Lz4Codec.java:95-104; `ZstdCodec` with zstd=True) writes when the
        application hands over `data` in write() calls of `write_size` bytes (0 = one call)."""
        b = _buf(data)
        cap = int(lib().fourmc_blockstream_bound(int(zstd), len(b), write_size))
        out = C.create_string_buffer(cap)
        rc = self._check(int(lib().fourmc_blockstream_compress_host(self._h, int(zstd), level, b, len(b), write_size, out, cap)))
        return out.raw[:rc]

    def decompress_blockstream(self, stream, capacity: int, zstd: bool = False) -> bytes:
        """What `Lz4Codec.createInputStream` (Lz4Codec.java:128-138) reads back; `capacity` bounds the output."""
        b = _buf(stream)
        out = C.create_string_buffer(max(capacity, 1))
        rc = self._check(int(lib().fourmc_blockstream_decompress_host(self._h, int(zstd), b, len(b), out, capacity)))
        return out.raw[:rc]

    # ---- block index, splits, line records (FourMcBlockIndex / FourMcInputFormat / FourMcLineRecordReader) ----
    def read_index(self, stream_bytes) -> list[int]:
        """Block offsets from the footer index of a whole .4mc / .4mz file (FourMcInputStream.readIndex)."""
        b = _buf(stream_bytes)
        n = self._check(int(lib().fourmc_read_index_host(self._h, b, len(b), None, 0)))
        arr = (C.c_int64 * max(n, 1))()
        self._check(int(lib().fourmc_read_index_host(self._h, b, len(b), arr, n)))
        return list(arr[:n])

    def read_split_lines(self, stream_bytes, start: int, length: int) -> bytes:
        """The records FourMcLineRecordReader returns for the split [start, start + length), concatenated."""
        b = _buf(stream_bytes)
        cap = (length // 4 + 8 * 1024 * 1024) * 16
        while True:
            out = C.create_string_buffer(cap)
            rc = int(lib().fourmc_read_split_lines_host(self._h, b, len(b), start, length, out, cap))
            if rc == E_OUTPUT and cap < (1 << 36):
                cap *= 4
                continue
            self._check(rc)
            return out.raw[:rc]

    def read_splits_lines(self, stream_bytes, splits) -> list[bytes]:
        """fourmc_read_splits_lines_host: the records of every (start, length) split, all blocks decoded as one batch."""
        b = _buf(stream_bytes)
        n = len(splits)
        st = (C.c_int64 * max(n, 1))(*[s for s, _ in splits])
        ln = (C.c_int64 * max(n, 1))(*[l for _, l in splits])
        offs = (C.c_int64 * (n + 1))()
        cap = sum(l for _, l in splits) * 8 + (n + 4) * 3 * BLOCKSIZE
        while True:
            out = C.create_string_buffer(cap)
            rc = int(lib().fourmc_read_splits_lines_host(self._h, b, len(b), n, st, ln, out, cap, offs))
            if rc == E_OUTPUT and cap < (1 << 36):
                cap *= 4
                continue
            self._check(rc)
            return [out.raw[offs[i]:offs[i + 1]] for i in range(n)]

    # ---- device-resident calls: raw device pointers (ints) and an optional CUDA stream handle ----
    def gen_device(self, d_out: int, n_pages: int, seed: int = 0x4D43, first_page: int = 0, kind: int = 0, stream=None):
        self._check(lib().fourmc_gen_device(self._h, stream, kind, seed, first_page, n_pages, d_out))

    def compress_device(self, d_in: int, n: int, d_out: int, out_capacity: int, d_out_size: int,
                        d_block_lens: int | None = None, level: int = 1, stream=None):
        self._check(lib().fourmc_4mc_compress_device(self._h, stream, level, d_in, n, d_out, out_capacity,
                                                     d_out_size, d_block_lens))

    def compress_span_device(self, d_in: int, n: int, d_span: int, span_capacity: int, d_span_size: int,
                             d_block_lens: int | None = None, level: int = 1, stream=None):
        self._check(lib().fourmc_4mc_compress_span_device(self._h, stream, level, d_in, n, d_span, span_capacity,
                                                          d_span_size, d_block_lens))

    def build_index_device(self, d_block_lens: int, n_blocks: int, d_header: int | None, d_tail: int, stream=None):
        self._check(lib().fourmc_4mc_build_index_device(self._h, stream, d_block_lens, n_blocks, d_header, d_tail))

    def decompress_device(self, d_in: int, n: int, d_out: int, out_capacity: int, d_result: int, stream=None):
        self._check(lib().fourmc_4mc_decompress_device(self._h, stream, d_in, n, d_out, out_capacity, d_result))

    def decompress_range_device(self, d_in: int, n: int, first_block: int, n_blocks: int, d_out: int, out_capacity: int,
                                d_result: int, stream=None, zstd: bool = False):
        """Blocks [first_block, first_block + n_blocks) of ONE stream (a rank's share)."""
        f = lib().fourmc_4mz_decompress_range_device if zstd else lib().fourmc_4mc_decompress_range_device
        self._check(f(self._h, stream, d_in, n, first_block, n_blocks, d_out, out_capacity, d_result))

    def xxh32_batch_device(self, n_items: int, d_base: int, d_off: int, d_len: int, d_out: int, seed: int = 0, stream=None):
        self._check(lib().fourmc_xxh32_batch_device(self._h, stream, n_items, d_base, d_off, d_len, seed, d_out))

    def compress_host_ptr(self, in_ptr: int, n: int, out_ptr: int, cap: int, level: int = 1) -> int:
        """fourmc_4mc_compress_host on raw host addresses (e.g. pinned torch tensors)."""
        return self._check(lib().fourmc_4mc_compress_host(self._h, level, in_ptr, n, out_ptr, cap))

    def decompress_host_ptr(self, in_ptr: int, n: int, out_ptr: int, cap: int) -> int:
        return self._check(lib().fourmc_4mc_decompress_host(self._h, in_ptr, n, out_ptr, cap))


# ---- reference-shaped facades ---------------------------------------------------------------

class Lz4Compressor:
    """Mirror of com.fing.compression.fourmc.Lz4Compressor's native methods
    (Lz4Compressor.java:314-321; native/jniCompressor.c:72-194)."""

    def __init__(self, ctx: Context, level: int = 1):
        self.ctx, self.level = ctx, level

    @staticmethod
    def compress_bound(n: int) -> int:
        return int(lib().fourmc_lz4_compress_bound(n))

    def compress_bytes_direct(self, uncompressed) -> bytes:
        out = self.ctx.lz4_compress(uncompressed, self.level)
        if out is None:
            raise FourMcError(0, "LZ4_compress returned: 0")      # jniCompressor.c:95-100 InternalError
        return out

    def xxhash32(self, buf, off: int, length: int, seed: int) -> int:
        return self.ctx.xxh32(bytes(buf[off:off + length]), seed)


class Lz4Decompressor:
    """Mirror of Lz4Decompressor's natives (Lz4Decompressor.java:287-290; native/jniDecompressor.c:67-118)."""

    def __init__(self, ctx: Context, direct_buffer_size: int = BLOCKSIZE):
        self.ctx, self.direct_buffer_size = ctx, direct_buffer_size

    def decompress_bytes_direct(self, compressed) -> bytes:
        rc, out = self.ctx.lz4_decompress_safe(compressed, self.direct_buffer_size)
        if rc < 0:
            raise FourMcError(rc, f"LZ4_decompress_safe returned: {rc}")     # jniDecompressor.c:93-97
        return out

    def xxhash32(self, buf, off: int, length: int, seed: int) -> int:
        return self.ctx.xxh32(bytes(buf[off:off + length]), seed)


class ZstdCompressor:
    """Mirror of ZstdCompressor's natives (ZstdCompressor.java; native/jniZstdCompressor.c:59-199)."""

    def __init__(self, ctx: Context, level: int = 1):
        self.ctx, self.level = ctx, level

    @staticmethod
    def compress_bound(n: int) -> int:
        return int(lib().fourmc_zstd_compress_bound(n))

    def compress_bytes_direct(self, uncompressed) -> bytes:
        out = self.ctx.zstd_compress(uncompressed, self.level)
        if out is None:
            raise FourMcError(-70, "ZSTD_compress returned: dstSize_tooSmall")   # jniZstdCompressor.c:95-100
        return out

    def xxhash32(self, buf, off: int, length: int, seed: int) -> int:
        return self.ctx.xxh32(bytes(buf[off:off + length]), seed)


class ZstdDecompressor:
    """Mirror of ZstdDecompressor's natives (ZstdDecompressor.java; native/jniZstdDecompressor.c:67-120)."""

    def __init__(self, ctx: Context, direct_buffer_size: int = BLOCKSIZE):
        self.ctx, self.direct_buffer_size = ctx, direct_buffer_size

    def decompress_bytes_direct(self, compressed) -> bytes:
        rc, out = self.ctx.zstd_decompress(compressed, self.direct_buffer_size)
        if rc < 0:
            raise FourMcError(rc, f"ZSTD_decompress returned: {rc}")         # jniZstdDecompressor.c:95-99
        return out

    def xxhash32(self, buf, off: int, length: int, seed: int) -> int:
        return self.ctx.xxh32(bytes(buf[off:off + length]), seed)


class FourMzCodec:
    """Whole-stream 4mz codec: FourMzCodec.java / native/4mc.c:389-553 (writer), :709-857 (reader)."""

    def __init__(self, ctx: Context, level: int = 1):
        self.ctx, self.level = ctx, level

    def compress(self, data) -> bytes:
        return self.ctx.compress_4mz(data, self.level)

    def decompress(self, stream) -> bytes:
        return self.ctx.decompress_4mz(stream)


class FourMcCodec:
    """Whole-stream codec: FourMcCodec.java:82-168 / native/4mc.c:220,896."""

    def __init__(self, ctx: Context, level: int = 1):
        self.ctx, self.level = ctx, level

    def compress(self, data) -> bytes:
        return self.ctx.compress_4mc(data, self.level)

    def decompress(self, stream) -> bytes:
        return self.ctx.decompress_4mc(stream)


class FourMcBlockIndex:
    """Mirror of com.fing.compression.fourmc.FourMcBlockIndex (FourMcBlockIndex.java:92-173) over the C-ABI."""
    NOT_FOUND = -1

    def __init__(self, offsets):
        self.offsets = list(offsets)
        self._arr = (C.c_int64 * max(len(self.offsets), 1))(*self.offsets)

    def find_next_position(self, pos: int) -> int:
        return int(lib().fourmc_index_find_next_position(self._arr, len(self.offsets), pos))

    def find_belonging_block_index(self, pos: int) -> int:
        return int(lib().fourmc_index_find_belonging_block(self._arr, len(self.offsets), pos))

    def align_slice_start_to_index(self, start: int, end: int) -> int:
        return int(lib().fourmc_index_align_slice_start(self._arr, len(self.offsets), start, end))

    def align_slice_end_to_index(self, end: int, file_size: int) -> int:
        return int(lib().fourmc_index_align_slice_end(self._arr, len(self.offsets), end, file_size))

    def plan_splits(self, file_size: int, split_size: int) -> list[tuple[int, int]]:
        """FourMcInputFormat.getSplits for one file: [(start, length)] (FourMcInputFormat.java:126-173)."""
        n = int(lib().fourmc_plan_splits(self._arr, len(self.offsets), file_size, split_size, None, None, 0))
        if n < 0:
            raise FourMcError(n, "fourmc_plan_splits")
        st, ln = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
        lib().fourmc_plan_splits(self._arr, len(self.offsets), file_size, split_size, st, ln, n)
        return [(int(st[i]), int(ln[i])) for i in range(n)]


# ---- block sharding across ranks ---------------------------------------------

def shard_blocks(n_blocks: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block range [lo, hi) of `rank`: r gets [r*ceil(n/G), (r+1)*ceil(n/G))."""
    per = -(-n_blocks // world_size) if n_blocks else 0
    lo = min(n_blocks, rank * per)
    return lo, min(n_blocks, lo + per)


def shard_splits(n_splits: int, world_size: int, rank: int) -> list[int]:
    """Indices of the input splits `rank` reads: round-robin, so that every rank gets splits from all over the file
    (10 000 splits over 8 GPUs).  Splits are independent -- no collective on the read path."""
    return list(range(rank, n_splits, world_size))


def span_base_offsets(span_sizes: list[int]) -> list[int]:
    """File offset of every rank's span in the single output stream: rank r starts at
    12 (file header) + the spans of ranks < r (native/4mc.c:285-293 block offsets)."""
    out, pos = [], 12
    for s in span_sizes:
        out.append(pos)
        pos += s
    return out
