"""Shared fixtures.  `-m "not gpu"` runs on a CPU box (oracle, goldens, host logic, ABI surface);
`-m gpu` holds the parity tests proper, all of which call through the C-ABI of lib4mcgpu.so."""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("4mc_b200")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on oracle/_build/liboracle.so (the CPU restatement; test infrastructure)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "oracle"], check=True)
    O = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
    O.fmo_xxh32.restype = C.c_uint32
    O.fmo_xxh32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    O.fmo_lz4_decompress_safe.restype = C.c_int
    O.fmo_lz4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    O.fmo_lz4_compress.restype = C.c_int
    O.fmo_lz4_compress.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    O.fmo_lz4_compress_bound.restype = C.c_int
    O.fmo_4mc_bound.restype = C.c_size_t
    O.fmo_4mc_bound.argtypes = [C.c_size_t]
    O.fmo_4mc_compress.restype = C.c_longlong
    O.fmo_4mc_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    O.fmo_4mc_decompress.restype = C.c_longlong
    O.fmo_4mc_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    O.fmo_zstd_decompress.restype = C.c_longlong
    O.fmo_zstd_decompress.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_longlong]
    O.fmo_4mz_decompress.restype = C.c_longlong
    O.fmo_4mz_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    O.fmo_4mc_read_index.restype = C.c_longlong
    O.fmo_4mc_read_index.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_int64), C.c_size_t]
    for f in ("fmo_index_find_next_position", "fmo_index_find_belonging_block"):
        getattr(O, f).restype = C.c_int64
        getattr(O, f).argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int64]
    for f in ("fmo_index_align_slice_start", "fmo_index_align_slice_end"):
        getattr(O, f).restype = C.c_int64
        getattr(O, f).argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int64, C.c_int64]
    return O


class OracleApi:
    """Convenience wrappers over the oracle used by several tests."""

    def __init__(self, O):
        self.O = O

    def xxh32(self, b, seed=0):
        return self.O.fmo_xxh32(bytes(b), len(b), seed)

    def lz4_decompress(self, src, cap):
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = self.O.fmo_lz4_decompress_safe(bytes(src), out, len(src), cap)
        return r, out.raw[:max(r, 0)]

    def lz4_compress(self, src):
        cap = len(src) + len(src) // 255 + 64
        out = C.create_string_buffer(cap)
        r = self.O.fmo_lz4_compress(bytes(src), out, len(src), cap)
        return out.raw[:r]

    def compress_4mc(self, data):
        cap = self.O.fmo_4mc_bound(len(data))
        out = C.create_string_buffer(cap)
        r = self.O.fmo_4mc_compress(bytes(data), len(data), out, cap)
        assert r > 0
        return out.raw[:r]

    def decompress_4mc(self, stream, cap):
        out = C.create_string_buffer(max(cap, 1))
        r = self.O.fmo_4mc_decompress(bytes(stream), len(stream), out, cap)
        return r, out.raw[:max(r, 0)]

    def zstd_decompress(self, frame, cap):
        """oracle/zstd_oracle.c: strict Zstandard frame decoder"""
        out = C.create_string_buffer(max(cap, 1) + 64)
        r = self.O.fmo_zstd_decompress(out, cap, bytes(frame), len(frame))
        return r, out.raw[:max(r, 0)]

    def decompress_4mz(self, stream, cap):
        out = C.create_string_buffer(max(cap, 1))
        r = self.O.fmo_4mz_decompress(bytes(stream), len(stream), out, cap)
        return r, out.raw[:max(r, 0)]


@pytest.fixture(scope="session")
def ora(oracle):
    return OracleApi(oracle)


@pytest.fixture(scope="session")
def ref():
    """The reference itself (oracle/_ref, built from /root/reference by oracle/Makefile); tests
    that need it are skipped where it was never built."""
    p = os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(p)
    R.XXH32.restype = C.c_uint32
    R.XXH32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    R.LZ4_compress_default.restype = C.c_int
    R.LZ4_decompress_safe.restype = C.c_int
    return R


@pytest.fixture(scope="session")
def ref_cli():
    p = os.path.join(ROOT, "oracle", "_ref", "4mc")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built")
    return p


def golden_bytes(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def lcg_buffer(n=1 << 20):
    """s=12345; s=s*1103515245+12345; b[i]=(s>>16)&0xFF"""
    import numpy as np
    out = np.empty(n, dtype=np.uint8)
    s = 12345
    for i in range(n):
        s = (s * 1103515245 + 12345) & 0xFFFFFFFF
        out[i] = (s >> 16) & 0xFF
    return out.tobytes()


@pytest.fixture(scope="session")
def lcg():
    return lcg_buffer()


def gen_logtext(pkg, nbytes, seed=0x4D43, first_page=0):
    pages = (nbytes + 4095) // 4096
    buf = C.create_string_buffer(max(pages, 1) * 4096)
    assert pkg.lib().fourmc_gen_host(0, seed, first_page, pages, buf) == 0
    return buf.raw[:nbytes]


def build_native(name, sources, extra=(), deps=()):
    """Compiles a test-only native helper into tests/_build/<name>.so with g++ (`deps`: headers that
    also trigger a rebuild)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".so")
    srcs = [os.path.join(ROOT, s) for s in sources]
    watch = srcs + [os.path.join(ROOT, d) for d in deps]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in watch):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out] + srcs + list(extra), check=True)
    return C.CDLL(out)


@pytest.fixture(scope="session")
def ctx(pkg):
    """A fourmc context on cuda:0 -- the GPU tests fail (not skip) if the library cannot get one."""
    c = pkg.Context(0)
    yield c
    c.close()


def _be32(v):
    return int(v).to_bytes(4, "big")


def make_4mc(ora, records, magic=b"4MC\0"):
    """A .4mc / .4mz stream from (announced_usize, payload) records (hand-built edge cases: the container
    fields as native/4mc.c:263-362 writes them, checksums from the oracle's XXH32)."""
    hdr = magic + _be32(1)
    out = hdr + _be32(ora.xxh32(hdr))
    lens = []
    for usize, payload in records:
        out += _be32(usize) + _be32(len(payload)) + _be32(ora.xxh32(payload)) + payload
        lens.append(12 + len(payload))
    out += bytes(12)
    fsize = 20 + 4 * len(records)
    foot = _be32(fsize) + _be32(1)
    for i in range(len(records)):
        foot += _be32(12 if i == 0 else lens[i - 1])
    foot += _be32(fsize) + magic
    return out + foot + _be32(ora.xxh32(foot))


def short_block_stream(ora):
    """Three blocks, the middle one announcing 100 bytes while its LZ4 payload (a 52-byte literal-only block)
    decodes to 50: the reference writes what the decoder returned (native/4mc.c:661-666)."""
    a = bytes(range(65, 91)) * 40                       # 1040 bytes, compressible
    lits = bytes((i * 7 + 3) & 0xFF for i in range(50))
    short = bytes([0xF0, 50 - 15]) + lits                # token: 15+ literals, no match
    c = bytes(reversed(a))
    return make_4mc(ora, [(len(a), ora.lz4_compress(a)), (100, short), (len(c), c)]), a + lits + c
