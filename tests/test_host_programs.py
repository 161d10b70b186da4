"""The host programs on top of the C-ABI: the `4mc` CLI (flags, exit codes, files the reference CLI can
read and vice versa) and the JNI library driven through a fake JNIEnv (tests/native/jni_mock.c)."""
import ctypes as C
import os
import subprocess

import pytest

from conftest import ROOT, golden_bytes, golden_json, gen_logtext

HOST = os.path.join(ROOT, "4mc_b200", "host")
CLI = os.path.join(HOST, "4mc")
JNI = os.path.join(HOST, "libhadoop-4mc.so")


@pytest.fixture(scope="module", autouse=True)
def build_host(pkg):
    pkg.build()
    subprocess.run(["make", "-C", HOST, "-s"], check=True)


def test_jni_library_exports_the_reference_symbols():
    """Exactly the 35 Java_* symbols of the shipped reference library (tests/golden/jni_symbols.txt,
    from `nm -D` of java/hadoop-4mc/src/main/resources/.../linux/amd64/libhadoop-4mc.so)."""
    want = open(os.path.join(ROOT, "tests", "golden", "jni_symbols.txt")).read().split()
    out = subprocess.run(["nm", "-D", JNI], capture_output=True, text=True, check=True).stdout
    have = sorted(l.split()[2] for l in out.splitlines() if " T Java_" in l)
    assert have == sorted(want) and len(have) == 35


def test_jni_table_layout_matches_the_jni_specification():
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", os.path.join(ROOT, "tests", "_build", "jni_mock.so"),
                    os.path.join(ROOT, "tests", "native", "jni_mock.c"), "-ldl"], check=True)
    M = C.CDLL(os.path.join(ROOT, "tests", "_build", "jni_mock.so"))
    offs = (C.c_int * 13)()
    M.mock_slot_offsets(offs)
    # slots 6, 14, 23, 94, 95, 100, 101, 109, 110, 167, 222, 223, 230, 8 bytes each
    assert list(offs) == [8 * s for s in (6, 14, 23, 94, 95, 100, 101, 109, 110, 167, 222, 223, 230)]


def test_cli_fails_loudly_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = tmp_path / "in.bin"
    p.write_bytes(b"hello")
    r = subprocess.run([CLI, "-f", str(p), str(tmp_path / "o.4mc")], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr
    assert subprocess.run([CLI, "-V"], capture_output=True).returncode == 0
    assert subprocess.run([CLI, "-x"], capture_output=True).returncode == 1          # bad usage -> exit 1


REF_CLI_ON_US = os.path.join(ROOT, "tests", "_build", "ref_4mccli_on_lib4mcgpu")


def test_reference_cli_source_links_against_this_library():
    """native/4mc.h:36-41: the four entry points the reference's own CLI calls are exported, so its
    native/4mccli.c -- unmodified, compiled where it lies -- links against lib4mcgpu.so (the binary travels to
    the GPU box and is run there by test_reference_cli_binary_runs_on_this_library)."""
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "4mc_b200", "lib4mcgpu.so")], capture_output=True, text=True, check=True).stdout
    have = {l.split()[2] for l in out.splitlines() if " T " in l}
    assert {"fourMCcompressFilename", "fourMcDecompressFileName", "fourMZcompressFilename", "fourMZDecompressFileName"} <= have
    ref_src = "/root/reference/native/4mccli.c"
    if not os.path.exists(ref_src):
        pytest.skip("reference sources not present")
    os.makedirs(os.path.dirname(REF_CLI_ON_US), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-w", "-I/root/reference/native", "-o", REF_CLI_ON_US, ref_src,
                    "-L" + os.path.join(ROOT, "4mc_b200"), "-l4mcgpu", "-Wl,-rpath," + os.path.join(ROOT, "4mc_b200"),
                    "-Wl,-rpath,$ORIGIN/../../4mc_b200"], check=True)
    assert subprocess.run([REF_CLI_ON_US, "-V"], capture_output=True).returncode == 0


# ---------------------------------------------------------------- on the GPU box

@pytest.mark.gpu
def test_reference_cli_binary_runs_on_this_library(ref_cli, pkg, tmp_path):
    """The reference's 4mccli.c linked against lib4mcgpu.so (built by the CPU test above): round trips of both
    containers, checked against the reference's own build."""
    if not os.path.exists(REF_CLI_ON_US):
        pytest.skip("tests/_build/ref_4mccli_on_lib4mcgpu was not built (no reference sources on the build box)")
    data = gen_logtext(pkg, 5 * 1024 * 1024 + 77)
    src = tmp_path / "d.bin"
    src.write_bytes(data)
    for flags, ext in (([], ".4mc"), (["-z"], ".4mz"), (["-3"], ".4mc")):
        comp, back, back2 = tmp_path / ("c" + ext), tmp_path / "back", tmp_path / "back2"
        assert subprocess.run([REF_CLI_ON_US, "-f", "-q", "-q"] + flags + [str(src), str(comp)]).returncode == 0
        zd = ["-z"] if ext == ".4mz" else []
        subprocess.run([ref_cli, "-f", "-q", "-q", "-d"] + zd + [str(comp), str(back)], check=True)
        assert subprocess.run([REF_CLI_ON_US, "-f", "-q", "-q", "-d"] + zd + [str(comp), str(back2)]).returncode == 0
        assert back.read_bytes() == data and back2.read_bytes() == data


@pytest.mark.gpu
def test_files_stream_through_in_slices(ref_cli, pkg, tmp_path):
    """File -> file in slices of many blocks through pinned bounce buffers (4mc_b200/csrc/fileio.h): an input of
    several slices plus a ragged tail, pipes on both ends, two concatenated streams, and a damaged block in the
    second slice -- whose predecessors reach the output before the error, like the serial reader."""
    n = 331 * 1024 * 1024 + 12345                            # > 2 writer slices (128 MiB), > 2 reader chunks (160 MiB)
    data = gen_logtext(pkg, n)
    src, ours, theirs = tmp_path / "big.bin", tmp_path / "big.4mc", tmp_path / "ref.4mc"
    src.write_bytes(data)
    assert subprocess.run([CLI, "-f", "-q", str(src), str(ours)]).returncode == 0
    subprocess.run([ref_cli, "-f", "-q", "-q", "-d", str(ours), str(tmp_path / "o1")], check=True)
    assert (tmp_path / "o1").read_bytes() == data
    subprocess.run([ref_cli, "-f", "-q", "-q", str(src), str(theirs)], check=True)
    assert subprocess.run([CLI, "-f", "-q", "-d", str(theirs), str(tmp_path / "o2")]).returncode == 0
    assert (tmp_path / "o2").read_bytes() == data
    # pipes: stdin -> stdout, both directions
    r = subprocess.run([CLI, "-q", "-c"], stdin=open(src, "rb"), capture_output=True)
    assert r.returncode == 0
    r2 = subprocess.run([CLI, "-q", "-c", "-d"], input=r.stdout, capture_output=True)
    assert r2.returncode == 0 and r2.stdout == data
    # two streams back to back (native/4mc.c:909-913)
    cat = tmp_path / "two.4mc"
    cat.write_bytes(ours.read_bytes() + golden_bytes("A.4mc"))
    assert subprocess.run([CLI, "-f", "-q", "-d", str(cat), str(tmp_path / "o3")]).returncode == 0
    assert (tmp_path / "o3").read_bytes() == data + b"A"
    # damage far into the file: same exit code and same partial output as the reference
    blob = bytearray(theirs.read_bytes())
    blob[len(blob) * 3 // 4] ^= 0x40
    bad = tmp_path / "bad.4mc"
    bad.write_bytes(bytes(blob))
    a = subprocess.run([CLI, "-f", "-q", "-d", str(bad), str(tmp_path / "p1")], capture_output=True).returncode
    b = subprocess.run([ref_cli, "-f", "-q", "-q", "-d", str(bad), str(tmp_path / "p2")], capture_output=True).returncode
    assert a == b == 4
    assert (tmp_path / "p1").read_bytes() == (tmp_path / "p2").read_bytes()

@pytest.mark.gpu
def test_cli_round_trip_and_cross_compatibility(ref_cli, pkg, tmp_path):
    data = gen_logtext(pkg, 9 * 1024 * 1024 + 1234) + bytes(70000)
    src = tmp_path / "d.bin"
    src.write_bytes(data)
    ours, theirs = tmp_path / "ours.4mc", tmp_path / "theirs.4mc"
    assert subprocess.run([CLI, "-f", "-q", "-1", str(src), str(ours)]).returncode == 0
    subprocess.run([ref_cli, "-f", "-q", "-q", "-1", str(src), str(theirs)], check=True)
    out1, out2, out3 = tmp_path / "o1", tmp_path / "o2", tmp_path / "o3"
    subprocess.run([ref_cli, "-f", "-q", "-q", "-d", str(ours), str(out1)], check=True)      # reference reads ours
    assert subprocess.run([CLI, "-f", "-q", "-d", str(theirs), str(out2)]).returncode == 0   # we read the reference's
    assert subprocess.run([CLI, "-f", "-q", "-d", str(ours), str(out3)]).returncode == 0
    assert out1.read_bytes() == data and out2.read_bytes() == data and out3.read_bytes() == data
    # -t (test) and stdin/stdout plumbing
    assert subprocess.run([CLI, "-q", "-t", str(ours)]).returncode == 0
    r = subprocess.run([CLI, "-q", "-c", "-d", str(theirs)], capture_output=True)
    assert r.returncode == 0 and r.stdout == data


@pytest.mark.gpu
def test_cli_exit_codes_match_reference(ref_cli, tmp_path):
    good = golden_bytes("logtext_128k.l1.4mc")
    cases = {}
    b = bytearray(good); b[100] ^= 1; cases["payload"] = bytes(b)
    b = bytearray(good); b[9] ^= 1; cases["header"] = bytes(b)
    b = bytearray(good); b[-1] ^= 1; cases["footer"] = bytes(b)
    cases["truncated"] = good[:20]
    cases["magic"] = good.replace(b"4MC\0", b"4MZ\0", 1)
    for name, blob in cases.items():
        p = tmp_path / (name + ".4mc")
        p.write_bytes(blob)
        ours = subprocess.run([CLI, "-f", "-q", "-d", str(p), str(tmp_path / "x")], capture_output=True).returncode
        ref = subprocess.run([ref_cli, "-f", "-q", "-d", str(p), str(tmp_path / "y")], capture_output=True).returncode
        assert ours == ref, name


@pytest.mark.gpu
def test_jni_shim_through_fake_jvm(ora, pkg):
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-o", os.path.join(ROOT, "tests", "_build", "jni_mock.so"),
                    os.path.join(ROOT, "tests", "native", "jni_mock.c"), "-ldl"], check=True)
    M = C.CDLL(os.path.join(ROOT, "tests", "_build", "jni_mock.so"))
    assert M.mock_open(JNI.encode()) == 0
    assert M.mock_bound(4 * 1024 * 1024) == 4210768
    data = gen_logtext(pkg, 4 * 1024 * 1024)
    msg = C.create_string_buffer(512)
    la, threw = C.c_int(), C.c_int()
    for which, lvl in ((0, 0), (1, 0), (2, 4), (2, 8)):
        src = C.create_string_buffer(data, len(data))
        dst = C.create_string_buffer(4210768)
        r = M.mock_compress(which, lvl, src, len(data), dst, C.byref(la), C.byref(threw), msg)
        assert r > 0 and threw.value == 0 and la.value == 0            # uncompressedDirectBufLen reset (jniCompressor.c:94)
        assert ora.lz4_decompress(dst.raw[:r], len(data)) == (len(data), data)
        out = C.create_string_buffer(4 * 1024 * 1024)
        comp = C.create_string_buffer(dst.raw[:r], r)
        d = M.mock_decompress(comp, r, out, 4 * 1024 * 1024, C.byref(la), C.byref(threw), msg)
        assert d == len(data) and out.raw[:d] == data and la.value == 0 and threw.value == 0
    # corrupt block -> InternalError with the reference's message format (jniDecompressor.c:93-97)
    bad = C.create_string_buffer(bytes.fromhex("1041 0900 C0") + b"0123456789ab", 17)
    out = C.create_string_buffer(64)
    d = M.mock_decompress(bad, 17, out, 17, C.byref(la), C.byref(threw), msg)
    assert d == -5 and threw.value == 1 and msg.value == b"java/lang/InternalError: LZ4_decompress_safe returned: -5"
    # xxhash32(byte[] buf, int off, int len, int seed) on all four classes
    buf = C.create_string_buffer(b"xx" + b"Nobody inspects the spammish repetition" + b"yy")
    for cls in range(4):
        assert M.mock_xxh(cls, buf, 2, 39, 0) & 0xFFFFFFFF == 0xE2293B2F
    # ZstdDecompressor.decompressBytesDirect (native/jniZstdDecompressor.c:68-101) on reference-made frames
    frame = next(z for z in golden_json("zstd_decode.json") if len(z["hex"]) > 4000 and z["runs"][0][1] > 0)
    comp = bytes.fromhex(frame["hex"])
    cap, ret, xxh = next(r for r in frame["runs"] if r[1] > 0)
    out = C.create_string_buffer(4 * 1024 * 1024)
    d = M.mock_zstd_decompress(C.create_string_buffer(comp, len(comp)), len(comp), out, 4 * 1024 * 1024, C.byref(la), C.byref(threw), msg)
    assert d == ret and ora.xxh32(out.raw[:d]) == xxh and la.value == 0 and threw.value == 0
    bad = comp[:len(comp) // 2]
    d = M.mock_zstd_decompress(C.create_string_buffer(bad, len(bad)), len(bad), out, 4 * 1024 * 1024, C.byref(la), C.byref(threw), msg)
    assert d < 0 and threw.value == 1 and msg.value == b"java/lang/InternalError: LZ4_decompress_safe returned: %d" % d
    # ZstdCompressor natives (native/jniZstdCompressor.c:72-172): frames the reference-shaped decoder restores
    assert M.mock_zstd_bound(4 * 1024 * 1024) == 4 * 1024 * 1024 + 16384
    for which, lvl in ((0, 0), (1, 0), (2, 6), (2, 12)):
        src = C.create_string_buffer(data, len(data))
        dst = C.create_string_buffer(4 * 1024 * 1024 + 16384)
        r = M.mock_zstd_compress(which, lvl, src, len(data), dst, C.byref(la), C.byref(threw), msg)
        assert 0 < r < len(data) // 2 and threw.value == 0 and la.value == 0
        out = C.create_string_buffer(4 * 1024 * 1024)
        d = M.mock_zstd_decompress(C.create_string_buffer(dst.raw[:r], r), r, out, 4 * 1024 * 1024, C.byref(la), C.byref(threw), msg)
        assert d == len(data) and out.raw[:d] == data and threw.value == 0
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref4mc.so")):
            R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref4mc.so"))
            R.ZSTD_decompress.restype = C.c_size_t
            R.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
            o2 = C.create_string_buffer(len(data))
            assert R.ZSTD_decompress(o2, len(data), dst.raw[:r], r) == len(data) and o2.raw == data
    # the streaming zstd natives resolve but are not built
    assert M.mock_zstd_stream_throws(msg) == 1 and b"InternalError" in msg.value


@pytest.mark.gpu
def test_cli_levels_like_the_reference(ref_cli, pkg, tmp_path):
    """-1 .. -4 and -z (native/4mccli.c:170-361, native/4mc.c:243-253 / :415-425): every level's output is read
    by the reference CLI, and the higher levels are smaller."""
    data = gen_logtext(pkg, 6 * 1024 * 1024 + 77, first_page=31)
    src = tmp_path / "d.bin"
    src.write_bytes(data)
    for z in ([], ["-z"]):
        sizes = []
        for lv in ("-1", "-2", "-3", "-4"):
            out, back = tmp_path / "o.bin", tmp_path / "b.bin"
            assert subprocess.run([CLI, "-f", "-q"] + z + [lv, str(src), str(out)]).returncode == 0
            subprocess.run([ref_cli, "-f", "-q", "-q"] + z + ["-d", str(out), str(back)], check=True)
            assert back.read_bytes() == data
            sizes.append(out.stat().st_size)
        assert sizes[3] < sizes[2] < sizes[1] < sizes[0], sizes
