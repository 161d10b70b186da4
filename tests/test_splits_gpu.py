"""Per-split reading (BASELINE.json configs[4]): FourMcInputFormat's split planner plus
FourMcLineRecordReader's record rules (java/hadoop-4mc/src/main/java/com/fing/mapreduce/
FourMcInputFormat.java:126-173, FourMcLineRecordReader.java:116-163) with the split's blocks decoded on the GPU.
Property: over the splits of a file every line of the input comes back exactly once, in order."""
import ctypes as C

import pytest

from conftest import gen_logtext, golden_bytes

pytestmark = pytest.mark.gpu
MIB = 1024 * 1024


def make_input(pkg):
    text = gen_logtext(pkg, 23 * MIB + 4567)
    text = text[:text.rfind(b"\n") + 1]
    long_line = b"L" * (9 * MIB + 13) + b"\n"                  # one record spanning three blocks
    return text[:10 * MIB] + long_line + text[10 * MIB:] + b"last line without terminator"


@pytest.mark.parametrize("codec", ["4mc", "4mz"])
def test_splits_return_every_line_once(ctx, pkg, oracle, codec):
    data = make_input(pkg)
    stream = ctx.compress_4mc(data) if codec == "4mc" else ctx.compress_4mz(data)
    offs = ctx.read_index(stream)
    n = len(offs)
    assert n == (len(data) + 4 * MIB - 1) // (4 * MIB) and offs[0] == 12
    arr = (C.c_int64 * n)()
    magic = 0x344D4300 if codec == "4mc" else 0x344D5A00
    assert oracle.fmo_4mc_read_index(stream, len(stream), magic, arr, n) == n and list(arr) == offs
    ix = pkg.FourMcBlockIndex(offs)
    for split_size in (len(stream), 5 * MIB + 17, 2 * MIB, 700 * 1024):
        splits = ix.plan_splits(len(stream), split_size)
        got = b"".join(ctx.read_split_lines(stream, s, ln) for s, ln in splits)
        assert got == data, (codec, split_size, len(splits))
    # a split without any block start yields nothing; a damaged footer is reported
    assert ctx.read_split_lines(stream, offs[1] + 1, 10) == b""
    bad = bytearray(stream); bad[-1] ^= 1
    with pytest.raises(pkg.FourMcError):
        ctx.read_index(bytes(bad))


def test_split_of_reference_written_file(ctx, pkg):
    stream = golden_bytes("logtext_128k.l3.4mc")
    want = golden_bytes("logtext_128k.bin")
    offs = ctx.read_index(stream)
    assert offs == [12]
    assert ctx.read_split_lines(stream, 0, len(stream)) == want
    assert ctx.read_index(golden_bytes("empty.4mc")) == []
    assert ctx.read_split_lines(golden_bytes("empty.4mc"), 0, 44) == b""
