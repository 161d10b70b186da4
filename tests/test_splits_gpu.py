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


def test_line_terminators_like_hadoop_line_reader(ctx, pkg):
    """LF, CR and CR LF all end a line (Hadoop's LineReader, behind FourMcLineRecordReader.java:136-163): every split
    against a model of the reader's rules on the decoded bytes -- skip the first line unless the split starts the
    file, finish the line that is open at the split's end."""
    import random
    rng = random.Random(9)
    words = gen_logtext(pkg, 64 * 1024).split(b"\n")
    parts = []
    size = 0
    while size < 13 * MIB:
        ln = words[rng.randrange(len(words))] + rng.choice([b"\n", b"\r\n", b"\r", b"\n"])
        parts.append(ln); size += len(ln)
    data = b"".join(parts)
    stream = ctx.compress_4mc(data)
    offs = ctx.read_index(stream)
    ix = pkg.FourMcBlockIndex(offs)

    def eol_end(pos):                                   # end of the first terminator at or after pos, None if there is none
        a, b = data.find(b"\n", pos), data.find(b"\r", pos)
        cands = [x for x in (a, b) if x >= 0]
        if not cands:
            return None
        i = min(cands)
        return i + 2 if data[i:i + 2] == b"\r\n" else i + 1

    for split_size in (3 * MIB, 1 * MIB + 11):
        splits = ix.plan_splits(len(stream), split_size)
        got_all = b""
        for s, ln in splits:
            b0 = offs.index(s) if s else 0
            b1 = len([o for o in offs if o < s + ln])
            u0, u1 = b0 * 4 * MIB, min(len(data), b1 * 4 * MIB)
            frm = 0 if s == 0 else eol_end(u0)
            to = len(data) if u1 >= len(data) else (eol_end(u1) or len(data))
            want = b"" if frm is None or to <= frm else data[frm:to]
            got = ctx.read_split_lines(stream, s, ln)
            assert got == want, (split_size, s, ln, len(got), len(want))
            got_all += got
        assert got_all == data


@pytest.mark.parametrize("codec", ["4mc", "4mz"])
def test_many_splits_in_one_call(ctx, pkg, codec):
    """fourmc_read_splits_lines_host: the blocks of all the splits decoded as one batch; per split the same records as
    the single-split call (including the three-block line, which takes the general path inside)."""
    data = make_input(pkg)
    stream = ctx.compress_4mc(data) if codec == "4mc" else ctx.compress_4mz(data)
    ix = pkg.FourMcBlockIndex(ctx.read_index(stream))
    for split_size in (2 * MIB, 700 * 1024, len(stream)):
        splits = ix.plan_splits(len(stream), split_size)
        one_by_one = [ctx.read_split_lines(stream, s, ln) for s, ln in splits]
        assert ctx.read_splits_lines(stream, splits) == one_by_one
        assert b"".join(one_by_one) == data
        picked = splits[::-2]                                   # any subset, any order
        assert ctx.read_splits_lines(stream, picked) == one_by_one[::-2]
    assert ctx.read_splits_lines(stream, []) == []


def test_split_of_reference_written_file(ctx, pkg):
    stream = golden_bytes("logtext_128k.l3.4mc")
    want = golden_bytes("logtext_128k.bin")
    offs = ctx.read_index(stream)
    assert offs == [12]
    assert ctx.read_split_lines(stream, 0, len(stream)) == want
    assert ctx.read_index(golden_bytes("empty.4mc")) == []
    assert ctx.read_split_lines(golden_bytes("empty.4mc"), 0, 44) == b""
