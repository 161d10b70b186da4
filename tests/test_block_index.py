"""TestFourMcBlockIndex (java/hadoop-4mc/src/test/java/com/fing/compression/fourmc/TestFourMcBlockIndex.java:20-84)
restated over the oracle's index functions -- the one thing the reference's own tests pin on this path."""
import ctypes as C


def _idx(*v):
    return (C.c_int64 * len(v))(*v), len(v)


def test_find_next_position(oracle):
    a, n = _idx(100, 200, 300, 400)
    f = oracle.fmo_index_find_next_position
    assert f(a, n, 100) == 100 and f(a, n, 110) == 200 and f(a, n, 210) == 300
    assert f(a, n, 401) == -1                     # NOT_FOUND beyond the last block


def test_find_belonging_block_index(oracle):
    a, n = _idx(100, 200, 300, 400)
    f = oracle.fmo_index_find_belonging_block
    assert f(a, n, 50) == -1
    assert [f(a, n, p) for p in (100, 110, 210, 300, 350, 400, 450)] == [0, 0, 1, 2, 2, 3, 3]


def test_align_slice(oracle):
    a, n = _idx(100, 200, 300, 400)
    assert oracle.fmo_index_align_slice_start(a, n, 0, 350) == 0
    assert oracle.fmo_index_align_slice_start(a, n, 100, 350) == 100
    assert oracle.fmo_index_align_slice_start(a, n, 310, 350) == -1
    assert oracle.fmo_index_align_slice_end(a, n, 350, 550) == 400
    assert oracle.fmo_index_align_slice_end(a, n, 250, 550) == 300
    assert oracle.fmo_index_align_slice_end(a, n, 450, 550) == 550
