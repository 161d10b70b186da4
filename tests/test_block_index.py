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


# ---- the product's own index functions (C-ABI of lib4mcgpu.so; no device needed) against the oracle ----

def test_product_index_matches_reference_test_cases(pkg):
    """The same TestFourMcBlockIndex cases through fourmc_index_* (FourMcBlockIndex.java:92-173)."""
    ix = pkg.FourMcBlockIndex([100, 200, 300, 400])
    assert [ix.find_next_position(p) for p in (100, 110, 210, 401)] == [100, 200, 300, -1]
    assert [ix.find_belonging_block_index(p) for p in (50, 100, 110, 210, 300, 350, 400, 450)] == [-1, 0, 0, 1, 2, 2, 3, 3]
    assert ix.align_slice_start_to_index(0, 350) == 0 and ix.align_slice_start_to_index(100, 350) == 100
    assert ix.align_slice_start_to_index(310, 350) == -1
    assert [ix.align_slice_end_to_index(e, 550) for e in (350, 250, 450)] == [400, 300, 550]
    assert pkg.FourMcBlockIndex([]).find_next_position(5) == -1 and pkg.FourMcBlockIndex([]).find_belonging_block_index(5) == -1


def test_product_index_matches_oracle_on_random_indexes(pkg, oracle):
    import random
    rng = random.Random(3)
    for _ in range(200):
        n = rng.randint(1, 40)
        offs, cur = [], 12
        for _i in range(n):
            offs.append(cur)
            cur += 12 + rng.randint(1, 5000)
        file_size = cur + 12 + 20 + 4 * n
        arr = (C.c_int64 * n)(*offs)
        ix = pkg.FourMcBlockIndex(offs)
        for _q in range(60):
            p = rng.choice(offs) + rng.choice((-1, 0, 1)) if rng.random() < 0.5 else rng.randint(0, file_size + 10)
            e = p + rng.randint(0, 20000)
            assert ix.find_next_position(p) == oracle.fmo_index_find_next_position(arr, n, p)
            assert ix.find_belonging_block_index(p) == oracle.fmo_index_find_belonging_block(arr, n, p)
            assert ix.align_slice_start_to_index(p, e) == oracle.fmo_index_align_slice_start(arr, n, p, e)
            assert ix.align_slice_end_to_index(p, file_size) == oracle.fmo_index_align_slice_end(arr, n, p, file_size)


def test_plan_splits_follows_the_input_format(pkg, oracle):
    """FourMcInputFormat.getSplits (FourMcInputFormat.java:126-173) over Hadoop's default splits
    (FileInputFormat: pieces of split_size while remaining / split_size > 1.1, then the rest)."""
    import random
    rng = random.Random(9)
    for _ in range(100):
        n = rng.randint(0, 60)
        offs, cur = [], 12
        for _i in range(n):
            offs.append(cur)
            cur += 12 + rng.randint(1, 3000)
        file_size = cur + 12 + 20 + 4 * n
        split = rng.randint(500, 40000)
        arr = (C.c_int64 * max(n, 1))(*offs)
        want, rem, pos = [], file_size, 0
        pieces = []
        while rem / split > 1.1:
            pieces.append((pos, split)); pos += split; rem -= split
        if rem:
            pieces.append((pos, rem))
        for s, ln in pieces:
            if n == 0:
                want.append((s, ln)); continue
            a = oracle.fmo_index_align_slice_start(arr, n, s, s + ln)
            b = oracle.fmo_index_align_slice_end(arr, n, s + ln, file_size)
            if a != -1 and b != -1:
                want.append((a, b - a))
        got = pkg.FourMcBlockIndex(offs).plan_splits(file_size, split)
        assert got == want
        if n:
            # every block start belongs to exactly one split
            owners = [sum(1 for a, ln in got if a <= o < a + ln) for o in offs]
            assert owners == [1] * n
