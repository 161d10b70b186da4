"""World-size-2 test of the sharded writer's host logic on CPU (gloo): contiguous block ranges per
rank, one all-gather of the block lengths, every rank writes its span at its scanned offset, rank 0
adds header + EOS + footer.  The block records come from the oracle here (no GPU); on the GPU box the
same plumbing runs in bench.py with the CUDA kernels and NCCL."""
import importlib
import os
import struct
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = 4 * 1024 * 1024


def _worker(rank, world, port, path, n_bytes):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest
    import ctypes as C
    pkg = importlib.import_module("4mc_b200")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
    O.fmo_lz4_compress.restype = C.c_int
    O.fmo_xxh32.restype = C.c_uint32
    O.fmo_xxh32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    data = conftest.gen_logtext(pkg, n_bytes)
    nb = (n_bytes + BLOCK - 1) // BLOCK
    lo, hi = pkg.shard_blocks(nb, world, rank)
    # this rank's span: block records back to back
    span = bytearray()
    lens = []
    for b in range(lo, hi):
        blk = data[b * BLOCK:(b + 1) * BLOCK]
        out = C.create_string_buffer(len(blk))
        c = O.fmo_lz4_compress(blk, out, len(blk), len(blk) - 1)
        payload = out.raw[:c] if c > 0 else blk
        span += struct.pack(">III", len(blk), len(payload), O.fmo_xxh32(payload, len(payload), 0)) + payload
        lens.append(12 + len(payload))
    # the one exchange: block lengths (padded to the largest shard) and span sizes
    per = -(-nb // world)
    mine = torch.zeros(per, dtype=torch.int64)
    mine[:len(lens)] = torch.tensor(lens, dtype=torch.int64)
    gathered = [torch.zeros(per, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    all_lens = []
    for r in range(world):
        l, h = pkg.shard_blocks(nb, world, r)
        all_lens += gathered[r][:h - l].tolist()
    sizes = [sum(gathered[r].tolist()) for r in range(world)]
    base = pkg.span_base_offsets(sizes)[rank]
    fd = os.open(path, os.O_RDWR | os.O_CREAT)
    os.pwrite(fd, bytes(span), base)
    if rank == 0:
        hdr = struct.pack(">II", 0x344D4300, 1)
        os.pwrite(fd, hdr + struct.pack(">I", O.fmo_xxh32(hdr, 8, 0)), 0)
        fsize = 20 + 4 * nb
        deltas = [12] + all_lens[:-1] if nb else []
        foot = struct.pack(">II", fsize, 1) + b"".join(struct.pack(">I", d) for d in deltas) + struct.pack(">II", fsize, 0x344D4300)
        foot += struct.pack(">I", O.fmo_xxh32(foot, len(foot), 0))
        os.pwrite(fd, bytes(12) + foot, 12 + sum(sizes))
    os.close(fd)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bytes", [3 * BLOCK + 12345, 5 * BLOCK])
def test_two_rank_sharded_writer(oracle, ora, pkg, tmp_path, n_bytes):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "sharded.4mc")
    mp.spawn(_worker, args=(2, port, path, n_bytes), nprocs=2, join=True)
    stream = open(path, "rb").read()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import gen_logtext
    data = gen_logtext(pkg, n_bytes)
    r, out = ora.decompress_4mc(stream, n_bytes)
    assert r == n_bytes and out == data
    # the footer index points at every block header
    import ctypes as C
    offs = (C.c_int64 * 16)()
    nb = oracle.fmo_4mc_read_index(stream, len(stream), 0x344D4300, offs, 16)
    assert nb == (n_bytes + BLOCK - 1) // BLOCK
    for i in range(nb):
        assert struct.unpack(">I", stream[offs[i]:offs[i] + 4])[0] == min(BLOCK, n_bytes - i * BLOCK)


def _reader_worker(rank, world, port, path, split_size, out_dir):
    """Config 5 plumbing on CPU: every rank plans the same splits from the footer index, takes its round-robin share
    and reads the blocks of its splits (oracle LZ4 here, the GPU batch decoder on the box); no collective on the data
    path -- the ranks only meet at the end to compare notes."""
    sys.path.insert(0, ROOT)
    import ctypes as C
    pkg = importlib.import_module("4mc_b200")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
    O.fmo_4mc_read_index.restype = C.c_longlong
    O.fmo_4mc_read_index.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_int64), C.c_size_t]
    O.fmo_lz4_decompress_safe.restype = C.c_int
    O.fmo_xxh32.restype = C.c_uint32
    O.fmo_xxh32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    stream = open(path, "rb").read()
    offs = (C.c_int64 * 4096)()
    nb = O.fmo_4mc_read_index(stream, len(stream), 0x344D4300, offs, 4096)
    index = pkg.FourMcBlockIndex(list(offs[:nb]))
    splits = index.plan_splits(len(stream), split_size)
    mine = pkg.shard_splits(len(splits), world, rank)
    seen = torch.zeros(nb, dtype=torch.int64)
    sums = torch.zeros(nb, dtype=torch.int64)
    for i in mine:
        start, length = splits[i]
        for b in range(nb):
            if start <= offs[b] < start + length:                  # the blocks whose header lies in the split
                u, c, ck = struct.unpack(">III", stream[offs[b]:offs[b] + 12])
                payload = stream[offs[b] + 12:offs[b] + 12 + c]
                assert O.fmo_xxh32(payload, c, 0) == ck
                if c == u:
                    raw = payload
                else:
                    out = C.create_string_buffer(u)
                    assert O.fmo_lz4_decompress_safe(payload, out, c, u) == u
                    raw = out.raw
                with open(os.path.join(out_dir, f"block{b:05d}"), "wb") as f:
                    f.write(raw)
                seen[b] += 1
                sums[b] = O.fmo_xxh32(raw, u, 0)
    dist.all_reduce(seen)
    dist.all_reduce(sums)
    assert seen.tolist() == [1] * nb                               # every block read by exactly one rank
    if rank == 0:
        torch.save({"sums": sums, "splits": splits, "per_rank": [pkg.shard_splits(len(splits), world, r) for r in range(world)]},
                   os.path.join(out_dir, "summary.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_split_reader(oracle, ora, pkg, tmp_path):
    import socket
    n_bytes = 9 * BLOCK + 4321
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import gen_logtext
    data = gen_logtext(pkg, n_bytes)
    path = str(tmp_path / "in.4mc")
    with open(path, "wb") as f:
        f.write(ora.compress_4mc(data))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_dir = str(tmp_path / "blocks")
    os.makedirs(out_dir)
    split_size = 3 * 1024 * 1024                                   # several splits, not aligned to blocks
    mp.spawn(_reader_worker, args=(2, port, path, split_size, out_dir), nprocs=2, join=True)
    summary = torch.load(os.path.join(out_dir, "summary.pt"))
    assert len(summary["splits"]) >= 4
    assert sorted(summary["per_rank"][0] + summary["per_rank"][1]) == list(range(len(summary["splits"])))
    assert all(len(p) >= len(summary["splits"]) // 2 for p in summary["per_rank"])
    got = b"".join(open(os.path.join(out_dir, f), "rb").read() for f in sorted(os.listdir(out_dir)) if f.startswith("block"))
    assert got == data
    for b, s in enumerate(summary["sums"].tolist()):
        assert s == ora.xxh32(data[b * BLOCK:(b + 1) * BLOCK])


def test_shard_helpers(pkg):
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 8):
            parts = [pkg.shard_splits(n, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            blocks = [pkg.shard_blocks(n, world, r) for r in range(world)]
            assert sum(h - l for l, h in blocks) == n
            assert all(blocks[r][1] == blocks[r + 1][0] or blocks[r + 1][0] == n for r in range(world - 1))
