"""Lane-by-lane emulation of the D2 word stage (4mc_b200/csrc/lz4_decode.cuh, lz4_copy_block, path A).

TEST / DEVELOPMENT AID: restates, with plain Python loops over the 32 lanes, which bytes of a batch travel as
aligned 32-bit words of the span, how a word shared by two neighbouring sequences is merged (the LEFT lane stores
it, OR-ing in the right lane's first word), and what is patched in afterwards with byte stores.  It shares no
code with the kernel; tests/test_host_logic.py runs it on real LZ4 sequences and on adversarial ones
(4-byte sequences, offsets 0..16, sources inside the span).

  run_batch(seqs, out_ref, dstbase, rnd, wlit=64, wmatch=32) -> (assembled bytes, op0, total)
      seqs: up to 32 tuples (lit, ml, off, op, literal bytes); out_ref: the expected output (source of the matches)
  parse_lz4_block(block bytes) -> list of (token position, lit, ml, off, op)
"""
import random

SPAN = 2048


def u32(x):
    return x & 0xffffffff


def parse_lz4_block(s):
    seqs = []
    c = len(s)
    ip = op = 0
    while ip < c:
        tp = ip
        tok = s[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = s[ip]; ip += 1; lit += b
                if b != 255:
                    break
        ip += lit
        if ip >= c:
            seqs.append((tp, lit, 0, 0, op)); op += lit
            break
        off = s[ip] | (s[ip + 1] << 8); ip += 2
        ml = tok & 15
        if ml == 15:
            while True:
                b = s[ip]; ip += 1; ml += b
                if b != 255:
                    break
        ml += 4
        seqs.append((tp, lit, ml, off, op)); op += lit + ml
    return seqs


def run_batch(seqs, out_ref, dstbase, rnd, wlit=64, wmatch=32):
    n = len(seqs)
    op0 = seqs[0][3]
    total = sum(s[0] + s[1] for s in seqs)
    assert n <= 32 and total <= SPAN
    shift = (dstbase + op0) & 15
    span = bytearray(rnd.randrange(256) for _ in range(SPAN + 64))       # whatever the last batch left there

    def P(x):
        return shift + (x - op0)

    def stw(idx, v):
        span[idx * 4:idx * 4 + 4] = int(v).to_bytes(4, "little")

    def masked_words(a, nbytes, srcbytes):
        """words of the span covering [a, a + nbytes), bytes outside zeroed (garbage around the source bytes)"""
        dd = a & 3
        nb = dd + nbytes
        nw = (nb + 3) >> 2
        raw = bytes(rnd.randrange(256) for _ in range(dd)) + srcbytes + bytes(rnd.randrange(256) for _ in range(8))
        out = []
        for w in range(nw):
            v = int.from_bytes(raw[4 * w:4 * w + 4], "little")
            if w == 0:
                v &= u32(0xffffffff << (8 * dd))
            if w == nw - 1:
                v &= 0xffffffff >> (8 * (4 * nw - nb))
            out.append(((a - dd) // 4 + w, v))
        return out

    lanes = []
    for lit, ml, off, op, lb in seqs:
        d = op + lit
        mstart = d - off
        fastm = ml > 0 and off != 0 and mstart + ml <= op0 and ml <= wmatch
        slowm = ml > 0 and not fastm
        rs, re = op, op + lit + ml
        if slowm:
            re = rs
        elif lit > wlit:
            rs = d
        if re - rs < 4:
            rs = re = op
        inn = re > rs
        lit_in = inn and rs == op and lit > 0
        lanes.append(dict(lit=lit, ml=ml, off=off, op=op, lb=lb, d=d, mstart=mstart, slowm=slowm, rs=rs, re=re, inn=inn,
                          lit_in=lit_in, m_in=inn and ml > 0, later=lit > 0 and not lit_in))
    for L in lanes:
        if not L["inn"]:
            continue
        words = []
        if L["lit_in"]:
            words += masked_words(P(L["op"]), L["lit"], L["lb"])
        if L["m_in"]:
            words += masked_words(P(L["d"]), L["ml"], bytes(out_ref[L["mstart"]:L["mstart"] + L["ml"]]))
        merged = []                                     # the word where the literals end and the match begins holds both
        for idx, v in words:
            if merged and merged[-1][0] == idx:
                merged[-1] = (idx, merged[-1][1] | v)
            else:
                merged.append((idx, v))
        assert all(merged[i + 1][0] == merged[i][0] + 1 for i in range(len(merged) - 1))
        L["words"] = merged
        L["hv"] = merged[0][1]
    for j, L in enumerate(lanes):
        if not L["inn"]:
            continue
        prev = lanes[j - 1] if j > 0 else None
        nxt = lanes[j + 1] if j + 1 < n else None
        head_skip = prev is not None and prev["inn"] and prev["re"] == L["rs"] and (P(L["rs"]) & 3) != 0
        adj = nxt is not None and nxt["inn"] and nxt["rs"] == L["re"] and (P(L["re"]) & 3) != 0
        w = L["words"]
        assert not (head_skip and len(w) < 2)
        for i, (idx, v) in enumerate(w):
            if i == 0 and head_skip:
                continue                                # the left neighbour stores this word
            if i == len(w) - 1 and adj:
                v |= nxt["hv"]
            stw(idx, v)
    # patches (byte stores): literals that did not travel as words, then the slow matches in destination order
    for L in lanes:
        if L["later"]:
            for i in range(L["lit"]):
                span[P(L["op"]) + i] = L["lb"][i]
    for L in lanes:
        if L["slowm"]:
            for i in range(L["ml"]):
                s = L["mstart"] + i
                span[P(L["d"]) + i] = 0 if L["off"] == 0 else (out_ref[s] if s < op0 else span[P(s)])
    return bytes(span[shift:shift + total]), op0, total


def adversarial_case(rnd):
    out = bytearray(rnd.randrange(256) for _ in range(300))
    seqs = []
    op = len(out)
    nl = rnd.randrange(1, 33)
    for j in range(nl):
        lit = rnd.choice([0, 0, 0, 1, 2, 3, 4, 5, 8, 9, 12, 30, 64, 65, 70])
        last = (j == nl - 1) and rnd.random() < 0.3
        ml = 0 if last else rnd.choice([4, 4, 5, 6, 7, 8, 9, 12, 16, 31, 32, 33, 40])
        lb = bytes(rnd.randrange(256) for _ in range(lit))
        d = op + lit
        off = min(rnd.choice([0, 1, 2, 3, 4, 7, 8, 15, 16, 33, 100, 250, d]), d) if ml else 0
        out += lb
        for i in range(ml):
            out.append(0 if off == 0 else out[d - off + i])
        seqs.append((lit, ml, off, op, lb))
        op += lit + ml
    return seqs, bytes(out)
