// fileio.h -- file -> file through the block kernels, streaming (host code of lib4mcgpu.so; included by capi.cu).
//
// The reference's fourMCcompressFilename / fourMcDecompressFileName (native/4mc.c:220-386, :896-934) and their
// MZ twins (:388-556, :936-966) stream a file one 4 MiB block at a time.  Here a reader thread fills pinned
// bounce buffers with slices of many blocks, the calling thread drives the GPU over each slice (upload,
// block kernels, download) and a writer thread drains the results, so a file of any size -- larger than
// host memory, or a pipe -- goes through in one pass at the speed of the slower of read() and write().
//
//   fourmc_compress_fd / fourmc_decompress_fd      descriptors in, descriptors out, status codes back
//   fourMCcompressFilename ... fourMZDecompressFileName   the reference's own four entry points
//       (native/4mc.h:36-41): same arguments, same console messages per display level, and -- like the
//       reference -- exit(1..4) on a fatal error (native/4mc.c:135-161), 0 on success.
#pragma once

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace {

constexpr size_t FIO_BLOCK = FOURMC_BLOCKSIZE;
constexpr int FIO_BUFS = 3;                            // bounce buffers per direction
constexpr size_t FIO_ENC_BLOCKS = 32;                  // blocks per slice, writer (128 MiB in)
constexpr size_t FIO_DEC_IN = (size_t)160 << 20;       // bytes per chunk, reader
constexpr size_t FIO_DEC_OUT_BLOCKS = 64;              // blocks per decode batch (256 MiB out)

template <class T>
class Chan {
    std::mutex m;
    std::condition_variable cv;
    std::deque<T> q;
    bool closed = false;

public:
    void push(const T &v) { { std::lock_guard<std::mutex> l(m); q.push_back(v); } cv.notify_one(); }
    void close() { { std::lock_guard<std::mutex> l(m); closed = true; } cv.notify_all(); }
    bool pop(T &v)
    {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return !q.empty() || closed; });
        if (q.empty()) return false;
        v = q.front(); q.pop_front();
        return true;
    }
};

struct Filled { int buf; size_t off, len; int eof; int err; };

// read until `want` bytes, end of file (returns what it got) or an error (-1)
long long read_full(int fd, uint8_t *p, size_t want)
{
    size_t got = 0;
    while (got < want) {
        const ssize_t r = read(fd, p + got, want - got);
        if (r < 0) { if (errno == EINTR) continue; return -1; }
        if (r == 0) break;
        got += (size_t)r;
    }
    return (long long)got;
}

bool write_full(int fd, const uint8_t *p, size_t n)
{
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}

// Regular files are read and written at explicit offsets by a few threads at once (one thread's read() / write()
// through the page cache moves 2-3 GB/s, less than the GPU side of this pipeline); pipes stay sequential.
constexpr int FIO_IO_THREADS = 4;

bool positional(int fd)
{
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return false;
    const int fl = fcntl(fd, F_GETFL);
    return fl >= 0 && !(fl & O_APPEND) && lseek(fd, 0, SEEK_CUR) >= 0;
}

// bytes read into p from file offset `off` (short only at the end of the file), or -1
long long pread_par(int fd, uint8_t *p, size_t want, uint64_t off)
{
    const size_t part = ((want + FIO_IO_THREADS - 1) / FIO_IO_THREADS + 4095) & ~(size_t)4095;
    long long got[FIO_IO_THREADS];
    std::thread th[FIO_IO_THREADS];
    int nt = 0;
    for (size_t b = 0; b < want; b += part, nt++) {
        const size_t len = std::min(part, want - b);
        th[nt] = std::thread([=, &got] {
            size_t g = 0;
            while (g < len) {
                const ssize_t r = pread(fd, p + b + g, len - g, (off_t)(off + b + g));
                if (r < 0) { if (errno == EINTR) continue; got[nt] = -1; return; }
                if (r == 0) break;
                g += (size_t)r;
            }
            got[nt] = (long long)g;
        });
    }
    long long total = 0;
    bool ended = false, bad = false;
    for (int i = 0; i < nt; i++) {
        th[i].join();
        if (got[i] < 0) bad = true;
        else if (!ended) { total += got[i]; if ((size_t)got[i] < std::min(part, want - (size_t)i * part)) ended = true; }
    }
    return bad ? -1 : total;
}

bool pwrite_par(int fd, const uint8_t *p, size_t n, uint64_t off)
{
    if (n < ((size_t)8 << 20)) {
        while (n) {
            const ssize_t w = pwrite(fd, p, n, (off_t)off);
            if (w < 0) { if (errno == EINTR) continue; return false; }
            p += w; n -= (size_t)w; off += (uint64_t)w;
        }
        return true;
    }
    const size_t part = ((n + FIO_IO_THREADS - 1) / FIO_IO_THREADS + 4095) & ~(size_t)4095;
    std::atomic<bool> ok{true};
    std::thread th[FIO_IO_THREADS];
    int nt = 0;
    for (size_t b = 0; b < n; b += part, nt++) {
        const size_t len = std::min(part, n - b);
        th[nt] = std::thread([=, &ok] {
            size_t g = 0;
            while (g < len) {
                const ssize_t w = pwrite(fd, p + b + g, len - g, (off_t)(off + b + g));
                if (w < 0) { if (errno == EINTR) continue; ok = false; return; }
                g += (size_t)w;
            }
        });
    }
    for (int i = 0; i < nt; i++) th[i].join();
    return ok;
}

struct PinSet {
    uint8_t *p[FIO_BUFS] = {};
    size_t cap = 0;
    ~PinSet() { for (auto q : p) if (q) cudaFreeHost(q); }
    bool alloc(size_t bytes)
    {
        cap = bytes;
        for (auto &q : p) if (cudaMallocHost(&q, bytes) != cudaSuccess) { q = nullptr; return false; }
        return true;
    }
};

// the writer thread: drains buffers in order
struct Drain {
    int fd;
    PinSet *bufs;
    Chan<Filled> filled;
    Chan<int> free_;
    std::atomic<bool> failed{false};
    std::thread th;
    bool pos = false;                                    // regular file: positional, parallel writes
    uint64_t at = 0;                                     // file offset of the next byte
    void start()
    {
        pos = positional(fd);
        if (pos) at = (uint64_t)lseek(fd, 0, SEEK_CUR);
        for (int i = 0; i < FIO_BUFS; i++) free_.push(i);
        th = std::thread([this] {
            Filled f;
            while (filled.pop(f)) {
                if (!failed && f.len) {
                    const bool ok = pos ? pwrite_par(fd, bufs->p[f.buf] + f.off, f.len, at) : write_full(fd, bufs->p[f.buf] + f.off, f.len);
                    if (!ok) failed = true;
                    at += f.len;
                }
                if (f.eof) free_.push(f.buf);            // eof = "last piece of this buffer"
            }
        });
    }
    // the calling thread's own small writes (header, footer) go through here, in order with the buffers
    bool write_now(const uint8_t *p, size_t n)
    {
        const bool ok = pos ? pwrite_par(fd, p, n, at) : write_full(fd, p, n);
        at += n;
        return ok;
    }
    void finish()
    {
        filled.close();
        if (th.joinable()) th.join();
        if (pos) lseek(fd, (off_t)at, SEEK_SET);
    }
    ~Drain() { finish(); }
};

uint32_t header_checksum(int codec) { return codec == CODEC_ZSTD ? 0x289A1C9Au : 0xA4B73443u; }      // XXH32(magic, version 1)

void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }

}  // namespace

extern "C" {

// native/4mc.c:263-362 (fourMCcompressFilename) / :431-531 (fourMZcompressFilename) over descriptors: header, one
// block record per 4 MiB read, end-of-stream mark, footer index.  Returns the bytes written or a negative
// FOURMC_E_* (FOURMC_E_INPUT: read error, FOURMC_E_OUTPUT: write error); *in_bytes (may be NULL) = bytes read.
long long fourmc_compress_fd(fourmc_ctx *ctx, int zstd, int level, int in_fd, int out_fd, uint64_t *in_bytes)
{
    if (!ctx || in_fd < 0 || out_fd < 0) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int codec = zstd ? CODEC_ZSTD : CODEC_LZ4;
    if (level < 1) level = 1;
    const size_t slice = FIO_ENC_BLOCKS * FIO_BLOCK, out_cap = slice + 12 * FIO_ENC_BLOCKS + 64;
    PinSet in, out;
    if (!in.alloc(slice) || !out.alloc(out_cap)) return fail(ctx, FOURMC_E_CUDA, "pinned bounce buffers");
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], slice + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], out_cap))) return r;
    if ((r = pinned_scratch(ctx, 4096 + FIO_ENC_BLOCKS * 4))) return r;

    // reader thread: whole slices, the last one short
    Chan<Filled> filled_in;
    Chan<int> free_in;
    for (int i = 0; i < FIO_BUFS; i++) free_in.push(i);
    const bool in_pos = positional(in_fd);
    const uint64_t in_base = in_pos ? (uint64_t)lseek(in_fd, 0, SEEK_CUR) : 0;
    std::thread reader([&] {
        int b;
        uint64_t at = in_base;
        while (free_in.pop(b)) {
            const long long n = in_pos ? pread_par(in_fd, in.p[b], slice, at) : read_full(in_fd, in.p[b], slice);
            if (n > 0) at += (uint64_t)n;
            filled_in.push(Filled{b, 0, n < 0 ? 0 : (size_t)n, n < (long long)slice, n < 0});
            if (n < (long long)slice) break;
        }
        filled_in.close();
    });
    Drain drain;
    drain.fd = out_fd; drain.bufs = &out;
    drain.start();
    auto stop = [&](long long code) { free_in.close(); reader.join(); drain.finish(); return code; };

    uint8_t hdr[12];
    put_be32(hdr, zstd ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC); put_be32(hdr + 4, FOURMC_VERSION); put_be32(hdr + 8, header_checksum(codec));
    if (!drain.write_now(hdr, 12)) return stop(FOURMC_E_OUTPUT);                          // :273 (nothing is queued yet)
    uint64_t total_in = 0, total_out = 12;
    std::vector<uint32_t> lens;
    cudaStream_t st = ctx->stream;
    EncWs &ws = ctx->enc[0];
    uint64_t *h_span = (uint64_t *)ctx->pinned;
    uint32_t *h_lens = (uint32_t *)((uint8_t *)ctx->pinned + 64);
    Filled f;
    while (filled_in.pop(f)) {
        if (f.err) return stop(FOURMC_E_INPUT);
        if (f.len) {
            const uint32_t cnt = blocks_of(f.len);
            if (cudaMemcpyAsync(ctx->stage_in[0].p, in.p[f.buf], f.len, cudaMemcpyHostToDevice, st) != cudaSuccess)
                return stop(fail(ctx, FOURMC_E_CUDA, "H2D"));
            if ((r = ensure(ctx, ws.lens, (size_t)cnt * 4))) return stop(r);
            if ((r = enc_span_codec(ctx, st, ws, codec, level, (const uint8_t *)ctx->stage_in[0].p, f.len,
                                    (uint8_t *)ctx->stage_out[0].p, 0, (uint32_t *)ws.lens.p, -1)))
                return stop(r);
            cudaMemcpyAsync(h_span, (uint8_t *)ws.misc.p + 8, 8, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(h_lens, ws.lens.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) return stop(fail(ctx, FOURMC_E_CUDA, "encode slice"));
            const size_t span = (size_t)*h_span;
            lens.insert(lens.end(), h_lens, h_lens + cnt);
            int ob;
            if (!drain.free_.pop(ob)) return stop(FOURMC_E_OUTPUT);
            if (cudaMemcpyAsync(out.p[ob], ctx->stage_out[0].p, span, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess)
                return stop(fail(ctx, FOURMC_E_CUDA, "D2H"));
            drain.filled.push(Filled{ob, 0, span, 1, 0});
            total_in += f.len; total_out += span;
        }
        if (!f.eof) free_in.push(f.buf);
    }
    reader.join();
    drain.filled.close();
    if (drain.th.joinable()) drain.th.join();
    if (drain.failed) return FOURMC_E_OUTPUT;                                             // :312 / :327
    // end-of-stream mark + footer, assembled on the device from the block lengths (:335-362)
    const uint32_t nb = (uint32_t)lens.size();
    const size_t tail_bytes = 12 + 20 + 4 * (size_t)nb;
    DevBuf &tb = ctx->dec[1].tables;
    if ((r = ensure(ctx, tb, (size_t)std::max<uint32_t>(nb, 1) * 4 + 64 + tail_bytes))) return r;
    uint8_t *d_tail = (uint8_t *)tb.p + (((size_t)std::max<uint32_t>(nb, 1) * 4 + 15) & ~(size_t)15);
    if (nb) CK(cudaMemcpyAsync(tb.p, lens.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    if ((r = build_index_impl(ctx, st, zstd ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC, (const uint32_t *)tb.p, nb, nullptr, d_tail))) return r;
    std::vector<uint8_t> tail(tail_bytes);
    CK(cudaMemcpyAsync(tail.data(), d_tail, tail_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const bool tail_ok = drain.write_now(tail.data(), tail_bytes);                        // :339 / :361
    drain.finish();
    if (!tail_ok) return FOURMC_E_OUTPUT;
    if (in_bytes) *in_bytes = total_in;
    return (long long)(total_out + tail_bytes);
}

// decodeFourMC / decodeFourMZ in the loop over concatenated streams (native/4mc.c:560-707, :709-857, :896-913) over
// descriptors.  Blocks are written as they decode, so -- like the serial reader -- everything that precedes a damaged
// block has reached the output when the error is returned.  Returns the decoded size or FOURMC_E_*.
long long fourmc_decompress_fd(fourmc_ctx *ctx, int zstd, int in_fd, int out_fd, uint64_t *in_bytes)
{
    if (!ctx || in_fd < 0 || out_fd < 0) return FOURMC_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int codec = zstd ? CODEC_ZSTD : CODEC_LZ4;
    const uint32_t magic = zstd ? FOURMC_MAGIC_4MZ : FOURMC_MAGIC_4MC;
    const size_t out_cap = FIO_DEC_OUT_BLOCKS * FIO_BLOCK;
    PinSet in, out;
    if (!in.alloc(FIO_DEC_IN) || !out.alloc(out_cap)) return fail(ctx, FOURMC_E_CUDA, "pinned bounce buffers");
    int r;
    if ((r = ensure(ctx, ctx->stage_in[0], FIO_DEC_IN + 64))) return r;
    if ((r = ensure(ctx, ctx->stage_out[0], out_cap + 64))) return r;

    // reader thread: fills the buffer it is handed from `off` (the unconsumed tail of the previous chunk sits below)
    struct Req { int buf; size_t off; };
    Chan<Req> reqs;
    Chan<Filled> filled_in;
    const bool in_pos = positional(in_fd);
    const uint64_t in_base = in_pos ? (uint64_t)lseek(in_fd, 0, SEEK_CUR) : 0;
    std::thread reader([&] {
        Req q;
        bool eof = false;
        uint64_t at = in_base;
        while (reqs.pop(q)) {
            long long n = 0;
            if (!eof) n = in_pos ? pread_par(in_fd, in.p[q.buf] + q.off, FIO_DEC_IN - q.off, at) : read_full(in_fd, in.p[q.buf] + q.off, FIO_DEC_IN - q.off);
            if (n > 0) at += (uint64_t)n;
            if (n < (long long)(FIO_DEC_IN - q.off)) eof = true;
            filled_in.push(Filled{q.buf, 0, q.off + (n < 0 ? 0 : (size_t)n), eof, n < 0});
        }
        filled_in.close();
    });
    Drain drain;
    drain.fd = out_fd; drain.bufs = &out;
    drain.start();
    auto stop = [&](long long code) { reqs.close(); reader.join(); drain.finish(); if (code >= 0 && drain.failed) return (long long)FOURMC_E_OUTPUT; return code; };

    struct Item { size_t off; uint32_t c, u, ck; };        // u == 0xffffffff: a footer (checksum only)
    std::vector<Item> items;
    cudaStream_t st = ctx->stream;
    DecWs &ws = ctx->dec[0];
    uint64_t total_out = 0, total_in = 0, stream_out = 0;
    std::vector<uint8_t> h_tab, h_status;
    std::vector<int32_t> h_size;

    // decodes items[i0, i1) of chunk buffer `cb` (consecutive in the chunk), writes their output; FOURMC_OK or the error
    auto run = [&](const uint8_t *cb, size_t i0, size_t i1) -> long long {
        if (i0 >= i1) return FOURMC_OK;
        const uint32_t cnt = (uint32_t)(i1 - i0);
        const size_t s0 = items[i0].off, s1 = items[i1 - 1].off + items[i1 - 1].c;
        h_tab.resize((size_t)cnt * 28);
        uint64_t *t_src = (uint64_t *)h_tab.data(), *t_dst = t_src + cnt;
        uint32_t *t_c = (uint32_t *)(t_dst + cnt), *t_u = t_c + cnt, *t_x = t_u + cnt;
        uint64_t dpos = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            const Item &it = items[i0 + i];
            t_src[i] = it.off - s0; t_dst[i] = dpos; t_c[i] = it.c; t_u[i] = it.u; t_x[i] = it.ck;
            if (it.u != 0xffffffffu) dpos += it.u;
        }
        int rr;
        if ((rr = ensure(ctx, ws.tables, (size_t)cnt * 33 + 64))) return rr;
        uint8_t *d_tb = (uint8_t *)ws.tables.p;
        CK(cudaMemcpyAsync(ctx->stage_in[0].p, cb + s0, s1 - s0, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_tb, h_tab.data(), h_tab.size(), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                                  // h_tab is pageable
        const uint64_t *d_so = (const uint64_t *)d_tb, *d_do = d_so + cnt;
        const uint32_t *d_c = (const uint32_t *)(d_do + cnt), *d_u = d_c + cnt, *d_x = d_u + cnt;
        int32_t *d_osz = (int32_t *)(d_x + cnt);
        uint8_t *d_st = (uint8_t *)(d_osz + cnt);
        if ((rr = dec_batch(ctx, st, ws, cnt, ctx->stage_in[0].p, d_so, d_c, d_u, d_x, 1, ctx->stage_out[0].p, d_do, d_osz, d_st, codec)))
            return rr;
        h_status.resize(cnt); h_size.resize(cnt);
        CK(cudaMemcpyAsync(h_status.data(), d_st, cnt, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_size.data(), d_osz, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        int ob = -1;
        if (dpos) {
            if (!drain.free_.pop(ob)) return FOURMC_E_OUTPUT;
            CK(cudaMemcpyAsync(out.p[ob], ctx->stage_out[0].p, (size_t)dpos, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        // in stream order: everything before the first failing item is written (:637-668)
        long long verdict = FOURMC_OK;
        bool exact = true;
        uint32_t good = 0;
        for (; good < cnt; good++) {
            if (h_status[good] != FOURMC_BLOCK_OK) { verdict = FOURMC_E_CONTENT; break; }
            if (items[i0 + good].u != 0xffffffffu && (uint32_t)h_size[good] != items[i0 + good].u) exact = false;
        }
        if (ob >= 0) {
            if (exact) {
                const uint64_t n = good == cnt ? dpos : t_dst[good];
                drain.filled.push(Filled{ob, 0, (size_t)n, 1, 0});
                total_out += n; stream_out += n;
            } else {                                                    // some block decoded short (:661-666): piece by piece
                for (uint32_t i = 0; i < good; i++) {
                    if (items[i0 + i].u == 0xffffffffu) continue;
                    drain.filled.push(Filled{ob, (size_t)t_dst[i], (size_t)h_size[i], 0, 0});
                    total_out += (uint64_t)h_size[i]; stream_out += (uint64_t)h_size[i];
                }
                drain.filled.push(Filled{ob, 0, 0, 1, 0});
            }
        }
        return verdict;
    };

    enum { HEADER, BLOCKS, FOOTER } state = HEADER;
    int next_buf = 1;
    reqs.push(Req{0, 0});
    long long result = FOURMC_OK;
    bool done = false;
    std::vector<uint8_t> foot_flags;                                    // per item: 1 = footer, 2 = footer with a bad version field
    while (!done) {
        Filled f;
        if (!filled_in.pop(f) || f.err) { result = FOURMC_E_INPUT; break; }
        const uint8_t *cb = in.p[f.buf];
        const size_t len = f.len;
        // ---- walk: the container fields of everything complete in this chunk (no payload work)
        size_t pos = 0;
        items.clear(); foot_flags.clear();
        long long deferred = FOURMC_OK;                                 // container error met by the walk: it counts after the items before it
        bool need_more = false, end_of_input = false;
        while (!need_more && deferred == FOURMC_OK) {
            const size_t avail = len - pos;
            if (state == HEADER) {
                if (avail == 0 && f.eof) { end_of_input = true; break; }                         // :867 end of input
                if (avail < 12 && !f.eof) { need_more = true; break; }
                if (avail < 4) { deferred = FOURMC_E_CONTENT; break; }                            // :868
                if (be32(cb + pos) != magic) { deferred = FOURMC_E_CONTENT; break; }              // :873
                if (avail < 12) { deferred = FOURMC_E_CONTENT; break; }                           // :577
                if (be32(cb + pos + 4) != FOURMC_VERSION || be32(cb + pos + 8) != header_checksum(codec)) { deferred = FOURMC_E_CONTENT; break; }   // :583-584
                pos += 12; state = BLOCKS;
            } else if (state == BLOCKS) {
                if (avail < 12) { if (f.eof) deferred = FOURMC_E_INPUT; else need_more = true; break; }   // :610
                const uint32_t u = be32(cb + pos), c = be32(cb + pos + 4), ck = be32(cb + pos + 8);
                if (u == 0 && c == 0 && ck == 0) { pos += 12; state = FOOTER; continue; }        // :616
                if (c > FOURMC_BLOCKSIZE) { deferred = FOURMC_E_CONTENT; break; }                 // :618
                if (avail - 12 < c) { if (f.eof) deferred = FOURMC_E_INPUT; else need_more = true; break; }   // :632
                if (u != c && u > FOURMC_BLOCKSIZE) { deferred = FOURMC_E_CONTENT; break; }       // :651
                items.push_back(Item{pos + 12, c, u, ck}); foot_flags.push_back(0);
                pos += 12 + (size_t)c;
            } else {
                if (avail < 4) { if (f.eof) deferred = FOURMC_E_GENERIC; else need_more = true; break; }   // :672
                const uint32_t fsize = be32(cb + pos);
                if (fsize < 4 || fsize > FIO_DEC_IN / 2) { deferred = FOURMC_E_INPUT; break; }    // :680 (no file holds such a footer)
                if (avail < fsize) { if (f.eof) deferred = FOURMC_E_INPUT; else need_more = true; break; }
                if (fsize < 8) { deferred = FOURMC_E_CONTENT; break; }
                items.push_back(Item{pos, fsize - 4, 0xffffffffu, be32(cb + pos + fsize - 4)});   // :685 checksum
                foot_flags.push_back(be32(cb + pos + 4) != 1 ? 2 : 1);                            // :687 version, after the checksum
                pos += fsize;
                state = HEADER;                                          // provided the stream decoded to something (below)
            }
        }
        total_in += pos;
        // ---- the next chunk is read while this one decodes: its unconsumed tail goes first
        const size_t tail = len - pos;
        const bool more_input = need_more && deferred == FOURMC_OK && !end_of_input;
        if (more_input) {
            memcpy(in.p[next_buf], cb + pos, tail);
            reqs.push(Req{next_buf, tail});
            next_buf = (next_buf + 1) % FIO_BUFS;
        }
        // ---- decode, in batches bounded by the output buffer; a footer closes its stream
        size_t i0 = 0;
        uint64_t batch_out = 0;
        for (size_t i = 0; i < items.size() && result == FOURMC_OK && !done; i++) {
            const bool foot = foot_flags[i] != 0;
            if (!foot && (batch_out + items[i].u > out_cap || i - i0 >= 4096)) {
                result = run(cb, i0, i);
                i0 = i; batch_out = 0;
                if (result != FOURMC_OK) break;
            }
            if (!foot) { batch_out += items[i].u; continue; }
            result = run(cb, i0, i + 1);
            i0 = i + 1; batch_out = 0;
            if (result != FOURMC_OK) break;
            if (foot_flags[i] == 2) { result = FOURMC_E_CONTENT; break; }
            if (stream_out == 0) done = true;                            // :909-913 `do {...} while (decodedSize)`
            stream_out = 0;
        }
        if (result == FOURMC_OK && !done) result = run(cb, i0, items.size());
        if (result == FOURMC_OK && !done && deferred != FOURMC_OK) result = deferred;
        if (result != FOURMC_OK || done || !more_input) break;
    }
    if (in_bytes) *in_bytes = total_in;
    return stop(result == FOURMC_OK ? (long long)total_out : result);
}

}  // extern "C"

// ---- the reference's own entry points (native/4mc.h:36-41) ------------------------------------------------------

namespace {

fourmc_ctx *file_ctx()
{
    static fourmc_ctx *c = nullptr;                  // one per process, like the CLI's lifetime
    static std::mutex m;
    std::lock_guard<std::mutex> l(m);
    if (!c && fourmc_ctx_create(&c, -1) != FOURMC_OK) c = nullptr;
    return c;
}

double fio_now()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + (double)t.tv_nsec * 1e-9;
}
// FOURMC_CLI_TIMING=1: wall-clock phases of a file call on stderr (tools/cli_file_timing.py)
void fio_timing(const char *what, double t0, double t1, double t2, uint64_t bytes)
{
    if (!getenv("FOURMC_CLI_TIMING")) return;
    fprintf(stderr, "timing: %s: context %.3f s, stream %.3f s = %.2f GB/s of %llu bytes\n", what, t1 - t0, t2 - t1,
            t2 > t1 ? (double)bytes / (t2 - t1) / 1e9 : 0.0, (unsigned long long)bytes);
}

#define FIO_SAY(l, ...) do { if (displayLevel >= (l)) fprintf(stderr, __VA_ARGS__); } while (0)
#define FIO_DIE(code, ...) do { FIO_SAY(1, __VA_ARGS__); FIO_SAY(1, "\n"); exit(code); } while (0)

// openIOFileHandles, native/4mc.c:163-211
void open_io(int displayLevel, int overwrite, const char *in_name, const char *out_name, int *in_fd, int *out_fd)
{
    if (!strcmp(in_name, "stdin")) { FIO_SAY(4, "Using stdin for input\n"); *in_fd = 0; }
    else *in_fd = open(in_name, O_RDONLY);
    if (!strcmp(out_name, "stdout")) { FIO_SAY(4, "Using stdout for output\n"); *out_fd = 1; }
    else {
        if (strcmp(out_name, "/dev/null") && access(out_name, F_OK) == 0 && !overwrite) {
            FIO_SAY(2, "Warning : %s already exists\n", out_name);
            FIO_SAY(2, "Overwrite ? (Y/N) : ");
            if (displayLevel <= 1) FIO_DIE(3, "Operation aborted : %s already exists", out_name);
            const int ch = getchar();
            if (ch != 'Y' && ch != 'y') FIO_DIE(3, "Operation aborted : %s already exists", out_name);
        }
        *out_fd = open(out_name, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    }
    if (*in_fd < 0) FIO_DIE(2, "Cannot open input file: %s", in_name);
    if (*out_fd < 0) FIO_DIE(3, "Cannot open output file: %s", out_name);
}

int compress_filename(int zstd, int displayLevel, int overwrite, const char *in_name, const char *out_name, int level)
{
    const clock_t t0 = clock();
    if (displayLevel == 2 && level > 1) displayLevel = 3;                                 // :237
    int in_fd, out_fd;
    open_io(displayLevel, overwrite, in_name, out_name, &in_fd, &out_fd);
    const double w0 = fio_now();
    fourmc_ctx *ctx = file_ctx();
    if (!ctx) FIO_DIE(1, "lib4mcgpu: no usable CUDA device (there is no CPU fallback)");
    uint64_t n = 0;
    const double w1 = fio_now();
    const long long c = fourmc_compress_fd(ctx, zstd, level, in_fd, out_fd, &n);
    if (in_fd != 0) close(in_fd);
    if (out_fd != 1) close(out_fd);
    fio_timing("compress", w0, w1, fio_now(), n);
    if (c == FOURMC_E_INPUT) FIO_DIE(2, "Read error : cannot read input");
    if (c == FOURMC_E_OUTPUT) FIO_DIE(3, "Write error : cannot write compressed block");
    if (c < 0) FIO_DIE(1, "Compression failed: %s", fourmc_last_error(ctx));
    FIO_SAY(2, "\r%79s\r", "");
    FIO_SAY(2, "Compressed (%s) %llu bytes into %llu bytes ==> %.2f%% (Ratio=%.3f)\n",
            level <= 1 ? "fast" : level == 2 ? "medium" : level == 3 ? "high" : "ultra", (unsigned long long)n, (unsigned long long)c,
            n ? (double)c / (double)n * 100 : 0.0, c ? (double)n / (double)c : 0.0);
    const double s = (double)(clock() - t0) / CLOCKS_PER_SEC;
    FIO_SAY(4, "Done in %.2f s ==> %.2f MB/s\n", s, s > 0 ? (double)n / s / 1024 / 1024 : 0.0);
    return 0;
}

int decompress_filename(int zstd, int displayLevel, int overwrite, const char *in_name, const char *out_name)
{
    const clock_t t0 = clock();
    int in_fd, out_fd;
    open_io(displayLevel, overwrite, in_name, out_name, &in_fd, &out_fd);
    const double w0 = fio_now();
    fourmc_ctx *ctx = file_ctx();
    if (!ctx) FIO_DIE(1, "lib4mcgpu: no usable CUDA device (there is no CPU fallback)");
    const double w1 = fio_now();
    const long long d = fourmc_decompress_fd(ctx, zstd, in_fd, out_fd, nullptr);
    if (in_fd != 0) close(in_fd);
    if (out_fd != 1) close(out_fd);
    fio_timing("decompress", w0, w1, fio_now(), d > 0 ? (uint64_t)d : 0);
    if (d == FOURMC_E_INPUT) FIO_DIE(2, "Read error : cannot read next block");
    if (d == FOURMC_E_OUTPUT) FIO_DIE(3, "Write error : cannot write decoded block");
    if (d == FOURMC_E_CONTENT) FIO_DIE(4, "Decoding Failed ! Corrupted input detected !");
    if (d < 0) FIO_DIE(1, "%s", d == FOURMC_E_GENERIC ? "Unreadable footer" : fourmc_last_error(ctx));
    FIO_SAY(2, "\r%79s\r", "");
    FIO_SAY(2, "Successfully decoded %llu bytes \n", (unsigned long long)d);
    const double s = (double)(clock() - t0) / CLOCKS_PER_SEC;
    FIO_SAY(4, "Done in %.2f s ==> %.2f MB/s\n", s, s > 0 ? (double)d / s / 1024 / 1024 : 0.0);
    return 0;
}

}  // namespace

extern "C" {

int fourMCcompressFilename(int displayLevel, int overwrite, char *input_filename, char *output_filename, int compressionlevel)
{
    return compress_filename(0, displayLevel, overwrite, input_filename, output_filename, compressionlevel);
}
int fourMcDecompressFileName(int displayLevel, int overwrite, char *input_filename, char *output_filename)
{
    return decompress_filename(0, displayLevel, overwrite, input_filename, output_filename);
}
int fourMZcompressFilename(int displayLevel, int overwrite, char *input_filename, char *output_filename, int compressionlevel)
{
    return compress_filename(1, displayLevel, overwrite, input_filename, output_filename, compressionlevel);
}
int fourMZDecompressFileName(int displayLevel, int overwrite, char *input_filename, char *output_filename)
{
    return decompress_filename(1, displayLevel, overwrite, input_filename, output_filename);
}

}  // extern "C"
