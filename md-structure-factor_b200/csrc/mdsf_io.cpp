// libmdsf_io.so: parallel inflate of the pieces of a traj npz member (see include/mdsf_io.h).
#include "../../include/mdsf_io.h"

#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace {
bool inflate_one(int fd, int64_t off, int64_t clen, void* dst, int64_t rlen, std::vector<unsigned char>& in) {
    if (clen < 0 || rlen < 0) return false;
    in.resize((size_t)clen);
    int64_t got = 0;
    while (got < clen) {
        const ssize_t r = pread(fd, in.data() + got, (size_t)(clen - got), (off_t)(off + got));
        if (r <= 0) return false;
        got += r;
    }
    z_stream zs{};
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    unsigned char sink[8];                       // a piece must not produce more than raw_len bytes
    int64_t ipos = 0, opos = 0;
    bool ok = true;
    for (;;) {
        const int64_t ileft = clen - ipos, oleft = rlen - opos;
        zs.next_in = in.data() + ipos;
        zs.avail_in = (uInt)std::min<int64_t>(ileft, 1 << 30);
        zs.next_out = oleft > 0 ? (unsigned char*)dst + opos : sink;
        zs.avail_out = oleft > 0 ? (uInt)std::min<int64_t>(oleft, 1 << 30) : (uInt)sizeof sink;
        const uInt ain = zs.avail_in, aout = zs.avail_out;
        const int rc = inflate(&zs, Z_SYNC_FLUSH);
        const int64_t din = ain - zs.avail_in, dout = aout - zs.avail_out;
        ipos += din;
        if (oleft > 0) opos += dout; else if (dout > 0) { ok = false; break; }
        if (rc == Z_STREAM_END) break;
        if (rc != Z_OK && rc != Z_BUF_ERROR) { ok = false; break; }
        if (ipos >= clen && opos >= rlen) break;     // a sync-flushed piece ends without a final block
        if (din == 0 && dout == 0) break;            // no progress: the size checks below decide
    }
    inflateEnd(&zs);
    return ok && opos == rlen && ipos == clen;
}
}  // namespace

extern "C" int mdsf_io_abi_version(void) { return 1; }

extern "C" int mdsf_io_inflate_pieces(int fd, int64_t n, const int64_t* file_off, const int64_t* comp_len,
                                      void* const* dst, const int64_t* raw_len, int threads) {
    if (n <= 0) return 0;
    if (!file_off || !comp_len || !dst || !raw_len) return -1;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, n));
    std::atomic<int64_t> next{0}, bad{n};
    auto work = [&]() {
        std::vector<unsigned char> in;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            if (!inflate_one(fd, file_off[i], comp_len[i], dst[i], raw_len[i], in)) {
                int64_t cur = bad.load();
                while (i < cur && !bad.compare_exchange_weak(cur, i)) {}
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    const int64_t b = bad.load();
    return b < n ? -(int)(b + 1) : 0;
}
