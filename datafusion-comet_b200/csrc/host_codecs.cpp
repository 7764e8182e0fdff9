#include "host_codecs.h"

#include <dlfcn.h>
#include <zlib.h>

#include <cstring>
#include <mutex>
#include <string>

#include "parquet.h"
#include "plan.h"

namespace cb200 {
namespace {
struct Zstd {
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool tried = false;
};
struct Lz4 {
    int (*decompress_safe)(const char*, char*, int, int) = nullptr;
    bool tried = false;
};
std::mutex g_mu;
Zstd g_zstd;
Lz4 g_lz4;

void* open_first(const char* const* names) {
    for (; *names; names++)
        if (void* h = dlopen(*names, RTLD_NOW | RTLD_LOCAL)) return h;
    return nullptr;
}
const Zstd& zstd() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_zstd.tried) {
        g_zstd.tried = true;
        static const char* const names[] = {"libzstd.so.1", "libzstd.so", nullptr};
        if (void* h = open_first(names)) {
            g_zstd.decompress = (size_t(*)(void*, size_t, const void*, size_t))dlsym(h, "ZSTD_decompress");
            g_zstd.is_error = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
        }
    }
    if (!g_zstd.decompress || !g_zstd.is_error) throw Unsupported("parquet codec ZSTD: libzstd.so.1 could not be loaded on this host");
    return g_zstd;
}
const Lz4& lz4() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_lz4.tried) {
        g_lz4.tried = true;
        static const char* const names[] = {"liblz4.so.1", "liblz4.so", nullptr};
        if (void* h = open_first(names)) g_lz4.decompress_safe = (int (*)(const char*, char*, int, int))dlsym(h, "LZ4_decompress_safe");
    }
    if (!g_lz4.decompress_safe) throw Unsupported("parquet codec LZ4: liblz4.so.1 could not be loaded on this host");
    return g_lz4;
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
} // namespace

bool host_codec_supported(int c) { return c == pq::GZIP || c == pq::LZ4 || c == pq::ZSTD || c == pq::LZ4_RAW; }

void host_decompress(int codec, const uint8_t* src, size_t n, uint8_t* dst, size_t unc) {
    if (unc == 0) return;
    if (codec == pq::ZSTD) {
        const Zstd& z = zstd();
        const size_t r = z.decompress(dst, unc, src, n);
        if (z.is_error(r) || r != unc) throw PlanError("parquet: malformed ZSTD page");
        return;
    }
    if (codec == pq::LZ4_RAW || codec == pq::LZ4) {
        const Lz4& l = lz4();
        if (n > 0x7fffffffu || unc > 0x7fffffffu) throw PlanError("parquet: LZ4 page larger than 2 GiB");
        if (codec == pq::LZ4 && n >= 8) {
            // the deprecated LZ4 codec as Hadoop writes it: repeated [u32 BE uncompressed][u32 BE compressed][raw block]
            size_t ip = 0, op = 0;
            bool ok = true;
            while (ip < n && ok) {
                if (ip + 8 > n) { ok = false; break; }
                const uint32_t u = be32(src + ip), c = be32(src + ip + 4);
                ip += 8;
                if (c > n - ip || u > unc - op) { ok = false; break; }
                if (l.decompress_safe((const char*)src + ip, (char*)dst + op, (int)c, (int)u) != (int)u) { ok = false; break; }
                ip += c;
                op += u;
            }
            if (ok && op == unc) return;
            // some writers put a raw block under the deprecated name: fall through and try that
        }
        if (l.decompress_safe((const char*)src, (char*)dst, (int)n, (int)unc) != (int)unc) throw PlanError("parquet: malformed LZ4 page");
        return;
    }
    if (codec == pq::GZIP) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15 + 32) != Z_OK) throw PlanError("parquet: zlib initialisation failed"); // +32: gzip or zlib header, detected
        zs.next_in = const_cast<Bytef*>(src);
        zs.avail_in = (uInt)n;
        zs.next_out = dst;
        zs.avail_out = (uInt)unc;
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.total_out == unc;
        inflateEnd(&zs);
        if (!ok) throw PlanError("parquet: malformed GZIP page");
        return;
    }
    throw Unsupported("parquet codec " + std::to_string(codec));
}

} // namespace cb200
