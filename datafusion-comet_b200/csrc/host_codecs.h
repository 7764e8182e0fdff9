// host_codecs.h -- page decompression on the HOST for the Parquet codecs the device does not decompress (ZSTD, LZ4 / LZ4_RAW, GZIP).
//
// The reference reads all of them through the third-party parquet crate (features `snap,lz4,zstd,flate2`, native/core/Cargo.toml:40).
// Here UNCOMPRESSED and SNAPPY pages are decoded entirely on the device; for the other codecs only the byte-level decompression
// runs on the host (the stated fallback of SURVEY.md 7 step 5) -- levels, dictionaries, PLAIN / RLE values are still decoded by the
// device kernels from the decompressed bytes.  libzstd / liblz4 ship without headers in this image: they are resolved with dlopen
// (like NCCL); zlib is linked.  A codec whose library is missing raises Unsupported, never a silent fallback.
#ifndef CB200_HOST_CODECS_H
#define CB200_HOST_CODECS_H
#include <cstddef>
#include <cstdint>

namespace cb200 {
bool host_codec_supported(int parquet_codec); // GZIP, LZ4, ZSTD, LZ4_RAW
// decompress exactly `unc` bytes; throws PlanError on malformed input, Unsupported when the codec library cannot be loaded
void host_decompress(int parquet_codec, const uint8_t* src, size_t n, uint8_t* dst, size_t unc);
} // namespace cb200
#endif
