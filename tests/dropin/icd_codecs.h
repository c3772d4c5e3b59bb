/*
 * icd_codecs.h -- a stand-in for libicd (github.com/lucianpls/libicd, not vendored by the reference and not in this
 * image), just enough of its interface for the reference's own cqb3.cpp to compile unmodified: binary PNM (P5 / P6,
 * maxval 255 or 65535) where libicd reads PNG / JPEG, and PNM again where it writes PNG. Test infrastructure only:
 * tests/test_dropin.py builds /root/reference/cqb3.cpp against include/QB3.h + libQB3.so with this header on the
 * include path and round-trips a file through the binary. No codec arithmetic lives here.
 */
#ifndef ICD_CODECS_STUB_H
#define ICD_CODECS_STUB_H

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ICD {

enum ICDDataType { ICDT_Unknown = 0, ICDT_Byte = 1, ICDT_UInt16 = 2, ICDT_Short = 3 };
inline size_t getTypeSize(ICDDataType dt) { return dt == ICDT_Byte ? 1 : 2; }

struct sz5 { size_t x, y, z, c, l; };

struct Raster {
    sz5 size = {0, 0, 1, 0, 0};
    ICDDataType dt = ICDT_Byte;
};

struct storage_manager {
    storage_manager() : buffer(nullptr), size(0) {}
    storage_manager(void *p, size_t n) : buffer(static_cast<char *>(p)), size(n) {}
    char *buffer;
    size_t size;
};

struct codec_params {
    explicit codec_params(const Raster &r) : raster(r), line_stride(0) { error_message[0] = 0; }
    size_t get_buffer_size() const { return raster.size.x * raster.size.y * raster.size.c * getTypeSize(raster.dt); }
    Raster raster;
    size_t line_stride;
    char error_message[1024];
};
struct png_params : codec_params {
    explicit png_params(const Raster &r) : codec_params(r), compression_level(6) {}
    int compression_level;
};

/* "P5" / "P6" header: returns the payload offset or 0 */
inline size_t pnm_header(const storage_manager &src, Raster &r)
{
    if (src.size < 8 || src.buffer[0] != 'P' || (src.buffer[1] != '5' && src.buffer[1] != '6')) return 0;
    size_t at = 2, vals[3], n = 0;
    while (n < 3 && at < src.size) {
        while (at < src.size && (src.buffer[at] == ' ' || src.buffer[at] == '\n' || src.buffer[at] == '\t' || src.buffer[at] == '\r')) at++;
        if (at < src.size && src.buffer[at] == '#') { while (at < src.size && src.buffer[at] != '\n') at++; continue; }
        size_t v = 0, digits = 0;
        while (at < src.size && src.buffer[at] >= '0' && src.buffer[at] <= '9') { v = 10 * v + (src.buffer[at++] - '0'); digits++; }
        if (!digits) return 0;
        vals[n++] = v;
    }
    if (n < 3 || at >= src.size) return 0;
    r.size.x = vals[0]; r.size.y = vals[1]; r.size.z = 1; r.size.l = 0;
    r.size.c = src.buffer[1] == '6' ? 3 : 1;
    r.dt = vals[2] > 255 ? ICDT_UInt16 : ICDT_Byte;
    return at + 1; /* one white space byte after maxval */
}

inline const char *image_peek(const storage_manager &src, Raster &r)
{
    return pnm_header(src, r) ? nullptr : "not a binary PNM file (this build reads P5 / P6 in place of PNG / JPEG)";
}

/* decodes into buffer, host byte order */
inline const char *stride_decode(codec_params &params, storage_manager &src, void *buffer)
{
    Raster r;
    const size_t off = pnm_header(src, r), n = params.get_buffer_size();
    if (!off || off + n > src.size) return "truncated PNM file";
    memcpy(buffer, src.buffer + off, n);
    if (r.dt != ICDT_Byte) { /* PNM samples are big endian */
        uint8_t *p = static_cast<uint8_t *>(buffer);
        for (size_t i = 0; i + 1 < n; i += 2) { const uint8_t t = p[i]; p[i] = p[i + 1]; p[i + 1] = t; }
    }
    return nullptr;
}

/* writes PNM; 16 bit input arrives big endian (cqb3.cpp swaps it for libicd's PNG writer), which is what PNM wants */
inline const char *png_encode(png_params &params, storage_manager &src, storage_manager &dst)
{
    const Raster &r = params.raster;
    if (r.size.c != 1 && r.size.c != 3) return "PNM holds one or three bands";
    char head[64];
    const int hl = snprintf(head, sizeof(head), "P%c\n%zu %zu\n%d\n", r.size.c == 3 ? '6' : '5', r.size.x, r.size.y,
                            r.dt == ICDT_Byte ? 255 : 65535);
    if ((size_t)hl + src.size > dst.size) return "output buffer too small";
    memcpy(dst.buffer, head, hl);
    memcpy(dst.buffer + hl, src.buffer, src.size);
    dst.size = hl + src.size;
    return nullptr;
}

} // namespace ICD
#endif
