// png.hpp -- PNG read (to RGB8) and write (RGB8 / GREY8) on top of zlib.
//
// The reference uses a vendored lodepng (reference BHRayTracer/LoadPNG/): textures are decoded
// with lodepng::decode(..., LCT_RGB) (Textures/Texture.cpp:76) = forced 8-bit RGB, alpha
// dropped, 16-bit samples reduced to their high byte; images are saved as 8-bit RGB / grey
// (Scenes/scene.h:634-644).  This file provides the same two conversions.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace bhrt { namespace host {

inline uint32_t png_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Decodes a non-interlaced PNG of any colour type / bit depth into tightly packed RGB8.
inline bool png_load_rgb8(const char *path, std::vector<uint8_t> &rgb, int &width, int &height, std::string *err = nullptr)
{
    auto fail = [&](const char *m) { if (err) *err = std::string(path) + ": " + m; return false; };
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail("cannot open");
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(fp);
    static const uint8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) return fail("not a PNG");
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool gotIHDR = false;
    while (pos + 12 <= file.size()) {
        uint32_t len = png_be32(&file[pos]);
        const uint8_t *type = &file[pos + 4];
        const uint8_t *data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return fail("truncated chunk");
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail("bad IHDR");
            w = png_be32(data); h = png_be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            gotIHDR = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!gotIHDR || w == 0 || h == 0) return fail("missing IHDR");
    if (interlace) return fail("interlaced PNG not supported");
    int channels;
    switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return fail("bad colour type");
    }
    size_t bpp_bits = (size_t)channels * depth;
    size_t stride = (w * bpp_bits + 7) / 8;
    size_t bpp = (bpp_bits + 7) / 8;            // filter byte distance
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return fail("inflate failed");
    // un-filter in place
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *in = &raw[(stride + 1) * y];
        uint8_t ft = in[0];
        in++;
        uint8_t *out = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= bpp ? out[x - bpp] : 0;
            int b = up ? up[x] : 0;
            int c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: {
                int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: return fail("bad filter type");
            }
            out[x] = (uint8_t)v;
        }
    }
    width = (int)w; height = (int)h;
    rgb.resize((size_t)w * h * 3);
    auto sample = [&](const uint8_t *row, size_t idx) -> unsigned {   // idx = sample index within the row
        if (depth == 8) return row[idx];
        if (depth == 16) return row[idx * 2];                          // high byte
        size_t bit = idx * depth;
        unsigned v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
        return v;
    };
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = &img[stride * y];
        uint8_t *o = &rgb[(size_t)y * w * 3];
        for (uint32_t x = 0; x < w; x++) {
            unsigned r, g, b;
            if (ctype == 3) {
                unsigned i = sample(row, x);
                if ((size_t)i * 3 + 2 < plte.size()) { r = plte[i * 3]; g = plte[i * 3 + 1]; b = plte[i * 3 + 2]; }
                else r = g = b = 0;
            } else if (ctype == 0 || ctype == 4) {
                unsigned v = sample(row, (size_t)x * channels);
                if (depth < 8) v = v * 255 / ((1u << depth) - 1);
                r = g = b = v;
            } else {
                r = sample(row, (size_t)x * channels);
                g = sample(row, (size_t)x * channels + 1);
                b = sample(row, (size_t)x * channels + 2);
            }
            o[x * 3] = (uint8_t)r; o[x * 3 + 1] = (uint8_t)g; o[x * 3 + 2] = (uint8_t)b;
        }
    }
    return true;
}

// Writes an 8-bit PNG with `comp` = 1 (grey) or 3 (RGB) components per pixel.
inline bool png_save(const char *path, const uint8_t *data, int width, int height, int comp)
{
    if (comp != 1 && comp != 3) return false;
    size_t stride = (size_t)width * comp;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;
        memcpy(&raw[(stride + 1) * y + 1], data + stride * y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp_data(clen);
    if (compress2(comp_data.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE *fp = fopen(path, "wb");
    if (!fp) return false;
    auto chunk = [&](const char *type, const uint8_t *d, uint32_t len) {
        uint8_t hdr[8] = { (uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                           (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3] };
        fwrite(hdr, 1, 8, fp);
        if (len) fwrite(d, 1, len, fp);
        uLong crc = crc32(0, hdr + 4, 4);
        if (len) crc = crc32(crc, d, len);
        uint8_t c[4] = { (uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc };
        fwrite(c, 1, 4, fp);
    };
    static const uint8_t sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    fwrite(sig, 1, 8, fp);
    uint8_t ihdr[13] = { (uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                         (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                         8, (uint8_t)(comp == 3 ? 2 : 0), 0, 0, 0 };
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp_data.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    fclose(fp);
    return true;
}

}}  // namespace
