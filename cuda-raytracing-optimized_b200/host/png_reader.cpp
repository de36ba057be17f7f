// png_reader.cpp -- PNG -> 8-bit RGB, the one image format the reference's scene loader reads.
//
// The reference decodes its nine texture files with the vendored stb_image.h (staircase_scene.h:103-118:
// stbi_load(file, &w, &h, &n, 3) after stbi_set_flip_vertically_on_load(true), then data[i] / 255.0f). This is that step
// rebuilt from the PNG and DEFLATE specifications (RFC 2083 / RFC 1950 / RFC 1951), not from stb's code: a table-free
// canonical-Huffman inflate, the five scanline filters, all colour types and bit depths, Adam7 interlacing. The conversions
// follow what stbi_load(.., 3) returns so that a texture decodes to the same bytes: grey -> r = g = b, alpha dropped,
// 16-bit samples reduced to their high byte, palette entries expanded.
#include "png_reader.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace crt {
namespace {

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t hold = 0;
    int bits = 0;
    bool overrun = false;
    uint32_t take(int count) { // LSB first (RFC 1951 3.1.1)
        while (bits < count) {
            uint32_t byte = 0;
            if (pos < n) byte = p[pos++];
            else overrun = true;
            hold |= byte << bits;
            bits += 8;
        }
        const uint32_t v = hold & ((count == 32) ? 0xFFFFFFFFu : ((1u << count) - 1u));
        hold = count == 32 ? 0 : hold >> count;
        bits -= count;
        return v;
    }
    void alignToByte() { hold = 0; bits = 0; }
};

// Canonical Huffman code (RFC 1951 3.2.2) decoded bit by bit from the per-length counts: first code of each length and the
// symbols sorted by (length, value). No lookup tables: textures are decoded once at scene load.
struct Huffman {
    uint16_t count[16] = {0};
    std::vector<uint16_t> symbols;
    bool build(const uint8_t* lengths, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[lengths[i]]++;
        count[0] = 0;
        int left = 1;
        for (int len = 1; len < 16; len++) {
            left = (left << 1) - count[len];
            if (left < 0) return false; // over-subscribed
        }
        uint16_t offset[16];
        offset[1] = 0;
        for (int len = 1; len < 15; len++) offset[len + 1] = offset[len] + count[len];
        symbols.assign(n, 0);
        for (int i = 0; i < n; i++)
            if (lengths[i]) symbols[offset[lengths[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; len++) {
            code |= (int)br.take(1);
            const int c = count[len];
            if (code - c < first) return symbols[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
            if (br.overrun) return -1;
        }
        return -1;
    }
};

const uint16_t kLengthBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLengthExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflateBlock(BitReader& br, const Huffman& lit, const Huffman& dist, std::vector<uint8_t>& out, size_t limit) {
    while (true) {
        if (out.size() > limit) return false; // more data than the image header allows: not a texture, stop inflating
        const int sym = lit.decode(br);
        if (sym < 0) return false;
        if (sym < 256) {
            out.push_back((uint8_t)sym);
        } else if (sym == 256) {
            return true;
        } else {
            const int li = sym - 257;
            if (li >= 29) return false;
            const int length = kLengthBase[li] + (int)br.take(kLengthExtra[li]);
            const int ds = dist.decode(br);
            if (ds < 0 || ds >= 30) return false;
            const size_t distance = kDistBase[ds] + br.take(kDistExtra[ds]);
            if (distance > out.size()) return false;
            const size_t from = out.size() - distance;
            for (int k = 0; k < length; k++) out.push_back(out[from + k]); // may overlap its own output: byte by byte
        }
        if (br.overrun) return false;
    }
}

// zlib stream (RFC 1950) holding deflate blocks (RFC 1951).
// `limit`: the largest output the caller accepts (the scanline bytes the PNG header implies); a stream that inflates to more fails.
bool inflateZlib(const uint8_t* data, size_t n, std::vector<uint8_t>& out, size_t limit) {
    if (n < 6) return false;
    if ((data[0] & 0x0F) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 0x20)) return false; // deflate, check bits, no preset dictionary
    BitReader br{data + 2, n - 2};
    bool last = false;
    while (!last) {
        last = br.take(1) != 0;
        const uint32_t type = br.take(2);
        if (type == 0) { // stored
            br.alignToByte();
            if (br.pos + 4 > br.n) return false;
            const uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || br.pos + len > br.n || out.size() + len > limit + 258) return false;
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
        } else if (type == 1) { // fixed codes (3.2.6)
            uint8_t l[288], d[30];
            for (int i = 0; i < 144; i++) l[i] = 8;
            for (int i = 144; i < 256; i++) l[i] = 9;
            for (int i = 256; i < 280; i++) l[i] = 7;
            for (int i = 280; i < 288; i++) l[i] = 8;
            for (int i = 0; i < 30; i++) d[i] = 5;
            Huffman lit, dist;
            lit.build(l, 288);
            dist.build(d, 30);
            if (!inflateBlock(br, lit, dist, out, limit)) return false;
        } else if (type == 2) { // dynamic codes (3.2.7)
            const int nlit = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
            if (nlit > 286 || ndist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (int i = 0; i < ncode; i++) cl[order[i]] = (uint8_t)br.take(3);
            Huffman code;
            if (!code.build(cl, 19)) return false;
            uint8_t lengths[286 + 30] = {0};
            int i = 0;
            while (i < nlit + ndist) {
                const int sym = code.decode(br);
                if (sym < 0) return false;
                if (sym < 16) {
                    lengths[i++] = (uint8_t)sym;
                } else {
                    int repeat;
                    uint8_t value = 0;
                    if (sym == 16) {
                        if (i == 0) return false;
                        value = lengths[i - 1];
                        repeat = 3 + (int)br.take(2);
                    } else if (sym == 17) {
                        repeat = 3 + (int)br.take(3);
                    } else {
                        repeat = 11 + (int)br.take(7);
                    }
                    if (i + repeat > nlit + ndist) return false;
                    while (repeat--) lengths[i++] = value;
                }
            }
            if (lengths[256] == 0) return false; // no end-of-block code
            Huffman lit, dist;
            if (!lit.build(lengths, nlit) || !dist.build(lengths + nlit, ndist)) return false;
            if (!inflateBlock(br, lit, dist, out, limit)) return false;
        } else {
            return false;
        }
        if (br.overrun) return false;
    }
    return true;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c) { // RFC 2083 6.6
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Reverses the scanline filters of one (sub)image in place. `rows` lines of 1 + stride bytes; bpp = bytes per complete pixel
// (at least 1).
bool unfilter(uint8_t* data, size_t rows, size_t stride, size_t bpp) {
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (size_t y = 0; y < rows; y++) {
        uint8_t* line = data + y * (stride + 1);
        const uint8_t type = line[0];
        uint8_t* cur = line + 1;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int add;
            switch (type) {
            case 0: add = 0; break;
            case 1: add = a; break;
            case 2: add = b; break;
            case 3: add = (a + b) >> 1; break;
            case 4: add = paeth(a, b, c); break;
            default: return false;
            }
            cur[i] = (uint8_t)(cur[i] + add);
        }
        prev = cur;
    }
    return true;
}

struct Header {
    uint32_t width = 0, height = 0;
    int depth = 0, colour = 0, interlace = 0;
    int channels() const { return colour == 0 ? 1 : colour == 2 ? 3 : colour == 3 ? 1 : colour == 4 ? 2 : 4; }
};

// One sample of a scanline as an 8-bit value (16-bit: high byte; 1/2/4-bit grey: scaled to 0..255; palette: the index).
inline uint8_t sampleAt(const uint8_t* line, const Header& h, size_t index) {
    switch (h.depth) {
    case 8: return line[index];
    case 16: return line[2 * index];
    default: {
        const int perByte = 8 / h.depth;
        const uint8_t byte = line[index / perByte];
        const int shift = 8 - h.depth * (int)(index % perByte + 1);
        const uint8_t v = (uint8_t)((byte >> shift) & ((1 << h.depth) - 1));
        return h.colour == 3 ? v : (uint8_t)(v * (255 / ((1 << h.depth) - 1)));
    }
    }
}

// Writes the pixels of one unfiltered (sub)image into the RGB frame at (x0 + i*dx, y0 + j*dy).
void scatterPass(const uint8_t* data, const Header& h, const std::vector<uint8_t>& palette, size_t w, size_t rows, size_t stride, size_t x0, size_t y0,
                 size_t dx, size_t dy, uint8_t* rgb) {
    const int ch = h.channels();
    for (size_t j = 0; j < rows; j++) {
        const uint8_t* line = data + j * (stride + 1) + 1;
        uint8_t* outRow = rgb + (y0 + j * dy) * (size_t)h.width * 3;
        for (size_t i = 0; i < w; i++) {
            uint8_t* px = outRow + (x0 + i * dx) * 3;
            if (h.colour == 3) {
                const size_t idx = sampleAt(line, h, i);
                const bool ok = (idx + 1) * 3 <= palette.size();
                px[0] = ok ? palette[3 * idx] : 0;
                px[1] = ok ? palette[3 * idx + 1] : 0;
                px[2] = ok ? palette[3 * idx + 2] : 0;
            } else if (ch <= 2) { // grey (+ alpha): r = g = b
                px[0] = px[1] = px[2] = sampleAt(line, h, i * ch);
            } else { // rgb (+ alpha)
                px[0] = sampleAt(line, h, i * ch);
                px[1] = sampleAt(line, h, i * ch + 1);
                px[2] = sampleAt(line, h, i * ch + 2);
            }
        }
    }
}

} // namespace

bool decodePng(const uint8_t* bytes, size_t n, bool flipVertically, int& width, int& height, std::vector<uint8_t>& rgb) {
    static const uint8_t kSignature[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || std::memcmp(bytes, kSignature, 8) != 0) return false;
    Header h;
    std::vector<uint8_t> idat, palette;
    bool haveHeader = false, ended = false;
    size_t pos = 8;
    while (pos + 12 <= n && !ended) {
        const uint32_t len = be32(bytes + pos);
        const uint8_t* type = bytes + pos + 4;
        const uint8_t* body = bytes + pos + 8;
        if (len > n - pos - 12) return false;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            h.width = be32(body);
            h.height = be32(body + 4);
            h.depth = body[8];
            h.colour = body[9];
            h.interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || h.interlace > 1) return false;
            haveHeader = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            ended = true;
        }
        pos += 12 + (size_t)len; // (chunk CRCs are not checked: a damaged stream fails in inflate or in the size check below)
    }
    if (!haveHeader || h.width == 0 || h.height == 0 || h.width > (1u << 24) || h.height > (1u << 24)) return false;
    const bool depthOk = (h.colour == 0 && (h.depth == 1 || h.depth == 2 || h.depth == 4 || h.depth == 8 || h.depth == 16)) ||
                         (h.colour == 3 && (h.depth == 1 || h.depth == 2 || h.depth == 4 || h.depth == 8)) ||
                         ((h.colour == 2 || h.colour == 4 || h.colour == 6) && (h.depth == 8 || h.depth == 16));
    if (!depthOk || (h.colour == 3 && palette.empty())) return false;

    const size_t bitsPerPixel = (size_t)h.depth * h.channels();
    // scanline bytes the header implies (filter byte + packed row, per row; Adam7: 1.875 filter bytes per image row and at most one
    // partial byte per pass row, i.e. under 4 more bytes per image row): what the zlib stream may inflate to
    const size_t rawLimit = (((size_t)h.width * bitsPerPixel + 7) / 8 + 1) * h.height + (h.interlace ? 4 * (size_t)h.height + 64 : 0);
    std::vector<uint8_t> raw;
    raw.reserve(rawLimit);
    if (!inflateZlib(idat.data(), idat.size(), raw, rawLimit)) return false;
    const size_t bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
    rgb.assign((size_t)h.width * h.height * 3, 0);
    if (!h.interlace) {
        const size_t stride = ((size_t)h.width * bitsPerPixel + 7) / 8;
        if (raw.size() < (stride + 1) * h.height) return false;
        if (!unfilter(raw.data(), h.height, stride, bpp)) return false;
        scatterPass(raw.data(), h, palette, h.width, h.height, stride, 0, 0, 1, 1, rgb.data());
    } else { // Adam7 (RFC 2083 2.6): seven sub-images, each filtered on its own
        static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
        size_t offset = 0;
        for (int p = 0; p < 7; p++) {
            const size_t w = (h.width + dx[p] - 1 - x0[p]) / dx[p], rows = (h.height + dy[p] - 1 - y0[p]) / dy[p];
            if ((size_t)x0[p] >= h.width || (size_t)y0[p] >= h.height || w == 0 || rows == 0) continue;
            const size_t stride = (w * bitsPerPixel + 7) / 8;
            if (raw.size() < offset + (stride + 1) * rows) return false;
            if (!unfilter(raw.data() + offset, rows, stride, bpp)) return false;
            scatterPass(raw.data() + offset, h, palette, w, rows, stride, x0[p], y0[p], dx[p], dy[p], rgb.data());
            offset += (stride + 1) * rows;
        }
    }
    if (flipVertically) { // stbi_set_flip_vertically_on_load(true), staircase_scene.h:120
        const size_t line = (size_t)h.width * 3;
        std::vector<uint8_t> tmp(line);
        for (size_t y = 0; y < h.height / 2; y++) {
            uint8_t* a = rgb.data() + y * line;
            uint8_t* b = rgb.data() + (h.height - 1 - y) * line;
            std::memcpy(tmp.data(), a, line);
            std::memcpy(a, b, line);
            std::memcpy(b, tmp.data(), line);
        }
    }
    width = (int)h.width;
    height = (int)h.height;
    return true;
}

bool readPngFile(const char* path, bool flipVertically, int& width, int& height, std::vector<uint8_t>& rgb) {
    std::FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::vector<uint8_t> bytes;
    uint8_t buf[65536];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + got);
    std::fclose(f);
    return decodePng(bytes.data(), bytes.size(), flipVertically, width, height, rgb);
}

} // namespace crt
