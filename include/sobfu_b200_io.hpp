// sobfu_b200_io.hpp -- host-side file formats of the sobfu application, dependency-free:
//   * PNG decoding (8/16-bit grey, grey+alpha, RGB, RGBA, palette; non-interlaced) incl. a zlib inflate -- replaces
//     cv::imread(path, CV_LOAD_IMAGE_ANYDEPTH | CV_LOAD_IMAGE_COLOR | CV_8U) of src/apps/demo.cpp:301-309 for the 16-bit depth
//     maps (millimetres), the colour frames and the object masks of the VolumeDeform / KillingFusion sequences
//   * PNG encoding (stored deflate blocks) for writing synthetic sequences
//   * legacy-VTK polydata writer equivalent to pcl::io::saveVTKFile (demo.cpp:237-246) and a VTK XML image-data writer for
//     the deformation field (demo.cpp:252-284)
//   * sorted directory listing (cv::glob + std::sort, demo.cpp:191-199)
// Nothing here touches the GPU; the header is usable without CUDA.
#pragma once
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sobfu_b200 {
namespace io {

struct io_error : std::runtime_error { explicit io_error(const std::string &w) : std::runtime_error(w) {} };

// ---------------------------------------------------------------------------------------------------------------
// inflate (RFC 1951) of a zlib stream (RFC 1950)
// ---------------------------------------------------------------------------------------------------------------
namespace detail {

struct BitReader {
    const uint8_t *p, *end;
    uint32_t buf = 0;
    int cnt = 0;
    BitReader(const uint8_t *b, const uint8_t *e) : p(b), end(e) {}
    uint32_t bits(int n) {
        while (cnt < n) {
            if (p >= end) throw io_error("inflate: unexpected end of stream");
            buf |= (uint32_t)(*p++) << cnt;
            cnt += 8;
        }
        const uint32_t v = n ? (buf & ((1u << n) - 1u)) : 0u;
        buf >>= n;
        cnt -= n;
        return v;
    }
    void align_byte() { buf = 0; cnt = 0; }
};

// canonical Huffman decoder: counts per length + symbols sorted by (length, value)
struct Huffman {
    uint16_t count[16];
    std::vector<uint16_t> symbol;
    void build(const uint8_t *lengths, int n) {
        std::memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i) ++count[lengths[i]];
        count[0] = 0;
        int left = 1;
        for (int len = 1; len < 16; ++len) {
            left = (left << 1) - count[len];
            if (left < 0) throw io_error("inflate: over-subscribed Huffman code");
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int len = 1; len < 15; ++len) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
        symbol.assign(n, 0);
        for (int i = 0; i < n; ++i)
            if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader &br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len) {
            code |= (int)br.bits(1);
            const int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        throw io_error("inflate: invalid Huffman code");
    }
};

inline void inflate_block(BitReader &br, const Huffman &lit, const Huffman &dist, std::vector<uint8_t> &out, size_t max_out) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
        const int sym = lit.decode(br);
        if (sym < 256) {
            if (out.size() >= max_out) throw io_error("inflate: output larger than expected");
            out.push_back((uint8_t)sym);
            continue;
        }
        if (sym == 256) return;
        const int li = sym - 257;
        if (li >= 29) throw io_error("inflate: invalid length symbol");
        const int len = lbase[li] + (int)br.bits(lext[li]);
        const int di = dist.decode(br);
        if (di >= 30) throw io_error("inflate: invalid distance symbol");
        const size_t d = dbase[di] + br.bits(dext[di]);
        if (d > out.size()) throw io_error("inflate: distance beyond the start of the output");
        if (out.size() + (size_t)len > max_out) throw io_error("inflate: output larger than expected");
        const size_t from = out.size() - d;
        for (int k = 0; k < len; ++k) out.push_back(out[from + k]);     // may overlap: byte by byte
    }
}

}  // namespace detail

// max_out bounds the output (a corrupt or hostile stream cannot make the reader allocate without limit); 0 = 1 GiB
inline std::vector<uint8_t> zlib_inflate(const uint8_t *data, size_t n, size_t size_hint = 0, size_t max_out = 0) {
    if (!max_out) max_out = (size_t)1 << 30;
    using namespace detail;
    if (n < 6) throw io_error("zlib: stream too short");
    if ((data[0] & 0x0f) != 8 || ((data[0] << 8 | data[1]) % 31) != 0 || (data[1] & 0x20)) throw io_error("zlib: bad header");
    BitReader br(data + 2, data + n);
    std::vector<uint8_t> out;
    out.reserve(std::min(size_hint, (size_t)64 << 20));
    Huffman fixed_lit, fixed_dist;
    bool have_fixed = false;
    for (bool last = false; !last;) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {
            br.align_byte();
            if (br.end - br.p < 4) throw io_error("inflate: truncated stored block");
            const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
            if ((len ^ 0xffffu) != nlen) throw io_error("inflate: stored block length check failed");
            br.p += 4;
            if ((size_t)(br.end - br.p) < len) throw io_error("inflate: truncated stored block");
            if (out.size() + len > max_out) throw io_error("inflate: output larger than expected");
            out.insert(out.end(), br.p, br.p + len);
            br.p += len;
        } else if (type == 1) {
            if (!have_fixed) {
                uint8_t l[288], d[30];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                for (int i = 0; i < 30; ++i) d[i] = 5;
                fixed_lit.build(l, 288);
                fixed_dist.build(d, 30);
                have_fixed = true;
            }
            inflate_block(br, fixed_lit, fixed_dist, out, max_out);
        } else if (type == 2) {
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (nlen > 286 || ndist > 30) throw io_error("inflate: bad code counts");
            uint8_t lengths[320];
            std::memset(lengths, 0, sizeof lengths);
            for (int i = 0; i < ncode; ++i) lengths[order[i]] = (uint8_t)br.bits(3);
            Huffman lencode;
            lencode.build(lengths, 19);
            std::memset(lengths, 0, sizeof lengths);
            for (int i = 0; i < nlen + ndist;) {
                const int sym = lencode.decode(br);
                if (sym < 16) { lengths[i++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) {
                    if (i == 0) throw io_error("inflate: repeat without a previous length");
                    val = lengths[i - 1];
                    rep = 3 + (int)br.bits(2);
                } else if (sym == 17) rep = 3 + (int)br.bits(3);
                else rep = 11 + (int)br.bits(7);
                if (i + rep > nlen + ndist) throw io_error("inflate: too many code lengths");
                while (rep--) lengths[i++] = (uint8_t)val;
            }
            if (lengths[256] == 0) throw io_error("inflate: no end-of-block code");
            Huffman lit, dist;
            lit.build(lengths, nlen);
            dist.build(lengths + nlen, ndist);
            inflate_block(br, lit, dist, out, max_out);
        } else {
            throw io_error("inflate: invalid block type");
        }
    }
    return out;
}

inline uint32_t adler32(const uint8_t *d, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; ++i) { a = (a + d[i]) % 65521u; b = (b + a) % 65521u; }
    return (b << 16) | a;
}
inline uint32_t crc32(const uint8_t *d, size_t n, uint32_t crc = 0) {
    struct Table {
        uint32_t v[256];
        Table() {
            for (uint32_t i = 0; i < 256; ++i) {
                uint32_t c = i;
                for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
                v[i] = c;
            }
        }
    };
    static const Table tab;                 // initialised once, thread-safe (C++11 magic static)
    const uint32_t *table = tab.v;
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}

// ---------------------------------------------------------------------------------------------------------------
// PNG
// ---------------------------------------------------------------------------------------------------------------
struct Image {
    int width = 0, height = 0, channels = 0, bit_depth = 0;   // channels as stored in the file (palette expanded to 3)
    std::vector<uint8_t> data;                                // row-major, interleaved, 16-bit samples in HOST byte order
    size_t step() const { return (size_t)width * channels * (bit_depth == 16 ? 2 : 1); }
    bool empty() const { return data.empty(); }
};

inline std::vector<uint8_t> read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw io_error("cannot open '" + path + "'");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

inline Image decode_png(const uint8_t *d, size_t n) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 8 || std::memcmp(d, sig, 8)) throw io_error("png: bad signature");
    auto be32 = [](const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
    Image im;
    int colour = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool end = false;
    for (size_t pos = 8; pos + 12 <= n && !end;) {
        const uint32_t len = be32(d + pos);
        if (pos + 12 + (size_t)len > n) throw io_error("png: truncated chunk");
        const uint8_t *type = d + pos + 4, *body = d + pos + 8;
        if (crc32(type, 4 + len) != be32(body + len)) throw io_error("png: chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) throw io_error("png: bad IHDR");
            im.width = (int)be32(body);
            im.height = (int)be32(body + 4);
            im.bit_depth = body[8];
            colour = body[9];
            interlace = body[12];
            if (body[10] != 0 || body[11] != 0) throw io_error("png: unknown compression / filter method");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (colour < 0 || im.width <= 0 || im.height <= 0) throw io_error("png: no IHDR");
    if (interlace) throw io_error("png: interlaced images are not supported");
    int ch;
    switch (colour) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 3: ch = 1; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: throw io_error("png: bad colour type");
    }
    const int bd = im.bit_depth;
    if (!(bd == 8 || bd == 16 || (bd < 8 && (colour == 0 || colour == 3) && (bd == 1 || bd == 2 || bd == 4)))) throw io_error("png: unsupported bit depth");
    if (colour == 3 && bd == 16) throw io_error("png: bad palette bit depth");
    const size_t bpp_bits = (size_t)ch * bd, stride = ((size_t)im.width * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
    if ((uint64_t)im.width * (uint64_t)im.height > ((uint64_t)1 << 28)) throw io_error("png: image too large");
    std::vector<uint8_t> raw = zlib_inflate(idat.data(), idat.size(), (stride + 1) * im.height, (stride + 1) * im.height);
    if (raw.size() < (stride + 1) * (size_t)im.height) throw io_error("png: image data too short");
    // undo the scanline filters in place
    std::vector<uint8_t> pix(stride * im.height);
    for (int y = 0; y < im.height; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *src = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &pix[stride * y];
        const uint8_t *up = y ? &pix[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? dst[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw io_error("png: bad filter type");
            }
            dst[x] = (uint8_t)v;
        }
    }
    // unpack to 8- or 16-bit samples
    if (colour == 3) {
        im.channels = 3;
        im.data.resize((size_t)im.width * im.height * 3);
        for (int y = 0; y < im.height; ++y)
            for (int x = 0; x < im.width; ++x) {
                const size_t bit = (size_t)x * bd;
                const int idx = bd == 8 ? pix[stride * y + x] : (pix[stride * y + bit / 8] >> (8 - bd - (bit % 8))) & ((1 << bd) - 1);
                if ((size_t)idx * 3 + 2 >= plte.size()) throw io_error("png: palette index out of range");
                std::memcpy(&im.data[((size_t)y * im.width + x) * 3], &plte[(size_t)idx * 3], 3);
            }
        im.bit_depth = 8;
    } else if (bd < 8) {
        im.channels = 1;
        im.data.resize((size_t)im.width * im.height);
        const int maxv = (1 << bd) - 1;
        for (int y = 0; y < im.height; ++y)
            for (int x = 0; x < im.width; ++x) {
                const size_t bit = (size_t)x * bd;
                const int v = (pix[stride * y + bit / 8] >> (8 - bd - (bit % 8))) & maxv;
                im.data[(size_t)y * im.width + x] = (uint8_t)(v * 255 / maxv);
            }
        im.bit_depth = 8;
    } else if (bd == 8) {
        im.channels = ch;
        im.data.swap(pix);
    } else {   // 16 bit: big endian in the file
        im.channels = ch;
        im.data.resize(pix.size());
        uint16_t *o = reinterpret_cast<uint16_t *>(im.data.data());
        for (size_t i = 0; i < pix.size() / 2; ++i) o[i] = (uint16_t)((pix[2 * i] << 8) | pix[2 * i + 1]);
    }
    return im;
}
inline Image read_png(const std::string &path) {
    const std::vector<uint8_t> f = read_file(path);
    try {
        return decode_png(f.data(), f.size());
    } catch (const io_error &e) {
        throw io_error(path + ": " + e.what());
    }
}

// PNG writer (8/16-bit, 1-4 channels), stored deflate blocks: exact and dependency-free, size is not a concern for depth maps
inline std::vector<uint8_t> encode_png(const void *pixels, int width, int height, int channels, int bit_depth, size_t step_bytes = 0) {
    if (!(bit_depth == 8 || bit_depth == 16) || channels < 1 || channels > 4) throw io_error("png: unsupported format for writing");
    static const int colour_of[5] = {0, 0, 4, 2, 6};
    const size_t row = (size_t)width * channels * (bit_depth / 8);
    if (!step_bytes) step_bytes = row;
    std::vector<uint8_t> raw((row + 1) * height);
    for (int y = 0; y < height; ++y) {
        uint8_t *dst = &raw[(row + 1) * y];
        *dst++ = 0;
        const uint8_t *src = static_cast<const uint8_t *>(pixels) + step_bytes * y;
        if (bit_depth == 8) std::memcpy(dst, src, row);
        else {
            const uint16_t *s = reinterpret_cast<const uint16_t *>(src);
            for (size_t i = 0; i < row / 2; ++i) { dst[2 * i] = (uint8_t)(s[i] >> 8); dst[2 * i + 1] = (uint8_t)(s[i] & 0xff); }
        }
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    for (size_t pos = 0; pos < raw.size() || pos == 0;) {
        const size_t len = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + len >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back((uint8_t)(len & 0xff)); z.push_back((uint8_t)(len >> 8));
        z.push_back((uint8_t)(~len & 0xff)); z.push_back((uint8_t)((~len >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
        pos += len;
        if (last) break;
    }
    const uint32_t ad = adler32(raw.data(), raw.size());
    for (int k = 3; k >= 0; --k) z.push_back((uint8_t)(ad >> (8 * k)));
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char *type, const std::vector<uint8_t> &body) {
        const uint32_t len = (uint32_t)body.size();
        for (int k = 3; k >= 0; --k) out.push_back((uint8_t)(len >> (8 * k)));
        std::vector<uint8_t> tb(type, type + 4);
        tb.insert(tb.end(), body.begin(), body.end());
        out.insert(out.end(), tb.begin(), tb.end());
        const uint32_t c = crc32(tb.data(), tb.size());
        for (int k = 3; k >= 0; --k) out.push_back((uint8_t)(c >> (8 * k)));
    };
    std::vector<uint8_t> ihdr(13, 0);
    for (int k = 0; k < 4; ++k) { ihdr[k] = (uint8_t)((uint32_t)width >> (8 * (3 - k))); ihdr[4 + k] = (uint8_t)((uint32_t)height >> (8 * (3 - k))); }
    ihdr[8] = (uint8_t)bit_depth;
    ihdr[9] = (uint8_t)colour_of[channels];
    chunk("IHDR", ihdr);
    chunk("IDAT", z);
    chunk("IEND", {});
    return out;
}
inline void write_png(const std::string &path, const void *pixels, int width, int height, int channels, int bit_depth, size_t step_bytes = 0) {
    const std::vector<uint8_t> f = encode_png(pixels, width, height, channels, bit_depth, step_bytes);
    std::ofstream o(path, std::ios::binary);
    if (!o || !o.write(reinterpret_cast<const char *>(f.data()), (std::streamsize)f.size())) throw io_error("cannot write '" + path + "'");
}

// ---------------------------------------------------------------------------------------------------------------
// file system helpers
// ---------------------------------------------------------------------------------------------------------------
inline bool path_exists(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_directory(const std::string &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool make_directory(const std::string &p) { return ::mkdir(p.c_str(), 0777) == 0; }     // true when it was created
// regular files of a directory as "dir/name", sorted (cv::glob(dir, out) + std::sort)
inline std::vector<std::string> list_files(const std::string &dir) {
    std::vector<std::string> out;
    DIR *d = ::opendir(dir.c_str());
    if (!d) return out;
    while (dirent *e = ::readdir(d)) {
        const std::string name = e->d_name;
        if (name == "." || name == "..") continue;
        const std::string full = dir + (dir.empty() || dir.back() == '/' ? "" : "/") + name;
        struct stat st;
        if (::stat(full.c_str(), &st) == 0 && S_ISREG(st.st_mode)) out.push_back(full);
    }
    ::closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// VTK
// ---------------------------------------------------------------------------------------------------------------
// Legacy ASCII polydata as pcl::io::saveVTKFile (pcl/io/vtk_io.cpp) lays it out: header, POINTS n float, one "x y z" per
// point, VERTICES n 2n with "1 i" per point, POLYGONS np total with "k i0 i1 ..." per polygon.
// xyz: n points, `stride_floats` floats apart (3 for packed xyz, 4 for pcl::PointXYZ / float4).
inline void write_vtk_polydata(const std::string &path, const float *xyz, size_t n_points, size_t stride_floats,
                               const uint32_t *polygons, size_t n_polygons, int verts_per_polygon = 3, int precision = 5) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw io_error("cannot write '" + path + "'");
    std::fprintf(f, "# vtk DataFile Version 3.0\nvtk output\nASCII\nDATASET POLYDATA\nPOINTS %zu float", n_points);
    char fmt[32];
    std::snprintf(fmt, sizeof fmt, "\n%%.%dg %%.%dg %%.%dg", precision, precision, precision);
    for (size_t i = 0; i < n_points; ++i) std::fprintf(f, fmt, xyz[i * stride_floats], xyz[i * stride_floats + 1], xyz[i * stride_floats + 2]);
    std::fprintf(f, "\nVERTICES %zu %zu", n_points, 2 * n_points);
    for (size_t i = 0; i < n_points; ++i) std::fprintf(f, "\n1 %zu", i);
    std::fprintf(f, "\nPOLYGONS %zu %zu", n_polygons, n_polygons * (size_t)(verts_per_polygon + 1));
    for (size_t i = 0; i < n_polygons; ++i) {
        std::fprintf(f, "\n%d", verts_per_polygon);
        for (int k = 0; k < verts_per_polygon; ++k) std::fprintf(f, " %u", polygons ? polygons[i * verts_per_polygon + k] : (uint32_t)(i * verts_per_polygon + k));
    }
    std::fprintf(f, "\n");
    if (std::fclose(f) != 0) throw io_error("cannot write '" + path + "'");
}

// VTK XML image data (.vti) with one float array of `components` per point, raw appended encoding -- what
// vtkXMLImageDataWriter produces for the deformation field of demo.cpp:252-284, uncompressed
inline void write_vti(const std::string &path, const float *data, int nx, int ny, int nz, int components, const char *name = "ImageScalars") {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw io_error("cannot write '" + path + "'");
    const uint64_t bytes = (uint64_t)nx * ny * nz * components * sizeof(float);
    std::fprintf(f,
                 "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
                 "  <ImageData WholeExtent=\"0 %d 0 %d 0 %d\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n    <Piece Extent=\"0 %d 0 %d 0 %d\">\n"
                 "      <PointData Scalars=\"%s\">\n        <DataArray type=\"Float32\" Name=\"%s\" NumberOfComponents=\"%d\" format=\"appended\" offset=\"0\"/>\n"
                 "      </PointData>\n      <CellData/>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n   _",
                 nx - 1, ny - 1, nz - 1, nx - 1, ny - 1, nz - 1, name, name, components);
    std::fwrite(&bytes, sizeof bytes, 1, f);
    std::fwrite(data, 1, (size_t)bytes, f);
    std::fprintf(f, "\n  </AppendedData>\n</VTKFile>\n");
    if (std::fclose(f) != 0) throw io_error("cannot write '" + path + "'");
}

// ---------------------------------------------------------------------------------------------------------------
// .ini files of the reference (params/*.ini): NAME=VALUE, '#' comments -- flat key/value view for callers that do not go
// through the boost::program_options-compatible layer (include/compat/boost/program_options.hpp)
// ---------------------------------------------------------------------------------------------------------------
inline std::vector<std::pair<std::string, std::string>> read_ini(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw io_error("cannot open '" + path + "'");
    std::vector<std::pair<std::string, std::string>> out;
    std::string line;
    auto trim = [](std::string s) {
        const size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    };
    while (std::getline(f, line)) {
        const size_t h = line.find('#');
        if (h != std::string::npos) line.erase(h);
        line = trim(line);
        if (line.empty()) continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) throw io_error(path + ": invalid line '" + line + "'");
        out.emplace_back(trim(line.substr(0, eq)), trim(line.substr(eq + 1)));
    }
    return out;
}

}  // namespace io
}  // namespace sobfu_b200
