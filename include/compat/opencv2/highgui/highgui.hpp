/*
 * cv::imread / cv::glob as the sobfu application uses them (src/apps/demo.cpp:191-199,301-309), on top of the
 * dependency-free PNG reader of sobfu_b200_io.hpp:
 *   imread(path, CV_LOAD_IMAGE_ANYDEPTH)  16-bit depth maps stay 16-bit (CV_16UC1), colour is reduced to grey
 *   imread(path, CV_LOAD_IMAGE_COLOR)     8-bit, 3 channels, B G R order
 *   imread(path, CV_8U) (== grey scale)   8-bit, 1 channel
 * A file that cannot be read gives an empty Mat (data == nullptr), as in OpenCV.  PNG only (what the VolumeDeform /
 * KillingFusion sequences ship); with OpenCV installed put it first on the include path.
 */
#pragma once
#include <opencv2/core/core.hpp>
#include <sobfu_b200_io.hpp>

#include <string>
#include <vector>

#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
#define CV_LOAD_IMAGE_ANYDEPTH 2
#define CV_LOAD_IMAGE_ANYCOLOR 4

namespace cv {

typedef std::string String;

enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1, IMREAD_ANYDEPTH = 2, IMREAD_ANYCOLOR = 4 };

/* every regular file of the directory (OpenCV's glob of a directory pattern is not recursive by default) */
inline void glob(const String &dir, std::vector<String> &result, bool /*recursive*/ = false) { result = sobfu_b200::io::list_files(dir); }

inline Mat imread(const String &path, int flags = IMREAD_COLOR) {
    sobfu_b200::io::Image im;
    try {
        im = sobfu_b200::io::read_png(path);
    } catch (const std::exception &) {
        return Mat();
    }
    const bool keep_depth = flags == IMREAD_UNCHANGED || (flags & IMREAD_ANYDEPTH);
    const bool want_colour = flags != IMREAD_UNCHANGED && (flags & IMREAD_COLOR);
    const int src_ch = im.channels, colour_ch = src_ch >= 3 ? 3 : 1;
    int out_ch = flags == IMREAD_UNCHANGED ? src_ch : (want_colour ? 3 : 1);
    const bool out16 = im.bit_depth == 16 && keep_depth;
    Mat m(im.height, im.width, CV_MAKETYPE(out16 ? CV_16U : CV_8U, out_ch));
    const size_t n = (size_t)im.width * im.height;
    auto sample = [&](size_t i, int c) -> unsigned {      /* channel c of pixel i at the OUTPUT depth */
        if (im.bit_depth == 16) {
            const unsigned v = reinterpret_cast<const uint16_t *>(im.data.data())[i * src_ch + c];
            return out16 ? v : (v >> 8);
        }
        return im.data[i * src_ch + c];
    };
    for (size_t i = 0; i < n; ++i) {
        unsigned px[4] = {0, 0, 0, 0};
        if (flags == IMREAD_UNCHANGED) {
            for (int c = 0; c < src_ch; ++c) px[c] = sample(i, c);
            if (src_ch >= 3) std::swap(px[0], px[2]);                       /* R G B (A) -> B G R (A) */
        } else if (want_colour) {
            if (colour_ch == 3) { px[0] = sample(i, 2); px[1] = sample(i, 1); px[2] = sample(i, 0); }
            else px[0] = px[1] = px[2] = sample(i, 0);
        } else {
            if (colour_ch == 3) {   /* OpenCV's fixed-point BGR -> grey: (R*4899 + G*9617 + B*1868 + 8192) >> 14 */
                px[0] = (sample(i, 0) * 4899u + sample(i, 1) * 9617u + sample(i, 2) * 1868u + 8192u) >> 14;
            } else px[0] = sample(i, 0);
        }
        if (out16) for (int c = 0; c < out_ch; ++c) reinterpret_cast<uint16_t *>(m.data)[i * out_ch + c] = (uint16_t)px[c];
        else for (int c = 0; c < out_ch; ++c) m.data[i * out_ch + c] = (unsigned char)px[c];
    }
    return m;
}

inline bool imwrite(const String &path, const Mat &m) {
    if (m.empty() || !(m.depth() == CV_8U || m.depth() == CV_16U)) return false;
    try {
        if (m.channels() >= 3) {      /* B G R (A) -> R G B (A) */
            Mat t = m.clone();
            const size_t n = (size_t)m.rows * m.cols;
            const int ch = m.channels();
            if (m.depth() == CV_8U) for (size_t i = 0; i < n; ++i) std::swap(t.data[i * ch], t.data[i * ch + 2]);
            else for (size_t i = 0; i < n; ++i) std::swap(reinterpret_cast<uint16_t *>(t.data)[i * ch], reinterpret_cast<uint16_t *>(t.data)[i * ch + 2]);
            sobfu_b200::io::write_png(path, t.data, t.cols, t.rows, ch, m.depth() == CV_16U ? 16 : 8, t.step);
        } else {
            sobfu_b200::io::write_png(path, m.data, m.cols, m.rows, m.channels(), m.depth() == CV_16U ? 16 : 8, m.step);
        }
    } catch (const std::exception &) {
        return false;
    }
    return true;
}

}  // namespace cv
