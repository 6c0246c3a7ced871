// Mat.cpp -- PPM/PGM I/O and the three colour conversions the host classes need, for builds without
// OpenCV (see Mat.hpp).  Channel order follows OpenCV: a 3-channel Mat is BGR.
#ifndef RIP_HAVE_OPENCV
#include "Mat.hpp"

#include <cstdio>
#include <fstream>
#include <stdexcept>

namespace cv {

static bool read_token(std::istream &in, std::string &tok)
{
    tok.clear();
    int c;
    while ((c = in.get()) != EOF) {
        if (c == '#') {  // comment to end of line
            while ((c = in.get()) != EOF && c != '\n') {}
            continue;
        }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
            if (!tok.empty()) return true;
            continue;
        }
        tok.push_back((char)c);
    }
    return !tok.empty();
}

// OpenCV's 8-bit BGR->gray: fixed point, 14 fractional bits, round to nearest
static inline uchar bgr2gray_px(int b, int g, int r) { return (uchar)((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14); }

Mat imread(const std::string &path, int flags)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return Mat();
    std::string magic, sw, sh, smax;
    if (!read_token(in, magic) || (magic != "P6" && magic != "P5")) return Mat();
    if (!read_token(in, sw) || !read_token(in, sh) || !read_token(in, smax)) return Mat();
    const int w = std::atoi(sw.c_str()), h = std::atoi(sh.c_str());
    if (w <= 0 || h <= 0 || std::atoi(smax.c_str()) != 255) return Mat();
    const int file_cn = magic == "P6" ? 3 : 1;
    std::vector<uchar> buf((size_t)w * h * file_cn);
    in.read(reinterpret_cast<char *>(buf.data()), (std::streamsize)buf.size());
    if ((size_t)in.gcount() != buf.size()) return Mat();

    const bool want_gray = (flags == IMREAD_GRAYSCALE) || (flags == IMREAD_UNCHANGED && file_cn == 1);
    Mat out(h, w, want_gray ? CV_8UC1 : CV_8UC3);
    const size_t n = (size_t)w * h;
    if (file_cn == 3 && !want_gray) {
        for (size_t i = 0; i < n; i++) {  // PPM stores RGB, cv::Mat holds BGR
            out.data[3 * i + 0] = buf[3 * i + 2];
            out.data[3 * i + 1] = buf[3 * i + 1];
            out.data[3 * i + 2] = buf[3 * i + 0];
        }
    } else if (file_cn == 3) {
        for (size_t i = 0; i < n; i++) out.data[i] = bgr2gray_px(buf[3 * i + 2], buf[3 * i + 1], buf[3 * i + 0]);
    } else if (want_gray) {
        std::memcpy(out.data, buf.data(), n);
    } else {
        for (size_t i = 0; i < n; i++) out.data[3 * i] = out.data[3 * i + 1] = out.data[3 * i + 2] = buf[i];
    }
    return out;
}

bool imwrite(const std::string &path, const Mat &img)
{
    if (img.empty()) return false;
    std::ofstream out(path, std::ios::binary);
    if (!out) return false;
    const size_t n = img.total();
    if (img.channels() == 1) {
        out << "P5\n" << img.cols << " " << img.rows << "\n255\n";
        out.write(reinterpret_cast<const char *>(img.data), (std::streamsize)n);
    } else {
        const int cn = img.channels();  // 3: BGR, 4: BGRA -> RGB on disk
        out << "P6\n" << img.cols << " " << img.rows << "\n255\n";
        std::vector<uchar> rgb(n * 3);
        for (size_t i = 0; i < n; i++) {
            rgb[3 * i + 0] = img.data[cn * i + 2];
            rgb[3 * i + 1] = img.data[cn * i + 1];
            rgb[3 * i + 2] = img.data[cn * i + 0];
        }
        out.write(reinterpret_cast<const char *>(rgb.data()), (std::streamsize)rgb.size());
    }
    return (bool)out;
}

void cvtColor(const Mat &src, Mat &dst, int code)
{
    const size_t n = src.total();
    Mat out;
    switch (code) {
    case COLOR_BGR2RGBA:
        if (src.channels() != 3) throw std::runtime_error("cvtColor(BGR2RGBA): source must have 3 channels");
        out.create(src.rows, src.cols, CV_8UC4);
        for (size_t i = 0; i < n; i++) {
            out.data[4 * i + 0] = src.data[3 * i + 2];
            out.data[4 * i + 1] = src.data[3 * i + 1];
            out.data[4 * i + 2] = src.data[3 * i + 0];
            out.data[4 * i + 3] = 255;
        }
        break;
    case COLOR_RGBA2BGR:
        if (src.channels() != 4) throw std::runtime_error("cvtColor(RGBA2BGR): source must have 4 channels");
        out.create(src.rows, src.cols, CV_8UC3);
        for (size_t i = 0; i < n; i++) {
            out.data[3 * i + 0] = src.data[4 * i + 2];
            out.data[3 * i + 1] = src.data[4 * i + 1];
            out.data[3 * i + 2] = src.data[4 * i + 0];
        }
        break;
    case COLOR_BGR2RGB:
        if (src.channels() != 3) throw std::runtime_error("cvtColor(BGR2RGB): source must have 3 channels");
        out.create(src.rows, src.cols, CV_8UC3);
        for (size_t i = 0; i < n; i++) {
            out.data[3 * i + 0] = src.data[3 * i + 2];
            out.data[3 * i + 1] = src.data[3 * i + 1];
            out.data[3 * i + 2] = src.data[3 * i + 0];
        }
        break;
    case COLOR_BGR2GRAY:   // 3 or 4 channels, first three are B,G,R
    case COLOR_RGBA2GRAY:  // first three are R,G,B
    {
        const int cn = src.channels();
        if (cn < 3) throw std::runtime_error("cvtColor(*2GRAY): source must have 3 or 4 channels");
        out.create(src.rows, src.cols, CV_8UC1);
        for (size_t i = 0; i < n; i++) {
            const uchar *p = src.data + cn * i;
            out.data[i] = code == COLOR_BGR2GRAY ? bgr2gray_px(p[0], p[1], p[2]) : bgr2gray_px(p[2], p[1], p[0]);
        }
        break;
    }
    default:
        throw std::runtime_error("cvtColor: conversion code not provided by the Mat shim");
    }
    dst = out;
}

}  // namespace cv
#endif
