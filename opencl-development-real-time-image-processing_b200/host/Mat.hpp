// Mat.hpp -- the slice of cv::Mat the reference's host classes touch (rows, cols, data, total(),
// channels(), ptr<T>(row), empty(), clone()) plus imread/imwrite for binary PPM/PGM, used ONLY when
// OpenCV's C++ headers are absent (they are absent in this image).  Build with -DRIP_HAVE_OPENCV to
// compile the host classes against the real <opencv2/opencv.hpp> instead.
#pragma once

#ifdef RIP_HAVE_OPENCV
#include <opencv2/opencv.hpp>
#else

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace cv {

typedef unsigned char uchar;

enum { CV_8UC1 = 0, CV_8UC3 = 16, CV_8UC4 = 24 };
enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { COLOR_BGR2RGBA = 2, COLOR_RGBA2BGR = 3, COLOR_BGR2RGB = 4, COLOR_RGB2BGR = 4, COLOR_BGR2GRAY = 6, COLOR_RGBA2GRAY = 11 };

inline int channels_of_type(int type) { return (type >> 3) + 1; }
inline int type_of_channels(int cn) { return (cn - 1) << 3; }

class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    // wraps caller-owned memory (no copy), like cv::Mat(rows, cols, type, void*)
    Mat(int r, int c, int type, void *external) : rows(r), cols(c), data(static_cast<uchar *>(external)), m_cn(channels_of_type(type)) {}

    void create(int r, int c, int type)
    {
        rows = r; cols = c; m_cn = channels_of_type(type);
        m_store = std::make_shared<std::vector<uchar>>((size_t)r * c * m_cn);
        data = m_store->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int channels() const { return m_cn; }
    int type() const { return type_of_channels(m_cn); }
    size_t total() const { return (size_t)rows * cols; }
    size_t step() const { return (size_t)cols * m_cn; }
    template <typename T> T *ptr(int row = 0) { return reinterpret_cast<T *>(data + (size_t)row * step()); }
    template <typename T> const T *ptr(int row = 0) const { return reinterpret_cast<const T *>(data + (size_t)row * step()); }
    Mat clone() const
    {
        Mat m(rows, cols, type());
        if (!empty()) std::memcpy(m.data, data, total() * m_cn);
        return m;
    }

private:
    int m_cn = 1;
    std::shared_ptr<std::vector<uchar>> m_store;
};

// binary PPM (P6, 3 channels stored RGB -> returned BGR like OpenCV) and PGM (P5)
Mat imread(const std::string &path, int flags = IMREAD_COLOR);
bool imwrite(const std::string &path, const Mat &img);
void cvtColor(const Mat &src, Mat &dst, int code);

}  // namespace cv
#endif
