// Comparator.cpp -- see Comparator.hpp.  Compile with -ffp-contract=off: the reference's CPU code is
// built without FMA contraction (no -march / -ffast-math in its CMake files), and the float sums below
// must round after every multiply and every add.
#include "Comparator.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

namespace {

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int reflect101(int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i)); }

void gray_rows(const cv::Mat &src, bool order_bgr, cv::Mat &dst)
{
    const int cn = src.channels();
    for (int y = 0; y < src.rows; y++) {
        const cv::uchar *s = src.ptr<cv::uchar>(y);
        cv::uchar *d = dst.ptr<cv::uchar>(y);
        for (int x = 0; x < src.cols; x++) {
            const cv::uchar c0 = s[cn * x], c1 = s[cn * x + 1], c2 = s[cn * x + 2];
            const double r = order_bgr ? c2 : c0, g = c1, b = order_bgr ? c0 : c2;
            d[x] = static_cast<cv::uchar>(0.299 * r + 0.587 * g + 0.114 * b);   // RT/src/Comparator.cpp:41
        }
    }
}

void blur_rows(const cv::Mat &src, const std::vector<float> &k, int ksize, cv::Mat &dst)
{
    const int cn = src.channels(), half = ksize / 2, W = src.cols, H = src.rows;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int c = 0; c < cn; c++) {
                float sum = 0.0f;
                for (int ky = -half; ky <= half; ky++) {
                    const cv::uchar *row = src.ptr<cv::uchar>(clampi(y + ky, 0, H - 1));
                    for (int kx = -half; kx <= half; kx++) {
                        const float px = (float)row[cn * clampi(x + kx, 0, W - 1) + c];
                        sum += px * k[(size_t)(ky + half) * ksize + (kx + half)];   // GaussianBlur.cpp:247-252
                    }
                }
                dst.ptr<cv::uchar>(y)[cn * x + c] = static_cast<cv::uchar>(std::min(std::max(sum, 0.0f), 255.0f));
            }
}

void sobel_rows(const cv::Mat &g, cv::Mat &dst)
{
    const int W = g.cols, H = g.rows;
    for (int y = 0; y < H; y++) {
        const cv::uchar *r0 = g.ptr<cv::uchar>(reflect101(y - 1, H)), *r1 = g.ptr<cv::uchar>(y), *r2 = g.ptr<cv::uchar>(reflect101(y + 1, H));
        cv::uchar *d = dst.ptr<cv::uchar>(y);
        for (int x = 0; x < W; x++) {
            const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
            const int gx = -r0[xl] + r0[xr] - 2 * r1[xl] + 2 * r1[xr] - r2[xl] + r2[xr];
            const int gy = -r0[xl] - 2 * r0[x] - r0[xr] + r2[xl] + 2 * r2[x] + r2[xr];
            const float m = std::sqrt((float)(gx * gx + gy * gy));                // cv::magnitude on CV_32F
            d[x] = (cv::uchar)std::min(255L, std::lrint(m));                      // convertTo(CV_8U): round half to even, saturate
        }
    }
}

template <typename F>
double time_iterations(int n, F &&body)
{
    n = std::max(n, 1);
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < n; i++) body();
    const auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / n;
}

}  // namespace

Comparator::Comparator(int num_methods, int num_iterations) : m_num_methods(num_methods), NUMBER_OF_ITERATIONS(num_iterations) {}

std::vector<float> Comparator::GaussianKernel(int kernel_size, float sigma)
{
    // exp() of a float argument is evaluated in double (unqualified ::exp), 2*M_PI*sigma*sigma likewise; each
    // weight is rounded to float, the sum accumulates in float, the normalisation is a float divide
    std::vector<float> k((size_t)kernel_size * kernel_size);
    const int half = kernel_size / 2;
    float sum = 0.0f;
    for (int y = -half; y <= half; y++)
        for (int x = -half; x <= half; x++) {
            const float arg = -(float)(x * x + y * y) / (2.0f * sigma * sigma);
            const float v = (float)(std::exp((double)arg) / (2.0 * M_PI * (double)sigma * (double)sigma));
            k[(size_t)(y + half) * kernel_size + (x + half)] = v;
            sum += v;
        }
    for (float &v : k) v /= sum;
    return k;
}

cv::Mat Comparator::PerformCPU_Grayscaling(std::string image_path, double &avg_cpu_execution_time, Logger &logger)
{
    cv::Mat image = cv::imread(image_path, cv::IMREAD_COLOR);   // BGR, like the reference (Comparator.cpp:14)
    if (image.empty()) {
        logger.log("Failed to load image " + image_path, Logger::LogLevel::ERROR);
        avg_cpu_execution_time = 0.0;
        return cv::Mat();
    }
    return PerformCPU_Grayscaling(image, true, avg_cpu_execution_time, logger);
}

cv::Mat Comparator::PerformCPU_Grayscaling(const cv::Mat &image, bool order_bgr, double &avg_cpu_execution_time, Logger &logger)
{
    if (image.empty() || image.channels() < 3) {
        logger.log("PerformCPU_Grayscaling needs a 3- or 4-channel image", Logger::LogLevel::ERROR);
        return cv::Mat();
    }
    cv::Mat out(image.rows, image.cols, cv::CV_8UC1);
    avg_cpu_execution_time = time_iterations(NUMBER_OF_ITERATIONS, [&] { gray_rows(image, order_bgr, out); });
    return out;
}

cv::Mat Comparator::PerformCPU_GaussianBlur(const cv::Mat &image, int kernel_size, float kernel_sigma, double &avg_cpu_execution_time,
                                            Logger &logger)
{
    if (image.empty() || kernel_size < 1 || !(kernel_size & 1)) {
        logger.log("PerformCPU_GaussianBlur needs a non-empty image and an odd kernel size", Logger::LogLevel::ERROR);
        return cv::Mat();
    }
    cv::Mat out(image.rows, image.cols, image.type());
    avg_cpu_execution_time = time_iterations(NUMBER_OF_ITERATIONS, [&] {
        const std::vector<float> k = GaussianKernel(kernel_size, kernel_sigma);   // regenerated inside the timed loop (GaussianBlur.cpp:230)
        blur_rows(image, k, kernel_size, out);
    });
    return out;
}

cv::Mat Comparator::PerformCPU_EdgeDetection(const cv::Mat &gray, double &avg_cpu_execution_time, Logger &logger)
{
    if (gray.empty() || gray.channels() != 1) {
        logger.log("PerformCPU_EdgeDetection needs a single-channel image", Logger::LogLevel::ERROR);
        return cv::Mat();
    }
    cv::Mat out(gray.rows, gray.cols, cv::CV_8UC1);
    avg_cpu_execution_time = time_iterations(NUMBER_OF_ITERATIONS, [&] { sobel_rows(gray, out); });
    return out;
}

cv::Mat Comparator::PerformCPU_Fused(const cv::Mat &image, bool order_bgr, int kernel_size, float kernel_sigma, double &avg_cpu_execution_time,
                                     Logger &logger)
{
    if (image.empty() || image.channels() < 3) {
        logger.log("PerformCPU_Fused needs a 3- or 4-channel image", Logger::LogLevel::ERROR);
        return cv::Mat();
    }
    // the composition of the three CPU stages with a u8 image between them
    cv::Mat gray(image.rows, image.cols, cv::CV_8UC1), blurred(image.rows, image.cols, cv::CV_8UC1), out(image.rows, image.cols, cv::CV_8UC1);
    avg_cpu_execution_time = time_iterations(NUMBER_OF_ITERATIONS, [&] {
        gray_rows(image, order_bgr, gray);
        const std::vector<float> k = GaussianKernel(kernel_size, kernel_sigma);
        blur_rows(gray, k, kernel_size, blurred);
        sobel_rows(blurred, out);
    });
    return out;
}

double Comparator::ComputeMAE(const cv::Mat &reference, const cv::Mat &result, Logger &logger)
{
    if (reference.empty() || result.empty() || reference.rows != result.rows || reference.cols != result.cols) {
        logger.log("ComputeMAE: images are empty or differ in size", Logger::LogLevel::ERROR);
        return -1.0;
    }
    // channel 0 of each (cv::mean(absdiff(...))[0], Comparator.cpp:97-100); a 4-channel result against a 1-channel
    // reference compares its first channel, which is the gray value in the (g,g,g,255) container
    const int ca = reference.channels(), cb = result.channels();
    double sum = 0.0;
    for (int y = 0; y < reference.rows; y++) {
        const cv::uchar *a = reference.ptr<cv::uchar>(y), *b = result.ptr<cv::uchar>(y);
        for (int x = 0; x < reference.cols; x++) sum += std::abs((int)a[ca * x] - (int)b[cb * x]);
    }
    return sum / (double)reference.total();
}

int Comparator::ComputeMaxAbs(const cv::Mat &reference, const cv::Mat &result, Logger &logger)
{
    if (reference.empty() || result.empty() || reference.rows != result.rows || reference.cols != result.cols || reference.channels() != result.channels()) {
        logger.log("ComputeMaxAbs: images are empty or differ in shape", Logger::LogLevel::ERROR);
        return -1;
    }
    int m = 0;
    const size_t n = reference.total() * reference.channels();
    for (size_t i = 0; i < n; i++) m = std::max(m, std::abs((int)reference.data[i] - (int)result.data[i]));
    return m;
}

Comparator::Report Comparator::CompareGPUvsCPU(const std::vector<unsigned char> &gpu_output, const cv::Mat &cpu_result, double cpu_ms, Logger &logger)
{
    Report rep;
    rep.cpu_ms = cpu_ms;
    const size_t px = cpu_result.total();
    const int cn = cpu_result.channels();
    if (cpu_result.empty() || px == 0) {
        logger.log("CompareGPUvsCPU: empty CPU result", Logger::LogLevel::ERROR);
        return rep;
    }
    // the GPU buffer either has the CPU result's layout, or it is the (g,g,g,255) gray container of a 1-channel result
    const bool container = cn == 1 && gpu_output.size() == px * 4;
    if (!container && gpu_output.size() != px * cn) {
        logger.log("CompareGPUvsCPU: GPU output has " + std::to_string(gpu_output.size()) + " bytes, CPU result " + std::to_string(px * cn),
                   Logger::LogLevel::ERROR);
        rep.mismatches = (size_t)-1;
        return rep;
    }
    double sum = 0.0;
    rep.bytes = px * cn;
    for (size_t i = 0; i < rep.bytes; i++) {
        const int d = std::abs((int)cpu_result.data[i] - (int)gpu_output[container ? 4 * i : i]);
        sum += d;
        rep.max_abs = std::max(rep.max_abs, d);
        rep.mismatches += d != 0;
    }
    rep.mae = sum / (double)rep.bytes;
    (void)m_num_methods;
    return rep;
}
