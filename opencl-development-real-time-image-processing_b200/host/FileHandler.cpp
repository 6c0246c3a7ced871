// FileHandler.cpp -- see FileHandler.hpp.
#include "FileHandler.hpp"

#include <algorithm>
#include <filesystem>
#include <fstream>

namespace fs = std::filesystem;

FileHandler::FileHandler() : SAVE_IMAGES(false), m_directory_name("output") {}

std::vector<std::string> FileHandler::LoadImages(const std::string &directory)
{
    std::vector<std::string> found;
    std::error_code ec;
    for (const auto &entry : fs::directory_iterator(directory, ec)) {
        if (!entry.is_regular_file()) continue;
        std::string ext = entry.path().extension().string();
        std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        if (ext == ".jpg" || ext == ".png" || ext == ".ppm" || ext == ".pgm") found.push_back(entry.path().string());
    }
    std::sort(found.begin(), found.end());
    m_image_paths.insert(m_image_paths.end(), found.begin(), found.end());
    return m_image_paths;
}

void FileHandler::SaveImages(std::string image_path, cv::Mat &opencl_output_image)
{
    if (!SAVE_IMAGES || opencl_output_image.empty()) return;
    std::error_code ec;
    fs::create_directories(m_directory_name, ec);
    const fs::path stem = fs::path(image_path).stem();
#ifdef RIP_HAVE_OPENCV
    const std::string out = (fs::path(m_directory_name) / (stem.string() + ".jpg")).string();
#else
    const std::string out = (fs::path(m_directory_name) / (stem.string() + (opencl_output_image.channels() == 1 ? ".pgm" : ".ppm"))).string();
#endif
    cv::imwrite(out, opencl_output_image);
}

void FileHandler::WriteResultsToCSV(const std::string &filename, std::vector<ResultRow> &results)
{
    std::ofstream f(filename);
    if (!f) return;
    // header of RT/src/FileHandler.cpp:28
    f << "Timestamp, Image, Resolution, Num_Iterations, avg_CPU_Time_ms, avg_OpenCL_Time_ms, avg_OpenCL_kernel_ms, "
         "avg_OpenCL_kernel_write_ms, avg_OpenCL_kernel_read_ms, avg_OpenCL_kernel_operation_ms, Error_MAE\n";
    for (const auto &r : results)
        f << std::get<0>(r) << ", " << std::get<1>(r) << ", " << std::get<2>(r) << ", " << std::get<3>(r) << ", " << std::get<4>(r) << ", "
          << std::get<5>(r) << ", " << std::get<6>(r) << ", " << std::get<7>(r) << ", " << std::get<8>(r) << ", " << std::get<9>(r) << ", "
          << std::get<10>(r) << "\n";
}

FileHandler::ExtendedRow FileHandler::Extend(const ResultRow &base, const std::string &method, int width, int height, int max_abs_err,
                                             double kernel_bytes_per_pixel, double hbm_peak_gbs, int n_gpus)
{
    ExtendedRow r;
    r.base = base;
    r.method = method;
    r.max_abs_err = max_abs_err;
    r.n_gpus = n_gpus;
    const double px = (double)width * height;
    const double op_ms = std::get<9>(base), kernel_ms = std::get<6>(base), e2e_ms = std::get<5>(base);
    if (op_ms > 0.0) r.mpix_s = px / (op_ms * 1e3);                 // pixels per microsecond = Mpixel/s
    if (e2e_ms > 0.0) r.fps = 1e3 / e2e_ms;                         // what a caller of PerformOpenCL sees
    if (kernel_ms > 0.0) r.gbps = px * kernel_bytes_per_pixel / (kernel_ms * 1e6);
    if (hbm_peak_gbs > 0.0) r.pct_hbm_peak = 100.0 * r.gbps / hbm_peak_gbs;
    return r;
}

void FileHandler::WriteExtendedResultsToCSV(const std::string &filename, const std::vector<ExtendedRow> &results)
{
    std::ofstream f(filename);
    if (!f) return;
    f << "Timestamp, Image, Resolution, Num_Iterations, avg_CPU_Time_ms, avg_OpenCL_Time_ms, avg_OpenCL_kernel_ms, "
         "avg_OpenCL_kernel_write_ms, avg_OpenCL_kernel_read_ms, avg_OpenCL_kernel_operation_ms, Error_MAE, "
         "Method, max_abs_err, Mpix_s, fps, GBps, pct_hbm_peak, n_gpus\n";
    for (const auto &e : results) {
        const ResultRow &r = e.base;
        f << std::get<0>(r) << ", " << std::get<1>(r) << ", " << std::get<2>(r) << ", " << std::get<3>(r) << ", " << std::get<4>(r) << ", "
          << std::get<5>(r) << ", " << std::get<6>(r) << ", " << std::get<7>(r) << ", " << std::get<8>(r) << ", " << std::get<9>(r) << ", "
          << std::get<10>(r) << ", " << e.method << ", " << e.max_abs_err << ", " << e.mpix_s << ", " << e.fps << ", " << e.gbps << ", "
          << e.pct_hbm_peak << ", " << e.n_gpus << "\n";
    }
}
