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
