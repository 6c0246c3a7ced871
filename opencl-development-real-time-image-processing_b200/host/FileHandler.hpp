// FileHandler.hpp -- drop-in for the reference's FileHandler (include/FileHandler.hpp:12-25;
// RT/src/FileHandler.cpp): directory scan, image save, results CSV with the reference's 11 columns.
#pragma once

#include <string>
#include <tuple>
#include <vector>

#include "Mat.hpp"

class FileHandler {
public:
    typedef std::tuple<std::string, std::string, std::string, int, double, double, double, double, double, double, double> ResultRow;

    FileHandler();
    // non-recursive scan for .jpg / .png (as the reference) plus .ppm / .pgm (what the built-in loader reads
    // when OpenCV is absent); like the reference the paths ACCUMULATE across calls (FileHandler.cpp:5-14)
    std::vector<std::string> LoadImages(const std::string &directory);
    void SaveImages(std::string image_path, cv::Mat &opencl_output_image);
    void WriteResultsToCSV(const std::string &filename, std::vector<ResultRow> &results);
    void SetSaveImages(bool save) { SAVE_IMAGES = save; }   // [new] the reference never initialises this member

private:
    bool SAVE_IMAGES;
    std::string m_directory_name;
    std::vector<std::string> m_image_paths;
};
