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

    // [new] the same table with the columns a GPU-versus-CPU harness needs: the reference's 11 columns
    // first and unchanged, so its plotting scripts (src/*/results/visualisation.py) still read the file, then the method,
    // the largest absolute error, throughput (kernel operation time: upload + kernel + download), frames per second,
    // algorithmic GB/s of the kernel alone and its share of the HBM peak, and how many devices served the call.
    struct ExtendedRow {
        ResultRow base;
        std::string method;
        int max_abs_err = 0;
        double mpix_s = 0.0, fps = 0.0, gbps = 0.0, pct_hbm_peak = 0.0;
        int n_gpus = 1;
    };
    // `kernel_bytes_per_pixel`: algorithmic bytes the kernel moves per pixel; `hbm_peak_gbs`: the measured
    // copy bandwidth of the device (MEASURED_PEAKS.json), 0 = leave the percentage empty
    static ExtendedRow Extend(const ResultRow &base, const std::string &method, int width, int height, int max_abs_err,
                              double kernel_bytes_per_pixel, double hbm_peak_gbs, int n_gpus);
    void WriteExtendedResultsToCSV(const std::string &filename, const std::vector<ExtendedRow> &results);
    void SetSaveImages(bool save) { SAVE_IMAGES = save; }   // [new] the reference never initialises this member

private:
    bool SAVE_IMAGES;
    std::string m_directory_name;
    std::vector<std::string> m_image_paths;
};
