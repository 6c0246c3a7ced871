// Comparator.hpp -- drop-in for the reference's Comparator (include/Comparator.hpp:8-23;
// RT/src/Comparator.cpp), the CPU side of the project's validation harness, extended as the north star
// asks into a GPU-versus-CPU correctness and timing harness.
//
// The CPU functions here are the COMPARISON side only: nothing in Controller / ProgramHandler / librip_cuda
// calls them, and there is no CPU fallback for the device path.  They restate the reference's CPU code
// (the definition of correct output):
//   gray   RT/src/Comparator.cpp:30-45        uchar(0.299*r + 0.587*g + 0.114*b) in double, truncation
//   blur   src/GaussianBlur/GaussianBlur.cpp:231-261   float32 K*K sum in ky-major/kx-minor order, clamp borders, truncation
//   Sobel  src/EdgeDetection/EdgeDetection.cpp:219-240 3x3 correlation, BORDER_REFLECT_101, magnitude, round-half-even, saturate
//   MAE    RT/src/Comparator.cpp:60-101       mean |a - b| over channel 0
#pragma once

#include <string>
#include <vector>

#include "Logger.hpp"
#include "Mat.hpp"

class Comparator {
public:
    struct Report {            // [new] one GPU-versus-CPU comparison
        double mae = 0.0;      // mean absolute error over all bytes
        int max_abs = 0;       // largest absolute difference
        size_t mismatches = 0; // bytes that differ
        size_t bytes = 0;
        double cpu_ms = 0.0;   // average CPU time per iteration
        bool exact() const { return mismatches == 0; }
    };

    Comparator(int num_methods, int num_iterations);
    // reference method: BGR image from disk -> CV_8UC1 gray; avg_cpu_execution_time in ms over the iterations
    cv::Mat PerformCPU_Grayscaling(std::string image_path, double &avg_cpu_execution_time, Logger &logger);

    // [new] the same CPU paths on in-memory images.  `order_bgr`: channel order of 3/4-channel input.
    cv::Mat PerformCPU_Grayscaling(const cv::Mat &image, bool order_bgr, double &avg_cpu_execution_time, Logger &logger);
    cv::Mat PerformCPU_GaussianBlur(const cv::Mat &image, int kernel_size, float kernel_sigma, double &avg_cpu_execution_time, Logger &logger);
    cv::Mat PerformCPU_EdgeDetection(const cv::Mat &gray, double &avg_cpu_execution_time, Logger &logger);
    cv::Mat PerformCPU_Fused(const cv::Mat &image, bool order_bgr, int kernel_size, float kernel_sigma, double &avg_cpu_execution_time, Logger &logger);

    // public here (private in the reference, include/Comparator.hpp:21)
    double ComputeMAE(const cv::Mat &reference, const cv::Mat &result, Logger &logger);
    int ComputeMaxAbs(const cv::Mat &reference, const cv::Mat &result, Logger &logger);
    // [new] compares a GPU output buffer (as returned by Controller / ProgramHandler) with a CPU result
    Report CompareGPUvsCPU(const std::vector<unsigned char> &gpu_output, const cv::Mat &cpu_result, double cpu_ms, Logger &logger);

    // the reference's weight generator (src/GaussianBlur/src/Controller.cpp:342-362), typed exactly like it
    static std::vector<float> GaussianKernel(int kernel_size, float sigma);

private:
    int m_num_methods;
    int NUMBER_OF_ITERATIONS;
};
