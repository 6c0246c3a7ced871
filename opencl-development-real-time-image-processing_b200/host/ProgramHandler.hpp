// ProgramHandler.hpp -- drop-in for the reference's orchestration class (include/ProgramHandler.hpp:6-45;
// RT/src/ProgramHandler.cpp): method string -> kernel, image load + BGR->RGBA, iteration loop, timing
// averages.  Same constructor and public methods; the method strings are "GRAYSCALE", "EDGE", "GAUSSIAN"
// and [new] "FUSED".
#pragma once

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "Controller.hpp"
#include "Logger.hpp"
#include "Mat.hpp"

class ProgramHandler {
public:
    ProgramHandler(int number_of_iterations, bool log_events, bool display_images, bool display_terminal_results, bool bypass_image_support,
                   int gaussian_kernel_size = 17, float gaussian_sigma = 6.0f);

    void InitLogger(Logger &logger, Logger::LogLevel level, bool save_to_file);
    void InitOpenCL(Controller &controller, cl_context *context, cl_command_queue *command_queue, cl_program *program, cl_kernel *kernel,
                    std::string method, Logger &logger);
    void AddKernels(std::vector<std::string> kernels, std::string kernel_index);
    void SetDeviceProperties(int platform_index, int device_index);

    // per-image benchmark (RT/src/ProgramHandler.cpp:144-257): N iterations, returns the last output
    std::vector<unsigned char> PerformOpenCL(Controller &controller, std::string image_path, cl_context *context, cl_command_queue *command_queue,
                                             cl_kernel *kernel, double &avg_opencl_execution_time, double &avg_opencl_kernel_write_time,
                                             double &avg_opencl_kernel_execution_time, double &avg_opencl_kernel_read_time,
                                             double &avg_opencl_kernel_operation, cl_int &width, cl_int &height, Logger &logger,
                                             std::string method);
    // per-frame variant (RT/src/ProgramHandler.cpp:259-329): input_frame is RGBA (4 channels)
    std::vector<unsigned char> PerformOpenCL(Controller &controller, const cv::Mat &input_frame, cl_context *context,
                                             cl_command_queue *command_queue, cl_kernel *kernel, cl_int &width, cl_int &height, Logger &logger,
                                             std::string method);

    // [new] streaming form of the per-frame variant: SubmitOpenCL converts the frame, queues it and returns a handle (0 = rejected)
    // at once; CollectOpenCL blocks until that frame is done and returns its output (and, if `events` is given, the six
    // profiling values).  The frame loop of RealtimeImageProcessing.cpp:325-418 with two or three frames in flight overlaps
    // the upload of frame i+1, the kernel of frame i and the download of frame i-1.  Frames complete in submission order.
    int SubmitOpenCL(Controller &controller, const cv::Mat &input_frame, cl_command_queue *command_queue, cl_int &width, cl_int &height,
                     Logger &logger, std::string method);
    std::vector<unsigned char> CollectOpenCL(Controller &controller, int handle, Logger &logger, std::vector<cl_ulong> *events = nullptr);

private:
    struct FrameSlot {   // (page-locking these containers with cudaHostRegister was measured in round 2: DMA out of / into registered
        std::vector<unsigned char> in, out;   // vector storage is slower on this host than the pipeline's own staging copy into cudaMallocHost memory,
    };                                        // which runs on the device's worker thread next to the caller: 1080p FUSED 1.25 ms vs 0.99 ms per frame)
    std::map<int, std::unique_ptr<FrameSlot>> m_frames;   // frames in flight, by Controller handle
    std::vector<std::unique_ptr<FrameSlot>> m_free_slots; // recycled containers (no allocation per frame in steady state)
    bool LOG_EVENTS;
    bool DISPLAY_IMAGES;
    bool DISPLAY_TERMINAL_RESULTS;
    bool BYPASS_IMAGE_SUPPORT;
    int NUMBER_OF_ITERATIONS;
    int PLATFORM_INDEX;
    int DEVICE_INDEX;
    int GAUSSIAN_KERNEL_SIZE;
    float GAUSSIAN_SIGMA;
    std::vector<std::string> METHOD;
    std::map<std::string, std::vector<std::string>> KERNELS;

    void GetImageOpenCL(std::string image_path, std::vector<unsigned char> *input_data, cl_int *width, cl_int *height, Logger &logger);
    void GetMatrix(const cv::Mat &input_frame, std::vector<unsigned char> *input_data, cl_int *width, cl_int *height, Logger &logger);
    bool Dispatch(Controller &controller, const std::string &method, cl_context *context, cl_command_queue *command_queue, cl_kernel *kernel,
                  std::vector<cl_ulong> *events, std::vector<unsigned char> *in, std::vector<unsigned char> *out, cl_int &width, cl_int &height,
                  Logger &logger);
};
