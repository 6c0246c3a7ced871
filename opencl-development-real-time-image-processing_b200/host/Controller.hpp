// Controller.hpp -- drop-in for the reference's Controller (include/Controller.hpp:16-68): same public
// method names, parameter order and container types, implemented on the CUDA C ABI (include/rip_cuda.h)
// instead of the OpenCL C API.  The OpenCL handle types in the signatures are the aliases of rip_compat.h.
//
//   reference (RT/ = src/RealtimeImageProcessing/)                  here
//   clGetPlatformIDs / clGetDeviceIDs   RT/src/Controller.cpp:13-64   rip_device_count / rip_device_name
//   clCreateContext                     :97-113                       rip_ctx_create over the device set
//   clCreateCommandQueue (profiling on) :115-129                      one device of the context + its streams
//   clCreateProgramWithSource + build   :131-179                      rip_module_load (kernels are compiled in)
//   clCreateKernel                      :181-191                      rip_kernel_get
//   PerformCL* = create buffers, write, NDRange, read, float->uchar   rip_process_host: cached pinned/device
//                (:429-744)                                           buffers, async copies, one kernel, u8 out
//
// Results are those of the reference's CPU paths (bit-exact), not of its OpenCL kernels.
// Additions beyond the reference are marked [new].
#pragma once

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "InfoPlatform.hpp"
#include "Logger.hpp"
#include "rip_compat.h"

class Controller {
public:
    Controller();
    // prints "Error: <name> (<code>)" and exits on err != CL_SUCCESS (RT/src/Controller.cpp:5-11)
    void CheckError(cl_int err, const char *name);

    std::vector<cl_platform_id> GetPlatforms();
    std::vector<cl_device_id> GetDevices(cl_platform_id platform);
    cl_bool GetImageSupport();                       // always CL_FALSE: the image2d kernel flavours are not reproduced
    void SetImageSupport(cl_bool image_support);     // kept for source compatibility; ignored

    cl_context CreateContext(cl_platform_id platform, std::vector<cl_device_id> devices);
    cl_command_queue CreateCommandQueue(cl_context context, cl_device_id device);
    cl_program CreateProgram(cl_context context, cl_device_id device, const char *filename);
    cl_kernel CreateKernel(cl_program program, const char *kernel_name);
    void DisplayPlatformInformation(cl_platform_id platform);
    void Cleanup(cl_context context = 0, cl_command_queue commandQueue = 0, cl_program program = 0, cl_kernel kernel = 0,
                 cl_sampler sampler = 0, cl_mem *mem_objects = 0, int num_mem_objects = 0);

    // input_data: RGBA, W*H*4 bytes.  output_data: (g,g,g,255) W*H*4 bytes (RT/src/ProgramHandler.cpp:185)
    void PerformCLImageGrayscaling(cl_context *context, cl_command_queue *command_queue, cl_kernel *kernel,
                                   std::vector<cl_ulong> *profiling_events, std::vector<unsigned char> *input_data,
                                   std::vector<unsigned char> *output_data, cl_int &width, cl_int &height, Logger &logger);
    // output_data: W*H bytes, Sobel magnitude of the gray image (RT/src/ProgramHandler.cpp:192)
    void PerformCLImageEdgeDetection(cl_context *context, cl_command_queue *command_queue, cl_kernel *kernel,
                                     std::vector<cl_ulong> *profiling_events, std::vector<unsigned char> *input_data,
                                     std::vector<unsigned char> *output_data, cl_int &width, cl_int &height, Logger &logger);
    // output_data: RGBA W*H*4 bytes, all four channels blurred (RT/src/ProgramHandler.cpp:199)
    void PerformCLGaussianBlur(int &kernel_size, float &kernel_sigma, cl_context *context, cl_command_queue *command_queue,
                               cl_kernel *kernel, std::vector<cl_ulong> *profiling_events, std::vector<unsigned char> *input_data,
                               std::vector<unsigned char> *output_data, cl_int &width, cl_int &height, Logger &logger);

    // [new] gray -> KxK Gaussian -> Sobel in one kernel; output_data: W*H bytes.  in_format: RIP_FMT_*.
    void PerformFused(int &kernel_size, float &kernel_sigma, cl_context *context, cl_command_queue *command_queue, cl_kernel *kernel,
                      std::vector<cl_ulong> *profiling_events, std::vector<unsigned char> *input_data,
                      std::vector<unsigned char> *output_data, cl_int &width, cl_int &height, Logger &logger,
                      int in_format = RIP_FMT_RGBA8);
    // [new] a batch of n_frames consecutive frames, sharded over all devices of the CONTEXT (contiguous
    // blocks of frames per device, no inter-device traffic).  method: "GRAYSCALE" | "EDGE" | "GAUSSIAN" | "FUSED".
    void PerformBatch(const std::string &method, cl_context *context, std::vector<cl_ulong> *profiling_events,
                      const unsigned char *frames, int n_frames, std::vector<unsigned char> *output_data, cl_int &width,
                      cl_int &height, Logger &logger, int in_format = RIP_FMT_RGBA8, int kernel_size = 5, float kernel_sigma = 1.0f,
                      size_t frames_bytes = 0 /* bytes behind `frames` (0 = not checked) */);
    // [new] one large frame split into row bands with halo, one band per device of the context (EDGE, FUSED)
    void PerformBanded(const std::string &method, cl_context *context, std::vector<cl_ulong> *profiling_events,
                       std::vector<unsigned char> *input_data, std::vector<unsigned char> *output_data, cl_int &width,
                       cl_int &height, Logger &logger, int in_format = RIP_FMT_RGBA8, int kernel_size = 5, float kernel_sigma = 1.0f);
    // [new] asynchronous per-frame form of the three operations (+ "FUSED"): SubmitFrame queues one frame and returns a handle
    // (0 = rejected, logged) at once; CollectFrame blocks until that frame is in output_data.  Frames complete in submission
    // order; with 2-3 frames in flight upload, kernel and download of consecutive frames overlap.
    int SubmitFrame(const std::string &method, cl_command_queue *command_queue, std::vector<unsigned char> *input_data,
                    std::vector<unsigned char> *output_data, cl_int &width, cl_int &height, Logger &logger, int in_format = RIP_FMT_RGBA8,
                    int kernel_size = 5, float kernel_sigma = 1.0f);
    bool CollectFrame(int handle, std::vector<cl_ulong> *profiling_events, Logger &logger);
    // [new] page-lock / release a container that is reused across calls (copied by DMA without a staging copy)
    bool PinHostBuffer(std::vector<unsigned char> *buffer);
    void UnpinHostBuffer(std::vector<unsigned char> *buffer);
    // [new] restrict GetDevices() to these CUDA ordinals (default: all visible devices)
    void SetDevices(const std::vector<int> &ordinals);

    // the reference generator, typed exactly like it (RT/src/Controller.cpp:352-372); public as in
    // src/GaussianBlur/include/Controller.hpp:28 so that the CPU comparison can share it
    std::vector<float> _GenerateGausianKernel(int kernel_size, float sigma);

private:
    cl_uint num_platforms, num_devices;
    cl_bool m_image_support;
    std::vector<int> m_ordinals;
    struct Pending {
        rip_ticket *ticket;
        std::vector<unsigned char> *output;
    };
    std::map<int, Pending> m_pending;
    int m_next_handle = 0;
    std::vector<void *> m_pinned;

    void _appendProfile(const uint64_t prof_ns[6], std::vector<cl_ulong> *profiling_events);
    void _run(const rip_op_desc &desc, rip_ctx *ctx, std::vector<cl_ulong> *profiling_events, const unsigned char *in, size_t in_bytes,
              std::vector<unsigned char> *output_data, int width, int height, int n_frames, bool banded, Logger &logger, const char *what);
};
