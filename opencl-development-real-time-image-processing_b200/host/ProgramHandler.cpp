// ProgramHandler.cpp -- see ProgramHandler.hpp.
#include "ProgramHandler.hpp"

#include <chrono>
#include <cstdlib>
#include <iostream>

ProgramHandler::ProgramHandler(int number_of_iterations, bool log_events, bool display_images, bool display_terminal_results,
                               bool bypass_image_support, int gaussian_kernel_size, float gaussian_sigma)
    : LOG_EVENTS(log_events), DISPLAY_IMAGES(display_images), DISPLAY_TERMINAL_RESULTS(display_terminal_results),
      BYPASS_IMAGE_SUPPORT(bypass_image_support), NUMBER_OF_ITERATIONS(number_of_iterations), PLATFORM_INDEX(0), DEVICE_INDEX(0),
      GAUSSIAN_KERNEL_SIZE(gaussian_kernel_size), GAUSSIAN_SIGMA(gaussian_sigma)
{
}

void ProgramHandler::InitLogger(Logger &logger, Logger::LogLevel level, bool save_to_file)
{
    logger.setLogLevel(level);
    logger.setTerminalDisplay(DISPLAY_TERMINAL_RESULTS);
    logger.setLogFile("RealtimeImageProcessing.log", save_to_file);
}

void ProgramHandler::AddKernels(std::vector<std::string> kernels, std::string kernel_index)
{
    KERNELS[kernel_index] = kernels;
    METHOD.push_back(kernel_index);
}

void ProgramHandler::SetDeviceProperties(int platform_index, int device_index)
{
    PLATFORM_INDEX = platform_index;
    DEVICE_INDEX = device_index;
}

void ProgramHandler::InitOpenCL(Controller &controller, cl_context *context, cl_command_queue *command_queue, cl_program *program,
                                cl_kernel *kernel, std::string method, Logger &logger)
{
    std::vector<cl_platform_id> platforms = controller.GetPlatforms();
    if (PLATFORM_INDEX < 0 || PLATFORM_INDEX >= (int)platforms.size()) controller.CheckError(CL_INVALID_VALUE, "platform index out of range");
    std::vector<cl_device_id> devices = controller.GetDevices(platforms[PLATFORM_INDEX]);
    if (DEVICE_INDEX < 0 || DEVICE_INDEX >= (int)devices.size()) controller.CheckError(CL_INVALID_VALUE, "device index out of range");
    if (DISPLAY_TERMINAL_RESULTS) controller.DisplayPlatformInformation(platforms[PLATFORM_INDEX]);

    // method -> kernel entry point (RT/src/ProgramHandler.cpp:69-78)
    std::string kernel_name;
    if (method == "GRAYSCALE") kernel_name = "grayscale";
    else if (method == "EDGE") kernel_name = "sobel_edge_detection";
    else if (method == "GAUSSIAN") kernel_name = "gaussian_blur";
    else if (method == "FUSED") kernel_name = "fused";
    else controller.CheckError(CL_INVALID_VALUE, ("unknown method " + method).c_str());

    // image2d support is never reported, so index 1 (the *_base flavour) is used whenever the caller registered
    // two files, exactly what BYPASS_IMAGE_SUPPORT = true selects in the reference (:81-103)
    std::string file = kernel_name;
    auto it = KERNELS.find(method);
    if (it != KERNELS.end() && !it->second.empty()) file = it->second.size() > 1 ? it->second[1] : it->second[0];
    controller.SetImageSupport(CL_FALSE);
    (void)BYPASS_IMAGE_SUPPORT;

    *context = controller.CreateContext(platforms[PLATFORM_INDEX], devices);
    *command_queue = controller.CreateCommandQueue(*context, devices[DEVICE_INDEX]);
    *program = controller.CreateProgram(*context, devices[DEVICE_INDEX], file.c_str());
    if (!*command_queue || !*program) controller.CheckError(CL_INVALID_VALUE, "CreateCommandQueue / CreateProgram");
    *kernel = controller.CreateKernel(*program, kernel_name.c_str());
    logger.log("Initialised CUDA path for method " + method + " (kernel " + kernel_name + ", image " + file + ")", Logger::LogLevel::INFO);
}

void ProgramHandler::GetImageOpenCL(std::string image_path, std::vector<unsigned char> *input_data, cl_int *width, cl_int *height,
                                    Logger &logger)
{
    cv::Mat image = cv::imread(image_path, cv::IMREAD_COLOR);
    if (image.empty()) {
        logger.log("Failed to load image " + image_path, Logger::LogLevel::ERROR);
        input_data->clear();
        *width = *height = 0;
        return;
    }
    cv::Mat rgba;
    cv::cvtColor(image, rgba, cv::COLOR_BGR2RGBA);   // every device op consumes RGBA (RT/src/ProgramHandler.cpp:127)
    GetMatrix(rgba, input_data, width, height, logger);
}

void ProgramHandler::GetMatrix(const cv::Mat &input_frame, std::vector<unsigned char> *input_data, cl_int *width, cl_int *height,
                               Logger &logger)
{
    if (input_frame.empty() || input_frame.channels() != 4) {
        logger.log("Input frame must be a non-empty 4-channel (RGBA) matrix", Logger::LogLevel::ERROR);
        input_data->clear();
        *width = *height = 0;
        return;
    }
    *width = input_frame.cols;
    *height = input_frame.rows;
    input_data->assign(input_frame.data, input_frame.data + input_frame.total() * 4);
}

bool ProgramHandler::Dispatch(Controller &controller, const std::string &method, cl_context *context, cl_command_queue *command_queue,
                              cl_kernel *kernel, std::vector<cl_ulong> *events, std::vector<unsigned char> *in,
                              std::vector<unsigned char> *out, cl_int &width, cl_int &height, Logger &logger)
{
    out->clear();
    if (method == "GRAYSCALE") controller.PerformCLImageGrayscaling(context, command_queue, kernel, events, in, out, width, height, logger);
    else if (method == "EDGE") controller.PerformCLImageEdgeDetection(context, command_queue, kernel, events, in, out, width, height, logger);
    else if (method == "GAUSSIAN")
        controller.PerformCLGaussianBlur(GAUSSIAN_KERNEL_SIZE, GAUSSIAN_SIGMA, context, command_queue, kernel, events, in, out, width, height, logger);
    else if (method == "FUSED")
        controller.PerformFused(GAUSSIAN_KERNEL_SIZE, GAUSSIAN_SIGMA, context, command_queue, kernel, events, in, out, width, height, logger);
    else {
        logger.log("Unknown method " + method, Logger::LogLevel::ERROR);
        return false;
    }
    return !out->empty();
}

std::vector<unsigned char> ProgramHandler::PerformOpenCL(Controller &controller, std::string image_path, cl_context *context,
                                                         cl_command_queue *command_queue, cl_kernel *kernel, double &avg_opencl_execution_time,
                                                         double &avg_opencl_kernel_write_time, double &avg_opencl_kernel_execution_time,
                                                         double &avg_opencl_kernel_read_time, double &avg_opencl_kernel_operation, cl_int &width,
                                                         cl_int &height, Logger &logger, std::string method)
{
    std::vector<unsigned char> input_data, output;
    GetImageOpenCL(image_path, &input_data, &width, &height, logger);
    avg_opencl_execution_time = avg_opencl_kernel_write_time = avg_opencl_kernel_execution_time = avg_opencl_kernel_read_time =
        avg_opencl_kernel_operation = 0.0;
    if (input_data.empty()) return output;

    double total_ms = 0.0, write_ms = 0.0, kernel_ms = 0.0, read_ms = 0.0;
    int done = 0;
    for (int i = 0; i < NUMBER_OF_ITERATIONS; i++) {
        std::vector<cl_ulong> events;   // per iteration: the reference never clears its vector and so averages iteration 0
        const auto t0 = std::chrono::high_resolution_clock::now();
        const bool ok = Dispatch(controller, method, context, command_queue, kernel, &events, &input_data, &output, width, height, logger);
        const auto t1 = std::chrono::high_resolution_clock::now();
        if (!ok || events.size() < 6) continue;
        total_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        write_ms += (events[1] - events[0]) * 1e-6;
        kernel_ms += (events[3] - events[2]) * 1e-6;
        read_ms += (events[5] - events[4]) * 1e-6;
        done++;
        if (LOG_EVENTS) logger.log("Iteration " + std::to_string(i) + " complete", Logger::LogLevel::INFO);
    }
    if (done) {
        avg_opencl_execution_time = total_ms / done;
        avg_opencl_kernel_write_time = write_ms / done;
        avg_opencl_kernel_execution_time = kernel_ms / done;
        avg_opencl_kernel_read_time = read_ms / done;
        avg_opencl_kernel_operation = avg_opencl_kernel_write_time + avg_opencl_kernel_execution_time + avg_opencl_kernel_read_time;
    }
    (void)DISPLAY_IMAGES;   // no display on a headless GPU box
    return output;
}

std::vector<unsigned char> ProgramHandler::PerformOpenCL(Controller &controller, const cv::Mat &input_frame, cl_context *context,
                                                         cl_command_queue *command_queue, cl_kernel *kernel, cl_int &width, cl_int &height,
                                                         Logger &logger, std::string method)
{
    std::vector<unsigned char> input_data, output;
    cl_int w = 0, h = 0;
    GetMatrix(input_frame, &input_data, &w, &h, logger);
    if (input_data.empty()) return output;
    if (width != 0 && height != 0 && (w != width || h != height)) {   // the reference asserts the frame matches the announced size (:268-272)
        logger.log("Frame is " + std::to_string(w) + "x" + std::to_string(h) + " but " + std::to_string(width) + "x" + std::to_string(height) +
                       " was announced", Logger::LogLevel::ERROR);
        return output;
    }
    width = w;
    height = h;
    std::vector<cl_ulong> events;
    Dispatch(controller, method, context, command_queue, kernel, &events, &input_data, &output, width, height, logger);
    return output;
}

int ProgramHandler::SubmitOpenCL(Controller &controller, const cv::Mat &input_frame, cl_command_queue *command_queue, cl_int &width,
                                 cl_int &height, Logger &logger, std::string method)
{
    std::unique_ptr<FrameSlot> slot;
    if (!m_free_slots.empty()) {
        slot = std::move(m_free_slots.back());
        m_free_slots.pop_back();
    } else {
        slot.reset(new FrameSlot());
    }
    cl_int w = 0, h = 0;
    GetMatrix(input_frame, &slot->in, &w, &h, logger);
    if (slot->in.empty()) return 0;
    if (width != 0 && height != 0 && (w != width || h != height)) {
        logger.log("Frame is " + std::to_string(w) + "x" + std::to_string(h) + " but " + std::to_string(width) + "x" + std::to_string(height) +
                       " was announced", Logger::LogLevel::ERROR);
        return 0;
    }
    width = w;
    height = h;
    const int handle = controller.SubmitFrame(method, command_queue, &slot->in, &slot->out, width, height, logger, RIP_FMT_RGBA8,
                                              GAUSSIAN_KERNEL_SIZE, GAUSSIAN_SIGMA);
    if (handle == 0) return 0;
    m_frames[handle] = std::move(slot);
    return handle;
}

std::vector<unsigned char> ProgramHandler::CollectOpenCL(Controller &controller, int handle, Logger &logger, std::vector<cl_ulong> *events)
{
    std::vector<unsigned char> output;
    auto it = m_frames.find(handle);
    if (it == m_frames.end()) {
        logger.log("CollectOpenCL: unknown handle " + std::to_string(handle), Logger::LogLevel::ERROR);
        return output;
    }
    std::unique_ptr<FrameSlot> slot = std::move(it->second);
    m_frames.erase(it);
    if (controller.CollectFrame(handle, events, logger)) output = slot->out;   // (the slot keeps its capacity for the next frame)
    if (m_free_slots.size() < 8) m_free_slots.push_back(std::move(slot));
    return output;
}
