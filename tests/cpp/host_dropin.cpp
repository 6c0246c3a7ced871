// host_dropin.cpp -- drives the host classes the way the reference's applications do
// (src/RealtimeImageProcessing/RealtimeImageProcessing.cpp:32-138 PerformOnImages; the stand-alone apps'
// main() loops, e.g. src/EdgeDetection/EdgeDetection.cpp:329-431): InitOpenCL per method, PerformOpenCL on
// each image of a directory, the CPU path of the Comparator, MAE, results CSV.
//
//   host_dropin <workdir> [--gpu]
//
// <workdir>/images/*.ppm are the inputs.  Always: the Comparator's CPU results are written to
// <workdir>/cpu_<method>_<image>.raw (the test compares them with the oracle).  With --gpu: every method
// also runs on the device through ProgramHandler / Controller and must match the CPU result bit for bit;
// the GPU outputs go to <workdir>/gpu_<method>_<image>.raw and the table to <workdir>/results.csv.
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "Comparator.hpp"
#include "Controller.hpp"
#include "FileHandler.hpp"
#include "Logger.hpp"
#include "ProgramHandler.hpp"

namespace fs = std::filesystem;

static void write_raw(const std::string &path, const unsigned char *p, size_t n)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(p), (std::streamsize)n);
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cerr << "usage: host_dropin <workdir> [--gpu]" << std::endl;
        return 2;
    }
    const std::string work = argv[1];
    const bool gpu = argc > 2 && std::string(argv[2]) == "--gpu";
    const int iterations = 2, ksize = 5;
    const float sigma = 1.5f;   // the reference's stand-alone default (GaussianBlur.cpp:15-16)

    Logger &logger = Logger::getInstance();
    ProgramHandler handler(iterations, false, false, false, true, ksize, sigma);
    handler.InitLogger(logger, Logger::LogLevel::ERROR, false);
    handler.SetDeviceProperties(0, 0);
    handler.AddKernels({"grayscale_images.cl", "grayscale_base.cl"}, "GRAYSCALE");
    handler.AddKernels({"edge_images.cl", "edge_base.cl"}, "EDGE");
    handler.AddKernels({"gaussian_images.cl", "gaussian_base.cl"}, "GAUSSIAN");
    handler.AddKernels({"fused"}, "FUSED");

    FileHandler files;
    Comparator comparator(4, iterations);
    Controller controller;
    std::vector<std::string> images = files.LoadImages(work + "/images");
    if (images.empty()) {
        std::cerr << "no images under " << work << "/images" << std::endl;
        return 2;
    }
    std::vector<FileHandler::ResultRow> results;
    int failures = 0;
    const char *methods[] = {"GRAYSCALE", "EDGE", "GAUSSIAN", "FUSED"};

    for (const char *method : methods) {
        cl_context context = 0;
        cl_command_queue queue = 0;
        cl_program program = 0;
        cl_kernel kernel = 0;
        if (gpu) handler.InitOpenCL(controller, &context, &queue, &program, &kernel, method, logger);
        for (const std::string &path : images) {
            const std::string tag = std::string(method) + "_" + fs::path(path).stem().string();
            // ---- CPU side (Comparator) ----
            cv::Mat bgr = cv::imread(path, cv::IMREAD_COLOR), rgba;
            cv::cvtColor(bgr, rgba, cv::COLOR_BGR2RGBA);
            double cpu_ms = 0.0, t = 0.0;
            cv::Mat cpu;
            if (!strcmp(method, "GRAYSCALE")) cpu = comparator.PerformCPU_Grayscaling(path, cpu_ms, logger);
            else if (!strcmp(method, "GAUSSIAN")) cpu = comparator.PerformCPU_GaussianBlur(rgba, ksize, sigma, cpu_ms, logger);
            else if (!strcmp(method, "EDGE")) {
                cv::Mat gray = comparator.PerformCPU_Grayscaling(rgba, false, t, logger);
                cpu = comparator.PerformCPU_EdgeDetection(gray, cpu_ms, logger);
            } else cpu = comparator.PerformCPU_Fused(rgba, false, ksize, sigma, cpu_ms, logger);
            if (cpu.empty()) { failures++; continue; }
            write_raw(work + "/cpu_" + tag + ".raw", cpu.data, cpu.total() * cpu.channels());
            if (!gpu) continue;
            // ---- device side (ProgramHandler -> Controller -> librip_cuda) ----
            double e2e = 0, wr = 0, ke = 0, rd = 0, op = 0;
            cl_int w = 0, h = 0;
            std::vector<unsigned char> out = handler.PerformOpenCL(controller, path, &context, &queue, &kernel, e2e, wr, ke, rd, op, w, h, logger, method);
            if (out.empty()) { std::cerr << tag << ": no GPU output" << std::endl; failures++; continue; }
            write_raw(work + "/gpu_" + tag + ".raw", out.data(), out.size());
            const Comparator::Report rep = comparator.CompareGPUvsCPU(out, cpu, cpu_ms, logger);
            // per-frame overload on the same frame must agree with the per-image one
            cl_int w2 = w, h2 = h;
            std::vector<unsigned char> out2 = handler.PerformOpenCL(controller, rgba, &context, &queue, &kernel, w2, h2, logger, method);
            const bool same = out2 == out;
            std::printf("%-10s %-12s %4dx%-4d cpu %8.3f ms  gpu e2e %7.3f ms (write %.3f kernel %.3f read %.3f)  MAE %g max %d mismatches %zu%s\n",
                        method, fs::path(path).stem().string().c_str(), w, h, cpu_ms, e2e, wr, ke, rd, rep.mae, rep.max_abs, rep.mismatches,
                        same ? "" : "  [per-frame overload differs]");
            if (!rep.exact() || !same) failures++;
            results.emplace_back(logger.getCurrentTime(), fs::path(path).filename().string(), std::to_string(w) + "x" + std::to_string(h),
                                 iterations, cpu_ms, e2e, ke, wr, rd, op, rep.mae);
        }
        if (gpu) controller.Cleanup(context, queue, program, kernel);
    }

    if (gpu) {
        // [new] batch API: 3 copies of the first image through every device of a context
        cv::Mat bgr = cv::imread(images[0], cv::IMREAD_COLOR), rgba;
        cv::cvtColor(bgr, rgba, cv::COLOR_BGR2RGBA);
        std::vector<unsigned char> batch;
        for (int i = 0; i < 3; i++) batch.insert(batch.end(), rgba.data, rgba.data + rgba.total() * 4);
        std::vector<cl_platform_id> platforms = controller.GetPlatforms();
        std::vector<cl_device_id> devices = controller.GetDevices(platforms[0]);
        cl_context context = controller.CreateContext(platforms[0], devices);
        std::vector<cl_ulong> events;
        std::vector<unsigned char> out;
        cl_int w = rgba.cols, h = rgba.rows;
        controller.PerformBatch("FUSED", &context, &events, batch.data(), 3, &out, w, h, logger, RIP_FMT_RGBA8, ksize, sigma);
        double t = 0;
        cv::Mat cpu = comparator.PerformCPU_Fused(rgba, false, ksize, sigma, t, logger);
        bool ok = out.size() == cpu.total() * 3 && events.size() == 6;
        for (int i = 0; ok && i < 3; i++) ok = !memcmp(out.data() + (size_t)i * cpu.total(), cpu.data, cpu.total());
        std::printf("PerformBatch FUSED x3 on %zu device(s): %s\n", devices.size(), ok ? "exact" : "MISMATCH");
        if (!ok) failures++;
        // [new] NV12 (the reference's camera format, RealtimeImageProcessing.cpp:153): the luma plane is the gray image
        {
            double t2 = 0;
            cv::Mat gray = comparator.PerformCPU_Grayscaling(rgba, false, t2, logger);
            const int hh = gray.rows & ~1, ww = gray.cols;
            std::vector<unsigned char> nv12((size_t)ww * hh * 3 / 2, 128), out_nv;
            memcpy(nv12.data(), gray.data, (size_t)ww * hh);
            cv::Mat luma(hh, ww, cv::CV_8UC1, nv12.data());
            cv::Mat cpu_nv = comparator.PerformCPU_EdgeDetection(comparator.PerformCPU_GaussianBlur(luma, ksize, sigma, t2, logger), t2, logger);
            cl_command_queue q = controller.CreateCommandQueue(context, devices[0]);
            cl_int w2 = ww, h2 = hh;
            int ks = ksize;
            float sg = sigma;
            controller.PerformFused(ks, sg, &context, &q, nullptr, &events, &nv12, &out_nv, w2, h2, logger, RIP_FMT_NV12);
            const bool ok_nv = out_nv.size() == (size_t)ww * hh && !memcmp(out_nv.data(), cpu_nv.data, out_nv.size());
            std::printf("PerformFused NV12 %dx%d: %s\n", ww, hh, ok_nv ? "exact" : "MISMATCH");
            if (!ok_nv) failures++;
            // a short NV12 buffer must be rejected, not read past its end
            std::vector<unsigned char> small((size_t)ww * hh, 0), out_small;
            controller.PerformFused(ks, sg, &context, &q, nullptr, &events, &small, &out_small, w2, h2, logger, RIP_FMT_NV12);
            if (!out_small.empty()) { std::printf("short NV12 buffer was accepted\n"); failures++; }
            controller.Cleanup(0, q);
        }
        // [new] streaming form: six frames, three in flight, every result equal to the blocking call's
        {
            cl_context c2 = 0; cl_command_queue q2 = 0; cl_program p2 = 0; cl_kernel k2 = 0;
            handler.InitOpenCL(controller, &c2, &q2, &p2, &k2, "FUSED", logger);
            cl_int w2 = rgba.cols, h2 = rgba.rows;
            std::vector<unsigned char> want = handler.PerformOpenCL(controller, rgba, &c2, &q2, &k2, w2, h2, logger, "FUSED");
            std::vector<int> inflight;
            int got = 0;
            bool ok_s = !want.empty();
            for (int i = 0; i < 6 && ok_s; i++) {
                const int hnd = handler.SubmitOpenCL(controller, rgba, &q2, w2, h2, logger, "FUSED");
                if (!hnd) { ok_s = false; break; }
                inflight.push_back(hnd);
                if (inflight.size() == 3) {
                    std::vector<cl_ulong> ev;
                    ok_s = handler.CollectOpenCL(controller, inflight.front(), logger, &ev) == want && ev.size() == 6;
                    inflight.erase(inflight.begin());
                    got++;
                }
            }
            for (int hnd : inflight) { ok_s = (handler.CollectOpenCL(controller, hnd, logger) == want) && ok_s; got++; }
            std::printf("SubmitOpenCL/CollectOpenCL x%d: %s\n", got, ok_s && got == 6 ? "exact" : "MISMATCH");
            if (!ok_s || got != 6) failures++;
            controller.Cleanup(c2, q2, p2, k2);
        }
        controller.Cleanup(context);
        files.WriteResultsToCSV(work + "/results.csv", results);
    }
    std::printf("%s\n", failures ? "FAILED" : "OK");
    return failures ? 1 : 0;
}
