// rip_headless.cpp -- the reference's applications without a window, a camera or OpenCV: a program, not a test.
//
//   rip_headless images <dir> [--iterations N] [--ksize K] [--sigma S] [--csv FILE] [--hbm-peak GBps] [--synthetic WxH]...
//       PerformOnImages (RT/RealtimeImageProcessing.cpp:32-138): every method (GRAYSCALE, EDGE, GAUSSIAN, FUSED) on every
//       image of <dir> (*.ppm / *.pgm; *.jpg / *.png where the host classes are built with OpenCV) through
//       ProgramHandler::PerformOpenCL, the Comparator's CPU path next to it, MAE and largest error, and the results
//       table with the reference's 11 columns plus max_abs_err, Mpix_s, fps, GBps, pct_hbm_peak, n_gpus
//       (FileHandler::WriteExtendedResultsToCSV).  --synthetic adds a generated frame of that size to the list.
//   rip_headless stream <WxH> [--frames N] [--inflight D] [--method M] [--ksize K] [--sigma S]
//       the per-frame loop of PerformOnCamera (RT/RealtimeImageProcessing.cpp:325-418) on synthetic RGBA frames: first the
//       blocking per-frame call (ProgramHandler::PerformOpenCL(cv::Mat), RT/src/ProgramHandler.cpp:259-329), then the
//       streaming pair SubmitOpenCL / CollectOpenCL with D frames in flight; prints ms per frame and frames per second.
//
// Build: tools/build_tools.sh (g++ against librip_host.so / librip_cuda.so).  Exit status 0 = every GPU result
// matched the CPU path bit for bit.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "Comparator.hpp"
#include "Controller.hpp"
#include "FileHandler.hpp"
#include "Logger.hpp"
#include "ProgramHandler.hpp"

namespace fs = std::filesystem;

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static const char *arg_value(int argc, char **argv, const char *name, const char *dflt)
{
    for (int i = 0; i + 1 < argc; i++)
        if (!strcmp(argv[i], name)) return argv[i + 1];
    return dflt;
}

static bool parse_size(const char *s, int *w, int *h) { return std::sscanf(s, "%dx%d", w, h) == 2 && *w > 0 && *h > 0; }

// a synthetic BGR frame: smooth structure plus noise (seeded), written as a PPM so that it takes the same path as a file
static std::string write_synthetic(const std::string &dir, int w, int h)
{
    cv::Mat bgr(h, w, cv::CV_8UC3);
    std::mt19937 rng(0xB200u + (unsigned)w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            unsigned char *p = bgr.data + ((size_t)y * w + x) * 3;
            const int base = (x * 255 / (w > 1 ? w - 1 : 1) + y * 255 / (h > 1 ? h - 1 : 1)) / 2;
            for (int c = 0; c < 3; c++) p[c] = (unsigned char)((base + (int)(rng() % 64) + 40 * c) & 255);
        }
    fs::create_directories(dir);
    const std::string path = dir + "/synthetic_" + std::to_string(w) + "x" + std::to_string(h) + ".ppm";
    cv::imwrite(path, bgr);
    return path;
}

static int run_images(int argc, char **argv)
{
    const std::string dir = argv[2];
    const int iterations = atoi(arg_value(argc, argv, "--iterations", "10"));
    const int ksize = atoi(arg_value(argc, argv, "--ksize", "5"));
    const float sigma = (float)atof(arg_value(argc, argv, "--sigma", "1.0"));
    const std::string csv = arg_value(argc, argv, "--csv", "results_extended.csv");
    const double hbm_peak = atof(arg_value(argc, argv, "--hbm-peak", "0"));

    Logger &logger = Logger::getInstance();
    ProgramHandler handler(iterations, false, false, false, true, ksize, sigma);
    handler.InitLogger(logger, Logger::LogLevel::ERROR, false);
    handler.SetDeviceProperties(0, 0);
    handler.AddKernels({"grayscale_images.cl", "grayscale_base.cl"}, "GRAYSCALE");
    handler.AddKernels({"edge_images.cl", "edge_base.cl"}, "EDGE");
    handler.AddKernels({"gaussian_images.cl", "gaussian_base.cl"}, "GAUSSIAN");
    handler.AddKernels({"fused"}, "FUSED");

    FileHandler files;
    Comparator comparator(4, iterations > 3 ? 3 : iterations);   // (the CPU side is the slow one: three iterations at most)
    Controller controller;
    std::vector<std::string> images;
    if (fs::is_directory(dir)) images = files.LoadImages(dir);
    for (int i = 3; i + 1 < argc; i++)
        if (!strcmp(argv[i], "--synthetic")) {
            int w = 0, h = 0;
            if (!parse_size(argv[i + 1], &w, &h)) { std::cerr << "bad --synthetic size " << argv[i + 1] << std::endl; return 2; }
            images.push_back(write_synthetic(dir, w, h));
        }
    if (images.empty()) { std::cerr << "no images under " << dir << " (and no --synthetic WxH)" << std::endl; return 2; }

    // algorithmic bytes per pixel of each kernel: RGBA in; (g,g,g,255) / RGBA / u8 out
    struct M { const char *name; double bytes_per_px; };
    const M methods[] = {{"GRAYSCALE", 8.0}, {"EDGE", 5.0}, {"GAUSSIAN", 8.0}, {"FUSED", 5.0}};
    std::vector<FileHandler::ExtendedRow> rows;
    int failures = 0;
    std::printf("%-10s %-28s %-11s %12s %12s %12s %10s %8s %7s\n", "method", "image", "resolution", "cpu ms", "gpu e2e ms", "kernel ms", "Mpix/s", "MAE", "max");
    for (const M &m : methods) {
        cl_context context = 0; cl_command_queue queue = 0; cl_program program = 0; cl_kernel kernel = 0;
        handler.InitOpenCL(controller, &context, &queue, &program, &kernel, m.name, logger);
        for (const std::string &path : images) {
            cv::Mat bgr = cv::imread(path, cv::IMREAD_COLOR), rgba;
            if (bgr.empty()) { std::cerr << "cannot read " << path << std::endl; failures++; continue; }
            cv::cvtColor(bgr, rgba, cv::COLOR_BGR2RGBA);
            double cpu_ms = 0.0, t = 0.0;
            cv::Mat cpu;
            if (!strcmp(m.name, "GRAYSCALE")) cpu = comparator.PerformCPU_Grayscaling(path, cpu_ms, logger);
            else if (!strcmp(m.name, "GAUSSIAN")) cpu = comparator.PerformCPU_GaussianBlur(rgba, ksize, sigma, cpu_ms, logger);
            else if (!strcmp(m.name, "EDGE")) {
                cv::Mat gray = comparator.PerformCPU_Grayscaling(rgba, false, t, logger);
                cpu = comparator.PerformCPU_EdgeDetection(gray, cpu_ms, logger);
                cpu_ms += t;   // the GPU method does both stages
            } else cpu = comparator.PerformCPU_Fused(rgba, false, ksize, sigma, cpu_ms, logger);
            double e2e = 0, wr = 0, ke = 0, rd = 0, op = 0;
            cl_int w = 0, h = 0;
            std::vector<unsigned char> out = handler.PerformOpenCL(controller, path, &context, &queue, &kernel, e2e, wr, ke, rd, op, w, h, logger, m.name);
            if (out.empty() || cpu.empty()) { std::cerr << m.name << " " << path << ": no output" << std::endl; failures++; continue; }
            const Comparator::Report rep = comparator.CompareGPUvsCPU(out, cpu, cpu_ms, logger);
            if (!rep.exact()) failures++;
            FileHandler::ResultRow base(logger.getCurrentTime(), fs::path(path).filename().string(), std::to_string(w) + "x" + std::to_string(h), iterations,
                                        cpu_ms, e2e, ke, wr, rd, op, rep.mae);
            rows.push_back(FileHandler::Extend(base, m.name, w, h, rep.max_abs, m.bytes_per_px, hbm_peak, 1));
            std::printf("%-10s %-28s %-11s %12.3f %12.3f %12.4f %10.0f %8g %7d%s\n", m.name, fs::path(path).filename().string().c_str(),
                        (std::to_string(w) + "x" + std::to_string(h)).c_str(), cpu_ms, e2e, ke, rows.back().mpix_s, rep.mae, rep.max_abs,
                        rep.exact() ? "" : "  MISMATCH");
        }
        controller.Cleanup(context, queue, program, kernel);
    }
    files.WriteExtendedResultsToCSV(csv, rows);
    std::printf("%zu rows -> %s; %s\n", rows.size(), csv.c_str(), failures ? "FAILED" : "every GPU result equals the CPU path");
    return failures ? 1 : 0;
}

static int run_stream(int argc, char **argv)
{
    int w = 0, h = 0;
    if (!parse_size(argv[2], &w, &h)) { std::cerr << "bad size " << argv[2] << std::endl; return 2; }
    const int frames = atoi(arg_value(argc, argv, "--frames", "200"));
    const int inflight = atoi(arg_value(argc, argv, "--inflight", "3"));
    const std::string method = arg_value(argc, argv, "--method", "FUSED");
    const int ksize = atoi(arg_value(argc, argv, "--ksize", "5"));
    const float sigma = (float)atof(arg_value(argc, argv, "--sigma", "1.0"));

    Logger &logger = Logger::getInstance();
    ProgramHandler handler(1, false, false, false, true, ksize, sigma);
    handler.InitLogger(logger, Logger::LogLevel::ERROR, false);
    handler.SetDeviceProperties(0, 0);
    handler.AddKernels({"grayscale_images.cl", "grayscale_base.cl"}, "GRAYSCALE");
    handler.AddKernels({"edge_images.cl", "edge_base.cl"}, "EDGE");
    handler.AddKernels({"gaussian_images.cl", "gaussian_base.cl"}, "GAUSSIAN");
    handler.AddKernels({"fused"}, "FUSED");
    Controller controller;
    cl_context context = 0; cl_command_queue queue = 0; cl_program program = 0; cl_kernel kernel = 0;
    handler.InitOpenCL(controller, &context, &queue, &program, &kernel, method, logger);

    // a few distinct RGBA frames, cycled (a camera would deliver a new cv::Mat per frame)
    const int n_src = 4;
    std::vector<cv::Mat> src;
    std::mt19937 rng(0xB200u);
    for (int i = 0; i < n_src; i++) {
        cv::Mat f(h, w, cv::CV_8UC4);
        for (size_t k = 0; k < (size_t)w * h; k++) {
            const uint32_t v = rng() | 0xff000000u;
            memcpy(f.data + 4 * k, &v, 4);
        }
        src.push_back(f);
    }
    cl_int cw = w, ch = h;
    // the blocking per-frame call
    std::vector<std::vector<unsigned char>> want(n_src);
    for (int i = 0; i < n_src; i++) want[i] = handler.PerformOpenCL(controller, src[i], &context, &queue, &kernel, cw, ch, logger, method);
    if (want[0].empty()) { std::cerr << "the per-frame call produced no output" << std::endl; return 1; }
    double t0 = now_ms();
    for (int i = 0; i < frames; i++) handler.PerformOpenCL(controller, src[i % n_src], &context, &queue, &kernel, cw, ch, logger, method);
    const double ms_blocking = (now_ms() - t0) / frames;
    // the streaming pair, `inflight` frames in flight
    std::vector<int> q;
    bool ok = true;
    int collected = 0;
    t0 = now_ms();
    for (int i = 0; i < frames; i++) {
        const int hnd = handler.SubmitOpenCL(controller, src[i % n_src], &queue, cw, ch, logger, method);
        if (!hnd) { ok = false; break; }
        q.push_back(hnd);
        if ((int)q.size() == inflight) {
            ok = (handler.CollectOpenCL(controller, q.front(), logger) == want[collected % n_src]) && ok;
            q.erase(q.begin());
            collected++;
        }
    }
    for (int hnd : q) { ok = (handler.CollectOpenCL(controller, hnd, logger) == want[collected % n_src]) && ok; collected++; }
    const double ms_stream = (now_ms() - t0) / frames;
    controller.Cleanup(context, queue, program, kernel);
    std::printf("%s %dx%d RGBA, %d frames through ProgramHandler (pageable cv::Mat in, std::vector out):\n", method.c_str(), w, h, frames);
    std::printf("  PerformOpenCL(cv::Mat), one frame at a time : %8.3f ms per frame  %8.1f frames/s  %9.0f Mpix/s\n", ms_blocking, 1e3 / ms_blocking,
                (double)w * h / ms_blocking / 1e3);
    std::printf("  SubmitOpenCL / CollectOpenCL, %d in flight    : %8.3f ms per frame  %8.1f frames/s  %9.0f Mpix/s  (%s)\n", inflight, ms_stream,
                1e3 / ms_stream, (double)w * h / ms_stream / 1e3, ok && collected == frames ? "every frame equals the blocking call's result" : "MISMATCH");
    return ok && collected == frames ? 0 : 1;
}

int main(int argc, char **argv)
{
    if (argc >= 3 && !strcmp(argv[1], "images")) return run_images(argc, argv);
    if (argc >= 3 && !strcmp(argv[1], "stream")) return run_stream(argc, argv);
    std::cerr << "usage: rip_headless images <dir> [--iterations N] [--ksize K] [--sigma S] [--csv FILE] [--hbm-peak GBps] [--synthetic WxH]...\n"
                 "       rip_headless stream <WxH> [--frames N] [--inflight D] [--method GRAYSCALE|EDGE|GAUSSIAN|FUSED] [--ksize K] [--sigma S]" << std::endl;
    return 2;
}
