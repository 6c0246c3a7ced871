// rip_api.cu -- the C ABI of librip_cuda.so (include/rip_cuda.h): discovery, handles, memory,
// device-resident ops and the host-buffer pipeline that Controller::PerformCL* drives.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rip_common.cuh"
#include "rip_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ============================================================================================
// error plumbing / counters
// ============================================================================================
namespace rip {

static thread_local char g_err[512] = "no error";
static std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return (int)e;
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count(int device)
{
    static int cache[64];
    static std::atomic<bool> have[64];
    if (device < 0 || device >= 64) return 148;
    if (!have[device].load(std::memory_order_acquire)) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cache[device] = v;
        have[device].store(true, std::memory_order_release);
    }
    return cache[device];
}

static int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RIP_ENODEV, "no CUDA device available (%s); librip_cuda has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(RIP_EINVAL, "device index %d out of range [0,%d)", device, n);
    return RIP_OK;
}

static int channels_of(int fmt)
{
    switch (fmt) {
    case RIP_FMT_GRAY8: case RIP_FMT_NV12: return 1;   // (NV12: the luma plane is the image)
    case RIP_FMT_RGB8: case RIP_FMT_BGR8: return 3;
    case RIP_FMT_RGBA8: case RIP_FMT_BGRA8: return 4;
    default: return 0;
    }
}

// bytes from one frame of a batch to the next
static size_t frame_bytes_of(int fmt, int width, int height)
{
    if (fmt == RIP_FMT_NV12) return (size_t)width * height * 3 / 2;
    return (size_t)width * height * channels_of(fmt);
}

static int load_weights(Weights &dst, int ksize, const float *weights, const char *who)
{
    if (ksize < 1 || ksize > RIP_MAX_KSIZE || (ksize & 1) == 0)
        return fail(RIP_EINVAL, "%s: kernel size must be odd and in [1,%d] (got %d)", who, RIP_MAX_KSIZE, ksize);
    if (!weights) return fail(RIP_EINVAL, "%s: weights is NULL", who);
    memcpy(dst.w, weights, sizeof(float) * ksize * ksize);
    return RIP_OK;
}

}  // namespace rip

using namespace rip;

// ============================================================================================
// handle types
// ============================================================================================
namespace {
constexpr int kSets = 3;  // chunk buffers in flight per device (H2D / kernel / D2H overlap)

struct BufSet {
    cudaStream_t stream = nullptr;
    void *d_in = nullptr, *d_out = nullptr, *d_ws = nullptr;
    size_t in_cap = 0, out_cap = 0, ws_cap = 0;
};

struct DevState {
    int device = 0;
    BufSet set[kSets];
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

int ensure(void **p, size_t *cap, size_t need)
{
    if (need <= *cap) return RIP_OK;
    if (*p) RIP_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    // grow in 1 MiB steps so slightly different frame sizes reuse the allocation
    const size_t want = (need + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
    RIP_CUDA(cudaMalloc(p, want));
    *cap = want;
    return RIP_OK;
}
}  // namespace

struct rip_ctx {
    std::vector<DevState> devs;
};
struct rip_module {
    rip_ctx *ctx;
    int op;
    std::string variant;
};
struct rip_kernel {
    rip_module *module;
    int op;
    std::string name;
};
struct rip_event {
    int device;
    cudaEvent_t ev;
};

// ============================================================================================
// discovery
// ============================================================================================
extern "C" int rip_abi_version(void) { return RIP_ABI_VERSION; }
extern "C" const char *rip_last_error_string(void) { return g_err; }

extern "C" int rip_launch_count(uint64_t *launches)
{
    if (!launches) return fail(RIP_EINVAL, "rip_launch_count: NULL");
    *launches = g_launches.load(std::memory_order_relaxed);
    return RIP_OK;
}

extern "C" int rip_device_count(int *count)
{
    if (!count) return fail(RIP_EINVAL, "rip_device_count: NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(RIP_ENODEV, "no CUDA device available (%s); librip_cuda has no CPU fallback", cudaGetErrorString(e));
    }
    *count = n;
    return RIP_OK;
}

extern "C" int rip_device_get_info(int device, rip_device_info *info)
{
    if (!info) return fail(RIP_EINVAL, "rip_device_get_info: NULL");
    if (int rc = check_device(device)) return rc;
    cudaDeviceProp p;
    RIP_CUDA(cudaGetDeviceProperties(&p, device));
    memset(info, 0, sizeof(*info));
    snprintf(info->name, sizeof(info->name), "%s", p.name);
    info->sm_count = p.multiProcessorCount;
    info->cc_major = p.major;
    info->cc_minor = p.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    info->clock_khz = khz;
    info->l2_bytes = p.l2CacheSize;
    info->global_mem_bytes = p.totalGlobalMem;
    info->smem_per_sm_bytes = p.sharedMemPerMultiprocessor;
    return RIP_OK;
}

extern "C" int rip_device_pci_bus_id(int device, char *buf, size_t buf_len)
{
    if (int rc = check_device(device)) return rc;
    if (!buf || buf_len < 13) return fail(RIP_EINVAL, "rip_device_pci_bus_id: buffer of at least 13 bytes needed");
    RIP_CUDA(cudaDeviceGetPCIBusId(buf, (int)buf_len, device));
    return RIP_OK;
}

extern "C" int rip_device_name(int device, char *buf, size_t buf_len)
{
    if (!buf || buf_len == 0) return fail(RIP_EINVAL, "rip_device_name: NULL buffer");
    rip_device_info info;
    if (int rc = rip_device_get_info(device, &info)) return rc;
    snprintf(buf, buf_len, "%s", info.name);
    return RIP_OK;
}

// ============================================================================================
// context / module / kernel
// ============================================================================================
extern "C" int rip_ctx_create(const int *devices, int n_devices, rip_ctx **out)
{
    if (!out) return fail(RIP_EINVAL, "rip_ctx_create: NULL");
    *out = nullptr;
    int count = 0;
    if (int rc = rip_device_count(&count)) return rc;
    if (count <= 0) return fail(RIP_ENODEV, "no CUDA device available; librip_cuda has no CPU fallback");
    std::vector<int> devs;
    if (!devices || n_devices <= 0) {
        devs.push_back(0);
    } else {
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= count)
                return fail(RIP_EINVAL, "rip_ctx_create: device %d out of range [0,%d)", devices[i], count);
            devs.push_back(devices[i]);
        }
    }
    rip_ctx *ctx = new (std::nothrow) rip_ctx();
    if (!ctx) return fail(RIP_ENOMEM, "rip_ctx_create: out of host memory");
    ctx->devs.resize(devs.size());
    for (size_t i = 0; i < devs.size(); i++) {
        DevState &d = ctx->devs[i];
        d.device = devs[i];
        DeviceGuard g(d.device);
        for (int k = 0; k < kSets; k++) {
            cudaError_t e = cudaStreamCreateWithFlags(&d.set[k].stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) {
                rip_ctx_destroy(ctx);
                return cuda_fail(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
            }
        }
        for (int k = 0; k < 4; k++) {
            cudaError_t e = cudaEventCreate(&d.ev[k]);
            if (e != cudaSuccess) {
                rip_ctx_destroy(ctx);
                return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__);
            }
        }
    }
    *out = ctx;
    return RIP_OK;
}

extern "C" int rip_ctx_destroy(rip_ctx *ctx)
{
    if (!ctx) return RIP_OK;
    for (DevState &d : ctx->devs) {
        DeviceGuard g(d.device);
        for (int k = 0; k < kSets; k++) {
            BufSet &b = d.set[k];
            if (b.stream) {
                cudaStreamSynchronize(b.stream);
                cudaStreamDestroy(b.stream);
            }
            if (b.d_in) cudaFree(b.d_in);
            if (b.d_out) cudaFree(b.d_out);
            if (b.d_ws) cudaFree(b.d_ws);
        }
        for (int k = 0; k < 4; k++)
            if (d.ev[k]) cudaEventDestroy(d.ev[k]);
    }
    delete ctx;
    return RIP_OK;
}

extern "C" int rip_ctx_device_count(const rip_ctx *ctx, int *n)
{
    if (!ctx || !n) return fail(RIP_EINVAL, "rip_ctx_device_count: NULL");
    *n = (int)ctx->devs.size();
    return RIP_OK;
}

extern "C" int rip_ctx_device(const rip_ctx *ctx, int index, int *device)
{
    if (!ctx || !device) return fail(RIP_EINVAL, "rip_ctx_device: NULL");
    if (index < 0 || index >= (int)ctx->devs.size()) return fail(RIP_EINVAL, "rip_ctx_device: index %d out of range", index);
    *device = ctx->devs[index].device;
    return RIP_OK;
}

static int op_of_variant(const std::string &v)
{
    // the reference's kernel file names (RealtimeImageProcessing.cpp:28-30) select the operation
    if (v.find("grayscale") != std::string::npos) return RIP_OP_GRAY;
    if (v.find("gaussian") != std::string::npos) return RIP_OP_GAUSSIAN;
    if (v.find("edge") != std::string::npos || v.find("sobel") != std::string::npos) return RIP_OP_EDGE;
    if (v.find("fused") != std::string::npos) return RIP_OP_FUSED;
    return -1;
}

extern "C" int rip_module_load(rip_ctx *ctx, const char *variant, rip_module **module)
{
    if (!ctx || !variant || !module) return fail(RIP_EINVAL, "rip_module_load: NULL");
    *module = nullptr;
    const int op = op_of_variant(variant);
    if (op < 0) return fail(RIP_EINVAL, "rip_module_load: unknown kernel variant '%s'", variant);
    *module = new (std::nothrow) rip_module{ctx, op, variant};
    return *module ? RIP_OK : fail(RIP_ENOMEM, "rip_module_load: out of host memory");
}

extern "C" int rip_module_release(rip_module *m)
{
    delete m;
    return RIP_OK;
}

extern "C" int rip_kernel_get(rip_module *module, const char *kernel_name, rip_kernel **kernel)
{
    if (!module || !kernel_name || !kernel) return fail(RIP_EINVAL, "rip_kernel_get: NULL");
    *kernel = nullptr;
    // entry-point names of the reference kernels (ProgramHandler.cpp:69-78)
    int op = -1;
    if (!strcmp(kernel_name, "grayscale")) op = RIP_OP_GRAY;
    else if (!strcmp(kernel_name, "gaussian_blur")) op = RIP_OP_GAUSSIAN;
    else if (!strcmp(kernel_name, "sobel_edge_detection")) op = RIP_OP_EDGE;
    else if (!strcmp(kernel_name, "fused")) op = RIP_OP_FUSED;
    if (op < 0 || op != module->op)
        return fail(RIP_EINVAL, "rip_kernel_get: module '%s' has no kernel '%s'", module->variant.c_str(), kernel_name);
    *kernel = new (std::nothrow) rip_kernel{module, op, kernel_name};
    return *kernel ? RIP_OK : fail(RIP_ENOMEM, "rip_kernel_get: out of host memory");
}

extern "C" int rip_kernel_release(rip_kernel *k)
{
    delete k;
    return RIP_OK;
}

extern "C" int rip_kernel_op(const rip_kernel *k, int *op)
{
    if (!k || !op) return fail(RIP_EINVAL, "rip_kernel_op: NULL");
    *op = k->op;
    return RIP_OK;
}

// ============================================================================================
// streams, events, memory
// ============================================================================================
extern "C" int rip_stream_create(int device, rip_stream *stream)
{
    if (!stream) return fail(RIP_EINVAL, "rip_stream_create: NULL");
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    cudaStream_t s;
    RIP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return RIP_OK;
}

extern "C" int rip_stream_destroy(int device, rip_stream stream)
{
    if (!stream) return RIP_OK;
    DeviceGuard g(device);
    RIP_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return RIP_OK;
}

extern "C" int rip_stream_sync(int device, rip_stream stream)
{
    DeviceGuard g(device);
    RIP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return RIP_OK;
}

extern "C" int rip_device_sync(int device)
{
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    RIP_CUDA(cudaDeviceSynchronize());
    return RIP_OK;
}

extern "C" int rip_event_create(int device, rip_event **event)
{
    if (!event) return fail(RIP_EINVAL, "rip_event_create: NULL");
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    cudaEvent_t e;
    RIP_CUDA(cudaEventCreate(&e));
    *event = new (std::nothrow) rip_event{device, e};
    return *event ? RIP_OK : fail(RIP_ENOMEM, "rip_event_create: out of host memory");
}

extern "C" int rip_event_destroy(rip_event *event)
{
    if (!event) return RIP_OK;
    DeviceGuard g(event->device);
    cudaEventDestroy(event->ev);
    delete event;
    return RIP_OK;
}

extern "C" int rip_event_record(rip_event *event, rip_stream stream)
{
    if (!event) return fail(RIP_EINVAL, "rip_event_record: NULL");
    DeviceGuard g(event->device);
    RIP_CUDA(cudaEventRecord(event->ev, (cudaStream_t)stream));
    return RIP_OK;
}

extern "C" int rip_event_sync(rip_event *event)
{
    if (!event) return fail(RIP_EINVAL, "rip_event_sync: NULL");
    DeviceGuard g(event->device);
    RIP_CUDA(cudaEventSynchronize(event->ev));
    return RIP_OK;
}

extern "C" int rip_event_elapsed_ns(rip_event *start, rip_event *stop, uint64_t *ns)
{
    if (!start || !stop || !ns) return fail(RIP_EINVAL, "rip_event_elapsed_ns: NULL");
    DeviceGuard g(start->device);
    float ms = 0.f;
    RIP_CUDA(cudaEventElapsedTime(&ms, start->ev, stop->ev));
    *ns = (uint64_t)llround((double)ms * 1e6);
    return RIP_OK;
}

extern "C" int rip_malloc_device(int device, size_t bytes, void **d_ptr)
{
    if (!d_ptr) return fail(RIP_EINVAL, "rip_malloc_device: NULL");
    *d_ptr = nullptr;
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    RIP_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return RIP_OK;
}

extern "C" int rip_free_device(int device, void *d_ptr)
{
    if (!d_ptr) return RIP_OK;
    DeviceGuard g(device);
    RIP_CUDA(cudaFree(d_ptr));
    return RIP_OK;
}

extern "C" int rip_malloc_pinned(size_t bytes, void **h_ptr)
{
    if (!h_ptr) return fail(RIP_EINVAL, "rip_malloc_pinned: NULL");
    *h_ptr = nullptr;
    if (int rc = check_device(0)) return rc;
    RIP_CUDA(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return RIP_OK;
}

extern "C" int rip_free_pinned(void *h_ptr)
{
    if (!h_ptr) return RIP_OK;
    RIP_CUDA(cudaFreeHost(h_ptr));
    return RIP_OK;
}

extern "C" int rip_memcpy_h2d_async(int device, void *d_dst, const void *h_src, size_t bytes, rip_stream stream)
{
    if (!d_dst || !h_src) return fail(RIP_EINVAL, "rip_memcpy_h2d_async: NULL");
    DeviceGuard g(device);
    RIP_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return RIP_OK;
}

extern "C" int rip_memcpy_d2h_async(int device, void *h_dst, const void *d_src, size_t bytes, rip_stream stream)
{
    if (!h_dst || !d_src) return fail(RIP_EINVAL, "rip_memcpy_d2h_async: NULL");
    DeviceGuard g(device);
    RIP_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return RIP_OK;
}

extern "C" int rip_memset_device_async(int device, void *d_dst, int value, size_t bytes, rip_stream stream)
{
    if (!d_dst) return fail(RIP_EINVAL, "rip_memset_device_async: NULL");
    DeviceGuard g(device);
    RIP_CUDA(cudaMemsetAsync(d_dst, value, bytes, (cudaStream_t)stream));
    return RIP_OK;
}

// ============================================================================================
// Gaussian weights.  Same typing as the reference generator (Controller.cpp:352-372): the exponent
// argument is a float quotient, exp() is the double overload, the 2*pi*sigma^2 divisor is double,
// each tap is rounded to float when stored, the running sum and the normalising divide are float.
// ============================================================================================
extern "C" int rip_gauss_weights(int ksize, float sigma, float *weights)
{
    if (!weights) return fail(RIP_EINVAL, "rip_gauss_weights: NULL");
    if (ksize < 1 || ksize > RIP_MAX_KSIZE || (ksize & 1) == 0)
        return fail(RIP_EINVAL, "rip_gauss_weights: kernel size must be odd and in [1,%d] (got %d)", RIP_MAX_KSIZE, ksize);
    if (!(sigma > 0.0f)) return fail(RIP_EINVAL, "rip_gauss_weights: sigma must be positive");
    const int r = ksize / 2;
    const float two_sigma_sq = 2 * sigma * sigma;
    const double norm = 2 * M_PI * sigma * sigma;
    volatile float total = 0.0f;  // volatile: keep the accumulation in float on every compiler
    float *w = weights;
    for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
            const float e = (float)(-(dx * dx + dy * dy)) / two_sigma_sq;
            const float tap = (float)(std::exp((double)e) / norm);
            *w++ = tap;
            total = total + tap;
        }
    const float t = total;
    for (int i = 0; i < ksize * ksize; i++) weights[i] = weights[i] / t;
    return RIP_OK;
}

// ============================================================================================
// device-resident ops
// ============================================================================================
static int check_image(const char *who, const void *in, const void *out, int w, int h, int n)
{
    if (!in || !out) return fail(RIP_EINVAL, "%s: NULL device pointer", who);
    if (w <= 0 || h <= 0 || n <= 0) return fail(RIP_EINVAL, "%s: width, height and n_frames must be positive (got %d x %d x %d)", who, w, h, n);
    return RIP_OK;
}

extern "C" int rip_gray(int device, rip_stream stream, const uint8_t *d_in, uint8_t *d_out, int width, int height,
                        int n_frames, int in_format, int out_mode)
{
    if (int rc = check_device(device)) return rc;
    if (int rc = check_image("rip_gray", d_in, d_out, width, height, n_frames)) return rc;
    if (out_mode != RIP_GRAY_OUT_U8 && out_mode != RIP_GRAY_OUT_RGBA) return fail(RIP_EINVAL, "rip_gray: bad out_mode %d", out_mode);
    DeviceGuard g(device);
    return launch_gray((cudaStream_t)stream, d_in, d_out, (long long)width * height * n_frames, in_format, out_mode, device);
}

extern "C" int rip_gauss(int device, rip_stream stream, const uint8_t *d_in, uint8_t *d_out, int width, int height,
                         int n_frames, int channels, int ksize, const float *weights)
{
    if (int rc = check_device(device)) return rc;
    if (int rc = check_image("rip_gauss", d_in, d_out, width, height, n_frames)) return rc;
    Weights wts;
    if (int rc = load_weights(wts, ksize, weights, "rip_gauss")) return rc;
    DeviceGuard g(device);
    return launch_blur((cudaStream_t)stream, d_in, d_out, width, height, n_frames, channels, ksize, wts, 0, height, 0, height);
}

extern "C" int rip_sobel(int device, rip_stream stream, const uint8_t *d_in, uint8_t *d_out, int width, int height,
                         int n_frames, int in_format)
{
    if (int rc = check_device(device)) return rc;
    if (int rc = check_image("rip_sobel", d_in, d_out, width, height, n_frames)) return rc;
    if (channels_of(in_format) == 0) return fail(RIP_EINVAL, "rip_sobel: unsupported input format %d", in_format);
    DeviceGuard g(device);
    if (fused_supported(width, height, in_format, 0, d_in, d_out))
        return launch_fused((cudaStream_t)stream, d_in, d_out, width, height, n_frames, in_format, false, nullptr, 0, height, 0,
                            height, device);
    if (in_format == RIP_FMT_NV12 && n_frames > 1)
        return fail(RIP_EUNSUPPORTED, "rip_sobel: NV12 batches need width %% 4 == 0, an even height and 4-byte aligned buffers");
    return launch_sobel((cudaStream_t)stream, d_in, d_out, width, height, n_frames, in_format == RIP_FMT_NV12 ? RIP_FMT_GRAY8 : in_format, 0,
                        height, 0, height);
}

extern "C" int rip_fused_workspace_bytes(int width, int in_rows, int n_frames, int ksize, size_t *bytes)
{
    if (!bytes) return fail(RIP_EINVAL, "rip_fused_workspace_bytes: NULL");
    if (width <= 0 || in_rows <= 0 || n_frames <= 0) return fail(RIP_EINVAL, "rip_fused_workspace_bytes: bad shape");
    // staged path (any ksize but 5, or shapes the single-kernel path rejects): gray band + blurred band
    (void)ksize;
    *bytes = (size_t)2 * width * in_rows * n_frames;
    return RIP_OK;
}

extern "C" int rip_fused(int device, rip_stream stream, const uint8_t *d_in, uint8_t *d_out, int width, int height,
                         int n_frames, int in_format, int ksize, const float *weights, int in_row0, int in_rows,
                         int out_row0, int out_rows, void *d_workspace, size_t workspace_bytes)
{
    if (int rc = check_device(device)) return rc;
    if (int rc = check_image("rip_fused", d_in, d_out, width, height, n_frames)) return rc;
    const int cn = channels_of(in_format);
    if (cn == 0) return fail(RIP_EINVAL, "rip_fused: unsupported input format %d", in_format);
    Weights wts;
    if (int rc = load_weights(wts, ksize, weights, "rip_fused")) return rc;
    const int half = ksize / 2;
    if (out_row0 < 0 || out_rows <= 0 || out_row0 + out_rows > height)
        return fail(RIP_EINVAL, "rip_fused: output rows [%d,%d) outside the image (height %d)", out_row0, out_row0 + out_rows, height);
    // blurred rows the Sobel stage needs, then gray rows the blur stage needs (both clipped to the image)
    const int b0 = max(0, out_row0 - 1), b1 = min(height, out_row0 + out_rows + 1);
    const int g0 = max(0, b0 - half), g1 = min(height, b1 + half);
    if (in_row0 < 0 || in_rows <= 0 || in_row0 > g0 || in_row0 + in_rows < g1 || in_row0 + in_rows > height)
        return fail(RIP_EINVAL, "rip_fused: input band [%d,%d) does not cover rows [%d,%d) needed for output rows [%d,%d)",
                    in_row0, in_row0 + in_rows, g0, g1, out_row0, out_row0 + out_rows);
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    {
        float g3[3], thr;  // single-kernel path: 5x5, aligned shape, and weights the guard band can cover
        if (ksize == 5 && fused_supported(width, height, in_format, 5, d_in, d_out) && fused_plan_weights(wts.w, g3, &thr))
            return launch_fused(s, d_in, d_out, width, height, n_frames, in_format, true, wts.w, in_row0, in_rows, out_row0, out_rows, device);
    }

    // staged path: gray band -> exact KxK blur -> Sobel, through the caller's workspace
    size_t need = 0;
    rip_fused_workspace_bytes(width, in_rows, n_frames, ksize, &need);
    if (!d_workspace || workspace_bytes < need)
        return fail(RIP_EINVAL, "rip_fused: this shape runs the staged path and needs %zu bytes of workspace (got %zu)", need, workspace_bytes);
    uint8_t *ws_gray = (uint8_t *)d_workspace;
    uint8_t *ws_blur = ws_gray + (size_t)width * in_rows * n_frames;
    if (cn == 1) {   // the input is the gray image already
        if (in_format == RIP_FMT_NV12 && n_frames > 1)
            return fail(RIP_EUNSUPPORTED, "rip_fused: NV12 batches run the single-kernel path only (5x5 weights, width %% 4 == 0, aligned buffers)");
        ws_gray = const_cast<uint8_t *>(d_in);
    } else if (int rc = launch_gray(s, d_in, ws_gray, (long long)width * in_rows * n_frames, in_format, RIP_GRAY_OUT_U8, device)) {
        return rc;
    }
    if (int rc = launch_blur(s, ws_gray, ws_blur, width, height, n_frames, 1, ksize, wts, in_row0, in_rows, b0, b1 - b0)) return rc;
    return launch_sobel(s, ws_blur, d_out, width, height, n_frames, RIP_FMT_GRAY8, b0, b1 - b0, out_row0, out_rows);
}

// ============================================================================================
// diagnostics
// ============================================================================================
static unsigned long long *g_d_slow = nullptr;

extern "C" int rip_debug_slow_path_stats(int device, int enable, uint64_t *slow_pixels)
{
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    if (slow_pixels) *slow_pixels = 0;
    if (g_d_slow && slow_pixels) {
        unsigned long long v = 0;
        RIP_CUDA(cudaDeviceSynchronize());
        RIP_CUDA(cudaMemcpy(&v, g_d_slow, sizeof(v), cudaMemcpyDeviceToHost));
        *slow_pixels = v;
    }
    if (enable && !g_d_slow) {
        RIP_CUDA(cudaMalloc(&g_d_slow, sizeof(unsigned long long)));
    }
    if (g_d_slow) RIP_CUDA(cudaMemset(g_d_slow, 0, sizeof(unsigned long long)));
    if (!enable && g_d_slow) {
        fused_set_slow_counter(nullptr);
        blur_sep_set_slow_counter(nullptr);
        RIP_CUDA(cudaFree(g_d_slow));
        g_d_slow = nullptr;
    }
    if (enable) {
        fused_set_slow_counter(g_d_slow);
        blur_sep_set_slow_counter(g_d_slow);
    }
    return RIP_OK;
}

extern "C" int rip_debug_selftest(int device, uint64_t *checked, uint64_t *mismatches)
{
    if (!checked || !mismatches) return fail(RIP_EINVAL, "rip_debug_selftest: NULL");
    if (int rc = check_device(device)) return rc;
    DeviceGuard g(device);
    unsigned long long c = 0, m = 0;
    if (int rc = fused_selftest(device, &c, &m)) return rc;
    *checked = c;
    *mismatches = m;
    return RIP_OK;
}

// ============================================================================================
// host-buffer pipeline
// ============================================================================================
extern "C" int rip_out_bytes_per_frame(const rip_op_desc *desc, int width, int height, size_t *bytes)
{
    if (!desc || !bytes) return fail(RIP_EINVAL, "rip_out_bytes_per_frame: NULL");
    const size_t px = (size_t)width * height;
    switch (desc->op) {
    case RIP_OP_GRAY: *bytes = desc->gray_out == RIP_GRAY_OUT_RGBA ? px * 4 : px; return RIP_OK;
    case RIP_OP_EDGE: case RIP_OP_FUSED: *bytes = px; return RIP_OK;
    case RIP_OP_GAUSSIAN: *bytes = px * channels_of(desc->in_format); return RIP_OK;
    default: return fail(RIP_EINVAL, "unknown op %d", desc->op);
    }
}

namespace {

bool fused_single_kernel(int W, int H, int fmt, int ksize, const float *weights, const void *d_in, const void *d_out)
{
    float g3[3], thr;
    return ksize == 5 && weights && fused_supported(W, H, fmt, 5, (const uint8_t *)d_in, (const uint8_t *)d_out) &&
           fused_plan_weights(weights, g3, &thr);
}

struct Job {
    const rip_op_desc *desc;
    int W, H;
    int cn;
    size_t in_frame_bytes, out_frame_bytes;
};

// enqueue one operation on device-resident frames (whole frames)
int enqueue_op(const Job &job, int device, const BufSet &b, int n_frames)
{
    const rip_op_desc &d = *job.desc;
    cudaStream_t s = b.stream;
    const uint8_t *in = (const uint8_t *)b.d_in;
    uint8_t *out = (uint8_t *)b.d_out;
    switch (d.op) {
    case RIP_OP_GRAY:
        return rip_gray(device, s, in, out, job.W, job.H, n_frames, d.in_format, d.gray_out);
    case RIP_OP_EDGE:
        return rip_sobel(device, s, in, out, job.W, job.H, n_frames, d.in_format);
    case RIP_OP_GAUSSIAN:
        return rip_gauss(device, s, in, out, job.W, job.H, n_frames, job.cn, d.ksize, d.weights);
    case RIP_OP_FUSED:
        return rip_fused(device, s, in, out, job.W, job.H, n_frames, d.in_format, d.ksize, d.weights, 0, job.H, 0, job.H,
                         b.d_ws, b.ws_cap);
    default:
        return fail(RIP_EINVAL, "unknown op %d", d.op);
    }
}

int validate_desc(const rip_op_desc *desc, int *cn_out)
{
    if (!desc) return fail(RIP_EINVAL, "rip_process_host: NULL descriptor");
    const int cn = channels_of(desc->in_format);
    if (cn == 0) return fail(RIP_EINVAL, "rip_process_host: unsupported input format %d", desc->in_format);
    switch (desc->op) {
    case RIP_OP_GRAY:
        if (cn < 3) return fail(RIP_EINVAL, "GRAYSCALE needs a colour input");
        break;
    case RIP_OP_EDGE:
        break;
    case RIP_OP_GAUSSIAN:
        if ((cn != 1 && cn != 4) || desc->in_format == RIP_FMT_NV12) return fail(RIP_EINVAL, "GAUSSIAN runs on GRAY8 or RGBA8/BGRA8 input");
        if (!desc->weights) return fail(RIP_EINVAL, "GAUSSIAN needs weights");
        break;
    case RIP_OP_FUSED:
        if (!desc->weights) return fail(RIP_EINVAL, "FUSED needs weights");
        break;
    default:
        return fail(RIP_EINVAL, "unknown op %d", desc->op);
    }
    *cn_out = cn;
    return RIP_OK;
}

// Process frames [f0, f1) of the batch on one device: chunks of frames cycle through kSets buffer
// sets, each with its own stream, so the H2D of chunk i+1 and the D2H of chunk i-1 overlap the
// kernel of chunk i.  If prof != NULL the first chunk is bracketed by events.
int run_device(DevState &dev, const Job &job, const uint8_t *h_in, uint8_t *h_out, int f0, int f1, double *prof_ms, char *err,
               size_t err_len)
{
    int rc = RIP_OK;
    {
        DeviceGuard g(dev.device);
        const int n = f1 - f0;
        // ~48 MiB of input per chunk keeps three chunks in flight without hoarding HBM
        int chunk = (int)((size_t)(48u << 20) / job.in_frame_bytes);
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;
        int k = 0;
        for (int c0 = 0; c0 < n && rc == RIP_OK; c0 += chunk, k++) {
            const int cf = (n - c0 < chunk) ? n - c0 : chunk;
            BufSet &b = dev.set[k % kSets];
            if (k >= kSets) rc = (int)cudaStreamSynchronize(b.stream);  // previous user of this set is done
            if (rc) { rc = cuda_fail((cudaError_t)rc, "cudaStreamSynchronize", __FILE__, __LINE__); break; }
            if ((rc = ensure(&b.d_in, &b.in_cap, job.in_frame_bytes * cf))) break;
            if ((rc = ensure(&b.d_out, &b.out_cap, job.out_frame_bytes * cf))) break;
            if (job.desc->op == RIP_OP_FUSED &&
                !fused_single_kernel(job.W, job.H, job.desc->in_format, job.desc->ksize, job.desc->weights, b.d_in, b.d_out)) {
                size_t ws_need = 0;  // staged path only
                rip_fused_workspace_bytes(job.W, job.H, cf, job.desc->ksize, &ws_need);
                if ((rc = ensure(&b.d_ws, &b.ws_cap, ws_need))) break;
            }
            const bool timed = prof_ms && k == 0;
            const uint8_t *src = h_in + (size_t)(f0 + c0) * job.in_frame_bytes;
            uint8_t *dst = h_out + (size_t)(f0 + c0) * job.out_frame_bytes;
            if (timed) cudaEventRecord(dev.ev[0], b.stream);
            if ((rc = (int)cudaMemcpyAsync(b.d_in, src, job.in_frame_bytes * cf, cudaMemcpyHostToDevice, b.stream))) {
                rc = cuda_fail((cudaError_t)rc, "cudaMemcpyAsync(H2D)", __FILE__, __LINE__);
                break;
            }
            if (timed) cudaEventRecord(dev.ev[1], b.stream);
            if ((rc = enqueue_op(job, dev.device, b, cf))) break;
            if (timed) cudaEventRecord(dev.ev[2], b.stream);
            if ((rc = (int)cudaMemcpyAsync(dst, b.d_out, job.out_frame_bytes * cf, cudaMemcpyDeviceToHost, b.stream))) {
                rc = cuda_fail((cudaError_t)rc, "cudaMemcpyAsync(D2H)", __FILE__, __LINE__);
                break;
            }
            if (timed) cudaEventRecord(dev.ev[3], b.stream);
        }
        for (int i = 0; i < kSets; i++) {
            cudaError_t e = cudaStreamSynchronize(dev.set[i].stream);
            if (e != cudaSuccess && rc == RIP_OK) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
        }
        if (rc == RIP_OK && prof_ms) {
            float a = 0, bms = 0, c = 0;
            cudaEventElapsedTime(&a, dev.ev[0], dev.ev[1]);
            cudaEventElapsedTime(&bms, dev.ev[1], dev.ev[2]);
            cudaEventElapsedTime(&c, dev.ev[2], dev.ev[3]);
            prof_ms[0] = a; prof_ms[1] = bms; prof_ms[2] = c;
        }
    }
    if (rc != RIP_OK && err) snprintf(err, err_len, "%s", rip_last_error_string());
    return rc;
}

void fill_prof(uint64_t prof_ns[6], const double ms[3])
{
    // cumulative offsets from the start of the call, in ns: [w0, w1, k0, k1, r0, r1]
    const double w = ms[0] * 1e6, k = ms[1] * 1e6, r = ms[2] * 1e6;
    prof_ns[0] = 0;
    prof_ns[1] = (uint64_t)llround(w);
    prof_ns[2] = prof_ns[1];
    prof_ns[3] = (uint64_t)llround(w + k);
    prof_ns[4] = prof_ns[3];
    prof_ns[5] = (uint64_t)llround(w + k + r);
}

}  // namespace

extern "C" int rip_shard_frames(int n_frames, int n_parts, int index, int *first, int *count)
{
    if (!first || !count || n_frames < 0 || n_parts <= 0 || index < 0 || index >= n_parts)
        return fail(RIP_EINVAL, "rip_shard_frames: bad arguments (%d frames, part %d of %d)", n_frames, index, n_parts);
    const int f0 = (int)((long long)n_frames * index / n_parts), f1 = (int)((long long)n_frames * (index + 1) / n_parts);
    *first = f0;
    *count = f1 - f0;
    return RIP_OK;
}

extern "C" int rip_band_rows(int height, int n_parts, int index, int halo, int *in_row0, int *in_rows, int *out_row0, int *out_rows)
{
    if (!in_row0 || !in_rows || !out_row0 || !out_rows || height <= 0 || n_parts <= 0 || n_parts > height || index < 0 ||
        index >= n_parts || halo < 0)
        return fail(RIP_EINVAL, "rip_band_rows: bad arguments (height %d, part %d of %d, halo %d)", height, index, n_parts, halo);
    const int o0 = (int)((long long)height * index / n_parts), o1 = (int)((long long)height * (index + 1) / n_parts);
    const int i0 = max(0, o0 - halo), i1 = min(height, o1 + halo);
    *out_row0 = o0; *out_rows = o1 - o0; *in_row0 = i0; *in_rows = i1 - i0;
    return RIP_OK;
}

extern "C" int rip_process_host(rip_ctx *ctx, const rip_op_desc *desc, const uint8_t *h_in, uint8_t *h_out, int width,
                                int height, int n_frames, uint64_t prof_ns[6])
{
    if (!ctx) return fail(RIP_EINVAL, "rip_process_host: NULL context");
    if (!h_in || !h_out) return fail(RIP_EINVAL, "rip_process_host: NULL host buffer");
    if (width <= 0 || height <= 0 || n_frames <= 0)
        return fail(RIP_EINVAL, "rip_process_host: width, height and n_frames must be positive (got %d x %d x %d)", width, height, n_frames);
    Job job;
    job.desc = desc;
    job.W = width;
    job.H = height;
    if (int rc = validate_desc(desc, &job.cn)) return rc;
    if (desc->in_format == RIP_FMT_NV12 && (height & 1)) return fail(RIP_EINVAL, "NV12 frames need an even height (got %d)", height);
    job.in_frame_bytes = frame_bytes_of(desc->in_format, width, height);
    if (int rc = rip_out_bytes_per_frame(desc, width, height, &job.out_frame_bytes)) return rc;

    const int nd = (int)ctx->devs.size();
    const int used = n_frames < nd ? n_frames : nd;
    double prof_ms[3] = {0, 0, 0};
    if (used == 1) {
        int rc = run_device(ctx->devs[0], job, h_in, h_out, 0, n_frames, prof_ns ? prof_ms : nullptr, nullptr, 0);
        if (rc == RIP_OK && prof_ns) fill_prof(prof_ns, prof_ms);
        return rc;
    }
    // contiguous blocks of frames per device; one host thread per device, no inter-device traffic
    std::vector<std::thread> th;
    std::vector<int> rcs(used, RIP_OK);
    std::vector<std::string> errs(used, std::string(512, '\0'));
    for (int i = 0; i < used; i++) {
        int f0 = 0, fc = 0;
        rip_shard_frames(n_frames, used, i, &f0, &fc);
        const int f1 = f0 + fc;
        th.emplace_back([&, i, f0, f1]() {
            rcs[i] = run_device(ctx->devs[i], job, h_in, h_out, f0, f1, (i == 0 && prof_ns) ? prof_ms : nullptr, &errs[i][0], errs[i].size());
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < used; i++)
        if (rcs[i] != RIP_OK) return fail(rcs[i], "device %d: %s", ctx->devs[i].device, errs[i].c_str());
    if (prof_ns) fill_prof(prof_ns, prof_ms);
    return RIP_OK;
}

extern "C" int rip_process_host_banded(rip_ctx *ctx, const rip_op_desc *desc, const uint8_t *h_in, uint8_t *h_out,
                                       int width, int height, uint64_t prof_ns[6])
{
    if (!ctx) return fail(RIP_EINVAL, "rip_process_host_banded: NULL context");
    if (!h_in || !h_out) return fail(RIP_EINVAL, "rip_process_host_banded: NULL host buffer");
    if (width <= 0 || height <= 0) return fail(RIP_EINVAL, "rip_process_host_banded: bad shape %d x %d", width, height);
    int cn = 0;
    if (int rc = validate_desc(desc, &cn)) return rc;
    if (desc->op != RIP_OP_FUSED && desc->op != RIP_OP_EDGE)
        return fail(RIP_EUNSUPPORTED, "row-band mode supports FUSED and EDGE only");
    const int halo = desc->op == RIP_OP_FUSED ? desc->ksize / 2 + 1 : 1;
    int nd = (int)ctx->devs.size();
    if (nd > height) nd = height;
    std::vector<std::thread> th;
    std::vector<int> rcs(nd, RIP_OK);
    std::vector<std::string> errs(nd, std::string(512, '\0'));
    double prof_ms[3] = {0, 0, 0};
    const size_t row_in = (size_t)width * cn;
    for (int i = 0; i < nd; i++) {
        th.emplace_back([&, i]() {
            DevState &dev = ctx->devs[i];
            int rc = RIP_OK;
            {
                DeviceGuard g(dev.device);
                int o0 = 0, on = 0, i0 = 0, in = 0;
                rip_band_rows(height, nd, i, halo, &i0, &in, &o0, &on);
                const int o1 = o0 + on, i1 = i0 + in;
                BufSet &b = dev.set[0];
                do {
                    if ((rc = ensure(&b.d_in, &b.in_cap, row_in * (i1 - i0)))) break;
                    if ((rc = ensure(&b.d_out, &b.out_cap, (size_t)width * (o1 - o0)))) break;
                    if (desc->op == RIP_OP_FUSED &&
                        !fused_single_kernel(width, height, desc->in_format, desc->ksize, desc->weights, b.d_in, b.d_out)) {
                        size_t ws_need = 0;  // staged path only
                        rip_fused_workspace_bytes(width, i1 - i0, 1, desc->ksize, &ws_need);
                        if ((rc = ensure(&b.d_ws, &b.ws_cap, ws_need))) break;
                    }
                    const bool timed = prof_ns && i == 0;
                    if (timed) cudaEventRecord(dev.ev[0], b.stream);
                    cudaError_t e = cudaMemcpyAsync(b.d_in, h_in + row_in * i0, row_in * (i1 - i0), cudaMemcpyHostToDevice, b.stream);
                    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(H2D)", __FILE__, __LINE__); break; }
                    if (timed) cudaEventRecord(dev.ev[1], b.stream);
                    if (desc->op == RIP_OP_FUSED) {
                        rc = rip_fused(dev.device, b.stream, (const uint8_t *)b.d_in, (uint8_t *)b.d_out, width, height, 1,
                                       desc->in_format, desc->ksize, desc->weights, i0, i1 - i0, o0, o1 - o0, b.d_ws, b.ws_cap);
                    } else if (fused_supported(width, height, desc->in_format, 0, (const uint8_t *)b.d_in, (uint8_t *)b.d_out)) {
                        rc = launch_fused(b.stream, (const uint8_t *)b.d_in, (uint8_t *)b.d_out, width, height, 1, desc->in_format, false,
                                          nullptr, i0, i1 - i0, o0, o1 - o0, dev.device);
                    } else {
                        rc = launch_sobel(b.stream, (const uint8_t *)b.d_in, (uint8_t *)b.d_out, width, height, 1,
                                          desc->in_format == RIP_FMT_NV12 ? RIP_FMT_GRAY8 : desc->in_format, i0,
                                          i1 - i0, o0, o1 - o0);
                    }
                    if (rc) break;
                    if (timed) cudaEventRecord(dev.ev[2], b.stream);
                    e = cudaMemcpyAsync(h_out + (size_t)width * o0, b.d_out, (size_t)width * (o1 - o0), cudaMemcpyDeviceToHost, b.stream);
                    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync(D2H)", __FILE__, __LINE__); break; }
                    if (timed) cudaEventRecord(dev.ev[3], b.stream);
                    e = cudaStreamSynchronize(b.stream);
                    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__); break; }
                    if (timed) {
                        float a = 0, k = 0, c = 0;
                        cudaEventElapsedTime(&a, dev.ev[0], dev.ev[1]);
                        cudaEventElapsedTime(&k, dev.ev[1], dev.ev[2]);
                        cudaEventElapsedTime(&c, dev.ev[2], dev.ev[3]);
                        prof_ms[0] = a; prof_ms[1] = k; prof_ms[2] = c;
                    }
                } while (0);
            }
            if (rc != RIP_OK) snprintf(&errs[i][0], errs[i].size(), "%s", rip_last_error_string());
            rcs[i] = rc;
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < nd; i++)
        if (rcs[i] != RIP_OK) return fail(rcs[i], "device %d: %s", ctx->devs[i].device, errs[i].c_str());
    if (prof_ns) fill_prof(prof_ns, prof_ms);
    return RIP_OK;
}
