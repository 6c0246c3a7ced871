// rip_compat.h -- the OpenCL names the reference's class signatures mention, re-pointed at the CUDA
// C ABI (include/rip_cuda.h), so call sites written against the reference
// (RealtimeImageProcessing.cpp:53,79; grayscale.cpp:150 ...) compile unchanged without libOpenCL.
//
//   cl_platform_id   -> the one "NVIDIA CUDA" platform
//   cl_device_id     -> a CUDA device ordinal
//   cl_context       -> rip_ctx over a device set (cached device buffers, chunk streams)
//   cl_command_queue -> a device of that context + its single-device pipeline
//   cl_program       -> rip_module  (kernel image; the .cl file name selects the variant)
//   cl_kernel        -> rip_kernel  (bound entry point)
#pragma once

#include <cstdint>

#include "rip_cuda.h"

#ifndef CL_SUCCESS
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_uint cl_bool;
typedef cl_uint cl_platform_info;
#define CL_SUCCESS 0
#define CL_DEVICE_NOT_FOUND (-1)
#define CL_INVALID_VALUE (-30)
#define CL_TRUE 1u
#define CL_FALSE 0u
#define CL_PLATFORM_PROFILE 0x0900
#define CL_PLATFORM_VERSION 0x0901
#define CL_PLATFORM_NAME 0x0902
#define CL_PLATFORM_VENDOR 0x0903

struct rip_platform_rec { int id; };
struct rip_device_rec { int ordinal; };
struct rip_context_rec {        // owns the multi-device pipeline
    rip_ctx *ctx;
    int n_devices;
    int devices[16];
};
struct rip_queue_rec {          // one device of a context, with its own single-device pipeline
    rip_context_rec *context;
    int ordinal;
    rip_ctx *ctx;
};

typedef rip_platform_rec *cl_platform_id;
typedef rip_device_rec *cl_device_id;
typedef rip_context_rec *cl_context;
typedef rip_queue_rec *cl_command_queue;
typedef rip_module *cl_program;
typedef rip_kernel *cl_kernel;
typedef void *cl_sampler;
typedef void *cl_mem;
typedef rip_event *cl_event;
#endif
