#include "InfoPlatform.hpp"

#include <cuda_runtime_api.h>

#include <iostream>
#include <sstream>

InfoPlatform::InfoPlatform(cl_platform_id id)
{
    for (cl_platform_info what : {CL_PLATFORM_PROFILE, CL_PLATFORM_NAME, CL_PLATFORM_VERSION, CL_PLATFORM_VENDOR})
        setPlatformInfo(what, retrievePlatformInfo(id, what, ""));
}

std::string InfoPlatform::retrievePlatformInfo(cl_platform_id, cl_platform_info name, std::string)
{
    switch (name) {
    case CL_PLATFORM_PROFILE: return "FULL_PROFILE (CUDA, sm_100a kernels compiled into librip_cuda)";
    case CL_PLATFORM_NAME: return "NVIDIA CUDA";
    case CL_PLATFORM_VENDOR: return "NVIDIA Corporation";
    case CL_PLATFORM_VERSION: {
        int rt = 0, drv = 0;
        cudaRuntimeGetVersion(&rt);
        cudaDriverGetVersion(&drv);
        std::ostringstream s;
        s << "CUDA runtime " << rt / 1000 << "." << (rt % 1000) / 10 << ", driver " << drv / 1000 << "." << (drv % 1000) / 10
          << ", rip ABI " << rip_abi_version();
        return s.str();
    }
    default: return "";
    }
}

void InfoPlatform::setPlatformInfo(cl_platform_info name, std::string info)
{
    if (name == CL_PLATFORM_PROFILE) m_profile = info;
    else if (name == CL_PLATFORM_NAME) m_name = info;
    else if (name == CL_PLATFORM_VERSION) m_version = info;
    else if (name == CL_PLATFORM_VENDOR) m_vendor = info;
}

std::string InfoPlatform::GetPlatformInfo(cl_platform_info name)
{
    if (name == CL_PLATFORM_PROFILE) return m_profile;
    if (name == CL_PLATFORM_NAME) return m_name;
    if (name == CL_PLATFORM_VERSION) return m_version;
    if (name == CL_PLATFORM_VENDOR) return m_vendor;
    return "";
}

void InfoPlatform::DisplaySinglePlatformInfo(cl_platform_id, cl_platform_info name, std::string str)
{
    std::cout << "\t" << str << ":\t" << GetPlatformInfo(name) << std::endl;
}

void InfoPlatform::Display()
{
    std::cout << "Platform information" << std::endl;
    DisplaySinglePlatformInfo(nullptr, CL_PLATFORM_PROFILE, "CL_PLATFORM_PROFILE");
    DisplaySinglePlatformInfo(nullptr, CL_PLATFORM_NAME, "CL_PLATFORM_NAME");
    DisplaySinglePlatformInfo(nullptr, CL_PLATFORM_VERSION, "CL_PLATFORM_VERSION");
    DisplaySinglePlatformInfo(nullptr, CL_PLATFORM_VENDOR, "CL_PLATFORM_VENDOR");
}
