// InfoPlatform.hpp -- same shape as the reference's clGetPlatformInfo pretty-printer
// (include/InfoPlatform.hpp:7-25), reporting the CUDA platform instead.
#pragma once

#include <string>

#include "rip_compat.h"

class InfoPlatform {
public:
    explicit InfoPlatform(cl_platform_id id);
    void DisplaySinglePlatformInfo(cl_platform_id id, cl_platform_info name, std::string str);
    void Display();
    std::string GetPlatformInfo(cl_platform_info name);

private:
    std::string retrievePlatformInfo(cl_platform_id id, cl_platform_info name, std::string str);
    void setPlatformInfo(cl_platform_info name, std::string info);

    std::string m_profile, m_name, m_version, m_vendor;
};
