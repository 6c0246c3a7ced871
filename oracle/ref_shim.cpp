// ref_shim.cpp -- ORACLE / TEST INFRASTRUCTURE ONLY.  A C entry point over the REFERENCE's own Gaussian weight
// generator, Controller::_GenerateGausianKernel -> _GenerateGaussianKernelBuffers
// (/root/reference/src/GaussianBlur/src/Controller.cpp:342-362, public in src/GaussianBlur/include/Controller.hpp:28).
// The reference sources are compiled where they lie (oracle/Makefile target _ref/librip_ref_weights.so); nothing of
// them is copied into this repository.  Used by tests/ to pin both the oracle's restatement and the product's
// rip_gauss_weights bit for bit, and by tools/make_golden.py to write the committed fixture
// tests/golden/ref_weights.json (the GPU box has no /root/reference).
#include <CL/cl.h>
#include <Controller.hpp>

extern "C" int rip_ref_gauss_weights(int ksize, float sigma, float *out)
{
    Controller c;
    c.SetImageSupport(CL_FALSE);   // the buffer flavour, as every published result (BYPASS_IMAGE_SUPPORT = true)
    const std::vector<float> k = c._GenerateGausianKernel(ksize, sigma);
    if ((int)k.size() != ksize * ksize) return -1;
    for (size_t i = 0; i < k.size(); i++) out[i] = k[i];
    return 0;
}
