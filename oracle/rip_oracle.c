/*
 * rip_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See rip_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  The reference's
 * CMake sets no -march / -ffast-math, so on x86-64 its CPU paths are strict IEEE SSE2 with no
 * FMA; -ffp-contract=off keeps this restatement the same on any host.
 */
#include "rip_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int pick_threads(int threads)
{
#ifdef _OPENMP
    if (threads <= 0) return omp_get_max_threads();
    return threads;
#else
    (void)threads;
    return 1;
#endif
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cv::BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba); a length-1 axis maps everything to 0,
 * as cv::borderInterpolate does. */
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

/* Comparator.cpp:36-41 -- int r,g,b promoted to double, left-to-right sum, C truncation. */
int rip_oracle_gray(const uint8_t *src, int w, int h, int cn, int order, uint8_t *dst, int threads)
{
    if (!src || !dst || w <= 0 || h <= 0 || (cn != 3 && cn != 4)) return -1;
    const int ri = (order == RIP_ORACLE_BGR) ? 2 : 0;
    const int bi = (order == RIP_ORACLE_BGR) ? 0 : 2;
    const int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t *row = src + (size_t)y * w * cn;
        uint8_t *out = dst + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int r = row[x * cn + ri];
            int g = row[x * cn + 1];
            int b = row[x * cn + bi];
            out[x] = (uint8_t)(0.299 * r + 0.587 * g + 0.114 * b);
        }
    }
    return 0;
}

/* GaussianBlur/src/Controller.cpp:342-362.  Typing matters: `-(x*x+y*y)` is int, `2*sigma*sigma`
 * is float, so the exponent argument is a float quotient; unqualified exp() binds to
 * ::exp(double) under g++/libstdc++; `2*M_PI*sigma*sigma` is double; the quotient is rounded to
 * float on assignment; the running sum and the final divide are float. */
int rip_oracle_gauss_weights(int ksize, float sigma, float *weights)
{
    if (!weights || ksize <= 0 || (ksize & 1) == 0) return -1;
    const int half = ksize / 2;
    float sum = 0.0f;
    for (int y = -half; y <= half; y++) {
        for (int x = -half; x <= half; x++) {
            float arg = (float)(-(x * x + y * y)) / (2 * sigma * sigma);
            float value = (float)(exp((double)arg) / (2 * M_PI * sigma * sigma));
            weights[(y + half) * ksize + (x + half)] = value;
            sum += value;
        }
    }
    for (int i = 0; i < ksize * ksize; i++) weights[i] /= sum;
    return 0;
}

/* GaussianBlur.cpp:234-258 -- one float accumulator per channel starting at 0.0f, taps visited
 * ky-major / kx-minor, coordinates clamped, `sum += px * weight` (u8 -> int -> float, product
 * rounded, then sum rounded), static_cast<uchar>(clamp(sum, 0, 255)). */
int rip_oracle_blur(const uint8_t *src, int w, int h, int cn, int ksize, const float *weights,
                    uint8_t *dst, int threads)
{
    if (!src || !dst || !weights || w <= 0 || h <= 0 || cn <= 0 || cn > 4 || ksize <= 0 ||
        (ksize & 1) == 0)
        return -1;
    const int half = ksize / 2;
    const int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int ky = -half; ky <= half; ky++) {
                const int ny = clampi(y + ky, 0, h - 1);
                for (int kx = -half; kx <= half; kx++) {
                    const int nx = clampi(x + kx, 0, w - 1);
                    const uint8_t *px = src + ((size_t)ny * w + nx) * cn;
                    const float weight = weights[(ky + half) * ksize + (kx + half)];
                    for (int c = 0; c < cn; c++) {
                        float prod = (float)(int)px[c] * weight;
                        sum[c] = sum[c] + prod;
                    }
                }
            }
            uint8_t *out = dst + ((size_t)y * w + x) * cn;
            for (int c = 0; c < cn; c++) {
                float v = sum[c];
                if (v < 0.0f) v = 0.0f;
                if (v > 255.0f) v = 255.0f;
                out[c] = (uint8_t)v;
            }
        }
    }
    return 0;
}

/* EdgeDetection.cpp:219-240.  filter2D is a correlation with the float kernels
 * {-1,0,1;-2,0,2;-1,0,1} / {-1,-2,-1;0,0,0;1,2,1}, default border BORDER_REFLECT_101; every
 * partial value is a small integer, so the float result equals the integer one.
 * cv::magnitude = sqrtf(gx*gx + gy*gy) with gx*gx+gy*gy < 2^24 (exact in float) and a correctly
 * rounded sqrt; convertTo(CV_8UC1) = saturate_cast<uchar>(cvRound(v)) i.e. round-half-even
 * (lrintf in the default rounding mode) then clamp to 255. */
int rip_oracle_sobel(const uint8_t *gray, int w, int h, uint8_t *dst, int threads)
{
    if (!gray || !dst || w <= 0 || h <= 0) return -1;
    const int nt = pick_threads(threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = gray + (size_t)reflect101(y - 1, h) * w;
        const uint8_t *r1 = gray + (size_t)y * w;
        const uint8_t *r2 = gray + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
            const int gx = -(int)r0[xl] + (int)r0[xr] - 2 * (int)r1[xl] + 2 * (int)r1[xr] -
                           (int)r2[xl] + (int)r2[xr];
            const int gy = -(int)r0[xl] - 2 * (int)r0[x] - (int)r0[xr] + (int)r2[xl] +
                           2 * (int)r2[x] + (int)r2[xr];
            const float mag = sqrtf((float)(gx * gx + gy * gy));
            long m = lrintf(mag);
            dst[(size_t)y * w + x] = (uint8_t)(m > 255 ? 255 : m);
        }
    }
    return 0;
}

int rip_oracle_fused(const uint8_t *src, int w, int h, int cn, int order, int ksize,
                     const float *weights, uint8_t *dst, uint8_t *scratch, int threads)
{
    if (w <= 0 || h <= 0) return -1;
    uint8_t *tmp = scratch ? scratch : (uint8_t *)malloc((size_t)2 * w * h);
    if (!tmp) return -2;
    uint8_t *g = tmp, *b = tmp + (size_t)w * h;
    int rc = rip_oracle_gray(src, w, h, cn, order, g, threads);
    if (!rc) rc = rip_oracle_blur(g, w, h, 1, ksize, weights, b, threads);
    if (!rc) rc = rip_oracle_sobel(b, w, h, dst, threads);
    if (!scratch) free(tmp);
    return rc;
}

double rip_oracle_mae_ch0(const uint8_t *a, const uint8_t *b, int w, int h, int cn)
{
    const size_t n = (size_t)w * h;
    unsigned long long acc = 0;
    for (size_t i = 0; i < n; i++) {
        int d = (int)a[i * cn] - (int)b[i * cn];
        acc += (unsigned)(d < 0 ? -d : d);
    }
    return n ? (double)acc / (double)n : 0.0;
}

int rip_oracle_max_abs(const uint8_t *a, const uint8_t *b, long n, long *n_diff)
{
    int mx = 0;
    long nd = 0;
    for (long i = 0; i < n; i++) {
        int d = (int)a[i] - (int)b[i];
        if (d < 0) d = -d;
        if (d) nd++;
        if (d > mx) mx = d;
    }
    if (n_diff) *n_diff = nd;
    return mx;
}

/* ------------------------------------------------------------------------------------------
 * OpenCL buffer-path semantics (float32, unfused).  Not part of the hot path being rebuilt:
 * they exist so the published Error_MAE values (CPU path vs OpenCL buffer path) can be replayed
 * as known-answer tests for the CPU-path restatements above.
 * ------------------------------------------------------------------------------------------ */
static inline float ocl_gray(const uint8_t *p)
{
    float s = 0.299f * (float)p[0];
    s = s + 0.587f * (float)p[1];
    s = s + 0.114f * (float)p[2];
    return s / 255.0f;
}

int rip_oracle_ocl_gray_rgba(const uint8_t *rgba, int w, int h, uint8_t *dst)
{
    if (!rgba || !dst || w <= 0 || h <= 0) return -1;
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        uint8_t g = (uint8_t)(ocl_gray(rgba + 4 * i) * 255.0f);
        dst[4 * i] = dst[4 * i + 1] = dst[4 * i + 2] = g;
        dst[4 * i + 3] = (uint8_t)(1.0f * 255.0f);
    }
    return 0;
}

int rip_oracle_ocl_sobel_rgba(const uint8_t *rgba, int w, int h, uint8_t *dst)
{
    static const int sx[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
    static const int sy[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
    if (!rgba || !dst || w <= 0 || h <= 0) return -1;
    memset(dst, 0, (size_t)w * h); /* fresh clCreateBuffer read back as 0 on the authors' device */
    for (int y = 1; y < h - 1; y++) {
        for (int x = 1; x < w - 1; x++) {
            float gx = 0.0f, gy = 0.0f;
            for (int ky = -1; ky <= 1; ky++) {
                for (int kx = -1; kx <= 1; kx++) {
                    float gray = ocl_gray(rgba + 4 * ((size_t)(y + ky) * w + (x + kx)));
                    gx = gx + gray * (float)sx[ky + 1][kx + 1];
                    gy = gy + gray * (float)sy[ky + 1][kx + 1];
                }
            }
            float xx = gx * gx, yy = gy * gy;
            float mag = sqrtf(xx + yy);
            if (mag < 0.0f) mag = 0.0f;
            if (mag > 1.0f) mag = 1.0f;
            dst[(size_t)y * w + x] = (uint8_t)(mag * 255.0f);
        }
    }
    return 0;
}

int rip_oracle_ocl_blur_rgba(const uint8_t *rgba, int w, int h, int ksize, const float *weights,
                             uint8_t *dst)
{
    if (!rgba || !dst || !weights || w <= 0 || h <= 0 || (ksize & 1) == 0) return -1;
    const int half = ksize / 2;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            float total = 0.0f;
            for (int ky = -half; ky <= half; ky++) {
                for (int kx = -half; kx <= half; kx++) {
                    const int nx = clampi(x + kx, 0, w - 1), ny = clampi(y + ky, 0, h - 1);
                    const uint8_t *px = rgba + 4 * ((size_t)ny * w + nx);
                    const float weight = weights[(ky + half) * ksize + (kx + half)];
                    for (int c = 0; c < 4; c++) {
                        float prod = (float)px[c] * weight;
                        sum[c] = sum[c] + prod;
                    }
                    total = total + weight;
                }
            }
            uint8_t *out = dst + 4 * ((size_t)y * w + x);
            for (int c = 0; c < 4; c++) out[c] = (uint8_t)(sum[c] / total);
        }
    }
    return 0;
}
