// rip_blur_sep.cu -- the stand-alone KxK Gaussian (replaces kernel/gaussian_base.cl; results defined by the
// reference's CPU path, GaussianBlur.cpp:231-261) at separable cost, bit-exact.
//
// The reference's value of a pixel/channel is trunc(clamp(S_ref, 0, 255)) with S_ref the sequential, unfused,
// ky-major fp32 sum of the K*K rounded products.  Its weights (Controller.cpp:342-362 of src/GaussianBlur)
// are a normalised Gaussian, i.e. separable up to float rounding.  This kernel evaluates a separable sum
//   S~ = sum_ky g[ky] * ( sum_kx g[kx] * p[y+ky][x+kx] ),     g[k] = sqrt(w[k][k]),
// (2K FMAs instead of K*K multiply-adds) and proves, per launch, a bound `band` on |S~ - S_ref|
// (plan_sep_blur below).  S~ is formed on top of a bias of 256, which puts floor(S~) in mantissa bits 15..22
// and the fraction in bits 0..14; a pixel whose fraction is farther than `band` from both 0 and 1 has
// floor(S~) == floor(S_ref).  The others (~0.1 % per channel for 5x5, a few % for 17x17) replay the
// reference's exact sequence from the same shared-memory tile.  Weights the bound cannot cover (not a
// symmetric separable kernel) run the plain exact kernel in rip_stages.cu.
//
// Layout: a block of 256 threads produces a 32 x 32 tile.  The tile with its halo is loaded once with
// clamp-to-edge coordinates (GaussianBlur.cpp:240-241), converted u8 -> fp32 once per loaded pixel (PRMT +
// FADD: no I2F), and kept in shared memory as float4 (RGBA) or float (gray); the horizontal pass writes a
// second shared array, the vertical pass reads it.  All shared-memory accesses of a warp are consecutive
// 16-byte (or 4-byte) words: conflict-free.  HBM traffic: 8 algorithmic bytes per RGBA pixel; the halo
// re-reads ((32+2h)^2 / 32^2) are L2 hits.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "rip_common.cuh"
#include "rip_internal.h"

namespace rip {

namespace {

constexpr int SEP_TW = 32, SEP_TH = 32, SEP_THREADS = 256;
constexpr int kSepFracBits = 15;
constexpr float kSepBias = 256.0f;

struct SepParams {
    const uint8_t *src;
    uint8_t *dst;
    int W, H, src_row0, src_rows, out_row0, out_rows, ksize;
    uint32_t zoff, zthr;        // replay iff ((bits << 17) + zoff) < zthr
    float g[RIP_MAX_KSIZE];     // separable taps
    float sgv[5], sgh[5], sbias;   // streaming 5x5 kernel: vertical taps * 2^75, horizontal taps * 2^74, bias 256 + a * 2^-15
    float sg1[RIP_MAX_KSIZE], sg2[RIP_MAX_KSIZE];   // streaming KxK kernel: first-pass taps * 2^75, second-pass taps * 2^74
    uint32_t f255, a255;        // streaming 5x5 kernel: the bit pattern of the fast sum of an all-255 window (0: test disabled) and flat[255] << 24
    float rw[25];               // streaming 5x5 kernel: the reference's 25 weights * 2^100 (replay on integer bit patterns, bs_replay1)
    uint8_t flat[256];          // flat[v] = the reference's result for a CONSTANT KxK window of value v (its own sequence, host-evaluated)
    unsigned long long *slow_counter;   // optional statistics (NULL in production)
};

template <int CN> struct SepPx;
template <> struct SepPx<4> {
    typedef float4 T;
    typedef uint32_t raw_t;
    static __device__ __forceinline__ uint32_t load_raw(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
    static __device__ __forceinline__ T load(const uint8_t *p) { return cvt(load_raw(p)); }
    static __device__ __forceinline__ T cvt(uint32_t v)
    {
        // 0x4B000000 | byte is the float 8388608 + byte
        const float m = 8388608.0f;
        return make_float4(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7441)) - m,
                           __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7442)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7443)) - m);
    }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T splat(float b) { return make_float4(b, b, b, b); }
    static __device__ __forceinline__ T fma(float g, T v, T a) { return make_float4(fmaf(g, v.x, a.x), fmaf(g, v.y, a.y), fmaf(g, v.z, a.z), fmaf(g, v.w, a.w)); }
    static __device__ __forceinline__ T mul(float g, T v) { return make_float4(g * v.x, g * v.y, g * v.z, g * v.w); }
    // the reference's step: acc = fl(acc + fl(v * w)), no FMA
    static __device__ __forceinline__ T ref_step(T a, T v, float w)
    {
        return make_float4(__fadd_rn(a.x, __fmul_rn(v.x, w)), __fadd_rn(a.y, __fmul_rn(v.y, w)), __fadd_rn(a.z, __fmul_rn(v.z, w)),
                           __fadd_rn(a.w, __fmul_rn(v.w, w)));
    }
    static __device__ __forceinline__ uint32_t zmin(T f, uint32_t zoff)
    {
        const uint32_t a = (__float_as_uint(f.x) << (32 - kSepFracBits)) + zoff, b = (__float_as_uint(f.y) << (32 - kSepFracBits)) + zoff;
        const uint32_t c = (__float_as_uint(f.z) << (32 - kSepFracBits)) + zoff, d = (__float_as_uint(f.w) << (32 - kSepFracBits)) + zoff;
        return min(__vimin3_u32(a, b, c), d);
    }
    // the same with the channels in `cm` (0xff in the byte of a channel that is constant over the whole tile) left out
    // ... and a channel whose fast sum is below the band's width (bits < zlow = the bits of 256 + a ulps) left out as well: its true sum is
    // >= 0 and below 1, the result is 0 either way (black sky: every all-zero window of a tile that is not constant sat on the replay list)
    static __device__ __forceinline__ uint32_t zmin_masked(T f, uint32_t zoff, uint32_t cm, uint32_t zlow)
    {
        const uint32_t a = ((__float_as_uint(f.x) << (32 - kSepFracBits)) + zoff) | (uint32_t)-(int)(cm & 1u) | (__float_as_uint(f.x) < zlow ? ~0u : 0u);
        const uint32_t b = ((__float_as_uint(f.y) << (32 - kSepFracBits)) + zoff) | (uint32_t)-(int)((cm >> 8) & 1u) | (__float_as_uint(f.y) < zlow ? ~0u : 0u);
        const uint32_t c = ((__float_as_uint(f.z) << (32 - kSepFracBits)) + zoff) | (uint32_t)-(int)((cm >> 16) & 1u) | (__float_as_uint(f.z) < zlow ? ~0u : 0u);
        const uint32_t d = ((__float_as_uint(f.w) << (32 - kSepFracBits)) + zoff) | (uint32_t)-(int)((cm >> 24) & 1u) | (__float_as_uint(f.w) < zlow ? ~0u : 0u);
        return min(__vimin3_u32(a, b, c), d);
    }
    static __device__ __forceinline__ uint32_t pack_fast(T f)   // floor(S~) of each channel: mantissa bits 15..22
    {
        return ((__float_as_uint(f.x) >> kSepFracBits) & 0xffu) | (((__float_as_uint(f.y) >> kSepFracBits) & 0xffu) << 8) |
               (((__float_as_uint(f.z) >> kSepFracBits) & 0xffu) << 16) | (((__float_as_uint(f.w) >> kSepFracBits) & 0xffu) << 24);
    }
    static __device__ __forceinline__ uint32_t pack_exact(T a)  // (uchar)clamp(sum, 0, 255), GaussianBlur.cpp:255-258
    {
        return (uint32_t)__float2int_rz(fminf(fmaxf(a.x, 0.f), 255.f)) | ((uint32_t)__float2int_rz(fminf(fmaxf(a.y, 0.f), 255.f)) << 8) |
               ((uint32_t)__float2int_rz(fminf(fmaxf(a.z, 0.f), 255.f)) << 16) | ((uint32_t)__float2int_rz(fminf(fmaxf(a.w, 0.f), 255.f)) << 24);
    }
    static __device__ __forceinline__ void store(uint8_t *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }
};
template <> struct SepPx<1> {
    typedef float T;
    typedef uint8_t raw_t;
    static __device__ __forceinline__ uint32_t load_raw(const uint8_t *p) { return (uint32_t)__ldg(p); }
    static __device__ __forceinline__ T cvt(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }
    static __device__ __forceinline__ T load(const uint8_t *p) { return cvt(load_raw(p)); }
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T splat(float b) { return b; }
    static __device__ __forceinline__ T fma(float g, T v, T a) { return fmaf(g, v, a); }
    static __device__ __forceinline__ T mul(float g, T v) { return g * v; }
    static __device__ __forceinline__ T ref_step(T a, T v, float w) { return __fadd_rn(a, __fmul_rn(v, w)); }
    static __device__ __forceinline__ uint32_t zmin(T f, uint32_t zoff) { return (__float_as_uint(f) << (32 - kSepFracBits)) + zoff; }
    static __device__ __forceinline__ uint32_t zmin_masked(T f, uint32_t zoff, uint32_t cm, uint32_t zlow)
    {
        return zmin(f, zoff) | (uint32_t)-(int)(cm & 1u) | (__float_as_uint(f) < zlow ? ~0u : 0u);
    }
    static __device__ __forceinline__ uint32_t pack_fast(T f) { return (__float_as_uint(f) >> kSepFracBits) & 0xffu; }
    static __device__ __forceinline__ uint32_t pack_exact(T a) { return (uint32_t)__float2int_rz(fminf(fmaxf(a, 0.f), 255.f)); }
    static __device__ __forceinline__ void store(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }
};

// KT = compile-time kernel size (unrolled tap loops), or 0 = the run-time size in p.ksize
template <int CN, int KT>
__global__ void __launch_bounds__(SEP_THREADS)
blur_sep_kernel(const __grid_constant__ SepParams p, const __grid_constant__ Weights wts)
{
    typedef SepPx<CN> P;
    typedef typename P::T T;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int K = KT ? KT : p.ksize;
    const int half = K >> 1;
    const int tws = (SEP_TW + 2 * half + 3) & ~3, th = SEP_TH + 2 * half;   // row pitch: a multiple of 4 elements
    const int tq = tws >> 2;
    T *tile = reinterpret_cast<T *>(smem_raw);   // [th][tws]  the input tile with halo, as fp32, columns de-interleaved (col())
    T *hbuf = tile + th * tws;                   // [th][SEP_TW] after the horizontal pass
    // Column c of a tile row lives at element (c % 4) * tq + c / 4: a thread of the horizontal pass produces FOUR adjacent
    // outputs (columns 4q .. 4q+3) from K + 3 loads, and with this layout the eight threads of a quarter warp read eight
    // consecutive 16-byte elements at every step.  (Round 1 produced one output per thread from K loads: the kernel was
    // bound by shared-memory bandwidth -- 2.5 K 16-byte loads per pixel -- not by arithmetic: 17x17 on 16 1080p frames
    // 2505 us, then 1244 us with the block-wide replay alone.)
    auto col = [tq](int c) { return (c & 3) * tq + (c >> 2); };

    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * SEP_TW, y0 = p.out_row0 + blockIdx.y * SEP_TH;
    const uint8_t *fsrc = p.src + (size_t)blockIdx.z * p.src_rows * p.W * CN;
    uint8_t *fdst = p.dst + (size_t)blockIdx.z * p.out_rows * p.W * CN;

    // ---- 1. tile with halo, clamp-to-edge; rows below the last row this band's outputs need are never
    //         consumed and may lie outside src: zero.  On the way: which channels are CONSTANT over the whole tile?
    //         Every window of such a channel is constant, so its result is p.flat[value] whatever the fast sum says -- and
    //         the fast sum of a constant window sits on an integer, inside the guard band, for EVERY pixel: the alpha
    //         channel of any real RGBA frame (255 throughout) and black sky sent every pixel of the frame through the
    //         K*K replay (5x5 on 16 1080p frames: 977 us instead of 219 us; 17x17: 9.2 ms instead of 1.0 ms).
    const int y_last = min(p.out_row0 + p.out_rows, p.H) - 1 + half;
    const uint32_t ref = P::load_raw(fsrc + ((size_t)(clampi(y0 - half, 0, p.H - 1) - p.src_row0) * p.W + clampi(x0 - half, 0, p.W - 1)) * CN);
    uint32_t dacc = 0;
    for (int ty = wrp; ty < th; ty += SEP_THREADS / 32) {
        const int gy_raw = y0 - half + ty;
        const bool row_ok = gy_raw <= y_last;
        const uint8_t *row = fsrc + (size_t)(clampi(gy_raw, 0, p.H - 1) - p.src_row0) * p.W * CN;
        for (int tx = lane; tx < SEP_TW + 2 * half; tx += 32) {
            const uint32_t raw = row_ok ? P::load_raw(row + (size_t)clampi(x0 - half + tx, 0, p.W - 1) * CN) : ref;
            dacc |= raw ^ ref;
            tile[ty * tws + col(tx)] = row_ok ? P::cvt(raw) : P::zero();
        }
    }
    __shared__ uint32_t s_diff[SEP_THREADS / 32];
    dacc = __reduce_or_sync(0xffffffffu, dacc);
    if (lane == 0) s_diff[wrp] = dacc;
    __syncthreads();
    uint32_t cm = 0, cexact = 0;   // 0xff in the byte of every tile-constant channel; those channels' exact results
    {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < SEP_THREADS / 32; i++) d |= s_diff[i];
#pragma unroll
        for (int c = 0; c < CN; c++)
            if (((d >> (8 * c)) & 0xffu) == 0u) {
                cm |= 0xffu << (8 * c);
                cexact |= (uint32_t)p.flat[(ref >> (8 * c)) & 0xffu] << (8 * c);
            }
    }

    // ---- 2. horizontal pass: hbuf[ty][x] = sum_k g[k] * tile[ty][x + k]   (one FMA chain per output, k ascending)
    if (KT) {
        const int q = lane & 7, rsub = lane >> 3;   // outputs 4q .. 4q+3 of row ty; four rows per warp instruction
        for (int ty = wrp * 4 + rsub; ty < th; ty += SEP_THREADS / 8) {
            const T *t = tile + ty * tws + q;
            T acc[4];
#pragma unroll
            for (int k = 0; k < (KT ? KT : 1) + 3; k++) {
                const T v = t[(k & 3) * tq + (k >> 2)];   // column 4q + k
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int tap = k - j;
                    if (tap == 0) acc[j] = P::mul(p.g[0], v);
                    else if (tap > 0 && tap < (KT ? KT : 1)) acc[j] = P::fma(p.g[tap], v, acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) hbuf[ty * SEP_TW + 4 * q + j] = acc[j];
        }
    } else {
        for (int ty = wrp; ty < th; ty += SEP_THREADS / 32) {
            const T *t = tile + ty * tws;
            T acc = P::mul(p.g[0], t[col(lane)]);
            for (int k = 1; k < K; k++) acc = P::fma(p.g[k], t[col(lane + k)], acc);
            hbuf[ty * SEP_TW + lane] = acc;
        }
    }
    __syncthreads();

    // ---- 3. vertical pass on top of the bias and guard band: a thread produces four vertically adjacent outputs of its
    //         column from K + 3 loads.  Every pixel stores its fast value; a pixel with a channel inside the band is
    //         appended to a per-block list in shared memory.
    // ---- 4. the list is replayed by the WHOLE block, one (pixel, channel) per thread.  Round 1 replayed inside the row loop:
    //         the flagged lane ran the K*K-step chain for its four channels while the 31 others waited -- for 17x17 (band
    //         +-88 ulps, 2 % of the pixels flagged, so every second warp-row held one) that was 1420 of the kernel's 1920
    //         lane-instructions per pixel.  Compacted, the same chains run on all lanes at once.
    uint16_t *list = reinterpret_cast<uint16_t *>(hbuf + th * SEP_TW);   // [SEP_TH * SEP_TW] entries oy << 5 | lane
    __shared__ uint32_t n_list;
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    const int x = x0 + lane;
    {
        const int oy0 = wrp * 4;   // SEP_TH = 32 rows = 8 warps x 4
        const T *h = hbuf + oy0 * SEP_TW + lane;
        T f[4];
        if (KT) {
#pragma unroll
            for (int k = 0; k < (KT ? KT : 1) + 3; k++) {
                const T v = h[k * SEP_TW];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int tap = k - j;
                    if (tap == 0) f[j] = P::fma(p.g[0], v, P::splat(kSepBias));
                    else if (tap > 0 && tap < (KT ? KT : 1)) f[j] = P::fma(p.g[tap], v, f[j]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                f[j] = P::fma(p.g[0], h[j * SEP_TW], P::splat(kSepBias));
                for (int k = 1; k < K; k++) f[j] = P::fma(p.g[k], h[(j + k) * SEP_TW], f[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int oy = oy0 + j, y = y0 + oy;
            if (y < p.out_row0 + p.out_rows && y < p.H && x < p.W) {
                P::store(fdst + ((size_t)(y - p.out_row0) * p.W + x) * CN, (P::pack_fast(f[j]) & ~cm) | cexact);
                if (P::zmin_masked(f[j], p.zoff, cm, 0x43800000u + (p.zoff >> (32 - kSepFracBits))) < p.zthr) list[atomicAdd(&n_list, 1u)] = (uint16_t)(oy << 5 | lane);
            }
        }
    }
    __syncthreads();   // (also orders the fast stores above before the exact stores below, for the threads of this block)
    const uint32_t n_items = n_list * CN;
    if (n_items && threadIdx.x == 0 && p.slow_counter) atomicAdd(p.slow_counter, (unsigned long long)n_list);
    const float *tf = reinterpret_cast<const float *>(tile);
#pragma unroll 1
    for (uint32_t it = threadIdx.x; it < n_items; it += SEP_THREADS) {
        const uint32_t e = list[it / CN], c = it % CN, oy = e >> 5, ox = e & 31u;
        if ((cm >> (8 * c)) & 1u) continue;   // (a tile-constant channel already holds its exact value)
        // the reference's sequence (GaussianBlur.cpp:236-258) for one channel: ky-major / kx-minor from 0.0f, unfused
        float acc = 0.0f;
        for (int ky = 0; ky < K; ky++) {
            const float *tr = tf + (size_t)(oy + ky) * tws * CN + c;
            const float *wr = wts.w + ky * K;
#pragma unroll 4
            for (int kx = 0; kx < K; kx++) acc = __fadd_rn(acc, __fmul_rn(tr[col((int)ox + kx) * CN], wr[kx]));
        }
        const uint32_t v = (uint32_t)__float2int_rz(fminf(fmaxf(acc, 0.f), 255.f));   // (uchar)clamp(sum, 0, 255), GaussianBlur.cpp:255-258
        fdst[((size_t)(y0 + (int)oy - p.out_row0) * p.W + (x0 + (int)ox)) * CN + c] = (uint8_t)v;
    }
}

#include "rip_blur_stream.cuh"
#include "rip_blur_streamk.cuh"

unsigned long long *g_sep_slow_counter = nullptr;

}  // namespace

void blur_sep_set_slow_counter(unsigned long long *d_counter) { g_sep_slow_counter = d_counter; }

// half an ulp of the fp32 binade that holds x (x > 0): the largest rounding error of a result whose magnitude is <= x
static double hulp(double x) { return x > 0.0 ? std::ldexp(1.0, (int)std::floor(std::log2(x)) - 24) : 0.0; }

// Separable taps and the guard band for them.  All weights >= 0, pixel values <= 255, W_k = w_0 + .. + w_k:
//   reference:  acc_k = fl(acc_{k-1} + fl(p_k w_k)), acc_0 = fl(p_0 w_0): every rounding is bounded by half an ulp of the
//               binade its result can reach (as in rip_fused.cu; round 1 used u |value|, up to twice as much):
//               |S_ref - S| <= sum_k hulp(255 w_k) + sum_{k>=1} hulp(255 W_k)
//   separable model:  |S - S_sep| <= 255 sum_ij |w_ij - g_i g_j|            (S_sep = the exact separable sum)
//   horizontal FMA chain: result k is <= 255 G_k (G_k = g_0 + .. + g_k) and rounded once: error <= sum_k hulp(255 G_k),
//               which the vertical taps scale by sum(g)
//   vertical FMA chain on top of the bias: every result lies in [256, 512) and is rounded on that binade's 2^-16
//               half-ulp: <= K 2^-16 (the last of them is the +1/2 in `a` below; counted here as well)
//   margin 2 % + 1e-6 on top.
static bool plan_sep_blur(const float *w, int K, float *g, double *band_out)
{
    const double up = 1.0 + 1e-6;   // (bounds that sit just below a power of two are pushed into the next binade: safe side)
    double sum = 0.0, b_ref = 0.0;
    for (int i = 0; i < K * K; i++) {
        if (!(w[i] >= 0.0f) || !std::isfinite(w[i])) return false;
        sum += (double)w[i];
        b_ref += hulp(255.0 * (double)w[i] * up);
        if (i >= 1) b_ref += hulp(255.0 * sum + 1e-3);   // (+1e-3: the computed partial sums carry their own errors)
    }
    if (!(sum > 0.0) || 255.0 * sum >= 255.9) return false;   // floor(S) must stay <= 255
    for (int k = 0; k < K; k++) g[k] = (float)std::sqrt((double)w[k * K + k]);
    double dev = 0.0, gsum = 0.0, gpart = 0.0, b_h = 0.0;
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) dev += std::fabs((double)w[i * K + j] - (double)g[i] * (double)g[j]);
    for (int k = 0; k < K; k++) {
        gsum += (double)g[k];
        gpart += (double)g[k];
        b_h += hulp(255.0 * gpart * up);
    }
    const double band = (255.0 * dev + b_ref + b_h * gsum + K * std::ldexp(1.0, -16)) * 1.02 + 1e-6;
    if (band > 0.05) return false;   // not (close to) a symmetric separable kernel: the exact kernel handles it
    *band_out = band;
    return true;
}

// flat[v]: the reference's result (GaussianBlur.cpp:236-258: float accumulator from 0.0f, ky-major / kx-minor, one rounded
// product and one rounded add per tap, clamp, truncate) for a constant KxK window of value v
static void plan_flat_bytes(const float *w, int K, uint8_t flat[256])
{
    for (int v = 0; v < 256; v++) {
        volatile float acc = 0.0f;   // volatile: every product and every add is rounded to float, none is fused
        for (int i = 0; i < K * K; i++) {
            volatile float prod = (float)v * w[i];
            acc = acc + prod;
        }
        float s = acc;
        s = s < 0.0f ? 0.0f : s > 255.0f ? 255.0f : s;
        flat[v] = (uint8_t)(int)s;
    }
}

// The alpha channel of every frame the reference uploads is 255 (cv::COLOR_BGR2RGBA, ProgramHandler.cpp:127 of RT/), so its window is
// constant everywhere and its fast sum sits inside the guard band for every pixel.  The streaming kernel recognises that case from the
// fast sum itself: every step of its two FMA chains is monotone in every pixel, so the sum F255 of an all-255 window is the largest
// value the chain can produce, and F == F255 holds for NO other window iff lowering any single input by one step lowers the result.
// That is checked here with the kernel's own operand values and operation order (vertical: fma(G4,q4, fma(G3,q3, fma(G2,q2,
// fma(G1,q1, G0*q0)))) on the denormal bit patterns; horizontal on top of the bias); if it fails, f255 = 0 disables the shortcut.
static void plan_stream_alpha(SepParams &p)
{
    auto den = [](uint32_t q) { float f; memcpy(&f, &q, 4); return f; };   // q * 2^-149
    auto vsum = [&](const uint32_t q[5]) {
        volatile float a = p.sgv[0] * den(q[0]);
        for (int k = 1; k < 5; k++) a = std::fmaf(p.sgv[k], den(q[k]), a);
        return (float)a;
    };
    auto hsum = [&](const float v[5]) {
        volatile float a = p.sbias;
        for (int k = 0; k < 5; k++) a = std::fmaf(p.sgh[k], v[k], a);
        return (float)a;
    };
    p.f255 = 0;
    p.a255 = (uint32_t)p.flat[255] << 24;
    uint32_t q[5] = {255, 255, 255, 255, 255};
    const float v255 = vsum(q);
    float vsec = 0.0f;   // the largest vertical sum of a column that is not all 255
    for (int k = 0; k < 5; k++) {
        q[k] = 254;
        const float v = vsum(q);
        q[k] = 255;
        if (!(v < v255)) return;
        vsec = v > vsec ? v : vsec;
    }
    float v[5] = {v255, v255, v255, v255, v255};
    const float f255 = hsum(v);
    for (int k = 0; k < 5; k++) {
        v[k] = vsec;
        const float f = hsum(v);
        v[k] = v255;
        if (!(f < f255)) return;
    }
    if (!(f255 >= 256.0f && f255 < 512.0f)) return;
    memcpy(&p.f255, &f255, 4);
}

// the same proof for the streaming KxK kernel, whose chains run the other way round: horizontal from g[0] (sg1), vertical on the bias (sg2)
static void plan_streamk_alpha(SepParams &p, int K)
{
    auto den = [](uint32_t q) { float f; memcpy(&f, &q, 4); return f; };   // q * 2^-149
    auto hsum = [&](int k_low) {   // the row sum with pixel k_low one step below 255 (k_low < 0: all 255)
        volatile float a = p.sg1[0] * den(k_low == 0 ? 254u : 255u);
        for (int k = 1; k < K; k++) a = std::fmaf(p.sg1[k], den(k_low == k ? 254u : 255u), a);
        return (float)a;
    };
    p.f255 = 0;
    p.a255 = (uint32_t)p.flat[255] << 24;
    const float h255 = hsum(-1);
    float hsec = 0.0f;   // the largest row sum of a row that is not all 255
    for (int k = 0; k < K; k++) {
        const float h = hsum(k);
        if (!(h < h255)) return;
        hsec = h > hsec ? h : hsec;
    }
    auto vsum = [&](int k_low) {
        volatile float a = p.sbias;
        for (int k = 0; k < K; k++) a = std::fmaf(p.sg2[k], k_low == k ? hsec : h255, a);
        return (float)a;
    };
    const float f255 = vsum(-1);
    for (int k = 0; k < K; k++)
        if (!(vsum(k) < f255)) return;
    if (!(f255 >= 256.0f && f255 < 512.0f)) return;
    memcpy(&p.f255, &f255, 4);
}

template <int CN>
static int launch_sep_cn(cudaStream_t s, const SepParams &p, const Weights &wts, dim3 grid, size_t smem)
{
    void (*kern)(SepParams, Weights);
    switch (p.ksize) {
    case 3:  kern = blur_sep_kernel<CN, 3>; break;
    case 5:  kern = blur_sep_kernel<CN, 5>; break;
    case 7:  kern = blur_sep_kernel<CN, 7>; break;
    case 9:  kern = blur_sep_kernel<CN, 9>; break;
    case 17: kern = blur_sep_kernel<CN, 17>; break;
    default: kern = blur_sep_kernel<CN, 0>; break;
    }
    if (smem > 48 * 1024) RIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, SEP_THREADS, smem, s>>>(p, wts);
    RIP_LAUNCH_CHECK();
    return RIP_OK;
}

// returns RIP_EUNSUPPORTED (without recording an error) when the weights or the shape need the exact kernel
int launch_blur_sep(cudaStream_t s, const uint8_t *src, uint8_t *dst, int W, int H, int n_frames, int cn, int ksize,
                    const Weights &wts, int src_row0, int src_rows, int out_row0, int out_rows)
{
    if ((cn != 1 && cn != 4) || ksize < 3) return RIP_EUNSUPPORTED;
    if (options().blur_exact) return RIP_EUNSUPPORTED;   // (tests compare the two kernels)
    if (cn == 4 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3u)) return RIP_EUNSUPPORTED;
    // Everything that depends on the weights alone -- separable taps and the guard band (O(K^2) libm calls), the constant-window table
    // (256 K^2 operations), the scaled taps and weights of the streaming kernels, the alpha shortcut's proof -- is planned once and kept
    // for the next launch: a caller's weights rarely change, and for one small frame the planning cost more than the kernel (17x17 on a
    // 680 x 1023 frame: 45 us per call against 30 us best).
    struct Plan {
        bool valid = false, ok = false, stream5_ok = false, streamk_ok = false;
        int ksize = 0;
        float w[RIP_MAX_TAPS];
        SepParams p;
        Weights rws;   // the weights * 2^100 (streaming KxK kernel)
    };
    static std::mutex mu;
    static Plan plan;
    SepParams p;
    Weights rws;
    bool stream_ok, streamk_ok;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!plan.valid || plan.ksize != ksize || memcmp(plan.w, wts.w, sizeof(float) * ksize * ksize) != 0) {
            plan.valid = true;
            plan.ksize = ksize;
            memcpy(plan.w, wts.w, sizeof(float) * ksize * ksize);
            SepParams &q = plan.p;
            memset(&q, 0, sizeof(q));
            double band = 0.0;
            plan.ok = plan_sep_blur(wts.w, ksize, q.g, &band);
            if (plan.ok) {
                q.ksize = ksize;
                plan_flat_bytes(wts.w, ksize, q.flat);
                // F = S~ + 256 is rounded to a multiple of ulp = 2^-15 (error <= ulp/2): a pixel whose 15 fraction bits are
                // >= a and <= 2^15 - 1 - a has S~ at least `band` away from every integer when a >= band / ulp + 1/2
                const double ulp = std::ldexp(1.0, -kSepFracBits);
                const uint32_t a = (uint32_t)std::ceil(band / ulp + 0.5);
                q.zoff = a << (32 - kSepFracBits);
                q.zthr = (2u * a) << (32 - kSepFracBits);
                for (int k = 0; k < ksize; k++) {
                    q.sg1[k] = std::ldexp(q.g[k], 75);
                    q.sg2[k] = std::ldexp(q.g[k], 74);
                }
                q.sbias = (float)(256.0 + a * ulp);
                // the streaming kernels' replay multiplies the pixel's integer bit pattern (q * 2^-149) by w * 2^100: the product
                // q w 2^-49 carries the reference's mantissa iff it is a normal float, i.e. for weights that are 0 or >= 2^-70 (any
                // Gaussian with sigma >= 0.3); anything else takes the tiled kernel
                bool scaled_ok = true;
                memset(&plan.rws, 0, sizeof(plan.rws));
                for (int i = 0; i < ksize * ksize; i++) {
                    if (wts.w[i] != 0.0f && wts.w[i] < std::ldexp(1.0f, -70)) scaled_ok = false;
                    plan.rws.w[i] = std::ldexp(wts.w[i], 100);
                }
                plan.stream5_ok = ksize == 5 && scaled_ok;
                plan.streamk_ok = (ksize == 9 || ksize == 17) && scaled_ok;
                if (ksize == 5) {
                    for (int k = 0; k < 5; k++) {
                        q.sgv[k] = q.sg1[k];
                        q.sgh[k] = q.sg2[k];
                    }
                    memcpy(q.rw, plan.rws.w, sizeof(q.rw));
                    plan_stream_alpha(q);
                } else if (plan.streamk_ok) {
                    plan_streamk_alpha(q, ksize);
                }
            }
        }
        if (!plan.ok) return RIP_EUNSUPPORTED;
        p = plan.p;
        stream_ok = plan.stream5_ok;
        streamk_ok = plan.streamk_ok;
        if (streamk_ok) rws = plan.rws;
    }
    p.src = src; p.dst = dst; p.W = W; p.H = H;
    p.src_row0 = src_row0; p.src_rows = src_rows; p.out_row0 = out_row0; p.out_rows = out_rows;
    p.slow_counter = g_sep_slow_counter;
    // Which kernel (one frame, measured, tiled / streaming, us -- profiles/r2x_blur_small.txt): 5x5 on 640x512 33 / 25, 680x1023 27 / 23,
    // 1080p 40 / 24, 4K 93 / 46; 9x9 on 1080p 44 / 44, 4K 125 / 87; 17x17 on 1080p 73 / 95, 4K 227 / 197.  The streaming kernels walk
    // K - 1 warm-up rows per segment, so they need enough rows per warp to pay for them: from 0.3 Mpx for 5x5, 3 Mpx for 9x9, 6 Mpx for 17x17.
    const long long n_px = (long long)n_frames * out_rows * W;
    const bool big = n_px >= 300000LL;
    if (cn == 4 && stream_ok && !options().blur_tiled && (big || options().blur_stream)) {
        // the streaming kernel: bands of 60 columns per warp, segments of rows sized so that the grid fills the GPU
        // (4 warm-up rows per segment)
        StreamGeo sg;
        const int n_bands = (W + kBsBand - 1) / kBsBand;
        sg.n_band_groups = (n_bands + kBsWarps - 1) / kBsWarps;
        int device = 0;
        cudaGetDevice(&device);
        const long long want = (long long)sm_count(device) * 16;
        int seg = 128;
        while (seg > 8 && (long long)n_frames * sg.n_band_groups * ((out_rows + seg - 1) / seg) < want) seg >>= 1;
        sg.seg_rows = seg < out_rows ? seg : out_rows;
        sg.n_segs = (out_rows + sg.seg_rows - 1) / sg.seg_rows;
        const long long blocks = (long long)n_frames * sg.n_segs * sg.n_band_groups;
        if (blocks > 0 && blocks <= 0x7fffffffLL) {
            blur_stream5_kernel<<<(unsigned)blocks, kBsWarps * 32, 0, s>>>(p, sg);
            RIP_LAUNCH_CHECK();
            return RIP_OK;
        }
    }
    // 9x9 and 17x17 RGBA (the reference's default): the accumulate-form streaming kernel for large inputs
    if (cn == 4 && streamk_ok && !options().blur_tiled && (n_px >= (ksize == 9 ? 3000000LL : 6000000LL) || options().blur_stream) && W <= 65536) {
        StreamGeo sg;
        const int n_bands = (W + kSkBand - 1) / kSkBand;
        sg.n_band_groups = (n_bands + kSkWarps - 1) / kSkWarps;
        int device = 0;
        cudaGetDevice(&device);
        // rows per segment: every segment pays K - 1 warm-up rows, and the grid runs in waves of `resident` blocks: take the split
        // with the smallest (number of waves) x (rows a block walks)
        const long long resident = (long long)sm_count(device) * (ksize <= 9 ? 5 : 4), per_seg = (long long)n_frames * sg.n_band_groups;
        long long best = -1;
        int best_n = 1;
        for (int n = 1; n <= 64 && (out_rows + n - 1) / n >= 2 * ksize; n++) {
            // (between the two models "blocks flow through the SMs" and "whole waves": a sparse last wave runs faster, not for free)
            const long long rows = (out_rows + n - 1) / n + ksize - 1, blocks_n = per_seg * n, waves = (blocks_n + resident - 1) / resident;
            const long long cost = rows * (blocks_n + waves * resident);
            if (best < 0 || cost < best) {
                best = cost;
                best_n = n;
            }
        }
        sg.seg_rows = (out_rows + best_n - 1) / best_n;
        if (sg.seg_rows > 8192) sg.seg_rows = 8192;   // (the list entries hold 14 bits of row offset)
        sg.n_segs = (out_rows + sg.seg_rows - 1) / sg.seg_rows;
        const long long blocks = (long long)n_frames * sg.n_segs * sg.n_band_groups;
        if (blocks > 0 && blocks <= 0x7fffffffLL) {
            if (ksize == 9) blur_streamk_kernel<9><<<(unsigned)blocks, kSkWarps * 32, 0, s>>>(p, rws, sg);
            else blur_streamk_kernel<17><<<(unsigned)blocks, kSkWarps * 32, 0, s>>>(p, rws, sg);
            RIP_LAUNCH_CHECK();
            return RIP_OK;
        }
    }
    const int half = ksize >> 1;
    const size_t elem = cn == 4 ? sizeof(float4) : sizeof(float);
    const size_t smem = ((size_t)(SEP_TH + 2 * half) * ((SEP_TW + 2 * half + 3) & ~3) + (size_t)(SEP_TH + 2 * half) * SEP_TW) * elem +
                        (size_t)SEP_TH * SEP_TW * sizeof(uint16_t);   // tile, horizontal pass, list of guard-band pixels
    const dim3 grid((W + SEP_TW - 1) / SEP_TW, (out_rows + SEP_TH - 1) / SEP_TH, n_frames);
    if (grid.y > 65535u || grid.z > 65535u) return RIP_EUNSUPPORTED;
    return cn == 4 ? launch_sep_cn<4>(s, p, wts, grid, smem) : launch_sep_cn<1>(s, p, wts, grid, smem);
}

}  // namespace rip
