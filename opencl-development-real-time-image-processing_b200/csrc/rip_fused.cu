// rip_fused.cu -- gray -> 5x5 Gaussian -> 3x3 Sobel in ONE kernel (one HBM round trip per frame),
// and the same kernel without the blur stage (gray -> Sobel, BASELINE config 3).
//
// Design (sm_100a, HBM/issue-bound integer+fp32 stencil; no tensor cores by design):
//   * every WARP is independent.  A lane owns 4 horizontally adjacent pixels; a warp covers 128
//     pixels of which the middle 120 (lanes 1..30) are outputs and the outer lanes are halo.
//     The warp slides DOWN a row segment, so the vertical halo costs 6 warm-up rows per segment
//     and the horizontal halo 8/128 of the lanes.  No shared memory, no block barriers.
//   * per new image row a lane loads its 12 (RGB) or 16 (RGBA) bytes with 32/128-bit loads that
//     are prefetched two rows ahead, converts to the reference's exact gray, and keeps the last 5
//     gray rows / 2 Sobel partial rows in REGISTERS (the loop is unrolled by 10 = lcm(5,2) so the
//     ring indices are compile-time constants and no register moves are needed).
//   * the horizontal neighbours (2 per side for the blur, 1 per side for Sobel) come from the
//     adjacent lanes with warp shuffles.
//   * exactness: the blurred value must equal the reference's sequential, unfused, 25-tap fp32
//     sum truncated to u8 (GaussianBlur.cpp:236-258).  The fast path evaluates a separable fp32
//     sum S~ and rounds with a magic-number add; a pixel whose S~ is closer to an integer than the
//     proven bound on |S~ - S_ref| (rip_fused_band below) is recomputed with the exact 25-tap
//     sequence (__fmul_rn/__fadd_rn in the reference order).  So the u8 blurred image, and with it
//     the Sobel output, is bit-exact, at separable cost on all but ~0.1 % of the pixels.
//   * Sobel: gx, gy from separable partial sums in fp32 (small integers, exact), magnitude via
//     sqrt.approx (its error is 8x below the distance of any integer's root to a rounding
//     boundary for results < 255.5), saturate, round with the magic-number add, pack 4 bytes,
//     one 32-bit store per lane per row.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>

#include "rip_common.cuh"
#include "rip_internal.h"

#ifndef RIP_LDG_PF
#define RIP_LDG_PF 1
#endif
#ifndef RIP_LDG_MINBLOCKS
#define RIP_LDG_MINBLOCKS 6
#endif
// x2 kernel (rip_fused_x2.cuh): prefetch depth in rows, resident blocks per SM, row-loop unroll
#ifndef RIP_X2_PF
#define RIP_X2_PF 2
#endif
#ifndef RIP_X2_MINB8
#define RIP_X2_MINB8 4
#endif
#ifndef RIP_X2_MINB4
#define RIP_X2_MINB4 6
#endif
#ifndef RIP_X2_UNROLL
#define RIP_X2_UNROLL 3
#endif
#ifndef RIP_X2_L2PF
#define RIP_X2_L2PF 6   // rows ahead of the register loads that prefetch.global.L2 runs (0 = off)
#endif

namespace rip {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kBandPx = 120;      // output pixels per warp per row (lanes 1..30 x 4 px)
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: x + kMagic rounds x to the nearest integer (ties to even)

struct FusedParams {
    const uint8_t *in;
    uint8_t *out;
    int W, H;
    int in_row0, in_rows, out_row0, out_rows;
    int seg_rows, n_segs, n_bands, n_band_groups;
    float g0, g1, g2;    // separable taps: w2d[ky][kx] ~= g[|ky|] * g[|kx|]
    float thr;           // slow path if |frac - 0.5| > thr  (thr = 0.5 - band)
    float w[25];         // exact 2-D weights for the slow path
    unsigned long long *slow_counter;  // optional statistics (NULL in production)
};

// ---- exact gray of 4 packed pixels -> 4 floats ------------------------------------------------
// t = 299r + 587g + 114b via two 2-way dot products per pixel; q = floor(t/1000) = hi32(t * 4294968)
// (exact for t <= 255000: 4294968*1000 - 2^32 = 704 and 255000*704 < 2^32); t % 1000 == 0 iff the
// low word of that product is < 2^18 (it is 704*q <= 179520 then, and >= 4294968 otherwise).
// Off the multiples of 1000 the reference's double expression (Comparator.cpp:41) truncates to q
// (it is >= 1e-3 away from an integer, the double rounding error is < 1e-12).  ON a multiple of
// 1000 the rounding of the three double products decides between q and q-1; since 114*b mod 1000
// has period 500 > 255, (r,g) determines that b uniquely, so one bit per (r,g) -- tabulated on
// the host by evaluating the reference expression itself -- says whether the result is q-1.
__device__ uint32_t d_gray_down[2048];  // bit (r<<8|g): the double evaluation lands below q

template <int CN, bool BGR>
__device__ __forceinline__ void gray4(const uint32_t *w, float f[4])
{
    constexpr uint32_t cA = BGR ? 114u : 299u, cB = 587u, cC = BGR ? 299u : 114u;  // weights of byte 0,1,2
    constexpr uint32_t AB = cA | (cB << 16), C0 = cC, zA = cA << 16, BC = cB | (cC << 16);
    uint32_t t[4], lo[4];
    if constexpr (CN == 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) t[j] = __dp2a_hi(C0, w[j], __dp2a_lo(AB, w[j], 0u));  // alpha x 0
    } else {
        // byte stream: p0 = w0.b0-2, p1 = w0.b3 w1.b0-1, p2 = w1.b2-3 w2.b0, p3 = w2.b1-3
        t[0] = __dp2a_hi(C0, w[0], __dp2a_lo(AB, w[0], 0u));
        t[1] = __dp2a_lo(BC, w[1], __dp2a_hi(zA, w[0], 0u));
        t[2] = __dp2a_lo(C0, w[2], __dp2a_hi(AB, w[1], 0u));
        t[3] = __dp2a_hi(BC, w[2], __dp2a_lo(zA, w[2], 0u));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const unsigned long long prod = (unsigned long long)t[j] * 4294968ull;
        lo[j] = (uint32_t)prod;
        f[j] = (float)(uint32_t)(prod >> 32);
    }
    // warp-uniform test (one vote) so the warp stays converged for the shuffles that follow
    if (__builtin_expect(__any_sync(0xffffffffu, min(min(lo[0], lo[1]), min(lo[2], lo[3])) < (1u << 18)), 0)) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (lo[j] < (1u << 18)) {
                uint32_t px;  // the pixel's three channel bytes in bits 0..23
                if constexpr (CN == 4) px = w[j];
                else px = j == 0 ? w[0] : j == 1 ? __funnelshift_r(w[0], w[1], 24) : j == 2 ? __funnelshift_r(w[1], w[2], 16) : (w[2] >> 8);
                const uint32_t r = BGR ? (px >> 16) & 0xffu : px & 0xffu, g = (px >> 8) & 0xffu;
                const uint32_t idx = (r << 8) | g;
                f[j] -= (float)((__ldg(&d_gray_down[idx >> 5]) >> (idx & 31u)) & 1u);
            }
        }
    }
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int CN>
struct RawRow {
    uint32_t w[CN];  // CN 32-bit words = 4 pixels of CN bytes
};

template <int CN>
__device__ __forceinline__ RawRow<CN> load_row(const uint8_t *p, bool valid)
{
    RawRow<CN> r;  // lanes outside the image keep stale registers: their pixels are never consumed
    if (valid) {
        if constexpr (CN == 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
            r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        } else {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
            r.w[0] = __ldg(q); r.w[1] = __ldg(q + 1); r.w[2] = __ldg(q + 2);
        }
    }
    return r;
}

constexpr unsigned FULL = 0xffffffffu;
constexpr int kRingRow = 128 + 8;  // floats per gray row in the per-warp shared ring (4 pad each side)

#include "rip_fused_x2.cuh"

// Per-warp sliding-window state, all in registers.  The row loop is NOT unrolled: the whole hot
// loop is ~3 KB of SASS and stays resident in the per-partition instruction cache (an earlier
// 5x-unrolled version was instruction-fetch bound, see profiles/).  To make a rolled loop possible
// the vertical blur runs in accumulate form -- each new gray row is added into the four pending
// blurred rows with FMAs whose destination is the *next* accumulator, so the shift costs no
// register moves -- and only the small Sobel ring is shifted explicitly.
template <int CN>
struct WarpState {
    float a0[4], a1[4], a2[4], a3[4];  // partial vertical sums of blurred rows r-2, r-1, r, r+1 (missing rows >= r)
    float X0[4], X1[4];                // D(yb-2) + 2 D(yb-1)  and  D(yb-1),  D(y) = b[x+1] - b[x-1] of blurred row y
    float S1[4], S2[4];                // S(yb-1), S(yb-2),    S(y) = b[x-1] + 2 b[x] + b[x+1]
    RawRow<CN> pre;                    // input row r+1, prefetched one step ahead
#if RIP_LDG_PF == 2
    RawRow<CN> pre2;                   // input row r+2
#endif
};

struct Geometry {
    const uint8_t *src;      // this lane's pixels in the input row that is prefetched next
    uint8_t *dst;            // this lane's pixels in the output row produced next (may point before the
                             // band during the warm-up rows; only dereferenced for valid rows)
    float *ring;             // this warp's gray ring [5][kRingRow] in shared memory
    float *ring_cur;         // row of the ring holding the newest gray row (this lane's 4 columns)
    uint32_t in_pitch;
    int lane, lane_last;
    bool left_edge, right_edge, in_img;
    uint32_t store_lane;
    int ys;                  // first output row of the segment
};

// Cold path, out of line: exact replay of the reference's 25-tap sum (GaussianBlur.cpp:236-258) for
// the pixels inside the guard band.  The gray rows yb-2..yb+2 sit in the warp's shared-memory ring
// (slot_new holds row yb+2), so a lane reads its +-2 neighbour columns directly.  Per flagged
// component the 25 products are summed ky-major / kx-minor from 0.0f with unfused multiply and
// add, clamped to [0,255] and truncated -- exactly the reference sequence.
__device__ __noinline__ float4 blur_exact(const float *ring, int slot_new, const float *w25, float4 b, uint32_t mask, int lane)
{
    __syncwarp();  // the newest row was just stored by the other lanes
    float out[4] = {b.x, b.y, b.z, b.w};
    const float *base = ring + 4 + 4 * lane - 2;  // column x-2 of component 0
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (mask & (1u << j)) {
            float acc = 0.f;
            int slot = slot_new;
#pragma unroll
            for (int ky = 0; ky < 5; ky++) {
                slot = slot == 4 ? 0 : slot + 1;  // oldest row first
                const float *row = base + slot * kRingRow + j;
#pragma unroll
                for (int kx = 0; kx < 5; kx++) acc = __fadd_rn(acc, __fmul_rn(row[kx], w25[ky * 5 + kx]));
            }
            out[j] = truncf(fminf(fmaxf(acc, 0.f), 255.f));
        }
    }
    __syncwarp();  // the ring slot of the oldest row is overwritten by the next step
    return make_float4(out[0], out[1], out[2], out[3]);
}

// One image row of the sliding window.
//   EDGE    the warp's band touches the left/right image border (clamp / reflect fix-ups in x)
//   STORE   the step produces an output row (false for the warm-up rows of a segment)
//   SPECIAL the step may be the first or last row of the frame (BORDER_REFLECT_101 in y); only the
//           first and last storing step of a segment are instantiated with it, so the main loop
//           carries no per-row border checks
template <int CN, bool BGR, bool BLUR, bool EDGE, bool STORE, bool SPECIAL>
__device__ __forceinline__ void step(WarpState<CN> &st, const FusedParams &p, Geometry &geo, int r)
{
    const int W = p.W, H = p.H, lane = geo.lane;
    // ---- 1. gray of the new row r; prefetch row r+1 (row index clamped to the rows of the band) ----
    float f[4];
    {
        const RawRow<CN> raw = st.pre;
#if RIP_LDG_PF == 2
        st.pre = st.pre2;
        if ((unsigned)(r + 1 - p.in_row0) < (unsigned)(p.in_rows - 1)) geo.src += geo.in_pitch;
        st.pre2 = load_row<CN>(geo.src, geo.in_img);
#else
        if ((unsigned)(r - p.in_row0) < (unsigned)(p.in_rows - 1)) geo.src += geo.in_pitch;
        st.pre = load_row<CN>(geo.src, geo.in_img);
#endif
        gray4<CN, BGR>(raw.w, f);
    }
    float b[4];  // blurred row yb as exact u8 values held in floats; without the blur stage: the gray row
    const int yb = BLUR ? r - 2 : r;
    if constexpr (BLUR) {
        // clamp-to-edge columns (GaussianBlur.cpp:240): x < 0 -> column 0, x >= W -> column W-1
        if constexpr (EDGE) {
            const float first = __shfl_sync(FULL, f[0], 1);
            const float last = __shfl_sync(FULL, f[3], min(geo.lane_last, 31));
            if (geo.left_edge && lane == 0) f[0] = f[1] = f[2] = f[3] = first;
            if (geo.right_edge && lane > geo.lane_last) f[0] = f[1] = f[2] = f[3] = last;
        }
        // park the gray row in the shared ring (only the cold exact replay reads it back)
        geo.ring_cur += kRingRow;
        if (geo.ring_cur == geo.ring + 5 * kRingRow + 4 + 4 * lane) geo.ring_cur -= 5 * kRingRow;
        *reinterpret_cast<float4 *>(geo.ring_cur) = make_float4(f[0], f[1], f[2], f[3]);
        // vertical pass, accumulate form: row r completes blurred row r-2
        float V[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            V[j] = fmaf(p.g2, f[j], st.a0[j]);
            st.a0[j] = fmaf(p.g1, f[j], st.a1[j]);
            st.a1[j] = fmaf(p.g0, f[j], st.a2[j]);
            st.a2[j] = fmaf(p.g1, f[j], st.a3[j]);
            st.a3[j] = p.g2 * f[j];
        }
        const float Vm2 = __shfl_up_sync(FULL, V[2], 1), Vm1 = __shfl_up_sync(FULL, V[3], 1);
        const float Vp4 = __shfl_down_sync(FULL, V[0], 1), Vp5 = __shfl_down_sync(FULL, V[1], 1);
        const float c[8] = {Vm2, Vm1, V[0], V[1], V[2], V[3], Vp4, Vp5};
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float e2 = c[j] + c[j + 4], e1 = c[j + 1] + c[j + 3];
            const float u = fmaf(p.g2, e2, fmaf(p.g1, e1, fmaf(p.g0, c[j + 2], -0.5f)));  // S~ - 0.5
            const float rr = u + kMagic;  // nearest integer to S~ - 0.5: floor(S~) outside the guard band
            b[j] = rr - kMagic;
            d[j] = fabsf(u - b[j]);       // |frac(S~) - 0.5|
        }
        const bool slow = fmaxf(fmaxf(d[0], d[1]), fmaxf(d[2], d[3])) > p.thr;
        if (__builtin_expect(__any_sync(FULL, slow), 0)) {
            const uint32_t mask = (d[0] > p.thr ? 1u : 0u) | (d[1] > p.thr ? 2u : 0u) | (d[2] > p.thr ? 4u : 0u) |
                                  (d[3] > p.thr ? 8u : 0u);
            const int slot = (int)(geo.ring_cur - (geo.ring + 4 + 4 * lane)) / kRingRow;
            const float4 fx = blur_exact(geo.ring, slot, p.w, make_float4(b[0], b[1], b[2], b[3]), mask, lane);
            b[0] = fx.x; b[1] = fx.y; b[2] = fx.z; b[3] = fx.w;
            if (p.slow_counter && mask) atomicAdd(p.slow_counter, (unsigned long long)__popc(mask));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = f[j];
    }

    // ---- 3. Sobel partial sums of row yb, BORDER_REFLECT_101 in x --------------------------------
    float bl = __shfl_up_sync(FULL, b[3], 1), br = __shfl_down_sync(FULL, b[0], 1);
    if constexpr (EDGE) {
        if (geo.left_edge && lane == 1) bl = b[1];                 // x = -1 -> x = 1
        if (geo.right_edge && lane == geo.lane_last) br = b[2];    // x = W  -> x = W-2
    }
    float Dc[4], Sc[4];
    {
        const float e[6] = {bl, b[0], b[1], b[2], b[3], br};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            Dc[j] = e[j + 2] - e[j];
            Sc[j] = fmaf(2.f, e[j + 1], e[j] + e[j + 2]);
        }
    }
    // ---- 4. output row yo = yb-1:  gx = D(yo-1) + 2 D(yo) + D(yo+1),  gy = S(yo+1) - S(yo-1) ------
    if constexpr (SPECIAL) {  // BORDER_REFLECT_101 in y
        if (yb == 1) {        // output row 0: row -1 -> row 1:  gx = 2 D(0) + 2 D(1), gy = 0
#pragma unroll
            for (int j = 0; j < 4; j++) { st.X0[j] = fmaf(2.f, st.X1[j], Dc[j]); st.S2[j] = Sc[j]; }
        }
        if (yb == H) {        // output row H-1: row H -> row H-2 (this step's input row was a dummy): D(H-2) = X0 - 2 X1
#pragma unroll
            for (int j = 0; j < 4; j++) { Dc[j] = fmaf(-2.f, st.X1[j], st.X0[j]); Sc[j] = st.S2[j]; }
        }
    }
    if constexpr (STORE) {
        float q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float gx = st.X0[j] + Dc[j];
            const float gy = Sc[j] - st.S2[j];
            const float m = sqrt_approx(fmaf(gx, gx, gy * gy));
            q[j] = fminf(m, 255.f) + kMagic;  // saturate, round half to even: result in the low byte
        }
        const uint32_t q01 = __byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x0040);
        const uint32_t q23 = __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x0040);
        // predicated store (no branch: lanes 0 and 31 are halo lanes and must not diverge here)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
                     :: "l"(geo.dst), "r"(__byte_perm(q01, q23, 0x5410)), "r"(geo.store_lane) : "memory");
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        st.X0[j] = fmaf(2.f, Dc[j], st.X1[j]);
        st.X1[j] = Dc[j];
        st.S2[j] = st.S1[j];
        st.S1[j] = Sc[j];
    }
    geo.dst += W;
}

#ifndef RIP_MAIN_UNROLL
#define RIP_MAIN_UNROLL 1
#endif
constexpr int kMainUnroll = RIP_MAIN_UNROLL;

template <int CN, bool BGR, bool BLUR, bool EDGE>
__device__ __forceinline__ void run_segment(WarpState<CN> &st, const FusedParams &p, Geometry &geo, int r, int r_last)
{
    constexpr int HALO = BLUR ? 3 : 1;
    const int r_store = geo.ys + HALO;  // first step that produces an output row
#pragma unroll 1
    for (; r < r_store; r++) step<CN, BGR, BLUR, EDGE, false, false>(st, p, geo, r);   // warm-up rows
    step<CN, BGR, BLUR, EDGE, true, true>(st, p, geo, r);                              // may be frame row 0 (and H-1)
    r++;
#pragma unroll kMainUnroll
    for (; r < r_last; r++) step<CN, BGR, BLUR, EDGE, true, false>(st, p, geo, r);     // main loop: no border checks
    if (r == r_last) step<CN, BGR, BLUR, EDGE, true, true>(st, p, geo, r);             // may be frame row H-1
}

template <int CN, bool BGR, bool BLUR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, RIP_LDG_MINBLOCKS)
fused_kernel(const __grid_constant__ FusedParams p)
{
    constexpr int HALO = BLUR ? 3 : 1;  // input rows above/below an output row

    __shared__ __align__(16) float ring[BLUR ? kWarpsPerBlock * 5 * kRingRow : 4];
    Geometry geo;
    geo.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    geo.ring = ring + (BLUR ? warp * 5 * kRingRow : 0);
    geo.ring_cur = geo.ring + 4 + 4 * geo.lane;
    int bid = blockIdx.x;
    const int bg = bid % p.n_band_groups; bid /= p.n_band_groups;
    const int seg = bid % p.n_segs;
    const int frame = bid / p.n_segs;
    const int band = bg * kWarpsPerBlock + warp;
    if (band >= p.n_bands) return;  // warp-uniform, and there are no block-level barriers

    const int W = p.W;
    const int xw0 = band * kBandPx;
    const int x = xw0 - 4 + 4 * geo.lane;        // first of this lane's 4 pixels
    geo.in_img = (x >= 0) && (x < W);            // W % 4 == 0: a lane is fully inside or fully outside
    geo.lane_last = (W - xw0) >> 2;              // lane holding pixels W-4..W-1 (may be > 31)
    geo.left_edge = (band == 0);
    geo.right_edge = (geo.lane_last <= 31);
    geo.ys = p.out_row0 + seg * p.seg_rows;
    const int ye = min(geo.ys + p.seg_rows, p.out_row0 + p.out_rows);
    geo.in_pitch = (uint32_t)W * CN;
    const uint8_t *in_base = p.in + (size_t)frame * p.in_rows * geo.in_pitch;
    geo.store_lane = ((geo.lane >= 1) && (geo.lane <= 30) && geo.in_img) ? 1u : 0u;

    WarpState<CN> st;
#pragma unroll
    for (int j = 0; j < 4; j++)
        st.a0[j] = st.a1[j] = st.a2[j] = st.a3[j] = st.X0[j] = st.X1[j] = st.S1[j] = st.S2[j] = 0.f;

    const int r_first = geo.ys - HALO, r_last = ye - 1 + HALO;
    const uint32_t xoff = geo.in_img ? (uint32_t)x * CN : 0u;
    // Row indices are clamped to the rows the input band holds.  The host guarantees the band
    // covers every row an output needs, and that it starts at row 0 / ends at row H-1 wherever the
    // clamp-to-edge rule (GaussianBlur.cpp:241) is actually exercised; other clamped rows are
    // read-ahead only and never consumed.
    geo.src = in_base + (size_t)min(max(r_first - p.in_row0, 0), p.in_rows - 1) * geo.in_pitch + xoff;
    st.pre = load_row<CN>(geo.src, geo.in_img);
#if RIP_LDG_PF == 2
    if ((unsigned)(r_first - p.in_row0) < (unsigned)(p.in_rows - 1)) geo.src += geo.in_pitch;
    st.pre2 = load_row<CN>(geo.src, geo.in_img);
#endif
    // output row produced by the step of input row r is r - HALO
    geo.dst = p.out + (size_t)frame * p.out_rows * W + (ptrdiff_t)(r_first - HALO - p.out_row0) * W + x;

    if (geo.left_edge || geo.right_edge) run_segment<CN, BGR, BLUR, true>(st, p, geo, r_first, r_last);
    else run_segment<CN, BGR, BLUR, false>(st, p, geo, r_first, r_last);
}


// =============================================================================================
// TMA variant (the main path): same arithmetic and the same rolled row loop, but the input rows are
// staged in shared memory by the Tensor Memory Accelerator instead of per-lane global loads:
//   * one elected lane issues a 2-D cp.async.bulk.tensor for a [TR rows x band] box of the warp's
//     column band into a per-warp, double-buffered ring; every lane waits on the stage's mbarrier
//     and then reads its own pixels with shared-memory loads.  The prefetch is 1-2 tiles (4-8 rows)
//     deep and costs no registers, so DRAM latency is off the critical path.
//   * the tensor is addressed in 32-bit elements; the TMA unit requires a box to START on a 16-byte
//     boundary of the row (measured: any other start raises "illegal instruction",
//     tools/tma_probe.cu).  For 3-byte pixels that is a multiple of 16 pixels, so the RGB box starts
//     at the band start rounded down to 16 px and is wide enough for the worst rounding; columns
//     outside the image are zero-filled by the TMA unit and replaced by the clamp-to-edge fix.
//   * NPX = pixels per lane: 4 (default) or 8 (fewer shuffles per pixel but a loop body that no
//     longer fits the instruction cache; kept for experiments, RIP_FUSED_NPX=8).
// =============================================================================================
constexpr int TR = 4;    // rows per TMA box
constexpr int NST = 2;   // stages in the per-warp ring

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

struct TmaParams {
    FusedParams f;
    int tile_words;    // 32-bit words per box row
    int stage_words;   // words per ring stage = TR * tile_words rounded up to 128 bytes
};

template <int NPX>
struct WarpStateT {
    float a0[NPX], a1[NPX], a2[NPX], a3[NPX];  // partial vertical sums of blurred rows r-2..r+1
    float X0[NPX], X1[NPX];                    // D(yb-2) + 2 D(yb-1), D(yb-1)
    float S1[NPX], S2[NPX];                    // S(yb-1), S(yb-2)
};

struct GeometryT {
    uint8_t *dst;            // this lane's pixels in the output row produced next
    float *ring;             // this warp's gray ring [5][32 * NPX + 8]
    float *ring_cur;         // this lane's columns in the ring row of the newest gray row
    const uint32_t *tiles;   // this warp's TMA ring [NST][stage_words]
    const uint32_t *src;     // this lane's words in the ring row of the input row consumed next
    uint64_t *bars;          // this warp's NST full barriers
    int lane, lane_last;
    bool left_edge, right_edge;
    uint32_t store_lane;
    int ys;
    int lane_word;           // word offset of this lane's pixels inside a box row
    int row_in_tile;         // row of the current tile that the next step consumes
    int k_cur;               // tile the consumer is in
    int tma_x, tma_y0;       // tensor coordinates of tile 0: word column, global row
};

// exact replay (see blur_exact); values travel by value so they stay in registers
template <int NPX>
struct FN {
    float v[NPX];
};

template <int NPX>
__device__ __noinline__ FN<NPX> blur_exact_n(const float *ring, int slot_new, const float *w25, FN<NPX> b, uint32_t mask, int lane)
{
    constexpr int kRow = 32 * NPX + 8;
    __syncwarp();
    const float *base = ring + 4 + NPX * lane - 2;
#pragma unroll
    for (int j = 0; j < NPX; j++) {
        if (mask & (1u << j)) {
            float acc = 0.f;
            int slot = slot_new;
#pragma unroll
            for (int ky = 0; ky < 5; ky++) {
                slot = slot == 4 ? 0 : slot + 1;
                const float *row = base + slot * kRow + j;
#pragma unroll
                for (int kx = 0; kx < 5; kx++) acc = __fadd_rn(acc, __fmul_rn(row[kx], w25[ky * 5 + kx]));
            }
            b.v[j] = truncf(fminf(fmaxf(acc, 0.f), 255.f));
        }
    }
    __syncwarp();
    return b;
}

// Advance to the next tile of the ring: refill the stage just left, wait for the one entered.  Out
// of line and by value (it runs once per TR rows): returns the new tile index; the caller rebuilds
// its row pointer from it.
__device__ __noinline__ int tma_next_tile(const uint32_t *tiles, uint64_t *bars, const CUtensorMap *map, int k_cur, int tma_x, int tma_y0,
                                          int tile_words, int stage_words, int lane)
{
    __syncwarp();  // every lane is done reading tile k_cur
    if (lane == 0) {
        const int kn = k_cur + NST, stage = k_cur % NST;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bars + stage, (uint32_t)(TR * tile_words * 4));
        tma_load_2d(const_cast<uint32_t *>(tiles) + stage * stage_words, map, tma_x, tma_y0 + kn * TR, bars + stage);
    }
    k_cur++;
    mbar_wait(bars + (k_cur % NST), (uint32_t)((k_cur / NST) & 1));
    return k_cur;
}

template <int NPX, int CN, bool BGR, bool BLUR, bool EDGE, bool STORE, bool SPECIAL>
__device__ __forceinline__ void step_t(WarpStateT<NPX> &st, const TmaParams &tp, const CUtensorMap *map, GeometryT &geo, int r)
{
    const FusedParams &p = tp.f;
    const int W = p.W, H = p.H, lane = geo.lane;
    constexpr int NW = NPX * CN / 4;        // words per lane per row
    constexpr int kRow = 32 * NPX + 8;      // floats per gray-ring row
    // ---- 1. this lane's pixels of row r from the TMA ring ----------------------------------------
    uint32_t raw[NW];
    if constexpr (NW == 3) {
        raw[0] = geo.src[0]; raw[1] = geo.src[1]; raw[2] = geo.src[2];
    } else if constexpr (NW == 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(geo.src);
        raw[0] = v.x; raw[1] = v.y; raw[2] = v.z; raw[3] = v.w;
    } else if constexpr (NW == 6) {
        const uint2 v0 = *reinterpret_cast<const uint2 *>(geo.src), v1 = *reinterpret_cast<const uint2 *>(geo.src + 2),
                    v2 = *reinterpret_cast<const uint2 *>(geo.src + 4);
        raw[0] = v0.x; raw[1] = v0.y; raw[2] = v1.x; raw[3] = v1.y; raw[4] = v2.x; raw[5] = v2.y;
    } else {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(geo.src), v1 = *reinterpret_cast<const uint4 *>(geo.src + 4);
        raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w; raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
    }
    // next step consumes row clamp(r+1): advance unless clamped to the first/last row of the band
    if ((unsigned)(r - p.in_row0) < (unsigned)(p.in_rows - 1)) {
        geo.src += tp.tile_words;
        if (++geo.row_in_tile == TR) {
            geo.k_cur = tma_next_tile(geo.tiles, geo.bars, map, geo.k_cur, geo.tma_x, geo.tma_y0, tp.tile_words, tp.stage_words, lane);
            geo.row_in_tile = 0;
            geo.src = geo.tiles + (geo.k_cur % NST) * tp.stage_words + geo.lane_word;
        }
    }
    float f[NPX];
    gray4<CN, BGR>(raw, f);
    if constexpr (NPX == 8) gray4<CN, BGR>(raw + NW / 2, f + 4);

    float b[NPX];
    const int yb = BLUR ? r - 2 : r;
    if constexpr (BLUR) {
        if constexpr (EDGE) {  // clamp-to-edge columns (GaussianBlur.cpp:240)
            const float first = __shfl_sync(FULL, f[0], 1);
            const float last = __shfl_sync(FULL, f[NPX - 1], min(geo.lane_last, 31));
            if (geo.left_edge && lane == 0) {
#pragma unroll
                for (int j = 0; j < NPX; j++) f[j] = first;
            }
            if (geo.right_edge && lane > geo.lane_last) {
#pragma unroll
                for (int j = 0; j < NPX; j++) f[j] = last;
            }
        }
        geo.ring_cur += kRow;
        if (geo.ring_cur == geo.ring + 5 * kRow + 4 + NPX * lane) geo.ring_cur -= 5 * kRow;
#pragma unroll
        for (int j = 0; j < NPX; j += 4) *reinterpret_cast<float4 *>(geo.ring_cur + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        float c[NPX + 4];
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            c[j + 2] = fmaf(p.g2, f[j], st.a0[j]);
            st.a0[j] = fmaf(p.g1, f[j], st.a1[j]);
            st.a1[j] = fmaf(p.g0, f[j], st.a2[j]);
            st.a2[j] = fmaf(p.g1, f[j], st.a3[j]);
            st.a3[j] = p.g2 * f[j];
        }
        c[0] = __shfl_up_sync(FULL, c[NPX], 1);
        c[1] = __shfl_up_sync(FULL, c[NPX + 1], 1);
        c[NPX + 2] = __shfl_down_sync(FULL, c[2], 1);
        c[NPX + 3] = __shfl_down_sync(FULL, c[3], 1);
        float d[NPX];
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const float e2 = c[j] + c[j + 4], e1 = c[j + 1] + c[j + 3];
            const float u = fmaf(p.g2, e2, fmaf(p.g1, e1, fmaf(p.g0, c[j + 2], -0.5f)));  // S~ - 0.5
            const float rr = u + kMagic;
            b[j] = rr - kMagic;
            d[j] = fabsf(u - b[j]);
        }
        float dm = d[0];
#pragma unroll
        for (int j = 1; j < NPX; j++) dm = fmaxf(dm, d[j]);
        if (__builtin_expect(__any_sync(FULL, dm > p.thr), 0)) {
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < NPX; j++) mask |= (d[j] > p.thr ? 1u : 0u) << j;
            FN<NPX> bv;
#pragma unroll
            for (int j = 0; j < NPX; j++) bv.v[j] = b[j];
            const int slot = (int)(geo.ring_cur - (geo.ring + 4 + NPX * lane)) / kRow;
            bv = blur_exact_n<NPX>(geo.ring, slot, p.w, bv, mask, lane);
#pragma unroll
            for (int j = 0; j < NPX; j++) b[j] = bv.v[j];
            if (p.slow_counter && mask) atomicAdd(p.slow_counter, (unsigned long long)__popc(mask));
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPX; j++) b[j] = f[j];
    }

    // ---- Sobel partial sums of row yb, BORDER_REFLECT_101 in x -----------------------------------
    float e[NPX + 2];
    e[0] = __shfl_up_sync(FULL, b[NPX - 1], 1);
    e[NPX + 1] = __shfl_down_sync(FULL, b[0], 1);
#pragma unroll
    for (int j = 0; j < NPX; j++) e[j + 1] = b[j];
    if constexpr (EDGE) {
        if (geo.left_edge && lane == 1) e[0] = b[1];
        if (geo.right_edge && lane == geo.lane_last) e[NPX + 1] = b[NPX - 2];
    }
    float Dc[NPX], Sc[NPX];
#pragma unroll
    for (int j = 0; j < NPX; j++) {
        Dc[j] = e[j + 2] - e[j];
        Sc[j] = fmaf(2.f, e[j + 1], e[j] + e[j + 2]);
    }
    if constexpr (SPECIAL) {  // BORDER_REFLECT_101 in y
        if (yb == 1) {
#pragma unroll
            for (int j = 0; j < NPX; j++) { st.X0[j] = fmaf(2.f, st.X1[j], Dc[j]); st.S2[j] = Sc[j]; }
        }
        if (yb == H) {
#pragma unroll
            for (int j = 0; j < NPX; j++) { Dc[j] = fmaf(-2.f, st.X1[j], st.X0[j]); Sc[j] = st.S2[j]; }
        }
    }
    if constexpr (STORE) {
        uint32_t q[NPX];
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const float gx = st.X0[j] + Dc[j];
            const float gy = Sc[j] - st.S2[j];
            const float m = sqrt_approx(fmaf(gx, gx, gy * gy));
            q[j] = __float_as_uint(fminf(m, 255.f) + kMagic);
        }
        const uint32_t lo = __byte_perm(__byte_perm(q[0], q[1], 0x0040), __byte_perm(q[2], q[3], 0x0040), 0x5410);
        if constexpr (NPX == 8) {
            const uint32_t hi = __byte_perm(__byte_perm(q[4], q[5], 0x0040), __byte_perm(q[6], q[7], 0x0040), 0x5410);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.v2.u32 [%0], {%1, %2};\n\t}"
                         ::"l"(geo.dst), "r"(lo), "r"(hi), "r"(geo.store_lane) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
                         ::"l"(geo.dst), "r"(lo), "r"(geo.store_lane) : "memory");
        }
    }
#pragma unroll
    for (int j = 0; j < NPX; j++) {
        st.X0[j] = fmaf(2.f, Dc[j], st.X1[j]);
        st.X1[j] = Dc[j];
        st.S2[j] = st.S1[j];
        st.S1[j] = Sc[j];
    }
    geo.dst += W;
}

template <int NPX, int CN, bool BGR, bool BLUR, bool EDGE>
__device__ __forceinline__ void run_segment_t(WarpStateT<NPX> &st, const TmaParams &tp, const CUtensorMap *map, GeometryT &geo, int r,
                                              int r_last)
{
    constexpr int HALO = BLUR ? 3 : 1;
    const int r_store = geo.ys + HALO;
#pragma unroll 1
    for (; r < r_store; r++) step_t<NPX, CN, BGR, BLUR, EDGE, false, false>(st, tp, map, geo, r);
    step_t<NPX, CN, BGR, BLUR, EDGE, true, true>(st, tp, map, geo, r);
    r++;
#pragma unroll 1
    for (; r < r_last; r++) step_t<NPX, CN, BGR, BLUR, EDGE, true, false>(st, tp, map, geo, r);
    if (r == r_last) step_t<NPX, CN, BGR, BLUR, EDGE, true, true>(st, tp, map, geo, r);
}

template <int NPX, int CN, bool BGR, bool BLUR>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
fused_tma_kernel(const __grid_constant__ TmaParams tp, const __grid_constant__ CUtensorMap map)
{
    constexpr int HALO = BLUR ? 3 : 1;
    constexpr int kRow = 32 * NPX + 8;
    constexpr int kBand = 30 * NPX;
    const FusedParams &p = tp.f;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5;
    const int tile_bytes = NST * tp.stage_words * 4;  // multiple of 128
    GeometryT geo;
    geo.lane = threadIdx.x & 31;
    geo.tiles = reinterpret_cast<const uint32_t *>(smem + warp * tile_bytes);
    geo.ring = reinterpret_cast<float *>(smem + kWarpsPerBlock * tile_bytes) + warp * (BLUR ? 5 * kRow : 0);
    geo.ring_cur = geo.ring + 4 + NPX * geo.lane;
    geo.bars = reinterpret_cast<uint64_t *>(smem + kWarpsPerBlock * tile_bytes + (BLUR ? kWarpsPerBlock * 5 * kRow * 4 : 0)) + warp * NST;

    int bid = blockIdx.x;
    const int bg = bid % p.n_band_groups; bid /= p.n_band_groups;
    const int seg = bid % p.n_segs;
    const int frame = bid / p.n_segs;
    const int band = bg * kWarpsPerBlock + warp;
    if (band >= p.n_bands) return;  // warp-uniform; all synchronisation below is per warp

    const int W = p.W;
    const int xw = band * kBand - NPX;                // first pixel of lane 0 (a halo lane)
    const int x = xw + NPX * geo.lane;
    const bool in_img = (x >= 0) && (x < W);
    geo.lane_last = (W - band * kBand) / NPX;
    geo.left_edge = (band == 0);
    geo.right_edge = (geo.lane_last <= 31);
    geo.ys = p.out_row0 + seg * p.seg_rows;
    const int ye = min(geo.ys + p.seg_rows, p.out_row0 + p.out_rows);
    geo.store_lane = ((geo.lane >= 1) && (geo.lane <= 30) && in_img) ? 1u : 0u;
    const int r_first = geo.ys - HALO, r_last = ye - 1 + HALO;
    const int t0 = min(max(r_first - p.in_row0, 0), p.in_rows - 1);   // band-relative source row of tile 0
    // box start: the band start rounded down to a 16-byte boundary of the row (16 px for 3-byte pixels)
    const int box_px = CN == 3 ? (xw & ~15) : xw;
    geo.tma_x = (box_px * CN) / 4;
    geo.lane_word = ((xw - box_px) * CN) / 4 + (NPX * CN / 4) * geo.lane;
    geo.tma_y0 = frame * p.in_rows + t0;
    geo.k_cur = 0;
    geo.row_in_tile = 0;
    geo.src = geo.tiles + geo.lane_word;
    geo.dst = p.out + (size_t)frame * p.out_rows * W + (ptrdiff_t)(r_first - HALO - p.out_row0) * W + x;

    if (geo.lane == 0) {
        for (int s = 0; s < NST; s++) mbar_init(geo.bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < NST; s++) {
            mbar_expect_tx(geo.bars + s, (uint32_t)(TR * tp.tile_words * 4));
            tma_load_2d(const_cast<uint32_t *>(geo.tiles) + s * tp.stage_words, &map, geo.tma_x, geo.tma_y0 + s * TR, geo.bars + s);
        }
    }
    __syncwarp();
    mbar_wait(geo.bars, 0);

    WarpStateT<NPX> st;
#pragma unroll
    for (int j = 0; j < NPX; j++)
        st.a0[j] = st.a1[j] = st.a2[j] = st.a3[j] = st.X0[j] = st.X1[j] = st.S1[j] = st.S2[j] = 0.f;

    if (geo.left_edge || geo.right_edge) run_segment_t<NPX, CN, BGR, BLUR, true>(st, tp, &map, geo, r_first, r_last);
    else run_segment_t<NPX, CN, BGR, BLUR, false>(st, tp, &map, geo, r_first, r_last);

    // drain: TMA loads still in flight target this block's shared memory; wait for them before exit
    for (int kk = geo.k_cur + 1; kk < geo.k_cur + NST; kk++) mbar_wait(geo.bars + (kk % NST), (uint32_t)((kk / NST) & 1));
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool fused_supported(int W, int H, int fmt, int ksize, const uint8_t *d_in, const uint8_t *d_out)
{
    if (ksize != 0 && ksize != 5) return false;
    if (W < 4 || (W & 3) || H < 2) return false;
    const int cn = (fmt == RIP_FMT_RGB8 || fmt == RIP_FMT_BGR8) ? 3 : (fmt == RIP_FMT_RGBA8 || fmt == RIP_FMT_BGRA8) ? 4 : 0;
    if (cn == 0) return false;
    const uintptr_t in_align = cn == 4 ? 15u : 3u;
    if ((reinterpret_cast<uintptr_t>(d_in) & in_align) || (reinterpret_cast<uintptr_t>(d_out) & 3u)) return false;
    if (getenv("RIP_DISABLE_FUSED")) return false;
    return true;
}

// Guard band for the fast path, and the separable taps that minimise it.  See DESIGN.md
// ("Exact blur at separable cost") for the derivation:
//   |S_ref - S| <= u * 255 * (sum_i w_i (25 - i) + sum_i w_i)      (sequential fp32 sum, u = 2^-24)
//   |S~    - S| <= 255 * sum|w_ij - g_i g_j|  +  9 u * 255 * sum_i w_i   (separable FMA evaluation)  The kernel compares |frac(S~) - 0.5| against 0.5 - band.
static bool plan_weights_band(const float *w25, float g[3], double *band_out)
{
    double sum = 0.0;
    for (int i = 0; i < 25; i++) {
        if (!(w25[i] >= 0.0f) || !std::isfinite(w25[i])) return false;
        // the x2 kernel feeds gray in as q * 2^-149 and carries the 2^149 in the weights; a product must
        // stay a normal float for its rounding to equal the reference's (rip_fused_x2.cuh)
        if (w25[i] != 0.0f && w25[i] < 8.470329472543003e-22f /* 2^-70 */) return false;
        sum += (double)w25[i];
    }
    if (!(sum > 0.0) || 255.0 * sum >= 255.9) return false;  // floor(S) must stay <= 255
    // symmetric separable fit from the diagonal: g_k = sqrt(w[k][k])
    double gd[3];
    for (int k = 0; k < 3; k++) gd[k] = std::sqrt((double)w25[(2 + k) * 5 + (2 + k)]);
    for (int k = 0; k < 3; k++) g[k] = (float)gd[k];
    double dev = 0.0;
    for (int ky = -2; ky <= 2; ky++)
        for (int kx = -2; kx <= 2; kx++)
            dev += std::fabs((double)w25[(ky + 2) * 5 + (kx + 2)] - (double)g[std::abs(ky)] * (double)g[std::abs(kx)]);
    // reference error: acc_k = fl(acc_{k-1} + fl(p_k w_k)); each add errs by <= u*|acc_k| and
    // acc_k <= 255 * (w_0 + .. + w_k), so the adds contribute <= u * 255 * sum_i w_i * (25 - i);
    // the 25 rounded products add <= u * 255 * sum(w).  Fast path: <= 9 u * 255 * sum(w).
    const double u = std::ldexp(1.0, -24);
    double cum = 0.0;
    for (int i = 0; i < 25; i++) cum += (double)w25[i] * (25 - i);
    const double band = 255.0 * dev + u * 255.0 * (cum + sum + 9.0 * sum) * 1.02 + 1e-6;
    if (band > 0.05) return false;  // weights are not (close to) a symmetric separable kernel
    *band_out = band;
    return true;
}

bool fused_plan_weights(const float *w25, float g[3], float *thr)
{
    double band;
    if (!plan_weights_band(w25, g, &band)) return false;
    *thr = (float)(0.5 - band);
    return true;
}

static int pick_seg_rows(int out_rows, int n_frames, int n_band_groups, int device, int resident_blocks = 6)
{
    if (const char *e = getenv("RIP_FUSED_SEG")) {
        const int v = atoi(e);
        if (v > 0) return v < out_rows ? v : out_rows;
    }
    // enough blocks for >= ~3 waves of (SMs x resident blocks), but segments of >= 64 rows so the
    // 6 warm-up rows stay below 10 %; never more than 256 rows (tail balance).  (Measured on 32 4K
    // frames: 64 rows 440 us, 128 rows 416 us, 270 rows 415 us.)
    const long long target_blocks = (long long)sm_count(device) * resident_blocks * 3;
    int seg = 256;
    while (seg > 64 && (long long)n_frames * n_band_groups * ((out_rows + seg - 1) / seg) < target_blocks) seg >>= 1;
    if (seg > out_rows) seg = out_rows;
    return seg;
}

static unsigned long long *g_slow_counter = nullptr;  // set by rip_debug_slow_path_stats

// d_gray_down, per device, filled once by evaluating the reference expression (Comparator.cpp:41)
// in double on the host for the 16 774 (r,g,b) triples whose 299r+587g+114b is a multiple of 1000.
static int ensure_gray_table(int device)
{
    static std::mutex mu;
    static bool done[64];
    if (device < 0 || device >= 64) return fail(RIP_EINVAL, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(mu);
    if (done[device]) return RIP_OK;
    static uint32_t bits[2048];
    static bool built = false;
    if (!built) {
        memset(bits, 0, sizeof(bits));
        for (int r = 0; r < 256; r++)
            for (int g = 0; g < 256; g++)
                for (int b = 0; b < 256; b++) {
                    const int t = 299 * r + 587 * g + 114 * b;
                    if (t % 1000) continue;
                    volatile double s = 0.299 * r;  // volatile: no contraction, strict left-to-right doubles
                    volatile double s2 = 0.587 * g;
                    volatile double s3 = 0.114 * b;
                    volatile double sum = s + s2;
                    sum = sum + s3;
                    if ((int)sum < t / 1000) bits[(r << 8 | g) >> 5] |= 1u << ((r << 8 | g) & 31);
                }
        built = true;
    }
    RIP_CUDA(cudaMemcpyToSymbol(d_gray_down, bits, sizeof(bits)));
    done[device] = true;
    return RIP_OK;
}

// ---- TMA variant -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        else
            cudaGetLastError();
    });
    return fn;
}

static bool tma_usable(int W, int cn, const uint8_t *d_in, const uint8_t *d_out)
{
    // Measured on B200 (profiles/): the TMA-staged variant retires ~13 % more instructions per row
    // (shared-memory reads + ring bookkeeping) and both variants are issue-bound, so the register /
    // shuffle kernel with direct global loads is the default.  RIP_FUSED_TMA=1 selects the TMA kernel.
    const char *want = getenv("RIP_FUSED_TMA");
    if (!want || atoi(want) == 0) return false;
    if ((W & 3) || W < 4) return false;                         // 4-pixel lanes
    if (((size_t)W * cn) & 15u) return false;                   // TMA global stride: multiple of 16 bytes
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 7u)) return false;
    return encode_tiled_fn() != nullptr;
}

template <int NPX, int CN, bool BGR>
static cudaError_t launch_tma_t(bool blur, dim3 grid, size_t smem, cudaStream_t s, const TmaParams &tp, const CUtensorMap &map)
{
    auto kern = blur ? fused_tma_kernel<NPX, CN, BGR, true> : fused_tma_kernel<NPX, CN, BGR, false>;
    if (smem > 48 * 1024) {  // per device and cheap, so set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, kWarpsPerBlock * 32, smem, s>>>(tp, map);
    return cudaSuccess;
}

template <int NPX>
static int launch_fused_tma_n(cudaStream_t s, FusedParams p, int n_frames, int fmt, bool with_blur, int device)
{
    const int cn = (fmt == RIP_FMT_RGB8 || fmt == RIP_FMT_BGR8) ? 3 : 4;
    constexpr int kBand = 30 * NPX, kRow = 32 * NPX + 8;
    TmaParams tp;
    // RGB: 32*NPX px of lanes + up to 12 px of start rounding, rounded up to 16 px; RGBA: exactly the lanes
    const int box_px = cn == 3 ? ((32 * NPX + 12 + 15) / 16) * 16 : 32 * NPX;
    tp.tile_words = box_px * cn / 4;
    tp.stage_words = ((TR * tp.tile_words * 4 + 127) / 128) * 128 / 4;
    p.n_bands = (p.W + kBand - 1) / kBand;
    p.n_band_groups = (p.n_bands + kWarpsPerBlock - 1) / kWarpsPerBlock;
    p.seg_rows = pick_seg_rows(p.out_rows, n_frames, p.n_band_groups, device);
    p.n_segs = (p.out_rows + p.seg_rows - 1) / p.seg_rows;
    tp.f = p;

    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)p.W * cn / 4, (cuuint64_t)n_frames * p.in_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)p.W * cn};
    const cuuint32_t box[2] = {(cuuint32_t)tp.tile_words, (cuuint32_t)TR};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode_tiled_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(p.in), gdim, gstride, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(RIP_EINVAL, "rip_fused: cuTensorMapEncodeTiled failed (%d)", (int)cr);

    const size_t smem = (size_t)kWarpsPerBlock * NST * tp.stage_words * 4 + (with_blur ? (size_t)kWarpsPerBlock * 5 * kRow * 4 : 0) +
                        (size_t)kWarpsPerBlock * NST * 8;
    const long long blocks = (long long)n_frames * p.n_segs * p.n_band_groups;
    if (blocks <= 0 || blocks > 0x7fffffffLL) return fail(RIP_EINVAL, "rip_fused: grid of %lld blocks is out of range", blocks);
    const dim3 grid((unsigned)blocks);
    cudaError_t e = cudaSuccess;
    switch (fmt) {
    case RIP_FMT_RGB8:  e = launch_tma_t<NPX, 3, false>(with_blur, grid, smem, s, tp, map); break;
    case RIP_FMT_BGR8:  e = launch_tma_t<NPX, 3, true>(with_blur, grid, smem, s, tp, map); break;
    case RIP_FMT_RGBA8: e = launch_tma_t<NPX, 4, false>(with_blur, grid, smem, s, tp, map); break;
    case RIP_FMT_BGRA8: e = launch_tma_t<NPX, 4, true>(with_blur, grid, smem, s, tp, map); break;
    default: return fail(RIP_EINVAL, "rip_fused: unsupported input format %d", fmt);
    }
    if (e != cudaSuccess) return cuda_fail(e, "fused_tma_kernel setup", __FILE__, __LINE__);
    RIP_LAUNCH_CHECK();
    return RIP_OK;
}

static int launch_fused_tma(cudaStream_t s, const FusedParams &p, int n_frames, int fmt, bool with_blur, int device)
{
    const char *e = getenv("RIP_FUSED_NPX");
    if (e && atoi(e) == 8 && (p.W & 7) == 0) return launch_fused_tma_n<8>(s, p, n_frames, fmt, with_blur, device);
    return launch_fused_tma_n<4>(s, p, n_frames, fmt, with_blur, device);
}

// ---- x2 kernel (the default) --------------------------------------------------------------------
template <int NPX, int CN, bool BGR>
static void launch_x2_t(bool blur, dim3 grid, cudaStream_t s, const X2Params &xp)
{
    if (blur) fused_x2_kernel<NPX, CN, BGR, true><<<grid, kWarpsPerBlock * 32, 0, s>>>(xp);
    else fused_x2_kernel<NPX, CN, BGR, false><<<grid, kWarpsPerBlock * 32, 0, s>>>(xp);
}

template <int NPX>
static int launch_fused_x2_n(cudaStream_t s, FusedParams p, int n_frames, int fmt, bool with_blur, double band, const float g[3], int device)
{
    constexpr int kBand = 30 * NPX;
    X2Params xp;
    memset(&xp, 0, sizeof(xp));
    p.n_bands = (p.W + kBand - 1) / kBand;
    p.n_band_groups = (p.n_bands + kWarpsPerBlock - 1) / kWarpsPerBlock;
    p.seg_rows = pick_seg_rows(p.out_rows, n_frames, p.n_band_groups, device, NPX == 8 ? RIP_X2_MINB8 : RIP_X2_MINB4);
    p.n_segs = (p.out_rows + p.seg_rows - 1) / p.seg_rows;
    xp.f = p;
    if (with_blur) {
        // gray enters the vertical pass as an integer bit pattern (value q * 2^-149): the vertical
        // taps carry 2^75 and the horizontal taps 2^74 (powers of two: every rounding is unchanged)
        xp.gv0 = std::ldexp(g[0], 75); xp.gv1 = std::ldexp(g[1], 75); xp.gv2 = std::ldexp(g[2], 75);
        xp.gh0 = std::ldexp(g[0], 74); xp.gh1 = std::ldexp(g[1], 74); xp.gh2 = std::ldexp(g[2], 74);
        // S~ + 256 is rounded to a multiple of ulp = 2^-15 (error <= ulp/2).  A pixel whose 15 fraction
        // bits are >= a and <= 2^15 - 1 - a has frac(S~) in [(a - 1/2) ulp, 1 - (a + 1/2) ulp], i.e. S~ is
        // >= band away from an integer when a >= band / ulp + 1/2; all others are replayed exactly.
        const double ulp = std::ldexp(1.0, -kFracBits);
        const uint32_t a = (uint32_t)std::ceil(band / ulp + 0.5);
        xp.zoff = a << (32 - kFracBits);
        xp.zthr = (2u * a) << (32 - kFracBits);
    }
    const long long blocks = (long long)n_frames * p.n_segs * p.n_band_groups;
    if (blocks <= 0 || blocks > 0x7fffffffLL) return fail(RIP_EINVAL, "rip_fused: grid of %lld blocks is out of range", blocks);
    const dim3 grid((unsigned)blocks);
    switch (fmt) {
    case RIP_FMT_RGB8:  launch_x2_t<NPX, 3, false>(with_blur, grid, s, xp); break;
    case RIP_FMT_BGR8:  launch_x2_t<NPX, 3, true>(with_blur, grid, s, xp); break;
    case RIP_FMT_RGBA8: launch_x2_t<NPX, 4, false>(with_blur, grid, s, xp); break;
    case RIP_FMT_BGRA8: launch_x2_t<NPX, 4, true>(with_blur, grid, s, xp); break;
    default: return fail(RIP_EINVAL, "rip_fused: unsupported input format %d", fmt);
    }
    RIP_LAUNCH_CHECK();
    return RIP_OK;
}

// which kernel runs: RIP_FUSED_KERNEL = x2 (default) | ldg | tma; RIP_FUSED_NPX = 8 | 4 pixels per lane
static int x2_npx(int W, int cn, const uint8_t *d_in, const uint8_t *d_out)
{
    const char *k = getenv("RIP_FUSED_KERNEL");
    if (k && strcmp(k, "x2") != 0) return 0;
    if (getenv("RIP_FUSED_TMA") && atoi(getenv("RIP_FUSED_TMA")) != 0) return 0;
    int want = 8;
    if (const char *e = getenv("RIP_FUSED_NPX")) want = atoi(e) == 4 ? 4 : 8;
    const uintptr_t in_align8 = cn == 4 ? 15u : 7u;
    const bool ok8 = (W & 7) == 0 && !(reinterpret_cast<uintptr_t>(d_in) & in_align8) && !(reinterpret_cast<uintptr_t>(d_out) & 7u);
    if (want == 8 && ok8) return 8;
    return 4;  // fused_supported() already guarantees W % 4 == 0 and the 4-pixel alignments
}

template <int CN, bool BGR>
static void launch_t(bool blur, dim3 grid, cudaStream_t s, const FusedParams &p)
{
    if (blur) fused_kernel<CN, BGR, true><<<grid, kWarpsPerBlock * 32, 0, s>>>(p);
    else fused_kernel<CN, BGR, false><<<grid, kWarpsPerBlock * 32, 0, s>>>(p);
}

int launch_fused(cudaStream_t s, const uint8_t *d_in, uint8_t *d_out, int W, int H, int n_frames, int fmt,
                 bool with_blur, const float *weights25, int in_row0, int in_rows, int out_row0, int out_rows,
                 int device)
{
    if (int rc = ensure_gray_table(device)) return rc;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.out = d_out; p.W = W; p.H = H;
    p.in_row0 = in_row0; p.in_rows = in_rows; p.out_row0 = out_row0; p.out_rows = out_rows;
    p.n_bands = (W + kBandPx - 1) / kBandPx;
    p.n_band_groups = (p.n_bands + kWarpsPerBlock - 1) / kWarpsPerBlock;
    p.seg_rows = pick_seg_rows(out_rows, n_frames, p.n_band_groups, device);
    p.n_segs = (out_rows + p.seg_rows - 1) / p.seg_rows;
    p.slow_counter = g_slow_counter;
    double band = 0.0;
    float g[3] = {0.f, 0.f, 0.f};
    if (with_blur) {
        if (!plan_weights_band(weights25, g, &band))
            return fail(RIP_EUNSUPPORTED, "rip_fused: weights are not a non-negative symmetric separable 5x5 kernel");
        p.thr = (float)(0.5 - band);
        p.g0 = g[0]; p.g1 = g[1]; p.g2 = g[2];
        memcpy(p.w, weights25, sizeof(float) * 25);
    }
    {
        const int cn = (fmt == RIP_FMT_RGB8 || fmt == RIP_FMT_BGR8) ? 3 : 4;
        const int npx = x2_npx(W, cn, d_in, d_out);
        if (npx == 8) return launch_fused_x2_n<8>(s, p, n_frames, fmt, with_blur, band, g, device);
        if (npx == 4) return launch_fused_x2_n<4>(s, p, n_frames, fmt, with_blur, band, g, device);
        if (tma_usable(W, cn, d_in, d_out)) return launch_fused_tma(s, p, n_frames, fmt, with_blur, device);
    }
    const long long blocks = (long long)n_frames * p.n_segs * p.n_band_groups;
    if (blocks <= 0 || blocks > 0x7fffffffLL) return fail(RIP_EINVAL, "rip_fused: grid of %lld blocks is out of range", blocks);
    const dim3 grid((unsigned)blocks);
    switch (fmt) {
    case RIP_FMT_RGB8:  launch_t<3, false>(with_blur, grid, s, p); break;
    case RIP_FMT_BGR8:  launch_t<3, true>(with_blur, grid, s, p); break;
    case RIP_FMT_RGBA8: launch_t<4, false>(with_blur, grid, s, p); break;
    case RIP_FMT_BGRA8: launch_t<4, true>(with_blur, grid, s, p); break;
    default: return fail(RIP_EINVAL, "rip_fused: unsupported input format %d", fmt);
    }
    RIP_LAUNCH_CHECK();
    return RIP_OK;
}

void fused_set_slow_counter(unsigned long long *d_counter) { g_slow_counter = d_counter; }

// ---------------------------------------------------------------------------------------------
// device self-test of the two arithmetic shortcuts the fused kernel relies on, exhaustively:
//   (1) min(255, rint(sqrt.approx(m2))) == min(255, rint(sqrt_rn(m2))) for every reachable
//       m2 = gx^2 + gy^2 <= 2 * 1020^2;
//   (2) the dp2a / mad.wide gray path == gray_exact() for all 2^24 (c0,c1,c2) triples, for the RGB
//       and RGBA packers in both channel orders.
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void selftest_sqrt_kernel(unsigned long long *bad)
{
    const unsigned m2_max = 2u * 1020u * 1020u;
    for (unsigned m2 = blockIdx.x * blockDim.x + threadIdx.x; m2 <= m2_max; m2 += gridDim.x * blockDim.x) {
        const float q = fminf(sqrt_approx((float)m2), 255.f) + kMagic;
        const unsigned got = __float_as_uint(q) & 0xffu;
        const unsigned want = (unsigned)min(__float2int_rn(__fsqrt_rn((float)m2)), 255);
        if (got != want) atomicAdd(bad, 1ull);
    }
}

template <int CN, bool BGR>
__global__ void selftest_gray_kernel(unsigned long long *bad)
{
    // thread q handles triples 4q..4q+3 (triple index i: c0 = i & 255, c1 = (i >> 8) & 255, c2 = i >> 16)
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < (1u << 22); q += gridDim.x * blockDim.x) {
        uint8_t bytes[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned i = 4 * q + j;
            bytes[CN * j + 0] = i & 255u; bytes[CN * j + 1] = (i >> 8) & 255u; bytes[CN * j + 2] = i >> 16;
            if (CN == 4) bytes[4 * j + 3] = (uint8_t)(i * 37u);  // alpha must be ignored
        }
        uint32_t w[CN];
#pragma unroll
        for (int k = 0; k < CN; k++)
            w[k] = bytes[4 * k] | (bytes[4 * k + 1] << 8) | (bytes[4 * k + 2] << 16) | ((uint32_t)bytes[4 * k + 3] << 24);
        float f[4];
        gray4<CN, BGR>(w, f);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned i = 4 * q + j;
            const unsigned c0 = i & 255u, c1 = (i >> 8) & 255u, c2 = i >> 16;
            const unsigned want = BGR ? gray_exact(c2, c1, c0) : gray_exact(c0, c1, c2);
            if (f[j] != (float)want) atomicAdd(bad, 1ull);
        }
    }
}

// the x2 kernel's versions of the same two shortcuts: (1) sqrt.approx, then a multiply by 2^-149 (or
// 2^-127 on the 2^-22-scaled values of the no-blur variant) whose denormal result IS the rounded
// integer, then I2IP.U8.S32.SAT; (2) IDP.2A + FMUL2.RM / FFMA2.RP gray with the cold (r,g) lookup.
__global__ void selftest_sqrt_x2_kernel(unsigned long long *bad)
{
    const unsigned m2_max = 2u * 1020u * 1020u;
    const float s44 = __uint_as_float((127u - 44u) << 23);  // 2^-44
    for (unsigned m2 = blockIdx.x * blockDim.x + threadIdx.x; m2 <= m2_max; m2 += gridDim.x * blockDim.x) {
        const unsigned want = (unsigned)min(__float2int_rn(__fsqrt_rn((float)m2)), 255);
        const u64 a = mul2(pk2(sqrt_approx((float)m2), sqrt_approx((float)m2 * s44)),
                           pk2(__uint_as_float(1u), __uint_as_float(0x00400000u)));
        const uint32_t packed = i2ip(hi2u(a), lo2u(a), 0u);
        if ((packed & 0xffu) != want || ((packed >> 8) & 0xffu) != want) atomicAdd(bad, 1ull);
    }
}

template <int NPX, int CN, bool BGR>
__global__ void selftest_gray_x2_kernel(unsigned long long *bad)
{
    constexpr int NP = NPX / 2, NW = NPX * CN / 4;
    __shared__ uint32_t table[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) table[i] = d_gray_down[i];
    __syncthreads();
    const uint32_t table_s = (uint32_t)__cvta_generic_to_shared(table);
    // thread q handles triples NPX*q .. NPX*q + NPX-1 (triple i: c0 = i & 255, c1 = (i >> 8) & 255, c2 = i >> 16)
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < (1u << 24) / NPX; q += gridDim.x * blockDim.x) {
        uint8_t bytes[NW * 4];
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const unsigned i = NPX * q + j;
            bytes[CN * j + 0] = i & 255u; bytes[CN * j + 1] = (i >> 8) & 255u; bytes[CN * j + 2] = i >> 16;
            if (CN == 4) bytes[4 * j + 3] = (uint8_t)(i * 37u);  // alpha must be ignored
        }
        uint32_t w[NW];
#pragma unroll
        for (int k = 0; k < NW; k++)
            w[k] = bytes[4 * k] | (bytes[4 * k + 1] << 8) | (bytes[4 * k + 2] << 16) | ((uint32_t)bytes[4 * k + 3] << 24);
        u64 Q[NP], E[NP];
        const uint32_t any = gray_x2<NPX, CN, BGR>(w, Q, E);
        if (any & 1u) gray_fix_x2<NPX, CN, BGR>(w, Q, E, table_s);
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const unsigned i = NPX * q + j;
            const unsigned c0 = i & 255u, c1 = (i >> 8) & 255u, c2 = i >> 16;
            const unsigned want = BGR ? gray_exact(c2, c1, c0) : gray_exact(c0, c1, c2);
            const unsigned got = j < NP ? lo2u(Q[j % NP]) : hi2u(Q[j % NP]);
            if (got != want) atomicAdd(bad, 1ull);
        }
    }
}

}  // namespace

int fused_selftest(int device, unsigned long long *checked, unsigned long long *mismatches)
{
    if (int rc = ensure_gray_table(device)) return rc;
    unsigned long long *d_bad = nullptr;
    RIP_CUDA(cudaMalloc(&d_bad, sizeof(*d_bad)));
    RIP_CUDA(cudaMemset(d_bad, 0, sizeof(*d_bad)));
    const int grid = sm_count(device) * 8;
    selftest_sqrt_kernel<<<grid, 256>>>(d_bad);
    selftest_gray_kernel<3, false><<<grid, 256>>>(d_bad);
    selftest_gray_kernel<3, true><<<grid, 256>>>(d_bad);
    selftest_gray_kernel<4, false><<<grid, 256>>>(d_bad);
    selftest_gray_kernel<4, true><<<grid, 256>>>(d_bad);
    selftest_sqrt_x2_kernel<<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<8, 3, false><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<8, 3, true><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<8, 4, false><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<8, 4, true><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<4, 3, false><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<4, 3, true><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<4, 4, false><<<grid, 256>>>(d_bad);
    selftest_gray_x2_kernel<4, 4, true><<<grid, 256>>>(d_bad);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    unsigned long long bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess) return cuda_fail(e, "fused_selftest", __FILE__, __LINE__);
    count_launch(14);
    *checked = 3ull * (2ull * 1020ull * 1020ull + 1ull) + 12ull * (1ull << 24);
    *mismatches = bad;
    return RIP_OK;
}

}  // namespace rip
