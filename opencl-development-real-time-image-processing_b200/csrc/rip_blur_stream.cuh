// rip_blur_stream.cuh -- the 5x5 RGBA Gaussian as a streaming kernel (the shape of the fused kernel: every warp is
// independent and slides down a segment of rows), for the stand-alone blur's common case (BASELINE config 2).
// Same arithmetic contract as blur_sep_kernel (rip_blur_sep.cu): separable fp32 sum on top of a bias of 256,
// mantissa-bit guard band, exact replay of the pixels inside it, bit-exact against GaussianBlur.cpp:231-261.
//
//   * a lane owns 2 horizontally adjacent pixels = 4 packed pairs (channels 0-1 and 2-3 of each pixel), so every
//     horizontal tap of a pair is again an aligned pair: blur runs on FFMA2 without pair construction;
//   * a warp covers 64 columns: lanes 1..30 produce 60 output columns, lanes 0 and 31 are the 2-pixel halo.
//     Clamp-to-edge (GaussianBlur.cpp:240-241) is applied when a pixel is LOADED (column and row indices are
//     clamped), so the halo lanes of an edge band simply hold copies of the edge column;
//   * vertical pass first, in accumulate form (row r completes output row r-2: four pending sums per pair in
//     registers); the completed row goes through a double-buffered 1 KB shared row so that a lane can read its
//     neighbours' pairs (2 STS.128 + 4 LDS.128 per lane per row instead of 16 shuffles); horizontal pass on top
//     of the bias;
//   * the raw pixels of the last 5 rows stay in a shared ring for the exact replay (lane-local: 25 taps x 4
//     channels in the reference's order).
// Per pixel: ~45 instructions against ~140 for the tiled kernel.  Included by rip_blur_sep.cu inside its
// anonymous namespace (shares SepParams, the guard-band constants and plan_sep_blur's bound: the vertical chain
// here plays the role of the tiled kernel's horizontal chain and vice versa, the bound is symmetric in them).

typedef unsigned long long bs_u64;

__device__ __forceinline__ bs_u64 bs_pk2(float lo, float hi) { bs_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ uint32_t bs_lo(bs_u64 v) { uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ uint32_t bs_hi(bs_u64 v) { uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ bs_u64 bs_fma2(bs_u64 a, bs_u64 b, bs_u64 c) { bs_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ bs_u64 bs_mul2(bs_u64 a, bs_u64 b) { bs_u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ bs_u64 bs_add2_rm(bs_u64 a, bs_u64 b) { bs_u64 d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ bs_u64 bs_pk2u(uint32_t lo, uint32_t hi) { bs_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

constexpr int kBsWarps = 4;          // warps per block, each an independent band
constexpr int kBsBand = 60;          // output columns per warp
constexpr int kBsRowB = 32 * 32;     // bytes of one shared V row (32 lanes x 4 pairs)

// Two channels of a packed pixel as a pair of INTEGER BIT PATTERNS: the isolated byte q, read as a float, is the denormal
// q * 2^-149, which the vertical FMAs consume at full rate; the vertical taps carry 2^75 and the horizontal taps 2^74
// (powers of two: every rounding is unchanged), so no conversion instruction runs at all (round 1: PRMT + FADD per channel).
__device__ __forceinline__ bs_u64 bs_cvt2(uint32_t px, uint32_t sel_lo, uint32_t sel_hi)
{
    return bs_pk2u(__byte_perm(px, 0u, sel_lo), __byte_perm(px, 0u, sel_hi));
}

// floor of the four channels of one pixel packed into a u32.  The biased sums lie in [256, 512); adding 2^23 - 256 rounded toward
// minus infinity leaves 2^23 + floor(sum - 256), whose low byte is the result: one packed add per channel pair, then a byte gather
// (round 1: a shift per channel before the gather).
__device__ __forceinline__ uint32_t bs_pack(bs_u64 c01, bs_u64 c23)
{
    const float m = 8388608.0f - kSepBias;
    const bs_u64 M = bs_pk2(m, m), t01 = bs_add2_rm(c01, M), t23 = bs_add2_rm(c23, M);
    return __byte_perm(__byte_perm(bs_lo(t01), bs_hi(t01), 0x0040), __byte_perm(bs_lo(t23), bs_hi(t23), 0x0040), 0x5410);
}

// the reference's sequence for one pixel (GaussianBlur.cpp:236-258): column `col` of the warp's ring, rows oldest first
__device__ __noinline__ uint32_t bs_replay(uint32_t ring, uint32_t cur_slot, uint32_t col, const Weights &wts, unsigned long long *slow_counter)
{
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (uint32_t ky = 0; ky < 5u; ky++) {
        uint32_t slot = cur_slot + 1u + ky;
        slot = slot >= 5u ? slot - 5u : slot;
        for (uint32_t kx = 0; kx < 5u; kx++) {
            uint32_t px;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(px) : "r"(ring + slot * 256u + 4u * (col + kx - 2u)));
            const float w = wts.w[ky * 5u + kx];
            a0 = __fadd_rn(a0, __fmul_rn((float)(px & 0xffu), w));
            a1 = __fadd_rn(a1, __fmul_rn((float)((px >> 8) & 0xffu), w));
            a2 = __fadd_rn(a2, __fmul_rn((float)((px >> 16) & 0xffu), w));
            a3 = __fadd_rn(a3, __fmul_rn((float)(px >> 24), w));
        }
    }
    if (slow_counter) atomicAdd(slow_counter, 1ull);
    return (uint32_t)__float2int_rz(fminf(fmaxf(a0, 0.f), 255.f)) | ((uint32_t)__float2int_rz(fminf(fmaxf(a1, 0.f), 255.f)) << 8) |
           ((uint32_t)__float2int_rz(fminf(fmaxf(a2, 0.f), 255.f)) << 16) | ((uint32_t)__float2int_rz(fminf(fmaxf(a3, 0.f), 255.f)) << 24);
}

// The fix of one lane-row, out of line: (1) which channels are constant over the 5 rows x 6 columns around the lane's two pixels?
// Their result is flat[value] whatever the fast sum says -- the fast sum of a constant window sits on an integer, inside the guard
// band, so the alpha channel of every real RGBA frame (255 throughout) and black or clipped regions flag every pixel.  (2) Only a
// pixel with a flagged channel that is NOT constant runs the 25-tap replay.  Before round 2's end every flagged pixel ran the replay for its four channels:
// 16 1080p frames with alpha = 255 took 977 us against 219 us for frames whose alpha is noise.
// Returns the patched outputs in .x / .y and in .z which pixels still need the replay (bit 0 / bit 1); the caller runs it, so that
// the call depth -- and with it the kernel's register allocation -- stays what it was.
__device__ __noinline__ uint3 bs_fix(uint32_t ring, uint32_t cur_slot, uint32_t lane, uint32_t fm /* flagged channels: bits 0-3 pixel 0, bits 4-7 pixel 1 */,
                                     const uint8_t *flat, uint32_t o0, uint32_t o1)
{
    const uint32_t a0 = ring + 8u * lane - 8u;   // columns 2 lane - 2 .. 2 lane + 3 of a ring row (lanes 1..30: inside the row)
    uint32_t ref, diff = 0;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ref) : "r"(a0 + cur_slot * 256u + 8u));
#pragma unroll 1
    for (uint32_t sl = 0; sl < 5u; sl++) {   // (rolled: the function must not raise the kernel's register count)
        uint32_t w[6];
#pragma unroll
        for (int k = 0; k < 3; k++) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[2 * k]), "=r"(w[2 * k + 1]) : "r"(a0 + sl * 256u + 8u * k));
#pragma unroll
        for (int k = 0; k < 6; k++) diff |= w[k] ^ ref;
    }
    uint32_t cb = 0, cexact = 0, cbits = 0;   // 0xff per constant channel, those channels' exact bytes, one bit per constant channel
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (((diff >> (8 * c)) & 0xffu) == 0u) {
            cb |= 0xffu << (8 * c);
            cbits |= 1u << c;
            cexact |= (uint32_t)flat[(ref >> (8 * c)) & 0xffu] << (8 * c);
        }
    uint32_t need = 0;
    if (fm & 0xfu & ~cbits) need |= 1u;
    else if (fm & 0xfu) o0 = (o0 & ~cb) | cexact;
    if ((fm >> 4) & ~cbits) need |= 2u;
    else if (fm >> 4) o1 = (o1 & ~cb) | cexact;
    return make_uint3(o0, o1, need);
}

struct StreamGeo {
    int seg_rows, n_segs, n_band_groups;
};

__global__ void __launch_bounds__(kBsWarps * 32)
blur_stream5_kernel(const __grid_constant__ SepParams p, const __grid_constant__ Weights wts, const StreamGeo sg)
{
    __shared__ __align__(16) uint32_t ring_s[kBsWarps][5][64];        // raw pixels of the last 5 rows, per warp
    __shared__ __align__(16) uint32_t vrow_s[kBsWarps][2][kBsRowB / 4];  // completed vertical sums, double-buffered

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int bid = blockIdx.x;
    const int bg = bid % sg.n_band_groups; bid /= sg.n_band_groups;
    const int seg = bid % sg.n_segs;
    const int frame = bid / sg.n_segs;
    const int band = bg * kBsWarps + (int)warp;
    const int xw0 = band * kBsBand;
    if (xw0 >= p.W) return;   // warp-uniform; no block-level barrier anywhere in this kernel

    const int ys = p.out_row0 + seg * sg.seg_rows, ye = min(ys + sg.seg_rows, p.out_row0 + p.out_rows);
    const int x0 = xw0 - 2 + 2 * (int)lane;                       // this lane's two pixels: x0, x0 + 1
    const int c0 = min(max(x0, 0), p.W - 1), c1 = min(max(x0 + 1, 0), p.W - 1);   // clamp-to-edge columns
    const uint8_t *fsrc = p.src + (size_t)frame * p.src_rows * p.W * 4;
    uint8_t *fdst = p.dst + (size_t)frame * p.out_rows * p.W * 4;
    const bool store0 = lane >= 1u && lane <= 30u && x0 < p.W, store1 = lane >= 1u && lane <= 30u && x0 + 1 < p.W;

    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(&ring_s[warp][0][0]);
    const uint32_t vrow = (uint32_t)__cvta_generic_to_shared(&vrow_s[warp][0][0]);
    // vertical taps * 2^75 (the pixels enter as q * 2^-149), horizontal taps * 2^74, and the bias with the guard band's lower edge
    // (a ulps of 2^-15, exact) riding in it: a channel is inside the band iff its fraction bits are below 2a, i.e.
    // (bits << 17) < zthr, and the masked value of every other channel is floor(S~).  All scaled on the host (launch_blur_sep), so
    // they reach the packed FMAs as scalar operands from the constant bank instead of occupying twenty registers.
    const bs_u64 G0 = bs_pk2(p.sgv[0], p.sgv[0]), G1 = bs_pk2(p.sgv[1], p.sgv[1]), G2 = bs_pk2(p.sgv[2], p.sgv[2]), G3 = bs_pk2(p.sgv[3], p.sgv[3]),
                 G4 = bs_pk2(p.sgv[4], p.sgv[4]);
    const bs_u64 H0 = bs_pk2(p.sgh[0], p.sgh[0]), H1 = bs_pk2(p.sgh[1], p.sgh[1]), H2 = bs_pk2(p.sgh[2], p.sgh[2]), H3 = bs_pk2(p.sgh[3], p.sgh[3]),
                 H4 = bs_pk2(p.sgh[4], p.sgh[4]);
    const bs_u64 BIAS = bs_pk2(p.sbias, p.sbias);

    // Five vertical accumulators per pair, one per output row in flight (slot = output row mod 5, counted from the
    // first row of the segment); the row loop is unrolled by five so that every accumulator is updated in place
    // (a shifting a0 <- a1 <- a2 <- a3 delay line costs 25 register moves per row in a rolled loop).
    bs_u64 acc[5][4];
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[k][j] = 0ull;

    // source rows: clamp to the image (GaussianBlur.cpp:241), then to the rows the band holds (read-ahead only).
    // The clamped index of row r + 1 is one more than that of row r iff row_lo <= r < row_hi, so the two
    // per-lane pointers just advance under that test.
    const int row_lo = max(0, p.src_row0), row_hi = min(p.H - 1, p.src_row0 + p.src_rows - 1);
    const uint32_t row_span = (uint32_t)max(row_hi - row_lo, 0);
    const uint32_t *pn0, *pn1;       // this lane's two pixels in the row loaded last
    {
        const int rr = min(max(min(max(ys - 2, 0), p.H - 1) - p.src_row0, 0), p.src_rows - 1);
        const uint32_t *row = reinterpret_cast<const uint32_t *>(fsrc + (size_t)rr * p.W * 4);
        pn0 = row + c0;
        pn1 = row + c1;
    }
    uint32_t n0 = __ldg(pn0), n1 = __ldg(pn1);   // the row loaded one step ahead
    uint32_t *orow = reinterpret_cast<uint32_t *>(fdst + (size_t)(ys - p.out_row0) * p.W * 4);

    // one row: PH = (r - (ys - 2)) mod 5; an output row lives in the slot of the phase at which it completes.
    // Row r is tap 4 of output r-2 (slot PH: completes now), tap 3 of r-1 (slot PH+1), tap 2 of r (PH+2), tap 1 of
    // r+1 (PH+3) and tap 0 of r+2 (slot PH+4, which completed one step ago: restart it).
    auto step = [&](auto ph_tag, int r) {
        constexpr int PH = decltype(ph_tag)::value;
        const uint32_t q0 = n0, q1 = n1;
        if ((uint32_t)(r - row_lo) < row_span) {   // row r + 1 is a new row (not a clamped repeat of row r)
            pn0 += p.W;
            pn1 += p.W;
        }
        n0 = __ldg(pn0);
        n1 = __ldg(pn1);
        // raw pixels into the ring (for the replay): ring slot = PH
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ring + PH * 256u + 8u * lane), "r"(q0), "r"(q1) : "memory");
        const bs_u64 Q[4] = {bs_cvt2(q0, 0x4440, 0x4441), bs_cvt2(q0, 0x4442, 0x4443), bs_cvt2(q1, 0x4440, 0x4441), bs_cvt2(q1, 0x4442, 0x4443)};
        bs_u64 V[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            V[j] = bs_fma2(G4, Q[j], acc[PH][j]);
            acc[(PH + 1) % 5][j] = bs_fma2(G3, Q[j], acc[(PH + 1) % 5][j]);
            acc[(PH + 2) % 5][j] = bs_fma2(G2, Q[j], acc[(PH + 2) % 5][j]);
            acc[(PH + 3) % 5][j] = bs_fma2(G1, Q[j], acc[(PH + 3) % 5][j]);
            acc[(PH + 4) % 5][j] = bs_mul2(G0, Q[j]);   // (that slot completed one step ago and is free)
        }
        const int y = r - 2;   // the output row this step completes
        if (y >= ys) {         // warp-uniform
            const uint32_t vb = vrow + (uint32_t)(r & 1) * kBsRowB;
            asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(vb + 32u * lane), "l"(V[0]), "l"(V[1]) : "memory");
            asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(vb + 32u * lane + 16u), "l"(V[2]), "l"(V[3]) : "memory");
            __syncwarp();
            bs_u64 L[4], R[4];   // the four pairs of the left / right neighbour lane
            const uint32_t la = vb + 32u * ((lane + 31u) & 31u), ra = vb + 32u * ((lane + 1u) & 31u);
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(L[0]), "=l"(L[1]) : "r"(la));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(L[2]), "=l"(L[3]) : "r"(la + 16u));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(R[0]), "=l"(R[1]) : "r"(ra));
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(R[2]), "=l"(R[3]) : "r"(ra + 16u));
            // pixel x0: taps x0-2 (L px0), x0-1 (L px1), x0 (own px0), x0+1 (own px1), x0+2 (R px0); channel pairs h = 0, 1
            bs_u64 F[4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                F[h] = bs_fma2(H4, R[h], bs_fma2(H3, V[2 + h], bs_fma2(H2, V[h], bs_fma2(H1, L[2 + h], bs_fma2(H0, L[h], BIAS)))));
                // pixel x0 + 1: taps x0-1 (L px1), x0 (own px0), x0+1 (own px1), x0+2 (R px0), x0+3 (R px1)
                F[2 + h] = bs_fma2(H4, R[2 + h], bs_fma2(H3, R[h], bs_fma2(H2, V[2 + h], bs_fma2(H1, V[h], bs_fma2(H0, L[2 + h], BIAS)))));
            }
            uint32_t z[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                z[2 * j] = bs_lo(F[j]) << (32 - kSepFracBits);
                z[2 * j + 1] = bs_hi(F[j]) << (32 - kSepFracBits);
            }
            const uint32_t zm0 = min(__vimin3_u32(z[0], z[1], z[2]), z[3]), zm1 = min(__vimin3_u32(z[4], z[5], z[6]), z[7]);
            uint32_t o0 = bs_pack(F[0], F[1]), o1 = bs_pack(F[2], F[3]);
            if (min(zm0, zm1) < p.zthr && store0) {   // lane-local fix (rare on textured content), only for pixels that are stored
                uint32_t fm = 0;   // (the bias carries the band's lower edge: inside iff the shifted fraction bits are below zthr)
#pragma unroll
                for (int k = 0; k < 8; k++) fm |= (z[k] < p.zthr ? 1u : 0u) << k;
                if (!store1) fm &= 0xfu;
                const uint3 o = bs_fix(ring, (uint32_t)PH, lane, fm, p.flat, o0, o1);
                o0 = (o.z & 1u) ? bs_replay(ring, (uint32_t)PH, 2u * lane, wts, p.slow_counter) : o.x;
                o1 = (o.z & 2u) ? bs_replay(ring, (uint32_t)PH, 2u * lane + 1u, wts, p.slow_counter) : o.y;
            }
            __syncwarp();   // the next step overwrites the ring's oldest row, which a replay above may still be reading
            if (store0) orow[x0] = o0;
            if (store1) orow[x0 + 1] = o1;
            orow += p.W;
        }
    };

    int r = ys - 2;
#pragma unroll 1
    for (; r + 4 <= ye + 1; r += 5) {
        step(std::integral_constant<int, 0>{}, r);
        step(std::integral_constant<int, 1>{}, r + 1);
        step(std::integral_constant<int, 2>{}, r + 2);
        step(std::integral_constant<int, 3>{}, r + 3);
        step(std::integral_constant<int, 4>{}, r + 4);
    }
    // the last 0..4 rows of the segment (phases 0.. in order)
    if (r <= ye + 1) step(std::integral_constant<int, 0>{}, r++);
    if (r <= ye + 1) step(std::integral_constant<int, 1>{}, r++);
    if (r <= ye + 1) step(std::integral_constant<int, 2>{}, r++);
    if (r <= ye + 1) step(std::integral_constant<int, 3>{}, r++);
}
