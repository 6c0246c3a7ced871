// rip_blur_streamk.cuh -- the KxK RGBA Gaussian for larger K (9x9, and 17x17 sigma 6: the reference's default,
// include/ProgramHandler.hpp:9) as a streaming kernel.  Same arithmetic contract as blur_sep_kernel (rip_blur_sep.cu): the fast
// value of a pixel/channel is bit-identical to the tiled kernel's (horizontal FMA chain from g[0], vertical chain on top of the
// bias, k ascending), so plan_sep_blur's bound holds unchanged; pixels inside the guard band are replayed with the reference's
// exact sequence (GaussianBlur.cpp:231-261).
//
//   * a warp is independent: it owns 32 output columns (one pixel = four channels = two packed pairs per lane) and slides down a
//     segment of rows; K - 1 warm-up rows per segment;
//   * per input row the 32 + K - 1 pixels of the row are converted once (integer bit patterns, no conversion instruction) and
//     staged in shared memory; a lane's horizontal sum reads K consecutive 16-byte pixels (conflict-free) -- 2 K FFMA2;
//   * the vertical pass runs in ACCUMULATE form: the row's horizontal sum is added into the K output rows in flight, which live
//     in registers as a delay line that moves by one slot per row -- acc[t] = fma(g, h, acc[t + 1]): the FMA's own destination does
//     the shift, so the row loop stays rolled (one copy of ~150 instructions; unrolled K times with every accumulator updated in
//     place, the 17x17 loop was 71 KB of code and the kernel waited for instruction fetches: 3.7 warps per issue slot stalled on
//     no_instruction, 810 us) -- 2 K FFMA2 and no shared-memory traffic at all.  The tiled kernel reads every operand of both passes from shared memory and recomputes the
//     horizontal pass for its halo rows (48 x 48 loaded, 48 x 32 filtered for a 32 x 32 tile): 526 lane-instructions per pixel;
//   * guard-band pixels are not replayed where they are found (a 289-step chain in one lane while 31 wait: with 1.5 % of the pixels
//     flagged, four warp-rows of ten would hold one) but appended, one (pixel, channel) per entry, to a per-warp list; whenever 32
//     entries are waiting the warp replays them side by side, reading the window from global memory (L2 hits: the rows were read
//     moments ago) as integer bit patterns against weights * 2^100;
//   * constant windows: per staged pixel and channel the kernel counts the consecutive rows in which the pixel equalled its right
//     neighbour and the pixel above (byte-parallel counters, two staged pixels per lane); eight ballots and a funnel shift per lane
//     turn them into "all K columns of my window have K such rows", i.e. the window is constant: those channels take flat[value] (the
//     reference's own sequence for a constant window, host-evaluated) and leave the guard-band test -- alpha = 255 of every real RGBA
//     frame, black sky, clipped highlights, and the dark plateaus of a decoded JPEG (the reference's Artemis_* images), which a
//     per-band tracker missed: 604 us instead of 121 us on Artemis_large1024, because every plateau pixel went through the replay in
//     the few warps that own those bands.
//
// Included by rip_blur_sep.cu inside its anonymous namespace, after rip_blur_stream.cuh (bs_* helpers, StreamGeo).

// byte C of m replicated... the top bit of byte C of m over all 32 bits (m holds 0x00 / 0xff bytes)
template <int C> __device__ __forceinline__ uint32_t bs_rep(uint32_t m)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(m), "n"(0x8888 + 0x1111 * C));
    return r;
}

constexpr int kSkWarps = 4;      // warps per block, each an independent band
constexpr int kSkBand = 32;      // output columns per warp
constexpr int kSkList = 160;     // list entries per warp: < 32 waiting + at most 128 new ones per row

// The reference's value of one channel of one pixel per lane, windows read from global memory.  item: x | (y - ys) << 16 | ch << 30.
template <int K>
__device__ __noinline__ void sk_replay(const SepParams &p, const float *rw /* weights * 2^100 */, const uint8_t *fsrc, uint8_t *fdst, int ys, uint32_t item,
                                       bool valid)
{
    constexpr int HALF = K / 2;
    __syncwarp();   // the fast values of these pixels were stored by other lanes
    const int x = (int)(item & 0xffffu), y = ys + (int)((item >> 16) & 0x3fffu);
    const uint32_t ch = item >> 30;
    const bool interior = x - HALF >= 0 && x + HALF <= p.W - 1;
    float a = 0.f;
    if (__all_sync(0xffffffffu, interior || !valid)) {
        if (valid) {
#pragma unroll 1
            for (int ky = 0; ky < K; ky++) {
                const uint8_t *row = fsrc + ((size_t)(clampi(y + ky - HALF, 0, p.H - 1) - p.src_row0) * p.W + (x - HALF)) * 4 + ch;
                const float *wr = rw + ky * K;
#pragma unroll
                for (int kx = 0; kx < K; kx++) a = __fadd_rn(a, __fmul_rn(__uint_as_float((uint32_t)__ldg(row + 4 * kx)), wr[kx]));
            }
        }
    } else if (valid) {
#pragma unroll 1
        for (int ky = 0; ky < K; ky++) {
            const uint8_t *row = fsrc + (size_t)(clampi(y + ky - HALF, 0, p.H - 1) - p.src_row0) * p.W * 4 + ch;
            const float *wr = rw + ky * K;
#pragma unroll 1
            for (int kx = 0; kx < K; kx++)
                a = __fadd_rn(a, __fmul_rn(__uint_as_float((uint32_t)__ldg(row + 4 * clampi(x + kx - HALF, 0, p.W - 1))), wr[kx]));
        }
    }
    if (valid) {
        // (uchar)clamp(sum, 0, 255), GaussianBlur.cpp:255-258; the sum carries 2^-49 and is >= 0
        fdst[((size_t)(y - p.out_row0) * p.W + x) * 4 + ch] = (uint8_t)__float2int_rz(fminf(a * 562949953421312.0f, 255.f));
    }
}

// The cold block of one row, out of line (the call site must stay small: the row loop is unrolled K times and lives in the
// instruction cache): append the flagged (pixel, channel) pairs of this row to the warp's list and replay while 32 are waiting.
// A channel whose fast value is 0 needs nothing: the true sum is >= 0 and below 1.  Returns the new count.
template <int K>
__device__ __noinline__ uint32_t sk_cold(const SepParams &p, const float *rw, int frame, int ys, uint32_t list, uint32_t cnt, uint32_t item0 /* x | (y - ys) << 16 */,
                                         uint32_t fm /* flagged channels */)
{
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (uint32_t c = 0; c < 4u; c++) {
        const bool f = (fm >> c) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (f) asm volatile("st.shared.u32 [%0], %1;" ::"r"(list + 4u * (cnt + (uint32_t)__popc(b & ((1u << lane) - 1u)))), "r"(item0 | (c << 30)) : "memory");
        cnt += (uint32_t)__popc(b);
    }
    __syncwarp();
    if (cnt >= 32u) {
        const uint8_t *fsrc = p.src + (size_t)frame * p.src_rows * p.W * 4;
        uint8_t *fdst = p.dst + (size_t)frame * p.out_rows * p.W * 4;
#pragma unroll 1
        while (cnt >= 32u) {
            cnt -= 32u;
            uint32_t item;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(item) : "r"(list + 4u * (cnt + lane)));
            if (p.slow_counter && lane == 0) atomicAdd(p.slow_counter, 32ull);
            sk_replay<K>(p, rw, fsrc, fdst, ys, item, true);
        }
    }
    return cnt;
}

template <int K>
__global__ void __launch_bounds__(kSkWarps * 32, K <= 9 ? 5 : 4)
blur_streamk_kernel(const __grid_constant__ SepParams p, const __grid_constant__ Weights rws, const StreamGeo sg)
{
    constexpr int HALF = K / 2, SW = kSkBand + 2 * HALF;               // staged pixels per row
    __shared__ __align__(16) uint32_t stage_s[kSkWarps][64 * 4];       // the converted row: 16 bytes (two pairs) per pixel; SW are used, lanes >= 2 HALF park their B pixel behind
    __shared__ uint32_t list_s[kSkWarps][kSkList];
    __shared__ float rw_s[K * K];   // the reference's weights * 2^100 for the replay (a generic load from the parameter bank per tap was its critical path)
    __shared__ __align__(4) uint8_t flat_s[256];   // flat[v] (a lookup by pixel value: per-lane addresses, which the constant bank serialises)
    for (int i = threadIdx.x; i < K * K; i += kSkWarps * 32) rw_s[i] = rws.w[i];
    for (int i = threadIdx.x; i < 64; i += kSkWarps * 32) reinterpret_cast<uint32_t *>(flat_s)[i] = reinterpret_cast<const uint32_t *>(p.flat)[i];
    __syncthreads();   // (the only block-level barrier: before any warp leaves)

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    int bid = blockIdx.x;
    const int bg = bid % sg.n_band_groups; bid /= sg.n_band_groups;
    const int seg = bid % sg.n_segs;
    const int frame = bid / sg.n_segs;
    const int xw0 = (bg * kSkWarps + (int)warp) * kSkBand;
    if (xw0 >= p.W) return;   // warp-uniform; no block-level barrier anywhere in this kernel

    const int ys = p.out_row0 + seg * sg.seg_rows, ye = min(ys + sg.seg_rows, p.out_row0 + p.out_rows);
    const int x = xw0 + (int)lane;
    const bool store = x < p.W;
    const uint8_t *fsrc = p.src + (size_t)frame * p.src_rows * p.W * 4;
    uint8_t *fdst = p.dst + (size_t)frame * p.out_rows * p.W * 4;
    const uint32_t list = (uint32_t)__cvta_generic_to_shared(&list_s[warp][0]);
    uint32_t st = (uint32_t)__cvta_generic_to_shared(&stage_s[warp][0]) + 16u * lane;   // this lane's first staged pixel
    asm volatile("" : "+r"(st));

    // staged columns: A = xw0 - HALF + lane (all lanes), B = A + 32 (lanes < 2 HALF), clamp-to-edge (GaussianBlur.cpp:240-241)
    const int ca = clampi(xw0 - HALF + (int)lane, 0, p.W - 1), cb = clampi(xw0 - HALF + 32 + (int)lane, 0, p.W - 1);
    const int row_lo = max(0, p.src_row0), row_hi = min(p.H - 1, p.src_row0 + p.src_rows - 1);
    const uint32_t row_span = (uint32_t)max(row_hi - row_lo, 0);
    const uint32_t *pa, *pb;
    {
        const int rr = min(max(min(max(ys - HALF, 0), p.H - 1) - p.src_row0, 0), p.src_rows - 1);
        const uint32_t *row = reinterpret_cast<const uint32_t *>(fsrc + (size_t)rr * p.W * 4);
        pa = row + ca;
        pb = row + ((int)lane < 2 * HALF ? cb : ca);
    }
    // rows are loaded two steps ahead
    uint32_t na = __ldg(pa), nb = __ldg(pb);
    {
        const size_t adv = (uint32_t)(ys - HALF - row_lo) < row_span ? (size_t)(uint32_t)p.W : (size_t)0;
        pa += adv;
        pb += adv;
    }
    uint32_t ma = __ldg(pa), mb = __ldg(pb);
    uint32_t *po = reinterpret_cast<uint32_t *>(fdst + (size_t)(ys - p.out_row0) * p.W * 4) + x;

    bs_u64 acc[K - 1][2];   // acc[t]: the output row that completes t + 1 rows from now
#pragma unroll
    for (int k = 0; k < K - 1; k++) acc[k][0] = acc[k][1] = 0ull;
    const bs_u64 BIAS = bs_pk2(p.sbias, p.sbias);

    // The constant-window tracker, per staged pixel and channel (one byte each): the number of consecutive rows, ending at the current
    // one, in which the pixel equalled its right neighbour and the pixel above it; saturates at 127.  K such rows in each of the K
    // columns of a window make the window constant.  (Until the end of round 2 the tracker was per BAND -- all 32 + K - 1 pixels of a row
    // one value: fine for alpha and for regions wider than the band, but the dark plateaus of a decoded JPEG are narrower.  On the
    // reference's Artemis_large1024 image, 5 % of whose 17x17 windows are constant and not black, every such pixel went through the
    // 289-tap replay, all in the few warps that own those bands: 604 us against 86 us for the tiled kernel.)
    uint32_t pva = na, pvb = nb, runA = 0, runB = 0, cnt = 0;
    // the pixel word the exact bytes `cex` were last looked up for (valid from the start: a sentinel such as ~na collides with a channel whose value
    // is the complement of the first row's -- 0 then 255 -- and leaves 0 in the output: found by tests/soak.py)
    uint32_t cexsrc = na, cex = (uint32_t)flat_s[na & 0xffu] | ((uint32_t)flat_s[(na >> 8) & 0xffu] << 8) | ((uint32_t)flat_s[(na >> 16) & 0xffu] << 16) |
                                ((uint32_t)flat_s[na >> 24] << 24);
    const int r_last = ye - 1 + HALF;
#pragma unroll 1
    for (int r = ys - HALF; r <= r_last; r++) {
        const uint32_t qa = na, qb = nb;
        na = ma;
        nb = mb;
        {   // row r + 2 is a new row (not a clamped repeat of row r + 1) iff row_lo <= r + 1 < row_hi
            const size_t adv = (uint32_t)(r + 1 - row_lo) < row_span ? (size_t)(uint32_t)p.W : (size_t)0;
            pa += adv;
            pb += adv;
        }
        ma = __ldg(pa);
        mb = __ldg(pb);
        // tracker: which channels of this lane's window (staged columns lane .. lane + K - 1, rows r - K + 1 .. r) are constant
        {   // (the counters run every row; the window test -- eight ballots -- only when some lane has a channel inside the band, below)
            const uint32_t ra = __shfl_down_sync(0xffffffffu, qa, 1), rb0 = __shfl_sync(0xffffffffu, qb, 0), rbn = __shfl_down_sync(0xffffffffu, qb, 1);
            const uint32_t right_a = lane == 31u ? rb0 : ra;   // staged column 32 is lane 0's B pixel
            const uint32_t ea = (qa ^ right_a) | (qa ^ pva), eb = (qb ^ rbn) | (qb ^ pvb);   // (B pixels of lanes >= 2 HALF - 1 are never looked at)
            pva = qa;
            pvb = qb;
            runA += 0x01010101u;
            runA -= (runA >> 7) & 0x01010101u;
            runA &= ~bs_nzb(ea);
            runB += 0x01010101u;
            runB -= (runB >> 7) & 0x01010101u;
            runB &= ~bs_nzb(eb);
        }
        // stage the converted pixels (integer bit patterns: q * 2^-149; the taps carry the powers of two back)
        const uint32_t sb = st;
        asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(sb), "l"(bs_cvt2(qa, 0x4440, 0x4441)), "l"(bs_cvt2(qa, 0x4442, 0x4443)) : "memory");
        asm volatile("st.shared.v2.b64 [%0+512], {%1, %2};" ::"r"(sb), "l"(bs_cvt2(qb, 0x4440, 0x4441)), "l"(bs_cvt2(qb, 0x4442, 0x4443)) : "memory");
        __syncwarp();
        // horizontal pass: pixel x reads staged pixels lane .. lane + K - 1
        bs_u64 h0, h1;
        {
            bs_u64 v0, v1;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(sb));
            const bs_u64 g = bs_pk2(p.sg1[0], p.sg1[0]);
            h0 = bs_mul2(g, v0);
            h1 = bs_mul2(g, v1);
        }
#pragma unroll
        for (int k = 1; k < K; k++) {
            bs_u64 v0, v1;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(sb + 16u * k));
            const bs_u64 g = bs_pk2(p.sg1[k], p.sg1[k]);
            h0 = bs_fma2(g, v0, h0);
            h1 = bs_fma2(g, v1, h1);
        }
        __syncwarp();   // (one staging buffer: the next row's stores wait for these loads)
        // vertical pass, accumulate form: row r is tap K-1 of output row r - HALF (acc[0]: completes now), tap K-2-t of the output row
        // that moves from acc[t + 1] to acc[t], and tap 0 of output row r + HALF, which starts on the bias in acc[K - 2]
        bs_u64 f0, f1;
        {
            const bs_u64 g = bs_pk2(p.sg2[K - 1], p.sg2[K - 1]);
            f0 = bs_fma2(g, h0, acc[0][0]);
            f1 = bs_fma2(g, h1, acc[0][1]);
        }
#pragma unroll
        for (int t = 0; t < K - 2; t++) {
            const bs_u64 g = bs_pk2(p.sg2[K - 2 - t], p.sg2[K - 2 - t]);
            acc[t][0] = bs_fma2(g, h0, acc[t + 1][0]);
            acc[t][1] = bs_fma2(g, h1, acc[t + 1][1]);
        }
        {
            const bs_u64 g = bs_pk2(p.sg2[0], p.sg2[0]);
            acc[K - 2][0] = bs_fma2(g, h0, BIAS);
            acc[K - 2][1] = bs_fma2(g, h1, BIAS);
        }
        if (r - HALF >= ys) {   // warp-uniform (false: one of the K - 1 warm-up rows of the segment)
            uint32_t z0 = bs_lo(f0) << (32 - kSepFracBits), z1 = bs_hi(f0) << (32 - kSepFracBits), z2 = bs_lo(f1) << (32 - kSepFracBits),
                     z3 = bs_hi(f1) << (32 - kSepFracBits);
            uint32_t o = bs_pack(f0, f1);
            // alpha: a fast sum equal to that of an all-255 window proves the window IS all 255 (plan_streamk_alpha, the argument of the 5x5
            // kernel): exact value from the table, out of the guard-band test -- the alpha channel of every frame the reference uploads
            if (bs_hi(f1) == p.f255) {
                z3 = 0xffffffffu;
                o = (o & 0x00ffffffu) | p.a255;
            }
            // (a channel whose fast value is 0 needs no fix: black regions -- colour bytes all 0, alpha out of the band -- do not get any further)
            bool flag = min(__vimin3_u32(z0, z1, z2), z3) < p.zthr && store && !((o & 0x00ffffffu) == 0u && z3 >= p.zthr);
            const bool any = __any_sync(0xffffffffu, flag);
            if (any) {   // warp-uniform: which channels of this lane's window (staged columns lane .. lane + K - 1, rows r - K + 1 .. r) are constant?
                uint32_t cm = 0;   // 0xff per constant channel
                const uint32_t ga = runA + (uint32_t)(128 - K) * 0x01010101u, gb = runB + (uint32_t)(128 - K) * 0x01010101u;   // bit 7 of a byte: K good rows
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t A = __ballot_sync(0xffffffffu, (ga >> (8 * c + 7)) & 1u), B = __ballot_sync(0xffffffffu, (gb >> (8 * c + 7)) & 1u);
                    const uint32_t w = __funnelshift_r(A, B, lane);   // bit k: staged column lane + k
                    if ((~w & ((1u << K) - 1u)) == 0u) cm |= 0xffu << (8 * c);
                }
                if (cm) {   // constant channels take the table's value (any pixel of the window gives it: this lane's A pixel is its first column) and leave the test
                    if ((qa ^ cexsrc) & cm) {
                        cexsrc = qa;
                        cex = (uint32_t)flat_s[qa & 0xffu] | ((uint32_t)flat_s[(qa >> 8) & 0xffu] << 8) | ((uint32_t)flat_s[(qa >> 16) & 0xffu] << 16) |
                              ((uint32_t)flat_s[qa >> 24] << 24);
                    }
                    o = (o & ~cm) | (cex & cm);
                    z0 |= bs_rep<0>(cm);
                    z1 |= bs_rep<1>(cm);
                    z2 |= bs_rep<2>(cm);
                    z3 |= bs_rep<3>(cm);
                    flag = min(__vimin3_u32(z0, z1, z2), z3) < p.zthr && store && !((o & 0x00ffffffu) == 0u && z3 >= p.zthr);
                }
            }
            if (store) *po = o;   // (before the cold block: a replay below may patch bytes of this very row)
            po += p.W;
            if (any && __any_sync(0xffffffffu, flag)) {
                uint32_t fm = 0;   // flagged channels whose fast value is not 0
                if (flag)
                    fm = (z0 < p.zthr && (o & 0xffu) ? 1u : 0u) | (z1 < p.zthr && (o & 0xff00u) ? 2u : 0u) | (z2 < p.zthr && (o & 0xff0000u) ? 4u : 0u) |
                         (z3 < p.zthr && (o & 0xff000000u) ? 8u : 0u);
                cnt = sk_cold<K>(p, rw_s, frame, ys, list, cnt, (uint32_t)x | ((uint32_t)(r - HALF - ys) << 16), fm);
            }
        }
    }

    // what is still waiting in the list
    __syncwarp();
#pragma unroll 1
    while (cnt) {
        const uint32_t n = cnt < 32u ? cnt : 32u;
        cnt -= n;
        uint32_t item = 0;
        if (lane < n) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(item) : "r"(list + 4u * (cnt + lane)));
        if (p.slow_counter && lane == 0) atomicAdd(p.slow_counter, (unsigned long long)n);
        sk_replay<K>(p, rw_s, fsrc, fdst, ys, item, lane < n);
    }
}
