// rip_fused_x2.cuh -- the production fused kernel: gray -> 5x5 Gaussian -> 3x3 Sobel (or gray -> Sobel) in
// one pass, written around what tools/pipe_probe.cu measured on B200 (profiles/pipe_probe_r1.txt):
//
//   pipe          lanes/clk/SM   used here for
//   FMA  (fp32)   128            FFMA2/FADD2/FMUL2 (two pixels per instruction: half the issue slots,
//                                same pipe time), IDP.2A (64/clk)
//   ALU           64             LOP3, VIMNMX3, I2IP, LEA (LEA measured at 128/clk)
//   XU            16             MUFU.SQRT only (I2F/F2I/FRND and IMAD.WIDE/.HI are slow: none are used)
//   LSU/shuffle   32             LDG, SHFL, STS, STG
//
// The kernel is FMA-pipe bound (~26 fp32 lane-operations per pixel), so everything that is not a
// multiply-add was moved off that pipe or removed:
//   * gray: t = 299r+587g+114b by IDP.2A; the INTEGER bit pattern of t, read as a float, is the
//     denormal t*2^-149, so floor(t/1000) is ONE multiply rounded toward -inf by the float just above
//     1/1000 (FMUL2.RM), and its result is again an integer bit pattern that feeds the blur's FFMAs
//     directly (denormal operands run at full rate; the taps carry the 2^149 back).  No IMAD.WIDE,
//     no I2F.  A second FMA rounded toward +inf with the float just below 1/1000 yields 1 exactly
//     on the multiples of 1000, the only triples where the reference's double arithmetic
//     (Comparator.cpp:41) can land one below t/1000 (looked up per (r,g), cold path).
//   * blur: separable fp32 fast path; S~ + 256 puts floor(S~) in mantissa bits 15..22 and the
//     fraction in bits 0..14, so the guard band test is one LEA + VIMNMX3 per pixel and the rounded
//     value 256+b one LOP3 -- no float subtractions.  Pixels inside the guard band are replayed
//     with the reference's exact 25-tap sequence (GaussianBlur.cpp:236-258) from a shared-memory ring.
//   * Sobel runs on the biased values 256+b (all sums stay exact in fp32), the magnitude's
//     round-half-even + saturation is an FMUL2 by 2^-149 (the result's bit pattern IS the integer)
//     followed by I2IP.U8.S32.SAT, which also packs the bytes.
//   * a lane owns NPX horizontally adjacent pixels, pixel j paired with pixel j+NPX/2 in one 64-bit
//     register so that every horizontal tap of a pair is again an aligned pair; the row loop is
//     unrolled by NB over NB input-row buffers (loads run NB rows ahead; 3, or 6 without the blur stage), so the buffer rotation
//     and the two-row Sobel delay line are register renames, not moves.
//
// Included by rip_fused.cu inside its anonymous namespace (shares FusedParams, the gray table and the
// exact replay conventions with the older kernels kept there for A/B runs).

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 pk2u(uint32_t lo, uint32_t hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ float lo2(u64 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float hi2(u64 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ uint32_t lo2u(u64 v) { uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ uint32_t hi2u(u64 v) { uint32_t lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2_rm(u64 a, u64 b) { u64 d; asm("mul.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2_rp(u64 a, u64 b, u64 c) { u64 d; asm("fma.rp.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// d = (c & 0xffff) << 16 | sat_u8(a) << 8 | sat_u8(b)
__device__ __forceinline__ uint32_t i2ip(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr float kInvK_up = 1.0000000474974513e-3f;   // 0x3A83126F: the float just ABOVE 1/1000
constexpr float kInvK_dn = 9.9999993108212948e-4f;   // 0x3A83126E: the float just BELOW 1/1000
constexpr float kBias = 256.0f;                      // [256, 512): ulp 2^-15, floor(S~) in mantissa bits 15..22
constexpr uint32_t kBiasMask = 0xffff8000u;
constexpr int kFracBits = 15;

struct X2Params {
    FusedParams f;
    float gv0, gv1, gv2;   // vertical taps   * 2^75  (gray enters as the integer bit pattern q = q * 2^-149)
    float gh0, gh1, gh2;   // horizontal taps * 2^74
    uint32_t zoff, zthr;   // guard band: pixel is replayed iff ((bits << 19) + zoff) < zthr  (unsigned, mod 2^32)
};

// ---- exact gray of NPX packed pixels -> NPX/2 pairs of integer bit patterns ---------------------
// returns the OR of the "t is a multiple of 1000" flags in bit 0
template <int NPX, int CN, bool BGR>
__device__ __forceinline__ uint32_t gray_x2(const uint32_t *w, u64 *Q, u64 *E)
{
    constexpr int NP = NPX / 2;
    if constexpr (CN == 1) {
        // the input IS the gray image (GRAY8, or the luma plane of NV12): isolate the bytes; an isolated byte is
        // already the integer bit pattern the later stages consume.  No exactness cases here.
#pragma unroll
        for (int j = 0; j < NP; j++) {
            Q[j] = pk2u(__byte_perm(w[j / 4], 0u, 0x4440u | (uint32_t)(j & 3)), __byte_perm(w[(j + NP) / 4], 0u, 0x4440u | (uint32_t)((j + NP) & 3)));
            E[j] = 0ull;
        }
        return 0u;
    } else {
    constexpr uint32_t cA = BGR ? 114u : 299u, cB = 587u, cC = BGR ? 299u : 114u;  // weights of byte 0,1,2
    constexpr uint32_t AB = cA | (cB << 16), C0 = cC, zA = cA << 16, BC = cB | (cC << 16);
    uint32_t t[NPX];
#pragma unroll
    for (int g = 0; g < NPX / 4; g++) {
        const uint32_t *v = w + g * CN;
        if constexpr (CN == 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) t[4 * g + j] = __dp2a_hi(C0, v[j], __dp2a_lo(AB, v[j], 0u));  // alpha x 0
        } else {
            // byte stream: p0 = v0.b0-2, p1 = v0.b3 v1.b0-1, p2 = v1.b2-3 v2.b0, p3 = v2.b1-3
            t[4 * g + 0] = __dp2a_hi(C0, v[0], __dp2a_lo(AB, v[0], 0u));
            t[4 * g + 1] = __dp2a_lo(BC, v[1], __dp2a_hi(zA, v[0], 0u));
            t[4 * g + 2] = __dp2a_lo(C0, v[2], __dp2a_hi(AB, v[1], 0u));
            t[4 * g + 3] = __dp2a_hi(BC, v[2], __dp2a_lo(zA, v[2], 0u));
        }
    }
    const u64 up = pk2(kInvK_up, kInvK_up), ndn = pk2(-kInvK_dn, -kInvK_dn);
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const u64 T = pk2u(t[j], t[j + NP]);
        Q[j] = mul2_rm(T, up);        // floor(t * up) = floor(t / 1000)          (t <= 255000)
        E[j] = fma2_rp(T, ndn, Q[j]); // ceil(q - t * dn): 1 iff t = 1000 q > 0, else -0 / +0
        any |= lo2u(E[j]) | hi2u(E[j]);
    }
    return any;
    }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_b64(uint32_t addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

// The cold paths below never write a register the hot path reads: they patch a copy of the lane's pairs
// in shared memory (pair layout: word 2j = pixel j, word 2j+1 = pixel j + NPX/2) and the hot path
// reloads the pairs with loads predicated on "this lane was flagged".  A cold block that modified the
// hot registers in place would make every one of them a phi and cost ~30 register moves per row.
// STRIDE = bytes between consecutive pairs: 8 in the gray ring (dense), 16 in the scratch copies (so that
// ptxas cannot merge their 64-bit stores into 128-bit ones, which need consecutive registers and cost moves
// that it hoists onto the hot path)
template <int NP, int STRIDE>
__device__ __forceinline__ void reload_pairs_if(u64 *v, uint32_t addr, uint32_t flag)
{
    if constexpr (NP == 4) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.b64 %0, [%4];\n\t@p ld.shared.b64 %1, [%4+%6];\n\t"
                     "@p ld.shared.b64 %2, [%4+%7];\n\t@p ld.shared.b64 %3, [%4+%8];\n\t}"
                     : "+l"(v[0]), "+l"(v[1]), "+l"(v[2]), "+l"(v[3]) : "r"(addr), "r"(flag), "n"(STRIDE), "n"(2 * STRIDE), "n"(3 * STRIDE) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.b64 %0, [%2];\n\t@p ld.shared.b64 %1, [%2+%4];\n\t}"
                     : "+l"(v[0]), "+l"(v[1]) : "r"(addr), "r"(flag), "n"(STRIDE) : "memory");
    }
}

// byte offset of pixel j's word inside a lane's copy of NPX/2 pairs
template <int NPX, int STRIDE>
__device__ __forceinline__ uint32_t pair_off(uint32_t j) { return (j & (NPX / 2 - 1)) * STRIDE + (j / (NPX / 2)) * 4u; }

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

// Cold (inline on purpose: a call would make ptxas shuffle the hot registers that sit in the callee's
// argument registers, on the hot path): for the pixels flagged in `m` (t is a multiple of 1000) look up
// the (r,g)-indexed bit that says whether the reference's double evaluation lands one below t/1000, and
// decrement the copy of that pixel's gray at `pairs`.  `raw` is a shared-memory copy of the lane's
// input bytes, so a run-time pixel index needs no select chains.
template <int NPX, int CN, bool BGR, int STRIDE>
__device__ __forceinline__ void gray_patch(uint32_t m, uint32_t pairs, uint32_t raw, uint32_t table /* d_gray_down staged in shared memory */)
{
#pragma unroll 1
    while (m) {
        const uint32_t j = (uint32_t)__ffs(m) - 1u;
        m &= m - 1;
        const uint32_t r = lds_u8(raw + CN * j + (BGR ? 2 : 0)), g = lds_u8(raw + CN * j + 1);
        const uint32_t idx = (r << 8) | g;
        const uint32_t down = (lds_u32(table + 4u * (idx >> 5)) >> (idx & 31u)) & 1u;
        const uint32_t a = pairs + pair_off<NPX, STRIDE>(j);
        sts_u32(a, lds_u32(a) - down);
    }
}

// flagged-pixel mask of a lane from the E pairs of gray_x2 (bit j = pixel j).  Every half of E is 1
// (flagged), +0 or -0, so a shift-add chain collects the bits (the sign bit of a -0 either leaves the
// word or lands in bit 31, which the final mask drops).
template <int NPX>
__device__ __forceinline__ uint32_t gray_flag_mask(const u64 *E)
{
    constexpr int NP = NPX / 2;
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < NP; j++) m += (lo2u(E[j]) << j) + (hi2u(E[j]) << (j + NP));
    return m & ((1u << NPX) - 1u);
}

// register-in / register-out variant for the self-test (not used by the kernel); `table` = shared-memory
// address of a staged copy of d_gray_down
template <int NPX, int CN, bool BGR>
__device__ __forceinline__ void gray_fix_x2(const uint32_t *w, u64 *Q, const u64 *E, uint32_t table)
{
    constexpr int NP = NPX / 2, NW = NPX * CN / 4;
    __shared__ __align__(16) uint32_t scratch[256 * 3 * NPX];
    const uint32_t pairs = (uint32_t)__cvta_generic_to_shared(scratch + threadIdx.x * 3 * NPX), raw = pairs + 8 * NPX;
#pragma unroll
    for (int j = 0; j < NP; j++) sts_b64(pairs + 16 * j, Q[j]);
#pragma unroll
    for (int k = 0; k < NW; k++) sts_u32(raw + 4 * k, w[k]);
    gray_patch<NPX, CN, BGR, 16>(gray_flag_mask<NPX>(E), pairs, raw, table);
    reload_pairs_if<NP, 16>(Q, pairs, 1u);
}

template <int NPX, int CN>
struct RawX {
    uint32_t w[NPX * CN / 4];
};

// Unpredicated: lanes outside the image read the start of the row (their pointer has x offset 0) and
// their pixels are never consumed.  A predicated load would tie each destination register to its
// previous value and turn the rotation of the prefetch registers into moves.
template <int NPX, int CN>
__device__ __forceinline__ void load_row_x2(RawX<NPX, CN> &r, const uint8_t *p)
{
    constexpr int NW = NPX * CN / 4;
    if constexpr (NW == 1) {
        r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
    } else if constexpr (NW == 2) {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(p));
        r.w[0] = a.x; r.w[1] = a.y;
    } else if constexpr (NW == 3) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
        r.w[0] = __ldg(q); r.w[1] = __ldg(q + 1); r.w[2] = __ldg(q + 2);
    } else if constexpr (NW == 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    } else if constexpr (NW == 6) {
        const uint2 *q = reinterpret_cast<const uint2 *>(p);
        const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y; r.w[4] = c.x; r.w[5] = c.y;
    } else {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        const uint4 a = __ldg(q), b = __ldg(q + 1);
        r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
    }
}

template <int NPX, int CN>
struct WarpX {
    u64 a0[NPX / 2], a1[NPX / 2], a2[NPX / 2], a3[NPX / 2];  // pending vertical sums of blurred rows r-2 .. r+1
    u64 F1[NPX / 2], F2[NPX / 2];                            // rows yb-1 and yb-2 of the image the Sobel stage reads
};

struct GeoX {
    const uint8_t *src;      // this lane's pixels in the input row that was loaded last
    uint8_t *dst;            // this lane's pixels in the output row produced next
    uint32_t ring_warp;      // shared-memory byte address of this warp's gray ring [5][32*NPX] (integer gray;
                             // within a lane's NPX words the pixels sit in pair order, see reload_pairs_if)
    uint32_t ring_cur;       // byte address of this lane's words in the ring row holding the newest gray row
    uint32_t patch;          // byte address of this lane's NPX words of scratch for the cold paths
    uint32_t w25;            // shared-memory byte address of the exact 2-D weights times 2^100 (for the replay)
    uint32_t scratch;        // shared-memory byte address of this warp's 32-float scratch row (for the replay)
    uint32_t table;          // shared-memory byte address of the block's copy of d_gray_down
    uint32_t in_pitch;
    int adv_lo, adv_n;       // the source pointer advances before the load of step r iff 0 <= r - adv_lo < adv_n
    uint32_t pf_off;         // byte offset from src of the line this lane prefetches into L2 (0: none)
    int lane;
    int cmin, cmax;          // first / last column of the band (0 = pixel 0 of lane 0) that lies inside the image
    bool e_left, e_right;    // this lane holds image column 0 / W-1 (border rules in x apply to it)
    uint32_t store_lane;
    int r_store, r_last;     // first / last step that produces an output row
};

// Cold (inline, see gray_patch): exact replay of the pixels inside the guard band, warp-cooperatively.
// The whole warp is held up by one flagged pixel anyway, so all of it works on that pixel: lane t < 25
// fetches tap t (ky = t / 5, kx = t % 5) of the pixel's 5x5 gray window from the warp's ring and forms
// the reference's rounded product; the products go through a 32-float scratch row to the pixel's owner
// lane, which adds them in the reference's order (GaussianBlur.cpp:236-258: ky-major / kx-minor from 0.0f,
// unfused), clamps, truncates and writes kBias + b over its copy of the pixel in `patch`.
// The ring holds gray as integer bit patterns (= q * 2^-149 as floats) and `w25` the weights times 2^100:
// fl(q*2^-149 * w*2^100) = fl(q * w) * 2^-49 exactly (same mantissa, results stay normal), likewise every
// partial sum, so the scaled chain rounds exactly like the reference's and needs no integer-to-float
// conversion.  Columns are clamped to [cmin, cmax] (clamp-to-edge, GaussianBlur.cpp:240).
// STATS: count the replayed pixels (rip_debug_slow_path_stats); a separate instantiation so that the production
// kernel does not carry the counting code in its three copies of this block
template <int NPX, bool STATS>
__device__ __forceinline__ void blur_replay_warp(const u64 *F, uint32_t patch, uint32_t scratch, uint32_t ring_warp, uint32_t ring_cur,
                                                 uint32_t w25, int cmin, int cmax, uint32_t zoff, uint32_t zthr,
                                                 unsigned long long *slow_counter)
{
    constexpr int NP = NPX / 2;
    constexpr uint32_t kRowB = 32 * NPX * 4;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t my = 0;   // this lane's pixels inside the guard band (bit j = pixel j)
#pragma unroll
    for (int j = 0; j < NP; j++) {
        my |= (((lo2u(F[j]) << (32 - kFracBits)) + zoff) < zthr ? 1u : 0u) << j;
        my |= (((hi2u(F[j]) << (32 - kFracBits)) + zoff) < zthr ? 1u : 0u) << (j + NP);
    }
    // of the two halo lanes only the pixel next to the band feeds an output (through the Sobel halo exchange)
    if (lane == 0u) my &= 1u << (NPX - 1);
    if (lane == 31u) my &= 1u;
    uint32_t lanes = __ballot_sync(FULL, my != 0u);
    // this lane's tap: ring row of ky (oldest row first: the slot after the newest), column offset, weight
    const uint32_t t = lane < 25u ? lane : 24u, ky = t / 5u, kx = t - 5u * ky;
    uint32_t slot = (ring_cur - ring_warp) / kRowB + 1u + ky;
    slot = slot >= 5u ? slot - 5u : slot;
    const uint32_t row = ring_warp + slot * kRowB;
    const float wt = lds_f32(w25 + 4u * t);
    const int dx = (int)kx - 2;
#pragma unroll 1
    while (lanes) {
        const uint32_t src = (uint32_t)__ffs(lanes) - 1u;
        lanes &= lanes - 1u;
        uint32_t m = __shfl_sync(FULL, my, src);
#pragma unroll 1
        while (m) {
            const uint32_t j = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
            // A pixel whose S~ is exactly 0 (bit pattern of kBias) needs no replay: the taps are non-negative, so
            // every gray value under a non-zero tap is 0 and the reference's sum is 0 as well (black regions would
            // otherwise pay ~100 instructions per pixel).  Every lane reads the owner's copy of F: uniform.
            if (__all_sync(FULL, lds_u32(patch + (uint32_t)(((int)src - (int)lane) * (3 * NPX * 4)) + pair_off<NPX, 16>(j)) == __float_as_uint(kBias))) continue;
            const uint32_t cc = (uint32_t)min(max((int)(NPX * src + j) + dx, cmin), cmax);
            const uint32_t g = lds_u32(row + 4u * (cc & ~(uint32_t)(NPX - 1)) + pair_off<NPX, 8>(cc & (NPX - 1)));
            sts_u32(scratch + 4u * lane, __float_as_uint(__fmul_rn(__uint_as_float(g), wt)));
            __syncwarp();
            if (lane == src) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 24; k += 4) {
                    float p0, p1, p2, p3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p0), "=f"(p1), "=f"(p2), "=f"(p3) : "r"(scratch + 4 * k));
                    acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, p0), p1), p2), p3);
                }
                acc = __fadd_rn(acc, lds_f32(scratch + 4 * 24));
                acc = __fmul_rn(acc, 562949953421312.0f);   // * 2^49: back to the reference's scale (exact)
                sts_u32(patch + pair_off<NPX, 16>(j), __float_as_uint(kBias + truncf(fminf(fmaxf(acc, 0.f), 255.f))));
                if constexpr (STATS) {
                    if (slow_counter) atomicAdd(slow_counter, 1ull);
                }
            }
            __syncwarp();
        }
    }
}

// One image row of the sliding window: consumes the input row held in `buf` (row r, clamped to the
// rows of the band), refills `buf` with row r + NB (NB = row buffers, see X2Cfg), and -- for r_store <= r <= r_last -- stores output
// row r - HALO.  The border rules are applied at run time (per-lane selects in x, two rare uniform
// branches in y), so this is the only copy of the row body; the caller unrolls it by three with three
// row buffers, which makes the buffer rotation and the two-row Sobel delay line register renames.
// What a copy of the row body has to check at run time:
//   SPECIAL  head and tail rows of a segment: does this row store, do the next input row and the prefetched
//            line exist, BORDER_REFLECT_101 in y.  The main loop's copies check none of that.
//   EDGE     the warp's band holds image column 0 and/or W-1: per-lane border selects in x.  Interior warps
//            (most) run copies without them.
template <int NPX, int CN, bool BGR, bool BLUR, bool SPECIAL, bool EDGE, bool STATS>
__device__ __forceinline__ void step_x2(WarpX<NPX, CN> &st, RawX<NPX, CN> &buf, const X2Params &xp, GeoX &geo, int r)
{
    constexpr int NP = NPX / 2;
    constexpr int kRowB = 32 * NPX * 4;  // bytes per ring row
    const FusedParams &p = xp.f;
    const int W = p.W, H = p.H;
    (void)H;

    // ---- 1. gray of row r; refill the buffer with row r + NB --------------------------------------
    u64 Q[NP];
    {
        u64 E[NP];
        const uint32_t flagged = gray_x2<NPX, CN, BGR>(buf.w, Q, E) & 1u;
        uint32_t pairs = geo.patch;
        if constexpr (BLUR) {
            // park the integer gray row in the shared ring (the exact replay reads it back)
            geo.ring_cur += kRowB;
            if (geo.ring_cur >= geo.ring_warp + 5 * kRowB) geo.ring_cur -= 5 * kRowB;
            pairs = geo.ring_cur;
#pragma unroll
            for (int j = 0; j < NP; j++) sts_b64(pairs + 8 * j, Q[j]);
        }
#ifndef RIP_X2_NOCOLD   // (experiment switch: hot path only, wrong results)
        if (CN != 1 && __builtin_expect(__any_sync(FULL, flagged), 0)) {
            if (flagged) {
                if constexpr (!BLUR) {
#pragma unroll
                    for (int j = 0; j < NP; j++) sts_b64(pairs + 16 * j, Q[j]);
                }
                constexpr int NW = NPX * CN / 4;
#pragma unroll
                for (int k = 0; k < NW; k++) sts_u32(geo.patch + 8 * NPX + 4 * k, buf.w[k]);
                gray_patch<NPX, CN, BGR, BLUR ? 8 : 16>(gray_flag_mask<NPX>(E), pairs, geo.patch + 8 * NPX, geo.table);
            }
            reload_pairs_if<NP, BLUR ? 8 : 16>(Q, pairs, flagged);
        }
#endif
        if constexpr (SPECIAL) {
            if ((unsigned)(r - geo.adv_lo) < (unsigned)geo.adv_n) geo.src += geo.in_pitch;
        } else {
            geo.src += geo.in_pitch;   // (the main loop stops short of the rows where the band ends)
        }
        load_row_x2<NPX, CN>(buf, geo.src);
#if RIP_X2_L2PF > 0
        // pull the warp's bytes of a row further down into L2 (one 128-byte line per lane; pf_off is 0 in
        // the lanes that have no line to fetch and at the rows the band does not hold)
        if (geo.pf_off != 0 && (!SPECIAL || (unsigned)(r - geo.adv_lo) + RIP_X2_L2PF < (unsigned)geo.adv_n))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(geo.src + geo.pf_off));
#endif
    }

    // F[j] = (f[j], f[j + NP]): the row the Sobel stage consumes (blurred row yb, biased by kBias, or
    // the gray row scaled to normal floats when there is no blur stage)
    u64 F[NP];
    const int yb = BLUR ? r - 2 : r;
    if constexpr (BLUR) {
        // vertical pass, accumulate form: row r completes blurred row r-2
        const u64 GV0 = pk2(xp.gv0, xp.gv0), GV1 = pk2(xp.gv1, xp.gv1), GV2 = pk2(xp.gv2, xp.gv2);
        u64 V[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) {
            V[j] = fma2(GV2, Q[j], st.a0[j]);
            st.a0[j] = fma2(GV1, Q[j], st.a1[j]);
            st.a1[j] = fma2(GV0, Q[j], st.a2[j]);
            st.a2[j] = fma2(GV1, Q[j], st.a3[j]);
            st.a3[j] = mul2(GV2, Q[j]);
        }
        if constexpr (SPECIAL) {
            // warm-up rows of a segment: the first blurred row any stored output reads (ys - 1) completes at step
            // r_store - 2; before that only the vertical accumulators matter (the vote tells the compiler that
            // the whole warp leaves together)
            if (__all_sync(FULL, r < geo.r_store - 2)) {
                geo.dst += W;
                return;
            }
        }
        // horizontal pass: P[k] = (c[k], c[k + NP]) with c[m] = V of pixel m - 2 (pixel m lives in
        // pair m % NP, half m / NP).  Clamp-to-edge columns (GaussianBlur.cpp:240): V is linear in the
        // gray column, so the clamp applies to V: left of column 0 / right of column W-1 repeat it.
        float Vm2 = __shfl_up_sync(FULL, hi2(V[NP - 2]), 1), Vm1 = __shfl_up_sync(FULL, hi2(V[NP - 1]), 1);
        float Vp0 = __shfl_down_sync(FULL, lo2(V[0]), 1), Vp1 = __shfl_down_sync(FULL, lo2(V[1]), 1);
        if constexpr (EDGE) {
            Vm2 = geo.e_left ? lo2(V[0]) : Vm2;
            Vm1 = geo.e_left ? lo2(V[0]) : Vm1;
            Vp0 = geo.e_right ? hi2(V[NP - 1]) : Vp0;
            Vp1 = geo.e_right ? hi2(V[NP - 1]) : Vp1;
        }
        u64 P[NP + 4];
        P[0] = pk2(Vm2, lo2(V[NP - 2]));   // (V[-2], V[NP-2])
        P[1] = pk2(Vm1, lo2(V[NP - 1]));   // (V[-1], V[NP-1])
#pragma unroll
        for (int j = 0; j < NP; j++) P[j + 2] = V[j];
        P[NP + 2] = pk2(hi2(V[0]), Vp0);   // (V[NP],   V[NPX])
        P[NP + 3] = pk2(hi2(V[1]), Vp1);   // (V[NP+1], V[NPX+1])
        const u64 GH0 = pk2(xp.gh0, xp.gh0), GH1 = pk2(xp.gh1, xp.gh1), GH2 = pk2(xp.gh2, xp.gh2);
        const u64 BIAS = pk2(kBias, kBias);
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const u64 e2 = add2(P[j], P[j + 4]), e1 = add2(P[j + 1], P[j + 3]);
            // S~ of pixels j, j + NP, plus the bias (it rides in the FMA chain): floor(S~) in bits 15..22, fraction below
            F[j] = fma2(GH2, e2, fma2(GH1, e1, fma2(GH0, P[j + 2], BIAS)));
        }
        // guard band: the fraction bits within a ulps of 0 (mod 2^kFracBits)
        uint32_t zmin = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < NP; j++)
            zmin = __vimin3_u32(zmin, (lo2u(F[j]) << (32 - kFracBits)) + xp.zoff, (hi2u(F[j]) << (32 - kFracBits)) + xp.zoff);
        const uint32_t flagged = zmin < xp.zthr ? 1u : 0u;
#ifndef RIP_X2_NOCOLD
        if (__builtin_expect(__any_sync(FULL, flagged), 0)) {
#pragma unroll
            for (int j = 0; j < NP; j++) sts_b64(geo.patch + 16 * j, F[j]);
            __syncwarp();  // the newest ring row was just stored by the other lanes
            blur_replay_warp<NPX, STATS>(F, geo.patch, geo.scratch, geo.ring_warp, geo.ring_cur, geo.w25, geo.cmin, geo.cmax, xp.zoff, xp.zthr,
                                  p.slow_counter);
            // (its trailing __syncwarp also orders the ring reads before the next step overwrites the oldest slot)
            reload_pairs_if<NP, 16>(F, geo.patch, flagged);
        }
#else
        if (flagged == 77u) F[0] = 0;
#endif
#pragma unroll
        for (int j = 0; j < NP; j++) F[j] = pk2u(lo2u(F[j]) & kBiasMask, hi2u(F[j]) & kBiasMask);   // kBias + floor(S)
    } else {
        const float sc = __uint_as_float(0x7f000000u);  // 2^127: q*2^-149 -> q*2^-22 (a normal float)
        const u64 SC = pk2(sc, sc);
#pragma unroll
        for (int j = 0; j < NP; j++) F[j] = mul2(Q[j], SC);
    }

    // ---- 3. Sobel, vertical pass first: output row yo = yb-1 reads rows yb-2, yb-1, yb -------------
    // BORDER_REFLECT_101 in y: row -1 -> row 1 (first output row), row H -> row H-2 (last output row;
    // this step's input row is a dummy then).  Only the SPECIAL copy of the row body carries these
    // checks; the caller runs it for the first and last trips of a segment only.
    if constexpr (SPECIAL) {
        if (yb == 1) {
#pragma unroll
            for (int j = 0; j < NP; j++) st.F2[j] = F[j];
        }
        if (yb == H) {
#pragma unroll
            for (int j = 0; j < NP; j++) F[j] = st.F2[j];
        }
    }
    const u64 TWO = pk2(2.f, 2.f);
    u64 Vs[NP], Vd[NP];   // Vs = f(yb-2) + 2 f(yb-1) + f(yb),  Vd = f(yb) - f(yb-2)
#pragma unroll
    for (int j = 0; j < NP; j++) {
        Vs[j] = fma2(TWO, st.F1[j], add2(st.F2[j], F[j]));
        Vd[j] = sub2(F[j], st.F2[j]);
    }
    // horizontal pass, BORDER_REFLECT_101 in x: gx = Vs[x+1] - Vs[x-1],  gy = Vd[x-1] + 2 Vd[x] + Vd[x+1]
    float sl = __shfl_up_sync(FULL, hi2(Vs[NP - 1]), 1), sr = __shfl_down_sync(FULL, lo2(Vs[0]), 1);
    float dl = __shfl_up_sync(FULL, hi2(Vd[NP - 1]), 1), dr = __shfl_down_sync(FULL, lo2(Vd[0]), 1);
    if constexpr (EDGE) {
        sl = geo.e_left ? lo2(Vs[1]) : sl;             // x = -1 -> x = 1
        dl = geo.e_left ? lo2(Vd[1]) : dl;
        sr = geo.e_right ? hi2(Vs[NP - 2]) : sr;       // x = W  -> x = W-2
        dr = geo.e_right ? hi2(Vd[NP - 2]) : dr;
    }
    // KS[k] = (e[k], e[k + NP]) with e[m] = Vs of pixel m - 1; KD likewise for Vd
    u64 KS[NP + 2], KD[NP + 2];
    KS[0] = pk2(sl, lo2(Vs[NP - 1]));
    KD[0] = pk2(dl, lo2(Vd[NP - 1]));
#pragma unroll
    for (int j = 0; j < NP; j++) { KS[j + 1] = Vs[j]; KD[j + 1] = Vd[j]; }
    KS[NP + 1] = pk2(hi2(Vs[0]), sr);
    KD[NP + 1] = pk2(hi2(Vd[0]), dr);
    {
        // m * OS has the integer round-half-even(m) as its bit pattern (denormal result)
        const float os = __uint_as_float(BLUR ? 1u /* 2^-149 */ : 0x00400000u /* 2^-127 */);
        const u64 OS = pk2(os, os);
        uint32_t q[NPX];
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const u64 gx = sub2(KS[j + 2], KS[j]);
            const u64 gy = fma2(TWO, KD[j + 1], add2(KD[j], KD[j + 2]));
            const u64 m2 = fma2(gx, gx, mul2(gy, gy));
            const u64 m = mul2(pk2(sqrt_approx(lo2(m2)), sqrt_approx(hi2(m2))), OS);
            q[j] = lo2u(m);
            q[j + NP] = hi2u(m);
        }
        const uint32_t ok = (!SPECIAL || (r >= geo.r_store && r <= geo.r_last)) ? geo.store_lane : 0u;
        const uint32_t w0 = i2ip(q[1], q[0], i2ip(q[3], q[2], 0u));
        if constexpr (NPX == 8) {
            const uint32_t w1 = i2ip(q[5], q[4], i2ip(q[7], q[6], 0u));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.v2.u32 [%0], {%1, %2};\n\t}"
                         ::"l"(geo.dst), "r"(w0), "r"(w1), "r"(ok) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
                         ::"l"(geo.dst), "r"(w0), "r"(ok) : "memory");
        }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) {
        st.F2[j] = st.F1[j];
        st.F1[j] = F[j];
    }
    geo.dst += W;
}

// The rows of one segment.  Head: the warm-up rows and the first storing row (it may be frame row 0), one row
// per trip with the SPECIAL copy of the row body and an explicit rotation of the three row buffers.  Main loop:
// NB rows per trip with the plain copy; the buffer rotation and the Sobel delay line are register renames
// there.  Tail: the remaining rows (the last may be frame row H-1), SPECIAL again.  Only the main loop is hot,
// so only its three copies of the row body need to stay in the instruction cache.
// NB = row buffers = how many rows ahead of their use the register loads run (and the unroll of the main loop:
// 3 with the blur stage; 6 without it, where a row is consumed twice as fast and load latency is the top stall).
template <int NPX, int CN, bool BGR, bool BLUR>
struct X2Cfg {
    static constexpr int NB = BLUR ? 3 : 6;
};

template <int NPX, int CN, bool BGR, bool BLUR, bool EDGE, int NB, bool STATS>
__device__ __forceinline__ void run_rows_x2(WarpX<NPX, CN> &st, RawX<NPX, CN> (&b)[NB], const X2Params &xp, GeoX &geo, int r)
{
    const FusedParams &p = xp.f;
#pragma unroll 1
    for (; r <= geo.r_store && r < geo.r_last; r++) {
        step_x2<NPX, CN, BGR, BLUR, true, EDGE, STATS>(st, b[0], xp, geo, r);
        const RawX<NPX, CN> t = b[0];
#pragma unroll
        for (int k = 0; k + 1 < NB; k++) b[k] = b[k + 1];
        b[NB - 1] = t;
    }
    // the main loop runs while all rows of a trip store, and the row they load (NB ahead) as well as the line
    // they prefetch (RIP_X2_L2PF further) lie inside the band: step s advances freely iff
    // s + NB - 1 + RIP_X2_L2PF < in_row0 + in_rows - 1
    const int r_main_last = min(geo.r_last - 1, p.in_row0 + p.in_rows - 1 - NB - RIP_X2_L2PF);
#pragma unroll 1
    for (; r + NB - 1 <= r_main_last; r += NB) {
#pragma unroll
        for (int k = 0; k < NB; k++) step_x2<NPX, CN, BGR, BLUR, false, EDGE, STATS>(st, b[k], xp, geo, r + k);
    }
#pragma unroll 1
    for (; r <= geo.r_last; r++) {
        step_x2<NPX, CN, BGR, BLUR, true, EDGE, STATS>(st, b[0], xp, geo, r);
        const RawX<NPX, CN> t = b[0];
#pragma unroll
        for (int k = 0; k + 1 < NB; k++) b[k] = b[k + 1];
        b[NB - 1] = t;
    }
}

template <int NPX, int CN, bool BGR, bool BLUR, bool STATS = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, NPX == 8 ? ((BLUR || CN == 1) ? RIP_X2_MINB8 : RIP_X2_MINB8_NOBLUR) : RIP_X2_MINB4)
fused_x2_kernel(const __grid_constant__ X2Params xp)
{
    constexpr int HALO = BLUR ? 3 : 1;  // input rows above/below an output row
    constexpr int kRowW = 32 * NPX;     // words per ring row
    constexpr int kBand = 30 * NPX;
    const FusedParams &p = xp.f;

    __shared__ __align__(16) uint32_t ring[BLUR ? kWarpsPerBlock * 5 * kRowW : 4];
    __shared__ __align__(16) uint32_t patch[kWarpsPerBlock * 3 * kRowW];   // per lane: NPX/2 pairs at a 16-byte stride + NPX words of raw input
    __shared__ float w25s[32];
    __shared__ __align__(16) float scratch[kWarpsPerBlock * 32];
    __shared__ __align__(16) uint32_t gray_down_s[2048];   // the (r,g) bit table of the gray fix: global-memory latency would stall the whole warp
    if constexpr (CN != 1) {
        // 8 KB per block: all loads of a thread in flight at once (a dependent load/store loop here was 3.9 % of
        // the kernel's stall samples: sixteen L2 round trips in sequence before the first row)
        static_assert(2048 % (4 * kWarpsPerBlock * 32) == 0, "table copy is unrolled");
        constexpr int kPer = 2048 / (4 * kWarpsPerBlock * 32);
        uint4 t[kPer];
#pragma unroll
        for (int k = 0; k < kPer; k++) t[k] = __ldg(reinterpret_cast<const uint4 *>(d_gray_down) + threadIdx.x + k * kWarpsPerBlock * 32);
#pragma unroll
        for (int k = 0; k < kPer; k++) reinterpret_cast<uint4 *>(gray_down_s)[threadIdx.x + k * kWarpsPerBlock * 32] = t[k];
    }
    if (threadIdx.x < 25) w25s[threadIdx.x] = p.w[threadIdx.x] * 1.2676506002282294e30f;   // * 2^100 (exact), see blur_replay_warp
    __syncthreads();  // the only block-level barrier: the warps are independent from here on

    GeoX geo;
    geo.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    geo.ring_warp = (uint32_t)__cvta_generic_to_shared(ring + (BLUR ? warp * 5 * kRowW : 0));
    geo.ring_cur = geo.ring_warp + NPX * 4 * geo.lane;
    geo.patch = (uint32_t)__cvta_generic_to_shared(patch + threadIdx.x * 3 * NPX);
    geo.w25 = (uint32_t)__cvta_generic_to_shared(w25s);
    geo.scratch = (uint32_t)__cvta_generic_to_shared(scratch + 32 * (threadIdx.x >> 5));
    geo.table = (uint32_t)__cvta_generic_to_shared(gray_down_s);
    int bid = blockIdx.x;
    const int bg = bid % p.n_band_groups; bid /= p.n_band_groups;
    const int seg = bid % p.n_segs;
    const int frame = bid / p.n_segs;
    const int band = bg * kWarpsPerBlock + warp;
    if (band >= p.n_bands) return;  // warp-uniform

    const int W = p.W;
    const int xw0 = band * kBand;
    const int x = xw0 - NPX + NPX * geo.lane;      // first of this lane's pixels
    const bool in_img = (x >= 0) && (x < W);       // W % NPX == 0: a lane is fully inside or fully outside
    const int lane_last = (W - xw0) / NPX;         // lane holding the last NPX pixels of the row (may be > 31)
    geo.e_left = (band == 0) && geo.lane == 1;
    geo.e_right = geo.lane == lane_last;
    geo.cmin = band == 0 ? NPX : 0;
    geo.cmax = min(lane_last, 31) * NPX + NPX - 1;
    const int ys = p.out_row0 + seg * p.seg_rows;
    const int ye = min(ys + p.seg_rows, p.out_row0 + p.out_rows);
    geo.r_store = ys + HALO;
    geo.r_last = ye - 1 + HALO;
    geo.in_pitch = (uint32_t)W * CN;
    const uint8_t *in_base = p.in + (size_t)frame * p.in_frame_bytes;
    geo.store_lane = ((geo.lane >= 1) && (geo.lane <= 30) && in_img) ? 1u : 0u;

    WarpX<NPX, CN> st;
#pragma unroll
    for (int j = 0; j < NPX / 2; j++) st.a0[j] = st.a1[j] = st.a2[j] = st.a3[j] = st.F1[j] = st.F2[j] = 0ull;

    // Rows are clamped to the rows the input band holds.  The host guarantees the band covers every
    // row an output needs, and that it starts at row 0 / ends at row H-1 wherever the clamp-to-edge rule
    // (GaussianBlur.cpp:241) is actually exercised; other clamped rows are read-ahead only.
    int r = ys - HALO;
    const uint32_t xoff = in_img ? (uint32_t)x * CN : 0u;
    constexpr int NB = X2Cfg<NPX, CN, BGR, BLUR>::NB;
    RawX<NPX, CN> b[NB];   // rows r .. r + NB - 1
#pragma unroll
    for (int k = 0; k < NB; k++) {
        const int i = min(max(r + k - p.in_row0, 0), p.in_rows - 1);
        geo.src = in_base + (size_t)i * geo.in_pitch + xoff;
        load_row_x2<NPX, CN>(b[k], geo.src);
    }
    // step r loads row r + NB = one past the row src points at: advance iff in_row0 <= r + NB - 1 < in_row0 + in_rows - 1
    geo.adv_lo = p.in_row0 - (NB - 1);
    geo.adv_n = p.in_rows - 1;
    {   // lanes 0..n-1 fetch the n consecutive 128-byte lines that hold the warp's NPX*CN*32 bytes of a row
        constexpr int kLines = (32 * NPX * CN + 127) / 128 + 1;
        const uint32_t lane0_to_me = (uint32_t)(NPX * CN) * (uint32_t)geo.lane;   // src points at this lane's pixels
        geo.pf_off = (in_img && geo.lane < kLines) ? (uint32_t)RIP_X2_L2PF * geo.in_pitch + 128u * (uint32_t)geo.lane - lane0_to_me : 0u;
        if (x - NPX * geo.lane < 0) geo.pf_off = 0u;   // (left-most band: lane 0 sits before the row; keep it simple)
    }
    // output row produced by the step of input row r is r - HALO
    geo.dst = p.out + (size_t)frame * p.out_rows * W + (ptrdiff_t)(r - HALO - p.out_row0) * W + x;

    // One copy of the row loops for every warp.  A second copy without the border selects for the warps whose
    // band touches neither image edge (-4 % instructions on their hot path) was measured at 525 us instead of
    // 418 us per 32 4K frames: with both copies live on an SM the main loops no longer fit the instruction
    // cache (no_instruction stalls 0.17 -> 1.86 warps per issue).  RIP_X2_INTERIOR re-enables it for experiments
    // (the vote tells the compiler what it cannot see: the choice is the same in every lane).
#ifdef RIP_X2_INTERIOR
    if (__any_sync(FULL, band == 0 || lane_last <= 31)) run_rows_x2<NPX, CN, BGR, BLUR, true, NB, STATS>(st, b, xp, geo, r);
    else run_rows_x2<NPX, CN, BGR, BLUR, false, NB, STATS>(st, b, xp, geo, r);
#else
    run_rows_x2<NPX, CN, BGR, BLUR, true, NB, STATS>(st, b, xp, geo, r);
#endif
}
