// tma_probe.cu -- stand-alone probe of the TMA primitives used by rip_fused.cu (debug aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %d (%s) at %s:%d\n", (int)e, cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

constexpr int TW = 192, TRR = 4, WARPS = 4, NSTG = 2;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, uint32_t *out, int x0, int rows, int mode)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *tiles = reinterpret_cast<uint32_t *>(smem) + warp * NSTG * TRR * TW;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + WARPS * NSTG * TRR * TW * 4) + warp * NSTG;
    if (mode == 1 && warp >= 2) return;  // early exit of some warps
    if (lane == 0) {
        for (int s = 0; s < NSTG; s++) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < NSTG; s++) {
            mbar_expect_tx(bars + s, TRR * TW * 4);
            tma_load_2d(tiles + s * TRR * TW, &map, x0 + warp * 180, blockIdx.x * rows + s * TRR, bars + s);
        }
    }
    __syncwarp();
    uint32_t acc = 0;
    const int ntiles = rows / TRR;
    for (int k = 0; k < ntiles; k++) {
        if (k > 0) {
            __syncwarp();
            if (lane == 0 && k - 1 + NSTG < ntiles + (mode == 2 ? 1 : 0)) {
                const int st = (k - 1) % NSTG;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bars + st, TRR * TW * 4);
                tma_load_2d(tiles + st * TRR * TW, &map, x0 + warp * 180, blockIdx.x * rows + (k - 1 + NSTG) * TRR, bars + st);
            }
        }
        mbar_wait(bars + (k % NSTG), (k / NSTG) & 1);
        for (int r = 0; r < TRR; r++)
            for (int i = 0; i < 6; i++) acc += tiles[((k % NSTG) * TRR + r) * TW + 6 * lane + i];
    }
    if (mode == 2) mbar_wait(bars + (ntiles % NSTG), (ntiles / NSTG) & 1);  // drain the extra tile
    out[(blockIdx.x * WARPS + warp) * 32 + lane] = acc;
}

int main(int argc, char **argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 3840, H = argc > 2 ? atoi(argv[2]) : 2160, mode = argc > 3 ? atoi(argv[3]) : 0;
    const int x0arg = argc > 4 ? atoi(argv[4]) : 0;
    const int rows = 64, nblk = H / rows;
    const size_t pitch = (size_t)W * 3;
    std::vector<uint8_t> h(pitch * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d; uint32_t *dout;
    CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, nblk * WARPS * 32 * 4));
    void *sym = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)W * 3 / 4, (cuuint64_t)H};
    cuuint64_t gstr[1] = {pitch};
    cuuint32_t box[2] = {TW, TRR}, es[2] = {1, 1};
    CUresult cr = ((Fn)sym)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d (q=%d) W=%d H=%d mode=%d\n", (int)cr, (int)q, W, H, mode);
    const size_t smem = WARPS * NSTG * TRR * TW * 4 + WARPS * NSTG * 8;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int x0 : {x0arg}) {
        probe<<<nblk, 128, smem>>>(map, dout, x0, rows, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("x0=%d -> %s\n", x0, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    std::vector<uint32_t> ho(nblk * WARPS * 32);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    // reference for x0 = 174 (last launch), block 1, warp 1, lane 3
    {
        const int x0 = x0arg, b = 1, wp = 1, ln = 3;
        uint32_t acc = 0;
        for (int r = 0; r < rows; r++)
            for (int i = 0; i < 6; i++) {
                long word = x0 + wp * 180 + 6 * ln + i; long row = b * rows + r;
                uint32_t v = 0;
                if (word >= 0 && word < (long)W * 3 / 4) memcpy(&v, &h[row * pitch + word * 4], 4);
                acc += v;
            }
        printf("check: got %u want %u\n", ho[(b * WARPS + wp) * 32 + ln], acc);
    }
    return 0;
}
