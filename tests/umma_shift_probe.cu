// Hardware probe (not part of the product): can a K-major swizzled tcgen05 A operand be read
// through a descriptor whose start address is shifted by a NON-multiple-of-8 number of rows,
// and with a stride between 8-row groups (SBO) that is not a multiple of the swizzle repeat?
// If yes, ONE (TH+2)x(TW+2) halo box can serve all nine taps of a 3x3 conv.
//
// A[r][c] rows are TMA-loaded (hardware swizzle) into shared memory; B is an identity block, so
// D[m][n] must equal A[row(m)][n].  Each case = (row shift, SBO in rows, base_offset field).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I stereo_depth_estimation_b200/csrc \
//        -o tests/umma_shift_probe.bin tests/umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"

using namespace sdn;

struct Case { int shift, sbo_rows, base_off; };
constexpr int MAX_CASES = 64;
struct Params {
    CUtensorMap a_map, b_map;
    Case cases[MAX_CASES];
    int ncases;
    float* out;   // [ncases][128][64]
};

template <int SW>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
    constexpr int KE = SW / 2;           // bf16 per row
    constexpr int NROWS = 256;
    constexpr uint32_t LAYOUT = SW == 128 ? 2u : 4u;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;                         // NROWS rows
    uint8_t* b_s = smem + NROWS * SW;            // 64 rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + 64 * SW);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) { ptx::tmem_alloc(tptr, 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(&bars[0], (NROWS + 64) * SW);
        ptx::tma_load_2d(a_s, &p.a_map, &bars[0], 0, 0);
        ptx::tma_load_2d(b_s, &p.b_map, &bars[0], 0, 0);
    }
    ptx::mbar_wait(&bars[0], 0);
    ptx::tc_fence_after();
    constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, 64, 0, 0);
    uint32_t ph = 0;
    for (int c = 0; c < p.ncases; ++c) {
        const Case cs = p.cases[c];
        if (threadIdx.x == 0) {
            const uint32_t a_addr = ptx::smem_u32(a_s) + cs.shift * SW;
            uint64_t adesc = ptx::make_smem_desc(a_addr, 16, cs.sbo_rows * SW, LAYOUT) | (uint64_t(cs.base_off & 7) << 49);
            const uint64_t bdesc = ptx::make_smem_desc(ptx::smem_u32(b_s), 16, 8 * SW, LAYOUT);
            for (int k = 0; k < KE / 16; ++k)
                ptx::tc_mma_bf16(tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), IDESC, k != 0 ? 1u : 0u);
            ptx::tc_commit(&bars[1]);
        }
        ptx::mbar_wait(&bars[1], ph);
        ph ^= 1;
        ptx::tc_fence_after();
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + ch * 32, v);
            ptx::tmem_ld_wait();
            float* o = p.out + (size_t(c) * 128 + warp * 32 + lane) * 64 + ch * 32;
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
        }
        ptx::tc_fence_before();
        __syncthreads();
        ptx::tc_fence_after();
    }
    if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

// ---- second probe: cost of back-to-back MMAs into the SAME accumulator vs several accumulators ----
// 64 MMAs (M = 128, K = 16, operands = whatever is in shared memory) issued by one thread; nacc
// accumulators used round-robin; cycles from first issue to the commit's mbarrier arrival.
template <int N, int NACC, int MODE>
__global__ void __launch_bounds__(128, 1) chain_kernel(long long* out, int slot) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bars[0], 1); ptx::fence_mbar_init(); }
    if (warp == 0) { ptx::tmem_alloc(tptr, 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tptr;
    if (MODE >= 2 ? warp == 0 : threadIdx.x == 0) {
        constexpr uint32_t IDESC = ptx::make_idesc_bf16(128, N, 0, 0);
        const uint64_t adesc = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024, 2u);
        const uint64_t bdesc = ptx::make_smem_desc(ptx::smem_u32(smem) + 32 * 1024, 16, 1024, 2u);
        uint32_t ph = 0;
        const uint32_t lead = (MODE >= 2) ? (ptx::elect_one() ? 1u : 0u) : 1u;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                if (MODE == 0)
                    ptx::tc_mma_bf16(tmem + uint32_t((i % NACC) * N), adesc + uint64_t(2 * (i & 3)), bdesc + uint64_t(2 * (i & 3)),
                                     IDESC, 1u);
                else if (MODE == 1)
                    ptx::tc_mma_bf16(tmem, adesc, bdesc, IDESC, 1u);
                else if (MODE == 2)
                    ptx::tc_mma_bf16_pred(tmem + uint32_t((i % NACC) * N), adesc + uint64_t(2 * (i & 3) + 64 * (i >> 2 & 3)),
                                          bdesc + uint64_t(2 * (i & 3)), IDESC, 1u, lead);
                else {
                    // MN-major operands (wgrad): MODE 3 = 64-byte swizzle atoms (32 channels), MODE 4 = 128-byte
                    constexpr uint32_t SW = MODE == 3 ? 64 : 128;
                    constexpr uint32_t LAY = MODE == 3 ? 4u : 2u;
                    constexpr uint32_t ID = ptx::make_idesc_bf16(128, N, 1, 1);
                    const uint64_t am = ptx::make_smem_desc(ptx::smem_u32(smem), 8 * SW, 8 * SW, LAY);
                    const uint64_t bm = ptx::make_smem_desc(ptx::smem_u32(smem) + 32 * 1024, 16 * SW, 8 * SW, LAY);
                    ptx::tc_mma_bf16_pred(tmem + uint32_t((i % NACC) * N), am + uint64_t((16 * SW / 16) * (i & 3)),
                                          bm + uint64_t((16 * SW / 16) * (i & 3)), ID, 1u, lead);
                }
            }
            const long long t1 = clock64();
            if (lead) ptx::tc_commit(&bars[0]);
            ptx::mbar_wait(&bars[0], ph);
            ph ^= 1;
            const long long t2 = clock64();
            if (rep == 2 && lead) { out[slot * 2] = t1 - t0; out[slot * 2 + 1] = t2 - t0; }
            if (MODE >= 2) __syncwarp();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int MODE>
static void run_chain1(long long* dout, int& slot) {
    const int smem = 1024 + 64 * 1024 + 256;
    cudaFuncSetAttribute(chain_kernel<N, NACC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    chain_kernel<N, NACC, MODE><<<1, 128, smem>>>(dout, slot);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, dout + slot * 2, sizeof h, cudaMemcpyDeviceToHost);
    printf("chain mode=%d N=%d accumulators=%d : issue %lld cyc, done %lld cyc -> %.1f cyc/MMA (%s)\n", MODE, N, NACC, h[0], h[1],
           h[1] / 64.0, cudaGetErrorString(e));
    ++slot;
}
template <int N>
static void run_chain(long long* dout, int& slot) {
    run_chain1<N, 1, 2>(dout, slot);
    run_chain1<N, 2, 3>(dout, slot);
    run_chain1<N, 2, 4>(dout, slot);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float aval(int r, int c) { return float((r * 7 + c * 3) % 251); }

template <int SW>
static int run(EncodeFn enc) {
    constexpr int KE = SW / 2, NROWS = 256;
    std::vector<__nv_bfloat16> ha(size_t(NROWS) * KE), hb(size_t(64) * KE);
    for (int r = 0; r < NROWS; ++r)
        for (int c = 0; c < KE; ++c) ha[size_t(r) * KE + c] = __float2bfloat16(aval(r, c));
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < KE; ++k) hb[size_t(n) * KE + k] = __float2bfloat16((n % KE) == k ? 1.f : 0.f);
    __nv_bfloat16 *da, *db;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    Params p{};
    const CUtensorMapSwizzle swz = SW == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[2] = {(cuuint64_t)KE, (cuuint64_t)NROWS};
        cuuint64_t strides[1] = {(cuuint64_t)KE * 2};
        cuuint32_t box[2] = {(cuuint32_t)KE, (cuuint32_t)NROWS};
        cuuint32_t es[2] = {1, 1};
        if (enc(&p.a_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode A failed\n"); return 1; }
        dims[1] = 64; box[1] = 64;
        if (enc(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode B failed\n"); return 1; }
    }
    std::vector<Case> cases;
    const int shifts[] = {0, 1, 2, 3, 4, 7, 8, 9};
    for (int sbo : {8, 10})
        for (int s : shifts)
            for (int bo = 0; bo < 8; ++bo) {
                // base_offset candidates: 0 and the "natural" phase of the start row
                if (bo != 0 && bo != (s & 7) && bo != ((s * SW / 128) & 7)) continue;
                if ((int)cases.size() < MAX_CASES) cases.push_back({s, sbo, bo});
            }
    p.ncases = (int)cases.size();
    for (int i = 0; i < p.ncases; ++i) p.cases[i] = cases[i];
    cudaMalloc(&p.out, size_t(p.ncases) * 128 * 64 * 4);
    cudaMemset(p.out, 0xFF, size_t(p.ncases) * 128 * 64 * 4);
    const int smem = 1024 + (NROWS + 64) * SW + 256;
    cudaFuncSetAttribute(probe_kernel<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<SW><<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("SW%d kernel failed: %s\n", SW, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(size_t(p.ncases) * 128 * 64);
    cudaMemcpy(out.data(), p.out, out.size() * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < p.ncases; ++i) {
        const Case cs = cases[i];
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; ++m) {
            const int row = cs.shift + (m / 8) * cs.sbo_rows + (m % 8);
            for (int n = 0; n < 64; ++n) {
                const float want = aval(row, n % KE);
                if (out[(size_t(i) * 128 + m) * 64 + n] != want) { ++bad; if (first_bad < 0) first_bad = m * 64 + n; }
            }
        }
        printf("SW%d shift=%d sbo_rows=%d base_off=%d : %s (%d wrong, first at m=%d n=%d)\n", SW, cs.shift, cs.sbo_rows,
               cs.base_off, bad == 0 ? "OK" : "MISMATCH", bad, first_bad < 0 ? -1 : first_bad / 64, first_bad < 0 ? -1 : first_bad % 64);
    }
    return 0;
}

int main() {
    cudaFree(0);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    int rc = 0;
    if (getenv("PROBE_SHIFT") != nullptr) { rc = run<128>(enc); rc |= run<64>(enc); }
    long long* dout;
    cudaMalloc(&dout, 64 * 2 * sizeof(long long));
    int slot = 0;
    run_chain<32>(dout, slot);
    run_chain<64>(dout, slot);
    run_chain<128>(dout, slot);
    run_chain<96>(dout, slot);
    run_chain<256>(dout, slot);
    return rc;
}
