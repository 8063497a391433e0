// Weight gradient of the 3x3 convs with Cout <= 64 (levels 1 and 2 of the U-Net, where the
// pixel count is large and the channel count small), swapped operand roles, one halo box.
//
//   dW[(dy, dx, ci)][co] = sum over pixels p of X[p + (dy, dx)][ci] * dY[p][co]
//   (autograd of nn.Conv2d at src/foundation_stereo_depth/model.py:36,39, reached by train.py:342)
//
// tcgen05 view: the contraction index K is the PIXEL, so both operands are MN-major.
//   A (M side) = X:  ONE TMA box of (TH+2) x (TW+2) = 18 x 10 pixels per channel atom (CA = 32 or 64
//                channels = one swizzle row).  The hardware swizzle is a function of the shared-memory
//                address only (tests/umma_shift_probe.cu), so the nine taps are nine views of that box:
//                start = (dy*10 + dx) rows in, stride between 8-pixel K groups = 10 rows.  The three
//                vertical taps are M-atoms LBO = 10 rows apart: CA = 32 -> one M = 128 MMA covers
//                (dy = 0,1,2 and a fourth, ignored, atom); CA = 64 -> M = 128 (dy = 0,1) + M = 64 (dy = 2).
//   B (N side) = dY: the 16 x 8 pixel tile, N = Cout.
// A tile is 128 pixels (eight K = 16 steps); every tap / k-step / unit offset is a compile-time constant
// added to one per-stage descriptor pair, and the MMAs are predicated on the elected lane, so the issue
// loop runs on the uniform datapath (measured MMA cost at M = 128: 34 + N/4 cycles, shared-memory bound).
// Split-K over pixel tiles; fp32 partials merged with red.global.add into the [k][co] workspace.
#pragma once
#include "ptx.cuh"
#include "wgrad_gemm.cuh"

namespace sdn {

template <int CA, int COUT, int NATOMS, int NDX = 3>
struct WtrCfg {
    static constexpr int SWB = CA * 2;                       // X swizzle row
    static constexpr int SWY = COUT * 2;                     // dY swizzle row
    static constexpr int TW = 8, TH = 16, KPIX = 128;
    static constexpr int BOX_W = TW + 2, BOX_H = TH + 2;
    static constexpr int Y_BYTES = KPIX * SWY;
    static constexpr int X_BYTES = ((BOX_W * BOX_H * SWB) + 1023) & ~1023;
    static constexpr int X_TX = BOX_W * BOX_H * SWB;         // bytes one X box actually delivers
    static constexpr int STAGE_BYTES = Y_BYTES + NATOMS * X_BYTES;
    static constexpr int COLS_PER_UNIT = COUT * (CA == 64 ? 2 : 1);
    static constexpr int COLS = NATOMS * NDX * COLS_PER_UNIT;
    static constexpr int TMEM_COLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    static_assert(COLS <= 512, "accumulators exceed TMEM");
    static constexpr int smem_bytes(int stages) { return 1024 + stages * STAGE_BYTES + 4 * 32 * 33 * 4 + 256; }
};

// NDX = 1: vertical taps only (the first layer in its row-halo form): the single horizontal position is the centre.
template <int CA, int COUT, int NATOMS, int NDX = 3>
__global__ void __launch_bounds__(192, 1) wgrad_tr_kernel(const __grid_constant__ WgradParams p) {
    using Cfg = WtrCfg<CA, COUT, NATOMS, NDX>;
    constexpr int SWB = Cfg::SWB, SWY = Cfg::SWY;
    constexpr uint32_t LAYOUT_X = (SWB == 128) ? 2u : 4u;
    constexpr uint32_t LAYOUT_Y = (SWY == 128) ? 2u : 4u;
    constexpr uint32_t ROW10 = Cfg::BOX_W * SWB;   // one image row of the box

    ptx::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    float* tsm_base = reinterpret_cast<float*>(smem + stages * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * Cfg::STAGE_BYTES + 4 * 32 * 33 * 4);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tfull_bar = bars + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int atom0 = blockIdx.y * NATOMS;
    const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.a_maps[0]);
        ptx::prefetch_tmap(&p.b_maps[0]);
        ptx::prefetch_tmap(&p.b_maps[1]);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    ptx::pdl_wait();

    const int my_tiles = (int)blockIdx.x < ptiles ? (ptiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ producer (converged warp)
        int s = 0;
        uint32_t ph = 0;
        int a_src[NATOMS], a_c0[NATOMS];
#pragma unroll
        for (int a = 0; a < NATOMS; ++a) {
            const int ca = atom0 + a;
            a_src[a] = ca < p.atoms_src0 ? 0 : 1;
            a_c0[a] = (a_src[a] == 0 ? ca : ca - p.atoms_src0) * CA;
        }
        ptx::TileWalker tw;
        for (tw.init(blockIdx.x, gridDim.x, ptiles, 1, p.tiles_x, p.tiles_y); tw.valid(); tw.next()) {
            const int x0 = tw.tx * Cfg::TW, y0 = tw.ty * Cfg::TH, n0 = tw.tn;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* y_dst = smem + s * Cfg::STAGE_BYTES;
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::Y_BYTES + NATOMS * Cfg::X_TX);
                ptx::tma_load_4d(y_dst, &p.a_maps[0], &full_bar[s], 0, x0, y0, n0);
#pragma unroll
                for (int a = 0; a < NATOMS; ++a)
                    ptx::tma_load_4d(y_dst + Cfg::Y_BYTES + a * Cfg::X_BYTES, &p.b_maps[a_src[a]], &full_bar[s], a_c0[a],
                                     x0 - 1, y0 - 1, n0);
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------- MMA issuer (converged, predicated)
        int s = 0;
        uint32_t ph = 0;
        constexpr uint32_t IDESC128 = ptx::make_idesc_bf16(128, COUT, 1, 1);
        constexpr uint32_t IDESC64 = ptx::make_idesc_bf16(64, COUT, 1, 1);
        const uint32_t smem0 = ptx::smem_u32(smem);
        // dY: N = COUT is one swizzle row wide (LBO unused); 8-pixel K groups are contiguous
        const uint64_t ydesc0 = ptx::make_smem_desc(smem0, 0, 8 * SWY, LAYOUT_Y);
        // X: vertical taps = M atoms one box row apart; 8-pixel K groups = consecutive image rows of the box
        const uint64_t xdesc0 = ptx::make_smem_desc(smem0 + Cfg::Y_BYTES, ROW10, ROW10, LAYOUT_X);
        constexpr uint32_t STAGE16 = Cfg::STAGE_BYTES >> 4;
        for (int it = 0; it < my_tiles; ++it) {
            ptx::mbar_wait(&full_bar[s], ph);
            ptx::tc_fence_after();
            const bool leader = ptx::elect_one();
            const uint32_t lead = leader ? 1u : 0u;
            const uint64_t ys = ydesc0 + uint64_t(uint32_t(s) * STAGE16);
            const uint64_t xs = xdesc0 + uint64_t(uint32_t(s) * STAGE16);
            const uint32_t first = it != 0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 16 pixels = image rows 2k, 2k+1 of the tile
                const uint32_t acc = k != 0 ? 1u : first;
#pragma unroll
                for (int a = 0; a < NATOMS; ++a) {
#pragma unroll
                    for (int dxl = 0; dxl < NDX; ++dxl) {
                        constexpr int DX0 = NDX == 1 ? 1 : 0;
                        const int dx = dxl + DX0;
                        const uint64_t ydesc = ys + uint64_t((k * 16 * SWY) >> 4);
                        const uint64_t xdesc = xs + uint64_t((a * Cfg::X_BYTES + (2 * k) * ROW10 + dx * SWB) >> 4);
                        const uint32_t d = tmem_base + uint32_t((a * NDX + dxl) * Cfg::COLS_PER_UNIT);
                        ptx::tc_mma_bf16_pred(d, xdesc, ydesc, IDESC128, acc, lead);
                        if (CA == 64)
                            ptx::tc_mma_bf16_pred(d + COUT, xdesc + uint64_t((2 * ROW10) >> 4), ydesc, IDESC64, acc, lead);
                    }
                }
            }
            if (leader) ptx::tc_commit(&empty_bar[s]);
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (ptx::elect_one()) ptx::tc_commit(tfull_bar);
    } else if (my_tiles > 0) {
        // ------------------------------------------------------------ epilogue (once per CTA)
        const int quarter = warp & 3;
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        float* out = p.out;
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
        constexpr int HALVES = CA == 64 ? 2 : 1;
        // transpose each warp's 32 rows x 32 columns through shared memory so that one red instruction
        // covers 32 CONSECUTIVE output channels of one workspace row
        float* tsm = tsm_base + (warp - 2) * (32 * 33);
        for (int g = 0; g < NATOMS * NDX; ++g) {
            const int ca = atom0 + g / NDX, dxi = NDX == 1 ? 1 : g % NDX;
            if (ca >= p.atoms_per_tap) break;
            for (int h = 0; h < HALVES; ++h) {
                for (int ch = 0; ch < COUT / 32; ++ch) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + g * Cfg::COLS_PER_UNIT + h * COUT + ch * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tsm[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int nrows = h == 0 ? 32 : 16;   // M = 64: 16 rows per lane quarter
                    for (int j = 0; j < nrows; ++j) {
                        int dyi, ci;
                        if (h == 0) { const int rr = quarter * 32 + j; dyi = rr / CA; ci = rr % CA; }
                        else { dyi = 2; ci = quarter * 16 + j; }
                        const int krow = (dyi * 3 + dxi) * p.cin_tot + ca * CA + ci;
                        if (dyi < 3 && krow < p.k_rows_valid)
                            atomicAdd(out + size_t(krow) * COUT + ch * 32 + lane, tsm[j * 33 + lane]);
                    }
                    __syncwarp();
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace sdn
