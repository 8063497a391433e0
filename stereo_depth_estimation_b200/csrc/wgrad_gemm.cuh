// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
// dW[k = (tap, ci)][co] = sum over pixels p of dY[p][co] * X[p + tap][ci]
// (the weight gradient of nn.Conv2d / nn.ConvTranspose2d in the reference model,
// src/foundation_stereo_depth/model.py:36,39,67-73, which the reference gets from
// autograd at train.py:342).
//
// The contraction index is the PIXEL, which is the slow dimension of our NHWC
// tensors, so both operands are "MN-major" for tcgen05.mma: a TMA box
// (64 channels, TW, TH, TN) lands in shared memory as [pixels][128 B] with the
// 128-byte swizzle, which is exactly the canonical MN-major SW128 atom column
// (8-row groups 1024 B apart = SBO, channel atoms LBO apart).
//
//   A (M side) = dY: one or two 64-channel atoms -> M = 128 (with one atom the
//                second half aliases the first via LBO = 0 and is ignored).
//   B (N side) = "units".  For a 3x3 conv a unit is (horizontal tap dx, channel
//                atom): ONE TMA box of (TH+2) x TW pixels whose three vertical taps
//                are the same buffer read at row offsets 0, TW, 2*TW.  TW is a
//                multiple of 8, so those offsets are whole 8-row swizzle groups
//                and the shifted views stay canonical: the three taps are three
//                N-atoms LBO = TW rows apart and cost ONE load instead of three.
//                For 1x1 / ConvTranspose2d a unit is a plain channel atom.
//   One CTA owns (co tile) x (up to U units) with all of their accumulators in
//   TMEM (<= 512 columns), so dY is fetched once per pixel tile for all taps.
// The kernel is bound by L2->SM request throughput, not by the tensor pipe; the
// layout above cuts the requests per pixel ~4x against one box per tap.
// Split-K over pixel tiles; fp32 partials merged with red.global.add.
#pragma once
#include "ptx.cuh"

namespace sdn {

struct alignas(64) WgradParams {
    CUtensorMap a_maps[4];  // dY variants (convT: the 4 output quadrants), box (64, TW, TH, TN)
    CUtensorMap b_maps[2];  // X sources, box (CA, TW, TH + 2*halo, TN)
    int a_atoms;            // 1 or 2 64-channel atoms on the M side
    int a_variants;         // 1, or 4 for ConvTranspose2d
    int m_tiles;            // ceil(Cout / 128)
    int halo;               // 1: 3x3 conv (ndy = 3 row-shifted taps per unit), 0: single tap
    int U;                  // units per CTA
    int total_units;        // halo: 3 * atoms_per_tap, else atoms_per_tap
    int atoms_per_tap;      // Cin_total / CA
    int atoms_src0;         // atoms that come from b_maps[0]
    int unit_groups;        // ceil(total_units / U)
    int tiles_x, tiles_y, tiles_n, TW, TH, TN;
    int kpix;               // TW*TH*TN
    int stages;
    int tmem_cols;          // power of two >= U * ndy * CA
    int cout;               // valid output channels (row length of the workspace)
    int cin_tot;
    int k_rows_valid;       // rows of the workspace that exist
    float* out;             // [a_variants][taps * cin_tot][cout] fp32, pre-zeroed
    int tr;                 // transposed roles (Cout <= 64, 3x3): M = (vertical tap, ci) rows of a unit, N = co
    long long* dbg;         // -DSDN_FORENSICS: clock64 stamps [role][tile < 16][event < 8] of CTA (0,0,0)
};

#ifdef SDN_FORENSICS
#define SDN_WDBG(role, tile, ev)                                                                          \
    do {                                                                                                  \
        if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tile) < 16)      \
            p.dbg[((role) * 16 + (tile)) * 8 + (ev)] = clock64();                                          \
    } while (0)
#else
#define SDN_WDBG(role, tile, ev) do { } while (0)
#endif

// TR (Cout <= 64, 3x3 only): the roles are swapped.  The MMA's M side is one unit's halo tile - its three
// row-shifted vertical taps are M-atoms LBO = TW rows apart, M = 128 = 4 x 32 channels (the 4th atom is
// garbage and ignored) or 2 x 64 (+ one M = 64 MMA for the third tap) - and the N side is dY with
// N = Cout.  A level-1 layer then issues M=128,N=32 MMAs (16 pipe cycles) instead of M=64,N=96 ones
// that the pipe charges as M=128 (48 cycles) with 3/4 of the rows wasted.
template <int SWB, bool HALO, bool TR = false>
__global__ void __launch_bounds__(192, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
    constexpr int CA = SWB / 2;  // channels per B atom
    constexpr uint32_t LAYOUT_B = (SWB == 128) ? 2u : 4u;

    ptx::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    constexpr int ndy = HALO ? 3 : 1;
    const int a_tile_bytes = TR ? p.kpix * p.cout * 2 : p.kpix * 128;   // TR: dY box is exactly Cout wide
    const int b_rows = HALO ? (p.TH + 2) * p.TW : p.kpix;
    const int b_tile_bytes = ((b_rows * SWB) + 1023) & ~1023;
    const int stage_bytes = p.a_atoms * a_tile_bytes + p.U * b_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tfull_bar = bars + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int variant = blockIdx.z / p.m_tiles;
    const int m_tile = blockIdx.z % p.m_tiles;
    const int unit0 = blockIdx.y * p.U;
    const int nunits = min(p.U, p.total_units - unit0);
    const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_smem, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
    ptx::pdl_wait();

    const int my_tiles = (int)blockIdx.x < ptiles ? (ptiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        {   // converged warp; one elected lane issues the arrive + TMA instructions
            if (lane == 0) {
                ptx::prefetch_tmap(&p.a_maps[variant]);
                ptx::prefetch_tmap(&p.b_maps[0]);
                ptx::prefetch_tmap(&p.b_maps[1]);
            }
            int s = 0;
            uint32_t ph = 0;
            const uint32_t tx_bytes = p.a_atoms * a_tile_bytes + nunits * (b_rows * SWB);
            // per-unit constants, decoded once (no divisions inside the pixel-tile loop)
            int u_src[8], u_c0[8], u_dx[8];
            for (int g = 0; g < nunits; ++g) {
                const int u = unit0 + g;
                const int dxi = HALO ? u / p.atoms_per_tap : 1;
                const int ca = HALO ? u % p.atoms_per_tap : u;
                u_src[g] = ca < p.atoms_src0 ? 0 : 1;
                u_c0[g] = (u_src[g] == 0 ? ca : ca - p.atoms_src0) * CA;
                u_dx[g] = dxi - 1;
            }
            ptx::TileWalker tw;
            int dbg_it = 0;
            for (tw.init(blockIdx.x, gridDim.x, ptiles, 1, p.tiles_x, p.tiles_y); tw.valid(); tw.next()) {
                const int x0 = tw.tx * p.TW, y0 = tw.ty * p.TH, n0 = tw.tn * p.TN;
                const int dbg_tile = dbg_it++;
                (void)dbg_tile;
                if (lane == 0) SDN_WDBG(0, dbg_tile, 0);
                ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                if (lane == 0) SDN_WDBG(0, dbg_tile, 1);
                uint8_t* a_dst = smem + s * stage_bytes;
                uint8_t* b_dst = a_dst + p.a_atoms * a_tile_bytes;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                    for (int i = 0; i < p.a_atoms; ++i)
                        ptx::tma_load_4d(a_dst + i * a_tile_bytes, &p.a_maps[variant], &full_bar[s],
                                         m_tile * 128 + i * 64, x0, y0, n0);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < nunits)
                            ptx::tma_load_4d(b_dst + g * b_tile_bytes, &p.b_maps[u_src[g]], &full_bar[s], u_c0[g],
                                             x0 + u_dx[g], y0 - (HALO ? 1 : 0), n0);
                }
                __syncwarp();
                if (lane == 0) SDN_WDBG(0, dbg_tile, 2);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp runs the loop converged (descriptor math stays in uniform registers);
            // only the tcgen05 instructions are issued by one elected lane
            int s = 0;
            uint32_t ph = 0;
            // one dY atom -> M = 64 (half the shared-memory operand reads of an aliased M = 128)
            // halo: one MMA per unit (N = 3 row-shifted atoms, LBO = one image row).  plain: the units are
            // whole tiles b_tile_bytes apart, so ONE MMA covers all of them (N = nunits * CA <= 256).
            const bool grouped = !HALO && nunits * CA <= 256;
            const uint32_t idesc = ptx::make_idesc_bf16(p.a_atoms == 2 ? 128 : 64, grouped ? nunits * CA : ndy * CA, 1, 1);
            const uint32_t a_lbo = p.a_atoms == 2 ? uint32_t(a_tile_bytes) : 0u;
            const uint32_t b_lbo = HALO ? uint32_t(p.TW * SWB) : uint32_t(b_tile_bytes);  // halo: one image row = one vertical tap
            // descriptors of stage 0; later stages / k-steps / units only add to the 14-bit address field
            const uint32_t smem0 = ptx::smem_u32(smem);
            const uint64_t adesc0 = ptx::make_smem_desc(smem0, a_lbo, 1024, 2u);
            const uint64_t bdesc0 = ptx::make_smem_desc(smem0 + p.a_atoms * a_tile_bytes, b_lbo, 8 * SWB, LAYOUT_B);
            const uint32_t b_unit16 = uint32_t(b_tile_bytes) >> 4;
            const uint32_t stage16 = uint32_t(stage_bytes) >> 4;
            constexpr uint32_t ncol = uint32_t(ndy * CA);
            const uint32_t tr_idesc128 = ptx::make_idesc_bf16(128, TR ? p.cout : 16, 1, 1);
            const uint32_t tr_idesc64 = ptx::make_idesc_bf16(64, TR ? p.cout : 16, 1, 1);
            const uint32_t tr_cout = uint32_t(p.cout);
            const uint32_t tr_dy_row = uint32_t(p.cout * 2);                      // dY atom row: 64 B or 128 B
            const uint64_t tr_ydesc0 = ptx::make_smem_desc(smem0, 0, 8 * tr_dy_row, p.cout == 64 ? 2u : 4u);
            const uint64_t tr_xdesc0 = ptx::make_smem_desc(smem0 + a_tile_bytes, uint32_t(p.TW * SWB), 8 * SWB, LAYOUT_B);
            const uint32_t tr_ystep16 = (16 * tr_dy_row) >> 4;
            const uint32_t tr_dy2_16 = uint32_t(2 * p.TW * SWB) >> 4;
            const int nmma = grouped ? 1 : nunits;
            for (int it = 0; it < my_tiles; ++it) {
                if (lane == 0) SDN_WDBG(1, it, 0);
                ptx::mbar_wait(&full_bar[s], ph);
                ptx::tc_fence_after();
                if (lane == 0) SDN_WDBG(1, it, 1);
                const uint64_t sa = adesc0 + uint64_t(s * stage16);
                const uint64_t sb = bdesc0 + uint64_t(s * stage16);
                // one election per stage.  Where it measured faster (32-channel swapped-role units, the deep
                // row-halo layers) the MMAs are PREDICATED on it and the warp stays converged; elsewhere the
                // elected lane branches around the whole sequence.
                constexpr bool PRED = TR ? (CA == 32) : HALO;
                const bool leader = ptx::elect_one();
                const uint32_t lead = leader ? 1u : 0u;
                if (PRED || leader) {
                    if (TR) {
                        // stage-0 descriptors + (stage, k-step, unit) offsets in the 14-bit address field
                        const uint64_t ys = tr_ydesc0 + uint64_t(s * stage16);
                        const uint64_t xs = tr_xdesc0 + uint64_t(s * stage16);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ydesc = ys + uint64_t(k * tr_ystep16);
                            const uint64_t xk = xs + uint64_t(k * ((16 * SWB) >> 4));
                            const uint32_t acc = (it | k) != 0 ? 1u : 0u;
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                if (g < nunits) {
                                    const uint64_t xdesc = xk + uint64_t(g * b_unit16);
                                    if (CA == 32) {
                                        ptx::tc_mma_bf16_pred(tmem_base + g * tr_cout, xdesc, ydesc, tr_idesc128, acc, lead);
                                    } else {
                                        ptx::tc_mma_bf16(tmem_base + g * 2 * tr_cout, xdesc, ydesc, tr_idesc128, acc);
                                        ptx::tc_mma_bf16(tmem_base + g * 2 * tr_cout + tr_cout, xdesc + uint64_t(tr_dy2_16),
                                                         ydesc, tr_idesc64, acc);
                                    }
                                }
                            }
                        }
                    } else {
                        // kpix / 16 K = 16 steps of two 8-row groups each (halo tiles: 64 pixels, plain tiles: 128)
                        const int ksteps = HALO ? 4 : (p.kpix >> 4);
#pragma unroll 4
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t adesc = sa + uint64_t(k * ((16 * 128) >> 4));
                            const uint64_t bk = sb + uint64_t(k * ((16 * SWB) >> 4));
                            const uint32_t acc = (it | k) != 0 ? 1u : 0u;
#pragma unroll
                            for (int g = 0; g < (HALO ? 5 : 8); ++g)
                                if (g < nmma) {
                                    if (PRED) ptx::tc_mma_bf16_pred(tmem_base + g * ncol, adesc, bk + uint64_t(g * b_unit16), idesc, acc, lead);
                                    else ptx::tc_mma_bf16(tmem_base + g * ncol, adesc, bk + uint64_t(g * b_unit16), idesc, acc);
                                }
                        }
                    }
                }
                if (lane == 0) SDN_WDBG(1, it, 2);
                if (leader) ptx::tc_commit(&empty_bar[s]);
                __syncwarp();
                if (lane == 0) SDN_WDBG(1, it, 3);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
            if (ptx::elect_one()) ptx::tc_commit(tfull_bar);
        }
    } else {
        const int quarter = warp & 3;
        // M = 128: accumulator row r lives in TMEM lane r.  M = 64 (cta_group::1): row r lives in lane
        // 32 * (r / 16) + r % 16, i.e. each warp's lane quarter holds 16 rows in its first 16 lanes.
        const bool m64 = p.a_atoms == 1;
        const int r = m64 ? quarter * 16 + lane : quarter * 32 + lane;
        const int co = m_tile * 128 + r;
        const bool row_ok = (m64 ? lane < 16 : true) && (co < p.cout);
        if (TR && my_tiles > 0) {
            ptx::mbar_wait(tfull_bar, 0);
            ptx::tc_fence_after();
            float* out = p.out;
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
            const int halves = CA == 64 ? 2 : 1;
            // every MMA has completed, so the pipeline stages are free: use them to transpose each warp's
            // 32 rows x 32 columns so that one red instruction covers 32 CONSECUTIVE output channels of one
            // workspace row (coalesced) instead of 32 different rows
            float* tsm = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
            for (int g = 0; g < nunits; ++g) {
                const int u = unit0 + g;
                const int dxi = u / p.atoms_per_tap, ca = u % p.atoms_per_tap;
                for (int h = 0; h < halves; ++h) {
                    for (int ch = 0; ch < p.cout / 32; ++ch) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(taddr + (g * halves + h) * p.cout + ch * 32, v);
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
                                atomicAdd(out + size_t(krow) * p.cout + ch * 32 + lane, tsm[j * 33 + lane]);
                        }
                        __syncwarp();
                    }
                }
            }
        } else if (my_tiles > 0) {
            ptx::mbar_wait(tfull_bar, 0);
            ptx::tc_fence_after();
            constexpr int taps = HALO ? 9 : 1;
            float* out = p.out + size_t(variant) * taps * p.cin_tot * p.cout;
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16);
            const int ncols = nunits * ndy * CA;
            for (int ch = 0; ch < ncols / 32; ++ch) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + ch * 32, v);
                ptx::tmem_ld_wait();
                if (row_ok) {
                    const int cn0 = ch * 32;
                    const int g = cn0 / (ndy * CA);
                    const int dyi = (cn0 % (ndy * CA)) / CA;
                    const int u = unit0 + g;
                    const int dxi = HALO ? u / p.atoms_per_tap : 0;
                    const int ca = HALO ? u % p.atoms_per_tap : u;
                    const int tap = HALO ? dyi * 3 + dxi : 0;
                    const int krow0 = tap * p.cin_tot + ca * CA + (cn0 % CA);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int krow = krow0 + j;
                        if (krow < p.k_rows_valid) atomicAdd(out + size_t(krow) * p.cout + co, __uint_as_float(v[j]));
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

}  // namespace sdn
