// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel covers every "pixels x channels"
// contraction of the stereo U-Net (reference model: src/foundation_stereo_depth/
// model.py:32-104):
//   * conv3x3 forward       (model.py:36,39)  : 9 (or 18, skip-concat) A-segments
//   * conv3x3 data gradient                    : same, weights flipped/transposed
//   * ConvTranspose2d 2x2/s2 forward (model.py:67-73): 1 segment, 4 strided D maps
//   * ConvTranspose2d data gradient            : 4 segments over 4 strided A maps
//   * the im2col'ed first layer (enc1.block.0) : 1 segment, K = 64
//
// GEMM view: D[M = 128 pixels, N = BLOCK_N channels] += A[M, K] * B[N, K]^T with
// both operands K-major bf16 in shared memory (TMA, hardware swizzle) and the
// fp32 accumulator in TMEM.  The A tile of one k-block is ONE 4-D TMA box
// (channels, TW, TH, TN) of an NHWC activation tensor, shifted by the filter
// tap; out-of-bounds rows/columns are zero-filled by TMA, which is exactly the
// conv's zero padding.  torch.cat([up, skip]) (model.py:89-95) never
// materialises: the K loop simply walks two tensor maps.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..5 = epilogue (TMEM -> registers -> bf16 -> swizzled smem ->
// TMA store), double-buffered accumulators so the epilogue of tile i overlaps the
// main loop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace sdn {

enum : int { CG_RELU = 1, CG_STATS = 2, CG_BRES = 4, CG_BSTATS = 8, CG_DBG_NOMMA = 64, CG_DBG_NOEPI = 128, CG_DBG_NOLOADA = 256, CG_DBG_NOSTORE = 512, CG_DBG_NOSTATS = 1024 };

struct CgSeg {
    int8_t map;  // index into a_maps
    int8_t dx;   // added to the tile's x0
    int8_t dy;   // added to the tile's y0
    int8_t pad_;
    int16_t c0;       // first channel in that map
    int16_t cblocks;  // k-blocks taken from this segment
};

constexpr int CG_MAX_SEGS = 18;

struct alignas(64) ConvGemmParams {
    CUtensorMap a_maps[4];
    CUtensorMap b_map;
    CUtensorMap d_maps[4];
    CgSeg segs[CG_MAX_SEGS];
    int nsegs;
    int kblocks_total;
    int tiles_x, tiles_y, tiles_n;  // M tiling of (W, H, batch)
    int TW, TH, TN;                 // pixel box, TW*TH*TN == 128
    int img_w, img_h, img_n;        // output extent (rows of a ragged box outside it are zeroed)
    int n_tiles;                    // N tiling
    int n_per_dmap;                 // output channels per destination map (a tile may span several maps)
    int n_total;                    // n_tiles * BLOCK_N
    int flags;
    int stages;
    int a_stage_bytes;      // HALO: bytes of one (TH+2) x TW halo box, 1024-aligned
    int b_res_bytes;        // CG_BRES: the whole packed weight matrix lives in shared memory (loaded once per CTA)
    int ups;                // HALO: units (k-blocks) per pipeline stage (1 or 3): fewer producer/MMA handshakes per tile
    const float* bias;      // [n_total] or nullptr; added before ReLU
    float* stats_partials;  // [gridDim.x][2 * n_total] when CG_STATS / CG_BSTATS
    // CG_BSTATS (data-gradient kernels): the tile just produced is dA of a conv+BN+ReLU layer; the epilogue also
    // TMA-loads the matching tile of that layer's pre-BN output y and reduces the two BatchNorm-backward sums
    //   s1[c] = sum dz,  q[c] = sum dz * (y - mean[c]),   dz = (y*scale[c] + shift[c] > 0) ? dA : 0
    // into stats_partials, so the separate reduction pass over (y, dA) disappears.
    CUtensorMap y_map;      // same box / swizzle as d_maps[0], over y
    const float* bs_scale;  // [n_total] each
    const float* bs_shift;
    const float* bs_mean;
    int ybuf;               // y staging buffers per epilogue group (1 or 2)
    long long* dbg;         // optional: block 0 records clock64() per role / tile / event (timing forensics)
};
// Timing forensics (clock64 stamps per role / tile, ablation flags) compile in only with
// -DSDN_FORENSICS: the tcgen05 issue path is sensitive to extra predicates.
#ifdef SDN_FORENSICS
#define SDN_DBG(role, tile, ev)                                                              \
    do {                                                                                     \
        if (p.dbg != nullptr && blockIdx.x == 0 && (tile) < 16)                              \
            p.dbg[((role) * 16 + (tile)) * 8 + (ev)] = clock64();                            \
    } while (0)
#define SDN_ABLATE(flag) ((p.flags & (flag)) != 0)
#else
#define SDN_DBG(role, tile, ev) do { } while (0)
#define SDN_ABLATE(flag) false
#endif

template <int SWA, int BLOCK_N, int SWD_SEL = 0, int NCTA = 1, int EGSEL = 0>
struct CgCfg {
    static constexpr int KB = SWA / 2;  // bf16 channels per k-block
    static constexpr int A_BYTES = 128 * SWA;
    // NCTA == 2 (CTA pair, tcgen05 cta_group::2): this CTA stages HALF of the tile's B rows
    static constexpr int BN_LOC = BLOCK_N / NCTA;
    static constexpr int B_BYTES = BN_LOC * SWA;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // staging / store swizzle: 64-channel blocks, except 32-channel blocks for N = 32 and for 32-channel
    // sources (so that a 64-wide tile can be split over two 32-channel destinations: dec1 dgrad)
    // (SWD_SEL = 64 forces 32-channel blocks: a 128-wide ConvTranspose2d tile over four 32-channel quadrants)
    static constexpr int SWD = SWD_SEL ? SWD_SEL : ((BLOCK_N >= 64 && SWA == 128) ? 128 : 64);
    static constexpr int DCH = SWD / 2;                   // channels per D block
    static constexpr int D_BLOCKS = BLOCK_N / DCH;
    static constexpr int D_BLOCK_BYTES = 128 * SWD;
    static constexpr int D_BYTES = 128 * BLOCK_N * 2;
    static constexpr int WPR = DCH / 2;     // 32-bit words per staging row
    static constexpr int RG = 128 / WPR;    // row groups in the stats pass
    // small-N tiles are epilogue-bound (the per-tile bookkeeping of a 128-thread epilogue is longer than
    // their main loop): two epilogue warpgroups, one per TMEM accumulator stage, take alternate tiles
    // (EGSEL = 2: N = 128 tiles with a SHORT K loop - the ConvTranspose2d GEMMs, 1-4 k-blocks per tile - have the same
    // problem: one epilogue group with one staging buffer waited for the previous tile's TMA store to drain)
    static constexpr int EG = EGSEL ? EGSEL : (BLOCK_N <= 64 ? 2 : 1);
    // -DSDN_TWO_MMA_WARPS: two MMA-issuing warps (one per TMEM stage / epilogue group), an experiment kept for
    // reference: the pipe is not idle between tiles, so interleaving two tiles only delays both epilogues
#ifdef SDN_TWO_MMA_WARPS
    static constexpr int NMMA = EG;
#else
    static constexpr int NMMA = 1;   // measured: two issuing warps do not help (32->32 level 1: 11 % slower, 64->64: 5 % faster)
#endif
    static constexpr int EPI0 = 32 * (1 + NMMA);   // first epilogue thread
    static constexpr int THREADS = EPI0 + 128 * EG;
    // N = 32: BatchNorm statistics accumulate in registers (64 per epilogue thread); wider tiles would spill
#ifdef SDN_NO_REGSTATS
    static constexpr bool REGSTATS = false;
#else
    static constexpr bool REGSTATS = BLOCK_N == 32;
#endif
    static constexpr int SCRATCH_BYTES = EG * RG * BLOCK_N * 2 * 4;
    static constexpr int ACC_BYTES = 2 * 512 * 4;
    static constexpr int TMEM_COLS = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);
    static constexpr int DBUF = (BLOCK_N <= 64 || EGSEL == 2) ? 2 : 1;   // staging buffers (small tiles: defer the store-read wait)
    static constexpr int NT = BLOCK_N >= 128 ? 512 / BLOCK_N : 1;    // N tiles a stats layer can have (Cout <= 512)
    static constexpr int smem_bytes(int stages) {
        return 1024 + stages * STAGE_BYTES + DBUF * D_BYTES + SCRATCH_BYTES + ACC_BYTES + 256;
    }
    // HALO: one stage = one halo A box + the three vertical-tap weight blocks
    static constexpr int smem_bytes_halo(int stages, int a_stage_bytes) {
        return 1024 + stages * (a_stage_bytes + 3 * B_BYTES) + DBUF * D_BYTES + SCRATCH_BYTES + ACC_BYTES + 256;
    }
};

// HALO = true (3x3 convs, one image per box, TW % 8 == 0): a "segment" is a
// (source, horizontal tap, channel block) unit.  Its A operand is ONE TMA box of
// (TH+2) x TW pixels; the three vertical taps read that buffer at row offsets
// 0, TW, 2*TW - whole 8-row swizzle groups, so the shifted descriptors stay
// canonical - and its B operand is one 3-D box holding the three taps' weight
// blocks.  L2->SM requests per tile drop from 9*(128+N) to 3*(TW*(TH+2)+3N).
//
// HALO = 2 ("box9", TW == 8, resident weights): the hardware swizzle is a pure function of the
// shared-memory ADDRESS (probed: tests/umma_shift_probe.cu), so a descriptor may start at any row
// and step between 8-row groups by any stride.  ONE (TH+2) x (TW+2) box per (source, channel
// block) then serves all NINE taps: tap (dy, dx) starts (dy*(TW+2) + dx) rows into the box and
// strides (TW+2) rows between the 8-pixel image-row segments.  L2->SM rows per tile: 180 instead of 432.
//
// HALO = 3 ("super-pixel", 32-channel sources and 32 output channels, i.e. the level-1 convs): with N = 32 a
// tcgen05.mma costs 42 cycles for 32 columns - the A operand is re-read from shared memory for every tap - so these
// layers run at a third of the pipe.  Two horizontally adjacent pixels are ONE 64-channel super-pixel of the same
// NHWC memory ([B,H,W,32] == [B,H,W/2,64]), so the kernel sees a (W/2) x H image with 64 input and 64 output
// "channels" (output pixel pair x 32).  Per vertical tap the 3x3 conv becomes: the centre super-pixel with a dense
// 64 x 64 weight block (all four pixel pairings are valid taps dx = p_in - p_out), the left neighbour with only its
// SECOND pixel feeding the FIRST output pixel (K = 32, N = 32, columns 0-31) and the right neighbour with only its
// first pixel feeding the second output pixel (K = 32, N = 32, columns 32-63).  24 MMAs per 256 pixels instead of
// 36, a third less A traffic, and one barrier handshake per 256 pixels instead of two.  Same box9 addressing: one
// (TH+2) x (TW+2) super-pixel box per source, 128-byte rows.  Weights: [unit][dy][128 rows][64 k], rows 0-63 centre,
// 64-95 left, 96-127 right (pack modes 11 / 12).
//
// NCTA = 2 (CTA pairs, N >= 128 tiles; HALO 0 / 1): the two CTAs of a cluster take the same tile position in two
// consecutive image groups (tn = 2 * pair_tn + rank) and run it as ONE M = 256 tcgen05.mma.cta_group::2 per k-step:
// each CTA loads its own A box and HALF of the weight rows, every TMA load of the pair completes on the LEADER's
// full barrier, the leader (cluster rank 0) issues, its commits are multicast to both CTAs' barriers, and the peer's
// epilogue warps release the accumulator stage on the leader's barrier through the cluster window.  The weight bytes
// that cross L2 -> SM per FLOP halve: these layers were pinned at the L2 throughput cap (ncu: 10.5-11.4 TB/s of
// lts2xbar traffic), not at the tensor pipe.
template <bool PAIR>
__device__ __forceinline__ void cg_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                       uint32_t issue) {
    if constexpr (PAIR) ptx::tc_mma_bf16_pair_pred(tmem_d, adesc, bdesc, idesc, accumulate, issue);
    else ptx::tc_mma_bf16_pred(tmem_d, adesc, bdesc, idesc, accumulate, issue);
}
template <bool PAIR>
__device__ __forceinline__ void cg_commit(uint64_t* bar) {
    if constexpr (PAIR) ptx::tc_commit_pair(bar);
    else ptx::tc_commit(bar);
}

template <int SWA, int BLOCK_N, int HALO, int SWD_SEL = 0, int NCTA = 1, int EGSEL = 0>
__global__ void __launch_bounds__((CgCfg<SWA, BLOCK_N, SWD_SEL, NCTA, EGSEL>::THREADS), 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = CgCfg<SWA, BLOCK_N, SWD_SEL, NCTA, EGSEL>;
    constexpr int KB = Cfg::KB;
    constexpr bool PAIR = NCTA == 2;
    static_assert(!PAIR || (HALO <= 2 && Cfg::NMMA == 1), "CTA pairs: plain / row-halo / box9 kernels");
    constexpr uint32_t LAYOUT_A = (SWA == 128) ? 2u : 4u;
    constexpr uint32_t SBO_A = 8 * SWA;
    constexpr uint32_t IDESC = ptx::make_idesc_bf16(128 * NCTA, BLOCK_N, 0, 0);

    ptx::pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    uint8_t* stage_base = smem;
    const bool bres = HALO != 0 && (p.flags & CG_BRES) != 0;
    const int a_bytes = HALO ? p.a_stage_bytes : Cfg::A_BYTES;
    const int unit_bytes = HALO ? p.a_stage_bytes + (bres ? 0 : 3 * Cfg::B_BYTES) : Cfg::STAGE_BYTES;
    const int ups = HALO ? p.ups : 1;
    const int stage_bytes = ups * unit_bytes;
    uint8_t* b_res = smem + stages * stage_bytes;                 // resident weights (CG_BRES), 1024-aligned
    uint8_t* stg0 = b_res + (bres ? p.b_res_bytes : 0);           // D staging (DBUF buffers), 1024-aligned
    const bool bstats = (p.flags & CG_BSTATS) != 0;
    uint8_t* ystg0 = stg0 + Cfg::DBUF * Cfg::D_BYTES;             // y tiles (CG_BSTATS), 1024-aligned
    const int ystg_bytes = bstats ? Cfg::EG * p.ybuf * Cfg::D_BYTES : 0;
    float* scratch = reinterpret_cast<float*>(ystg0 + ystg_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ystg0 + ystg_bytes + Cfg::SCRATCH_BYTES + Cfg::ACC_BYTES);
    uint64_t* full_bar = bars;         // [8]
    uint64_t* empty_bar = bars + 8;    // [8]
    uint64_t* tfull_bar = bars + 16;   // [2]
    uint64_t* tempty_bar = bars + 18;  // [2]
    uint64_t* bres_bar = bars + 20;    // resident-weights arrival
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 22);
    uint64_t* ybar_all = bars + 24;    // [EG][2] y-tile arrival (CG_BSTATS)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // tile walking: a CTA pair is one worker (p.tiles_n counts PAIRS of image groups then)
    const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const int wid = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
    const int wstep = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int num_tiles = m_tiles * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], 4 * NCTA);   // the epilogue warps of BOTH CTAs of a pair
        }
        ptx::mbar_init(bres_bar, 1);
        for (int i = 0; i < 4; ++i) ptx::mbar_init(&ybar_all[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.a_maps[i]);
        ptx::prefetch_tmap(&p.b_map);
        for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.d_maps[i]);
        if (bstats) ptx::prefetch_tmap(&p.y_map);
    }
    if (warp == 1) {
        if (PAIR) { ptx::tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS); ptx::tmem_relinquish_pair(); }
        else { ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
    ptx::pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        // converged warp; one elected lane issues the mbarrier arrive and the TMA instructions
        {
            int s = 0;
            uint32_t ph = 0;
            int dbg_it = 0;
            if (HALO == 3) {
                if (ptx::elect_one()) {
                    // [unit][dy] slabs of 128 rows x 128 bytes (centre 64, left 32, right 32 rows)
                    const int slabs = p.kblocks_total * 3;
                    ptx::mbar_arrive_expect_tx(bres_bar, uint32_t(slabs) * 128 * 128);
                    for (int u = 0; u < slabs; ++u) ptx::tma_load_2d(b_res + u * 128 * 128, &p.b_map, bres_bar, 0, u * 128);
                }
            } else if (bres && ptx::elect_one()) {
                // every k-block's weight slabs, once: [unit][dy][BLOCK_N][KB] ([unit][dy][dx].. for box9)
                const int slabs = p.kblocks_total * (HALO == 2 ? 3 : 1);   // 3-block TMA boxes
                if (!PAIR || cta_rank == 0) ptx::mbar_arrive_expect_tx(bres_bar, uint32_t(NCTA * slabs) * 3 * Cfg::B_BYTES);
                const uint32_t bres_lead = PAIR ? ptx::mapa_u32(bres_bar, 0) : 0u;
                for (int u = 0; u < slabs; ++u) {
                    if (PAIR) ptx::tma_load_3d_pair(b_res + u * 3 * Cfg::B_BYTES, &p.b_map, bres_lead, 0, int(cta_rank) * Cfg::BN_LOC, u * 3);
                    else ptx::tma_load_3d(b_res + u * 3 * Cfg::B_BYTES, &p.b_map, bres_bar, 0, 0, u * 3);
                }
            }
            const uint32_t full_lead0 = PAIR ? ptx::mapa_u32(&full_bar[0], 0) : 0u;   // the leader's full barriers
            const bool tx_owner = !PAIR || cta_rank == 0;
            ptx::TileWalker tw;
            for (tw.init(wid, wstep, num_tiles, p.n_tiles, p.tiles_x, p.tiles_y); tw.valid(); tw.next()) {
                const int n_tile = tw.n_tile;
                const int x0 = tw.tx * p.TW, y0 = tw.ty * p.TH, n0 = (PAIR ? 2 * tw.tn + int(cta_rank) : tw.tn) * p.TN;
                int kcount = 0;
                const int dbg_tile = dbg_it++;
                if (lane == 0) SDN_DBG(0, dbg_tile, 0);
                int sub = 0;   // unit slot inside the current stage
                for (int sg = 0; sg < p.nsegs; ++sg) {
                    const CgSeg seg = p.segs[sg];
                    for (int cb = 0; cb < seg.cblocks; ++cb) {
                        uint8_t* a_dst = stage_base + s * stage_bytes + sub * unit_bytes;
                        uint8_t* b_dst = a_dst + a_bytes;
                        if (HALO == 2 || HALO == 3) {
                            const uint32_t a_box = uint32_t((p.TW + 2) * (p.TH + 2) * SWA);
                            const bool noa = SDN_ABLATE(CG_DBG_NOLOADA);
                            if (sub == 0) ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                            if (ptx::elect_one()) {
                                if (sub == 0 && tx_owner) ptx::mbar_arrive_expect_tx(&full_bar[s], uint32_t(NCTA * ups) * (noa ? 0 : a_box));
                                if (PAIR)
                                    ptx::tma_load_4d_pair(a_dst, &p.a_maps[seg.map], full_lead0 + uint32_t(s) * 8u, seg.c0 + cb * KB,
                                                          x0 - 1, y0 - 1, n0);
                                else if (!noa)
                                    ptx::tma_load_4d(a_dst, &p.a_maps[seg.map], &full_bar[s], seg.c0 + cb * KB, x0 - 1,
                                                     y0 - 1, n0);
                            }
                        } else if (HALO == 1) {
                            const uint32_t a_box = uint32_t(p.TW * (p.TH + 2) * SWA);
                            const bool noa = SDN_ABLATE(CG_DBG_NOLOADA);
                            if (sub == 0) ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                            if (ptx::elect_one()) {
                                if (sub == 0 && tx_owner)
                                    ptx::mbar_arrive_expect_tx(
                                        &full_bar[s], uint32_t(NCTA * ups) * ((noa ? 0 : a_box) + (bres ? 0 : 3 * Cfg::B_BYTES)));
                                if (PAIR) {
                                    const uint32_t fb = full_lead0 + uint32_t(s) * 8u;
                                    ptx::tma_load_4d_pair(a_dst, &p.a_maps[seg.map], fb, seg.c0 + cb * KB, x0 + seg.dx, y0 - 1, n0);
                                    if (!bres)
                                        ptx::tma_load_3d_pair(b_dst, &p.b_map, fb, 0, n_tile * BLOCK_N + int(cta_rank) * Cfg::BN_LOC,
                                                              kcount * 3);
                                } else {
                                    if (!noa)
                                        ptx::tma_load_4d(a_dst, &p.a_maps[seg.map], &full_bar[s], seg.c0 + cb * KB,
                                                         x0 + seg.dx, y0 - 1, n0);
                                    if (!bres)
                                        ptx::tma_load_3d(b_dst, &p.b_map, &full_bar[s], 0, n_tile * BLOCK_N, kcount * 3);
                                }
                            }
                        } else {
                            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                            if (ptx::elect_one()) {
                                if (tx_owner) ptx::mbar_arrive_expect_tx(&full_bar[s], NCTA * Cfg::STAGE_BYTES);
                                if (PAIR) {
                                    const uint32_t fb = full_lead0 + uint32_t(s) * 8u;
                                    ptx::tma_load_4d_pair(a_dst, &p.a_maps[seg.map], fb, seg.c0 + cb * KB, x0 + seg.dx, y0 + seg.dy, n0);
                                    ptx::tma_load_2d_pair(b_dst, &p.b_map, fb, kcount * KB, n_tile * BLOCK_N + int(cta_rank) * Cfg::BN_LOC);
                                } else {
                                    ptx::tma_load_4d(a_dst, &p.a_maps[seg.map], &full_bar[s], seg.c0 + cb * KB, x0 + seg.dx,
                                                     y0 + seg.dy, n0);
                                    ptx::tma_load_2d(b_dst, &p.b_map, &full_bar[s], kcount * KB, n_tile * BLOCK_N);
                                }
                            }
                        }
                        __syncwarp();
                        ++kcount;
                        if (lane == 0 && kcount <= 7) SDN_DBG(0, dbg_tile, kcount);
                        if (++sub == ups) {
                            sub = 0;
                            if (++s == stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp <= Cfg::NMMA) {
        // ---------------------------------------------------------- MMA issuer(s)
        // The whole warp runs the loop converged so the descriptor arithmetic stays in uniform
        // registers; only the tcgen05 instructions themselves are issued by one elected lane.
        // NMMA == 2: warp 1 takes the even tiles of this CTA (TMEM stage 0), warp 2 the odd ones (stage 1).
        {
            const int mw = warp - 1;
            const int stages_per_tile = (p.kblocks_total + ups - 1) / ups;
            // Two consumers are only safe while neither can run a whole ring ahead of the producer: an
            // mbarrier parity wait two phases early returns true at once.  Both tiles' stages must fit.
            const bool dual = Cfg::NMMA == 2 && 2 * stages_per_tile <= stages;
            const int NMMA = dual ? 2 : 1;
            int s = 0;
            uint32_t ph = 0;
            int a = dual ? mw : 0;
            uint32_t aph = 0;
            int my_tiles = wid < num_tiles ? (num_tiles - wid + wstep - 1) / wstep : 0;
            if (!dual && mw != 0) my_tiles = 0;   // the second MMA warp idles
            if (PAIR && cta_rank != 0) my_tiles = 0;   // only the pair's leader issues
            // skip the pipeline stages of the tiles the other MMA warp consumes
            auto skip_stages = [&](int n) {
                for (int i = 0; i < n; ++i)
                    if (++s == stages) { s = 0; ph ^= 1; }
            };
            if (NMMA == 2) skip_stages(mw * stages_per_tile);
            if (bres && my_tiles > 0) ptx::mbar_wait(bres_bar, 0);
            const uint32_t b_res_addr = ptx::smem_u32(b_res);
            for (int it = mw; it < my_tiles; it += NMMA) {
                if (lane == 0) SDN_DBG(1, it, 0);
                ptx::mbar_wait(&tempty_bar[a], aph ^ 1);
                ptx::tc_fence_after();
                if (lane == 0) SDN_DBG(1, it, 1);
                const uint32_t tmem_d = tmem_base + a * BLOCK_N;
                for (int kb = 0; kb < p.kblocks_total; kb += ups) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st_addr = ptx::smem_u32(stage_base + s * stage_bytes);
                    const bool leader = ptx::elect_one();   // one election per stage
                    if (SDN_ABLATE(CG_DBG_NOMMA)) {
                    } else if (HALO == 2) {
                        constexpr uint32_t ROW_PITCH = 10 * SWA;   // box row = TW + 2 = 10 pixels
                        const uint32_t lead = leader ? 1u : 0u;
                        for (int j = 0; j < ups; ++j) {
                            // one descriptor pair per unit; every tap / k-step is a compile-time offset of the
                            // 14-bit (address >> 4) field, and the issue is predicated, not branched
                            const uint64_t a0 = ptx::make_smem_desc(st_addr + j * unit_bytes, 16, ROW_PITCH, LAYOUT_A);
                            const uint64_t b0 =
                                ptx::make_smem_desc(b_res_addr + (kb + j) * 9 * Cfg::B_BYTES, 16, SBO_A, LAYOUT_A);
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                                for (int k = 0; k < KB / 16; ++k)
                                    cg_mma<PAIR>(
                                        tmem_d, a0 + uint64_t(((tap / 3) * ROW_PITCH + (tap % 3) * SWA) / 16 + 2 * k),
                                        b0 + uint64_t(tap * Cfg::B_BYTES / 16 + 2 * k), IDESC,
                                        (tap | k) != 0 ? 1u : ((kb | j) != 0 ? 1u : 0u), lead);
                            }
                        }
                    } else if (HALO == 3) {
                        constexpr uint32_t ROW_PITCH = 10 * 128;   // box row = TW + 2 = 10 super-pixels of 128 bytes
                        constexpr uint32_t IDESC32 = ptx::make_idesc_bf16(128, 32, 0, 0);
                        const uint32_t lead = leader ? 1u : 0u;
                        for (int j = 0; j < ups; ++j) {
                            const uint64_t a0 = ptx::make_smem_desc(st_addr + j * unit_bytes, 16, ROW_PITCH, 2u);
                            const uint64_t b0 = ptx::make_smem_desc(b_res_addr + (kb + j) * 3 * 128 * 128, 16, 1024, 2u);
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy) {
                                const uint32_t first = (dy != 0 || (kb | j) != 0) ? 1u : 0u;
                                // centre super-pixel: K = 64, N = 64 (all four pixel pairings)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::tc_mma_bf16_pred(tmem_d, a0 + uint64_t((dy * ROW_PITCH + 128) / 16 + 2 * k),
                                                          b0 + uint64_t(dy * (128 * 128 / 16) + 2 * k), IDESC,
                                                          k != 0 ? 1u : first, lead);
                                // left neighbour: its second pixel (bytes 64..127) -> first output pixel (columns 0..31)
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    ptx::tc_mma_bf16_pred(tmem_d, a0 + uint64_t((dy * ROW_PITCH + 64) / 16 + 2 * k),
                                                          b0 + uint64_t(dy * (128 * 128 / 16) + (64 * 128 + 64) / 16 + 2 * k),
                                                          IDESC32, 1u, lead);
                                // right neighbour: its first pixel (bytes 0..63) -> second output pixel (columns 32..63)
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    ptx::tc_mma_bf16_pred(tmem_d + 32, a0 + uint64_t((dy * ROW_PITCH + 2 * 128) / 16 + 2 * k),
                                                          b0 + uint64_t(dy * (128 * 128 / 16) + (96 * 128) / 16 + 2 * k),
                                                          IDESC32, 1u, lead);
                            }
                        }
                    } else if (HALO == 1) {
                        const uint32_t lead = leader ? 1u : 0u;
                        const uint64_t dy_step = uint64_t(uint32_t(p.TW * SWA) >> 4);
                        for (int j = 0; j < ups; ++j) {
                            const uint32_t a_addr = st_addr + j * unit_bytes;
                            const uint32_t b_addr = bres ? b_res_addr + (kb + j) * 3 * Cfg::B_BYTES : a_addr + a_bytes;
                            const uint64_t a0 = ptx::make_smem_desc(a_addr, 16, SBO_A, LAYOUT_A);
                            const uint64_t b0 = ptx::make_smem_desc(b_addr, 16, SBO_A, LAYOUT_A);
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                                for (int k = 0; k < KB / 16; ++k)
                                    cg_mma<PAIR>(tmem_d, a0 + dy * dy_step + uint64_t(2 * k),
                                                 b0 + uint64_t(dy * Cfg::B_BYTES / 16 + 2 * k), IDESC,
                                                 (dy | k) != 0 ? 1u : ((kb | j) != 0 ? 1u : 0u), lead);
                            }
                        }
                    } else {
                        const uint32_t a_addr = st_addr;
                        const uint32_t b_addr = a_addr + a_bytes;
                        const uint64_t adesc = ptx::make_smem_desc(a_addr, 16, SBO_A, LAYOUT_A);
                        const uint64_t bdesc = ptx::make_smem_desc(b_addr, 16, SBO_A, LAYOUT_A);
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k) {
                            // +32 bytes (= 16 bf16) along K inside the swizzle atom: +2 in the >>4 address field
                            cg_mma<PAIR>(tmem_d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), IDESC,
                                         k != 0 ? 1u : (kb != 0 ? 1u : 0u), leader ? 1u : 0u);
                        }
                    }
                    if (leader) cg_commit<PAIR>(&empty_bar[s]);
                    __syncwarp();
                    if (lane == 0 && kb < 5) SDN_DBG(1, it, 2 + kb);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                if (ptx::elect_one()) cg_commit<PAIR>(&tfull_bar[a]);
                if (lane == 0) SDN_DBG(1, it, 7);
                if (NMMA == 2) {
                    aph ^= 1;
                    skip_stages(stages_per_tile);
                } else {
                    a ^= 1;
                    if (a == 0) aph ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        constexpr int EG = Cfg::EG;
        const int eg = EG == 2 ? int(threadIdx.x - Cfg::EPI0) >> 7 : 0;   // epilogue group (== its TMEM stage when EG == 2)
        const int te = (threadIdx.x - Cfg::EPI0) & 127;                   // 0..127 inside the group
        const bool dbg_lead = threadIdx.x == Cfg::EPI0;
        (void)dbg_lead;
        const int quarter = warp & 3;        // TMEM lane quarter this warp may read
        const int r = quarter * 32 + lane;   // tile row == pixel index in the box
        const bool do_stats = (p.flags & CG_STATS) != 0 && !SDN_ABLATE(CG_DBG_NOSTATS);
        const bool do_relu = (p.flags & CG_RELU) != 0;
        int a = eg;
        uint32_t aph = 0;
        const int rw = r % p.TW, rh = (r / p.TW) % p.TH, rn = r / (p.TW * p.TH);  // this thread's pixel in the box
        constexpr int STAT_ROWS = 128 / Cfg::RG;
        const int st_w = te % Cfg::WPR, st_rg = te / Cfg::WPR;
        float st_acc[Cfg::NT][Cfg::D_BLOCKS][4];
#pragma unroll
        for (int i = 0; i < Cfg::NT; ++i)
#pragma unroll
            for (int j = 0; j < Cfg::D_BLOCKS; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) st_acc[i][j][k] = 0.f;
        // N = 32: statistics stay in REGISTERS - this thread's row, every channel - across all of
        // its tiles; no shared-memory pass per tile.  (N >= 128 keeps the column-slice pass below.)
        constexpr int RSN = Cfg::REGSTATS ? BLOCK_N : 1;
        float rs[RSN], rq[RSN];
#pragma unroll
        for (int j = 0; j < RSN; ++j) { rs[j] = 0.f; rq[j] = 0.f; }
        int sbuf = 0;
        int dbg_it = 0;
        const int dmap_div = p.n_per_dmap;   // channels per destination map
        // CG_BSTATS: y tiles of this group's tiles, fetched `ybuf` tiles ahead by the group's first thread
        uint8_t* ystg = ystg0 + eg * p.ybuf * Cfg::D_BYTES;
        uint64_t* ybar = ybar_all + eg * 2;
        const int ybufs = bstats ? p.ybuf : 1;
        ptx::TileWalker twy;
        auto issue_y = [&](int buf) {
            ptx::mbar_arrive_expect_tx(&ybar[buf], Cfg::D_BYTES);
            const int c0 = twy.n_tile * BLOCK_N;
#pragma unroll
            for (int cbk = 0; cbk < Cfg::D_BLOCKS; ++cbk)
                ptx::tma_load_4d(ystg + buf * Cfg::D_BYTES + cbk * Cfg::D_BLOCK_BYTES, &p.y_map, &ybar[buf],
                                 c0 + cbk * Cfg::DCH, twy.tx * p.TW, twy.ty * p.TH,
                                 (PAIR ? 2 * twy.tn + int(cta_rank) : twy.tn) * p.TN);
        };
        const int tn_mul = PAIR ? 2 : 1, tn_add = PAIR ? int(cta_rank) : 0;   // this CTA's image group of the pair's tile
        if (bstats && te == 0) {
            twy.init(wid + eg * wstep, EG * wstep, num_tiles, p.n_tiles, p.tiles_x, p.tiles_y);
            for (int b = 0; b < ybufs && twy.valid(); ++b, twy.next()) issue_y(b);
        }
        int yk = 0;   // tiles this group has processed
        ptx::TileWalker tw;
        const uint32_t tempty_lead0 = PAIR ? ptx::mapa_u32(&tempty_bar[0], 0) : 0u;
        for (tw.init(wid + eg * wstep, EG * wstep, num_tiles, p.n_tiles, p.tiles_x, p.tiles_y); tw.valid();
             tw.next()) {
            const int n_tile = tw.n_tile;
            const int x0 = tw.tx * p.TW, y0 = tw.ty * p.TH, n0 = (tw.tn * tn_mul + tn_add) * p.TN;

            // A box row outside the image still sees in-image neighbours through the
            // shifted taps, so its accumulator is not zero: zero it (the TMA store
            // clips it anyway, but the BatchNorm statistics must not see it).
            const bool row_in_image = (x0 + rw < p.img_w) && (y0 + rh < p.img_h) && (n0 + rn < p.img_n);
            uint8_t* stg = stg0 + (EG == 2 ? eg : sbuf) * Cfg::D_BYTES;
            const int dbg_tile = dbg_it++;
            if (dbg_lead) SDN_DBG(2, dbg_tile, 0);
            // the store issued DBUF tiles ago has finished reading this staging buffer
            if (te == 0) {
                if (EG == 1 && Cfg::DBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else ptx::tma_store_wait_read0();   // EG == 2: this group's previous store, a whole tile ago
            }
            ptx::named_bar_sync(1 + eg, 128);
            if (dbg_lead) SDN_DBG(2, dbg_tile, 1);
            // every thread of the group is past the previous tile's statistics pass: its y buffer is free
            if (bstats && te == 0 && yk >= 1 && twy.valid()) { issue_y((yk - 1) % ybufs); twy.next(); }

            ptx::mbar_wait(&tfull_bar[a], aph);
            ptx::tc_fence_after();
            if (dbg_lead) SDN_DBG(2, dbg_tile, 2);
            const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + a * BLOCK_N;
            if (!SDN_ABLATE(CG_DBG_NOEPI))
#pragma unroll
            for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(taddr + ch * 32, v);
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (p.bias != nullptr) {
                    const float* b = p.bias + n_tile * BLOCK_N + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldg(b + j);
                }
                if (do_relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (!row_in_image) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = 0.f;
                }
                const int cbk = (ch * 32) / Cfg::DCH;
                const int j0 = ((ch * 32) % Cfg::DCH) / 8;
                uint8_t* rowp = stg + cbk * Cfg::D_BLOCK_BYTES + r * Cfg::SWD;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 q;
                    q.x = ptx::pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
                    q.y = ptx::pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
                    q.z = ptx::pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
                    q.w = ptx::pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
                    const int sw = (Cfg::SWD == 128) ? ((j0 + i) ^ (r & 7)) : ((j0 + i) ^ ((r >> 1) & 3));
                    *reinterpret_cast<uint4*>(rowp + (sw << 4)) = q;
                    if (Cfg::REGSTATS && do_stats) {
                        // sums of the bf16-ROUNDED values (what the next kernel reads back)
                        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float lo = __uint_as_float(w4[h] << 16), hi = __uint_as_float(w4[h] & 0xFFFF0000u);
                            const int c = Cfg::REGSTATS ? ch * 32 + 8 * i + 2 * h : 0;
                            rs[c % RSN] += lo; rq[c % RSN] = fmaf(lo, lo, rq[c % RSN]);
                            rs[(c + 1) % RSN] += hi; rq[(c + 1) % RSN] = fmaf(hi, hi, rq[(c + 1) % RSN]);
                        }
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && cta_rank != 0) ptx::mbar_arrive_cluster(tempty_lead0 + uint32_t(a) * 8u);
                else ptx::mbar_arrive(&tempty_bar[a]);
            }
            if (dbg_lead) SDN_DBG(2, dbg_tile, 3);
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1 + eg, 128);
            if (dbg_lead) SDN_DBG(2, dbg_tile, 4);

            if (te == 0 && !SDN_ABLATE(CG_DBG_NOSTORE)) {
                // channel block -> (destination map, channel inside it) without a division (<= 4 maps)
                int dmap = 0, cbase = n_tile * BLOCK_N;
                while (cbase >= dmap_div) { cbase -= dmap_div; ++dmap; }
#pragma unroll
                for (int cbk = 0; cbk < Cfg::D_BLOCKS; ++cbk) {
                    ptx::tma_store_4d(&p.d_maps[dmap], stg + cbk * Cfg::D_BLOCK_BYTES, cbase, x0, y0, n0);
                    cbase += Cfg::DCH;
                    if (cbase >= dmap_div) { cbase -= dmap_div; ++dmap; }
                }
                ptx::tma_store_commit();
            }
            if (dbg_lead) SDN_DBG(2, dbg_tile, 6);
            if (bstats) {
                // BatchNorm-backward sums of the layer whose output gradient this tile is (see ConvGemmParams):
                // each thread owns one 32-bit word column (2 channels) of one row group; dA from the staged bf16
                // values (what the apply pass will read back), y from the TMA-loaded tile at the same offsets.
                const int yb = yk % ybufs;
                ptx::mbar_wait(&ybar[yb], uint32_t(yk / ybufs) & 1u);
#pragma unroll
                for (int cbk = 0; cbk < Cfg::D_BLOCKS; ++cbk) {
                    const int ch0 = n_tile * BLOCK_N + cbk * Cfg::DCH + 2 * st_w;
                    const float sc0 = __ldg(p.bs_scale + ch0), sc1 = __ldg(p.bs_scale + ch0 + 1);
                    const float sh0 = __ldg(p.bs_shift + ch0), sh1 = __ldg(p.bs_shift + ch0 + 1);
                    const float mu0 = __ldg(p.bs_mean + ch0), mu1 = __ldg(p.bs_mean + ch0 + 1);
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                    const uint8_t* blk = stg + cbk * Cfg::D_BLOCK_BYTES;
                    const uint8_t* yblk = ystg + yb * Cfg::D_BYTES + cbk * Cfg::D_BLOCK_BYTES;
#pragma unroll 8
                    for (int rr = 0; rr < STAT_ROWS; ++rr) {
                        const int row = st_rg * STAT_ROWS + rr;
                        const int j = st_w >> 2;
                        const int sw = (Cfg::SWD == 128) ? (j ^ (row & 7)) : (j ^ ((row >> 1) & 3));
                        const int off = row * Cfg::SWD + (sw << 4) + ((st_w & 3) << 2);
                        const uint32_t u = *reinterpret_cast<const uint32_t*>(blk + off);
                        const uint32_t yv = *reinterpret_cast<const uint32_t*>(yblk + off);
                        const float g_lo = __uint_as_float(u << 16), g_hi = __uint_as_float(u & 0xFFFF0000u);
                        const float y_lo = __uint_as_float(yv << 16), y_hi = __uint_as_float(yv & 0xFFFF0000u);
                        const float d_lo = fmaf(y_lo, sc0, sh0) > 0.f ? g_lo : 0.f;
                        const float d_hi = fmaf(y_hi, sc1, sh1) > 0.f ? g_hi : 0.f;
                        s0 += d_lo; q0 = fmaf(d_lo, y_lo - mu0, q0);
                        s1 += d_hi; q1 = fmaf(d_hi, y_hi - mu1, q1);
                    }
#pragma unroll
                    for (int nt = 0; nt < Cfg::NT; ++nt)
                        if (nt == n_tile || Cfg::NT == 1) {
                            st_acc[nt][cbk][0] += s0; st_acc[nt][cbk][1] += s1;
                            st_acc[nt][cbk][2] += q0; st_acc[nt][cbk][3] += q1;
                        }
                }
            } else if (!Cfg::REGSTATS && do_stats) {
                // per-channel sum / sum of squares of the bf16 values just staged (== what the
                // next kernel reads back).  Each thread owns one 32-bit word column (2 channels)
                // of one row group and keeps its partials in REGISTERS across tiles; the
                // cross-thread reduction happens once, after the tile loop.
#pragma unroll
                for (int cbk = 0; cbk < Cfg::D_BLOCKS; ++cbk) {
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                    const uint8_t* blk = stg + cbk * Cfg::D_BLOCK_BYTES;
#pragma unroll 8
                    for (int rr = 0; rr < STAT_ROWS; ++rr) {
                        const int row = st_rg * STAT_ROWS + rr;
                        const int j = st_w >> 2;
                        const int sw = (Cfg::SWD == 128) ? (j ^ (row & 7)) : (j ^ ((row >> 1) & 3));
                        const uint32_t u =
                            *reinterpret_cast<const uint32_t*>(blk + row * Cfg::SWD + (sw << 4) + ((st_w & 3) << 2));
                        const float lo = __uint_as_float(u << 16);
                        const float hi = __uint_as_float(u & 0xFFFF0000u);
                        s0 += lo; q0 = fmaf(lo, lo, q0);
                        s1 += hi; q1 = fmaf(hi, hi, q1);
                    }
#pragma unroll
                    for (int nt = 0; nt < Cfg::NT; ++nt)
                        if (nt == n_tile || Cfg::NT == 1) {
                            st_acc[nt][cbk][0] += s0; st_acc[nt][cbk][1] += s1;
                            st_acc[nt][cbk][2] += q0; st_acc[nt][cbk][3] += q1;
                        }
                }
            }
            if (dbg_lead) SDN_DBG(2, dbg_tile, 5);
            sbuf = (Cfg::DBUF == 2) ? (sbuf ^ 1) : 0;
            ++yk;
            if (EG == 2) {
                aph ^= 1;
            } else {
                a ^= 1;
                if (a == 0) aph ^= 1;
            }
        }
        if (Cfg::REGSTATS && do_stats && !bstats) {
            // one reduction per kernel: 32 rows by warp shuffle, the group's 4 warps through scratch
            float* dst = p.stats_partials + size_t(blockIdx.x * EG + eg) * 2 * p.n_total;
            scratch += eg * (Cfg::SCRATCH_BYTES / 4 / EG);
            const int wq = (threadIdx.x >> 5) & 3;
#pragma unroll
            for (int c = 0; c < RSN; ++c) {
                float a = rs[c], b = rq[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane == 0) { scratch[wq * 2 * BLOCK_N + c] = a; scratch[wq * 2 * BLOCK_N + BLOCK_N + c] = b; }
            }
            ptx::named_bar_sync(3 + eg, 128);
            for (int o = te; o < 2 * BLOCK_N; o += 128) {
                const float v = (scratch[o] + scratch[2 * BLOCK_N + o]) + (scratch[4 * BLOCK_N + o] + scratch[6 * BLOCK_N + o]);
                const int kind = o / BLOCK_N, c = o % BLOCK_N;
                if (c < p.n_total) dst[kind * p.n_total + c] = v;
            }
        } else if (do_stats || bstats) {
            // one cross-row-group reduction per kernel: scratch[rg][channel] -> per-CTA partials
            float* dst = p.stats_partials + size_t(blockIdx.x * EG + eg) * 2 * p.n_total;
            scratch += eg * (Cfg::SCRATCH_BYTES / 4 / EG);
#pragma unroll
            for (int nt = 0; nt < Cfg::NT; ++nt) {
                if (nt * BLOCK_N >= p.n_total) break;
                ptx::named_bar_sync(3 + eg, 128);
#pragma unroll
                for (int cbk = 0; cbk < Cfg::D_BLOCKS; ++cbk) {
                    const int chn = cbk * Cfg::DCH + 2 * st_w;
                    scratch[st_rg * BLOCK_N + chn] = st_acc[nt][cbk][0];
                    scratch[st_rg * BLOCK_N + chn + 1] = st_acc[nt][cbk][1];
                    scratch[Cfg::RG * BLOCK_N + st_rg * BLOCK_N + chn] = st_acc[nt][cbk][2];
                    scratch[Cfg::RG * BLOCK_N + st_rg * BLOCK_N + chn + 1] = st_acc[nt][cbk][3];
                }
                ptx::named_bar_sync(3 + eg, 128);
                for (int c = te; c < BLOCK_N; c += 128) {
                    float sum = 0.f, sq = 0.f;
#pragma unroll
                    for (int g = 0; g < Cfg::RG; ++g) {
                        sum += scratch[g * BLOCK_N + c];
                        sq += scratch[Cfg::RG * BLOCK_N + g * BLOCK_N + c];
                    }
                    dst[nt * BLOCK_N + c] = sum;
                    dst[p.n_total + nt * BLOCK_N + c] = sq;
                }
            }
        }
        if (te == 0) ptx::tma_store_wait0();
    }

    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync_all();   // the leader's MMAs read the peer's shared memory and signal its barriers
    else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace sdn
