// Host-side executor of the B200-native stereo U-Net step and its C ABI
// (include/sdn.h).  It owns the layer plan, the TMA tensor maps, the bf16
// operand cache and the activation workspace; it launches only the kernels in
// conv_gemm.cuh / wgrad_gemm.cuh / elementwise.cuh / preprocess.cuh.  There is
// no library GEMM/conv call and no CPU fallback anywhere on this path.
//
// Topology mirrored (not copied) from the reference model
// src/foundation_stereo_depth/model.py:48-104:
//   enc1..enc4, bottleneck, (up4,dec4) .. (up1,dec1), disparity/logvar heads.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is resolved at run time (sdn_comm_init), never linked

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <vector>

#include "../../include/sdn.h"
#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "preprocess.cuh"
#include "live.cuh"
#include "wgrad_gemm.cuh"
#include "wgrad_tr.cuh"

using namespace sdn;

static_assert(sizeof(sdn_aug_params) == sizeof(AugParams), "ABI struct mismatch");

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}
#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define SDN_OK(call)              \
    do {                          \
        int r_ = (call);          \
        if (r_ != 0) return r_;   \
    } while (0)
// every C-ABI entry: select the context's device and its launch mode for this host thread
#define SDN_ENTER(c)                            \
    do {                                        \
        CUDA_OK(cudaSetDevice((c)->device));    \
        t_pdl = (c)->pdl;                       \
    } while (0)

// ------------------------------------------------------- tensor-map encode
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

static CUtensorMapSwizzle swz(int bytes) {
    return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                   : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// 4-D bf16 map over (C, W, H, N) with explicit element strides for W, H, N.
static int encode4(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW, long long sH,
                   long long sN, int bC, int bW, int bH, int bN, int swizzle_bytes) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
    cuuint32_t box[4] = {(cuuint32_t)bC, (cuuint32_t)bW, (cuuint32_t)bH, (cuuint32_t)bN};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail("cuTensorMapEncodeTiled(4d) failed: %d (C=%d W=%d H=%d N=%d box=%d,%d,%d,%d sw=%d)", (int)r, C, W,
                    H, N, bC, bW, bH, bN, swizzle_bytes);
    return 0;
}
static int encode2(CUtensorMap* m, const void* base, int K, int Nrows, int bK, int bN, int swizzle_bytes) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Nrows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)bK, (cuuint32_t)bN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail("cuTensorMapEncodeTiled(2d) failed: %d (K=%d N=%d box=%d,%d sw=%d)", (int)r, K, Nrows, bK, bN,
                    swizzle_bytes);
    return 0;
}

// 3-D weight map for the row-halo kernels: (KB, N rows, K blocks); box (KB, bN, 3)
static int encode3(CUtensorMap* m, const void* base, int KB, int Nrows, int kblocks, int bN, int swizzle_bytes) {
    cuuint64_t dims[3] = {(cuuint64_t)KB, (cuuint64_t)Nrows, (cuuint64_t)kblocks};
    cuuint64_t strides[2] = {(cuuint64_t)kblocks * KB * 2, (cuuint64_t)KB * 2};
    cuuint32_t box[3] = {(cuuint32_t)KB, (cuuint32_t)bN, 3};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail("cuTensorMapEncodeTiled(3d) failed: %d (KB=%d N=%d kb=%d box=%d sw=%d)", (int)r, KB, Nrows, kblocks,
                    bN, swizzle_bytes);
    return 0;
}

// ------------------------------------------------------------- tile shapes
struct Tile {
    int TW, TH, TN;
};
// Pixel box (TW x TH x TN images) with TW*TH*TN == target that wastes the least
// work on ragged edges; ties -> fewer images per box, then the squarest box.
static Tile choose_tile(int W, int H, int B, int target) {
    Tile best{target, 1, 1};
    double best_cost = 1e300;
    for (int TN = 1; TN <= target; TN *= 2)
        for (int TW = 2; TW * TN <= target; TW *= 2) {
            const int TH = target / (TW * TN);
            if (TH < 1 || TW * TH * TN != target) continue;
            if (TW > 256 || TH > 256) continue;
            const double padded = double((W + TW - 1) / TW * TW) * double((H + TH - 1) / TH * TH) *
                                  double((B + TN - 1) / TN * TN);
            const double cost = padded * 1e6 + TN * 1e3 + double((TW + 2) * (TH + 2));
            if (cost < best_cost) { best_cost = cost; best = Tile{TW, TH, TN}; }
        }
    return best;
}

// --------------------------------------------------------------- tensors
struct Act {
    bf16* p = nullptr;
    int C = 0, H = 0, W = 0;  // per image; batch is the context's current B
    size_t elems(int B) const { return (size_t)B * H * W * C; }
};

struct GemmOp {
    ConvGemmParams p;
    int swa = 128, block_n = 32, grid = 1, smem = 0;
    bool swd64 = false;  // 128-wide tile stored as four 32-channel blocks (ConvTranspose2d with Cout = 32)
    int eg = 1;    // epilogue warpgroups of the kernel (partial-statistics rows per CTA)
    int ncta = 1;  // 2: CTA pairs (tcgen05 cta_group::2, cluster of two CTAs, each stages half of the weight rows)
    int halo = 0;  // 3x3 convs: 1 = row-halo A boxes (one per horizontal tap), 2 = one box for all nine taps; the packed weights use the matching K order
};
struct WgradOp {
    WgradParams p;
    int tr2_natoms = 0;   // > 0: wgrad_tr_kernel<CA, Cout, natoms> (3x3, Cout <= 64)
    int tr2_ndx = 3;      // 1: vertical taps only (first layer, row-halo form)
    int swb = 128, smem = 0;
    dim3 grid;
};

struct ConvL {  // conv3x3 + BatchNorm + ReLU
    int cin = 0, cout = 0, lvl = 0;
    int nsrc = 1;
    const Act* src[2] = {nullptr, nullptr};
    bool pooled_out = false;   // block output that is also max-pooled
    bool first = false;        // enc1.block.0: im2col input, K = 64
    int p_w = 0, p_gamma = 0, p_beta = 0, bn = 0;
    Act y, a, pool, dy, ga, gp;
    bf16 *wf = nullptr, *wd = nullptr;
    float *scale = nullptr, *shift = nullptr, *mean = nullptr, *rstd = nullptr, *c1 = nullptr, *c2 = nullptr;
    float* wg = nullptr;  // fp32 weight-gradient workspace [9*cin][cout]
    unsigned short* amax = nullptr;  // pooled layers: 2-bit arg-max per (quad, channel), 8 channels per entry
    GemmOp fprop, fprop_eval, dgrad;   // fprop_eval: BN folded, epilogue = +shift, ReLU, writes `a` directly
    WgradOp wgrad;
    bool has_dgrad = true;
    // the kernel that writes `ga` (the next layer's dgrad, or a ConvTranspose2d dgrad) also reduced this layer's
    // BatchNorm-backward sums (CG_BSTATS): bn_backward skips its reduction pass and finalizes from these partials
    bool bwd_stats_fused = false;
    int bwd_stats_parts = 0;
    int bwd_stats_fold = 1;   // 2: the producer ran in the super-pixel view (partial rows hold 2 x C columns)
};
struct UpL {  // ConvTranspose2d(k=2, s=2)
    bool wgrad_pairs = false;   // dY variants of the weight gradient are quadrant pairs (workspace layout of unpack mode 5)
    int cin = 0, cout = 0, lvl_in = 0;
    const Act* src = nullptr;
    int src_layer = 0;
    int p_w = 0, p_b = 0;
    Act u, gu;
    bf16 *wf = nullptr, *wd = nullptr;
    float* bias4 = nullptr;
    float* wg = nullptr;  // [4][cin][cout]
    float* bg = nullptr;  // [cout]
    GemmOp fprop, dgrad;
    WgradOp wgrad;
    bool bias_fused = false;  // bias gradient comes from the column sums of the dgrad that writes `gu`
};

struct sdn_ctx {
    int device = 0, maxB = 0, H = 0, W = 0, num_sms = 148;
    int B = 0;  // batch the tensor maps are currently encoded for
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    ConvL conv[18];
    UpL up[4];
    Act x0;     // im2col of the network input [B,H,W,64]
    float* stats_partials = nullptr;  // [num_sms][2*512]
    float* bwd_partials = nullptr;    // [BWD_BLOCKS][2*512]
    float* head_grads = nullptr;      // 66 floats
    float* wg_all = nullptr;          // all weight-gradient workspaces, contiguous
    size_t wg_all_bytes = 0;
    unsigned long long* n_local = nullptr;  // device u64
    float* gray_part = nullptr;
    float* blur_tmp = nullptr;
    float* aug_stage = nullptr;   // device copy of pinned-host augmentation parameters (2*maxB structs)
    float* view_mean = nullptr;   // per-view gray mean (2*maxB floats)
    size_t blur_tmp_elems = 0;
    const float* params[SDN_NUM_PARAMS] = {};
    float* grads[SDN_NUM_PARAMS] = {};
    float* bn_rm[SDN_NUM_BN] = {};
    float* bn_rv[SDN_NUM_BN] = {};
    int64_t* bn_nbt[SDN_NUM_BN] = {};
    bool have_params = false, have_forward_train = false;
    bool pre_only = false;  // SDN_CTX_PREPROCESS_ONLY: no network workspace
    long long* dbg = nullptr;  // timing forensics buffer (3 roles x 16 tiles x 8 events)
    int accumulate = 0;
    int pdl = 0;   // programmatic dependent launch for this context's kernels (small per-GPU batches only)
    int64_t launches = 0;
    // optional per-op timing (CUDA events on the launching stream)
    bool prof = false;
    // Backward overlap: weight gradients run on a low-priority side stream, next to the memory-bound
    // BatchNorm backward of the following layer (dgrad -> BN backward is the critical path; a wgrad and a
    // dgrad cannot share an SM, a wgrad and the shared-memory-free BN kernels can).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pack = nullptr;
    bool side_dirty = false;
    // Data parallelism (sdn_comm_init): one NCCL communicator per context; gradient buckets are all-reduced on
    // `comm_stream`, forked from the caller's stream after each backward stage and joined at the end of the step.
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_bucket = nullptr, ev_comm = nullptr;
    struct ProfRec {
        const char* name;
        int layer;
        cudaEvent_t a, b;
        double flops, bytes;
    };
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
};

// Times everything enqueued on `st` during its lifetime when profiling is on.
struct ProfScope {
    sdn_ctx* c;
    cudaStream_t st;
    int idx = -1;
    ProfScope(sdn_ctx* c_, cudaStream_t st_, const char* name, int layer, double flops, double bytes) : c(c_), st(st_) {
        if (!c->prof) return;
        sdn_ctx::ProfRec r{name, layer, c->next_event(), c->next_event(), flops, bytes};
        cudaEventRecord(r.a, st);
        idx = (int)c->recs.size();
        c->recs.push_back(r);
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(c->recs[idx].b, st);
    }
};

static const int BWD_BLOCKS = 592;  // 4 x 148

// level geometry
static inline int lvl_h(const sdn_ctx* c, int lvl) { return c->H >> (lvl - 1); }
static inline int lvl_w(const sdn_ctx* c, int lvl) { return c->W >> (lvl - 1); }

// -------------------------------------------------------------- launchers
// Every kernel is launched with programmatic stream serialization (PDL): each kernel signals
// `griddepcontrol.launch_dependents` at entry and blocks on `griddepcontrol.wait` before it touches
// memory, so the NEXT kernel's launch and prologue (barrier init, TMEM allocation, descriptor prefetch)
// overlap the tail of the previous one.  SDN_PDL=0 / 1 forces it off / on.
// measured: +1.3 % at 32 pairs per GPU, neutral at 64, -1 % at 128, -4 % at 256 (parked dependents take SM
// resources from long memory-bound kernels), so the default follows the batch
// The switch belongs to the CONTEXT (sdn_ctx::pdl, chosen from its batch); every C-ABI entry copies it into this
// thread-local before it launches anything, so two contexts / two host threads never see each other's choice.
static thread_local int t_pdl = 0;
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static thread_local int t_cluster = 1;   // set by launch_cg around a CTA-pair launch (cluster dimension x)
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const int pdl_env = env_int("SDN_PDL", -1);
    const int pdl = pdl_env >= 0 ? pdl_env : t_pdl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (t_cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)t_cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

template <int SWA, int BN, int HALO>
static int launch_cg_t(sdn_ctx* c, const GemmOp& op, cudaStream_t st) {
    launch_k(conv_gemm_kernel<SWA, BN, HALO>, op.grid, CgCfg<SWA, BN>::THREADS, op.smem, st, op.p);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}
template <int SWA, int BN, int HALO, int EGSEL = 0>
static int launch_cg_pair(sdn_ctx* c, const GemmOp& op, cudaStream_t st) {
    t_cluster = 2;
    const cudaError_t e = launch_k(conv_gemm_kernel<SWA, BN, HALO, 0, 2, EGSEL>, op.grid, CgCfg<SWA, BN, 0, 2, EGSEL>::THREADS, op.smem, st, op.p);
    t_cluster = 1;
    ++c->launches;
    CUDA_OK(e);
    CUDA_OK(cudaGetLastError());
    return 0;
}
static int launch_cg(sdn_ctx* c, const GemmOp& op, cudaStream_t st) {
    if (op.eg == 2 && op.block_n == 128) {
        if (!(op.swa == 128 && op.halo <= 1 && !op.swd64)) return fail("two epilogue groups at N = 128: plain / row-halo kernels only");
        if (op.halo == 1 && op.ncta == 2) return launch_cg_pair<128, 128, 1, 2>(c, op, st);
        if (op.ncta == 2) return fail("two epilogue groups at N = 128: no pair form of the plain kernel");
        if (op.halo == 1) launch_k(conv_gemm_kernel<128, 128, 1, 0, 1, 2>, op.grid, CgCfg<128, 128, 0, 1, 2>::THREADS, op.smem, st, op.p);
        else launch_k(conv_gemm_kernel<128, 128, 0, 0, 1, 2>, op.grid, CgCfg<128, 128, 0, 1, 2>::THREADS, op.smem, st, op.p);
        ++c->launches;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (op.ncta == 2) {
        if (op.swa == 128 && op.block_n == 256 && op.halo == 0) return launch_cg_pair<128, 256, 0>(c, op, st);
        if (op.swa == 128 && op.block_n == 128 && op.halo == 0) return launch_cg_pair<128, 128, 0>(c, op, st);
        if (op.swa == 128 && op.block_n == 128 && op.halo == 1) return launch_cg_pair<128, 128, 1>(c, op, st);
        if (op.swa == 64 && op.block_n == 64 && op.halo == 2) return launch_cg_pair<64, 64, 2>(c, op, st);
        if (op.swa == 128 && op.block_n == 64 && op.halo == 2) return launch_cg_pair<128, 64, 2>(c, op, st);
        return fail("no CTA-pair conv_gemm instantiation for swizzle %d, BLOCK_N %d, halo %d", op.swa, op.block_n, op.halo);
    }
    if (op.swd64) {
        if (op.swa != 128 || op.block_n != 128 || op.halo != 0) return fail("swd64 needs the plain <128,128> kernel");
        launch_k(conv_gemm_kernel<128, 128, 0, 64>, op.grid, CgCfg<128, 128, 64>::THREADS, op.smem, st, op.p);
        ++c->launches;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (op.halo == 3) return launch_cg_t<128, 64, 3>(c, op, st);
    if (op.halo == 2) {
        if (op.swa == 128) {
            switch (op.block_n) {
                case 32: return launch_cg_t<128, 32, 2>(c, op, st);
                case 64: return launch_cg_t<128, 64, 2>(c, op, st);
            }
        } else if (op.swa == 64) {
            switch (op.block_n) {
                case 32: return launch_cg_t<64, 32, 2>(c, op, st);
                case 64: return launch_cg_t<64, 64, 2>(c, op, st);
            }
        }
        return fail("no box9 conv_gemm instantiation for swizzle %d, BLOCK_N %d", op.swa, op.block_n);
    }
    if (op.halo) {
        if (op.swa == 128) {
            switch (op.block_n) {
                case 32: return launch_cg_t<128, 32, 1>(c, op, st);
                case 64: return launch_cg_t<128, 64, 1>(c, op, st);
                case 128: return launch_cg_t<128, 128, 1>(c, op, st);
            }
        } else if (op.swa == 64) {
            switch (op.block_n) {
                case 32: return launch_cg_t<64, 32, 1>(c, op, st);
                case 64: return launch_cg_t<64, 64, 1>(c, op, st);
            }
        }
        return fail("no row-halo conv_gemm instantiation for swizzle %d, BLOCK_N %d", op.swa, op.block_n);
    }
    if (op.swa == 128) {
        switch (op.block_n) {
            case 32: return launch_cg_t<128, 32, 0>(c, op, st);
            case 64: return launch_cg_t<128, 64, 0>(c, op, st);
            case 128: return launch_cg_t<128, 128, 0>(c, op, st);
            case 256: return launch_cg_t<128, 256, 0>(c, op, st);
        }
    } else if (op.swa == 64) {
        switch (op.block_n) {
            case 32: return launch_cg_t<64, 32, 0>(c, op, st);
            case 64: return launch_cg_t<64, 64, 0>(c, op, st);
        }
    }
    return fail("no conv_gemm instantiation for swizzle %d, BLOCK_N %d", op.swa, op.block_n);
}
static int launch_wg(sdn_ctx* c, const WgradOp& op, cudaStream_t st) {
    if (op.tr2_natoms > 0) {
        const int ca = op.swb / 2;
        if (op.tr2_ndx == 1) {
            if (!(ca == 32 && op.p.cout == 32 && op.tr2_natoms == 1)) return fail("vertical-only wgrad_tr needs CA 32, Cout 32, one atom");
            launch_k(wgrad_tr_kernel<32, 32, 1, 1>, op.grid, 192, op.smem, st, op.p);
        } else if (ca == 32 && op.p.cout == 32 && op.tr2_natoms == 1) launch_k(wgrad_tr_kernel<32, 32, 1>, op.grid, 192, op.smem, st, op.p);
        else if (ca == 32 && op.p.cout == 32 && op.tr2_natoms == 2) launch_k(wgrad_tr_kernel<32, 32, 2>, op.grid, 192, op.smem, st, op.p);
        else if (ca == 32 && op.p.cout == 64 && op.tr2_natoms == 1) launch_k(wgrad_tr_kernel<32, 64, 1>, op.grid, 192, op.smem, st, op.p);
        else if (ca == 64 && op.p.cout == 64 && op.tr2_natoms == 1) launch_k(wgrad_tr_kernel<64, 64, 1>, op.grid, 192, op.smem, st, op.p);
        else if (ca == 64 && op.p.cout == 32 && op.tr2_natoms == 1) launch_k(wgrad_tr_kernel<64, 32, 1>, op.grid, 192, op.smem, st, op.p);
        else return fail("no wgrad_tr instantiation for CA %d, Cout %d, atoms %d", ca, op.p.cout, op.tr2_natoms);
        ++c->launches;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (op.swb == 128) {
        if (op.p.tr) launch_k(wgrad_gemm_kernel<128, true, true>, op.grid, 192, op.smem, st, op.p);
        else if (op.p.halo) launch_k(wgrad_gemm_kernel<128, true>, op.grid, 192, op.smem, st, op.p);
        else launch_k(wgrad_gemm_kernel<128, false>, op.grid, 192, op.smem, st, op.p);
    } else {
        if (op.p.tr) launch_k(wgrad_gemm_kernel<64, true, true>, op.grid, 192, op.smem, st, op.p);
        else if (op.p.halo) launch_k(wgrad_gemm_kernel<64, true>, op.grid, 192, op.smem, st, op.p);
        else launch_k(wgrad_gemm_kernel<64, false>, op.grid, 192, op.smem, st, op.p);
    }
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}
static int g_max_pairs = 0;   // co-resident CTA pairs of the pair kernels (cudaOccupancyMaxActiveClusters)
static int set_smem_attrs() {
    const int big = 227 * 1024;
#define SDN_SMEM_ATTR(...) CUDA_OK(cudaFuncSetAttribute((conv_gemm_kernel<__VA_ARGS__>), cudaFuncAttributeMaxDynamicSharedMemorySize, big))
    SDN_SMEM_ATTR(128, 32, 0); SDN_SMEM_ATTR(128, 64, 0); SDN_SMEM_ATTR(128, 128, 0);
    SDN_SMEM_ATTR(128, 256, 0); SDN_SMEM_ATTR(64, 32, 0); SDN_SMEM_ATTR(64, 64, 0);
    SDN_SMEM_ATTR(128, 32, 1); SDN_SMEM_ATTR(128, 64, 1); SDN_SMEM_ATTR(128, 128, 1);
    SDN_SMEM_ATTR(64, 32, 1); SDN_SMEM_ATTR(64, 64, 1);
    SDN_SMEM_ATTR(128, 32, 2); SDN_SMEM_ATTR(128, 64, 2); SDN_SMEM_ATTR(64, 32, 2); SDN_SMEM_ATTR(64, 64, 2);
    SDN_SMEM_ATTR(128, 64, 3);
    SDN_SMEM_ATTR(128, 128, 0, 64);
    SDN_SMEM_ATTR(128, 256, 0, 0, 2); SDN_SMEM_ATTR(128, 128, 0, 0, 2); SDN_SMEM_ATTR(128, 128, 1, 0, 2);
    SDN_SMEM_ATTR(128, 128, 0, 0, 1, 2); SDN_SMEM_ATTR(128, 128, 1, 0, 1, 2); SDN_SMEM_ATTR(128, 128, 1, 0, 2, 2);
    SDN_SMEM_ATTR(64, 64, 2, 0, 2); SDN_SMEM_ATTR(128, 64, 2, 0, 2);
#undef SDN_SMEM_ATTR
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(CgCfg<128, 256, 0, 2>::THREADS); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<128, 256, 0, 0, 2>, &cfg) == cudaSuccess && n > 0) g_max_pairs = n;
        else (void)cudaGetLastError();
    }
    CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute(wgrad_gemm_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_gemm_kernel<128, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_gemm_kernel<64, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<32, 32, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<32, 32, 1, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<32, 32, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<32, 64, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<64, 64, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute((wgrad_tr_kernel<64, 32, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CUDA_OK(cudaFuncSetAttribute(decode_resize_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_OK(cudaFuncSetAttribute(decode_resize_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return 0;
}

static int cg_smem_halo(int swa, int bn, int stages, int a_stage_bytes) {
    if (swa == 128) {
        if (bn == 32) return CgCfg<128, 32>::smem_bytes_halo(stages, a_stage_bytes);
        if (bn == 64) return CgCfg<128, 64>::smem_bytes_halo(stages, a_stage_bytes);
        return CgCfg<128, 128>::smem_bytes_halo(stages, a_stage_bytes);
    }
    if (bn == 32) return CgCfg<64, 32>::smem_bytes_halo(stages, a_stage_bytes);
    return CgCfg<64, 64>::smem_bytes_halo(stages, a_stage_bytes);
}
static int cg_smem(int swa, int bn, int stages) {
    if (swa == 128) {
        if (bn == 32) return CgCfg<128, 32>::smem_bytes(stages);
        if (bn == 64) return CgCfg<128, 64>::smem_bytes(stages);
        if (bn == 128) return CgCfg<128, 128>::smem_bytes(stages);
        return CgCfg<128, 256>::smem_bytes(stages);
    }
    if (bn == 32) return CgCfg<64, 32>::smem_bytes(stages);
    return CgCfg<64, 64>::smem_bytes(stages);
}

// Source description for one GEMM A-segment group: an activation tensor (or a
// strided view of one) restricted to channels [c_off, c_off + C).
struct SrcView {
    const bf16* base;   // already offset to the view's first element
    int C;              // channels in the view
    int W, H;           // logical extent of the view
    long long sW, sH, sN;  // element strides
};
static SrcView full_view(const Act& a, int c_off = 0, int C = -1) {
    SrcView v;
    v.base = a.p + c_off;
    v.C = C < 0 ? a.C : C;
    v.W = a.W; v.H = a.H;
    v.sW = a.C; v.sH = (long long)a.W * a.C; v.sN = (long long)a.H * a.W * a.C;
    return v;
}
// quadrant (i, j) of a 2x-upsampled tensor u[B, 2h, 2w, C], seen at the low resolution
static SrcView quad_view(const Act& u, int q) {
    const int i = q >> 1, j = q & 1;
    SrcView v;
    v.base = u.p + ((long long)i * u.W + j) * u.C;
    v.C = u.C;
    v.W = u.W / 2; v.H = u.H / 2;
    v.sW = 2LL * u.C; v.sH = 2LL * u.W * u.C; v.sN = (long long)u.H * u.W * u.C;
    return v;
}

// Row-parity i of a 2x-upsampled tensor u[B, 2h, 2w, C], seen at the low resolution with the two horizontally adjacent
// output pixels (quadrants (i,0) and (i,1)) as ONE pixel of 2C contiguous channels: the same bytes as quad_view(u, 2i)
// and quad_view(u, 2i+1), but in boxes with 2C-channel rows (half the TMA boxes, twice the contiguous run).
static SrcView pair_view(const Act& u, int i) {
    SrcView v;
    v.base = u.p + (long long)i * u.W * u.C;
    v.C = 2 * u.C;
    v.W = u.W / 2; v.H = u.H / 2;
    v.sW = 2LL * u.C; v.sH = 2LL * u.W * u.C; v.sN = (long long)u.H * u.W * u.C;
    return v;
}

// Fill a GemmOp.  segs: list of (view index, dx, dy); every view contributes all
// of its channels per segment.  dviews: destination views, each n_per_dmap wide.
struct SegSpec {
    int view, dx, dy;
};
// SDN_FIRST_ROWS (default 1): the first conv runs as a row-halo 3x1 conv over a 32-channel tensor holding the
// three horizontal taps (im2col_rows_kernel) instead of a K = 64 GEMM over a full 64-channel im2col
static int first_rows() {
    static const int v = env_int("SDN_FIRST_ROWS", 1);
    return v;
}

// BatchNorm-backward statistics fused into a data-gradient epilogue (CG_BSTATS): the layer whose dA the op writes
struct BStatSpec {
    const Act* y;
    const float *scale, *shift, *mean;
};

// The same tensor seen as super-pixels: [B,H,W,32] == [B,H,W/2,64]
static SrcView super_view(const SrcView& v) {
    SrcView s = v;
    s.C = 64; s.W = v.W / 2; s.sW = 2 * v.sW;
    return s;
}

// conv_gemm HALO = 3 (see conv_gemm.cuh): 3x3 conv, every source 32 channels, 32 output channels, one destination.
// The kernel runs <128, 64, 3> on the super-pixel views; bmat holds the mode-11 / mode-12 packing.
static int build_gemm_superpixel(sdn_ctx* c, GemmOp& op, int B, const std::vector<SrcView>& aviews, const bf16* bmat,
                                 const SrcView& dview, const float* bias, int flags, float* stats_partials,
                                 const BStatSpec* bs) {
    memset(&op.p, 0, sizeof op.p);
    ConvGemmParams& p = op.p;
    op.swa = 128; op.block_n = 64; op.halo = 3; op.swd64 = false; op.eg = 2; op.ncta = 1;
    const SrcView d = super_view(dview);
    const int W = d.W, H = d.H, units = (int)aviews.size();
    p.TW = 8; p.TH = 16; p.TN = 1;
    p.tiles_x = (W + 7) / 8; p.tiles_y = (H + 15) / 16; p.tiles_n = B;
    p.img_w = W; p.img_h = H; p.img_n = B;
    for (int i = 0; i < units; ++i) {
        const SrcView v = super_view(aviews[i]);
        SDN_OK(encode4(&p.a_maps[i], v.base, 64, v.W, v.H, B, v.sW, v.sH, v.sN, 64, 10, 18, 1, 128));
        p.segs[i].map = (int8_t)i; p.segs[i].dx = 0; p.segs[i].dy = 0; p.segs[i].c0 = 0; p.segs[i].cblocks = 1;
    }
    for (int i = units; i < 4; ++i) p.a_maps[i] = p.a_maps[0];
    p.nsegs = units;
    p.kblocks_total = units;
    p.a_stage_bytes = (10 * 18 * 128 + 1023) & ~1023;
    SDN_OK(encode2(&p.b_map, bmat, 64, units * 3 * 128, 64, 128, 128));
    SDN_OK(encode4(&p.d_maps[0], d.base, 64, d.W, d.H, B, d.sW, d.sH, d.sN, 64, 8, 16, 1, 128));
    for (int i = 1; i < 4; ++i) p.d_maps[i] = p.d_maps[0];
    int ybytes = 0;
    p.y_map = p.d_maps[0];
    p.ybuf = 1;
    if (bs != nullptr) {
        const SrcView y = super_view(full_view(*bs->y));
        SDN_OK(encode4(&p.y_map, y.base, 64, y.W, y.H, B, y.sW, y.sH, y.sN, 64, 8, 16, 1, 128));
        flags |= CG_BSTATS;
        stats_partials = c->stats_partials;
        p.bs_scale = bs->scale; p.bs_shift = bs->shift; p.bs_mean = bs->mean;   // (duplicated at [32, 64) by the finalize)
        p.ybuf = 2;
        ybytes = 2 * p.ybuf * 128 * 64 * 2;
    }
    p.n_tiles = 1; p.n_per_dmap = 64; p.n_total = 64;
    p.flags = flags | CG_BRES;
    p.bias = bias;
    p.stats_partials = stats_partials;
    p.b_res_bytes = units * 3 * 128 * 128;
    int fixed = cg_smem_halo(128, 64, 0, p.a_stage_bytes) + ybytes;
    if (bs != nullptr && (220 * 1024 - fixed - p.b_res_bytes) / p.a_stage_bytes < 3) {
        p.ybuf = 1; fixed -= ybytes / 2; ybytes /= 2;
    }
    const int budget = 220 * 1024 - fixed - p.b_res_bytes;
    p.ups = (units <= 2 && budget / (units * p.a_stage_bytes) >= 3) ? units : 1;
    p.stages = std::max(2, std::min(8, budget / (p.ups * p.a_stage_bytes)));
    op.smem = fixed + p.b_res_bytes + p.stages * p.ups * p.a_stage_bytes;
    if (op.smem > 227 * 1024) return fail("build_gemm_superpixel: shared memory %d", op.smem);
    op.grid = std::max(1, std::min(p.tiles_x * p.tiles_y * p.tiles_n, c->num_sms));
    return 0;
}
static int build_gemm(sdn_ctx* c, GemmOp& op, int B, const std::vector<SrcView>& aviews,
                      const std::vector<SegSpec>& segs_in, const bf16* bmat, int n_total,
                      const std::vector<SrcView>& dviews, int n_per_dmap, const float* bias, int flags,
                      float* stats_partials, bool conv3x3 = false, int dx_taps = 3, const BStatSpec* bs = nullptr) {
    static const int g_halo_max_n = env_int("SDN_HALO_MAXN", 128);   // largest BLOCK_N that uses the row-halo kernel (0 disables)
    std::vector<SegSpec> segs = segs_in;
    if (aviews.empty() || aviews.size() > 4 || dviews.empty() || dviews.size() > 4 || segs.size() > CG_MAX_SEGS)
        return fail("build_gemm: bad view/segment counts");
    memset(&op.p, 0, sizeof op.p);
    {
        static const int sp_on = env_int("SDN_SUPERPIX", 1), box9_ok = env_int("SDN_BOX9", 1);
        bool sp = sp_on && box9_ok && conv3x3 && dx_taps == 3 && n_total == 32 && n_per_dmap == 32 && dviews.size() == 1 &&
                  dviews[0].C == 32 && dviews[0].W % 16 == 0 && aviews.size() <= 2 && (bs == nullptr || bs->y->C == 32);
        for (const SrcView& v : aviews) sp = sp && v.C == 32 && v.W == dviews[0].W && v.sW == 32;
        if (sp) return build_gemm_superpixel(c, op, B, aviews, bmat, dviews[0], bias, flags, stats_partials, bs);
    }
    bool all64 = true;
    for (const SrcView& v : aviews) {
        if (v.C % 32 != 0) return fail("build_gemm: source channels %d not a multiple of 32", v.C);
        if (v.C % 64 != 0) all64 = false;
    }
    op.swa = all64 ? 128 : 64;
    int bn = std::min(n_per_dmap, 256);
    // narrow destinations (dgrad of a concat input): one tile spans both, so the A operand is fetched once
    op.swd64 = false;
    if (n_per_dmap < 128 && n_total > n_per_dmap) {
        int wide = std::min(n_total, op.swa == 64 ? 64 : 128);
        const int wide_dch = (wide >= 64 && op.swa == 128) ? 64 : 32;   // store block of that tile width
        if (wide % n_per_dmap == 0 && n_per_dmap % wide_dch == 0) bn = wide;
        else if (!conv3x3 && op.swa == 128 && wide == 128 && n_per_dmap == 32 && (flags & CG_STATS) == 0) {
            bn = 128;            // four 32-channel quadrants from one tile: the A operand is read once, not 4x
            op.swd64 = true;
        }
    }
    if (op.swa == 64 && bn > 64) bn = 64;
    // short K loops (ConvTranspose2d GEMMs: 1-4 k-blocks per tile) are epilogue-bound: N = 128 tiles with two
    // epilogue groups and two staging buffers instead of one N = 256 / N = 128 tile drained by one group
    int kb_est = 0;
    for (const SegSpec& sg : segs_in) kb_est += aviews[sg.view].C / 64;
    static const int eg2_on = env_int("SDN_EG2", 1);
    const bool short_k = eg2_on && !conv3x3 && op.swa == 128 && kb_est <= 4 && !op.swd64 && n_per_dmap % 64 == 0;
    if (short_k && bn == 256) bn = 128;
    const int W = dviews[0].W, H = dviews[0].H;
    if (bn == 256 && n_per_dmap % 128 == 0) {
        // Wave quantisation on the small levels: a persistent grid of num_sms CTAs runs ceil(jobs / num_sms)
        // rounds.  Halving the N tile doubles the jobs at half the cost each (N = 128 MMAs keep the pipe as busy
        // as N = 256 ones: 67 vs 131 cycles per K = 16 step) and can save most of a nearly-empty last round:
        // 32 pairs at 15x20: 150 jobs = 2 rounds -> 300 half-jobs = 1.5; at 30x40: 3 -> 2.5.  The A tile is then
        // fetched once per N tile (L2-resident at these sizes), hence the 5 % bar.
        static const int nsplit_on = env_int("SDN_NSPLIT", 1);
        const long long m_jobs = ((long long)W * H * B + 127) / 128;
        const long long r256 = (m_jobs * (n_total / 256) + c->num_sms - 1) / c->num_sms * 2;
        const long long r128 = (m_jobs * (n_total / 128) + c->num_sms - 1) / c->num_sms;
        if (nsplit_on && r128 * 1.03 < 0.95 * r256) bn = 128;
    }
    static const int bstats_on = env_int("SDN_BSTATS", 1);
    if (bs != nullptr && !(bstats_on && dviews.size() == 1 && n_per_dmap == n_total && bn == n_total && bn <= 128 && !op.swd64))
        bs = nullptr;   // wide / split tiles keep the separate reduction pass (levels 4-5: few bytes)
    if (!(flags & CG_STATS) && bs == nullptr) {
        // latency regime (few pixels): narrower N tiles spread the K loop over more CTAs
        const long long m_est = ((long long)W * H * B + 127) / 128;
        while (bn > 32 && m_est * (n_total / bn) < c->num_sms / 2) {
            if (op.swd64) { op.swd64 = false; bn = std::min(n_per_dmap, 256); }   // back to one tile per quadrant
            else bn /= 2;
        }
    }
    if (n_per_dmap % bn != 0 && bn % n_per_dmap != 0)
        return fail("build_gemm: N %d and BLOCK_N %d do not nest", n_per_dmap, bn);
    op.block_n = bn;
    op.halo = (conv3x3 && bn <= g_halo_max_n) ? 1 : 0;
    Tile t = choose_tile(W, H, B, 128);
    static const int box9_on = env_int("SDN_BOX9", 1), bres_max = env_int("SDN_BRES_MAXKB", 80) * 1024;
    int smem_budget = 220 * 1024;
    bool halfk = false;
    {
        // box9: one (TH+2)x(TW+2) box per (source, channel block) feeds all nine taps; needs the whole
        // packed weight matrix resident in shared memory and 8-pixel-wide tiles
        int cin_tot = 0;
        for (const SrcView& v : aviews) cin_tot += v.C;
        // (two pipeline stages of one 18x10 box must still fit next to the weights and the store staging)
        const int box9_stage = (10 * 18 * op.swa + 1023) & ~1023;
        const int box9_need = (bn <= 64 ? cg_smem_halo(op.swa, bn, 0, box9_stage) : 1 << 30) + 9 * cin_tot * bn * 2 + 2 * box9_stage;
        const bool box9_shape = op.halo && dx_taps == 3 && box9_on && bn == n_total && bn <= 64 && W % 8 == 0;
        if (box9_shape && 9 * cin_tot * bn * 2 <= bres_max && box9_need <= 227 * 1024) {
            op.halo = 2;
        } else if (box9_shape && op.swa == 128 && n_per_dmap == n_total && bs == nullptr) {
            // 128 input channels -> 64 (dec2.0 forward, enc3.0 data gradient): the 147 KB weight matrix does not fit
            // next to 128-byte-row stages, and the row-halo form re-streams it from L2 for every tile (252 KB of
            // L2->SM traffic per 128 pixels: these two layers ran at 46 % of the pipe where their same-FLOP siblings
            // reach 66-70 %).  With HALF k-blocks (32 channels, 64-byte swizzle rows: 11.5 KB per 18x10 box) three
            // pipeline stages fit next to the resident weights: 46 KB per tile.
            static const int half_on = env_int("SDN_BOX9_HALFK", 1), half_max = env_int("SDN_BOX9_HALFK_MAXKB", 150) * 1024;
            const int stage64 = (10 * 18 * 64 + 1023) & ~1023;
            const int need64 = cg_smem_halo(64, bn, 0, stage64) + 9 * cin_tot * bn * 2 + 3 * stage64;
            // a CTA pair keeps half of the weight rows per CTA: full 64-channel k-blocks fit again (two units of 36 MMAs
            // per tile instead of four of 18, 128-byte TMA rows): dec2.0 forward 0.69 -> 0.62 ms over the half-k-block
            // pair form; SDN_BOX9_PAIRS=2 keeps the half k-blocks
            static const int pairfull = env_int("SDN_BOX9_PAIRS", 1) == 1 && env_int("SDN_CTA2", 1);
            if (pairfull && bn == 64 && (B % 2 == 0 || B >= 16) && 9 * cin_tot * bn <= bres_max &&
                cg_smem_halo(128, bn, 0, box9_stage) + 9 * cin_tot * bn + 3 * box9_stage <= 220 * 1024) {
                op.halo = 2;
                halfk = true;      // (selects the pair form below; the k-blocks stay whole)
            } else if (half_on && 9 * cin_tot * bn * 2 <= half_max && need64 <= 226 * 1024) {
                op.swa = 64;
                op.halo = 2;
                smem_budget = 226 * 1024;
                halfk = true;
            }
        }
    }
    const int KB = op.swa / 2;
    if (op.halo == 2) {
        t = Tile{8, 16, 1};
        segs.clear();
        for (int sv = 0; sv < (int)aviews.size(); ++sv) segs.push_back({sv, 0, 0});
    } else if (op.halo) {
        // one image per box and TW % 8 == 0 so the vertical-tap row shifts are whole swizzle groups
        double best = 1e300;
        static const int force_tw = env_int("SDN_HALO_TW", 0);
        for (int TW = 8; TW <= 32; TW *= 2) {
            if (force_tw && TW != force_tw) continue;
            const int TH = 128 / TW;
            const double padded = double((W + TW - 1) / TW * TW) * double((H + TH - 1) / TH * TH);
            const double cost = padded * double(TH + 2) / double(TH);
            if (cost < best) { best = cost; t = Tile{TW, TH, 1}; }
        }
        // units: (horizontal tap, source); each contributes all of its channel blocks, 3 vertical taps each
        segs.clear();
        for (int dx = (dx_taps == 3 ? -1 : 0); dx <= (dx_taps == 3 ? 1 : 0); ++dx)
            for (int sv = 0; sv < (int)aviews.size(); ++sv) segs.push_back({sv, dx, 0});
    }
    ConvGemmParams& p = op.p;
    p.TW = t.TW; p.TH = t.TH; p.TN = t.TN;
    p.tiles_x = (W + t.TW - 1) / t.TW;
    p.tiles_y = (H + t.TH - 1) / t.TH;
    p.tiles_n = (B + t.TN - 1) / t.TN;
    // CTA pairs (conv_gemm NCTA = 2): wide tiles, the two CTAs take consecutive image groups of one tile position
    static const int cta2_on = env_int("SDN_CTA2", 1);
    op.ncta = 1;
    // (short K loops lose to the pair's extra handshakes - measured: 64 -> 128 3x3 at N = 128 and the level-1/2
    // ConvTranspose2d GEMMs got 10-25 % slower, everything with K * N >= 512 * 256 got 20-35 % faster)
    int k_total = 0;
    for (const SegSpec& sg : segs) k_total += aviews[sg.view].C * (op.halo ? 3 : 1);
    static const int cta2_min_kn = env_int("SDN_CTA2_MIN_KN", 512 * 256);
    // Row-halo N = 128 tiles with K = 576 (enc3.0 forward, the 128-wide data gradient of dec2.0): one epilogue group
    // needs longer for a 128 x 128 tile (TMEM loads, conversion, statistics, store) than the 36 MMAs of the main loop
    // take, so these two layers were EPILOGUE-bound and the pair form alone made them 10-15 % slower; with TWO epilogue
    // groups on alternate tiles AND the pair form (weights resident: 74 KB per CTA) they gain 15-20 %.  Measured
    // beside it: two groups without the pair form change nothing, a second staging buffer alone changes nothing, and
    // at K = 1152 two groups cost 7 % (fewer pipeline stages).  Not with CG_BSTATS: two groups' y tiles plus two
    // staging buffers leave no room for the pipeline stages.
    static const int eg2h_on = env_int("SDN_EG2_HALO", 1);
    const bool eg2_halo = eg2h_on && cta2_on && op.halo == 1 && bn == 128 && op.swa == 128 && k_total <= 576 && !op.swd64 &&
                          bs == nullptr && (p.tiles_n % 2 == 0 || p.tiles_n >= 16);
    // box9 kernels in the pair form: measured on every box9 layer - the two half-k-block layers (128 -> 64: dec2.0
    // forward, enc3.0 data gradient) gain 23-24 % (each CTA keeps half of the 147 KB weights, which leaves room for
    // nine pipeline stages instead of three), every other box9 layer is unchanged (64 -> 64 forward) or 40-50 % SLOWER
    // (the CG_BSTATS data gradients, the 32-channel-source kernels): pairs only where the weights do not fit otherwise
    static const int box9_pair = env_int("SDN_BOX9_PAIRS", 1);
    const bool pair_box9 = box9_pair && cta2_on && op.halo == 2 && halfk && bn == 64 && (p.tiles_n % 2 == 0 || p.tiles_n >= 16);
    if (pair_box9 || (cta2_on && bn >= 128 && op.halo <= 1 && !op.swd64 && op.swa == 128 && (p.tiles_n % 2 == 0 || p.tiles_n >= 16) &&
        ((long long)k_total * bn >= cta2_min_kn || eg2_halo))) {
        op.ncta = 2;
        p.tiles_n = (p.tiles_n + 1) / 2;   // the kernel walks pairs of image groups
    }
    const int bn_loc = bn / op.ncta;       // weight rows this CTA stages
    op.eg = (bn <= 64 || (short_k && bn == 128 && op.halo == 0 && op.ncta == 1) || eg2_halo) ? 2 : 1;
    p.img_w = W; p.img_h = H; p.img_n = B;
    for (size_t i = 0; i < aviews.size(); ++i) {
        const SrcView& v = aviews[i];
        SDN_OK(encode4(&p.a_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, KB, op.halo == 2 ? t.TW + 2 : t.TW,
                       op.halo ? t.TH + 2 : t.TH, t.TN, op.swa));
    }
    p.a_stage_bytes = ((op.halo == 2 ? t.TW + 2 : t.TW) * (t.TH + 2) * op.swa + 1023) & ~1023;
    for (size_t i = aviews.size(); i < 4; ++i) p.a_maps[i] = p.a_maps[0];
    int kblocks = 0;
    p.nsegs = (int)segs.size();
    for (size_t s = 0; s < segs.size(); ++s) {
        CgSeg& g = p.segs[s];
        g.map = (int8_t)segs[s].view;
        g.dx = (int8_t)segs[s].dx;
        g.dy = (int8_t)segs[s].dy;
        g.c0 = 0;
        g.cblocks = (int16_t)(aviews[segs[s].view].C / KB);
        kblocks += g.cblocks;
    }
    p.kblocks_total = kblocks;
    if (op.halo) SDN_OK(encode3(&p.b_map, bmat, KB, n_total, kblocks * (op.halo == 2 ? 9 : 3), bn_loc, op.swa));
    else SDN_OK(encode2(&p.b_map, bmat, kblocks * KB, n_total, KB, bn_loc, op.swa));
    const int swd = op.swd64 ? 64 : ((bn >= 64 && op.swa == 128) ? 128 : 64);
    const int dch = swd / 2;
    if (n_per_dmap % dch != 0) return fail("build_gemm: destination width %d not a multiple of the %d-channel store block", n_per_dmap, dch);
    for (size_t i = 0; i < dviews.size(); ++i) {
        const SrcView& v = dviews[i];
        SDN_OK(encode4(&p.d_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, dch, t.TW, t.TH, t.TN, swd));
    }
    for (size_t i = dviews.size(); i < 4; ++i) p.d_maps[i] = p.d_maps[0];
    int ybytes = 0;
    if (bs != nullptr) {
        const SrcView v = full_view(*bs->y);
        if (v.C != n_total || v.W != W || v.H != H) return fail("build_gemm: BatchNorm-backward statistics need y of the destination's shape");
        SDN_OK(encode4(&p.y_map, v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, dch, t.TW, t.TH, t.TN, swd));
        flags |= CG_BSTATS;
        stats_partials = c->stats_partials;
        p.bs_scale = bs->scale; p.bs_shift = bs->shift; p.bs_mean = bs->mean;
        // small tiles are epilogue-bound: fetch y two tiles ahead; wide tiles hide one fetch behind their main loop
        p.ybuf = bn <= 64 ? 2 : 1;
        ybytes = op.eg * p.ybuf * 128 * bn * 2;
    } else {
        p.y_map = p.d_maps[0];
        p.ybuf = 1;
    }
    p.n_tiles = n_total / bn;
    p.n_per_dmap = n_per_dmap;
    p.n_total = n_total;
    p.flags = flags;
    p.bias = bias;
    p.stats_partials = stats_partials;
    if ((flags & CG_STATS) && (n_total > 512 || (p.n_tiles > 1 && bn < 128))) return fail("build_gemm: stats need n_total <= 512 and N tiles of >= 128 channels");
    int stages = 8;
    if (op.halo == 2) {
        const int b_total = kblocks * 9 * bn_loc * op.swa;
        int fixed = cg_smem_halo(op.swa, bn, 0, p.a_stage_bytes) + ybytes;
        if (bs != nullptr && p.ybuf == 2 && (smem_budget - fixed - b_total) / p.a_stage_bytes < 3) {
            p.ybuf = 1;            // big resident weights (64 -> 64): keep three pipeline stages instead
            fixed -= ybytes / 2;
            ybytes /= 2;
        }
        p.flags |= CG_BRES;
        p.b_res_bytes = b_total;
        const int budget = smem_budget - fixed - b_total;
        // every unit of a tile in one stage when at least three such stages fit
        p.ups = (kblocks <= 3 && budget / (kblocks * p.a_stage_bytes) >= 3) ? kblocks : 1;
        stages = std::max(2, std::min(8, budget / (p.ups * p.a_stage_bytes)));
        op.smem = fixed + b_total + stages * p.ups * p.a_stage_bytes;
        if (op.smem > 227 * 1024) return fail("build_gemm: box9 shared memory %d", op.smem);
    } else if (op.halo) {
        // weights resident in shared memory when the whole packed matrix of this N tile fits;
        // three units per pipeline stage when >= 4 such stages still fit (fewer handshakes per tile)
        const int b_total = kblocks * 3 * bn_loc * op.swa;
        static const int ups_on = env_int("SDN_UPS", 3);
        const int fixed = (op.eg == 2 && bn == 128 ? CgCfg<128, 128, 0, 1, 2>::smem_bytes_halo(0, p.a_stage_bytes)
                                                   : cg_smem_halo(op.swa, bn, 0, p.a_stage_bytes)) + ybytes;   // staging, scratch, barriers
        const bool res = p.n_tiles == 1 && b_total <= bres_max;
        if (res) { p.flags |= CG_BRES; p.b_res_bytes = b_total; }
        const int unit_bytes = p.a_stage_bytes + (res ? 0 : 3 * bn_loc * op.swa);
        const int budget = 220 * 1024 - fixed - (res ? b_total : 0);
        p.ups = (ups_on == 3 && kblocks % 3 == 0 && budget / (3 * unit_bytes) >= 4) ? 3 : 1;
        stages = std::max(2, std::min(8, budget / (p.ups * unit_bytes)));
        op.smem = fixed + (res ? b_total : 0) + stages * p.ups * unit_bytes;
        if (op.smem > 227 * 1024) return fail("build_gemm: row-halo shared memory %d", op.smem);
    } else {
        if (op.swd64) {
            while (stages > 2 && CgCfg<128, 128, 64>::smem_bytes(stages) > 220 * 1024) --stages;
            op.smem = CgCfg<128, 128, 64>::smem_bytes(stages);
        } else {
            const int stage_bytes = 128 * op.swa + bn_loc * op.swa;
            const int fixed = (op.eg == 2 && bn == 128) ? CgCfg<128, 128, 0, 1, 2>::smem_bytes(0) : cg_smem(op.swa, bn, 0);
            while (stages > 2 && fixed + stages * stage_bytes + ybytes > 220 * 1024) --stages;
            op.smem = fixed + stages * stage_bytes + ybytes;
        }
    }
    p.stages = stages;
    const int num_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;
    op.grid = op.ncta == 2 ? 2 * std::max(1, std::min(num_tiles, g_max_pairs > 0 ? g_max_pairs : c->num_sms / 2))
                           : std::max(1, std::min(num_tiles, c->num_sms));
    return 0;
}

// Weight-gradient op.  avariants: dY views (1, or the 4 quadrants for convT);
// bsrc: 1 or 2 X sources; taps 9 (row-halo units, see wgrad_gemm.cuh) or 1.
static int build_wgrad(sdn_ctx* c, WgradOp& op, int B, const std::vector<SrcView>& avariants, int cout,
                       const std::vector<SrcView>& bsrc, int taps, float* out, int k_rows_valid, int ndx = 3) {
    memset(&op.p, 0, sizeof op.p);
    WgradParams& p = op.p;
    bool all64 = true;
    int cin_tot = 0;
    for (const SrcView& v : bsrc) {
        if (v.C % 32 != 0) return fail("build_wgrad: source channels %d not a multiple of 32", v.C);
        if (v.C % 64 != 0) all64 = false;
        cin_tot += v.C;
    }
    op.swb = all64 ? 128 : 64;
    const int CA = op.swb / 2;
    const int W = avariants[0].W, H = avariants[0].H;
    p.halo = taps == 9 ? 1 : 0;
    op.tr2_natoms = 0;
    static const int tr2_on = env_int("SDN_WGRAD_TR2", 1);
    if (tr2_on && p.halo && (cout == 32 || cout == 64) && W % 8 == 0 && avariants.size() == 1 && avariants[0].C == cout) {
        // levels 1-2 (small Cout, many pixels): swapped roles, one halo box per channel atom (wgrad_tr.cuh)
        const int atoms = cin_tot / CA;
        const int natoms = CA == 32 ? std::min(atoms, 2) : 1;
        const SrcView& y = avariants[0];
        SDN_OK(encode4(&p.a_maps[0], y.base, y.C, y.W, y.H, B, y.sW, y.sH, y.sN, cout, 8, 16, 1, cout * 2));
        for (int i = 1; i < 4; ++i) p.a_maps[i] = p.a_maps[0];
        for (size_t i = 0; i < bsrc.size(); ++i) {
            const SrcView& v = bsrc[i];
            SDN_OK(encode4(&p.b_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, CA, 10, 18, 1, op.swb));
        }
        if (bsrc.size() == 1) p.b_maps[1] = p.b_maps[0];
        p.TW = 8; p.TH = 16; p.TN = 1; p.kpix = 128;
        p.tiles_x = W / 8; p.tiles_y = (H + 15) / 16; p.tiles_n = B;
        p.atoms_per_tap = atoms;
        p.atoms_src0 = bsrc[0].C / CA;
        p.unit_groups = (atoms + natoms - 1) / natoms;
        p.cout = cout; p.cin_tot = cin_tot; p.k_rows_valid = k_rows_valid; p.out = out; p.tr = 2;
        const int y_bytes = 128 * cout * 2;
        const int x_bytes = (10 * 18 * op.swb + 1023) & ~1023;
        const int stage_bytes = y_bytes + natoms * x_bytes;
        const int fixed = 1024 + 4 * 32 * 33 * 4 + 256;
        p.stages = std::max(2, std::min(8, (220 * 1024 - fixed) / stage_bytes));
        op.smem = fixed + p.stages * stage_bytes;
        op.tr2_natoms = natoms;
        op.tr2_ndx = ndx;
        const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;
        static const int waves2 = env_int("SDN_WGRAD_WAVES", 1);
        int split = std::max(1, std::min((waves2 * c->num_sms) / p.unit_groups, ptiles));
        op.grid = dim3(split, p.unit_groups, 1);
        return 0;
    }
    Tile t;
    int plain_kpix = 64;
    if (p.halo) {
        // one image per box, TW a multiple of 8 (row shifts must be whole swizzle groups)
        double best = 1e300;
        t = Tile{8, 8, 1};
        for (int TW = 8; TW <= 64; TW *= 2) {
            const int TH = 64 / TW;
            const double padded = double((W + TW - 1) / TW * TW) * double((H + TH - 1) / TH * TH);
            const double cost = padded * double(TH + 2) / double(TH);
            if (cost < best) { best = cost; t = Tile{TW, TH, 1}; }
        }
    } else {
        // plain (1x1 / ConvTranspose2d / im2col'ed first layer) tiles: 128 pixels halve the per-tile handshakes
        // (only where there are plenty of tiles: the small levels need the parallelism of 64-pixel tiles more)
        plain_kpix = (long long)W * H * B / 128 >= 32LL * c->num_sms ? 128 : 64;
        t = choose_tile(W, H, B, plain_kpix);
    }
    p.TW = t.TW; p.TH = t.TH; p.TN = t.TN;
    p.kpix = p.halo ? 64 : plain_kpix;
    p.tiles_x = (W + t.TW - 1) / t.TW;
    p.tiles_y = (H + t.TH - 1) / t.TH;
    p.tiles_n = (B + t.TN - 1) / t.TN;
    p.a_variants = (int)avariants.size();
    p.a_atoms = cout >= 128 ? 2 : 1;
    p.m_tiles = (cout + 127) / 128;
    static const int tr_on = env_int("SDN_WGRAD_TR", 1);
    // measured: the swapped roles win when a unit is 64 channels wide or when there are >= 6 units
    // (two sources); 3 units of 32 channels are issue-bound either way and stay in the plain layout
    p.tr = (tr_on && p.halo && (cout == 32 || cout == 64) && (tr_on == 2 || CA == 64 || 3 * (cin_tot / CA) >= 6)) ? 1 : 0;
    for (size_t i = 0; i < avariants.size(); ++i) {
        const SrcView& v = avariants[i];
        if (p.tr) SDN_OK(encode4(&p.a_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, cout, t.TW, t.TH, t.TN, cout * 2));
        else SDN_OK(encode4(&p.a_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, 64, t.TW, t.TH, t.TN, 128));
    }
    for (size_t i = avariants.size(); i < 4; ++i) p.a_maps[i] = p.a_maps[0];
    const int b_rows = p.halo ? (t.TH + 2) * t.TW : p.kpix;
    for (size_t i = 0; i < bsrc.size(); ++i) {
        const SrcView& v = bsrc[i];
        SDN_OK(encode4(&p.b_maps[i], v.base, v.C, v.W, v.H, B, v.sW, v.sH, v.sN, CA, t.TW, p.halo ? t.TH + 2 : t.TH,
                       t.TN, op.swb));
    }
    if (bsrc.size() == 1) p.b_maps[1] = p.b_maps[0];
    p.atoms_per_tap = cin_tot / CA;
    p.atoms_src0 = bsrc[0].C / CA;
    const int ndy = p.halo ? 3 : 1;
    p.total_units = (p.halo ? 3 : 1) * p.atoms_per_tap;
    const int a_bytes = p.tr ? p.kpix * cout * 2 : p.a_atoms * p.kpix * 128;
    const int b_tile_bytes = (b_rows * op.swb + 1023) & ~1023;
    const int cols_per_unit = p.tr ? cout * (CA == 64 ? 2 : 1) : ndy * CA;
    int umax = std::min(p.total_units, (p.halo ? 512 : 256) / cols_per_unit);   // plain units share one N <= 256 MMA
    umax = std::min(umax, 8);
    while (umax > 1 && 3 * (a_bytes + umax * b_tile_bytes) + 2048 > 200 * 1024) --umax;  // keep >= 3 stages
    p.unit_groups = (p.total_units + umax - 1) / umax;
    p.U = (p.total_units + p.unit_groups - 1) / p.unit_groups;
    p.unit_groups = (p.total_units + p.U - 1) / p.U;
    int cols = 32;
    while (cols < p.U * cols_per_unit) cols *= 2;
    p.tmem_cols = cols;
    p.cout = cout;
    p.cin_tot = cin_tot;
    p.k_rows_valid = k_rows_valid;
    p.out = out;
    const int stage_bytes = a_bytes + p.U * b_tile_bytes;
    int stages = 8;
    while (stages > 2 && stages * stage_bytes + 2048 > 200 * 1024) --stages;
    p.stages = stages;
    op.smem = stages * stage_bytes + 2048;
    const int ptiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int ctas_per_split = p.unit_groups * p.m_tiles * p.a_variants;
    // one wave measured best with the current kernels (256 pairs: +1.2 %, 32 pairs: +4.6 % over two waves:
    // half the split-K partial sums to merge); the older kernels preferred two
    static const int waves = env_int("SDN_WGRAD_WAVES", 1);
    // ConvTranspose2d: the four quadrant variants of a pixel tile read the SAME source tile.  With one wave
    // all four run side by side (blockIdx.z = variant, same blockIdx.x = same tiles), so the source comes from
    // DRAM once and from L2 three times; with two waves it is streamed from DRAM twice.
    static const int convt_waves = env_int("SDN_CONVT_WGRAD_WAVES", 1);
    const int w_eff = p.a_variants == 4 ? convt_waves : waves;
    int split = (w_eff * c->num_sms) / ctas_per_split;   // never spill into a partial extra wave (1 CTA per SM)
    split = std::max(1, std::min(split, ptiles));
    op.grid = dim3(split, p.unit_groups, p.m_tiles * p.a_variants);
    return 0;
}

// ------------------------------------------------------------------- plan
static const int kBlockCin[9] = {6, 32, 64, 128, 256, 512, 256, 128, 64};      // enc1..4, bott, dec4..1 (block.0 in)
static const int kBlockCout[9] = {32, 64, 128, 256, 512, 256, 128, 64, 32};
static const int kBlockLvl[9] = {1, 2, 3, 4, 5, 4, 3, 2, 1};
// first parameter index of each block in StereoUNet.parameters() order
static const int kBlockParam0[9] = {0, 6, 12, 18, 24, 32, 40, 48, 56};
static const int kUpParam0[4] = {30, 38, 46, 54};  // up4, up3, up2, up1

static void carve(uint8_t*& cur, size_t bytes, void** out) {
    *out = cur;
    cur += (bytes + 1023) & ~size_t(1023);
}

static int plan_and_alloc(sdn_ctx* c) {
    // Pass 1 computes sizes with cur = nullptr offsets, pass 2 assigns pointers.
    for (int pass = 0; pass < 2; ++pass) {
        uint8_t* cur = pass == 0 ? nullptr : c->ws;
        const int B = c->maxB;
        auto act = [&](Act& a, int C, int lvl) {
            a.C = C; a.H = lvl_h(c, lvl); a.W = lvl_w(c, lvl);
            carve(cur, a.elems(B) * sizeof(bf16), (void**)&a.p);
        };
        size_t wg_off = 0;
        if (!c->pre_only) act(c->x0, 64, 1);
        for (int b = 0; b < 9 && !c->pre_only; ++b) {
            for (int h = 0; h < 2; ++h) {
                ConvL& L = c->conv[2 * b + h];
                L.cin = h == 0 ? kBlockCin[b] : kBlockCout[b];
                L.cout = kBlockCout[b];
                L.lvl = kBlockLvl[b];
                L.first = (b == 0 && h == 0);
                L.pooled_out = (h == 1 && b < 4);
                L.p_w = kBlockParam0[b] + 3 * h;
                L.p_gamma = L.p_w + 1;
                L.p_beta = L.p_w + 2;
                L.bn = 2 * b + h;
                L.has_dgrad = !L.first;
                act(L.y, L.cout, L.lvl);
                act(L.a, L.cout, L.lvl);
                act(L.dy, L.cout, L.lvl);
                act(L.ga, L.cout, L.lvl);
                if (L.pooled_out) {
                    act(L.pool, L.cout, L.lvl + 1); act(L.gp, L.cout, L.lvl + 1);
                    carve(cur, L.pool.elems(B) / 8 * sizeof(unsigned short), (void**)&L.amax);
                }
                const int kdim = L.first ? 288 : 9 * L.cin;   // first layer: up to 9 x 32 workspace rows (row-halo form)
                // (32 -> 32 / 64 -> 32 layers may use the super-pixel packing: 3 x 128 x 64 per 32-channel source)
                const size_t wcap = std::max((size_t)L.cout * kdim, L.cout == 32 ? (size_t)((L.cin + 31) / 32) * 3 * 128 * 64 : 0);
                carve(cur, wcap * sizeof(bf16), (void**)&L.wf);
                carve(cur, wcap * sizeof(bf16), (void**)&L.wd);
                float* vecs = nullptr;
                carve(cur, 6 * 512 * sizeof(float), (void**)&vecs);
                L.scale = vecs; L.shift = vecs + 512; L.mean = vecs + 1024; L.rstd = vecs + 1536;
                L.c1 = vecs + 2048; L.c2 = vecs + 2560;
                wg_off += (size_t)kdim * L.cout;
            }
        }
        for (int k = 0; k < 4 && !c->pre_only; ++k) {
            UpL& U = c->up[k];
            U.cin = kBlockCout[4 + k];       // 512, 256, 128, 64
            U.cout = U.cin / 2;
            U.lvl_in = 5 - k;
            U.src_layer = 9 + 2 * k;         // bottleneck.3, dec4.3, dec3.3, dec2.3
            U.p_w = kUpParam0[k];
            U.p_b = U.p_w + 1;
            act(U.u, U.cout, U.lvl_in - 1);
            act(U.gu, U.cout, U.lvl_in - 1);
            carve(cur, (size_t)4 * U.cin * U.cout * sizeof(bf16), (void**)&U.wf);
            carve(cur, (size_t)4 * U.cin * U.cout * sizeof(bf16), (void**)&U.wd);
            carve(cur, (size_t)4 * U.cout * sizeof(float), (void**)&U.bias4);
            wg_off += (size_t)4 * U.cin * U.cout + U.cout;
        }
        // contiguous fp32 weight-gradient workspace (one memset per backward)
        carve(cur, wg_off * sizeof(float) + 1024, (void**)&c->wg_all);
        c->wg_all_bytes = wg_off * sizeof(float) + 1024;
        if (pass == 1 && !c->pre_only) {
            float* w = c->wg_all;
            for (int i = 0; i < 18; ++i) {
                ConvL& L = c->conv[i];
                L.wg = w;
                w += (size_t)(L.first ? 288 : 9 * L.cin) * L.cout;
            }
            for (int k = 0; k < 4; ++k) {
                UpL& U = c->up[k];
                U.wg = w; w += (size_t)4 * U.cin * U.cout;
                U.bg = w; w += U.cout;
            }
        }
        carve(cur, (size_t)c->num_sms * 2 * 512 * sizeof(float), (void**)&c->stats_partials);
        carve(cur, (size_t)BWD_BLOCKS * 2 * 512 * sizeof(float), (void**)&c->bwd_partials);
        carve(cur, 128 * sizeof(float), (void**)&c->head_grads);
        carve(cur, 3 * 16 * 8 * sizeof(long long), (void**)&c->dbg);
        carve(cur, 64, (void**)&c->n_local);
        const int parts = ((c->W + 127) / 128) * ((c->H + PRE_ROWS - 1) / PRE_ROWS);
        carve(cur, (size_t)2 * B * parts * sizeof(float), (void**)&c->gray_part);
        c->blur_tmp_elems = (size_t)2 * B * 3 * c->H * c->W;
        carve(cur, c->blur_tmp_elems * sizeof(float), (void**)&c->blur_tmp);
        carve(cur, (size_t)2 * B * sizeof(AugParams), (void**)&c->aug_stage);
        carve(cur, (size_t)2 * B * sizeof(float), (void**)&c->view_mean);
        if (pass == 0) {
            c->ws_bytes = (size_t)(cur - (uint8_t*)nullptr) + 4096;
            CUDA_OK(cudaMalloc((void**)&c->ws, c->ws_bytes));
            CUDA_OK(cudaMemset(c->ws, 0, c->ws_bytes));
        }
    }
    if (c->pre_only) return 0;
    // wiring of conv inputs
    for (int b = 0; b < 9; ++b) {
        ConvL& L0 = c->conv[2 * b];
        ConvL& L1 = c->conv[2 * b + 1];
        L1.nsrc = 1; L1.src[0] = &L0.a;
        if (b == 0) { L0.nsrc = 1; L0.src[0] = &c->x0; }
        else if (b <= 4) { L0.nsrc = 1; L0.src[0] = &c->conv[2 * b - 1].pool; }
        else {
            const int k = b - 5;                    // dec4..dec1 <-> up[k]
            const int skip = 7 - 2 * k;             // enc4.3, enc3.3, enc2.3, enc1.3
            L0.nsrc = 2; L0.src[0] = &c->up[k].u; L0.src[1] = &c->conv[skip].a;
        }
    }
    for (int k = 0; k < 4; ++k) c->up[k].src = &c->conv[c->up[k].src_layer].a;
    return 0;
}

// (Re)encode every tensor map for batch B.
static int prepare_batch(sdn_ctx* c, int B) {
    if (B == c->B) return 0;
    if (B < 1 || B > c->maxB) return fail("batch %d outside [1, %d]", B, c->maxB);
    static const SegSpec k3x3[9] = {{0, -1, -1}, {0, 0, -1}, {0, 1, -1}, {0, -1, 0}, {0, 0, 0},
                                    {0, 1, 0},   {0, -1, 1}, {0, 0, 1},  {0, 1, 1}};
    for (int i = 0; i < 18; ++i) {
        ConvL& L = c->conv[i];
        // ---- forward
        std::vector<SrcView> av;
        std::vector<SegSpec> segs;
        if (L.first) {
            c->x0.C = first_rows() ? 32 : 64;
            av.push_back(full_view(c->x0));
            segs.push_back({0, 0, 0});
        } else {
            for (int s = 0; s < L.nsrc; ++s) av.push_back(full_view(*L.src[s]));
            for (int tap = 0; tap < 9; ++tap)
                for (int s = 0; s < L.nsrc; ++s) segs.push_back({s, k3x3[tap].dx, k3x3[tap].dy});
        }
        const bool rows_first = L.first && first_rows();
        SDN_OK(build_gemm(c, L.fprop, B, av, segs, L.wf, L.cout, {full_view(L.y)}, L.cout, nullptr, CG_STATS,
                          c->stats_partials, !L.first || rows_first, rows_first ? 1 : 3));
        SDN_OK(build_gemm(c, L.fprop_eval, B, av, segs, L.wf, L.cout, {full_view(L.a)}, L.cout, L.shift, CG_RELU,
                          nullptr, !L.first || rows_first, rows_first ? 1 : 3));
        // ---- data gradient: conv3x3 of dy with flipped / transposed weights
        if (L.has_dgrad) {
            std::vector<SegSpec> dsegs(k3x3, k3x3 + 9);
            std::vector<SrcView> dv;
            int n_per = L.cin;
            if (L.nsrc == 2) {
                const int k = (i - 10) / 2;
                const int skip = 7 - 2 * k;
                dv.push_back(full_view(c->up[k].gu));
                dv.push_back(full_view(c->conv[skip].ga));
                n_per = L.cin / 2;
            } else if (i % 2 == 1) {
                dv.push_back(full_view(c->conv[i - 1].ga));
            } else {
                dv.push_back(full_view(c->conv[i - 1].gp));
            }
            // i odd: the destination is dA of conv[i-1] (never a pooled layer): its BatchNorm-backward sums come
            // out of this kernel's epilogue
            BStatSpec bspec{};
            ConvL* target = (i % 2 == 1) ? &c->conv[i - 1] : nullptr;
            if (target != nullptr) bspec = BStatSpec{&target->y, target->scale, target->shift, target->mean};
            SDN_OK(build_gemm(c, L.dgrad, B, {full_view(L.dy)}, dsegs, L.wd, L.cin, dv, n_per, nullptr, 0, nullptr, true, 3,
                              target ? &bspec : nullptr));
            if (target != nullptr) {
                target->bwd_stats_fused = (L.dgrad.p.flags & CG_BSTATS) != 0;
                target->bwd_stats_parts = L.dgrad.grid * L.dgrad.eg;
                target->bwd_stats_fold = L.dgrad.halo == 3 ? 2 : 1;
            }
            if (L.nsrc == 2) {
                // the first destination is the up-conv output gradient: its per-channel column sums are the
                // ConvTranspose2d bias gradient, and the epilogue can produce them like BatchNorm statistics
                UpL& U = c->up[(i - 10) / 2];
                U.bias_fused = L.dgrad.p.n_tiles == 1 && L.dgrad.block_n <= 128;
                if (U.bias_fused) { L.dgrad.p.flags |= CG_STATS; L.dgrad.p.stats_partials = c->stats_partials; }
            }
        }
        // ---- weight gradient
        std::vector<SrcView> bs;
        if (L.first) bs.push_back(full_view(c->x0));
        else for (int s = 0; s < L.nsrc; ++s) bs.push_back(full_view(*L.src[s]));
        if (rows_first) SDN_OK(build_wgrad(c, L.wgrad, B, {full_view(L.dy)}, L.cout, bs, 9, L.wg, 288, 1));
        else SDN_OK(build_wgrad(c, L.wgrad, B, {full_view(L.dy)}, L.cout, bs, L.first ? 1 : 9, L.wg, L.first ? 64 : 9 * L.cin));
        if (rows_first && L.wgrad.tr2_natoms == 0) return fail("first layer: the row-halo form needs the wgrad_tr kernel");
        if (rows_first && (L.fprop.halo != 1 || L.fprop_eval.halo != 1)) return fail("first layer: the row-halo form needs the row-halo conv kernel");
    }
    for (int k = 0; k < 4; ++k) {
        UpL& U = c->up[k];
        std::vector<SrcView> quads_u, quads_gu;
        for (int q = 0; q < 4; ++q) { quads_u.push_back(quad_view(U.u, q)); quads_gu.push_back(quad_view(U.gu, q)); }
        // the GEMM columns are [quadrant q = 2i + j][co]: quadrants (i,0), (i,1) are adjacent in N and in memory
        static const int pair_on = env_int("SDN_CONVT_PAIRS", 1);
        std::vector<SrcView> pairs_u = {pair_view(U.u, 0), pair_view(U.u, 1)}, pairs_gu = {pair_view(U.gu, 0), pair_view(U.gu, 1)};
        if (pair_on)
            SDN_OK(build_gemm(c, U.fprop, B, {full_view(*U.src)}, {{0, 0, 0}}, U.wf, 4 * U.cout, pairs_u, 2 * U.cout, U.bias4,
                              0, nullptr));
        else
        SDN_OK(build_gemm(c, U.fprop, B, {full_view(*U.src)}, {{0, 0, 0}}, U.wf, 4 * U.cout, quads_u, U.cout, U.bias4,
                          0, nullptr));
        std::vector<SegSpec> qsegs = {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}, {3, 0, 0}};
        ConvL& T = c->conv[U.src_layer];     // bottleneck.3 / dec4.3 / dec3.3 / dec2.3: not pooled
        const BStatSpec bspec{&T.y, T.scale, T.shift, T.mean};
        if (pair_on)
            SDN_OK(build_gemm(c, U.dgrad, B, pairs_gu, {{0, 0, 0}, {1, 0, 0}}, U.wd, U.cin, {full_view(T.ga)}, U.cin, nullptr, 0,
                              nullptr, false, 3, &bspec));
        else
        SDN_OK(build_gemm(c, U.dgrad, B, quads_gu, qsegs, U.wd, U.cin, {full_view(T.ga)}, U.cin, nullptr, 0, nullptr, false, 3,
                          &bspec));
        T.bwd_stats_fused = (U.dgrad.p.flags & CG_BSTATS) != 0;
        T.bwd_stats_parts = U.dgrad.grid * U.dgrad.eg;
        T.bwd_stats_fold = 1;
        // weight gradient: the same pairing halves the dY variants (dense 128-byte rows instead of half-filled
        // ones, the source tile read twice instead of four times); workspace [2][cin][2 * cout] (unpack mode 5)
        static const int wpair_on = env_int("SDN_CONVT_WGRAD_PAIRS", 1);
        U.wgrad_pairs = pair_on && wpair_on;
        if (U.wgrad_pairs) SDN_OK(build_wgrad(c, U.wgrad, B, pairs_gu, 2 * U.cout, {full_view(*U.src)}, 1, U.wg, U.cin));
        else SDN_OK(build_wgrad(c, U.wgrad, B, quads_gu, U.cout, {full_view(*U.src)}, 1, U.wg, U.cin));
    }
    c->B = B;
    return 0;
}

// Grid of a grid-stride kernel = exactly the blocks that can be resident (occupancy x SMs): a larger grid
// runs a partial last wave, and the tail of a 1.3-wave launch is a third of its time.
template <typename K>
static int occ_grid(const sdn_ctx* c, K kernel, long long work_items, int block) {
    static thread_local std::map<const void*, int> cache;
    const void* key = reinterpret_cast<const void*>(kernel);
    auto it = cache.find(key);
    int per_sm;
    if (it == cache.end()) {
        per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
        cache[key] = per_sm;
    } else {
        per_sm = it->second;
    }
    const long long need = (work_items + block - 1) / block;
    return (int)std::max(1LL, std::min(need, (long long)c->num_sms * per_sm));
}

static int zero_fill(sdn_ctx* c, void* p, size_t bytes, cudaStream_t st) {
    if (bytes % 4 != 0) return fail("zero_fill: %zu bytes not a multiple of 4", bytes);
    const size_t words = bytes / 4;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((words / 4 + 255) / 256, (size_t)c->num_sms * 8));
    launch_k(zero_u32_kernel, grid, 256, 0, st, reinterpret_cast<uint32_t*>(p), words);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// bf16 operand cache <- fp32 parameters: one launch for every layer
// part 0: the first PACK_SHALLOW conv layers (needed at once), part 1: everything else, part 2: all
static const int PACK_SHALLOW = 4;
static int pack_params(sdn_ctx* c, bool training, cudaStream_t st, int part = 2) {
    ProfScope ps(c, st, "pack_weights", 0, 0.0, part == 0 ? 0.0 : 7763938.0 * (4 + 2) * (training ? 2 : 1));
    const bool fold = !training;   // eval: BatchNorm scale folded into the forward weights
    PackTable t;
    t.n = 0;
    t.total = 0;
    auto add = [&](const float* w, void* dst, const float* oscale, int mode, int Co, int Ci, int Kpad, int count) {
        PackEntry& e = t.e[t.n++];
        e.w = w; e.dst = dst; e.oscale = oscale; e.mode = mode; e.Co = Co; e.Ci = Ci; e.Kpad = Kpad;
        e.start = t.total;
        t.total += (count + 7) & ~7;     // pack_all_kernel works in runs of 8 elements
    };
    for (int i = 0; i < 18; ++i) {
        if ((part == 0 && i >= PACK_SHALLOW) || (part == 1 && i < PACK_SHALLOW)) continue;
        ConvL& L = c->conv[i];
        const float* w = c->params[L.p_w];
        if (L.first && first_rows()) {
            add(w, L.wf, fold ? L.scale : nullptr, 10, L.cout, L.cin, 32, L.cout * 96);
        } else if (L.first) {
            add(w, L.wf, fold ? L.scale : nullptr, 2, L.cout, L.cin, 64, L.cout * 64);
        } else {
            const int n = 9 * L.cin * L.cout;
            static const int fmode[4] = {0, 5, 8, 11}, dmode[4] = {1, 6, 9, 12};
            const GemmOp& fop = fold ? L.fprop_eval : L.fprop;
            // super-pixel packing: [unit][dy][128 rows][64 k] per 32-channel source (see pack_value modes 11 / 12)
            add(w, L.wf, fold ? L.scale : nullptr, fmode[fop.halo], L.cout, L.cin, fop.swa / 2,
                fop.halo == 3 ? (L.cin / 32) * 3 * 128 * 64 : n);
            if (training) add(w, L.wd, nullptr, dmode[L.dgrad.halo], L.cout, L.cin, L.dgrad.swa / 2,
                              L.dgrad.halo == 3 ? 3 * 128 * 64 : n);
        }
    }
    for (int k = 0; k < 4 && part != 0; ++k) {
        UpL& U = c->up[k];
        const int n = 4 * U.cin * U.cout;
        add(c->params[U.p_w], U.wf, nullptr, 3, U.cout, U.cin, 0, n);
        if (training) add(c->params[U.p_w], U.wd, nullptr, 4, U.cout, U.cin, 0, n);
        add(c->params[U.p_b], U.bias4, nullptr, 7, U.cout, 0, 0, 4 * U.cout);
    }
    launch_k(pack_all_kernel, part == 0 ? c->num_sms : c->num_sms * 4, 256, 0, st, t);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int run_bn_relu(sdn_ctx* c, ConvL& L, int B, cudaStream_t st) {
    const int H = L.y.H, W = L.y.W, C = L.cout;
    if (L.pooled_out) {
        const long long items = (long long)B * (H / 2) * (W / 2) * (C / 8);
        launch_k(bn_relu_pool_kernel<true>, occ_grid(c, bn_relu_pool_kernel<true>, items, 256), 256, 0, st, L.y.p, L.scale, L.shift, L.a.p, L.pool.p,
                                                                         L.amax, B, H, W, C);
    } else {
        const long long items = (long long)B * H * W * (C / 8);
        launch_k(bn_relu_pool_kernel<false>, occ_grid(c, bn_relu_pool_kernel<false>, items, 256), 256, 0, st, L.y.p, L.scale, L.shift, L.a.p, nullptr,
                                                                          nullptr, B, H, W, C);
    }
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int forward_impl(sdn_ctx* c, const float* x, float* disp, float* logvar, int B, int training, int dirty,
                        cudaStream_t st) {
    if (c->pre_only) return fail("sdn_forward: context was created with SDN_CTX_PREPROCESS_ONLY");
    if (!c->have_params) return fail("sdn_forward: call sdn_set_params first");
    SDN_OK(prepare_batch(c, B));
    c->pdl = t_pdl = (long long)B * c->H * c->W <= 48LL * 240 * 320 ? 1 : 0;
    if (!training && dirty) {
        // eval: scale/shift from the running statistics, folded into the weights at pack time
        for (int i = 0; i < 18; ++i) {
            ConvL& L = c->conv[i];
            launch_k(bn_prepare_eval_kernel, (L.cout + 127) / 128, 128, 0, st, L.cout, c->params[L.p_gamma],
                                                                         c->params[L.p_beta], c->bn_rm[L.bn],
                                                                         c->bn_rv[L.bn], 1e-5f, L.scale, L.shift);
            ++c->launches;
        }
    }
    // The deep layers hold 99 % of the packed bytes and are not needed before enc3: their packing runs on the side
    // stream next to the first-layer taps kernel and the level-1/2 convs (0.09 ms per step whatever the batch,
    // i.e. 1.5 % of a 32-pair step), the main stream joins right before conv 4.
    static const int pack_overlap = env_int("SDN_PACK_OVERLAP", 0);   // measured: the low-priority side stream delays the join (6.11 -> 6.17 ms at 32 pairs)
    bool pack_pending = false;
    if (dirty) {
        if (pack_overlap && training && c->side != nullptr && !c->prof) {
            CUDA_OK(cudaEventRecord(c->ev_fork, st));
            CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
            SDN_OK(pack_params(c, true, c->side, 1));
            CUDA_OK(cudaEventRecord(c->ev_pack, c->side));
            SDN_OK(pack_params(c, true, st, 0));
            pack_pending = true;
        } else {
            SDN_OK(pack_params(c, training != 0, st));
        }
    }
    const int H = c->H, W = c->W;
    {
        const double px = (double)B * H * W;
        ProfScope ps(c, st, "im2col_first", 0, 0.0, px * (6 * 4 + c->x0.C * 2));
        if (first_rows())
            launch_k(im2col_rows_kernel<6>, B * ((H + IM2COL_ROWS - 1) / IM2COL_ROWS) * ((W + IM2COL_PX - 1) / IM2COL_PX), 256, 0, st,
                     x, c->x0.p, B, H, W);
        else
            launch_k(im2col_first_kernel<6>, B * ((H + IM2COL_ROWS - 1) / IM2COL_ROWS) * ((W + IM2COL_PX - 1) / IM2COL_PX), 256, 0, st,
                     x, c->x0.p, B, H, W);
        ++c->launches;
    }
    CUDA_OK(cudaGetLastError());
    for (int i = 0; i < 18; ++i) {
        ConvL& L = c->conv[i];
        if (pack_pending && i == PACK_SHALLOW) {
            CUDA_OK(cudaStreamWaitEvent(st, c->ev_pack, 0));
            pack_pending = false;
        }
        if (i >= 10 && i % 2 == 0) {
            UpL& U = c->up[(i - 10) / 2];
            const double px = (double)B * U.src->H * U.src->W;
            ProfScope ps(c, st, "convT_fprop", 100 + (i - 10) / 2, 2.0 * px * U.cin * 4 * U.cout,
                         px * (U.cin + 4 * U.cout) * 2);
            SDN_OK(launch_cg(c, U.fprop, st));
        }
        if (!training) {
            const double px = (double)B * L.y.H * L.y.W;
            {
                ProfScope ps(c, st, "conv_fprop_eval", i, 2.0 * px * L.cout * 9 * L.cin,
                             px * ((L.first ? c->x0.C : L.cin) + L.cout) * 2);
                SDN_OK(launch_cg(c, L.fprop_eval, st));
            }
            if (L.pooled_out) {
                ProfScope ps(c, st, "maxpool", i, 0.0, px * L.cout * 2 * 1.25);
                const long long items = (long long)B * (L.y.H / 2) * (L.y.W / 2) * (L.cout / 8);
                launch_k(maxpool2x2_kernel, occ_grid(c, maxpool2x2_kernel, items, 256), 256, 0, st, L.a.p, L.pool.p, B, L.y.H, L.y.W, L.cout);
                ++c->launches;
            }
            continue;
        }
        GemmOp op = L.fprop;
        op.p.flags = (op.p.flags & ~CG_STATS) | CG_STATS;
#ifdef SDN_FORENSICS
        {
            // timing experiments only (results are garbage): never compiled into the product library
            static const int dbg = env_int("SDN_DEBUG_ABLATE", 0) &
                                   (CG_DBG_NOMMA | CG_DBG_NOEPI | CG_DBG_NOLOADA | CG_DBG_NOSTORE | CG_DBG_NOSTATS);
            op.p.flags |= dbg;
            static const int dbg_layer = env_int("SDN_DEBUG_TRACE_LAYER", -1);
            op.p.dbg = (i == dbg_layer) ? c->dbg : nullptr;
        }
#endif
        {
            const double px = (double)B * L.y.H * L.y.W;
            ProfScope ps(c, st, "conv_fprop", i, 2.0 * px * L.cout * 9 * L.cin, px * ((L.first ? c->x0.C : L.cin) + L.cout) * 2);
            SDN_OK(launch_cg(c, op, st));
        }
        // (layer 17 only finalizes its statistics here: its BatchNorm+ReLU runs inside the head kernel)
        ProfScope ps(c, st, "bn_relu_pool", i, 0.0,
                     i == 17 ? 0.0 : (double)B * L.y.H * L.y.W * L.cout * 2 * (L.pooled_out ? 2.25 : 2.0));
        if (training) {
            const double count = (double)B * L.y.H * L.y.W;
            launch_k(bn_finalize_train_kernel, (L.cout * 32 + 255) / 256, 256, 0, st, 
                c->stats_partials, op.grid * op.eg, L.cout, count, c->params[L.p_gamma], c->params[L.p_beta], c->bn_rm[L.bn],
                c->bn_rv[L.bn], (long long*)c->bn_nbt[L.bn], 1e-5f, 0.1f, L.scale, L.shift, L.mean, L.rstd, op.halo == 3 ? 2 : 1);
            ++c->launches;
        }
        // dec1.block.3 (the last conv layer): its activated output feeds only the two 1x1 heads, which apply
        // scale / shift / ReLU on the fly from y (head_kernel): no BatchNorm+ReLU pass, no `a` tensor
        if (i != 17) SDN_OK(run_bn_relu(c, L, B, st));
    }
    if (disp != nullptr) {
        const long long npix = (long long)B * H * W;
        ProfScope ps(c, st, "head_fwd", 0, 0.0, (double)npix * (64 + (logvar ? 8 : 4)));
        const ConvL& T = c->conv[17];
        launch_k(head_kernel<0>, occ_grid(c, head_kernel<0>, 2 * npix, HEAD_THREADS), HEAD_THREADS, 0, st, training ? T.y.p : T.a.p, c->params[62], c->params[63],
                                                              c->params[64], c->params[65], disp, logvar, nullptr,
                                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                              nullptr, nullptr, npix, training ? T.scale : nullptr,
                                                              training ? T.shift : nullptr, training ? T.mean : nullptr, nullptr);
        ++c->launches;
        CUDA_OK(cudaGetLastError());
    }
    c->have_forward_train = training != 0;
    return 0;
}

// -------------------------------------------------------------- backward
static int bn_backward(sdn_ctx* c, ConvL& L, int B, cudaStream_t st) {
    const int H = L.y.H, W = L.y.W, C = L.cout;
    const double count = (double)B * H * W;
    const bool pool = L.pooled_out;
    const long long items = pool ? (long long)B * (H / 2) * (W / 2) * (C / 8) : (long long)B * H * W * (C / 8);
    if (L.bwd_stats_fused) {
        // the kernel that produced `ga` already reduced sum(dz) and sum(dz * (y - mean)) per CTA (CG_BSTATS)
        launch_k(bn_bwd_finalize_kernel, (C * 32 + 255) / 256, 256, 0, st, c->stats_partials, L.bwd_stats_parts, C, count,
                 L.c1, L.c2, c->grads[L.p_gamma], c->grads[L.p_beta], c->accumulate, (const float*)L.rstd, L.bwd_stats_fold);
        ++c->launches;
    } else {
        int grid = pool ? occ_grid(c, bn_bwd_reduce_kernel<true>, items, 256) : occ_grid(c, bn_bwd_reduce_kernel<false>, items, 256);
        grid = std::min(grid, BWD_BLOCKS);
        if (pool)
            launch_k(bn_bwd_reduce_kernel<true>, grid, 256, 0, st, L.y.p, L.ga.p, L.gp.p, L.amax, L.scale, L.shift, L.mean, L.rstd,
                                                             c->bwd_partials, B, H, W, C);
        else
            launch_k(bn_bwd_reduce_kernel<false>, grid, 256, 0, st, L.y.p, L.ga.p, nullptr, nullptr, L.scale, L.shift, L.mean, L.rstd,
                                                              c->bwd_partials, B, H, W, C);
        ++c->launches;
        launch_k(bn_bwd_finalize_kernel, (C * 32 + 255) / 256, 256, 0, st, c->bwd_partials, grid, C, count, L.c1, L.c2,
                                                                c->grads[L.p_gamma], c->grads[L.p_beta], c->accumulate, (const float*)L.rstd, 1);
        ++c->launches;
    }
    if (pool)
        launch_k(bn_bwd_apply_kernel<true>, occ_grid(c, bn_bwd_apply_kernel<true>, items, 256), 256, 0, st, L.y.p, L.ga.p, L.gp.p, L.amax, L.scale, L.shift, L.mean, L.rstd, L.c1,
                                                         L.c2, L.dy.p, B, H, W, C);
    else
        launch_k(bn_bwd_apply_kernel<false>, occ_grid(c, bn_bwd_apply_kernel<false>, items, 256), 256, 0, st, L.y.p, L.ga.p, nullptr, nullptr, L.scale, L.shift, L.mean, L.rstd,
                                                          L.c1, L.c2, L.dy.p, B, H, W, C);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int conv_backward(sdn_ctx* c, int i, int B, cudaStream_t st) {
    ConvL& L = c->conv[i];
    const double px = (double)B * L.y.H * L.y.W;
    {
        // reduce: y + g (+ gp/4); apply: y + g (+ gp/4) + dy
        // reduce: y + g (+ gp/4) [skipped when the producer of g reduced the sums]; apply: y + g (+ gp/4) + dy
        ProfScope ps(c, st, "bn_bwd", i, 0.0, px * L.cout * 2 * (L.pooled_out ? 5.5 : (L.bwd_stats_fused ? 3.0 : 5.0)));
        SDN_OK(bn_backward(c, L, B, st));
    }
    {
#ifdef SDN_FORENSICS
        static const int wdbg_layer = env_int("SDN_DEBUG_TRACE_WGRAD", -1);
        L.wgrad.p.dbg = (i == wdbg_layer) ? c->dbg : nullptr;
#endif
        static const int overlap_env = env_int("SDN_WGRAD_OVERLAP", 1);
        // per-op profiling times kernels with events on `st`: keep everything there while it is on
        const bool overlap = overlap_env && c->side != nullptr && !c->prof && L.has_dgrad;
        cudaStream_t ws = st;
        if (overlap) {
            CUDA_OK(cudaEventRecord(c->ev_fork, st));            // dy of this layer is ready
            CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
            ws = c->side;
            c->side_dirty = true;
        }
        // the data gradient first: it heads the critical path (dgrad -> BN backward of the next layer)
        if (overlap) SDN_OK(launch_cg(c, L.dgrad, st));
        {
            ProfScope ps(c, ws, "conv_wgrad", i, 2.0 * px * L.cout * 9 * L.cin, px * ((L.first ? c->x0.C : L.cin) + L.cout) * 2);
            SDN_OK(launch_wg(c, L.wgrad, ws));
        }
        const bool dgrad_reads_y = L.has_dgrad && (L.dgrad.p.flags & CG_BSTATS) != 0;   // + the target layer's y tile
        ProfScope ps2(c, st, L.has_dgrad ? "conv_dgrad" : "grad_unpack", i, L.has_dgrad ? 2.0 * px * L.cout * 9 * L.cin : 0.0,
                      L.has_dgrad ? px * (L.cin * (dgrad_reads_y ? 2 : 1) + L.cout) * 2 : 0.0);
        if (c->grads[L.p_w] != nullptr) {
            const int n = 9 * L.cin * L.cout;
            launch_k(unpack_grad_kernel, occ_grid(c, unpack_grad_kernel, n, 256), 256, 0, ws, L.wg, c->grads[L.p_w], L.first ? (first_rows() ? 4 : 2) : 0, L.cout, L.cin,
                     c->accumulate);
            ++c->launches;
        }
        if (L.has_dgrad && !overlap) SDN_OK(launch_cg(c, L.dgrad, st));
    }
    if (L.has_dgrad && L.nsrc == 2 && c->up[(i - 10) / 2].bias_fused) {
        UpL& U = c->up[(i - 10) / 2];
        launch_k(colsum_partials_kernel, (U.cout * 32 + 255) / 256, 256, 0, st, c->stats_partials,
                 L.dgrad.grid * L.dgrad.eg, L.dgrad.p.n_total, U.cout, U.bg);
        ++c->launches;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int up_backward(sdn_ctx* c, int k, int B, cudaStream_t st) {
    UpL& U = c->up[k];
    const long long npix = (long long)B * U.gu.H * U.gu.W;
    dim3 g(U.cout / 8, (unsigned)std::max(1LL, std::min((npix + 255) / 256, (long long)(c->num_sms * 2))));
    const double pxin = (double)B * U.src->H * U.src->W;
    if (!U.bias_fused) {
        ProfScope ps(c, st, "convT_bias_grad", 100 + k, 0.0, (double)npix * U.cout * 2);
        launch_k(colsum_kernel, g, 256, 0, st, U.gu.p, npix, U.cout, U.bg, 0);
        ++c->launches;
    }
    // same overlap as the 3x3 convs: the ConvTranspose2d data gradient goes first on `st`, its weight gradient
    // (and the bias / weight unpack) run next to whatever follows on the low-priority side stream
    static const int overlap_env = env_int("SDN_WGRAD_OVERLAP", 1);
    const bool overlap = overlap_env && c->side != nullptr && !c->prof;
    cudaStream_t ws = st;
    if (overlap) {
        CUDA_OK(cudaEventRecord(c->ev_fork, st));
        CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        ws = c->side;
        c->side_dirty = true;
        SDN_OK(launch_cg(c, U.dgrad, st));
    }
    {
        ProfScope ps(c, ws, "convT_wgrad", 100 + k, 2.0 * pxin * U.cin * 4 * U.cout, pxin * (U.cin + 4 * U.cout) * 2);
        SDN_OK(launch_wg(c, U.wgrad, ws));
    }
    ProfScope ps2(c, st, "convT_dgrad", 100 + k, 2.0 * pxin * U.cin * 4 * U.cout,
                  pxin * (U.cin * ((U.dgrad.p.flags & CG_BSTATS) ? 2 : 1) + 4 * U.cout) * 2);
    if (c->grads[U.p_w] != nullptr) {
        const int n = 4 * U.cin * U.cout;
        launch_k(unpack_grad_kernel, occ_grid(c, unpack_grad_kernel, n, 256), 256, 0, ws, U.wg, c->grads[U.p_w], U.wgrad_pairs ? 5 : 3, U.cout, U.cin, c->accumulate);
        ++c->launches;
    }
    if (c->grads[U.p_b] != nullptr) {
        launch_k(copy_f32_kernel, 1, 256, 0, ws, U.bg, c->grads[U.p_b], U.cout, c->accumulate);
        ++c->launches;
    }
    if (!overlap) SDN_OK(launch_cg(c, U.dgrad, st));
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int head_grads_out(sdn_ctx* c, cudaStream_t st) {
    // head_grads: [0,32) dW_d, [32] db_d, [33,65) dW_l, [65] db_l  -> params 62..65
    const int off[4] = {0, 32, 33, 65};
    const int len[4] = {32, 1, 32, 1};
    for (int j = 0; j < 4; ++j)
        if (c->grads[62 + j] != nullptr) {
            launch_k(copy_f32_kernel, 1, 32, 0, st, c->head_grads + off[j], c->grads[62 + j], len[j], c->accumulate);
            ++c->launches;
        }
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int backward_prologue(sdn_ctx* c, int accumulate, cudaStream_t st) {
    if (!c->have_forward_train) return fail("backward without a training-mode forward");
    c->accumulate = accumulate;
    SDN_OK(zero_fill(c, c->wg_all, c->wg_all_bytes, st));
    SDN_OK(zero_fill(c, c->head_grads, 128 * sizeof(float), st));
    return 0;
}

// Photometric chain + blur + noise in place on input[B,6,H,W] (dataset.py:248-270), after a kernel that left the
// per-tile gray partial sums in gray_part.
static int run_augment(sdn_ctx* c, float* input, int B, const AugParams* aug, float* gray_part, float* blur_tmp, int parts,
                       cudaStream_t st) {
    const int H = c->H, W = c->W;
    ProfScope ps2(c, st, "pre_augment", 0, 0.0, (double)B * H * W * 6 * 4 * 2);
    if (((long long)H * W) % 4 == 0) {
        launch_k(view_mean_kernel, (2 * B * 32 + 255) / 256, 256, 0, st, (const float*)gray_part, parts, 2 * B,
                 1.0 / ((double)H * (double)W), c->view_mean);
        launch_k(augment_point4_kernel, dim3((H * W / 4 + 255) / 256, 2 * B), 256, 0, st, input, B, H, W, aug,
                 (const float*)c->view_mean, blur_tmp);
        c->launches += 2;
    } else {
        launch_k(augment_point_kernel, dim3((H * W + 255) / 256, 2 * B), 256, 0, st, input, B, H, W, aug, gray_part, parts, blur_tmp);
        ++c->launches;
    }
    launch_k(blur_noise_kernel<5>, dim3((W + 31) / 32, (H + 7) / 8, 2 * B), dim3(32, 8), 0, st, input, B, H, W, aug, blur_tmp);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int preprocess_chunk(sdn_ctx* c, const uint8_t* left, const uint8_t* right, const uint8_t* disparity, int B,
                            int Hs, int Ws, const AugParams* aug, float* input, float* target, uint8_t* mask,
                            unsigned long long* valid_count, unsigned flags, float* gray_part, float* blur_tmp,
                            cudaStream_t st) {
    const int H = c->H, W = c->W;
    const int xblocks = (W + 127) / 128, yblocks = (H + PRE_ROWS - 1) / PRE_ROWS;
    const int parts = xblocks * yblocks;
    // 3 uint8 sources read + fp32 input/target + u8 mask written per sample
    // decode+resize moves the SURVEY 8(d) bytes; the augmentation passes re-read / re-write the fp32 views
    {
    ProfScope ps(c, st, "pre_decode_resize", 0, 0.0, (double)B * (3.0 * Hs * Ws * 3 + (double)H * W * (6 * 4 + 4 + 1)));
    // shared-memory staged kernel when the source rows are 16-byte aligned and the footprint fits
    const int max_rows = (int)std::ceil((double)PRE_ROWS * Hs / H) + 3;
    const int row_bytes = (((int)std::ceil(128.0 * Ws / W) + 3) * 3 + 47) & ~15;
    const size_t pre_smem = (size_t)3 * max_rows * row_bytes;      // one tile buffer: one tile per CTA (see the kernel)
    const bool aligned = ((Ws * 3) % 16 == 0) && (((uintptr_t)left | (uintptr_t)right | (uintptr_t)disparity) % 16 == 0);
    const bool staged = aligned && pre_smem <= 200 * 1024 && !(flags & SDN_PREPROCESS_DIRECT);
    if (staged) {
        const int grid = parts * B;
        if (flags & SDN_RESIZE_FOURTERM)
            launch_k(decode_resize_smem_kernel<true>, grid, PRE_THREADS, pre_smem, st,
                left, right, disparity, B, Hs, Ws, H, W, input, target, mask, valid_count, aug,
                aug ? gray_part : nullptr, parts, max_rows, row_bytes);
        else
            launch_k(decode_resize_smem_kernel<false>, grid, PRE_THREADS, pre_smem, st,
                left, right, disparity, B, Hs, Ws, H, W, input, target, mask, valid_count, aug,
                aug ? gray_part : nullptr, parts, max_rows, row_bytes);
    } else if (flags & SDN_RESIZE_FOURTERM)
        launch_k(decode_resize_kernel<true>, dim3(parts, B), 128, 0, st, left, right, disparity, B, Hs, Ws, H, W, input,
                                                                   target, mask, valid_count, aug,
                                                                   aug ? gray_part : nullptr, parts);
    else
        launch_k(decode_resize_kernel<false>, dim3(parts, B), 128, 0, st, left, right, disparity, B, Hs, Ws, H, W, input,
                                                                    target, mask, valid_count, aug,
                                                                    aug ? gray_part : nullptr, parts);
    ++c->launches;
    }
    if (aug != nullptr) SDN_OK(run_augment(c, input, B, aug, gray_part, blur_tmp, parts, st));
    CUDA_OK(cudaGetLastError());
    return 0;
}



// ------------------------------------------------------------------- NCCL
// Resolved with dlopen at sdn_comm_init time: inside a PyTorch process "libnccl.so.2" is already mapped (torch's
// bundled copy) and RTLD_NOLOAD returns exactly that one, so the process never holds two NCCL versions; a plain
// C host gets the system library.  No link-time dependency: single-GPU users never need NCCL at all.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.handle != nullptr) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return fail("sdn_comm: cannot load libnccl.so.2 (%s)", dlerror());
    NcclApi a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString)
        return fail("sdn_comm: libnccl.so.2 lacks a required symbol");
    a.handle = h;
    g_nccl = a;
    return 0;
}
#define NCCL_OK(call)                                                                                     \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// elements of parameter i in StereoUNet.parameters() order (model.py:59-77)
static long long param_numel(const sdn_ctx* c, int i) {
    for (int l = 0; l < 18; ++l) {
        const ConvL& L = c->conv[l];
        if (i == L.p_w) return 9LL * L.cin * L.cout;
        if (i == L.p_gamma || i == L.p_beta) return L.cout;
    }
    for (int k = 0; k < 4; ++k) {
        const UpL& U = c->up[k];
        if (i == U.p_w) return 4LL * U.cin * U.cout;
        if (i == U.p_b) return U.cout;
    }
    return (i == 62 || i == 64) ? 32 : 1;   // the two 1x1 heads: weight [1,32,1,1], bias [1]
}

// The gradient destinations of backward stage `stage` as ONE contiguous fp32 range (the all-reduce bucket).
static int stage_bucket(sdn_ctx* c, int stage, float** base, long long* count) {
    int first = 0, num = 0;
    SDN_OK(sdn_stage_param_range(stage, &first, &num));
    long long total = 0;
    for (int i = first; i < first + num; ++i) {
        if (c->grads[i] == nullptr) return fail("data-parallel step: gradient destination %d is not bound", i);
        if (c->grads[i] != c->grads[first] + total)
            return fail("data-parallel step: gradients %d..%d must be consecutive views of one flat buffer "
                        "(parameters() order), parameter %d is not", first, first + num - 1, i);
        total += param_numel(c, i);
    }
    *base = c->grads[first];
    *count = total;
    return 0;
}

// ------------------------------------------------------------------ C ABI
extern "C" {

const char* sdn_last_error(void) { return g_err.c_str(); }
int sdn_version(void) { return 200; }

int sdn_create(sdn_ctx** out, int device, int max_batch, int H, int W, unsigned flags) {
    if (out == nullptr) return fail("sdn_create: out is NULL");
    const bool pre_only_req = (flags & SDN_CTX_PREPROCESS_ONLY) != 0;
    if (H < 1 || W < 1) return fail("H and W must be positive (got %dx%d)", H, W);
    if (!pre_only_req && (H % 16 != 0 || W % 16 != 0)) return fail("H and W must be positive multiples of 16 (got %dx%d)", H, W);
    if (max_batch < 1) return fail("max_batch must be >= 1");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("device %d not available (%d devices)", device, ndev);
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("sdn requires an sm_100a (B200) device; device %d is sm_%d%d", device, prop.major, prop.minor);
    SDN_OK(load_encode());
    SDN_OK(set_smem_attrs());
    sdn_ctx* c = new sdn_ctx();
    c->device = device; c->maxB = max_batch; c->H = H; c->W = W;
    c->num_sms = prop.multiProcessorCount;
    c->pre_only = (flags & SDN_CTX_PREPROCESS_ONLY) != 0;
    int r = plan_and_alloc(c);
    if (r != 0) { if (c->ws) cudaFree(c->ws); delete c; return r; }
    if (!c->pre_only) {
        int lo = 0, hi = 0;
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_OK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, lo));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
    }
    *out = c;
    return 0;
}

int sdn_destroy(sdn_ctx* c) {
    if (c == nullptr) return 0;
    cudaSetDevice(c->device);
    sdn_comm_destroy(c);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_pack) cudaEventDestroy(c->ev_pack);
    if (c->ws) cudaFree(c->ws);
    delete c;
    return 0;
}

int64_t sdn_workspace_bytes(const sdn_ctx* c) { return c ? (int64_t)c->ws_bytes : 0; }
int64_t sdn_launch_count(const sdn_ctx* c) { return c ? c->launches : 0; }

int sdn_set_params(sdn_ctx* c, const float* const* params, float* const* grads, float* const* rm, float* const* rv,
                   int64_t* const* nbt) {
    if (c == nullptr || params == nullptr) return fail("sdn_set_params: NULL argument");
    for (int i = 0; i < SDN_NUM_PARAMS; ++i) {
        if (params[i] == nullptr) return fail("sdn_set_params: param %d is NULL", i);
        c->params[i] = params[i];
        c->grads[i] = grads ? grads[i] : nullptr;
    }
    for (int i = 0; i < SDN_NUM_BN; ++i) {
        if (rm == nullptr || rv == nullptr || rm[i] == nullptr || rv[i] == nullptr)
            return fail("sdn_set_params: BatchNorm buffer %d is NULL", i);
        c->bn_rm[i] = rm[i];
        c->bn_rv[i] = rv[i];
        c->bn_nbt[i] = nbt ? nbt[i] : nullptr;
    }
    c->have_params = true;
    return 0;
}

int sdn_forward(sdn_ctx* c, const float* x, float* disp, float* logvar, int B, int training, int params_dirty,
                void* stream) {
    if (c == nullptr || x == nullptr) return fail("sdn_forward: NULL argument");
    SDN_ENTER(c);
    return forward_impl(c, x, disp, logvar, B, training, params_dirty, (cudaStream_t)stream);
}

int sdn_backward_begin(sdn_ctx* c, const float* g_disp, const float* g_logvar, int accumulate, void* stream) {
    if (c == nullptr || g_disp == nullptr) return fail("sdn_backward_begin: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    SDN_OK(backward_prologue(c, accumulate, st));
    const long long npix = (long long)c->B * c->H * c->W;
    ProfScope ps(c, st, "head_bwd", 0, 0.0, (double)npix * (64 + 8 + 64));
    ConvL& T = c->conv[17];
    const int hgrid = occ_grid(c, head_kernel<1>, 2 * npix, HEAD_THREADS);
    launch_k(head_kernel<1>, hgrid, HEAD_THREADS, 0, st, T.y.p, c->params[62], c->params[63], c->params[64],
                                                          c->params[65], nullptr, nullptr, g_disp, g_logvar, nullptr,
                                                          nullptr, nullptr, nullptr, nullptr, T.ga.p,
                                                          c->head_grads, npix, T.scale, T.shift, T.mean, c->stats_partials);
    T.bwd_stats_fused = true; T.bwd_stats_parts = hgrid; T.bwd_stats_fold = 1;
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return head_grads_out(c, st);
}

int sdn_count_valid(sdn_ctx* c, const float* target, const uint8_t* mask, int B, unsigned long long* count_out,
                    void* stream) {
    if (c == nullptr || target == nullptr || mask == nullptr || count_out == nullptr)
        return fail("sdn_count_valid: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    SDN_OK(zero_fill(c, count_out, sizeof(unsigned long long), st));
    const long long npix = (long long)B * c->H * c->W;
    launch_k(mask_count_kernel, occ_grid(c, mask_count_kernel, npix, 256), 256, 0, st, target, mask, npix, count_out);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int sdn_loss_begin(sdn_ctx* c, const float* target, const uint8_t* mask, float* disp, float* logvar, double* sums4,
                   unsigned long long* count, const unsigned long long* n_norm_dev, int with_backward, int accumulate,
                   void* stream) {
    if (c == nullptr || target == nullptr || mask == nullptr || sums4 == nullptr || count == nullptr)
        return fail("sdn_loss_begin: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    if (c->B < 1) return fail("sdn_loss_begin: no forward has run");
    const long long npix = (long long)c->B * c->H * c->W;
    if (with_backward) {
        SDN_OK(backward_prologue(c, accumulate, st));
        if (n_norm_dev == nullptr) {
            SDN_OK(sdn_count_valid(c, target, mask, c->B, c->n_local, stream));
            n_norm_dev = c->n_local;
        }
    } else {
        SDN_OK(zero_fill(c, c->head_grads, 128 * sizeof(float), st));
        SDN_OK(zero_fill(c, c->n_local, sizeof(unsigned long long), st));
        n_norm_dev = c->n_local;  // zero -> gradients are zero and unused
    }
    // the gradient buffer of dec1's output doubles as scratch on the metrics-only path
    ProfScope ps(c, st, "head_loss", 0, 0.0, (double)npix * (64 + 4 + 1 + 64));
    ConvL& T = c->conv[17];
    const bool train_fwd = c->have_forward_train;     // the forward just run was a training forward: d1 = y (see forward_impl)
    const int hgrid = occ_grid(c, head_kernel<2>, 2 * npix, HEAD_THREADS);
    launch_k(head_kernel<2>, hgrid, HEAD_THREADS, 0, st, train_fwd ? T.y.p : T.a.p, c->params[62], c->params[63], c->params[64],
                                                          c->params[65], disp, logvar, nullptr, nullptr, target, mask,
                                                          n_norm_dev, sums4, count, T.ga.p, c->head_grads,
                                                          npix, train_fwd ? T.scale : nullptr, train_fwd ? T.shift : nullptr,
                                                          train_fwd ? T.mean : nullptr, with_backward ? c->stats_partials : nullptr);
    if (with_backward) { T.bwd_stats_fused = true; T.bwd_stats_parts = hgrid; T.bwd_stats_fold = 1; }
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    if (with_backward) return head_grads_out(c, st);
    return 0;
}

int sdn_stage_param_range(int stage, int* first, int* num) {
    // heads + dec1..dec3 + up1..up3 | dec4 + up4 | bottleneck | enc4 + enc3 | enc2 + enc1.  The last bucket is the
    // only one whose all-reduce cannot hide behind later backward work, so it is the smallest (0.26 MB).
    static const int f[SDN_NUM_STAGES] = {38, 30, 24, 12, 0};
    static const int n[SDN_NUM_STAGES] = {28, 8, 6, 12, 12};
    if (stage < 0 || stage >= SDN_NUM_STAGES || first == nullptr || num == nullptr)
        return fail("sdn_stage_param_range: bad stage %d", stage);
    *first = f[stage];
    *num = n[stage];
    return 0;
}

int sdn_backward_stage(sdn_ctx* c, int stage, void* stream) {
    if (c == nullptr) return fail("sdn_backward_stage: NULL ctx");
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    if (!c->have_forward_train) return fail("backward without a training-mode forward");
    const int B = c->B;
    switch (stage) {
        case 0:  // dec1, up1, dec2, up2, dec3, up3 (+ heads, done in *_begin)
            SDN_OK(conv_backward(c, 17, B, st)); SDN_OK(conv_backward(c, 16, B, st)); SDN_OK(up_backward(c, 3, B, st));
            SDN_OK(conv_backward(c, 15, B, st)); SDN_OK(conv_backward(c, 14, B, st)); SDN_OK(up_backward(c, 2, B, st));
            SDN_OK(conv_backward(c, 13, B, st)); SDN_OK(conv_backward(c, 12, B, st)); SDN_OK(up_backward(c, 1, B, st));
            break;
        case 1:  // dec4, up4
            SDN_OK(conv_backward(c, 11, B, st)); SDN_OK(conv_backward(c, 10, B, st)); SDN_OK(up_backward(c, 0, B, st));
            break;
        case 2:  // bottleneck
            SDN_OK(conv_backward(c, 9, B, st)); SDN_OK(conv_backward(c, 8, B, st));
            break;
        case 3:  // enc4, enc3
            for (int i = 7; i >= 4; --i) SDN_OK(conv_backward(c, i, B, st));
            break;
        case 4:  // enc2, enc1
            for (int i = 3; i >= 0; --i) SDN_OK(conv_backward(c, i, B, st));
            break;
        default:
            return fail("sdn_backward_stage: bad stage %d", stage);
    }
    if (c->side_dirty) {   // the stage's weight gradients must be complete before the caller reduces / applies them
        CUDA_OK(cudaEventRecord(c->ev_join, c->side));
        CUDA_OK(cudaStreamWaitEvent(st, c->ev_join, 0));
        c->side_dirty = false;
    }
    return 0;
}

// ------------------------------------------------------------ data parallel
int sdn_comm_unique_id(void* out128) {
    if (out128 == nullptr) return fail("sdn_comm_unique_id: NULL argument");
    SDN_OK(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

int sdn_comm_init(sdn_ctx* c, const void* id128, int rank, int world) {
    if (c == nullptr || id128 == nullptr) return fail("sdn_comm_init: NULL argument");
    if (world < 1 || rank < 0 || rank >= world) return fail("sdn_comm_init: rank %d outside world %d", rank, world);
    if (c->comm != nullptr) return fail("sdn_comm_init: the context already has a communicator");
    SDN_ENTER(c);
    SDN_OK(load_nccl());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_OK(g_nccl.CommInitRank(&c->comm, world, id, rank));
    c->comm_rank = rank; c->comm_world = world;
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_bucket, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    return 0;
}

int sdn_comm_destroy(sdn_ctx* c) {
    if (c == nullptr) return 0;
    if (c->comm != nullptr) {
        cudaSetDevice(c->device);
        if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    if (c->comm_stream) { cudaStreamDestroy(c->comm_stream); c->comm_stream = nullptr; }
    if (c->ev_bucket) { cudaEventDestroy(c->ev_bucket); c->ev_bucket = nullptr; }
    if (c->ev_comm) { cudaEventDestroy(c->ev_comm); c->ev_comm = nullptr; }
    c->comm_world = 1; c->comm_rank = 0;
    return 0;
}

int sdn_comm_world(const sdn_ctx* c) { return c ? c->comm_world : 0; }

int sdn_comm_allreduce(sdn_ctx* c, void* buf, int64_t count, int dtype, void* stream) {
    if (c == nullptr || buf == nullptr || count < 0) return fail("sdn_comm_allreduce: bad argument");
    if (c->comm == nullptr) return c->comm_world == 1 ? 0 : fail("sdn_comm_allreduce: no communicator");
    SDN_ENTER(c);
    const ncclDataType_t t = dtype == SDN_F32 ? ncclFloat32 : dtype == SDN_F64 ? ncclFloat64 : ncclUint64;
    if (dtype != SDN_F32 && dtype != SDN_F64 && dtype != SDN_U64) return fail("sdn_comm_allreduce: dtype %d", dtype);
    NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)count, t, ncclSum, c->comm, (cudaStream_t)stream));
    return 0;
}

// One call = the loop body of run_epoch for one assembled batch (train.py:325-342, everything but
// optimizer.step()): forward -> valid count (all-reduced: the loss normaliser is the GLOBAL count) -> fused
// loss + metric sums + head backward -> four backward stages, each stage's gradient bucket all-reduced on the
// communicator stream while the next stage runs -> join.
int sdn_train_step(sdn_ctx* c, const float* x, const float* target, const uint8_t* mask, int B, int params_dirty,
                   double* sums4, unsigned long long* count, unsigned long long* n_norm, unsigned flags,
                   void* stream) {
    if (c == nullptr || x == nullptr || target == nullptr || mask == nullptr || sums4 == nullptr || count == nullptr ||
        n_norm == nullptr)
        return fail("sdn_train_step: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    const bool dp = c->comm != nullptr && c->comm_world > 1;
    float* bucket[SDN_NUM_STAGES] = {};
    long long bucket_n[SDN_NUM_STAGES] = {};
    if (dp)
        for (int s = 0; s < SDN_NUM_STAGES; ++s) SDN_OK(stage_bucket(c, s, &bucket[s], &bucket_n[s]));
    const bool overlap = dp && !(flags & SDN_STEP_NO_OVERLAP) && !c->prof;
    // The loss normaliser is the GLOBAL valid count.  It depends only on the batch, so its all-reduce (8 bytes, but a
    // rendezvous of all ranks) runs on the communicator stream UNDER the forward; the main stream picks it up right
    // before the loss kernel.  On the main stream it would make every rank wait for the slowest one after each forward.
    const bool count_early = overlap && (flags & SDN_STEP_HAVE_COUNT);
    if (count_early) {
        CUDA_OK(cudaEventRecord(c->ev_bucket, st));
        CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_bucket, 0));
        NCCL_OK(g_nccl.AllReduce(n_norm, n_norm, 1, ncclUint64, ncclSum, c->comm, c->comm_stream));
        CUDA_OK(cudaEventRecord(c->ev_comm, c->comm_stream));
    }
    SDN_OK(forward_impl(c, x, nullptr, nullptr, B, 1, params_dirty, st));
    if (!(flags & SDN_STEP_HAVE_COUNT)) SDN_OK(sdn_count_valid(c, target, mask, B, n_norm, stream));
    if (count_early) {
        CUDA_OK(cudaStreamWaitEvent(st, c->ev_comm, 0));
    } else if (dp) {
        ProfScope ps(c, st, "allreduce_count", 0, 0.0, 8.0);
        NCCL_OK(g_nccl.AllReduce(n_norm, n_norm, 1, ncclUint64, ncclSum, c->comm, st));
    }
    SDN_OK(sdn_loss_begin(c, target, mask, nullptr, nullptr, sums4, count, n_norm, 1, 0, stream));
    for (int s = 0; s < SDN_NUM_STAGES; ++s) {
        SDN_OK(sdn_backward_stage(c, s, stream));
        if (!dp) continue;
        cudaStream_t cs = st;
        if (overlap) {
            CUDA_OK(cudaEventRecord(c->ev_bucket, st));
            CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_bucket, 0));
            cs = c->comm_stream;
        }
        ProfScope ps(c, cs, "allreduce_grads", s, 0.0, (double)bucket_n[s] * 4.0);
        NCCL_OK(g_nccl.AllReduce(bucket[s], bucket[s], (size_t)bucket_n[s], ncclFloat32, ncclSum, c->comm, cs));
    }
    if (overlap) {
        CUDA_OK(cudaEventRecord(c->ev_comm, c->comm_stream));
        CUDA_OK(cudaStreamWaitEvent(st, c->ev_comm, 0));
    }
    return 0;
}

// Validation / preview forward (run_epoch with optimizer=None, train.py:618; log_epoch_previews, train.py:268-272):
// eval-mode forward + the same five metric sums, no gradient.  disp / logvar are optional outputs.
int sdn_eval_step(sdn_ctx* c, const float* x, const float* target, const uint8_t* mask, int B, int params_dirty,
                  float* disp, float* logvar, double* sums4, unsigned long long* count, void* stream) {
    if (c == nullptr || x == nullptr || target == nullptr || mask == nullptr || sums4 == nullptr || count == nullptr)
        return fail("sdn_eval_step: NULL argument");
    SDN_ENTER(c);
    SDN_OK(forward_impl(c, x, nullptr, nullptr, B, 0, params_dirty, (cudaStream_t)stream));
    return sdn_loss_begin(c, target, mask, disp, logvar, sums4, count, nullptr, 0, 0, stream);
}

int sdn_preprocess(sdn_ctx* c, const uint8_t* left, const uint8_t* right, const uint8_t* disparity, int B, int Hs,
                   int Ws, const sdn_aug_params* aug_dev, float* input, float* target, uint8_t* mask,
                   unsigned long long* valid_count, unsigned flags, void* stream) {
    if (c == nullptr || left == nullptr || right == nullptr || disparity == nullptr || input == nullptr ||
        target == nullptr || mask == nullptr)
        return fail("sdn_preprocess: NULL argument");
    if (B < 1 || B > c->maxB) return fail("sdn_preprocess: batch %d outside [1, %d]", B, c->maxB);
    if (Hs < 1 || Ws < 1) return fail("sdn_preprocess: bad source size %dx%d", Hs, Ws);
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    c->pdl = t_pdl = (long long)B * c->H * c->W <= 48LL * 240 * 320 ? 1 : 0;
    if (valid_count != nullptr) SDN_OK(zero_fill(c, valid_count, sizeof(unsigned long long), st));
    const AugParams* aug_all = reinterpret_cast<const AugParams*>(aug_dev);
    if (aug_all != nullptr && (flags & SDN_PREPROCESS_AUG_HOST)) {
        const int words = 2 * B * (int)(sizeof(AugParams) / 4);
        launch_k(copy_f32_kernel, (words + 255) / 256, 256, 0, st, reinterpret_cast<const float*>(aug_dev), c->aug_stage, words, 0);
        ++c->launches;
        aug_all = reinterpret_cast<const AugParams*>(c->aug_stage);
    }
    // SDN_PRE_CHUNK=n sends the samples through in chunks whose fp32 views stay L2-resident between the
    // decode/resize pass and the in-place augmentation pass.  It paid while the augmentation was ALU-bound
    // at ~1100 instructions per pixel; with the special-function-unit version the extra launches cost more
    // than the DRAM round trip (chunk 16: 1.94 ms, whole batch: 1.51 ms at 256 pairs), so the default is off.
    static const int chunk_cfg = env_int("SDN_PRE_CHUNK", 0);
    const int chunk = (aug_all != nullptr && chunk_cfg > 0) ? chunk_cfg : B;
    const size_t src_img = (size_t)Hs * Ws * 3, plane = (size_t)c->H * c->W;
    const int parts = ((c->W + 127) / 128) * ((c->H + PRE_ROWS - 1) / PRE_ROWS);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        SDN_OK(preprocess_chunk(c, left + b0 * src_img, right + b0 * src_img, disparity + b0 * src_img, nb, Hs, Ws,
                                aug_all ? aug_all + 2 * b0 : nullptr, input + (size_t)b0 * 6 * plane,
                                target + (size_t)b0 * plane, mask + (size_t)b0 * plane, valid_count, flags,
                                c->gray_part + (size_t)2 * b0 * parts, c->blur_tmp + (size_t)2 * b0 * 3 * plane, st));
    }
    return 0;
}

// FoundationStereoDataset.__getitem__ on a cache hit (dataset.py:86-106, 272-311): see live.cuh
int sdn_preprocess_cached(sdn_ctx* c, const uint8_t* left, const uint8_t* right, const uint16_t* disparity_f16, int B,
                          const sdn_aug_params* aug_dev, float* input, float* target, uint8_t* mask,
                          unsigned long long* valid_count, unsigned flags, void* stream) {
    if (c == nullptr || left == nullptr || right == nullptr || disparity_f16 == nullptr || input == nullptr ||
        target == nullptr || mask == nullptr)
        return fail("sdn_preprocess_cached: NULL argument");
    if (B < 1 || B > c->maxB) return fail("sdn_preprocess_cached: batch %d outside [1, %d]", B, c->maxB);
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    c->pdl = t_pdl = (long long)B * c->H * c->W <= 48LL * 240 * 320 ? 1 : 0;
    if (valid_count != nullptr) SDN_OK(zero_fill(c, valid_count, sizeof(unsigned long long), st));
    const AugParams* aug = reinterpret_cast<const AugParams*>(aug_dev);
    if (aug != nullptr && (flags & SDN_PREPROCESS_AUG_HOST)) {
        const int words = 2 * B * (int)(sizeof(AugParams) / 4);
        launch_k(copy_f32_kernel, (words + 255) / 256, 256, 0, st, reinterpret_cast<const float*>(aug_dev), c->aug_stage, words, 0);
        ++c->launches;
        aug = reinterpret_cast<const AugParams*>(c->aug_stage);
    }
    const int H = c->H, W = c->W;
    const int parts = ((W + 127) / 128) * ((H + PRE_ROWS - 1) / PRE_ROWS);
    {
        // 2 x u8 HWC + f16 read; fp32 input / target + u8 mask written
        ProfScope ps(c, st, "pre_cached", 0, 0.0, (double)B * H * W * (6 + 2 + 6 * 4 + 4 + 1));
        launch_k(cached_assemble_kernel, dim3(parts, B), 128, 0, st, left, right, reinterpret_cast<const __half*>(disparity_f16), B, H,
                 W, input, target, mask, valid_count, aug, aug ? c->gray_part : nullptr, parts);
        ++c->launches;
    }
    if (aug != nullptr) SDN_OK(run_augment(c, input, B, aug, c->gray_part, c->blur_tmp, parts, st));
    CUDA_OK(cudaGetLastError());
    return 0;
}

// preprocess_rgb x 2 + cat (live_camera/depth_live_dl.py:225-229, 516-520): see live.cuh
int sdn_live_preprocess(sdn_ctx* c, const uint8_t* frame_left_bgr, const uint8_t* frame_right_bgr, int Hs, int Ws,
                        float* input, void* stream) {
    if (c == nullptr || frame_left_bgr == nullptr || frame_right_bgr == nullptr || input == nullptr)
        return fail("sdn_live_preprocess: NULL argument");
    if (Hs < 1 || Ws < 1) return fail("sdn_live_preprocess: bad frame size %dx%d", Hs, Ws);
    SDN_ENTER(c);
    const int H = c->H, W = c->W;
    // resize.cpp: inv_scale = dsize / ssize; scale = 1. / inv_scale (both double)
    const double scale_x = 1.0 / ((double)W / (double)Ws), scale_y = 1.0 / ((double)H / (double)Hs);
    ProfScope ps(c, (cudaStream_t)stream, "live_pre", 0, 0.0, 2.0 * ((double)Hs * Ws * 3 + (double)H * W * 12));
    launch_k(live_preprocess_kernel, dim3((W + 127) / 128, H, 2), 128, 0, (cudaStream_t)stream, frame_left_bgr, frame_right_bgr, Hs,
             Ws, H, W, scale_x, scale_y, input);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// EMA (depth_live_dl.py:531-538), disparity_to_depth (:371-377), confidence_from_logvar (:380-381): see live.cuh
int sdn_live_postprocess(sdn_ctx* c, const float* disp, const float* logvar, int64_t n_pixels, float* ema_state,
                         int ema_has_state, double ema_alpha, double focal_px, double baseline_m, float* disp_out,
                         float* depth_out, float* conf_out, void* stream) {
    if (c == nullptr || disp == nullptr || n_pixels < 1) return fail("sdn_live_postprocess: bad argument");
    if (conf_out != nullptr && logvar == nullptr) return fail("sdn_live_postprocess: confidence needs logvar");
    SDN_ENTER(c);
    const int ema_mode = (ema_state != nullptr && ema_alpha > 0.0) ? (ema_has_state ? 2 : 1) : 0;
    ProfScope ps(c, (cudaStream_t)stream, "live_post", 0, 0.0, (double)n_pixels * 4 * 6);
    launch_k(live_postprocess_kernel, occ_grid(c, live_postprocess_kernel, n_pixels, 256), 256, 0, (cudaStream_t)stream, disp, logvar,
             (long long)n_pixels, ema_state, ema_mode, (float)ema_alpha, (float)(1.0 - ema_alpha),
             (float)(focal_px * baseline_m), disp_out, depth_out, conf_out);
    ++c->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int sdn_adamw_step(sdn_ctx* c, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, int n, double lr, double beta1, double beta2,
                   double eps, double weight_decay, long long* step_dev, const unsigned long long* gate_dev,
                   void* stream) {
    if (c == nullptr || params == nullptr || grads == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr ||
        numel == nullptr || step_dev == nullptr)
        return fail("sdn_adamw_step: NULL argument");
    if (n < 1 || n > 66) return fail("sdn_adamw_step: n = %d outside [1, 66]", n);
    cudaStream_t st = (cudaStream_t)stream;
    SDN_ENTER(c);
    AdamTable t;
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i];
        t.start[i] = (int)total;
        total += numel[i];
    }
    if (total > 0x7fffffffLL) return fail("sdn_adamw_step: too many elements");
    t.start[n] = (int)total;
    t.n = n;
    ProfScope ps(c, st, "adamw", 0, 0.0, (double)total * 28.0);
    launch_k(adamw_step_count_kernel, 1, 1, 0, st, step_dev, gate_dev);
    // hyper-parameters arrive as Python doubles; derived scalars are rounded to fp32 once, like torch's scalars
    launch_k(adamw_all_kernel, c->num_sms * 4, 256, 0, st, t, lr, beta1, beta2, (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                     (float)eps, (float)(1.0 - lr * weight_decay), step_dev, gate_dev);
    c->launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int sdn_debug_trace(sdn_ctx* c, long long* host_out) {
    if (c == nullptr || host_out == nullptr) return fail("sdn_debug_trace: NULL argument");
    SDN_ENTER(c);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(host_out, c->dbg, 3 * 16 * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

int sdn_profile_enable(sdn_ctx* c, int enable) {
    if (c == nullptr) return fail("sdn_profile_enable: NULL ctx");
    c->prof = enable != 0;
    c->recs.clear();
    c->ev_used = 0;
    return 0;
}

// CSV "name,layer,calls,total_ms,flops,bytes" aggregated per (name, layer) since enable; resets the records.
int sdn_profile_dump(sdn_ctx* c, char* buf, int64_t capacity) {
    if (c == nullptr || buf == nullptr || capacity < 64) return fail("sdn_profile_dump: bad argument");
    SDN_ENTER(c);
    CUDA_OK(cudaDeviceSynchronize());
    struct Agg { const char* name; int layer; int calls; double ms, flops, bytes; };
    std::vector<Agg> aggs;
    for (const auto& r : c->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        Agg* hit = nullptr;
        for (auto& a : aggs) if (a.layer == r.layer && strcmp(a.name, r.name) == 0) { hit = &a; break; }
        if (!hit) { aggs.push_back(Agg{r.name, r.layer, 0, 0.0, 0.0, 0.0}); hit = &aggs.back(); }
        hit->calls += 1; hit->ms += ms; hit->flops += r.flops; hit->bytes += r.bytes;
    }
    std::string out = "name,layer,calls,total_ms,flops,bytes\n";
    char line[256];
    for (const auto& a : aggs) {
        snprintf(line, sizeof line, "%s,%d,%d,%.6f,%.6e,%.6e\n", a.name, a.layer, a.calls, a.ms, a.flops, a.bytes);
        out += line;
    }
    c->recs.clear();
    c->ev_used = 0;
    if ((int64_t)out.size() + 1 > capacity) return fail("sdn_profile_dump: buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

int sdn_debug_read(sdn_ctx* c, int which, int kind, float* host_out, int64_t capacity, int* dims4) {
    if (c == nullptr || host_out == nullptr || dims4 == nullptr) return fail("sdn_debug_read: NULL argument");
    const Act* a = nullptr;
    if (which >= 0 && which < 18) {
        ConvL& L = c->conv[which];
        a = kind == 0 ? &L.y : kind == 1 ? &L.a : kind == 2 ? &L.dy : kind == 3 ? &L.ga
            : (kind == 4 && L.pooled_out) ? &L.gp : (kind == 5 && L.pooled_out) ? &L.pool : nullptr;
    } else if (which >= 100 && which < 104) {
        UpL& U = c->up[which - 100];
        a = kind == 0 ? &U.u : kind == 3 ? &U.gu : nullptr;
    } else if (which == 200) {
        a = &c->x0;
    }
    if (a == nullptr || a->p == nullptr) return fail("sdn_debug_read: bad selector %d/%d", which, kind);
    const size_t n = a->elems(c->B);
    dims4[0] = c->B; dims4[1] = a->H; dims4[2] = a->W; dims4[3] = a->C;
    if ((int64_t)n > capacity) return fail("sdn_debug_read: capacity %lld < %zu", (long long)capacity, n);
    SDN_ENTER(c);
    CUDA_OK(cudaDeviceSynchronize());
    std::vector<uint16_t> tmp(n);
    CUDA_OK(cudaMemcpy(tmp.data(), a->p, n * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        uint32_t u = (uint32_t)tmp[i] << 16;
        float f;
        memcpy(&f, &u, 4);
        host_out[i] = f;
    }
    return 0;
}

}  // extern "C"
