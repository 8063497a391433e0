// Memory-bound kernels around the tensor-core convolutions: weight packing,
// first-layer im2col, BatchNorm (train statistics / eval), ReLU, MaxPool, the
// two 1x1 heads with softplus / clamp, the heteroscedastic Laplace loss and the
// backward counterparts.  All activations are NHWC bf16; every kernel moves
// 16-byte vectors (8 channels) per thread and is sized as a grid-stride loop
// over a multiple of the SM count.
//
// Reference semantics (cited per kernel):
//   ConvBlock conv->BN->ReLU x2 : src/foundation_stereo_depth/model.py:32-45
//   MaxPool2d(2)                : model.py:59,83-86
//   heads, softplus, clamp      : model.py:76-77,98,103
//   loss + metric sums          : src/foundation_stereo_depth/train.py:329-357
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdn {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(u[i] << 16);
        f[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 q;
    q.x = pack2(f[0], f[1]);
    q.y = pack2(f[2], f[3]);
    q.z = pack2(f[4], f[5]);
    q.w = pack2(f[6], f[7]);
    return q;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ------------------------------------------------------------------ packing
// mode 0: conv3x3 forward   dst[co][tap*Ci + ci]          = W[co][ci][tap]
// mode 1: conv3x3 dgrad     dst[ci][tap*Co + co]          = W[co][ci][8 - tap]
// mode 2: first layer       dst[co][k], k = tap*Ci + ci<54 = W[co][ci][tap], zero-padded to Kpad
// mode 3: convT forward     dst[(q*Co + co)][ci]          = W[ci][co][q]      (W is [Ci][Co][2][2])
// mode 4: convT dgrad       dst[ci][q*Co + co]            = W[ci][co][q]
// row-halo K order (conv_gemm HALO): k = ((dx*blocks + cb)*3 + dy)*KB + c, channel = cb*KB + c, tap = dy*3 + dx
// mode 5: conv3x3 forward   dst[co][k]                    = W[co][channel][tap]           (Kpad = KB)
// mode 6: conv3x3 dgrad     dst[ci][k]                    = W[channel][ci][8 - tap]       (Kpad = KB)
// box9 K order (conv_gemm HALO = 2): k = (cb*9 + tap)*KB + c
// mode 8: conv3x3 forward   dst[co][k] = W[co][channel][tap];  mode 9: dgrad  dst[ci][k] = W[channel][ci][8 - tap]
// oscale (modes 0, 2, 5, 8 only): per-output-channel factor folded into the packed weights
// (eval-mode BatchNorm: w' = w * gamma / sqrt(running_var + eps)); nullptr = 1.
// Channel counts of every 3x3 / 2x2 layer are powers of two (32..512): the index decoding uses shifts and
// masks (plus divisions by the constants 3 and 9); runtime integer divisions made this kernel ALU-bound.
__device__ __forceinline__ float pack_value(const float* __restrict__ w, int mode, int Co, int Ci, int Kpad,
                                            const float* __restrict__ oscale, int i) {
    float v = 0.f;
    const int lCi = 31 - __clz(Ci), lCo = 31 - __clz(Co);
    if (mode == 0) {
        const int q = i >> lCi, co = q / 9, tap = q - 9 * co, ci = i & (Ci - 1);
        v = w[(co * Ci + ci) * 9 + tap];
        if (oscale != nullptr) v *= oscale[co];
    } else if (mode == 1) {
        const int q = i >> lCo, ci = q / 9, tap = q - 9 * ci, co = i & (Co - 1);
        v = w[(co * Ci + ci) * 9 + (8 - tap)];
    } else if (mode == 2) {
        const int co = i / Kpad, k = i % Kpad;
        if (k < 9 * Ci) {
            const int tap = k / Ci, ci = k % Ci;
            v = w[(co * Ci + ci) * 9 + tap];
            if (oscale != nullptr) v *= oscale[co];
        }
    } else if (mode == 5 || mode == 6 || mode == 8 || mode == 9) {
        const int KB = Kpad, lKB = 31 - __clz(KB);
        const bool fwd = mode == 5 || mode == 8;
        const int lk = fwd ? lCi : lCo;             // channels on the K side (a power of two)
        const int q = i >> lk, row = q / 9;
        const int k = i - ((row * 9) << lk);
        const int c = k & (KB - 1), u = k >> lKB;
        int chan, tap;
        if (mode <= 6) {   // row-halo order: u = (dx*blocks + cb)*3 + dy
            const int dy = u % 3, unit = u / 3, lb = lk - lKB;
            const int dx = unit >> lb, cb = unit & ((1 << lb) - 1);
            chan = (cb << lKB) + c; tap = dy * 3 + dx;
        } else {           // box9 order: u = cb*9 + tap
            const int cb = u / 9;
            tap = u - 9 * cb; chan = (cb << lKB) + c;
        }
        v = fwd ? w[(row * Ci + chan) * 9 + tap] : w[(chan * Ci + row) * 9 + (8 - tap)];
        if (fwd && oscale != nullptr) v *= oscale[row];
    } else if (mode == 10) {
        // first layer as a row-halo 3x1 conv over [.., 32]: dst[co][dy*32 + dx*Ci + c] = W[co][c][dy*3 + dx]
        const int co = i / 96, k = i % 96, dy = k / 32, cc = k % 32;
        if (cc < 3 * Ci) {
            v = w[(co * Ci + cc % Ci) * 9 + dy * 3 + cc / Ci];
            if (oscale != nullptr) v *= oscale[co];
        }
    } else if (mode == 11 || mode == 12) {
        // super-pixel layout (conv_gemm HALO = 3; Cout == 32, 32-channel sources): dst[unit][dy][row 0..127][k 0..63]
        //   rows   0..63 : centre  (output pixel po = row >> 5, channel = row & 31; k: input pixel pi = k >> 5, c = k & 31,
        //                           horizontal tap dx = pi - po + 1)
        //   rows  64..95 : left neighbour, only its second pixel (k >= 32) feeds output pixel 0 (dx = 0)
        //   rows 96..127 : right neighbour, only its first pixel (k < 32) feeds output pixel 1 (dx = 2)
        // mode 11 (forward): value = W[row_ch][unit*32 + c][dy][dx]; mode 12 (data gradient, single 32-channel
        // source): value = W[c][row_ch][2 - dy][2 - dx]   (flipped taps, transposed channels)
        const int k = i & 63, row = (i >> 6) & 127, slab = i >> 13, dy = slab % 3, unit = slab / 3;
        const int pi = k >> 5, c = k & 31;
        int och = row & 31, dx = -1;
        if (row < 64) dx = pi - (row >> 5) + 1;
        else if (row < 96) dx = pi == 1 ? 0 : -1;
        else dx = pi == 0 ? 2 : -1;
        if (dx >= 0) {
            if (mode == 11) {
                v = w[(och * Ci + unit * 32 + c) * 9 + dy * 3 + dx];
                if (oscale != nullptr) v *= oscale[och];
            } else {
                v = w[(c * Ci + och) * 9 + (2 - dy) * 3 + (2 - dx)];
            }
        }
    } else if (mode == 3) {
        const int ci = i & (Ci - 1), row = i >> lCi, q = row >> lCo, co = row & (Co - 1);
        v = w[(ci * Co + co) * 4 + q];
    } else if (mode == 4) {
        const int co = i & (Co - 1), t = i >> lCo, q = t & 3, ci = t >> 2;
        v = w[(ci * Co + co) * 4 + q];
    } else {  // mode 7: bias replicated over the 4 convT quadrants (fp32 destination)
        v = w[i & (Co - 1)];
    }
    return v;
}

// Every layer's packing in ONE launch (the table rides in the kernel parameters).
struct PackEntry {
    const float* w;
    void* dst;
    const float* oscale;
    int mode, Co, Ci, Kpad;
    int start;  // first flat index of this entry
};
struct PackTable {
    PackEntry e[52];
    int n;
    int total;
};
// Source offset, source stride and output-channel row of the run of 8 consecutive destination elements that
// starts at flat index i (a multiple of 8): in every conv3x3 / ConvTranspose2d packing the innermost destination
// index is a channel, so the 8 values are one strided gather and share the index decoding (the per-element
// decode made this kernel ALU-bound: 0.09 ms for 62 MB).  Returns false for the small irregular modes.
__device__ __forceinline__ bool pack_run(int mode, int Co, int Ci, int Kpad, int i, int& src, int& stride, int& orow) {
    const int lCi = 31 - __clz(Ci), lCo = 31 - __clz(Co);
    orow = -1;
    if (mode == 0) {
        const int q = i >> lCi, co = q / 9, tap = q - 9 * co, ci = i & (Ci - 1);
        src = (co * Ci + ci) * 9 + tap; stride = 9; orow = co;
    } else if (mode == 1) {
        const int q = i >> lCo, ci = q / 9, tap = q - 9 * ci, co = i & (Co - 1);
        src = (co * Ci + ci) * 9 + (8 - tap); stride = Ci * 9;
    } else if (mode == 5 || mode == 6 || mode == 8 || mode == 9) {
        const int KB = Kpad, lKB = 31 - __clz(KB);
        const bool fwd = mode == 5 || mode == 8;
        const int lk = fwd ? lCi : lCo;
        const int q = i >> lk, row = q / 9;
        const int k = i - ((row * 9) << lk);
        const int c = k & (KB - 1), u = k >> lKB;
        int chan, tap;
        if (mode <= 6) {
            const int dy = u % 3, unit = u / 3, lb = lk - lKB;
            const int dx = unit >> lb, cb = unit & ((1 << lb) - 1);
            chan = (cb << lKB) + c; tap = dy * 3 + dx;
        } else {
            const int cb = u / 9;
            tap = u - 9 * cb; chan = (cb << lKB) + c;
        }
        if (fwd) { src = (row * Ci + chan) * 9 + tap; stride = 9; orow = row; }
        else { src = (chan * Ci + row) * 9 + (8 - tap); stride = Ci * 9; }
    } else if (mode == 3) {
        const int ci = i & (Ci - 1), row = i >> lCi, q = row >> lCo, co = row & (Co - 1);
        src = (ci * Co + co) * 4 + q; stride = Co * 4;
    } else if (mode == 4) {
        const int co = i & (Co - 1), t = i >> lCo, q = t & 3, ci = t >> 2;
        src = (ci * Co + co) * 4 + q; stride = 4;
    } else {
        return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) pack_all_kernel(const __grid_constant__ PackTable t) {
    SDN_PDL_ENTRY();
    __shared__ int starts[53];
    if (threadIdx.x <= t.n) starts[threadIdx.x] = threadIdx.x < t.n ? t.e[threadIdx.x].start : t.total;
    __syncthreads();
    // one thread = 8 consecutive destination elements (every entry's start and length are multiples of 8)
    const int runs = t.total >> 3;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < runs; r += gridDim.x * blockDim.x) {
        const int idx = r << 3;
        int lo = 0, hi = t.n - 1;
        while (lo < hi) {  // last entry with start <= idx
            const int mid = (lo + hi + 1) >> 1;
            if (starts[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        const PackEntry& e = t.e[lo];
        const int i0 = idx - e.start;
        float v[8];
        int src, stride, orow;
        if (pack_run(e.mode, e.Co, e.Ci, e.Kpad, i0, src, stride, orow)) {
            const float sc = (orow >= 0 && e.oscale != nullptr) ? e.oscale[orow] : 1.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(e.w + src + j * stride) * sc;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = pack_value(e.w, e.mode, e.Co, e.Ci, e.Kpad, e.oscale, i0 + j);
        }
        if (e.mode == 7) {
            float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.dst) + i0);
            d[0] = make_float4(v[0], v[1], v[2], v[3]);
            d[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.dst) + i0) = pack8(v);
        }
    }
}

// bias replicated over the 4 quadrants of a ConvTranspose2d GEMM: dst[q*Co + co] = b[co]
__global__ void tile_bias_kernel(const float* __restrict__ b, float* __restrict__ dst, int Co, int reps) {
    SDN_PDL_ENTRY();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Co * reps; i += gridDim.x * blockDim.x) dst[i] = b[i % Co];
}

// ------------------------------------------------------- first-layer im2col
// x fp32 NCHW [B,Cin,H,W] (the reference sample layout, dataset.py:305-311) ->
// bf16 [B,H,W,64] with k = (dy*3+dx)*Cin + c, zero padding at the border and
// zeros for k >= 9*Cin.  One block = one output row segment of IM2COL_PX pixels:
// the 3 x Cin input row segments are staged in shared memory with coalesced
// loads, then every thread emits 16-byte chunks (8 consecutive k of one pixel),
// so 8 consecutive threads write one pixel's 128 contiguous bytes.
constexpr int IM2COL_PX = 160;   // pixels per block in x
constexpr int IM2COL_ROWS = 4;   // output rows per block (they share the staged input rows)
constexpr int IM2COL_MAXC = 8;

template <int Cin>
__global__ void __launch_bounds__(256) im2col_first_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B,
                                                           int H, int W) {
    SDN_PDL_ENTRY();
    // +1: one always-zero cell that the k >= 9*Cin padding lanes read (branch-free gather)
    __shared__ float tile[Cin * (IM2COL_ROWS + 2) * (IM2COL_PX + 2) + 1];
    static_assert(Cin <= IM2COL_MAXC, "first-layer channel count");
    constexpr int K = 9 * Cin;
    constexpr int TR = IM2COL_ROWS + 2;
    constexpr int PITCH = IM2COL_PX + 2;
    constexpr int ZERO = Cin * TR * PITCH;
    const int xblocks = (W + IM2COL_PX - 1) / IM2COL_PX;
    const int yblocks = (H + IM2COL_ROWS - 1) / IM2COL_ROWS;
    const int xb = blockIdx.x % xblocks;
    const int yb = (blockIdx.x / xblocks) % yblocks;
    const int n = blockIdx.x / (xblocks * yblocks);
    const int x0 = xb * IM2COL_PX, y0 = yb * IM2COL_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // stage: one warp per (channel, row) line, lanes stride over the columns (no div / mod per element)
    for (int row = warp; row < Cin * TR; row += 8) {
        const int c = row / TR, r = row % TR;   // compile-time divisors
        const int yy = y0 + r - 1;
        const bool row_ok = yy >= 0 && yy < H;
        const float* src = x + (((size_t)n * Cin + c) * H + (row_ok ? yy : 0)) * W;
#pragma unroll
        for (int t = 0; t < (PITCH + 31) / 32; ++t) {
            const int col = lane + 32 * t;
            const int xx = x0 + col - 1;
            if (col < PITCH) tile[row * PITCH + col] = (row_ok && xx >= 0 && xx < W) ? __ldg(src + xx) : 0.f;
        }
    }
    if (threadIdx.x == 0) tile[ZERO] = 0.f;
    __syncthreads();
    // gather: a thread always emits the same 16-byte chunk g (8 consecutive k) of its pixels, so the eight
    // shared-memory offsets of that chunk are loop invariants
    const int g = threadIdx.x & 7;
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = g * 8 + j;
        const int tap = k / Cin, c = k % Cin;
        off[j] = k < K ? (c * TR + tap / 3) * PITCH + tap % 3 : -1;
    }
    const int npx = min(IM2COL_PX, W - x0);
    const int nrows = min(IM2COL_ROWS, H - y0);
    for (int ry = 0; ry < nrows; ++ry) {
        bf16* orow = out + (((size_t)n * H + y0 + ry) * W + x0) * 64 + g * 8;
        for (int px = threadIdx.x >> 3; px < npx; px += 32) {
            const int base = ry * PITCH + px;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = tile[off[j] >= 0 ? off[j] + base : ZERO];
            *reinterpret_cast<uint4*>(orow + (size_t)px * 64) = pack8(f);
        }
    }
}

// Horizontal-taps-only variant: bf16 [B,H,W,32] with k' = dx*Cin + c (dx = 0,1,2 <-> x-1, x, x+1), zeros for
// k' >= 3*Cin.  The first conv then runs as a row-halo 3x1 conv over this 32-channel tensor (the three
// vertical taps are row shifts of one TMA box), which halves the bytes of the im2col'ed input: 64 instead of
// 128 bytes per pixel written here and read back by the first layer's fprop and wgrad.
template <int Cin>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B,
                                                          int H, int W) {
    SDN_PDL_ENTRY();
    __shared__ float tile[Cin * IM2COL_ROWS * (IM2COL_PX + 2) + 1];
    static_assert(3 * Cin <= 32, "first-layer channel count");
    constexpr int PITCH = IM2COL_PX + 2;
    constexpr int ZERO = Cin * IM2COL_ROWS * PITCH;
    const int xblocks = (W + IM2COL_PX - 1) / IM2COL_PX;
    const int yblocks = (H + IM2COL_ROWS - 1) / IM2COL_ROWS;
    const int xb = blockIdx.x % xblocks;
    const int yb = (blockIdx.x / xblocks) % yblocks;
    const int n = blockIdx.x / (xblocks * yblocks);
    const int x0 = xb * IM2COL_PX, y0 = yb * IM2COL_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < Cin * IM2COL_ROWS; row += 8) {
        const int c = row / IM2COL_ROWS, r = row % IM2COL_ROWS;
        const int yy = y0 + r;
        const bool row_ok = yy < H;
        const float* src = x + (((size_t)n * Cin + c) * H + (row_ok ? yy : 0)) * W;
#pragma unroll
        for (int t = 0; t < (PITCH + 31) / 32; ++t) {
            const int col = lane + 32 * t;
            const int xx = x0 + col - 1;
            if (col < PITCH) tile[row * PITCH + col] = (row_ok && xx >= 0 && xx < W) ? __ldg(src + xx) : 0.f;
        }
    }
    if (threadIdx.x == 0) tile[ZERO] = 0.f;
    __syncthreads();
    const int g = threadIdx.x & 3;   // 16-byte chunk (8 consecutive k') of the pixel's 64 bytes
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = g * 8 + j;
        off[j] = k < 3 * Cin ? ((k % Cin) * IM2COL_ROWS) * PITCH + k / Cin : -1;
    }
    const int npx = min(IM2COL_PX, W - x0);
    const int nrows = min(IM2COL_ROWS, H - y0);
    for (int ry = 0; ry < nrows; ++ry) {
        bf16* orow = out + (((size_t)n * H + y0 + ry) * W + x0) * 32 + g * 8;
        for (int px = threadIdx.x >> 2; px < npx; px += 64) {
            const int base = ry * PITCH + px;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = tile[off[j] >= 0 ? off[j] + base : ZERO];
            *reinterpret_cast<uint4*>(orow + (size_t)px * 32) = pack8(f);
        }
    }
}

// -------------------------------------------------------------- BatchNorm
// Training statistics (nn.BatchNorm2d in train mode, model.py:37,40): biased
// variance for normalisation, unbiased into running_var, momentum 0.1,
// num_batches_tracked += 1.  partials = per-CTA (sum, sumsq) rows written by the
// conv epilogue; summed here in fp64 in a fixed order (deterministic).
__global__ void bn_finalize_train_kernel(const float* __restrict__ partials, int nparts, int C, double count,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         long long* __restrict__ num_batches, float eps, float momentum,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ mean_out, float* __restrict__ rstd_out, int fold) {
    SDN_PDL_ENTRY();
    // one warp per channel: lanes stride over the partial rows, fixed-shape butterfly -> deterministic
    // fold == 2: the producing kernel ran in the super-pixel view (conv_gemm HALO = 3): partial rows hold 2*C
    // "channels" and real channel c is the sum of columns c and c + C
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c == 0 && lane == 0 && num_batches != nullptr) *num_batches += 1;
    if (c >= C) return;
    const int Ck = C * fold;
    double s = 0.0, q = 0.0;
    // eight rows per lane in flight: this 1-block-per-8-channels kernel sits between a conv and its BatchNorm pass and
    // is pure L2 latency (one row per round trip cost ~8 us per launch, 36 launches per step: 5 % of a 32-pair step)
    for (int base = lane; base < nparts; base += 32 * 8) {
        float vs[8], vq[8], ws[8], wq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = base + 32 * k;
            const bool ok = i < nparts;
            const float* row = partials + (size_t)(ok ? i : 0) * 2 * Ck;
            vs[k] = ok ? __ldg(row + c) : 0.f; vq[k] = ok ? __ldg(row + Ck + c) : 0.f;
            ws[k] = (ok && fold == 2) ? __ldg(row + C + c) : 0.f; wq[k] = (ok && fold == 2) ? __ldg(row + Ck + C + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { s += (double)vs[k]; q += (double)vq[k]; s += (double)ws[k]; q += (double)wq[k]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane != 0) return;
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    mean_out[c] = (float)mean;
    rstd_out[c] = rstd;
    if (C == 32) {   // duplicates for kernels that see this layer as 64 super-pixel channels
        scale[c + 32] = sc; shift[c + 32] = shift[c]; mean_out[c + 32] = (float)mean; rstd_out[c + 32] = rstd;
    }
    if (running_mean != nullptr) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// Eval mode: normalise with the running statistics.
__global__ void bn_prepare_eval_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       float eps, float* __restrict__ scale, float* __restrict__ shift) {
    SDN_PDL_ENTRY();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float rstd = 1.f / sqrtf(running_var[c] + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - running_mean[c] * sc;
    if (C == 32) { scale[c + 32] = sc; shift[c + 32] = shift[c]; }   // super-pixel view (conv_gemm HALO = 3)
}

// Walks the 2x2 quads q0, q0 + qstep, ... of a [B, H/2, W/2] quad grid WITHOUT per-item divisions (the 64-bit
// div / mod chain of the index decode was ~450 of the ~750 instructions per item of the pooled kernels).
struct QuadWalker {
    int x2, y2, n, dx, dy, dn, W2, H2;
    __device__ __forceinline__ void init(long long q0, long long qstep, int W2_, int H2_) {
        W2 = W2_; H2 = H2_;
        x2 = int(q0 % W2); long long m = q0 / W2; y2 = int(m % H2); n = int(m / H2);
        dx = int(qstep % W2); m = qstep / W2; dy = int(m % H2); dn = int(m / H2);
    }
    __device__ __forceinline__ void next() {
        x2 += dx; int c = x2 >= W2 ? 1 : 0; x2 -= c * W2;
        y2 += dy + c; c = y2 >= H2 ? 1 : 0; y2 -= c * H2;
        n += dn + c;
    }
};

// a = relu(y*scale + shift), optionally also p = maxpool2x2(a).
// POOL: one thread = a 2x2 pixel quad x 8 channels.  Otherwise one pixel x 8 channels.
template <bool POOL>
__global__ void bn_relu_pool_kernel(const bf16* __restrict__ y, const float* __restrict__ scale,
                                    const float* __restrict__ shift, bf16* __restrict__ a, bf16* __restrict__ pooled,
                                    unsigned short* __restrict__ amax, int B, int H, int W, int C) {
    SDN_PDL_ENTRY();
    const int CG = C >> 3;
    // the grid stride (gridDim.x * 256) is a multiple of CG (<= 64), so a thread keeps ONE channel group: its
    // scale / shift live in registers.  Re-loading them per item cost 16 scalar loads (8 cache lines each for
    // wide layers) per 16 bytes of payload and made the 128..512-channel layers LSU-bound.
    const int cg = int(((long long)blockIdx.x * blockDim.x + threadIdx.x) % CG);
    float sc[8], sh[8];
    {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8) + 1);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8)), h1 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8) + 1);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
    }
    if (POOL) {
        const int H2 = H >> 1, W2 = W >> 1;
        const long long total = (long long)B * H2 * W2 * CG;
        const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
        QuadWalker qw;
        qw.init(tid0 / CG, stride / CG, W2, H2);     // (the stride is a multiple of CG: this thread keeps its channel group)
        for (long long i = tid0; i < total; i += stride, qw.next()) {
            const int x2 = qw.x2, y2 = qw.y2, n = qw.n;
            float mx[8];
            unsigned am = 0;   // 2 bits per channel: quad position of the FIRST maximum (nn.MaxPool2d routing)
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[j] = 0.f;  // relu output >= 0
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const long long pix = ((long long)n * H + (2 * y2 + (d >> 1))) * W + (2 * x2 + (d & 1));
                float f[8];
                unpack8(ldg16(y + pix * C + cg * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                }
                const uint4 o = pack8(f);
                *reinterpret_cast<uint4*>(a + pix * C + cg * 8) = o;
                float g[8];
                unpack8(o, g);  // pool the bf16-rounded values (what the consumers will see)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (d > 0 && g[j] > mx[j]) { mx[j] = g[j]; am = (am & ~(3u << (2 * j))) | (unsigned(d) << (2 * j)); }
                    else if (d == 0) mx[j] = g[j];
            }
            *reinterpret_cast<uint4*>(pooled + (((long long)n * H2 + y2) * W2 + x2) * C + cg * 8) = pack8(mx);
            if (amax != nullptr) amax[i] = (unsigned short)am;   // the backward pass routes by it instead of recomputing
        }
    } else {
        const long long total = (long long)B * H * W * CG;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += (long long)gridDim.x * blockDim.x) {
            float f[8];
            unpack8(ldg16(y + i * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            *reinterpret_cast<uint4*>(a + i * 8) = pack8(f);
        }
    }
}

// 2x2 max pool of an NHWC bf16 tensor (eval path: BN+ReLU already applied by the conv epilogue).
__global__ void maxpool2x2_kernel(const bf16* __restrict__ a, bf16* __restrict__ pooled, int B, int H, int W, int C) {
    SDN_PDL_ENTRY();
    const int CG = C >> 3, H2 = H >> 1, W2 = W >> 1;
    const long long total = (long long)B * H2 * W2 * CG;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cg = int(i % CG);
        const long long qd = i / CG;
        const int x2 = int(qd % W2);
        const int y2 = int((qd / W2) % H2);
        const int n = int(qd / ((long long)W2 * H2));
        const long long p00 = ((long long)n * H + 2 * y2) * W + 2 * x2;
        float m[8], f[8];
        unpack8(ldg16(a + p00 * C + cg * 8), m);
        const long long nb[3] = {p00 + 1, p00 + W, p00 + W + 1};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            unpack8(ldg16(a + nb[d] * C + cg * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
        }
        *reinterpret_cast<uint4*>(pooled + (((long long)n * H2 + y2) * W2 + x2) * C + cg * 8) = pack8(m);
    }
}

// ---------------------------------------------------- BatchNorm+ReLU backward
// Inputs: y (pre-BN conv output), g (gradient w.r.t. the block's ReLU output,
// full resolution) and optionally gp (gradient w.r.t. the 2x2-max-pooled
// output, half resolution) which is routed to the first maximum of each quad in
// row-major order (nn.MaxPool2d backward, model.py:59).  dz = relu'(z) * dA.
// Pass 1 (reduce): per-channel sum(dz), sum(dz * xhat)  -> per-block partials.
// Pass 2 (apply) : dy = scale * (dz - c1 - xhat * c2)  with c1 = sum(dz)/M,
//                  c2 = sum(dz*xhat)/M (training mode batch-norm backward).
// One work item = 8 channels of one pixel (or of one 2x2 quad when POOL).  APPLY = false:
// accumulate the two per-channel sums; APPLY = true: write dy.  The quad variant keeps only
// the raw bf16 y vectors and the running arg-max in registers (two passes over the quad).
// Packed fp32 pairs (FFMA2 / FADD2 on sm_100): these kernels are ISSUE-bound, not bandwidth-bound (the pooled
// variants ran ~14 instructions per element at 70-79 % of HBM), so channel pairs (2k, 2k+1) travel as float2.
__device__ __forceinline__ void unpack8_2(const uint4& q, float2 (&f)[4]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
}
__device__ __forceinline__ uint4 pack8_2(const float2 (&f)[4]) {
    uint4 q;
    q.x = pack2(f[0].x, f[0].y); q.y = pack2(f[1].x, f[1].y); q.z = pack2(f[2].x, f[2].y); q.w = pack2(f[3].x, f[3].y);
    return q;
}
// per-thread coefficients of one 8-channel group: z = y*sc + sh (ReLU gate), ym = y + nmu (= y - mean);
// apply:  dy = sc*dz + kb*ym + kc  with  kb = -sc*c2*rstd, kc = -sc*c1   (== sc * (dz - c1 - xhat*c2))
struct BnBwdCoef {
    float2 sc[4], sh[4], nmu[4], kb[4], kc[4];
};
template <bool APPLY>
__device__ __forceinline__ void bn_bwd_coef_load(BnBwdCoef& k, int cg, const float* __restrict__ scale,
                                                 const float* __restrict__ shift, const float* __restrict__ mean,
                                                 const float* __restrict__ rstd, const float* __restrict__ c1,
                                                 const float* __restrict__ c2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = cg * 8 + 2 * j;
        k.sc[j] = make_float2(scale[c], scale[c + 1]);
        k.sh[j] = make_float2(shift[c], shift[c + 1]);
        k.nmu[j] = make_float2(-mean[c], -mean[c + 1]);
        if (APPLY) {
            k.kb[j] = make_float2(-k.sc[j].x * c2[c] * rstd[c], -k.sc[j].y * c2[c + 1] * rstd[c + 1]);
            k.kc[j] = make_float2(-k.sc[j].x * c1[c], -k.sc[j].y * c1[c + 1]);
        }
    }
}
// one pixel x 8 channels: yq / gq raw bf16 vectors, `route` = pooled gradient to add where the arg-max mask says so
template <bool APPLY>
__device__ __forceinline__ uint4 bn_bwd_px(const BnBwdCoef& k, const uint4& yq, const uint4& gq, float2 (&s1)[4],
                                           float2 (&s2)[4]) {
    float2 y2[4], g2[4], o[4];
    unpack8_2(yq, y2);
    unpack8_2(gq, g2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 z = __ffma2_rn(y2[j], k.sc[j], k.sh[j]);
        const float2 dz = make_float2(z.x > 0.f ? g2[j].x : 0.f, z.y > 0.f ? g2[j].y : 0.f);
        const float2 ym = __fadd2_rn(y2[j], k.nmu[j]);
        if (APPLY) o[j] = __ffma2_rn(k.kb[j], ym, __ffma2_rn(k.sc[j], dz, k.kc[j]));
        else { s1[j] = __fadd2_rn(s1[j], dz); s2[j] = __ffma2_rn(dz, ym, s2[j]); }
    }
    return APPLY ? pack8_2(o) : uint4{0u, 0u, 0u, 0u};
}
// One work item = 8 channels of one pixel (or of one 2x2 quad when POOL).  APPLY = false: accumulate
// s1 = sum dz and s2 = sum dz * (y - mean) (the finalize kernel multiplies s2 by rstd); APPLY = true: write dy.
template <bool POOL, bool APPLY>
__device__ __forceinline__ void bn_bwd_item(const bf16* __restrict__ y, const bf16* __restrict__ g,
                                            const bf16* __restrict__ gp, const unsigned short* __restrict__ amax,
                                            const BnBwdCoef& k, long long i, int cg, const QuadWalker& qw, int H, int W,
                                            int C, float2 (&s1)[4], float2 (&s2)[4], bf16* __restrict__ dy) {
    if (POOL) {
        const int H2 = H >> 1, W2 = W >> 1;
        const int x2 = qw.x2, y2 = qw.y2, n = qw.n;
        const long long p00 = ((long long)n * H + 2 * y2) * W + 2 * x2;
        const long long pix[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
        // the forward stored which quad position won each channel's max (first maximum, bf16 values)
        const unsigned am = __ldg(amax + i);
        uint4 yr[4], gr[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            yr[d] = ldg16(y + pix[d] * C + cg * 8);
            gr[d] = ldg16(g + pix[d] * C + cg * 8);
        }
        float2 gp2[4];
        unpack8_2(ldg16(gp + (((long long)n * H2 + y2) * W2 + x2) * C + cg * 8), gp2);
        unsigned amj[8];     // each channel's 2-bit field, left in place: compared with d << (2 * j)
#pragma unroll
        for (int j = 0; j < 8; ++j) amj[j] = am & (3u << (2 * j));
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            // route the pooled gradient: dA = g + (arg-max == d ? gp : 0), added in fp32 like autograd's accumulation
            float2 g2[4];
            unpack8_2(gr[d], g2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (amj[2 * j] == (unsigned(d) << (4 * j))) g2[j].x += gp2[j].x;
                if (amj[2 * j + 1] == (unsigned(d) << (4 * j + 2))) g2[j].y += gp2[j].y;
            }
            float2 y2v[4], o[4];
            unpack8_2(yr[d], y2v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 z = __ffma2_rn(y2v[j], k.sc[j], k.sh[j]);
                const float2 dz = make_float2(z.x > 0.f ? g2[j].x : 0.f, z.y > 0.f ? g2[j].y : 0.f);
                const float2 ym = __fadd2_rn(y2v[j], k.nmu[j]);
                if (APPLY) o[j] = __ffma2_rn(k.kb[j], ym, __ffma2_rn(k.sc[j], dz, k.kc[j]));
                else { s1[j] = __fadd2_rn(s1[j], dz); s2[j] = __ffma2_rn(dz, ym, s2[j]); }
            }
            if (APPLY) *reinterpret_cast<uint4*>(dy + pix[d] * C + cg * 8) = pack8_2(o);
        }
    } else {
        const uint4 o = bn_bwd_px<APPLY>(k, ldg16(y + i * 8), ldg16(g + i * 8), s1, s2);
        if (APPLY) *reinterpret_cast<uint4*>(dy + i * 8) = o;
    }
}

template <bool POOL>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) bn_bwd_reduce_kernel(const bf16* __restrict__ y, const bf16* __restrict__ g,
                                                            const bf16* __restrict__ gp,
                                                            const unsigned short* __restrict__ amax,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            float* __restrict__ partials, int B, int H, int W, int C) {
    SDN_PDL_ENTRY();
    const int CG = C >> 3;
    const long long total = POOL ? (long long)B * (H >> 1) * (W >> 1) * CG : (long long)B * H * W * CG;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = int(tid0 % CG);  // fixed per thread: the stride is a multiple of CG
    BnBwdCoef k;
    bn_bwd_coef_load<false>(k, cg, scale, shift, mean, rstd, nullptr, nullptr);
    float2 s1[4], s2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s1[j] = make_float2(0.f, 0.f); s2[j] = make_float2(0.f, 0.f); }
    const long long stride = (long long)gridDim.x * blockDim.x;
    QuadWalker qw;
    if (POOL) qw.init(tid0 / CG, stride / CG, W >> 1, H >> 1);
    for (long long i = tid0; i < total; i += stride) {
        bn_bwd_item<POOL, false>(y, g, gp, amax, k, i, cg, qw, H, W, C, s1, s2, nullptr);
        if (POOL) qw.next();
    }
    __shared__ float red[256][17];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[threadIdx.x][2 * j] = s1[j].x; red[threadIdx.x][2 * j + 1] = s1[j].y;
        red[threadIdx.x][8 + 2 * j] = s2[j].x; red[threadIdx.x][8 + 2 * j + 1] = s2[j].y;
    }
    __syncthreads();
    // thread t < 2*C reduces one (sum kind, channel) over the threads that share its channel group
    for (int o = threadIdx.x; o < 2 * C; o += blockDim.x) {
        const int kind = o / C, c = o % C, g8 = c >> 3, j = c & 7;
        float acc = 0.f;
        for (int t = g8; t < (int)blockDim.x; t += CG) acc += red[t][kind * 8 + j];
        partials[(size_t)blockIdx.x * 2 * C + o] = acc;
    }
}

// partials -> c1, c2 and the affine parameter gradients (dgamma = sum(dz*xhat), dbeta = sum(dz)).
// rstd_scale (nullable): the second partial column holds sum(dz * (y - mean)) instead of sum(dz * xhat) (the
// fused conv-epilogue reduction, CG_BSTATS): multiply by rstd[c] here.
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count,
                                       float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate, const float* __restrict__ rstd_scale,
                                       int fold) {
    SDN_PDL_ENTRY();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per channel
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    const int Ck = C * fold;     // fold == 2: super-pixel partial rows (see bn_finalize_train_kernel)
    double s1 = 0.0, s2 = 0.0;
    for (int base = lane; base < nparts; base += 32 * 8) {      // eight rows per lane in flight (see bn_finalize_train_kernel)
        float va[8], vb[8], wa[8], wb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = base + 32 * k;
            const bool ok = i < nparts;
            const float* row = partials + (size_t)(ok ? i : 0) * 2 * Ck;
            va[k] = ok ? __ldg(row + c) : 0.f; vb[k] = ok ? __ldg(row + Ck + c) : 0.f;
            wa[k] = (ok && fold == 2) ? __ldg(row + C + c) : 0.f; wb[k] = (ok && fold == 2) ? __ldg(row + Ck + C + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { s1 += (double)va[k]; s2 += (double)vb[k]; s1 += (double)wa[k]; s2 += (double)wb[k]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane != 0) return;
    if (rstd_scale != nullptr) s2 *= (double)rstd_scale[c];
    c1[c] = (float)(s1 / count);
    c2[c] = (float)(s2 / count);
    if (dgamma != nullptr) {
        if (accumulate) { dgamma[c] += (float)s2; dbeta[c] += (float)s1; }
        else { dgamma[c] = (float)s2; dbeta[c] = (float)s1; }
    }
}

template <bool POOL>
__global__ void __launch_bounds__(256, POOL ? 2 : 3) bn_bwd_apply_kernel(const bf16* __restrict__ y, const bf16* __restrict__ g,
                                                           const bf16* __restrict__ gp,
                                                           const unsigned short* __restrict__ amax,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const float* __restrict__ c1, const float* __restrict__ c2,
                                                           bf16* __restrict__ dy, int B, int H, int W, int C) {
    SDN_PDL_ENTRY();
    const int CG = C >> 3;
    const long long total = POOL ? (long long)B * (H >> 1) * (W >> 1) * CG : (long long)B * H * W * CG;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int cg = int(tid0 % CG);
    BnBwdCoef k;
    bn_bwd_coef_load<true>(k, cg, scale, shift, mean, rstd, c1, c2);
    float2 s1[4], s2[4];   // (unused by the apply pass)
    if (POOL) {
        QuadWalker qw;
        qw.init(tid0 / CG, stride / CG, W >> 1, H >> 1);
        for (long long i = tid0; i < total; i += stride, qw.next())
            bn_bwd_item<true, true>(y, g, gp, amax, k, i, cg, qw, H, W, C, s1, s2, dy);
    } else {
        // two items per iteration, all four 16-byte loads issued before the first use: at three blocks per SM one
        // item per thread kept only 24 KB per SM in flight, below what the HBM latency needs
        long long i = tid0;
        for (; i + stride < total; i += 2 * stride) {
            const uint4 y0 = ldg16(y + i * 8), g0 = ldg16(g + i * 8);
            const uint4 y1 = ldg16(y + (i + stride) * 8), g1 = ldg16(g + (i + stride) * 8);
            *reinterpret_cast<uint4*>(dy + i * 8) = bn_bwd_px<true>(k, y0, g0, s1, s2);
            *reinterpret_cast<uint4*>(dy + (i + stride) * 8) = bn_bwd_px<true>(k, y1, g1, s1, s2);
        }
        if (i < total) *reinterpret_cast<uint4*>(dy + i * 8) = bn_bwd_px<true>(k, ldg16(y + i * 8), ldg16(g + i * 8), s1, s2);
    }
}

// ConvTranspose2d bias gradient from the per-CTA column sums that the producing dgrad's epilogue
// already wrote (CG_STATS partial rows [nparts][2 * n_total], sums first): one warp per channel.
__global__ void colsum_partials_kernel(const float* __restrict__ partials, int nparts, int n_total, int C,
                                       float* __restrict__ out) {
    SDN_PDL_ENTRY();
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0;
    for (int i = lane; i < nparts; i += 32) s += (double)partials[(size_t)i * 2 * n_total + c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[c] = (float)s;
}

// Per-channel column sum of an NHWC bf16 tensor (ConvTranspose2d bias gradient).
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long npix, int C,
                                                     float* __restrict__ out, int accumulate) {
    SDN_PDL_ENTRY();
    // one block per 8-channel group slice; grid.x = C/8, grid.y = slices (atomics merge slices)
    const int cg = blockIdx.x;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (long long p = (long long)blockIdx.y * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.y * blockDim.x) {
        float f[8];
        unpack8(ldg16(x + p * C + cg * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
    __shared__ float red[8][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = s[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        atomicAdd(out + cg * 8 + threadIdx.x, v);
    }
    (void)accumulate;
}

// ------------------------------------------------------------ heads + loss
// z_d = w_d . d1 + b_d ; disparity = softplus(z_d)        (model.py:76,98)
// z_l = w_l . d1 + b_l ; logvar    = clamp(z_l, -6, 3)    (model.py:77,103)
// MODE 0: forward only.
// MODE 1: backward with external output gradients g_disp / g_logvar (the
//         nn.Module drop-in path: the loss lives in train.py:329-340).
// MODE 2: fused heteroscedastic Laplace loss (train.py:329-357): forward,
//         5 metric sums, and the loss gradient seeded with 1/n, n read from device.
// sums layout (fp64, atomically accumulated across blocks and across steps, pre-zeroed by the caller: the
// reference accumulates its running sums in Python doubles, train.py:345-357):
//   [0] sum nll  [1] sum |diff|  [2] sum diff^2  [3] sum exp(0.5*logvar); the valid count is a separate u64
// head_grads layout: [0,32) dW_d  [32] db_d  [33,65) dW_l  [65] db_l
// Train mode (bn_scale != nullptr): d1 is the PRE-BatchNorm output y of dec1.block.3 and the kernel applies
// scale / shift / ReLU itself - the post-activation tensor of the last conv layer has no other consumer, so it is never
// written (one full BatchNorm+ReLU pass over a level-1 tensor less).  With bn_partials != nullptr (MODE 1, 2 with
// backward) the kernel also reduces that layer's BatchNorm-backward sums  s1 = sum dz, q = sum dz * (y - mean),
// dz = (z > 0) ? dA : 0,  from the gradient it has just produced: per-block rows [2 * 32] for
// bn_bwd_finalize_kernel, so the layer's separate reduction pass disappears as well.
constexpr int HEAD_THREADS = 384;
template <int MODE>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const bf16* __restrict__ d1, const float* __restrict__ w_d,
                                                   const float* __restrict__ b_d, const float* __restrict__ w_l,
                                                   const float* __restrict__ b_l, float* __restrict__ disp,
                                                   float* __restrict__ logvar, const float* __restrict__ g_disp,
                                                   const float* __restrict__ g_logvar,
                                                   const float* __restrict__ target,
                                                   const uint8_t* __restrict__ mask,
                                                   const unsigned long long* __restrict__ n_valid,
                                                   double* __restrict__ sums,
                                                   unsigned long long* __restrict__ count_out,
                                                   bf16* __restrict__ g_d1, float* __restrict__ head_grads,
                                                   long long npix, const float* __restrict__ bn_scale,
                                                   const float* __restrict__ bn_shift, const float* __restrict__ bn_mean,
                                                   float* __restrict__ bn_partials) {
    SDN_PDL_ENTRY();
    // TWO threads per pixel (an even / odd lane pair), 16 channels each: the per-channel accumulators (head weight
    // gradients, BatchNorm-backward sums) split over the pair, which keeps the kernel out of register spills and lets
    // two blocks share an SM; the two half dot products meet through one shuffle.
    constexpr int NW = HEAD_THREADS / 32;
    __shared__ float red[NW][136];
    __shared__ float s_par[5][32];     // head weights and BatchNorm parameters (registers are for the accumulators)
    const int half = threadIdx.x & 1, c0 = half * 16;
    const bool bn_in = bn_scale != nullptr;
    const bool bn_stats = MODE != 0 && bn_partials != nullptr;
    if (threadIdx.x < 32) {
        s_par[0][threadIdx.x] = w_d[threadIdx.x]; s_par[1][threadIdx.x] = w_l[threadIdx.x];
        s_par[2][threadIdx.x] = bn_in ? bn_scale[threadIdx.x] : 1.f;
        s_par[3][threadIdx.x] = bn_in ? bn_shift[threadIdx.x] : 0.f;
        s_par[4][threadIdx.x] = bn_in ? bn_mean[threadIdx.x] : 0.f;
    }
    __syncthreads();
    const float *wd = s_par[0] + c0, *wl = s_par[1] + c0, *sc = s_par[2] + c0, *sh = s_par[3] + c0, *mu = s_par[4] + c0;
    const float bd = b_d[0], bl = b_l[0];
    float inv_n = 0.f;
    if (MODE == 2) {
        const unsigned long long n = n_valid[0];
        inv_n = n > 0ull ? 1.f / (float)n : 0.f;
    }
    // acc: [0,16) dW_d, [16,32) dW_l of this thread's channels; 32 db_d, 33 db_l, 34..38 sum nll, |diff|, diff^2,
    // exp(.5 logvar), count (counted by the even lane only)
    float acc[39], bs1[16], bq[16];
    if (MODE != 0) {
#pragma unroll
        for (int j = 0; j < 39; ++j) acc[j] = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { bs1[j] = 0.f; bq[j] = 0.f; }
    }
    // The warp stays converged for the pair shuffles: the loop runs while ANY pair of the warp has a pixel left and
    // a pair past the end only idles (pair-local shuffle masks measured 15 % slower).
    const unsigned pmask = 0xffffffffu;
    const long long p_first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const long long p_step = ((long long)gridDim.x * blockDim.x) >> 1;
    // the NEXT pixel's operands are requested before the current one is processed (a DRAM round trip per pixel
    // cannot be hidden otherwise at this occupancy)
    uint4 nxt[2];
    float nxt_tg = 0.f;
    uint8_t nxt_mk = 0;
    if (p_first < npix) {
        nxt[0] = ldg16(d1 + p_first * 32 + c0);
        nxt[1] = ldg16(d1 + p_first * 32 + c0 + 8);
        if (MODE == 2) { nxt_tg = __ldg(target + p_first); nxt_mk = __ldg(mask + p_first); }
    }
    for (long long p = p_first; __any_sync(0xffffffffu, p < npix); p += p_step) {
        const bool active = p < npix;
        const uint4 cur0 = nxt[0], cur1 = nxt[1];
        const float cur_tg = nxt_tg;
        const uint8_t cur_mk = active ? nxt_mk : (uint8_t)0;
        if (p + p_step < npix) {
            nxt[0] = ldg16(d1 + (p + p_step) * 32 + c0);
            nxt[1] = ldg16(d1 + (p + p_step) * 32 + c0 + 8);
            if (MODE == 2) { nxt_tg = __ldg(target + p + p_step); nxt_mk = __ldg(mask + p + p_step); }
        }
        float yv[16], f[16];
        {
            float t8[8];
            unpack8(cur0, t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) yv[j] = t8[j];
            unpack8(cur1, t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) yv[8 + j] = t8[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)   // a = relu(y * scale + shift), kept in fp32 (this tensor is never stored)
            f[j] = bn_in ? fmaxf(fmaf(yv[j], sc[j], sh[j]), 0.f) : yv[j];
        float zd = 0.f, zl = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { zd = fmaf(f[j], wd[j], zd); zl = fmaf(f[j], wl[j], zl); }
        // fixed summation order (channels 0-15, then 16-31, then the bias) on both lanes of the pair
        const float zd_o = __shfl_xor_sync(pmask, zd, 1), zl_o = __shfl_xor_sync(pmask, zl, 1);
        zd = (half ? zd_o + zd : zd + zd_o) + bd;
        zl = (half ? zl_o + zl : zl + zl_o) + bl;
        // the scalar part of the pixel (softplus / clamp, loss, gradient seeds) runs on the even lane only; the odd
        // lane receives the two seeds through shuffles
        float dzd = 0.f, dzl = 0.f;
        if (half == 0 && active) {
            // softplus / sigmoid / exp on the special-function unit (the scalar chain of this lane is the kernel's
            // critical path): log1p(e) as its series below e = exp(-5) (error e^4/4 < 1e-7 relative), __logf(1 + e)
            // above it (<= 1e-5 relative); -DSDN_AUG_PRECISE restores libm
#ifdef SDN_AUG_PRECISE
            const float ez = expf(fminf(zd, 20.f));
            const float dsp = zd > 20.f ? zd : log1pf(ez);
#else
            const float ez = __expf(fminf(zd, 20.f));
            const float dsp = zd > 20.f ? zd : (zd < -5.f ? ez * (1.f - ez * (0.5f - ez * (1.f / 3.f))) : __logf(1.f + ez));
#endif
            const float lv = fminf(fmaxf(zl, -6.f), 3.f);
            if (MODE != 1) {
                if (disp != nullptr) disp[p] = dsp;
                if (logvar != nullptr) logvar[p] = lv;
            }
            if (MODE != 0) {
                float gd, gl;
                if (MODE == 1) {
                    gd = g_disp[p];
                    gl = g_logvar != nullptr ? g_logvar[p] : 0.f;
                } else {
                    const float tg = cur_tg;
                    const bool m = (cur_mk != 0) && isfinite(tg);
                    gd = 0.f; gl = 0.f;
                    if (m) {
                        const float diff = dsp - tg;
                        const float ad = fabsf(diff);
#ifdef SDN_AUG_PRECISE
                        const float e = expf(-lv);
#else
                        const float e = __expf(-lv);
#endif
                        acc[34] += ad * e + lv;
                        acc[35] += ad;
                        acc[36] = fmaf(diff, diff, acc[36]);
#ifdef SDN_AUG_PRECISE
                        acc[37] += expf(0.5f * lv);
#else
                        acc[37] += rsqrtf(e);          // exp(lv / 2) = exp(-lv)^(-1/2)
#endif
                        acc[38] += 1.f;
                        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                        gd = sg * e * inv_n;
                        gl = (1.f - ad * e) * inv_n;
                    }
                }
#ifdef SDN_AUG_PRECISE
                const float sig = zd > 20.f ? 1.f : 1.f / (1.f + expf(-zd));
#else
                const float sig = zd > 20.f ? 1.f : __fdividef(ez, 1.f + ez);     // e^z / (1 + e^z)
#endif
                dzd = gd * sig;
                dzl = (zl >= -6.f && zl <= 3.f) ? gl : 0.f;
            }
        }
        if (MODE == 0) continue;
        dzd = __shfl_sync(pmask, dzd, (threadIdx.x & 31) & ~1);
        dzl = __shfl_sync(pmask, dzl, (threadIdx.x & 31) & ~1);
        if (!active) continue;      // (after the last shuffle of the iteration)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(dzd, wd[v * 8 + j], dzl * wl[v * 8 + j]);
            const uint4 packed = pack8(o);
            *reinterpret_cast<uint4*>(g_d1 + p * 32 + c0 + v * 8) = packed;
            if (bn_stats) {
                float g8[8];
                unpack8(packed, g8);          // the bf16 values the apply pass will read back
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ch = v * 8 + j;
                    const float dz = f[ch] > 0.f ? g8[j] : 0.f;      // relu(z) > 0 <=> z > 0
                    bs1[ch] += dz;
                    bq[ch] = fmaf(dz, yv[ch] - mu[ch], bq[ch]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc[j] = fmaf(dzd, f[j], acc[j]); acc[16 + j] = fmaf(dzl, f[j], acc[16 + j]); }
        if (half == 0) { acc[32] += dzd; acc[33] += dzl; }
    }
    if (MODE == 0) return;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // per-warp reduction over the 16 lanes of each parity (offsets 2..16), then the 8 warps through shared memory.
    // red row layout: [0,32) dW_d by channel, [32,64) dW_l, 64 db_d, 65 db_l, 66..70 sums, [72,104) BatchNorm s1, [104,136) q
#pragma unroll
    for (int j = 0; j < 39; ++j) {
        float v = acc[j];
        for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (j >= 32) v += __shfl_xor_sync(0xffffffffu, v, 1);      // scalars: both parities (the odd lanes hold zeros)
        if (lane < 2) {
            if (j < 16) red[wrp][lane * 16 + j] = v;
            else if (j < 32) red[wrp][32 + lane * 16 + (j - 16)] = v;
            else if (lane == 0) red[wrp][64 + (j - 32)] = v;
        }
    }
    if (bn_stats) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float a = bs1[j], b = bq[j];
            for (int o = 16; o > 1; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            if (lane < 2) { red[wrp][72 + lane * 16 + j] = a; red[wrp][104 + lane * 16 + j] = b; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 71) {
        float v = 0.f;
        for (int w = 0; w < NW; ++w) v += red[w][threadIdx.x];
        // head_grads layout: [0,32) dW_d  [32] db_d  [33,65) dW_l  [65] db_l
        if (threadIdx.x < 32) atomicAdd(head_grads + threadIdx.x, v);
        else if (threadIdx.x < 64) atomicAdd(head_grads + 33 + (threadIdx.x - 32), v);
        else if (threadIdx.x == 64) atomicAdd(head_grads + 32, v);
        else if (threadIdx.x == 65) atomicAdd(head_grads + 65, v);
        else if (MODE == 2) {
            if (threadIdx.x < 70) atomicAdd(sums + (threadIdx.x - 66), (double)v);
            else atomicAdd(count_out, (unsigned long long)v);  // block partial < 2^24: exact
        }
    }
    if (bn_stats && threadIdx.x >= 128 && threadIdx.x < 192) {
        // this block's BatchNorm-backward partial row [s1[32] | q[32]] for bn_bwd_finalize_kernel
        const int t = threadIdx.x - 128;
        float v = 0.f;
        for (int w = 0; w < NW; ++w) v += red[w][72 + t];
        bn_partials[(size_t)blockIdx.x * 64 + t] = v;
    }
}

// n = number of pixels with valid_mask & isfinite(target) (train.py:329-330), as a u64 on device.
__global__ void __launch_bounds__(256) mask_count_kernel(const float* __restrict__ target,
                                                         const uint8_t* __restrict__ mask, long long npix,
                                                         unsigned long long* __restrict__ n_out) {
    SDN_PDL_ENTRY();
    unsigned int c = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (long long)gridDim.x * blockDim.x)
        c += (mask[p] != 0 && isfinite(target[p])) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ unsigned int red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(n_out, (unsigned long long)t);
    }
}

// ------------------------------------------------------------------- AdamW
// torch.optim.AdamW (decoupled weight decay, train.py:578) over all 66 parameter tensors in ONE launch,
// same per-element operation order as torch's _single_tensor_adamw.  `gate` (device u64, e.g. the global
// valid-pixel count) == 0 skips the whole update on the device, which is the reference's "no valid pixel
// -> no optimizer step" rule (train.py:331-332) without a host round trip.  step_dev counts applied steps.
struct AdamTable {
    float* p[66];
    const float* g[66];
    float* m[66];
    float* v[66];
    int start[67];
    int n;
};
__global__ void adamw_step_count_kernel(long long* step_dev, const unsigned long long* gate) {
    SDN_PDL_ENTRY();
    if (gate == nullptr || *gate != 0ull) *step_dev += 1;
}
__global__ void __launch_bounds__(256) adamw_all_kernel(const __grid_constant__ AdamTable t, double lr, double beta1,
                                                        double beta2, float one_minus_b1, float one_minus_b2,
                                                        float eps, float decay,
                                                        const long long* __restrict__ step_dev,
                                                        const unsigned long long* __restrict__ gate) {
    SDN_PDL_ENTRY();
    if (gate != nullptr && *gate == 0ull) return;
    __shared__ int starts[67];
    if (threadIdx.x <= t.n) starts[threadIdx.x] = t.start[threadIdx.x];
    __syncthreads();
    const double step = (double)(*step_dev);
    const float b2f = (float)beta2;
    const float step_size = (float)(lr / (1.0 - pow(beta1, step)));
    const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, step));
    const int total = starts[t.n];
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int lo = 0, hi = t.n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (starts[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        const int i = idx - starts[lo];
        const float g = t.g[lo][i];
        float p = t.p[lo][i] * decay;
        float m = t.m[lo][i];
        m = m + one_minus_b1 * (g - m);                        // lerp_
        const float v = t.v[lo][i] * b2f + one_minus_b2 * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p = p - step_size * (m / denom);
        t.p[lo][i] = p; t.m[lo][i] = m; t.v[lo][i] = v;
    }
}

// ------------------------------------------------------------- grad unpack
// workspace layouts (written by wgrad_gemm_kernel) -> torch parameter layouts.
// mode 0: conv3x3   ws[(tap*Ci + ci)][Co]  -> grad[Co][Ci][3][3]
// mode 2: first     ws[k][Co], k=tap*Ci+ci -> grad[Co][Ci][3][3]   (same formula, ws has >= 9*Ci rows)
// mode 3: convT     ws[q][ci][Co]          -> grad[Ci][Co][2][2]
// mode 4: first layer in the row-halo form (see im2col_rows_kernel)
__global__ void unpack_grad_kernel(const float* __restrict__ ws, float* __restrict__ grad, int mode, int Co, int Ci,
                                   int accumulate) {
    SDN_PDL_ENTRY();
    const int total = (mode == 3 || mode == 5) ? 4 * Co * Ci : 9 * Co * Ci;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float v;
        if (mode == 3) {
            const int ci = i / (Co * 4), rem = i % (Co * 4), co = rem / 4, q = rem % 4;
            v = ws[((size_t)q * Ci + ci) * Co + co];
        } else if (mode == 5) {
            // ConvTranspose2d, quadrant PAIRS as the dY variants: ws[i2][ci][j * Co + co], quadrant q = 2 * i2 + j
            const int ci = i / (Co * 4), rem = i % (Co * 4), co = rem / 4, q = rem % 4;
            v = ws[((size_t)(q >> 1) * Ci + ci) * (2 * Co) + (q & 1) * Co + co];
        } else if (mode == 4) {
            // first layer, row-halo form: ws[((dy*3 + 1)*32 + dx*Ci + c)][Co] -> grad[Co][Ci][3][3]
            const int co = i / (Ci * 9), rem = i % (Ci * 9), ci = rem / 9, tap = rem % 9;
            v = ws[((size_t)((tap / 3) * 3 + 1) * 32 + (tap % 3) * Ci + ci) * Co + co];
        } else {
            const int co = i / (Ci * 9), rem = i % (Ci * 9), ci = rem / 9, tap = rem % 9;
            v = ws[((size_t)tap * Ci + ci) * Co + co];
        }
        if (accumulate) grad[i] += v; else grad[i] = v;
    }
}

// Zero fills as kernels: cudaMemsetAsync may be served by a copy engine, where it would queue behind a
// bulk host->device prefetch running on another stream.
__global__ void zero_u64_kernel(unsigned long long* __restrict__ p, int n) {
    SDN_PDL_ENTRY();
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0ull;
}
__global__ void __launch_bounds__(256) zero_u32_kernel(uint32_t* __restrict__ p, size_t n) {
    SDN_PDL_ENTRY();
    const size_t n4 = n / 4;
    uint4* p4 = reinterpret_cast<uint4*>(p);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x)
        p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = n4 * 4 + blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) p[i] = 0u;
}
__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int accumulate) {
    SDN_PDL_ENTRY();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (accumulate) dst[i] += src[i]; else dst[i] = src[i];
    }
}

}  // namespace sdn
