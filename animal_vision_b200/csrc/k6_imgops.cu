// K6: generic float32 image operators -- the building blocks of every path that is NOT a fused uint8
// kernel: float / wide-integer frames of Cat and HoneyBee, HoneyBee's hsi_downsample route and large
// blur sigmas, and the UV species that are compositions of these steps (reference uv_helpers.py).
//
//   avb_img_to_float01   uv_helpers.py:15-23 to_float01 / animals/animal_utils.py:41-50 get_normalized_image
//   avb_img_resample     cv2.resize as uv_helpers.py:57-64 / :84-99 / :155-183 use it (INTER_AREA, INTER_LINEAR,
//                        INTER_CUBIC): one axis at a time with host-built (index, weight) tap tables, the
//                        horizontal pass first as OpenCV does
//   avb_img_blur         cv2.GaussianBlur(BORDER_REFLECT101) with explicit taps (uv_helpers.py:67-73), any radius
//   avb_img_stats        per frame and channel min / max / mean (safe_norm :47-53, von Kries :195-206)
//   avb_img_percentile   numpy.percentile(method="linear") of strided planes: exact radix select
//
// Images are device float32, packed [n, H, W, C] (C <= 4), contiguous unless strides are passed.  These
// kernels are plain grid-stride CUDA: the fused uint8 kernels (K1-K3) are the throughput path, this file
// is the generality path; every step still runs on the GPU (no CPU fallback anywhere).
#include <algorithm>

#include "avb_common.cuh"

namespace avb {
namespace img {

constexpr unsigned FULL = 0xffffffffu;

static unsigned grid_for(long long items, int per_sm = 16) {
    const long long want = (items + 255) / 256, cap = (long long)sm_count() * per_sm;
    return (unsigned)std::max<long long>(1, std::min(want, cap));
}

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// order-preserving map float -> uint32 (negative floats included) and back
__device__ __forceinline__ uint32_t ord_bits(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ------------------------------------------------------------------ frame maximum (u8 or f32 input)
template <class T>
__global__ void __launch_bounds__(256) frame_max_kernel(const T *__restrict__ in, long long per_frame, long long frame_stride, uint32_t *maxbits) {
    const T *f = in + (long long)blockIdx.y * frame_stride;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < per_frame; i += (long long)gridDim.x * 256) m = fmaxf(m, (float)f[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits + blockIdx.y, __float_as_uint(m));
}

struct ToFloatP {
    const void *in; float *out;
    long long per_frame, total;
    int in_is_u8, mode;
    const uint32_t *maxbits;
};
__global__ void __launch_bounds__(256) to_float01_kernel(const __grid_constant__ ToFloatP p) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.total; i += (long long)gridDim.x * 256) {
        const float mx = __uint_as_float(p.maxbits[i / p.per_frame]);
        float v = p.in_is_u8 ? (float)static_cast<const uint8_t *>(p.in)[i] : static_cast<const float *>(p.in)[i];
        if (p.mode == AVB_IMG_NORM_UV) {
            // uv_helpers.py:15-23: uint8 is always divided; other dtypes only when max > 1.001 (then clipped)
            if (p.in_is_u8) v = __fdiv_rn(v, 255.0f);
            else if (mx > 1.001f) v = clip01(__fdiv_rn(v, 255.0f));
        } else {
            // animal_utils.py:41-50: divide when max > 1, always clip
            if (mx > 1.0f) v = __fdiv_rn(v, 255.0f);
            v = clip01(v);
        }
        p.out[i] = v;
    }
}

// ------------------------------------------------------------------ resample along one axis
struct ResampleP {
    const float *in; float *out;
    int n, Hout, Wout, C, taps, axis;
    long long in_fs, in_rs;          // element strides of the input (frame, row); pixels are packed (C floats)
    const int *idx; const float *w;  // [len(axis)][taps]
};
// grid (ceil(Wout / 256), Hout, n): a thread owns one output pixel (all C channels share its taps) -- no per-element
// 64-bit division, the tap index / weight loads are amortised over the channels
__global__ void __launch_bounds__(256) resample_kernel(const __grid_constant__ ResampleP p) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= p.Wout) return;
    const float *base = p.in + (long long)blockIdx.z * p.in_fs;
    float *o = p.out + (((long long)blockIdx.z * p.Hout + y) * p.Wout + x) * p.C;
    const int sel = p.axis == 0 ? x : y;
    const int *ix = p.idx + (long long)sel * p.taps;
    const float *wt = p.w + (long long)sel * p.taps;
    for (int c0 = 0; c0 < p.C; c0 += 4) {                              // four channels at a time keep the accumulators in registers
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int nc = min(4, p.C - c0);
        for (int t = 0; t < p.taps; ++t) {
            const float wv = __ldg(wt + t);
            const long long s = __ldg(ix + t);
            const float *q = p.axis == 0 ? base + (long long)y * p.in_rs + s * p.C + c0 : base + s * p.in_rs + (long long)x * p.C + c0;
            for (int c = 0; c < nc; ++c) acc[c] = fmaf(wv, q[c], acc[c]);
        }
        for (int c = 0; c < nc; ++c) o[c0 + c] = acc[c];
    }
}

// ------------------------------------------------------------------ blur along one axis (REFLECT_101)
struct BlurP {
    const float *in; float *out;
    int n, H, W, C, R, axis;
    const float *taps;               // device, 2R+1
};
__global__ void __launch_bounds__(256) blur_kernel(const __grid_constant__ BlurP p) {
    extern __shared__ float taps_s[];
    for (int i = threadIdx.x; i < 2 * p.R + 1; i += 256) taps_s[i] = __ldg(p.taps + i);
    __syncthreads();
    const long long total = (long long)p.n * p.H * p.W * p.C;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % p.C);
        const long long px = e / p.C;
        const int x = (int)(px % p.W), y = (int)((px / p.W) % p.H);
        const long long fbase = (px / ((long long)p.H * p.W)) * p.H * p.W;
        float acc = 0.f;
        if (p.axis == 0) {
            const long long rbase = fbase + (long long)y * p.W;
            if (x >= p.R && x + p.R < p.W) {
                const float *q = p.in + (rbase + x - p.R) * p.C + c;
                for (int k = 0; k <= 2 * p.R; ++k) acc = fmaf(taps_s[k], q[(long long)k * p.C], acc);
            } else {
                for (int k = -p.R; k <= p.R; ++k) acc = fmaf(taps_s[k + p.R], p.in[(rbase + reflect101(x + k, p.W)) * p.C + c], acc);
            }
        } else {
            if (y >= p.R && y + p.R < p.H) {
                const float *q = p.in + (fbase + (long long)(y - p.R) * p.W + x) * p.C + c;
                for (int k = 0; k <= 2 * p.R; ++k) acc = fmaf(taps_s[k], q[(long long)k * p.W * p.C], acc);
            } else {
                for (int k = -p.R; k <= p.R; ++k) acc = fmaf(taps_s[k + p.R], p.in[(fbase + (long long)reflect101(y + k, p.H) * p.W + x) * p.C + c], acc);
            }
        }
        p.out[e] = acc;
    }
}

// The same separable correlation as ONE tiled kernel (C <= 4, the planes the UV species blur): a CTA owns 32 rows x 64 pixels,
// stages them with their (Ry, Rx) halo in shared memory (REFLECT_101 resolved while loading), runs the row pass into a
// second shared tile and the column pass from there -- the intermediate never goes to HBM and every input element is read
// once instead of 2R+1 times through L1.  Tap order and fused multiply-adds are those of blur_kernel: identical bits.
constexpr int BT_X = 64, BT_Y = 32;
struct BlurTileP {
    const float *in; float *out;
    int n, H, W, C, Rx, Ry;
    const float *taps_x, *taps_y;
};
__global__ void __launch_bounds__(256) blur_tile_kernel(const __grid_constant__ BlurTileP p) {
    extern __shared__ float bt_smem[];
    const int C = p.C, Rx = p.Rx, Ry = p.Ry;
    const int SW = (BT_X + 2 * Rx) * C, TW = BT_X * C, SH = BT_Y + 2 * Ry;
    float *S = bt_smem;                      // [SH][SW] input tile with halo
    float *T = S + SH * SW;                  // [SH][TW] after the row pass
    float *tx = T + SH * TW, *ty = tx + 2 * Rx + 1;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * Rx + 1; i += 256) tx[i] = __ldg(p.taps_x + i);
    for (int i = tid; i < 2 * Ry + 1; i += 256) ty[i] = __ldg(p.taps_y + i);
    const int x0 = blockIdx.x * BT_X, y0 = blockIdx.y * BT_Y;
    const float *f = p.in + (long long)blockIdx.z * p.H * p.W * C;
    const int tr = tid >> 6, tc = tid & 63;                            // 4 rows x 64 lanes: no per-element division anywhere
    for (int r = tr; r < SH; r += 4) {
        const float *row = f + (long long)reflect101(y0 - Ry + r, p.H) * p.W * C;
        for (int cx = tc; cx < BT_X + 2 * Rx; cx += 64) {
            const float *q = row + (long long)reflect101(x0 - Rx + cx, p.W) * C;
            for (int c = 0; c < C; ++c) S[r * SW + cx * C + c] = __ldg(q + c);
        }
    }
    __syncthreads();
    for (int r = tr; r < SH; r += 4) {
        for (int i = tc; i < TW; i += 64) {
            const float *q = S + r * SW + i;
            float acc = 0.f;
            for (int k = 0; k <= 2 * Rx; ++k) acc = fmaf(tx[k], q[k * C], acc);
            T[r * TW + i] = acc;
        }
    }
    __syncthreads();
    float *o = p.out + (long long)blockIdx.z * p.H * p.W * C;
    const int live = min(TW, (p.W - x0) * C);                           // elements of a tile row inside the image
    for (int r = tr; r < BT_Y && y0 + r < p.H; r += 4) {
        float *orow = o + ((long long)(y0 + r) * p.W + x0) * C;
        for (int i = tc; i < live; i += 64) {
            const float *q = T + r * TW + i;
            float acc = 0.f;
            for (int k = 0; k <= 2 * Ry; ++k) acc = fmaf(ty[k], q[k * TW], acc);
            orow[i] = acc;
        }
    }
}

// ------------------------------------------------------------------ per frame / channel min, max, sum
struct StatsAcc {                    // scratch per (frame, channel)
    uint32_t mn, mx;
    double sum;
};
constexpr int STATS_MAX_C = 16;       // mantis_shrimp.py:167-171: ten band maps normalised at once
__global__ void __launch_bounds__(256) stats_kernel(const float *__restrict__ in, long long npx, int C, StatsAcc *acc) {
    const int frame = blockIdx.y;
    const float *f = in + (long long)frame * npx * C;
    uint32_t mn[STATS_MAX_C], mx[STATS_MAX_C];
    double sm[STATS_MAX_C];
#pragma unroll
    for (int c = 0; c < STATS_MAX_C; ++c) { mn[c] = 0xffffffffu; mx[c] = 0u; sm[c] = 0.0; }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npx; i += (long long)gridDim.x * 256) {
#pragma unroll
        for (int c = 0; c < STATS_MAX_C; ++c) {
            if (c < C) {
                const float v = f[i * C + c];
                const uint32_t o = ord_bits(v);
                mn[c] = min(mn[c], o); mx[c] = max(mx[c], o);
                sm[c] += (double)v;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < STATS_MAX_C; ++c) {
        if (c >= C) break;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = min(mn[c], __shfl_xor_sync(FULL, mn[c], o));
            mx[c] = max(mx[c], __shfl_xor_sync(FULL, mx[c], o));
            sm[c] += __shfl_xor_sync(FULL, sm[c], o);
        }
        if ((threadIdx.x & 31) == 0) {
            StatsAcc *a = acc + (long long)frame * C + c;
            atomicMin(&a->mn, mn[c]); atomicMax(&a->mx, mx[c]); atomicAdd(&a->sum, sm[c]);
        }
    }
}
__global__ void stats_init_kernel(StatsAcc *acc, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { acc[i].mn = 0xffffffffu; acc[i].mx = 0u; acc[i].sum = 0.0; }
}
__global__ void stats_finish_kernel(const StatsAcc *acc, int count, double npx, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        out[4 * i] = ord_float(acc[i].mn);
        out[4 * i + 1] = ord_float(acc[i].mx);
        out[4 * i + 2] = (float)(acc[i].sum / npx);
        out[4 * i + 3] = 0.f;
    }
}

// ------------------------------------------------------------------ cat binocular wide-FOV warp on float frames
// animals/cat_widevision_utils.py:46-99 on a float32 [0,1] frame: cv2.remap INTER_LINEAR (map coordinate quantised
// to 1/32 px, two taps, BORDER_CONSTANT 0; the row map is the identity) for the two eye views, cos^2 blend,
// division by wL + wR + 1e-8, clip.  Table layout as avb_cat_u8's warp_dev: xL, xR, wL, wR, ws, 1/ws (6*W floats).
struct CatWarpP {
    const float *in; float *out; const float *tab;
    int n, H, W;
};
__device__ __forceinline__ void warp_tap(const float *row, int W, float xs, float &c0, float &c1, float &c2) {
    const int sx = __float2int_rn(xs * 32.0f);
    const int ix = sx >> 5;
    const float f = (float)(sx & 31) * (1.0f / 32.0f), w0 = 1.0f - f;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if ((unsigned)ix < (unsigned)W) { a0 = row[3 * ix]; a1 = row[3 * ix + 1]; a2 = row[3 * ix + 2]; }
    if ((unsigned)(ix + 1) < (unsigned)W) { b0 = row[3 * ix + 3]; b1 = row[3 * ix + 4]; b2 = row[3 * ix + 5]; }
    c0 = __fadd_rn(__fmul_rn(a0, w0), __fmul_rn(b0, f));
    c1 = __fadd_rn(__fmul_rn(a1, w0), __fmul_rn(b1, f));
    c2 = __fadd_rn(__fmul_rn(a2, w0), __fmul_rn(b2, f));
}
__global__ void __launch_bounds__(256) cat_warp_kernel(const __grid_constant__ CatWarpP p) {
    const long long npx = (long long)p.n * p.H * p.W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npx; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % p.W);
        const float *row = p.in + (i - x) * 3;
        const float wL = __ldg(p.tab + 2 * p.W + x), wR = __ldg(p.tab + 3 * p.W + x), ws = __ldg(p.tab + 4 * p.W + x);
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (wL != 0.0f) warp_tap(row, p.W, __ldg(p.tab + x), l0, l1, l2);
        if (wR != 0.0f) warp_tap(row, p.W, __ldg(p.tab + p.W + x), r0, r1, r2);
        p.out[3 * i] = clip01(__fdiv_rn(__fadd_rn(__fmul_rn(l0, wL), __fmul_rn(r0, wR)), ws));
        p.out[3 * i + 1] = clip01(__fdiv_rn(__fadd_rn(__fmul_rn(l1, wL), __fmul_rn(r1, wR)), ws));
        p.out[3 * i + 2] = clip01(__fdiv_rn(__fadd_rn(__fmul_rn(l2, wL), __fmul_rn(r2, wR)), ws));
    }
}

// ------------------------------------------------------------------ von Kries: x / max(stat, eps) per frame and channel
struct DivideP {
    const float *in; float *out; const float *stats;     // stats: [n][C][4] = min, max, mean, 0 (avb_img_stats)
    long long npx, total;
    int C, which;
    float eps;
};
__global__ void __launch_bounds__(256) divide_kernel(const __grid_constant__ DivideP p) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < p.total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % p.C);
        const long long f = e / (p.npx * p.C);
        const float s = fmaxf(p.stats[(f * p.C + c) * 4 + p.which], p.eps);     // uv_helpers.py:195-206
        p.out[e] = __fdiv_rn(p.in[e], s);
    }
}

// ------------------------------------------------------------------ exact percentiles
constexpr int PCT_MAX = 16, PCT_BINS = 2048;
struct PctState {
    uint32_t prefix, rank, cnt_le, min_gt;
};
struct PctP {
    const float *in;
    long long npx, stride;
    int nreq, level;
    long long off[PCT_MAX], k_lo[PCT_MAX], k_hi[PCT_MAX];
    float gamma[PCT_MAX];
    uint32_t *hist;                  // [nreq][PCT_BINS]
    PctState *st;                    // [nreq]
    float *out;                      // [nreq]
};
__device__ __constant__ int PCT_SHIFT[3] = {21, 10, 0};
__device__ __constant__ int PCT_NBITS[3] = {11, 11, 10};

__global__ void pct_init_kernel(const __grid_constant__ PctP p) {
    const int r = blockIdx.x;
    for (int i = threadIdx.x; i < PCT_BINS; i += blockDim.x) p.hist[r * PCT_BINS + i] = 0u;
    if (threadIdx.x == 0) p.st[r] = PctState{0u, (uint32_t)p.k_lo[r], 0u, 0xffffffffu};
}

// level pass: histogram of this level's digit among the values that match the prefix found so far
__global__ void __launch_bounds__(256) pct_hist_kernel(const __grid_constant__ PctP p) {
    __shared__ uint32_t hs[PCT_BINS];
    const int r = blockIdx.y, lv = p.level;
    for (int i = threadIdx.x; i < PCT_BINS; i += 256) hs[i] = 0u;
    __syncthreads();
    const int sh = PCT_SHIFT[lv], nb = PCT_NBITS[lv];
    const uint32_t mask = (1u << nb) - 1u, prefix = p.st[r].prefix;
    const float *src = p.in + p.off[r];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.npx; i += (long long)gridDim.x * 256) {
        const uint32_t b = ord_bits(src[i * p.stride]);
        if (lv == 0 || (b >> (sh + nb)) == prefix) atomicAdd(&hs[(b >> sh) & mask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PCT_BINS; i += 256)
        if (hs[i]) atomicAdd(&p.hist[r * PCT_BINS + i], hs[i]);
}
// pick the bin that holds the rank, extend the prefix, clear the histogram for the next level
__global__ void __launch_bounds__(1024) pct_pick_kernel(const __grid_constant__ PctP p) {
    __shared__ uint32_t part[1024];
    __shared__ uint32_t found[2];
    const int r = blockIdx.x, tid = threadIdx.x, nb = PCT_NBITS[p.level], nbins = 1 << nb;
    uint32_t *h = p.hist + r * PCT_BINS;
    const uint32_t a = 2 * tid < nbins ? h[2 * tid] : 0u, b = 2 * tid + 1 < nbins ? h[2 * tid + 1] : 0u;
    part[tid] = a + b;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
    }
    __syncthreads();
    const uint32_t rank = p.st[r].rank, run = part[tid];
    if (rank >= run && rank < run + a) { found[0] = 2 * tid; found[1] = rank - run; }
    else if (rank >= run + a && rank < run + a + b) { found[0] = 2 * tid + 1; found[1] = rank - run - a; }
    __syncthreads();
    if (tid == 0) { p.st[r].prefix = (p.st[r].prefix << nb) | found[0]; p.st[r].rank = found[1]; }
    h[2 * tid] = 0u; h[2 * tid + 1] = 0u;
}
// a = value of rank k_lo is known: count values <= a and the smallest value above it
__global__ void __launch_bounds__(256) pct_upper_kernel(const __grid_constant__ PctP p) {
    const int r = blockIdx.y;
    const uint32_t a = p.st[r].prefix;
    const float *src = p.in + p.off[r];
    uint32_t le = 0, mg = 0xffffffffu;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.npx; i += (long long)gridDim.x * 256) {
        const uint32_t b = ord_bits(src[i * p.stride]);
        if (b <= a) ++le; else mg = min(mg, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        le += __shfl_xor_sync(FULL, le, o);
        mg = min(mg, __shfl_xor_sync(FULL, mg, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&p.st[r].cnt_le, le); atomicMin(&p.st[r].min_gt, mg); }
}
__global__ void pct_finish_kernel(const __grid_constant__ PctP p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.nreq) return;
    const PctState s = p.st[r];
    uint32_t hi = s.prefix;
    if ((uint32_t)p.k_hi[r] >= s.cnt_le && s.min_gt != 0xffffffffu) hi = s.min_gt;
    p.out[r] = numpy_lerp(ord_float(s.prefix), ord_float(hi), p.gamma[r]);
}

}  // namespace img
}  // namespace avb

using namespace avb;
using namespace avb::img;

extern "C" int avb_img_to_float01(const void *in_dev, int in_is_u8, float *out_dev, int n, int64_t per_frame, int mode,
                                  uint32_t *scratch_dev, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && scratch_dev, "null pointer");
    AVB_REQUIRE(n > 0 && n <= 65535 && per_frame > 0, "bad geometry");
    AVB_REQUIRE(mode == AVB_IMG_NORM_UV || mode == AVB_IMG_NORM_MAMMAL, "unknown mode");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AVB_CUDA_OK(cudaMemsetAsync(scratch_dev, 0, sizeof(uint32_t) * n, st));
    const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((per_frame + 255) / 256, sm_count() * 8 / n + 1));
    {
        AVB_TIMED("k6_frame_max", st);
        if (in_is_u8) frame_max_kernel<uint8_t><<<dim3(bx, n), 256, 0, st>>>(static_cast<const uint8_t *>(in_dev), per_frame, per_frame, scratch_dev);
        else frame_max_kernel<float><<<dim3(bx, n), 256, 0, st>>>(static_cast<const float *>(in_dev), per_frame, per_frame, scratch_dev);
    }
    ToFloatP p{in_dev, out_dev, per_frame, per_frame * n, in_is_u8, mode, scratch_dev};
    {
        AVB_TIMED("k6_to_float01", st);
        to_float01_kernel<<<grid_for(p.total), 256, 0, st>>>(p);
    }
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_img_resample(const float *in_dev, float *out_dev, int n, int Hout, int Wout, int C, int axis,
                                int64_t in_frame_stride, int64_t in_row_stride, const int32_t *idx_dev, const float *w_dev,
                                int taps, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && idx_dev && w_dev, "null pointer");
    AVB_REQUIRE(n > 0 && Hout > 0 && Wout > 0 && C >= 1 && C <= 64 && taps >= 1 && (axis == 0 || axis == 1), "bad geometry");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ResampleP p{in_dev, out_dev, n, Hout, Wout, C, taps, axis, in_frame_stride, in_row_stride, idx_dev, w_dev};
    AVB_TIMED(axis == 0 ? "k6_resample_x" : "k6_resample_y", st);
    AVB_REQUIRE(Hout <= 65535 && n <= 65535, "frame height / batch too large for one launch");
    resample_kernel<<<dim3((Wout + 255) / 256, Hout, n), 256, 0, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_img_blur(const float *in_dev, float *out_dev, float *tmp_dev, int n, int H, int W, int C,
                            const float *taps_x_dev, int kx, const float *taps_y_dev, int ky, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && tmp_dev, "null pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0 && C >= 1 && C <= 64, "bad geometry");
    AVB_REQUIRE(kx >= 1 && (kx & 1) && ky >= 1 && (ky & 1) && kx <= 4097 && ky <= 4097 && taps_x_dev && taps_y_dev, "tap counts must be odd");
    AVB_REQUIRE(tmp_dev != in_dev && tmp_dev != out_dev, "tmp must not alias in / out");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned g = grid_for((long long)n * H * W * C);
    AVB_TIMED("k6_blur", st);
    {
        const int Rx = kx / 2, Ry = ky / 2;
        const size_t smem = sizeof(float) * ((size_t)(BT_Y + 2 * Ry) * (BT_X + 2 * Rx) * C + (size_t)(BT_Y + 2 * Ry) * BT_X * C + kx + ky);
        if (C <= 4 && smem <= 100 * 1024 && n <= 65535 && in_dev != out_dev) {
            static SmemOptIn optin;
            AVB_CUDA_OK(optin.ensure(blur_tile_kernel, 100 * 1024));
            BlurTileP tp{in_dev, out_dev, n, H, W, C, Rx, Ry, taps_x_dev, taps_y_dev};
            blur_tile_kernel<<<dim3((W + BT_X - 1) / BT_X, (H + BT_Y - 1) / BT_Y, n), 256, smem, st>>>(tp);
            AVB_CUDA_OK(cudaGetLastError());
            return AVB_OK;
        }
    }
    BlurP p{in_dev, tmp_dev, n, H, W, C, kx / 2, 0, taps_x_dev};          // rows first, as cv2.GaussianBlur
    blur_kernel<<<g, 256, sizeof(float) * kx, st>>>(p);
    p.in = tmp_dev; p.out = out_dev; p.R = ky / 2; p.axis = 1; p.taps = taps_y_dev;
    blur_kernel<<<g, 256, sizeof(float) * ky, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_img_stats(const float *in_dev, int n, int64_t npx, int C, float *out_dev, void *scratch_dev, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && scratch_dev, "null pointer");
    AVB_REQUIRE(n > 0 && n <= 65535 && npx > 0 && C >= 1 && C <= STATS_MAX_C, "bad geometry");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StatsAcc *acc = static_cast<StatsAcc *>(scratch_dev);
    const int count = n * C;
    AVB_TIMED("k6_stats", st);
    stats_init_kernel<<<(count + 127) / 128, 128, 0, st>>>(acc, count);
    const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((npx + 255) / 256, sm_count() * 8 / n + 1));
    stats_kernel<<<dim3(bx, n), 256, 0, st>>>(in_dev, npx, C, acc);
    stats_finish_kernel<<<(count + 127) / 128, 128, 0, st>>>(acc, count, (double)npx, out_dev);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_cat_warp_f32(const float *in01_dev, float *out_dev, int n, int H, int W, const float *warp_dev, avb_stream_t stream) {
    AVB_REQUIRE(in01_dev && out_dev && warp_dev, "null pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0 && in01_dev != out_dev, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CatWarpP p{in01_dev, out_dev, warp_dev, n, H, W};
    AVB_TIMED("k6_cat_warp", st);
    cat_warp_kernel<<<grid_for((long long)n * H * W), 256, 0, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_img_divide_channels(const float *in_dev, float *out_dev, int n, int64_t npx, int C, const float *stats_dev,
                                       int which, float eps, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && stats_dev, "null pointer");
    AVB_REQUIRE(n > 0 && npx > 0 && C >= 1 && C <= 4 && (which == AVB_STAT_MAX || which == AVB_STAT_MEAN), "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DivideP p{in_dev, out_dev, stats_dev, npx, (long long)n * npx * C, C, which, eps};
    AVB_TIMED("k6_divide", st);
    divide_kernel<<<grid_for(p.total), 256, 0, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int64_t avb_img_percentile_scratch_bytes(int nreq) {
    return nreq > 0 ? (int64_t)nreq * (PCT_BINS * sizeof(uint32_t) + sizeof(PctState)) + 256 : 0;
}

extern "C" int avb_img_percentile(const float *in_dev, int64_t npx, int64_t stride, const int64_t *offsets_host,
                                  const double *q_host, int nreq, float *out_dev, void *scratch_dev, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && offsets_host && q_host && out_dev && scratch_dev, "null pointer");
    AVB_REQUIRE(npx > 0 && npx < (1LL << 32) && stride >= 1 && nreq >= 1, "bad geometry");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int r0 = 0; r0 < nreq; r0 += PCT_MAX) {
        const int nr = std::min(PCT_MAX, nreq - r0);
        PctP p{};
        p.in = in_dev; p.npx = npx; p.stride = stride; p.nreq = nr;
        for (int r = 0; r < nr; ++r) {
            const double q = q_host[r0 + r];
            AVB_REQUIRE(q >= 0.0 && q <= 100.0, "percentile outside [0, 100]");
            const PctIndex pi = numpy_percentile_index(q, npx);     // float32 virtual index, as NumPy computes it
            p.off[r] = offsets_host[r0 + r];
            p.k_lo[r] = pi.k_lo; p.k_hi[r] = pi.k_hi; p.gamma[r] = pi.gamma;
        }
        p.hist = static_cast<uint32_t *>(scratch_dev);
        p.st = reinterpret_cast<PctState *>(static_cast<uint8_t *>(scratch_dev) + ((size_t)nr * PCT_BINS * sizeof(uint32_t) + 255) / 256 * 256);
        p.out = out_dev + r0;
        AVB_TIMED("k6_percentile", st);
        pct_init_kernel<<<nr, 256, 0, st>>>(p);
        const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((npx + 255) / 256, sm_count() * 8 / nr + 1));
        for (int lv = 0; lv < 3; ++lv) {
            p.level = lv;
            pct_hist_kernel<<<dim3(bx, nr), 256, 0, st>>>(p);
            pct_pick_kernel<<<nr, 1024, 0, st>>>(p);
        }
        pct_upper_kernel<<<dim3(bx, nr), 256, 0, st>>>(p);
        pct_finish_kernel<<<1, 32, 0, st>>>(p);
    }
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}
