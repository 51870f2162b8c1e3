// Float-frame path of the dichromat mammals: "float in => float [0,1] out, no quantisation"
// (reference animals/dog.py:56-59; animals/animal_utils.py:41-50 get_normalized_image).
//
// The renderers only ever produce uint8 frames (renderers/video.py:95, cv2.imread), which take the
// fused K1 / K2 / K2s kernels.  A caller that hands `visualize` a float (or a wider integer) frame
// gets the same recipe in plain fp32, one straightforward kernel per step -- no LUTs, because a
// float frame has no 256-entry alphabet:
//   frame max (the data-dependent "/255 only if max > 1" branch)  ->  normalise, clip, sRGB decode
//   (powf), 3x3 [per-row 3x3 for the streak species], [S-cone row gain]  ->  separable Gaussian
//   (REFLECT_101, OpenCV taps) or the per-row streak filter  ->  [chroma compression], clip, sRGB
//   encode (powf), clip [, x*255+0.5 truncation for integer callers].
// Frames are packed float32 [n, H, W, 3].  Parity bar: <= 1e-5 relative on the fp32 result.
#include <algorithm>

#include "avb_common.cuh"

namespace avb {
namespace f32path {

constexpr int TAB = 56;          // floats per streak row-table entry (tables.streak_row_table)
constexpr int TAB_CENTRE = 16;   // index of the centre tap

__device__ __forceinline__ float srgb_decode(float v) {          // animal_utils.py:5-11, float32 like NumPy
    return v <= 0.04045f ? __fdiv_rn(v, 12.92f) : powf(__fdiv_rn(v + 0.055f, 1.055f), 2.4f);
}
__device__ __forceinline__ float srgb_encode(float v) {          // animal_utils.py:13-19 (1/2.4 as a float32 scalar)
    return v <= 0.0031308f ? 12.92f * v : __fmaf_rn(1.055f, powf(v, 0.41666666f), -0.055f);
}
__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ int reflect(int i, int n) { return reflect101(i, n); }

// ---- per-frame maximum: bit pattern of max(v, 0) (non-negative floats order like unsigned ints)
__global__ void __launch_bounds__(256) frame_max_kernel(const float *__restrict__ in, long long per_frame, uint32_t *maxbits) {
    const float *f = in + (long long)blockIdx.y * per_frame;
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < per_frame; i += (long long)gridDim.x * 256) m = fmaxf(m, f[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(maxbits + blockIdx.y, __float_as_uint(m));
}

struct ProduceP {
    const float *in; float *out;
    int n, H, W;
    float M[9];
    const float *row_gain;      // [H] or null (Rat, animal_utils.py:206-259)
    const float *row_tab;       // [H][TAB] or null: per-row 3x3 at [33..42) (streak species)
    const uint32_t *maxbits;    // [n]
    int final_encode;           // 1: no spatial filter follows -> clip + encode here
    int quantize;
};

__global__ void __launch_bounds__(256) produce_kernel(const __grid_constant__ ProduceP p) {
    const long long npx = (long long)p.n * p.H * p.W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npx; i += (long long)gridDim.x * 256) {
        const int frame = (int)(i / ((long long)p.H * p.W));
        const int y = (int)((i / p.W) % p.H);
        const bool div = __uint_as_float(p.maxbits[frame]) > 1.0f;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = p.in[3 * i + c];
            if (div) s = __fdiv_rn(s, 255.0f);
            v[c] = srgb_decode(clip01(s));
        }
        const float *m = p.row_tab ? p.row_tab + (long long)y * TAB + 33 : p.M;
        float o0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
        float o1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
        float o2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
        if (p.row_gain) o2 = clip01(__fmul_rn(o2, p.row_gain[y]));
        if (p.final_encode) {
            o0 = clip01(srgb_encode(clip01(o0))); o1 = clip01(srgb_encode(clip01(o1))); o2 = clip01(srgb_encode(clip01(o2)));
            if (p.quantize) { o0 = truncf(o0 * 255.0f + 0.5f); o1 = truncf(o1 * 255.0f + 0.5f); o2 = truncf(o2 * 255.0f + 0.5f); }
        }
        p.out[3 * i] = o0; p.out[3 * i + 1] = o1; p.out[3 * i + 2] = o2;
    }
}

// ---- separable Gaussian on packed float32 HWC3, REFLECT_101 (cv2.GaussianBlur on CV_32FC3)
struct BlurP {
    const float *in; float *out;
    int n, H, W, R, vertical;
    float taps[33];
};
__global__ void __launch_bounds__(256) blur_kernel(const __grid_constant__ BlurP p) {
    const long long total = (long long)p.n * p.H * p.W * 3;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % 3);
        const long long px = e / 3;
        const int x = (int)(px % p.W), y = (int)((px / p.W) % p.H);
        const long long fbase = (px / ((long long)p.H * p.W)) * p.H * p.W;
        float acc = 0.f;
        for (int k = -p.R; k <= p.R; ++k) {
            const long long q = p.vertical ? fbase + (long long)reflect(y + k, p.H) * p.W + x : fbase + (long long)y * p.W + reflect(x + k, p.W);
            acc = fmaf(p.taps[k + p.R], p.in[3 * q + c], acc);
        }
        p.out[e] = acc;
    }
}

// ---- per-row streak filter (rows independent; taps centred at TAB_CENTRE, radius at [42])
struct StreakP {
    const float *in; float *out; const float *row_tab;
    int n, H, W;
};
__global__ void __launch_bounds__(256) streak_kernel(const __grid_constant__ StreakP p) {
    const long long total = (long long)p.n * p.H * p.W * 3;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int c = (int)(e % 3);
        const long long px = e / 3;
        const int x = (int)(px % p.W), y = (int)((px / p.W) % p.H);
        const long long rbase = (px / p.W) * p.W;
        const float *tab = p.row_tab + (long long)y * TAB;
        const int R = (int)tab[42];
        float acc = 0.f;
        for (int k = -R; k <= R; ++k) acc = fmaf(tab[TAB_CENTRE + k], p.in[3 * (rbase + reflect(x + k, p.W)) + c], acc);
        p.out[e] = acc;
    }
}

// ---- tail: [chroma compression, animal_utils.py:174-181] -> clip -> OETF -> clip [-> quantise]
struct TailP {
    const float *in; float *out; long long npx;
    int chroma_on; float chroma_keep; int quantize;
};
__global__ void __launch_bounds__(256) tail_kernel(const __grid_constant__ TailP p) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.npx; i += (long long)gridDim.x * 256) {
        float v0 = p.in[3 * i], v1 = p.in[3 * i + 1], v2 = p.in[3 * i + 2];
        if (p.chroma_on) {
            // gray = mean over channels (float32), out = gray + (x - gray) * (1 - strength)
            const float g = __fdiv_rn(__fadd_rn(__fadd_rn(v0, v1), v2), 3.0f);
            v0 = __fadd_rn(g, __fmul_rn(__fsub_rn(v0, g), p.chroma_keep));
            v1 = __fadd_rn(g, __fmul_rn(__fsub_rn(v1, g), p.chroma_keep));
            v2 = __fadd_rn(g, __fmul_rn(__fsub_rn(v2, g), p.chroma_keep));
        }
        v0 = clip01(srgb_encode(clip01(v0))); v1 = clip01(srgb_encode(clip01(v1))); v2 = clip01(srgb_encode(clip01(v2)));
        if (p.quantize) { v0 = truncf(v0 * 255.0f + 0.5f); v1 = truncf(v1 * 255.0f + 0.5f); v2 = truncf(v2 * 255.0f + 0.5f); }
        p.out[3 * i] = v0; p.out[3 * i + 1] = v1; p.out[3 * i + 2] = v2;
    }
}

static unsigned grid_for(long long items) {
    const long long want = (items + 255) / 256, cap = (long long)sm_count() * 16;
    return (unsigned)std::max<long long>(1, std::min(want, cap));
}

}  // namespace f32path
}  // namespace avb

using namespace avb;
using namespace avb::f32path;

extern "C" int avb_dichromat_f32(const float *in, float *out, float *tmp, int n, int H, int W, const float *m_host,
                                 int kind, const float *taps_host, int ksize, const float *row_tab_dev,
                                 const float *row_gain_dev, float chroma, int quantize, uint32_t *maxbits_dev,
                                 avb_stream_t stream) {
    AVB_REQUIRE(in && out && tmp && m_host && maxbits_dev, "null pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0, "bad frame geometry");
    AVB_REQUIRE(kind == AVB_F32_POINT || kind == AVB_F32_GAUSS || kind == AVB_F32_STREAK, "unknown kind");
    AVB_REQUIRE(kind != AVB_F32_GAUSS || (taps_host && ksize >= 1 && ksize <= 33 && (ksize & 1)), "Gaussian needs an odd tap count <= 33");
    AVB_REQUIRE(kind != AVB_F32_STREAK || row_tab_dev, "streak needs the per-row table");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long npx = (long long)n * H * W;
    AVB_CUDA_OK(cudaMemsetAsync(maxbits_dev, 0, sizeof(uint32_t) * n, st));
    {
        AVB_TIMED("k5_frame_max", st);
        const long long per = (long long)H * W * 3;
        frame_max_kernel<<<dim3((unsigned)std::max<long long>(1, std::min<long long>((per + 255) / 256, sm_count() * 8 / n + 1)), n), 256, 0, st>>>(in, per, maxbits_dev);
    }
    ProduceP pp{};
    pp.in = in; pp.n = n; pp.H = H; pp.W = W;
    for (int i = 0; i < 9; ++i) pp.M[i] = m_host[i];
    pp.row_gain = row_gain_dev; pp.row_tab = kind == AVB_F32_STREAK ? row_tab_dev : nullptr; pp.maxbits = maxbits_dev;
    pp.final_encode = kind == AVB_F32_POINT; pp.quantize = quantize;
    pp.out = kind == AVB_F32_POINT ? out : tmp;
    {
        AVB_TIMED("k5_produce", st);
        produce_kernel<<<grid_for(npx), 256, 0, st>>>(pp);
    }
    if (kind == AVB_F32_POINT) {
        AVB_CUDA_OK(cudaGetLastError());
        return AVB_OK;
    }
    const float *filtered;
    if (kind == AVB_F32_GAUSS) {
        BlurP bp{};
        bp.n = n; bp.H = H; bp.W = W; bp.R = ksize / 2;
        for (int i = 0; i < ksize; ++i) bp.taps[i] = taps_host[i];
        AVB_TIMED("k5_gauss", st);
        bp.in = tmp; bp.out = out; bp.vertical = 0;
        blur_kernel<<<grid_for(npx * 3), 256, 0, st>>>(bp);
        bp.in = out; bp.out = tmp; bp.vertical = 1;
        blur_kernel<<<grid_for(npx * 3), 256, 0, st>>>(bp);
        filtered = tmp;
    } else {
        StreakP sp{tmp, out, row_tab_dev, n, H, W};
        AVB_TIMED("k5_streak", st);
        streak_kernel<<<grid_for(npx * 3), 256, 0, st>>>(sp);
        filtered = out;
    }
    TailP tp{filtered, out, npx, chroma != 0.f, (float)(1.0 - (double)chroma), quantize};
    {
        AVB_TIMED("k5_tail", st);
        tail_kernel<<<grid_for(npx), 256, 0, st>>>(tp);
    }
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}
