// K3: the UV path -- RGB -> analytic 31-band spectrum -> photoreceptor catches -> von Kries
// adaptation -> small acuity blur -> opponent (HSV-like) mapping with two GLOBAL 95th percentiles
// -> sRGB encode, without ever materialising the H x W x 31 hyperspectral cube
// (257 MB per 1080p frame in the reference: classic_rgb_to_hsi.py:47-82, honeybee.py:126-135).
//
// The global statistics force several passes over the frame; every pass RE-COMPUTES the receptor
// catches from the uint8 input (3 B/px, L2-resident after the first pass) instead of round-tripping
// fp32 planes through HBM:
//   A  catches -> per-frame max (white patch) / sum (gray world)
//   B1 adapted + blurred catches -> (radius, L) -> histogram of float bits [30:20]
//   B2 ... bits [19:9] of the values sharing the rank's 11-bit prefix
//   B3 ... bits [8:0]  -> the two order statistics around rank 0.95 (N-1), exact -> np.percentile
//   C  catches -> hue/sat/val -> HSV->RGB -> clip -> OETF -> uint8
// (a one-CTA "scan" kernel between the histogram passes turns counts into the next prefix).
//
// Receptor catches come either from the per-pixel 31-band sum in registers ("bands" mode, the
// reference's own order of operations) or from the algebraically identical 3x3 (the whole chain
// lobes -> illuminant -> sensitivities is linear; SURVEY.md 8a-11: 1.9e-7 relative difference).
#include "avb_common.cuh"

namespace avb {

constexpr int UV_TW = 64, UV_TH = 16, UV_THREADS = 256;
constexpr int UV_MAX_BANDS = 160;
constexpr int UV_BINS = 2048;
constexpr int UV_MAX_BLUR_R = 2;

// per-frame statistics block in the caller's workspace
struct UvFrameStats {
    uint32_t max_bits[3];        // white patch: max of each raw catch (non-negative floats as bits)
    uint32_t pad0;
    double sum[3];               // gray world
    uint32_t prefix[2][2];       // [quantity][lo/hi rank]: bits fixed so far
    uint32_t remaining[2][2];    // rank within the current prefix
    float pct[2];                // the two percentiles (radius, L)
    uint32_t pad1[2];
    uint32_t hist1[2][UV_BINS];          // level 1: [quantity]
    uint32_t hist23[2][2][2][UV_BINS];   // levels 2,3: [level-2][quantity][lo/hi]
};

struct UvParams {
    FrameIO io;
    const float *lut;            // decode LUT (device, 256)
    float M3[9];                 // collapsed receptor matrix: catch k = sum_c M3[3k+c] * lin[c]
    const float *bands;          // bands mode: [B][8] = g0,g1,g2 (lobe of input channel c), E, s0,s1,s2 (sensitivities), 0
    int n_bands;                 // 0 -> collapsed mode
    float denom_eps;             // lobe normaliser + 1e-8 (float32, as torch computes it)
    int adapt;                   // 0 none, 1 white patch (max), 2 gray world (mean)
    float eps;                   // 1e-8
    int blur_r;                  // 0..2
    float blur_taps[2 * UV_MAX_BLUR_R + 1];
    UvFrameStats *stats;         // [n]
    const uint32_t *enc;
    long long k_lo, k_hi;        // order-statistic ranks of the percentile
    double gamma;                // interpolation weight
    float *dbg_catches;          // optional [n][H][W][3] raw catches (test hook), else nullptr
};

// ------------------------------------------------------------------ per-pixel receptor catches
struct Catcher {
    const float *lut_s;
    const UvParams *p;
    __device__ __forceinline__ void raw(const uint8_t *q, float &u, float &b, float &g) const {
        const float c0 = lut_s[q[0]], c1 = lut_s[q[1]], c2 = lut_s[q[2]];
        if (p->n_bands == 0) {
            u = p->M3[0] * c0 + p->M3[1] * c1 + p->M3[2] * c2;
            b = p->M3[3] * c0 + p->M3[4] * c1 + p->M3[5] * c2;
            g = p->M3[6] * c0 + p->M3[7] * c1 + p->M3[8] * c2;
        } else {
            // classic_rgb_to_hsi.py:70-78, honeybee.py:126-135 in the reference's own order:
            // spec = (g2*c2 + g1*c1 + g0*c0) / (denom+1e-8);  rad = spec * E;  catch += rad * s
            float au = 0.f, ab = 0.f, ag = 0.f;
            const float4 *t = reinterpret_cast<const float4 *>(p->bands);
            for (int l = 0; l < p->n_bands; ++l) {
                const float4 lo = __ldg(t + 2 * l), hi = __ldg(t + 2 * l + 1);
                float spec = __fadd_rn(__fadd_rn(__fmul_rn(lo.z, c2), __fmul_rn(lo.y, c1)), __fmul_rn(lo.x, c0));
                spec = fmaxf(__fdiv_rn(spec, p->denom_eps), 0.f);
                const float rad = __fmul_rn(spec, lo.w);
                au = fmaf(rad, hi.x, au);
                ab = fmaf(rad, hi.y, ab);
                ag = fmaf(rad, hi.z, ag);
            }
            u = au; b = ab; g = ag;
        }
    }
};

__device__ __forceinline__ void adapt_scales(const UvParams &p, const UvFrameStats &st, long long npx, float (&w)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (p.adapt == 1) w[k] = fmaxf(__uint_as_float(st.max_bits[k]), p.eps);          // uv_helpers.py:195-199
        else if (p.adapt == 2) w[k] = fmaxf((float)(st.sum[k] / (double)npx), p.eps);    // uv_helpers.py:202-206
        else w[k] = 1.0f;
    }
}

// ------------------------------------------------------------------ pass A: maxima / sums
__global__ void __launch_bounds__(256) uv_stats_kernel(const __grid_constant__ UvParams p) {
    __shared__ float lut_s[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_s[i] = __ldg(p.lut + i);
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *src = p.io.in + (int64_t)frame * p.io.in_fs;
    Catcher cat{lut_s, &p};
    float mx[3] = {0.f, 0.f, 0.f};
    double sm[3] = {0.0, 0.0, 0.0};
    const int W = p.io.W;
    const long long npx = (long long)p.io.H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        float u, b, g;
        cat.raw(src + (int64_t)y * p.io.in_rs + 3 * x, u, b, g);
        if (p.dbg_catches) {
            float *d = p.dbg_catches + ((int64_t)frame * npx + i) * 3;
            d[0] = u; d[1] = b; d[2] = g;
        }
        mx[0] = fmaxf(mx[0], u); mx[1] = fmaxf(mx[1], b); mx[2] = fmaxf(mx[2], g);
        sm[0] += u; sm[1] += b; sm[2] += g;
    }
    UvFrameStats &st = p.stats[frame];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float m = mx[k];
        double s = sm[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (p.adapt == 1) atomicMax(&st.max_bits[k], __float_as_uint(fmaxf(m, 0.f)));
            if (p.adapt == 2) atomicAdd(&st.sum[k], s);
        }
    }
}

// ------------------------------------------------------------------ tile machinery for passes B*, C
// A CTA owns a UV_TW x UV_TH tile; adapted catches for the tile plus a blur_r halo (REFLECT_101)
// go to shared memory, then every thread blurs and maps its pixels.
struct UvTile {
    float *pl;       // [3][TH+2r][TW+2r] planes
    int pw, ph, r;
};

template <int PASS>   // 1,2,3: histogram level; 4: map + encode
__global__ void __launch_bounds__(UV_THREADS) uv_tile_kernel(const __grid_constant__ UvParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *lut_s = reinterpret_cast<float *>(smem_raw);
    float *pl = lut_s + 256;
    const int r = p.blur_r;
    const int pw = UV_TW + 2 * r, ph = UV_TH + 2 * r;
    uint32_t *tail = reinterpret_cast<uint32_t *>(pl + 3 * pw * ph);   // histograms (B) or encode table + stage (C)

    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int H = p.io.H, W = p.io.W;
    const int x0 = blockIdx.x * UV_TW, y0 = blockIdx.y * UV_TH;
    const uint8_t *src = p.io.in + (int64_t)frame * p.io.in_fs;
    UvFrameStats &st = p.stats[frame];
    const long long npx = (long long)H * W;

    for (int i = tid; i < 256; i += UV_THREADS) lut_s[i] = __ldg(p.lut + i);
    constexpr int NH = (PASS == 1) ? 2 : 4;
    if (PASS <= 3) {
        for (int i = tid; i < NH * UV_BINS; i += UV_THREADS) tail[i] = 0;
    } else {
        copy_to_smem(tail, p.enc, min((int)AVB_ENC_TABLE_MAX, ENC_HEADER + (int)__ldg(p.enc + 2)));
    }
    __syncthreads();

    float ws[3];
    adapt_scales(p, st, npx, ws);
    Catcher cat{lut_s, &p};
    for (int i = tid; i < pw * ph; i += UV_THREADS) {
        const int ty = i / pw, tx = i - ty * pw;
        const int y = reflect101(y0 - r + ty, H), x = reflect101(x0 - r + tx, W);
        float u, b, g;
        cat.raw(src + (int64_t)y * p.io.in_rs + 3 * x, u, b, g);
        pl[i] = __fdiv_rn(u, ws[0]);
        pl[pw * ph + i] = __fdiv_rn(b, ws[1]);
        pl[2 * pw * ph + i] = __fdiv_rn(g, ws[2]);
    }
    __syncthreads();

    float pr = 0.f, pL = 0.f;
    EncTable enc{};
    uint8_t *stage = nullptr;
    if (PASS == 4) {
        pr = st.pct[0] + p.eps;      // uv_mappers.py:61-62: percentile + eps, float32
        pL = st.pct[1] + p.eps;
        enc = enc_view(tail);
        stage = reinterpret_cast<uint8_t *>(tail + AVB_ENC_TABLE_MAX);
    }
    uint32_t pre[2][2];
    if (PASS == 2 || PASS == 3) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int t = 0; t < 2; ++t) pre[q][t] = st.prefix[q][t];
    }

    for (int i = tid; i < UV_TW * UV_TH; i += UV_THREADS) {
        const int ty = i / UV_TW, tx = i - ty * UV_TW;
        const int y = y0 + ty, x = x0 + tx;
        const bool inside = (y < H) && (x < W);
        float c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float *q = pl + k * pw * ph + (ty + r) * pw + (tx + r);
            if (r == 0) {
                c[k] = q[0];
            } else {
                // separable correlation, rows (x) first then columns, as cv2.GaussianBlur does
                float acc = 0.f;
                for (int dy = -r; dy <= r; ++dy) {
                    float row = 0.f;
                    for (int dx = -r; dx <= r; ++dx) row = fmaf(p.blur_taps[dx + r], q[dy * pw + dx], row);
                    acc = fmaf(p.blur_taps[dy + r], row, acc);
                }
                c[k] = acc;
            }
        }
        // uv_mappers.py:53-60
        const float U = c[0], B = c[1], G = c[2];
        const float O1 = G - B, O2 = B - U;
        const float L = __fdiv_rn(__fadd_rn(__fadd_rn(U, B), G), 3.0f);
        const float radius = sqrtf(__fadd_rn(__fmul_rn(O1, O1), __fmul_rn(O2, O2)));
        if (PASS <= 3) {
            if (inside) {
                const uint32_t bits[2] = {__float_as_uint(fmaxf(radius, 0.f)), __float_as_uint(fmaxf(L, 0.f))};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (PASS == 1) {
                        atomicAdd(&tail[q * UV_BINS + (bits[q] >> 20)], 1u);
                    } else {
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (PASS == 2) {
                                if ((bits[q] >> 20) == pre[q][t]) atomicAdd(&tail[(q * 2 + t) * UV_BINS + ((bits[q] >> 9) & 0x7ffu)], 1u);
                            } else {
                                if ((bits[q] >> 9) == pre[q][t]) atomicAdd(&tail[(q * 2 + t) * UV_BINS + (bits[q] & 0x1ffu)], 1u);
                            }
                        }
                    }
                }
            }
        } else {
            // hue / sat / val and hsv_to_rgb (uv_mappers.py:14-26, :57-64)
            const float PI_F = 3.14159274101257324f;          // float32(np.pi)
            const float hue = __fdiv_rn(__fadd_rn(atan2f(O2, O1), PI_F), 6.28318548202514648f);
            const float sat = __saturatef(__fdiv_rn(radius, pr));
            const float val = __saturatef(__fdiv_rn(L, pL));
            const float h6 = __fmul_rn(hue, 6.0f);
            const float fl = floorf(h6);
            const float f = h6 - fl;
            int sext = (int)fl % 6;
            if (sext < 0) sext += 6;
            // NumPy promotes f = h*6 - int32 to float64, so q and t are float64 products rounded
            // once by the final astype(float32); p stays float32 (uv_mappers.py:18-21, :64)
            const float pp_ = __fmul_rn(val, __fsub_rn(1.0f, sat));
            const float qq = (float)((double)val * (1.0 - (double)f * (double)sat));
            const float tt = (float)((double)val * (1.0 - (1.0 - (double)f) * (double)sat));
            float R_, G_, B_;
            switch (sext) {
                case 0: R_ = val; G_ = tt; B_ = pp_; break;
                case 1: R_ = qq; G_ = val; B_ = pp_; break;
                case 2: R_ = pp_; G_ = val; B_ = tt; break;
                case 3: R_ = pp_; G_ = qq; B_ = val; break;
                case 4: R_ = tt; G_ = pp_; B_ = val; break;
                default: R_ = val; G_ = pp_; B_ = qq; break;
            }
            stage[i * 3 + 0] = (uint8_t)encode_u8(enc, R_);
            stage[i * 3 + 1] = (uint8_t)encode_u8(enc, G_);
            stage[i * 3 + 2] = (uint8_t)encode_u8(enc, B_);
        }
    }
    __syncthreads();

    if (PASS <= 3) {
        uint32_t *gh = (PASS == 1) ? &st.hist1[0][0] : &st.hist23[PASS - 2][0][0][0];
        for (int i = tid; i < NH * UV_BINS; i += UV_THREADS) {
            const uint32_t v = tail[i];
            if (v) atomicAdd(gh + i, v);
        }
    } else {
        uint8_t *dst = p.io.out + (int64_t)frame * p.io.out_fs;
        const bool vec_ok = (x0 + UV_TW <= W) && ((p.io.out_rs & 15) == 0) && ((p.io.out_fs & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.io.out) & 15) == 0);
        if (vec_ok) {
            constexpr int VPR = UV_TW * 3 / 16;
            for (int i = tid; i < UV_TH * VPR; i += UV_THREADS) {
                const int ty = i / VPR, q = i - ty * VPR;
                if (y0 + ty < H)
                    *reinterpret_cast<uint4 *>(dst + (int64_t)(y0 + ty) * p.io.out_rs + (int64_t)x0 * 3 + q * 16) =
                        reinterpret_cast<const uint4 *>(stage + ty * UV_TW * 3)[q];
            }
        } else {
            const int nbytes = min(UV_TW, W - x0) * 3;
            for (int i = tid; i < UV_TH * UV_TW * 3; i += UV_THREADS) {
                const int ty = i / (UV_TW * 3), b = i - ty * (UV_TW * 3);
                if (y0 + ty < H && b < nbytes) dst[(int64_t)(y0 + ty) * p.io.out_rs + (int64_t)x0 * 3 + b] = stage[i];
            }
        }
    }
}

// ------------------------------------------------------------------ scan: counts -> next prefix
// One CTA per frame.  LEVEL 1: from hist1 pick, for each quantity and each of the two ranks, the
// bin holding that rank.  LEVEL 2/3: same inside hist23.  After LEVEL 3 the full 31-bit patterns
// of both order statistics are known and the percentile is their linear interpolation
// (numpy.percentile, method "linear").
template <int LEVEL>
__global__ void __launch_bounds__(256) uv_scan_kernel(const __grid_constant__ UvParams p) {
    UvFrameStats &st = p.stats[blockIdx.x];
    __shared__ uint32_t part[256];
    __shared__ uint32_t found_bin[2][2], found_rem[2][2];
    const int tid = threadIdx.x;
    constexpr int PER = UV_BINS / 256;
    for (int q = 0; q < 2; ++q)
        for (int t = 0; t < 2; ++t) {
            const uint32_t *h = (LEVEL == 1) ? st.hist1[q] : st.hist23[LEVEL - 2][q][t];
            const uint32_t rank = (LEVEL == 1) ? (uint32_t)(t == 0 ? p.k_lo : p.k_hi) : st.remaining[q][t];
            uint32_t loc[PER], s = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) { loc[j] = h[tid * PER + j]; s += loc[j]; }
            part[tid] = s;
            __syncthreads();
            // exclusive prefix over the 256 partial sums (serial in one thread: 256 adds, negligible)
            if (tid == 0) {
                uint32_t run = 0;
                for (int i = 0; i < 256; ++i) { const uint32_t v = part[i]; part[i] = run; run += v; }
            }
            __syncthreads();
            uint32_t run = part[tid];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                if (rank >= run && rank < run + loc[j]) { found_bin[q][t] = tid * PER + j; found_rem[q][t] = rank - run; }
                run += loc[j];
            }
            __syncthreads();
        }
    if (tid < 4) {
        const int q = tid >> 1, t = tid & 1;
        const uint32_t bin = found_bin[q][t];
        if (LEVEL == 1) st.prefix[q][t] = bin;
        else if (LEVEL == 2) st.prefix[q][t] = (st.prefix[q][t] << 11) | bin;
        else st.prefix[q][t] = (st.prefix[q][t] << 9) | bin;
        st.remaining[q][t] = found_rem[q][t];
    }
    __syncthreads();
    if (LEVEL == 3 && tid < 2) {
        const double a = (double)__uint_as_float(st.prefix[tid][0]), b = (double)__uint_as_float(st.prefix[tid][1]);
        st.pct[tid] = (float)(a + (b - a) * p.gamma);
    }
}

static size_t uv_tile_smem(int r, int pass) {
    const int pw = UV_TW + 2 * r, ph = UV_TH + 2 * r;
    size_t s = (256 + 3 * (size_t)pw * ph) * 4;
    if (pass == 1) s += 2 * UV_BINS * 4;
    else if (pass <= 3) s += 4 * UV_BINS * 4;
    else s += AVB_ENC_TABLE_MAX * 4 + UV_TW * UV_TH * 3;
    return s;
}

template <int PASS>
static int launch_tile(const UvParams &p, cudaStream_t st) {
    const size_t smem = uv_tile_smem(p.blur_r, PASS);
    AVB_CUDA_OK(cudaFuncSetAttribute(uv_tile_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.io.W + UV_TW - 1) / UV_TW, (p.io.H + UV_TH - 1) / UV_TH, p.io.n);
    static const char *names[5] = {"", "k3_uv_hist1", "k3_uv_hist2", "k3_uv_hist3", "k3_uv_map"};
    AVB_TIMED(names[PASS], st);
    uv_tile_kernel<PASS><<<grid, UV_THREADS, smem, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

}  // namespace avb

using namespace avb;

extern "C" int64_t avb_uv_workspace_bytes(int n) { return n > 0 ? (int64_t)n * (int64_t)sizeof(UvFrameStats) : 0; }

extern "C" int avb_uv_opponent_u8(const uint8_t *in, uint8_t *out, int n, int H, int W,
                                  int64_t in_frame_stride, int64_t in_row_stride,
                                  int64_t out_frame_stride, int64_t out_row_stride,
                                  const float *dec_dev, const uint32_t *enc_dev,
                                  const float *m3_host, const float *bands_dev, int n_bands, float denom_eps,
                                  int adapt_mode, const float *blur_taps_host, int blur_ksize, float percentile,
                                  void *workspace_dev, float *dbg_catches_dev, avb_stream_t stream) {
    UvParams p{};
    p.io = FrameIO{in, out, in_frame_stride, in_row_stride, out_frame_stride, out_row_stride, n, H, W};
    AVB_REQUIRE(in && out, "null frame pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0, "bad frame geometry");
    AVB_REQUIRE(in_row_stride >= 3LL * W && out_row_stride >= 3LL * W, "row stride smaller than 3*W");
    AVB_REQUIRE(dec_dev && enc_dev && m3_host && workspace_dev, "null table / workspace pointer");
    AVB_REQUIRE(n_bands >= 0 && n_bands <= UV_MAX_BANDS && (n_bands == 0 || bands_dev), "bad band table");
    AVB_REQUIRE(adapt_mode >= 0 && adapt_mode <= 2, "adapt_mode must be 0 (none), 1 (white patch) or 2 (gray world)");
    AVB_REQUIRE(blur_ksize == 0 || ((blur_ksize & 1) && blur_ksize <= 2 * UV_MAX_BLUR_R + 1 && blur_taps_host),
                "blur ksize must be 0, 3 or 5");
    AVB_REQUIRE(percentile >= 0.f && percentile <= 100.f, "percentile out of range");
    AVB_REQUIRE((long long)H * W < (1LL << 31), "frame too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.lut = dec_dev;
    p.enc = enc_dev;
    for (int i = 0; i < 9; ++i) p.M3[i] = m3_host[i];
    p.bands = bands_dev;
    p.n_bands = n_bands;
    p.denom_eps = denom_eps;
    p.adapt = adapt_mode;
    p.eps = 1e-8f;
    p.blur_r = blur_ksize / 2;
    for (int i = 0; i < blur_ksize; ++i) p.blur_taps[i] = blur_taps_host[i];
    p.stats = static_cast<UvFrameStats *>(workspace_dev);
    p.dbg_catches = dbg_catches_dev;
    // numpy.percentile(method="linear"): virtual index q/100 * (N-1)
    const long long npx = (long long)H * W;
    const double vi = ((double)percentile / 100.0) * (double)(npx - 1);
    p.k_lo = (long long)vi;
    p.k_hi = p.k_lo + 1 < npx ? p.k_lo + 1 : p.k_lo;
    p.gamma = vi - (double)p.k_lo;

    AVB_CUDA_OK(cudaMemsetAsync(workspace_dev, 0, sizeof(UvFrameStats) * (size_t)n, st));
    if (adapt_mode != 0 || dbg_catches_dev) {
        const long long blocks = (npx + 256 * 8 - 1) / (256 * 8);
        dim3 grid((unsigned)(blocks < 4096 ? blocks : 4096), n);
        AVB_TIMED("k3_uv_stats", st);
        uv_stats_kernel<<<grid, 256, 0, st>>>(p);
        AVB_CUDA_OK(cudaGetLastError());
    }
    if (int e = launch_tile<1>(p, st)) return e;
    { AVB_TIMED("k3_uv_scan", st); uv_scan_kernel<1><<<n, 256, 0, st>>>(p); }
    if (int e = launch_tile<2>(p, st)) return e;
    { AVB_TIMED("k3_uv_scan", st); uv_scan_kernel<2><<<n, 256, 0, st>>>(p); }
    if (int e = launch_tile<3>(p, st)) return e;
    { AVB_TIMED("k3_uv_scan", st); uv_scan_kernel<3><<<n, 256, 0, st>>>(p); }
    AVB_CUDA_OK(cudaGetLastError());
    return launch_tile<4>(p, st);
}
