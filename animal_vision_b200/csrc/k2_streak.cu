// K2s: the per-row "streak" acuity blur of the grazing mammals (cow, deer, goat, horse, kangaroo,
// sheep, panda, rabbit, pig), fused with decode -> 3x3 in front and [chroma compression] -> clip ->
// sRGB encode -> uint8 behind it.  6 B/px of HBM traffic, no fp32 intermediate leaves the SM.
//
// What the reference's apply_anisotropic_acuity_blur_with_streak (animals/animal_utils.py:147-172)
// actually computes (pinned bit-exact by tests/test_oracle_golden.py): every image
// row is handed to OpenCV as a W x 3 single-channel matrix, so
//   pass 1 blurs along x AND across the three colour channels with the taps of sigmaX(y),
//   pass 2 blurs along x again with the taps of sigmaY(y),
// and nothing ever mixes rows.  Rows are therefore independent 1-D problems with per-row filters.
// The host (tables.streak_row_table) folds, per row and in float64,
//   * the 3-wide REFLECT_101 channel mix and the species' dichromat matrix into ONE 3x3, and
//   * the two x passes into ONE symmetric tap vector (REFLECT_101 extension commutes with a
//     symmetric filter, so the composition is exact up to float rounding),
// so the device does: LUT decode -> per-row 3x3 -> one horizontal correlation -> tail.
#include "avb_common.cuh"

namespace avb {

constexpr int ST_TW = 512;                 // output pixels per CTA per row
constexpr int ST_RMAX = 16;                // combined radius limit (ksize <= 33)
constexpr int ST_THREADS = 192;            // 3 channels x 64 groups of 8 outputs
constexpr int ST_ROWS = 16;                // rows a CTA walks down
constexpr int ST_TAB = 56;                 // floats per row-table entry: 33 taps, 9 matrix, radius, 3x2 P_y, 2x3 Q, flag
constexpr int ST_PW = ST_TW + 2 * ST_RMAX; // produced columns
constexpr int ST_PITCH = ST_PW + 4;        // 548: 16 B aligned rows, odd multiple of 16 B
constexpr int ST_OPITCH = ST_TW + 4;

struct StreakParams {
    FrameIO io;
    const float *lut;        // decode LUT
    const uint32_t *enc;
    const float *row_tab;    // [H][ST_TAB]
    uint32_t *flags;
    int fixup;
    int chroma_on;
    float chroma_keep;       // float32(1 - strength), animal_utils.py:181
};

__device__ __forceinline__ uint32_t st_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_smem(dst)), "l"(src), "r"(bytes), "r"(st_smem(bar)) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = st_smem(bar);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}

template <int R>
__device__ __forceinline__ void streak_row_blur(const float *P, float *O, const float *taps_s, int ch, int g) {
    // outputs x = 8g .. 8g+7 of channel ch; window P[8g + (16-R) .. 8g + (16-R) + 8 + 2R)
    constexpr int OFF = ST_RMAX - R;               // multiple of 4 for R in {4,8,12,16}
    constexpr int NW4 = (8 + 2 * R) / 4;
    const float4 *w4 = reinterpret_cast<const float4 *>(P + ch * ST_PITCH + 8 * g + OFF);
    float v[4 * NW4];
#pragma unroll
    for (int q = 0; q < NW4; ++q) {
        const float4 t = w4[q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) {
        const float tk = taps_s[OFF + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(tk, v[j + k], acc[j]);
    }
    float4 *d4 = reinterpret_cast<float4 *>(O + ch * ST_OPITCH + 8 * g);
    d4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

__global__ void __launch_bounds__(ST_THREADS) streak_kernel(const __grid_constant__ StreakParams p) {
    __shared__ __align__(16) float P[3 * ST_PITCH];      // produced row (planar, with x halo)
    __shared__ __align__(16) float O[3 * ST_OPITCH];     // blurred row (planar)
    __shared__ __align__(16) uint32_t stage[ST_TW * 3 / 4];
    __shared__ float lut_s[256];
    __shared__ float tab_s[ST_TAB];
    // interior strips: the packed bytes of a row (strip + 16-px halo each side = 1632 B) are staged by the TMA
    // engine one row ahead (cp.async.bulk, mbarrier completion) instead of being gathered byte by byte
    __shared__ __align__(16) uint8_t raw_s[2][3 * ST_PW];
    __shared__ __align__(8) uint64_t rbar[2];
    __shared__ uint32_t enc_s[AVB_ENC_TABLE_MAX];

    const int frame = blockIdx.z;
    if (p.fixup && p.flags[frame] != 0) return;
    const int tid = threadIdx.x;
    const int H = p.io.H, W = p.io.W;
    const int x0 = blockIdx.x * ST_TW;
    const int y_begin = blockIdx.y * ST_ROWS, y_end = min(H, y_begin + ST_ROWS);

    for (int i = tid; i < 256; i += ST_THREADS) lut_s[i] = __ldg(p.lut + i);
    copy_to_smem(enc_s, p.enc, min((int)AVB_ENC_TABLE_MAX, ENC_HEADER + (int)__ldg(p.enc + 2)));
    __syncthreads();
    const EncTable enc = enc_view(enc_s);

    const uint8_t *src = p.io.in + (int64_t)frame * p.io.in_fs;
    uint8_t *dst = p.io.out + (int64_t)frame * p.io.out_fs;
    const int npx = min(ST_TW, W - x0);
    const bool vec_ok = (npx == ST_TW) && ((p.io.out_rs & 15) == 0) && ((p.io.out_fs & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.io.out) & 15) == 0);
    uint32_t seen = 0;
    const bool tma = x0 >= ST_RMAX && x0 + ST_TW + ST_RMAX <= W && ((p.io.in_rs & 15) == 0) && ((p.io.in_fs & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.io.in) & 15) == 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem(&rbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem(&rbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto raw_issue = [&](int y, int buf) {          // one thread
        st_bulk_load(raw_s[buf], src + (int64_t)y * p.io.in_rs + 3 * (x0 - ST_RMAX), 3 * ST_PW, &rbar[buf]);
    };
    uint32_t rphase = 0;
    if (tma && tid == 0 && y_begin < y_end) raw_issue(y_begin, 0);

    for (int y = y_begin; y < y_end; ++y) {
        if (tid < ST_TAB) tab_s[tid] = __ldg(p.row_tab + (int64_t)y * ST_TAB + tid);
        __syncthreads();      // also fences the previous row's readers of P / O / stage / raw_s
        const int buf = (y - y_begin) & 1;
        if (tma) {
            if (tid == 0 && y + 1 < y_end) raw_issue(y + 1, buf ^ 1);      // its buffer was read two barriers ago
            st_mbar_wait(&rbar[buf], (rphase >> buf) & 1u);
            rphase ^= 1u << buf;
        }
        // ---- produce: decode -> per-row 3x3 (channel mix x dichromat), or -- every dichromat matrix has
        // rank 2 -- the two planes Q lin whose 3x2 expansion P_y follows the filter (a third fewer taps)
        const bool two = tab_s[55] != 0.f;                 // uniform: the same for every row of the table
        {
            float m[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) m[i] = two ? (i < 6 ? tab_s[49 + i] : 0.f) : tab_s[33 + i];
            const uint8_t *row = src + (int64_t)y * p.io.in_rs;
            for (int i = tid; i < ST_PW; i += ST_THREADS) {
                const uint8_t *q = tma ? raw_s[buf] + 3 * i : row + 3 * reflect101(x0 - ST_RMAX + i, W);
                const uint32_t b0 = q[0], b1 = q[1], b2 = q[2];
                seen |= b0 | b1 | b2;
                const float l0 = lut_s[b0], l1 = lut_s[b1], l2 = lut_s[b2];
                P[i] = m[0] * l0 + m[1] * l1 + m[2] * l2;
                P[ST_PITCH + i] = m[3] * l0 + m[4] * l1 + m[5] * l2;
                if (!two) P[2 * ST_PITCH + i] = m[6] * l0 + m[7] * l1 + m[8] * l2;
            }
        }
        __syncthreads();
        // ---- horizontal correlation with the row's combined taps (block-uniform radius class)
        {
            const int r = (int)tab_s[42];
            const int ch = tid >> 6, g = tid & 63;
            if (!two || ch < 2) {
                if (r <= 4) streak_row_blur<4>(P, O, tab_s, ch, g);
                else if (r <= 8) streak_row_blur<8>(P, O, tab_s, ch, g);
                else if (r <= 12) streak_row_blur<12>(P, O, tab_s, ch, g);
                else streak_row_blur<16>(P, O, tab_s, ch, g);
            }
        }
        __syncthreads();
        // ---- tail: [chroma compression] -> encode; 4 pixels (12 bytes) per thread
        if (tid < ST_TW / 4) {
            const float4 a = reinterpret_cast<const float4 *>(O)[tid];
            const float4 b = reinterpret_cast<const float4 *>(O + ST_OPITCH)[tid];
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!two) c = reinterpret_cast<const float4 *>(O + 2 * ST_OPITCH)[tid];
            float v[4][3] = {{a.x, b.x, c.x}, {a.y, b.y, c.y}, {a.z, b.z, c.z}, {a.w, b.w, c.w}};
            if (two) {                                      // expand: rgb = P_y (plane0, plane1)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float p0 = v[j][0], p1 = v[j][1];
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[j][k] = fmaf(tab_s[43 + 2 * k + 1], p1, tab_s[43 + 2 * k] * p0);
                }
            }
            uint32_t bytes[12];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (p.chroma_on) {
                    // animal_utils.py:174-181: gray + (lin - gray) * (1 - strength), float32
                    const float gray = __fdiv_rn(__fadd_rn(__fadd_rn(v[j][0], v[j][1]), v[j][2]), 3.0f);
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[j][k] = __fadd_rn(gray, __fmul_rn(__fsub_rn(v[j][k], gray), p.chroma_keep));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) bytes[3 * j + k] = encode_u8(enc, v[j][k]);
            }
#pragma unroll
            for (int wd = 0; wd < 3; ++wd)
                stage[3 * tid + wd] = bytes[4 * wd] | (bytes[4 * wd + 1] << 8) | (bytes[4 * wd + 2] << 16) | (bytes[4 * wd + 3] << 24);
        }
        __syncthreads();
        // ---- store
        uint8_t *orow = dst + (int64_t)y * p.io.out_rs + (int64_t)x0 * 3;
        if (vec_ok) {
            for (int i = tid; i < ST_TW * 3 / 16; i += ST_THREADS)
                reinterpret_cast<uint4 *>(orow)[i] = reinterpret_cast<const uint4 *>(stage)[i];
        } else {
            const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage);
            for (int i = tid; i < npx * 3; i += ST_THREADS) orow[i] = sb[i];
        }
    }
    if (p.flags != nullptr && !p.fixup) {
        if (__any_sync(0xffffffffu, (seen & 0xfeu) != 0) && (tid & 31) == 0) p.flags[frame] = 1u;
    }
}

static int launch_streak(const StreakParams &p, cudaStream_t st) {
    dim3 grid((p.io.W + ST_TW - 1) / ST_TW, (p.io.H + ST_ROWS - 1) / ST_ROWS, p.io.n);
    AVB_TIMED(p.fixup ? "k2_streak_fixup" : "k2_streak", st);
    streak_kernel<<<grid, ST_THREADS, 0, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

}  // namespace avb

using namespace avb;

extern "C" int avb_streak_blur_u8(const uint8_t *in, uint8_t *out, int n, int H, int W,
                                  int64_t in_frame_stride, int64_t in_row_stride,
                                  int64_t out_frame_stride, int64_t out_row_stride,
                                  const float *dec_dev, const float *dec_raw_dev, const uint32_t *enc_dev,
                                  const float *row_tab_dev, float chroma_strength,
                                  int norm_mode, uint32_t *flags_dev, avb_stream_t stream) {
    StreakParams p{};
    p.io = FrameIO{in, out, in_frame_stride, in_row_stride, out_frame_stride, out_row_stride, n, H, W};
    AVB_REQUIRE(in && out, "null frame pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0, "bad frame geometry");
    AVB_REQUIRE(n <= 65535 && (H + ST_ROWS - 1) / ST_ROWS <= 65535, "batch or frame height too large for one launch");
    AVB_REQUIRE(in_row_stride >= 3LL * W && out_row_stride >= 3LL * W, "row stride smaller than 3*W");
    AVB_REQUIRE(dec_dev && enc_dev && row_tab_dev, "null table pointer");
    AVB_REQUIRE(chroma_strength >= 0.f && chroma_strength <= 1.f, "chroma strength out of [0,1]");
    AVB_REQUIRE(norm_mode == AVB_NORM_DIV255 || (norm_mode == AVB_NORM_AUTO && dec_raw_dev && flags_dev),
                "AVB_NORM_AUTO needs dec_raw_dev and flags_dev");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.lut = dec_dev;
    p.enc = enc_dev;
    p.row_tab = row_tab_dev;
    p.chroma_on = chroma_strength > 0.f;
    p.chroma_keep = (float)(1.0 - (double)chroma_strength);
    if (norm_mode == AVB_NORM_AUTO) {
        AVB_CUDA_OK(cudaMemsetAsync(flags_dev, 0, sizeof(uint32_t) * n, st));
        p.flags = flags_dev;
    }
    p.fixup = 0;
    if (int e = launch_streak(p, st)) return e;
    if (norm_mode == AVB_NORM_AUTO) {
        p.fixup = 1;
        p.lut = dec_raw_dev;
        if (int e = launch_streak(p, st)) return e;
    }
    return AVB_OK;
}
