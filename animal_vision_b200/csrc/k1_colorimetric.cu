// K1: fused per-pixel colorimetric kernel -- uint8 sRGB -> LUT decode -> 3x3 -> [row gain on
// channel 2] -> clip -> sRGB encode -> uint8.  HBM bound: 3 B/px in, 3 B/px out, nothing else.
//
// Packed RGB does not divide into 16-byte vectors (16 B = 5 1/3 px), so a warp moves 1536 B =
// 512 px at a time: three fully coalesced 128-bit loads per lane into a per-warp shared-memory
// slab, then each lane owns 48 contiguous bytes (16 whole pixels) of the slab, transforms them in
// registers and the slab goes back out with three coalesced 128-bit stores.
#include "avb_common.cuh"

namespace avb {

#ifndef K1_THREADS_N
#define K1_THREADS_N 256
#endif
#ifndef K1_CTAS_PER_SM
#define K1_CTAS_PER_SM 2
#endif
constexpr int K1_THREADS = K1_THREADS_N;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_UNIT = 1536;          // bytes per warp iteration
constexpr int K1_ENC_SMEM = AVB_ENC_TABLE_MAX;

struct K1Params {
    FrameIO io;
    Mat3 M;
    const float *lut;
    const uint32_t *enc;
    const float *row_gain;   // nullptr or [H]
    uint32_t *flags;
    int fixup;
    int units_per_frame;     // ceil(3*W*H / 1536)
};

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[12], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xffu; }

template <bool GAIN>
__device__ __forceinline__ void transform16(uint32_t (&w)[12], const float *lut, const EncTable &enc, const float (&m)[9],
                                            const float *row_gain, int64_t first_px, int W) {
    uint32_t o[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = 0;
    int y = 0, xr = 0;
    if (GAIN) {
        y = (int)(first_px / W);
        xr = (int)(first_px - (int64_t)y * W);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float l0 = lut[byte_of(w, 3 * j)], l1 = lut[byte_of(w, 3 * j + 1)], l2 = lut[byte_of(w, 3 * j + 2)];
        const float r0 = m[0] * l0 + m[1] * l1 + m[2] * l2;
        const float r1 = m[3] * l0 + m[4] * l1 + m[5] * l2;
        float r2 = m[6] * l0 + m[7] * l1 + m[8] * l2;
        if (GAIN) {
            // animal_utils.py:254: out[...,2] = clip(out[...,2] * w, 0, 1)
            const int yy = y + ((xr + j >= W) ? 1 : 0);   // contiguous path requires W >= 16: one wrap at most
            r2 = __fmul_rn(r2, __ldg(row_gain + yy));
        }
        const uint32_t e0 = encode_u8(enc, r0), e1 = encode_u8(enc, r1), e2 = encode_u8(enc, r2);
        o[(3 * j) >> 2] |= e0 << (8 * ((3 * j) & 3));
        o[(3 * j + 1) >> 2] |= e1 << (8 * ((3 * j + 1) & 3));
        o[(3 * j + 2) >> 2] |= e2 << (8 * ((3 * j + 2) & 3));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = o[i];
}

// ---------------------------------------------------------------- bulk-copy (TMA) helpers
__device__ __forceinline__ uint32_t k1_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k1_mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(k1_smem(bar)));
}
__device__ __forceinline__ void k1_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k1_smem(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(k1_smem(dst)), "l"(src), "r"(bytes), "r"(k1_smem(bar)) : "memory");
}
__device__ __forceinline__ void k1_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = k1_smem(bar);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void k1_bulk_store(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(k1_smem(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Fast path: frames are contiguous (row stride == 3*W) and 16-byte aligned.  Every warp runs its own
// double-buffered pipeline over 1536-byte units: the TMA engine (cp.async.bulk, completion on an
// mbarrier) fills slab i+1 while the lanes transform slab i, and finished slabs leave through bulk
// stores -- no thread ever waits on an HBM load it issued itself.  The 256-entry decode LUT is
// replicated once per shared-memory bank (index v*32 + lane), so the three decode lookups of a pixel
// are conflict-free whatever the image content (noise frames made the single copy 3-4-way conflicted).
constexpr int K1_LUT_REP = 32;
struct K1Smem {
    uint4 in[K1_WARPS][2][96];
    uint4 out[K1_WARPS][2][96];
    float lut[256 * K1_LUT_REP];
    uint32_t enc[K1_ENC_SMEM];
    uint64_t bar[K1_WARPS][2];
};

template <bool GAIN>
__device__ __forceinline__ void transform16r(uint32_t (&w)[12], const float *lut_lane, const EncTable &enc, const float (&m)[9],
                                             const float *row_gain, int64_t first_px, int W) {
    uint32_t o[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = 0;
    int y = 0, xr = 0;
    if (GAIN) {
        y = (int)(first_px / W);
        xr = (int)(first_px - (int64_t)y * W);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float l0 = lut_lane[byte_of(w, 3 * j) * K1_LUT_REP], l1 = lut_lane[byte_of(w, 3 * j + 1) * K1_LUT_REP],
                    l2 = lut_lane[byte_of(w, 3 * j + 2) * K1_LUT_REP];
        const float r0 = m[0] * l0 + m[1] * l1 + m[2] * l2;
        const float r1 = m[3] * l0 + m[4] * l1 + m[5] * l2;
        float r2 = m[6] * l0 + m[7] * l1 + m[8] * l2;
        if (GAIN) {
            const int yy = y + ((xr + j >= W) ? 1 : 0);   // contiguous path requires W >= 16: one wrap at most
            r2 = __fmul_rn(r2, __ldg(row_gain + yy));
        }
        const uint32_t e0 = encode_u8(enc, r0), e1 = encode_u8(enc, r1), e2 = encode_u8(enc, r2);
        o[(3 * j) >> 2] |= e0 << (8 * ((3 * j) & 3));
        o[(3 * j + 1) >> 2] |= e1 << (8 * ((3 * j + 1) & 3));
        o[(3 * j + 2) >> 2] |= e2 << (8 * ((3 * j + 2) & 3));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = o[i];
}

template <bool GAIN>
__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM) k1_contig_kernel(const __grid_constant__ K1Params p) {
    extern __shared__ __align__(128) uint8_t k1_dsm[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_dsm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (p.fixup) {          // second pass of AVB_NORM_AUTO: almost always nothing to redo -- leave before the tables are built
        bool todo = false;
        for (int i = lane; i < p.io.n; i += 32) todo |= p.flags[i] == 0;
        if (!__any_sync(0xffffffffu, todo)) return;
    }
    for (int i = tid; i < 256 * K1_LUT_REP; i += K1_THREADS) sm.lut[i] = __ldg(p.lut + (i >> 5));
    copy_to_smem(sm.enc, p.enc, min(K1_ENC_SMEM, ENC_HEADER + (int)__ldg(p.enc + 2)));
    if (lane == 0) {
        k1_mbar_init(&sm.bar[warp][0]);
        k1_mbar_init(&sm.bar[warp][1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const EncTable enc = enc_view(sm.enc);
    const float *lut_lane = sm.lut + lane;
    float m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = p.M.m[i];

    const int64_t frame_bytes = 3LL * p.io.W * p.io.H;
    const int64_t stride = (int64_t)gridDim.x * K1_WARPS;
    // A unit is (frame, index within the frame); positions advance incrementally (no 64-bit division per
    // unit).  `full` units go through the slab pipeline, frame tails (< 1536 B) are done in place.
    const int upf = p.units_per_frame;
    struct Pos { int frame, uif; };
    auto advance = [&](Pos &q) {
        int64_t v = (int64_t)q.uif + stride;
        while (v >= upf) { v -= upf; ++q.frame; }
        q.uif = (int)v;
    };
    auto is_full = [&](const Pos &q) { return frame_bytes - (int64_t)q.uif * K1_UNIT >= K1_UNIT; };
    auto is_skip = [&](const Pos &q) { return p.fixup && p.flags[q.frame] != 0; };
    auto prefetch = [&](const Pos &q, int slot) {       // lane 0 only
        if (q.frame >= p.io.n) return;
        if (is_full(q) && !is_skip(q))
            k1_bulk_load(sm.in[warp][slot], p.io.in + (int64_t)q.frame * p.io.in_fs + (int64_t)q.uif * K1_UNIT, K1_UNIT, &sm.bar[warp][slot]);
    };
    uint32_t seen = 0, phases = 0;                     // bit s: parity the next wait on slot s expects
    int flagged_frame = -1, slot = 0;
    Pos cur, nxt;
    {
        const int64_t u0 = (int64_t)blockIdx.x * K1_WARPS + warp;
        cur.frame = (int)(u0 / upf);
        cur.uif = (int)(u0 - (int64_t)cur.frame * upf);
    }
    if (lane == 0) prefetch(cur, 0);
    nxt = cur;
    for (; cur.frame < p.io.n; cur = nxt) {
        advance(nxt);
        const int frame = cur.frame;
        const int64_t off = (int64_t)cur.uif * K1_UNIT;
        const bool full = is_full(cur), skip = is_skip(cur);
        if (lane == 0) prefetch(nxt, slot ^ (full && !skip ? 1 : 0));     // a unit that uses no slab leaves the slot to its successor
        if (skip) continue;
        if (frame != flagged_frame) {   // flush the per-frame "byte >= 2" accumulator
            if (p.flags && !p.fixup && flagged_frame >= 0 && __any_sync(0xffffffffu, (seen & 0xfefefefeu) != 0) && lane == 0)
                p.flags[flagged_frame] = 1u;
            seen = 0;
            flagged_frame = frame;
        }
        const uint8_t *src = p.io.in + (int64_t)frame * p.io.in_fs + off;
        uint8_t *dst = p.io.out + (int64_t)frame * p.io.out_fs + off;
        if (full) {
            k1_mbar_wait(&sm.bar[warp][slot], (phases >> slot) & 1u);
            phases ^= 1u << slot;
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint4 t = sm.in[warp][slot][lane * 3 + k];
                w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) seen |= w[i];
            transform16r<GAIN>(w, lut_lane, enc, m, p.row_gain, (off + lane * 48) / 3, p.io.W);
            // the bulk store that last read out[slot] (two units ago) must have drained it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 3; ++k) sm.out[warp][slot][lane * 3 + k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) k1_bulk_store(dst, sm.out[warp][slot], K1_UNIT);
            slot ^= 1;
        } else {
            // frame tail (< 1536 B): whole pixels, one per lane-iteration
            const int npx = (int)((frame_bytes - off) / 3);
            const int64_t px0 = off / 3;
            for (int j = lane; j < npx; j += 32) {
                const uint32_t b0 = src[3 * j], b1 = src[3 * j + 1], b2 = src[3 * j + 2];
                seen |= b0 | b1 | b2;
                const float l0 = lut_lane[b0 * K1_LUT_REP], l1 = lut_lane[b1 * K1_LUT_REP], l2 = lut_lane[b2 * K1_LUT_REP];
                const float r0 = m[0] * l0 + m[1] * l1 + m[2] * l2;
                const float r1 = m[3] * l0 + m[4] * l1 + m[5] * l2;
                float r2 = m[6] * l0 + m[7] * l1 + m[8] * l2;
                if (GAIN) r2 = __fmul_rn(r2, __ldg(p.row_gain + (int)((px0 + j) / p.io.W)));
                dst[3 * j] = (uint8_t)encode_u8(enc, r0);
                dst[3 * j + 1] = (uint8_t)encode_u8(enc, r1);
                dst[3 * j + 2] = (uint8_t)encode_u8(enc, r2);
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // slabs stay valid until read out
    if (p.flags && !p.fixup && flagged_frame >= 0 && __any_sync(0xffffffffu, (seen & 0xfefefefeu) != 0) && lane == 0)
        p.flags[flagged_frame] = 1u;
}

// Generic path: arbitrary strides / alignment, one pixel per thread-iteration.
template <bool GAIN>
__global__ void __launch_bounds__(K1_THREADS) k1_strided_kernel(const __grid_constant__ K1Params p) {
    __shared__ float lut_s[256];
    __shared__ uint32_t enc_s[K1_ENC_SMEM];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += K1_THREADS) lut_s[i] = __ldg(p.lut + i);
    copy_to_smem(enc_s, p.enc, min(K1_ENC_SMEM, ENC_HEADER + (int)__ldg(p.enc + 2)));
    __syncthreads();
    const EncTable enc = enc_view(enc_s);
    const int frame = blockIdx.z;
    if (p.fixup && p.flags[frame] != 0) return;
    const uint8_t *src = p.io.in + (int64_t)frame * p.io.in_fs;
    uint8_t *dst = p.io.out + (int64_t)frame * p.io.out_fs;
    uint32_t seen = 0;
    for (int y = blockIdx.y; y < p.io.H; y += gridDim.y) {
        const float g = GAIN ? __ldg(p.row_gain + y) : 1.f;
        for (int x = blockIdx.x * K1_THREADS + tid; x < p.io.W; x += gridDim.x * K1_THREADS) {
            const uint8_t *q = src + (int64_t)y * p.io.in_rs + 3 * x;
            const uint32_t b0 = q[0], b1 = q[1], b2 = q[2];
            seen |= b0 | b1 | b2;
            const float l0 = lut_s[b0], l1 = lut_s[b1], l2 = lut_s[b2];
            const float r0 = p.M.m[0] * l0 + p.M.m[1] * l1 + p.M.m[2] * l2;
            const float r1 = p.M.m[3] * l0 + p.M.m[4] * l1 + p.M.m[5] * l2;
            float r2 = p.M.m[6] * l0 + p.M.m[7] * l1 + p.M.m[8] * l2;
            if (GAIN) r2 = __fmul_rn(r2, g);
            uint8_t *o = dst + (int64_t)y * p.io.out_rs + 3 * x;
            o[0] = (uint8_t)encode_u8(enc, r0);
            o[1] = (uint8_t)encode_u8(enc, r1);
            o[2] = (uint8_t)encode_u8(enc, r2);
        }
    }
    if (p.flags && !p.fixup && __any_sync(0xffffffffu, (seen & 0xfeu) != 0) && (tid & 31) == 0) p.flags[frame] = 1u;
}

static int launch_k1(const K1Params &p, cudaStream_t st) {
    const FrameIO &io = p.io;
    AVB_TIMED(p.fixup ? "k1_colorimetric_fixup" : "k1_colorimetric", st);
    const bool contig = io.W >= 16 && io.in_rs == 3LL * io.W && io.out_rs == 3LL * io.W && (io.in_fs & 15) == 0 && (io.out_fs & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(io.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(io.out) & 15) == 0;
    if (contig) {
        const int64_t total_units = (int64_t)p.units_per_frame * io.n;
        int64_t blocks = (total_units + K1_WARPS - 1) / K1_WARPS;
        const int64_t cap = (int64_t)sm_count() * K1_CTAS_PER_SM;   // persistent: every resident CTA walks its share of the units
        if (blocks > cap) blocks = cap;
        static SmemOptIn opt_gain, opt_plain;
        AVB_CUDA_OK(opt_gain.ensure(k1_contig_kernel<true>, (int)sizeof(K1Smem)));
        AVB_CUDA_OK(opt_plain.ensure(k1_contig_kernel<false>, (int)sizeof(K1Smem)));
        if (p.row_gain) k1_contig_kernel<true><<<(unsigned)blocks, K1_THREADS, sizeof(K1Smem), st>>>(p);
        else k1_contig_kernel<false><<<(unsigned)blocks, K1_THREADS, sizeof(K1Smem), st>>>(p);
    } else {
        dim3 grid((io.W + K1_THREADS - 1) / K1_THREADS, io.H < 1024 ? io.H : 1024, io.n);
        if (p.row_gain) k1_strided_kernel<true><<<grid, K1_THREADS, 0, st>>>(p);
        else k1_strided_kernel<false><<<grid, K1_THREADS, 0, st>>>(p);
    }
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

}  // namespace avb

using namespace avb;

extern "C" int avb_colorimetric_u8(const uint8_t *in, uint8_t *out, int n, int H, int W,
                                   int64_t in_frame_stride, int64_t in_row_stride,
                                   int64_t out_frame_stride, int64_t out_row_stride,
                                   const float *dec_dev, const float *dec_raw_dev, const uint32_t *enc_dev,
                                   const float *m_host, const float *row_gain_dev,
                                   int norm_mode, uint32_t *flags_dev, avb_stream_t stream) {
    K1Params p{};
    p.io = FrameIO{in, out, in_frame_stride, in_row_stride, out_frame_stride, out_row_stride, n, H, W};
    AVB_REQUIRE(in && out, "null frame pointer");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0, "bad frame geometry");
    AVB_REQUIRE(in_row_stride >= 3LL * W && out_row_stride >= 3LL * W, "row stride smaller than 3*W");
    AVB_REQUIRE(dec_dev && enc_dev && m_host, "null table pointer");
    AVB_REQUIRE(norm_mode == AVB_NORM_DIV255 || (norm_mode == AVB_NORM_AUTO && dec_raw_dev && flags_dev),
                "AVB_NORM_AUTO needs dec_raw_dev and flags_dev");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < 9; ++i) p.M.m[i] = m_host[i];
    p.lut = dec_dev;
    p.enc = enc_dev;
    p.row_gain = row_gain_dev;
    p.units_per_frame = (int)((3LL * W * H + K1_UNIT - 1) / K1_UNIT);
    if (norm_mode == AVB_NORM_AUTO) {
        AVB_CUDA_OK(cudaMemsetAsync(flags_dev, 0, sizeof(uint32_t) * n, st));
        p.flags = flags_dev;
    }
    p.fixup = 0;
    if (int e = launch_k1(p, st)) return e;
    if (norm_mode == AVB_NORM_AUTO) {
        p.fixup = 1;
        p.lut = dec_raw_dev;
        if (int e = launch_k1(p, st)) return e;
    }
    return AVB_OK;
}
