// K7: fused per-pixel programs for the UV species + cv2.remap on float32 frames.
//
// The fifteen UV species and MantisShrimp (animals/reindeer.py, goldfish.py, ... mantis_shrimp.py) are chains of
// 100-200 NumPy element-wise operations between a handful of spatial operators (resize, blur, Sobel, remap) and
// global reductions (min / max / percentile).  The reference materialises a full-frame float32 temporary per
// operation; here every maximal run of element-wise operations is ONE launch: the host (animal_vision_b200/lazy.py)
// records the species' arithmetic as an expression DAG, linearises it into a register program and this kernel
// executes the program per pixel -- operands come straight from the source planes / per-row and per-column tables /
// per-frame device scalars (reduction results never visit the host), results go straight to the output planes
// (float32, or the final uint8 frame).  Every instruction is one IEEE float32 operation in the reference's order
// (no contraction, no re-association), so the result equals NumPy's up to the library transcendental functions.
//
// The register file lives in shared memory ([reg][thread], conflict free), instructions are fetched with uniform
// loads; a program is a few hundred bytes and is cached on the device by content.
#include <algorithm>

#include "avb_common.cuh"

namespace avb {

constexpr int VM_THREADS = 256;

struct VmParams {
    const avb_vm_ins *prog;
    int n_ins, n_regs;
    int n, H, W;
    avb_vm_src src[AVB_VM_MAX_SRC];
    avb_vm_dst dst[AVB_VM_MAX_DST];
};

// x^e for x in (0, 1], e > 0 on the SFU: ex2(e * lg2(x)).  lg2.approx is good to 2^-22 absolute on this range, so the
// relative error of the power is <= e * ln2 * 2^-22 ~ 4e-7 (e = 2.4): far inside the 1e-5 budget of the float
// intermediates, and 100x cheaper than libm powf, which dominated the programs (six sRGB conversions per pixel).
__device__ __forceinline__ float vm_pow_unit(float x, float e) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e * l));
    return r;
}
__device__ __forceinline__ float vm_srgb_dec(float s) {      // uv_helpers.py:33-37 / classic_rgb_to_hsi.py:16-22
    if (s <= 0.04045f) return __fdiv_rn(s, 12.92f);
    const float b = __fdiv_rn(__fadd_rn(s, 0.055f), 1.055f);
    return b <= 1.0f ? vm_pow_unit(b, 2.4f) : powf(b, 2.4f);     // bicubic overshoot above 1: the library routine
}
__device__ __forceinline__ float vm_srgb_enc(float l) {      // uv_helpers.py:40-44 (1/2.4 acts as a float32 scalar)
    if (l <= 0.0031308f) return __fmul_rn(l, 12.92f);
    const float pw = l <= 1.0f ? vm_pow_unit(l, 0.41666666f) : powf(l, 0.41666666f);
    return __fsub_rn(__fmul_rn(1.055f, pw), 0.055f);
}

// Every thread runs the program on VM_PX pixels at once (pixel j of a thread is VM_THREADS apart from pixel j-1, so each
// of the VM_PX loads / stores of an instruction is coalesced across the warp): fetch + decode + dispatch are paid once
// per VM_PX results.  Register file r[reg][j][thread] and the program itself live in shared memory.
constexpr int VM_PX = 4;

template <class F>
__device__ __forceinline__ void vm_un(float *r, int dst, int a, F f) {
#pragma unroll
    for (int j = 0; j < VM_PX; ++j) r[(dst * VM_PX + j) * VM_THREADS] = f(r[(a * VM_PX + j) * VM_THREADS]);
}
template <class F>
__device__ __forceinline__ void vm_bin(float *r, int dst, int a, int b, F f) {
#pragma unroll
    for (int j = 0; j < VM_PX; ++j) r[(dst * VM_PX + j) * VM_THREADS] = f(r[(a * VM_PX + j) * VM_THREADS], r[(b * VM_PX + j) * VM_THREADS]);
}

__global__ void __launch_bounds__(VM_THREADS) vm_kernel(const __grid_constant__ VmParams p) {
    extern __shared__ __align__(16) float vm_smem[];       // [n_ins] uint2 program | [n_regs][VM_PX][VM_THREADS] registers
    uint2 *prog = reinterpret_cast<uint2 *>(vm_smem);
    float *r = vm_smem + 2 * p.n_ins + threadIdx.x;
    for (int i = threadIdx.x; i < p.n_ins; i += VM_THREADS) prog[i] = __ldg(reinterpret_cast<const uint2 *>(p.prog) + i);
    __syncthreads();
    const long long npx = (long long)p.H * p.W, total = npx * p.n;
    for (long long base = (long long)blockIdx.x * (VM_THREADS * VM_PX); base < total; base += (long long)gridDim.x * (VM_THREADS * VM_PX)) {
        int frame[VM_PX], y[VM_PX], x[VM_PX];
        long long pix[VM_PX];
        bool live[VM_PX];
#pragma unroll
        for (int j = 0; j < VM_PX; ++j) {
            long long gi = base + j * VM_THREADS + threadIdx.x;
            live[j] = gi < total;
            gi = live[j] ? gi : total - 1;                  // dead lanes recompute the last pixel and do not store
            frame[j] = (int)(gi / npx);
            pix[j] = gi - (long long)frame[j] * npx;
            y[j] = (int)(pix[j] / p.W);
            x[j] = (int)(pix[j] - (long long)y[j] * p.W);
        }
        for (int pc = 0; pc < p.n_ins; ++pc) {
            const uint2 raw = prog[pc];
            const int op = raw.x & 0xffu, dst = (raw.x >> 8) & 0xffu, a = (raw.x >> 16) & 0xffu, b = raw.x >> 24;
            switch (op) {
                case AVB_VM_LOAD: {
                    const avb_vm_src &s = p.src[a];
                    const float *ptr = static_cast<const float *>(s.ptr);
#pragma unroll
                    for (int j = 0; j < VM_PX; ++j) {
                        long long idx;
                        switch (s.kind) {
                            case AVB_VM_SRC_PLANE: idx = (long long)frame[j] * s.frame_stride + pix[j] * s.pix_stride + b; break;
                            case AVB_VM_SRC_ROW: idx = (long long)y[j] * s.pix_stride + b; break;
                            case AVB_VM_SRC_COL: idx = (long long)x[j] * s.pix_stride + b; break;
                            default: idx = (long long)frame[j] * s.frame_stride + b; break;      // per-frame scalars
                        }
                        r[(dst * VM_PX + j) * VM_THREADS] = __ldg(ptr + idx);
                    }
                    break;
                }
                case AVB_VM_CONST: {
                    const float v = __uint_as_float(raw.y);
#pragma unroll
                    for (int j = 0; j < VM_PX; ++j) r[(dst * VM_PX + j) * VM_THREADS] = v;
                    break;
                }
                case AVB_VM_MOV: vm_un(r, dst, a, [](float u) { return u; }); break;
                case AVB_VM_ADD: vm_bin(r, dst, a, b, [](float u, float v) { return __fadd_rn(u, v); }); break;
                case AVB_VM_SUB: vm_bin(r, dst, a, b, [](float u, float v) { return __fsub_rn(u, v); }); break;
                case AVB_VM_MUL: vm_bin(r, dst, a, b, [](float u, float v) { return __fmul_rn(u, v); }); break;
                case AVB_VM_DIV: vm_bin(r, dst, a, b, [](float u, float v) { return __fdiv_rn(u, v); }); break;
                case AVB_VM_MIN: vm_bin(r, dst, a, b, [](float u, float v) { return fminf(u, v); }); break;
                case AVB_VM_MAX: vm_bin(r, dst, a, b, [](float u, float v) { return fmaxf(u, v); }); break;
                case AVB_VM_POW: vm_bin(r, dst, a, b, [](float u, float v) { return powf(u, v); }); break;
                case AVB_VM_ATAN2: vm_bin(r, dst, a, b, [](float u, float v) { return atan2f(u, v); }); break;
                case AVB_VM_GT: vm_bin(r, dst, a, b, [](float u, float v) { return u > v ? 1.f : 0.f; }); break;
                case AVB_VM_GE: vm_bin(r, dst, a, b, [](float u, float v) { return u >= v ? 1.f : 0.f; }); break;
                case AVB_VM_LT: vm_bin(r, dst, a, b, [](float u, float v) { return u < v ? 1.f : 0.f; }); break;
                case AVB_VM_LE: vm_bin(r, dst, a, b, [](float u, float v) { return u <= v ? 1.f : 0.f; }); break;
                case AVB_VM_NEG: vm_un(r, dst, a, [](float u) { return -u; }); break;
                case AVB_VM_ABS: vm_un(r, dst, a, [](float u) { return fabsf(u); }); break;
                case AVB_VM_SQRT: vm_un(r, dst, a, [](float u) { return __fsqrt_rn(u); }); break;
                case AVB_VM_EXP: vm_un(r, dst, a, [](float u) { return expf(u); }); break;
                case AVB_VM_SIN: vm_un(r, dst, a, [](float u) { return sinf(u); }); break;
                case AVB_VM_COS: vm_un(r, dst, a, [](float u) { return cosf(u); }); break;
                case AVB_VM_FLOOR: vm_un(r, dst, a, [](float u) { return floorf(u); }); break;
                case AVB_VM_SRGB_DEC: vm_un(r, dst, a, [](float u) { return vm_srgb_dec(u); }); break;
                case AVB_VM_SRGB_ENC: vm_un(r, dst, a, [](float u) { return vm_srgb_enc(u); }); break;
                case AVB_VM_QUANT:                                                         // uv_helpers.py:26-30
                    vm_un(r, dst, a, [](float u) { return truncf(fminf(fmaxf(__fadd_rn(__fmul_rn(u, 255.0f), 0.5f), 0.f), 255.f)); });
                    break;
                case AVB_VM_ADDI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fadd_rn(u, m); }); break; }
                case AVB_VM_SUBI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fsub_rn(u, m); }); break; }
                case AVB_VM_RSUBI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fsub_rn(m, u); }); break; }
                case AVB_VM_MULI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fmul_rn(u, m); }); break; }
                case AVB_VM_DIVI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fdiv_rn(u, m); }); break; }
                case AVB_VM_RDIVI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return __fdiv_rn(m, u); }); break; }
                case AVB_VM_MINI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return fminf(u, m); }); break; }
                case AVB_VM_MAXI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return fmaxf(u, m); }); break; }
                case AVB_VM_POWI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return powf(u, m); }); break; }
                case AVB_VM_GTI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return u > m ? 1.f : 0.f; }); break; }
                case AVB_VM_GEI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return u >= m ? 1.f : 0.f; }); break; }
                case AVB_VM_LTI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return u < m ? 1.f : 0.f; }); break; }
                case AVB_VM_LEI: { const float m = __uint_as_float(raw.y); vm_un(r, dst, a, [m](float u) { return u <= m ? 1.f : 0.f; }); break; }
                case AVB_VM_SELECT: {
                    const int c = raw.y & 0xffu;
#pragma unroll
                    for (int j = 0; j < VM_PX; ++j)
                        r[(dst * VM_PX + j) * VM_THREADS] = r[(a * VM_PX + j) * VM_THREADS] != 0.f ? r[(b * VM_PX + j) * VM_THREADS] : r[(c * VM_PX + j) * VM_THREADS];
                    break;
                }
                case AVB_VM_STORE: {
                    const avb_vm_dst &d = p.dst[b];
#pragma unroll
                    for (int j = 0; j < VM_PX; ++j) {
                        if (!live[j]) continue;
                        const long long idx = (long long)frame[j] * d.frame_stride + (long long)y[j] * d.row_stride + (long long)x[j] * d.pix_stride + raw.y;
                        const float val = r[(a * VM_PX + j) * VM_THREADS];
                        if (d.kind == AVB_VM_DST_U8) static_cast<uint8_t *>(d.ptr)[idx] = (uint8_t)val;
                        else static_cast<float *>(d.ptr)[idx] = val;
                    }
                    break;
                }
                default: break;
            }
        }
    }
}

// ---- cv2.remap(INTER_LINEAR, BORDER_REFLECT_101) on float32 frames with float32 maps (animals/anableps.py:224-237).
// OpenCV quantises the map to 1/32 px (INTER_BITS = 5): sx = rint(32 x), integer part sx >> 5, fraction (sx & 31) / 32,
// the four taps weighted by the float products of the 1-D weights.
struct RemapP {
    const float *in;
    float *out;
    const float *mx, *my;
    int n, H, W, C;
};
__global__ void __launch_bounds__(256) remap_kernel(const __grid_constant__ RemapP p) {
    const long long npx = (long long)p.H * p.W, total = npx * p.n;
    for (long long gi = (long long)blockIdx.x * 256 + threadIdx.x; gi < total; gi += (long long)gridDim.x * 256) {
        const int frame = (int)(gi / npx);
        const long long pix = gi - (long long)frame * npx;
        const int sx = __float2int_rn(__ldg(p.mx + pix) * 32.0f), sy = __float2int_rn(__ldg(p.my + pix) * 32.0f);
        const int ix = sx >> 5, iy = sy >> 5;
        const float fx = (float)(sx & 31) * (1.0f / 32.0f), fy = (float)(sy & 31) * (1.0f / 32.0f);
        const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx), w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
        const int x0 = reflect101(ix, p.W), x1 = reflect101(ix + 1, p.W), y0 = reflect101(iy, p.H), y1 = reflect101(iy + 1, p.H);
        const float *f = p.in + (long long)frame * npx * p.C;
        float *o = p.out + gi * p.C;
        for (int c = 0; c < p.C; ++c) {
            const float a = f[((long long)y0 * p.W + x0) * p.C + c], b = f[((long long)y0 * p.W + x1) * p.C + c];
            const float d = f[((long long)y1 * p.W + x0) * p.C + c], e = f[((long long)y1 * p.W + x1) * p.C + c];
            o[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, w00), __fmul_rn(b, w01)), __fmul_rn(d, w10)), __fmul_rn(e, w11));
        }
    }
}

}  // namespace avb

using namespace avb;

extern "C" int avb_vm_run(const avb_vm_ins *prog_dev, int n_ins, int n_regs, int n, int H, int W,
                          const avb_vm_src *src_host, int n_src, const avb_vm_dst *dst_host, int n_dst, avb_stream_t stream) {
    AVB_REQUIRE(prog_dev && n_ins > 0 && n_ins <= AVB_VM_MAX_INS, "bad program");
    AVB_REQUIRE(n_regs > 0 && n_regs <= AVB_VM_MAX_REGS, "bad register count");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0, "bad frame geometry");
    AVB_REQUIRE(n_src >= 0 && n_src <= AVB_VM_MAX_SRC && n_dst > 0 && n_dst <= AVB_VM_MAX_DST, "too many sources / destinations");
    AVB_REQUIRE((n_src == 0 || src_host) && dst_host, "null descriptor table");
    VmParams p{};
    p.prog = prog_dev; p.n_ins = n_ins; p.n_regs = n_regs; p.n = n; p.H = H; p.W = W;
    for (int i = 0; i < n_src; ++i) {
        AVB_REQUIRE(src_host[i].ptr && src_host[i].kind >= AVB_VM_SRC_PLANE && src_host[i].kind <= AVB_VM_SRC_FRAME, "bad source descriptor");
        p.src[i] = src_host[i];
    }
    for (int i = 0; i < n_dst; ++i) {
        AVB_REQUIRE(dst_host[i].ptr && (dst_host[i].kind == AVB_VM_DST_F32 || dst_host[i].kind == AVB_VM_DST_U8), "bad destination descriptor");
        p.dst[i] = dst_host[i];
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)n_ins * sizeof(uint2) + (size_t)n_regs * VM_PX * VM_THREADS * sizeof(float);
    static SmemOptIn optin;
    const int smem_max = AVB_VM_MAX_INS * (int)sizeof(uint2) + AVB_VM_MAX_REGS * VM_PX * VM_THREADS * (int)sizeof(float);
    if (smem > 48 * 1024) AVB_CUDA_OK(optin.ensure(vm_kernel, smem_max));
    const long long total = (long long)n * H * W, tiles = (total + VM_THREADS * VM_PX - 1) / (VM_THREADS * VM_PX);
    const int per_sm = std::max(1, std::min(8, (int)((220 * 1024) / (smem + 1024))));
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(tiles, (long long)sm_count() * per_sm));
    AVB_TIMED("k7_vm", st);
    vm_kernel<<<grid, VM_THREADS, smem, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}

extern "C" int avb_img_remap(const float *in_dev, float *out_dev, int n, int H, int W, int C, const float *mapx_dev,
                             const float *mapy_dev, avb_stream_t stream) {
    AVB_REQUIRE(in_dev && out_dev && mapx_dev && mapy_dev, "null pointer");
    AVB_REQUIRE(in_dev != out_dev, "remap cannot run in place");
    AVB_REQUIRE(n > 0 && H > 0 && W > 0 && C > 0 && C <= 16, "bad geometry");
    RemapP p{in_dev, out_dev, mapx_dev, mapy_dev, n, H, W, C};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)n * H * W;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm_count() * 16));
    AVB_TIMED("k7_remap", st);
    remap_kernel<<<grid, 256, 0, st>>>(p);
    AVB_CUDA_OK(cudaGetLastError());
    return AVB_OK;
}
