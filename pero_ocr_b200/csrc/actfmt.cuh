// Activation record formats shared by every kernel that produces or consumes an NHWC activation tensor.
//
// A pixel (or sequence frame) with C channels is one contiguous record:
//   ACT_F16       [C x fp16]                                     2C bytes   single-pass fp16 operands
//   ACT_F16_HILO  [C x fp16 hi][C x fp16 lo]                     4C bytes   lo = fp16(x - hi); 3 fp16 MMA passes
//   ACT_F16_F8    [C x fp16 hi][C x e5m2 lo'][C x e5m2 hi8]      4C bytes   lo' = e5m2((x - hi) * 2^11), hi8 = e5m2(hi)
//
// ACT_F16_F8 is the operand format of the "fp16 + fp8 correction" precision: a product a*b with a = ah + al,
// b = bh + bl is evaluated as ah*bh on the fp16 tensor pipe plus the two first-order correction terms al*bh + ah*bl
// on the fp8 pipe at twice the rate.  The correction terms are ~2^-11 of the product, so the 3-bit e5m2 significand
// leaves a relative error of ~2^-14 -- 8x below single-pass fp16 -- for 2 pass-equivalents instead of 3.
// The accumulator is kept at scale 2^11: fp16 weights are stored as bh * 2^11, the fp8 planes as e5m2(bh) and
// e5m2(bl * 2^11) (engine.cu: build_gemm), and every epilogue multiplies by 2^-11 before the bias.
// In the K dimension of the fp8 pass the activation row is [lo'(C) | hi8(C)] and the weight row is
// [e5m2(bh)(C) | e5m2(bl * 2^11)(C)], i.e. one contraction of length 2C.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

enum ActFormat : int { ACT_F16 = 0, ACT_F16_HILO = 1, ACT_F16_F8 = 2 };

constexpr float kF8Scale = 2048.f;  // 2^11: one fp16 ulp of a value in [1, 2) maps to 1.0 .. 2.0 in the lo' plane

__host__ __device__ inline int act_planes(int fmt) { return fmt == ACT_F16 ? 1 : 2; }  // fp16-sized planes per record

__host__ __device__ inline uint8_t f32_to_e5m2(float v) {
    return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2));
}
__host__ __device__ inline float e5m2_to_f32(uint8_t b) {
    // e5m2 is the upper byte of an IEEE fp16
    const __half_raw hr = {static_cast<unsigned short>(static_cast<unsigned short>(b) << 8)};
    return __half2float(__half(hr));
}

#ifdef __CUDACC__
// two floats -> packed e5m2x2 (low byte = first value)
__device__ __forceinline__ uint32_t pack_e5m2x2(float a, float b) {
    return static_cast<uint32_t>(__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E5M2));
}
__device__ __forceinline__ uint32_t pack_e5m2x4(float a, float b, float c, float d) {
    return pack_e5m2x2(a, b) | (pack_e5m2x2(c, d) << 16);
}

// Scalar store of channel i of a C-channel record (slow paths: LayerNorm, attention, cross-check kernels).
__device__ __forceinline__ void act_store(__half* rec, int C, int i, float v, int fmt) {
    const __half hi = __float2half_rn(v);
    rec[i] = hi;
    if (fmt == ACT_F16_HILO) {
        rec[C + i] = __float2half_rn(v - __half2float(hi));
    } else if (fmt == ACT_F16_F8) {
        uint8_t* b = reinterpret_cast<uint8_t*>(rec + C);
        b[i] = f32_to_e5m2((v - __half2float(hi)) * kF8Scale);
        b[C + i] = f32_to_e5m2(__half2float(hi));
    }
}
// Value represented by channel i of a record (debug read-back).
__device__ __forceinline__ float act_load(const __half* rec, int C, int i, int fmt) {
    float v = __half2float(rec[i]);
    if (fmt == ACT_F16_HILO) v += __half2float(rec[C + i]);
    else if (fmt == ACT_F16_F8) v += e5m2_to_f32(reinterpret_cast<const uint8_t*>(rec + C)[i]) * (1.f / kF8Scale);
    return v;
}

// Tensor-core epilogue: 32 consecutive channels [n0, n0 + 32) of one pixel.  r = raw fp32 accumulators (after any
// max-pool), out = acc * acc_scale + bias -> activation -> optional affine -> record planes, 16-byte stores.
__device__ __forceinline__ float epi_act(float v, int act, float slope) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : slope * v;
    return v;
}
// r -> packed record words of the 32 channels: ph = 16 words fp16 hi; pl = 16 words fp16 lo (HILO) or 8 words lo' +
// 8 words hi8 (F8).
__device__ __forceinline__ void epi_pack32(const uint32_t (&r)[32], int n0, float acc_scale, const float* s_bias,
                                           const float* s_scale, const float* s_shift, bool has_affine, int act,
                                           float slope, int fmt, uint32_t (&ph)[16], uint32_t (&pl)[16]) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + j);
        float v[4] = {fmaf(__uint_as_float(r[j]), acc_scale, b4.x), fmaf(__uint_as_float(r[j + 1]), acc_scale, b4.y),
                      fmaf(__uint_as_float(r[j + 2]), acc_scale, b4.z), fmaf(__uint_as_float(r[j + 3]), acc_scale, b4.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = epi_act(v[e], act, slope);
        if (has_affine) {
            const float4 a4 = *reinterpret_cast<const float4*>(s_scale + n0 + j);
            const float4 s4 = *reinterpret_cast<const float4*>(s_shift + n0 + j);
            v[0] = fmaf(v[0], a4.x, s4.x); v[1] = fmaf(v[1], a4.y, s4.y);
            v[2] = fmaf(v[2], a4.z, s4.z); v[3] = fmaf(v[3], a4.w, s4.w);
        }
        const __half2 h01 = __floats2half2_rn(v[0], v[1]);
        const __half2 h23 = __floats2half2_rn(v[2], v[3]);
        ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&h01);
        ph[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        if (fmt != ACT_F16) {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            if (fmt == ACT_F16_HILO) {
                const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y);
                const __half2 l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
                pl[j >> 1] = *reinterpret_cast<const uint32_t*>(&l01);
                pl[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l23);
            } else {
                pl[j >> 2] = pack_e5m2x4((v[0] - f01.x) * kF8Scale, (v[1] - f01.y) * kF8Scale,
                                         (v[2] - f23.x) * kF8Scale, (v[3] - f23.y) * kF8Scale);
                pl[8 + (j >> 2)] = pack_e5m2x4(f01.x, f01.y, f23.x, f23.y);
            }
        }
    }
}

__device__ __forceinline__ void epi_store32(const uint32_t (&r)[32], int n0, float acc_scale, const float* s_bias,
                                            const float* s_scale, const float* s_shift, bool has_affine, int act,
                                            float slope, __half* orow, int cout, int fmt) {
    uint32_t ph[16], pl[16];
    epi_pack32(r, n0, acc_scale, s_bias, s_scale, s_shift, has_affine, act, slope, fmt, ph, pl);
    uint4* dst = reinterpret_cast<uint4*>(orow + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
    if (fmt == ACT_F16_HILO) {
        uint4* dl = reinterpret_cast<uint4*>(orow + cout + n0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dl[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
    } else if (fmt == ACT_F16_F8) {
        uint8_t* b = reinterpret_cast<uint8_t*>(orow + cout);
        uint4* dlo = reinterpret_cast<uint4*>(b + n0);
        uint4* dh8 = reinterpret_cast<uint4*>(b + cout + n0);
        dlo[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        dlo[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
        dh8[0] = make_uint4(pl[8], pl[9], pl[10], pl[11]);
        dh8[1] = make_uint4(pl[12], pl[13], pl[14], pl[15]);
    }
}

// 256-bit variant (st.global.v8.b32, sm_100): in epi_store32 one store instruction touches 32 records a record-stride
// apart and fills half a 32-byte sector of each; here every instruction fills whole sectors.  (Round 1 bought the same
// with a detour through shared memory; measured against both, this is the fastest: profiles/r02u_store_ab.json.)
// Needs 32-byte aligned record pieces: n0 % 32 == 0 and cout % 32 == 0 (16 for the fp16 planes).
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void epi_store32_v8(const uint32_t (&r)[32], int n0, float acc_scale, const float* s_bias,
                                               const float* s_scale, const float* s_shift, bool has_affine, int act,
                                               float slope, __half* orow, int cout, int fmt, bool skip_lo = false) {
    uint32_t ph[16], pl[16];
    epi_pack32(r, n0, acc_scale, s_bias, s_scale, s_shift, has_affine, act, slope, fmt, ph, pl);
    uint8_t* rec = reinterpret_cast<uint8_t*>(orow);
    st_global_v8(rec + n0 * 2, ph);
    st_global_v8(rec + n0 * 2 + 32, ph + 8);
    if (fmt == ACT_F16_HILO) {
        st_global_v8(rec + cout * 2 + n0 * 2, pl);
        st_global_v8(rec + cout * 2 + n0 * 2 + 32, pl + 8);
    } else if (fmt == ACT_F16_F8) {
        if (!skip_lo) st_global_v8(rec + cout * 2 + n0, pl);      // lo': unread by a weight-side-only consumer
        st_global_v8(rec + cout * 3 + n0, pl + 8);
    }
}

#endif
