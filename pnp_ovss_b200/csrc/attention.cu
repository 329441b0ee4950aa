// attention.cu -- ViT self-attention softmax(Q K^T / sqrt(d)) V at fp32-grade accuracy on the fp16 tensor cores (VIT:93-119, the
// attention of every encoder block; SURVEY 8f.1: the model pass).
//
// Same idea as the 3xFP16 GEMM operands (tf32x3.cu): every fp32 operand is split exactly into x = h + l 2^-11 (h = fp16(x),
// l = fp16((x - h) 2^11)) and each product a b is evaluated as  a_h b_h + 2^-11 (a_l b_h + a_h b_l)  -- three fp16
// mma.sync.m16n8k16 with fp32 accumulation, the main term and the two corrections in separate accumulators (the tensor core
// adds with truncation: the 2^-11-sized corrections must not sit in the main chain).  That holds for both contractions:
// S = Q K^T and O = P V, with the probabilities P split the same way after the fp32 online softmax.
// Two passes.  (1) q, k, v are split once into fp16 (hi, lo) pairs in a workspace (q pre-scaled for an exp2 softmax).
// (2) Flash-attention structure: one CTA = 64 query rows of one (image, head), 4 warps x 16 rows; K and V tiles of 64 keys
// stream through a double-buffered cp.async pipeline in shared memory; nothing of size L x L touches HBM.
// HBM traffic: q, k, v read once, their splits written once and read once (K/V tiles hit in L2 across the 7 query tiles of a
// head), O written once.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace pnp {

constexpr int kAttD = 64;          // head dimension (ViT-L/16: 1024 / 16)
constexpr int kAttTile = 64;       // keys per shared-memory tile = query rows per CTA
constexpr int kAttPitch = 72;      // halves per shared row: 144 B, 16-byte aligned, conflict-free for ldmatrix and quad LDS.32

__device__ __forceinline__ void mma_16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(unsigned &r0, unsigned &r1, unsigned &r2, unsigned &r3, const __half *row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"((unsigned)__cvta_generic_to_shared(row_ptr)));
}

__device__ __forceinline__ void ldmatrix_x4_trans(unsigned &r0, unsigned &r1, unsigned &r2, unsigned &r3, const __half *row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"((unsigned)__cvta_generic_to_shared(row_ptr)));
}

__device__ __forceinline__ void ldmatrix_x2_trans(unsigned &r0, unsigned &r1, const __half *row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r0), "=r"(r1)
                 : "r"((unsigned)__cvta_generic_to_shared(row_ptr)));
}

// (x0, x1) -> packed half2 of the hi parts and of the 2^11-scaled lo parts (packed cvt.rn.f16x2.f32 conversions)
__device__ __forceinline__ void split2(float x0, float x1, unsigned &hi, unsigned &lo) {
    const __half2 hh = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn((x0 - hf.x) * 2048.0f, (x1 - hf.y) * 2048.0f);
    hi = *reinterpret_cast<const unsigned *>(&hh);
    lo = *reinterpret_cast<const unsigned *>(&ll);
}

// ---- pass 1: fp32 q, k, v -> fp16 (hi, lo) pairs, [3, 2, B, H, Lp, 64] with Lp = L rounded up to the tile (rows past L zero).
// q is pre-multiplied by softmax_scale * log2(e) (the softmax runs on exp2), k and v by in_scale (a power of two: exact).
// Done once per (image, head) instead of once per query tile inside the main kernel (7 query tiles share every K/V tile).
__global__ void __launch_bounds__(256) attention_split_kernel(const float *__restrict__ qkv, __half *__restrict__ ws, int L, int Lp, int H,
                                                              int B, float in_scale, int *__restrict__ flag) {
    // grid = (Lp, B): one CTA per token row; its k and v parts are 2 x H x 64 contiguous floats of qkv.  No per-item divisions.
    const size_t plane = (size_t)B * H * Lp * kAttD;                 // one [B,H,Lp,64] fp16 array
    const int l = blockIdx.x, b = blockIdx.y;
    const int per_which = H * (kAttD / 4);                           // float4 pieces of the k (or v) part of one row
    bool bad = false;
    for (int i = threadIdx.x; i < 2 * per_which; i += blockDim.x) {
        const int which = 1 + (i >= per_which), j = i - (which - 1) * per_which;   // 1 k, 2 v (q is split by the main kernel)
        const int h = j >> 4, c4 = (j & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < L) v = ldg_stream4(qkv + (((size_t)b * L + l) * 3 + which) * H * kAttD + 4 * j);
        v.x *= in_scale; v.y *= in_scale; v.z *= in_scale; v.w *= in_scale;
        bad = bad || !(fabsf(v.x) <= 65504.f && fabsf(v.y) <= 65504.f && fabsf(v.z) <= 65504.f && fabsf(v.w) <= 65504.f);
        unsigned h01, l01, h23, l23;
        split2(v.x, v.y, h01, l01);
        split2(v.z, v.w, h23, l23);
        const size_t o = (((size_t)b * H + h) * Lp + l) * kAttD + c4;
        *reinterpret_cast<uint2 *>(ws + (size_t)(2 * which) * plane + o) = make_uint2(h01, h23);
        *reinterpret_cast<uint2 *>(ws + (size_t)(2 * which + 1) * plane + o) = make_uint2(l01, l23);
    }
    if (bad && flag) *flag = 1;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- pass 2: one CTA = 64 query rows of one (image, head); K/V tiles stream through a double-buffered cp.async pipeline.
// out [B, L, H*64] fp32.
// out [B, L, H*64] fp32 and / or out3 = its fp16 [h*hi_scale | l | h] operand split ([B*L, 3*H*64], see tf32x3.cu) for the
// projection GEMM that follows.
template <bool KX4, bool VX4>
__global__ void __launch_bounds__(128, 2) attention_fp16x3_kernel(const float *__restrict__ qkv, const __half *__restrict__ ws,
                                                                  float *__restrict__ out, __half *__restrict__ out3, int L, int Lp,
                                                                  int H, int B, float q_scale, float hi_scale, int *__restrict__ flag) {
    extern __shared__ __align__(16) __half att_smem[];   // [2 stages][kh, kl, vh, vl][64][72]
    constexpr int kArr = kAttTile * kAttPitch;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttTile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const size_t plane = (size_t)B * H * Lp * kAttD;
    const size_t head = ((size_t)b * H + h) * Lp * kAttD;
    const __half *src[4] = {ws + 2 * plane + head, ws + 3 * plane + head, ws + 4 * plane + head, ws + 5 * plane + head};

    auto load_tile = [&](int stage, int k0) {   // 4 arrays x 64 rows x 128 B = 2048 16-byte pieces, 16 per thread
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            __half *dst = att_smem + (stage * 4 + a) * kArr;
            const __half *s = src[a] + (size_t)k0 * kAttD;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = threadIdx.x + 128 * i;
                const int row = idx >> 3, c8 = (idx & 7) * 8;
                cp_async16(dst + row * kAttPitch + c8, s + row * kAttD + c8);
            }
        }
        cp_async_commit();
    };
    load_tile(0, 0);

    // ---- this warp's 16 query rows as A fragments (hi and lo), pre-multiplied by in_scale * softmax_scale * log2(e)
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const size_t row_stride = (size_t)3 * H * kAttD;
    const float *qb = qkv + (size_t)b * L * row_stride + (size_t)h * kAttD;
    unsigned qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int c = 16 * ks + 2 * tig;
        float2 v00 = make_float2(0.f, 0.f), v10 = v00, v01 = v00, v11 = v00;
        if (r0 < L) {
            v00 = *reinterpret_cast<const float2 *>(qb + (size_t)r0 * row_stride + c);
            v01 = *reinterpret_cast<const float2 *>(qb + (size_t)r0 * row_stride + c + 8);
        }
        if (r1 < L) {
            v10 = *reinterpret_cast<const float2 *>(qb + (size_t)r1 * row_stride + c);
            v11 = *reinterpret_cast<const float2 *>(qb + (size_t)r1 * row_stride + c + 8);
        }
        split2(v00.x * q_scale, v00.y * q_scale, qh[ks][0], ql[ks][0]);
        split2(v10.x * q_scale, v10.y * q_scale, qh[ks][1], ql[ks][1]);
        split2(v01.x * q_scale, v01.y * q_scale, qh[ks][2], ql[ks][2]);
        split2(v11.x * q_scale, v11.y * q_scale, qh[ks][3], ql[ks][3]);
    }
    float o_main[8][4], o_corr[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o_main[nt][i] = o_corr[nt][i] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const int n_tiles = Lp / kAttTile;
    for (int t = 0; t < n_tiles; ++t) {
        const int k0 = t * kAttTile, stage = t & 1;
        if (t + 1 < n_tiles) {
            load_tile(stage ^ 1, k0 + kAttTile);   // the buffer read two iterations ago: its readers passed the barrier below
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const __half *s_kh = att_smem + (stage * 4 + 0) * kArr, *s_kl = att_smem + (stage * 4 + 1) * kArr;
        const __half *s_vh = att_smem + (stage * 4 + 2) * kArr, *s_vl = att_smem + (stage * 4 + 3) * kArr;

        // ---- S = Q K^T for 16 rows x 64 keys: main and correction accumulators; consecutive mma hit different accumulators
        float s_main[8][4], s_corr[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s_main[nt][i] = s_corr[nt][i] = 0.f;
        if (KX4) {
            // B fragments of K^T through ldmatrix.x4 over the row-major [key][dim] tile: matrices (keys 8nt.., dims 16ks..+7),
            // (same keys, dims +8) = b0, b1 of k-step ks, then the same pair for k-step ks+1; lane l supplies row (l & 7) of matrix l >> 3
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {           // k-steps 2kp, 2kp+1
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int off = (8 * nt + (lane & 7)) * kAttPitch + 32 * kp + 8 * (lane >> 3);
                    unsigned h0, h1, h2, h3, l0_, l1_, l2_, l3_;
                    ldmatrix_x4(h0, h1, h2, h3, s_kh + off);
                    ldmatrix_x4(l0_, l1_, l2_, l3_, s_kl + off);
                    mma_16816(s_main[nt], qh[2 * kp], h0, h1);
                    mma_16816(s_corr[nt], ql[2 * kp], h0, h1);
                    mma_16816(s_corr[nt], qh[2 * kp], l0_, l1_);
                    mma_16816(s_main[nt], qh[2 * kp + 1], h2, h3);
                    mma_16816(s_corr[nt], ql[2 * kp + 1], h2, h3);
                    mma_16816(s_corr[nt], qh[2 * kp + 1], l2_, l3_);
                }
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const __half *krow_h = s_kh + (8 * nt + g) * kAttPitch + 2 * tig + 16 * ks;
                    const __half *krow_l = s_kl + (8 * nt + g) * kAttPitch + 2 * tig + 16 * ks;
                    const unsigned bh0 = *reinterpret_cast<const unsigned *>(krow_h), bh1 = *reinterpret_cast<const unsigned *>(krow_h + 8);
                    const unsigned bl0 = *reinterpret_cast<const unsigned *>(krow_l), bl1 = *reinterpret_cast<const unsigned *>(krow_l + 8);
                    mma_16816(s_main[nt], qh[ks], bh0, bh1);
                    mma_16816(s_corr[nt], ql[ks], bh0, bh1);
                    mma_16816(s_corr[nt], qh[ks], bl0, bl1);
                }
            }
        }
        // ---- online softmax in the exp2 domain (fp32)
        float mx0 = -INFINITY, mx1 = -INFINITY;
        const bool tail = k0 + kAttTile > L;   // only the last tile holds keys past L (uniform branch)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s_main[nt][i] = fmaf(s_corr[nt][i], 1.0f / 2048.0f, s_main[nt][i]);
            if (tail) {
                const int key = k0 + 8 * nt + 2 * tig;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (key + (i & 1) >= L) s_main[nt][i] = -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s_main[nt][0], s_main[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s_main[nt][2], s_main[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: every tile holds at least one key < L
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s_main[nt][0] = exp2f(s_main[nt][0] - mn0);
            s_main[nt][1] = exp2f(s_main[nt][1] - mn0);
            s_main[nt][2] = exp2f(s_main[nt][2] - mn1);
            s_main[nt][3] = exp2f(s_main[nt][3] - mn1);
            sum0 += s_main[nt][0] + s_main[nt][1];
            sum1 += s_main[nt][2] + s_main[nt][3];
            o_main[nt][0] *= a0; o_main[nt][1] *= a0; o_main[nt][2] *= a1; o_main[nt][3] *= a1;
            o_corr[nt][0] *= a0; o_corr[nt][1] *= a0; o_corr[nt][2] *= a1; o_corr[nt][3] *= a1;
        }
        l0 = l0 * a0 + sum0;   // per-thread partial sums; reduced over the quad at the end
        l1 = l1 * a1 + sum1;

        // ---- O += P V : P (this thread's C fragments) becomes the A fragments of the next mma, split into hi / lo
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // keys 16 ks .. 16 ks + 15
            unsigned ph[4], pl[4];
            split2(s_main[2 * ks][0], s_main[2 * ks][1], ph[0], pl[0]);
            split2(s_main[2 * ks][2], s_main[2 * ks][3], ph[1], pl[1]);
            split2(s_main[2 * ks + 1][0], s_main[2 * ks + 1][1], ph[2], pl[2]);
            split2(s_main[2 * ks + 1][2], s_main[2 * ks + 1][3], ph[3], pl[3]);
            // B fragments of V[keys 16ks.., dims 8nt..]: ldmatrix.trans over the row-major [key][dim] tile; lanes 0-15 give the rows
            if (VX4) {
                // x4.trans: lanes 0-15 give the 16 key rows at dims 8nt.. (b0, b1 of n-tile nt), lanes 16-31 the same rows at dims
                // 8(nt+1).. (b0, b1 of n-tile nt+1)
                const int voff = (16 * ks + (lane & 15)) * kAttPitch + 8 * (lane >> 4);
#pragma unroll
                for (int nt = 0; nt < 8; nt += 2) {
                    unsigned vh0, vh1, vh2, vh3, vl0, vl1, vl2, vl3;
                    ldmatrix_x4_trans(vh0, vh1, vh2, vh3, s_vh + voff + 8 * nt);
                    ldmatrix_x4_trans(vl0, vl1, vl2, vl3, s_vl + voff + 8 * nt);
                    mma_16816(o_main[nt], ph, vh0, vh1);
                    mma_16816(o_corr[nt], pl, vh0, vh1);
                    mma_16816(o_corr[nt], ph, vl0, vl1);
                    mma_16816(o_main[nt + 1], ph, vh2, vh3);
                    mma_16816(o_corr[nt + 1], pl, vh2, vh3);
                    mma_16816(o_corr[nt + 1], ph, vl2, vl3);
                }
            } else {
                const int vrow = 16 * ks + (lane & 15);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    unsigned vh0, vh1, vl0, vl1;
                    ldmatrix_x2_trans(vh0, vh1, s_vh + vrow * kAttPitch + 8 * nt);
                    ldmatrix_x2_trans(vl0, vl1, s_vl + vrow * kAttPitch + 8 * nt);
                    mma_16816(o_main[nt], ph, vh0, vh1);
                    mma_16816(o_corr[nt], pl, vh0, vh1);
                    mma_16816(o_corr[nt], ph, vl0, vl1);
                }
            }
        }
        __syncthreads();   // everyone is done with this stage before the next iteration's prefetch overwrites the other one's successor
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int Dm = H * kAttD;
    bool bad = false;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = h * kAttD + 8 * nt + 2 * tig;
#pragma unroll
        for (int half_row = 0; half_row < 2; ++half_row) {
            const int r = half_row ? r1 : r0;
            if (r >= L) continue;
            const float inv = half_row ? inv1 : inv0;
            const float x0 = fmaf(o_corr[nt][2 * half_row], 1.0f / 2048.0f, o_main[nt][2 * half_row]) * inv;
            const float x1 = fmaf(o_corr[nt][2 * half_row + 1], 1.0f / 2048.0f, o_main[nt][2 * half_row + 1]) * inv;
            const size_t m = (size_t)b * L + r;
            if (out) *reinterpret_cast<float2 *>(out + m * Dm + c) = make_float2(x0, x1);
            if (out3) {   // [h * hi_scale | l | h] of the attention output, ready for the projection GEMM
                unsigned hh, ll;
                split2(x0, x1, hh, ll);
                __half2 hs = __hmul2(*reinterpret_cast<__half2 *>(&hh), __float2half2_rn(hi_scale));   // a power of two: exact
                bad = bad || !(fabsf(x0 * hi_scale) <= 65504.f && fabsf(x1 * hi_scale) <= 65504.f);
                __half *row = out3 + m * 3 * Dm;
                *reinterpret_cast<unsigned *>(row + c) = *reinterpret_cast<unsigned *>(&hs);
                *reinterpret_cast<unsigned *>(row + Dm + c) = ll;
                *reinterpret_cast<unsigned *>(row + 2 * Dm + c) = hh;
            }
        }
    }
    if (bad && flag) *flag = 1;
}

// attention_tc5.cu: the same contraction on tcgen05.mma / tensor memory (the default main kernel)
int launch_attention_tc5(const float *qkv, const __half *ws, float *out, __half *out3, int L, int Lp, int H, int B, float q_scale,
                         float hi_scale, int *flag, cudaStream_t st);

}  // namespace pnp

using namespace pnp;

extern "C" size_t pnp_attention_fp16x3_workspace_bytes(int B, int L, int H, int D) {
    if (B < 0 || L < 1 || H < 1 || D != kAttD) return 0;
    const size_t Lp = (size_t)ceil_div(L, kAttTile) * kAttTile;
    return 6 * (size_t)B * H * Lp * kAttD * sizeof(__half);   // (planes 2..5 = k_hi, k_lo, v_hi, v_lo are used)
}

extern "C" int pnp_attention_fp16x3(const float *qkv, float in_scale, float softmax_scale, float *out, uint16_t *out3, float out3_hi_scale,
                                    void *workspace, size_t workspace_bytes, int *overflow_flag, int B, int L, int H, int D,
                                    pnp_stream_t stream) {
    if (!qkv || (!out && !out3) || (out3 && (reinterpret_cast<uintptr_t>(out3) & 3)) || !workspace || B < 0 || L < 1 || H < 1 || D != kAttD || B > 65535 || H > 65535 ||
        (reinterpret_cast<uintptr_t>(qkv) & 15) || (out && (reinterpret_cast<uintptr_t>(out) & 7)) || (reinterpret_cast<uintptr_t>(workspace) & 15) ||
        !(in_scale > 0.f))
        return PNP_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < pnp_attention_fp16x3_workspace_bytes(B, L, H, D)) return PNP_ERR_WORKSPACE;
    if (B == 0) return PNP_OK;
    cudaStream_t st = as_stream(stream);
    const int Lp = ceil_div(L, kAttTile) * kAttTile;
    const size_t smem = 2 * 4 * (size_t)kAttTile * kAttPitch * sizeof(__half);   // 73 728 B: two CTAs per SM
    // fragment loads: bit 0 = K through ldmatrix.x4 (else quad LDS.32), bit 1 = V through ldmatrix.x4.trans (else x2.trans).  All four
    // run within 1.5 % of each other (0.353-0.359 ms per call at 35 x 16 x 442, profiles/experiments/r2_attention_variants.py): the
    // kernel is bound by the mma pipe and its latency at 2 CTAs per SM, not by fragment loads.  Default 2 (229 registers).
    static const int variant = getenv("PNP_ATT_VARIANT") ? atoi(getenv("PNP_ATT_VARIANT")) : 2;
    __half *ws = reinterpret_cast<__half *>(workspace);
    const bool timed = prof::on(kAttention, st);
    if (timed) prof::begin(kAttention, st);
    attention_split_kernel<<<dim3(Lp, B), 256, 0, st>>>(qkv, ws, L, Lp, H, B, in_scale, overflow_flag);
    const dim3 grid_main(Lp / kAttTile, H, B);
    const float q_scale = in_scale * softmax_scale * 1.4426950408889634f;
    __half *o3 = reinterpret_cast<__half *>(out3);
    // main kernel: tcgen05.mma with TMEM accumulators (attention_tc5.cu) unless PNP_ATT_TCGEN05=0 asks for the mma.sync one below
    const char *env_tc5 = getenv("PNP_ATT_TCGEN05");   // read per call: tests compare the two kernels in one process
    const int use_tc5 = env_tc5 ? atoi(env_tc5) : 1;
    if (use_tc5) {
        const int rc = launch_attention_tc5(qkv, ws, out, o3, L, Lp, H, B, q_scale, out3_hi_scale, overflow_flag, st);
        if (timed) prof::end(kAttention, st);
        return rc;
    }
#define PNP_ATT(KX, VX)                                                                                                       \
    do {                                                                                                                      \
        cudaError_t e = allow_smem(attention_fp16x3_kernel<KX, VX>, smem); \
        if (e != cudaSuccess) return cuda_err(e);                                                                             \
        attention_fp16x3_kernel<KX, VX><<<grid_main, 128, smem, st>>>(qkv, ws, out, o3, L, Lp, H, B, q_scale, out3_hi_scale, overflow_flag); \
    } while (0)
    switch (variant & 3) {
        case 1: PNP_ATT(true, false); break;
        case 2: PNP_ATT(false, true); break;
        case 3: PNP_ATT(true, true); break;
        default: PNP_ATT(false, false); break;
    }
#undef PNP_ATT
    if (timed) prof::end(kAttention, st);
    return launch_status();
}
