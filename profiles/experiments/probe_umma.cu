// probe_umma.cu -- one-CTA tcgen05.mma probe: which shared-memory descriptor conventions does the hardware accept for the
// un-swizzled (INTERLEAVE) canonical layouts?  D[128,64] = A[128,K] * B[64,K]^T, fp16 inputs, fp32 accumulator in TMEM.
// Operand A is K-major; operand B is tested K-major ([n][k], k contiguous) and MN-major ([k][n], n contiguous); each with the
// LBO/SBO fields as documented (variant 0) and swapped (variant 1).  Prints the max error of every combination.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma probe_umma.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#include <vector>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128) probe_kernel(const __half *__restrict__ A, const __half *__restrict__ B, float *__restrict__ D,
                                                    int b_mn_major, int a_var, int b_var, int a_tmem, int *__restrict__ err) {
    __shared__ __align__(128) __half sA[M * K];
    __shared__ __align__(128) __half sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    constexpr uint32_t LBO = 128, SBO = (K / 8) * 128;
    // A: element (r, k) -> (r/8) SBO + (k/8) LBO + (r%8) 16 + (k%8) 2
    for (int i = t; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        const uint32_t off = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2;
        sA[off / 2] = A[i];
    }
    for (int i = t; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        uint32_t off;
        if (!b_mn_major) off = (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2;
        else off = (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2;   // 8 n contiguous, 8 k rows per core matrix
        sB[off / 2] = B[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (a_tmem) {   // A row t as packed fp16 pairs in TMEM columns [64, 96): column c = elements (2c, 2c+1), low half first
        uint32_t r[32];
        for (int c = 0; c < 32; ++c) {
            const unsigned short lo = __half_as_ushort(A[t * K + 2 * c]), hi = __half_as_ushort(A[t * K + 2 * c + 1]);
            r[c] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
            ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
              "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
              "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
              "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (t == 0) {
        // idesc: c_format F32 (1 << 4), a/b F16, a K-major, b major bit 16, N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(b_mn_major ? 1 : 0) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int s = 0; s < K / 16; ++s) {
            const uint32_t a_addr = smem_u32(sA) + s * 2 * LBO, b_addr = smem_u32(sB) + s * 2 * LBO;
            const uint64_t da = a_var ? make_desc(a_addr, SBO, LBO) : make_desc(a_addr, LBO, SBO);
            const uint64_t db = b_var ? make_desc(b_addr, SBO, LBO) : make_desc(b_addr, LBO, SBO);
            const uint32_t acc = s > 0;
            if (a_tmem) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                             "r"(tmem + 64 + 8 * s), "l"(db), "r"(idesc), "r"(acc));
                continue;
            }
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // bounded wait on phase 0
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(&bar)), "r"(0));
    }
    if (!done) { if (t == 0) *err = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (done) {
        uint32_t r[32];
        for (int c0 = 0; c0 < N; c0 += 32) {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[t * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float *dD; int *dErr;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dErr, 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
    for (int a_tm = 0; a_tm < 2; ++a_tm)
    for (int b_mn = 0; b_mn < 2; ++b_mn)
        for (int av = 0; av < 1; ++av)
            for (int bv = 0; bv < 1; ++bv) {
                cudaMemset(dD, 0, M * N * 4); cudaMemset(dErr, 0, 4);
                probe_kernel<<<1, 128>>>(dA, dB, dD, b_mn, av, bv, a_tm, dErr);
                cudaError_t e = cudaDeviceSynchronize();
                int herr = 0;
                if (e == cudaSuccess) {
                    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
                    cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
                }
                double mx = 0;
                for (int i = 0; i < M * N; ++i) mx = fmax(mx, fabs((double)out[i] - ref[i]));
                printf("A in %s  B %s-major  a_var %d  b_var %d : cuda %s  timeout %d  max|err| %.3e\n", a_tm ? "TMEM" : "smem", b_mn ? "MN" : "K", av, bv, cudaGetErrorString(e),
                       herr, mx);
                if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
            }
    return 0;
}
