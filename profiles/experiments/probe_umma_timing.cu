// probe_umma_timing.cu -- what does one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) cost on a B200, by operand source,
// shared-memory layout and N?  One CTA, one issuing thread, REP dependent MMAs into one accumulator, clock64 around
// "first issue .. commit observed".  Also checks the SWIZZLE_128B descriptors (K-major A and B, MN-major B) against a CPU product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma_timing probe_umma_timing.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int M = 128, K = 64, NMAX = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout 0: no swizzle (core matrices 8 rows x 16 B, LBO 128 along K, SBO 1024 along M/N); layout 2: SWIZZLE_128B (rows of 128 B,
// 16-byte pieces XOR-ed with row & 7, SBO 1024 = 8 rows)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int swz, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(swz ? 2 : 0) << 61);
}

// mode bits: 1 = A from TMEM, 2 = SWIZZLE_128B layouts, 4 = B MN-major
__global__ void __launch_bounds__(128) probe(const __half *__restrict__ A, const __half *__restrict__ B, float *__restrict__ D, int N, int mode,
                                             int rep, long long *__restrict__ cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    const bool a_tmem = mode & 1, swz = mode & 2, b_mn = mode & 4;
    __half *sA = reinterpret_cast<__half *>(smem), *sB = reinterpret_cast<__half *>(smem + 16384);
    for (int i = t; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        const uint32_t off = swz ? r * 128 + (((k / 8) ^ (r & 7)) * 16) + (k % 8) * 2 : (r / 8) * 1024 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
        sA[off / 2] = A[i];
    }
    for (int i = t; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        uint32_t off;
        if (!b_mn) off = swz ? n * 128 + (((k / 8) ^ (n & 7)) * 16) + (k % 8) * 2 : (n / 8) * 1024 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
        else if (!swz) off = (n / 8) * 1024 + (k / 8) * 128 + (k % 8) * 16 + (n % 8) * 2;
        else off = (n / 64) * 8192 + k * 128 + ((((n % 64) / 8) ^ (k & 7)) * 16) + (n % 8) * 2;   // [k][64 n] rows of 128 B, 64-column blocks 8 KB apart
        sB[off / 2] = B[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (a_tmem) {   // A row t as packed fp16 pairs in TMEM columns [256, 288)
        uint32_t r[32];
        for (int c = 0; c < 32; ++c) r[c] = (uint32_t)__half_as_ushort(A[t * K + 2 * c]) | ((uint32_t)__half_as_ushort(A[t * K + 2 * c + 1]) << 16);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
            ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
              "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
              "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
              "r"(r[31]), "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    long long t0 = 0;
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(b_mn ? 1 : 0) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        t0 = clock64();
        for (int it = 0; it < rep; ++it) {
            for (int s = 0; s < K / 16; ++s) {
                // k-step s: no swizzle -> two 128-byte core matrices further; swizzled K-major -> 32 bytes further inside the 128-byte row;
                // swizzled MN-major -> 16 key rows = 2 KB further
                const uint32_t a_addr = smem_u32(sA) + (swz ? 32 * s : 256 * s);
                const uint32_t b_addr = smem_u32(sB) + (b_mn ? (swz ? 2048 * s : 256 * s) : (swz ? 32 * s : 256 * s));
                const uint64_t da = make_desc(a_addr, swz, swz ? 16 : 128, 1024);
                const uint64_t db = make_desc(b_addr, swz, (swz && b_mn) ? 8192 : (swz ? 16 : 128), 1024);
                const uint32_t acc = (it | s) > 0;
                if (a_tmem)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                                 "r"(tmem + 256 + 8 * s), "l"(db), "r"(idesc), "r"(acc));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                                 "l"(da), "l"(db), "r"(idesc), "r"(acc));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        cycles[1] = clock64() - t0;   // issue time alone
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
    if (t == 0) cycles[0] = done ? clock64() - t0 : -1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (done && rep == 1) {
        uint32_t r[32];
        for (int c0 = 0; c0 < N; c0 += 32) {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
                  "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[t * NMAX + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    std::vector<__half> hA(M * K), hB(NMAX * K);
    std::vector<float> fA(M * K), fB(NMAX * K), out(M * NMAX);
    srand(1);
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fA[i] = __half2float(hA[i]); }
    for (int i = 0; i < NMAX * K; ++i) { hB[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float *dD; long long *dC;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, NMAX * K * 2); cudaMalloc(&dD, M * NMAX * 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), NMAX * K * 2, cudaMemcpyHostToDevice);
    const size_t smem = 16384 + 32768;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("mode: +1 A from TMEM, +2 SWIZZLE_128B, +4 B MN-major; one MMA = M128 x N x K16\n");
    for (int mode = 0; mode < 8; ++mode)
        for (int N : {64, 128, 256}) {
            if ((mode & 4) && N == 256) continue;   // MN-major stacked test stops at two 64-column blocks
            // correctness: 1 repetition
            cudaMemset(dD, 0, M * NMAX * 4);
            probe<<<1, 128, smem>>>(dA, dB, dD, N, mode, 1, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dD, M * NMAX * 4, cudaMemcpyDeviceToHost);
            double mx = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double s = 0;
                    for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
                    mx = fmax(mx, fabs(s - out[m * NMAX + n]));
                }
            // timing: 64 repetitions x 4 k-steps = 256 dependent MMAs
            long long cyc[2] = {0, 0};
            for (int r = 0; r < 2; ++r) {
                probe<<<1, 128, smem>>>(dA, dB, dD, N, mode, 64, dC);
                cudaDeviceSynchronize();
                cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
            }
            printf("mode %d (%s A, %s, B %s-major)  N %3d : max|err| %.2e   %6.1f cycles per MMA (issue alone %6.1f)\n", mode, (mode & 1) ? "TMEM" : "smem",
                   (mode & 2) ? "SW128 " : "no swz", (mode & 4) ? "MN" : "K ", N, mx, cyc[0] / 256.0, cyc[1] / 256.0);
        }
    return 0;
}
