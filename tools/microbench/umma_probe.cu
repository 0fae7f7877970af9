// Developer probe (GPU box): one 128 x N x K tcgen05.mma (kind::f16, bf16 inputs, f32 accumulate in TMEM) from operands
// laid out by plain st.shared in the no-swizzle K-major canonical layout, checked against a CPU product. Used once to pin
// the shared-memory descriptor fields (LBO / SBO meaning) and the instruction descriptor before the fused deformation
// kernel was written on top of the same primitives.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
//   ./umma_probe N K variant        variant 0: LBO = K-chunk stride, SBO = 8-row-group stride; 1: swapped
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    return d;                        // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

__global__ void __launch_bounds__(128) umma_probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                  float* __restrict__ D, int N, int K, int variant, int reps)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 128 * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5;

    // A: row r = tid, K/8 chunks of 16 B; physical offset = chunk * (128 * 16) + r * 16
    for (int kc = 0; kc < K / 8; ++kc)
        *reinterpret_cast<uint4*>(sA + kc * 128 * 16 + tid * 16) = *reinterpret_cast<const uint4*>(A + (size_t)tid * K + kc * 8);
    for (int i = tid; i < N * (K / 8); i += 128) {
        int n = i % N, kc = i / N;
        *reinterpret_cast<uint4*>(sB + kc * N * 16 + n * 16) = *reinterpret_cast<const uint4*>(B + (size_t)n * K + kc * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // variants 2/3: the A operand lives in TMEM (lane = row, 32-bit column c = K elements 2c | 2c+1), written by tcgen05.st
    const uint32_t tmem_a = tmem + 256;
    if (variant == 2 || variant == 3 || variant == 5) {
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
            for (int i = 0; i < 8; ++i) {
                uint32_t e0 = reinterpret_cast<const uint16_t*>(A)[(size_t)tid * K + 2 * (c0 + i)];
                uint32_t e1 = reinterpret_cast<const uint16_t*>(A)[(size_t)tid * K + 2 * (c0 + i) + 1];
                v[i] = variant != 3 ? (e0 | (e1 << 16)) : (e1 | (e0 << 16));
            }
            uint32_t taddr = tmem_a + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t strideA = 128 * 16, strideB = (uint32_t)N * 16;   // K-chunk stride in bytes
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (variant == 4 || variant == 5) {
            // warp-uniform issue: all 32 lanes of warp 0 run the loop, one elected lane executes the MMA
            if (warp == 0) {
                uint32_t elected;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
                uint64_t da = make_desc(smem_u32(sA), strideA, 128), db = make_desc(smem_u32(sB), strideB, 128);
                const uint64_t ia = (2 * strideA) >> 4, ib = (2 * strideB) >> 4;
                const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
                for (int ks = 0; ks < K / 16; ++ks) {
                    if (elected && variant == 5)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                     :: "r"(tm), "r"(tm + 256 + ks * 8), "l"(db), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
                    else if (elected)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     :: "r"(tm), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
                    da += ia; db += ib;
                }
                if (elected)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
                __syncwarp();
            }
        } else
        if (tid == 0) {
            for (int ks = 0; ks < K / 16; ++ks) {
                uint64_t da = variant == 0 ? make_desc(smem_u32(sA) + ks * 2 * strideA, strideA, 128)
                                           : make_desc(smem_u32(sA) + ks * 2 * strideA, 128, strideA);
                uint64_t db = variant == 0 ? make_desc(smem_u32(sB) + ks * 2 * strideB, strideB, 128)
                                           : make_desc(smem_u32(sB) + ks * 2 * strideB, 128, strideB);
                uint32_t acc = ks > 0;
                if (variant == 2 || variant == 3) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 :: "r"(tmem), "r"(tmem_a + ks * 8), "l"(make_desc(smem_u32(sB) + ks * 2 * strideB, strideB, 128)),
                                    "r"(idesc), "r"(acc) : "memory");
                    continue;
                }
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        }
        uint32_t done = 0;
        for (long long spin = 0; !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
            if (spin > (1ll << 22)) { if (tid == 0) printf("mbarrier wait timed out\n"); __trap(); }
        }
        phase ^= 1;
    }
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    if (tid == 0 && reps > 1) printf("clocks per %d-step MMA chain + commit + wait: %.1f\n", K / 16, double(t1 - t0) / reps);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}

int main(int argc, char** argv)
{
    int N = argc > 1 ? atoi(argv[1]) : 128, K = argc > 2 ? atoi(argv[2]) : 128, variant = argc > 3 ? atoi(argv[3]) : 0;
    int reps = argc > 4 ? atoi(argv[4]) : 1;
    std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K);
    srand(7);
    for (int i = 0; i < 128 * K; ++i) { float v = float((rand() % 17) - 8); fA[i] = v; hA[i] = __float2bfloat16(v); }
    for (int i = 0; i < N * K; ++i) { float v = float((rand() % 13) - 6) * 0.5f; fB[i] = v; hB[i] = __float2bfloat16(v); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    size_t smem = (128 + N) * K * 2 + 1024;
    CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, variant, reps);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[r * K + k] * fB[n * K + k];
            double e = fabs(s - hD[r * N + n]);
            if (!(e <= 1e-3)) { if (bad < 4) printf("  mismatch r=%d n=%d got %g want %g\n", r, n, hD[r * N + n], s); ++bad; }
            if (e > maxerr) maxerr = e;
        }
    printf("N=%d K=%d variant=%d: %s (%d / %d mismatches, max err %g)\n", N, K, variant, bad ? "FAIL" : "OK", bad, 128 * N, maxerr);
    return bad ? 1 : 0;
}
