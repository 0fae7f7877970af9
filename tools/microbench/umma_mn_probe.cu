// Developer probe (GPU box): 128 x N x K tcgen05.mma (kind::f16, bf16, f32 accumulate) with BOTH operands MN-major in
// shared memory (no swizzle) — the layout a weight-gradient GEMM wants, because its K dimension is the batch row and
// the row-major [row][feature] operands then go to shared memory with 16-byte stores and no transposition.
// Element (mn, k) of an operand lives at   (mn % 8) * 2 + (mn / 8) * MNSTRIDE + (k % 8) * 16 + (k / 8) * 128   bytes
// (an 8 x 8 core matrix = 8 k-rows of 16 bytes, 8 consecutive mn each), MNSTRIDE = K * 16.
// Pins which descriptor field carries which stride, and the two "major" bits of the instruction descriptor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn_probe umma_mn_probe.cu
//   ./umma_mn_probe N K variant [reps]
//      variant 0: no swizzle, LBO field = k-block stride (128), SBO field = mn-block stride; 1: the two swapped
//      variant 2: 128-byte swizzle — element (mn, k) at (mn % 8) * 2 + (((mn / 8) % 8) ^ (k % 8)) * 16 + (k % 8) * 128 +
//                 (k / 8) * 1024 + (mn / 64) * (K / 8) * 1024, LBO field = 64-column block stride, SBO field = 1024; 3: swapped
//      reps > 1 prints the clocks of one K/16-step MMA chain + commit + wait
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// A: [K][128] row-major bf16 (k = row), B: [K][N] row-major bf16
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                             float* __restrict__ D, int N, int K, int variant, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool sw = variant >= 2;
    const uint32_t mnstride = (uint32_t)K * 16;          // no swizzle: bytes between groups of 8 columns
    const uint32_t blk = (uint32_t)(K / 8) * 1024;       // swizzled: bytes between blocks of 64 columns
    uint8_t* sA = smem;
    uint8_t* sB = smem + (sw ? 2 * blk : 16 * mnstride);
    auto offset = [&](int k, int g) -> uint32_t {
        return sw ? (uint32_t)(g / 8) * blk + (k / 8) * 1024 + (k % 8) * 128 + ((g % 8) ^ (k % 8)) * 16
                  : (uint32_t)g * mnstride + (k / 8) * 128 + (k % 8) * 16;
    };
    for (int u = tid; u < K * 16; u += 128) {          // unit = (row k, group of 8 columns)
        const int k = u / 16, g = u % 16;
        *reinterpret_cast<uint4*>(sA + offset(k, g)) = *reinterpret_cast<const uint4*>(A + (size_t)k * 128 + g * 8);
    }
    for (int u = tid; u < K * (N / 8); u += 128) {
        const int k = u / (N / 8), g = u % (N / 8);
        *reinterpret_cast<uint4*>(sB + offset(k, g)) = *reinterpret_cast<const uint4*>(B + (size_t)k * N + g * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256u));
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
    // D = f32, A = B = bf16, A and B MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    const long long c_start = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t step = sw ? 2048u : 256u;                       // 16 k = two k-blocks
                const uint32_t a = smem_u32(sA) + ks * step, b = smem_u32(sB) + ks * step;
                uint64_t da, db;
                if (variant == 0) { da = make_desc(a, 128, mnstride); db = make_desc(b, 128, mnstride); }
                else if (variant == 1) { da = make_desc(a, mnstride, 128); db = make_desc(b, mnstride, 128); }
                else if (variant == 2) { da = make_desc(a, blk, 1024) | ((uint64_t)2 << 61); db = make_desc(b, blk, 1024) | ((uint64_t)2 << 61); }
                else { da = make_desc(a, 1024, blk) | ((uint64_t)2 << 61); db = make_desc(b, 1024, blk) | ((uint64_t)2 << 61); }
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
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
    if (tid == 0 && reps > 1) printf("  clocks per %d-MMA chain + commit + wait: %.1f\n", K / 16, double(clock64() - c_start) / reps);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u));
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128, K = argc > 2 ? atoi(argv[2]) : 64, variant = argc > 3 ? atoi(argv[3]) : 0;
    const int reps = argc > 4 ? atoi(argv[4]) : 1;
    std::vector<__nv_bfloat16> hA(K * 128), hB(K * N);
    std::vector<float> fA(K * 128), fB(K * N);
    srand(11);
    for (int i = 0; i < K * 128; ++i) { float v = float((rand() % 17) - 8); fA[i] = v; hA[i] = __float2bfloat16(v); }
    for (int i = 0; i < K * N; ++i) { float v = float((rand() % 13) - 6) * 0.5f; fB[i] = v; hB[i] = __float2bfloat16(v); }
    __nv_bfloat16 *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * N * 4));
    const size_t smem = variant >= 2 ? (size_t)(2 + (N + 63) / 64) * (K / 8) * 1024 + 1024 : (size_t)(16 + N / 8) * K * 16 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(dA, dB, dD, N, K, variant, reps);
    CK(cudaGetLastError());
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d variant=%d: kernel failed: %s\n", N, K, variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[k * 128 + m] * fB[k * N + n];
            if (!(fabs(s - hD[m * N + n]) <= 1e-3)) { if (bad < 3) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * N + n], s); ++bad; }
        }
    printf("N=%d K=%d variant=%d: %s (%d / %d mismatches)\n", N, K, variant, bad ? "FAIL" : "OK", bad, 128 * N);
    return bad ? 1 : 0;
}
