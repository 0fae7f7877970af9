// Per-frame deformation -> rasterizer hand-off: everything the reference's
// GaussianModel.get_deformation_eval does between the cached plane features and the rasterizer's inputs, as one
// selection pass plus ONE fused tcgen05 kernel.
//
// Reference (scene/saro_gaussian.py):
//   :871-881  distance = t - temporal_pos; state = exp(-4 (distance / lifespan)^2); select state > 0.001
//   :875-876  feature = cat(hexplane_feature, [d, sin(d), cos(d), sin(2d), cos(2d), sin(4d), cos(4d), sin(8d), cos(8d)])
//             (get_embedder(4), :922-969)
//   :104,108,110  motion_mlp / rot_mlp / shs_mlp = Linear(in,128) ReLU Linear(128,128) ReLU Linear(128, 3 | 7 | 48)
//   :883-885  means3D  = xyz[sel] + motion_mlp(feature)
//   :889-897  rotation = normalize(rotation[sel] + rot_mlp(feature)[:, :4]);  scale = exp(scaling[sel] + rot_mlp(feature)[:, 4:])
//   :903-905  opacity  = sigmoid(opacity[sel]) * state
//   :911-915  shs      = cat(features_dc, features_rest)[sel] + shs_mlp(feature).reshape(-1, 16, 3)
// In the reference this is ~40 PyTorch kernels (boolean-mask gathers, 9 fp32 GEMMs with the hidden activations written
// to and re-read from HBM, elementwise epilogues, cat) inside the timed region of the test-time render loop
// (renderer/__init__.py:188-203).
//
// B200 design.  The three MLPs are GEMM-shaped (145 kFLOP per Gaussian), so they run on the 5th-generation tensor
// cores.  Persistent kernel, one 512-thread CTA per SM:
//   * a CTA owns ONE of the three MLPs for its whole life and keeps that MLP's weights resident in shared memory
//     (bf16 hi + lo planes in the UMMA no-swizzle K-major canonical layout, 116 KB); CTAs are split between the MLPs
//     in proportion to their measured cost per tile;
//   * two independent row groups of 128 Gaussians (UMMA M = 128 = TMEM lanes), two threads per row.  TMEM is fully
//     allocated: per group 128 accumulator columns + 64 + 64 columns that hold the A operand (hi / lo bf16 planes,
//     written with tcgen05.st).  tcgen05.mma takes A from TMEM and B from shared memory; hidden activations never touch
//     shared memory or HBM.  While one group's layer is in the tensor core the other group runs its epilogue;
//   * per tile: each thread builds its share of the layer-1 operand row (plane features + time embedding), the group's
//     first warp issues the layer (all lanes run the unrolled sequence, one elected lane per tcgen05.mma), and after the
//     tcgen05.commit -> mbarrier hand-shake every thread reads its half of its accumulator row with tcgen05.ld,
//     applies bias + ReLU in packed f32x2 and writes the next layer's operand;
//   * fp32 fidelity on bf16 tensor cores: every operand x is split as x = hi + lo (two bf16 values, 16 significand
//     bits) and each product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation — measured 7e-6 of the output
//     range against float64, 13x inside the 1e-4 parity bar, at 3 MMAs per product instead of the 8x slower fp32 SIMT
//     path (plain bf16 gives 4e-3);
//   * the last layer's epilogue applies the residual add and activations and writes the rasterizer's inputs directly;
//     the 48-float SH rows are gathered and written back coalesced through a shared staging buffer.
// Descriptor encodings and the TMEM operand packing were pinned on hardware with tools/microbench/umma_probe.cu.
#include "../../include/saro_gs_b200.h"
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace sgs_deform {

constexpr int ROWS = 128;            // Gaussians per tile = UMMA M = TMEM lanes
constexpr int GROUP_THREADS = 256;   // two threads per row
constexpr int HID = 128;             // hidden width (args.deform_hidden_dim)
constexpr int K1 = 48;               // padded input width (feature dim + 9 <= 48)
constexpr int TIME_DIMS = 9;         // get_embedder(4): x + 4 x (sin, cos)
constexpr int N3_MAX = 48;
// packed weight image of one MLP (bytes); every matrix in canonical layout: elem (n, k) at (k/8)*(N*16) + n*16 + (k%8)*2
constexpr int OFF_W1HI = 0;
constexpr int OFF_W1LO = OFF_W1HI + K1 * HID * 2;
constexpr int OFF_W2HI = OFF_W1LO + K1 * HID * 2;
constexpr int OFF_W2LO = OFF_W2HI + HID * HID * 2;
constexpr int OFF_W3 = OFF_W2LO + HID * HID * 2;               // [hi rows ; lo rows] stacked along N
constexpr int OFF_B1 = OFF_W3 + 2 * HID * N3_MAX * 2;
constexpr int OFF_B2 = OFF_B1 + HID * 4;
constexpr int OFF_B3 = OFF_B2 + HID * 4;
constexpr int IMG_BYTES = OFF_B3 + N3_MAX * 4;                 // 115 904
constexpr int IMG_PAD = (IMG_BYTES + 1023) / 1024 * 1024;      // 116 736
constexpr int STAGE_STRIDE = 49;                               // floats per staged SH row (48 + 1: conflict-free by row and by column)
constexpr int STAGE_BYTES = ROWS * STAGE_STRIDE * 4;           // 25 088 per row group
constexpr int SMEM_BYTES = IMG_PAD + 2 * STAGE_BYTES + 64;     // weights + SH staging + two mbarriers + the TMEM base slot

// measured k-clocks per tile of each MLP class (SGS_DEFORM_PROFILE=1); drives the CTA split
constexpr float COST_MOTION = 10.7f, COST_ROT = 11.9f, COST_SHS = 15.4f;

__host__ __device__ constexpr int n3_real(int mlp) { return mlp == 0 ? 3 : mlp == 1 ? 7 : 48; }
__host__ __device__ constexpr int n3_pad(int mlp) { return mlp == 2 ? 48 : 16; }

struct Alive {   // saro_gaussian.py:872-873,878 and :757-759
    const float* tpos; const float* life; float t;
    __device__ __forceinline__ float state(int i) const {
        const float d = t - tpos[i];
        const float q = d / life[i];
        return expf(-4.f * (q * q));
    }
    __device__ __forceinline__ bool operator()(int i) const { return state(i) > 0.001f; }
};

// ------------------------------------------------------------------------------------------------ weight packing
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__global__ void pack_mlp_kernel(int mlp, int in_dim, const float* __restrict__ W1, const float* __restrict__ b1,
                                const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                                const float* __restrict__ b3, uint8_t* __restrict__ img) {
    const int n3r = n3_real(mlp), n3p = n3_pad(mlp);
    const int total = HID * K1 + HID * HID + 2 * n3p * HID + 2 * HID + N3_MAX;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int i = e;
        if (i < HID * K1) {                         // W1 [128][in_dim] -> [128][K1]
            const int n = i / K1, k = i % K1;
            __nv_bfloat16 hi, lo;
            split_bf16(k < in_dim ? W1[n * in_dim + k] : 0.f, hi, lo);
            const int off = (k / 8) * (HID * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W1HI + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W1LO + off) = lo;
            continue;
        }
        i -= HID * K1;
        if (i < HID * HID) {
            const int n = i / HID, k = i % HID;
            __nv_bfloat16 hi, lo;
            split_bf16(W2[n * HID + k], hi, lo);
            const int off = (k / 8) * (HID * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W2HI + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W2LO + off) = lo;
            continue;
        }
        i -= HID * HID;
        if (i < 2 * n3p * HID) {                    // W3: hi and lo planes stacked along N -> one [2 n3][128] operand
            const int ns = i / HID, k = i % HID, n = ns % n3p;
            __nv_bfloat16 hi, lo;
            split_bf16(n < n3r ? W3[n * HID + k] : 0.f, hi, lo);
            const int off = (k / 8) * (2 * n3p * 16) + ns * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W3 + off) = ns < n3p ? hi : lo;
            continue;
        }
        i -= 2 * n3p * HID;
        if (i < HID) { reinterpret_cast<float*>(img + OFF_B1)[i] = b1[i]; continue; }
        i -= HID;
        if (i < HID) { reinterpret_cast<float*>(img + OFF_B2)[i] = b2[i]; continue; }
        i -= HID;
        reinterpret_cast<float*>(img + OFF_B3)[i] = i < n3r ? b3[i] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// B operand: shared-memory matrix descriptor, no swizzle, K-major: LBO = byte stride between the two 8-wide K chunks
// of one instruction, SBO = byte stride between 8-row groups (128: rows are packed 16 B apart).
// A operand: TMEM, lane = row, 32-bit column c = K elements (2c | 2c+1 << 16).
// Both pinned on hardware by tools/microbench/umma_probe.cu (variants 0 and 2).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// One layer: D[128 x N] = A[128 x 16*KSTEPS] * B[N x 16*KSTEPS]^T as hi*hi + hi*lo + lo*hi.  Called by ALL lanes of
// the issuing warp with warp-uniform arguments (descriptor arithmetic stays off the critical path; the probe measured
// 64 clocks per N = 128 instruction this way against 160 when a single divergent thread builds the operands);
// one lane is elected per instruction.
template <int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t acc_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int N, uint32_t mbar) {
    const uint32_t idesc = umma_idesc(N);
    const uint32_t b_chunk = (uint32_t)N * 16;
    const uint64_t b_step = (uint64_t)((2 * b_chunk) >> 4);
#pragma unroll
    for (int prod = 0; prod < 3; ++prod) {
        const uint32_t a = prod == 2 ? a_lo : a_hi;
        uint64_t db = umma_desc(prod == 1 ? b_lo : b_hi, b_chunk);
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            // elect.sync and the predicated MMA in one block: ptxas then emits a single predicated UTCHMMA instead of a
            // per-active-thread serialisation loop around it
            if ((prod | ks) != 0)
                asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t}"
                             :: "r"(acc_tmem), "r"(a + ks * 8), "l"(db), "r"(idesc) : "memory");
            else
                asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 0;\n\t}"
                             :: "r"(acc_tmem), "r"(a + ks * 8), "l"(db), "r"(idesc) : "memory");
            db += b_step;
        }
    }
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(mbar) : "memory");
    __syncwarp();
}

// Last layer (N3 = 16 or 48 outputs: far below the tensor pipe's appetite, so the instruction count is what costs).
// W3's hi and lo planes are stacked along N: A_hi x [W_hi ; W_lo] leaves hi*hi in accumulator columns [0, N3) and
// hi*lo in [N3, 2 N3) with ONE instruction per K step, A_lo x W_hi accumulates into [0, N3): 16 instructions
// instead of 24; the output epilogue adds the two column blocks.
__device__ __forceinline__ void issue_last_layer(uint32_t acc_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b, int N3,
                                                 uint32_t mbar) {
    const uint32_t chunk = (uint32_t)(2 * N3) * 16;
    const uint64_t b_step = (uint64_t)((2 * chunk) >> 4);
#pragma unroll
    for (int prod = 0; prod < 2; ++prod) {
        const uint32_t a = prod ? a_lo : a_hi;
        const uint32_t idesc = umma_idesc(prod ? N3 : 2 * N3);
        uint64_t db = umma_desc(b, chunk);
#pragma unroll
        for (int ks = 0; ks < HID / 16; ++ks) {
            if ((prod | ks) != 0)
                asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t}"
                             :: "r"(acc_tmem), "r"(a + ks * 8), "l"(db), "r"(idesc) : "memory");
            else
                asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                             "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 0;\n\t}"
                             :: "r"(acc_tmem), "r"(a + ks * 8), "l"(db), "r"(idesc) : "memory");
            db += b_step;
        }
    }
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                 "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(mbar) : "memory");
    __syncwarp();
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
        if (spin > (1u << 24)) __trap();   // a lost commit must fail loudly, never hang the device
    }
}

__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {    // caller issues tcgen05.wait::ld
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&w)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

// 8 consecutive K values of this thread's row -> 4 + 4 packed words of the hi / lo bf16 planes of the TMEM A operand
__device__ __forceinline__ void store_chunk(uint32_t t_hi, uint32_t t_lo, int kc, const float2 (&v)[4]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __float22bfloat162_rn(v[i]);
        const float2 lo = __ffma2_rn(__bfloat1622float2(hh), make_float2(-1.f, -1.f), v[i]);     // v - hi, exact
        const __nv_bfloat162 ll = __float22bfloat162_rn(lo);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    tmem_st4(t_hi + kc * 4, h);
    tmem_st4(t_lo + kc * 4, l);
}

// order this thread's TMEM stores before the MMA another thread is about to issue, then meet the group
__device__ __forceinline__ void publish_operand(int group) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" :: "r"(1 + group), "r"(GROUP_THREADS) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

struct EvalParams {
    int feat_dim;
    float timestamp;
    const float *xyz, *rotation, *scaling, *opacity, *features_dc, *features_rest, *tpos, *life, *feat;
    const uint8_t* packed;       // 3 images of IMG_BYTES
    const int* index;            // [count] source rows, ascending
    const int* count;
    float *o_means3D, *o_rot, *o_scale, *o_opacity, *o_shs;
    int ctas_shs, ctas_motion;   // CTA split between the MLPs (the rest work on rot)
    long long* phase_clocks;     // developer profiling (SGS_DEFORM_PROFILE=1): [3 MLPs][16 phases] clocks of one thread, or NULL
};

// hidden layer epilogue for this thread's half of the columns: accumulator row -> + bias, ReLU -> next layer's A
// operand (TMEM, bf16 hi / lo planes).  The next 16 columns are in flight while the current 16 are converted.
__device__ __forceinline__ void hidden_epilogue(uint32_t t_acc, uint32_t t_hi, uint32_t t_lo, const float* __restrict__ bias,
                                                int half) {
    const int cbase = half * (HID / 2);
    uint32_t r[2][16];
    tmem_ld16_async(t_acc + cbase, r[0]);
#pragma unroll
    for (int it = 0; it < HID / 32; ++it) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (it + 1 < HID / 32) tmem_ld16_async(t_acc + cbase + (it + 1) * 16, r[(it + 1) & 1]);
        const int c0 = cbase + it * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 b = *reinterpret_cast<const float2*>(bias + c0 + h * 8 + 2 * i);
                const float2 y = __fadd2_rn(make_float2(__uint_as_float(r[it & 1][h * 8 + 2 * i]),
                                                        __uint_as_float(r[it & 1][h * 8 + 2 * i + 1])), b);
                x[i] = make_float2(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f));
            }
            store_chunk(t_hi, t_lo, c0 / 8 + h, x);
        }
    }
}

// Two threads share a row: `half` 0 owns K chunks 0, 2, 4 of the layer-1 operand and the low half of every column
// range, `half` 1 the others.
struct RowInputs {
    float tpos;           // raw load; timestamp - tpos is formed at the point of use so the load never stalls the prefetch
    float4 f[4];          // this thread's (up to two) 8-wide plane-feature chunks
};

__device__ __forceinline__ void load_row_inputs(const EvalParams& p, int src, int nf, int half, RowInputs& r) {
    r.tpos = __ldg(p.tpos + src);
    const float4* frow = reinterpret_cast<const float4*>(p.feat + (size_t)src * p.feat_dim);
#pragma unroll
    for (int c = 0; c < 2; ++c)
        if (2 * c + half < nf) {
            r.f[2 * c] = __ldg(frow + 2 * (2 * c + half));
            r.f[2 * c + 1] = __ldg(frow + 2 * (2 * c + half) + 1);
        }
}

__device__ __forceinline__ int load_src(const EvalParams& p, int tile, int tiles, int row, int count) {
    if (tile >= tiles) return 0;
    const int j = tile * ROWS + row;
    return __ldg(p.index + (j < count ? j : count - 1));
}

#define DF_TICK(i)                                                              \
    if (prof) {                                                                 \
        const long long t_ = clock64();                                         \
        prof[i] += t_ - t_prev;                                                 \
        t_prev = t_;                                                            \
    }

template <int MLP>
__device__ __forceinline__ void run_mlp(const EvalParams& p, const uint8_t* img, float* stage, int group, int worker, int workers,
                                        uint32_t tmem_group, uint32_t mbar) {
    const int gt = threadIdx.x & (GROUP_THREADS - 1);
    const int row = gt & (ROWS - 1), half = gt >> 7, warp_in_group = gt >> 5;
    const bool issuer = warp_in_group == 0;
    const uint32_t lane_base = (uint32_t)((warp_in_group & 3) * 32) << 16;
    const uint32_t acc_u = tmem_group, hi_u = tmem_group + 128, lo_u = tmem_group + 192;     // lane-0 addresses (MMA)
    const uint32_t t_acc = acc_u + lane_base, t_hi = hi_u + lane_base, t_lo = lo_u + lane_base;
    const uint32_t simg = smem_u32(img);
    const float* b1 = reinterpret_cast<const float*>(img + OFF_B1);
    const float* b2 = reinterpret_cast<const float*>(img + OFF_B2);
    const float* b3 = reinterpret_cast<const float*>(img + OFF_B3);
    const int count = *p.count;
    const int tiles = (count + ROWS - 1) / ROWS;
    const int nf = p.feat_dim >> 3;                    // feature chunks of 8 (feat_dim is a multiple of 8, <= 32)
    constexpr int N3 = n3_pad(MLP);
    uint32_t phase = 0;
    long long* prof = (p.phase_clocks && worker == 0 && threadIdx.x == 0) ? p.phase_clocks + MLP * 16 : nullptr;
    long long t_prev = clock64();

    // software pipeline over tiles: source index two tiles ahead, row inputs one tile ahead
    RowInputs cur;
    int src = load_src(p, worker, tiles, row, count);
    int src_next = load_src(p, worker + workers, tiles, row, count);
    if (worker < tiles) load_row_inputs(p, src, nf, half, cur);

#pragma unroll 1
    for (int tile = worker; tile < tiles; tile += workers) {
        const int j = tile * ROWS + row;
        const bool valid = j < count;
        const float d = p.timestamp - cur.tpos;

        // ---- layer-1 operand: [plane feature | time embedding | 0 padding]  (saro_gaussian.py:875-876, :939-969)
        {
            // [d, sin d, cos d, sin 2d, cos 2d, sin 4d, cos 4d, sin 8d, cos 8d]: one accurate sincos, then three
            // double-angle steps (error <= 1e-6, far below the bf16 hi/lo split of the operand)
            float emb[TIME_DIMS];
            emb[0] = d;
            sincosf(d, &emb[1], &emb[2]);
#pragma unroll
            for (int f = 1; f < 4; ++f) {
                emb[1 + 2 * f] = 2.f * emb[2 * f - 1] * emb[2 * f];
                emb[2 + 2 * f] = 1.f - 2.f * emb[2 * f - 1] * emb[2 * f - 1];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int kc = 2 * c + half;
                float2 x[4];
                if (c < 2 && kc < nf) {
                    const float4 u = cur.f[2 * (c < 2 ? c : 0)], w = cur.f[2 * (c < 2 ? c : 0) + 1];
                    x[0] = make_float2(u.x, u.y); x[1] = make_float2(u.z, u.w);
                    x[2] = make_float2(w.x, w.y); x[3] = make_float2(w.z, w.w);
                } else if (kc == nf) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(emb[2 * i], emb[2 * i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(0.f, 0.f);
                    if (kc == nf + 1) x[0].x = emb[8];
                }
                store_chunk(t_hi, t_lo, kc, x);
            }
        }
        DF_TICK(0)
        publish_operand(group);
        DF_TICK(1)
        if (issuer) issue_layer<K1 / 16>(acc_u, hi_u, lo_u, simg + OFF_W1HI, simg + OFF_W1LO, HID, mbar);
        DF_TICK(2)

        // ---- while the tensor core works: next tile's inputs and this tile's residual bases (consumed much later)
        float sh0[16], sh1[16];                           // shs class: this warp's 16 rows of the SH block, in flight
        if (MLP == 2) {
            // the [16][3] SH block of 16 rows per warp, one row per pass: lanes walk the row's 48 contiguous floats
            // (3 dc + 45 rest), so the gather is coalesced; all 32 loads are issued before anything waits on them
            const int sub = warp_in_group >> 2, lane = gt & 31;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int r_src = __shfl_sync(0xffffffffu, src, sub * 16 + i);
                const float* rest = p.features_rest + (size_t)r_src * 45;
                sh0[i] = __ldg(lane < 3 ? p.features_dc + (size_t)r_src * 3 + lane : rest + (lane - 3));
                sh1[i] = lane < 16 ? __ldg(rest + 29 + lane) : 0.f;
            }
        }
        RowInputs nxt = cur;
        if (tile + workers < tiles) load_row_inputs(p, src_next, nf, half, nxt);
        const int src_next2 = load_src(p, tile + 2 * workers, tiles, row, count);
        float base[4];
        float life = 1.f;
        if (MLP == 0) {
            if (half == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) base[c] = __ldg(p.xyz + (size_t)src * 3 + c);
            }
        } else if (MLP == 1) {
            if (half == 0) {
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.rotation) + src);
                base[0] = q0.x; base[1] = q0.y; base[2] = q0.z; base[3] = q0.w;
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) base[c] = __ldg(p.scaling + (size_t)src * 3 + c);
                base[3] = __ldg(p.opacity + src);
                life = __ldg(p.life + src);
            }
        } else {
            // staged in shared memory (49-float row stride) until the output epilogue
            const int sub = warp_in_group >> 2, lane = gt & 31;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float* srow = stage + ((warp_in_group & 3) * 32 + sub * 16 + i) * STAGE_STRIDE;
                srow[lane] = sh0[i];
                if (lane < 16) srow[32 + lane] = sh1[i];
            }
        }

        DF_TICK(3)
        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        DF_TICK(4)
        hidden_epilogue(t_acc, t_hi, t_lo, b1, half);
        DF_TICK(5)
        publish_operand(group);
        DF_TICK(6)
        if (issuer) issue_layer<HID / 16>(acc_u, hi_u, lo_u, simg + OFF_W2HI, simg + OFF_W2LO, HID, mbar);
        DF_TICK(7)
        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        DF_TICK(8)
        hidden_epilogue(t_acc, t_hi, t_lo, b2, half);
        DF_TICK(9)
        publish_operand(group);
        DF_TICK(10)
        if (issuer) issue_last_layer(acc_u, hi_u, lo_u, simg + OFF_W3, N3, mbar);
        DF_TICK(11)
        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        DF_TICK(12)

        // ---- output epilogue: residual + activation, written in the rasterizer's input layout
        if (MLP == 0) {                                  // means3D = xyz + motion            (:883-885)
            float r[8], r2[8];
            tmem_ld8(t_acc, r);
            tmem_ld8(t_acc + N3, r2);
            if (valid && half == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) p.o_means3D[(size_t)j * 3 + c] = base[c] + ((r[c] + r2[c]) + b3[c]);
            }
        } else if (MLP == 1) {                           // rotation, scale, opacity         (:889-897, :903-905)
            float r[8], r2[8];
            tmem_ld8(t_acc, r);
            tmem_ld8(t_acc + N3, r2);
#pragma unroll
            for (int c = 0; c < 8; ++c) r[c] += r2[c];
            if (valid && half == 0) {
                const float qx = base[0] + (r[0] + b3[0]), qy = base[1] + (r[1] + b3[1]);
                const float qz = base[2] + (r[2] + b3[2]), qw = base[3] + (r[3] + b3[3]);
                const float nrm = fmaxf(sqrtf(qx * qx + qy * qy + qz * qz + qw * qw), 1e-12f);   // F.normalize eps
                reinterpret_cast<float4*>(p.o_rot)[j] = make_float4(qx / nrm, qy / nrm, qz / nrm, qw / nrm);
            } else if (valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c) p.o_scale[(size_t)j * 3 + c] = expf(base[c] + (r[4 + c] + b3[4 + c]));
                const float q = d / life;                                    // :872-873, same arithmetic as the selection
                p.o_opacity[j] = (1.f / (1.f + expf(-base[3]))) * expf(-4.f * (q * q));
            }
        } else {                                         // shs = cat(dc, rest) + residual   (:911-915)
            float* mine = stage + row * STAGE_STRIDE + half * 24;
#pragma unroll
            for (int c0 = 0; c0 < 24; c0 += 8) {
                float r[8], r2[8];
                tmem_ld8(t_acc + half * 24 + c0, r);
                tmem_ld8(t_acc + N3 + half * 24 + c0, r2);
                const float* bb = b3 + half * 24 + c0;
#pragma unroll
                for (int i = 0; i < 8; ++i) mine[c0 + i] += (r[i] + r2[i]) + bb[i];
            }
            asm volatile("bar.sync %0, %1;" :: "r"(1 + group), "r"(GROUP_THREADS) : "memory");
            // the tile's 128 x 48 outputs are one contiguous block: coalesced float4 stores
            float4* out = reinterpret_cast<float4*>(p.o_shs + (size_t)tile * ROWS * 48);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int e4 = gt + k * GROUP_THREADS;
                const int o_row = (e4 * 4) / 48, c = (e4 * 4) % 48;
                const float* sp = stage + o_row * STAGE_STRIDE + c;
                if (tile * ROWS + o_row < count) out[e4] = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        }
        cur = nxt;
        src = src_next;
        src_next = src_next2;
        DF_TICK(13)
        if (prof) prof[15] += 1;
    }
}

// 512 threads = two independent row groups (128 rows each, two threads per row) that share one MLP's resident
// weights: while one group's layer is in the tensor core the other group runs its epilogue, so the chain latency of
// one tile hides behind the other's.
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) deform_mlp_kernel(const EvalParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* img = smem;
    float* stage_all = reinterpret_cast<float*>(smem + IMG_PAD);
    uint64_t* mbar_p = reinterpret_cast<uint64_t*>(smem + IMG_PAD + 2 * STAGE_BYTES);          // one per group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_p + 2);

    const int tid = threadIdx.x, warp = tid >> 5, group = tid / GROUP_THREADS;
    // CTAs are split between the three MLPs in proportion to their cost per tile (shs first: it is the longest)
    int mlp = 2, first = 0, ctas = p.ctas_shs;
    if ((int)blockIdx.x >= p.ctas_shs + p.ctas_motion) { mlp = 1; first = p.ctas_shs + p.ctas_motion; ctas = (int)gridDim.x - first; }
    else if ((int)blockIdx.x >= p.ctas_shs) { mlp = 0; first = p.ctas_shs; ctas = p.ctas_motion; }
    const int worker = ((int)blockIdx.x - first) * 2 + group;
    const int workers = ctas * 2;
    float* stage = stage_all + group * (STAGE_BYTES / 4);

    {   // resident weights: linear copy of this MLP's packed image
        const uint4* src = reinterpret_cast<const uint4*>(p.packed + (size_t)mlp * IMG_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(img);
        for (int i = tid; i < IMG_BYTES / 16; i += 2 * GROUP_THREADS) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar_p)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar_p + 1)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_group = tmem + (uint32_t)group * 256;    // acc [0,128) | operand hi [128,192) | operand lo [192,256)
    const uint32_t mbar = smem_u32(mbar_p + group);

    if (mlp == 0) run_mlp<0>(p, img, stage, group, worker, workers, tmem_group, mbar);
    else if (mlp == 1) run_mlp<1>(p, img, stage, group, worker, workers, tmem_group, mbar);
    else run_mlp<2>(p, img, stage, group, worker, workers, tmem_group, mbar);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}


// =====================================================================================================================
// Training path (training half): GaussianModel.get_deformation (scene/saro_gaussian.py:
// 779-847) evaluates, per view and over ALL Gaussians, opacity_mlp on the plane feature (lifespan, :782), the three
// deformation MLPs on [feature | time embedding of t - temporal_pos] (:812,:819,:845) and again on the base feature
// [feature | embedding of 0] (:796-803: regularisation residuals, real_xyz) — up to seven MLP evaluations — and
// autograd runs their backward.  Here every evaluation is a "job" of ONE persistent tcgen05 launch (CTAs split between
// the jobs by cost); forward jobs write raw outputs, the ReLU sign bits (16 bytes per row and layer) and, for the
// weight-gradient kernel, the hidden activations.  The data-gradient chain  dy -> (.W3) o mask2 -> (.W2) o mask1 ->
// (.W1[:, :F]) -> d feature  is the SAME three-layer pipeline with an image packed from the transposed weights and a
// mask multiply in place of bias + ReLU, so it runs through the same code (BWD = true).
constexpr int MAX_JOBS = 8;
constexpr int TRAIN_SMEM_BYTES = IMG_PAD + 64;

struct TrainJob {
    const uint8_t* img;            // packed image (forward: from W; backward: from the transposed W)
    const float* in;               // backward: dL/d out [N][n_io]
    float* out;                    // forward: raw outputs [N][n_io]; backward: dL/d feature [N][feat_dim]
    uint8_t* save_a; uint8_t* save_b;   // operand planes (16 column groups) or NULL: forward h1, h2; backward d h2, d h1 (after the mask)
    uint8_t* save_in;              // operand planes of the layer-1 operand: forward [feature | embedding | 0] (6 groups),
                                   // backward dL/d out (6 groups for 48 outputs, 2 for up to 8), or NULL
    uint2* mask_a; uint2* mask_b;  // [N][2]: sign bits of the two hidden layers, 64 columns per entry (forward writes:
                                   // a = layer 1, b = layer 2; backward reads: a = layer 2, b = layer 1)
    int n_io;
    int n3p;                       // forward 16 | 48, backward 32
    int zero_time;                 // forward: the base feature (time embedding of 0)
    int cta_first, cta_count;
};
struct TrainParams {
    int N, feat_dim, n_jobs;
    float timestamp;
    const float* tpos; const float* feat;
    TrainJob jobs[MAX_JOBS];
};

// general packer: forward image of an MLP in_w -> 128 -> hid2 -> n_out (zero padded to 48 / 128 / n3p), or the image
// of its data-gradient chain n_out -> hid2 -> 128 -> feat_dim (transposed weights, no biases, n3p = 32)
__global__ void pack_general_kernel(int backward, int in_w, int hid2, int n_out, int feat_dim, int n3p,
                                    const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                                    const float* __restrict__ b2, const float* __restrict__ W3, const float* __restrict__ b3,
                                    uint8_t* __restrict__ img) {
    const int total = HID * K1 + HID * HID + 2 * n3p * HID + 2 * HID + N3_MAX;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        int i = e;
        if (i < HID * K1) {
            const int n = i / K1, k = i % K1;
            float v;
            if (!backward) v = k < in_w ? W1[n * in_w + k] : 0.f;
            else v = (k < n_out && n < hid2) ? W3[k * hid2 + n] : 0.f;
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            const int off = (k / 8) * (HID * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W1HI + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W1LO + off) = lo;
            continue;
        }
        i -= HID * K1;
        if (i < HID * HID) {
            const int n = i / HID, k = i % HID;
            float v;
            if (!backward) v = n < hid2 ? W2[n * HID + k] : 0.f;
            else v = k < hid2 ? W2[k * HID + n] : 0.f;
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            const int off = (k / 8) * (HID * 16) + n * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W2HI + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W2LO + off) = lo;
            continue;
        }
        i -= HID * HID;
        if (i < 2 * n3p * HID) {
            const int ns = i / HID, k = i % HID, n = ns % n3p;
            float v;
            if (!backward) v = (n < n_out && k < hid2) ? W3[n * hid2 + k] : 0.f;
            else v = n < feat_dim ? W1[k * in_w + n] : 0.f;
            __nv_bfloat16 hi, lo;
            split_bf16(v, hi, lo);
            const int off = (k / 8) * (2 * n3p * 16) + ns * 16 + (k % 8) * 2;
            *reinterpret_cast<__nv_bfloat16*>(img + OFF_W3 + off) = ns < n3p ? hi : lo;
            continue;
        }
        i -= 2 * n3p * HID;
        if (i < HID) { reinterpret_cast<float*>(img + OFF_B1)[i] = backward ? 0.f : b1[i]; continue; }
        i -= HID;
        if (i < HID) { reinterpret_cast<float*>(img + OFF_B2)[i] = (backward || i >= hid2) ? 0.f : b2[i]; continue; }
        i -= HID;
        reinterpret_cast<float*>(img + OFF_B3)[i] = (backward || i >= n_out) ? 0.f : b3[i];
    }
}

// 256-bit global accesses (sm_100: LDG / STG .256): one full 32-byte sector per thread and instruction — a thread
// that owns 8 consecutive floats of a row moves them in one request instead of two half-sector ones
__device__ __forceinline__ void st_global_v8(float* dst, const float2 (&x)[4]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(dst), "f"(x[0].x), "f"(x[0].y), "f"(x[1].x), "f"(x[1].y),
                 "f"(x[2].x), "f"(x[2].y), "f"(x[3].x), "f"(x[3].y) : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* src, float4& u, float4& w) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w), "=f"(w.x),
                 "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(src));
}

// "Operand planes": what the weight-gradient kernel consumes.  A [rows][8 G] matrix is stored per 32-row tile as a hi
// plane then a lo plane (bf16 x = hi + lo), each G groups of [32 rows][8 columns] — i.e. already in the MN-major
// shared-memory form of tcgen05.mma, so the weight-gradient kernel moves tiles from HBM to the tensor core by TMA alone.
// The producing kernels have the hi / lo split in registers anyway (it is their own next-layer operand), and a warp's
// 32 rows of one group are 512 contiguous bytes: perfectly coalesced 16-byte stores.
// Byte offset of (row R, group g, plane p) = ((R / 32) * 2 + p) * G * 512 + g * 512 + (R % 32) * 16.
__device__ __forceinline__ uint8_t* planes_row(uint8_t* base, size_t R, int G) {
    return base ? base + (R >> 5) * (size_t)(2 * G * 512) + (R & 31) * 16 : nullptr;
}

// store_chunk + optional copy of the two 16-byte plane rows
__device__ __forceinline__ void store_chunk_save(uint32_t t_hi, uint32_t t_lo, int kc, const float2 (&v)[4], uint8_t* save_row, int G) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __float22bfloat162_rn(v[i]);
        const float2 lo = __ffma2_rn(__bfloat1622float2(hh), make_float2(-1.f, -1.f), v[i]);
        const __nv_bfloat162 ll = __float22bfloat162_rn(lo);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    tmem_st4(t_hi + kc * 4, h);
    tmem_st4(t_lo + kc * 4, l);
    if (save_row != nullptr) {
        *reinterpret_cast<uint4*>(save_row + kc * 512) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(save_row + G * 512 + kc * 512) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// hidden layer epilogue of the training kernels.  Forward: + bias, ReLU, sign bits out.  Backward: multiply by the
// saved sign bits.  Both: next layer's TMEM operand, optional operand planes for the weight-gradient kernel (rows past
// N are written as zeros: the planes are allocated in whole 128-row tiles).
template <bool BWD>
__device__ __forceinline__ void hidden_epilogue_train(uint32_t t_acc, uint32_t t_hi, uint32_t t_lo, const float* __restrict__ bias,
                                                      int half, bool valid, uint8_t* __restrict__ save_row, uint2* __restrict__ mask_slot,
                                                      uint2 mask_in) {
    const int cbase = half * (HID / 2);
    uint32_t mw[2] = {BWD ? mask_in.x : 0u, BWD ? mask_in.y : 0u};
    uint32_t r[2][16];
    tmem_ld16_async(t_acc + cbase, r[0]);
#pragma unroll
    for (int it = 0; it < HID / 32; ++it) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (it + 1 < HID / 32) tmem_ld16_async(t_acc + cbase + (it + 1) * 16, r[(it + 1) & 1]);
        const int c0 = cbase + it * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int bit = (it * 16 + h * 8 + 2 * i) & 31, word = it >> 1;      // 64 columns -> two 32-bit words
                const float2 a = make_float2(__uint_as_float(r[it & 1][h * 8 + 2 * i]), __uint_as_float(r[it & 1][h * 8 + 2 * i + 1]));
                if (!BWD) {
                    const float2 b = *reinterpret_cast<const float2*>(bias + c0 + h * 8 + 2 * i);
                    const float2 y = __fadd2_rn(a, b);
                    x[i] = valid ? make_float2(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f)) : make_float2(0.f, 0.f);
                    mw[word] |= (y.x > 0.f ? 1u : 0u) << bit;
                    mw[word] |= (y.y > 0.f ? 1u : 0u) << (bit + 1);
                } else {
                    x[i] = make_float2((mw[word] >> bit) & 1u ? a.x : 0.f, (mw[word] >> (bit + 1)) & 1u ? a.y : 0.f);
                }
            }
            store_chunk_save(t_hi, t_lo, c0 / 8 + h, x, save_row, HID / 8);
        }
    }
    if (!BWD && mask_slot != nullptr && valid) *mask_slot = make_uint2(mw[0], mw[1]);
}

struct TrainRow {
    float tpos;
    float4 v[6];       // forward: up to two 8-wide feature chunks in v[0..3]; backward: three 8-wide dy chunks
};

template <bool BWD, int K1STEPS>
__device__ __forceinline__ void load_train_row(const TrainParams& p, const TrainJob& job, int src, int nf, int half, TrainRow& r) {
    if (!BWD) {
        r.tpos = __ldg(p.tpos + src);
        const float4* frow = reinterpret_cast<const float4*>(p.feat + (size_t)src * p.feat_dim);
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (2 * c + half < nf) {
                r.v[2 * c] = __ldg(frow + 2 * (2 * c + half));
                r.v[2 * c + 1] = __ldg(frow + 2 * (2 * c + half) + 1);
            }
    } else if (K1STEPS == 3) {                 // 48 gradient columns, 16-byte aligned rows
        const float4* q = reinterpret_cast<const float4*>(job.in + (size_t)src * 48);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            r.v[2 * c] = __ldg(q + 2 * (2 * c + half));
            r.v[2 * c + 1] = __ldg(q + 2 * (2 * c + half) + 1);
        }
    } else {                                   // at most 8 gradient columns: chunk 0, owned by half 0
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = (half == 0 && i < job.n_io) ? __ldg(job.in + (size_t)src * job.n_io + i) : 0.f;
        r.v[0] = make_float4(t[0], t[1], t[2], t[3]);
        r.v[1] = make_float4(t[4], t[5], t[6], t[7]);
    }
}

template <bool BWD, int N3P, int K1STEPS>
__device__ __forceinline__ void run_train(const TrainParams& p, const TrainJob& job, const uint8_t* img, int group, int worker,
                                          int workers, uint32_t tmem_group, uint32_t mbar) {
    const int gt = threadIdx.x & (GROUP_THREADS - 1);
    const int row = gt & (ROWS - 1), half = gt >> 7, warp_in_group = gt >> 5;
    const bool issuer = warp_in_group == 0;
    const uint32_t lane_base = (uint32_t)((warp_in_group & 3) * 32) << 16;
    const uint32_t acc_u = tmem_group, hi_u = tmem_group + 128, lo_u = tmem_group + 192;
    const uint32_t t_acc = acc_u + lane_base, t_hi = hi_u + lane_base, t_lo = lo_u + lane_base;
    const uint32_t simg = smem_u32(img);
    const float* b1 = reinterpret_cast<const float*>(img + OFF_B1);
    const float* b2 = reinterpret_cast<const float*>(img + OFF_B2);
    const float* b3 = reinterpret_cast<const float*>(img + OFF_B3);
    const int N = p.N;
    const int tiles = (N + ROWS - 1) / ROWS;
    const int nf = p.feat_dim >> 3;
    uint32_t phase = 0;

    TrainRow cur;
    if (worker < tiles) {
        const int j0 = worker * ROWS + row;
        load_train_row<BWD, K1STEPS>(p, job, j0 < N ? j0 : N - 1, nf, half, cur);
    }
#pragma unroll 1
    for (int tile = worker; tile < tiles; tile += workers) {
        const int j = tile * ROWS + row;
        const bool valid = j < N;
        const size_t jr = (size_t)(valid ? j : N - 1);

        // ---- layer-1 operand
        uint8_t* const in_row = planes_row(job.save_in, (size_t)j, K1STEPS == 3 ? 6 : 2);
        if (!BWD) {
            const float d = job.zero_time ? 0.f : p.timestamp - cur.tpos;       // saro_gaussian.py:788-794
            float emb[TIME_DIMS];
            emb[0] = d;
            sincosf(d, &emb[1], &emb[2]);
#pragma unroll
            for (int f = 1; f < 4; ++f) {
                emb[1 + 2 * f] = 2.f * emb[2 * f - 1] * emb[2 * f];
                emb[2 + 2 * f] = 1.f - 2.f * emb[2 * f - 1] * emb[2 * f - 1];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int kc = 2 * c + half;
                float2 x[4];
                if (c < 2 && kc < nf) {
                    const float4 u = cur.v[2 * (c < 2 ? c : 0)], w = cur.v[2 * (c < 2 ? c : 0) + 1];
                    x[0] = make_float2(u.x, u.y); x[1] = make_float2(u.z, u.w);
                    x[2] = make_float2(w.x, w.y); x[3] = make_float2(w.z, w.w);
                } else if (kc == nf) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(emb[2 * i], emb[2 * i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(0.f, 0.f);
                    if (kc == nf + 1) x[0].x = emb[8];
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(0.f, 0.f);
                }
                store_chunk_save(t_hi, t_lo, kc, x, in_row, 6);
            }
        } else {
#pragma unroll
            for (int c = 0; c < (K1STEPS == 3 ? 3 : 1); ++c) {
                const int kc = 2 * c + half;
                const float4 u = cur.v[2 * c], w = cur.v[2 * c + 1];
                float2 x[4] = {make_float2(u.x, u.y), make_float2(u.z, u.w), make_float2(w.x, w.y), make_float2(w.z, w.w)};
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = make_float2(0.f, 0.f);
                }
                store_chunk_save(t_hi, t_lo, kc, x, in_row, K1STEPS == 3 ? 6 : 2);
            }
        }
        publish_operand(group);
        if (issuer) issue_layer<K1STEPS>(acc_u, hi_u, lo_u, simg + OFF_W1HI, simg + OFF_W1LO, HID, mbar);

        // ---- while the tensor core works: the next tile's inputs, this tile's sign bits (backward)
        TrainRow nxt = cur;
        if (tile + workers < tiles) {
            const int jn = (tile + workers) * ROWS + row;
            load_train_row<BWD, K1STEPS>(p, job, jn < N ? jn : N - 1, nf, half, nxt);
        }
        uint2 m_a = make_uint2(0u, 0u), m_b = make_uint2(0u, 0u);
        if (BWD && valid) {
            m_a = __ldg(job.mask_a + jr * 2 + half);
            m_b = __ldg(job.mask_b + jr * 2 + half);
        }

        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        hidden_epilogue_train<BWD>(t_acc, t_hi, t_lo, b1, half, valid, planes_row(job.save_a, (size_t)j, HID / 8),
                                   job.mask_a ? job.mask_a + jr * 2 + half : nullptr, m_a);
        publish_operand(group);
        if (issuer) issue_layer<HID / 16>(acc_u, hi_u, lo_u, simg + OFF_W2HI, simg + OFF_W2LO, HID, mbar);
        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        hidden_epilogue_train<BWD>(t_acc, t_hi, t_lo, b2, half, valid, planes_row(job.save_b, (size_t)j, HID / 8),
                                   job.mask_b ? job.mask_b + jr * 2 + half : nullptr, m_b);
        publish_operand(group);
        if (issuer) issue_last_layer(acc_u, hi_u, lo_u, simg + OFF_W3, N3P, mbar);
        mbar_wait(mbar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- output epilogue: accumulator columns [0, N3P) hold hi*hi + lo*hi, [N3P, 2 N3P) hold hi*lo
        if (N3P == 16) {
            float r[8], r2[8];
            tmem_ld8(t_acc, r);
            tmem_ld8(t_acc + N3P, r2);
            if (valid && half == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < job.n_io) job.out[jr * job.n_io + c] = (r[c] + r2[c]) + b3[c];
            }
        } else {
            constexpr int PER = N3P / 2;                         // 24 (forward, 48 outputs) or 16 (backward, <= 32 feature columns)
            const int width = BWD ? p.feat_dim : 48;
#pragma unroll
            for (int c0 = 0; c0 < PER; c0 += 8) {
                float r[8], r2[8];
                tmem_ld8(t_acc + half * PER + c0, r);
                tmem_ld8(t_acc + N3P + half * PER + c0, r2);
                const int col = half * PER + c0;
                if (valid && col < width) {
                    const float* bb = b3 + col;
                    const float2 y[4] = {make_float2((r[0] + r2[0]) + bb[0], (r[1] + r2[1]) + bb[1]), make_float2((r[2] + r2[2]) + bb[2], (r[3] + r2[3]) + bb[3]),
                                         make_float2((r[4] + r2[4]) + bb[4], (r[5] + r2[5]) + bb[5]), make_float2((r[6] + r2[6]) + bb[6], (r[7] + r2[7]) + bb[7])};
                    st_global_v8(job.out + jr * width + col, y);
                }
            }
        }
        cur = nxt;
    }
}

template <bool BWD>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) deform_train_kernel(const __grid_constant__ TrainParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* img = smem;
    uint64_t* mbar_p = reinterpret_cast<uint64_t*>(smem + IMG_PAD);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_p + 2);
    const int tid = threadIdx.x, warp = tid >> 5, group = tid / GROUP_THREADS;
    int ji = 0;
    while (ji + 1 < p.n_jobs && (int)blockIdx.x >= p.jobs[ji].cta_first + p.jobs[ji].cta_count) ++ji;
    const TrainJob& job = p.jobs[ji];
    const int worker = ((int)blockIdx.x - job.cta_first) * 2 + group;
    const int workers = job.cta_count * 2;
    {
        const uint4* src = reinterpret_cast<const uint4*>(job.img);
        uint4* dst = reinterpret_cast<uint4*>(img);
        for (int i = tid; i < IMG_BYTES / 16; i += 2 * GROUP_THREADS) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar_p)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar_p + 1)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t tmem_group = tmem + (uint32_t)group * 256;
    const uint32_t mbar = smem_u32(mbar_p + group);

    if (!BWD) {
        if (job.n3p == 16) run_train<false, 16, K1 / 16>(p, job, img, group, worker, workers, tmem_group, mbar);
        else run_train<false, 48, K1 / 16>(p, job, img, group, worker, workers, tmem_group, mbar);
    } else {
        if (job.n_io <= 8) run_train<true, 32, 1>(p, job, img, group, worker, workers, tmem_group, mbar);
        else run_train<true, 32, 3>(p, job, img, group, worker, workers, tmem_group, mbar);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u));
}


// ---------------------------------------------------------------------------------------------------------------------
// Weight gradients of the training path: dW = G^T X summed over the N rows, for every (job, layer).  The batch row is
// the GEMM's K dimension.  Both operands were emitted by the producing kernels as "operand planes" (above): 32-row
// tiles whose bytes ARE the MN-major shared-memory operand form of tcgen05.mma (descriptor fields pinned on hardware by
// tools/microbench/umma_mn_probe.cu), so this kernel has no arithmetic at all: one thread streams tiles HBM -> shared
// memory with TMA bulk copies (one contiguous block per operand and tile, a six-deep ring), one warp issues per tile
// 6 MMAs (hi*hi + hi*lo + lo*hi over two K-steps, f32 accumulate in TMEM across the CTA's whole tile range) plus 4 small
// ones against a constant-one operand (bias gradients), and at the end the CTA's partial [128][144] goes to global;
// a reduction kernel adds the few partials of a task in a fixed order.  Earlier versions converted float32 operands
// inside this kernel (through registers, then through a TMA-fed staging ring): both ran at ~3.3 TB/s, bound by the
// conversion's issue slots and block barriers, not by DRAM (ncu: DRAM 32 %, tensor pipe 4 %).
constexpr int WG_KT = 32;                       // batch rows per tile (2 MMA K-steps)
constexpr int WG_STAGES = 6;
constexpr int WG_THREADS = 128;
constexpr int WG_CTAS_PER_SM = 1;
constexpr int WG_MAX_TASKS = 24;
constexpr int WG_GROUP_BYTES = WG_KT * 16;      // 512: one column group of one plane
constexpr int WG_A_BYTES = 2 * 16 * WG_GROUP_BYTES;          // 16 384: hi + lo planes of a 128-column operand
constexpr int WG_STAGE_BYTES = 2 * WG_A_BYTES;  // A tile + B tile (<= 128 columns)
constexpr int WG_ONES_BYTES = 2 * WG_GROUP_BYTES;            // N = 16 operand: column 0 = 1, the rest 0
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 128;     // 197 760
constexpr int WG_BIAS_COL = 128;                // accumulator column of the bias gradient
constexpr int WG_OUT_STRIDE = 144;              // floats per row of a partial

struct WTask {
    const uint8_t* A;        // operand planes, 16 groups: becomes the M dimension
    const uint8_t* B;        // operand planes, gb groups (2, 6 or 16)
    float* partial;          // [cta_count][128][WG_OUT_STRIDE]
    int gb;
    int bias;                // also accumulate sum_r A[r][m] into column WG_BIAS_COL
    int cta_first, cta_count;
};
struct WParams {
    int tiles, n_tasks;      // 32-row tiles of the (padded) operands
    WTask tasks[WG_MAX_TASKS];
};

// 1-D bulk copy global -> shared (TMA), completion counted in bytes on an mbarrier
__device__ __forceinline__ void wg_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t wg_desc(uint32_t saddr) {      // MN-major, no swizzle: LBO = 128 (k-blocks), SBO = 512 (column groups)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(WG_GROUP_BYTES >> 4) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(WG_THREADS, WG_CTAS_PER_SM) deform_wgrad_kernel(const __grid_constant__ WParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ones = smem + WG_STAGES * WG_STAGE_BYTES;
    uint64_t* full_p = reinterpret_cast<uint64_t*>(ones + WG_ONES_BYTES);      // [WG_STAGES] "tile landed"
    uint64_t* done_p = full_p + WG_STAGES;                                      // [WG_STAGES] "MMAs of this stage finished"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_p + WG_STAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ti = 0;
    while (ti + 1 < p.n_tasks && (int)blockIdx.x >= p.tasks[ti].cta_first + p.tasks[ti].cta_count) ++ti;
    const WTask& task = p.tasks[ti];
    const int local = (int)blockIdx.x - task.cta_first;
    const int t0 = (int)((long long)p.tiles * local / task.cta_count), t1 = (int)((long long)p.tiles * (local + 1) / task.cta_count);
    const int n_t = t1 - t0;
    const uint32_t b_bytes = (uint32_t)task.gb * 2 * WG_GROUP_BYTES;
    const int nmma = task.gb * 8;

    for (int u = tid; u < WG_ONES_BYTES / 16; u += WG_THREADS)                  // group 0: column 0 = bf16(1.0) in every row
        reinterpret_cast<uint4*>(ones)[u] = make_uint4(u < WG_KT ? 0x3F80u : 0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < 2 * WG_STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(full_p + s)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 1 && lane == 0) {
        // ---- producer: tile t0 + it into stage it % WG_STAGES, as soon as the MMAs that read the stage have finished
        for (int it = 0; it < n_t; ++it) {
            const int s = it % WG_STAGES;
            if (it >= WG_STAGES) mbar_wait(smem_u32(done_p + s), (uint32_t)(it / WG_STAGES - 1) & 1u);
            const uint32_t bar = smem_u32(full_p + s);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)WG_A_BYTES + b_bytes) : "memory");
            const uint32_t dst = smem_u32(smem + s * WG_STAGE_BYTES);
            wg_bulk_load(dst, task.A + (size_t)(t0 + it) * WG_A_BYTES, WG_A_BYTES, bar);
            wg_bulk_load(dst + WG_A_BYTES, task.B + (size_t)(t0 + it) * b_bytes, b_bytes, bar);
        }
    } else if (warp == 0) {
        // ---- consumer: all lanes run the sequence, one elected lane issues (single predicated UTCHMMA per instruction)
        // D = f32, A = B = bf16, both MN-major (bits 15, 16), N at bit 17, M = 128 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(nmma >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t one_addr = smem_u32(ones);
        for (int it = 0; it < n_t; ++it) {
            const int s = it % WG_STAGES;
            mbar_wait(smem_u32(full_p + s), (uint32_t)(it / WG_STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(smem + s * WG_STAGE_BYTES), a_lo = a_hi + 16 * WG_GROUP_BYTES;
            const uint32_t b_hi = a_hi + WG_A_BYTES, b_lo = b_hi + task.gb * WG_GROUP_BYTES;
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                const uint32_t a = prod == 2 ? a_lo : a_hi, b = prod == 1 ? b_lo : b_hi;
#pragma unroll
                for (int ks = 0; ks < WG_KT / 16; ++ks) {
                    const uint32_t acc = (it > 0 || prod > 0 || ks > 0) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred pe, pa;\n\tsetp.ne.b32 pa, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
                                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
                                 :: "r"(tmem), "l"(wg_desc(a + ks * 256)), "l"(wg_desc(b + ks * 256)), "r"(idesc), "r"(acc) : "memory");
                }
            }
            if (task.bias) {
#pragma unroll
                for (int prod = 0; prod < 2; ++prod)
#pragma unroll
                    for (int ks = 0; ks < WG_KT / 16; ++ks) {
                        const uint32_t acc = (it > 0 || prod > 0 || ks > 0) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred pe, pa;\n\tsetp.ne.b32 pa, %4, 0;\n\telect.sync _|pe, 0xffffffff;\n\t"
                                     "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
                                     :: "r"(tmem + WG_BIAS_COL), "l"(wg_desc((prod ? a_lo : a_hi) + ks * 256)), "l"(wg_desc(one_addr + ks * 256)),
                                        "r"(idesc_b), "r"(acc) : "memory");
                    }
            }
            asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
                         "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                         :: "r"(smem_u32(done_p + s)) : "memory");
            __syncwarp();
        }
        // every commit the producer did not consume must still be observed before the accumulator is read
        for (int back = n_t < WG_STAGES ? n_t : WG_STAGES; back >= 1; --back) {
            const int it = n_t - back;
            mbar_wait(smem_u32(done_p + it % WG_STAGES), (uint32_t)(it / WG_STAGES) & 1u);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // partial [128][144]: warp w reads TMEM lanes 32 w ...; columns [0, nmma) and the bias block
    float* out = task.partial + (size_t)local * 128 * WG_OUT_STRIDE + (size_t)(warp * 32 + lane) * WG_OUT_STRIDE;
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < WG_OUT_STRIDE; c0 += 8) {
        if (c0 >= nmma && !(task.bias && c0 == WG_BIAS_COL)) continue;
        float v[8];
        if (n_t > 0) tmem_ld8(t_row + c0, v);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        reinterpret_cast<float4*>(out + c0)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(out + c0)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u));
}

// Adds the per-CTA partials of each task (fixed order: deterministic) and writes the parameter-shaped gradients.
// One block per (task, m); thread n owns column n.  dW[m][n] for n < cols (or dW[n][m] when transposed: the last layer's
// GEMM yields the transposed weight gradient), db[m] from column bias_col.  accumulate = 1 adds to what is there (the
// second job of an MLP that is evaluated twice); such tasks run in a second launch.
struct WReduceTask {
    const float* partial; int count;
    float* dW; int ldw; int rows; int cols; int transposed;
    float* db; int bias_col;
    int accumulate;
};
struct WReduceParams { int n_tasks; WReduceTask tasks[WG_MAX_TASKS]; };

__global__ void __launch_bounds__(160) deform_wgrad_reduce_kernel(const __grid_constant__ WReduceParams p) {
    const WReduceTask& t = p.tasks[blockIdx.x >> 7];
    const int m = blockIdx.x & 127, n = threadIdx.x;
    if (m >= t.rows || n >= WG_OUT_STRIDE) return;
    const bool is_w = n < t.cols, is_b = t.db != nullptr && n == t.bias_col;
    if (!is_w && !is_b) return;
    float sum = 0.f;
    for (int c = 0; c < t.count; ++c) sum += t.partial[((size_t)c * 128 + m) * WG_OUT_STRIDE + n];
    float* dst = is_w ? (t.transposed ? t.dW + (size_t)n * t.ldw + m : t.dW + (size_t)m * t.ldw + n) : t.db + m;
    *dst = t.accumulate ? *dst + sum : sum;
}


// ---------------------------------------------------------------------------------------------------------------------
// Elementwise epilogues of the training path (scene/saro_gaussian.py:782-831), one thread per Gaussian, forward and
// backward: lifespan from the opacity_mlp output, survival state, opacity, rotation, scale, position, real_xyz.  In the
// reference these are ~25 PyTorch kernels forward and as many in autograd's backward, each over a few bytes per Gaussian.
struct EpiParams {
    int N;
    float timestamp, min_scale;
    const float *life_raw, *motion_raw, *rot_raw, *motion_base_raw;      // MLP outputs [N][1], [N][3], [N][7], [N][3] (may be NULL)
    const float *xyz, *rotation, *scaling, *opacity, *tpos;              // model parameters / get_temporalpos
    float *o_motion, *o_rot, *o_scale, *o_opacity, *o_lifespan, *o_real_xyz;
    // backward
    const float *g_motion, *g_rot, *g_scale, *g_opacity, *g_lifespan;    // any may be NULL (no gradient arrived)
    float *d_life_raw, *d_rot_raw, *d_rotation, *d_scaling, *d_opacity, *d_tpos;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) deform_epilogue_fwd_kernel(const __grid_constant__ EpiParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    // :782-785  lifespan = (1 - min_scale) * (1 - sigmoid(z)) + min_scale
    const float life = (1.f - p.min_scale) * (1.f - sigmoidf_(p.life_raw[i])) + p.min_scale;
    p.o_lifespan[i] = life;
    // :788-789, :757-759  state = exp(-4 ((t - tpos) / lifespan)^2);  :830-831 opacity = sigmoid(o) * state
    const float u = (p.timestamp - p.tpos[i]) / life;
    p.o_opacity[i] = sigmoidf_(p.opacity[i]) * expf(-4.f * (u * u));
    float q[4], nrm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float x = p.xyz[(size_t)i * 3 + c];
        p.o_motion[(size_t)i * 3 + c] = x + p.motion_raw[(size_t)i * 3 + c];                               // :807-809
        if (p.motion_base_raw) p.o_real_xyz[(size_t)i * 3 + c] = x + p.motion_base_raw[(size_t)i * 3 + c];  // :803-804
        p.o_scale[(size_t)i * 3 + c] = expf(p.scaling[(size_t)i * 3 + c] + p.rot_raw[(size_t)i * 7 + 4 + c]);   // :819-821
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        q[c] = p.rotation[(size_t)i * 4 + c] + p.rot_raw[(size_t)i * 7 + c];                               // :813-817
        nrm += q[c] * q[c];
    }
    nrm = fmaxf(sqrtf(nrm), 1e-12f);                                                                        // F.normalize eps
#pragma unroll
    for (int c = 0; c < 4; ++c) p.o_rot[(size_t)i * 4 + c] = q[c] / nrm;
}

__global__ void __launch_bounds__(256) deform_epilogue_bwd_kernel(const __grid_constant__ EpiParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    // ---- opacity / lifespan / temporal position
    const float sz = sigmoidf_(p.life_raw[i]);
    const float life = (1.f - p.min_scale) * (1.f - sz) + p.min_scale;
    const float d = p.timestamp - p.tpos[i];
    const float u = d / life;
    const float state = expf(-4.f * (u * u));
    const float so = sigmoidf_(p.opacity[i]);
    const float g_op = p.g_opacity ? p.g_opacity[i] : 0.f;
    p.d_opacity[i] = g_op * state * so * (1.f - so);
    const float g_u = g_op * so * state * (-8.f * u);
    p.d_tpos[i] = -(g_u / life);
    const float g_life = -g_u * u / life + (p.g_lifespan ? p.g_lifespan[i] : 0.f);
    p.d_life_raw[i] = g_life * (1.f - p.min_scale) * (-(sz * (1.f - sz)));
    // ---- rotation: y = q / max(|q|, eps)  ->  dq = (g - (g . y) y) / |q|
    float q[4], g[4], nrm = 0.f, gy = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        q[c] = p.rotation[(size_t)i * 4 + c] + p.rot_raw[(size_t)i * 7 + c];
        g[c] = p.g_rot ? p.g_rot[(size_t)i * 4 + c] : 0.f;
        nrm += q[c] * q[c];
    }
    nrm = sqrtf(nrm);
    const float den = fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int c = 0; c < 4; ++c) gy += g[c] * (q[c] / den);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float dq = nrm > 1e-12f ? (g[c] - gy * (q[c] / den)) / den : g[c] / den;     // below eps the divisor is the constant
        p.d_rotation[(size_t)i * 4 + c] = dq;
        p.d_rot_raw[(size_t)i * 7 + c] = dq;
    }
    // ---- scale: y = exp(s)  ->  ds = g y
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float y = expf(p.scaling[(size_t)i * 3 + c] + p.rot_raw[(size_t)i * 7 + 4 + c]);
        const float ds = (p.g_scale ? p.g_scale[(size_t)i * 3 + c] : 0.f) * y;
        p.d_scaling[(size_t)i * 3 + c] = ds;
        p.d_rot_raw[(size_t)i * 7 + 4 + c] = ds;
    }
}

std::mutex g_mu;
int* g_pinned_count = nullptr;
cudaEvent_t g_count_ready = nullptr;
int g_sm_count = 0;
bool g_attr_set = false;

size_t select_temp_bytes(int N) {
    static std::mutex mu;
    static int cached_n = -1;
    static size_t cached_bytes = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (N == cached_n) return cached_bytes;
    size_t bytes = 0;
    cub::DeviceSelect::If(nullptr, bytes, thrust::counting_iterator<int>(0), (int*)nullptr, (int*)nullptr, N,
                          Alive{nullptr, nullptr, 0.f});
    cached_n = N;
    cached_bytes = bytes;
    return bytes;
}
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace sgs_deform

extern "C" {

size_t sgs_deform_packed_bytes(void) { return 3 * (size_t)sgs_deform::IMG_BYTES; }

size_t sgs_deform_workspace_bytes(int N) {
    if (N <= 0) return 256;
    return sgs_deform::align256((size_t)N * 4) + 256 + sgs_deform::align256(sgs_deform::select_temp_bytes(N));
}

int sgs_deform_pack_mlp(int mlp, int in_dim, const float* W1, const float* b1, const float* W2, const float* b2,
                        const float* W3, const float* b3, void* packed, void* stream) {
    using namespace sgs_deform;
    if (mlp < 0 || mlp > 2 || in_dim <= TIME_DIMS || in_dim > K1 || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !packed)
        return SGS_ERR_INVALID_ARGUMENT;
    pack_mlp_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(mlp, in_dim, W1, b1, W2, b2, W3, b3,
                                                          reinterpret_cast<uint8_t*>(packed) + (size_t)mlp * IMG_BYTES);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int64_t sgs_deform_eval(int N, int feat_dim, float timestamp, const float* xyz, const float* rotation, const float* scaling,
                        const float* opacity, const float* features_dc, const float* features_rest,
                        const float* temporal_pos, const float* lifespan, const float* hexplane_feature,
                        const void* packed, void* workspace, size_t workspace_bytes, float* out_means3D,
                        float* out_rotations, float* out_scales, float* out_opacity, float* out_shs, void* stream) {
    using namespace sgs_deform;
    if (N < 0 || feat_dim <= 0 || (feat_dim & 7) || feat_dim > 32) return SGS_ERR_INVALID_ARGUMENT;
    if (N == 0) return 0;
    if (!xyz || !rotation || !scaling || !opacity || !features_dc || !features_rest || !temporal_pos || !lifespan ||
        !hexplane_feature || !packed || !workspace || !out_means3D || !out_rotations || !out_scales || !out_opacity || !out_shs)
        return SGS_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < sgs_deform_workspace_bytes(N)) return SGS_ERR_INVALID_ARGUMENT;
    // the kernel reads rows of hexplane_feature / rotation and writes out_rotations / out_shs with 128-bit accesses:
    // the alignment contract stated in include/saro_gs_b200.h is checked here instead of faulting on the device
    if ((reinterpret_cast<size_t>(hexplane_feature) | reinterpret_cast<size_t>(rotation) |
         reinterpret_cast<size_t>(out_rotations) | reinterpret_cast<size_t>(out_shs)) & 15)
        return SGS_ERR_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_pinned_count) {
        if (cudaHostAlloc(&g_pinned_count, 64, cudaHostAllocDefault) != cudaSuccess) return SGS_ERR_ALLOC;
        if (cudaEventCreateWithFlags(&g_count_ready, cudaEventDisableTiming) != cudaSuccess) return SGS_ERR_CUDA;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    if (!g_attr_set) {
        if (cudaFuncSetAttribute(deform_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
            return SGS_ERR_CUDA;
        g_attr_set = true;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    int* index = reinterpret_cast<int*>(ws);
    int* count = reinterpret_cast<int*>(ws + align256((size_t)N * 4));
    void* temp = ws + align256((size_t)N * 4) + 256;
    size_t temp_bytes = select_temp_bytes(N);
    const Alive alive{temporal_pos, lifespan, timestamp};
    if (cub::DeviceSelect::If(temp, temp_bytes, thrust::counting_iterator<int>(0), index, count, N, alive, s) != cudaSuccess)
        return SGS_ERR_CUDA;
    if (cudaMemcpyAsync(g_pinned_count, count, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) return SGS_ERR_CUDA;
    if (cudaEventRecord(g_count_ready, s) != cudaSuccess) return SGS_ERR_CUDA;

    EvalParams p;
    p.feat_dim = feat_dim; p.timestamp = timestamp;
    p.xyz = xyz; p.rotation = rotation; p.scaling = scaling; p.opacity = opacity;
    p.features_dc = features_dc; p.features_rest = features_rest; p.tpos = temporal_pos; p.life = lifespan;
    p.feat = hexplane_feature; p.packed = reinterpret_cast<const uint8_t*>(packed); p.index = index; p.count = count;
    p.o_means3D = out_means3D; p.o_rot = out_rotations; p.o_scale = out_scales; p.o_opacity = out_opacity; p.o_shs = out_shs;
    const int tiles_max = (N + ROWS - 1) / ROWS;
    int grid = g_sm_count > 0 ? g_sm_count : 148;
    if (grid > 3 * ((tiles_max + 1) / 2)) grid = 3 * ((tiles_max + 1) / 2);
    // measured clocks per tile (SGS_DEFORM_PROFILE=1): motion : rot : shs = COST_MOTION : COST_ROT : COST_SHS
    p.ctas_shs = (int)(grid * (COST_SHS / (COST_SHS + COST_MOTION + COST_ROT)) + 0.5f);
    p.ctas_motion = (int)(grid * (COST_MOTION / (COST_SHS + COST_MOTION + COST_ROT)) + 0.5f);
    if (p.ctas_shs < 1) p.ctas_shs = 1;
    if (p.ctas_motion < 1) p.ctas_motion = 1;
    while (p.ctas_shs + p.ctas_motion >= grid) { if (p.ctas_shs > 1) --p.ctas_shs; else --p.ctas_motion; }
    static const bool profile = getenv("SGS_DEFORM_PROFILE") != nullptr;
    static long long* d_prof = nullptr;
    p.phase_clocks = nullptr;
    if (profile) {
        if (!d_prof && cudaMalloc(&d_prof, 48 * sizeof(long long)) != cudaSuccess) return SGS_ERR_ALLOC;
        cudaMemsetAsync(d_prof, 0, 48 * sizeof(long long), s);
        p.phase_clocks = d_prof;
    }
    deform_mlp_kernel<<<grid, 2 * GROUP_THREADS, SMEM_BYTES, s>>>(p);
    if (cudaGetLastError() != cudaSuccess) return SGS_ERR_CUDA;
    if (profile) {
        long long h[48];
        cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        static const char* names[14] = {"build", "publish1", "issue1", "prefetch", "wait1", "epi1", "publish2", "issue2", "wait2",
                                        "epi2", "publish3", "issue3", "wait3", "final"};
        for (int m = 0; m < 3; ++m) {
            const double tiles = (double)(h[m * 16 + 15] > 0 ? h[m * 16 + 15] : 1);
            fprintf(stderr, "[sgs_deform profile] mlp %d, %lld tiles, clocks/tile:", m, h[m * 16 + 15]);
            double tot = 0;
            for (int i = 0; i < 14; ++i) { fprintf(stderr, " %s %.0f", names[i], h[m * 16 + i] / tiles); tot += h[m * 16 + i] / tiles; }
            fprintf(stderr, " | total %.0f\n", tot);
        }
    }
    // the host needs the number of selected Gaussians to shape the rasterizer call; it is ready as soon as the
    // selection pass is, while the MLP kernel keeps running
    if (cudaEventSynchronize(g_count_ready) != cudaSuccess) return SGS_ERR_CUDA;
    return (int64_t)*g_pinned_count;
}


size_t sgs_deform_image_bytes(void) { return (size_t)sgs_deform::IMG_BYTES; }

int sgs_deform_pack_general(int backward, int in_w, int hid2, int n_out, int feat_dim, const float* W1, const float* b1,
                            const float* W2, const float* b2, const float* W3, const float* b3, void* image, void* stream) {
    using namespace sgs_deform;
    if (in_w <= 0 || in_w > K1 || hid2 <= 0 || hid2 > HID || n_out <= 0 || n_out > N3_MAX || (n_out > 8 && n_out != 48) ||
        feat_dim <= 0 || (feat_dim & 7) || feat_dim > 32 || feat_dim > in_w || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !image)
        return SGS_ERR_INVALID_ARGUMENT;
    const int n3p = backward ? 32 : (n_out <= 8 ? 16 : 48);
    pack_general_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(backward, in_w, hid2, n_out, feat_dim, n3p, W1, b1, W2, b2, W3, b3,
                                                              reinterpret_cast<uint8_t*>(image));
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

static int sgs_deform_train_launch(bool backward, int N, int feat_dim, float timestamp, const float* temporal_pos,
                                   const float* feature, int n_jobs, const sgs_mlp_job_t* jobs, void* stream) {
    using namespace sgs_deform;
    if (N < 0 || feat_dim <= 0 || (feat_dim & 7) || feat_dim > 32 || n_jobs <= 0 || n_jobs > MAX_JOBS || !jobs)
        return SGS_ERR_INVALID_ARGUMENT;
    if (N == 0) return 0;
    if (!backward && (!temporal_pos || !feature || (reinterpret_cast<size_t>(feature) & 15))) return SGS_ERR_INVALID_ARGUMENT;
    static int sm_count = 0;
    static bool attr_set[2] = {false, false};
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (sm_count == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        }
        if (!attr_set[backward]) {
            const cudaError_t e = backward
                ? cudaFuncSetAttribute(deform_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAIN_SMEM_BYTES)
                : cudaFuncSetAttribute(deform_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRAIN_SMEM_BYTES);
            if (e != cudaSuccess) return SGS_ERR_CUDA;
            attr_set[backward] = true;
        }
    }
    TrainParams p;
    p.N = N; p.feat_dim = feat_dim; p.n_jobs = n_jobs; p.timestamp = timestamp; p.tpos = temporal_pos; p.feat = feature;
    float cost[MAX_JOBS], total = 0.f;
    for (int i = 0; i < n_jobs; ++i) {
        const sgs_mlp_job_t& j = jobs[i];
        if (!j.packed || !j.out || j.n_io <= 0 || (j.n_io > 8 && j.n_io != 48)) return SGS_ERR_INVALID_ARGUMENT;
        if (backward && (!j.in || !j.mask_a || !j.mask_b)) return SGS_ERR_INVALID_ARGUMENT;
        if ((reinterpret_cast<size_t>(j.in) | reinterpret_cast<size_t>(j.mask_a) | reinterpret_cast<size_t>(j.mask_b)) & 15)
            return SGS_ERR_INVALID_ARGUMENT;
        if ((reinterpret_cast<size_t>(j.save_a) | reinterpret_cast<size_t>(j.save_b) | reinterpret_cast<size_t>(j.save_in)) & 15)
            return SGS_ERR_INVALID_ARGUMENT;
        if (((backward || j.n_io > 8) ? reinterpret_cast<size_t>(j.out) : 0) & 31)            // 256-bit stores
            return SGS_ERR_INVALID_ARGUMENT;
        TrainJob& t = p.jobs[i];
        t.img = reinterpret_cast<const uint8_t*>(j.packed);
        t.in = j.in; t.out = j.out;
        t.save_a = reinterpret_cast<uint8_t*>(j.save_a); t.save_b = reinterpret_cast<uint8_t*>(j.save_b);
        t.save_in = reinterpret_cast<uint8_t*>(j.save_in);
        t.mask_a = reinterpret_cast<uint2*>(j.mask_a); t.mask_b = reinterpret_cast<uint2*>(j.mask_b);
        t.n_io = j.n_io; t.n3p = backward ? 32 : (j.n_io <= 8 ? 16 : 48); t.zero_time = j.zero_time;
        cost[i] = (j.n_io > 8 ? COST_SHS : COST_ROT) + ((j.save_a || j.save_b) ? 2.f : 0.f);
        total += cost[i];
    }
    const int tiles = (N + ROWS - 1) / ROWS;
    int grid = sm_count > 0 ? sm_count : 148;
    const int useful = n_jobs * ((tiles + 1) / 2);
    if (grid > useful) grid = useful;
    if (grid < n_jobs) grid = n_jobs;
    // CTAs in proportion to the jobs' cost per tile, at least one each (largest remainder to the costliest jobs)
    int given = 0;
    for (int i = 0; i < n_jobs; ++i) {
        int c = (int)(grid * (cost[i] / total));
        if (c < 1) c = 1;
        p.jobs[i].cta_count = c;
        given += c;
    }
    for (int i = 0; given < grid; i = (i + 1) % n_jobs) { ++p.jobs[i].cta_count; ++given; }
    for (int i = 0; given > grid; i = (i + 1) % n_jobs)
        if (p.jobs[i].cta_count > 1) { --p.jobs[i].cta_count; --given; }
    int first = 0;
    for (int i = 0; i < n_jobs; ++i) { p.jobs[i].cta_first = first; first += p.jobs[i].cta_count; }
    cudaStream_t s = (cudaStream_t)stream;
    if (backward) deform_train_kernel<true><<<grid, 2 * GROUP_THREADS, TRAIN_SMEM_BYTES, s>>>(p);
    else deform_train_kernel<false><<<grid, 2 * GROUP_THREADS, TRAIN_SMEM_BYTES, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_deform_train_forward(int N, int feat_dim, float timestamp, const float* temporal_pos, const float* feature, int n_jobs,
                             const sgs_mlp_job_t* jobs, void* stream) {
    return sgs_deform_train_launch(false, N, feat_dim, timestamp, temporal_pos, feature, n_jobs, jobs, stream);
}

int sgs_deform_train_backward(int N, int feat_dim, int n_jobs, const sgs_mlp_job_t* jobs, void* stream) {
    return sgs_deform_train_launch(true, N, feat_dim, 0.f, nullptr, nullptr, n_jobs, jobs, stream);
}


int sgs_deform_wgrad_max_ctas(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    return (sms > 0 ? sms : 148) * sgs_deform::WG_CTAS_PER_SM;
}

size_t sgs_deform_wgrad_partial_floats(void) { return (size_t)128 * sgs_deform::WG_OUT_STRIDE; }

size_t sgs_deform_planes_bytes(int N, int groups) {
    if (N <= 0 || groups <= 0) return 0;
    return (size_t)((N + sgs_deform::ROWS - 1) / sgs_deform::ROWS) * 4 * 2 * (size_t)groups * 512;
}

int sgs_deform_wgrad(int N, int n_tasks, const sgs_wgrad_task_t* tasks, float* partials, void* stream) {
    using namespace sgs_deform;
    if (N <= 0 || n_tasks <= 0 || n_tasks > WG_MAX_TASKS || !tasks || !partials) return SGS_ERR_INVALID_ARGUMENT;
    static bool attr_set = false;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!attr_set) {
            if (cudaFuncSetAttribute(deform_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES) != cudaSuccess)
                return SGS_ERR_CUDA;
            attr_set = true;
        }
    }
    const int tiles = (N + ROWS - 1) / ROWS * (ROWS / WG_KT);          // the planes are allocated in whole 128-row tiles
    int grid = sgs_deform_wgrad_max_ctas();
    if (grid < n_tasks) return SGS_ERR_INVALID_ARGUMENT;
    if ((long long)grid > (long long)n_tasks * tiles) grid = n_tasks * tiles;
    WParams p;
    p.tiles = tiles; p.n_tasks = n_tasks;
    float cost[WG_MAX_TASKS], total = 0.f;
    for (int i = 0; i < n_tasks; ++i) {
        const sgs_wgrad_task_t& t = tasks[i];
        if (!t.A || !t.B || (t.groups_b != 2 && t.groups_b != 6 && t.groups_b != 16) ||
            ((reinterpret_cast<size_t>(t.A) | reinterpret_cast<size_t>(t.B)) & 15) || !t.dW || t.rows <= 0 || t.rows > 128 ||
            t.cols <= 0 || t.cols > t.groups_b * 8)
            return SGS_ERR_INVALID_ARGUMENT;
        WTask& w = p.tasks[i];
        w.A = reinterpret_cast<const uint8_t*>(t.A); w.B = reinterpret_cast<const uint8_t*>(t.B);
        w.gb = t.groups_b; w.bias = t.db != nullptr;
        cost[i] = 16.f + (float)t.groups_b + 4.f;      // bytes per tile, plus a constant for the per-tile MMA issue
        total += cost[i];
    }
    int given = 0;
    for (int i = 0; i < n_tasks; ++i) {
        int c = (int)(grid * (cost[i] / total));
        if (c < 1) c = 1;
        if (c > tiles) c = tiles;
        p.tasks[i].cta_count = c;
        given += c;
    }
    for (int i = 0, guard = 0; given < grid && guard < 4 * grid; i = (i + 1) % n_tasks, ++guard)
        if (p.tasks[i].cta_count < tiles) { ++p.tasks[i].cta_count; ++given; }
    for (int i = 0; given > grid; i = (i + 1) % n_tasks)
        if (p.tasks[i].cta_count > 1) { --p.tasks[i].cta_count; --given; }
    grid = given;
    int first = 0;
    for (int i = 0; i < n_tasks; ++i) {
        p.tasks[i].cta_first = first;
        p.tasks[i].partial = partials + (size_t)first * 128 * WG_OUT_STRIDE;
        first += p.tasks[i].cta_count;
    }
    cudaStream_t s = (cudaStream_t)stream;
    deform_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, s>>>(p);
    if (cudaGetLastError() != cudaSuccess) return SGS_ERR_CUDA;
    for (int pass = 0; pass < 2; ++pass) {
        WReduceParams r;
        r.n_tasks = 0;
        for (int i = 0; i < n_tasks; ++i) {
            if ((tasks[i].accumulate != 0) != (pass == 1)) continue;
            WReduceTask& q = r.tasks[r.n_tasks++];
            q.partial = p.tasks[i].partial; q.count = p.tasks[i].cta_count;
            q.dW = tasks[i].dW; q.ldw = tasks[i].ldw; q.rows = tasks[i].rows; q.cols = tasks[i].cols; q.transposed = tasks[i].transposed;
            q.db = tasks[i].db; q.bias_col = WG_BIAS_COL; q.accumulate = pass;
        }
        if (r.n_tasks > 0) deform_wgrad_reduce_kernel<<<r.n_tasks * 128, 160, 0, s>>>(r);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}


int sgs_deform_train_epilogue_forward(int N, float timestamp, float min_scale, const float* life_raw, const float* motion_raw,
                                      const float* rot_raw, const float* motion_base_raw, const float* xyz, const float* rotation,
                                      const float* scaling, const float* opacity, const float* temporal_pos, float* out_means3D,
                                      float* out_rotations, float* out_scales, float* out_opacity, float* out_lifespan,
                                      float* out_real_xyz, void* stream) {
    using namespace sgs_deform;
    if (N < 0) return SGS_ERR_INVALID_ARGUMENT;
    if (N == 0) return 0;
    if (!life_raw || !motion_raw || !rot_raw || !xyz || !rotation || !scaling || !opacity || !temporal_pos || !out_means3D ||
        !out_rotations || !out_scales || !out_opacity || !out_lifespan || (motion_base_raw && !out_real_xyz))
        return SGS_ERR_INVALID_ARGUMENT;
    EpiParams p = {};
    p.N = N; p.timestamp = timestamp; p.min_scale = min_scale;
    p.life_raw = life_raw; p.motion_raw = motion_raw; p.rot_raw = rot_raw; p.motion_base_raw = motion_base_raw;
    p.xyz = xyz; p.rotation = rotation; p.scaling = scaling; p.opacity = opacity; p.tpos = temporal_pos;
    p.o_motion = out_means3D; p.o_rot = out_rotations; p.o_scale = out_scales; p.o_opacity = out_opacity;
    p.o_lifespan = out_lifespan; p.o_real_xyz = out_real_xyz;
    deform_epilogue_fwd_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

int sgs_deform_train_epilogue_backward(int N, float timestamp, float min_scale, const float* life_raw, const float* rot_raw,
                                       const float* rotation, const float* scaling, const float* opacity, const float* temporal_pos,
                                       const float* g_rotations, const float* g_scales, const float* g_opacity,
                                       const float* g_lifespan, float* d_life_raw, float* d_rot_raw, float* d_rotation,
                                       float* d_scaling, float* d_opacity, float* d_temporal_pos, void* stream) {
    using namespace sgs_deform;
    if (N < 0) return SGS_ERR_INVALID_ARGUMENT;
    if (N == 0) return 0;
    if (!life_raw || !rot_raw || !rotation || !scaling || !opacity || !temporal_pos || !d_life_raw || !d_rot_raw || !d_rotation ||
        !d_scaling || !d_opacity || !d_temporal_pos)
        return SGS_ERR_INVALID_ARGUMENT;
    EpiParams p = {};
    p.N = N; p.timestamp = timestamp; p.min_scale = min_scale;
    p.life_raw = life_raw; p.rot_raw = rot_raw; p.rotation = rotation; p.scaling = scaling; p.opacity = opacity; p.tpos = temporal_pos;
    p.g_rot = g_rotations; p.g_scale = g_scales; p.g_opacity = g_opacity; p.g_lifespan = g_lifespan;
    p.d_life_raw = d_life_raw; p.d_rot_raw = d_rot_raw; p.d_rotation = d_rotation; p.d_scaling = d_scaling; p.d_opacity = d_opacity;
    p.d_tpos = d_temporal_pos;
    deform_epilogue_bwd_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : SGS_ERR_CUDA;
}

}  // extern "C"
