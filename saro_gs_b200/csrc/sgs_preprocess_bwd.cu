// Backward per-Gaussian preprocess: ONE fused kernel for what the reference runs as
// computeCov2DCUDA + preprocessCUDA (backward) + nine torch::zeros fills.
//
// Math follows
//   dL/dconic -> dL/dcov2D -> dL/dcov3D, dL/dT -> dL/dJ -> dL/dt -> dL/dmean3D   $R/cuda_rasterizer/backward.cu:144-274
//   dL/dmean2D -> dL/dmean3D through the projection                             $R/cuda_rasterizer/backward.cu:365-387
//   SH backward (clamp mask, dL/dsh, view-direction term into dL/dmean3D)       $R/cuda_rasterizer/backward.cu:20-139
//   dL/dcov3D -> dL/dscale, dL/dquaternion (no normalisation Jacobian)          $R/cuda_rasterizer/backward.cu:278-341
//
// Inputs: `acc` [P][12] = per-Gaussian MOMENT sums produced by the backward render kernel over all
// blended pixel x instance pairs, with g = G dL/dalpha, w = alpha T, d = mean2D - pixel:
//   slots 0..5 = sum g dx, sum g dy, sum g dx^2, sum g dx dy, sum g dy^2, sum g ; 6..8 = sum w dL/dpix[c].
// The per-Gaussian factors of $R/cuda_rasterizer/backward.cu:536-554 (opacity, conic, 0.5W/0.5H, -0.5) are
// applied here, once per Gaussian, instead of once per pair.
// Every output element is written here (zeros for culled Gaussians), so callers pass
// uninitialised tensors — no memset traffic for the ~300 B/Gaussian of outputs.
#include "sgs_common.cuh"

namespace sgs {

struct F3 {
    float x, y, z;
};
__forceinline__ __device__ F3 operator*(float s, const F3& a) { return {s * a.x, s * a.y, s * a.z}; }
__forceinline__ __device__ F3 operator+(const F3& a, const F3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__forceinline__ __device__ float dot3(const F3& a, const F3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// d normalize(v) / d v applied to dv   ($R/cuda_rasterizer/auxiliary.h:107-117)
__forceinline__ __device__ float3 dnormvdv3(float3 v, float3 dv) {
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

#define SGS_PRE_THREADS 128
#define SGS_SH_ROW4 12          // float4 per 16-coefficient SH row (192 B)
#define SGS_SH_PAD4 13          // padded row stride in shared memory (conflict-free 128-bit accesses)

template <bool VEC_SH>
__global__ void __launch_bounds__(SGS_PRE_THREADS)
preprocess_bwd_kernel(int P, const __grid_constant__ ViewParams vp, const float* __restrict__ means3D,
                      const int* __restrict__ radii, const float* __restrict__ shs,
                      const float* __restrict__ scales, const float* __restrict__ rotations,
                      const float* __restrict__ cov3Ds, const uint8_t* __restrict__ clamped,
                      const float4* __restrict__ conic_opacity, float* __restrict__ acc, float* __restrict__ dL_dmean2D,
                      float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                      float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dscale, float* __restrict__ dL_drot, int rot_vec, const DensifySink sink) {
    __shared__ ViewSmem cam;
    // SH rows travel through shared memory so that both the 192-B reads and the 192-B gradient writes of a
    // warp are fully coalesced (32 consecutive rows = 6 KB contiguous)
    __shared__ float4 s_sh[VEC_SH ? (SGS_PRE_THREADS / 32) * 32 * SGS_SH_PAD4 : 1];
    stage_view(cam, vp);     // camera constants: inputs of the call, not written by the predecessor
    pdl_wait();              // launched programmatically dependent on the backward render kernel (writes `acc`)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < P;
    const int M = vp.sh_coeffs;
    // latency-bound kernel (ncu: 51 % of the samples wait on global loads at 23 % occupancy): every load of this
    // Gaussian is issued up front, unconditionally, so that all of them are in flight together
    int pre_radius = 0;
    float4 pre_a0 = {0.f, 0.f, 0.f, 0.f}, pre_a1 = pre_a0, pre_a2 = pre_a0, pre_co = pre_a0, pre_q = pre_a0;
    float3 pre_mean = {0.f, 0.f, 0.f}, pre_sc = {0.f, 0.f, 0.f};
    float pre_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint8_t pre_cm = 0;
    if (valid) {
        pre_radius = radii[idx];
        float4* arow = reinterpret_cast<float4*>(acc + (size_t)idx * 12);
        pre_a0 = arow[0]; pre_a1 = arow[1]; pre_a2 = arow[2];
        pre_co = conic_opacity[idx];
        pre_mean = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
#pragma unroll
        for (int i = 0; i < 6; i++) pre_cov[i] = cov3Ds[6 * (size_t)idx + i];
        pre_cm = clamped[idx];
        if (scales != nullptr) {
            pre_sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
            if (rot_vec) pre_q = reinterpret_cast<const float4*>(rotations)[idx];
            else pre_q = make_float4(rotations[4 * idx], rotations[4 * idx + 1], rotations[4 * idx + 2], rotations[4 * idx + 3]);
        }
    }
    const bool visible = valid && pre_radius > 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* rows = s_sh + (VEC_SH ? warp * 32 * SGS_SH_PAD4 : 0);
    const int first_row = blockIdx.x * blockDim.x + warp * 32;
    const int nrows = min(32, P - first_row);
    if (VEC_SH && shs != nullptr && nrows > 0) {
        const float4* src = reinterpret_cast<const float4*>(shs + (size_t)first_row * 48);
        const int total = nrows * SGS_SH_ROW4;
#pragma unroll
        for (int it = 0; it < SGS_SH_ROW4; it++) {
            const int e = it * 32 + lane;
            if (e < total) {
                const int row = e / SGS_SH_ROW4, c = e - row * SGS_SH_ROW4;
                rows[row * SGS_SH_PAD4 + c] = __ldg(src + e);
            }
        }
        __syncwarp();
    }

    float o_mean2D[3] = {0.f, 0.f, 0.f};
    float o_opacity = 0.f;
    float o_color[3] = {0.f, 0.f, 0.f};
    float o_mean3D[3] = {0.f, 0.f, 0.f};
    float o_cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float o_scale[3] = {0.f, 0.f, 0.f};
    float o_rot[4] = {0.f, 0.f, 0.f, 0.f};
    F3 o_sh[16];
#pragma unroll
    for (int k = 0; k < 16; k++) o_sh[k] = {0.f, 0.f, 0.f};

    if (visible) {
        const float4 a0 = pre_a0, a1 = pre_a1, a2 = pre_a2;
        const float4 co = pre_co;   // A, B, C, opacity
        // $R/cuda_rasterizer/backward.cu:460-461 (double product rounded to float once)
        const float ddelx_dx = (float)(0.5 * vp.W), ddely_dy = (float)(0.5 * vp.H);
        const float Sx = a0.x, Sy = a0.y, Sxx = a0.z, Sxy = a0.w, Syy = a1.x;
        o_mean2D[0] = -co.w * ddelx_dx * (co.x * Sx + co.y * Sy);
        o_mean2D[1] = -co.w * ddely_dy * (co.z * Sy + co.y * Sx);
        const float hw = -0.5f * co.w;
        const float3 dL_dconic = {hw * Sxx, hw * Sxy, hw * Syy};
        o_opacity = a1.y;
        o_color[0] = a1.z;
        o_color[1] = a1.w;
        o_color[2] = a2.x;

        const float3 mean = pre_mean;
        const float* cov3D = pre_cov;
        const float* view = cam.view;
        const float h_x = vp.focal_x, h_y = vp.focal_y;

        // ---- cov2D / conic backward ---------------------------------------------------------
        float3 t = xform_point_4x3(mean, view);
        const float limx = 1.3f * vp.tan_fovx;
        const float limy = 1.3f * vp.tan_fovy;
        const float txtz = t.x / t.z;
        const float tytz = t.y / t.z;
        t.x = min(limx, max(-limx, txtz)) * t.z;
        t.y = min(limy, max(-limy, tytz)) * t.z;
        const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
        const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

        Mat3 J = mat3_cols(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z),
                           0, 0, 0);
        Mat3 Wm = mat3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
        Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
        Mat3 Tm = mat3_mul(Wm, J);
        Mat3 cov2D = mat3_mul(mat3_mul(mat3_T(Tm), mat3_T(Vrk)), Tm);

        const float a = cov2D.c[0][0] + 0.3f;
        const float b = cov2D.c[0][1];
        const float c = cov2D.c[1][1] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
            dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
            dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

            const float T00 = Tm.c[0][0], T01 = Tm.c[0][1], T02 = Tm.c[0][2];
            const float T10 = Tm.c[1][0], T11 = Tm.c[1][1], T12 = Tm.c[1][2];
            o_cov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            o_cov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            o_cov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            o_cov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            o_cov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            o_cov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        }

        // dL/dT (upper 2x3), with V = Vrk (symmetric): rowk(T) . colj(V)
        float dL_dT0[3], dL_dT1[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float r0 = Tm.c[0][0] * Vrk.c[j][0] + Tm.c[0][1] * Vrk.c[j][1] + Tm.c[0][2] * Vrk.c[j][2];
            const float r1 = Tm.c[1][0] * Vrk.c[j][0] + Tm.c[1][1] * Vrk.c[j][1] + Tm.c[1][2] * Vrk.c[j][2];
            dL_dT0[j] = 2 * r0 * dL_da + r1 * dL_db;
            dL_dT1[j] = 2 * r1 * dL_dc + r0 * dL_db;
        }
        const float dL_dJ00 = Wm.c[0][0] * dL_dT0[0] + Wm.c[0][1] * dL_dT0[1] + Wm.c[0][2] * dL_dT0[2];
        const float dL_dJ02 = Wm.c[2][0] * dL_dT0[0] + Wm.c[2][1] * dL_dT0[1] + Wm.c[2][2] * dL_dT0[2];
        const float dL_dJ11 = Wm.c[1][0] * dL_dT1[0] + Wm.c[1][1] * dL_dT1[1] + Wm.c[1][2] * dL_dT1[2];
        const float dL_dJ12 = Wm.c[2][0] * dL_dT1[0] + Wm.c[2][1] * dL_dT1[1] + Wm.c[2][2] * dL_dT1[2];

        const float tz = 1.f / t.z;
        const float tz2 = tz * tz;
        const float tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                             (2 * h_y * t.y) * tz3 * dL_dJ12;
        // t = view . mean  =>  dL/dmean = view_rot^T . dL/dt
        float dmx = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        float dmy = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        float dmz = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        // ---- projection backward ------------------------------------------------------------
        {
            const float* proj = cam.proj;
            const float4 m_hom = xform_point_4x4(mean, proj);
            const float m_w = 1.0f / (m_hom.w + 0.0000001f);
            const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
            const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
            const float gx = o_mean2D[0], gy = o_mean2D[1];
            dmx += (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
            dmy += (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
            dmz += (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
        }

        // ---- SH backward ----------------------------------------------------------------------
        if (shs != nullptr) {
            const int deg = vp.sh_degree;
            const float3 dir_orig = {mean.x - cam.campos[0], mean.y - cam.campos[1], mean.z - cam.campos[2]};
            const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
            const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;

            F3 sh[16];
            if (VEC_SH) {
                float f[48];
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    const float4 v = rows[lane * SGS_SH_PAD4 + k];
                    f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < 16; k++) sh[k] = {f[3 * k], f[3 * k + 1], f[3 * k + 2]};
            } else {
                const float* row = shs + (size_t)idx * M * 3;
                const int ncoef = (deg + 1) * (deg + 1);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k < ncoef) sh[k] = {row[3 * k], row[3 * k + 1], row[3 * k + 2]};
                    else sh[k] = {0.f, 0.f, 0.f};
                }
            }
            const uint8_t cm = pre_cm;
            F3 dL_dRGB = {o_color[0], o_color[1], o_color[2]};
            dL_dRGB.x *= (cm & 1) ? 0 : 1;
            dL_dRGB.y *= (cm & 2) ? 0 : 1;
            dL_dRGB.z *= (cm & 4) ? 0 : 1;

            F3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
            o_sh[0] = SGS_SH_C0 * dL_dRGB;
            if (deg > 0) {
                o_sh[1] = (-SGS_SH_C1 * y) * dL_dRGB;
                o_sh[2] = (SGS_SH_C1 * z) * dL_dRGB;
                o_sh[3] = (-SGS_SH_C1 * x) * dL_dRGB;
                dRGBdx = -SGS_SH_C1 * sh[3];
                dRGBdy = -SGS_SH_C1 * sh[1];
                dRGBdz = SGS_SH_C1 * sh[2];
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z;
                    const float xy = x * y, yz = y * z, xz = x * z;
                    o_sh[4] = (SGS_SH_C2_0 * xy) * dL_dRGB;
                    o_sh[5] = (SGS_SH_C2_1 * yz) * dL_dRGB;
                    o_sh[6] = (SGS_SH_C2_2 * (2.f * zz - xx - yy)) * dL_dRGB;
                    o_sh[7] = (SGS_SH_C2_3 * xz) * dL_dRGB;
                    o_sh[8] = (SGS_SH_C2_4 * (xx - yy)) * dL_dRGB;
                    dRGBdx = dRGBdx + (SGS_SH_C2_0 * y * sh[4] + SGS_SH_C2_2 * 2.f * -x * sh[6] + SGS_SH_C2_3 * z * sh[7] +
                                       SGS_SH_C2_4 * 2.f * x * sh[8]);
                    dRGBdy = dRGBdy + (SGS_SH_C2_0 * x * sh[4] + SGS_SH_C2_1 * z * sh[5] + SGS_SH_C2_2 * 2.f * -y * sh[6] +
                                       SGS_SH_C2_4 * 2.f * -y * sh[8]);
                    dRGBdz = dRGBdz + (SGS_SH_C2_1 * y * sh[5] + SGS_SH_C2_2 * 2.f * 2.f * z * sh[6] + SGS_SH_C2_3 * x * sh[7]);
                    if (deg > 2) {
                        o_sh[9] = (SGS_SH_C3_0 * y * (3.f * xx - yy)) * dL_dRGB;
                        o_sh[10] = (SGS_SH_C3_1 * xy * z) * dL_dRGB;
                        o_sh[11] = (SGS_SH_C3_2 * y * (4.f * zz - xx - yy)) * dL_dRGB;
                        o_sh[12] = (SGS_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB;
                        o_sh[13] = (SGS_SH_C3_4 * x * (4.f * zz - xx - yy)) * dL_dRGB;
                        o_sh[14] = (SGS_SH_C3_5 * z * (xx - yy)) * dL_dRGB;
                        o_sh[15] = (SGS_SH_C3_6 * x * (xx - 3.f * yy)) * dL_dRGB;
                        dRGBdx = dRGBdx + ((SGS_SH_C3_0 * 3.f * 2.f * xy) * sh[9] + (SGS_SH_C3_1 * yz) * sh[10] +
                                           (SGS_SH_C3_2 * -2.f * xy) * sh[11] + (SGS_SH_C3_3 * -3.f * 2.f * xz) * sh[12] +
                                           (SGS_SH_C3_4 * (-3.f * xx + 4.f * zz - yy)) * sh[13] +
                                           (SGS_SH_C3_5 * 2.f * xz) * sh[14] + (SGS_SH_C3_6 * 3.f * (xx - yy)) * sh[15]);
                        dRGBdy = dRGBdy + ((SGS_SH_C3_0 * 3.f * (xx - yy)) * sh[9] + (SGS_SH_C3_1 * xz) * sh[10] +
                                           (SGS_SH_C3_2 * (-3.f * yy + 4.f * zz - xx)) * sh[11] +
                                           (SGS_SH_C3_3 * -3.f * 2.f * yz) * sh[12] + (SGS_SH_C3_4 * -2.f * xy) * sh[13] +
                                           (SGS_SH_C3_5 * -2.f * yz) * sh[14] + (SGS_SH_C3_6 * -3.f * 2.f * xy) * sh[15]);
                        dRGBdz = dRGBdz + ((SGS_SH_C3_1 * xy) * sh[10] + (SGS_SH_C3_2 * 4.f * 2.f * yz) * sh[11] +
                                           (SGS_SH_C3_3 * 3.f * (2.f * zz - xx - yy)) * sh[12] +
                                           (SGS_SH_C3_4 * 4.f * 2.f * xz) * sh[13] + (SGS_SH_C3_5 * (xx - yy)) * sh[14]);
                    }
                }
            }
            const float3 dL_ddir = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
            const float3 dm = dnormvdv3(dir_orig, dL_ddir);
            dmx += dm.x;
            dmy += dm.y;
            dmz += dm.z;
        }
        o_mean3D[0] = dmx;
        o_mean3D[1] = dmy;
        o_mean3D[2] = dmz;

        // ---- cov3D -> scale / rotation --------------------------------------------------------
        if (scales != nullptr) {
            const float3 sc = pre_sc;
            const float4 q = pre_q;
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            Mat3 R = mat3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                               2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                               2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
            const float s0 = vp.scale_modifier * sc.x, s1 = vp.scale_modifier * sc.y, s2 = vp.scale_modifier * sc.z;
            // M = S R in column-major terms: M.c[i][j] = s_j * R.c[i][j]
            Mat3 Mm;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                Mm.c[i][0] = s0 * R.c[i][0];
                Mm.c[i][1] = s1 * R.c[i][1];
                Mm.c[i][2] = s2 * R.c[i][2];
            }
            Mat3 dSig = mat3_cols(o_cov[0], 0.5f * o_cov[1], 0.5f * o_cov[2], 0.5f * o_cov[1], o_cov[3], 0.5f * o_cov[4],
                                  0.5f * o_cov[2], 0.5f * o_cov[4], o_cov[5]);
            Mat3 M2;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) M2.c[i][j] = 2.0f * Mm.c[i][j];
            Mat3 dL_dM = mat3_mul(M2, dSig);
            Mat3 Rt = mat3_T(R);
            Mat3 dMt = mat3_T(dL_dM);
            o_scale[0] = Rt.c[0][0] * dMt.c[0][0] + Rt.c[0][1] * dMt.c[0][1] + Rt.c[0][2] * dMt.c[0][2];
            o_scale[1] = Rt.c[1][0] * dMt.c[1][0] + Rt.c[1][1] * dMt.c[1][1] + Rt.c[1][2] * dMt.c[1][2];
            o_scale[2] = Rt.c[2][0] * dMt.c[2][0] + Rt.c[2][1] * dMt.c[2][1] + Rt.c[2][2] * dMt.c[2][2];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                dMt.c[0][j] *= s0;
                dMt.c[1][j] *= s1;
                dMt.c[2][j] *= s2;
            }
            o_rot[0] = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) +
                       2 * x * (dMt.c[1][2] - dMt.c[2][1]);
            o_rot[1] = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) +
                       2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
            o_rot[2] = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) +
                       2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
            o_rot[3] = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) +
                       2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
        }
    }

    // ---- write everything (zeros when culled) --------------------------------------------------
    if (valid) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            dL_dmean2D[3 * idx + i] = o_mean2D[i];
            dL_dcolor[3 * idx + i] = o_color[i];
            dL_dmean3D[3 * idx + i] = o_mean3D[i];
            dL_dscale[3 * idx + i] = o_scale[i];
        }
        dL_dopacity[idx] = o_opacity;
        // leave the moment accumulator of this Gaussian zeroed again (only visible Gaussians ever receive atomics): a
        // second backward pass over the same forward state starts clean without any memset (sgs_common.cuh: GeomState)
        if (pre_radius > 0) {
            float4* arow = reinterpret_cast<float4*>(acc + (size_t)idx * 12);
            arow[0] = arow[1] = arow[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // densification statistics of this view (train.py:211-215 of the reference) as an epilogue: the means2D gradient
        // and the radius are in registers, so sgs_densify_add_view costs no launch and no re-read (csrc/sgs_densify.cu
        // holds the stand-alone kernel with the same arithmetic).  Culled Gaussians would add 0 / 0 / max(., 0).
        if (sink.grad_sum && pre_radius > 0) {
            sink.grad_sum[idx] += sqrtf(o_mean2D[0] * o_mean2D[0] + o_mean2D[1] * o_mean2D[1]);
            sink.vis_count[idx] += 1;
            sink.radii_max[idx] = max(sink.radii_max[idx], pre_radius);
        }
#pragma unroll
        for (int i = 0; i < 6; i++) dL_dcov3D[6 * idx + i] = o_cov[i];
        if (rot_vec) {
            reinterpret_cast<float4*>(dL_drot)[idx] = make_float4(o_rot[0], o_rot[1], o_rot[2], o_rot[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) dL_drot[4 * idx + i] = o_rot[i];
        }
    }
    if (M > 0) {
        if (VEC_SH) {
            // every lane has consumed its input row: reuse the buffer for the gradient rows
            __syncwarp();
            float f[48];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                f[3 * k] = o_sh[k].x; f[3 * k + 1] = o_sh[k].y; f[3 * k + 2] = o_sh[k].z;
            }
#pragma unroll
            for (int k = 0; k < 12; k++)
                rows[lane * SGS_SH_PAD4 + k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
            __syncwarp();
            if (nrows > 0) {
                float4* dst = reinterpret_cast<float4*>(dL_dsh + (size_t)first_row * 48);
                const int total = nrows * SGS_SH_ROW4;
#pragma unroll
                for (int it = 0; it < SGS_SH_ROW4; it++) {
                    const int e = it * 32 + lane;
                    if (e < total) {
                        const int row = e / SGS_SH_ROW4, c = e - row * SGS_SH_ROW4;
                        dst[e] = rows[row * SGS_SH_PAD4 + c];
                    }
                }
            }
        } else if (valid) {
            float* row = dL_dsh + (size_t)idx * M * 3;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k < M) {
                    row[3 * k] = o_sh[k].x; row[3 * k + 1] = o_sh[k].y; row[3 * k + 2] = o_sh[k].z;
                }
            }
            for (int k = 16; k < M; k++) {
                row[3 * k] = 0.f; row[3 * k + 1] = 0.f; row[3 * k + 2] = 0.f;
            }
        }
    }
}

void launch_preprocess_bwd(int P, const ViewParams& vp, const float* means3D, const int* radii, const float* shs,
                           const float* scales, const float* rotations, const float* cov3D, GeomState g,
                           float* acc, float* dL_dmean2D, float* dL_dopacity, float* dL_dcolor,
                           float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                           DensifySink sink, cudaStream_t s) {
    if (P <= 0) return;
    const int block = SGS_PRE_THREADS, grid = (P + block - 1) / block;
    const bool vec = (shs != nullptr) && vp.sh_coeffs == 16 && ((reinterpret_cast<size_t>(shs) & 15) == 0) &&
                     ((reinterpret_cast<size_t>(dL_dsh) & 15) == 0);
    // 128-bit accesses to the caller's rotations / dL_drot only when both pointers are 16-byte aligned
    const int rot_vec = (((reinterpret_cast<size_t>(rotations) | reinterpret_cast<size_t>(dL_drot)) & 15) == 0) ? 1 : 0;
    if (vec)
        launch_pdl(preprocess_bwd_kernel<true>, dim3(grid), dim3(block), 0, s, P, vp, means3D, radii, shs, scales, rotations,
                   cov3D, (const uint8_t*)g.clamped, (const float4*)g.conic_opacity, acc, dL_dmean2D, dL_dopacity, dL_dcolor,
                   dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, rot_vec, sink);
    else
        launch_pdl(preprocess_bwd_kernel<false>, dim3(grid), dim3(block), 0, s, P, vp, means3D, radii, shs, scales, rotations,
                   cov3D, (const uint8_t*)g.clamped, (const float4*)g.conic_opacity, acc, dL_dmean2D, dL_dopacity, dL_dcolor,
                   dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, rot_vec, sink);
}

}  // namespace sgs
