// preprocess_bwd.cu — per-surfel backward: raw gradients of T / mean2D / normal / colour from
// the arena are chained to scale, rotation, position and SH coefficients; every output row is
// written here (zeros for culled surfels), so no output needs a prior memset.
//
// Behavioural reference: preprocessCUDA + compute_transmat_aabb + computeColorFromSH of
// rast/cuda_rasterizer/backward.cu:614-669, :471-612, :22-141 and quat_to_rotmat_vjp
// auxiliary.h:245-289. Contract quirks kept on purpose: the transform is
// rebuilt with scale_modifier = 1 and with W,H = int(focal*tan*2); the quaternion vjp does not
// differentiate the normalisation; dL_dmeans2D is overwritten with the densification proxy.
#include "kernels.cuh"
#include "splat_math.cuh"

namespace mrgs {

namespace {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Gradient of the SH colour w.r.t. the coefficients (written to dsh, M triples) and w.r.t. the
// surfel position (returned). g = dL/dRGB with clamped channels already zeroed.
struct Sh3View {  // coefficient triples over a flat float array (constant indices stay in registers)
    const float* p;
    __device__ __forceinline__ V3 operator[](int k) const { return {p[3 * k], p[3 * k + 1], p[3 * k + 2]}; }
};
struct Sh3Out {
    float* p;
    struct Ref {
        float* q;
        __device__ __forceinline__ void operator=(V3 v) { q[0] = v.x; q[1] = v.y; q[2] = v.z; }
    };
    __device__ __forceinline__ Ref operator[](int k) const { return {p + 3 * k}; }
};

template <bool kZeroTail>
__device__ __forceinline__ V3 sh_backward(int deg, int M, const float* __restrict__ sh_raw, V3 pos, V3 cam, V3 g,
                                          float* __restrict__ dsh) {
    const Sh3View sh{sh_raw};
    const Sh3Out out{dsh};
    const V3 d0 = {pos.x - cam.x, pos.y - cam.y, pos.z - cam.z};
    const float inv_len = 1.0f / sqrtf(dot(d0, d0));
    const float x = d0.x * inv_len, y = d0.y * inv_len, z = d0.z * inv_len;

    V3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
    out[0] = SH_C0 * g;
    int written = 1;
    if (deg > 0) {
        out[1] = (-SH_C1 * y) * g;
        out[2] = (SH_C1 * z) * g;
        out[3] = (-SH_C1 * x) * g;
        written = 4;
        dx = -SH_C1 * sh[3];
        dy = -SH_C1 * sh[1];
        dz = SH_C1 * sh[2];
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            out[4] = (SH_C2[0] * xy) * g;
            out[5] = (SH_C2[1] * yz) * g;
            out[6] = (SH_C2[2] * (2.f * zz - xx - yy)) * g;
            out[7] = (SH_C2[3] * xz) * g;
            out[8] = (SH_C2[4] * (xx - yy)) * g;
            written = 9;
            dx = dx + (SH_C2[0] * y) * sh[4] + (SH_C2[2] * 2.f * -x) * sh[6] + (SH_C2[3] * z) * sh[7] +
                 (SH_C2[4] * 2.f * x) * sh[8];
            dy = dy + (SH_C2[0] * x) * sh[4] + (SH_C2[1] * z) * sh[5] + (SH_C2[2] * 2.f * -y) * sh[6] +
                 (SH_C2[4] * 2.f * -y) * sh[8];
            dz = dz + (SH_C2[1] * y) * sh[5] + (SH_C2[2] * 4.f * z) * sh[6] + (SH_C2[3] * x) * sh[7];
            if (deg > 2) {
                out[9] = (SH_C3[0] * y * (3.f * xx - yy)) * g;
                out[10] = (SH_C3[1] * xy * z) * g;
                out[11] = (SH_C3[2] * y * (4.f * zz - xx - yy)) * g;
                out[12] = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * g;
                out[13] = (SH_C3[4] * x * (4.f * zz - xx - yy)) * g;
                out[14] = (SH_C3[5] * z * (xx - yy)) * g;
                out[15] = (SH_C3[6] * x * (xx - 3.f * yy)) * g;
                written = 16;
                dx = dx + (SH_C3[0] * 6.f * xy) * sh[9] + (SH_C3[1] * yz) * sh[10] +
                     (SH_C3[2] * -2.f * xy) * sh[11] + (SH_C3[3] * -6.f * xz) * sh[12] +
                     (SH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * sh[13] + (SH_C3[5] * 2.f * xz) * sh[14] +
                     (SH_C3[6] * 3.f * (xx - yy)) * sh[15];
                dy = dy + (SH_C3[0] * 3.f * (xx - yy)) * sh[9] + (SH_C3[1] * xz) * sh[10] +
                     (SH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * sh[11] + (SH_C3[3] * -6.f * yz) * sh[12] +
                     (SH_C3[4] * -2.f * xy) * sh[13] + (SH_C3[5] * -2.f * yz) * sh[14] +
                     (SH_C3[6] * -6.f * xy) * sh[15];
                dz = dz + (SH_C3[1] * xy) * sh[10] + (SH_C3[2] * 8.f * yz) * sh[11] +
                     (SH_C3[3] * 3.f * (2.f * zz - xx - yy)) * sh[12] + (SH_C3[4] * 8.f * xz) * sh[13] +
                     (SH_C3[5] * (xx - yy)) * sh[14];
            }
        }
    }
    if (kZeroTail)
        for (int k = written; k < M; ++k) out[k] = V3{0.f, 0.f, 0.f};

    // through dir = d0 / |d0|
    const V3 ddir = {dot(dx, g), dot(dy, g), dot(dz, g)};
    const float s2 = dot(d0, d0);
    const float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
    V3 r;
    r.x = ((s2 - d0.x * d0.x) * ddir.x - d0.y * d0.x * ddir.y - d0.z * d0.x * ddir.z) * inv32;
    r.y = (-d0.x * d0.y * ddir.x + (s2 - d0.y * d0.y) * ddir.y - d0.z * d0.y * ddir.z) * inv32;
    r.z = (-d0.x * d0.z * ddir.x - d0.y * d0.z * ddir.y + (s2 - d0.z * d0.z) * ddir.z) * inv32;
    return r;
}

// kAcc: ADD the parameter gradients to what the output buffers hold (gradient accumulation over the views of a
// batch fused into this kernel: no separate add pass, no freshly written temporaries) and leave the rows of
// culled surfels untouched. dL_dmeans2D, the per-view densification proxy, is always overwritten.
template <bool kAcc> __device__ __forceinline__ void put(float* dst, float v) { *dst = kAcc ? (*dst + v) : v; }
template <bool kAcc> __device__ __forceinline__ void put(float2* dst, float2 v) {
    if (kAcc) { const float2 o = *dst; v.x += o.x; v.y += o.y; }
    *dst = v;
}
template <bool kAcc> __device__ __forceinline__ void put(float4* dst, float4 v) {
    if (kAcc) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    *dst = v;
}

template <bool kAcc, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) preprocess_bwd_kernel(const PreprocessBwdParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    const bool visible = p.radii[idx] > 0;
    if (kAcc && !visible) {
        p.dL_dmeans2D[3 * idx + 0] = 0.f;
        p.dL_dmeans2D[3 * idx + 1] = 0.f;
        p.dL_dmeans2D[3 * idx + 2] = 0.f;
        return;
    }
    const float* g = p.grad_arena + (size_t)idx * p.grad_stride;

    float dT[9], dm2x = 0.f, dm2y = 0.f, dopa = 0.f, dn[3] = {0.f, 0.f, 0.f}, dcol[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 9; ++i) dT[i] = 0.f;
    if (visible) {
        // rows are 16-byte aligned: read the 18 geometry/colour values as float4s
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const float4 a = g4[0], b = g4[1], c = g4[2], d = g4[3], e = g4[4];
        dT[0] = a.x; dT[1] = a.y; dT[2] = a.z; dT[3] = a.w;
        dT[4] = b.x; dT[5] = b.y; dT[6] = b.z; dT[7] = b.w;
        dT[8] = c.x; dm2x = c.y; dm2y = c.z; dopa = c.w;
        dn[0] = d.x; dn[1] = d.y; dn[2] = d.z; dcol[0] = d.w;
        dcol[1] = e.x; dcol[2] = e.y;
    }

    // pass-through outputs
    put<kAcc>(p.dL_dopacity + idx, dopa);
    if (p.dL_dcolors != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) put<kAcc>(p.dL_dcolors + 3 * idx + c, dcol[c]);
    }
    if ((p.S & 3) == 0) {
        // arena features start at float 18: not 16-byte aligned, so gather scalars, store vectors
        for (int c = 0; c < p.S; c += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (visible) v = make_float4(g[kGradFeature + c], g[kGradFeature + c + 1], g[kGradFeature + c + 2], g[kGradFeature + c + 3]);
            put<kAcc>(reinterpret_cast<float4*>(p.dL_dfeatures + (size_t)idx * p.S) + (c >> 2), v);
        }
    } else {
        for (int c = 0; c < p.S; ++c)
            put<kAcc>(p.dL_dfeatures + (size_t)idx * p.S + c, visible ? g[kGradFeature + c] : 0.f);
    }

    float dmean3[3] = {0.f, 0.f, 0.f}, dscale[2] = {0.f, 0.f}, drot[4] = {0.f, 0.f, 0.f, 0.f};
    float proxy_x = 0.f, proxy_y = 0.f;

    if (visible) {
        const bool precomp = (p.scales == nullptr);
        float T[9];
        float P3[3][4];  // world2ndc * ndc2pix, three columns of 4
        float Rm[3][3];  // rotation columns
        float sx = 0.f, sy = 0.f, q[4] = {0.f, 0.f, 0.f, 0.f};
        Vec3 pos = {0.f, 0.f, 0.f};
        Vec3 normal = {0.f, 0.f, 0.f};
        if (precomp) {
#pragma unroll
            for (int i = 0; i < 9; ++i) T[i] = p.transMat_precomp[9 * idx + i];
        } else {
            pos = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
            sx = p.scales[2 * idx];
            sy = p.scales[2 * idx + 1];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = p.rotations[4 * idx + i];
            const Rot3 R = quat_to_rot(q[0], q[1], q[2], q[3]);
            Rm[0][0] = R.c0.x; Rm[0][1] = R.c0.y; Rm[0][2] = R.c0.z;
            Rm[1][0] = R.c1.x; Rm[1][1] = R.c1.y; Rm[1][2] = R.c1.z;
            Rm[2][0] = R.c2.x; Rm[2][1] = R.c2.y; Rm[2][2] = R.c2.z;
            const float w2 = 0.5f * (float)p.W, wm = 0.5f * (float)(p.W - 1);
            const float h2 = 0.5f * (float)p.H, hm = 0.5f * (float)(p.H - 1);
            const float* pm = p.projmatrix;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // world2ndc[j][k] = pm[4k+j]
                P3[0][k] = pm[4 * k + 0] * w2 + pm[4 * k + 3] * wm;
                P3[1][k] = pm[4 * k + 1] * h2 + pm[4 * k + 3] * hm;
                P3[2][k] = pm[4 * k + 3];
            }
            // M columns: (sx*R0, 0), (sy*R1, 0), (pos, 1);  T[c][r] = sum_k M[r][k] * P3[c][k]
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T[3 * c + 0] = sx * (Rm[0][0] * P3[c][0] + Rm[0][1] * P3[c][1] + Rm[0][2] * P3[c][2]);
                T[3 * c + 1] = sy * (Rm[1][0] * P3[c][0] + Rm[1][1] * P3[c][1] + Rm[1][2] * P3[c][2]);
                T[3 * c + 2] = pos.x * P3[c][0] + pos.y * P3[c][1] + pos.z * P3[c][2] + P3[c][3];
            }
            normal = view_vector(p.viewmatrix, R.c2);
        }

        // low-pass-filter branch: dL/dmean2D -> dL/dT through the (cutoff-free) centre formula
        if (dm2x != 0.f || dm2y != 0.f) {
            const float t0 = T[6], t1 = T[7], t2 = T[8];
            const float distance = t0 * t0 + t1 * t1 - t2 * t2;
            const float f = 1.0f / distance;
            const float c0 = f - 2.f * f * f * t0 * t0;
            const float c1 = f - 2.f * f * f * t1 * t1;
            const float c2 = f + 2.f * f * f * t2 * t2;
            dT[0] += dm2x * f * t0;
            dT[1] += dm2x * f * t1;
            dT[2] += dm2x * -f * t2;
            dT[3] += dm2y * f * t0;
            dT[4] += dm2y * f * t1;
            dT[5] += dm2y * -f * t2;
            dT[6] += dm2x * T[0] * c0 + dm2y * T[3] * c0;
            dT[7] += dm2x * T[1] * c1 + dm2y * T[4] * c1;
            dT[8] += dm2x * -T[2] * c2 + dm2y * -T[5] * c2;
        }

        const float depth = p.rec[(size_t)idx * kGeomFloats + 8];  // forward Tw.z
        if (!precomp) {
            // dL/dM[r][k] = sum_c dL/dT[c][r] * P3[c][k]
            float dM[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    dM[r][k] = dT[r] * P3[0][k] + dT[3 + r] * P3[1][k] + dT[6 + r] * P3[2][k];

            const float* vm = p.viewmatrix;
            float dtn[3] = {vm[0] * dn[0] + vm[1] * dn[1] + vm[2] * dn[2],
                            vm[4] * dn[0] + vm[5] * dn[1] + vm[6] * dn[2],
                            vm[8] * dn[0] + vm[9] * dn[1] + vm[10] * dn[2]};
            const Vec3 pv = view_point(vm, pos);
            const float dd = __fmaf_rn(pv.z, normal.z, __fmaf_rn(pv.x, normal.x, __fmul_rn(pv.y, normal.y)));
            const float flip = (dd < 0.0f) ? 1.0f : -1.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) dtn[k] *= flip;

            // dL/dR columns
            float dR[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dR[0][k] = dM[0][k] * sx;
                dR[1][k] = dM[1][k] * sy;
                dR[2][k] = dtn[k];
            }
            dscale[0] = dM[0][0] * Rm[0][0] + dM[0][1] * Rm[0][1] + dM[0][2] * Rm[0][2];
            dscale[1] = dM[1][0] * Rm[1][0] + dM[1][1] * Rm[1][1] + dM[1][2] * Rm[1][2];
            dmean3[0] = dM[2][0];
            dmean3[1] = dM[2][1];
            dmean3[2] = dM[2][2];

            // quaternion vjp on the normalised quaternion (normalisation itself not differentiated)
            const float s = rsqrtf(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
            const float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
            drot[0] = 2.f * (x * (dR[1][2] - dR[2][1]) + y * (dR[2][0] - dR[0][2]) + z * (dR[0][1] - dR[1][0]));
            drot[1] = 2.f * (-2.f * x * (dR[1][1] + dR[2][2]) + y * (dR[0][1] + dR[1][0]) +
                             z * (dR[0][2] + dR[2][0]) + w * (dR[1][2] - dR[2][1]));
            drot[2] = 2.f * (x * (dR[0][1] + dR[1][0]) - 2.f * y * (dR[0][0] + dR[2][2]) +
                             z * (dR[1][2] + dR[2][1]) + w * (dR[2][0] - dR[0][2]));
            drot[3] = 2.f * (x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1]) -
                             2.f * z * (dR[0][0] + dR[1][1]) + w * (dR[0][1] - dR[1][0]));

            // densification proxy uses the RAW dT (the mean2D term is not written back here)
            proxy_x = g[kGradT + 2] * depth * 0.5f * (float)p.W;
            proxy_y = g[kGradT + 5] * depth * 0.5f * (float)p.H;
        } else {
            proxy_x = dT[2] * depth * 0.5f * (float)p.W;
            proxy_y = dT[5] * depth * 0.5f * (float)p.H;
        }
    }

    // dL_dtransMat: raw accumulation for the scale/rotation path, augmented for precomputed T
    if (p.dL_dtransMat != nullptr) {
        const bool precomp = (p.scales == nullptr);
#pragma unroll
        for (int i = 0; i < 9; ++i)
            put<kAcc>(p.dL_dtransMat + 9 * idx + i, visible ? (precomp ? dT[i] : g[kGradT + i]) : 0.f);
    }

    // SH coefficients and their contribution to the position gradient
    if (p.shs != nullptr) {
        float* dsh = p.dL_dsh + (size_t)idx * p.M * 3;
        if (visible) {
            const uint8_t cl = p.clamped[idx];
            const V3 gc = {(cl & 1) ? 0.f : dcol[0], (cl & 2) ? 0.f : dcol[1], (cl & 4) ? 0.f : dcol[2]};
            const V3 pos = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
            const V3 cam = {p.campos[0], p.campos[1], p.campos[2]};
            V3 dpos;
            if (p.M == 16) {
                // 192-byte rows: move them as 12 float4 each way, keep the 2x48 values in registers
                float sh_l[48], dsh_l[48];
                const float4* src = reinterpret_cast<const float4*>(p.shs + (size_t)idx * 48);
#pragma unroll
                for (int q = 0; q < 12; ++q) {
                    const float4 v = __ldg(src + q);
                    sh_l[4 * q] = v.x; sh_l[4 * q + 1] = v.y; sh_l[4 * q + 2] = v.z; sh_l[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < 48; ++k) dsh_l[k] = 0.f;
                dpos = sh_backward<false>(p.D, 16, sh_l, pos, cam, gc, dsh_l);
                float4* dst = reinterpret_cast<float4*>(dsh);
#pragma unroll
                for (int q = 0; q < 12; ++q)
                    put<kAcc>(dst + q, make_float4(dsh_l[4 * q], dsh_l[4 * q + 1], dsh_l[4 * q + 2], dsh_l[4 * q + 3]));
            } else if (kAcc) {
                // generic coefficient count (M <= 16, checked by the caller): stage the row, then add it
                float dsh_l[48];
                for (int k = 0; k < 48; ++k) dsh_l[k] = 0.f;
                dpos = sh_backward<false>(p.D, p.M, p.shs + (size_t)idx * p.M * 3, pos, cam, gc, dsh_l);
                for (int k = 0; k < p.M * 3; ++k) dsh[k] += dsh_l[k];
            } else {
                dpos = sh_backward<true>(p.D, p.M, p.shs + (size_t)idx * p.M * 3, pos, cam, gc, dsh);
            }
            dmean3[0] += dpos.x;
            dmean3[1] += dpos.y;
            dmean3[2] += dpos.z;
        } else if (p.M == 16) {
            float4* dst = reinterpret_cast<float4*>(dsh);
#pragma unroll
            for (int q = 0; q < 12; ++q) dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int k = 0; k < p.M * 3; ++k) dsh[k] = 0.f;
        }
    }

#pragma unroll
    for (int k = 0; k < 3; ++k) put<kAcc>(p.dL_dmeans3D + 3 * idx + k, dmean3[k]);
    p.dL_dmeans2D[3 * idx + 0] = proxy_x;
    p.dL_dmeans2D[3 * idx + 1] = proxy_y;
    p.dL_dmeans2D[3 * idx + 2] = 0.f;
    if (p.dL_dscales != nullptr)
        put<kAcc>(reinterpret_cast<float2*>(p.dL_dscales) + idx, make_float2(dscale[0], dscale[1]));
    if (p.dL_drotations != nullptr)
        put<kAcc>(reinterpret_cast<float4*>(p.dL_drotations) + idx, make_float4(drot[0], drot[1], drot[2], drot[3]));
}

}  // namespace

namespace {
template <int BS, int MINB>
void launch_pre_bwd(const PreprocessBwdParams& p, cudaStream_t stream) {
    if (p.accumulate)
        preprocess_bwd_kernel<true, BS, MINB><<<(p.P + BS - 1) / BS, BS, 0, stream>>>(p);
    else
        preprocess_bwd_kernel<false, BS, MINB><<<(p.P + BS - 1) / BS, BS, 0, stream>>>(p);
}
}  // namespace

void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream) {
    // 128-thread CTAs, 4 per SM: same 128-register budget as 256x2 but finer-grained slot reuse (0.206 -> 0.196 ms
    // at C3); capping the registers lower spills the SH backward and is far slower (96 regs 0.33 ms, 64 regs 0.62 ms).
    launch_pre_bwd<128, 4>(p, stream);
}

}  // namespace mrgs
