// features.cu — per-surfel feature preparation in front of the rasterizer, one kernel
// forward and one backward instead of ~20 eager P-sized torch kernels plus their autograd tape.
//
// Behavioural reference (what is computed, not how):
//   activations                       scene/gaussian_model.py:56-76, :236-266      exp / normalize / sigmoid
//   surfel normal                     scene/gaussian_model.py:48-54, :269-285      third column of R(q/|q|)
//   flip_align_view, safe_normalize   utils/general_utils.py:179-190
//   reflection, indirect SH, cat      gaussian_renderer/__init__.py:334-353        eval_sh(3, ., reflection), clamp_min 0
//   eval_sh                           utils/sh_utils.py:57-112
//
// Own design: one thread per surfel, 128 surfels per CTA. The 45 higher-order indirect-SH floats of a
// surfel have a 180-byte stride, so a CTA moves its 128x45 block between HBM and shared memory with
// consecutive (fully coalesced) accesses and every thread works on its own row (row stride 45 words: odd,
// conflict free). The backward reuses the tile: coefficients in, their gradients out, in place.
#include "kernels.cuh"

namespace mrgs {

namespace {

constexpr int kFeatThreads = 128;
constexpr int kRest = 45;  // 15 coefficients x rgb

constexpr float kC0 = 0.28209479177387814f;
constexpr float kC1 = 0.4886025119029199f;
__device__ constexpr float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                          0.5462742152960396f};
__device__ constexpr float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                          -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// degree-3 real SH basis at (x,y,z), indices 1..15 (index 0 is the constant kC0)
__device__ __forceinline__ void sh3_basis(float x, float y, float z, float* b) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[0] = kC0;
    b[1] = -kC1 * y;
    b[2] = kC1 * z;
    b[3] = -kC1 * x;
    b[4] = kC2[0] * xy;
    b[5] = kC2[1] * yz;
    b[6] = kC2[2] * (2.0f * zz - xx - yy);
    b[7] = kC2[3] * xz;
    b[8] = kC2[4] * (xx - yy);
    b[9] = kC3[0] * y * (3.0f * xx - yy);
    b[10] = kC3[1] * xy * z;
    b[11] = kC3[2] * y * (4.0f * zz - xx - yy);
    b[12] = kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = kC3[4] * x * (4.0f * zz - xx - yy);
    b[14] = kC3[5] * z * (xx - yy);
    b[15] = kC3[6] * x * (xx - 3.0f * yy);
}

// d basis / d(x,y,z)
__device__ __forceinline__ void sh3_basis_grad(float x, float y, float z, float* bx, float* by, float* bz) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    bx[0] = 0.f;                      by[0] = 0.f;                           bz[0] = 0.f;
    bx[1] = 0.f;                      by[1] = -kC1;                          bz[1] = 0.f;
    bx[2] = 0.f;                      by[2] = 0.f;                           bz[2] = kC1;
    bx[3] = -kC1;                     by[3] = 0.f;                           bz[3] = 0.f;
    bx[4] = kC2[0] * y;               by[4] = kC2[0] * x;                    bz[4] = 0.f;
    bx[5] = 0.f;                      by[5] = kC2[1] * z;                    bz[5] = kC2[1] * y;
    bx[6] = -2.0f * kC2[2] * x;       by[6] = -2.0f * kC2[2] * y;            bz[6] = 4.0f * kC2[2] * z;
    bx[7] = kC2[3] * z;               by[7] = 0.f;                           bz[7] = kC2[3] * x;
    bx[8] = 2.0f * kC2[4] * x;        by[8] = -2.0f * kC2[4] * y;            bz[8] = 0.f;
    bx[9] = kC3[0] * 6.0f * xy;       by[9] = kC3[0] * 3.0f * (xx - yy);     bz[9] = 0.f;
    bx[10] = kC3[1] * yz;             by[10] = kC3[1] * xz;                  bz[10] = kC3[1] * xy;
    bx[11] = -2.0f * kC3[2] * xy;     by[11] = kC3[2] * (4.0f * zz - xx - 3.0f * yy);  bz[11] = 8.0f * kC3[2] * yz;
    bx[12] = -6.0f * kC3[3] * xz;     by[12] = -6.0f * kC3[3] * yz;          bz[12] = kC3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
    bx[13] = kC3[4] * (4.0f * zz - 3.0f * xx - yy);  by[13] = -2.0f * kC3[4] * xy;     bz[13] = 8.0f * kC3[4] * xz;
    bx[14] = 2.0f * kC3[5] * xz;      by[14] = -2.0f * kC3[5] * yz;          bz[14] = kC3[5] * (xx - yy);
    bx[15] = 3.0f * kC3[6] * (xx - yy);              by[15] = -6.0f * kC3[6] * xy;     bz[15] = 0.f;
}

// geometry shared by forward and backward: view direction, flipped unit normal, reflection
struct FeatGeom {
    float dirn[3], inv_dlen;   // (xyz - campos) / |.|, 1/|.|
    float qh[4], inv_qlen;     // q / |q| (no epsilon, build_rotation), 1/|q|
    float sign;                // +1 / -1 from flip_align_view
    float n[3], inv_mlen;      // safe-normalised flipped normal, 1/max(|m|, 1e-20)
    float ndw;                 // n . w_o
    float refl[3];
};

__device__ __forceinline__ FeatGeom feat_geometry(const float* xyz, const float4 q, const float* cam) {
    FeatGeom g;
    const float dx = xyz[0] - cam[0], dy = xyz[1] - cam[1], dz = xyz[2] - cam[2];
    const float dlen = sqrtf(dx * dx + dy * dy + dz * dz);
    g.inv_dlen = 1.0f / dlen;
    g.dirn[0] = dx / dlen; g.dirn[1] = dy / dlen; g.dirn[2] = dz / dlen;
    const float qlen = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    g.inv_qlen = 1.0f / qlen;
    const float r = q.x / qlen, x = q.y / qlen, y = q.z / qlen, z = q.w / qlen;
    g.qh[0] = r; g.qh[1] = x; g.qh[2] = y; g.qh[3] = z;
    float m0 = 2.0f * (x * z + r * y), m1 = 2.0f * (y * z - r * x), m2 = 1.0f - 2.0f * (x * x + y * y);
    const float dot = -(m0 * g.dirn[0] + m1 * g.dirn[1] + m2 * g.dirn[2]);
    g.sign = dot >= 0.0f ? 1.0f : -1.0f;
    m0 *= g.sign; m1 *= g.sign; m2 *= g.sign;
    const float mlen = fmaxf(sqrtf(m0 * m0 + m1 * m1 + m2 * m2), 1e-20f);
    g.inv_mlen = 1.0f / mlen;
    g.n[0] = m0 / mlen; g.n[1] = m1 / mlen; g.n[2] = m2 / mlen;
    const float w0 = -g.dirn[0], w1 = -g.dirn[1], w2 = -g.dirn[2];
    g.ndw = g.n[0] * w0 + g.n[1] * w1 + g.n[2] * w2;
    g.refl[0] = 2.0f * g.ndw * g.n[0] - w0;
    g.refl[1] = 2.0f * g.ndw * g.n[1] - w1;
    g.refl[2] = 2.0f * g.ndw * g.n[2] - w2;
    return g;
}

__device__ __forceinline__ void load_rest_tile(float* s_rest, const float* __restrict__ rest, int base, int P) {
    const int rows = min(kFeatThreads, P - base);
    const float* src = rest + (size_t)base * kRest;
    for (int i = threadIdx.x; i < rows * kRest; i += kFeatThreads) s_rest[i] = src[i];
}

__global__ void __launch_bounds__(kFeatThreads) surfel_features_fwd_kernel(const MrgsSurfelFeatureArgs a) {
    __shared__ float s_rest[kFeatThreads * kRest];
    const int base = blockIdx.x * kFeatThreads;
    load_rest_tile(s_rest, a.indirect_rest, base, a.P);
    __syncthreads();
    const int i = base + threadIdx.x;
    if (i >= a.P) return;

    const float2 sc = reinterpret_cast<const float2*>(a.scaling)[i];
    reinterpret_cast<float2*>(a.scales)[i] = make_float2(expf(sc.x), expf(sc.y));
    const float4 q = reinterpret_cast<const float4*>(a.rotation)[i];
    {
        const float denom = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);  // F.normalize
        reinterpret_cast<float4*>(a.rotations)[i] = make_float4(q.x / denom, q.y / denom, q.z / denom, q.w / denom);
    }
    a.opacities[i] = sigmoidf(a.opacity[i]);

    const float xyz[3] = {a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2]};
    const float cam[3] = {a.campos[0], a.campos[1], a.campos[2]};
    const FeatGeom g = feat_geometry(xyz, q, cam);
    float b[16];
    sh3_basis(g.refl[0], g.refl[1], g.refl[2], b);
    const float* row = s_rest + threadIdx.x * kRest;
    float ind[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // the reference's left-to-right sum (sh_utils.py:79-104)
        float r = b[0] * a.indirect_dc[3 * i + c];
#pragma unroll
        for (int k = 1; k < 16; ++k) r += b[k] * row[(k - 1) * 3 + c];
        ind[c] = fmaxf(r, 0.0f);
    }
    float4* f = reinterpret_cast<float4*>(a.features) + 2 * (size_t)i;
    f[0] = make_float4(sigmoidf(a.refl_strength[i]), sigmoidf(a.roughness[i]), sigmoidf(a.ori_color[3 * i]),
                       sigmoidf(a.ori_color[3 * i + 1]));
    f[1] = make_float4(sigmoidf(a.ori_color[3 * i + 2]), ind[0], ind[1], ind[2]);
}

__global__ void __launch_bounds__(kFeatThreads) surfel_features_bwd_kernel(const MrgsSurfelFeatureArgs a) {
    __shared__ float s_rest[kFeatThreads * kRest];
    const int base = blockIdx.x * kFeatThreads;
    load_rest_tile(s_rest, a.indirect_rest, base, a.P);
    __syncthreads();
    const int i = base + threadIdx.x;
    if (i < a.P) {
        // activations
        const float2 sc = reinterpret_cast<const float2*>(a.scaling)[i];
        const float2 gs = reinterpret_cast<const float2*>(a.dL_dscales)[i];
        reinterpret_cast<float2*>(a.dL_dscaling)[i] = make_float2(gs.x * expf(sc.x), gs.y * expf(sc.y));
        {
            const float o = sigmoidf(a.opacity[i]);
            a.dL_dopacity[i] = a.dL_dopacities[i] * o * (1.0f - o);
        }
        const float4* gf4 = reinterpret_cast<const float4*>(a.dL_dfeatures) + 2 * (size_t)i;
        const float4 gf0 = gf4[0], gf1 = gf4[1];
        {
            const float s0 = sigmoidf(a.refl_strength[i]), s1 = sigmoidf(a.roughness[i]);
            a.dL_drefl_strength[i] = gf0.x * s0 * (1.0f - s0);
            a.dL_droughness[i] = gf0.y * s1 * (1.0f - s1);
            const float c0 = sigmoidf(a.ori_color[3 * i]), c1 = sigmoidf(a.ori_color[3 * i + 1]),
                        c2 = sigmoidf(a.ori_color[3 * i + 2]);
            a.dL_dori_color[3 * i] = gf0.z * c0 * (1.0f - c0);
            a.dL_dori_color[3 * i + 1] = gf0.w * c1 * (1.0f - c1);
            a.dL_dori_color[3 * i + 2] = gf1.x * c2 * (1.0f - c2);
        }

        const float4 q = reinterpret_cast<const float4*>(a.rotation)[i];
        const float xyz[3] = {a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2]};
        const float cam[3] = {a.campos[0], a.campos[1], a.campos[2]};
        const FeatGeom g = feat_geometry(xyz, q, cam);

        // indirect = clamp_min(sum_k b_k(refl) sh_k, 0): gradients to the coefficients and to refl
        float b[16], bx[16], by[16], bz[16];
        sh3_basis(g.refl[0], g.refl[1], g.refl[2], b);
        sh3_basis_grad(g.refl[0], g.refl[1], g.refl[2], bx, by, bz);
        float* row = s_rest + threadIdx.x * kRest;
        const float gin[3] = {gf1.y, gf1.z, gf1.w};
        float dr[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float dc = a.indirect_dc[3 * i + c];
            float r = b[0] * dc;
#pragma unroll
            for (int k = 1; k < 16; ++k) r += b[k] * row[(k - 1) * 3 + c];
            const float gc = r >= 0.0f ? gin[c] : 0.0f;   // clamp_min passes the gradient where x >= 0
            a.dL_dindirect_dc[3 * i + c] = gc * b[0];
#pragma unroll
            for (int k = 1; k < 16; ++k) {
                const float coef = row[(k - 1) * 3 + c];
                dr[0] += gc * coef * bx[k];
                dr[1] += gc * coef * by[k];
                dr[2] += gc * coef * bz[k];
                row[(k - 1) * 3 + c] = gc * b[k];          // in place: the tile now carries dL/d(rest)
            }
        }
        // refl = 2 (n.w) n - w,  w = -dirn
        const float w[3] = {-g.dirn[0], -g.dirn[1], -g.dirn[2]};
        const float drn = dr[0] * g.n[0] + dr[1] * g.n[1] + dr[2] * g.n[2];
        float dn[3], dw[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dn[c] = 2.0f * (drn * w[c] + g.ndw * dr[c]);
            dw[c] = 2.0f * drn * g.n[c] - dr[c];
        }
        // n = m / max(|m|, eps); m = sign * raw
        const float ndn = g.n[0] * dn[0] + g.n[1] * dn[1] + g.n[2] * dn[2];
        float draw[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) draw[c] = g.sign * (dn[c] - g.n[c] * ndn) * g.inv_mlen;
        // raw = (2(xz + ry), 2(yz - rx), 1 - 2(xx + yy)) of the unit quaternion (r,x,y,z)
        const float r_ = g.qh[0], x_ = g.qh[1], y_ = g.qh[2], z_ = g.qh[3];
        float dqh[4];
        dqh[0] = 2.0f * (y_ * draw[0] - x_ * draw[1]);
        dqh[1] = 2.0f * (z_ * draw[0] - r_ * draw[1]) - 4.0f * x_ * draw[2];
        dqh[2] = 2.0f * (r_ * draw[0] + z_ * draw[1]) - 4.0f * y_ * draw[2];
        dqh[3] = 2.0f * (x_ * draw[0] + y_ * draw[1]);
        const float qdot = g.qh[0] * dqh[0] + g.qh[1] * dqh[1] + g.qh[2] * dqh[2] + g.qh[3] * dqh[3];
        float dq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) dq[c] = (dqh[c] - g.qh[c] * qdot) * g.inv_qlen;
        // + the rasterizer's gradient through rotations = F.normalize(q)
        {
            const float4 gr = reinterpret_cast<const float4*>(a.dL_drotations)[i];
            const float qlen = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
            if (qlen > 1e-12f) {
                const float inv = 1.0f / qlen;
                const float h[4] = {q.x * inv, q.y * inv, q.z * inv, q.w * inv};
                const float gg[4] = {gr.x, gr.y, gr.z, gr.w};
                const float hd = h[0] * gg[0] + h[1] * gg[1] + h[2] * gg[2] + h[3] * gg[3];
#pragma unroll
                for (int c = 0; c < 4; ++c) dq[c] += (gg[c] - h[c] * hd) * inv;
            } else {
                dq[0] += gr.x * 1e12f; dq[1] += gr.y * 1e12f; dq[2] += gr.z * 1e12f; dq[3] += gr.w * 1e12f;
            }
        }
        reinterpret_cast<float4*>(a.dL_drotation)[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        // dirn = d / |d| with w = -dirn (the flip's sign has no gradient)
        const float dd[3] = {-dw[0], -dw[1], -dw[2]};
        const float ddn = g.dirn[0] * dd[0] + g.dirn[1] * dd[1] + g.dirn[2] * dd[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) a.dL_dxyz[3 * i + c] = (dd[c] - g.dirn[c] * ddn) * g.inv_dlen;
    }
    __syncthreads();
    const int rows = min(kFeatThreads, a.P - base);
    float* dst = a.dL_dindirect_rest + (size_t)base * kRest;
    for (int k = threadIdx.x; k < rows * kRest; k += kFeatThreads) dst[k] = s_rest[k];
}

}  // namespace

int launch_surfel_features(const MrgsSurfelFeatureArgs* a, bool backward, cudaStream_t stream) {
    if (a == nullptr || a->P < 0) {
        set_error("surfel_features: bad arguments");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->P == 0) return MRGS_OK;
    const bool in_ok = a->campos && a->xyz && a->scaling && a->rotation && a->opacity && a->refl_strength &&
                       a->roughness && a->ori_color && a->indirect_dc && a->indirect_rest;
    if (!in_ok) {
        set_error("surfel_features: null input");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const int blocks = (a->P + kFeatThreads - 1) / kFeatThreads;
    if (!backward) {
        if (!(a->scales && a->rotations && a->opacities && a->features)) {
            set_error("surfel_features_forward: null output");
            return MRGS_ERR_INVALID_ARGUMENT;
        }
        surfel_features_fwd_kernel<<<blocks, kFeatThreads, 0, stream>>>(*a);
    } else {
        const bool ok = a->dL_dscales && a->dL_drotations && a->dL_dopacities && a->dL_dfeatures && a->dL_dxyz &&
                        a->dL_dscaling && a->dL_drotation && a->dL_dopacity && a->dL_drefl_strength &&
                        a->dL_droughness && a->dL_dori_color && a->dL_dindirect_dc && a->dL_dindirect_rest;
        if (!ok) {
            set_error("surfel_features_backward: null gradient pointer");
            return MRGS_ERR_INVALID_ARGUMENT;
        }
        surfel_features_bwd_kernel<<<blocks, kFeatThreads, 0, stream>>>(*a);
    }
    return MRGS_OK;
}

}  // namespace mrgs
