// cube_sample.cuh — cube-map addressing and seamless bilinear fetch shared by the shading and the
// cubemap-prefilter kernels. Semantics restate nvdiffrast's dr.texture(boundary_mode='cube') and the
// reference's cube_to_dir table (scene/light_utils.py:24-31, scene/renderutils/c_src/cubemap.cu:32-46).
#pragma once
#include "common.cuh"

namespace mrgs {

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// direction -> (face, u, v) with u,v in [0,1]; also returns the projection Jacobian pieces
struct FaceUV {
    int face;
    float u, v;
    // u = su * a * m + .5, v = sv * b * m + .5 with m = 0.5/|c|; (ia, ib, ic) index x/y/z
    int ia, ib, ic;
    float su, sv, m, csign;
};

__device__ __forceinline__ FaceUV dir_to_face(F3 d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    FaceUV r;
    float a, b, c;
    if (az > fmaxf(ax, ay)) {
        r.face = 4; c = d.z; a = d.x; b = d.y; r.ia = 0; r.ib = 1; r.ic = 2;
    } else if (ay > ax) {
        r.face = 2; c = d.y; a = d.x; b = d.z; r.ia = 0; r.ib = 2; r.ic = 1;
    } else {
        r.face = 0; c = d.x; a = d.z; b = d.y; r.ia = 2; r.ib = 1; r.ic = 0;
    }
    if (c < 0.f) r.face += 1;
    r.csign = c < 0.f ? -1.f : 1.f;
    r.m = 0.5f / fabsf(c);
    // sign table of the reference's cube_to_dir (scene/light_utils.py:24-31)
    r.su = (r.face == 0 || r.face == 5) ? -1.f : 1.f;
    r.sv = (r.face == 2) ? 1.f : -1.f;
    r.u = fminf(fmaxf(r.su * a * r.m + 0.5f, 0.f), 1.f);
    r.v = fminf(fmaxf(r.sv * b * r.m + 0.5f, 0.f), 1.f);
    return r;
}

__device__ __forceinline__ F3 face_to_dir(int face, float fx, float fy) {
    switch (face) {
        case 0: return {1.f, -fy, -fx};
        case 1: return {-1.f, -fy, fx};
        case 2: return {fx, 1.f, fy};
        case 3: return {fx, -1.f, -fy};
        case 4: return {fx, -fy, 1.f};
        default: return {-fx, -fy, -1.f};
    }
}

// texel (ix,iy) of `face`, possibly one step outside it -> linear texel index, or -1 at a corner
__device__ __forceinline__ int resolve_texel(int face, int ix, int iy, int res) {
    const bool ox = ix < 0 || ix >= res, oy = iy < 0 || iy >= res;
    if (ox && oy) return -1;
    if (ox || oy) {
        const float fx = 2.f * ((float)ix + 0.5f) / (float)res - 1.f;
        const float fy = 2.f * ((float)iy + 0.5f) / (float)res - 1.f;
        const FaceUV n = dir_to_face(face_to_dir(face, fx, fy));
        face = n.face;
        ix = min(res - 1, (int)(n.u * (float)res));
        iy = min(res - 1, (int)(n.v * (float)res));
    }
    return (face * res + iy) * res + ix;
}

struct Bilinear {
    int idx[4];
    float w[4];           // effective per-texel weights (corner share already redistributed)
    F3 val;
    F3 dval_du, dval_dv;  // derivative w.r.t. the face coordinates u,v in [0,1]
};

__device__ __forceinline__ F3 load3(const float* __restrict__ tex, int idx) {
    const float* p = tex + 3 * (size_t)idx;
    return {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}

template <bool GRAD>
__device__ __forceinline__ void cube_bilinear(const float* __restrict__ tex, int res, int face, float u,
                                              float v, Bilinear& o) {
    const float U = u * (float)res - 0.5f, V = v * (float)res - 0.5f;
    const float fU = floorf(U), fV = floorf(V);
    const int iu0 = (int)fU, iv0 = (int)fV;
    const float fu = U - fU, fv = V - fV;
    o.idx[0] = resolve_texel(face, iu0, iv0, res);
    o.idx[1] = resolve_texel(face, iu0 + 1, iv0, res);
    o.idx[2] = resolve_texel(face, iu0, iv0 + 1, res);
    o.idx[3] = resolve_texel(face, iu0 + 1, iv0 + 1, res);
    F3 a[4];
    F3 sum = {0.f, 0.f, 0.f};
    int missing = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (o.idx[k] >= 0) {
            a[k] = load3(tex, o.idx[k]);
            sum = sum + a[k];
        } else {
            a[k] = {0.f, 0.f, 0.f};
            missing = k;
        }
    }
    const float w0[4] = {(1.f - fu) * (1.f - fv), fu * (1.f - fv), (1.f - fu) * fv, fu * fv};
#pragma unroll
    for (int k = 0; k < 4; ++k) o.w[k] = w0[k];
    if (missing >= 0) {
        const F3 avg = 0.33333333f * sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k == missing) {
                a[k] = avg;
                o.w[k] = 0.f;
            } else {
                o.w[k] = w0[k] + 0.33333333f * w0[missing];
            }
        }
    }
    o.val = w0[0] * a[0] + w0[1] * a[1] + w0[2] * a[2] + w0[3] * a[3];
    if (GRAD) {
        o.dval_du = (float)res * ((1.f - fv) * (a[1] - a[0]) + fv * (a[3] - a[2]));
        o.dval_dv = (float)res * ((1.f - fu) * (a[2] - a[0]) + fu * (a[3] - a[1]));
    }
}

}  // namespace mrgs
