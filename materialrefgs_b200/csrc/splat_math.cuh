// splat_math.cuh — the fp32 arithmetic of the 2DGS ray-splat path with every rounding made
// explicit (__fmaf_rn / __fmul_rn / __fadd_rn), so that tile keys, radii and per-pixel alpha
// sequences are bit-identical to the reference binary regardless of how nvcc would contract
// a natural-looking expression.
//
// The operation order of each function was read from the sm_100a SASS of the reference
// kernels (nvcc 12.9) and cross-checked against the source it came from:
//   view_point / view_vector      <- transformPoint4x3 / transformVec4x3  auxiliary.h:80-109
//   quat_to_rot                   <- quat_to_rotmat                       auxiliary.h:220-242
//   surfel_transmat               <- compute_transmat                     forward.cu:77-125
//   surfel_aabb                   <- compute_aabb                         forward.cu:129-159
//   tile_rect                     <- getRect                              auxiliary.h:68-78
//   ray_splat                     <- renderCUDA inner loop                forward.cu:371-399,
//                                                                         backward.cu:302-328
#pragma once
#include "common.cuh"

namespace mrgs {

struct Vec3 {
    float x, y, z;
};

// a*b + c*d + e*f as the reference binary evaluates a left-to-right sum of three products
// whose FIRST two products are (a*b) fused, (c*d) rounded:  fma(e,f, fma(a,b, rn(c*d)))
__device__ __forceinline__ float dot3_second_rounded(float a, float b, float c, float d, float e,
                                                     float f) {
    return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}

// p_view = p * viewmatrix (row-vector convention, m[col*4+row]) + translation
__device__ __forceinline__ Vec3 view_point(const float* __restrict__ m, Vec3 p) {
    Vec3 r;
    r.x = __fadd_rn(dot3_second_rounded(m[0], p.x, m[4], p.y, m[8], p.z), m[12]);
    r.y = __fadd_rn(dot3_second_rounded(m[1], p.x, m[5], p.y, m[9], p.z), m[13]);
    r.z = __fadd_rn(dot3_second_rounded(m[2], p.x, m[6], p.y, m[10], p.z), m[14]);
    return r;
}

__device__ __forceinline__ float view_depth(const float* __restrict__ m, Vec3 p) {
    return __fadd_rn(dot3_second_rounded(m[2], p.x, m[6], p.y, m[10], p.z), m[14]);
}

__device__ __forceinline__ Vec3 view_vector(const float* __restrict__ m, Vec3 v) {
    Vec3 r;
    r.x = dot3_second_rounded(m[0], v.x, m[4], v.y, m[8], v.z);
    r.y = dot3_second_rounded(m[1], v.x, m[5], v.y, m[9], v.z);
    r.z = dot3_second_rounded(m[2], v.x, m[6], v.y, m[10], v.z);
    return r;
}

// Rotation matrix columns c0,c1,c2 from a (w,x,y,z)-ordered quaternion q[0..3].
struct Rot3 {
    Vec3 c0, c1, c2;
};

__device__ __forceinline__ Rot3 quat_to_rot(float q0, float q1, float q2, float q3) {
    // |q|^2 is summed as fma(q2,q2, fma(q1,q1, fma(q0,q0, rn(q3*q3))))
    const float n2 =
        __fmaf_rn(q2, q2, __fmaf_rn(q1, q1, __fmaf_rn(q0, q0, __fmul_rn(q3, q3))));
    const float s = rsqrtf(n2);
    const float w = __fmul_rn(q0, s);
    const float x = __fmul_rn(q1, s);
    const float y = __fmul_rn(q2, s);
    const float z = __fmul_rn(q3, s);
    const float wz = __fmul_rn(w, z), wx = __fmul_rn(w, x), wy = __fmul_rn(w, y);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float h01 = __fmaf_rn(x, y, wz);   // x*y + w*z
    const float h10 = __fmaf_rn(x, y, -wz);  // x*y - w*z
    const float h12 = __fmaf_rn(y, z, wx);   // y*z + w*x
    const float h21 = __fmaf_rn(y, z, -wx);  // y*z - w*x
    const float h02 = __fmaf_rn(x, z, -wy);  // x*z - w*y
    const float h20 = __fmaf_rn(x, z, wy);   // x*z + w*y
    const float d00 = __fadd_rn(yy, zz);
    const float d11 = __fmaf_rn(x, x, zz);
    const float d22 = __fmaf_rn(x, x, yy);
    Rot3 R;
    R.c0 = {__fadd_rn(1.0f, -__fadd_rn(d00, d00)), __fadd_rn(h01, h01), __fadd_rn(h02, h02)};
    R.c1 = {__fadd_rn(h10, h10), __fadd_rn(1.0f, -__fadd_rn(d11, d11)), __fadd_rn(h12, h12)};
    R.c2 = {__fadd_rn(h20, h20), __fadd_rn(h21, h21), __fadd_rn(1.0f, -__fadd_rn(d22, d22))};
    return R;
}

// T (3 rows Tu,Tv,Tw stored row-wise in T[9]) = (splat2world)^T * world2ndc * ndc2pix, and the
// view-space normal (before the DUAL_VISIABLE flip).
__device__ __forceinline__ void surfel_transmat(Vec3 p, float sx, float sy, float mod, float q0,
                                                float q1, float q2, float q3,
                                                const float* __restrict__ pm,
                                                const float* __restrict__ vm, int W, int H,
                                                float* T, Vec3& normal) {
    const Rot3 R = quat_to_rot(q0, q1, q2, q3);
    const float msx = __fmul_rn(sx, mod);
    const float msy = __fmul_rn(sy, mod);
    const Vec3 L0 = {__fmul_rn(msx, R.c0.x), __fmul_rn(msx, R.c0.y), __fmul_rn(msx, R.c0.z)};
    const Vec3 L1 = {__fmul_rn(msy, R.c1.x), __fmul_rn(msy, R.c1.y), __fmul_rn(msy, R.c1.z)};

    // A[c][r]: only columns c = 0,1,3 of (splat2world^T * world2ndc) survive ndc2pix
    float A0[3], A1[3], A3[3];
    A0[0] = dot3_second_rounded(L0.x, pm[0], L0.y, pm[4], L0.z, pm[8]);
    A0[1] = dot3_second_rounded(L1.x, pm[0], L1.y, pm[4], L1.z, pm[8]);
    A0[2] = __fadd_rn(dot3_second_rounded(pm[0], p.x, pm[4], p.y, pm[8], p.z), pm[12]);
    A1[0] = dot3_second_rounded(L0.x, pm[1], L0.y, pm[5], L0.z, pm[9]);
    A1[1] = dot3_second_rounded(L1.x, pm[1], L1.y, pm[5], L1.z, pm[9]);
    A1[2] = __fadd_rn(dot3_second_rounded(pm[1], p.x, pm[5], p.y, pm[9], p.z), pm[13]);
    A3[0] = dot3_second_rounded(L0.x, pm[3], L0.y, pm[7], L0.z, pm[11]);
    A3[1] = dot3_second_rounded(L1.x, pm[3], L1.y, pm[7], L1.z, pm[11]);
    A3[2] = __fadd_rn(dot3_second_rounded(pm[3], p.x, pm[7], p.y, pm[11], p.z), pm[15]);

    const float w2 = __fmul_rn((float)W, 0.5f), wm = __fmul_rn((float)(W - 1), 0.5f);
    const float h2 = __fmul_rn((float)H, 0.5f), hm = __fmul_rn((float)(H - 1), 0.5f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T[0 + r] = __fmaf_rn(wm, A3[r], __fmul_rn(w2, A0[r]));
        T[3 + r] = __fmaf_rn(hm, A3[r], __fmul_rn(h2, A1[r]));
        T[6 + r] = A3[r];
    }
    normal = view_vector(vm, R.c2);
}

// Centre and 3-sigma half extent of the projected surfel. Returns false for a degenerate T.
__device__ __forceinline__ bool surfel_aabb(const float* T, float& cx, float& cy, float& ex,
                                            float& ey) {
    const float t20 = T[6], t21 = T[7], t22 = T[8];
    const float d = __fmaf_rn(-t22, t22,
                              __fmaf_rn(__fmul_rn(t20, t20), 9.0f,
                                        __fmul_rn(__fmul_rn(t21, t21), 9.0f)));
    if (d == 0.0f) return false;
    const float rc = __fdiv_rn(1.0f, d);
    const float f9 = __fmul_rn(rc, 9.0f);
    const float a0 = __fmul_rn(f9, T[0]), a1 = __fmul_rn(f9, T[1]), a2 = __fmul_rn(rc, -T[2]);
    const float b0 = __fmul_rn(f9, T[3]), b1 = __fmul_rn(f9, T[4]), b2 = __fmul_rn(rc, -T[5]);
    // sums of three products here round the FIRST product and fuse the other two
    cx = __fmaf_rn(a2, t22, __fmaf_rn(a1, t21, __fmul_rn(a0, t20)));
    const float tx = __fmaf_rn(a2, T[2], __fmaf_rn(a1, T[1], __fmul_rn(a0, T[0])));
    cy = __fmaf_rn(b2, t22, __fmaf_rn(b1, t21, __fmul_rn(b0, t20)));
    const float ty = __fmaf_rn(b2, T[5], __fmaf_rn(b1, T[4], __fmul_rn(b0, T[3])));
    const float hx = __fmaf_rn(cx, cx, -tx);
    const float hy = __fmaf_rn(cy, cy, -ty);
    ex = sqrtf(fmaxf(hx, 1e-4f));
    ey = sqrtf(fmaxf(hy, 1e-4f));
    return true;
}

__device__ __forceinline__ void tile_rect(float cx, float cy, int radius, int grid_x, int grid_y,
                                          int& x0, int& y0, int& x1, int& y1) {
    const float rf = (float)radius;
    x0 = min(grid_x, max(0, (int)__fmul_rn(__fadd_rn(cx, -rf), 0.0625f)));
    y0 = min(grid_y, max(0, (int)__fmul_rn(__fadd_rn(cy, -rf), 0.0625f)));
    x1 = min(grid_x, max(0, (int)__fmul_rn(
                                __fadd_rn(__fadd_rn(__fadd_rn(cx, rf), 16.0f), -1.0f), 0.0625f)));
    y1 = min(grid_y, max(0, (int)__fmul_rn(
                                __fadd_rn(__fadd_rn(__fadd_rn(cy, rf), 16.0f), -1.0f), 0.0625f)));
}

// Conservative support of one surfel's alpha. A pair is blended only if
//   alpha = min(0.99, opacity * exp(-min(rho3d, rho2d)/2)) >= 1/255  <=>  min(rho3d, rho2d) <= 2 ln(255 opacity) =: tau_exact.
// tau = tau_exact * 1.001 + 0.001 leaves a margin ~1000x larger than the rounding of the fp32 path
// (divisions, expf <= 2 ulp, logf), so "rho3d > tau and rho2d > tau" PROVES the exact evaluation would
// have skipped the pair; results stay bit-identical (tests/test_raster_gpu.py compares n_contrib and
// final_T bit-for-bit against the reference with the culling active).
// The returned box bounds, in pixel coordinates, {rho3d <= tau} (projected ellipse, the reference's
// compute_aabb formula with cutoff^2 = tau instead of 9) united with {rho2d <= tau} (disc of radius
// sqrt(tau/2) around the 3-sigma centre), inflated by half a pixel. Surfels whose tau-ellipse reaches
// the camera plane (unbounded projection) or whose box is numerically doubtful are never culled.
// `diag` receives the same kind of bound along the two diagonals, (umin, vmin, umax, vmax) with u = x + y and
// v = x - y: together with the box it is an octagon (8-DOP) around the support region, which for thin surfels
// lying diagonally on screen is far tighter than the box alone. The extent of the projected ellipse along
// any screen direction n follows from the same formula with the row n.x * T[0..2] + n.y * T[3..5].
__device__ __forceinline__ float4 alpha_support_bounds(const float* T, float cx, float cy, float opacity,
                                                       float& tau, float4& diag) {
    const float inf = __int_as_float(0x7f800000);
    const float4 everything = make_float4(-inf, -inf, inf, inf);
    const float4 nothing = make_float4(inf, inf, -inf, -inf);
    diag = everything;
    tau = 2.0f * logf(255.0f * opacity) * 1.001f + 0.001f;
    if (!(tau > 0.0f)) {  // opacity < 1/255 (or NaN): alpha can never reach 1/255
        if (!(tau <= 0.0f)) tau = inf;  // NaN opacity: disable every shortcut
        if (tau != inf) diag = nothing;
        return (tau == inf) ? everything : nothing;
    }
    const float t6 = T[6], t7 = T[7], t8 = T[8];
    const float d = tau * (t6 * t6 + t7 * t7) - t8 * t8;
    if (!(d < -1e-4f * t8 * t8)) return everything;
    const float rc = 1.0f / d;
    const float fx = tau * rc, fz = -rc;
    // centre and squared half extent of the tau-ellipse along the screen direction whose T row is (r0, r1, r2)
    auto extent = [&](float r0, float r1, float r2, float& centre, float& half2) {
        centre = fx * (r0 * t6 + r1 * t7) + fz * r2 * t8;
        half2 = centre * centre - (fx * (r0 * r0 + r1 * r1) + fz * r2 * r2);
    };
    float ex_c, ey_c, eu_c, ev_c, hx, hy, hu, hv;
    extent(T[0], T[1], T[2], ex_c, hx);
    extent(T[3], T[4], T[5], ey_c, hy);
    if (!(hx >= 0.0f) || !(hy >= 0.0f)) return everything;
    const float ex = sqrtf(hx), ey = sqrtf(hy);
    const float rlp = sqrtf(0.5f * tau);
    const float mx = 0.5f + 1e-3f * (ex + rlp + fabsf(ex_c - cx));
    const float my = 0.5f + 1e-3f * (ey + rlp + fabsf(ey_c - cy));
    float4 bb;
    bb.x = fminf(ex_c - ex, cx - rlp) - mx;
    bb.y = fminf(ey_c - ey, cy - rlp) - my;
    bb.z = fmaxf(ex_c + ex, cx + rlp) + mx;
    bb.w = fmaxf(ey_c + ey, cy + rlp) + my;
    if (!(bb.x <= bb.z) || !(bb.y <= bb.w)) return everything;  // NaN guard
    extent(T[0] + T[3], T[1] + T[4], T[2] + T[5], eu_c, hu);
    extent(T[0] - T[3], T[1] - T[4], T[2] - T[5], ev_c, hv);
    if (hu >= 0.0f && hv >= 0.0f) {
        const float eu = sqrtf(hu), ev = sqrtf(hv);
        const float rd = rlp * 1.41421366f;  // the disc's half extent along a (1, +-1) direction, rounded up
        const float cu = cx + cy, cv = cx - cy;
        const float mu = 1.0f + 2e-3f * (eu + rd + fabsf(eu_c - cu));
        const float mv = 1.0f + 2e-3f * (ev + rd + fabsf(ev_c - cv));
        float4 dd;
        dd.x = fminf(eu_c - eu, cu - rd) - mu;
        dd.y = fminf(ev_c - ev, cv - rd) - mv;
        dd.z = fmaxf(eu_c + eu, cu + rd) + mu;
        dd.w = fmaxf(ev_c + ev, cv + rd) + mv;
        if ((dd.x <= dd.z) && (dd.y <= dd.w)) diag = dd;
    }
    return bb;
}

// Per (pixel, surfel) ray-splat evaluation shared by the forward and backward blend kernels.
struct SplatHit {
    float kx, ky, kz, lx, ly, lz;  // the two homogeneous planes
    float pz;                      // cross(k,l).z
    float sx, sy;                  // intersection in splat uv
    float dx, dy;                  // centre - pixel
    float rho3d, rho2d;
    float depth;
    float G;      // exp(-rho/2)
    float alpha;  // min(0.99, opacity * G)
};

// Returns false when the reference would `continue` (p.z == 0, depth < near, power > 0,
// alpha < 1/255). g0 = (Tu.xyz, Tw.x), g1 = (Tv.xyz, Tw.y), g2 = (Tw.z, xy.x, xy.y, opacity).
// `tau` is the surfel's conservative alpha-support threshold (alpha_support_bounds).
__device__ __forceinline__ bool ray_splat(const float4 g0, const float4 g1, const float4 g2, const float tau,
                                          float pxf, float pyf, SplatHit& h) {
    const float Twx = g0.w, Twy = g1.w, Twz = g2.x;
    h.kx = __fmaf_rn(pxf, Twx, -g0.x);
    h.ky = __fmaf_rn(pxf, Twy, -g0.y);
    h.kz = __fmaf_rn(pxf, Twz, -g0.z);
    h.lx = __fmaf_rn(pyf, Twx, -g1.x);
    h.ly = __fmaf_rn(pyf, Twy, -g1.y);
    h.lz = __fmaf_rn(pyf, Twz, -g1.z);
    h.pz = __fmaf_rn(h.kx, h.ly, -__fmul_rn(h.ky, h.lx));
    if (h.pz == 0.0f) return false;
    const float ppx = __fmaf_rn(h.ky, h.lz, -__fmul_rn(h.kz, h.ly));
    const float ppy = __fmaf_rn(h.kz, h.lx, -__fmul_rn(h.kx, h.lz));
    h.dx = __fadd_rn(g2.y, -pxf);
    h.dy = __fadd_rn(g2.z, -pyf);
    const float d2 = __fmaf_rn(h.dy, h.dy, __fmul_rn(h.dx, h.dx));
    h.rho2d = __fadd_rn(d2, d2);
    // provably-safe early rejection before the two IEEE divisions and the exp (see alpha_support_bounds)
    if (h.rho2d > tau && (ppx * ppx + ppy * ppy) > tau * (h.pz * h.pz)) return false;
    h.sx = __fdiv_rn(ppx, h.pz);
    h.sy = __fdiv_rn(ppy, h.pz);
    h.rho3d = __fmaf_rn(h.sx, h.sx, __fmul_rn(h.sy, h.sy));
    // !(rho3d > rho2d) in the binary is `rho3d <= rho2d` in the source; a NaN picks Tw.z
    h.depth = (h.rho3d <= h.rho2d)
                  ? __fadd_rn(Twz, __fmaf_rn(Twx, h.sx, __fmul_rn(Twy, h.sy)))
                  : Twz;
    if (h.depth < kNear) return false;  // a NaN depth is NOT skipped (same as the reference)
    const float power = __fmul_rn(fminf(h.rho3d, h.rho2d), -0.5f);
    if (power > 0.0f) return false;
    h.G = expf(power);
    h.alpha = fminf(__fmul_rn(g2.w, h.G), kAlphaMax);
    if (h.alpha < kAlphaMin) return false;
    return true;
}

// depth -> [0,1] distortion coordinate  m = far/(far-near) * (1 - near/depth)
// Only the distortion / M1 / M2 outputs depend on it (1e-4 bar, no skip decision), so the division is the
// 2-ulp fast one instead of the ~10-instruction IEEE sequence.
__device__ __forceinline__ float distortion_coord(float depth) {
    return __fmul_rn(__fadd_rn(__fdividef(-kNear, depth), 1.0f), kFarOverRange);
}

}  // namespace mrgs
