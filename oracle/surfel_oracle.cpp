// surfel_oracle.cpp — CPU restatement of the reference surfel rasterizer. TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this file's library; nothing under materialrefgs_b200/ does. It is a checker, never the
// thing shipped or measured as the product.
//
// What it restates (rast/ = submodules/diff-surfel-rasterization/ of the MaterialRefGS tree):
//   oracle_preprocess        preprocessCUDA            rast/cuda_rasterizer/forward.cu:163-266
//                            in_frustum / getRect      rast/cuda_rasterizer/auxiliary.h:192-217, :68-78
//                            compute_transmat/aabb     forward.cu:77-159, computeColorFromSH :22-73
//   oracle_bin               duplicateWithKeys + stable (tile|depth) sort + identifyTileRanges
//                            rast/cuda_rasterizer/rasterizer_impl.cu:72-140, :306-324
//   oracle_render_forward    renderCUDA (fwd)          forward.cu:272-463
//   oracle_render_backward   renderCUDA (bwd)          rast/cuda_rasterizer/backward.cu:145-468
//   oracle_preprocess_backward preprocessCUDA (bwd), compute_transmat_aabb, SH bwd
//                            backward.cu:614-669, :471-612, :22-141; quat vjp auxiliary.h:245-289
//
// Pinning: the reference ships no golden vectors for this path, so the
// oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/raster_*.npz were
// produced on a B200 by the unmodified reference extension (oracle/_ref, built by
// oracle/build_ref.sh) with tests/golden/make_golden.py, and tests/test_oracle_cpu.py checks
// this file against them. Arithmetic is fp32 with the same fused-multiply-add placement as the
// reference binary where that was read from its SASS; rsqrt and exp use the host libm (the GPU
// uses MUFU approximations), so a handful of radii / contributor counts may differ by one and
// the tests bound that fraction instead of demanding bit equality on the CPU side. Gradients
// are accumulated in double (the reference's float atomics are order-dependent).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int TILE = 16;
constexpr float NEAR_N = 0.2f, FAR_N = 100.0f;
constexpr float FAR_OVER_RANGE = FAR_N / (FAR_N - NEAR_N);
constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                            -1.0925484305920792f, 0.5462742152960396f};
constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                            0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                            -0.5900435899266435f};

inline float fma3(float a, float b, float c, float d, float e, float f) {
    // a*b + c*d + e*f with (c*d) rounded first, the others fused (transformPoint4x3 pattern)
    return fmaf(e, f, fmaf(a, b, c * d));
}

struct Hit {
    float kx, ky, kz, lx, ly, lz, pz, sx, sy, dx, dy, rho3d, rho2d, depth, G, alpha;
};

// forward.cu:371-399 / backward.cu:302-328. T9 = Tu,Tv,Tw rows.
inline bool ray_splat(const float* T9, float cx, float cy, float opa, float px, float py, Hit& h) {
    const float* Tu = T9;
    const float* Tv = T9 + 3;
    const float* Tw = T9 + 6;
    h.kx = fmaf(px, Tw[0], -Tu[0]); h.ky = fmaf(px, Tw[1], -Tu[1]); h.kz = fmaf(px, Tw[2], -Tu[2]);
    h.lx = fmaf(py, Tw[0], -Tv[0]); h.ly = fmaf(py, Tw[1], -Tv[1]); h.lz = fmaf(py, Tw[2], -Tv[2]);
    h.pz = fmaf(h.kx, h.ly, -(h.ky * h.lx));
    if (h.pz == 0.0f) return false;
    const float ppx = fmaf(h.ky, h.lz, -(h.kz * h.ly));
    const float ppy = fmaf(h.kz, h.lx, -(h.kx * h.lz));
    h.sx = ppx / h.pz;
    h.sy = ppy / h.pz;
    h.rho3d = fmaf(h.sx, h.sx, h.sy * h.sy);
    h.dx = cx - px;
    h.dy = cy - py;
    const float d2 = fmaf(h.dy, h.dy, h.dx * h.dx);
    h.rho2d = d2 + d2;  // FilterInvSquare = 2
    h.depth = (h.rho3d <= h.rho2d) ? Tw[2] + fmaf(Tw[0], h.sx, Tw[1] * h.sy) : Tw[2];
    if (h.depth < NEAR_N) return false;
    const float power = std::fmin(h.rho3d, h.rho2d) * -0.5f;
    if (power > 0.0f) return false;
    h.G = expf(power);
    h.alpha = std::fmin(0.99f, opa * h.G);
    if (h.alpha < 1.0f / 255.0f) return false;
    return true;
}

inline float dist_coord(float depth) { return ((-NEAR_N / depth) + 1.0f) * FAR_OVER_RANGE; }

struct Rot {
    float c[3][3];  // c[col][row]
};

inline Rot quat_to_rot(const float* q) {  // auxiliary.h:220-242, q = (w,x,y,z) in memory order
    const float n2 = fmaf(q[2], q[2], fmaf(q[1], q[1], fmaf(q[0], q[0], q[3] * q[3])));
    const float s = 1.0f / sqrtf(n2);
    const float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    const float wz = w * z, wx = w * x, wy = w * y, yy = y * y, zz = z * z;
    Rot R;
    const float d00 = yy + zz, d11 = fmaf(x, x, zz), d22 = fmaf(x, x, yy);
    const float h01 = fmaf(x, y, wz), h02 = fmaf(x, z, -wy), h10 = fmaf(x, y, -wz);
    const float h12 = fmaf(y, z, wx), h20 = fmaf(x, z, wy), h21 = fmaf(y, z, -wx);
    R.c[0][0] = 1.0f - (d00 + d00); R.c[0][1] = h01 + h01; R.c[0][2] = h02 + h02;
    R.c[1][0] = h10 + h10; R.c[1][1] = 1.0f - (d11 + d11); R.c[1][2] = h12 + h12;
    R.c[2][0] = h20 + h20; R.c[2][1] = h21 + h21; R.c[2][2] = 1.0f - (d22 + d22);
    return R;
}

inline void tile_rect(float cx, float cy, int r, int gx, int gy, int* rc) {  // auxiliary.h:68-78
    const float rf = (float)r;
    auto clampi = [](int v, int hi) { return std::min(hi, std::max(0, v)); };
    rc[0] = clampi((int)((cx - rf) * 0.0625f), gx);
    rc[1] = clampi((int)((cy - rf) * 0.0625f), gy);
    rc[2] = clampi((int)((((cx + rf) + 16.0f) - 1.0f) * 0.0625f), gx);
    rc[3] = clampi((int)((((cy + rf) + 16.0f) - 1.0f) * 0.0625f), gy);
}

}  // namespace

struct OracleScene {
    int32_t P, S, D, M, W, H;
    float tan_fovx, tan_fovy, scale_modifier;
    const float* background;       // [3]
    const float* means3D;          // [P,3]
    const float* shs;              // [P,M,3] or null
    const float* colors_precomp;   // [P,3] or null
    const float* features;         // [P,S]
    const float* opacities;        // [P]
    const float* scales;           // [P,2] or null
    const float* rotations;        // [P,4] or null
    const float* transMat_precomp; // [P,9] or null
    const float* viewmatrix;       // [16]
    const float* projmatrix;       // [16]
    const float* campos;           // [3]
};

// Per-surfel arrays, all caller-allocated with P rows (reference GeometryState, a3 in SURVEY 8a).
struct OracleGeom {
    int32_t* radii;          // [P]
    float* depths;           // [P]
    float* means2D;          // [P,2]
    float* transMat;         // [P,9]
    float* normal_opacity;   // [P,4]
    float* rgb;              // [P,3]
    uint8_t* clamped;        // [P,3]
    uint32_t* tiles_touched; // [P]
};

ORACLE_API int oracle_abi_version() { return 1; }

// Returns R = sum(tiles_touched).
ORACLE_API int64_t oracle_preprocess(const OracleScene* sc, OracleGeom* g) {
    const int P = sc->P, W = sc->W, H = sc->H;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float* vm = sc->viewmatrix;
    const float* pm = sc->projmatrix;
    int64_t R = 0;
#pragma omp parallel for reduction(+ : R) schedule(static)
    for (int i = 0; i < P; ++i) {
        g->radii[i] = 0;
        g->tiles_touched[i] = 0;
        const float* p = sc->means3D + 3 * i;
        // in_frustum: only the view-space depth test survives (auxiliary.h:207)
        const float pvx = fma3(vm[0], p[0], vm[4], p[1], vm[8], p[2]) + vm[12];
        const float pvy = fma3(vm[1], p[0], vm[5], p[1], vm[9], p[2]) + vm[13];
        const float pvz = fma3(vm[2], p[0], vm[6], p[1], vm[10], p[2]) + vm[14];
        if (pvz <= 0.2f) continue;

        float T[9], n[3];
        if (sc->transMat_precomp == nullptr) {
            // compute_transmat, forward.cu:77-125
            const Rot R3 = quat_to_rot(sc->rotations + 4 * i);
            const float msx = sc->scales[2 * i] * sc->scale_modifier;
            const float msy = sc->scales[2 * i + 1] * sc->scale_modifier;
            float L0[3], L1[3];
            for (int r = 0; r < 3; ++r) { L0[r] = msx * R3.c[0][r]; L1[r] = msy * R3.c[1][r]; }
            float A[4][3];
            for (int c = 0; c < 4; ++c) {
                A[c][0] = fma3(L0[0], pm[c], L0[1], pm[4 + c], L0[2], pm[8 + c]);
                A[c][1] = fma3(L1[0], pm[c], L1[1], pm[4 + c], L1[2], pm[8 + c]);
                A[c][2] = fma3(pm[c], p[0], pm[4 + c], p[1], pm[8 + c], p[2]) + pm[12 + c];
            }
            const float w2 = (float)W * 0.5f, wm = (float)(W - 1) * 0.5f;
            const float h2 = (float)H * 0.5f, hm = (float)(H - 1) * 0.5f;
            for (int r = 0; r < 3; ++r) {
                T[r] = fmaf(wm, A[3][r], w2 * A[0][r]);
                T[3 + r] = fmaf(hm, A[3][r], h2 * A[1][r]);
                T[6 + r] = A[3][r];
            }
            for (int r = 0; r < 3; ++r)
                n[r] = fma3(vm[r], R3.c[2][0], vm[4 + r], R3.c[2][1], vm[8 + r], R3.c[2][2]);
        } else {
            std::memcpy(T, sc->transMat_precomp + 9 * i, sizeof(T));
            n[0] = 0.f; n[1] = 0.f; n[2] = 1.f;
        }
        // DUAL_VISIABLE flip, forward.cu:224-229
        const float d = fmaf(pvz, n[2], fmaf(pvx, n[0], pvy * n[1]));
        if (d == 0.0f) continue;
        const float flip = d < 0.0f ? 1.0f : -1.0f;
        for (int r = 0; r < 3; ++r) n[r] *= flip;

        // compute_aabb with cutoff 3, forward.cu:129-159
        const float dist = fmaf(-T[8], T[8], fmaf(T[6] * T[6], 9.0f, (T[7] * T[7]) * 9.0f));
        if (dist == 0.0f) continue;
        const float rc = 1.0f / dist, f9 = rc * 9.0f;
        const float a0 = f9 * T[0], a1 = f9 * T[1], a2 = rc * -T[2];
        const float b0 = f9 * T[3], b1 = f9 * T[4], b2 = rc * -T[5];
        const float cx = fmaf(a2, T[8], fmaf(a1, T[7], a0 * T[6]));
        const float tx = fmaf(a2, T[2], fmaf(a1, T[1], a0 * T[0]));
        const float cy = fmaf(b2, T[8], fmaf(b1, T[7], b0 * T[6]));
        const float ty = fmaf(b2, T[5], fmaf(b1, T[4], b0 * T[3]));
        const float ex = sqrtf(std::fmax(fmaf(cx, cx, -tx), 1e-4f));
        const float ey = sqrtf(std::fmax(fmaf(cy, cy, -ty), 1e-4f));
        const int radius = (int)ceilf(std::fmax(ex, ey));
        int rcT[4];
        tile_rect(cx, cy, radius, gx, gy, rcT);
        const int touched = (rcT[2] - rcT[0]) * (rcT[3] - rcT[1]);
        if (touched == 0) continue;

        if (sc->colors_precomp == nullptr) {
            // computeColorFromSH, forward.cu:22-73
            const float* sh = sc->shs + (size_t)i * sc->M * 3;
            const float dx0 = p[0] - sc->campos[0], dy0 = p[1] - sc->campos[1], dz0 = p[2] - sc->campos[2];
            const float len = sqrtf(fmaf(dz0, dz0, fmaf(dx0, dx0, dy0 * dy0)));
            const float x = dx0 / len, y = dy0 / len, z = dz0 / len;
            float k[16];
            k[0] = SH_C0;
            int nk = 1;
            if (sc->D > 0) {
                k[1] = -(SH_C1 * y); k[2] = SH_C1 * z; k[3] = -(SH_C1 * x);
                nk = 4;
                if (sc->D > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    k[4] = SH_C2[0] * xy; k[5] = SH_C2[1] * yz; k[6] = SH_C2[2] * (((zz + zz) - xx) - yy);
                    k[7] = SH_C2[3] * xz; k[8] = SH_C2[4] * (xx - yy);
                    nk = 9;
                    if (sc->D > 2) {
                        const float f4 = fmaf(4.0f, zz, -xx) - yy;
                        k[9] = (SH_C3[0] * y) * fmaf(3.0f, xx, -yy);
                        k[10] = (SH_C3[1] * xy) * z;
                        k[11] = (SH_C3[2] * y) * f4;
                        k[12] = (SH_C3[3] * z) * fmaf(-3.0f, yy, fmaf(-3.0f, xx, zz + zz));
                        k[13] = (SH_C3[4] * x) * f4;
                        k[14] = (SH_C3[5] * z) * (xx - yy);
                        k[15] = (SH_C3[6] * x) * fmaf(-3.0f, yy, xx);
                        nk = 16;
                    }
                }
            }
            for (int c = 0; c < 3; ++c) {
                float r = k[0] * sh[c];
                for (int j = 1; j < nk; ++j) r = fmaf(k[j], sh[3 * j + c], r);
                r += 0.5f;
                g->clamped[3 * i + c] = r < 0.0f;
                g->rgb[3 * i + c] = std::fmax(r, 0.0f);
            }
        } else {
            for (int c = 0; c < 3; ++c) {
                g->rgb[3 * i + c] = sc->colors_precomp[3 * i + c];
                g->clamped[3 * i + c] = 0;
            }
        }
        g->depths[i] = pvz;
        g->radii[i] = radius;
        g->means2D[2 * i] = cx;
        g->means2D[2 * i + 1] = cy;
        std::memcpy(g->transMat + 9 * i, T, sizeof(T));
        g->normal_opacity[4 * i + 0] = n[0];
        g->normal_opacity[4 * i + 1] = n[1];
        g->normal_opacity[4 * i + 2] = n[2];
        g->normal_opacity[4 * i + 3] = sc->opacities[i];
        g->tiles_touched[i] = (uint32_t)touched;
        R += touched;
    }
    return R;
}

// Emits (tile<<32 | depth bits, id) in surfel order, stable-sorts by key, writes per-tile ranges.
ORACLE_API int oracle_bin(const OracleScene* sc, const OracleGeom* g, int64_t R, uint64_t* keys,
                          uint32_t* point_list, uint32_t* ranges /* [tiles,2] */) {
    const int gx = (sc->W + TILE - 1) / TILE, gy = (sc->H + TILE - 1) / TILE;
    std::vector<std::pair<uint64_t, uint32_t>> kv;
    kv.reserve((size_t)R);
    for (int i = 0; i < sc->P; ++i) {
        if (g->radii[i] <= 0) continue;
        int rc[4];
        tile_rect(g->means2D[2 * i], g->means2D[2 * i + 1], g->radii[i], gx, gy, rc);
        uint32_t bits;
        std::memcpy(&bits, &g->depths[i], 4);
        for (int y = rc[1]; y < rc[3]; ++y)
            for (int x = rc[0]; x < rc[2]; ++x)
                kv.emplace_back(((uint64_t)(uint32_t)(y * gx + x) << 32) | bits, (uint32_t)i);
    }
    if ((int64_t)kv.size() != R) return 1;
    std::stable_sort(kv.begin(), kv.end(),
                     [](const auto& a, const auto& b) { return a.first < b.first; });
    std::memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    for (int64_t i = 0; i < R; ++i) {
        keys[i] = kv[i].first;
        point_list[i] = kv[i].second;
        const uint32_t t = (uint32_t)(kv[i].first >> 32);
        if (i == 0 || t != (uint32_t)(kv[i - 1].first >> 32)) {
            ranges[2 * t] = (uint32_t)i;
            if (i > 0) ranges[2 * (uint32_t)(kv[i - 1].first >> 32) + 1] = (uint32_t)i;
        }
        if (i == R - 1) ranges[2 * t + 1] = (uint32_t)R;
    }
    return 0;
}

// tile_step > 1 renders only every tile_step-th tile (bounded CPU-baseline sample); returns the
// number of tiles rendered. Outputs are planar [C,H,W]; final_T is [3,H,W] (T, M1, M2),
// n_contrib [2,H,W] (last, median).
ORACLE_API int oracle_render_forward(const OracleScene* sc, const OracleGeom* g,
                                     const uint32_t* point_list, const uint32_t* ranges,
                                     int tile_step, float* out_color, float* out_feature,
                                     float* out_others, float* final_T, uint32_t* n_contrib) {
    const int W = sc->W, H = sc->H, S = sc->S;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    int rendered = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rendered)
    for (int tile = 0; tile < gx * gy; ++tile) {
        if (tile_step > 1 && (tile % tile_step) != 0) continue;
        ++rendered;
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
        for (int yy = 0; yy < TILE; ++yy)
            for (int xx = 0; xx < TILE; ++xx) {
                const int px = tx * TILE + xx, py = ty * TILE + yy;
                if (px >= W || py >= H) continue;
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0}, F[24] = {0}, N[3] = {0, 0, 0};
                float D = 0, M1 = 0, M2 = 0, dist = 0, med_depth = 0;
                uint32_t contributor = 0, last = 0, med = 0;
                for (uint32_t e = beg; e < end; ++e) {
                    ++contributor;
                    const uint32_t id = point_list[e];
                    Hit h;
                    if (!ray_splat(g->transMat + 9 * id, g->means2D[2 * id], g->means2D[2 * id + 1],
                                   g->normal_opacity[4 * id + 3], pxf, pyf, h))
                        continue;
                    const float test_T = T * (1.0f - h.alpha);
                    if (test_T < 0.0001f) break;  // done = true
                    const float w = h.alpha * T;
                    const float A = 1.0f - T;
                    const float m = dist_coord(h.depth), mm = m * m;
                    dist = fmaf(w, fmaf(-M1, m + m, fmaf(A, mm, M2)), dist);
                    D = fmaf(h.depth, w, D);
                    M1 = fmaf(w, m, M1);
                    M2 = fmaf(w, mm, M2);
                    if (T > 0.5f) { med_depth = h.depth; med = contributor; }
                    for (int c = 0; c < 3; ++c) N[c] = fmaf(g->normal_opacity[4 * id + c], w, N[c]);
                    for (int c = 0; c < 3; ++c) C[c] = fmaf(w, g->rgb[3 * id + c], C[c]);
                    for (int c = 0; c < S; ++c) F[c] = fmaf(w, sc->features[(size_t)id * S + c], F[c]);
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                final_T[pix] = T; final_T[HW + pix] = M1; final_T[2 * HW + pix] = M2;
                n_contrib[pix] = last; n_contrib[HW + pix] = med;
                for (int c = 0; c < 3; ++c) out_color[c * HW + pix] = fmaf(T, sc->background[c], C[c]);
                for (int c = 0; c < S; ++c) out_feature[c * HW + pix] = F[c];
                out_others[0 * HW + pix] = D;
                out_others[1 * HW + pix] = 1.0f - T;
                for (int c = 0; c < 3; ++c) out_others[(2 + c) * HW + pix] = N[c];
                out_others[5 * HW + pix] = med_depth;
                out_others[6 * HW + pix] = dist;
            }
    }
    return rendered;
}

// Raw per-surfel gradients in double: dT[P,9], dmean2D[P,2], dopacity[P], dnormal[P,3],
// dcolor[P,3], dfeature[P,S]. Must be zero-initialised by the caller.
struct OracleRawGrads {
    double* dT; double* dmean2D; double* dopacity; double* dnormal; double* dcolor; double* dfeature;
};

ORACLE_API int oracle_render_backward(const OracleScene* sc, const OracleGeom* g,
                                      const uint32_t* point_list, const uint32_t* ranges,
                                      int tile_step, const float* final_T, const uint32_t* n_contrib,
                                      const float* dL_dcolor, const float* dL_dfeature,
                                      const float* dL_dothers, OracleRawGrads* out) {
    const int W = sc->W, H = sc->H, S = sc->S;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t HW = (size_t)H * W;
    int rendered = 0;
    // tiles write to shared surfels: serialise the accumulation per thread-private buffers would
    // cost P*26 doubles per thread, so use atomics on doubles instead
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rendered)
    for (int tile = 0; tile < gx * gy; ++tile) {
        if (tile_step > 1 && (tile % tile_step) != 0) continue;
        ++rendered;
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t beg = ranges[2 * tile];
        auto add = [](double* p, double v) {
#pragma omp atomic
            *p += v;
        };
        for (int yy = 0; yy < TILE; ++yy)
            for (int xx = 0; xx < TILE; ++xx) {
                const int px = tx * TILE + xx, py = ty * TILE + yy;
                if (px >= W || py >= H) continue;
                const float pxf = (float)px, pyf = (float)py;
                const size_t pix = (size_t)py * W + px;
                const float T_final = final_T[pix], final_D = final_T[HW + pix], final_D2 = final_T[2 * HW + pix];
                const float final_A = 1.0f - T_final;
                const int last = (int)n_contrib[pix], median = (int)n_contrib[HW + pix];
                float dpix[3 + 24];
                for (int c = 0; c < 3; ++c) dpix[c] = dL_dcolor[c * HW + pix];
                for (int c = 0; c < S; ++c) dpix[3 + c] = dL_dfeature[c * HW + pix];
                const float dL_ddepth = dL_dothers[0 * HW + pix], dL_daccum = dL_dothers[1 * HW + pix];
                const float dL_dn[3] = {dL_dothers[2 * HW + pix], dL_dothers[3 * HW + pix], dL_dothers[4 * HW + pix]};
                const float dL_dmedian = dL_dothers[5 * HW + pix], dL_dreg = dL_dothers[6 * HW + pix];
                float bg_dot = 0;
                for (int c = 0; c < 3; ++c) bg_dot += sc->background[c] * dpix[c];

                float T = T_final, last_alpha = 0, accum[27] = {0}, lastv[27] = {0};
                float last_depth = 0, accum_depth = 0, accum_alpha = 0, last_dL_dT = 0;
                float ln[3] = {0, 0, 0}, an[3] = {0, 0, 0};
                for (int e = last - 1; e >= 0; --e) {  // backward.cu:288-294
                    const uint32_t id = point_list[beg + e];
                    const float* T9 = g->transMat + 9 * id;
                    const float opa = g->normal_opacity[4 * id + 3];
                    Hit h;
                    if (!ray_splat(T9, g->means2D[2 * id], g->means2D[2 * id + 1], opa, pxf, pyf, h)) continue;
                    const float alpha = h.alpha, G = h.G;
                    T = T / (1.0f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0;
                    for (int c = 0; c < 3 + S; ++c) {
                        const float v = c < 3 ? g->rgb[3 * id + c] : sc->features[(size_t)id * S + (c - 3)];
                        accum[c] = last_alpha * lastv[c] + (1.0f - last_alpha) * accum[c];
                        lastv[c] = v;
                        dL_dalpha += (v - accum[c]) * dpix[c];
                        if (c < 3) add(&out->dcolor[3 * (size_t)id + c], (double)(w * dpix[c]));
                        else add(&out->dfeature[(size_t)id * S + (c - 3)], (double)(w * dpix[c]));
                    }
                    const float c_d = h.depth;
                    const float m_d = dist_coord(c_d);
                    const float dmd_dd = (FAR_N * NEAR_N) / ((FAR_N - NEAR_N) * c_d * c_d);
                    float dL_dz = 0;
                    if (e == median - 1) dL_dz += dL_dmedian;
                    const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
                    const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;
                    accum_depth = last_alpha * last_depth + (1.f - last_alpha) * accum_depth;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth) * dL_ddepth;
                    accum_alpha = last_alpha * 1.0f + (1.f - last_alpha) * accum_alpha;
                    dL_dalpha += (1 - accum_alpha) * dL_daccum;
                    for (int c = 0; c < 3; ++c) {
                        const float nv = g->normal_opacity[4 * id + c];
                        an[c] = last_alpha * ln[c] + (1.f - last_alpha) * an[c];
                        ln[c] = nv;
                        dL_dalpha += (nv - an[c]) * dL_dn[c];
                        add(&out->dnormal[3 * (size_t)id + c], (double)(w * dL_dn[c]));
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = opa * dL_dalpha;
                    dL_dz += w * dL_ddepth;
                    if (h.rho3d <= h.rho2d) {  // backward.cu:424-454
                        const float* Tw = T9 + 6;
                        const float dsx = dL_dG * -G * h.sx + dL_dz * Tw[0];
                        const float dsy = dL_dG * -G * h.sy + dL_dz * Tw[1];
                        const float ax = dsx / h.pz, ay = dsy / h.pz;
                        const float dp[3] = {ax, ay, -(ax * h.sx + ay * h.sy)};
                        const float k[3] = {h.kx, h.ky, h.kz}, l[3] = {h.lx, h.ly, h.lz};
                        const float dk[3] = {l[1] * dp[2] - l[2] * dp[1], l[2] * dp[0] - l[0] * dp[2], l[0] * dp[1] - l[1] * dp[0]};
                        const float dl[3] = {dp[1] * k[2] - dp[2] * k[1], dp[2] * k[0] - dp[0] * k[2], dp[0] * k[1] - dp[1] * k[0]};
                        const float dzT[3] = {h.sx, h.sy, 1.0f};
                        for (int c = 0; c < 3; ++c) {
                            add(&out->dT[9 * (size_t)id + c], (double)-dk[c]);
                            add(&out->dT[9 * (size_t)id + 3 + c], (double)-dl[c]);
                            add(&out->dT[9 * (size_t)id + 6 + c], (double)(pxf * dk[c] + pyf * dl[c] + dL_dz * dzT[c]));
                        }
                    } else {  // backward.cu:455-462
                        add(&out->dmean2D[2 * (size_t)id], (double)(dL_dG * (-G * 2.0f * h.dx)));
                        add(&out->dmean2D[2 * (size_t)id + 1], (double)(dL_dG * (-G * 2.0f * h.dy)));
                        add(&out->dT[9 * (size_t)id + 8], (double)dL_dz);
                    }
                    add(&out->dopacity[id], (double)(G * dL_dalpha));
                }
            }
    }
    return rendered;
}

struct OracleGrads {  // final outputs, float, caller-allocated and ZERO-INITIALISED
    float* dL_dmeans2D;   // [P,3]
    float* dL_dmeans3D;   // [P,3]
    float* dL_dtransMat;  // [P,9]
    float* dL_dsh;        // [P,M,3]
    float* dL_dscales;    // [P,2]
    float* dL_drotations; // [P,4]
};

ORACLE_API int oracle_preprocess_backward(const OracleScene* sc, const OracleGeom* g,
                                          const OracleRawGrads* raw, OracleGrads* out) {
    const int P = sc->P, M = sc->M;
    // backward.cu:646-647 with focal = size / (2 tan), rasterizer_impl.cu:398-399
    const float focal_x = sc->W / (2.0f * sc->tan_fovx), focal_y = sc->H / (2.0f * sc->tan_fovy);
    volatile float fw = focal_x * sc->tan_fovx, fh = focal_y * sc->tan_fovy;
    const int W = (int)(fw * 2.0f), H = (int)(fh * 2.0f);
    const float* vm = sc->viewmatrix;
    const float* pm = sc->projmatrix;
    const bool precomp = sc->scales == nullptr;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        if (!(g->radii[i] > 0)) continue;
        double dT[9];
        for (int k = 0; k < 9; ++k) dT[k] = raw->dT[9 * (size_t)i + k];
        double T[9], P3[3][4], Rm[3][3];
        float sx = 0, sy = 0;
        const float* p = sc->means3D + 3 * i;
        double nrm[3] = {0, 0, 0};
        if (precomp) {
            for (int k = 0; k < 9; ++k) T[k] = sc->transMat_precomp[9 * i + k];
        } else {
            const Rot R3 = quat_to_rot(sc->rotations + 4 * i);
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Rm[c][r] = R3.c[c][r];
            sx = sc->scales[2 * i]; sy = sc->scales[2 * i + 1];  // scale_modifier ignored (backward.cu:509)
            const double w2 = 0.5 * W, wm = 0.5 * (W - 1), h2 = 0.5 * H, hm = 0.5 * (H - 1);
            for (int k = 0; k < 4; ++k) {
                P3[0][k] = pm[4 * k + 0] * w2 + pm[4 * k + 3] * wm;
                P3[1][k] = pm[4 * k + 1] * h2 + pm[4 * k + 3] * hm;
                P3[2][k] = pm[4 * k + 3];
            }
            for (int c = 0; c < 3; ++c) {
                T[3 * c + 0] = sx * (Rm[0][0] * P3[c][0] + Rm[0][1] * P3[c][1] + Rm[0][2] * P3[c][2]);
                T[3 * c + 1] = sy * (Rm[1][0] * P3[c][0] + Rm[1][1] * P3[c][1] + Rm[1][2] * P3[c][2]);
                T[3 * c + 2] = p[0] * P3[c][0] + p[1] * P3[c][1] + p[2] * P3[c][2] + P3[c][3];
            }
            for (int r = 0; r < 3; ++r) nrm[r] = vm[r] * Rm[2][0] + vm[4 + r] * Rm[2][1] + vm[8 + r] * Rm[2][2];
        }
        const double m2x = raw->dmean2D[2 * (size_t)i], m2y = raw->dmean2D[2 * (size_t)i + 1];
        if (m2x != 0 || m2y != 0) {  // backward.cu:543-582
            const double t0 = T[6], t1 = T[7], t2 = T[8];
            const double f = 1.0 / (t0 * t0 + t1 * t1 - t2 * t2);
            const double c0 = f - 2 * f * f * t0 * t0, c1 = f - 2 * f * f * t1 * t1, c2 = f + 2 * f * f * t2 * t2;
            dT[0] += m2x * f * t0; dT[1] += m2x * f * t1; dT[2] += m2x * -f * t2;
            dT[3] += m2y * f * t0; dT[4] += m2y * f * t1; dT[5] += m2y * -f * t2;
            dT[6] += m2x * T[0] * c0 + m2y * T[3] * c0;
            dT[7] += m2x * T[1] * c1 + m2y * T[4] * c1;
            dT[8] += m2x * -T[2] * c2 + m2y * -T[5] * c2;
        }
        const double depth = g->transMat[9 * i + 8];
        double dmean[3] = {0, 0, 0};
        if (!precomp) {
            double dM[3][3];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k)
                    dM[r][k] = dT[r] * P3[0][k] + dT[3 + r] * P3[1][k] + dT[6 + r] * P3[2][k];
            const double* dn = raw->dnormal + 3 * (size_t)i;
            double dtn[3] = {vm[0] * dn[0] + vm[1] * dn[1] + vm[2] * dn[2],
                             vm[4] * dn[0] + vm[5] * dn[1] + vm[6] * dn[2],
                             vm[8] * dn[0] + vm[9] * dn[1] + vm[10] * dn[2]};
            const double pvx = vm[0] * p[0] + vm[4] * p[1] + vm[8] * p[2] + vm[12];
            const double pvy = vm[1] * p[0] + vm[5] * p[1] + vm[9] * p[2] + vm[13];
            const double pvz = vm[2] * p[0] + vm[6] * p[1] + vm[10] * p[2] + vm[14];
            const double flip = (pvx * nrm[0] + pvy * nrm[1] + pvz * nrm[2]) < 0 ? 1.0 : -1.0;
            for (int k = 0; k < 3; ++k) dtn[k] *= flip;
            double dR[3][3];
            for (int k = 0; k < 3; ++k) { dR[0][k] = dM[0][k] * sx; dR[1][k] = dM[1][k] * sy; dR[2][k] = dtn[k]; }
            out->dL_dscales[2 * i] = (float)(dM[0][0] * Rm[0][0] + dM[0][1] * Rm[0][1] + dM[0][2] * Rm[0][2]);
            out->dL_dscales[2 * i + 1] = (float)(dM[1][0] * Rm[1][0] + dM[1][1] * Rm[1][1] + dM[1][2] * Rm[1][2]);
            for (int k = 0; k < 3; ++k) dmean[k] = dM[2][k];
            const float* q = sc->rotations + 4 * i;  // quat_to_rotmat_vjp, auxiliary.h:245-289
            const double s = 1.0 / std::sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3]);
            const double w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
            float* dq = out->dL_drotations + 4 * i;
            dq[0] = (float)(2 * (x * (dR[1][2] - dR[2][1]) + y * (dR[2][0] - dR[0][2]) + z * (dR[0][1] - dR[1][0])));
            dq[1] = (float)(2 * (-2 * x * (dR[1][1] + dR[2][2]) + y * (dR[0][1] + dR[1][0]) + z * (dR[0][2] + dR[2][0]) + w * (dR[1][2] - dR[2][1])));
            dq[2] = (float)(2 * (x * (dR[0][1] + dR[1][0]) - 2 * y * (dR[0][0] + dR[2][2]) + z * (dR[1][2] + dR[2][1]) + w * (dR[2][0] - dR[0][2])));
            dq[3] = (float)(2 * (x * (dR[0][2] + dR[2][0]) + y * (dR[1][2] + dR[2][1]) - 2 * z * (dR[0][0] + dR[1][1]) + w * (dR[0][1] - dR[1][0])));
            for (int k = 0; k < 9; ++k) out->dL_dtransMat[9 * i + k] = (float)raw->dT[9 * (size_t)i + k];
            out->dL_dmeans2D[3 * i] = (float)(raw->dT[9 * (size_t)i + 2] * depth * 0.5 * W);   // backward.cu:666-668
            out->dL_dmeans2D[3 * i + 1] = (float)(raw->dT[9 * (size_t)i + 5] * depth * 0.5 * H);
        } else {
            for (int k = 0; k < 9; ++k) out->dL_dtransMat[9 * i + k] = (float)dT[k];
            out->dL_dmeans2D[3 * i] = (float)(dT[2] * depth * 0.5 * W);
            out->dL_dmeans2D[3 * i + 1] = (float)(dT[5] * depth * 0.5 * H);
        }
        if (sc->shs != nullptr) {  // backward.cu:22-141
            const float* sh = sc->shs + (size_t)i * M * 3;
            float* dsh = out->dL_dsh + (size_t)i * M * 3;
            double gcol[3];
            for (int c = 0; c < 3; ++c) gcol[c] = g->clamped[3 * i + c] ? 0.0 : raw->dcolor[3 * (size_t)i + c];
            const double d0[3] = {(double)p[0] - sc->campos[0], (double)p[1] - sc->campos[1], (double)p[2] - sc->campos[2]};
            const double s2 = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
            const double il = 1.0 / std::sqrt(s2);
            const double x = d0[0] * il, y = d0[1] * il, z = d0[2] * il;
            double k[16] = {0}, kx[16] = {0}, ky[16] = {0}, kz[16] = {0};  // basis and its derivatives
            k[0] = SH_C0;
            int nk = 1;
            if (sc->D > 0) {
                nk = 4;
                k[1] = -SH_C1 * y; k[2] = SH_C1 * z; k[3] = -SH_C1 * x;
                ky[1] = -SH_C1; kz[2] = SH_C1; kx[3] = -SH_C1;
                if (sc->D > 1) {
                    nk = 9;
                    const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    k[4] = SH_C2[0] * xy; k[5] = SH_C2[1] * yz; k[6] = SH_C2[2] * (2 * zz - xx - yy);
                    k[7] = SH_C2[3] * xz; k[8] = SH_C2[4] * (xx - yy);
                    kx[4] = SH_C2[0] * y; ky[4] = SH_C2[0] * x;
                    ky[5] = SH_C2[1] * z; kz[5] = SH_C2[1] * y;
                    kx[6] = SH_C2[2] * -2 * x; ky[6] = SH_C2[2] * -2 * y; kz[6] = SH_C2[2] * 4 * z;
                    kx[7] = SH_C2[3] * z; kz[7] = SH_C2[3] * x;
                    kx[8] = SH_C2[4] * 2 * x; ky[8] = SH_C2[4] * -2 * y;
                    if (sc->D > 2) {
                        nk = 16;
                        k[9] = SH_C3[0] * y * (3 * xx - yy); k[10] = SH_C3[1] * xy * z;
                        k[11] = SH_C3[2] * y * (4 * zz - xx - yy); k[12] = SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy);
                        k[13] = SH_C3[4] * x * (4 * zz - xx - yy); k[14] = SH_C3[5] * z * (xx - yy);
                        k[15] = SH_C3[6] * x * (xx - 3 * yy);
                        kx[9] = SH_C3[0] * 6 * xy; ky[9] = SH_C3[0] * 3 * (xx - yy);
                        kx[10] = SH_C3[1] * yz; ky[10] = SH_C3[1] * xz; kz[10] = SH_C3[1] * xy;
                        kx[11] = SH_C3[2] * -2 * xy; ky[11] = SH_C3[2] * (-3 * yy + 4 * zz - xx); kz[11] = SH_C3[2] * 8 * yz;
                        kx[12] = SH_C3[3] * -6 * xz; ky[12] = SH_C3[3] * -6 * yz; kz[12] = SH_C3[3] * 3 * (2 * zz - xx - yy);
                        kx[13] = SH_C3[4] * (-3 * xx + 4 * zz - yy); ky[13] = SH_C3[4] * -2 * xy; kz[13] = SH_C3[4] * 8 * xz;
                        kx[14] = SH_C3[5] * 2 * xz; ky[14] = SH_C3[5] * -2 * yz; kz[14] = SH_C3[5] * (xx - yy);
                        kx[15] = SH_C3[6] * 3 * (xx - yy); ky[15] = SH_C3[6] * -6 * xy;
                    }
                }
            }
            double ddir[3] = {0, 0, 0};
            for (int j = 0; j < nk; ++j)
                for (int c = 0; c < 3; ++c) {
                    dsh[3 * j + c] = (float)(k[j] * gcol[c]);
                    ddir[0] += kx[j] * sh[3 * j + c] * gcol[c];
                    ddir[1] += ky[j] * sh[3 * j + c] * gcol[c];
                    ddir[2] += kz[j] * sh[3 * j + c] * gcol[c];
                }
            const double inv32 = 1.0 / std::sqrt(s2 * s2 * s2);  // dnormvdv, auxiliary.h:129-139
            dmean[0] += ((s2 - d0[0] * d0[0]) * ddir[0] - d0[1] * d0[0] * ddir[1] - d0[2] * d0[0] * ddir[2]) * inv32;
            dmean[1] += (-d0[0] * d0[1] * ddir[0] + (s2 - d0[1] * d0[1]) * ddir[1] - d0[2] * d0[1] * ddir[2]) * inv32;
            dmean[2] += (-d0[0] * d0[2] * ddir[0] - d0[1] * d0[2] * ddir[1] + (s2 - d0[2] * d0[2]) * ddir[2]) * inv32;
        }
        for (int k = 0; k < 3; ++k) out->dL_dmeans3D[3 * i + k] = (float)dmean[k];
    }
    return 0;
}

ORACLE_API int oracle_num_threads() {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
