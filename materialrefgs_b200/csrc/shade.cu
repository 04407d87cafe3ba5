// shade.cu — fused deferred split-sum PBR shading of the rasterized G-buffer, forward and backward.
//
// One thread per pixel does what the reference spreads over ~35 eager torch kernels and two
// nvdiffrast texture ops:
//   normal to world space and /alpha      gaussian_renderer/__init__.py:46-48, :419-420
//   camera ray, reflection, N.V           utils/refl_utils.py:54-73, :95-98, :367-371
//   FG LUT bilinear-clamp fetch           utils/refl_utils.py:373-374
//   roughness -> mip level                scene/light.py:88-96
//   seamless trilinear cube fetch+sigmoid scene/light.py:118-129
//   specular weight / specular / final    utils/refl_utils.py:377, :400-401; __init__.py:433-445
// Texture semantics restate nvdiffrast's dr.texture (not vendored by the reference):
// texel centres at (i+0.5)/size, u*size-0.5 taps, LUT taps clamped, cube taps that leave a face
// continue on the adjacent face, the missing tap at a cube corner is the mean of the other three,
// mip level = clamp(bias, 0, L-1) with linear blending between floor and floor+1.
#include <algorithm>

#include "cube_sample.cuh"
#include "kernels.cuh"

namespace mrgs {

namespace {

struct MipLevel {
    int l0, l1;
    float f;
    float dlevel_drough;
};

// EnvLight.get_mip (scene/light.py:88-96) followed by dr.texture's level clamp
__device__ __forceinline__ MipLevel rough_to_level(float r, float min_r, float max_r, int L) {
    float lvl, dl;
    if (r < max_r) {
        lvl = (fminf(fmaxf(r, min_r), max_r) - min_r) / (max_r - min_r) * (float)(L - 2);
        dl = (r >= min_r) ? (float)(L - 2) / (max_r - min_r) : 0.f;
    } else {
        lvl = (fminf(fmaxf(r, max_r), 1.0f) - max_r) / (1.0f - max_r) + (float)(L - 2);
        dl = (r <= 1.0f) ? 1.0f / (1.0f - max_r) : 0.f;
    }
    MipLevel m;
    const float c = fminf(fmaxf(lvl, 0.f), (float)(L - 1));
    m.l0 = (int)floorf(c);
    m.l1 = min(m.l0 + 1, L - 1);
    m.f = (m.l1 == m.l0) ? 0.f : c - (float)m.l0;
    m.dlevel_drough = (lvl < 0.f || lvl > (float)(L - 1)) ? 0.f : dl;
    return m;
}

// bilinear fetch of the [256,256,2] LUT with clamp boundary; uv.x -> width, uv.y -> height
struct LutFetch {
    float fx, fy;          // the two channels
    float dfx_du, dfx_dv, dfy_du, dfy_dv;
};
__device__ __forceinline__ LutFetch lut_fetch(const float* __restrict__ lut, float u, float v) {
    constexpr int N = 256;
    const float Uc = fminf(fmaxf(u * N - 0.5f, 0.f), (float)(N - 1));
    const float Vc = fminf(fmaxf(v * N - 0.5f, 0.f), (float)(N - 1));
    const bool u_in = (u * N - 0.5f) >= 0.f && (u * N - 0.5f) <= (float)(N - 1);
    const bool v_in = (v * N - 0.5f) >= 0.f && (v * N - 0.5f) <= (float)(N - 1);
    const int x0 = (int)floorf(Uc), y0 = (int)floorf(Vc);
    const int x1 = min(x0 + 1, N - 1), y1 = min(y0 + 1, N - 1);
    const float fu = Uc - (float)x0, fv = Vc - (float)y0;
    const float2 a00 = __ldg(reinterpret_cast<const float2*>(lut) + y0 * N + x0);
    const float2 a10 = __ldg(reinterpret_cast<const float2*>(lut) + y0 * N + x1);
    const float2 a01 = __ldg(reinterpret_cast<const float2*>(lut) + y1 * N + x0);
    const float2 a11 = __ldg(reinterpret_cast<const float2*>(lut) + y1 * N + x1);
    LutFetch r;
    const float w00 = (1.f - fu) * (1.f - fv), w10 = fu * (1.f - fv), w01 = (1.f - fu) * fv, w11 = fu * fv;
    r.fx = w00 * a00.x + w10 * a10.x + w01 * a01.x + w11 * a11.x;
    r.fy = w00 * a00.y + w10 * a10.y + w01 * a01.y + w11 * a11.y;
    const float su = u_in ? (float)N : 0.f, sv = v_in ? (float)N : 0.f;
    r.dfx_du = su * ((1.f - fv) * (a10.x - a00.x) + fv * (a11.x - a01.x));
    r.dfy_du = su * ((1.f - fv) * (a10.y - a00.y) + fv * (a11.y - a01.y));
    r.dfx_dv = sv * ((1.f - fu) * (a01.x - a00.x) + fu * (a11.x - a10.x));
    r.dfy_dv = sv * ((1.f - fu) * (a01.y - a00.y) + fu * (a11.y - a10.y));
    return r;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr float kSrgbEps = 1.1920929e-07f;
__device__ __forceinline__ float linear_to_srgb(float x) {  // utils/graphics_utils.py:102-110
    const float s0 = (323.0f / 25.0f) * x;
    const float s1 = (211.0f * powf(fmaxf(x, kSrgbEps), 5.0f / 12.0f) - 11.0f) / 200.0f;
    return x <= 0.0031308f ? s0 : s1;
}
__device__ __forceinline__ float dlinear_to_srgb(float x) {
    if (x <= 0.0031308f) return 323.0f / 25.0f;
    return (211.0f / 200.0f) * (5.0f / 12.0f) * powf(x, -7.0f / 12.0f);
}

struct PixelShade {
    // inputs
    F3 base, albedo, nv;
    float rs, ro, A;
    // intermediates
    F3 nw, n, wo, r, rr, Ld, t, sw, spec, diff, fin_lin;
    float Ac, ndv, rl;
    LutFetch fg;
    MipLevel mip;
    FaceUV fuv;
    Bilinear b0, b1;
};

template <bool GRAD>
__device__ __forceinline__ void shade_pixel(const MrgsShadeArgs& p, int x, int y, size_t pix, size_t HW,
                                            PixelShade& s) {
    s.base = {p.base_color[pix], p.base_color[HW + pix], p.base_color[2 * HW + pix]};
    s.rs = p.features[pix];
    s.ro = p.features[HW + pix];
    s.albedo = {p.features[2 * HW + pix], p.features[3 * HW + pix], p.features[4 * HW + pix]};
    s.A = p.allmap[kAlphaOff * HW + pix];
    s.nv = {p.allmap[(kNormalOff + 0) * HW + pix], p.allmap[(kNormalOff + 1) * HW + pix],
            p.allmap[(kNormalOff + 2) * HW + pix]};
    const float* Q = p.normal_matrix;
    s.nw = {Q[0] * s.nv.x + Q[1] * s.nv.y + Q[2] * s.nv.z, Q[3] * s.nv.x + Q[4] * s.nv.y + Q[5] * s.nv.z,
            Q[6] * s.nv.x + Q[7] * s.nv.y + Q[8] * s.nv.z};
    s.Ac = fmaxf(s.A, 1e-6f);
    s.n = (1.0f / s.Ac) * s.nw;
    const float* M = p.ray_matrix;
    const float fx = (float)x, fy = (float)y;
    F3 d = {M[0] * fx + M[1] * fy + M[2], M[3] * fx + M[4] * fy + M[5], M[6] * fx + M[7] * fy + M[8]};
    d = (1.0f / sqrtf(dot(d, d))) * d;
    s.wo = {-d.x, -d.y, -d.z};
    s.ndv = dot(s.n, s.wo);
    s.r = (2.0f * s.ndv) * s.n - s.wo;
    s.rl = fmaxf(sqrtf(dot(s.r, s.r)), 1e-20f);
    s.rr = (1.0f / s.rl) * s.r;

    s.fg = lut_fetch(p.lut, fminf(fmaxf(s.ndv, 0.f), 1.f), fminf(fmaxf(s.ro, 0.f), 1.f));

    s.mip = rough_to_level(s.ro, p.min_roughness, p.max_roughness, p.num_levels);
    s.fuv = dir_to_face(s.rr);
    cube_bilinear<GRAD>(p.levels[s.mip.l0], p.base_res >> s.mip.l0, s.fuv.face, s.fuv.u, s.fuv.v, s.b0);
    s.t = s.b0.val;
    if (s.mip.l1 != s.mip.l0) {
        cube_bilinear<GRAD>(p.levels[s.mip.l1], p.base_res >> s.mip.l1, s.fuv.face, s.fuv.u, s.fuv.v, s.b1);
        s.t = (1.0f - s.mip.f) * s.b0.val + s.mip.f * s.b1.val;
    }
    s.Ld = {sigmoidf(s.t.x), sigmoidf(s.t.y), sigmoidf(s.t.z)};

    const float k0 = 0.04f * (1.0f - s.rs);
    s.sw = {(k0 + s.albedo.x * s.rs) * s.fg.fx + s.fg.fy, (k0 + s.albedo.y * s.rs) * s.fg.fx + s.fg.fy,
            (k0 + s.albedo.z * s.rs) * s.fg.fx + s.fg.fy};
    s.spec = {s.Ld.x * s.A * s.sw.x, s.Ld.y * s.A * s.sw.y, s.Ld.z * s.A * s.sw.z};
    s.diff = (1.0f - s.rs) * s.base;
    s.fin_lin = s.diff + s.spec;
}

__device__ __forceinline__ void store3(float* out, size_t pix, size_t HW, F3 v) {
    if (out == nullptr) return;
    out[pix] = v.x;
    out[HW + pix] = v.y;
    out[2 * HW + pix] = v.z;
}
__device__ __forceinline__ F3 load3p(const float* in, size_t pix, size_t HW) {
    if (in == nullptr) return {0.f, 0.f, 0.f};
    return {in[pix], in[HW + pix], in[2 * HW + pix]};
}

__global__ void __launch_bounds__(256) shade_fwd_kernel(const MrgsShadeArgs p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.width || y >= p.height) return;
    const size_t HW = (size_t)p.width * p.height, pix = (size_t)y * p.width + x;
    PixelShade s;
    shade_pixel<false>(p, x, y, pix, HW, s);
    F3 fin = s.fin_lin;
    if (p.srgb) fin = {linear_to_srgb(fin.x), linear_to_srgb(fin.y), linear_to_srgb(fin.z)};
    const float om = 1.0f - s.A;
    fin = {fin.x + p.background[0] * om, fin.y + p.background[1] * om, fin.z + p.background[2] * om};
    store3(p.out_final, pix, HW, fin);
    store3(p.out_specular, pix, HW, s.spec);
    store3(p.out_direct, pix, HW, s.Ld);
    store3(p.out_normal, pix, HW, s.nw);
    store3(p.out_diffuse, pix, HW, s.diff);
}

// gradient levels are float4 per texel (rgb + unused pad) so one tap is ONE 16-byte vector reduction
// (red.global.add.v4.f32, sm_90+) instead of three scalar atomics
__device__ __forceinline__ void scatter_bilinear(float* __restrict__ grad_tex, const Bilinear& b, F3 g) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (b.idx[k] < 0 || b.w[k] == 0.f) continue;
        float4* q = reinterpret_cast<float4*>(grad_tex) + (size_t)b.idx[k];
        atomicAdd(q, make_float4(b.w[k] * g.x, b.w[k] * g.y, b.w[k] * g.z, 0.0f));
    }
}

// gradient of the face coordinates (u, v) back to the direction: u = su*a*m + .5, v = sv*b*m + .5, m = .5/|c|
__device__ __forceinline__ void face_uv_backward(const FaceUV& f, F3 d, float g_u, float g_v, float g_d[3]) {
    g_d[0] = g_d[1] = g_d[2] = 0.f;
    const float dv[3] = {d.x, d.y, d.z};
    const float a = dv[f.ia], b = dv[f.ib], c = dv[f.ic];
    const bool u_free = f.u > 0.f && f.u < 1.f, v_free = f.v > 0.f && f.v < 1.f;
    const float gu = u_free ? g_u : 0.f, gv = v_free ? g_v : 0.f;
    g_d[f.ia] += gu * f.su * f.m;
    g_d[f.ib] += gv * f.sv * f.m;
    g_d[f.ic] += -(gu * f.su * a + gv * f.sv * b) * f.m / fabsf(c) * f.csign;
}

__global__ void __launch_bounds__(256) shade_bwd_kernel(const MrgsShadeArgs p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.width || y >= p.height) return;
    const size_t HW = (size_t)p.width * p.height, pix = (size_t)y * p.width + x;
    PixelShade s;
    shade_pixel<true>(p, x, y, pix, HW, s);

    const F3 g_fin = load3p(p.dL_dfinal, pix, HW);
    const F3 g_spec_out = load3p(p.dL_dspecular, pix, HW);
    const F3 g_diff_out = load3p(p.dL_ddiffuse, pix, HW);
    const F3 g_nw_out = load3p(p.dL_dnormal, pix, HW);

    float g_A = -(g_fin.x * p.background[0] + g_fin.y * p.background[1] + g_fin.z * p.background[2]);
    F3 g_lin = g_fin;
    if (p.srgb)
        g_lin = {g_fin.x * dlinear_to_srgb(s.fin_lin.x), g_fin.y * dlinear_to_srgb(s.fin_lin.y),
                 g_fin.z * dlinear_to_srgb(s.fin_lin.z)};
    const F3 g_diff = g_lin + g_diff_out;
    const F3 g_spec = g_lin + g_spec_out;

    float g_rs = -dot(s.base, g_diff);
    const F3 g_base = (1.0f - s.rs) * g_diff;

    const F3 g_Ld = {g_spec.x * s.A * s.sw.x, g_spec.y * s.A * s.sw.y, g_spec.z * s.A * s.sw.z};
    g_A += g_spec.x * s.Ld.x * s.sw.x + g_spec.y * s.Ld.y * s.sw.y + g_spec.z * s.Ld.z * s.sw.z;
    const F3 g_sw = {g_spec.x * s.Ld.x * s.A, g_spec.y * s.Ld.y * s.A, g_spec.z * s.Ld.z * s.A};

    g_rs += s.fg.fx * (g_sw.x * (s.albedo.x - 0.04f) + g_sw.y * (s.albedo.y - 0.04f) + g_sw.z * (s.albedo.z - 0.04f));
    const F3 g_albedo = (s.rs * s.fg.fx) * g_sw;
    const float k0 = 0.04f * (1.0f - s.rs);
    const float g_fgx = g_sw.x * (k0 + s.albedo.x * s.rs) + g_sw.y * (k0 + s.albedo.y * s.rs) + g_sw.z * (k0 + s.albedo.z * s.rs);
    const float g_fgy = g_sw.x + g_sw.y + g_sw.z;

    // LUT -> N.V and roughness (torch.clamp passes the gradient on the closed interval)
    float g_ndv = 0.f, g_ro = 0.f;
    if (s.ndv >= 0.f && s.ndv <= 1.f) g_ndv = g_fgx * s.fg.dfx_du + g_fgy * s.fg.dfy_du;
    if (s.ro >= 0.f && s.ro <= 1.f) g_ro = g_fgx * s.fg.dfx_dv + g_fgy * s.fg.dfy_dv;

    // environment: sigmoid, mip blend, texels, level, face coordinates
    const F3 g_t = {g_Ld.x * s.Ld.x * (1.0f - s.Ld.x), g_Ld.y * s.Ld.y * (1.0f - s.Ld.y),
                    g_Ld.z * s.Ld.z * (1.0f - s.Ld.z)};
    float g_u, g_v;
    if (s.mip.l1 != s.mip.l0) {
        const float f = s.mip.f;
        if (p.dL_dlevels[s.mip.l0]) scatter_bilinear(p.dL_dlevels[s.mip.l0], s.b0, (1.0f - f) * g_t);
        if (p.dL_dlevels[s.mip.l1]) scatter_bilinear(p.dL_dlevels[s.mip.l1], s.b1, f * g_t);
        g_ro += dot(g_t, s.b1.val - s.b0.val) * s.mip.dlevel_drough;
        g_u = dot(g_t, (1.0f - f) * s.b0.dval_du + f * s.b1.dval_du);
        g_v = dot(g_t, (1.0f - f) * s.b0.dval_dv + f * s.b1.dval_dv);
    } else {
        if (p.dL_dlevels[s.mip.l0]) scatter_bilinear(p.dL_dlevels[s.mip.l0], s.b0, g_t);
        g_u = dot(g_t, s.b0.dval_du);
        g_v = dot(g_t, s.b0.dval_dv);
    }
    float g_rr[3];
    face_uv_backward(s.fuv, s.rr, g_u, g_v, g_rr);
    // rr = r / max(|r|, eps)
    F3 g_r = {0.f, 0.f, 0.f};
    {
        const F3 grr = {g_rr[0], g_rr[1], g_rr[2]};
        if (s.rl > 1e-20f)
            g_r = (1.0f / s.rl) * (grr - dot(s.rr, grr) * s.rr);
        else
            g_r = (1.0f / s.rl) * grr;
    }
    // r = 2 n (n.wo) - wo ; ndv = n.wo
    F3 g_n = (2.0f * s.ndv) * g_r;
    g_ndv += 2.0f * dot(s.n, g_r);
    g_n = g_n + g_ndv * s.wo;
    // n = nw / max(A, 1e-6)
    F3 g_nw = (1.0f / s.Ac) * g_n;
    if (s.A >= 1e-6f) g_A += -dot(s.n, g_n) / s.Ac;
    g_nw = g_nw + g_nw_out;
    const float* Q = p.normal_matrix;
    const F3 g_nv = {Q[0] * g_nw.x + Q[3] * g_nw.y + Q[6] * g_nw.z, Q[1] * g_nw.x + Q[4] * g_nw.y + Q[7] * g_nw.z,
                     Q[2] * g_nw.x + Q[5] * g_nw.y + Q[8] * g_nw.z};

    store3(p.dL_dbase_color, pix, HW, g_base);
    if (p.dL_dfeatures) {
        p.dL_dfeatures[pix] = g_rs;
        p.dL_dfeatures[HW + pix] = g_ro;
        p.dL_dfeatures[2 * HW + pix] = g_albedo.x;
        p.dL_dfeatures[3 * HW + pix] = g_albedo.y;
        p.dL_dfeatures[4 * HW + pix] = g_albedo.z;
    }
    if (p.dL_dallmap) {
        p.dL_dallmap[kAlphaOff * HW + pix] = g_A;
        p.dL_dallmap[(kNormalOff + 0) * HW + pix] = g_nv.x;
        p.dL_dallmap[(kNormalOff + 1) * HW + pix] = g_nv.y;
        p.dL_dallmap[(kNormalOff + 2) * HW + pix] = g_nv.z;
    }
}

__global__ void __launch_bounds__(256)
envlight_query_kernel(const MrgsShadeArgs p, long long n, const float* __restrict__ dirs,
                      const float* __restrict__ roughness, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F3 d = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
    const FaceUV f = dir_to_face(d);
    Bilinear b0, b1;
    F3 t;
    if (roughness != nullptr) {
        const MipLevel m = rough_to_level(roughness[i], p.min_roughness, p.max_roughness, p.num_levels);
        cube_bilinear<false>(p.levels[m.l0], p.base_res >> m.l0, f.face, f.u, f.v, b0);
        t = b0.val;
        if (m.l1 != m.l0) {
            cube_bilinear<false>(p.levels[m.l1], p.base_res >> m.l1, f.face, f.u, f.v, b1);
            t = (1.0f - m.f) * b0.val + m.f * b1.val;
        }
    } else {
        cube_bilinear<false>(p.levels[0], p.base_res, f.face, f.u, f.v, b0);
        t = b0.val;
    }
    out[3 * i] = sigmoidf(t.x);
    out[3 * i + 1] = sigmoidf(t.y);
    out[3 * i + 2] = sigmoidf(t.z);
}

// Backward of envlight_query_kernel: texel gradients (float4 per texel, like shade_bwd), and optionally the
// gradients of the directions and of the roughness. SMALL: the chain is ONE level of <= kSmallTexels texels (the
// 6x16x16 diffuse map every surfel of render_volume samples): gradients are summed in shared memory and each CTA
// adds its tile to global memory once, instead of n x 4 atomics on the same 1536 addresses.
constexpr int kSmallTexels = 6 * 16 * 16;

template <bool SMALL>
__global__ void __launch_bounds__(256)
envlight_query_bwd_kernel(const MrgsShadeArgs p, long long n, const float* __restrict__ dirs,
                          const float* __restrict__ roughness, const float* __restrict__ dL_dout,
                          float* __restrict__ dL_ddirs, float* __restrict__ dL_droughness) {
    __shared__ float s_acc[SMALL ? kSmallTexels * 3 : 1];
    const int texels0 = 6 * p.base_res * p.base_res;
    if (SMALL) {
        for (int k = threadIdx.x; k < texels0 * 3; k += blockDim.x) s_acc[k] = 0.f;
        __syncthreads();
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const F3 d = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
        const FaceUV f = dir_to_face(d);
        Bilinear b0, b1;
        MipLevel m;
        m.l0 = m.l1 = 0;
        m.f = 0.f;
        m.dlevel_drough = 0.f;
        if (roughness != nullptr) m = rough_to_level(roughness[i], p.min_roughness, p.max_roughness, p.num_levels);
        cube_bilinear<true>(p.levels[m.l0], p.base_res >> m.l0, f.face, f.u, f.v, b0);
        F3 t = b0.val;
        const bool two = m.l1 != m.l0;
        if (two) {
            cube_bilinear<true>(p.levels[m.l1], p.base_res >> m.l1, f.face, f.u, f.v, b1);
            t = (1.0f - m.f) * b0.val + m.f * b1.val;
        }
        const F3 o = {sigmoidf(t.x), sigmoidf(t.y), sigmoidf(t.z)};
        const F3 g = {dL_dout[3 * i], dL_dout[3 * i + 1], dL_dout[3 * i + 2]};
        const F3 g_t = {g.x * o.x * (1.0f - o.x), g.y * o.y * (1.0f - o.y), g.z * o.z * (1.0f - o.z)};
        float g_u, g_v;
        if (two) {
            if (p.dL_dlevels[m.l0]) scatter_bilinear(p.dL_dlevels[m.l0], b0, (1.0f - m.f) * g_t);
            if (p.dL_dlevels[m.l1]) scatter_bilinear(p.dL_dlevels[m.l1], b1, m.f * g_t);
            if (dL_droughness) dL_droughness[i] = dot(g_t, b1.val - b0.val) * m.dlevel_drough;
            g_u = dot(g_t, (1.0f - m.f) * b0.dval_du + m.f * b1.dval_du);
            g_v = dot(g_t, (1.0f - m.f) * b0.dval_dv + m.f * b1.dval_dv);
        } else {
            if (SMALL) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (b0.idx[k] < 0 || b0.w[k] == 0.f) continue;
                    atomicAdd(&s_acc[3 * b0.idx[k] + 0], b0.w[k] * g_t.x);
                    atomicAdd(&s_acc[3 * b0.idx[k] + 1], b0.w[k] * g_t.y);
                    atomicAdd(&s_acc[3 * b0.idx[k] + 2], b0.w[k] * g_t.z);
                }
            } else if (p.dL_dlevels[m.l0]) {
                scatter_bilinear(p.dL_dlevels[m.l0], b0, g_t);
            }
            if (dL_droughness) dL_droughness[i] = 0.f;
            g_u = dot(g_t, b0.dval_du);
            g_v = dot(g_t, b0.dval_dv);
        }
        if (dL_ddirs) {
            float g_d[3];
            face_uv_backward(f, d, g_u, g_v, g_d);
            dL_ddirs[3 * i] = g_d[0];
            dL_ddirs[3 * i + 1] = g_d[1];
            dL_ddirs[3 * i + 2] = g_d[2];
        }
    }
    if (SMALL) {
        __syncthreads();
        float4* out = reinterpret_cast<float4*>(p.dL_dlevels[0]);
        for (int k = threadIdx.x; k < texels0; k += blockDim.x) {
            const float x = s_acc[3 * k], y = s_acc[3 * k + 1], z = s_acc[3 * k + 2];
            if (x != 0.f || y != 0.f || z != 0.f) atomicAdd(out + k, make_float4(x, y, z, 0.f));
        }
    }
}

// ---- per-surfel split-sum colours of the volume-rendering stage ---------------------------------------------------
// get_full_color_volume / get_full_color_volume_indirect (utils/refl_utils.py:426-490, visibility = 1) as ONE kernel
// pair instead of ~25 eager P-sized torch kernels around two dr.texture calls:
//   w_o = safe_normalize(campos - xyz); NdotV = w_o . n; rr = safe_normalize(2 n NdotV - w_o)
//   diffuse  = sigmoid(tex(diffuse map, n)) * (1 - refl) * albedo
//   direct   = sigmoid(tex(specular chain, rr, roughness))
//   specular = direct * ((0.04 (1 - refl) + albedo refl) * fg.x + fg.y)
// `fg` is ONE pair for all surfels: the reference indexes fg[0] on the [N,2] LUT result (:445, :481), i.e. the first
// surfel's pair; the host mirror evaluates it and gets its gradient back as the 2-float sum `dL_dfg`.
struct SurfelShade {
    F3 n, wo, r, rr, albedo, dl, Ld, sw;
    float rs, ro, ndv, rl, vl;
    MipLevel mip;
    FaceUV fn, fr;
    Bilinear bn, b0, b1;
};

template <bool GRAD>
__device__ __forceinline__ void shade_one_surfel(const MrgsSurfelShadeArgs& p, long long i, float fgx, float fgy,
                                                 SurfelShade& s) {
    const F3 x = {p.xyz[3 * i], p.xyz[3 * i + 1], p.xyz[3 * i + 2]};
    s.n = {p.normals[3 * i], p.normals[3 * i + 1], p.normals[3 * i + 2]};
    s.albedo = {p.albedo[3 * i], p.albedo[3 * i + 1], p.albedo[3 * i + 2]};
    s.rs = p.refl_strength[i];
    s.ro = p.roughness[i];
    const F3 v = {p.campos[0] - x.x, p.campos[1] - x.y, p.campos[2] - x.z};
    s.vl = fmaxf(sqrtf(dot(v, v)), 1e-20f);
    s.wo = (1.0f / s.vl) * v;
    s.ndv = dot(s.wo, s.n);
    s.r = (2.0f * s.ndv) * s.n - s.wo;
    s.rl = fmaxf(sqrtf(dot(s.r, s.r)), 1e-20f);
    s.rr = (1.0f / s.rl) * s.r;
    // diffuse: plain bilinear fetch of the cosine-convolved map in the direction of the normal
    s.fn = dir_to_face(s.n);
    cube_bilinear<GRAD>(p.diffuse_map, p.diffuse_res, s.fn.face, s.fn.u, s.fn.v, s.bn);
    s.dl = {sigmoidf(s.bn.val.x), sigmoidf(s.bn.val.y), sigmoidf(s.bn.val.z)};
    // specular: trilinear fetch of the GGX chain in the reflected direction
    s.mip = rough_to_level(s.ro, p.chain.min_roughness, p.chain.max_roughness, p.chain.num_levels);
    s.fr = dir_to_face(s.rr);
    cube_bilinear<GRAD>(p.chain.levels[s.mip.l0], p.chain.base_res >> s.mip.l0, s.fr.face, s.fr.u, s.fr.v, s.b0);
    F3 t = s.b0.val;
    if (s.mip.l1 != s.mip.l0) {
        cube_bilinear<GRAD>(p.chain.levels[s.mip.l1], p.chain.base_res >> s.mip.l1, s.fr.face, s.fr.u, s.fr.v, s.b1);
        t = (1.0f - s.mip.f) * s.b0.val + s.mip.f * s.b1.val;
    }
    s.Ld = {sigmoidf(t.x), sigmoidf(t.y), sigmoidf(t.z)};
    const float k0 = 0.04f * (1.0f - s.rs);
    s.sw = {(k0 + s.albedo.x * s.rs) * fgx + fgy, (k0 + s.albedo.y * s.rs) * fgx + fgy,
            (k0 + s.albedo.z * s.rs) * fgx + fgy};
}

__device__ __forceinline__ void store3i(float* out, long long i, F3 v) {
    if (out == nullptr) return;
    out[3 * i] = v.x;
    out[3 * i + 1] = v.y;
    out[3 * i + 2] = v.z;
}
__device__ __forceinline__ F3 load3i(const float* in, long long i) {
    if (in == nullptr) return {0.f, 0.f, 0.f};
    return {in[3 * i], in[3 * i + 1], in[3 * i + 2]};
}

__global__ void __launch_bounds__(256) surfel_shade_fwd_kernel(const MrgsSurfelShadeArgs p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.P) return;
    SurfelShade s;
    shade_one_surfel<false>(p, i, p.fg[0], p.fg[1], s);
    const float om = 1.0f - s.rs;
    store3i(p.diffuse, i, {s.dl.x * om * s.albedo.x, s.dl.y * om * s.albedo.y, s.dl.z * om * s.albedo.z});
    store3i(p.specular, i, {s.Ld.x * s.sw.x, s.Ld.y * s.sw.y, s.Ld.z * s.sw.z});
    store3i(p.direct_light, i, s.Ld);
}

// The 6 x 16^2 diffuse map is sampled by every surfel: its texel gradients are summed in shared memory and flushed
// once per CTA (grid-stride loop over few, fat CTAs), like envlight_query_bwd_kernel<true>.
template <bool SMALL>
__global__ void __launch_bounds__(256) surfel_shade_bwd_kernel(const MrgsSurfelShadeArgs p) {
    __shared__ float s_acc[SMALL ? kSmallTexels * 3 : 1];
    __shared__ float s_fg[2][8];
    const int texels_d = 6 * p.diffuse_res * p.diffuse_res;
    if (SMALL) {
        for (int k = threadIdx.x; k < texels_d * 3; k += blockDim.x) s_acc[k] = 0.f;
        __syncthreads();
    }
    const float fgx = p.fg[0], fgy = p.fg[1];
    float acc_fgx = 0.f, acc_fgy = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.P; i += (long long)gridDim.x * blockDim.x) {
        SurfelShade s;
        shade_one_surfel<true>(p, i, fgx, fgy, s);
        const F3 g_diff = load3i(p.dL_ddiffuse, i), g_spec = load3i(p.dL_dspecular, i), g_dir = load3i(p.dL_ddirect, i);
        const float om = 1.0f - s.rs;
        // diffuse = dl * (1 - rs) * albedo
        const F3 g_dl = {g_diff.x * om * s.albedo.x, g_diff.y * om * s.albedo.y, g_diff.z * om * s.albedo.z};
        float g_rs = -(g_diff.x * s.dl.x * s.albedo.x + g_diff.y * s.dl.y * s.albedo.y + g_diff.z * s.dl.z * s.albedo.z);
        F3 g_albedo = {g_diff.x * s.dl.x * om, g_diff.y * s.dl.y * om, g_diff.z * s.dl.z * om};
        // specular = Ld * sw, direct = Ld
        const F3 g_Ld = {g_spec.x * s.sw.x + g_dir.x, g_spec.y * s.sw.y + g_dir.y, g_spec.z * s.sw.z + g_dir.z};
        const F3 g_sw = {g_spec.x * s.Ld.x, g_spec.y * s.Ld.y, g_spec.z * s.Ld.z};
        g_rs += fgx * (g_sw.x * (s.albedo.x - 0.04f) + g_sw.y * (s.albedo.y - 0.04f) + g_sw.z * (s.albedo.z - 0.04f));
        g_albedo = g_albedo + (s.rs * fgx) * g_sw;
        const float k0 = 0.04f * om;
        acc_fgx += g_sw.x * (k0 + s.albedo.x * s.rs) + g_sw.y * (k0 + s.albedo.y * s.rs) + g_sw.z * (k0 + s.albedo.z * s.rs);
        acc_fgy += g_sw.x + g_sw.y + g_sw.z;
        // specular chain: sigmoid, mip blend, texels, level, face coordinates
        const F3 g_t = {g_Ld.x * s.Ld.x * (1.0f - s.Ld.x), g_Ld.y * s.Ld.y * (1.0f - s.Ld.y), g_Ld.z * s.Ld.z * (1.0f - s.Ld.z)};
        float g_ro = 0.f, g_u, g_v;
        if (s.mip.l1 != s.mip.l0) {
            const float fm = s.mip.f;
            if (p.chain.dL_dlevels[s.mip.l0]) scatter_bilinear(p.chain.dL_dlevels[s.mip.l0], s.b0, (1.0f - fm) * g_t);
            if (p.chain.dL_dlevels[s.mip.l1]) scatter_bilinear(p.chain.dL_dlevels[s.mip.l1], s.b1, fm * g_t);
            g_ro = dot(g_t, s.b1.val - s.b0.val) * s.mip.dlevel_drough;
            g_u = dot(g_t, (1.0f - fm) * s.b0.dval_du + fm * s.b1.dval_du);
            g_v = dot(g_t, (1.0f - fm) * s.b0.dval_dv + fm * s.b1.dval_dv);
        } else {
            if (p.chain.dL_dlevels[s.mip.l0]) scatter_bilinear(p.chain.dL_dlevels[s.mip.l0], s.b0, g_t);
            g_u = dot(g_t, s.b0.dval_du);
            g_v = dot(g_t, s.b0.dval_dv);
        }
        float g_rr[3];
        face_uv_backward(s.fr, s.rr, g_u, g_v, g_rr);
        F3 g_r;
        {
            const F3 grr = {g_rr[0], g_rr[1], g_rr[2]};
            g_r = (s.rl > 1e-20f) ? (1.0f / s.rl) * (grr - dot(s.rr, grr) * s.rr) : (1.0f / s.rl) * grr;
        }
        // r = 2 n (n.wo) - wo ; ndv = n.wo
        F3 g_n = (2.0f * s.ndv) * g_r;
        const float g_ndv = 2.0f * dot(s.n, g_r);
        g_n = g_n + g_ndv * s.wo;
        F3 g_wo = g_ndv * s.n - g_r;
        // diffuse map: sigmoid, texels, face coordinates of the normal
        const F3 g_td = {g_dl.x * s.dl.x * (1.0f - s.dl.x), g_dl.y * s.dl.y * (1.0f - s.dl.y), g_dl.z * s.dl.z * (1.0f - s.dl.z)};
        if (SMALL) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (s.bn.idx[k] < 0 || s.bn.w[k] == 0.f) continue;
                atomicAdd(&s_acc[3 * s.bn.idx[k] + 0], s.bn.w[k] * g_td.x);
                atomicAdd(&s_acc[3 * s.bn.idx[k] + 1], s.bn.w[k] * g_td.y);
                atomicAdd(&s_acc[3 * s.bn.idx[k] + 2], s.bn.w[k] * g_td.z);
            }
        } else if (p.dL_ddiffuse_map) {
            scatter_bilinear(p.dL_ddiffuse_map, s.bn, g_td);
        }
        float g_nd[3];
        face_uv_backward(s.fn, s.n, dot(g_td, s.bn.dval_du), dot(g_td, s.bn.dval_dv), g_nd);
        g_n = g_n + F3{g_nd[0], g_nd[1], g_nd[2]};
        // wo = v / max(|v|, eps), v = campos - xyz
        const F3 g_vv = (s.vl > 1e-20f) ? (1.0f / s.vl) * (g_wo - dot(s.wo, g_wo) * s.wo) : (1.0f / s.vl) * g_wo;
        store3i(p.dL_dxyz, i, {-g_vv.x, -g_vv.y, -g_vv.z});
        store3i(p.dL_dnormals, i, g_n);
        store3i(p.dL_dalbedo, i, g_albedo);
        if (p.dL_drefl_strength) p.dL_drefl_strength[i] = g_rs;
        if (p.dL_droughness) p.dL_droughness[i] = g_ro;
    }
    // the shared FG pair: block sums, one atomic pair per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_fgx += __shfl_xor_sync(0xffffffffu, acc_fgx, o);
        acc_fgy += __shfl_xor_sync(0xffffffffu, acc_fgy, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_fg[0][threadIdx.x >> 5] = acc_fgx;
        s_fg[1][threadIdx.x >> 5] = acc_fgy;
    }
    __syncthreads();
    if (threadIdx.x == 0 && p.dL_dfg) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += s_fg[0][w]; b += s_fg[1][w]; }
        atomicAdd(p.dL_dfg, a);
        atomicAdd(p.dL_dfg + 1, b);
    }
    if (SMALL) {
        float4* out = reinterpret_cast<float4*>(p.dL_ddiffuse_map);
        for (int k = threadIdx.x; k < texels_d; k += blockDim.x) {
            const float x = s_acc[3 * k], y = s_acc[3 * k + 1], z = s_acc[3 * k + 2];
            if (x != 0.f || y != 0.f || z != 0.f) atomicAdd(out + k, make_float4(x, y, z, 0.f));
        }
    }
}

int validate(const MrgsShadeArgs* a, const char* who, bool need_maps, int min_levels = 2) {
    if (a == nullptr) {
        set_error("%s: null args", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->num_levels < min_levels || a->num_levels > MRGS_MAX_MIP_LEVELS || a->base_res <= 0 ||
        (a->base_res >> (a->num_levels - 1)) < 1) {
        set_error("%s: bad mip chain (levels=%d, base_res=%d)", who, a->num_levels, a->base_res);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    for (int l = 0; l < a->num_levels; ++l)
        if (a->levels[l] == nullptr) {
            set_error("%s: level %d is null", who, l);
            return MRGS_ERR_INVALID_ARGUMENT;
        }
    if (need_maps && (a->width <= 0 || a->height <= 0 || !a->base_color || !a->features || !a->allmap || !a->lut)) {
        set_error("%s: missing G-buffer / LUT pointers or bad size %dx%d", who, a->width, a->height);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    return MRGS_OK;
}

}  // namespace

int launch_shade(const MrgsShadeArgs* a, bool backward, cudaStream_t stream) {
    int st = validate(a, backward ? "mrgs_shade_backward" : "mrgs_shade_forward", true);
    if (st != MRGS_OK) return st;
    const dim3 grid((a->width + 31) / 32, (a->height + 7) / 8);
    if (backward)
        shade_bwd_kernel<<<grid, 256, 0, stream>>>(*a);
    else
        shade_fwd_kernel<<<grid, 256, 0, stream>>>(*a);
    return MRGS_OK;
}

int launch_envlight_query(const MrgsShadeArgs* a, long long n, const float* dirs, const float* roughness,
                          float* out, cudaStream_t stream) {
    int st = validate(a, "mrgs_envlight_query", false, roughness ? 2 : 1);
    if (st != MRGS_OK) return st;
    if (n <= 0) return MRGS_OK;
    envlight_query_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(*a, n, dirs, roughness, out);
    return MRGS_OK;
}

int launch_envlight_query_bwd(const MrgsShadeArgs* a, long long n, const float* dirs, const float* roughness,
                              const float* dL_dout, float* dL_ddirs, float* dL_droughness, cudaStream_t stream) {
    int st = validate(a, "mrgs_envlight_query_backward", false, roughness ? 2 : 1);
    if (st != MRGS_OK) return st;
    if (n <= 0) return MRGS_OK;
    const bool small = roughness == nullptr && a->dL_dlevels[0] != nullptr && 6 * a->base_res * a->base_res <= kSmallTexels;
    if (small) {
        // few, fat CTAs: every CTA flushes its shared-memory tile once
        const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 148 * 4);
        envlight_query_bwd_kernel<true><<<blocks, 256, 0, stream>>>(*a, n, dirs, roughness, dL_dout, dL_ddirs, dL_droughness);
    } else {
        envlight_query_bwd_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(*a, n, dirs, roughness, dL_dout,
                                                                                        dL_ddirs, dL_droughness);
    }
    return MRGS_OK;
}

}  // namespace mrgs

// ---- pseudo surface depth + depth_to_normal --------------------------
// compute_2dgs_normal_and_regularizations gaussian_renderer/__init__.py:50-78 and depth_to_normal /
// depths_to_points utils/point_utils.py:9-37, fused: one thread per pixel evaluates the surface depth of
// itself and of its 4 neighbours straight from the rasterizer's allmap (no intermediate maps), builds the
// world points depth * (A (x,y,1)) + o and the normal of the centred differences.
namespace mrgs {
namespace {

struct DepthNormalParams {
    int W, H;
    float depth_ratio;
    float A[9];   // row-major: ray = A * (x, y, 1)
    float o[3];   // camera centre
    const float* allmap;
    float* surf_depth;        // [1,H,W]
    float* surf_normal;       // [3,H,W]
    const float* g_depth;     // backward inputs (may be null)
    const float* g_normal;
    float* g_allmap;          // planes 0 and 5 written, plane 1 accumulated (+=)
};

__device__ __forceinline__ float nan_to_zero(float v) { return (isnan(v) || isinf(v)) ? 0.0f : v; }

__device__ __forceinline__ float surf_depth_at(const DepthNormalParams& p, int x, int y) {
    const size_t HW = (size_t)p.W * p.H, pix = (size_t)y * p.W + x;
    const float expected = nan_to_zero(p.allmap[kDepthOff * HW + pix] / p.allmap[kAlphaOff * HW + pix]);
    const float median = nan_to_zero(p.allmap[kMidDepthOff * HW + pix]);
    return expected * (1.0f - p.depth_ratio) + p.depth_ratio * median;
}

__device__ __forceinline__ F3 ray_at(const DepthNormalParams& p, int x, int y) {
    const float fx = (float)x, fy = (float)y;
    return {p.A[0] * fx + p.A[1] * fy + p.A[2], p.A[3] * fx + p.A[4] * fy + p.A[5], p.A[6] * fx + p.A[7] * fy + p.A[8]};
}

__device__ __forceinline__ F3 cross3(F3 a, F3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// normal at interior centre (x,y) from the four neighbour depths; optionally its vjp w.r.t. them
struct CentreNormal {
    F3 n;            // normalised
    F3 rd, ru, rl, rr;  // rays of the down(y+1) / up(y-1) / left(x-1) / right(x+1) neighbours
    F3 dx, dy, c;
    float len;
};
__device__ __forceinline__ CentreNormal centre_normal(const DepthNormalParams& p, int x, int y, float d_dn,
                                                      float d_up, float d_lf, float d_rt) {
    CentreNormal r;
    r.rd = ray_at(p, x, y + 1);
    r.ru = ray_at(p, x, y - 1);
    r.rl = ray_at(p, x - 1, y);
    r.rr = ray_at(p, x + 1, y);
    r.dx = d_dn * r.rd - d_up * r.ru;   // points[y+1] - points[y-1] (the origin cancels)
    r.dy = d_rt * r.rr - d_lf * r.rl;   // points[x+1] - points[x-1]
    r.c = cross3(r.dx, r.dy);
    r.len = fmaxf(sqrtf(dot(r.c, r.c)), 1e-12f);  // F.normalize eps
    r.n = (1.0f / r.len) * r.c;
    return r;
}

__global__ void __launch_bounds__(256) depth_normal_fwd_kernel(const DepthNormalParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.W || y >= p.H) return;
    const size_t HW = (size_t)p.W * p.H, pix = (size_t)y * p.W + x;
    p.surf_depth[pix] = surf_depth_at(p, x, y);
    F3 n = {0.f, 0.f, 0.f};
    if (x >= 1 && y >= 1 && x < p.W - 1 && y < p.H - 1) {
        const CentreNormal c = centre_normal(p, x, y, surf_depth_at(p, x, y + 1), surf_depth_at(p, x, y - 1),
                                             surf_depth_at(p, x - 1, y), surf_depth_at(p, x + 1, y));
        const float a = p.allmap[kAlphaOff * HW + pix];  // alpha.detach()
        n = a * c.n;
    }
    p.surf_normal[pix] = n.x;
    p.surf_normal[HW + pix] = n.y;
    p.surf_normal[2 * HW + pix] = n.z;
}

// gradient of the loss w.r.t. the surface depth of pixel (x,y) through the normal of centre (cx,cy)
__device__ __forceinline__ float depth_grad_via_centre(const DepthNormalParams& p, int cx, int cy, int x, int y) {
    if (cx < 1 || cy < 1 || cx >= p.W - 1 || cy >= p.H - 1) return 0.0f;
    const size_t HW = (size_t)p.W * p.H, cpix = (size_t)cy * p.W + cx;
    const float a = p.allmap[kAlphaOff * HW + cpix];
    const F3 g = {a * p.g_normal[cpix], a * p.g_normal[HW + cpix], a * p.g_normal[2 * HW + cpix]};
    if (g.x == 0.f && g.y == 0.f && g.z == 0.f) return 0.0f;
    const CentreNormal c = centre_normal(p, cx, cy, surf_depth_at(p, cx, cy + 1), surf_depth_at(p, cx, cy - 1),
                                         surf_depth_at(p, cx - 1, cy), surf_depth_at(p, cx + 1, cy));
    // n = c / max(|c|, eps)
    F3 gc;
    if (sqrtf(dot(c.c, c.c)) > 1e-12f)
        gc = (1.0f / c.len) * (g - dot(c.n, g) * c.n);
    else
        gc = (1.0f / c.len) * g;
    // c = dx x dy:  d/d(dx) = dy x gc,  d/d(dy) = gc x dx
    const F3 g_dx = cross3(c.dy, gc), g_dy = cross3(gc, c.dx);
    if (x == cx && y == cy + 1) return dot(g_dx, c.rd);
    if (x == cx && y == cy - 1) return -dot(g_dx, c.ru);
    if (x == cx + 1 && y == cy) return dot(g_dy, c.rr);
    return -dot(g_dy, c.rl);
}

__global__ void __launch_bounds__(256) depth_normal_bwd_kernel(const DepthNormalParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.W || y >= p.H) return;
    const size_t HW = (size_t)p.W * p.H, pix = (size_t)y * p.W + x;
    float g = p.g_depth ? p.g_depth[pix] : 0.0f;
    if (p.g_normal) {
        g += depth_grad_via_centre(p, x, y - 1, x, y);  // this pixel is the "down" neighbour of (x, y-1)
        g += depth_grad_via_centre(p, x, y + 1, x, y);
        g += depth_grad_via_centre(p, x - 1, y, x, y);
        g += depth_grad_via_centre(p, x + 1, y, x, y);
    }
    // surf_depth = (1-ratio) * nan_to_num(D / alpha) + ratio * nan_to_num(median)
    const float D = p.allmap[kDepthOff * HW + pix], a = p.allmap[kAlphaOff * HW + pix];
    const float q = D / a;
    const bool q_ok = !(isnan(q) || isinf(q));
    const float ge = q_ok ? g * (1.0f - p.depth_ratio) : 0.0f;
    const float med = p.allmap[kMidDepthOff * HW + pix];
    // alpha == 0 (no contributor): the reference's autograd produces NaN here (0/0); nothing reads it, we write 0
    p.g_allmap[kDepthOff * HW + pix] = q_ok ? ge / a : 0.0f;
    p.g_allmap[kAlphaOff * HW + pix] += q_ok ? -ge * q / a : 0.0f;
    p.g_allmap[kMidDepthOff * HW + pix] = (isnan(med) || isinf(med)) ? 0.0f : g * p.depth_ratio;
}

}  // namespace

int launch_depth_normal(bool backward, int W, int H, float depth_ratio, const float* A, const float* o,
                        const float* allmap, float* surf_depth, float* surf_normal, const float* g_depth,
                        const float* g_normal, float* g_allmap, cudaStream_t stream) {
    if (W <= 0 || H <= 0 || allmap == nullptr || (!backward && (!surf_depth || !surf_normal)) ||
        (backward && !g_allmap)) {
        set_error("mrgs_depth_normal: bad arguments");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    DepthNormalParams p;
    p.W = W; p.H = H; p.depth_ratio = depth_ratio;
    for (int i = 0; i < 9; ++i) p.A[i] = A[i];
    for (int i = 0; i < 3; ++i) p.o[i] = o[i];
    p.allmap = allmap; p.surf_depth = surf_depth; p.surf_normal = surf_normal;
    p.g_depth = g_depth; p.g_normal = g_normal; p.g_allmap = g_allmap;
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    if (backward)
        depth_normal_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    else
        depth_normal_fwd_kernel<<<grid, 256, 0, stream>>>(p);
    return MRGS_OK;
}

int launch_surfel_shade(const MrgsSurfelShadeArgs* a, bool backward, cudaStream_t stream) {
    const char* who = backward ? "mrgs_surfel_shade_backward" : "mrgs_surfel_shade_forward";
    if (a == nullptr) {
        set_error("%s: null args", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    int st = validate(&a->chain, who, false, 2);
    if (st != MRGS_OK) return st;
    if (a->P < 0 || a->diffuse_res <= 0 || !a->diffuse_map || !a->fg || !a->campos ||
        (a->P > 0 && (!a->xyz || !a->normals || !a->albedo || !a->refl_strength || !a->roughness))) {
        set_error("%s: missing per-surfel inputs, diffuse map, fg pair or campos (P=%d)", who, a->P);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->P == 0) return MRGS_OK;
    if (!backward) {
        surfel_shade_fwd_kernel<<<(a->P + 255) / 256, 256, 0, stream>>>(*a);
        return MRGS_OK;
    }
    const unsigned blocks = (unsigned)std::min<long long>(((long long)a->P + 255) / 256, 148 * 4);
    if (a->dL_ddiffuse_map != nullptr && 6 * a->diffuse_res * a->diffuse_res <= kSmallTexels)
        surfel_shade_bwd_kernel<true><<<blocks, 256, 0, stream>>>(*a);
    else
        surfel_shade_bwd_kernel<false><<<blocks, 256, 0, stream>>>(*a);
    return MRGS_OK;
}

}  // namespace mrgs
