// preprocess.cu — per-surfel forward stage: near cull, 2DGS ray-splat transform T, view-space
// normal, screen-space AABB / radius / tile rectangle, SH -> RGB, and packing of the two
// per-surfel records the blend kernels gather from.
//
// Behavioural reference: preprocessCUDA rast/cuda_rasterizer/forward.cu:163-266 with
// in_frustum auxiliary.h:192-217 and computeColorFromSH forward.cu:22-73; checkFrustum
// rasterizer_impl.cu:56-68. The data layout (one 64-byte geometry record + one padded
// colour/feature record per surfel) is this library's own.
#include "kernels.cuh"
#include "splat_math.cuh"

namespace mrgs {

namespace {

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2_0 = 1.0925484305920792f;
constexpr float SH_C2_1 = -1.0925484305920792f;
constexpr float SH_C2_2 = 0.31539156525252005f;
constexpr float SH_C2_3 = -1.0925484305920792f;
constexpr float SH_C2_4 = 0.5462742152960396f;
constexpr float SH_C3_0 = -0.5900435899266435f;
constexpr float SH_C3_1 = 2.890611442640554f;
constexpr float SH_C3_2 = -0.4570457994644658f;
constexpr float SH_C3_3 = 0.3731763325901154f;
constexpr float SH_C3_4 = -0.4570457994644658f;
constexpr float SH_C3_5 = 1.445305721320277f;
constexpr float SH_C3_6 = -0.5900435899266435f;

// View-dependent colour of one surfel. `sh` points at this surfel's M coefficient triples.
// Every term is folded into the running sum with one fma, the order the reference binary uses.
__device__ __forceinline__ void sh_to_rgb(int deg, const float* __restrict__ sh, Vec3 pos,
                                          Vec3 cam, float rgb[3], unsigned& clamped_bits) {
    const float dx0 = pos.x - cam.x, dy0 = pos.y - cam.y, dz0 = pos.z - cam.z;
    const float len = sqrtf(__fmaf_rn(dz0, dz0, __fmaf_rn(dx0, dx0, __fmul_rn(dy0, dy0))));
    const float x = __fdiv_rn(dx0, len), y = __fdiv_rn(dy0, len), z = __fdiv_rn(dz0, len);

    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = __fmul_rn(SH_C0, sh[c]);
    if (deg > 0) {
        const float c1y = __fmul_rn(SH_C1, y), c1z = __fmul_rn(SH_C1, z), c1x = __fmul_rn(SH_C1, x);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            r[c] = __fmaf_rn(-c1y, sh[3 + c], r[c]);
            r[c] = __fmaf_rn(c1z, sh[6 + c], r[c]);
            r[c] = __fmaf_rn(-c1x, sh[9 + c], r[c]);
        }
        if (deg > 1) {
            const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
            const float k4 = __fmul_rn(SH_C2_0, xy);
            const float k5 = __fmul_rn(SH_C2_1, yz);
            const float k6 = __fmul_rn(SH_C2_2, __fadd_rn(__fadd_rn(__fadd_rn(zz, zz), -xx), -yy));
            const float k7 = __fmul_rn(SH_C2_3, xz);
            const float k8 = __fmul_rn(SH_C2_4, __fadd_rn(xx, -yy));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[c] = __fmaf_rn(k4, sh[12 + c], r[c]);
                r[c] = __fmaf_rn(k5, sh[15 + c], r[c]);
                r[c] = __fmaf_rn(k6, sh[18 + c], r[c]);
                r[c] = __fmaf_rn(k7, sh[21 + c], r[c]);
                r[c] = __fmaf_rn(k8, sh[24 + c], r[c]);
            }
            if (deg > 2) {
                const float k9 = __fmul_rn(__fmul_rn(SH_C3_0, y), __fmaf_rn(3.0f, xx, -yy));
                const float k10 = __fmul_rn(__fmul_rn(SH_C3_1, xy), z);
                const float f4zz = __fadd_rn(__fmaf_rn(4.0f, zz, -xx), -yy);
                const float k11 = __fmul_rn(__fmul_rn(SH_C3_2, y), f4zz);
                const float k12 = __fmul_rn(
                    __fmul_rn(SH_C3_3, z),
                    __fmaf_rn(-3.0f, yy, __fmaf_rn(-3.0f, xx, __fadd_rn(zz, zz))));
                const float k13 = __fmul_rn(__fmul_rn(SH_C3_4, x), f4zz);
                const float k14 = __fmul_rn(__fmul_rn(SH_C3_5, z), __fadd_rn(xx, -yy));
                const float k15 = __fmul_rn(__fmul_rn(SH_C3_6, x), __fmaf_rn(-3.0f, yy, xx));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    r[c] = __fmaf_rn(k9, sh[27 + c], r[c]);
                    r[c] = __fmaf_rn(k10, sh[30 + c], r[c]);
                    r[c] = __fmaf_rn(k11, sh[33 + c], r[c]);
                    r[c] = __fmaf_rn(k12, sh[36 + c], r[c]);
                    r[c] = __fmaf_rn(k13, sh[39 + c], r[c]);
                    r[c] = __fmaf_rn(k14, sh[42 + c], r[c]);
                    r[c] = __fmaf_rn(k15, sh[45 + c], r[c]);
                }
            }
        }
    }
    clamped_bits = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r[c] = __fadd_rn(r[c], 0.5f);
        if (r[c] < 0.0f) clamped_bits |= (1u << c);
        rgb[c] = fmaxf(r[c], 0.0f);
    }
}

template <int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) preprocess_fwd_kernel(PreprocessParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    p.radii[idx] = 0;
    p.tiles_touched[idx] = 0;
    p.sort_key[idx] = 0xffffffffu;

    const Vec3 pos = {p.means3D[3 * idx], p.means3D[3 * idx + 1], p.means3D[3 * idx + 2]};
    const Vec3 pv = view_point(p.viewmatrix, pos);
    if (pv.z <= kNear) {
        if (p.prefiltered) __trap();
        return;
    }

    float T[9];
    Vec3 normal;
    if (p.transMat_precomp == nullptr) {
        const float2 sc = reinterpret_cast<const float2*>(p.scales)[idx];
        const float4 q = reinterpret_cast<const float4*>(p.rotations)[idx];
        surfel_transmat(pos, sc.x, sc.y, p.scale_modifier, q.x, q.y, q.z, q.w, p.projmatrix,
                        p.viewmatrix, p.W, p.H, T, normal);
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) T[i] = p.transMat_precomp[9 * idx + i];
        normal = {0.0f, 0.0f, 1.0f};
    }

    // dual-visible surfels: orient the normal towards the camera (forward.cu:224-229)
    const float d = __fmaf_rn(pv.z, normal.z, __fmaf_rn(pv.x, normal.x, __fmul_rn(pv.y, normal.y)));
    if (d == 0.0f) return;
    const float flip = (d < 0.0f) ? 1.0f : -1.0f;
    normal.x *= flip;
    normal.y *= flip;
    normal.z *= flip;

    float cx, cy, ex, ey;
    if (!surfel_aabb(T, cx, cy, ex, ey)) return;
    const int radius = (int)ceilf(fmaxf(ex, ey));

    int x0, y0, x1, y1;
    tile_rect(cx, cy, radius, p.grid_x, p.grid_y, x0, y0, x1, y1);
    const unsigned touched = (unsigned)(x1 - x0) * (unsigned)(y1 - y0);
    if (touched == 0) return;

    // colour / feature record
    float* cf = p.cf + (size_t)idx * p.cf_stride;
    unsigned clamped_bits = 0;
    float rgb[3];
    if (p.colors_precomp == nullptr) {
        const Vec3 cam = {p.campos[0], p.campos[1], p.campos[2]};
        if (p.M == 16) {  // 192-byte rows as 12 float4 loads, coefficients stay in registers
            float sh_l[48];
            const float4* src = reinterpret_cast<const float4*>(p.shs + (size_t)idx * 48);
            const int nq = (p.D >= 3) ? 12 : (p.D == 2) ? 7 : (p.D == 1) ? 3 : 1;
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < nq) v = __ldg(src + q);
                sh_l[4 * q] = v.x; sh_l[4 * q + 1] = v.y; sh_l[4 * q + 2] = v.z; sh_l[4 * q + 3] = v.w;
            }
            sh_to_rgb(p.D, sh_l, pos, cam, rgb, clamped_bits);
        } else {
            sh_to_rgb(p.D, p.shs + (size_t)idx * p.M * 3, pos, cam, rgb, clamped_bits);
        }
    } else {
        rgb[0] = p.colors_precomp[3 * idx];
        rgb[1] = p.colors_precomp[3 * idx + 1];
        rgb[2] = p.colors_precomp[3 * idx + 2];
    }
    {
        // rgb | features | zero padding, written as float4s
        const float* f = p.features + (size_t)idx * p.S;
        for (int base = 0; base < p.cf_stride; base += 4) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = base + k;
                v[k] = (c < 3) ? rgb[c] : ((c - 3 < p.S) ? f[c - 3] : 0.0f);
            }
            reinterpret_cast<float4*>(cf)[base >> 2] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    p.clamped[idx] = (uint8_t)clamped_bits;

    const float opacity = p.opacities[idx];
    float tau;
    float4 bd;
    const float4 bb = alpha_support_bounds(T, cx, cy, opacity, tau, bd);

    float4* rec = reinterpret_cast<float4*>(p.rec + (size_t)idx * kGeomFloats);
    rec[0] = make_float4(T[0], T[1], T[2], T[6]);
    rec[1] = make_float4(T[3], T[4], T[5], T[7]);
    rec[2] = make_float4(T[8], cx, cy, opacity);
    rec[3] = make_float4(normal.x, normal.y, normal.z, tau);
    p.depth[idx] = pv.z;
    p.sort_key[idx] = __float_as_uint(pv.z);
    p.bbox[2 * idx] = bb;       // one 32-byte sector per surfel: box, then the diagonal slabs
    p.bbox[2 * idx + 1] = bd;

    p.rect[idx] = make_uint2((unsigned)x0 | ((unsigned)y0 << 16), (unsigned)x1 | ((unsigned)y1 << 16));
    p.radii[idx] = radius;
    p.tiles_touched[idx] = touched;
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const Vec3 pos = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    present[idx] = view_depth(viewmatrix, pos) > kNear ? 1 : 0;
}

}  // namespace

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream) {
    // 128-thread CTAs at 6 per SM (80 registers): 24 resident warps instead of 16 for this HBM-bound kernel.
    // Measured at C3: 256 threads 0.110 ms, 128x4 / 128x5 0.096, 128x6 0.093, 128x8 (64 registers, spills) 0.103.
    preprocess_fwd_kernel<128, 6><<<(p.P + 127) / 128, 128, 0, stream>>>(p);
}

void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                         cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
}

}  // namespace mrgs
