// render_fwd.cu — tile-blended forward: colour, S material channels, depth, alpha, normal,
// median depth and distortion per pixel, plus the per-pixel state the backward needs.
//
// Behavioural reference: renderCUDA rast/cuda_rasterizer/forward.cu:272-463 (per-pixel
// arithmetic and skip/termination tests are reproduced bit-for-bit through ray_splat()).
//
// Own design — WARP-AUTONOMOUS traversal, no CTA-wide barriers:
//   * one CTA per 16x16 tile, each of its 8 warps owns an 8x4 pixel block and walks the tile's
//     instance list on its own, 32 instances per step (lane <-> instance);
//   * per step every lane fetches its instance id and the surfel's conservative alpha-support box
//     (16 B, computed once per surfel in preprocess) and tests it against the warp's pixel block:
//     the cull costs 1/32 of an instruction stream per (warp, instance) instead of a full ray-splat
//     evaluation by all 32 pixels. Survivors (ballot) have their packed 64-B geometry and
//     colour/feature records staged into the warp's private shared-memory slots as 16-byte
//     vectors, then the 32 pixels blend them in list order with broadcast LDS.128 reads;
//   * the next step's ids/boxes are prefetched while the current survivors are blended;
//   * a warp stops as soon as its 32 pixels are saturated (the reference needs all 256 of the tile);
//   * culling and the in-lane early rejection are conservative by construction (splat_math.cuh), so
//     n_contrib / final_T stay bit-identical to the reference;
//   * per-pixel state is stored tile-major so each warp store is one full 128-byte line.
#include "kernels.cuh"
#include "splat_math.cuh"

namespace mrgs {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;

template <int NQ>
__global__ void __launch_bounds__(kTilePixels) render_fwd_kernel(const RenderFwdParams p) {
    __shared__ float4 s_g[kWarpsPerTile][4][32];
    __shared__ float4 s_cf[kWarpsPerTile][NQ][32];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int px = blockIdx.x * kTileX + slot_x(tid);
    const int py = blockIdx.y * kTileY + slot_y(tid);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    // pixel block of this warp (inclusive bounds, pixel-index coordinates)
    const float bx0 = (float)(blockIdx.x * kTileX + (warp & 1) * 8), bx1 = bx0 + 7.0f;
    const float by0 = (float)(blockIdx.y * kTileY + (warp >> 1) * 4), by1 = by0 + 3.0f;

    const uint2 range = p.ranges[tile];
    const int count = (int)(range.y - range.x);
    const uint32_t* __restrict__ list = p.point_list + range.x;

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0, median_contributor = 0;
    float2 acc2[NQ * 2];  // channel pairs; packed fma.rn.f32x2 keeps each half an IEEE fma (bit-identical)
#pragma unroll
    for (int c = 0; c < NQ * 2; ++c) acc2[c] = make_float2(0.0f, 0.0f);
    float N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f;
    float median_depth = 0.f;

    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(p.rec);
    const float4* __restrict__ cf4 = reinterpret_cast<const float4*>(p.cf);

    // software pipeline: (id, box) of the step after the current one
    uint32_t id_next = 0;
    float4 bb_next = make_float4(0.f, 0.f, -1.f, -1.f);
    if (lane < count) {
        id_next = list[lane];
        bb_next = p.bbox[2 * id_next];
    }

    for (int base = 0; base < count; base += 32) {
        if (__all_sync(kFullMask, done)) break;
        const uint32_t id = id_next;
        const float4 bb = bb_next;
        const int e_next = base + 32 + lane;
        if (e_next < count) {
            id_next = list[e_next];
            bb_next = p.bbox[2 * id_next];
        }
        // (the backward also tests the diagonal slabs stored next to the box; measured here the extra
        // gather costs the forward more than the pairs it removes: 0.458 -> 0.469 ms)
        const bool keep = (base + lane < count) && !(bb.x > bx1 || bb.z < bx0 || bb.y > by1 || bb.w < by0);
        unsigned mask = __ballot_sync(kFullMask, keep);
        if (mask == 0) continue;
        if (keep) {
            const float4* r = rec4 + (size_t)id * (kGeomFloats / 4);
            s_g[warp][0][lane] = r[0];
            s_g[warp][1][lane] = r[1];
            s_g[warp][2][lane] = r[2];
            s_g[warp][3][lane] = r[3];
            const float4* c = cf4 + (size_t)id * NQ;
#pragma unroll
            for (int q = 0; q < NQ; ++q) s_cf[warp][q][lane] = c[q];
        }
        __syncwarp();

        while (mask != 0 && !done) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float4 g3 = s_g[warp][3][j];
            SplatHit h;
            if (!ray_splat(s_g[warp][0][j], s_g[warp][1][j], s_g[warp][2][j], g3.w, pxf, pyf, h)) continue;
            const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -h.alpha));
            if (test_T < kTMin) {
                done = true;
                continue;
            }
            const uint32_t contributor = (uint32_t)(base + j + 1);
            const float w = __fmul_rn(h.alpha, T);

            const float A = __fadd_rn(1.0f, -T);
            const float m = distortion_coord(h.depth);
            const float mm = __fmul_rn(m, m);
            const float err = __fmaf_rn(-M1, __fadd_rn(m, m), __fmaf_rn(A, mm, M2));
            distortion = __fmaf_rn(w, err, distortion);
            D = __fmaf_rn(h.depth, w, D);
            M1 = __fmaf_rn(w, m, M1);
            M2 = __fmaf_rn(w, mm, M2);
            if (T > 0.5f) {
                median_depth = h.depth;
                median_contributor = contributor;
            }
            N0 = __fmaf_rn(g3.x, w, N0);
            N1 = __fmaf_rn(g3.y, w, N1);
            N2 = __fmaf_rn(g3.z, w, N2);
            const float2 w2 = make_float2(w, w);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 v = s_cf[warp][q][j];
                acc2[2 * q + 0] = __ffma2_rn(w2, make_float2(v.x, v.y), acc2[2 * q + 0]);
                acc2[2 * q + 1] = __ffma2_rn(w2, make_float2(v.z, v.w), acc2[2 * q + 1]);
            }
            T = test_T;
            last_contributor = contributor;
        }
        __syncwarp();  // all lanes are past their reads before the slots are overwritten
    }

    // per-pixel state for the backward, tile-major planes: T, M1, M2, n_contrib, median index
    float* st = p.state + (size_t)tile * (5 * kTilePixels) + tid;
    st[0 * kTilePixels] = T;
    st[1 * kTilePixels] = M1;
    st[2 * kTilePixels] = M2;
    reinterpret_cast<uint32_t*>(st)[3 * kTilePixels] = last_contributor;
    reinterpret_cast<uint32_t*>(st)[4 * kTilePixels] = median_contributor;

    float acc[NQ * 4];
#pragma unroll
    for (int c = 0; c < NQ * 2; ++c) {
        acc[2 * c] = acc2[c].x;
        acc[2 * c + 1] = acc2[c].y;
    }
    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.out_color[c * HW + pix] = __fmaf_rn(T, p.background[c], acc[c]);
#pragma unroll
        for (int c = 0; c < NQ * 4 - 3; ++c)
            if (c < p.S) p.out_feature[c * HW + pix] = acc[3 + c];
        p.out_others[kDepthOff * HW + pix] = D;
        p.out_others[kAlphaOff * HW + pix] = __fadd_rn(1.0f, -T);
        p.out_others[(kNormalOff + 0) * HW + pix] = N0;
        p.out_others[(kNormalOff + 1) * HW + pix] = N1;
        p.out_others[(kNormalOff + 2) * HW + pix] = N2;
        p.out_others[kMidDepthOff * HW + pix] = median_depth;
        p.out_others[kDistortionOff * HW + pix] = distortion;
    }
}

}  // namespace

int launch_render_fwd(const RenderFwdParams& p, cudaStream_t stream) {
    const dim3 grid(p.grid_x, p.grid_y);
    switch (p.cf_stride / 4) {
#define MRGS_CASE(NQ)                                                   \
    case NQ:                                                            \
        render_fwd_kernel<NQ><<<grid, kTilePixels, 0, stream>>>(p);     \
        break;
        MRGS_CASE(1)
        MRGS_CASE(2)
        MRGS_CASE(3)
        MRGS_CASE(4)
        MRGS_CASE(5)
        MRGS_CASE(6)
        MRGS_CASE(7)
#undef MRGS_CASE
        default:
            set_error("render_fwd: unsupported feature count S=%d (max %d)", p.S, MRGS_MAX_FEATURES);
            return MRGS_ERR_UNSUPPORTED;
    }
    return MRGS_OK;
}

}  // namespace mrgs
