// render_fwd.cu — tile-blended forward: colour, S material channels, depth, alpha, normal,
// median depth and distortion per pixel, plus the per-pixel state the backward needs.
//
// Behavioural reference: renderCUDA rast/cuda_rasterizer/forward.cu:272-463 (per-pixel
// arithmetic and skip/termination tests are reproduced bit-for-bit through ray_splat()).
// Own design: one CTA per 16x16 tile with each warp owning an 8x4 pixel block; a batch of 256
// instances is staged into shared memory as whole 16-byte vectors of the packed geometry and
// colour/feature records (no per-pair global gathers in the blend loop); per-pixel state is
// kept tile-major so every warp store is one full 128-byte line.
#include "kernels.cuh"
#include "splat_math.cuh"

namespace mrgs {

namespace {

template <int NQ>
__global__ void __launch_bounds__(kTilePixels) render_fwd_kernel(const RenderFwdParams p) {
    __shared__ float4 s_g0[kBatch];
    __shared__ float4 s_g1[kBatch];
    __shared__ float4 s_g2[kBatch];
    __shared__ float4 s_g3[kBatch];
    __shared__ float4 s_cf[NQ][kBatch];

    const int tid = threadIdx.x;
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int px = blockIdx.x * kTileX + slot_x(tid);
    const int py = blockIdx.y * kTileY + slot_y(tid);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = p.ranges[tile];
    const int count = (int)(range.y - range.x);
    const int rounds = (count + kBatch - 1) / kBatch;
    int toDo = count;

    bool done = !inside;
    float T = 1.0f;
    uint32_t contributor = 0, last_contributor = 0, median_contributor = 0;
    float acc[NQ * 4];
#pragma unroll
    for (int c = 0; c < NQ * 4; ++c) acc[c] = 0.0f;
    float N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f, M1 = 0.f, M2 = 0.f, distortion = 0.f;
    float median_depth = 0.f;

    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(p.rec);
    const float4* __restrict__ cf4 = reinterpret_cast<const float4*>(p.cf);

    for (int i = 0; i < rounds; ++i, toDo -= kBatch) {
        if (__syncthreads_count(done) == kTilePixels) break;

        const int progress = i * kBatch + tid;
        if (progress < count) {
            const uint32_t id = p.point_list[range.x + progress];
            const float4* r = rec4 + (size_t)id * (kGeomFloats / 4);
            s_g0[tid] = r[0];
            s_g1[tid] = r[1];
            s_g2[tid] = r[2];
            s_g3[tid] = r[3];
            const float4* c = cf4 + (size_t)id * NQ;
#pragma unroll
            for (int q = 0; q < NQ; ++q) s_cf[q][tid] = c[q];
        }
        __syncthreads();

        const int n = min(kBatch, toDo);
        for (int j = 0; !done && j < n; ++j) {
            ++contributor;
            SplatHit h;
            if (!ray_splat(s_g0[j], s_g1[j], s_g2[j], pxf, pyf, h)) continue;
            const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -h.alpha));
            if (test_T < kTMin) {
                done = true;
                continue;
            }
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
            const float4 g3 = s_g3[j];
            N0 = __fmaf_rn(g3.x, w, N0);
            N1 = __fmaf_rn(g3.y, w, N1);
            N2 = __fmaf_rn(g3.z, w, N2);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 v = s_cf[q][j];
                acc[4 * q + 0] = __fmaf_rn(w, v.x, acc[4 * q + 0]);
                acc[4 * q + 1] = __fmaf_rn(w, v.y, acc[4 * q + 1]);
                acc[4 * q + 2] = __fmaf_rn(w, v.z, acc[4 * q + 2]);
                acc[4 * q + 3] = __fmaf_rn(w, v.w, acc[4 * q + 3]);
            }
            T = test_T;
            last_contributor = contributor;
        }
    }

    // per-pixel state for the backward, tile-major planes: T, M1, M2, n_contrib, median index
    float* st = p.state + (size_t)tile * (5 * kTilePixels) + tid;
    st[0 * kTilePixels] = T;
    st[1 * kTilePixels] = M1;
    st[2 * kTilePixels] = M2;
    reinterpret_cast<uint32_t*>(st)[3 * kTilePixels] = last_contributor;
    reinterpret_cast<uint32_t*>(st)[4 * kTilePixels] = median_contributor;

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
