// geomloss.cu — the geometric regularisers of the reference's calculate_loss (second half) as ONE
// forward and ONE backward kernel, plus get_img_grad_weight.
//
// Behavioural reference (utils/loss_utils.py):
//   :165-172  normal consistency  (image_weight * |surf_normal - rend_normal|.sum(0)).mean()
//                                 or, without a weight map, (1 - (rend_normal * surf_normal).sum(0)).mean()
//   :176-178  distortion          rend_dist.mean()                       (lambda_dist is applied by the caller)
//   :121-122  first_order_edge_aware_loss(data, img) =
//                 (|spatial_gradient(data)| * exp(-|spatial_gradient(img)|)).sum(1).mean()
//             used on rend_normal (:183) and on surf_depth (:191), img = the ground-truth image
//   :127-139  get_img_grad_weight
// spatial_gradient is kornia 0.7.3's (requirements.txt:45): normalised 3x3 Sobel ([1 2 1]^T x [-1 0 1] / 8 and its
// transpose), replicate padding, outputs (d/dx, d/dy) per channel.
//
// Own design: one thread per pixel, 3x3 stencils straight out of L1 (the maps are 2.5 MB each, 12 of them), the four
// sums leave as one float4 partial per CTA and are added in a fixed order in double by a second tiny kernel
// (deterministic). The forward stores sign(grad data) * exp(-|grad img|) per pixel and direction (8 planes), so the
// backward is a 3x3 GATHER of those coefficients through the adjoint of the clamped stencil: no atomics.
#include "kernels.cuh"

namespace mrgs {

namespace {

constexpr int kGLx = 32, kGLy = 8;  // pixel block of a CTA

struct GeomLossParams {
    int H, W;
    unsigned flags;                 // MRGS_GEOM_* terms to evaluate
    const float* rend_normal;       // [3,H,W]
    const float* surf_normal;       // [3,H,W]
    const float* rend_dist;         // [H,W]
    const float* surf_depth;        // [H,W]
    const float* gt;                // [3,H,W]
    const float* weight;            // [H,W] or null
    float* coef;                    // [8,H,W]: normal c (x, y) for c = 0..2, then depth (x, y)
    float4* partials;
    const float* upstream;          // device [4]
    float* d_rend_normal;
    float* d_surf_normal;
    float* d_rend_dist;
    float* d_surf_depth;
};

__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// normalised Sobel pair at (y, x) of one plane with replicate padding
__device__ __forceinline__ void sobel(const float* __restrict__ pl, int y, int x, int H, int W, float& gx, float& gy) {
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const float* r0 = pl + (size_t)ym * W;
    const float* r1 = pl + (size_t)y * W;
    const float* r2 = pl + (size_t)yp * W;
    const float a00 = r0[xm], a01 = r0[x], a02 = r0[xp];
    const float a10 = r1[xm], a12 = r1[xp];
    const float a20 = r2[xm], a21 = r2[x], a22 = r2[xp];
    gx = ((a02 - a00) + 2.0f * (a12 - a10) + (a22 - a20)) * 0.125f;
    gy = ((a20 - a00) + 2.0f * (a21 - a01) + (a22 - a02)) * 0.125f;
}

__global__ void __launch_bounds__(kGLx * kGLy) geometry_loss_fwd_kernel(const GeomLossParams p) {
    __shared__ float4 s_red[kGLx * kGLy / 32];
    const int x = blockIdx.x * kGLx + threadIdx.x, y = blockIdx.y * kGLy + threadIdx.y;
    const int tid = threadIdx.y * kGLx + threadIdx.x;
    const size_t plane = (size_t)p.H * p.W;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < p.W && y < p.H) {
        const size_t o = (size_t)y * p.W + x;
        if (p.flags & MRGS_GEOM_NORMAL) {
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float rn = p.rend_normal[c * plane + o], sn = p.surf_normal[c * plane + o];
                acc += p.weight ? fabsf(sn - rn) : rn * sn;
            }
            t.x = p.weight ? p.weight[o] * acc : 1.0f - acc;
        }
        if (p.flags & MRGS_GEOM_DIST) t.y = p.rend_dist[o];
        if (p.flags & (MRGS_GEOM_NORMAL_SMOOTH | MRGS_GEOM_DEPTH_SMOOTH)) {
            float ex[3], ey[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float gx, gy;
                sobel(p.gt + c * plane, y, x, p.H, p.W, gx, gy);
                ex[c] = expf(-fabsf(gx));
                ey[c] = expf(-fabsf(gy));
            }
            if (p.flags & MRGS_GEOM_NORMAL_SMOOTH) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float gx, gy;
                    sobel(p.rend_normal + c * plane, y, x, p.H, p.W, gx, gy);
                    t.z += fabsf(gx) * ex[c] + fabsf(gy) * ey[c];
                    if (p.coef) {
                        p.coef[(2 * c + 0) * plane + o] = sgn(gx) * ex[c];
                        p.coef[(2 * c + 1) * plane + o] = sgn(gy) * ey[c];
                    }
                }
            }
            if (p.flags & MRGS_GEOM_DEPTH_SMOOTH) {
                float gx, gy;
                sobel(p.surf_depth, y, x, p.H, p.W, gx, gy);
                const float Ex = (ex[0] + ex[1]) + ex[2], Ey = (ey[0] + ey[1]) + ey[2];
                t.w = fabsf(gx) * Ex + fabsf(gy) * Ey;
                if (p.coef) {
                    p.coef[6 * plane + o] = sgn(gx) * Ex;
                    p.coef[7 * plane + o] = sgn(gy) * Ey;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, o);
        t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = t;
    __syncthreads();
    if (tid == 0) {
        float4 s = s_red[0];
        for (int w = 1; w < kGLx * kGLy / 32; ++w) { s.x += s_red[w].x; s.y += s_red[w].y; s.z += s_red[w].z; s.w += s_red[w].w; }
        p.partials[blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) geometry_loss_finish_kernel(const float4* __restrict__ partials, int n, double inv_hw,
                                                                   float* __restrict__ out4) {
    __shared__ double s[4][256];
    double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float4 v = partials[i];
        a += v.x; b += v.y; c += v.z; d += v.w;
    }
    s[0][threadIdx.x] = a; s[1][threadIdx.x] = b; s[2][threadIdx.x] = c; s[3][threadIdx.x] = d;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int k = 0; k < 4; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out4[0] = (float)(s[0][0] * inv_hw);
        out4[1] = (float)(s[1][0] * inv_hw);
        out4[2] = (float)(s[2][0] * inv_hw / 3.0);   // .sum(1).mean() over [3,H,W]
        out4[3] = (float)(s[3][0] * inv_hw / 3.0);
    }
}

// Adjoint of the clamped 3-tap stencils along one axis: for the three candidate output positions q = p-1, p, p+1,
// wa[q] = sum of the smoothing taps [1 2 1] of q that read input p, wb[q] = the same for the difference taps [-1 0 1].
__device__ __forceinline__ void adjoint_taps(int p, int n, float wa[3], float wb[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int q = p + d - 1;
        float a = 0.0f, b = 0.0f;
        if (q >= 0 && q < n) {
            if (max(q - 1, 0) == p) { a += 1.0f; b -= 1.0f; }
            if (q == p) a += 2.0f;
            if (min(q + 1, n - 1) == p) { a += 1.0f; b += 1.0f; }
        }
        wa[d] = a;
        wb[d] = b;
    }
}

__global__ void __launch_bounds__(kGLx * kGLy) geometry_loss_bwd_kernel(const GeomLossParams p) {
    const int x = blockIdx.x * kGLx + threadIdx.x, y = blockIdx.y * kGLy + threadIdx.y;
    if (x >= p.W || y >= p.H) return;
    const size_t plane = (size_t)p.H * p.W;
    const size_t o = (size_t)y * p.W + x;
    const float inv_hw = 1.0f / (float)plane;
    const float u0 = p.upstream[0] * inv_hw, u1 = p.upstream[1] * inv_hw;
    const float u2 = p.upstream[2] * inv_hw * (1.0f / 3.0f), u3 = p.upstream[3] * inv_hw * (1.0f / 3.0f);

    float d_rn[3] = {0.f, 0.f, 0.f}, d_sn[3] = {0.f, 0.f, 0.f}, d_depth = 0.0f;
    if (p.flags & MRGS_GEOM_NORMAL) {
        const float w = p.weight ? p.weight[o] : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float rn = p.rend_normal[c * plane + o], sn = p.surf_normal[c * plane + o];
            if (p.weight) {
                const float s = sgn(sn - rn) * w * u0;
                d_sn[c] = s;
                d_rn[c] = -s;
            } else {
                d_sn[c] = -rn * u0;
                d_rn[c] = -sn * u0;
            }
        }
    }
    if (p.flags & (MRGS_GEOM_NORMAL_SMOOTH | MRGS_GEOM_DEPTH_SMOOTH)) {
        float ay[3], by[3], ax[3], bx[3];
        adjoint_taps(y, p.H, ay, by);
        adjoint_taps(x, p.W, ax, bx);
        float gn[3] = {0.f, 0.f, 0.f}, gd = 0.0f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int qy = y + dy - 1;
            if (qy < 0 || qy >= p.H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int qx = x + dx - 1;
                if (qx < 0 || qx >= p.W) continue;
                const float wx = ay[dy] * bx[dx], wy = by[dy] * ax[dx];   // d out_x(q) / d in(p), d out_y(q) / d in(p) (x 8)
                const size_t q = (size_t)qy * p.W + qx;
                if (p.flags & MRGS_GEOM_NORMAL_SMOOTH) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        gn[c] += wx * p.coef[(2 * c) * plane + q] + wy * p.coef[(2 * c + 1) * plane + q];
                }
                if (p.flags & MRGS_GEOM_DEPTH_SMOOTH) gd += wx * p.coef[6 * plane + q] + wy * p.coef[7 * plane + q];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) d_rn[c] += gn[c] * (0.125f * u2);
        d_depth = gd * (0.125f * u3);
    }
    if (p.d_rend_normal)
#pragma unroll
        for (int c = 0; c < 3; ++c) p.d_rend_normal[c * plane + o] = d_rn[c];
    if (p.d_surf_normal)
#pragma unroll
        for (int c = 0; c < 3; ++c) p.d_surf_normal[c * plane + o] = d_sn[c];
    if (p.d_rend_dist) p.d_rend_dist[o] = (p.flags & MRGS_GEOM_DIST) ? u1 : 0.0f;
    if (p.d_surf_depth) p.d_surf_depth[o] = d_depth;
}

GeomLossParams to_params(const MrgsGeometryLossArgs* a) {
    GeomLossParams p{};
    p.H = a->height; p.W = a->width; p.flags = a->terms;
    p.rend_normal = a->rend_normal; p.surf_normal = a->surf_normal; p.rend_dist = a->rend_dist;
    p.surf_depth = a->surf_depth; p.gt = a->gt_image; p.weight = a->image_weight;
    p.coef = a->coef; p.partials = reinterpret_cast<float4*>(a->partials);
    p.upstream = a->upstream;
    p.d_rend_normal = a->dL_drend_normal; p.d_surf_normal = a->dL_dsurf_normal;
    p.d_rend_dist = a->dL_drend_dist; p.d_surf_depth = a->dL_dsurf_depth;
    return p;
}

// ---- get_img_grad_weight -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGLx * kGLy) img_grad_kernel(const float* __restrict__ img, int C, int H, int W,
                                                               float* __restrict__ out, unsigned* __restrict__ minmax) {
    // interior pixel (y, x), 1 <= y < H-1, 1 <= x < W-1: max over the two directions of the channel-mean central difference
    const int x = 1 + blockIdx.x * kGLx + threadIdx.x, y = 1 + blockIdx.y * kGLy + threadIdx.y;
    const size_t plane = (size_t)H * W;
    float g = 0.0f;
    const bool in = x < W - 1 && y < H - 1;
    if (in) {
        float sx = 0.0f, sy = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float* pl = img + c * plane;
            sx += fabsf(pl[(size_t)y * W + x + 1] - pl[(size_t)y * W + x - 1]);
            sy += fabsf(pl[(size_t)(y - 1) * W + x] - pl[(size_t)(y + 1) * W + x]);
        }
        g = fmaxf(sx / (float)C, sy / (float)C);
        out[(size_t)y * W + x] = g;
    }
    // values are >= 0, so their bit patterns order like unsigned integers
    unsigned lo = in ? __float_as_uint(g) : 0xffffffffu, hi = in ? __float_as_uint(g) : 0u;
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != 0xffffffffu) atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

__global__ void __launch_bounds__(256) img_grad_normalise_kernel(int H, int W, float* __restrict__ out,
                                                                 const unsigned* __restrict__ minmax) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const float mn = __uint_as_float(minmax[0]), mx = __uint_as_float(minmax[1]);
    const bool interior = x >= 1 && x < W - 1 && y >= 1 && y < H - 1;
    out[i] = interior ? (out[i] - mn) / (mx - mn) : 1.0f;   // F.pad(..., value=1.0)
}

}  // namespace

size_t geometry_loss_partials_count(int H, int W) {
    return (size_t)((H + kGLy - 1) / kGLy) * ((W + kGLx - 1) / kGLx);
}

int launch_geometry_loss(const MrgsGeometryLossArgs* a, bool backward, cudaStream_t stream) {
    const GeomLossParams p = to_params(a);
    const dim3 grid((p.W + kGLx - 1) / kGLx, (p.H + kGLy - 1) / kGLy), block(kGLx, kGLy);
    if (!backward) {
        geometry_loss_fwd_kernel<<<grid, block, 0, stream>>>(p);
        geometry_loss_finish_kernel<<<1, 256, 0, stream>>>(p.partials, (int)geometry_loss_partials_count(p.H, p.W),
                                                           1.0 / ((double)p.H * p.W), a->out4);
    } else {
        geometry_loss_bwd_kernel<<<grid, block, 0, stream>>>(p);
    }
    return MRGS_OK;
}

int launch_img_grad_weight(const float* img, int C, int H, int W, float* out, void* scratch8, cudaStream_t stream) {
    unsigned* mm = reinterpret_cast<unsigned*>(scratch8);
    cudaMemsetAsync(mm, 0xff, 4, stream);
    cudaMemsetAsync(mm + 1, 0x00, 4, stream);
    const dim3 grid((W - 2 + kGLx - 1) / kGLx, (H - 2 + kGLy - 1) / kGLy), block(kGLx, kGLy);
    img_grad_kernel<<<grid, block, 0, stream>>>(img, C, H, W, out, mm);
    const size_t n = (size_t)H * W;
    img_grad_normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(H, W, out, mm);
    return MRGS_OK;
}

}  // namespace mrgs
