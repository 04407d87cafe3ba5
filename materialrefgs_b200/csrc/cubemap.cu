// cubemap.cu — EnvLight.build_mips on the device: 2x2 mip reduction (and the reference's
// non-adjoint bilinear backward), GGX-prefiltered specular levels with cached per-texel bounds, and
// the cosine-convolved diffuse level.
//
// Behavioural reference (scene/renderutils = nvdiffrec's renderutils as carried by the reference):
//   cubemap_mip fwd/bwd            scene/light_utils.py:66-80
//   pixel_area / cube_to_dir       scene/renderutils/c_src/cubemap.cu:17-46
//   SpecularBoundsKernel           c_src/cubemap.cu:183-246   (incl. its 16x16 interval culling)
//   SpecularCubemapFwd/BwdKernel   c_src/cubemap.cu:248-354   ndfGGX :176-181
//   DiffuseCubemapFwd/BwdKernel    c_src/cubemap.cu:110-171
// Layout: cubemaps are float [6][res][res][C] (NHWC as in the reference); bounds are int32
// [6][res][res][6][4] = (xmin,xmax,ymin,ymax) per source face (the reference stores them as floats).
#include "kernels.cuh"
#include "prefilter_math.cuh"

namespace mrgs {

namespace {

__global__ void mip_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int res_out, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = 6 * res_out * res_out * C;
    if (i >= total) return;
    const int c = i % C;
    const int x = (i / C) % res_out;
    const int y = (i / (C * res_out)) % res_out;
    const int s = i / (C * res_out * res_out);
    const int ri = res_out * 2;
    const float* p = in + ((size_t)(s * ri + 2 * y) * ri + 2 * x) * C + c;
    // avg_pool2d sums the window then scales (ATen accumulates in row-major window order)
    out[i] = (p[0] + p[C] + p[(size_t)ri * C] + p[(size_t)ri * C + C]) * 0.25f;
}

// din(fine texel) = seamless bilinear fetch of (dout * 0.25) at the fine texel's direction
__global__ void mip_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int res_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int rf = res_out * 2;
    if (i >= 6 * rf * rf) return;
    const int x = i % rf, y = (i / rf) % rf, s = i / (rf * rf);
    const F3 d = texel_dir(x, y, s, rf);
    const FaceUV f = dir_to_face(d);
    Bilinear b;
    cube_bilinear<false>(dout, res_out, f.face, f.u, f.v, b);
    din[3 * (size_t)i + 0] = 0.25f * b.val.x;
    din[3 * (size_t)i + 1] = 0.25f * b.val.y;
    din[3 * (size_t)i + 2] = 0.25f * b.val.z;
}

__global__ void __launch_bounds__(64) specular_bounds_kernel(int N, float cutoff, int4* __restrict__ bounds) {
    const int px = blockIdx.x * 8 + (threadIdx.x & 7);
    const int py = blockIdx.y * 8 + (threadIdx.x >> 3);
    const int pz = blockIdx.z;
    if (px >= N || py >= N) return;
    const F3 V = texel_dir(px, py, pz, N);
    constexpr int TS = 16;
    for (int s = 0; s < 6; ++s) {
        int min_x = N - 1, max_x = 0, min_y = N - 1, max_y = 0;
        for (int tx = 0; tx < (N + TS - 1) / TS; ++tx) {
            for (int ty = 0; ty < (N + TS - 1) / TS; ++ty) {
                const int tsx = tx * TS, tsy = ty * TS;
                const int tex = min((tx + 1) * TS, N), tey = min((ty + 1) * TS, N);
                const F3 L0 = texel_dir(tsx, tsy, s, N), L1 = texel_dir(tex, tsy, s, N);
                const F3 L2 = texel_dir(tsx, tey, s, N), L3 = texel_dir(tex, tey, s, N);
                const float minx = fminf(fminf(L0.x, L1.x), fminf(L2.x, L3.x)), maxx = fmaxf(fmaxf(L0.x, L1.x), fmaxf(L2.x, L3.x));
                const float miny = fminf(fminf(L0.y, L1.y), fminf(L2.y, L3.y)), maxy = fmaxf(fmaxf(L0.y, L1.y), fmaxf(L2.y, L3.y));
                const float minz = fminf(fminf(L0.z, L1.z), fminf(L2.z, L3.z)), maxz = fmaxf(fmaxf(L0.z, L1.z), fmaxf(L2.z, L3.z));
                const float maxdp = fmaxf(minx * V.x, maxx * V.x) + fmaxf(miny * V.y, maxy * V.y) + fmaxf(minz * V.z, maxz * V.z);
                if (maxdp >= cutoff) {
                    for (int y = tsy; y < tey; ++y)
                        for (int x = tsx; x < tex; ++x) {
                            const F3 L = texel_dir(x, y, s, N);
                            if (dot(L, V) >= cutoff) {
                                min_x = min(min_x, x);
                                max_x = max(max_x, x);
                                min_y = min(min_y, y);
                                max_y = max(max_y, y);
                            }
                        }
                }
            }
        }
        bounds[((size_t)(pz * N + py) * N + px) * 6 + s] = make_int4(min_x, max_x, min_y, max_y);
    }
}

template <bool BACKWARD>
__global__ void __launch_bounds__(64)
specular_kernel(int N, float roughness, float cutoff, const int4* __restrict__ bounds,
                const float* __restrict__ cubemap, float* __restrict__ out4, const float* __restrict__ dout4,
                float* __restrict__ dcubemap) {
    const int px = blockIdx.x * 8 + (threadIdx.x & 7);
    const int py = blockIdx.y * 8 + (threadIdx.x >> 3);
    const int pz = blockIdx.z;
    if (px >= N || py >= N) return;
    const size_t o = (size_t)(pz * N + py) * N + px;
    const F3 V = texel_dir(px, py, pz, N);
    const float alpha = roughness * roughness;
    const float alphaSqr = alpha * alpha;
    F3 grad = {0.f, 0.f, 0.f};
    if (BACKWARD) {
        grad = {dout4[4 * o], dout4[4 * o + 1], dout4[4 * o + 2]};
        if (grad.x == 0.0f && grad.y == 0.0f && grad.z == 0.0f) return;
    }
    float wsum = 0.0f;
    F3 col = {0.f, 0.f, 0.f};
    for (int s = 0; s < 6; ++s) {
        const int4 b = bounds[o * 6 + s];
        if (b.x > b.y) continue;
        for (int y = b.z; y <= b.w; ++y)
            for (int x = b.x; x <= b.y; ++x) {
                const F3 L = texel_dir(x, y, s, N);
                const float LdotV = dot(L, V);
                if (LdotV < cutoff) continue;
                const float w = specular_tap_weight(V, L, LdotV, alphaSqr, pixel_area(x, y, N));
                const size_t t = ((size_t)(s * N + y) * N + x) * 3;
                if (BACKWARD) {
                    atomicAdd(dcubemap + t + 0, grad.x * w);
                    atomicAdd(dcubemap + t + 1, grad.y * w);
                    atomicAdd(dcubemap + t + 2, grad.z * w);
                } else {
                    col.x += cubemap[t] * w;
                    col.y += cubemap[t + 1] * w;
                    col.z += cubemap[t + 2] * w;
                    wsum += w;
                }
            }
    }
    if (!BACKWARD) {
        out4[4 * o] = col.x;
        out4[4 * o + 1] = col.y;
        out4[4 * o + 2] = col.z;
        out4[4 * o + 3] = wsum;
    }
}

template <bool BACKWARD>
__global__ void __launch_bounds__(64)
diffuse_kernel(int N, const float* __restrict__ cubemap, float* __restrict__ out, const float* __restrict__ dout,
               float* __restrict__ dcubemap) {
    const int px = blockIdx.x * 8 + (threadIdx.x & 7);
    const int py = blockIdx.y * 8 + (threadIdx.x >> 3);
    const int pz = blockIdx.z;
    if (px >= N || py >= N) return;
    const size_t o = ((size_t)(pz * N + py) * N + px) * 3;
    const F3 Nn = texel_dir(px, py, pz, N);
    F3 grad = {0.f, 0.f, 0.f};
    if (BACKWARD) {
        grad = {dout[o], dout[o + 1], dout[o + 2]};
        if (grad.x == 0.0f && grad.y == 0.0f && grad.z == 0.0f) return;
    }
    F3 col = {0.f, 0.f, 0.f};
    for (int s = 0; s < 6; ++s)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                const F3 L = texel_dir(x, y, s, N);
                const float w = diffuse_tap_weight(Nn, L, pixel_area(x, y, N));
                const size_t t = ((size_t)(s * N + y) * N + x) * 3;
                if (BACKWARD) {
                    atomicAdd(dcubemap + t + 0, grad.x * w);
                    atomicAdd(dcubemap + t + 1, grad.y * w);
                    atomicAdd(dcubemap + t + 2, grad.z * w);
                } else {
                    col.x += cubemap[t] * w;
                    col.y += cubemap[t + 1] * w;
                    col.z += cubemap[t + 2] * w;
                }
            }
    if (!BACKWARD) {
        out[o] = col.x;
        out[o + 1] = col.y;
        out[o + 2] = col.z;
    }
}

}  // namespace

void launch_cubemap_mip_fwd(const float* in, float* out, int res_out, int C, cudaStream_t stream) {
    const int total = 6 * res_out * res_out * C;
    mip_fwd_kernel<<<(total + 255) / 256, 256, 0, stream>>>(in, out, res_out, C);
}
void launch_cubemap_mip_bwd(const float* dout, float* din, int res_out, cudaStream_t stream) {
    const int total = 6 * 4 * res_out * res_out;
    mip_bwd_kernel<<<(total + 255) / 256, 256, 0, stream>>>(dout, din, res_out);
}
void launch_specular_bounds(int N, float cutoff, int32_t* bounds, cudaStream_t stream) {
    const dim3 grid((N + 7) / 8, (N + 7) / 8, 6);
    specular_bounds_kernel<<<grid, 64, 0, stream>>>(N, cutoff, reinterpret_cast<int4*>(bounds));
}
void launch_specular_cubemap(bool backward, int N, float roughness, float cutoff, const int32_t* bounds,
                             const float* cubemap, float* out4, const float* dout4, float* dcubemap,
                             cudaStream_t stream) {
    const dim3 grid((N + 7) / 8, (N + 7) / 8, 6);
    const int4* b = reinterpret_cast<const int4*>(bounds);
    if (backward)
        specular_kernel<true><<<grid, 64, 0, stream>>>(N, roughness, cutoff, b, cubemap, out4, dout4, dcubemap);
    else
        specular_kernel<false><<<grid, 64, 0, stream>>>(N, roughness, cutoff, b, cubemap, out4, dout4, dcubemap);
}
void launch_diffuse_cubemap(bool backward, int N, const float* cubemap, float* out, const float* dout,
                            float* dcubemap, cudaStream_t stream) {
    const dim3 grid((N + 7) / 8, (N + 7) / 8, 6);
    if (backward)
        diffuse_kernel<true><<<grid, 64, 0, stream>>>(N, cubemap, out, dout, dcubemap);
    else
        diffuse_kernel<false><<<grid, 64, 0, stream>>>(N, cubemap, out, dout, dcubemap);
}

}  // namespace mrgs
