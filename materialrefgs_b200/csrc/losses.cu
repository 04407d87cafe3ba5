// losses.cu — photometric loss terms on the rendered image: mean |x - y| and the mean of the
// 11x11 Gaussian-window SSIM map, forward and backward, each as ONE tiled kernel instead of five grouped
// convolutions + ~15 elementwise kernels and their autograd tape.
//
// Behavioural reference: utils/loss_utils.py:22-23 (l1_loss), :28-30 + :83-119 (gaussian window sigma 1.5,
// zero-padded conv2d, C1 = 0.01^2, C2 = 0.03^2, ssim_map.mean()); used by calculate_loss :155-157.
//
// Own design: a CTA owns a 16x16 pixel tile of one channel. The (16+10)^2 halo of both images is staged in shared
// memory once, the five moments (x, y, xx, yy, xy) are filtered separably (rows, then columns) out of shared
// memory, the SSIM value and its three partial derivatives are formed in registers, the tile's two sums leave as
// one partial per CTA (summed in a fixed order by a second tiny kernel: deterministic). The backward filters the
// three derivative maps the same way and combines them with x and y at the pixel.
#include "kernels.cuh"

namespace mrgs {

namespace {

constexpr int kLT = 16;            // tile edge
constexpr int kLR = 5;             // window radius
constexpr int kLH = kLT + 2 * kLR; // tile + halo
constexpr float kSsimC1 = 0.01f * 0.01f;
constexpr float kSsimC2 = 0.03f * 0.03f;

struct GaussWindow {
    float g[2 * kLR + 1];
};

struct LossParams {
    const float* img;
    const float* gt;
    int C, H, W;
    GaussWindow win;
    float* maps;       // [3][C][H][W]: dS/dmu1 (raw moments fixed), dS/dsigma1_sq, dS/dsigma12; may be null
    float2* partials;  // one (sum |x-y|, sum ssim) per CTA
    // backward
    const float* upstream;  // device [2]: dL/d(l1 mean), dL/d(ssim mean)
    float* dimg;
};

__global__ void __launch_bounds__(kLT * kLT) photometric_fwd_kernel(const LossParams p) {
    __shared__ float s_x[kLH][kLH + 1], s_y[kLH][kLH + 1];
    __shared__ float s_h[5][kLH][kLT + 1];
    __shared__ float2 s_red[kLT * kLT / 32];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kLT + tx;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT, c = blockIdx.z;
    const size_t plane = (size_t)p.H * p.W;
    const float* img = p.img + c * plane;
    const float* gt = p.gt + c * plane;
    for (int i = tid; i < kLH * kLH; i += kLT * kLT) {
        const int ly = i / kLH, lx = i - ly * kLH;
        const int gy = y0 + ly - kLR, gx = x0 + lx - kLR;
        const bool in = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;   // conv2d zero padding
        s_x[ly][lx] = in ? img[(size_t)gy * p.W + gx] : 0.0f;
        s_y[ly][lx] = in ? gt[(size_t)gy * p.W + gx] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kLH * kLT; i += kLT * kLT) {
        const int ly = i / kLT, lx = i - ly * kLT;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * kLR; ++k) {
            const float w = p.win.g[k], a = s_x[ly][lx + k], b = s_y[ly][lx + k];
            m1 += w * a; m2 += w * b; e11 += w * a * a; e22 += w * b * b; e12 += w * a * b;
        }
        s_h[0][ly][lx] = m1; s_h[1][ly][lx] = m2; s_h[2][ly][lx] = e11; s_h[3][ly][lx] = e22; s_h[4][ly][lx] = e12;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * kLR; ++k) {
        const float w = p.win.g[k];
        mu1 += w * s_h[0][ty + k][tx]; mu2 += w * s_h[1][ty + k][tx];
        e11 += w * s_h[2][ty + k][tx]; e22 += w * s_h[3][ty + k][tx]; e12 += w * s_h[4][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    const bool inside = gx < p.W && gy < p.H;
    float l1 = 0.f, ssim = 0.f;
    if (inside) {
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
        const float a = 2.0f * mu12 + kSsimC1, b = 2.0f * s12 + kSsimC2;
        const float cc = mu1_sq + mu2_sq + kSsimC1, d = s1 + s2 + kSsimC2;
        ssim = (a * b) / (cc * d);
        l1 = fabsf(s_x[ty + kLR][tx + kLR] - s_y[ty + kLR][tx + kLR]);
        if (p.maps != nullptr) {
            const float inv_cd = 1.0f / (cc * d);
            const float dS_dmu1_sigma = 2.0f * b * inv_cd * (mu2 - a * mu1 / cc);   // sigma terms held fixed
            const float dS_ds1 = -a * b * inv_cd / d;
            const float dS_ds12 = 2.0f * a * inv_cd;
            const size_t o = c * plane + (size_t)gy * p.W + gx;
            const size_t chw = (size_t)p.C * plane;
            // raw moments fixed: sigma1_sq = E11 - mu1^2, sigma12 = E12 - mu1 mu2
            p.maps[o] = dS_dmu1_sigma - 2.0f * mu1 * dS_ds1 - mu2 * dS_ds12;
            p.maps[chw + o] = dS_ds1;
            p.maps[2 * chw + o] = dS_ds12;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = make_float2(l1, ssim);
    __syncthreads();
    if (tid == 0) {
        float2 t = make_float2(0.f, 0.f);
        for (int w = 0; w < kLT * kLT / 32; ++w) { t.x += s_red[w].x; t.y += s_red[w].y; }
        p.partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

// fixed-order sum of the per-CTA partials in double, divided by the element count
__global__ void __launch_bounds__(256) photometric_finish_kernel(const float2* __restrict__ partials, int n, double inv_count,
                                                                 float* __restrict__ out2) {
    __shared__ double s_a[256], s_b[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { a += partials[i].x; b += partials[i].y; }
    s_a[threadIdx.x] = a; s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_b[threadIdx.x] += s_b[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = (float)(s_a[0] * inv_count); out2[1] = (float)(s_b[0] * inv_count); }
}

__global__ void __launch_bounds__(kLT * kLT) photometric_bwd_kernel(const LossParams p) {
    __shared__ float s_m[3][kLH][kLH + 1];
    __shared__ float s_h[3][kLH][kLT + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kLT + tx;
    const int x0 = blockIdx.x * kLT, y0 = blockIdx.y * kLT, c = blockIdx.z;
    const size_t plane = (size_t)p.H * p.W, chw = (size_t)p.C * plane;
    for (int i = tid; i < kLH * kLH; i += kLT * kLT) {
        const int ly = i / kLH, lx = i - ly * kLH;
        const int gy = y0 + ly - kLR, gx = x0 + lx - kLR;
        const bool in = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;   // no SSIM pixel exists outside the image
        const size_t o = c * plane + (size_t)gy * p.W + gx;
#pragma unroll
        for (int m = 0; m < 3; ++m) s_m[m][ly][lx] = in ? p.maps[m * chw + o] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kLH * kLT; i += kLT * kLT) {
        const int ly = i / kLT, lx = i - ly * kLT;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k <= 2 * kLR; ++k) {
            const float w = p.win.g[k];
            a += w * s_m[0][ly][lx + k]; b += w * s_m[1][ly][lx + k]; d += w * s_m[2][ly][lx + k];
        }
        s_h[0][ly][lx] = a; s_h[1][ly][lx] = b; s_h[2][ly][lx] = d;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= p.W || gy >= p.H) return;
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * kLR; ++k) {
        const float w = p.win.g[k];
        a += w * s_h[0][ty + k][tx]; b += w * s_h[1][ty + k][tx]; d += w * s_h[2][ty + k][tx];
    }
    const size_t o = c * plane + (size_t)gy * p.W + gx;
    const float x = p.img[o], y = p.gt[o];
    const float inv_n = 1.0f / (float)chw;
    const float g_l1 = p.upstream[0] * inv_n, g_ssim = p.upstream[1] * inv_n;
    const float diff = x - y;
    const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);   // torch.abs backward: sign(0) = 0
    p.dimg[o] = g_l1 * sgn + g_ssim * (a + 2.0f * x * b + y * d);
}

GaussWindow make_window() {
    // utils/loss_utils.py:28-30: torch.Tensor([exp(-(x-5)^2 / (2*1.5^2))]) / sum, in float32. The SSIM map is
    // sensitive to the LAST BIT of these weights (sigma = E[x^2] - mu^2 cancels, so a 1e-7 error in the
    // weight sum moves the mean SSIM by ~3e-6), hence the exact float32 values torch produces, as hex floats.
    static const float k[2 * kLR + 1] = {0x1.0d956cp-10f, 0x1.f1fe02p-8f, 0x1.26eb18p-5f, 0x1.bff0fep-4f,
                                         0x1.b43c3ep-3f,  0x1.106560p-2f, 0x1.b43c3ep-3f, 0x1.bff0fep-4f,
                                         0x1.26eb18p-5f,  0x1.f1fe02p-8f, 0x1.0d956cp-10f};
    GaussWindow w;
    for (int i = 0; i <= 2 * kLR; ++i) w.g[i] = k[i];
    return w;
}

}  // namespace

size_t photometric_partials_count(int C, int H, int W) {
    return (size_t)C * ((H + kLT - 1) / kLT) * ((W + kLT - 1) / kLT);
}

int launch_photometric_fwd(const float* img, const float* gt, int C, int H, int W, float* maps, float* partials,
                           float* out2, cudaStream_t stream) {
    LossParams p{};
    p.img = img; p.gt = gt; p.C = C; p.H = H; p.W = W; p.win = make_window();
    p.maps = maps; p.partials = reinterpret_cast<float2*>(partials);
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, C), block(kLT, kLT);
    photometric_fwd_kernel<<<grid, block, 0, stream>>>(p);
    photometric_finish_kernel<<<1, 256, 0, stream>>>(p.partials, (int)photometric_partials_count(C, H, W),
                                                     1.0 / ((double)C * H * W), out2);
    return MRGS_OK;
}

int launch_photometric_bwd(const float* img, const float* gt, const float* maps, int C, int H, int W,
                           const float* upstream, float* dimg, cudaStream_t stream) {
    LossParams p{};
    p.img = img; p.gt = gt; p.C = C; p.H = H; p.W = W; p.win = make_window();
    p.maps = const_cast<float*>(maps); p.upstream = upstream; p.dimg = dimg;
    const dim3 grid((W + kLT - 1) / kLT, (H + kLT - 1) / kLT, C), block(kLT, kLT);
    photometric_bwd_kernel<<<grid, block, 0, stream>>>(p);
    return MRGS_OK;
}

}  // namespace mrgs
