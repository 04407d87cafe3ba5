// prefilter_math.cuh — the per-tap arithmetic of the cubemap prefilter, shared by the cold-path kernels
// (cubemap.cu) and the plan builder (prefilter.cu). Expression order follows the reference so that the
// GGX weight at roughness 0.08 (where d = 1 - c^2 (1 - alpha^2) cancels to ~4e-5) rounds the same way:
//   pixel_area / cube_to_dir       scene/renderutils/c_src/cubemap.cu:17-46
//   ndfGGX                         c_src/cubemap.cu:176-181
//   specular tap weight            c_src/cubemap.cu:282-292
//   diffuse tap weight             c_src/cubemap.cu:130-134
#pragma once
#include "cube_sample.cuh"

namespace mrgs {

__device__ __forceinline__ float pixel_area(int x, int y, int N) {
    if (N > 1) {
        const int H = N / 2;
        x = abs(x - H);
        y = abs(y - H);
        const float dx = atanf((float)(x + 1) / (float)H) - atanf((float)x / (float)H);
        const float dy = atanf((float)(y + 1) / (float)H) - atanf((float)y / (float)H);
        return dx * dy;
    }
    return 1.0f;
}

__device__ __forceinline__ F3 normalize_safe(F3 v) {
    const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (l > 0.0f) return {v.x / l, v.y / l, v.z / l};
    return {0.f, 0.f, 0.f};
}

__device__ __forceinline__ F3 texel_dir(int x, int y, int side, int N) {
    const float fx = 2.0f * (((float)x + 0.5f) / (float)N) - 1.0f;
    const float fy = 2.0f * (((float)y + 0.5f) / (float)N) - 1.0f;
    return normalize_safe(face_to_dir(side, fx, fy));
}

__device__ __forceinline__ float ndf_ggx(float alphaSqr, float cosTheta) {
    const float c = fminf(fmaxf(cosTheta, 0.0f), 1.0f);
    const float d = (c * alphaSqr - c) * c + 1.0f;
    // the reference divides by the double constant M_PI (c_src/cubemap.cu:180): keep that rounding
    return (float)((double)alphaSqr / ((double)(d * d) * 3.14159265358979323846));
}

// weight of source texel (x, y) with direction L in the GGX lobe around V, given dot(L, V) >= cutoff
__device__ __forceinline__ float specular_tap_weight(F3 V, F3 L, float LdotV, float alphaSqr, float area) {
    const F3 Hh = normalize_safe(L + V);
    const float wiDotN = fmaxf(LdotV, 0.0f);
    const float VdotH = fmaxf(dot(V, Hh), 0.0f);
    return wiDotN * ndf_ggx(alphaSqr, VdotH) * area / 4.0f;
}

__device__ __forceinline__ float diffuse_tap_weight(F3 Nn, F3 L, float area) {
    const float costheta = fminf(fmaxf(dot(Nn, L), 0.0f), 0.999f);
    return costheta * area / 3.141592f;
}

}  // namespace mrgs
