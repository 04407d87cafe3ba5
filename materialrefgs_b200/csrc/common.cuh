// common.cuh — shared declarations for libmrgs.so (sm_100a only).
//
// The numerical contract (thresholds, tile size, channel layout) follows the reference
// rasterizer: rast/cuda_rasterizer/auxiliary.h:20-41, config.h:17-20 (rast/ =
// submodules/diff-surfel-rasterization/ of the MaterialRefGS tree).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mrgs.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmrgs is written for sm_100a (B200) only"
#endif

namespace mrgs {

constexpr int kTileX = MRGS_TILE_X;
constexpr int kTileY = MRGS_TILE_Y;
constexpr int kTilePixels = MRGS_TILE_PIXELS;
constexpr int kBatch = 256;          // surfel instances staged per tile per round
constexpr int kWarpsPerTile = 8;
constexpr int kGeomFloats = MRGS_GEOM_FLOATS;

// blend thresholds (auxiliary.h:39-41, forward.cu:396-404)
constexpr float kNear = 0.2f;
constexpr float kFar = 100.0f;
constexpr float kFarOverRange = kFar / (kFar - kNear);  // 1.00200403f, folded in fp32 like nvcc does
constexpr float kAlphaMax = 0.99f;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTMin = 0.0001f;

// allmap channel order (auxiliary.h:25-29)
constexpr int kDepthOff = 0;
constexpr int kAlphaOff = 1;
constexpr int kNormalOff = 2;
constexpr int kMidDepthOff = 5;
constexpr int kDistortionOff = 6;

// raw gradient arena row: [dT(9) | dmean2D(2) | dopacity(1) | dnormal(3) | dcolor(3) | dfeature(S)]
constexpr int kGradT = 0;
constexpr int kGradMean2D = 9;
constexpr int kGradOpacity = 11;
constexpr int kGradNormal = 12;
constexpr int kGradColor = 15;
constexpr int kGradFeature = 18;

__host__ __device__ constexpr int round_up4(int x) { return (x + 3) & ~3; }
__host__ __device__ constexpr int cf_stride(int S) { return round_up4(3 + S); }
// 15 geometry gradients + the padded colour/feature channels, rounded up to whole float4s
__host__ __device__ constexpr int grad_stride(int S) { return 16 + cf_stride(S); }

// pixel <-> slot mapping inside a 16x16 tile: warp w owns an 8x4 pixel block, blocks are laid
// out 2 (x) by 4 (y); lane l is pixel (l&7, l>>3) of its block.
__host__ __device__ inline int slot_x(int slot) { return ((slot >> 5) & 1) * 8 + (slot & 7); }
__host__ __device__ inline int slot_y(int slot) { return (slot >> 6) * 4 + ((slot >> 3) & 3); }
__host__ __device__ inline int slot_of(int x, int y) {
    return ((y >> 2) * 2 + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7);
}

void set_error(const char* fmt, ...);
int check_launch(const char* what, cudaStream_t stream, bool debug);

}  // namespace mrgs

#define MRGS_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            mrgs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                      \
            return MRGS_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

#define MRGS_LAUNCH_OK(what, stream, debug)                              \
    do {                                                                 \
        int _s = mrgs::check_launch(what, stream, debug);                \
        if (_s != MRGS_OK) return _s;                                    \
    } while (0)
