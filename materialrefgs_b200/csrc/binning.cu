// binning.cu — surfel x tile instance generation, (tile | depth) ordering, per-tile ranges.
//
// Behavioural reference: duplicateWithKeys rasterizer_impl.cu:72-113, the SortPairs call
// :309-314 with getHigherMsb :37-52, identifyTileRanges :118-140 and the InclusiveSum :283.
// Required result (bit-exact): instances ordered by (tile id, depth bits) with ties kept in
// emission order, i.e. ascending surfel id, then row-major tile order of one surfel.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.cuh"

namespace mrgs {

uint32_t higher_msb(uint32_t n) {
    // smallest b such that n >> b == 0, found the way the reference's bisection does
    uint32_t msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb)
            msb += step;
        else
            msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

size_t scan_temp_bytes(int P) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return bytes;
}

size_t sort_temp_bytes(int64_t R) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)R);
    return bytes;
}

int run_inclusive_scan(const uint32_t* in, uint32_t* out, int P, void* temp, size_t temp_bytes,
                       cudaStream_t stream) {
    MRGS_CUDA_OK(cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, P, stream));
    return MRGS_OK;
}

int run_sort_pairs(const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                   uint32_t* vals_out, int R, int end_bit, void* temp, size_t temp_bytes,
                   cudaStream_t stream) {
    MRGS_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in,
                                                 vals_out, R, 0, end_bit, stream));
    return MRGS_OK;
}

namespace {

__global__ void __launch_bounds__(256)
duplicate_with_keys_kernel(int P, const float* __restrict__ depth, const uint2* __restrict__ rect,
                           const int* __restrict__ radii, const uint32_t* __restrict__ offsets,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ values, int grid_x) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    if (radii[idx] <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const uint2 r = rect[idx];
    const int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff, y1 = r.y >> 16;
    const uint32_t depth_bits = __float_as_uint(depth[idx]);
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * grid_x + x) << 32) | depth_bits;
            keys[off] = key;
            values[off] = (uint32_t)idx;
            ++off;
        }
    }
}

__global__ void __launch_bounds__(256)
identify_tile_ranges_kernel(int R, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R) return;
    const uint32_t tile = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[tile].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (tile != prev) {
            ranges[prev].y = idx;
            ranges[tile].x = idx;
        }
    }
    if (idx == R - 1) ranges[tile].y = R;
}

}  // namespace

void launch_duplicate_with_keys(int P, const float* depth, const uint2* rect, const int* radii,
                                const uint32_t* offsets, uint64_t* keys, uint32_t* values,
                                int grid_x, cudaStream_t stream) {
    duplicate_with_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, depth, rect, radii, offsets,
                                                                    keys, values, grid_x);
}

void launch_identify_tile_ranges(int R, const uint64_t* keys, uint2* ranges, cudaStream_t stream) {
    identify_tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, keys, ranges);
}

}  // namespace mrgs
