// kernels.cuh — parameter blocks and launchers of the individual pipeline stages.
#pragma once
#include "common.cuh"

namespace mrgs {

struct PreprocessParams {
    int P, S, D, M, W, H;
    int grid_x, grid_y;
    int cf_stride;
    float scale_modifier;
    int prefiltered;
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* opacities;
    const float* shs;
    const float* colors_precomp;
    const float* features;
    const float* transMat_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    // outputs
    int* radii;
    float* rec;
    float* cf;
    uint8_t* clamped;
    uint32_t* tiles_touched;
    uint2* rect;
    float* depth;
    float4* bbox;
    uint32_t* sort_key;  // depth bits, 0xffffffff for culled surfels
};

void launch_preprocess_fwd(const PreprocessParams& p, cudaStream_t stream);
void launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                         cudaStream_t stream);

// ---- binning (all hand-written, see binning.cu) ------------------------------------------------
uint32_t higher_msb(uint32_t n);
int sort_blocks(int64_t n);
size_t sort_hist_bytes(int64_t n);
int scan_blocks(int n);
// depth-sorts the P surfels (keys_a holds the 32-bit keys, values are implicit ids); *order = sorted ids
int depth_sort(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int P, uint32_t* block_hist,
               cudaStream_t stream, uint32_t** order, int* launches);
// offsets_incl[k] = inclusive prefix sum of tiles_touched[order[k]]
int offsets_in_order(const uint32_t* order, const uint32_t* tiles_touched, int P, uint32_t* block_sums,
                     uint32_t* offsets_incl, cudaStream_t stream);
void launch_emit_instances(int P, const uint32_t* order, const uint32_t* tiles_touched, const uint2* rect,
                           const uint32_t* offsets_incl, uint16_t* keys, uint32_t* values, int grid_x,
                           uint32_t capacity, cudaStream_t stream);
// R_dev != nullptr: R is only a capacity (grids, scratch); the kernels read the real count from *R_dev
int tile_sort(uint16_t* keys_a, uint16_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int R, int tile_bits,
              uint32_t* block_hist, cudaStream_t stream, uint16_t** keys_sorted, uint32_t** vals_sorted, int* launches,
              const uint32_t* R_dev);
void launch_identify_tile_ranges(int R, const uint32_t* R_dev, const uint16_t* keys, uint2* ranges, cudaStream_t stream);
bool radix_lookback_enabled();

// ---- tile blend ----------------------------------------------------------------------------
struct RenderFwdParams {
    int S, W, H, grid_x, grid_y, cf_stride;
    const uint2* ranges;
    const uint32_t* point_list;
    const float* rec;
    const float* cf;
    const float4* bbox;
    const float* background;
    float* state;  // [tiles][5][256]
    float* out_color;
    float* out_feature;
    float* out_others;
};
int launch_render_fwd(const RenderFwdParams& p, cudaStream_t stream);

struct RenderBwdParams {
    int S, W, H, grid_x, grid_y, cf_stride, grad_stride;
    const uint2* ranges;
    const uint32_t* point_list;
    const float* rec;
    const float* cf;
    const float4* bbox;
    const float* background;
    const float* state;
    const float* dL_dcolor;
    const float* dL_dfeature;
    const float* dL_dothers;
    float* grad_arena;  // [P][grad_stride], zero-initialised
};
int launch_render_bwd(const RenderBwdParams& p, cudaStream_t stream);

struct PreprocessBwdParams {
    int P, S, D, M, W, H;  // W,H here are the reference's recomputed int(focal*tan*2) values
    int cf_stride, grad_stride;
    int accumulate;  // add to the parameter-gradient outputs instead of overwriting them
    const float* means3D;
    const float* scales;
    const float* rotations;
    const float* shs;
    const float* transMat_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    const int* radii;
    const float* rec;
    const uint8_t* clamped;
    const float* grad_arena;
    float* dL_dmeans2D;
    float* dL_dcolors;
    float* dL_dfeatures;
    float* dL_dopacity;
    float* dL_dmeans3D;
    float* dL_dtransMat;
    float* dL_dsh;
    float* dL_dscales;
    float* dL_drotations;
};
void launch_preprocess_bwd(const PreprocessBwdParams& p, cudaStream_t stream);

// ---- deferred shading ------------------------------------------------------------------------
int launch_shade(const MrgsShadeArgs* a, bool backward, cudaStream_t stream);
size_t photometric_partials_count(int C, int H, int W);
int launch_photometric_fwd(const float* img, const float* gt, int C, int H, int W, float* maps, float* partials,
                           float* out2, cudaStream_t stream);
int launch_photometric_bwd(const float* img, const float* gt, const float* maps, int C, int H, int W,
                           const float* upstream, float* dimg, cudaStream_t stream);
size_t geometry_loss_partials_count(int H, int W);
int launch_geometry_loss(const MrgsGeometryLossArgs* a, bool backward, cudaStream_t stream);
int launch_img_grad_weight(const float* img, int C, int H, int W, float* out, void* scratch8, cudaStream_t stream);
int launch_surfel_features(const MrgsSurfelFeatureArgs* a, bool backward, cudaStream_t stream);
int launch_envlight_query(const MrgsShadeArgs* a, long long n, const float* dirs, const float* roughness,
                          float* out, cudaStream_t stream);
int launch_surfel_shade(const MrgsSurfelShadeArgs* a, bool backward, cudaStream_t stream);
int launch_envlight_query_bwd(const MrgsShadeArgs* a, long long n, const float* dirs, const float* roughness,
                              const float* dL_dout, float* dL_ddirs, float* dL_droughness, cudaStream_t stream);

int launch_depth_normal(bool backward, int W, int H, float depth_ratio, const float* A, const float* o,
                        const float* allmap, float* surf_depth, float* surf_normal, const float* g_depth,
                        const float* g_normal, float* g_allmap, cudaStream_t stream);

// ---- cubemap prefilter (EnvLight.build_mips) ---------------------------------------------------
void launch_cubemap_mip_fwd(const float* in, float* out, int res_out, int C, cudaStream_t stream);
void launch_cubemap_mip_bwd(const float* dout, float* din, int res_out, cudaStream_t stream);
void launch_specular_bounds(int N, float cutoff, int32_t* bounds, cudaStream_t stream);
void launch_specular_cubemap(bool backward, int N, float roughness, float cutoff, const int32_t* bounds,
                             const float* cubemap, float* out4, const float* dout4, float* dcubemap,
                             cudaStream_t stream);
void launch_diffuse_cubemap(bool backward, int N, const float* cubemap, float* out, const float* dout,
                            float* dcubemap, cudaStream_t stream);

// ---- prefilter plans (prefilter.cu) ---------------------------------------------------------------
void launch_texel_table(int N, float* table, cudaStream_t stream);
int prefilter_patch_count(int N, int G, int PW);
int launch_prefilter_build(const MrgsPrefilterBuildArgs* a, bool fill, cudaStream_t stream);
int launch_prefilter_apply(const MrgsPrefilterJob* jobs, int num_jobs, int max_ctas, cudaStream_t stream);
int launch_mip_pyramid(const float* base3, int res, int num_levels, float* const* levels4, cudaStream_t stream,
                       int* launches);
void launch_mip_bwd_acc(const float* g_coarse, const float* g_coarse2, float* g_fine, int res_coarse, int accumulate,
                        cudaStream_t stream);

}  // namespace mrgs
