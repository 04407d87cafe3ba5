/*
 * mrgs.h — C ABI of libmrgs.so, the B200-native (sm_100a) surfel-splatting render path.
 *
 * This library is a drop-in for the compute side of MaterialRefGS's
 * submodules/diff-surfel-rasterization and of the split-sum deferred shading in
 * utils/refl_utils.py + scene/light.py + scene/renderutils (cubemap prefilter).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer unless the field name starts with `host_`;
 *   - all arithmetic data is fp32, ids/counters are 32-bit, sort keys 64-bit;
 *   - matrices use the reference's row-vector convention: the 16 floats are indexed
 *     m[col*4+row] (rast/cuda_rasterizer/auxiliary.h:80-99);
 *   - every entry point takes the CUDA stream to launch on (`void*` == cudaStream_t) and
 *     returns an int status: 0 = ok, MRGS_ERR_* otherwise. `mrgs_last_error()` returns a
 *     human-readable message for the calling thread;
 *   - no torch types cross this boundary; scratch memory is owned by the caller.
 *
 * Reference interface each entry point replaces (paths relative to the reference root,
 * rast/ = submodules/diff-surfel-rasterization/):
 *   mrgs_forward        <- CudaRasterizer::Rasterizer::forward   rast/cuda_rasterizer/rasterizer.h:33,
 *                          called from RasterizeGaussiansCUDA      rast/rasterize_points.cu:41-144
 *   mrgs_backward       <- CudaRasterizer::Rasterizer::backward  rast/cuda_rasterizer/rasterizer.h:61,
 *                          called from RasterizeGaussiansBackwardCUDA rast/rasterize_points.cu:146-252
 *   mrgs_mark_visible   <- CudaRasterizer::Rasterizer::markVisible rast/cuda_rasterizer/rasterizer.h:26,
 *                          called from markVisible                 rast/rasterize_points.cu:254-273
 *   mrgs_*_bytes        <- required<GeometryState/ImageState/BinningState>() rast/cuda_rasterizer/rasterizer_impl.h:68-74
 *   mrgs_shade_forward/backward <- get_specular_color_surfel utils/refl_utils.py:364-419,
 *                          EnvLight.get_mip/__call__ scene/light.py:88-129 and the compositing in
 *                          render_surfel gaussian_renderer/__init__.py:419-445
 *   mrgs_cubemap_*      <- scene/renderutils/c_src/cubemap.cu:110-354 + scene/light_utils.py:66-80
 *   mrgs_envlight_query(_backward) <- EnvLight.__call__ scene/light.py:98-129 (nvdiffrast dr.texture fwd/bwd + sigmoid)
 *   mrgs_surfel_shade_* <- get_full_color_volume(_indirect) utils/refl_utils.py:426-490 (per-surfel shading of render_volume)
 *   mrgs_depth_normal_* <- depth_to_normal utils/point_utils.py:9-37 + compute_2dgs_normal_and_regularizations
 *                          gaussian_renderer/__init__.py:42-90
 *   mrgs_surfel_features_* <- the activated getters + get_normal + eval_sh of gaussian_renderer/__init__.py:259-353
 *   mrgs_photometric_*, mrgs_geometry_loss_*, mrgs_img_grad_weight <- calculate_loss / get_img_grad_weight
 *                          utils/loss_utils.py:22-23, :83-139, :142-228
 *   mrgs_densify_stats  <- GaussianModel.add_densification_stats scene/gaussian_model.py:1059-1061
 *
 * ABI history: 5 = gradient sink (accumulate) in mrgs_backward; 6 = mrgs_geometry_loss_*, mrgs_img_grad_weight,
 * mrgs_envlight_query_backward, mrgs_surfel_shade_*, single-level chains accepted by mrgs_envlight_query;
 * 7 = prefilter plans (mrgs_prefilter_*), mrgs_mip_pyramid_forward, mrgs_mip_chain_backward;
 * 8 = MrgsForwardArgs.no_wait / count_out (capture-safe forward), mrgs_profile_collect_captured;
 * 9 = MrgsPrefilterJob.patch_begin / patch_end; 10 = mrgs_prefilter_apply(max_ctas).
 */
#ifndef MRGS_H_INCLUDED
#define MRGS_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRGS_ABI_VERSION 10

#if defined(__GNUC__)
#define MRGS_API __attribute__((visibility("default")))
#else
#define MRGS_API
#endif

/* compile-time constants of the numerical contract (rast/cuda_rasterizer/config.h:17-20,
 * auxiliary.h:23-41) */
#define MRGS_TILE_X 16
#define MRGS_TILE_Y 16
#define MRGS_TILE_PIXELS 256
#define MRGS_MAX_FEATURES 24
#define MRGS_NUM_CHANNELS 3
#define MRGS_NUM_OTHERS 7 /* depth, alpha, normal xyz, median depth, distortion */

/* status codes */
#define MRGS_OK 0
#define MRGS_ERR_INVALID_ARGUMENT 1
#define MRGS_ERR_CUDA 2
#define MRGS_ERR_WORKSPACE 3
#define MRGS_ERR_UNSUPPORTED 4

/* floats per surfel of the two packed per-surfel records inside the geometry buffer */
#define MRGS_GEOM_FLOATS 16

/* Allocation callback used for the one scratch buffer whose size is only known after
 * preprocessing (the number of surfel x tile instances R). Mirrors the reference's
 * std::function<char*(size_t)> trick (rast/rasterize_points.cu:33-39). Must return a
 * device pointer aligned to >= 256 bytes that stays valid until the matching backward. */
typedef void* (*mrgs_alloc_fn)(void* ctx, size_t bytes);

/* Byte offsets of the sub-arrays inside the caller-owned scratch buffers. Exposed so tests
 * can decode tile keys / sorted ids / ranges / contributor counts bit-for-bit. */
typedef struct MrgsGeomLayout {
    size_t rec;           /* float [P][16]: Tu.xyz Tw.x | Tv.xyz Tw.y | Tw.z xy.x xy.y opacity | n.xyz tau    */
    size_t cf;            /* float [P][cf_stride]: rgb(3) features(S) zero padding                       */
    size_t clamped;       /* uint8 [P]: bit c set when SH colour channel c was clamped to 0             */
    size_t tiles_touched; /* uint32[P]                                                                    */
    size_t point_offsets; /* uint32[P] inclusive prefix sum of tiles_touched IN DEPTH-SORTED surfel order    */
    size_t rect;          /* uint32[P][2]: (min.x | min.y<<16), (max.x | max.y<<16) tile rectangle        */
    size_t depth;         /* float [P]: view-space depth p_view.z (the low 32 key bits)                   */
    size_t bbox;          /* float [P][8]: conservative pixel bounds of the region where the surfel's alpha can
                             reach 1/255, as an octagon: (xmin,ymin,xmax,ymax) then (umin,vmin,umax,vmax) with
                             u = x+y, v = x-y; used for per-warp culling                                    */
    size_t sort_keys;     /* uint32[2][P] ping-pong depth keys (depth bits, 0xffffffff = culled)          */
    size_t sort_vals;     /* uint32[2][P] ping-pong surfel ids; [0] ends up holding the depth order       */
    size_t scan_temp;     /* radix-pass histograms / scan block sums                                      */
    size_t scan_temp_bytes;
    size_t total;         /* bytes required                                                               */
    int32_t cf_stride;    /* floats per surfel in `cf` (3+S rounded up to a multiple of 4)                */
} MrgsGeomLayout;

typedef struct MrgsImageLayout {
    size_t state;     /* float/uint32 [tiles][5][256]: per tile, planes final_T, M1, M2, n_contrib, median_contrib;
                         pixel slot inside a tile = mrgs pixel order (see mrgs_tile_slot)                */
    size_t ranges;    /* uint32[tiles][2]                                                                */
    size_t total;
} MrgsImageLayout;

typedef struct MrgsBinningLayout {
    size_t point_list;           /* uint32[R] surfel ids sorted by (tile, depth, id)       */
    size_t point_list_unsorted;  /* uint32[R] ping-pong scratch                           */
    size_t keys;                 /* uint16[R] tile id of every sorted instance            */
    size_t keys_unsorted;        /* uint16[R] ping-pong scratch                           */
    size_t sort_temp;
    size_t sort_temp_bytes;
    size_t total;
} MrgsBinningLayout;

MRGS_API int mrgs_abi_version(void);
MRGS_API const char* mrgs_last_error(void);

MRGS_API int mrgs_geom_layout(int32_t P, int32_t S, MrgsGeomLayout* out);
MRGS_API int mrgs_image_layout(int32_t width, int32_t height, MrgsImageLayout* out);
MRGS_API int mrgs_binning_layout(int64_t R, MrgsBinningLayout* out);
MRGS_API size_t mrgs_geom_bytes(int32_t P, int32_t S);
MRGS_API size_t mrgs_image_bytes(int32_t width, int32_t height);
MRGS_API size_t mrgs_binning_bytes(int64_t R);

/* Slot (0..255) of pixel (x,y) inside its 16x16 tile in the image-state planes. */
MRGS_API int mrgs_tile_slot(int32_t x_in_tile, int32_t y_in_tile);

typedef struct MrgsForwardArgs {
    int32_t P;              /* number of surfels                                        */
    int32_t S;              /* extra feature channels, 0..MRGS_MAX_FEATURES             */
    int32_t sh_degree;      /* active SH degree D                                       */
    int32_t sh_coeffs;      /* M: coefficients per channel in `shs` (0 if no SH)        */
    int32_t width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t prefiltered;
    int32_t debug;          /* !=0: synchronise + check after every launch              */
    const float* background;       /* [3]                                               */
    const float* means3D;          /* [P,3]                                             */
    const float* shs;              /* [P,M,3]  or NULL                                  */
    const float* colors_precomp;   /* [P,3]    or NULL (exactly one of shs / this)      */
    const float* features;         /* [P,S]    or NULL when S == 0                      */
    const float* opacities;        /* [P]                                               */
    const float* scales;           /* [P,2]    or NULL                                  */
    const float* rotations;        /* [P,4]    or NULL                                  */
    const float* transMat_precomp; /* [P,9]    or NULL (exactly one of scale+rot / this)*/
    const float* viewmatrix;       /* [16]                                              */
    const float* projmatrix;       /* [16]                                              */
    const float* campos;           /* [3]                                               */
    float* out_color;              /* [3,H,W]                                           */
    float* out_feature;            /* [S,H,W]                                           */
    float* out_others;             /* [7,H,W]                                           */
    int32_t* radii;                /* [P]                                               */
    void* geom_buffer;   size_t geom_bytes;
    void* image_buffer;  size_t image_bytes;
    mrgs_alloc_fn binning_alloc;   void* binning_ctx;
    /* Optional, optimistic binning. binning_scratch is a device buffer of at least
     * mrgs_binning_bytes(binning_capacity) bytes the caller guesses to be large enough (e.g. 1.25 x the
     * previous frame's R). With it the call enqueues instance expansion, tile sort, ranges and the blend
     * BEFORE it waits for R (the kernels read the count from device memory), so the host wait overlaps
     * queued GPU work instead of draining the stream. If R > binning_capacity the call falls back to
     * binning_alloc and redoes those stages; results are identical either way. 0 / NULL = exact path only. */
    void* binning_scratch;  size_t binning_scratch_bytes;  int64_t binning_capacity;
    /* results */
    int32_t num_rendered;          /* R, also the reference's first return value        */
    void* binning_buffer;          /* binning_scratch, or what binning_alloc returned (NULL if R == 0) */
    int64_t binning_capacity_used; /* instance count binning_buffer is laid out for (mrgs_binning_layout) */
    /* Capture-safe mode (CUDA-graph capture of a whole view): with no_wait != 0 the call needs the optimistic buffer
     * above, enqueues everything, NEVER waits and never calls binning_alloc; num_rendered then returns
     * binning_capacity (the layout key for the backward) and the real R is copied to *count_out (optional, PINNED host
     * int32) by the stream - the caller checks R <= binning_capacity after the stream (or a replay of the graph) has
     * run; if that fails the frame's results are invalid and must be redone with a larger capacity. */
    int32_t no_wait;
    int32_t* count_out;
} MrgsForwardArgs;

typedef struct MrgsBackwardArgs {
    int32_t P, S, sh_degree, sh_coeffs, width, height;
    float tan_fovx, tan_fovy, scale_modifier;
    int32_t debug;
    int32_t num_rendered;          /* R from the forward                                */
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* features;
    const float* scales;
    const float* rotations;
    const float* transMat_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    const int32_t* radii;
    const void* geom_buffer;
    const void* binning_buffer;
    const void* image_buffer;
    const float* dL_dout_color;    /* [3,H,W]                                           */
    const float* dL_dout_feature;  /* [S,H,W]                                           */
    const float* dL_dout_others;   /* [7,H,W]                                           */
    /* outputs: every element is written by the call (no pre-zeroing needed) */
    float* dL_dmeans2D;   /* [P,3]  densification proxy in .xy, 0 in .z                 */
    float* dL_dcolors;    /* [P,3]  or NULL when not wanted                              */
    float* dL_dfeatures;  /* [P,S]                                                      */
    float* dL_dopacity;   /* [P]                                                        */
    float* dL_dmeans3D;   /* [P,3]                                                      */
    float* dL_dtransMat;  /* [P,9]  or NULL when not wanted                              */
    float* dL_dsh;        /* [P,M,3]                                                    */
    float* dL_dscales;    /* [P,2]                                                      */
    float* dL_drotations; /* [P,4]                                                      */
    /* scratch: raw per-surfel gradient arena, >= mrgs_grad_arena_bytes(P,S) */
    void* grad_arena;     size_t grad_arena_bytes;
    /* != 0: gradient accumulation fused into the per-surfel backward. dL_dmeans3D, dL_dsh, dL_dfeatures,
     * dL_dopacity, dL_dscales, dL_drotations (and dL_dcolors / dL_dtransMat when given) are ADDED to, rows of
     * culled surfels are left untouched; dL_dmeans2D (the per-view densification proxy) is still overwritten.
     * Lets a view batch accumulate straight into the buffer that is all-reduced. Needs M <= 16. */
    int32_t accumulate;
} MrgsBackwardArgs;

MRGS_API size_t mrgs_grad_arena_bytes(int32_t P, int32_t S);
/* floats per surfel in the raw gradient arena and the field offsets inside one row */
MRGS_API int32_t mrgs_grad_arena_stride(int32_t S);

/* Optional per-stage timing: when enabled, every entry point brackets each pipeline stage with
 * CUDA events on the launching stream. mrgs_profile_read() waits for the pending events and
 * returns accumulated milliseconds and call counts per stage since the last reset; it returns the
 * number of stages. mrgs_launch_count() = kernels of THIS library launched since the last reset
 * (CUB launches inside the scan / sort stages are not counted). */
#define MRGS_STAGE_PREPROCESS_FWD 0
#define MRGS_STAGE_SCAN 1
#define MRGS_STAGE_DUPLICATE 2
#define MRGS_STAGE_SORT 3
#define MRGS_STAGE_RANGES 4
#define MRGS_STAGE_RENDER_FWD 5
#define MRGS_STAGE_RENDER_BWD 6
#define MRGS_STAGE_PREPROCESS_BWD 7
#define MRGS_STAGE_SHADE_FWD 8
#define MRGS_STAGE_SHADE_BWD 9
#define MRGS_STAGE_CUBEMAP 10
#define MRGS_STAGE_DEPTH_SORT 11
#define MRGS_STAGE_PREFILTER_FWD 12
#define MRGS_STAGE_PREFILTER_BWD 13
#define MRGS_STAGE_COUNT 14
MRGS_API void mrgs_profile_enable(int32_t on);
MRGS_API void mrgs_profile_reset(void);
MRGS_API int mrgs_profile_read(double* ms, int64_t* calls, int32_t n);
MRGS_API int64_t mrgs_launch_count(void);

/* ---- fused deferred split-sum shading ---------------------------------------------------------
 * One launch replaces get_specular_color_surfel (utils/refl_utils.py:364-419: camera rays, reflection,
 * FG-LUT fetch, EnvLight mip query, specular weight) and the compositing of render_surfel
 * (gaussian_renderer/__init__.py:372-376, :419-420, :433-445: normal to world space, /alpha,
 * (1-refl)*base + specular, optional sRGB, + bg*(1-alpha)).
 *   features planes: 0 refl strength, 1 roughness, 2-4 albedo (more planes may follow, ignored)
 *   allmap planes:   1 alpha, 2-4 view-space normal (auxiliary.h:25-29)
 *   levels[l]:       prefiltered cubemap level l, float [6][res>>l][res>>l][3] in LOGIT space
 *                    (the fetch is followed by a sigmoid, scene/light.py:129)
 *   lut:             float [256][256][2] split-sum DFG table (assets/bsdf_256_256.bin)
 *   ray_matrix:      row-major 3x3 M with ray_dir = normalize(M * (x, y, 1)) for integer pixel (x,y)
 *   normal_matrix:   row-major 3x3 Q with n_world = Q * n_view  (world_view_transform[:3,:3])
 * Texture semantics follow nvdiffrast's dr.texture (linear clamp for the LUT, seamless
 * linear-mipmap-linear cube fetch with mip level = per-pixel bias). */
#define MRGS_MAX_MIP_LEVELS 12
typedef struct MrgsShadeArgs {
    int32_t width, height;
    int32_t num_levels;            /* L                                                    */
    int32_t base_res;              /* resolution of level 0                                */
    int32_t srgb;                  /* apply linear_to_srgb to the final image              */
    float min_roughness, max_roughness;
    float ray_matrix[9];
    float normal_matrix[9];
    const float* background;       /* [3] device pointer                                   */
    const float* base_color;       /* [3,H,W]                                              */
    const float* features;         /* [>=5,H,W]                                            */
    const float* allmap;           /* [7,H,W]                                              */
    const float* lut;
    const float* levels[MRGS_MAX_MIP_LEVELS];
    /* forward outputs, each [3,H,W]; any may be NULL */
    float* out_final;
    float* out_specular;
    float* out_direct;             /* direct_light                                         */
    float* out_normal;             /* world-space rend_normal                              */
    float* out_diffuse;            /* (1-refl)*base                                        */
    /* backward inputs ([3,H,W], NULL = zero) */
    const float* dL_dfinal;
    const float* dL_dspecular;
    const float* dL_ddiffuse;
    const float* dL_dnormal;
    /* backward outputs: fully written planes */
    float* dL_dbase_color;         /* [3,H,W]                                              */
    float* dL_dfeatures;           /* [>=5,H,W]: planes 0..4 written                       */
    float* dL_dallmap;             /* [7,H,W]:   planes 1..4 written                       */
    float* dL_dlevels[MRGS_MAX_MIP_LEVELS]; /* float [6][res>>l][res>>l][4] (rgb + pad, 16-byte texels for
                                               vector atomics); accumulated, zero them first  */
} MrgsShadeArgs;

MRGS_API int mrgs_shade_forward(const MrgsShadeArgs* args, void* stream);
MRGS_API int mrgs_shade_backward(const MrgsShadeArgs* args, void* stream);

/* EnvLight.__call__ (scene/light.py:98-129) for arbitrary directions: out[i] = sigmoid(cube fetch of
 * dirs[i] at mip level get_mip(roughness[i])) (mode "specular"), or a plain bilinear fetch of
 * levels[0] when roughness == NULL (modes "diffuse"/"pure_env"; the chain may then hold one level).
 * n directions, out [n,3]. */
MRGS_API int mrgs_envlight_query(const MrgsShadeArgs* chain, int64_t n, const float* dirs,
                                 const float* roughness, float* out, void* stream);
/* Backward of mrgs_envlight_query (nvdiffrast's dr.texture backward + the sigmoid, scene/light.py:108-129): with
 * dL_dout [n,3], texel gradients are ADDED into chain->dL_dlevels[l] (float4 per texel, rgb + pad, like
 * mrgs_shade_backward; a NULL level is skipped), dL_ddirs [n,3] and dL_droughness [n] are written when non-NULL.
 * With roughness == NULL the chain may hold a single level (modes "diffuse" / "pure_env"). */
MRGS_API int mrgs_envlight_query_backward(const MrgsShadeArgs* chain, int64_t n, const float* dirs,
                                          const float* roughness, const float* dL_dout, float* dL_ddirs,
                                          float* dL_droughness, void* stream);

/* Per-SURFEL split-sum colours of the volume-rendering stage: get_full_color_volume / get_full_color_volume_indirect
 * (utils/refl_utils.py:426-490, visibility = 1) as one kernel pair. For every surfel i:
 *   w_o = safe_normalize(campos - xyz); NdotV = w_o . n; rr = safe_normalize(2 n NdotV - w_o)
 *   diffuse      = sigmoid(cube fetch of diffuse_map in direction n) * (1 - refl) * albedo
 *   direct_light = sigmoid(cube fetch of the chain in direction rr at get_mip(roughness))
 *   specular     = direct_light * ((0.04 (1 - refl) + albedo refl) * fg[0] + fg[1])
 * `fg` (device [2]) is ONE pair for all surfels: the reference indexes `fg[0]` on the [N,2] LUT result, i.e. the FIRST
 * surfel's pair (:445, :481) — the caller evaluates it; the backward returns its gradient as the sum dL_dfg [2]
 * (ACCUMULATED, zero it first). `chain` supplies levels / num_levels / base_res / min,max roughness / dL_dlevels
 * (float4 texels, accumulated). dL_ddiffuse_map is float4 per texel, accumulated. Outputs [P,3] (direct_light may be
 * NULL); backward inputs NULL = zero; backward outputs are fully written when non-NULL. */
typedef struct MrgsSurfelShadeArgs {
    int32_t P;
    int32_t diffuse_res;
    MrgsShadeArgs chain;
    const float* diffuse_map;      /* [6,diffuse_res,diffuse_res,3] */
    const float* campos;           /* device [3] */
    const float* fg;               /* device [2] */
    const float *xyz, *normals, *albedo, *refl_strength, *roughness;
    float *diffuse, *specular, *direct_light;
    const float *dL_ddiffuse, *dL_dspecular, *dL_ddirect;
    float *dL_dxyz, *dL_dnormals, *dL_dalbedo, *dL_drefl_strength, *dL_droughness;
    float* dL_ddiffuse_map;
    float* dL_dfg;
} MrgsSurfelShadeArgs;
MRGS_API int mrgs_surfel_shade_forward(const MrgsSurfelShadeArgs* args, void* stream);
MRGS_API int mrgs_surfel_shade_backward(const MrgsSurfelShadeArgs* args, void* stream);

/* ---- pseudo surface depth + depth_to_normal ------------------------------------------------------
 * compute_2dgs_normal_and_regularizations (gaussian_renderer/__init__.py:50-78) + depth_to_normal
 * (utils/point_utils.py:9-37) in one launch:
 *   surf_depth  = (1-ratio) * nan_to_num(allmap[0]/allmap[1]) + ratio * nan_to_num(allmap[5])     [1,H,W]
 *   surf_normal = normalize(cross(P[y+1]-P[y-1], P[x+1]-P[x-1])) * allmap[1] (detached), 0 on the border,
 *                 with world points P(x,y) = surf_depth * (ray_matrix * (x,y,1)) + origin             [3,H,W]
 * ray_matrix (row-major 3x3) = c2w rotation * inverse pinhole intrinsics with principal point (W/2,H/2).
 * Backward: dL_dallmap planes 0 and 5 are written, plane 1 is ACCUMULATED (the shading backward owns it). */
MRGS_API int mrgs_depth_normal_forward(int32_t width, int32_t height, float depth_ratio, const float* host_ray_matrix,
                                       const float* host_origin, const float* allmap, float* surf_depth,
                                       float* surf_normal, void* stream);
MRGS_API int mrgs_depth_normal_backward(int32_t width, int32_t height, float depth_ratio, const float* host_ray_matrix,
                                        const float* host_origin, const float* allmap, const float* dL_dsurf_depth,
                                        const float* dL_dsurf_normal, float* dL_dallmap, void* stream);

/* ---- EnvLight.build_mips on the device ----------------------------------------------------------
 * Cubemaps are float [6][res][res][C]. Replaces cubemap_mip (scene/light_utils.py:66-80) and the
 * renderutils_plugin ops diffuse_cubemap_fwd/bwd, specular_bounds, specular_cubemap_fwd/bwd
 * (scene/renderutils/c_src/torch_bindings.cpp:740-890, kernels in c_src/cubemap.cu:110-354).
 *   mip forward : out[res/2] = 2x2 average of in[res]
 *   mip backward: din[res] = seamless bilinear cube fetch of 0.25*dout[res/2] at the fine texel
 *                 directions (the reference's non-adjoint backward), C must be 3
 *   bounds      : int32 [6][res][res][6][4] = (xmin,xmax,ymin,ymax) of the texels of each source face
 *                 inside the cone dot(L,N) >= costheta_cutoff
 *   specular fwd: out4 [6][res][res][4] = (sum w*rgb, sum w); the caller divides rgb by sum w
 *   specular bwd: dcubemap [6][res][res][3] is zeroed and then accumulated from dout4's rgb
 *   diffuse     : cosine-weighted convolution over the whole cube (16x16 level in the reference) */
MRGS_API int mrgs_cubemap_mip_forward(const float* in, float* out, int32_t res_in, int32_t channels, void* stream);
MRGS_API int mrgs_cubemap_mip_backward(const float* dout, float* din, int32_t res_out, void* stream);
MRGS_API int mrgs_specular_bounds(int32_t res, float costheta_cutoff, int32_t* bounds, void* stream);
MRGS_API int mrgs_specular_cubemap_forward(const float* cubemap, const int32_t* bounds, int32_t res, float roughness,
                                           float costheta_cutoff, float* out4, void* stream);
MRGS_API int mrgs_specular_cubemap_backward(const float* cubemap, const int32_t* bounds, int32_t res, float roughness,
                                            float costheta_cutoff, const float* dout4, float* dcubemap, void* stream);
MRGS_API int mrgs_diffuse_cubemap_forward(const float* cubemap, int32_t res, float* out, void* stream);
MRGS_API int mrgs_diffuse_cubemap_backward(const float* cubemap, int32_t res, const float* dout, float* dcubemap,
                                           void* stream);

/* ---- EnvLight.build_mips as a precomputed sparse operator ("prefilter plan") --------------------------
 * The reference runs build_mips() every training iteration (train_refnerf.py:1155-1163, scene/light.py:72-86):
 * cubemap_mip down to min_res, ru.diffuse_cubemap on the smallest level and ru.specular_cubemap(level, roughness,
 * cutoff) on every level (scene/renderutils/ops.py:391-458, kernels c_src/cubemap.cu:110-354). For a fixed
 * (res, roughness, cutoff) those two ops are fixed linear maps of the cubemap; a plan stores one of them (or its
 * transpose, for the backward) as per-destination-texel tap lists with the tap weights the reference would compute
 * (same expression order, same loop domain = its cached bounds incl. the 16x16 tile culling), already divided by
 * the reference-order weight sum. Applying a plan is a gather: no atomics, no zero fill, every weight read once.
 *
 * A plan is built in two passes over caller-owned memory:
 *   1. mrgs_prefilter_plan_count: per patch (one warp's destination texels, see below) the number of source-row
 *      segments, weight rows and taps; for MRGS_PREFILTER_SPECULAR also wsum[texel] (the un-normalised weight sum
 *      = channel 3 of specular_cubemap_fwd's output) unless wsum is NULL (counts only: used to pick the patch shape).
 *   2. the caller turns the counts into exclusive prefix sums (patch_seg_begin / patch_slot_begin, patches+1
 *      entries), allocates seg_desc [segments][2] int32, spans [segments][32] uint16, weights
 *      [rows][rows_per_lane][32] float and calls mrgs_prefilter_plan_fill.
 * kinds: _SPECULAR   dst = GGX-prefiltered level, src = raw level            (specular_cubemap fwd, normalised)
 *        _SPECULAR_T dst = d(raw level),          src = d(prefiltered level) (specular_cubemap bwd incl. the
 *                    division by wsum); needs wsum and bounds of every texel; full_search = 1 searches the whole
 *                    cube for every texel (needed when the reference's tile culling is not conservative: the caller
 *                    compares the tap totals of both orientations)
 *        _DIFFUSE / _DIFFUSE_T  the cosine convolution and its transpose (no bounds, no normalisation)
 * Patch shape: patch_width (32, 16 or 8 lanes) x (32 / patch_width * rows_per_lane) texels of one face, a lane
 * owning rows_per_lane (1 or 2) vertically adjacent texels; (32, 1) is the linear layout (32 consecutive texels in
 * memory order) and works for any res, the others need res divisible by the block's width and height. Padded
 * weight slots (a segment is as long as its longest lane) depend on the shape: callers run the count pass per shape
 * and keep the smallest. */
#define MRGS_PREFILTER_SPECULAR 0
#define MRGS_PREFILTER_SPECULAR_T 1
#define MRGS_PREFILTER_DIFFUSE 2
#define MRGS_PREFILTER_DIFFUSE_T 3
#define MRGS_PREFILTER_MAX_JOBS 8

typedef struct MrgsPrefilterPlan {
    int32_t res;
    int32_t rows_per_lane;
    int32_t patch_width;
    int32_t* patch_seg_begin;
    int32_t* patch_slot_begin;
    int32_t* seg_desc;
    uint16_t* spans;
    float* weights;
} MrgsPrefilterPlan;

typedef struct MrgsPrefilterBuildArgs {
    int32_t kind;
    int32_t res;
    int32_t rows_per_lane;
    int32_t patch_width;
    int32_t full_search;
    float roughness;
    float costheta_cutoff;
    const float* texel_table;   /* [6][res][res][4] from mrgs_prefilter_texel_table */
    const int32_t* bounds;      /* [6][res][res][6][4] from mrgs_specular_bounds (specular kinds) */
    float* wsum;                /* [6][res][res]: written by the count pass of _SPECULAR, read otherwise */
    int32_t* seg_count;         /* count pass outputs, one entry per patch */
    int32_t* slot_count;
    int32_t* tap_count;
    MrgsPrefilterPlan plan;     /* fill pass: prefix sums in, seg_desc / spans / weights out */
} MrgsPrefilterBuildArgs;

/* one gather: dst[texel] = sum over the texel's taps of weight * src[tap]; rgb at src/dst_stride floats per texel
 * (3 or 4). nan_where_zero (optional, [texels]): destination texels whose entry is 0 receive NaN, which is what
 * the reference's rgb / wsum yields for an empty cone. */
typedef struct MrgsPrefilterJob {
    MrgsPrefilterPlan plan;
    const float* src;
    float* dst;
    const float* nan_where_zero;
    int32_t src_stride;
    int32_t dst_stride;
    /* patches [patch_begin, patch_end) of the plan only (0, 0 = all): a rank of a view-sharded step applies its share
     * of every level into a zero-filled buffer and the ranks sum the buffers (materialrefgs_b200/prefilter.py) */
    int32_t patch_begin;
    int32_t patch_end;
} MrgsPrefilterJob;

/* number of patches (= warps of the gather), or -1 when the shape does not tile a res x res face */
MRGS_API int32_t mrgs_prefilter_patch_count(int32_t res, int32_t rows_per_lane, int32_t patch_width);
MRGS_API int mrgs_prefilter_texel_table(int32_t res, float* table, void* stream);
MRGS_API int mrgs_prefilter_plan_count(const MrgsPrefilterBuildArgs* args, void* stream);
MRGS_API int mrgs_prefilter_plan_fill(const MrgsPrefilterBuildArgs* args, void* stream);
/* All jobs (at most MRGS_PREFILTER_MAX_JOBS: the levels of a chain + its diffuse map) in ONE launch, in the order
 * given (put the jobs with the longest tap lists first). backward != 0 only selects the profiling stage.
 * max_ctas > 0 caps the grid (the warps then stride over the patches): the gather is HBM-bound and leaves most issue
 * slots idle, the tile-blend kernels are the opposite - with e.g. one CTA per SM on a second stream the gather runs in
 * their background instead of in front of them. 0 = one warp per patch. */
MRGS_API int mrgs_prefilter_apply(const MrgsPrefilterJob* jobs, int32_t num_jobs, int32_t backward, int32_t max_ctas,
                                  void* stream);
/* base [6][res][res][3] -> levels4[0] = float4-padded copy of it, levels4[l] = [6][res>>l][res>>l][4] 2x2 averages
 * (cubemap_mip forward, scene/light_utils.py:69-71, for the whole chain in one launch per 5 levels).
 * levels4 is a HOST array of num_levels device pointers; res must be divisible by 2^(num_levels-1). */
MRGS_API int mrgs_mip_pyramid_forward(const float* base, int32_t res, int32_t num_levels, float* const* levels4,
                                      void* stream);
/* grads3[l] ([6][res>>l][res>>l][3], HOST array of device pointers) holds d(raw level l) from the level's own
 * prefilter; on return grads3[l] += cubemap_mip backward (the reference's non-adjoint bilinear fetch,
 * scene/light_utils.py:72-80) of the completed grads3[l+1], from the coarsest level down, so grads3[0] is the
 * gradient of the base cubemap. extra_last3 (optional) is added to the coarsest level first (the diffuse map's
 * gradient, which the reference's autograd sums into the same tensor). */
MRGS_API int mrgs_mip_chain_backward(int32_t res, int32_t num_levels, float* const* grads3, const float* extra_last3,
                                     void* stream);

/* Stage timing under CUDA-graph replay: stage scopes that run while their stream is being captured record their
 * events as external-event nodes, so every replay re-records them; this call waits for the captured stages' end
 * events and adds one sample each (call it after a replay has been enqueued and before the next one is). */
MRGS_API void mrgs_profile_collect_captured(void);

MRGS_API int mrgs_forward(MrgsForwardArgs* args, void* stream);
MRGS_API int mrgs_backward(const MrgsBackwardArgs* args, void* stream);
MRGS_API int mrgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, void* stream);

/* Per-surfel feature preparation in front of the rasterizer, replacing the ~20 eager
 * torch kernels of gaussian_renderer/__init__.py:259-266, :334-353 with scene/gaussian_model.py:236-303:
 *   scales    = exp(scaling)                                   [P,2]
 *   rotations = rotation / max(|rotation|, 1e-12)              [P,4]
 *   opacities = sigmoid(opacity)                               [P,1]
 *   features  = ( sigmoid(refl_strength), sigmoid(roughness), sigmoid(ori_color)[3],
 *                 clamp_min(eval_sh(3, cat(indirect_dc, indirect_rest), reflection), 0)[3] )   [P,8]
 * with reflection = 2 (n.w_o) n - w_o, w_o = -(xyz - campos)/|.|, n = the surfel normal (third column of
 * R(rotation/|rotation|)) flipped towards the camera (flip_align_view) and safe-normalised.
 * The backward takes the gradients of the four outputs and writes those of the nine raw parameters
 * (every element is written). All pointers are device pointers of contiguous fp32; campos is [3];
 * indirect_dc is [P,1,3], indirect_rest [P,15,3]. The forward ignores the dL_* members. */
typedef struct MrgsSurfelFeatureArgs {
    int32_t P;
    const float* campos;
    const float *xyz, *scaling, *rotation, *opacity, *refl_strength, *roughness, *ori_color, *indirect_dc,
        *indirect_rest;
    float *scales, *rotations, *opacities, *features;                                  /* forward outputs  */
    const float *dL_dscales, *dL_drotations, *dL_dopacities, *dL_dfeatures;            /* backward inputs  */
    float *dL_dxyz, *dL_dscaling, *dL_drotation, *dL_dopacity, *dL_drefl_strength, *dL_droughness, *dL_dori_color,
        *dL_dindirect_dc, *dL_dindirect_rest;                                          /* backward outputs */
} MrgsSurfelFeatureArgs;
MRGS_API int mrgs_surfel_features_forward(const MrgsSurfelFeatureArgs* args, void* stream);
MRGS_API int mrgs_surfel_features_backward(const MrgsSurfelFeatureArgs* args, void* stream);

/* Photometric loss terms of calculate_loss (utils/loss_utils.py:155-157) on [C,H,W] images:
 *   out2[0] = mean |img - gt|                                  (l1_loss, loss_utils.py:22-23)
 *   out2[1] = mean SSIM map, 11x11 Gaussian window sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2
 *             (ssim / _ssim, loss_utils.py:83-119, size_average = True)
 * forward : `maps` ([3][C][H][W], may be NULL when no gradient is wanted) receives the three per-pixel SSIM
 *           derivative maps the backward filters; `partials` is scratch of mrgs_photometric_partials_bytes().
 * backward: upstream = device [2] = (dL/d out2[0], dL/d out2[1]); dimg [C][H][W] is written (gt has no gradient).
 * Sums are formed in a fixed order (deterministic). */
MRGS_API size_t mrgs_photometric_partials_bytes(int32_t channels, int32_t height, int32_t width);
MRGS_API int mrgs_photometric_forward(const float* img, const float* gt, int32_t channels, int32_t height, int32_t width,
                                      float* maps, float* partials, float* out2, void* stream);
MRGS_API int mrgs_photometric_backward(const float* img, const float* gt, const float* maps, int32_t channels,
                                       int32_t height, int32_t width, const float* upstream, float* dimg, void* stream);

/* Geometric regularisers of calculate_loss (utils/loss_utils.py:160-197) on one view's maps, ONE
 * forward and ONE backward kernel. `terms` selects what is evaluated (members of the other terms may be NULL):
 *   MRGS_GEOM_NORMAL         out4[0] = mean_{HW}( image_weight * sum_c |surf_normal_c - rend_normal_c| )   (:169)
 *                                      or, with image_weight == NULL, mean_{HW}( 1 - sum_c rend_normal_c surf_normal_c ) (:171-172)
 *   MRGS_GEOM_DIST           out4[1] = mean_{HW}( rend_dist )                                               (:177)
 *   MRGS_GEOM_NORMAL_SMOOTH  out4[2] = first_order_edge_aware_loss(rend_normal, gt_image)                   (:121-122, :183)
 *   MRGS_GEOM_DEPTH_SMOOTH   out4[3] = first_order_edge_aware_loss(surf_depth,  gt_image)                   (:191)
 * first_order_edge_aware_loss(data, img) = mean_{3HW}( sum_{d in x,y} |grad_d data| * exp(-|grad_d img|) ) with
 * kornia 0.7.3's spatial_gradient (normalised 3x3 Sobel, replicate padding). Terms not selected give 0.
 * forward : `coef` ([8][H][W] scratch, may be NULL when no gradient is wanted) keeps sign(grad data) * exp(-|grad img|);
 *           `partials` is scratch of mrgs_geometry_loss_partials_bytes(); out4 is device [4].
 * backward: upstream = device [4] = dL/d out4; every non-NULL dL_d* map is written in full (zeros where a term is off).
 *           gt_image and image_weight have no gradient (the reference detaches them). Deterministic. */
#define MRGS_GEOM_NORMAL 1u
#define MRGS_GEOM_DIST 2u
#define MRGS_GEOM_NORMAL_SMOOTH 4u
#define MRGS_GEOM_DEPTH_SMOOTH 8u
typedef struct MrgsGeometryLossArgs {
    int32_t height, width;
    uint32_t terms;
    const float* rend_normal;   /* [3,H,W] */
    const float* surf_normal;   /* [3,H,W] */
    const float* rend_dist;     /* [1,H,W] */
    const float* surf_depth;    /* [1,H,W] */
    const float* gt_image;      /* [3,H,W] */
    const float* image_weight;  /* [H,W] or NULL */
    float* coef;                /* [8,H,W] scratch, forward writes / backward reads */
    float* partials;            /* scratch */
    float* out4;                /* forward output  */
    const float* upstream;      /* backward input  */
    float *dL_drend_normal, *dL_dsurf_normal, *dL_drend_dist, *dL_dsurf_depth;   /* backward outputs, each may be NULL */
} MrgsGeometryLossArgs;
MRGS_API size_t mrgs_geometry_loss_partials_bytes(int32_t height, int32_t width);
MRGS_API int mrgs_geometry_loss_forward(const MrgsGeometryLossArgs* args, void* stream);
MRGS_API int mrgs_geometry_loss_backward(const MrgsGeometryLossArgs* args, void* stream);

/* get_img_grad_weight (utils/loss_utils.py:127-139): per interior pixel max( mean_c |right - left|, mean_c |top - bottom| ),
 * normalised by the interior's min and max, border padded with 1. img [C,H,W] -> out [H,W]; H, W >= 3;
 * scratch8 = 8 bytes of device memory. */
MRGS_API int mrgs_img_grad_weight(const float* img, int32_t channels, int32_t height, int32_t width, float* out,
                                  void* scratch8, void* stream);

/* Densification statistics of one rendered view, one fused pass over the P surfels
 * (GaussianModel.add_densification_stats scene/gaussian_model.py:1059-1061 and the max_radii2D update
 * train_refnerf.py:1416-1418). For every surfel with radii > 0:
 *   stats[i][0] += |dL_dmeans2D[i]|         (xyz_gradient_accum; all three components like torch.norm(grad, dim=-1))
 *   stats[i][1] += 1                        (denom)
 *   max_radii[i] = max(max_radii[i], radii[i])
 * dL_dmeans2D is [P,3] (the screen-space gradient the backward returns), stats [P,2], max_radii int32 [P]. */
MRGS_API int mrgs_densify_stats(int32_t P, const float* dL_dmeans2D, const int32_t* radii, float* stats,
                                int32_t* max_radii, void* stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* MRGS_H_INCLUDED */
