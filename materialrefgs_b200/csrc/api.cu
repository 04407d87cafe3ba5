// api.cu — extern "C" entry points of libmrgs.so (see include/mrgs.h for the contract and the
// reference interface each one replaces).
#include <stdarg.h>
#include <string.h>

#include "kernels.cuh"

namespace mrgs {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what, cudaStream_t stream, bool debug) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MRGS_ERR_CUDA;
    }
    return MRGS_OK;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// pinned host slot for the one scalar that has to reach the host (R)
static int32_t* pinned_slot() {
    static thread_local int32_t* slot = nullptr;
    if (slot == nullptr) {
        if (cudaHostAlloc((void**)&slot, 64, cudaHostAllocDefault) != cudaSuccess) slot = nullptr;
    }
    return slot;
}

static cudaEvent_t r_ready_event() {
    // one event per (thread, device): an event can only be recorded on a stream of the device it was created on
    constexpr int kMaxDevices = 64;
    static thread_local cudaEvent_t ev[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    if (ev[dev] == nullptr) {
        if (cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) ev[dev] = nullptr;
    }
    return ev[dev];
}

// ---- optional per-stage timing (CUDA events on the launching stream) + own-kernel launch count
struct Profiler {
    bool enabled = false;
    cudaEvent_t ev[MRGS_STAGE_COUNT][2];
    bool created = false;
    bool pending[MRGS_STAGE_COUNT] = {};
    bool captured[MRGS_STAGE_COUNT] = {};   // recorded inside a stream capture: re-recorded by every graph replay
    double ms[MRGS_STAGE_COUNT] = {};
    long long calls[MRGS_STAGE_COUNT] = {};
    long long launches = 0;
};
static Profiler g_prof;

static void prof_collect_one(int s) {
    if (!g_prof.pending[s]) return;
    float t = 0.f;
    if (cudaEventSynchronize(g_prof.ev[s][1]) == cudaSuccess &&
        cudaEventElapsedTime(&t, g_prof.ev[s][0], g_prof.ev[s][1]) == cudaSuccess) {
        g_prof.ms[s] += t;
        g_prof.calls[s] += 1;
    }
    g_prof.pending[s] = false;
}

static void prof_collect() {
    for (int s = 0; s < MRGS_STAGE_COUNT; ++s) prof_collect_one(s);
}

struct StageScope {
    int stage;
    cudaStream_t stream;
    StageScope(int s, cudaStream_t st, int own_launches) : stage(s), stream(st) {
        g_prof.launches += own_launches;
        if (!g_prof.enabled) return;
        if (!g_prof.created) {
            for (int i = 0; i < MRGS_STAGE_COUNT; ++i) {
                cudaEventCreate(&g_prof.ev[i][0]);
                cudaEventCreate(&g_prof.ev[i][1]);
            }
            g_prof.created = true;
        }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        capturing = cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
        if (capturing) {   // an external-event node: every replay of the graph records it again
            g_prof.pending[stage] = false;
            cudaEventRecordWithFlags(g_prof.ev[stage][0], stream, cudaEventRecordExternal);
            return;
        }
        // one in-flight measurement per stage: only THIS stage's previous pair is waited for (it finished a
        // whole step ago), never the still-running tail of the previous step
        prof_collect_one(stage);
        cudaEventRecord(g_prof.ev[stage][0], stream);
    }
    ~StageScope() {
        if (!g_prof.enabled) return;
        if (capturing) {
            cudaEventRecordWithFlags(g_prof.ev[stage][1], stream, cudaEventRecordExternal);
            g_prof.captured[stage] = true;
            return;
        }
        cudaEventRecord(g_prof.ev[stage][1], stream);
        g_prof.pending[stage] = true;
    }
    bool capturing = false;
};

}  // namespace mrgs

using namespace mrgs;

extern "C" {

int mrgs_abi_version(void) { return MRGS_ABI_VERSION; }
const char* mrgs_last_error(void) { return g_error; }
int mrgs_tile_slot(int32_t x, int32_t y) { return slot_of(x, y); }
int32_t mrgs_grad_arena_stride(int32_t S) { return grad_stride(S); }
size_t mrgs_grad_arena_bytes(int32_t P, int32_t S) {
    return align_up((size_t)P * grad_stride(S) * sizeof(float));
}

void mrgs_profile_enable(int32_t on) { g_prof.enabled = on != 0; }
void mrgs_profile_collect_captured(void) {
    for (int s = 0; s < MRGS_STAGE_COUNT; ++s) {
        if (!g_prof.captured[s]) continue;
        float t = 0.f;
        if (cudaEventSynchronize(g_prof.ev[s][1]) == cudaSuccess &&
            cudaEventElapsedTime(&t, g_prof.ev[s][0], g_prof.ev[s][1]) == cudaSuccess && t >= 0.f) {
            g_prof.ms[s] += t;
            g_prof.calls[s] += 1;
        }
    }
}
void mrgs_profile_reset(void) {
    prof_collect();
    for (int s = 0; s < MRGS_STAGE_COUNT; ++s) {
        g_prof.ms[s] = 0.0;
        g_prof.calls[s] = 0;
    }
    g_prof.launches = 0;
}
int mrgs_profile_read(double* ms, int64_t* calls, int32_t n) {
    prof_collect();
    for (int s = 0; s < n && s < MRGS_STAGE_COUNT; ++s) {
        if (ms) ms[s] = g_prof.ms[s];
        if (calls) calls[s] = g_prof.calls[s];
    }
    return MRGS_STAGE_COUNT;
}
int64_t mrgs_launch_count(void) { return g_prof.launches; }

int mrgs_geom_layout(int32_t P, int32_t S, MrgsGeomLayout* out) {
    if (out == nullptr || P < 0 || S < 0 || S > MRGS_MAX_FEATURES) {
        set_error("mrgs_geom_layout: bad arguments (P=%d, S=%d)", P, S);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const size_t n = (size_t)P;
    size_t off = 0;
    out->cf_stride = cf_stride(S);
    out->rec = off;           off = align_up(off + n * kGeomFloats * sizeof(float));
    out->cf = off;            off = align_up(off + n * out->cf_stride * sizeof(float));
    out->clamped = off;       off = align_up(off + n);
    out->tiles_touched = off; off = align_up(off + n * sizeof(uint32_t));
    out->point_offsets = off; off = align_up(off + n * sizeof(uint32_t));
    out->rect = off;          off = align_up(off + n * sizeof(uint2));
    out->depth = off;         off = align_up(off + n * sizeof(float));
    out->bbox = off;          off = align_up(off + 2 * n * sizeof(float4));
    out->sort_keys = off;     off = align_up(off + 2 * n * sizeof(uint32_t));
    out->sort_vals = off;     off = align_up(off + 2 * n * sizeof(uint32_t));
    out->scan_temp = off;
    out->scan_temp_bytes = sort_hist_bytes(P > 0 ? P : 1) + (size_t)scan_blocks(P > 0 ? P : 1) * sizeof(uint32_t);
    off = align_up(off + out->scan_temp_bytes);
    out->total = off;
    return MRGS_OK;
}

int mrgs_image_layout(int32_t width, int32_t height, MrgsImageLayout* out) {
    if (out == nullptr || width <= 0 || height <= 0) {
        set_error("mrgs_image_layout: bad arguments (%dx%d)", width, height);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const size_t tiles = (size_t)((width + kTileX - 1) / kTileX) * ((height + kTileY - 1) / kTileY);
    size_t off = 0;
    out->state = off;  off = align_up(off + tiles * 5 * kTilePixels * sizeof(float));
    out->ranges = off; off = align_up(off + tiles * sizeof(uint2));
    out->total = off;
    return MRGS_OK;
}

int mrgs_binning_layout(int64_t R, MrgsBinningLayout* out) {
    if (out == nullptr || R < 0 || R > 0x7fffffff) {
        set_error("mrgs_binning_layout: bad instance count %lld", (long long)R);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const size_t n = (size_t)R;
    size_t off = 0;
    out->point_list = off;          off = align_up(off + n * sizeof(uint32_t));
    out->point_list_unsorted = off; off = align_up(off + n * sizeof(uint32_t));
    out->keys = off;                off = align_up(off + n * sizeof(uint16_t));
    out->keys_unsorted = off;       off = align_up(off + n * sizeof(uint16_t));
    out->sort_temp = off;
    out->sort_temp_bytes = sort_hist_bytes(R > 0 ? R : 1);
    off = align_up(off + out->sort_temp_bytes);
    out->total = off;
    return MRGS_OK;
}

size_t mrgs_geom_bytes(int32_t P, int32_t S) {
    MrgsGeomLayout l;
    return mrgs_geom_layout(P, S, &l) == MRGS_OK ? l.total : 0;
}
size_t mrgs_image_bytes(int32_t w, int32_t h) {
    MrgsImageLayout l;
    return mrgs_image_layout(w, h, &l) == MRGS_OK ? l.total : 0;
}
size_t mrgs_binning_bytes(int64_t R) {
    MrgsBinningLayout l;
    return mrgs_binning_layout(R, &l) == MRGS_OK ? l.total : 0;
}

int mrgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                      const float* projmatrix, uint8_t* present, void* stream_) {
    (void)projmatrix;  // the reference computes p_hom from it but never uses the result
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("mrgs_mark_visible: bad arguments");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return MRGS_OK;
    launch_mark_visible(P, means3D, viewmatrix, present, stream);
    MRGS_LAUNCH_OK("mark_visible", stream, false);
    return MRGS_OK;
}

namespace {
__global__ void densify_stats_kernel(int P, const float* __restrict__ g, const int32_t* __restrict__ radii,
                                     float2* __restrict__ stats, int32_t* __restrict__ max_radii) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int32_t r = radii[i];
    if (r <= 0) return;
    // torch.norm(viewspace_point_tensor.grad[filter], dim=-1): all three components (gaussian_model.py:1060); the
    // rasterizer leaves z at zero
    const float gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
    float2 s = stats[i];
    s.x += sqrtf(gx * gx + gy * gy + gz * gz);
    s.y += 1.0f;
    stats[i] = s;
    max_radii[i] = max(max_radii[i], r);
}
}  // namespace

int mrgs_surfel_features_forward(const MrgsSurfelFeatureArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int st = launch_surfel_features(a, false, stream);
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("surfel_features_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_surfel_features_backward(const MrgsSurfelFeatureArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int st = launch_surfel_features(a, true, stream);
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("surfel_features_bwd", stream, false);
    return MRGS_OK;
}

size_t mrgs_photometric_partials_bytes(int32_t C, int32_t H, int32_t W) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    return photometric_partials_count(C, H, W) * sizeof(float2);
}

int mrgs_photometric_forward(const float* img, const float* gt, int32_t C, int32_t H, int32_t W, float* maps,
                             float* partials, float* out2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535 || !img || !gt || !partials || !out2) {
        set_error("mrgs_photometric_forward: bad arguments (C=%d H=%d W=%d)", C, H, W);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    launch_photometric_fwd(img, gt, C, H, W, maps, partials, out2, stream);
    MRGS_LAUNCH_OK("photometric_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_photometric_backward(const float* img, const float* gt, const float* maps, int32_t C, int32_t H, int32_t W,
                              const float* upstream, float* dimg, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535 || !img || !gt || !maps || !upstream || !dimg) {
        set_error("mrgs_photometric_backward: bad arguments (C=%d H=%d W=%d)", C, H, W);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    launch_photometric_bwd(img, gt, maps, C, H, W, upstream, dimg, stream);
    MRGS_LAUNCH_OK("photometric_bwd", stream, false);
    return MRGS_OK;
}

size_t mrgs_geometry_loss_partials_bytes(int32_t H, int32_t W) {
    if (H <= 0 || W <= 0) return 0;
    return geometry_loss_partials_count(H, W) * sizeof(float4);
}

static bool geometry_loss_args_ok(const MrgsGeometryLossArgs* a, bool backward) {
    if (!a || a->height <= 0 || a->width <= 0) return false;
    const unsigned t = a->terms;
    if (t & ~(MRGS_GEOM_NORMAL | MRGS_GEOM_DIST | MRGS_GEOM_NORMAL_SMOOTH | MRGS_GEOM_DEPTH_SMOOTH)) return false;
    if ((t & MRGS_GEOM_NORMAL) && (!a->rend_normal || !a->surf_normal)) return false;
    if ((t & MRGS_GEOM_DIST) && !a->rend_dist) return false;
    if ((t & MRGS_GEOM_NORMAL_SMOOTH) && (!a->rend_normal || !a->gt_image)) return false;
    if ((t & MRGS_GEOM_DEPTH_SMOOTH) && (!a->surf_depth || !a->gt_image)) return false;
    if (!backward) return a->partials && a->out4;
    if ((t & (MRGS_GEOM_NORMAL_SMOOTH | MRGS_GEOM_DEPTH_SMOOTH)) && !a->coef) return false;
    return a->upstream != nullptr;
}

int mrgs_geometry_loss_forward(const MrgsGeometryLossArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!geometry_loss_args_ok(a, false)) {
        set_error("mrgs_geometry_loss_forward: bad arguments (a map of a selected term is NULL, or H/W <= 0)");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    launch_geometry_loss(a, false, stream);
    MRGS_LAUNCH_OK("geometry_loss_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_geometry_loss_backward(const MrgsGeometryLossArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!geometry_loss_args_ok(a, true)) {
        set_error("mrgs_geometry_loss_backward: bad arguments (a map of a selected term, coef or upstream is NULL)");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    launch_geometry_loss(a, true, stream);
    MRGS_LAUNCH_OK("geometry_loss_bwd", stream, false);
    return MRGS_OK;
}

int mrgs_img_grad_weight(const float* img, int32_t C, int32_t H, int32_t W, float* out, void* scratch8, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!img || !out || !scratch8 || C <= 0 || H < 3 || W < 3) {
        set_error("mrgs_img_grad_weight: bad arguments (C=%d H=%d W=%d; H, W >= 3)", C, H, W);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    launch_img_grad_weight(img, C, H, W, out, scratch8, stream);
    MRGS_LAUNCH_OK("img_grad_weight", stream, false);
    return MRGS_OK;
}

int mrgs_densify_stats(int32_t P, const float* dL_dmeans2D, const int32_t* radii, float* stats,
                       int32_t* max_radii, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || (P > 0 && (!dL_dmeans2D || !radii || !stats || !max_radii))) {
        set_error("mrgs_densify_stats: bad arguments");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return MRGS_OK;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, dL_dmeans2D, radii,
                                                              reinterpret_cast<float2*>(stats), max_radii);
    MRGS_LAUNCH_OK("densify_stats", stream, false);
    return MRGS_OK;
}

int mrgs_forward(MrgsForwardArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (a == nullptr) {
        set_error("mrgs_forward: null args");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    a->num_rendered = 0;
    a->binning_buffer = nullptr;
    a->binning_capacity_used = 0;
    if (a->P < 0 || a->width <= 0 || a->height <= 0) {
        set_error("mrgs_forward: bad sizes P=%d W=%d H=%d", a->P, a->width, a->height);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->S < 0 || a->S > MRGS_MAX_FEATURES) {
        set_error("mrgs_forward: S=%d feature channels, supported range is 0..%d", a->S,
                  MRGS_MAX_FEATURES);
        return MRGS_ERR_UNSUPPORTED;
    }
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr)) {
        set_error("mrgs_forward: provide exactly one of shs / colors_precomp");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const bool has_sr = a->scales != nullptr && a->rotations != nullptr;
    if (has_sr == (a->transMat_precomp != nullptr)) {
        set_error("mrgs_forward: provide exactly one of scales+rotations / transMat_precomp");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->shs != nullptr && a->sh_coeffs < (a->sh_degree + 1) * (a->sh_degree + 1)) {
        set_error("mrgs_forward: sh_degree %d needs %d coefficients, got %d", a->sh_degree,
                  (a->sh_degree + 1) * (a->sh_degree + 1), a->sh_coeffs);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->S > 0 && a->features == nullptr) {
        set_error("mrgs_forward: S=%d but features is null", a->S);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const bool debug = a->debug != 0;
    const int W = a->width, H = a->height;
    const int grid_x = (W + kTileX - 1) / kTileX, grid_y = (H + kTileY - 1) / kTileY;
    const int tiles = grid_x * grid_y;
    if ((long long)grid_x * grid_y > 0xffff) {
        set_error("mrgs_forward: %d x %d tiles exceed the 16-bit tile id (max 65535 tiles)", grid_x, grid_y);
        return MRGS_ERR_UNSUPPORTED;
    }

    MrgsImageLayout il;
    if (mrgs_image_layout(W, H, &il) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
    if (a->image_buffer == nullptr || a->image_bytes < il.total) {
        set_error("mrgs_forward: image buffer too small (%zu < %zu)", a->image_bytes, il.total);
        return MRGS_ERR_WORKSPACE;
    }
    char* img = (char*)a->image_buffer;
    uint2* ranges = (uint2*)(img + il.ranges);
    float* state = (float*)(img + il.state);

    RenderFwdParams rp{};
    rp.S = a->S; rp.W = W; rp.H = H; rp.grid_x = grid_x; rp.grid_y = grid_y;
    rp.cf_stride = cf_stride(a->S);
    rp.ranges = ranges;
    rp.background = a->background;
    rp.state = state;
    rp.out_color = a->out_color;
    rp.out_feature = a->out_feature;
    rp.out_others = a->out_others;

    MRGS_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), stream));

    int R = 0;
    if (a->P > 0) {
        MrgsGeomLayout gl;
        if (mrgs_geom_layout(a->P, a->S, &gl) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
        if (a->geom_buffer == nullptr || a->geom_bytes < gl.total) {
            set_error("mrgs_forward: geometry buffer too small (%zu < %zu)", a->geom_bytes, gl.total);
            return MRGS_ERR_WORKSPACE;
        }
        char* geom = (char*)a->geom_buffer;
        PreprocessParams pp{};
        pp.P = a->P; pp.S = a->S; pp.D = a->sh_degree; pp.M = a->sh_coeffs; pp.W = W; pp.H = H;
        pp.grid_x = grid_x; pp.grid_y = grid_y; pp.cf_stride = gl.cf_stride;
        pp.scale_modifier = a->scale_modifier; pp.prefiltered = a->prefiltered;
        pp.means3D = a->means3D; pp.scales = a->scales; pp.rotations = a->rotations;
        pp.opacities = a->opacities; pp.shs = a->shs; pp.colors_precomp = a->colors_precomp;
        pp.features = a->features; pp.transMat_precomp = a->transMat_precomp;
        pp.viewmatrix = a->viewmatrix; pp.projmatrix = a->projmatrix; pp.campos = a->campos;
        pp.radii = a->radii;
        pp.rec = (float*)(geom + gl.rec);
        pp.cf = (float*)(geom + gl.cf);
        pp.clamped = (uint8_t*)(geom + gl.clamped);
        pp.tiles_touched = (uint32_t*)(geom + gl.tiles_touched);
        pp.rect = (uint2*)(geom + gl.rect);
        pp.depth = (float*)(geom + gl.depth);
        pp.bbox = (float4*)(geom + gl.bbox);
        pp.sort_key = (uint32_t*)(geom + gl.sort_keys);
        uint32_t* offsets = (uint32_t*)(geom + gl.point_offsets);
        uint32_t* sort_keys_a = pp.sort_key;
        uint32_t* sort_keys_b = sort_keys_a + a->P;
        uint32_t* sort_vals_a = (uint32_t*)(geom + gl.sort_vals);
        uint32_t* sort_vals_b = sort_vals_a + a->P;
        uint32_t* hist = (uint32_t*)(geom + gl.scan_temp);
        uint32_t* block_sums = (uint32_t*)(geom + gl.scan_temp + sort_hist_bytes(a->P));

        {
            StageScope sc(MRGS_STAGE_PREPROCESS_FWD, stream, 1);
            launch_preprocess_fwd(pp, stream);
        }
        MRGS_LAUNCH_OK("preprocess_fwd", stream, debug);

        // depth order of the surfels, then instance offsets in that order
        uint32_t* order = nullptr;
        {
            int launches = 0, sort_status;
            {
                StageScope sc(MRGS_STAGE_DEPTH_SORT, stream, 0);
                sort_status = depth_sort(sort_keys_a, sort_keys_b, sort_vals_a, sort_vals_b, a->P, hist, stream, &order, &launches);
                g_prof.launches += launches;
            }
            if (sort_status != MRGS_OK) return sort_status;
        }
        MRGS_LAUNCH_OK("depth_sort", stream, debug);
        {
            StageScope sc(MRGS_STAGE_SCAN, stream, 2);
            offsets_in_order(order, pp.tiles_touched, a->P, block_sums, offsets, stream);
        }
        MRGS_LAUNCH_OK("scan", stream, debug);

        rp.rec = pp.rec;
        rp.cf = pp.cf;
        rp.bbox = pp.bbox;

        // R has to reach the host: it sizes the binning buffer and is the call's first result (same point
        // as rasterizer_impl.cu:287, but through pinned memory on the caller's stream).
        int32_t* slot = pinned_slot();
        cudaEvent_t r_ready = r_ready_event();
        if (slot == nullptr || r_ready == nullptr) {
            set_error("mrgs_forward: cudaHostAlloc / cudaEventCreate failed");
            return MRGS_ERR_CUDA;
        }
        const uint32_t* R_dev = offsets + a->P - 1;
        if (!a->no_wait) {
            MRGS_CUDA_OK(cudaMemcpyAsync(slot, R_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            MRGS_CUDA_OK(cudaEventRecord(r_ready, stream));
        }

        // instance expansion, tile sort, tile ranges and the blend for a binning buffer laid out for
        // `cap` instances; with dev_count the kernels take the real count from the device
        auto bin_and_render = [&](char* bin, int64_t cap, const uint32_t* dev_count) -> int {
            MrgsBinningLayout bl;
            if (mrgs_binning_layout(cap, &bl) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
            uint16_t* keys_a = (uint16_t*)(bin + bl.keys);
            uint16_t* keys_b = (uint16_t*)(bin + bl.keys_unsorted);
            uint32_t* vals_a = (uint32_t*)(bin + bl.point_list);
            uint32_t* vals_b = (uint32_t*)(bin + bl.point_list_unsorted);
            {
                StageScope sc(MRGS_STAGE_DUPLICATE, stream, 1);
                launch_emit_instances(a->P, order, pp.tiles_touched, pp.rect, offsets, keys_a, vals_a, grid_x,
                                      (uint32_t)cap, stream);
            }
            MRGS_LAUNCH_OK("emit_instances", stream, debug);
            // two stable passes over the tile bits only; an even pass count leaves the result in A
            const int tile_bits = (int)higher_msb((uint32_t)tiles);
            uint16_t* keys_sorted = nullptr;
            uint32_t* vals_sorted = nullptr;
            {
                int launches = 0, sort_status;
                {
                    StageScope sc(MRGS_STAGE_SORT, stream, 0);
                    sort_status = tile_sort(keys_a, keys_b, vals_a, vals_b, (int)cap, tile_bits < 9 ? 9 : tile_bits,
                                            (uint32_t*)(bin + bl.sort_temp), stream, &keys_sorted, &vals_sorted, &launches,
                                            dev_count);
                    g_prof.launches += launches;
                }
                if (sort_status != MRGS_OK) return sort_status;
            }
            MRGS_LAUNCH_OK("tile_sort", stream, debug);
            if (vals_sorted != vals_a) {
                set_error("mrgs_forward: internal error, tile sort result not in buffer A");
                return MRGS_ERR_CUDA;
            }
            {
                StageScope sc(MRGS_STAGE_RANGES, stream, 1);
                launch_identify_tile_ranges((int)cap, dev_count, keys_sorted, ranges, stream);
            }
            MRGS_LAUNCH_OK("identify_tile_ranges", stream, debug);
            rp.point_list = vals_sorted;
            int st;
            {
                StageScope sc(MRGS_STAGE_RENDER_FWD, stream, 1);
                st = launch_render_fwd(rp, stream);
            }
            if (st != MRGS_OK) return st;
            MRGS_LAUNCH_OK("render_fwd", stream, debug);
            return MRGS_OK;
        };

        // Optimistic path: the caller lent a buffer sized for binning_capacity instances, so everything up
        // to the blend is enqueued BEFORE the host waits for R; the wait then overlaps queued GPU work
        // instead of draining the stream. If R turns out larger, the exact path below redoes the binning.
        bool rendered = false;
        const int64_t cap = a->binning_capacity;
        if (cap > 0 && cap < (1ll << 30) && a->binning_scratch != nullptr && radix_lookback_enabled() &&
            a->binning_scratch_bytes >= mrgs_binning_bytes(cap)) {
            const int st = bin_and_render((char*)a->binning_scratch, cap, R_dev);
            if (st != MRGS_OK) return st;
            rendered = true;
        }
        if (a->no_wait) {   // capture-safe: nothing below may wait on the stream
            if (!rendered) {
                set_error("mrgs_forward: no_wait needs binning_scratch/binning_capacity (and the look-back sort)");
                return MRGS_ERR_INVALID_ARGUMENT;
            }
            if (a->count_out != nullptr)
                MRGS_CUDA_OK(cudaMemcpyAsync(a->count_out, R_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            a->binning_buffer = a->binning_scratch;
            a->binning_capacity_used = cap;
            a->num_rendered = (int32_t)cap;
            return MRGS_OK;
        }
        MRGS_CUDA_OK(cudaEventSynchronize(r_ready));
        R = *slot;
        if (R < 0) {
            set_error("mrgs_forward: instance count overflow (R=%d)", R);
            return MRGS_ERR_UNSUPPORTED;
        }
        if (rendered && R <= cap) {
            a->binning_buffer = a->binning_scratch;
            a->binning_capacity_used = cap;
        } else if (R > 0) {
            if (a->binning_alloc == nullptr) {
                set_error("mrgs_forward: binning_alloc callback is null");
                return MRGS_ERR_INVALID_ARGUMENT;
            }
            const size_t need = mrgs_binning_bytes(R);
            char* bin = (char*)a->binning_alloc(a->binning_ctx, need);
            if (bin == nullptr) {
                set_error("mrgs_forward: binning_alloc(%zu) returned null", need);
                return MRGS_ERR_WORKSPACE;
            }
            if (rendered) MRGS_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), stream));
            const int st = bin_and_render(bin, R, nullptr);
            if (st != MRGS_OK) return st;
            a->binning_buffer = bin;
            a->binning_capacity_used = R;
            rendered = true;
        }
        a->num_rendered = R;
        if (rendered) return MRGS_OK;
    }
    a->num_rendered = R;

    // no surfels or no instances: the blend still writes background / zero maps
    int st;
    {
        StageScope sc(MRGS_STAGE_RENDER_FWD, stream, 1);
        st = launch_render_fwd(rp, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("render_fwd", stream, debug);
    return MRGS_OK;
}

int mrgs_shade_forward(const MrgsShadeArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_FWD, stream, 1);
        st = launch_shade(a, false, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("shade_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_shade_backward(const MrgsShadeArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_BWD, stream, 1);
        st = launch_shade(a, true, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("shade_bwd", stream, false);
    return MRGS_OK;
}

int mrgs_envlight_query(const MrgsShadeArgs* chain, int64_t n, const float* dirs, const float* roughness,
                        float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n > 0 && (dirs == nullptr || out == nullptr)) {
        set_error("mrgs_envlight_query: null dirs/out");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_FWD, stream, 1);
        st = launch_envlight_query(chain, n, dirs, roughness, out, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("envlight_query", stream, false);
    return MRGS_OK;
}

int mrgs_envlight_query_backward(const MrgsShadeArgs* chain, int64_t n, const float* dirs, const float* roughness,
                                 const float* dL_dout, float* dL_ddirs, float* dL_droughness, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n > 0 && (dirs == nullptr || dL_dout == nullptr)) {
        set_error("mrgs_envlight_query_backward: null dirs/dL_dout");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_BWD, stream, 1);
        st = launch_envlight_query_bwd(chain, n, dirs, roughness, dL_dout, dL_ddirs, dL_droughness, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("envlight_query_bwd", stream, false);
    return MRGS_OK;
}

int mrgs_surfel_shade_forward(const MrgsSurfelShadeArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_FWD, stream, 1);
        st = launch_surfel_shade(a, false, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("surfel_shade_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_surfel_shade_backward(const MrgsSurfelShadeArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_BWD, stream, 1);
        st = launch_surfel_shade(a, true, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("surfel_shade_bwd", stream, false);
    return MRGS_OK;
}

int mrgs_depth_normal_forward(int32_t width, int32_t height, float depth_ratio, const float* host_ray_matrix,
                              const float* host_origin, const float* allmap, float* surf_depth, float* surf_normal,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!host_ray_matrix || !host_origin) {
        set_error("mrgs_depth_normal_forward: null camera");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_FWD, stream, 1);
        st = launch_depth_normal(false, width, height, depth_ratio, host_ray_matrix, host_origin, allmap, surf_depth,
                                 surf_normal, nullptr, nullptr, nullptr, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("depth_normal_fwd", stream, false);
    return MRGS_OK;
}

int mrgs_depth_normal_backward(int32_t width, int32_t height, float depth_ratio, const float* host_ray_matrix,
                               const float* host_origin, const float* allmap, const float* dL_dsurf_depth,
                               const float* dL_dsurf_normal, float* dL_dallmap, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!host_ray_matrix || !host_origin) {
        set_error("mrgs_depth_normal_backward: null camera");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    int st;
    {
        StageScope sc(MRGS_STAGE_SHADE_BWD, stream, 1);
        st = launch_depth_normal(true, width, height, depth_ratio, host_ray_matrix, host_origin, allmap, nullptr, nullptr,
                                 dL_dsurf_depth, dL_dsurf_normal, dL_dallmap, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("depth_normal_bwd", stream, false);
    return MRGS_OK;
}

#define MRGS_CUBE_CHECK(cond, who)                                  \
    if (!(cond)) {                                                  \
        set_error("%s: bad arguments", who);                        \
        return MRGS_ERR_INVALID_ARGUMENT;                           \
    }

int mrgs_cubemap_mip_forward(const float* in, float* out, int32_t res_in, int32_t channels, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(in && out && res_in >= 2 && (res_in % 2) == 0 && channels > 0, "mrgs_cubemap_mip_forward");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_cubemap_mip_fwd(in, out, res_in / 2, channels, stream);
    }
    MRGS_LAUNCH_OK("cubemap_mip_fwd", stream, false);
    return MRGS_OK;
}
int mrgs_cubemap_mip_backward(const float* dout, float* din, int32_t res_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(dout && din && res_out >= 1, "mrgs_cubemap_mip_backward");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_cubemap_mip_bwd(dout, din, res_out, stream);
    }
    MRGS_LAUNCH_OK("cubemap_mip_bwd", stream, false);
    return MRGS_OK;
}
int mrgs_specular_bounds(int32_t res, float cutoff, int32_t* bounds, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(bounds && res >= 1, "mrgs_specular_bounds");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_specular_bounds(res, cutoff, bounds, stream);
    }
    MRGS_LAUNCH_OK("specular_bounds", stream, false);
    return MRGS_OK;
}
int mrgs_specular_cubemap_forward(const float* cubemap, const int32_t* bounds, int32_t res, float roughness,
                                  float cutoff, float* out4, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(cubemap && bounds && out4 && res >= 1, "mrgs_specular_cubemap_forward");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_specular_cubemap(false, res, roughness, cutoff, bounds, cubemap, out4, nullptr, nullptr, stream);
    }
    MRGS_LAUNCH_OK("specular_cubemap_fwd", stream, false);
    return MRGS_OK;
}
int mrgs_specular_cubemap_backward(const float* cubemap, const int32_t* bounds, int32_t res, float roughness,
                                   float cutoff, const float* dout4, float* dcubemap, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(bounds && dout4 && dcubemap && res >= 1, "mrgs_specular_cubemap_backward");
    MRGS_CUDA_OK(cudaMemsetAsync(dcubemap, 0, (size_t)6 * res * res * 3 * sizeof(float), stream));
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_specular_cubemap(true, res, roughness, cutoff, bounds, cubemap, nullptr, dout4, dcubemap, stream);
    }
    MRGS_LAUNCH_OK("specular_cubemap_bwd", stream, false);
    return MRGS_OK;
}
int mrgs_diffuse_cubemap_forward(const float* cubemap, int32_t res, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(cubemap && out && res >= 1, "mrgs_diffuse_cubemap_forward");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_diffuse_cubemap(false, res, cubemap, out, nullptr, nullptr, stream);
    }
    MRGS_LAUNCH_OK("diffuse_cubemap_fwd", stream, false);
    return MRGS_OK;
}
int mrgs_diffuse_cubemap_backward(const float* cubemap, int32_t res, const float* dout, float* dcubemap,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(dout && dcubemap && res >= 1, "mrgs_diffuse_cubemap_backward");
    MRGS_CUDA_OK(cudaMemsetAsync(dcubemap, 0, (size_t)6 * res * res * 3 * sizeof(float), stream));
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_diffuse_cubemap(true, res, cubemap, nullptr, dout, dcubemap, stream);
    }
    MRGS_LAUNCH_OK("diffuse_cubemap_bwd", stream, false);
    return MRGS_OK;
}

int32_t mrgs_prefilter_patch_count(int32_t res, int32_t rows_per_lane, int32_t patch_width) {
    return prefilter_patch_count(res, rows_per_lane, patch_width);
}
int mrgs_prefilter_texel_table(int32_t res, float* table, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(table && res >= 1 && res <= 32768, "mrgs_prefilter_texel_table");
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        launch_texel_table(res, table, stream);
    }
    MRGS_LAUNCH_OK("prefilter_texel_table", stream, false);
    return MRGS_OK;
}
static int prefilter_build_check(const MrgsPrefilterBuildArgs* a, bool fill, const char* who) {
    if (a == nullptr || a->res < 1 || a->res > 32768 || a->texel_table == nullptr || a->kind < 0 || a->kind > 3 ||
        mrgs_prefilter_patch_count(a->res, a->rows_per_lane, a->patch_width) < 0) {
        set_error("%s: bad arguments", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    const bool spec = a->kind == MRGS_PREFILTER_SPECULAR || a->kind == MRGS_PREFILTER_SPECULAR_T;
    // (the count pass of _SPECULAR may run without wsum: shape selection only needs the counts)
    if (spec && (a->bounds == nullptr || (a->wsum == nullptr && (fill || a->kind == MRGS_PREFILTER_SPECULAR_T)))) {
        set_error("%s: the specular kinds need bounds and wsum", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (!fill && (!a->seg_count || !a->slot_count || !a->tap_count)) {
        set_error("%s: null count outputs", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (fill && (!a->plan.patch_seg_begin || !a->plan.patch_slot_begin || !a->plan.seg_desc || !a->plan.spans ||
                 !a->plan.weights)) {
        set_error("%s: incomplete plan storage", who);
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    return MRGS_OK;
}
int mrgs_prefilter_plan_count(const MrgsPrefilterBuildArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = prefilter_build_check(a, false, "mrgs_prefilter_plan_count");
    if (st != MRGS_OK) return st;
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        st = launch_prefilter_build(a, false, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("prefilter_plan_count", stream, false);
    return MRGS_OK;
}
int mrgs_prefilter_plan_fill(const MrgsPrefilterBuildArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int st = prefilter_build_check(a, true, "mrgs_prefilter_plan_fill");
    if (st != MRGS_OK) return st;
    {
        StageScope sc(MRGS_STAGE_CUBEMAP, stream, 1);
        st = launch_prefilter_build(a, true, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("prefilter_plan_fill", stream, false);
    return MRGS_OK;
}
int mrgs_prefilter_apply(const MrgsPrefilterJob* jobs, int32_t num_jobs, int32_t backward, int32_t max_ctas, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(jobs && num_jobs >= 1 && num_jobs <= MRGS_PREFILTER_MAX_JOBS, "mrgs_prefilter_apply");
    for (int k = 0; k < num_jobs; ++k) {
        const MrgsPrefilterJob& j = jobs[k];
        const bool ok = j.src && j.dst && (j.src_stride == 3 || j.src_stride == 4) &&
                        (j.dst_stride == 3 || j.dst_stride == 4) && j.plan.patch_seg_begin && j.plan.patch_slot_begin &&
                        j.plan.seg_desc && j.plan.spans && j.plan.weights &&
                        mrgs_prefilter_patch_count(j.plan.res, j.plan.rows_per_lane, j.plan.patch_width) >= 0 &&
                        (j.src_stride != 4 || ((uintptr_t)j.src & 15) == 0);
        if (!ok) {
            set_error("mrgs_prefilter_apply: job %d is malformed", k);
            return MRGS_ERR_INVALID_ARGUMENT;
        }
    }
    int st;
    {
        StageScope sc(backward ? MRGS_STAGE_PREFILTER_BWD : MRGS_STAGE_PREFILTER_FWD, stream, 1);
        st = launch_prefilter_apply(jobs, num_jobs, max_ctas, stream);
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("prefilter_apply", stream, false);
    return MRGS_OK;
}
int mrgs_mip_pyramid_forward(const float* base, int32_t res, int32_t num_levels, float* const* levels4, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(base && levels4 && res >= 1 && num_levels >= 1 && num_levels <= MRGS_MAX_MIP_LEVELS &&
                        (res % (1 << (num_levels - 1))) == 0, "mrgs_mip_pyramid_forward");
    for (int l = 0; l < num_levels; ++l) MRGS_CUBE_CHECK(levels4[l] != nullptr, "mrgs_mip_pyramid_forward");
    int st, launches = 0;
    {
        StageScope sc(MRGS_STAGE_PREFILTER_FWD, stream, 0);
        st = launch_mip_pyramid(base, res, num_levels, levels4, stream, &launches);
        g_prof.launches += launches;
    }
    if (st != MRGS_OK) return st;
    MRGS_LAUNCH_OK("mip_pyramid", stream, false);
    return MRGS_OK;
}
int mrgs_mip_chain_backward(int32_t res, int32_t num_levels, float* const* grads3, const float* extra_last3,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MRGS_CUBE_CHECK(grads3 && res >= 1 && num_levels >= 1 && num_levels <= MRGS_MAX_MIP_LEVELS &&
                        (res % (1 << (num_levels - 1))) == 0, "mrgs_mip_chain_backward");
    for (int l = 0; l < num_levels; ++l) MRGS_CUBE_CHECK(grads3[l] != nullptr, "mrgs_mip_chain_backward");
    {
        StageScope sc(MRGS_STAGE_PREFILTER_BWD, stream, num_levels - 1);
        for (int l = num_levels - 2; l >= 0; --l)
            launch_mip_bwd_acc(grads3[l + 1], l == num_levels - 2 ? extra_last3 : nullptr, grads3[l], res >> (l + 1), 1,
                               stream);
    }
    MRGS_LAUNCH_OK("mip_chain_backward", stream, false);
    return MRGS_OK;
}

int mrgs_backward(const MrgsBackwardArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (a == nullptr) {
        set_error("mrgs_backward: null args");
        return MRGS_ERR_INVALID_ARGUMENT;
    }
    if (a->P <= 0) return MRGS_OK;
    if (a->S < 0 || a->S > MRGS_MAX_FEATURES) {
        set_error("mrgs_backward: S=%d feature channels, supported range is 0..%d", a->S,
                  MRGS_MAX_FEATURES);
        return MRGS_ERR_UNSUPPORTED;
    }
    const bool debug = a->debug != 0;
    const int W = a->width, H = a->height;
    const int grid_x = (W + kTileX - 1) / kTileX, grid_y = (H + kTileY - 1) / kTileY;

    MrgsGeomLayout gl;
    MrgsImageLayout il;
    MrgsBinningLayout bl;
    if (mrgs_geom_layout(a->P, a->S, &gl) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
    if (mrgs_image_layout(W, H, &il) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
    if (mrgs_binning_layout(a->num_rendered, &bl) != MRGS_OK) return MRGS_ERR_INVALID_ARGUMENT;
    const size_t arena_bytes = mrgs_grad_arena_bytes(a->P, a->S);
    if (a->grad_arena == nullptr || a->grad_arena_bytes < arena_bytes) {
        set_error("mrgs_backward: gradient arena too small (%zu < %zu)", a->grad_arena_bytes,
                  arena_bytes);
        return MRGS_ERR_WORKSPACE;
    }
    const char* geom = (const char*)a->geom_buffer;
    const char* img = (const char*)a->image_buffer;
    const char* bin = (const char*)a->binning_buffer;
    if (geom == nullptr || img == nullptr || (a->num_rendered > 0 && bin == nullptr)) {
        set_error("mrgs_backward: missing forward buffers");
        return MRGS_ERR_INVALID_ARGUMENT;
    }

    MRGS_CUDA_OK(cudaMemsetAsync(a->grad_arena, 0, arena_bytes, stream));

    if (a->num_rendered > 0) {
        RenderBwdParams rp{};
        rp.S = a->S; rp.W = W; rp.H = H; rp.grid_x = grid_x; rp.grid_y = grid_y;
        rp.cf_stride = gl.cf_stride; rp.grad_stride = grad_stride(a->S);
        rp.ranges = (const uint2*)(img + il.ranges);
        rp.point_list = (const uint32_t*)(bin + bl.point_list);
        rp.rec = (const float*)(geom + gl.rec);
        rp.cf = (const float*)(geom + gl.cf);
        rp.bbox = (const float4*)(geom + gl.bbox);
        rp.background = a->background;
        rp.state = (const float*)(img + il.state);
        rp.dL_dcolor = a->dL_dout_color;
        rp.dL_dfeature = a->dL_dout_feature;
        rp.dL_dothers = a->dL_dout_others;
        rp.grad_arena = (float*)a->grad_arena;
        int st;
        {
            StageScope sc(MRGS_STAGE_RENDER_BWD, stream, 1);
            st = launch_render_bwd(rp, stream);
        }
        if (st != MRGS_OK) return st;
        MRGS_LAUNCH_OK("render_bwd", stream, debug);
    }

    // the reference rebuilds W,H in the backward from focal*tan*2 in fp32 (backward.cu:646-647,
    // rasterizer_impl.cu:398-399); the truncation can give size-1 and is part of the contract
    const float focal_y = H / (2.0f * a->tan_fovy);
    const float focal_x = W / (2.0f * a->tan_fovx);
    volatile float fw = focal_x * a->tan_fovx;
    volatile float fh = focal_y * a->tan_fovy;
    const int Wb = (int)(fw * 2.0f);
    const int Hb = (int)(fh * 2.0f);

    PreprocessBwdParams pb{};
    pb.P = a->P; pb.S = a->S; pb.D = a->sh_degree; pb.M = a->sh_coeffs; pb.W = Wb; pb.H = Hb;
    pb.cf_stride = gl.cf_stride; pb.grad_stride = grad_stride(a->S);
    pb.accumulate = a->accumulate != 0;
    if (pb.accumulate && a->shs != nullptr && a->sh_coeffs > 16) {
        set_error("mrgs_backward: accumulate needs sh_coeffs <= 16, got %d", a->sh_coeffs);
        return MRGS_ERR_UNSUPPORTED;
    }
    pb.means3D = a->means3D; pb.scales = a->scales; pb.rotations = a->rotations; pb.shs = a->shs;
    pb.transMat_precomp = a->transMat_precomp;
    pb.viewmatrix = a->viewmatrix; pb.projmatrix = a->projmatrix; pb.campos = a->campos;
    pb.radii = a->radii;
    pb.rec = (const float*)(geom + gl.rec);
    pb.clamped = (const uint8_t*)(geom + gl.clamped);
    pb.grad_arena = (const float*)a->grad_arena;
    pb.dL_dmeans2D = a->dL_dmeans2D; pb.dL_dcolors = a->dL_dcolors; pb.dL_dfeatures = a->dL_dfeatures;
    pb.dL_dopacity = a->dL_dopacity; pb.dL_dmeans3D = a->dL_dmeans3D; pb.dL_dtransMat = a->dL_dtransMat;
    pb.dL_dsh = a->dL_dsh; pb.dL_dscales = a->dL_dscales; pb.dL_drotations = a->dL_drotations;
    {
        StageScope sc(MRGS_STAGE_PREPROCESS_BWD, stream, 1);
        launch_preprocess_bwd(pb, stream);
    }
    MRGS_LAUNCH_OK("preprocess_bwd", stream, debug);
    return MRGS_OK;
}

}  // extern "C"
