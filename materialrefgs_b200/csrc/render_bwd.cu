// render_bwd.cu — tile-blended backward: re-traverses each tile's instance list back to front
// and accumulates the raw per-surfel gradients (dT[9], dmean2D[2], dopacity, dnormal[3],
// dcolour[3], dfeature[S]) into the gradient arena.
//
// Behavioural reference: renderCUDA rast/cuda_rasterizer/backward.cu:145-468 (same skip tests,
// same recurrences). Own design:
//   * warp-autonomous traversal without CTA barriers (see render_fwd.cu): each warp (8x4 pixel block)
//     starts at its own largest `last_contributor` — instances nobody blended are neither staged nor
//     evaluated — walks back to front 32 instances per step, culls them against the surfel's
//     conservative alpha-support box (1/32 of an instruction stream per instance) and stages only
//     the survivors' records in its private shared-memory slots;
//   * gradients of one instance are summed across the 32 pixels of a warp by transposing them
//     through a padded shared-memory tile (one STS per value, then lane c sums row c with 8
//     conflict-free LDS.128) and leave the warp as ONE coalesced red.global.add per instance
//     instead of 16+S same-address atomics per pixel;
//   * a warp skips the reduction when none of its pixels is touched by the instance.
#include <cstdlib>

#include <atomic>

#include "kernels.cuh"
#include "splat_math.cuh"

namespace mrgs {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

constexpr int kRedStride = 36;  // floats per value row: 32 lanes + 4 pad keeps LDS.128 conflict-free

template <int NQ, int WPC>
constexpr int bwd_smem_bytes() {
    // per warp: 4 geometry vectors + NQ colour vectors + ids for 32 staged instances, and the
    // [NV][36] transposition buffer of the warp reduction
    return WPC * ((4 + NQ) * 32 * 16 + 32 * 4 + (kGradColor + NQ * 4) * kRedStride * 4);
}

// WPC warps per CTA: the warps of a tile never synchronise with each other, so a tile can be split over 8 / WPC CTAs
// (blockIdx.z) — the register file then holds a non-integer number of TILES per SM (6 CTAs of 4 warps at 80
// registers = 24 warps instead of 2 CTAs of 8 = 16), and a finished warp frees its slot without waiting for its tile.
template <int NQ, int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB)
render_bwd_kernel(const RenderBwdParams p) {
    constexpr int NC = NQ * 4;           // colour + feature (+ padding) channels
    constexpr int NV = kGradColor + NC;  // values per instance incl. padding channels
    constexpr bool kTwoPass = NV > 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 (*s_g)[4][32] = reinterpret_cast<float4 (*)[4][32]>(smem_raw);
    float4 (*s_cf)[NQ][32] = reinterpret_cast<float4 (*)[NQ][32]>(smem_raw + WPC * 4 * 32 * 16);
    uint32_t (*s_id)[32] = reinterpret_cast<uint32_t (*)[32]>(smem_raw + WPC * (4 + NQ) * 32 * 16);
    float* s_red = reinterpret_cast<float*>(smem_raw + WPC * ((4 + NQ) * 32 * 16 + 32 * 4)) +
                   (threadIdx.x >> 5) * (NV * kRedStride);

    const int tid = blockIdx.z * (WPC * 32) + threadIdx.x;   // slot of this thread's pixel in the tile
    const int lane = tid & 31, warp = tid >> 5;               // warp: which 8x4 block of the tile
    const int lw = threadIdx.x >> 5;                          // this warp's shared-memory slot set
    const int tile = blockIdx.y * p.grid_x + blockIdx.x;
    const int px = blockIdx.x * kTileX + slot_x(tid);
    const int py = blockIdx.y * kTileY + slot_y(tid);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;
    const float bx0 = (float)(blockIdx.x * kTileX + (warp & 1) * 8), bx1 = bx0 + 7.0f;
    const float by0 = (float)(blockIdx.y * kTileY + (warp >> 1) * 4), by1 = by0 + 3.0f;
    const float bu0 = bx0 + by0, bu1 = bx1 + by1, bv0 = bx0 - by1, bv1 = bx1 - by0;  // diagonal extents

    const uint2 range = p.ranges[tile];
    const uint32_t* __restrict__ list = p.point_list + range.x;

    const float* st = p.state + (size_t)tile * (5 * kTilePixels) + tid;
    const float T_final = inside ? st[0] : 0.0f;
    const float final_D = inside ? st[1 * kTilePixels] : 0.0f;
    const float final_D2 = inside ? st[2 * kTilePixels] : 0.0f;
    const int last_contributor = inside ? (int)reinterpret_cast<const uint32_t*>(st)[3 * kTilePixels] : 0;
    const int median_contributor = inside ? (int)reinterpret_cast<const uint32_t*>(st)[4 * kTilePixels] : 0;
    const float final_A = 1.0f - T_final;

    // entries [0, warp_last) of the tile's list were blended by at least one pixel of this warp
    const int warp_last = __reduce_max_sync(kFull, last_contributor);
    if (warp_last == 0) return;

    float2 dL_dpix[NC / 2];  // channel pairs (2c, 2c+1)
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) dL_dpix[c] = make_float2(0.0f, 0.0f);
    float dL_ddepth = 0.f, dL_daccum = 0.f, dL_dreg = 0.f, dL_dmedian = 0.f;
    float dL_dn0 = 0.f, dL_dn1 = 0.f, dL_dn2 = 0.f;
    float bg_dot_dpixel = 0.f;
    if (inside) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float g = 0.0f;
            if (c < 3) {
                g = p.dL_dcolor[c * HW + pix];
                bg_dot_dpixel += p.background[c] * g;
            } else if (c - 3 < p.S) {
                g = p.dL_dfeature[(c - 3) * HW + pix];
            }
            if (c & 1)
                dL_dpix[c >> 1].y = g;
            else
                dL_dpix[c >> 1].x = g;
        }
        dL_ddepth = p.dL_dothers[kDepthOff * HW + pix];
        dL_daccum = p.dL_dothers[kAlphaOff * HW + pix];
        dL_dn0 = p.dL_dothers[(kNormalOff + 0) * HW + pix];
        dL_dn1 = p.dL_dothers[(kNormalOff + 1) * HW + pix];
        dL_dn2 = p.dL_dothers[(kNormalOff + 2) * HW + pix];
        dL_dmedian = p.dL_dothers[kMidDepthOff * HW + pix];
        dL_dreg = p.dL_dothers[kDistortionOff * HW + pix];
    }

    // Blend recurrences in closed form. The reference walks back to front keeping, per channel, the colour
    // accumulated BEHIND the current surfel (accum_rec) and adds (c - accum_rec) * dL_dpixel * T to dL_dalpha.
    // With q_i = sum over ALL blended channels of value_i * dL_dchannel (colour, features, depth, normal, alpha)
    // and w_k = alpha_k T_k this is   T_i q_i - B_i / (1 - alpha_i),   B_i = sum_{k behind i} w_k q_k :
    // one dot product per pair and ONE scalar running sum instead of 2 vectors of per-channel state
    // (gradient-only arithmetic: every skip decision stays exact, the 1e-3 bar is checked by the parity tests).
    float T = T_final;
    float behind = 0.f;       // B
    float last_dL_dT = 0.f;

    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(p.rec);
    const float4* __restrict__ cf4 = reinterpret_cast<const float4*>(p.cf);

    // back-to-front in steps of 32 list entries; lane l of a step holds entry hi-1-l, so walking the
    // survivor mask from bit 0 upwards visits entries in descending order
    // software pipeline: the NEXT step's id is loaded and its 32-byte bound (box + diagonal slabs, one sector) is
    // pulled into L1 with a prefetch while the current survivors are processed — a prefetch holds no registers,
    // where keeping the two float4 in flight cost 8 per thread across the whole gradient math
    uint32_t id_next = 0;
    if (warp_last - 1 - lane >= 0) {
        id_next = list[warp_last - 1 - lane];
        prefetch_l1(p.bbox + 2 * (size_t)id_next);
    }

    for (int hi = warp_last; hi > 0; hi -= 32) {
        const uint32_t id = id_next;
        const int e_mine = hi - 1 - lane;
        const int e_next = e_mine - 32;
        float4 bb = make_float4(0.f, 0.f, -1.f, -1.f), bd = bb;
        if (e_mine >= 0) {
            bb = p.bbox[2 * (size_t)id];
            bd = p.bbox[2 * (size_t)id + 1];
        }
        if (e_next >= 0) {
            id_next = list[e_next];
            prefetch_l1(p.bbox + 2 * (size_t)id_next);
        }
        const bool keep = (e_mine >= 0) && !(bb.x > bx1 || bb.z < bx0 || bb.y > by1 || bb.w < by0) &&
                          !(bd.x > bu1 || bd.z < bu0 || bd.y > bv1 || bd.w < bv0);
        unsigned mask = __ballot_sync(kFull, keep);
        if (mask == 0) continue;
        if (keep) {
            s_id[lw][lane] = id;
            const float4* r = rec4 + (size_t)id * (kGeomFloats / 4);
            s_g[lw][0][lane] = r[0];
            s_g[lw][1][lane] = r[1];
            s_g[lw][2][lane] = r[2];
            s_g[lw][3][lane] = r[3];
            const float4* c = cf4 + (size_t)id * NQ;
#pragma unroll
            for (int q = 0; q < NQ; ++q) s_cf[lw][q][lane] = c[q];
        }
        __syncwarp();

        while (mask != 0) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const int e = hi - 1 - j;
            const float4 g0 = s_g[lw][0][j], g1 = s_g[lw][1][j], g2 = s_g[lw][2][j];
            const float4 g3 = s_g[lw][3][j];
            SplatHit h;
            const bool valid = (e < last_contributor) && ray_splat(g0, g1, g2, g3.w, pxf, pyf, h);
            if (!__any_sync(kFull, valid)) continue;

            // v: the 12 geometric values that exist only for pixels the instance touches. The colour/feature and
            // normal gradients are all w * (a per-pixel constant): only w leaves the branch (0 for untouched pixels)
            // and the products are formed at the store, so they never occupy registers across the gradient math.
            float v[kGradNormal];
#pragma unroll
            for (int c = 0; c < kGradNormal; ++c) v[c] = 0.0f;
            float w_out = 0.0f;

            if (valid) {
                const float alpha = h.alpha, G = h.G;
                // gradients tolerate approximate reciprocals (1e-3 bar); every skip decision above is exact
                const float inv_1ma = __fdividef(1.0f, 1.0f - alpha);
                T = T * inv_1ma;
                const float w = alpha * T;
                float q;
                {
                    // two channels per instruction (Blackwell packed fp32: fma.rn.f32x2)
                    float2 dot2 = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        const float4 cv = s_cf[lw][k][j];
                        dot2 = __ffma2_rn(make_float2(cv.x, cv.y), dL_dpix[2 * k], dot2);
                        dot2 = __ffma2_rn(make_float2(cv.z, cv.w), dL_dpix[2 * k + 1], dot2);
                    }
                    q = dot2.x + dot2.y;
                }

                const float c_d = h.depth;
                const float m_d = distortion_coord(c_d);
                const float dmd_dd = ((kFar * kNear) / (kFar - kNear)) * __fdividef(1.0f, c_d * c_d);
                float dL_dz = 0.0f;
                if (e == median_contributor - 1) dL_dz += dL_dmedian;
                const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                const float dist_term = dL_dweight - last_dL_dT;
                last_dL_dT = dL_dweight * alpha + (1.0f - alpha) * last_dL_dT;
                const float dL_dmd = 2.0f * w * (m_d * final_A - final_D) * dL_dreg;
                dL_dz += dL_dmd * dmd_dd;

                // depth, alpha ("colour" 1) and normal are blended channels like the colours
                q += c_d * dL_ddepth + dL_daccum + g3.x * dL_dn0 + g3.y * dL_dn1 + g3.z * dL_dn2;
                w_out = w;

                float dL_dalpha = T * (q + dist_term) - inv_1ma * behind;
                behind += w * q;
                if (bg_dot_dpixel != 0.0f) dL_dalpha += (-T_final * inv_1ma) * bg_dot_dpixel;

                const float dL_dG = g2.w * dL_dalpha;
                dL_dz += w * dL_ddepth;

                if (h.rho3d <= h.rho2d) {
                    const float Twx = g0.w, Twy = g1.w;
                    const float dL_dsx = dL_dG * -G * h.sx + dL_dz * Twx;
                    const float dL_dsy = dL_dG * -G * h.sy + dL_dz * Twy;
                    const float inv_pz = __fdividef(1.0f, h.pz);
                    const float dsx_pz = dL_dsx * inv_pz;
                    const float dsy_pz = dL_dsy * inv_pz;
                    const float dpx = dsx_pz, dpy = dsy_pz, dpz = -(dsx_pz * h.sx + dsy_pz * h.sy);
                    // dL_dk = cross(l, dL_dp), dL_dl = cross(dL_dp, k)
                    const float dkx = h.ly * dpz - h.lz * dpy;
                    const float dky = h.lz * dpx - h.lx * dpz;
                    const float dkz = h.lx * dpy - h.ly * dpx;
                    const float dlx = dpy * h.kz - dpz * h.ky;
                    const float dly = dpz * h.kx - dpx * h.kz;
                    const float dlz = dpx * h.ky - dpy * h.kx;
                    v[kGradT + 0] = -dkx;
                    v[kGradT + 1] = -dky;
                    v[kGradT + 2] = -dkz;
                    v[kGradT + 3] = -dlx;
                    v[kGradT + 4] = -dly;
                    v[kGradT + 5] = -dlz;
                    v[kGradT + 6] = pxf * dkx + pyf * dlx + dL_dz * h.sx;
                    v[kGradT + 7] = pxf * dky + pyf * dly + dL_dz * h.sy;
                    v[kGradT + 8] = pxf * dkz + pyf * dlz + dL_dz;
                } else {
                    const float dG_ddelx = -G * 2.0f * h.dx;
                    const float dG_ddely = -G * 2.0f * h.dy;
                    v[kGradMean2D + 0] = dL_dG * dG_ddelx;
                    v[kGradMean2D + 1] = dL_dG * dG_ddely;
                    v[kGradT + 8] = dL_dz;
                }
                v[kGradOpacity] = G * dL_dalpha;
            }

            // transpose through shared memory: lane c sums value c over the 32 pixels (8 LDS.128 + 31 FADD)
            // and the warp leaves ONE coalesced red.global.add per arena row
#pragma unroll
            for (int c = 0; c < kGradNormal; ++c) s_red[c * kRedStride + lane] = v[c];
            s_red[(kGradNormal + 0) * kRedStride + lane] = w_out * dL_dn0;
            s_red[(kGradNormal + 1) * kRedStride + lane] = w_out * dL_dn1;
            s_red[(kGradNormal + 2) * kRedStride + lane] = w_out * dL_dn2;
            {
                const float2 w2 = make_float2(w_out, w_out);
#pragma unroll
                for (int k = 0; k < NC / 2; ++k) {
                    const float2 g = __fmul2_rn(w2, dL_dpix[k]);   // mul.rn.f32x2
                    s_red[(kGradColor + 2 * k + 0) * kRedStride + lane] = g.x;
                    s_red[(kGradColor + 2 * k + 1) * kRedStride + lane] = g.y;
                }
            }
            __syncwarp();
            float* row = p.grad_arena + (size_t)s_id[lw][j] * p.grad_stride;
            const int n_live = kGradFeature + p.S;  // padding channels carry no gradient
#pragma unroll
            for (int pass = 0; pass < (kTwoPass ? 2 : 1); ++pass) {
                const int c = pass * 32 + lane;
                if (c < n_live) {
                    const float4* r4 = reinterpret_cast<const float4*>(s_red + c * kRedStride);
                    float part[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 t = r4[q];
                        part[q] = (t.x + t.y) + (t.z + t.w);
                    }
                    const float total = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
                    atomicAdd(row + c, total);
                }
            }
            __syncwarp();
        }
        __syncwarp();  // all lanes are past their reads before the slots are overwritten
    }
}

}  // namespace

namespace {

template <int NQ, int WPC, int MINB>
void launch_variant(const RenderBwdParams& p, cudaStream_t stream) {
    // the attribute is per device: one flag per device (a process may drive several GPUs); set only when the
    // variant needs more than the 48 KB every device grants by default
    constexpr int kMaxDevices = 64;
    static std::atomic<bool> configured[kMaxDevices];
    int dev = 0;
    if (bwd_smem_bytes<NQ, WPC>() > 48 * 1024 && cudaGetDevice(&dev) == cudaSuccess) {
        if (dev < 0 || dev >= kMaxDevices || !configured[dev].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(render_bwd_kernel<NQ, WPC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bwd_smem_bytes<NQ, WPC>());
            if (dev >= 0 && dev < kMaxDevices) configured[dev].store(true, std::memory_order_release);
        }
    }
    const dim3 grid(p.grid_x, p.grid_y, kWarpsPerTile / WPC);
    render_bwd_kernel<NQ, WPC, MINB><<<grid, WPC * 32, bwd_smem_bytes<NQ, WPC>(), stream>>>(p);
}

// MRGS_BWD_SPLIT=82 selects the unsplit launch (8 warps per CTA, 2 CTAs per SM) of the S <= 9 kernel for A/B
// measurements. Measured at C3 (profiles/r01_v7_summary.md), first with 127 registers of natural demand: 8x2 1.029 ms,
// 4x4 0.997, 4x5 (96 regs) 0.952, 1x20 0.952, 2x10 0.995, 4x6 (80 regs, 108 B spill) 1.058; then with the register diet
// (products formed at the store, L1 prefetch of the next bounds: 106 registers of natural demand): 4x5 0.888,
// 4x6 (80 regs, 4 B spill) 0.872, 4x7 (72 regs, 28 B spill) 0.945.
bool bwd_unsplit() {
    static const bool v = [] {
        const char* e = getenv("MRGS_BWD_SPLIT");
        return e != nullptr && atoi(e) == 82;
    }();
    return v;
}

}  // namespace

int launch_render_bwd(const RenderBwdParams& p, cudaStream_t stream) {
    switch (p.cf_stride / 4) {
        case 1: launch_variant<1, 4, 6>(p, stream); break;
        case 2: launch_variant<2, 4, 6>(p, stream); break;
        case 3:
            if (bwd_unsplit())
                launch_variant<3, 8, 2>(p, stream);
            else
                launch_variant<3, 4, 6>(p, stream);
            break;
        case 4: launch_variant<4, 4, 2>(p, stream); break;
        case 5: launch_variant<5, 4, 2>(p, stream); break;
        case 6: launch_variant<6, 4, 2>(p, stream); break;
        case 7: launch_variant<7, 4, 2>(p, stream); break;
        default:
            set_error("render_bwd: unsupported feature count S=%d (max %d)", p.S, MRGS_MAX_FEATURES);
            return MRGS_ERR_UNSUPPORTED;
    }
    return MRGS_OK;
}

}  // namespace mrgs
