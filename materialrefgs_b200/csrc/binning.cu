// binning.cu — surfel x tile instance generation and (tile | depth) ordering, all hand-written.
//
// Behavioural reference: duplicateWithKeys rasterizer_impl.cu:72-113, the 64-bit
// cub::DeviceRadixSort::SortPairs :309-314 (+ getHigherMsb :37-52), identifyTileRanges :118-140 and
// the InclusiveSum :283. Required result (bit-exact): instances ordered by (tile id, depth bits)
// with ties kept in emission order, i.e. ascending surfel id.
//
// Own design. The reference sorts R instances on 44-45 key bits (6 onesweep passes over 12-byte
// pairs). The same order is produced here with far less traffic by splitting the key:
//   1. sort the P surfels ONCE by their 32 depth bits (culled surfels get key 0xffffffff); LSD radix,
//      stable, input in ascending id order  =>  order by (depth, id);
//   2. prefix-sum tiles_touched in that order, emit every surfel's (tile, id) instances in that order;
//   3. stable LSD radix sort of the R instances on the 12-13 tile-id bits only (2 passes, 16-bit keys)
//      =>  order by (tile, depth, id), identical to the reference's 64-bit sort;
//   4. tile ranges from the sorted 16-bit keys.
// One radix pass = per-block digit histogram (+ global digit totals) -> one CTA per digit row turns the
// counts into global offsets -> stable scatter in which each warp owns a contiguous run of the block's
// chunk, ranks its items with ballot-built peer masks + running per-(warp, digit) counters (no atomics
// in the ranking), stages them in digit order in shared memory and writes coalesced runs.
#include <stdlib.h>

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

namespace {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerThread = 16;
constexpr int kChunk = kSortThreads * kItemsPerThread;  // 4096 items per block
constexpr int kWarpRun = 32 * kItemsPerThread;          // 512 consecutive items per warp
constexpr int kMaxBins = 256;

// Lanes of the warp holding the same digit (match.any semantics, but built from one ballot per digit
// bit: the MATCH instruction itself is an order of magnitude slower than 8 VOTE + 8 LOP3).
__device__ __forceinline__ unsigned warp_peers(uint32_t d, int bits, bool valid) {
    unsigned peers = __ballot_sync(kFullMask, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < bits) {
            const bool bit = (d >> b) & 1u;
            const unsigned v = __ballot_sync(kFullMask, bit);
            peers &= bit ? v : ~v;
        }
    }
    return peers;
}

// per-block digit counts [digit][block] and global digit totals
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const KeyT* __restrict__ keys, int n, int shift, int bits, uint32_t* __restrict__ block_hist,
                  uint32_t* __restrict__ digit_total, int num_blocks) {
    __shared__ uint32_t s_hist[kMaxBins];
    const int bins = 1 << bits;
    for (int i = threadIdx.x; i < bins; i += kSortThreads) s_hist[i] = 0;
    __syncthreads();
    const int base = blockIdx.x * kChunk;
    const uint32_t mask = (uint32_t)bins - 1;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
#pragma unroll 4
    for (int k = 0; k < kItemsPerThread; ++k) {
        const int i = base + k * kSortThreads + threadIdx.x;
        const bool valid = i < n;
        const uint32_t d = valid ? (((uint32_t)keys[i] >> shift) & mask) : 0u;
        // neighbouring instances share digits: one shared atomic per distinct digit of the warp
        const unsigned peers = warp_peers(d, bits, valid);
        if (valid && (peers & lt) == 0) atomicAdd(&s_hist[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += kSortThreads) {
        const uint32_t c = s_hist[i];
        block_hist[(size_t)i * num_blocks + blockIdx.x] = c;
        if (c) atomicAdd(&digit_total[i], c);
    }
}

// one CTA per digit row: block_hist[d][*] -> exclusive prefix over blocks + start of digit d
__global__ void __launch_bounds__(256)
radix_row_scan_kernel(uint32_t* __restrict__ block_hist, const uint32_t* __restrict__ digit_total, int bins,
                      int num_blocks) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_start;
    const int d = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // start of this digit = sum of the totals of all smaller digits (bins <= 256: one value per thread)
    {
        uint32_t v = (threadIdx.x < d && threadIdx.x < bins) ? digit_total[threadIdx.x] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        if (lane == 0) s_warp[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += s_warp[w];
            s_start = t;
        }
        __syncthreads();
    }
    uint32_t* row = block_hist + (size_t)d * num_blocks;
    uint32_t carry = s_start;
    for (int base = 0; base < num_blocks; base += 256 * 4) {
        const int i0 = base + threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < num_blocks) ? row[i0 + k] : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();  // s_warp reuse
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t c = s_warp[w];
            if (w < warp) before += c;
            total += c;
        }
        uint32_t run = carry + before + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < num_blocks) row[i0 + k] = run;
            run += v[k];
        }
        carry += total;
    }
}

// Digit histograms of ALL passes from one read of the keys: hist[pass][digit] (global totals).
struct PassPlan {
    int passes;
    int shift[4];
    int bits[4];
};

// Every thread owns 16 CONSECUTIVE keys (one or two 16-byte loads per 8 keys) and run-length compresses
// each pass's digit sequence in registers: consecutive instances belong to the same surfel (row-major
// tiles), so the high tile bits repeat and one shared atomic covers a whole run; lanes are 16 items apart,
// i.e. mostly different surfels, so simultaneous atomics rarely collide. No ballots.
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
radix_multi_hist_kernel(const KeyT* __restrict__ keys, int n, const uint32_t* __restrict__ n_dev, PassPlan plan,
                        uint32_t* __restrict__ totals) {
    if (n_dev != nullptr) n = min(n, (int)*n_dev);   // launched for a capacity, the count lives on the device
    if ((long long)blockIdx.x * kChunk >= n) return;
    __shared__ uint32_t s_hist[4][kMaxBins];
    for (int i = threadIdx.x; i < 4 * kMaxBins; i += kSortThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int first = blockIdx.x * kChunk + threadIdx.x * kItemsPerThread;
    uint32_t key[kItemsPerThread];
    constexpr int kPerVec = 16 / (int)sizeof(KeyT);
    if (first + kItemsPerThread <= n) {
#pragma unroll
        for (int v = 0; v < kItemsPerThread / kPerVec; ++v) {
            const uint4 raw = *reinterpret_cast<const uint4*>(keys + first + v * kPerVec);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < kPerVec; ++e) {
                if (sizeof(KeyT) == 4)
                    key[v * kPerVec + e] = w[e];
                else
                    key[v * kPerVec + e] = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) key[k] = (first + k < n) ? (uint32_t)keys[first + k] : 0u;
    }
    const int count = max(0, min(kItemsPerThread, n - first));
    for (int ps = 0; ps < plan.passes; ++ps) {
        const int shift = plan.shift[ps];
        const uint32_t mask = (1u << plan.bits[ps]) - 1u;
        uint32_t run_d = 0, run_n = 0;
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) {
            if (k < count) {
                const uint32_t d = (key[k] >> shift) & mask;
                if (run_n != 0 && d != run_d) {
                    atomicAdd(&s_hist[ps][run_d], run_n);
                    run_n = 0;
                }
                run_d = d;
                ++run_n;
            }
        }
        if (run_n != 0) atomicAdd(&s_hist[ps][run_d], run_n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * kMaxBins; i += kSortThreads) {
        const uint32_t c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&totals[i], c);
    }
}

// totals[pass][256] -> starts[pass][256] (exclusive scan over digits), one CTA of 256 threads per pass
__global__ void __launch_bounds__(256) radix_digit_scan_kernel(const uint32_t* __restrict__ totals, uint32_t* __restrict__ starts) {
    __shared__ uint32_t s_warp[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = totals[blockIdx.x * kMaxBins + threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
        if (w < warp) before += s_warp[w];
    starts[blockIdx.x * kMaxBins + threadIdx.x] = before + incl - v;
}

constexpr uint32_t kFlagPartial = 1u << 30, kFlagInclusive = 2u << 30, kValueMask = (1u << 30) - 1u;

// Stable scatter of one radix pass. Item order inside a block: warp w owns items
// [w*512, (w+1)*512) of the chunk, walked 32 at a time, so (warp, step, lane) is ascending item index.
// Items are first placed at their block-local sorted position in shared memory, then written out so
// that every digit's run leaves as contiguous, coalesced stores.
// LOOKBACK = true makes the pass a single kernel (onesweep style): blocks take their chunk index from a
// ticket, publish their digit counts in `status[block][digit]` (2 flag bits + 30 value bits in ONE word) and
// obtain the exclusive prefix over earlier blocks by decoupled look-back; `block_off` then holds the
// per-digit global starts. LOOKBACK = false expects block_off[digit][block] from a separate scan.
template <typename KeyT, bool IOTA, bool LOOKBACK>
__global__ void __launch_bounds__(kSortThreads, 4)
radix_scatter_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift, int bits,
                     const uint32_t* __restrict__ block_off, int num_blocks, uint32_t* __restrict__ status,
                     uint32_t* __restrict__ ticket, const uint32_t* __restrict__ n_dev) {
    const int bins = 1 << bits;
    __shared__ int s_block;
    if (LOOKBACK) {
        if (threadIdx.x == 0) s_block = (int)atomicAdd(ticket, 1u);
        __syncthreads();
    }
    const int block = LOOKBACK ? s_block : (int)blockIdx.x;
    if (n_dev != nullptr) n = min(n, (int)*n_dev);   // grid sized for a capacity: surplus blocks leave at once
    if ((long long)block * kChunk >= n) return;
    __shared__ uint32_t s_cnt[kSortWarps][kMaxBins];  // per-warp digit counts -> local rank bases
    __shared__ uint32_t s_lstart[kMaxBins];           // block-local start of every digit
    __shared__ uint32_t s_gbase[kMaxBins];            // global start of this block's run of every digit
    __shared__ uint32_t s_wsum[kSortWarps];
    __shared__ KeyT s_keys[kChunk];
    __shared__ uint32_t s_vals[kChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kMaxBins; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    const uint32_t mask = (uint32_t)bins - 1;
    const int bbase = block * kChunk;
    const int wbase = bbase + warp * kWarpRun;
    KeyT key[kItemsPerThread];
    const unsigned lt = (1u << lane) - 1u;

    // phase 1: per-warp digit counts AND every item's rank inside its warp's run. The leader of each
    // peer group bumps the warp's digit counter with one shared atomic; the value it gets back is the
    // number of same-digit items in the warp's earlier steps (a warp's atomics to one address retire in
    // program order), so rank = old + (peers before me in this step). Steps are independent apart from
    // that atomic chain: all ballots overlap.
    uint16_t rank[kItemsPerThread];
#pragma unroll
    for (int k = 0; k < kItemsPerThread; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? keys_in[i] : (KeyT)0;
        const uint32_t d = ((uint32_t)key[k] >> shift) & mask;
        const unsigned peers = warp_peers(d, bits, valid);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) old = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(peers));
        old = __shfl_sync(kFullMask, old, leader < 0 ? 0 : leader);
        rank[k] = (uint16_t)(old + __popc(peers & lt));
    }
    __syncthreads();
    // phase 2: block-local exclusive scan over digits (bins <= 256 = one digit per thread), then the
    // per-warp bases inside each digit's local run
    {
        uint32_t tot = 0;
        const int d = threadIdx.x;
        if (d < bins) {
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) tot += s_cnt[w][d];
        }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
            if (w < warp) before += s_wsum[w];
        if (d < bins) {
            uint32_t run = before + incl - tot;
            s_lstart[d] = run;
            if (LOOKBACK) {
                uint32_t* mine = status + (size_t)block * kMaxBins + d;
                volatile uint32_t* st = status;
                if (block > 0) atomicExch(mine, kFlagPartial | tot);
                uint32_t excl = 0;
                int guard = 0;
                for (int b = block - 1; b >= 0;) {
                    const uint32_t sv = st[(size_t)b * kMaxBins + d];
                    const uint32_t flag = sv & ~kValueMask;
                    if (flag == 0) {
                        if (++guard > (1 << 24)) __trap();  // never expected: fail the launch loudly instead of hanging or sorting wrongly
                        continue;
                    }
                    excl += sv & kValueMask;
                    if (flag == kFlagInclusive) break;
                    --b;
                }
                atomicExch(mine, kFlagInclusive | (excl + tot));
                s_gbase[d] = block_off[d] + excl;
            } else {
                s_gbase[d] = block_off[(size_t)d * num_blocks + block];
            }
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t c = s_cnt[w][d];
                s_cnt[w][d] = run;
                run += c;
            }
        }
    }
    __syncthreads();
    // phase 3: place at the block-local sorted position (s_cnt now holds each warp's base per digit)
#pragma unroll
    for (int k = 0; k < kItemsPerThread; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = ((uint32_t)key[k] >> shift) & mask;
            const uint32_t pos = s_cnt[warp][d] + rank[k];
            s_keys[pos] = key[k];
            s_vals[pos] = IOTA ? (uint32_t)i : vals_in[i];
        }
    }
    __syncthreads();
    // phase 4: coalesced write-out, local position -> global position of the same digit run
    const int count = min(kChunk, n - bbase);
    for (int pos = threadIdx.x; pos < count; pos += kSortThreads) {
        const KeyT kk = s_keys[pos];
        const uint32_t d = ((uint32_t)kk >> shift) & mask;
        const uint32_t g = s_gbase[d] + ((uint32_t)pos - s_lstart[d]);
        keys_out[g] = kk;
        vals_out[g] = s_vals[pos];
    }
}

static bool use_lookback() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MRGS_RADIX_LOOKBACK");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// scratch layout (uint32): totals[4][256] | starts[4][256] | tickets[64] | per-pass block tables
constexpr size_t kScratchHeader = 4 * kMaxBins + 4 * kMaxBins + 64;

template <typename KeyT>
int radix_sort_pairs(KeyT* keys_a, KeyT* keys_b, uint32_t* vals_a, uint32_t* vals_b, int n, int total_bits,
                     bool first_pass_iota, uint32_t* scratch, cudaStream_t stream, KeyT** keys_final,
                     uint32_t** vals_final, const uint32_t* n_dev = nullptr) {
    // as few passes of <= 8 bits as possible, evenly split
    const int passes = (total_bits + 7) / 8;
    const int num_blocks = (n + kChunk - 1) / kChunk;
    PassPlan plan{};
    plan.passes = passes;
    for (int pass = 0, shift = 0; pass < passes; ++pass) {
        plan.bits[pass] = (total_bits - shift + (passes - pass) - 1) / (passes - pass);
        plan.shift[pass] = shift;
        shift += plan.bits[pass];
    }
    uint32_t* totals = scratch;
    uint32_t* starts = scratch + 4 * kMaxBins;
    uint32_t* tickets = scratch + 8 * kMaxBins;
    uint32_t* tables = scratch + kScratchHeader;
    const size_t table = (size_t)kMaxBins * num_blocks;
    KeyT* kin = keys_a;
    KeyT* kout = keys_b;
    uint32_t* vin = vals_a;
    uint32_t* vout = vals_b;
    int launches = 0;
    // the look-back status word keeps 30 value bits: larger inputs take the three-kernel passes (callers that need the
    // device-side count, i.e. the optimistic path, cap their capacity below 2^30: api.cu)
    const bool lookback = use_lookback() && (n < (1 << 30) || n_dev != nullptr);
    if (lookback) {
        if (cudaMemsetAsync(scratch, 0, (kScratchHeader + (size_t)passes * table) * sizeof(uint32_t), stream) != cudaSuccess)
            return -1;
        radix_multi_hist_kernel<KeyT><<<num_blocks, kSortThreads, 0, stream>>>(kin, n, n_dev, plan, totals);
        radix_digit_scan_kernel<<<passes, 256, 0, stream>>>(totals, starts);
        launches += 2;
        for (int pass = 0; pass < passes; ++pass) {
            uint32_t* status = tables + (size_t)pass * table;
            if (pass == 0 && first_pass_iota)
                radix_scatter_kernel<KeyT, true, true><<<num_blocks, kSortThreads, 0, stream>>>(
                    kin, vin, kout, vout, n, plan.shift[pass], plan.bits[pass], starts + pass * kMaxBins, num_blocks, status,
                    tickets + pass, n_dev);
            else
                radix_scatter_kernel<KeyT, false, true><<<num_blocks, kSortThreads, 0, stream>>>(
                    kin, vin, kout, vout, n, plan.shift[pass], plan.bits[pass], starts + pass * kMaxBins, num_blocks, status,
                    tickets + pass, n_dev);
            ++launches;
            KeyT* tk = kin; kin = kout; kout = tk;
            uint32_t* tv = vin; vin = vout; vout = tv;
        }
    } else {
        if (cudaMemsetAsync(totals, 0, 4 * kMaxBins * sizeof(uint32_t), stream) != cudaSuccess) return -1;
        for (int pass = 0; pass < passes; ++pass) {
            const int bits = plan.bits[pass], shift = plan.shift[pass];
            uint32_t* tot = totals + (size_t)pass * kMaxBins;
            radix_hist_kernel<KeyT><<<num_blocks, kSortThreads, 0, stream>>>(kin, n, shift, bits, tables, tot, num_blocks);
            radix_row_scan_kernel<<<1 << bits, 256, 0, stream>>>(tables, tot, 1 << bits, num_blocks);
            if (pass == 0 && first_pass_iota)
                radix_scatter_kernel<KeyT, true, false><<<num_blocks, kSortThreads, 0, stream>>>(
                    kin, vin, kout, vout, n, shift, bits, tables, num_blocks, nullptr, nullptr, nullptr);
            else
                radix_scatter_kernel<KeyT, false, false><<<num_blocks, kSortThreads, 0, stream>>>(
                    kin, vin, kout, vout, n, shift, bits, tables, num_blocks, nullptr, nullptr, nullptr);
            launches += 3;
            KeyT* tk = kin; kin = kout; kout = tk;
            uint32_t* tv = vin; vin = vout; vout = tv;
        }
    }
    *keys_final = kin;
    *vals_final = vin;
    return launches;
}

// inclusive scan of tiles_touched gathered through the depth-sorted surfel order -------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t mine, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < warp) before += c;
        sum += c;
    }
    total = sum;
    return before + incl - mine;
}

__global__ void __launch_bounds__(kScanThreads)
gather_reduce_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched, int n,
                     uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) mine += tiles_touched[order[base + k]];
    uint32_t total;
    block_exclusive_scan_256(mine, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// `block_sums` are the RAW per-block totals of gather_reduce_kernel: every block adds up the totals of the blocks
// before it on its own (a few hundred values, one or two per thread) — no separate scan launch in between.
__global__ void __launch_bounds__(kScanThreads)
gather_scan_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched, int n,
                   const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ offsets_incl) {
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ uint32_t s_before[kScanThreads / 32];
    uint32_t before = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanThreads) before += block_sums[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(kFullMask, before, o);
    if ((threadIdx.x & 31) == 0) s_before[threadIdx.x >> 5] = before;
    __syncthreads();
    uint32_t block_offset = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) block_offset += s_before[w];
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? tiles_touched[order[base + k]] : 0u;
        mine += v[k];
    }
    uint32_t total;
    uint32_t run = block_offset + block_exclusive_scan_256(mine, s_warp, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        run += v[k];
        if (base + k < n) offsets_incl[base + k] = run;
    }
}

// A warp expands 32 consecutive surfels of the depth order. Their instances form ONE contiguous
// output segment, so lanes stride over output positions (coalesced 2-byte / 4-byte stores) and find the
// owning surfel with a 5-step binary search over the warp's 32 inclusive offsets.
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched,
                      const uint2* __restrict__ rect, const uint32_t* __restrict__ offsets_incl,
                      uint16_t* __restrict__ keys, uint32_t* __restrict__ values, int grid_x, uint32_t capacity) {
    __shared__ uint32_t s_end[8][32];
    __shared__ uint32_t s_id[8][32];
    __shared__ uint2 s_rect[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = (blockIdx.x * 8 + warp) * 32;
    if (k0 >= P) return;
    const int k = k0 + lane;
    const uint32_t seg_begin = (k0 == 0) ? 0u : offsets_incl[k0 - 1];
    uint32_t end = seg_begin;
    uint32_t id = 0;
    if (k < P) {
        end = offsets_incl[k];
        id = order[k];
    }
    const int n_valid = min(32, P - k0);
    const uint32_t seg_end = __shfl_sync(kFullMask, end, n_valid - 1);
    if (k >= P) end = seg_end;  // lanes past the last surfel own an empty range at the segment end
    const uint32_t prev_end = __shfl_up_sync(kFullMask, end, 1);
    const uint32_t begin_mine = lane ? prev_end : seg_begin;
    s_end[warp][lane] = end;
    s_id[warp][lane] = id;
    if (k < P && end > begin_mine) s_rect[warp][lane] = rect[id];
    __syncwarp();
    const uint32_t stop = min(seg_end, capacity);   // never write past the buffer the caller sized
    for (uint32_t pos = seg_begin + lane; pos < stop; pos += 32) {
        // first lane j with s_end[j] > pos
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (s_end[warp][lo + step - 1] <= pos) lo += step;
        const uint32_t start_j = lo ? s_end[warp][lo - 1] : seg_begin;
        const uint2 r = s_rect[warp][lo];
        const int x0 = r.x & 0xffff, y0 = r.x >> 16, x1 = r.y & 0xffff;
        const int w = x1 - x0;
        const int t = (int)(pos - start_j);
        const int ty = t / w, tx = t - ty * w;
        keys[pos] = (uint16_t)((y0 + ty) * grid_x + x0 + tx);
        values[pos] = s_id[warp][lo];
    }
}

// Tile boundaries in the sorted 16-bit tile ids. Every thread owns 8 consecutive keys (one 16-byte load) plus the
// key before them, so the 2-byte keys are read with full-width coalesced accesses.
__global__ void __launch_bounds__(256)
identify_tile_ranges_kernel(int R, const uint32_t* __restrict__ n_dev, const uint16_t* __restrict__ keys,
                            uint2* __restrict__ ranges) {
    if (n_dev != nullptr) R = min(R, (int)*n_dev);
    const int first = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (first >= R) return;
    uint32_t k[8];
    if (first + 8 <= R) {
        const uint4 raw = *reinterpret_cast<const uint4*>(keys + first);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) k[e] = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) k[e] = (first + e < R) ? (uint32_t)keys[first + e] : 0u;
    }
    uint32_t prev = (first == 0) ? 0u : (uint32_t)keys[first - 1];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int idx = first + e;
        if (idx < R) {
            const uint32_t tile = k[e];
            if (idx == 0) {
                ranges[tile].x = 0;
            } else if (tile != prev) {
                ranges[prev].y = idx;
                ranges[tile].x = idx;
            }
            if (idx == R - 1) ranges[tile].y = R;
            prev = tile;
        }
    }
}

}  // namespace

int sort_blocks(int64_t n) { return (int)((n + kChunk - 1) / kChunk); }
size_t sort_hist_bytes(int64_t n) {
    const size_t blocks = (size_t)(sort_blocks(n) > 0 ? sort_blocks(n) : 1);
    return (kScratchHeader + 4 * (size_t)kMaxBins * blocks) * sizeof(uint32_t);
}
int scan_blocks(int n) { return (n + kScanChunk - 1) / kScanChunk; }

int depth_sort(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int P, uint32_t* block_hist,
               cudaStream_t stream, uint32_t** order, int* launches) {
    uint32_t* kf;
    *launches = radix_sort_pairs<uint32_t>(keys_a, keys_b, vals_a, vals_b, P, 32, true, block_hist, stream, &kf, order);
    if (*launches < 0) {
        *launches = 0;
        set_error("depth_sort: cudaMemsetAsync of the sort scratch failed");
        return MRGS_ERR_CUDA;
    }
    return MRGS_OK;
}

int offsets_in_order(const uint32_t* order, const uint32_t* tiles_touched, int P, uint32_t* block_sums,
                     uint32_t* offsets_incl, cudaStream_t stream) {
    const int nb = scan_blocks(P);
    gather_reduce_kernel<<<nb, kScanThreads, 0, stream>>>(order, tiles_touched, P, block_sums);
    gather_scan_kernel<<<nb, kScanThreads, 0, stream>>>(order, tiles_touched, P, block_sums, offsets_incl);
    return MRGS_OK;
}

void launch_emit_instances(int P, const uint32_t* order, const uint32_t* tiles_touched, const uint2* rect,
                           const uint32_t* offsets_incl, uint16_t* keys, uint32_t* values, int grid_x,
                           uint32_t capacity, cudaStream_t stream) {
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, order, tiles_touched, rect, offsets_incl, keys, values,
                                                              grid_x, capacity);
}

int tile_sort(uint16_t* keys_a, uint16_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int R, int tile_bits,
              uint32_t* block_hist, cudaStream_t stream, uint16_t** keys_sorted, uint32_t** vals_sorted, int* launches,
              const uint32_t* R_dev) {
    *launches = radix_sort_pairs<uint16_t>(keys_a, keys_b, vals_a, vals_b, R, tile_bits, false, block_hist, stream,
                                           keys_sorted, vals_sorted, R_dev);
    if (*launches < 0) {
        *launches = 0;
        set_error("tile_sort: cudaMemsetAsync of the sort scratch failed");
        return MRGS_ERR_CUDA;
    }
    return MRGS_OK;
}

void launch_identify_tile_ranges(int R, const uint32_t* R_dev, const uint16_t* keys, uint2* ranges, cudaStream_t stream) {
    identify_tile_ranges_kernel<<<(R + 2047) / 2048, 256, 0, stream>>>(R, R_dev, keys, ranges);
}

bool radix_lookback_enabled() { return use_lookback(); }

}  // namespace mrgs
