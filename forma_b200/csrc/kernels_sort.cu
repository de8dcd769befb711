// Stage 3: on-device LSD radix sort of the packed u64 pixel segments on key
// bits [20, 64) — replaces crumsort::ParCrumSort (cpu/rasterizer.rs:162-164)
// and the WGSL block-merge sort (gpu/conveyor_sort/sort.wgsl).
//
// * Plan: the 44 key bits are three fields (tile_y | tile_x | layer, or
//   tile_y | layer | tile_x for the painter's carry pass). Only the low bits of
//   each field that are actually set in some key take part: a 1-thread kernel
//   turns the OR of all keys into a *pass plan* — digits of <= 8 bits over the
//   concatenation of the used field bits (paris@4K: 16 + 8 + 8 = 32 bits = 4
//   passes instead of 6). The plan lives in device memory; pass kernels beyond
//   the planned count return immediately, so no host round trip is needed.
// * One upfront histogram kernel reads the keys once and counts the digits of
//   all planned passes in shared memory.
// * Every pass is one "onesweep" kernel: a CTA ranks a tile of keys with warp
//   match/popc (stable), obtains its global digit offsets by decoupled
//   look-back over the tiles before it, stages the tile in shared memory in
//   digit order and writes it out with coalesced stores: keys are read once and
//   written once per pass (16 B/key).
//
// The same kernels sort (key, u32 payload) pairs for the painter's cell and
// entry tables (smaller tiles: those sorts are latency-, not bandwidth-bound).
#include "cuda_common.cuh"
#include "kernels.h"

namespace forma {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kMaxPasses = 6;  // ceil(44 / 8)

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;

// One digit = up to three bit runs of the key, concatenated.
struct DigitSpec {
    uint8_t shift[3];
    uint8_t width[3];
    uint8_t lsh[3];
    uint8_t bits;
};
struct SortPlan {
    uint32_t n_passes;
    uint32_t total_bits;
    DigitSpec pass[kMaxPasses];
};

__device__ __forceinline__ uint32_t digit_of(uint64_t key, const DigitSpec& d) {
    uint32_t v = ((uint32_t)(key >> d.shift[0]) & ((1u << d.width[0]) - 1u)) << d.lsh[0];
    v |= ((uint32_t)(key >> d.shift[1]) & ((1u << d.width[1]) - 1u)) << d.lsh[1];
    v |= ((uint32_t)(key >> d.shift[2]) & ((1u << d.width[2]) - 1u)) << d.lsh[2];
    return v;
}

// OR of all keys (grid-stride), for sorts whose producer did not compute it.
__global__ void __launch_bounds__(256) key_or_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                     unsigned long long* __restrict__ key_or) {
    uint64_t acc = 0;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) acc |= keys[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(kFullMask, acc, o);
    if (lane_id() == 0 && acc) atomicOr(key_or, (unsigned long long)acc);
}

// Field f occupies key bits [pos[f], pos[f] + maxw[f]); f = 0 is least significant.
__global__ void sort_plan_kernel(const unsigned long long* __restrict__ key_or, uint64_t extra_or, uint3 pos, uint3 maxw,
                                 SortPlan* __restrict__ plan) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t m = *key_or | extra_or;
    const uint32_t p[3] = {pos.x, pos.y, pos.z}, mw[3] = {maxw.x, maxw.y, maxw.z};
    uint32_t w[3], total = 0;
    for (int f = 0; f < 3; ++f) {
        uint32_t field = (uint32_t)(m >> p[f]) & ((1u << mw[f]) - 1u);
        w[f] = field ? 32u - (uint32_t)__clz((int)field) : 0u;
        total += w[f];
    }
    uint32_t n_passes = (total + kRadixBits - 1) / kRadixBits;
    SortPlan out;
    out.n_passes = n_passes;
    out.total_bits = total;
    // Distribute the bits evenly over the passes (e.g. 38 bits -> 8,8,8,7,7).
    uint32_t lo = 0;  // position in the virtual (compacted) key
    for (uint32_t k = 0; k < kMaxPasses; ++k) {
        DigitSpec d;
        for (int r = 0; r < 3; ++r) d.shift[r] = d.width[r] = d.lsh[r] = 0;
        d.bits = 0;
        if (k < n_passes) {
            uint32_t remaining = total - lo, left = n_passes - k;
            uint32_t bits = (remaining + left - 1) / left;
            d.bits = (uint8_t)bits;
            uint32_t hi = lo + bits, base = 0, got = 0;
            int r = 0;
            for (int f = 0; f < 3; ++f) {  // intersect [lo, hi) with field f's slice [base, base + w[f])
                uint32_t a = max(lo, base), b = min(hi, base + w[f]);
                if (b > a) {
                    d.shift[r] = (uint8_t)(p[f] + (a - base));
                    d.width[r] = (uint8_t)(b - a);
                    d.lsh[r] = (uint8_t)got;
                    got += b - a;
                    ++r;
                }
                base += w[f];
            }
            lo = hi;
        }
        out.pass[k] = d;
    }
    *plan = out;
}

// Counts every digit of every planned pass in one read of the keys.
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                                const SortPlan* __restrict__ plan_g,
                                                                uint32_t* __restrict__ hist /*[passes][256]*/) {
    __shared__ uint32_t s_hist[kMaxPasses][kRadix];
    __shared__ SortPlan plan;
    if (threadIdx.x == 0) plan = *plan_g;
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kSortThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t np = plan.n_passes;
    if (np == 0) return;
    uint32_t stride = gridDim.x * kSortThreads;
    for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
        uint64_t k = keys[i];
        for (uint32_t p = 0; p < np; ++p) {
            // Warp-aggregated increment: neighbouring segments share tile / layer digits.
            uint32_t d = digit_of(k, plan.pass[p]);
            uint32_t peers = __match_any_sync(__activemask(), d);
            if ((uint32_t)(__ffs(peers) - 1) == lane_id()) atomicAdd(&s_hist[p][d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kSortThreads) {
        uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// Exclusive scan of each pass's 256-bin histogram (one CTA per pass).
__global__ void __launch_bounds__(kRadix) radix_scan_hist_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_tot[kRadix / 32];
    uint32_t* h = hist + blockIdx.x * kRadix;
    uint32_t v = h[threadIdx.x];
    uint32_t incl = warp_inclusive_scan(v);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) base += warp_tot[w];
    h[threadIdx.x] = base + incl - v;
}

template <bool kPairs, int kItems>
__global__ void __launch_bounds__(kSortThreads, kItems == 16 ? 4 : 6)
    onesweep_pass_kernel(uint64_t* __restrict__ buf_a, uint64_t* __restrict__ buf_b, uint32_t* __restrict__ val_a,
                         uint32_t* __restrict__ val_b, uint32_t n, uint32_t pass, const SortPlan* __restrict__ plan_g,
                         const uint32_t* __restrict__ hist /*[passes][256], exclusive*/,
                         uint32_t* __restrict__ lookback_all /*[passes][tiles][256], zeroed*/,
                         uint32_t* __restrict__ tile_counters, uint32_t tiles) {
    constexpr int kTileKeys = kSortThreads * kItems;
    __shared__ uint64_t s_keys[kTileKeys];
    __shared__ uint32_t s_warp_hist[kSortWarps][kRadix];
    __shared__ uint32_t s_digit_start[kRadix];
    __shared__ uint32_t s_global_base[kRadix];
    __shared__ uint32_t s_warp_tot[kSortWarps];
    __shared__ uint32_t s_tile;
    __shared__ DigitSpec s_spec;

    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    if (pass >= plan_g->n_passes) return;  // pass not planned: nothing to do
    if (t == 0) {
        s_tile = atomicAdd(tile_counters + pass, 1u);
        s_spec = plan_g->pass[pass];
    }
    for (int i = t; i < kSortWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const DigitSpec spec = s_spec;
    const uint64_t* __restrict__ keys_in = (pass & 1u) ? buf_b : buf_a;
    uint64_t* __restrict__ keys_out = (pass & 1u) ? buf_a : buf_b;
    const uint32_t* __restrict__ vals_in = (pass & 1u) ? val_b : val_a;
    uint32_t* __restrict__ vals_out = (pass & 1u) ? val_a : val_b;
    const uint32_t* __restrict__ global_offsets = hist + pass * kRadix;
    volatile uint32_t* lb = lookback_all + (size_t)pass * tiles * kRadix;

    const uint32_t tile = s_tile;
    const uint32_t base = tile * (uint32_t)kTileKeys;
    const uint32_t valid = min((uint32_t)kTileKeys, n - base);

    // Warp-striped load: warp w owns keys [w*32*kItems, (w+1)*32*kItems) of the tile.
    uint64_t key[kItems];
    const uint32_t warp_base = base + warp * (32u * kItems);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        uint32_t idx = warp_base + i * 32u + lane;
        key[i] = idx < n ? keys_in[idx] : ~0ull;
    }

    // Stable rank of every key among the keys of its warp with the same digit.
    // Out-of-range slots of the last tile get the largest digit so that they
    // rank after every real key.
    uint32_t rank[kItems];
    uint32_t dig[kItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t max_digit = (1u << spec.bits) - 1u;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        uint32_t idx = warp_base + i * 32u + lane;
        uint32_t d = idx < n ? digit_of(key[i], spec) : max_digit;
        dig[i] = d;
        uint32_t peers = __match_any_sync(kFullMask, d);
        uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(kFullMask, old, leader);
        rank[i] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // Thread t owns digit t: exclusive offsets of each warp, tile count.
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        uint32_t c = s_warp_hist[w][t];
        s_warp_hist[w][t] = count;
        count += c;
    }
    // Exclusive scan of the tile's digit counts (local layout in shared memory).
    uint32_t incl = warp_inclusive_scan(count);
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t dstart = incl - count;
    for (uint32_t w = 0; w < warp; ++w) dstart += s_warp_tot[w];
    s_digit_start[t] = dstart;

    // Decoupled look-back: exclusive count of digit t over all previous tiles.
    // The padding slots of the last tile are counted under max_digit; no tile
    // comes after the last one, so nobody consumes that aggregate.
    {
        uint32_t prefix = 0;
        if (tile == 0) {
            lb[t] = kFlagInclusive | count;
        } else {
            lb[tile * kRadix + t] = kFlagAggregate | count;
            int32_t p = (int32_t)tile - 1;
            while (true) {
                uint32_t v = lb[(uint32_t)p * kRadix + t];
                uint32_t flag = v & kFlagMask;
                if (flag == 0) continue;  // not published yet
                prefix += v & kValueMask;
                if (flag == kFlagInclusive) break;
                --p;
            }
            lb[tile * kRadix + t] = kFlagInclusive | (prefix + count);
        }
        s_global_base[t] = global_offsets[t] + prefix - dstart;
    }
    __syncthreads();

    // Stage the tile in shared memory in digit order.
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        uint32_t pos = s_digit_start[dig[i]] + s_warp_hist[warp][dig[i]] + rank[i];
        s_keys[pos] = key[i];
        rank[i] = pos;
    }
    __syncthreads();

    uint32_t out_idx[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        uint32_t p = t + k * kSortThreads;
        out_idx[k] = 0xFFFFFFFFu;
        if (p < valid) {
            uint64_t kk = s_keys[p];
            uint32_t o = s_global_base[digit_of(kk, spec)] + p;
            keys_out[o] = kk;
            out_idx[k] = o;
        }
    }
    if (kPairs) {
        __syncthreads();
        uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys);
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            uint32_t idx = warp_base + i * 32u + lane;
            if (idx < n) s_vals[rank[i]] = vals_in[idx];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            uint32_t p = t + k * kSortThreads;
            if (out_idx[k] != 0xFFFFFFFFu) vals_out[out_idx[k]] = s_vals[p];
        }
    }
}

// After an odd number of passes the data sits in the scratch buffers.
__global__ void __launch_bounds__(256) sort_copy_back_kernel(uint64_t* __restrict__ keys, const uint64_t* __restrict__ keys_tmp,
                                                             uint32_t* __restrict__ vals, const uint32_t* __restrict__ vals_tmp,
                                                             uint32_t n, const SortPlan* __restrict__ plan) {
    if ((plan->n_passes & 1u) == 0u) return;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        keys[i] = keys_tmp[i];
        if (vals) vals[i] = vals_tmp[i];
    }
}

static uint32_t tiles_for(uint32_t n, int items) { return (n + kSortThreads * items - 1) / (kSortThreads * items); }
static int items_for(uint32_t n) { return n >= (1u << 21) ? 16 : 4; }

// scratch layout (u32 words): plan (64 words) | key_or (2) | pad (2) | hist[6][256] | tile_counter[6 + pad 2]
//                             | lookback[6][tiles][256]
size_t radix_scratch_bytes(uint32_t n) {
    size_t words = 64 + 4 + (size_t)kMaxPasses * kRadix + 8 + (size_t)kMaxPasses * tiles_for(n, items_for(n)) * kRadix;
    return words * sizeof(uint32_t) + 256;
}

template <bool kPairs, int kItems>
static void launch_passes(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, uint32_t n,
                          const SortPlan* plan, const uint32_t* hist, uint32_t* lookback, uint32_t* counters,
                          uint32_t tiles, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {  // let 4-6 CTAs of 11-43 KB share one SM's shared memory
        cudaFuncSetAttribute(onesweep_pass_kernel<kPairs, kItems>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    for (uint32_t p = 0; p < (uint32_t)kMaxPasses; ++p)
        onesweep_pass_kernel<kPairs, kItems><<<tiles, kSortThreads, 0, stream>>>(keys, keys_tmp, vals, vals_tmp, n, p, plan,
                                                                                hist, lookback, counters, tiles);
}

int launch_radix_sort(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, uint32_t n,
                      const KeyLayout& layout, const unsigned long long* key_or_device, void* scratch,
                      cudaStream_t stream) {
    if (n < 2) return 0;
    static_assert(sizeof(SortPlan) <= 64 * sizeof(uint32_t), "plan must fit its scratch slot");
    const int items = items_for(n);
    const uint32_t tiles = tiles_for(n, items);
    uint32_t* words = static_cast<uint32_t*>(scratch);
    SortPlan* plan = reinterpret_cast<SortPlan*>(words);
    unsigned long long* key_or = reinterpret_cast<unsigned long long*>(words + 64);
    uint32_t* hist = words + 68;
    uint32_t* counters = hist + kMaxPasses * kRadix;
    uint32_t* lookback = counters + 8;
    size_t total_words = 68 + (size_t)kMaxPasses * kRadix + 8 + (size_t)kMaxPasses * tiles * kRadix;
    cudaMemsetAsync(scratch, 0, total_words * sizeof(uint32_t), stream);
    int launches = 0;
    if (!key_or_device) {
        key_or_kernel<<<min((n + 255u) / 256u, 148u * 8u), 256, 0, stream>>>(keys, n, key_or);
        key_or_device = key_or;
        ++launches;
    }
    sort_plan_kernel<<<1, 32, 0, stream>>>(key_or_device, layout.extra_or, make_uint3(layout.pos[0], layout.pos[1], layout.pos[2]),
                                           make_uint3(layout.maxw[0], layout.maxw[1], layout.maxw[2]), plan);
    uint32_t hist_blocks = min(tiles_for(n, 16) * 4u, 148u * 8u);
    radix_hist_kernel<<<hist_blocks, kSortThreads, 0, stream>>>(keys, n, plan, hist);
    radix_scan_hist_kernel<<<kMaxPasses, kRadix, 0, stream>>>(hist);
    launches += 3;
    if (vals) {
        if (items == 16) launch_passes<true, 16>(keys, keys_tmp, vals, vals_tmp, n, plan, hist, lookback, counters, tiles, stream);
        else launch_passes<true, 4>(keys, keys_tmp, vals, vals_tmp, n, plan, hist, lookback, counters, tiles, stream);
    } else {
        if (items == 16) launch_passes<false, 16>(keys, keys_tmp, nullptr, nullptr, n, plan, hist, lookback, counters, tiles, stream);
        else launch_passes<false, 4>(keys, keys_tmp, nullptr, nullptr, n, plan, hist, lookback, counters, tiles, stream);
    }
    launches += kMaxPasses;
    sort_copy_back_kernel<<<min((n + 255u) / 256u, 148u * 16u), 256, 0, stream>>>(keys, keys_tmp, vals, vals_tmp, n, plan);
    ++launches;
    return launches;
}

}  // namespace forma
