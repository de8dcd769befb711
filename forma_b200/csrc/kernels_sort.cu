// Stage 3: on-device LSD radix sort of the packed u64 pixel segments on key
// bits [20, 64) — replaces crumsort::ParCrumSort (cpu/rasterizer.rs:162-164)
// and the WGSL block-merge sort (gpu/conveyor_sort/sort.wgsl).
//
// One upfront histogram kernel reads the keys once and counts all digits of
// all passes in shared memory; every pass is then a single "onesweep" kernel:
// a CTA ranks a 4096-key tile with warp-level match/popc (stable), obtains its
// global digit offsets by decoupled look-back over the tiles before it, stages
// the tile in shared memory in digit order and writes it out with coalesced
// stores. Per pass the keys are read once and written once (16 B/key).
//
// The same kernels sort (key, u32 payload) pairs for the painter's cell and
// entry tables.
#include "cuda_common.cuh"
#include "kernels.h"

namespace forma {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys
constexpr int kNumPasses = (64 - kSortShift + kRadixBits - 1) / kRadixBits;  // 6

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int pass) {
    return (uint32_t)(key >> (kSortShift + pass * kRadixBits)) & (kRadix - 1);
}

// Counts every digit of every pass in one read of the keys.
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                                uint32_t* __restrict__ hist /*[passes][256]*/) {
    __shared__ uint32_t s_hist[kNumPasses][kRadix];
    for (int i = threadIdx.x; i < kNumPasses * kRadix; i += kSortThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t stride = gridDim.x * kSortThreads;
    for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += stride) {
        uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < kNumPasses; ++p) atomicAdd(&s_hist[p][digit_of(k, p)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNumPasses * kRadix; i += kSortThreads) {
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

template <bool kPairs>
__global__ void __launch_bounds__(kSortThreads)
    onesweep_pass_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                         const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, uint32_t n, int pass,
                         const uint32_t* __restrict__ global_offsets /*[256], exclusive*/,
                         uint32_t* __restrict__ lookback /*[tiles][256], zeroed*/, uint32_t* __restrict__ tile_counter) {
    __shared__ uint64_t s_keys[kSortTile];
    __shared__ uint32_t s_warp_hist[kSortWarps][kRadix];
    __shared__ uint32_t s_digit_start[kRadix];
    __shared__ uint32_t s_global_base[kRadix];
    __shared__ uint32_t s_warp_tot[kSortWarps];
    __shared__ uint32_t s_tile;

    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31u;
    if (t == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = t; i < kSortWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * (uint32_t)kSortTile;
    const uint32_t valid = min((uint32_t)kSortTile, n - base);

    // Warp-striped load: warp w owns keys [w*512, (w+1)*512) of the tile.
    uint64_t key[kSortItems];
    const uint32_t warp_base = base + warp * (32u * kSortItems);
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        uint32_t idx = warp_base + i * 32u + lane;
        key[i] = idx < n ? keys_in[idx] : ~0ull;
    }

    // Stable rank of every key among the keys of its warp with the same digit.
    uint32_t rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        uint32_t d = digit_of(key[i], pass);
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
    {
        volatile uint32_t* lb = lookback;
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
    for (int i = 0; i < kSortItems; ++i) {
        uint32_t d = digit_of(key[i], pass);
        uint32_t pos = s_digit_start[d] + s_warp_hist[warp][d] + rank[i];
        s_keys[pos] = key[i];
        rank[i] = pos;
    }
    __syncthreads();

    uint32_t out_idx[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        uint32_t p = t + k * kSortThreads;
        out_idx[k] = 0xFFFFFFFFu;
        if (p < valid) {
            uint64_t kk = s_keys[p];
            uint32_t o = s_global_base[digit_of(kk, pass)] + p;
            keys_out[o] = kk;
            out_idx[k] = o;
        }
    }
    if (kPairs) {
        __syncthreads();
        uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys);
#pragma unroll
        for (int i = 0; i < kSortItems; ++i) {
            uint32_t idx = warp_base + i * 32u + lane;
            if (idx < n) s_vals[rank[i]] = vals_in[idx];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            uint32_t p = t + k * kSortThreads;
            if (out_idx[k] != 0xFFFFFFFFu) vals_out[out_idx[k]] = s_vals[p];
        }
    }
}

static uint32_t num_tiles(uint32_t n) { return (n + kSortTile - 1) / kSortTile; }

// scratch layout: hist[passes][256] | tile_counter[passes] | lookback[passes][tiles][256]
size_t radix_scratch_bytes(uint32_t n) {
    size_t words = (size_t)kNumPasses * kRadix + kNumPasses + (size_t)kNumPasses * num_tiles(n) * kRadix;
    return words * sizeof(uint32_t) + 256;
}

int launch_radix_sort(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, uint32_t n,
                      void* scratch, cudaStream_t stream) {
    if (n < 2) return 0;
    uint32_t tiles = num_tiles(n);
    uint32_t* hist = static_cast<uint32_t*>(scratch);
    uint32_t* counters = hist + kNumPasses * kRadix;
    uint32_t* lookback = counters + kNumPasses;
    size_t words = (size_t)kNumPasses * kRadix + kNumPasses + (size_t)kNumPasses * tiles * kRadix;
    cudaMemsetAsync(scratch, 0, words * sizeof(uint32_t), stream);
    int launches = 0;
    uint32_t hist_blocks = min(tiles * 4u, 148u * 8u);
    radix_hist_kernel<<<hist_blocks, kSortThreads, 0, stream>>>(keys, n, hist);
    radix_scan_hist_kernel<<<kNumPasses, kRadix, 0, stream>>>(hist);
    launches += 2;
    uint64_t* kin = keys;
    uint64_t* kout = keys_tmp;
    uint32_t* vin = vals;
    uint32_t* vout = vals_tmp;
    for (int p = 0; p < kNumPasses; ++p) {
        if (vals)
            onesweep_pass_kernel<true><<<tiles, kSortThreads, 0, stream>>>(
                kin, kout, vin, vout, n, p, hist + p * kRadix, lookback + (size_t)p * tiles * kRadix, counters + p);
        else
            onesweep_pass_kernel<false><<<tiles, kSortThreads, 0, stream>>>(
                kin, kout, nullptr, nullptr, n, p, hist + p * kRadix, lookback + (size_t)p * tiles * kRadix,
                counters + p);
        ++launches;
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    // kNumPasses is even: the sorted data is back in `keys` / `vals`.
    static_assert(kNumPasses % 2 == 0, "ping-pong must end in the caller's buffer");
    return launches;
}

}  // namespace forma
