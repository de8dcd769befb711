// Stage 4: per-tile coverage accumulation and compositing to RGBA8 — replaces
// forma/src/cpu/painter/{mod.rs, layer_workbench/, styling.rs}.
//
// The reference walks each tile row left to right, carrying every layer's
// winding "cover" (16 x i8, one per pixel row) from tile to tile in a queue
// (cpu/painter/mod.rs:486-568, layer_workbench/mod.rs:196-342). To paint tiles
// independently the carries are materialised first:
//
//   cells     runs of sorted segments with equal (tile_y, tile_x, layer)
//             (cells_scan: one pass, head bits -> cell_start by decoupled look-back)
//   covers    per cell: sum of segment covers by local_y (wrapping i8), its key (same pass)
//   re-sort   cell ids stably by the layer bits only -> (layer, tile_y, tile_x)
//   carries   per (tile_y, layer) group a running sum -> carry-in of every cell
//             and "carry-only" entries for the tiles a layer spans without
//             segments (layer_workbench/mod.rs:213-234,328-336), generated in
//             that order and stably sorted by the tile digits only
//   entries   cells merged with the carry-only entries by rank into 64-byte
//             records ordered by (tile_y, tile_x, layer), + per-tile ranges
//   paint     kernels_painter.cu: one warp per tile
//
// Integer semantics: areas wrap at i16 and covers at i8 in the reference;
// sums are formed in i32 / packed bytes and truncated where the reference
// widens them (truncation commutes with wrapping addition).
#include "paint_common.cuh"

namespace forma {

// ---------------------------------------------------------------------------
// Cells
// ---------------------------------------------------------------------------
constexpr int kCellThreads = 256;

__device__ __forceinline__ bool is_cell_head(const uint64_t* __restrict__ segs, uint32_t i) {
    return i == 0 || (segs[i] >> kSortShift) != (segs[i - 1] >> kSortShift);
}

// A CTA scans kCellItems x 256 consecutive segments (warp-striped so that every
// load is coalesced and the element order inside a warp is item-major).
constexpr int kCellItems = 8;
constexpr int kCellTile = kCellThreads * kCellItems;

// Cells in one pass over the sorted segments. A CTA takes a tile of 2048 segments (by
// ticket), keeps them in registers and
//   1. marks the heads (a segment starts a cell when its key differs from its predecessor's),
//   2. gets the number of cells before its tile by decoupled look-back (state[t] = flag |
//      running count; state[tiles] = ticket counter; 32 predecessors per round),
//   3. writes cell_start for its heads,
//   4. sums the covers of every cell that lies entirely inside the tile, by local_y, in shared
//      memory (acc_segment's cover part + cover_carry, cpu/painter/mod.rs:257-271,
//      layer_workbench/mod.rs:218-224) and writes the cell's records: its cover (16 x i8),
//      its key and the (tile_y, layer, tile_x) key of the carry pass.
// Cells that cross a tile boundary (at most two per tile) are left to cells_boundary_kernel.
// Positions >= cap are not written: the kernel may run before the host knows the cell count
// (Renderer::render repeats both kernels after growing the buffers if it was too small).
constexpr unsigned long long kCellAggregate = 1ull << 62, kCellInclusive = 2ull << 62, kCellFlags = 3ull << 62;
constexpr uint32_t kCellSlots = 512;  // cells accumulated per round (a tile with more takes several rounds)

__device__ __forceinline__ void write_cell_records(const PaintScene& S, uint32_t c, uint64_t first_seg, const int32_t* acc16,
                                                   uint64_t* __restrict__ cell_key, uint4* __restrict__ cell_cover,
                                                   uint64_t* __restrict__ key2, uint32_t* __restrict__ perm) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        w[q] = ((uint32_t)acc16[4 * q] & 0xFFu) | (((uint32_t)acc16[4 * q + 1] & 0xFFu) << 8) |
               (((uint32_t)acc16[4 * q + 2] & 0xFFu) << 16) | (((uint32_t)acc16[4 * q + 3] & 0xFFu) << 24);
    const uint64_t ck = (first_seg >> kSortShift) << kSortShift;  // the cell's key = its first segment's
    cell_key[c] = ck;
    const int32_t ty = (int32_t)key_ty(ck) - 1, tx = (int32_t)key_tx(ck) - 1;
    const bool relevant = !(ty < (int32_t)S.ty_lo || ty >= (int32_t)S.ty_hi || tx >= (int32_t)S.tx_hi);
    perm[c] = c;
    cell_cover[c] = make_uint4(w[0], w[1], w[2], w[3]);
    key2[c] = relevant ? make_key2(ck) : sentinel_key(S.tiles_y);
}

__global__ void __launch_bounds__(kCellThreads)
    cells_kernel(PaintScene S, const uint64_t* __restrict__ segs, uint32_t n, unsigned long long* __restrict__ state,
                 uint32_t tiles, uint32_t* __restrict__ cell_start, uint32_t cap, uint32_t* __restrict__ n_cells_out,
                 uint64_t* __restrict__ cell_key, uint4* __restrict__ cell_cover, uint64_t* __restrict__ key2,
                 uint32_t* __restrict__ perm) {
    __shared__ uint32_t warp_cnt[kCellThreads / 32];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_prefix;
    __shared__ int32_t s_acc[kCellSlots][16];
    __shared__ uint64_t s_first[kCellSlots];  // first segment of the cell in each slot
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(&state[tiles], 1ull);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * kCellTile;
    const uint32_t base = tile_base + warp * (32u * kCellItems);
    // 1. segments -> registers, heads
    uint64_t seg[kCellItems];
    uint32_t masks[kCellItems];
    uint32_t cnt = 0;
    uint64_t carry_key = ~0ull;  // key of the element before the warp's first (none: every key differs from it)
    if (base > 0 && base < n) carry_key = segs[base - 1] >> kSortShift;
#pragma unroll
    for (int k = 0; k < kCellItems; ++k) {
        const uint32_t i = base + k * 32u + lane;
        seg[k] = i < n ? segs[i] : 0ull;
        const uint64_t key = seg[k] >> kSortShift;
        uint64_t prev = __shfl_up_sync(kFullMask, key, 1);
        if (lane == 0) prev = carry_key;
        const bool head = i < n && (i == 0 || key != prev);
        masks[k] = __ballot_sync(kFullMask, head);
        cnt += __popc(masks[k]);
        carry_key = __shfl_sync(kFullMask, key, 31);
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    uint32_t warp_off = 0, tile_sum = 0;
#pragma unroll
    for (int w = 0; w < kCellThreads / 32; ++w) {
        if ((uint32_t)w < warp) warp_off += warp_cnt[w];
        tile_sum += warp_cnt[w];
    }
    // 2. publish this tile's count at once: successors can start summing while we accumulate
    if (threadIdx.x == 0) {
        volatile unsigned long long* st = state;
        st[tile] = (tile == 0 ? kCellInclusive : kCellAggregate) | tile_sum;
    }
    // rank of every element's cell inside the tile (-1: the cell that continues from the previous tile)
    int32_t local[kCellItems];
    {
        uint32_t pos = warp_off;
#pragma unroll
        for (int k = 0; k < kCellItems; ++k) {
            local[k] = (int32_t)(pos + __popc(masks[k] & (0xFFFFFFFFu >> (31u - lane)))) - 1;  // heads at or before this lane
            pos += __popc(masks[k]);
        }
    }
    // The tile's last cell is complete iff the next tile starts with a head (or there is no next element).
    const uint32_t tile_end = min(tile_base + (uint32_t)kCellTile, n);
    bool last_complete = true;
    if (tile_end < n) last_complete = (segs[tile_end] >> kSortShift) != (segs[tile_end - 1u] >> kSortShift);
    const uint32_t n_complete = tile_sum - ((tile_sum > 0u && !last_complete) ? 1u : 0u);  // cells [0, n_complete) of the tile

    // 3. covers of the first kCellSlots complete cells — before the look-back: its latency hides behind this
    auto accumulate = [&](uint32_t r0, uint32_t nr) {
        for (uint32_t i = threadIdx.x; i < nr * 16u; i += kCellThreads) (&s_acc[0][0])[i] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kCellItems; ++k) {
            const uint32_t i = base + k * 32u + lane;
            const int32_t slot = local[k] - (int32_t)r0;
            if (i < n && slot >= 0 && slot < (int32_t)nr) {
                const uint64_t sg = seg[k];
                const uint32_t ly = (uint32_t)(sg >> 12) & 15u;
                const int32_t cv = (int32_t)(((uint32_t)sg & 0x3Fu) ^ 0x20u) - 0x20;
                atomicAdd(&s_acc[slot][ly], cv);
                if ((masks[k] >> lane) & 1u) s_first[slot] = sg;
            }
        }
        __syncthreads();
    };
    const uint32_t nr0 = min(kCellSlots, n_complete);
    accumulate(0u, nr0);

    // 4. look-back (first warp): 32 predecessors per round, until one holds an inclusive prefix
    if (warp == 0) {
        volatile unsigned long long* st = state;
        unsigned long long prefix = 0;
        if (tile != 0) {
            int32_t p = (int32_t)tile - 1;
            while (true) {
                const int32_t idx = p - (int32_t)lane;
                unsigned long long v = kCellInclusive;  // before the first tile: an inclusive prefix of 0
                if (idx >= 0) {
                    do {
                        v = st[idx];
                    } while ((v & kCellFlags) == 0);
                }
                const uint32_t incl = __ballot_sync(kFullMask, (v & kCellFlags) == kCellInclusive);
                const uint32_t upto = incl ? (uint32_t)__ffs((int)incl) : 32u;  // lanes [0, upto) count
                unsigned long long part = lane < upto ? (v & ~kCellFlags) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFullMask, part, o);
                prefix += part;
                if (incl) break;
                p -= 32;
            }
            if (lane == 0) st[tile] = kCellInclusive | (prefix + tile_sum);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile + 1 == tiles) {
                const uint32_t total = (uint32_t)(prefix + tile_sum);
                n_cells_out[0] = total;
                if (total < cap) cell_start[total] = n;  // one-past-the-end sentinel
            }
        }
    }
    __syncthreads();
    const uint32_t cells_before = (uint32_t)s_prefix;

    // 5. cell_start of the heads, records of the complete cells
#pragma unroll
    for (int k = 0; k < kCellItems; ++k) {
        if ((masks[k] >> lane) & 1u) {
            const uint32_t q = cells_before + (uint32_t)local[k];
            if (q < cap) cell_start[q] = base + k * 32u + lane;
        }
    }
    for (uint32_t r0 = 0; r0 < n_complete; r0 += kCellSlots) {
        const uint32_t nr = min(kCellSlots, n_complete - r0);
        if (r0) accumulate(r0, nr);  // a tile with more than kCellSlots cells: further rounds
        for (uint32_t sl = threadIdx.x; sl < nr; sl += kCellThreads) {
            const uint32_t c = cells_before + r0 + sl;
            if (c < cap) write_cell_records(S, c, s_first[sl], s_acc[sl], cell_key, cell_cover, key2, perm);
        }
        __syncthreads();
    }
}

// The cells that cross a tile boundary of cells_kernel: one warp per boundary b (between
// tiles b - 1 and b). The warp acts when the element after the boundary continues a cell whose
// head lies in tile b - 1 (a cell that spans whole tiles is taken at its first boundary only)
// and sums the covers over the cell's whole segment range.
__global__ void __launch_bounds__(kCellThreads)
    cells_boundary_kernel(PaintScene S, const uint64_t* __restrict__ segs, uint32_t n, const unsigned long long* __restrict__ state,
                          uint32_t tiles, const uint32_t* __restrict__ cell_start, uint32_t cap,
                          const uint32_t* __restrict__ n_cells_ptr, uint64_t* __restrict__ cell_key, uint4* __restrict__ cell_cover,
                          uint64_t* __restrict__ key2, uint32_t* __restrict__ perm) {
    const uint32_t b = blockIdx.x * (kCellThreads / 32) + (threadIdx.x >> 5) + 1u, lane = threadIdx.x & 31u;
    if (b >= tiles) return;
    const uint32_t n_cells = *n_cells_ptr;
    if (n_cells >= cap) return;  // cell_start is incomplete: the host repeats both kernels
    const uint32_t first = b * kCellTile;  // first element after the boundary
    if (first >= n || (segs[first] >> kSortShift) != (segs[first - 1u] >> kSortShift)) return;  // a head: nothing crosses
    const uint32_t incl_prev = (uint32_t)(state[b - 1u] & ~kCellFlags);                       // cells through tile b - 1
    const uint32_t incl_prev2 = b >= 2u ? (uint32_t)(state[b - 2u] & ~kCellFlags) : 0u;
    if (incl_prev == incl_prev2) return;  // tile b - 1 holds no head: an earlier boundary owns this cell
    const uint32_t c = incl_prev - 1u;
    const uint32_t s0 = cell_start[c], s1 = cell_start[c + 1u];
    int32_t acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0;
    for (uint32_t i = s0 + lane; i < s1; i += 32u) {
        const uint64_t sg = segs[i];
        const uint32_t ly = (uint32_t)(sg >> 12) & 15u;
        const int32_t cv = (int32_t)(((uint32_t)sg & 0x3Fu) ^ 0x20u) - 0x20;
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] += (ly == (uint32_t)r) ? cv : 0;
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(kFullMask, acc[r], o);
    }
    if (lane == 0) write_cell_records(S, c, segs[s0], acc, cell_key, cell_cover, key2, perm);
}

// ---------------------------------------------------------------------------
// Carries
// ---------------------------------------------------------------------------
// Counts kept on the device (DevCounts, kernels.h): a frame whose tables are built without a
// host read-back launches the kernels below over upper bounds and lets them pick up the real
// counts here. A count above its bound (the buffers are too small) turns every later kernel of
// the frame into a no-op; the host sees it at the end of the frame and builds the tables again.
__device__ __forceinline__ void resolve_cells(const DevCounts& dc, uint32_t& n_cells) {
    if (!dc.cells) return;
    const uint32_t c = *dc.cells;
    n_cells = c > dc.cell_bound ? 0u : c;
}
__device__ __forceinline__ void resolve_counts(const DevCounts& dc, uint32_t& n_cells, uint32_t& n_gaps) {
    if (!dc.cells) return;
    const uint32_t c = *dc.cells, g = *dc.gaps;
    const bool ok = c <= dc.cell_bound && g <= dc.gap_bound;
    n_cells = ok ? c : 0u;
    n_gaps = ok ? g : 0u;
}

// One thread per (tile_y, layer) group head walks its group (cells sorted by
// tile_x), producing each cell's carry-in, the running carry after it and the
// number of carry-only entries to create before the next cell.
__global__ void carry_scan_kernel(PaintScene S, const uint64_t* __restrict__ key2, const uint32_t* __restrict__ perm,
                                  const uint4* __restrict__ cell_cover, uint32_t n_cells, uint4* __restrict__ carry_in,
                                  uint4* __restrict__ carry_after, uint32_t* __restrict__ gap_count, DevCounts dc) {
    resolve_cells(dc, n_cells);
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cells) return;
    uint64_t k = key2[j];
    if (k == sentinel_key(S.tiles_y)) {  // irrelevant cell: no carries, no entries
        gap_count[j] = 0;
        return;
    }
    if (j > 0 && (key2[j - 1] >> 32) == (k >> 32)) return;  // not a group head
    const uint32_t layer = key2_layer(k);
    const uint32_t fill_rule = fill_rule_of(S, layer);
    const int32_t ty = (int32_t)key_ty(k) - 1;
    const bool row_painted = ty >= (int32_t)S.ty_lo && ty < (int32_t)S.ty_hi;
    uint4 run = make_uint4(0u, 0u, 0u, 0u);
    while (true) {
        uint32_t c = perm[j];
        carry_in[c] = run;
        run = cover_add(run, cell_cover[c]);
        carry_after[j] = run;
        int32_t t = (int32_t)key2_tx(k) - 1;
        bool has_next = j + 1 < n_cells && (key2[j + 1] >> 32) == (k >> 32);
        int32_t next_t = has_next ? (int32_t)key2_tx(key2[j + 1]) - 1 : (int32_t)S.tx_hi;
        uint32_t gaps = 0;
        if (row_painted) {
            int32_t lo = max(t + 1, (int32_t)S.tx_lo), hi = min(next_t, (int32_t)S.tx_hi);
            if (!cover_is_empty(run, fill_rule)) {
                gaps = hi > lo ? (uint32_t)(hi - lo) : 0u;
            } else if (t < (int32_t)S.tx_lo && next_t > (int32_t)S.tx_lo && S.tx_lo < S.tx_hi) {
                // covers_left_of_row: a layer with segments left of the first
                // painted tile is queued for it even when its cover sums to zero
                // (cpu/painter/mod.rs:501-522).
                gaps = 1;
            }
        }
        gap_count[j] = gaps;
        if (!has_next) break;
        ++j;
        k = key2[j];
    }
}

// Carry-only entries (payload = n_cells + gap id), generated in (layer, tile_y,
// tile_x) order: a stable sort on the tile digits alone then orders them by
// (tile_y, tile_x, layer).
__global__ void gap_fill_kernel(PaintScene S, const uint64_t* __restrict__ key2, const uint32_t* __restrict__ perm,
                                const uint64_t* __restrict__ cell_key, const uint4* __restrict__ carry_after,
                                const uint32_t* __restrict__ gap_offset /* exclusive scan of the gap counts */,
                                uint32_t n_cells, uint64_t* __restrict__ gkey, uint32_t* __restrict__ gid,
                                uint4* __restrict__ gap_carry, const uint32_t* __restrict__ n_gaps_ptr, uint32_t cap,
                                DevCounts dc) {
    // One thread per carry-only entry q: its source cell is the last one (in carry
    // order) whose exclusive offset is <= q — cells without entries share their
    // offset with the next cell and sort before it.
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_gaps = *n_gaps_ptr;
    if (n_gaps > cap || q >= n_gaps) return;  // cap: launched before the host knows the count
    resolve_cells(dc, n_cells);
    if (n_cells == 0u) return;
    uint32_t lo = 0, hi = n_cells;  // first j with gap_offset[j] > q
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (gap_offset[mid] <= q) lo = mid + 1u;
        else hi = mid;
    }
    const uint32_t j = lo - 1u;
    const uint32_t r = q - gap_offset[j];
    const uint64_t ck = cell_key[perm[j]];
    const int32_t tx = (int32_t)key_tx(ck) - 1;
    const int32_t first = max(tx + 1, (int32_t)S.tx_lo);
    gkey[q] = (ck & ~(0xFFFull << 41)) | ((uint64_t)(uint32_t)(first + (int32_t)r + 1) << 41);
    gid[q] = n_cells + q;
    gap_carry[q] = carry_after[j];
}

// Entries = cells (sorted by construction) merged with the sorted carry-only
// entries; both lists are strictly increasing and share no key, so an element's
// position is its own index plus its rank in the other list.
__device__ __forceinline__ uint32_t lower_bound_key(const uint64_t* __restrict__ a, uint32_t n, uint64_t k) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// Besides the sorted key, every entry gets a self-contained 64-byte record (its
// segment range, carry, style bits and colour) and its initial optimizer flags,
// so that the painter reaches everything it needs with one round of loads per
// 32 entries instead of chasing key -> id -> cell -> style pointers per tile.
__global__ void merge_entries_kernel(PaintScene S, const uint64_t* __restrict__ cell_key, uint32_t n_cells,
                                     const uint64_t* __restrict__ gkey, const uint32_t* __restrict__ gid, uint32_t n_gaps,
                                     const uint32_t* __restrict__ cell_start, const uint4* __restrict__ carry_in,
                                     const uint4* __restrict__ gap_carry, uint64_t* __restrict__ ekey,
                                     EntryRec* __restrict__ recs, uint8_t* __restrict__ eflags, DevCounts dc) {
    resolve_counts(dc, n_cells, n_gaps);
    // Merge-path style partition: the keys of a CTA's 256 consecutive elements of
    // one list bracket a short range of the other list, found once per CTA (two
    // full bisections by threads 0 and 1); every thread then bisects that range only.
    __shared__ uint32_t s_bound[2];
    const uint32_t i0 = blockIdx.x * blockDim.x, i = i0 + threadIdx.x;
    const uint32_t n_all = n_cells + n_gaps;
    if (i0 >= n_all) return;  // (whole CTA; only with device-side counts)
    const bool cell_block = i0 + blockDim.x <= n_cells, gap_block = i0 >= n_cells;  // else: the one mixed CTA
    if (threadIdx.x < 2u) {
        const uint32_t last = min(i0 + blockDim.x, n_all) - 1u;
        uint32_t b = threadIdx.x == 0u ? 0u : (cell_block ? n_gaps : n_cells);
        if (cell_block) b = lower_bound_key(gkey, n_gaps, cell_key[threadIdx.x == 0u ? i0 : last]);
        else if (gap_block) b = lower_bound_key(cell_key, n_cells, gkey[(threadIdx.x == 0u ? i0 : last) - n_cells]);
        s_bound[threadIdx.x] = b;
    }
    __syncthreads();
    if (i >= n_all) return;
    uint64_t k;
    uint32_t pos;
    EntryRec r;
    if (i < n_cells) {
        k = cell_key[i];
        const uint32_t lo = cell_block ? s_bound[0] : 0u, hi = cell_block ? s_bound[1] : n_gaps;
        pos = i + lo + lower_bound_key(gkey + lo, hi - lo, k);
        r.seg0 = cell_start[i];
        r.seg1 = cell_start[i + 1];
        r.carry = carry_in[i];
    } else {
        uint32_t g = i - n_cells;
        k = gkey[g];
        const uint32_t lo = gap_block ? s_bound[0] : 0u, hi = gap_block ? s_bound[1] : n_cells;
        pos = g + lo + lower_bound_key(cell_key + lo, hi - lo, k);
        r.seg0 = r.seg1 = 0;
        r.carry = gap_carry[gid[g] - n_cells];
    }
    r.layer = key_layer(k);
    r.slot = r.layer < S.n_orders ? S.order_to_style[r.layer] : -1;
    uint32_t flags = 0;
    if (r.slot >= 0) {
        const StyleRec& st = S.styles[r.slot];
        const bool unchanged = S.unchanged && S.unchanged[r.layer];
        r.meta = pack_style_meta(st, unchanged);
        r.clip_layers = st.clip_layers;
        r.color[0] = st.color[0];
        r.color[1] = st.color[1];
        r.color[2] = st.color[2];
        r.color[3] = st.color[3];
        if (r.seg1 > r.seg0) flags |= kFlagHasSegs;
        else if (cover_is_full(r.carry, st.fill_rule & 1u)) flags |= kFlagFull;  // layer_is_full, mod.rs:171-182
        if (unchanged) flags |= kFlagUnchanged;
        if (st.func == 1u || st.is_clipped) flags |= kFlagClipish;
        if (st.func == 0u && st.is_clipped) flags |= kFlagClippedDraw;
        if (st.func == 0u && st.fill_type == 0u && st.blend_mode == 0u && st.color[3] == 1.0f) flags |= kFlagOpaque;
    } else {
        r.meta = 0;
        r.clip_layers = 0;
        r.color[0] = r.color[1] = r.color[2] = r.color[3] = 0.0f;
        flags = kFlagMaskedOut;
    }
    r.flags0 = flags;
    r.pad = 0;
    ekey[pos] = k;
    recs[pos] = r;
    eflags[pos] = (uint8_t)flags;
}

// Per painted tile: [begin, end) of its entries in the sorted entry list, written by the
// thread of the tile's last entry (which bisects for the first one); the same thread files
// a tile with many entries in its class list (see paint_common.cuh: kHeavyMin).
__global__ void tile_index_kernel(PaintScene S, const uint64_t* __restrict__ ekey, uint32_t n_entries,
                                  uint2* __restrict__ tile_range, uint32_t* __restrict__ heavy /* [classes][tiles] or null */,
                                  uint32_t* __restrict__ heavy_count, DevCounts dc) {
    if (dc.cells) {
        uint32_t c = 0, g = 0;
        resolve_counts(dc, c, g);
        n_entries = c + g;
    }
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_entries) return;
    const uint64_t k = ekey[p];
    // Entries outside the render target (biased tile coordinate 0 = tile -1, or
    // beyond the last tile) belong to no painted tile.
    if (key_ty(k) == 0u || key_ty(k) > S.tiles_y || key_tx(k) == 0u || key_tx(k) > S.tiles_x) return;
    const uint64_t tile_bits = k >> 41;
    if (p + 1 != n_entries && (ekey[p + 1] >> 41) == tile_bits) return;  // not the tile's last entry
    uint32_t lo = 0, hi = p;  // first entry of the tile: the first q with ekey[q] >> 41 >= tile_bits
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((ekey[mid] >> 41) < tile_bits) lo = mid + 1u;
        else hi = mid;
    }
    const uint32_t tid = (key_ty(k) - 1u) * S.tiles_x + (key_tx(k) - 1u);
    tile_range[tid] = make_uint2(lo, p + 1u);
    const uint32_t count = p + 1u - lo;
    if (heavy && count >= kHeavyMin) {
        const int c = heavy_class(count);
        heavy[(size_t)c * S.tiles_x * S.tiles_y + atomicAdd(&heavy_count[c], 1u)] = tid;
    }
}

// Cost of every tile row of the frame just rendered: 32 x its (tile, layer) entries + its
// pixel segments (paint time follows the entries, sort / table time the segments). The
// multi-GPU band split of the next frame is balanced on these. One warp per row.
__global__ void row_cost_kernel(const uint2* __restrict__ tile_range, uint32_t tiles_x, uint32_t tiles_y,
                                const uint64_t* __restrict__ segs, uint32_t n, unsigned long long* __restrict__ out,
                                unsigned long long* __restrict__ seg_out) {
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (row >= tiles_y) return;
    uint32_t entries = 0;
    for (uint32_t tx = lane; tx < tiles_x; tx += 32u) {
        const uint2 r = tile_range[(size_t)row * tiles_x + tx];
        entries += r.y - r.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) entries += __shfl_xor_sync(kFullMask, entries, o);
    if (lane == 0) {
        auto lower = [&](uint64_t key) {  // first segment with key >= `key`
            uint32_t lo = 0, hi = n;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (segs[mid] < key) lo = mid + 1u;
                else hi = mid;
            }
            return lo;
        };
        const uint32_t s0 = lower((uint64_t)(row + 1u) << 53), s1 = lower((uint64_t)(row + 2u) << 53);
        out[row] = 32ull * entries + (unsigned long long)(s1 - s0);
        // Pixel segments and entries of the row. The first row also takes the segments above the
        // frame (tile_y clamped to -1), the last row those below it (they sort behind every
        // painted tile), so that the rows of a frame add up to its segment count.
        if (seg_out) {
            seg_out[row] = (row + 1u == tiles_y ? n : s1) - (row == 0u ? 0u : s0);
            seg_out[tiles_y + row] = entries;
        }
    }
}

void launch_row_costs(const uint2* tile_range, uint32_t tiles_x, uint32_t tiles_y, const uint64_t* segs, uint32_t n,
                      unsigned long long* out, cudaStream_t st, unsigned long long* seg_out) {
    if (tiles_y) row_cost_kernel<<<(tiles_y + 7) / 8, 256, 0, st>>>(tile_range, tiles_x, tiles_y, segs, n, out, seg_out);
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
uint32_t cell_num_blocks(uint32_t n) { return (n + kCellTile - 1) / kCellTile; }

size_t cells_state_words(uint32_t n) { return (size_t)cell_num_blocks(n) + 2; }

void launch_cells(const PaintScene& S, const uint64_t* segs, uint32_t n, unsigned long long* state, uint32_t* cell_start,
                  uint32_t cap, uint32_t* n_cells_out, uint64_t* cell_key, uint4* cell_cover, uint64_t* key2, uint32_t* perm,
                  cudaStream_t st) {
    const uint32_t tiles = cell_num_blocks(n);
    cudaMemsetAsync(state, 0, (tiles + 1) * sizeof(unsigned long long), st);
    cells_kernel<<<tiles, kCellThreads, 0, st>>>(S, segs, n, state, tiles, cell_start, cap, n_cells_out, cell_key, cell_cover, key2,
                                                 perm);
    if (tiles > 1)
        cells_boundary_kernel<<<(tiles - 1 + kCellThreads / 32 - 1) / (kCellThreads / 32), kCellThreads, 0, st>>>(
            S, segs, n, state, tiles, cell_start, cap, n_cells_out, cell_key, cell_cover, key2, perm);
}

// Largest values the three key fields can take in the pair sorts (sentinel
// included); the sort plan only spends passes on bits below these bounds.
// The cells arrive sorted by (tile_y, tile_x, layer); a stable sort on the layer
// bits alone leaves them ordered by (layer, tile_y, tile_x), which is all the
// carry scan needs: every (tile_y, layer) group contiguous, tile_x ascending.
SortPlan carry_sort_plan(const PaintScene& S) {  // fields, least significant first: tile_x, layer, tile_y
    const uint64_t bound[3] = {0u, S.n_orders ? S.n_orders - 1u : 0u, 0u};
    return make_sort_plan(carry_key_layout(), bound);
}
// Carry-only entries: tile digits only (see gap_fill_kernel).
SortPlan gap_sort_plan(const PaintScene& S) {  // layer, tile_x, tile_y
    const uint64_t bound[3] = {0u, S.tx_hi, S.ty_hi};
    return make_sort_plan(segment_key_layout(), bound);
}

void launch_carry_scan(const PaintScene& S, const uint64_t* key2, const uint32_t* perm, const uint4* cell_cover,
                       uint32_t n_cells, uint4* carry_in, uint4* carry_after, uint32_t* gap_count, cudaStream_t st,
                       const DevCounts& dc) {
    carry_scan_kernel<<<(n_cells + 127) / 128, 128, 0, st>>>(S, key2, perm, cell_cover, n_cells, carry_in, carry_after,
                                                              gap_count, dc);
}

void launch_gap_fill(const PaintScene& S, const uint64_t* key2, const uint32_t* perm, const uint64_t* cell_key,
                     const uint4* carry_after, const uint32_t* gap_offset, uint32_t n_cells,
                     uint64_t* gkey, uint32_t* gid, uint4* gap_carry, const uint32_t* n_gaps_ptr, uint32_t cap,
                     uint32_t grid_gaps, cudaStream_t st, const DevCounts& dc) {
    if (!grid_gaps || !n_cells) return;
    gap_fill_kernel<<<(grid_gaps + 127) / 128, 128, 0, st>>>(S, key2, perm, cell_key, carry_after, gap_offset, n_cells,
                                                              gkey, gid, gap_carry, n_gaps_ptr, cap, dc);
}

void launch_merge_entries(const PaintScene& S, const uint64_t* cell_key, uint32_t n_cells, const uint64_t* gkey,
                          const uint32_t* gid, uint32_t n_gaps, const uint32_t* cell_start, const uint4* carry_in,
                          const uint4* gap_carry, uint64_t* ekey, EntryRec* recs, uint8_t* eflags, cudaStream_t st,
                          const DevCounts& dc) {
    uint32_t n = n_cells + n_gaps;
    if (n)
        merge_entries_kernel<<<(n + 255) / 256, 256, 0, st>>>(S, cell_key, n_cells, gkey, gid, n_gaps, cell_start, carry_in,
                                                               gap_carry, ekey, recs, eflags, dc);
}

void launch_tile_index(const PaintScene& S, const uint64_t* ekey, uint32_t n_entries, uint2* tile_range, uint32_t* heavy,
                       uint32_t* heavy_count, cudaStream_t st, const DevCounts& dc) {
    cudaMemsetAsync(tile_range, 0, (size_t)S.tiles_x * S.tiles_y * sizeof(uint2), st);
    cudaMemsetAsync(heavy_count, 0, kHeavyClasses * sizeof(uint32_t), st);
    if (n_entries)
        tile_index_kernel<<<(n_entries + 255) / 256, 256, 0, st>>>(S, ekey, n_entries, tile_range, heavy, heavy_count, dc);
}


}  // namespace forma
