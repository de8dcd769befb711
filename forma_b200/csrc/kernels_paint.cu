// Stage 4: per-tile coverage accumulation and compositing to RGBA8 — replaces
// forma/src/cpu/painter/{mod.rs, layer_workbench/, styling.rs}.
//
// The reference walks each tile row left to right, carrying every layer's
// winding "cover" (16 x i8, one per pixel row) from tile to tile in a queue
// (cpu/painter/mod.rs:486-568, layer_workbench/mod.rs:196-342). To paint tiles
// independently the carries are materialised first:
//
//   cells     runs of sorted segments with equal (tile_y, tile_x, layer)
//   covers    per cell: sum of segment covers by local_y (wrapping i8)
//   re-sort   cell ids by (tile_y, layer, tile_x)            [pair radix sort]
//   carries   per (tile_y, layer) group a running sum -> carry-in of every cell
//             and "carry-only" entries for the tiles a layer spans without
//             segments (layer_workbench/mod.rs:213-234,328-336)
//   entries   cells ∪ carry-only entries, sorted by (tile_y, tile_x, layer)
//   paint     one warp per tile; lane l owns column l/2, rows 8*(l%2)..+8 —
//             exactly one f32x8 of the reference (cpu/painter/mod.rs:234-244)
//
// Integer semantics: areas wrap at i16 and covers at i8 in the reference;
// sums are formed in i32 / packed bytes and truncated where the reference
// widens them (truncation commutes with wrapping addition).
#include "cuda_common.cuh"
#include "kernels.h"
#include "paint_math.cuh"

namespace forma {

// --- key helpers (keys keep the segment layout; low 20 bits are zero) ---------
__device__ __forceinline__ uint32_t key_ty(uint64_t k) { return (uint32_t)(k >> 53) & 0x7FFu; }   // biased (+1)
__device__ __forceinline__ uint32_t key_tx(uint64_t k) { return (uint32_t)(k >> 41) & 0xFFFu; }   // biased (+1)
__device__ __forceinline__ uint32_t key_layer(uint64_t k) { return (uint32_t)(k >> 20) & 0x1FFFFFu; }
// (tile_y, layer, tile_x) ordering key in the same [20, 64) bit window.
__device__ __forceinline__ uint64_t make_key2(uint64_t k) {
    return ((uint64_t)key_ty(k) << 53) | ((uint64_t)key_layer(k) << 32) | ((uint64_t)key_tx(k) << 20);
}
__device__ __forceinline__ uint32_t key2_tx(uint64_t k2) { return (uint32_t)(k2 >> 20) & 0xFFFu; }
__device__ __forceinline__ uint32_t key2_layer(uint64_t k2) { return (uint32_t)(k2 >> 32) & 0x1FFFFFu; }

// --- packed 16 x i8 covers -----------------------------------------------------
__device__ __forceinline__ uint4 cover_add(uint4 a, uint4 b) {
    return make_uint4(__vadd4(a.x, b.x), __vadd4(a.y, b.y), __vadd4(a.z, b.z), __vadd4(a.w, b.w));
}
// Cover::is_empty / is_full, cpu/painter/mod.rs:187-214.
__device__ __forceinline__ bool cover_is_empty(uint4 c, uint32_t fill_rule) {
    uint32_t any = c.x | c.y | c.z | c.w;
    return fill_rule == 0u ? any == 0u : (any & 0x1F1F1F1Fu) == 0u;
}
__device__ __forceinline__ bool cover_is_full(uint4 c, uint32_t fill_rule) {
    if (fill_rule == 0u) {
        const uint32_t k = 0x10101010u;
        return __vabs4(c.x) == k && __vabs4(c.y) == k && __vabs4(c.z) == k && __vabs4(c.w) == k;
    }
    const uint32_t m = 0x1F1F1F1Fu, k = 0x10101010u;
    return (c.x & m) == k && (c.y & m) == k && (c.z & m) == k && (c.w & m) == k;
}

__device__ __forceinline__ uint32_t fill_rule_of(const PaintScene& S, uint32_t layer) {
    int32_t slot = layer < S.n_orders ? S.order_to_style[layer] : -1;
    return slot >= 0 ? S.styles[slot].fill_rule : 0u;
}

// ---------------------------------------------------------------------------
// Cells
// ---------------------------------------------------------------------------
constexpr int kCellThreads = 256;

__device__ __forceinline__ bool is_cell_head(const uint64_t* __restrict__ segs, uint32_t i) {
    return i == 0 || (segs[i] >> kSortShift) != (segs[i - 1] >> kSortShift);
}

__global__ void __launch_bounds__(kCellThreads)
    cell_count_kernel(const uint64_t* __restrict__ segs, uint32_t n, uint32_t* __restrict__ block_counts) {
    __shared__ uint32_t warp_cnt[kCellThreads / 32];
    uint32_t i = blockIdx.x * kCellThreads + threadIdx.x;
    bool head = i < n && is_cell_head(segs, i);
    uint32_t b = __ballot_sync(kFullMask, head);
    if (lane_id() == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kCellThreads / 32; ++w) s += warp_cnt[w];
        block_counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kCellThreads)
    cell_write_kernel(const uint64_t* __restrict__ segs, uint32_t n, const uint32_t* __restrict__ block_offsets,
                      uint32_t* __restrict__ cell_start, uint64_t* __restrict__ cell_key, uint32_t n_cells) {
    __shared__ uint32_t warp_cnt[kCellThreads / 32];
    uint32_t i = blockIdx.x * kCellThreads + threadIdx.x;
    bool head = i < n && is_cell_head(segs, i);
    uint32_t b = __ballot_sync(kFullMask, head);
    if (lane_id() == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (head) {
        uint32_t pos = block_offsets[blockIdx.x] + __popc(b & ((1u << lane_id()) - 1u));
        for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) pos += warp_cnt[w];
        cell_start[pos] = i;
        cell_key[pos] = (segs[i] >> kSortShift) << kSortShift;
    }
    if (i == 0) cell_start[n_cells] = n;
}

// Per cell: sum of covers by local_y (acc_segment's cover part + cover_carry,
// cpu/painter/mod.rs:257-271, layer_workbench/mod.rs:218-224), and the
// (tile_y, layer, tile_x) key for the carry pass.
__global__ void cell_cover_kernel(const uint64_t* __restrict__ segs, const uint32_t* __restrict__ cell_start,
                                  const uint64_t* __restrict__ cell_key, uint32_t n_cells, uint4* __restrict__ cell_cover,
                                  uint64_t* __restrict__ key2, uint32_t* __restrict__ perm) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    uint32_t s0 = cell_start[c], s1 = cell_start[c + 1];
    uint32_t acc[4] = {0u, 0u, 0u, 0u};
    for (uint32_t i = s0; i < s1; ++i) {
        uint64_t s = segs[i];
        uint32_t ly = (uint32_t)(s >> 12) & 15u;
        uint32_t cv = (uint32_t)s & 0x3Fu;
        cv = (cv ^ 0x20u) - 0x20u;  // sign-extend 6 bits
        uint32_t v = (cv & 0xFFu) << (8u * (ly & 3u));
        uint32_t w = ly >> 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = __vadd4(acc[k], w == (uint32_t)k ? v : 0u);
    }
    cell_cover[c] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    key2[c] = make_key2(cell_key[c]);
    perm[c] = c;
}

// ---------------------------------------------------------------------------
// Carries
// ---------------------------------------------------------------------------
// One thread per (tile_y, layer) group head walks its group (cells sorted by
// tile_x), producing each cell's carry-in, the running carry after it and the
// number of carry-only entries to create before the next cell.
__global__ void carry_scan_kernel(PaintScene S, const uint64_t* __restrict__ key2, const uint32_t* __restrict__ perm,
                                  const uint4* __restrict__ cell_cover, uint32_t n_cells, uint4* __restrict__ carry_in,
                                  uint4* __restrict__ carry_after, uint32_t* __restrict__ gap_count) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cells) return;
    uint64_t k = key2[j];
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

// Entries: one per cell (payload = cell id) + carry-only entries (payload =
// n_cells + gap id). Cells that are not painted get the key ~0 (sorted last).
__global__ void entry_fill_kernel(PaintScene S, const uint64_t* __restrict__ key2, const uint32_t* __restrict__ perm,
                                  const uint64_t* __restrict__ cell_key, const uint4* __restrict__ carry_after,
                                  const uint32_t* __restrict__ gap_count, const uint32_t* __restrict__ gap_offset,
                                  uint32_t n_cells, uint64_t* __restrict__ ekey, uint32_t* __restrict__ eid,
                                  uint4* __restrict__ gap_carry) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_cells) return;
    uint32_t c = perm[j];
    uint64_t ck = cell_key[c];
    int32_t ty = (int32_t)key_ty(ck) - 1, tx = (int32_t)key_tx(ck) - 1;
    bool painted = ty >= (int32_t)S.ty_lo && ty < (int32_t)S.ty_hi && tx >= (int32_t)S.tx_lo && tx < (int32_t)S.tx_hi;
    ekey[c] = painted ? ck : ~0ull;
    eid[c] = c;
    uint32_t g = gap_count[j];
    if (g) {
        uint32_t off = gap_offset[j];
        uint4 carry = carry_after[j];
        int32_t first = max(tx + 1, (int32_t)S.tx_lo);
        for (uint32_t r = 0; r < g; ++r) {
            uint64_t key = (ck & ~(0xFFFull << 41)) | ((uint64_t)(uint32_t)(first + (int32_t)r + 1) << 41);
            ekey[n_cells + off + r] = key;
            eid[n_cells + off + r] = n_cells + off + r;
            gap_carry[off + r] = carry;
        }
    }
}

// Per painted tile: [begin, end) of its entries in the sorted entry list.
__global__ void tile_range_kernel(PaintScene S, const uint64_t* __restrict__ ekey, uint32_t n_entries,
                                  uint32_t* __restrict__ tile_begin, uint32_t* __restrict__ tile_end) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_entries) return;
    uint64_t k = ekey[p];
    if (k == ~0ull) return;
    uint64_t tile_bits = k >> 41;
    uint32_t tid = (key_ty(k) - 1u) * S.tiles_x + (key_tx(k) - 1u);
    if (p == 0 || (ekey[p - 1] >> 41) != tile_bits) tile_begin[tid] = p;
    if (p + 1 == n_entries || (ekey[p + 1] >> 41) != tile_bits) tile_end[tid] = p + 1;
}

// ---------------------------------------------------------------------------
// Painter
// ---------------------------------------------------------------------------
constexpr int kPaintWarps = 4;

struct EntryRef {
    uint32_t layer;
    uint32_t seg0, seg1;
    uint4 carry;
};

struct PaintInputs {
    const uint64_t* segs;
    const uint64_t* ekey;        // sorted entries
    const uint32_t* eid;
    const uint32_t* cell_start;
    const uint4* carry_in;       // by cell
    const uint4* gap_carry;      // by gap id
    uint32_t n_cells;
    const uint32_t* tile_begin;
    const uint32_t* tile_end;
    uint8_t* eflags;             // per sorted entry scratch (optimizer passes)
    uint8_t* framebuffer;
};

__device__ __forceinline__ EntryRef load_entry(const PaintInputs& in, uint32_t p) {
    EntryRef e;
    e.layer = key_layer(in.ekey[p]);
    uint32_t id = in.eid[p];
    if (id < in.n_cells) {
        e.seg0 = in.cell_start[id];
        e.seg1 = in.cell_start[id + 1];
        e.carry = in.carry_in[id];
    } else {
        e.seg0 = e.seg1 = 0;
        e.carry = in.gap_carry[id - in.n_cells];
    }
    return e;
}

constexpr uint8_t kFlagHasSegs = 1, kFlagFull = 2, kFlagMaskedOut = 4, kFlagSkipClip = 8;

__device__ __forceinline__ const StyleRec& style_of(const PaintScene& S, uint32_t layer) {
    return S.styles[S.order_to_style[layer]];
}

__device__ __forceinline__ void store_tile_solid(const PaintScene& S, uint8_t* fb, uint32_t tx, uint32_t ty, uint32_t lane,
                                                 uint32_t rgba) {
    uint32_t px = tx * 16u + (lane >> 1);
    uint32_t py0 = ty * 16u + (lane & 1u) * 8u;
    if (px >= S.width) return;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        uint32_t py = py0 + l;
        if (py < S.height) *reinterpret_cast<uint32_t*>(fb + (size_t)py * S.stride + (size_t)px * 4u) = rgba;
    }
}

__global__ void __launch_bounds__(kPaintWarps * 32) paint_kernel(PaintScene S, PaintInputs in, uint32_t n_tiles) {
    __shared__ int32_t s_area[kPaintWarps][256];
    __shared__ int32_t s_cover[kPaintWarps][256];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t tile_lin = blockIdx.x * kPaintWarps + warp;
    if (tile_lin >= n_tiles) return;
    const uint32_t ntx = S.tx_hi - S.tx_lo;
    const uint32_t ty = S.ty_lo + tile_lin / ntx, tx = S.tx_lo + tile_lin % ntx;
    const uint32_t tid = ty * S.tiles_x + tx;
    const uint32_t b = in.tile_begin[tid], e = in.tile_end[tid];
    int32_t* area = s_area[warp];
    int32_t* cover = s_cover[warp];
    const Rgba clear{S.clear[0], S.clear[1], S.clear[2], S.clear[3]};

    // ---- optimizer passes (layer_workbench/passes/*.rs) ----------------------
    // Pass A: per-entry facts.
    bool any_clip = false;
    for (uint32_t p0 = b; p0 < e; p0 += 32u) {
        uint32_t p = p0 + lane;
        bool clipish = false;
        if (p < e) {
            EntryRef er = load_entry(in, p);
            const StyleRec& st = style_of(S, er.layer);
            uint8_t f = 0;
            if (er.seg1 > er.seg0) f |= kFlagHasSegs;
            else if (cover_is_full(er.carry, st.fill_rule)) f |= kFlagFull;  // layer_is_full, mod.rs:171-182
            in.eflags[p] = f;
            clipish = st.func == 1u || (st.func == 0u && st.is_clipped);
        }
        any_clip |= __any_sync(kFullMask, clipish);
    }
    __syncwarp();

    // Pass B: skip_trivial_clips (sequential; only tiles that contain clips).
    if (any_clip) {
        if (lane == 0) {
            bool has_clip = false, clip_full = false, clip_used = false;
            uint32_t clip_last = 0, clip_i = 0;
            for (uint32_t p = b; p < e; ++p) {
                uint32_t id = key_layer(in.ekey[p]);
                const StyleRec& st = style_of(S, id);
                uint8_t f = in.eflags[p];
                if (st.func == 1u) {
                    clip_full = (f & kFlagFull) != 0;
                    clip_last = id + st.clip_layers;
                    clip_i = p;
                    clip_used = false;
                    has_clip = true;
                    if (clip_full) f |= kFlagMaskedOut;
                }
                if (st.func == 0u && st.is_clipped) {
                    if (has_clip && id <= clip_last) {
                        if (clip_full) f |= kFlagSkipClip;
                        else clip_used = true;
                    } else {
                        f |= kFlagMaskedOut;
                    }
                }
                in.eflags[p] = f;
                if (has_clip && id > clip_last) {
                    has_clip = false;
                    if (!clip_used) in.eflags[clip_i] |= kFlagMaskedOut;
                }
            }
            if (has_clip && !clip_used) in.eflags[clip_i] |= kFlagMaskedOut;
        }
        __syncwarp();
    }

    // Pass C: skip_fully_covered_layers — top-most full, unclipped, opaque
    // `Over` solid layer culls everything below it.
    uint32_t first_paint = b;     // MaskedVec::skip_until
    bool incomplete = false;      // an "interesting" incomplete cover above the opaque layer
    bool have_opaque = false;
    for (uint32_t hi = e; hi > b && !have_opaque;) {
        uint32_t lo = hi - b >= 32u ? hi - 32u : b;
        uint32_t p = lo + lane;
        bool inc = false, cand = false;
        if (p < hi) {
            uint8_t f = in.eflags[p];
            if (!(f & kFlagMaskedOut)) {
                const StyleRec& st = style_of(S, key_layer(in.ekey[p]));
                bool clipped = st.func == 0u && st.is_clipped && !(f & kFlagSkipClip);
                if (clipped || !(f & kFlagFull)) inc = true;
                else if (st.func == 0u && st.fill_type == 0u && st.blend_mode == 0u && st.color[3] == 1.0f) cand = true;
            }
        }
        uint32_t cand_mask = __ballot_sync(kFullMask, cand);
        uint32_t inc_mask = __ballot_sync(kFullMask, inc);
        if (cand_mask) {
            uint32_t top = 31u - (uint32_t)__clz((int)cand_mask);
            have_opaque = true;
            first_paint = lo + top;
            if (top < 31u && (inc_mask >> (top + 1u)) != 0u) incomplete = true;
        } else if (inc_mask) {
            incomplete = true;
        }
        hi = lo;
    }

    if (!incomplete) {
        // Every visible layer is full: fold with the scalar blend and emit a
        // solid tile (skip_fully_covered_layers.rs:81-118, mod.rs:686-704).
        Rgba dst = clear;
        uint32_t p = first_paint;
        if (have_opaque) {
            const StyleRec& st = style_of(S, key_layer(in.ekey[p]));
            dst = Rgba{st.color[0], st.color[1], st.color[2], st.color[3]};
            ++p;
        }
        bool solid = true;
        for (; p < e; ++p) {
            if (in.eflags[p] & kFlagMaskedOut) continue;
            const StyleRec& st = style_of(S, key_layer(in.ekey[p]));
            if (st.func == 0u && st.fill_type == 0u) {
                dst = sblend::blend(st.blend_mode, dst, Rgba{st.color[0], st.color[1], st.color[2], st.color[3]});
            } else {
                solid = false;
                break;
            }
        }
        if (solid) {
            store_tile_solid(S, in.framebuffer, tx, ty, lane, solid_to_srgb_bytes(dst, S.channels));
            return;
        }
    }

    // ---- paint (layer_workbench/mod.rs:301-337, cpu/painter/mod.rs:290-347) ---
    const uint32_t x = lane >> 1, half = lane & 1u;
    float dr[8], dg[8], db[8], da[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        dr[l] = clear.r; dg[l] = clear.g; db[l] = clear.b; da[l] = clear.a;
    }
    bool clip_active = false;
    uint32_t clip_last = 0;
    float clip_mask[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) clip_mask[l] = 0.0f;
    // The cells of this warp start (and are kept) zeroed.
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        area[lane * 8 + l] = 0;
        cover[lane * 8 + l] = 0;
    }
    __syncwarp();
    const float fx = (float)(x + tx * 16u);
    const float fy = (float)(half * 8u + ty * 16u);

    for (uint32_t p = first_paint; p < e; ++p) {
        const uint8_t flags = in.eflags[p];
        if (flags & kFlagMaskedOut) continue;
        const EntryRef er = load_entry(in, p);
        const StyleRec& st = style_of(S, er.layer);
        const uint32_t fill_rule = st.fill_rule;

        // acc_segment: scatter-add the cell's segments (cpu/painter/mod.rs:257-271).
        int32_t a8[8];
        uint32_t run_lo, run_hi;  // running covers of rows 0-3 / 4-7 of this lane's half, packed i8
        if (er.seg1 > er.seg0) {
            for (uint32_t i = er.seg0 + lane; i < er.seg1; i += 32u) {
                uint64_t s = in.segs[i];
                uint32_t cell = ((uint32_t)(s >> 16) & 15u) * 16u + ((uint32_t)(s >> 12) & 15u);
                int32_t cv = (int32_t)(((uint32_t)s & 0x3Fu) ^ 0x20u) - 0x20;
                int32_t dam = (int32_t)((uint32_t)(s >> 6) & 0x3Fu);
                atomicAdd(&area[cell], dam * cv);
                atomicAdd(&cover[cell], cv);
            }
            __syncwarp();
            uint32_t c_lo = 0, c_hi = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                int idx = x * 16 + half * 8 + l;
                a8[l] = (int32_t)(int16_t)area[idx];
                uint32_t cb = (uint32_t)cover[idx] & 0xFFu;
                if (l < 4) c_lo |= cb << (8 * l);
                else c_hi |= cb << (8 * (l - 4));
                area[idx] = 0;
                cover[idx] = 0;
            }
            // Exclusive prefix over columns x' < x (same half): lanes l-2, l-4, ...
            uint32_t i_lo = c_lo, i_hi = c_hi;
#pragma unroll
            for (int o = 2; o < 32; o <<= 1) {
                uint32_t n_lo = __shfl_up_sync(kFullMask, i_lo, o);
                uint32_t n_hi = __shfl_up_sync(kFullMask, i_hi, o);
                if (lane >= (uint32_t)o) {
                    i_lo = __vadd4(i_lo, n_lo);
                    i_hi = __vadd4(i_hi, n_hi);
                }
            }
            uint32_t e_lo = __shfl_up_sync(kFullMask, i_lo, 2);
            uint32_t e_hi = __shfl_up_sync(kFullMask, i_hi, 2);
            if (lane < 2u) e_lo = e_hi = 0u;
            run_lo = __vadd4(e_lo, half ? er.carry.z : er.carry.x);
            run_hi = __vadd4(e_hi, half ? er.carry.w : er.carry.y);
            __syncwarp();
        } else {
#pragma unroll
            for (int l = 0; l < 8; ++l) a8[l] = 0;
            run_lo = half ? er.carry.z : er.carry.x;
            run_hi = half ? er.carry.w : er.carry.y;
        }

        if (clip_active && clip_last < er.layer) clip_active = false;  // mod.rs:302-306

        float cov[8];
        bool all_zero = true;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            uint32_t byte = ((l < 4 ? run_lo : run_hi) >> (8 * (l & 3))) & 0xFFu;
            int32_t doubled = 32 * (int32_t)(int8_t)byte + a8[l];  // compute_doubled_areas, mod.rs:388-404
            cov[l] = coverage_of(doubled, fill_rule);
            all_zero = all_zero && (cov[l] == 0.0f);
        }

        if (st.func == 1u) {  // clip_at, mod.rs:449-464
            if (!clip_active) {
                clip_active = true;
                clip_last = er.layer + st.clip_layers;
            }
#pragma unroll
            for (int l = 0; l < 8; ++l) clip_mask[l] = cov[l];
            continue;
        }
        const bool apply_clip = st.is_clipped && !(flags & kFlagSkipClip);
        if (all_zero) continue;                       // mod.rs:317-319 (whole f32x8 is zero)
        if (apply_clip && !clip_active) continue;     // mod.rs:321-323

        const uint32_t mode = st.blend_mode;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            float fill[4];
            if (st.fill_type == 0u) {
                fill[0] = st.color[0]; fill[1] = st.color[1]; fill[2] = st.color[2]; fill[3] = st.color[3];
            } else if (st.fill_type == 1u) {
                gradient_at(st, S.stops, fx, fy, l, fill);
            } else {
                texture_at(st, S.texels, fx, fy, l, fill);
            }
            // blend_at, mod.rs:406-447
            float sa = fill[3] * cov[l];
            if (apply_clip) sa *= clip_mask[l];
            float bl[3];
            vblend::blend(mode, dr[l], dg[l], db[l], fill[0], fill[1], fill[2], bl);
            float inv_dst_a = 1.0f - da[l];
            float inv_dst_a_src_a = inv_dst_a * sa;
            float inv_src_a = 1.0f - sa;
            float dst_a_src_a = da[l] * sa;
            float cr = fmaf(fill[0], inv_dst_a_src_a, bl[0] * dst_a_src_a);
            float cg = fmaf(fill[1], inv_dst_a_src_a, bl[1] * dst_a_src_a);
            float cb = fmaf(fill[2], inv_dst_a_src_a, bl[2] * dst_a_src_a);
            dr[l] = fmaf(dr[l], inv_src_a, cr);
            dg[l] = fmaf(dg[l], inv_src_a, cg);
            db[l] = fmaf(db[l], inv_src_a, cb);
            da[l] = fmaf(da[l], inv_src_a, sa);
        }
    }

    // compute_srgb + LinearLayout::write (mod.rs:466-483, layout/mod.rs:265-282).
    const uint32_t px = tx * 16u + x;
    if (px < S.width) {
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            uint32_t py = ty * 16u + half * 8u + l;
            if (py < S.height) {
                uint32_t rgba = pixel_to_srgb_bytes(dr[l], dg[l], db[l], da[l], S.channels);
                *reinterpret_cast<uint32_t*>(in.framebuffer + (size_t)py * S.stride + (size_t)px * 4u) = rgba;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Host launchers
// ---------------------------------------------------------------------------
uint32_t cell_num_blocks(uint32_t n) { return (n + kCellThreads - 1) / kCellThreads; }

void launch_cell_count(const uint64_t* segs, uint32_t n, uint32_t* block_counts, uint32_t* total, cudaStream_t st) {
    uint32_t nb = cell_num_blocks(n);
    cell_count_kernel<<<nb, kCellThreads, 0, st>>>(segs, n, block_counts);
    launch_scan_u32(block_counts, nb, total, st);
}

void launch_cell_write(const uint64_t* segs, uint32_t n, const uint32_t* block_offsets, uint32_t* cell_start,
                       uint64_t* cell_key, uint32_t n_cells, cudaStream_t st) {
    cell_write_kernel<<<cell_num_blocks(n), kCellThreads, 0, st>>>(segs, n, block_offsets, cell_start, cell_key, n_cells);
}

void launch_cell_cover(const uint64_t* segs, const uint32_t* cell_start, const uint64_t* cell_key, uint32_t n_cells,
                       uint4* cell_cover, uint64_t* key2, uint32_t* perm, cudaStream_t st) {
    cell_cover_kernel<<<(n_cells + 127) / 128, 128, 0, st>>>(segs, cell_start, cell_key, n_cells, cell_cover, key2, perm);
}

void launch_carry_scan(const PaintScene& S, const uint64_t* key2, const uint32_t* perm, const uint4* cell_cover,
                       uint32_t n_cells, uint4* carry_in, uint4* carry_after, uint32_t* gap_count, cudaStream_t st) {
    carry_scan_kernel<<<(n_cells + 127) / 128, 128, 0, st>>>(S, key2, perm, cell_cover, n_cells, carry_in, carry_after,
                                                              gap_count);
}

void launch_entry_fill(const PaintScene& S, const uint64_t* key2, const uint32_t* perm, const uint64_t* cell_key,
                       const uint4* carry_after, const uint32_t* gap_count, const uint32_t* gap_offset, uint32_t n_cells,
                       uint64_t* ekey, uint32_t* eid, uint4* gap_carry, cudaStream_t st) {
    entry_fill_kernel<<<(n_cells + 127) / 128, 128, 0, st>>>(S, key2, perm, cell_key, carry_after, gap_count, gap_offset,
                                                              n_cells, ekey, eid, gap_carry);
}

void launch_tile_ranges(const PaintScene& S, const uint64_t* ekey, uint32_t n_entries, uint32_t* tile_begin,
                        uint32_t* tile_end, cudaStream_t st) {
    size_t bytes = (size_t)S.tiles_x * S.tiles_y * sizeof(uint32_t);
    cudaMemsetAsync(tile_begin, 0, bytes, st);
    cudaMemsetAsync(tile_end, 0, bytes, st);
    if (n_entries) tile_range_kernel<<<(n_entries + 255) / 256, 256, 0, st>>>(S, ekey, n_entries, tile_begin, tile_end);
}

void launch_paint(const PaintScene& S, const uint64_t* segs, const uint64_t* ekey, const uint32_t* eid,
                  const uint32_t* cell_start, const uint4* carry_in, const uint4* gap_carry, uint32_t n_cells,
                  const uint32_t* tile_begin, const uint32_t* tile_end, uint8_t* eflags, uint8_t* framebuffer,
                  cudaStream_t st) {
    if (S.tx_hi <= S.tx_lo || S.ty_hi <= S.ty_lo) return;
    uint32_t n_tiles = (S.tx_hi - S.tx_lo) * (S.ty_hi - S.ty_lo);
    PaintInputs in{segs, ekey, eid, cell_start, carry_in, gap_carry, n_cells, tile_begin, tile_end, eflags, framebuffer};
    paint_kernel<<<(n_tiles + kPaintWarps - 1) / kPaintWarps, kPaintWarps * 32, 0, st>>>(S, in, n_tiles);
}

}  // namespace forma
