// Stage 4b: the painter — one warp per 16x16 tile. Replaces
// LayerWorkbench::drive_tile_painting + the optimiser passes
// (cpu/painter/layer_workbench/mod.rs:280-342, passes/*.rs),
// Painter::paint_layer / blend_at / clip_at / compute_srgb
// (cpu/painter/mod.rs:290-483) and LinearLayout::write
// (cpu/buffer/layout/mod.rs:265-282).
//
// Lane l owns pixel column x = l / 2 and rows 8*(l % 2) .. +8 of the tile:
// exactly one f32x8 of the reference (cpu/painter/mod.rs:234-244), so the
// reference's only cross-lane rule ("skip the f32x8 when all 8 coverages are
// zero", mod.rs:317-319) is a per-thread test here.
//
// Warps are persistent: each takes the next tile from an atomic counter, so
// heavy tiles (dozens of translucent layers) do not hold back a whole CTA.
// The per-layer metadata of a tile (segment range, carry-in, packed style) is
// fetched by 32 lanes at once and handed round with shuffles, so the layer loop
// itself contains no dependent global loads besides the segment words, whose
// first chunk is prefetched one layer ahead.
#include "paint_common.cuh"
#include "paint_math.cuh"

namespace forma {

struct PaintInputs {
    const uint64_t* segs;
    const EntryRec* recs;        // sorted entries (kernels_tables.cu: merge_entries_kernel)
    const uint32_t* tile_begin;
    const uint32_t* tile_end;
    uint8_t* eflags;             // per sorted entry: optimizer flags, initialised with EntryRec::flags0
    uint8_t* framebuffer;
    uint32_t* tile_counter;
};

// Per-entry header, one entry per lane.
struct EntryHdr {
    uint32_t layer, seg0, seg1;
    uint4 carry;
    int32_t slot;
    // packed style: fill_rule | func<<1 | is_clipped<<2 | fill_type<<3 | blend_mode<<5
    uint32_t meta;
    uint32_t clip_layers;
    float color[4];
};

__device__ __forceinline__ uint32_t meta_fill_rule(uint32_t m) { return m & 1u; }
__device__ __forceinline__ uint32_t meta_func(uint32_t m) { return (m >> 1) & 1u; }
__device__ __forceinline__ bool meta_is_clipped(uint32_t m) { return (m >> 2) & 1u; }
__device__ __forceinline__ uint32_t meta_fill_type(uint32_t m) { return (m >> 3) & 3u; }
__device__ __forceinline__ uint32_t meta_blend(uint32_t m) { return (m >> 5) & 15u; }
__device__ __forceinline__ bool meta_unchanged(uint32_t m) { return (m >> 9) & 1u; }

__device__ __forceinline__ EntryHdr load_hdr(const PaintInputs& in, uint32_t p) {
    const uint4* q = reinterpret_cast<const uint4*>(in.recs + p);  // four independent 16-byte loads
    const uint4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    EntryHdr h;
    h.layer = q0.x;
    h.seg0 = q0.y;
    h.seg1 = q0.z;
    h.meta = q0.w;
    h.carry = q1;
    h.color[0] = __uint_as_float(q2.x);
    h.color[1] = __uint_as_float(q2.y);
    h.color[2] = __uint_as_float(q2.z);
    h.color[3] = __uint_as_float(q2.w);
    h.slot = (int32_t)q3.x;
    h.clip_layers = q3.y;
    return h;
}

__device__ __forceinline__ EntryHdr bcast_hdr(const EntryHdr& h, int src) {
    EntryHdr o;
    o.layer = __shfl_sync(kFullMask, h.layer, src);
    o.seg0 = __shfl_sync(kFullMask, h.seg0, src);
    o.seg1 = __shfl_sync(kFullMask, h.seg1, src);
    o.carry.x = __shfl_sync(kFullMask, h.carry.x, src);
    o.carry.y = __shfl_sync(kFullMask, h.carry.y, src);
    o.carry.z = __shfl_sync(kFullMask, h.carry.z, src);
    o.carry.w = __shfl_sync(kFullMask, h.carry.w, src);
    o.slot = __shfl_sync(kFullMask, h.slot, src);
    o.meta = __shfl_sync(kFullMask, h.meta, src);
    o.clip_layers = __shfl_sync(kFullMask, h.clip_layers, src);
#pragma unroll
    for (int k = 0; k < 4; ++k) o.color[k] = __shfl_sync(kFullMask, h.color[k], src);
    return o;
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

// Shared-memory cell of pixel (local_x, local_y): the 8 cells of lane l are at
// l + 32*k, so a warp's row-k access touches 32 consecutive words (no bank
// conflicts when the lanes read back / re-zero their own cells).
__device__ __forceinline__ uint32_t cell_index(uint32_t lx, uint32_t ly) { return (ly & 7u) * 32u + lx * 2u + (ly >> 3); }

// blend_at for one pixel once its fill colour is known (cpu/painter/mod.rs:420-447).
__device__ __forceinline__ float4 blend_fill(uint32_t mode, const float fill[4], float coverage, bool apply_clip, float clip,
                                             float4 dst) {
    float sa = fill[3] * coverage;
    if (apply_clip) sa *= clip;
    float bl[3];
    vblend::blend(mode, dst.x, dst.y, dst.z, fill[0], fill[1], fill[2], bl);
    float inv_dst_a = 1.0f - dst.w;
    float inv_dst_a_src_a = inv_dst_a * sa;
    float inv_src_a = 1.0f - sa;
    float dst_a_src_a = dst.w * sa;
    float cr = fmaf(fill[0], inv_dst_a_src_a, bl[0] * dst_a_src_a);
    float cg = fmaf(fill[1], inv_dst_a_src_a, bl[1] * dst_a_src_a);
    float cb = fmaf(fill[2], inv_dst_a_src_a, bl[2] * dst_a_src_a);
    return make_float4(fmaf(dst.x, inv_src_a, cr), fmaf(dst.y, inv_src_a, cg), fmaf(dst.z, inv_src_a, cb),
                       fmaf(dst.w, inv_src_a, sa));
}

// One pixel of blend_at (cpu/painter/mod.rs:406-447) for any fill / blend
// mode; only instantiated inside blend_column_generic (inlining it eight times
// per layer into the kernel made the kernel 15k instructions long).
__device__ __forceinline__ float4 blend_pixel_generic(const StyleRec* __restrict__ st, const StopRec* __restrict__ stops,
                                                   const uint16_t* __restrict__ texels, float fx, float fy, int l,
                                                   float coverage, bool apply_clip, float clip, float4 dst) {
    float fill[4];
    if (st->fill_type == 0u) {
        fill[0] = st->color[0]; fill[1] = st->color[1]; fill[2] = st->color[2]; fill[3] = st->color[3];
    } else if (st->fill_type == 1u) {
        gradient_at(*st, stops, fx, fy, l, fill);
    } else {
        texture_at(*st, texels, fx, fy, l, fill);
    }
    return blend_fill(st->blend_mode, fill, coverage, apply_clip, clip, dst);
}

// The eight pixels of a lane (one f32x8) in one call: px = r[8] g[8] b[8] a[8] in
// local memory. One call per layer instead of eight keeps the register
// save / restore traffic around the call out of the pixel loop; the style record
// (and, for gradients of up to four stops, the stops) are loaded once per call.
__device__ __noinline__ void blend_column_generic(const StyleRec* __restrict__ st_ptr, const StopRec* __restrict__ stops,
                                                  const uint16_t* __restrict__ texels, float fx, float fy,
                                                  const float* __restrict__ cov, bool apply_clip,
                                                  const float* __restrict__ clip /* stride 32 */, float* __restrict__ px) {
    const StyleRec s = *st_ptr;
    if (s.fill_type == 1u && s.stop_count <= 4u) {
        const GradientSetup g = gradient_setup(s, stops);
#pragma unroll 2  // measured: 1 -> 10.5 ms, 2 -> 7.4 ms, 4 -> 10.2 ms (spills) on circles8k
        for (int l = 0; l < 8; ++l) {
            float fill[4];
            gradient_at_small(s, g, fx, fy, l, fill);
            float4 d = blend_fill(s.blend_mode, fill, cov[l], apply_clip, apply_clip ? clip[l * 32] : 1.0f,
                                  make_float4(px[l], px[8 + l], px[16 + l], px[24 + l]));
            px[l] = d.x;
            px[8 + l] = d.y;
            px[16 + l] = d.z;
            px[24 + l] = d.w;
        }
        return;
    }
#pragma unroll 2
    for (int l = 0; l < 8; ++l) {
        float4 d = make_float4(px[l], px[8 + l], px[16 + l], px[24 + l]);
        d = blend_pixel_generic(&s, stops, texels, fx, fy, l, cov[l], apply_clip, apply_clip ? clip[l * 32] : 1.0f, d);
        px[l] = d.x;
        px[8 + l] = d.y;
        px[16 + l] = d.z;
        px[24 + l] = d.w;
    }
}

// Slab mapping (paint_kernel<_, true>): the eight pixels of a lane lie in one row, pixel j
// at x0 + 2 j; `active` bit j says whether this lane blends pixel j (its f32x8 of the
// reference has a non-zero coverage, cpu/painter/mod.rs:317-319). (fy, l) is the f32x8 base
// row and the row inside it, exactly the operands gradient_at / texture_at combine.
__device__ __noinline__ void blend_row_generic(const StyleRec* __restrict__ st_ptr, const StopRec* __restrict__ stops,
                                               const uint16_t* __restrict__ texels, uint32_t x0, float fy, int l,
                                               const float* __restrict__ cov, uint32_t active, bool apply_clip,
                                               const float* __restrict__ clip /* stride 32 */, float* __restrict__ px) {
    const StyleRec s = *st_ptr;
    if (s.fill_type == 1u && s.stop_count <= 4u) {
        const GradientSetup g = gradient_setup(s, stops);
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {
            if (!((active >> j) & 1u)) continue;
            float fill[4];
            gradient_at_small(s, g, (float)(x0 + 2u * (uint32_t)j), fy, l, fill);
            float4 d = blend_fill(s.blend_mode, fill, cov[j], apply_clip, apply_clip ? clip[j * 32] : 1.0f,
                                  make_float4(px[j], px[8 + j], px[16 + j], px[24 + j]));
            px[j] = d.x;
            px[8 + j] = d.y;
            px[16 + j] = d.z;
            px[24 + j] = d.w;
        }
        return;
    }
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        if (!((active >> j) & 1u)) continue;
        float4 d = make_float4(px[j], px[8 + j], px[16 + j], px[24 + j]);
        d = blend_pixel_generic(&s, stops, texels, (float)(x0 + 2u * (uint32_t)j), fy, l, cov[j], apply_clip,
                                apply_clip ? clip[j * 32] : 1.0f, d);
        px[j] = d.x;
        px[8 + j] = d.y;
        px[16 + j] = d.z;
        px[24 + j] = d.w;
    }
}

__device__ __noinline__ uint32_t srgb_bytes_any_order(float r, float g, float b, float a, const uint32_t* ch) {
    return pixel_to_srgb_bytes(r, g, b, a, ch);
}

// The scalar blend of the solid-tile fold (all 16 modes) stays out of line too.
__device__ __noinline__ Rgba blend_solid(uint32_t mode, Rgba dst, Rgba src) { return sblend::blend(mode, dst, src); }

constexpr int kPaintWarpsPerBlock = 2;

// kMinBlocks trades registers for resident warps (8 -> 128 regs, 10 -> 96 regs).
//
// kSlab selects the pixel <-> lane mapping of the per-entry work (everything per tile is the
// same code):
//   false  lane = (column x = lane / 2, rows 8 (lane % 2) .. +8): one f32x8 of the reference per
//          lane; every entry costs the same ~550 warp instructions whatever it covers.
//   true   lane = (column parity p = lane / 16, row r = lane % 16); the tile is walked in eight
//          "slabs" of two columns, left to right, carrying the running cover of the lane's row
//          in a register. Slabs left / right of the entry's segments only see the carry / the
//          final cover, slabs without any coverage are skipped by the whole warp, and only the
//          cells of touched slabs are read and re-zeroed: the cost follows the covered area
//          (profiles/r1_paint_kernel_analysis.md). Per-pixel arithmetic is the same.
//          EXPERIMENTAL: selected with FORMA_PAINT_KERNEL=slab only; written after the GPU
//          budget of round 1 was spent, so it has not run on a device yet.
template <int kMinBlocks, bool kSlab>
__global__ void __launch_bounds__(kPaintWarpsPerBlock * 32, kMinBlocks) paint_kernel(PaintScene S, PaintInputs in, uint32_t n_tiles) {
    __shared__ int32_t s_area[kPaintWarpsPerBlock][256];
    __shared__ int32_t s_cover[kPaintWarpsPerBlock][256];
    __shared__ float s_clip[kPaintWarpsPerBlock][256];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    int32_t* area = s_area[warp];
    int32_t* cover = s_cover[warp];
    const Rgba clear{S.clear[0], S.clear[1], S.clear[2], S.clear[3]};
    const bool rgba_order = S.channels[0] == 0u && S.channels[1] == 1u && S.channels[2] == 2u && S.channels[3] == 3u;
    const uint32_t ntx = S.tx_hi - S.tx_lo;
    const uint32_t x = lane >> 1, half = lane & 1u;
    // The cells of this warp start (and are kept) zeroed.
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        area[l * 32 + lane] = 0;
        cover[l * 32 + lane] = 0;
    }
    __syncwarp();

    // Tile tickets are drawn two tiles ahead and the entry range of the next tile is
    // loaded while the current one is painted, so a warp never waits for the
    // counter or for tile_begin / tile_end between tiles.
    uint32_t cur = 0, next_raw = 0, cur_b = 0, cur_e = 0;
    if (lane == 0) cur = atomicAdd(in.tile_counter, 1u);
    cur = __shfl_sync(kFullMask, cur, 0);
    if (lane == 0) next_raw = atomicAdd(in.tile_counter, 1u);
    if (cur < n_tiles) {
        const uint32_t t0 = (S.ty_lo + cur / ntx) * S.tiles_x + S.tx_lo + cur % ntx;
        cur_b = in.tile_begin[t0];
        cur_e = in.tile_end[t0];
    }
    while (cur < n_tiles) {
        const uint32_t tile_lin = cur, b = cur_b, e = cur_e;
        {
            const uint32_t nxt = __shfl_sync(kFullMask, next_raw, 0);
            uint32_t nb = 0, ne = 0;
            if (nxt < n_tiles) {
                const uint32_t t1 = (S.ty_lo + nxt / ntx) * S.tiles_x + S.tx_lo + nxt % ntx;
                nb = in.tile_begin[t1];
                ne = in.tile_end[t1];
            }
            if (lane == 0) next_raw = atomicAdd(in.tile_counter, 1u);
            cur = nxt;
            cur_b = nb;
            cur_e = ne;
        }
        const uint32_t ty = S.ty_lo + tile_lin / ntx, tx = S.tx_lo + tile_lin % ntx;
        const uint32_t tid = ty * S.tiles_x + tx;

        // ---- optimizer passes (layer_workbench/passes/*.rs) ------------------
        // Pass A: per-entry facts, 32 entries at a time.
        // (has-segments / full / unchanged facts were computed when the entries were built.)
        bool any_clip = false, all_unchanged = true;
        for (uint32_t p0 = b; p0 < e; p0 += 32u) {
            uint32_t p = p0 + lane;
            uint32_t f = p < e ? (uint32_t)in.eflags[p] : kFlagUnchanged;
            any_clip |= __any_sync(kFullMask, (f & kFlagClipish) != 0u);
            all_unchanged = all_unchanged && !__any_sync(kFullMask, (f & kFlagUnchanged) == 0u);
        }

        // tile_unchanged pass (passes/tile_unchanged.rs:24-57) — only with a layer cache.
        const bool use_cache = S.cache_tiles != nullptr;
        uint32_t cache_x = 0, cache_solid = 0;
        bool layers_were_removed = true;  // PassesSharedState::reset
        if (use_cache) {
            uint2 c = S.cache_tiles[tid];
            const uint32_t layers = (e - b) & 0xFFFFFFu;
            const bool had = (c.x >> 31) != 0u;
            const uint32_t previous = c.x & 0xFFFFFFu;
            cache_solid = c.y;
            cache_x = (1u << 31) | (c.x & (1u << 30)) | layers;  // update_layer_count(Some(layers))
            bool is_unchanged = false;
            if (had) {
                layers_were_removed = layers < previous;
                is_unchanged = previous == layers && all_unchanged;
            }
            if (S.clear_unchanged && is_unchanged) {  // TileWriteOp::None
                if (lane == 0) S.cache_tiles[tid] = make_uint2(cache_x, cache_solid);
                continue;
            }
        }

        // Pass B: skip_trivial_clips (sequential; only tiles that contain clips).
        if (any_clip) {
            if (lane == 0) {
                bool has_clip = false, clip_full = false, clip_used = false;
                uint32_t clip_last = 0, clip_i = 0;
                for (uint32_t p = b; p < e; ++p) {
                    uint32_t f = in.eflags[p];
                    if (!(f & kFlagClipish)) {  // neither a clip nor a clipped layer
                        if (has_clip && in.recs[p].layer > clip_last) {
                            has_clip = false;
                            if (!clip_used) in.eflags[clip_i] |= (uint8_t)kFlagMaskedOut;
                        }
                        continue;
                    }
                    const uint32_t id = in.recs[p].layer;
                    if (!(f & kFlagClippedDraw)) {  // Func::Clip
                        clip_full = (f & kFlagFull) != 0;
                        clip_last = id + in.recs[p].clip_layers;
                        clip_i = p;
                        clip_used = false;
                        has_clip = true;
                        if (clip_full) f |= kFlagMaskedOut;
                    }
                    if (f & kFlagClippedDraw) {
                        if (has_clip && id <= clip_last) {
                            if (clip_full) f |= kFlagSkipClip;
                            else clip_used = true;
                        } else {
                            f |= kFlagMaskedOut;
                        }
                    }
                    in.eflags[p] = (uint8_t)f;
                    if (has_clip && id > clip_last) {
                        has_clip = false;
                        if (!clip_used) in.eflags[clip_i] |= (uint8_t)kFlagMaskedOut;
                    }
                }
                if (has_clip && !clip_used) in.eflags[clip_i] |= (uint8_t)kFlagMaskedOut;
            }
            __syncwarp();
        }

        // Pass C: skip_fully_covered_layers — the top-most full, unclipped, opaque
        // `Over` solid layer culls everything below it.
        uint32_t first_paint = b;  // MaskedVec::skip_until
        bool incomplete = false;   // an "interesting" incomplete cover at or above the opaque layer
        bool have_opaque = false;
        bool visible_unchanged = !layers_were_removed;  // skip_fully_covered_layers.rs:38-47
        for (uint32_t hi = e; hi > b && !have_opaque;) {
            uint32_t lo = hi - b >= 32u ? hi - 32u : b;
            uint32_t p = lo + lane;
            bool inc = false, cand = false, changed = false;
            if (p < hi) {
                uint32_t f = in.eflags[p];
                if (!(f & kFlagMaskedOut)) {
                    bool clipped = (f & kFlagClippedDraw) && !(f & kFlagSkipClip);
                    if (clipped || !(f & kFlagFull)) inc = true;
                    else if (f & kFlagOpaque) cand = true;
                    changed = !(f & kFlagUnchanged);
                }
            }
            uint32_t cand_mask = __ballot_sync(kFullMask, cand);
            uint32_t inc_mask = __ballot_sync(kFullMask, inc);
            uint32_t changed_mask = __ballot_sync(kFullMask, changed);
            if (cand_mask) {
                uint32_t top = 31u - (uint32_t)__clz((int)cand_mask);
                have_opaque = true;
                first_paint = lo + top;
                if (top < 31u && (inc_mask >> (top + 1u)) != 0u) incomplete = true;
                if ((changed_mask >> top) != 0u) visible_unchanged = false;  // layers visited before the break
            } else {
                if (inc_mask) incomplete = true;
                if (changed_mask) visible_unchanged = false;
            }
            hi = lo;
        }
        if (use_cache && have_opaque && !incomplete && visible_unchanged) {
            // Everything visible is unchanged: nothing to draw (skip_fully_covered_layers.rs:86-89).
            if (lane == 0) S.cache_tiles[tid] = make_uint2(cache_x, cache_solid);
            continue;
        }

        if (!incomplete) {
            // Every visible layer is full: fold with the scalar blend and emit a
            // solid tile (skip_fully_covered_layers.rs:81-118, mod.rs:686-704).
            Rgba dst = clear;
            uint32_t p = first_paint;
            if (have_opaque) {
                const EntryRec& r = in.recs[p];
                dst = Rgba{r.color[0], r.color[1], r.color[2], r.color[3]};
                ++p;
            }
            bool solid = true;
            for (; p < e; ++p) {
                if (in.eflags[p] & kFlagMaskedOut) continue;
                const EntryRec& r = in.recs[p];
                if (meta_func(r.meta) == 0u && meta_fill_type(r.meta) == 0u) {
                    dst = blend_solid(meta_blend(r.meta), dst, Rgba{r.color[0], r.color[1], r.color[2], r.color[3]});
                } else {
                    solid = false;
                    break;
                }
            }
            if (solid) {
                const uint32_t bytes = solid_to_srgb_bytes(dst, S.channels);
                // CachedTile::convert_optimizer_op (cpu/painter/mod.rs:686-704): the same
                // solid colour as last frame is not written again.
                const bool same = use_cache && ((cache_x >> 30) & 1u) && cache_solid == bytes;
                if (!same) store_tile_solid(S, in.framebuffer, tx, ty, lane, bytes);
                if (lane == 0) {
                    if (use_cache) S.cache_tiles[tid] = make_uint2(cache_x | (1u << 30), bytes);
                    if (!same && S.written_list) S.written_list[atomicAdd(S.written_count, 1u)] = tid;
                }
                continue;
            }
        }
        if (lane == 0) {
            // update_solid_color(None): the tile is painted
            if (use_cache) S.cache_tiles[tid] = make_uint2(cache_x & ~(1u << 30), cache_solid);
            if (S.written_list) S.written_list[atomicAdd(S.written_count, 1u)] = tid;
        }

        // ---- paint (layer_workbench/mod.rs:301-337, cpu/painter/mod.rs:290-347) ---
        float dr[8], dg[8], db[8], da[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            dr[l] = clear.r; dg[l] = clear.g; db[l] = clear.b; da[l] = clear.a;
        }
        bool clip_active = false;
        uint32_t clip_last = 0;
        // The clip mask lives in shared memory (lane-private slots l*32 + lane): it
        // is only touched by tiles that contain clip layers.
        float* clip_mask = s_clip[warp] + lane;
        const float fx = (float)(x + tx * 16u);
        const float fy = (float)(half * 8u + ty * 16u);

        for (uint32_t p0 = first_paint; p0 < e; p0 += 32u) {
            const uint32_t cnt = min(32u, e - p0);
            EntryHdr mine{};
            uint32_t my_flags = kFlagMaskedOut;
            if (lane < cnt) {
                mine = load_hdr(in, p0 + lane);
                my_flags = in.eflags[p0 + lane];
            }
            // Prefetch the first segment chunk of the first entry of this group.
            uint32_t nflags = __shfl_sync(kFullMask, my_flags, 0);
            uint64_t pre = 0;
            {
                uint32_t s0 = __shfl_sync(kFullMask, mine.seg0, 0), s1 = __shfl_sync(kFullMask, mine.seg1, 0);
                if (!(nflags & kFlagMaskedOut) && s0 + lane < s1) pre = in.segs[s0 + lane];
            }

            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t flags = nflags;
                const uint64_t first_seg = pre;
                if (k + 1 < cnt) {  // start fetching the next entry's segments
                    nflags = __shfl_sync(kFullMask, my_flags, (int)k + 1);
                    uint32_t s0 = __shfl_sync(kFullMask, mine.seg0, (int)k + 1);
                    uint32_t s1 = __shfl_sync(kFullMask, mine.seg1, (int)k + 1);
                    pre = 0;
                    if (!(nflags & kFlagMaskedOut) && s0 + lane < s1) pre = in.segs[s0 + lane];
                }
                if (flags & kFlagMaskedOut) continue;
                const EntryHdr er = bcast_hdr(mine, (int)k);
                const uint32_t fill_rule = meta_fill_rule(er.meta);

                if constexpr (kSlab) {
                    // ---- slab walk (see the comment above the kernel) --------------------
                    const uint32_t row = lane & 15u, par = lane >> 4;
                    // acc_segment (cpu/painter/mod.rs:257-271) into column-major cells: the cell of
                    // (column 2 j + par, row) is word 32 j + lane, so a slab is one conflict-free access.
                    uint32_t x_lo = 16u, x_hi = 0u;  // columns this entry's segments touch
                    if (er.seg1 > er.seg0) {
                        for (uint32_t i = er.seg0 + lane; i < er.seg1; i += 32u) {
                            uint64_t s = (i < er.seg0 + 32u) ? first_seg : in.segs[i];
                            const uint32_t lx = (uint32_t)(s >> 16) & 15u, ly = (uint32_t)(s >> 12) & 15u;
                            int32_t cv = (int32_t)(((uint32_t)s & 0x3Fu) ^ 0x20u) - 0x20;
                            int32_t dam = (int32_t)((uint32_t)(s >> 6) & 0x3Fu);
                            atomicAdd(&area[lx * 16u + ly], dam * cv);
                            atomicAdd(&cover[lx * 16u + ly], cv);
                            x_lo = min(x_lo, lx);
                            x_hi = max(x_hi, lx);
                        }
                        x_lo = __reduce_min_sync(kFullMask, x_lo);
                        x_hi = __reduce_max_sync(kFullMask, x_hi);
                        __syncwarp();
                    }
                    const bool has_cells = x_lo <= x_hi;
                    const uint32_t s_first = x_lo >> 1, s_last = x_hi >> 1;  // slabs with cells (if has_cells)

                    if (clip_active && clip_last < er.layer) clip_active = false;  // mod.rs:302-306
                    const bool is_clip = meta_func(er.meta) == 1u;
                    if (is_clip && !clip_active) {  // clip_at, mod.rs:449-464
                        clip_active = true;
                        clip_last = er.layer + er.clip_layers;
                    }
                    const bool apply_clip = meta_is_clipped(er.meta) && !(flags & kFlagSkipClip);
                    const bool draws = !is_clip && !(apply_clip && !clip_active);  // mod.rs:321-323

                    // Running cover of this lane's row over the columns left of the current slab
                    // (i8, wrapping like the reference's lanes), starting from the carry-in.
                    const uint32_t cw = row < 4u ? er.carry.x : row < 8u ? er.carry.y : row < 12u ? er.carry.z : er.carry.w;
                    int32_t run = (int32_t)(int8_t)((cw >> (8u * (row & 3u))) & 0xFFu);
                    // Slabs without cells see 32 * run only: one coverage before the cells, one after.
                    float cov_flat = coverage_of(32 * run, fill_rule);
                    uint32_t nz_flat = __ballot_sync(kFullMask, cov_flat != 0.0f);

                    float cov[8];
                    uint32_t act = 0u;  // bit j: this lane's f32x8 of slab j has a non-zero coverage
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint32_t nz;
                        if (has_cells && (uint32_t)j >= s_first && (uint32_t)j <= s_last) {
                            const int idx = j * 32 + (int)lane;
                            const int32_t a = (int32_t)(int16_t)area[idx];
                            const int32_t c = (int32_t)(int8_t)cover[idx];
                            area[idx] = 0;
                            cover[idx] = 0;
                            const int32_t c_other = __shfl_xor_sync(kFullMask, c, 16);
                            // column 2 j sees the covers left of the slab, column 2 j + 1 also column 2 j's
                            const int32_t here = (int32_t)(int8_t)(run + (par ? c_other : 0));
                            cov[j] = coverage_of(32 * here + a, fill_rule);  // compute_doubled_areas, mod.rs:388-404
                            run = (int32_t)(int8_t)(run + c + c_other);
                            nz = __ballot_sync(kFullMask, cov[j] != 0.0f);
                            if ((uint32_t)j == s_last) {  // right of the cells only the final cover counts
                                cov_flat = coverage_of(32 * run, fill_rule);
                                nz_flat = __ballot_sync(kFullMask, cov_flat != 0.0f);
                            }
                        } else {
                            cov[j] = cov_flat;
                            nz = nz_flat;
                        }
                        if (is_clip) clip_mask[j * 32] = cov[j];
                        if ((nz >> (lane & 24u)) & 0xFFu) act |= 1u << j;
                    }
                    if (has_cells) __syncwarp();
                    if (!draws) continue;
                    const uint32_t slabs = __reduce_or_sync(kFullMask, act);  // slabs somebody covers (warp-uniform)
                    if (!slabs) continue;

                    const uint32_t mode = meta_blend(er.meta);
                    const uint32_t fill_type = meta_fill_type(er.meta);
                    if (fill_type == 0u && mode == 0u) {  // blend_at, mod.rs:406-447, solid `Over`
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (!((slabs >> j) & 1u)) continue;  // nobody covers slab j: skipped by the whole warp
                            if (!((act >> j) & 1u)) continue;    // this lane's f32x8 is all zero (mod.rs:317-319)
                            float sa = er.color[3] * cov[j];
                            if (apply_clip) sa *= clip_mask[j * 32];
                            float inv_dst_a_src_a = (1.0f - da[j]) * sa;
                            float inv_src_a = 1.0f - sa;
                            float dst_a_src_a = da[j] * sa;
                            float cr = fmaf(er.color[0], inv_dst_a_src_a, er.color[0] * dst_a_src_a);
                            float cg = fmaf(er.color[1], inv_dst_a_src_a, er.color[1] * dst_a_src_a);
                            float cb = fmaf(er.color[2], inv_dst_a_src_a, er.color[2] * dst_a_src_a);
                            dr[j] = fmaf(dr[j], inv_src_a, cr);
                            dg[j] = fmaf(dg[j], inv_src_a, cg);
                            db[j] = fmaf(db[j], inv_src_a, cb);
                            da[j] = fmaf(da[j], inv_src_a, sa);
                        }
                    } else if (act) {
                        const StyleRec* st = &S.styles[er.slot];
                        float px[32];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            px[j] = dr[j]; px[8 + j] = dg[j]; px[16 + j] = db[j]; px[24 + j] = da[j];
                        }
                        blend_row_generic(st, S.stops, S.texels, tx * 16u + par, (float)((row >> 3) * 8u + ty * 16u), (int)(row & 7u), cov,
                                          act, apply_clip, clip_mask, px);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            dr[j] = px[j]; dg[j] = px[8 + j]; db[j] = px[16 + j]; da[j] = px[24 + j];
                        }
                    }
                } else {
                    // acc_segment: scatter-add the cell's segments (cpu/painter/mod.rs:257-271).
                    int32_t a8[8];
                    uint32_t run_lo, run_hi;  // running covers of rows 0-3 / 4-7 of this lane's half, packed i8
                    if (er.seg1 > er.seg0) {
                        for (uint32_t i = er.seg0 + lane; i < er.seg1; i += 32u) {
                            uint64_t s = (i < er.seg0 + 32u) ? first_seg : in.segs[i];
                            uint32_t cell = cell_index((uint32_t)(s >> 16) & 15u, (uint32_t)(s >> 12) & 15u);
                            int32_t cv = (int32_t)(((uint32_t)s & 0x3Fu) ^ 0x20u) - 0x20;
                            int32_t dam = (int32_t)((uint32_t)(s >> 6) & 0x3Fu);
                            atomicAdd(&area[cell], dam * cv);
                            atomicAdd(&cover[cell], cv);
                        }
                        __syncwarp();
                        uint32_t c_lo = 0, c_hi = 0;
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            int idx = l * 32 + (int)lane;  // == cell_index(x, half * 8 + l)
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

                    if (meta_func(er.meta) == 1u) {  // clip_at, mod.rs:449-464
                        if (!clip_active) {
                            clip_active = true;
                            clip_last = er.layer + er.clip_layers;
                        }
#pragma unroll
                        for (int l = 0; l < 8; ++l) clip_mask[l * 32] = cov[l];
                        continue;
                    }
                    const bool apply_clip = meta_is_clipped(er.meta) && !(flags & kFlagSkipClip);
                    if (all_zero) continue;                    // mod.rs:317-319 (whole f32x8 is zero)
                    if (apply_clip && !clip_active) continue;  // mod.rs:321-323

                    const uint32_t mode = meta_blend(er.meta);
                    const uint32_t fill_type = meta_fill_type(er.meta);
                    // blend_at, mod.rs:406-447. The mode / fill dispatch is hoisted out of
                    // the pixel loop: a solid `Over` layer (by far the most common) is
                    // straight-line code; everything else goes through one out-of-line
                    // helper per pixel so that the kernel stays small enough for the
                    // instruction cache.
                    if (fill_type == 0u && mode == 0u) {
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            float sa = er.color[3] * cov[l];
                            if (apply_clip) sa *= clip_mask[l * 32];
                            float inv_dst_a_src_a = (1.0f - da[l]) * sa;
                            float inv_src_a = 1.0f - sa;
                            float dst_a_src_a = da[l] * sa;
                            float cr = fmaf(er.color[0], inv_dst_a_src_a, er.color[0] * dst_a_src_a);
                            float cg = fmaf(er.color[1], inv_dst_a_src_a, er.color[1] * dst_a_src_a);
                            float cb = fmaf(er.color[2], inv_dst_a_src_a, er.color[2] * dst_a_src_a);
                            dr[l] = fmaf(dr[l], inv_src_a, cr);
                            dg[l] = fmaf(dg[l], inv_src_a, cg);
                            db[l] = fmaf(db[l], inv_src_a, cb);
                            da[l] = fmaf(da[l], inv_src_a, sa);
                        }
                    } else {
                        const StyleRec* st = &S.styles[er.slot];
                        float px[32], cv[8];
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            px[l] = dr[l]; px[8 + l] = dg[l]; px[16 + l] = db[l]; px[24 + l] = da[l];
                            cv[l] = cov[l];
                        }
                        blend_column_generic(st, S.stops, S.texels, fx, fy, cv, apply_clip, clip_mask, px);
#pragma unroll
                        for (int l = 0; l < 8; ++l) {
                            dr[l] = px[l]; dg[l] = px[8 + l]; db[l] = px[16 + l]; da[l] = px[24 + l];
                        }
                    }
                }
            }
        }

        // compute_srgb + LinearLayout::write (mod.rs:466-483, layout/mod.rs:265-282).
        if constexpr (kSlab) {
            // Pixel j of a lane is (2 j + par, row): transpose through the (now idle) clip-mask
            // words so that a store instruction writes two whole 64-byte tile rows.
            uint32_t* stage = reinterpret_cast<uint32_t*>(s_clip[warp]);
            const uint32_t row = lane & 15u, par = lane >> 4;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                stage[row * 16u + 2u * (uint32_t)j + par] =
                    rgba_order ? pixel_to_srgb_bytes_rgba(dr[j], dg[j], db[j], da[j])
                               : srgb_bytes_any_order(dr[j], dg[j], db[j], da[j], S.channels);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t w = (uint32_t)k * 32u + lane;
                const uint32_t py = ty * 16u + (w >> 4), qx = tx * 16u + (w & 15u);
                if (qx < S.width && py < S.height)
                    *reinterpret_cast<uint32_t*>(in.framebuffer + (size_t)py * S.stride + (size_t)qx * 4u) = stage[w];
            }
            __syncwarp();
            continue;
        }
        const uint32_t px = tx * 16u + x;
        if (px < S.width) {
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                uint32_t py = ty * 16u + half * 8u + l;
                if (py < S.height) {
                    // RGBA order (kernel-uniform) needs no channel selection; other orders take the
                    // out-of-line generic conversion.
                    uint32_t rgba = rgba_order ? pixel_to_srgb_bytes_rgba(dr[l], dg[l], db[l], da[l])
                                               : srgb_bytes_any_order(dr[l], dg[l], db[l], da[l], S.channels);
                    *reinterpret_cast<uint32_t*>(in.framebuffer + (size_t)py * S.stride + (size_t)px * 4u) = rgba;
                }
            }
        }
    }
}

// Packs the tiles named in `list` (written by paint_kernel) into 1 KB records,
// row-major 16x16 RGBA8, so that a damaged frame costs a device->host copy
// proportional to the damage. One warp per tile.
__global__ void __launch_bounds__(256) gather_tiles_kernel(const uint8_t* __restrict__ fb, uint32_t stride, uint32_t width,
                                                           uint32_t height, uint32_t tiles_x,
                                                           const uint32_t* __restrict__ list,
                                                           const uint32_t* __restrict__ count, uint32_t* __restrict__ out) {
    const uint32_t n = *count;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = blockIdx.x * 8u + (threadIdx.x >> 5); i < n; i += gridDim.x * 8u) {
        const uint32_t tid = list[i];
        const uint32_t x0 = (tid % tiles_x) * 16u, y0 = (tid / tiles_x) * 16u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t p = (uint32_t)k * 32u + lane;  // pixel index in the tile, row-major
            const uint32_t x = x0 + (p & 15u), y = y0 + (p >> 4);
            uint32_t v = 0;
            if (x < width && y < height) v = *reinterpret_cast<const uint32_t*>(fb + (size_t)y * stride + (size_t)x * 4u);
            out[(size_t)i * 256u + p] = v;
        }
    }
}

void launch_gather_tiles(const PaintScene& S, const uint8_t* framebuffer, uint32_t* packed, cudaStream_t st) {
    gather_tiles_kernel<<<148 * 4, 256, 0, st>>>(framebuffer, S.stride, S.width, S.height, S.tiles_x, S.written_list,
                                                  S.written_count, packed);
}

void launch_paint(const PaintScene& S, const uint64_t* segs, const EntryRec* recs, const uint32_t* tile_begin,
                  const uint32_t* tile_end, uint8_t* eflags, uint8_t* framebuffer, uint32_t* tile_counter, cudaStream_t st) {
    if (S.tx_hi <= S.tx_lo || S.ty_hi <= S.ty_lo) return;
    uint32_t n_tiles = (S.tx_hi - S.tx_lo) * (S.ty_hi - S.ty_lo);
    cudaMemsetAsync(tile_counter, 0, sizeof(uint32_t), st);
    PaintInputs in{segs, recs, tile_begin, tile_end, eflags, framebuffer, tile_counter};
    // Persistent warps: enough CTAs to fill every SM at the kernel's occupancy.
    // FORMA_PAINT_REGS=96 selects the 96-register build (default: 128 registers,
    // measured 17 % faster on paris@4K: fewer spills beat the extra warps).
    // FORMA_PAINT_KERNEL=slab selects the experimental slab mapping (see paint_kernel).
    static int blocks_per_sm = 0, variant = 0;
    if (!blocks_per_sm) {
        const char* e = getenv("FORMA_PAINT_REGS");
        const char* k = getenv("FORMA_PAINT_KERNEL");
        variant = (k && !strcmp(k, "slab")) ? 1 : (e && atoi(e) == 96) ? 10 : 8;
        if (variant == 8) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, paint_kernel<8, false>, kPaintWarpsPerBlock * 32, 0);
        else if (variant == 10) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, paint_kernel<10, false>, kPaintWarpsPerBlock * 32, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, paint_kernel<8, true>, kPaintWarpsPerBlock * 32, 0);
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t want = (uint32_t)(blocks_per_sm * sms);
    uint32_t need = (n_tiles + kPaintWarpsPerBlock - 1) / kPaintWarpsPerBlock;
    if (variant == 8) paint_kernel<8, false><<<min(want, need), kPaintWarpsPerBlock * 32, 0, st>>>(S, in, n_tiles);
    else if (variant == 10) paint_kernel<10, false><<<min(want, need), kPaintWarpsPerBlock * 32, 0, st>>>(S, in, n_tiles);
    else paint_kernel<8, true><<<min(want, need), kPaintWarpsPerBlock * 32, 0, st>>>(S, in, n_tiles);
}

}  // namespace forma
