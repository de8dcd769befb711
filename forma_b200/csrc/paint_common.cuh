// Helpers shared by the painter-table kernels (kernels_tables.cu) and the
// painter itself (kernels_painter.cu): key field access, packed 16 x i8 covers.
#pragma once

#include "cuda_common.cuh"
#include "kernels.h"

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

// --- per-entry optimizer flags (layer_workbench passes) -----------------------------
constexpr uint32_t kFlagHasSegs = 1, kFlagFull = 2, kFlagMaskedOut = 4, kFlagSkipClip = 8, kFlagUnchanged = 16;
constexpr uint32_t kFlagClipish = 32;      // Func::Clip, or a Draw layer with is_clipped
constexpr uint32_t kFlagOpaque = 64;       // Draw, solid fill, BlendMode::Over, alpha == 1
constexpr uint32_t kFlagClippedDraw = 128; // Draw layer with is_clipped

// Tiles with at least kHeavyMin entries are listed by class (class c: 2^(c+4) <= entries <
// 2^(c+5), the last class open-ended) so that the painter starts them first.
constexpr int kHeavyClasses = kHeavyListClasses;
constexpr uint32_t kHeavyMin = 16;
__device__ __forceinline__ int heavy_class(uint32_t entries) {
    const int c = 27 - __clz((int)entries);  // floor(log2) - 4
    return c > kHeavyClasses - 1 ? kHeavyClasses - 1 : c;
}

// packed style: fill_rule | func<<1 | is_clipped<<2 | fill_type<<3 | blend_mode<<5 | unchanged<<9 |
//               small gradient (<= 4 stops: the painter uses its GradRec)<<10
constexpr uint32_t kMetaSmallGradient = 1u << 10;
__device__ __forceinline__ uint32_t pack_style_meta(const StyleRec& st, bool unchanged) {
    return (st.fill_rule & 1u) | ((st.func & 1u) << 1) | ((st.is_clipped ? 1u : 0u) << 2) | ((st.fill_type & 3u) << 3) |
           ((st.blend_mode & 15u) << 5) | (unchanged ? (1u << 9) : 0u) |
           ((st.fill_type == 1u && st.stop_count >= 2u && st.stop_count <= 4u) ? kMetaSmallGradient : 0u);
}

__device__ __forceinline__ uint32_t fill_rule_of(const PaintScene& S, uint32_t layer) {
    int32_t slot = layer < S.n_orders ? S.order_to_style[layer] : -1;
    return slot >= 0 ? S.styles[slot].fill_rule : 0u;
}


// Cells that cannot influence a painted tile (rows outside the painted band,
// columns right of it) get this key in both pair sorts: it sorts last and keeps
// the sort keys inside host-known bounds (tile_y field = tiles_y + 1).
__device__ __host__ __forceinline__ uint64_t sentinel_key(uint32_t tiles_y) { return (uint64_t)(tiles_y + 1u) << 53; }

}  // namespace forma
