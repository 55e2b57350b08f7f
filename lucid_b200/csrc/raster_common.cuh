// raster_common.cuh -- device code shared by the three raster kernels (raster_bins.cu, raster_sort.cu,
// raster_shade.cu): scanline evaluation, block depth keys, half-block records, the warp sort, the texture filter,
// the per-sample shading arithmetic and the per-pixel reduction window.
//
// Replaces data/shaders/raster_low.glsl, raster_high.glsl, shared/raster.glsl and shared/shading.glsl.  Same
// results, different decomposition:
//   k_raster_bins   a persistent 256-thread CTA per bin walks the bin's triangles once and appends a record to the
//                   list of every 8x4 half-block (HIGH) or 8x8 block (LOW) the triangle covers (the reference
//                   builds per-row lists and filters them per block column: raster_low.glsl:22-32,58-63,81-106,
//                   raster_high.glsl:54-144); every non-empty list becomes a work item with a slice of the
//                   sorted-entry stream.
//   k_block_sort    warp per item: depth keys from the centroid of the covered pixels, register-tile bitonic sort,
//                   depth ties by triangle index (deterministic image), then the item's entries are written IN
//                   SORTED ORDER as two contiguous 16-byte planes: (triangle, pixel masks) and (depth plane,
//                   constant colour).
//   k_block_shade   warp per item, lane = pixel: streams the item's sorted entries, a 32x32 bit transpose of the
//                   pixel masks hands every pixel lane its own sample list (instead of regrouping shaded samples
//                   through shared-memory atomics and shuffles, raster.glsl:358-396), samples are shaded from
//                   per-entry data staged in shared memory, 3-entry window and blend as in shading.glsl:186-314.
#pragma once
#include "common.cuh"

namespace lucid {

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int SEGMENT_SIZE = 256;
constexpr int MAX_BLOCK_TRIS = 256;		 // raster_low.glsl:17
constexpr int MAX_HBLOCK_TRIS = 4096;	 // raster_high.glsl:27
constexpr int HB_LIST_CAP = MAX_HBLOCK_TRIS; // records per half-block list (HIGH)
constexpr int BLOCK_WARPS = 4;			 // warps per CTA of k_block_sort / k_block_shade
constexpr u32 AUX_VARYING = 0x00ffffffu; // alpha 0 with colour bits set: never produced by shadeConstant

__device__ __forceinline__ const int *cntc(const Params &p, int which) {
	return p.counts + (size_t)which * p.bin_count;
}
// list `sub` of a bin: a fixed slot (32 x 4096 x 8 bytes per bin: HIGH lists of 4096 8-byte records, LOW lists of 256
// 16-byte records), or -- LUCID_CREATE_COMPACT_LISTS -- where k_raster_bins put it in the pool
__device__ __forceinline__ unsigned char *blockList(const Params &p, int bin_id, int sub, bool high) {
	if(p.compact_lists)
		return reinterpret_cast<unsigned char *>(p.block_lists) + (size_t)p.list_offsets[bin_id * 32 + sub] * 8;
	return reinterpret_cast<unsigned char *>(p.block_lists) + (size_t)bin_id * BIN_LIST_BYTES +
		   (high ? (size_t)sub * HB_LIST_CAP * 8 : (size_t)sub * MAX_BLOCK_TRIS * 16);
}
__device__ __forceinline__ unsigned char *binLists(const Params &p, int bin_id) {
	return reinterpret_cast<unsigned char *>(p.block_lists) + (size_t)bin_id * BIN_LIST_BYTES;
}

// work items of the block stages are queued by size class (entries of the list), heaviest class first
constexpr int ITEM_CLASSES = 5;
__device__ __forceinline__ int itemClass(int entries) {
#ifndef RB_CL0
#define RB_CL0 384
#define RB_CL1 160
#define RB_CL2 64
#define RB_CL3 24
#endif
	const int limits[ITEM_CLASSES - 1] = {RB_CL0, RB_CL1, RB_CL2, RB_CL3};
	int k = ITEM_CLASSES - 1;
#pragma unroll
	for(int c = ITEM_CLASSES - 2; c >= 0; c--)
		if(entries > limits[c])
			k = c;
	return k;
}
// work_counters: [0] bins taken (k_raster_bins) [1] items taken by k_block_sort [2] items taken by k_block_shade
// [3..7] items per size class [8] sorted-stream entries handed out
constexpr int WC_BINS = 0, WC_SORT = 1, WC_SHADE = 2, WC_CLASS = 3, WC_STREAM = 8, WC_LIST_POOL = 9, WC_COUNT = 12;

// A work item: item = bin << 6 | high << 5 | block, its list length and the start of its slice of the sorted stream
struct WorkItem {
	u32 item, count, offset, pad;
};
// queue order: class by class, heaviest first; i-th item overall
__device__ __forceinline__ uint4 fetchWorkItem(const Params &p, u32 i, const u32 (&class_end)[ITEM_CLASSES]) {
	int k = 0;
	u32 first = 0;
#pragma unroll
	for(int c = 0; c < ITEM_CLASSES - 1; c++)
		if(i >= class_end[c])
			k = c + 1, first = class_end[c];
	return __ldcg(p.block_items + (size_t)k * p.block_items_cap + (i - first));
}

__device__ __forceinline__ uint4 *workItemSlot(const Params &p, u32 i, const u32 (&class_end)[ITEM_CLASSES]) {
	int k = 0;
	u32 first = 0;
#pragma unroll
	for(int c = 0; c < ITEM_CLASSES - 1; c++)
		if(i >= class_end[c])
			k = c + 1, first = class_end[c];
	return p.block_items + (size_t)k * p.block_items_cap + (i - first);
}
// LUCID_OPT_OPAQUE_PREPASS is honoured by the front-to-back blend only (lucid_abi.h)
__host__ __device__ __forceinline__ bool opaquePrepass(const Params &p) {
	return (p.opts & LUCID_OPT_OPAQUE_PREPASS) != 0 && (p.opts & (LUCID_OPT_ADDITIVE_BLENDING | LUCID_OPT_ALPHA_THRESHOLD)) == 0 &&
		   p.opaque_depth != nullptr;
}

// ------------------------------------------------------------------------------------------------
// scanline evaluation (scanline.glsl:13-26, raster.glsl:116-140)

struct RowScan {
	float scan[3], step[3];
	u32 xneg;
};

// four pixel rows: 5-bit xmin / xmax per row and the mask of touched 8-pixel columns
__device__ __forceinline__ void rasterBinStep(RowScan &r, u32 &min_bits, u32 &max_bits, u32 &bx_mask) {
	const float inf = __int_as_float(0x7f800000);
	min_bits = max_bits = bx_mask = 0;
#pragma unroll
	for(int row = 0; row < 4; row++) {
		float mn0 = (r.xneg & 1) ? -inf : r.scan[0], mx0 = (r.xneg & 1) ? r.scan[0] : inf;
		float mn1 = (r.xneg & 2) ? -inf : r.scan[1], mx1 = (r.xneg & 2) ? r.scan[1] : inf;
		float mn2 = (r.xneg & 4) ? -inf : r.scan[2], mx2 = (r.xneg & 4) ? r.scan[2] : inf;
		int imin = f2i(fmaxf(fmaxf(mn0, mn1), fmaxf(mn2, 0.0f)));
		int imax = f2i(fminf(fminf(mx0, mx1), fminf(mx2, float(BIN_SIZE)))) - 1;
		if(imin > imax)
			imin = BIN_SIZE - 1, imax = 0;
		r.scan[0] += r.step[0], r.scan[1] += r.step[1], r.scan[2] += r.step[2];
		min_bits |= (u32)imin << (5 * row);
		max_bits |= (u32)imax << (5 * row);
		bx_mask |= (0xfu << (imin >> 3)) & (0xfu >> (3 - (imax >> 3)));
	}
	bx_mask &= 0xfu;
}

// raster.glsl:170-176
__device__ __forceinline__ u32 blockDepth(uint4 d, float cx, float cy, float range) {
	float ray_pos = __uint_as_float(d.x) * cx + (__uint_as_float(d.y) * cy + __uint_as_float(d.z));
	float depth = range * saturatef(rsqrt_rn(ray_pos + 1.0f));
	return f2u(depth);
}

// ------------------------------------------------------------------------------------------------
// shading (shading.glsl:64-184)

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }

// Filter definition.  The reference leaves filtering to the Vulkan sampler (textureGrad on a trilinear sampler,
// lucid_base.h:30-31, shading.glsl:153-158).  Here the texture unit does it: the atlases are CUDA mipmapped arrays
// behind texture objects (RGBA8 unorm, normalised coordinates, wrap, linear + mip-linear, no anisotropy) and a
// sample is one tex2DLod.  Its arithmetic -- 8-bit weights split level -> x -> y, 16-bit unorm texels -- is
// restated in integers in the CPU checker (oracle/lucid_oracle.cpp textureUnitSample; fitted and verified bit for
// bit with tools/hwtex/), so the images still compare exactly.  lod: log2 of the larger screen-space derivative in
// texels, taken piecewise linearly from the exponent / mantissa bits of its square; one rounding per operation
// (it addresses the texture, like a coordinate).
__device__ __forceinline__ float4 sampleTexture(const Params &p, int slot, float u, float v, float dudx, float dvdx,
												float dudy, float dvdy) {
	if(p.tex_object[slot] == 0)
		return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	const float w0 = float(p.tex_width[slot]), h0 = float(p.tex_height[slot]);
	const float ax = dudx * w0, ay = dvdx * h0, bx = dudy * w0, by = dvdy * h0;
	const float rho2 = fmaxf(ax * ax + ay * ay, bx * bx + by * by);
	float lod = 0.0f;
	if(rho2 > 1.0f)
		lod = float((int)(__float_as_uint(rho2) - 0x3f800000u)) * (0.5f / 8388608.0f);
	lod = clampf(lod, 0.0f, float(p.tex_levels[slot] - 1));
	return tex2DLod<float4>((cudaTextureObject_t)p.tex_object[slot], u, v, lod);
}

// ------------------------------------------------------------------------------------------------
// shadeSample (shading.glsl:107-184) in two steps.  Everything the function reads of a triangle -- its record,
// the quad's vertex attributes, the instance's uv rectangle -- depends on the list entry only, so it is gathered
// ONCE per entry into a 128-byte stage (eight 16-byte words, shared memory in k_block_shade) and the per-sample
// part reads the stage instead of walking three dependent global loads per sample (getTriangleParams,
// getTriangleVertex*, g_instance_uv_rects; shading.glsl:64-105).
//   [0] bary0 (edge0.xyz, param0)   [1] bary1 (edge1.xyz, param1)
//   [2] flags | instance << 16, instance RGBA8, Lambert term of the flat normal (float bits), -
//   [3] tex0.xy, tex1.xy   [4] tex2.xy, -, -   [5] instance uv rectangle
//   [6] vertex colours c0 c1 c2, -   [7] vertex normals n0 n1 n2, -
constexpr int STAGE_WORDS = 8; // uint4 per entry

__device__ __forceinline__ float lambertTerm(const LightTerms &lt, F3 normal) {
	const float ndl = __fmaf_rn(lt.msun[0], normal.x, __fmaf_rn(lt.msun[1], normal.y, lt.msun[2] * normal.z));
	return fmaxf(0.0f, __fmaf_rn(ndl, 0.7f, 0.3f));
}

template <typename Dst> __device__ __forceinline__ void stageEntry(const Params &p, const LightTerms &lt, u32 tri_idx, Dst &&dst) {
	const uint4 *rec = reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx);
	const uint4 dq = __ldg(rec), misc = __ldg(rec + 1);
	dst[0] = __ldg(rec + 2), dst[1] = __ldg(rec + 3);
	const u32 flags = dq.w & 0xffffu, instance_id = dq.w >> 16;
	const u32 second = tri_idx & 1, quad_idx = tri_idx >> 1;
	float lv = 0.0f;
	if(!(flags & LUCID_INST_HAS_VERTEX_NORMALS))
		lv = lambertTerm(lt, decodeNormalUint(misc.x));
	dst[2] = make_uint4(dq.w, misc.y, __float_as_uint(lv), 0u);
	if(flags & LUCID_INST_HAS_ALBEDO_TEXTURE) {
		const uint4 q0 = __ldg(p.quad_uv + (size_t)quad_idx * 2), q1 = __ldg(p.quad_uv + (size_t)quad_idx * 2 + 1);
		dst[3] = make_uint4(q0.x, q0.y, second == 0 ? q0.z : q1.x, second == 0 ? q0.w : q1.y);
		dst[4] = make_uint4(second == 0 ? q1.x : q1.z, second == 0 ? q1.y : q1.w, 0u, 0u);
		if(flags & LUCID_INST_HAS_UV_RECT) {
			const float4 r = __ldg(p.inst_uv_rects + instance_id);
			dst[5] = make_uint4(__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w));
		}
	}
	if(flags & LUCID_INST_HAS_VERTEX_COLORS) {
		uint4 c = make_uint4(0, 0, 0, 0);
		if(p.vertex_colors)
			c = __ldg(p.quad_colors + quad_idx);
		dst[6] = make_uint4(c.x, second ? c.z : c.y, second ? c.w : c.z, 0u);
	}
	if(flags & LUCID_INST_HAS_VERTEX_NORMALS) {
		uint4 n = make_uint4(0, 0, 0, 0);
		if(p.vertex_normals)
			n = __ldg(p.quad_normals + quad_idx);
		dst[7] = make_uint4(n.x, second ? n.z : n.y, second ? n.w : n.z, 0u);
	}
}

// the per-sample part; dplane = the triangle's depth plane (xyz).  Returns the RGBA8 sample (0 = no sample).
template <typename Src>
__device__ __forceinline__ u32 shadeStaged(const Params &p, const ColourTables &tab, const LightTerms &lt, Src &&e, float dx,
										   float dy, float dz, float px, float py) {
	const uint4 b0q = e[0], b1q = e[1], m = e[2];
	const u32 flags = m.x & 0xffffu;
	const float e0x = __uint_as_float(b0q.x), e0y = __uint_as_float(b0q.y), e0z = __uint_as_float(b0q.z);
	const float e1x = __uint_as_float(b1q.x), e1y = __uint_as_float(b1q.y), e1z = __uint_as_float(b1q.z);
	// the sample depth orders the blending: one rounding per operation, as in the reference; everything below it
	// is colour (fused multiply-adds, see the colour contract in common.cuh)
	const float inv_ray_pos = dx * px + (dy * py + dz);
	const float ray_pos = rcp(inv_ray_pos);
	const float e0 = __fmaf_rn(e0x, px, __fmaf_rn(e0y, py, e0z));
	const float e1 = __fmaf_rn(e1x, px, __fmaf_rn(e1y, py, e1z));
	float b0 = e0 * ray_pos, b1 = e1 * ray_pos;

	float bdx0 = 0, bdx1 = 0, bdy0 = 0, bdy1 = 0;
	const bool textured = (flags & LUCID_INST_HAS_ALBEDO_TEXTURE) != 0;
	if(textured) {
		const float ray_posx = rcp(inv_ray_pos + dx);
		const float ray_posy = rcp(inv_ray_pos + dy);
		bdx0 = __fmaf_rn(e0 + e0x, ray_posx, -b0), bdx1 = __fmaf_rn(e1 + e1x, ray_posx, -b1);
		bdy0 = __fmaf_rn(e0 + e0y, ray_posy, -b0), bdy1 = __fmaf_rn(e1 + e1y, ray_posy, -b1);
	}
	b0 -= __uint_as_float(b0q.w), b1 -= __uint_as_float(b1q.w);

	float4 color = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	if(flags & LUCID_INST_HAS_COLOR)
		color = decodeRGBA8(m.y);

	if(textured) {
		const uint4 q0 = e[3], q1 = e[4];
		const float t0x = __uint_as_float(q0.x), t0y = __uint_as_float(q0.y);
		const float t1x = __uint_as_float(q0.z), t1y = __uint_as_float(q0.w);
		const float t2x = __uint_as_float(q1.x), t2y = __uint_as_float(q1.y);
		float u = __fmaf_rn(b0, t1x, __fmaf_rn(b1, t2x, t0x)), v = __fmaf_rn(b0, t1y, __fmaf_rn(b1, t2y, t0y));
		float dudx = __fmaf_rn(bdx0, t1x, bdx1 * t2x), dvdx = __fmaf_rn(bdx0, t1y, bdx1 * t2y);
		float dudy = __fmaf_rn(bdy0, t1x, bdy1 * t2x), dvdy = __fmaf_rn(bdy0, t1y, bdy1 * t2y);
		if(flags & LUCID_INST_HAS_UV_RECT) {
			const uint4 rq = e[5];
			const float rx = __uint_as_float(rq.x), ry = __uint_as_float(rq.y), rz = __uint_as_float(rq.z), rw = __uint_as_float(rq.w);
			u = __fmaf_rn(rz, fractf(u), rx), v = __fmaf_rn(rw, fractf(v), ry);
			dudx *= rz, dvdx *= rw, dudy *= rz, dvdy *= rw;
		}
		const bool tex_opaque = (flags & LUCID_INST_TEX_OPAQUE) != 0;
		float4 tc = sampleTexture(p, tex_opaque ? 0 : 1, u, v, dudx, dvdx, dudy, dvdy);
		if(tex_opaque)
			tc.w = 1.0f;
		color.x *= tc.x, color.y *= tc.y, color.z *= tc.z, color.w *= tc.w;
	}
	if(flags & LUCID_INST_HAS_VERTEX_COLORS) {
		const uint4 c = e[6];
		const float4 c0 = decodeRGBA8(c.x), c1 = decodeRGBA8(c.y), c2 = decodeRGBA8(c.z);
		const float w0 = 1.0f - b0 - b1;
		color.x *= __fmaf_rn(w0, c0.x, __fmaf_rn(b0, c1.x, b1 * c2.x));
		color.y *= __fmaf_rn(w0, c0.y, __fmaf_rn(b0, c1.y, b1 * c2.y));
		color.z *= __fmaf_rn(w0, c0.z, __fmaf_rn(b0, c1.z, b1 * c2.z));
		color.w *= __fmaf_rn(w0, c0.w, __fmaf_rn(b0, c1.w, b1 * c2.w));
	}
	if(color.w == 0.0f)
		return 0;

	float lv = __uint_as_float(m.z);
	if(flags & LUCID_INST_HAS_VERTEX_NORMALS) {
		const uint4 n = e[7];
		const F3 n0 = decodeNormalUint(n.x);
		const F3 n1 = decodeNormalUint(n.y) - n0, n2 = decodeNormalUint(n.z) - n0;
		lv = lambertTerm(lt, mk3(__fmaf_rn(b0, n1.x, __fmaf_rn(b1, n2.x, n0.x)), __fmaf_rn(b0, n1.y, __fmaf_rn(b1, n2.y, n0.y)),
								 __fmaf_rn(b0, n1.z, __fmaf_rn(b1, n2.z, n0.z))));
	}
	color.x = finalShadeFast(tab, color.x, __fmaf_rn(lt.sun[0], lv, lt.amb[0]));
	color.y = finalShadeFast(tab, color.y, __fmaf_rn(lt.sun[1], lv, lt.amb[1]));
	color.z = finalShadeFast(tab, color.z, __fmaf_rn(lt.sun[2], lv, lt.amb[2]));
	return encodeRGBA8(color);
}
// ------------------------------------------------------------------------------------------------
// per-pixel reduction: 3-entry insertion window (shading.glsl:186-314)

struct Reducer {
	float d0, d1, d2, d3;
	u32 c0, c1, c2;
	float trans;
	float r, g, b;
	u32 invalid;
};
__device__ __forceinline__ void reducerInit(Reducer &s) {
	s.d0 = s.d1 = s.d2 = s.d3 = 999999999.0f;
	s.c0 = s.c1 = s.c2 = 0;
	s.trans = 1.0f;
	s.r = s.g = s.b = 0.0f;
	s.invalid = 0;
}
__device__ __forceinline__ void reducerBlend(Reducer &s, u32 c, bool additive) {
	float4 cc = decodeRGBA8(c);
	if(additive) {
		s.r = __fmaf_rn(cc.x, cc.w, s.r), s.g = __fmaf_rn(cc.y, cc.w, s.g), s.b = __fmaf_rn(cc.z, cc.w, s.b);
	} else {
		const float wt = cc.w * s.trans;
		s.r = __fmaf_rn(cc.x, wt, s.r), s.g = __fmaf_rn(cc.y, wt, s.g), s.b = __fmaf_rn(cc.z, wt, s.b);
		s.trans = __fmaf_rn(-cc.w, s.trans, s.trans);
	}
}
__device__ __forceinline__ void reducerPush(Reducer &s, u32 color, float depth, bool additive,
											bool vis_errors) {
	if(depth > s.d0) {
		u32 tc = color;
		color = s.c0, s.c0 = tc;
		float td = depth;
		depth = s.d0, s.d0 = td;
		if(s.d0 > s.d1) {
			tc = s.c0, s.c0 = s.c1, s.c1 = tc;
			td = s.d0, s.d0 = s.d1, s.d1 = td;
			if(s.d1 > s.d2) {
				tc = s.c1, s.c1 = s.c2, s.c2 = tc;
				td = s.d1, s.d1 = s.d2, s.d2 = td;
				if(vis_errors && s.d2 > s.d3) {
					// the window was too small for this pixel (shading.glsl:258-265)
					s.invalid++;
					s.r = 1.0f, s.g = 0.0f, s.b = 0.0f, s.trans = 0.0f;
					return;
				}
			}
		}
	}
	s.d3 = s.d2, s.d2 = s.d1, s.d1 = s.d0, s.d0 = depth;
	if(s.c2 != 0)
		reducerBlend(s, s.c2, additive);
	s.c2 = s.c1, s.c1 = s.c0, s.c0 = color;
}

// ------------------------------------------------------------------------------------------------
// warp-level sort of u32 keys in shared memory (ascending)
//
// Bitonic network in its "mirrored" form: the first step of every merge level pairs element e with
// e ^ (k - 1), the remaining steps pair e with e ^ j, and every compare-exchange moves the smaller
// key to the lower index -- no per-run direction.  A lane holds K consecutive keys in registers
// (element e = lane * K + r): steps with a partner distance below K are register-to-register
// min/max pairs, the others one shuffle per key.  Tiles of 256 keys (K = 8) are sorted entirely in
// registers; only the steps with distance >= 256 of larger lists go through shared memory.

// The steps between registers of one lane are unrolled (static register indices); the steps across
// lanes run as loops over the lane distance, which keeps the code of the four tile sizes small
// enough to stay in the instruction cache next to the shading loop.
template <int K> __device__ __forceinline__ void sortRegsInLane(u32 (&v)[K]) { // distances K/2 .. 1
#pragma unroll
	for(int j = K / 2; j >= 1; j >>= 1) {
#pragma unroll
		for(int r = 0; r < K; r++)
			if((r & j) == 0) {
				u32 lo = min(v[r], v[r | j]), hi = max(v[r], v[r | j]);
				v[r] = lo, v[r | j] = hi;
			}
	}
}
template <int K> __device__ __forceinline__ void sortRegsAcrossLanes(u32 (&v)[K], int first_lm, u32 lane) {
#pragma unroll 1
	for(int lm = first_lm; lm >= 1; lm >>= 1) {
		const bool lower = (lane & lm) == 0;
#pragma unroll
		for(int r = 0; r < K; r++) {
			u32 o = __shfl_xor_sync(0xffffffffu, v[r], lm);
			v[r] = lower ? min(v[r], o) : max(v[r], o);
		}
	}
}
// merge steps with partner distances first_j, first_j / 2, ... 1 (first_j >= K)
template <int K> __device__ __forceinline__ void sortRegsMergeSteps(u32 (&v)[K], int first_j, u32 lane) {
	sortRegsAcrossLanes<K>(v, first_j / K, lane);
	sortRegsInLane<K>(v);
}

// full sort of the 32 * K keys held by the warp
template <int K> __device__ __forceinline__ void sortRegs(u32 (&v)[K], u32 lane) {
	// merge levels inside a lane (k <= K)
#pragma unroll
	for(int k = 2; k <= K; k <<= 1) {
#pragma unroll
		for(int r = 0; r < K; r++)
			if((r & (k >> 1)) == 0) {
				const int q = r ^ (k - 1);
				u32 lo = min(v[r], v[q]), hi = max(v[r], v[q]);
				v[r] = lo, v[q] = hi;
			}
#pragma unroll
		for(int j = k >> 2; j >= 1; j >>= 1) {
#pragma unroll
			for(int r = 0; r < K; r++)
				if((r & j) == 0) {
					u32 lo = min(v[r], v[r | j]), hi = max(v[r], v[r | j]);
					v[r] = lo, v[r | j] = hi;
				}
		}
	}
	// merge levels k = 2 K top: element e pairs with e ^ (k - 1), i.e. lane ^ (2 top - 1), register r ^ (K - 1)
#pragma unroll 1
	for(int top = 1; top < 32; top <<= 1) {
		const bool lower = (lane & top) == 0;
		u32 o[K];
#pragma unroll
		for(int r = 0; r < K; r++)
			o[r] = __shfl_xor_sync(0xffffffffu, v[r ^ (K - 1)], 2 * top - 1);
#pragma unroll
		for(int r = 0; r < K; r++)
			v[r] = lower ? min(v[r], o[r]) : max(v[r], o[r]);
		sortRegsAcrossLanes<K>(v, top >> 1, lane);
		sortRegsInLane<K>(v);
	}
}

template <int K> __device__ __forceinline__ void sortSingleTile(u32 *keys, int n, u32 lane) {
	u32 v[K];
#pragma unroll
	for(int r = 0; r < K; r++) {
		int e = lane * K + r;
		v[r] = e < n ? keys[e] : 0xffffffffu;
	}
	sortRegs<K>(v, lane);
#pragma unroll
	for(int r = 0; r < K; r++) {
		int e = lane * K + r;
		if(e < n)
			keys[e] = v[r];
	}
}

// compare-exchange steps of a merge level whose partner distance is at least one tile (256 keys)
__device__ __forceinline__ void mirrorStep(u32 *keys, int padded, int k, u32 lane) {
	const int half = k >> 1;
	for(int i = lane; i < (padded >> 1); i += 32) {
		int blk = i / half, idx = i - blk * half;
		int lo = blk * k + idx, hi = blk * k + (k - 1 - idx);
		u32 a = keys[lo], b = keys[hi];
		if(a > b)
			keys[lo] = b, keys[hi] = a;
	}
	__syncwarp();
}
__device__ __forceinline__ void distanceStep(u32 *keys, int padded, int j, u32 lane) {
	for(int i = lane; i < (padded >> 1); i += 32) {
		int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
		u32 a = keys[lo], b = keys[hi];
		if(a > b)
			keys[lo] = b, keys[hi] = a;
	}
	__syncwarp();
}
// the remaining steps (distance 128..1) of every 256-key tile, or a full sort of every tile
template <bool FULL_SORT> __device__ __forceinline__ void tileSteps(u32 *keys, int padded, u32 lane) {
	uint4 *tiles = reinterpret_cast<uint4 *>(keys);
	for(int base = 0; base < padded; base += 256) {
		uint4 a = tiles[(base >> 2) + lane * 2], b = tiles[(base >> 2) + lane * 2 + 1];
		u32 v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
		if(FULL_SORT)
			sortRegs<8>(v, lane);
		else
			sortRegsMergeSteps<8>(v, 128, lane);
		tiles[(base >> 2) + lane * 2] = make_uint4(v[0], v[1], v[2], v[3]);
		tiles[(base >> 2) + lane * 2 + 1] = make_uint4(v[4], v[5], v[6], v[7]);
	}
	__syncwarp();
}

// keys[0..n) ascending in shared memory; the array has room for n rounded up to a power of two
static __device__ __noinline__ void warpSortShared(u32 *keys, int n) {
	const u32 lane = laneId();
	if(n <= 1)
		return;
	if(n <= 32)
		sortSingleTile<1>(keys, n, lane);
	else if(n <= 64)
		sortSingleTile<2>(keys, n, lane);
	else if(n <= 128)
		sortSingleTile<4>(keys, n, lane);
	else if(n <= 256)
		sortSingleTile<8>(keys, n, lane);
	else {
		int padded = 512;
		while(padded < n)
			padded <<= 1;
		for(int i = n + lane; i < padded; i += 32)
			keys[i] = 0xffffffffu;
		__syncwarp();
		tileSteps<true>(keys, padded, lane);
		for(int k = 512; k <= padded; k <<= 1) {
			mirrorStep(keys, padded, k, lane);
			for(int j = k >> 2; j >= 256; j >>= 1)
				distanceStep(keys, padded, j, lane);
			tileSteps<false>(keys, padded, lane);
		}
	}
	__syncwarp();
}

// Lists longer than the shared-memory key array (up to the reference's 4096 per half-block) are
// sorted in an L2-resident global array: 1024-key blocks are staged through shared memory, only
// the steps with a partner distance of 1024 or more touch global memory directly.
constexpr int SMEM_KEYS = 1024;
static __device__ __noinline__ void warpSortLarge(u32 *gkeys, int n, u32 *skeys) {
	const u32 lane = laneId();
	int padded = 2 * SMEM_KEYS;
	while(padded < n)
		padded <<= 1;
	for(int i = n + lane; i < padded; i += 32)
		gkeys[i] = 0xffffffffu;
	__syncwarp();
	auto stage = [&](int base, bool load) {
		for(int i = lane; i < SMEM_KEYS; i += 32) {
			if(load)
				skeys[i] = __ldcg(gkeys + base + i);
			else
				__stcg(gkeys + base + i, skeys[i]);
		}
		__syncwarp();
	};
	for(int base = 0; base < padded; base += SMEM_KEYS) {
		stage(base, true);
		warpSortShared(skeys, SMEM_KEYS);
		stage(base, false);
	}
	for(int k = 2 * SMEM_KEYS; k <= padded; k <<= 1) {
		const int half = k >> 1;
		for(int i = lane; i < (padded >> 1); i += 32) {
			int blk = i / half, idx = i - blk * half;
			int lo = blk * k + idx, hi = blk * k + (k - 1 - idx);
			u32 a = __ldcg(gkeys + lo), b = __ldcg(gkeys + hi);
			if(a > b)
				__stcg(gkeys + lo, b), __stcg(gkeys + hi, a);
		}
		__syncwarp();
		for(int j = k >> 2; j >= SMEM_KEYS; j >>= 1) {
			for(int i = lane; i < (padded >> 1); i += 32) {
				int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
				u32 a = __ldcg(gkeys + lo), b = __ldcg(gkeys + hi);
				if(a > b)
					__stcg(gkeys + lo, b), __stcg(gkeys + hi, a);
			}
			__syncwarp();
		}
		for(int base = 0; base < padded; base += SMEM_KEYS) {
			stage(base, true);
			distanceStep(skeys, SMEM_KEYS, 512, lane);
			distanceStep(skeys, SMEM_KEYS, 256, lane);
			tileSteps<false>(skeys, SMEM_KEYS, lane);
			stage(base, false);
		}
	}
}

// Entries with equal quantised depth are ordered by triangle index.  The low key bits only make
// keys unique (they are list positions that depend on atomic arrival order); this pass makes the
// final order -- and therefore the image -- independent of them.  Runs of equal depth are mostly
// short: the lane that finds the start of a run sorts it by insertion, 32 runs at a time.  A long run
// (coplanar stacks, geometry so far away that the 18 / 22 key bits no longer separate it) would make
// that one lane's work quadratic: a run of long_run entries or more (TIE_RUN_LONG_SMEM when triOf reads shared
// memory, TIE_RUN_LONG when every comparison is a global load) is handed to onLongRun(first, length) instead and
// stays in arrival order here -- k_tie_runs (raster_sort.cu) ranks it in the sorted-entry stream.
// (Ranking long runs inside this kernel was built first: correct, but the extra code took the block sort of the
// 10M-triangle scene from 0.52 to 0.65 ms -- the kernel sits at the edge of the instruction cache.)
constexpr int TIE_RUN_LONG = 24, TIE_RUN_LONG_SMEM = 96;
template <typename TriOf, typename OnLongRun>
__device__ void warpFixDepthTies(u32 *keys, int n, int slot_bits, int long_run, TriOf triOf, OnLongRun onLongRun) {
	const int lane = laneId();
	const u32 slot_mask = (1u << slot_bits) - 1u;
	for(int i0 = 0; i0 + 1 < n; i0 += 32) {
		const int i = i0 + lane;
		bool start = false;
		if(i + 1 < n) {
			u32 d = keys[i] >> slot_bits;
			start = d == (keys[i + 1] >> slot_bits) && (i == 0 || d != (keys[i - 1] >> slot_bits));
		}
		if(start) {
			const u32 d = keys[i] >> slot_bits;
			for(int e = i + 1; e < n && (keys[e] >> slot_bits) == d; e++) {
				if(e - i >= long_run) {
					int end = e;
					while(end < n && (keys[end] >> slot_bits) == d)
						end++;
					onLongRun(i, end - i);
					break;
				}
				u32 ke = keys[e], te = triOf(ke & slot_mask);
				int q = e;
				while(q > i) {
					u32 kq = keys[q - 1];
					if(triOf(kq & slot_mask) <= te)
						break;
					keys[q] = kq;
					q--;
				}
				keys[q] = ke;
			}
		}
	}
	__syncwarp();
}

// ------------------------------------------------------------------------------------------------
// half-block records
//
// Phase A of a bin turns every (triangle, 4-row group, 8-pixel column) with coverage into one
// record appended to that half-block's list: the triangle index and, for each of the 4 pixel rows,
// the first covered x (3 bits) and the number of covered pixels (4 bits) -- the same content as the
// reference's half-block tri record (raster.glsl:152-161), kept per half-block from the start so
// that phase B never filters a row list.

// HIGH record (8 bytes): x = tri_idx | (mins & 0xff) << 24, y = mins >> 8 | maxs << 12, where
// mins / maxs are the bin-wide 5-bit spans of the group's four rows (raster.glsl:116-140).
// LOW record (16 bytes): x = tri_idx, y = mins of rows 0-3 | (maxs 0-3) << 20 (low 12 bits),
// z = mins of rows 4-7 | (maxs 4-7) << 20 (low 12 bits), w = the two maxs' high 8 bits.
__device__ __forceinline__ uint2 packHighRecord(u32 tri_idx, u32 mins, u32 maxs) {
	return make_uint2(tri_idx | (mins << 24), (mins >> 8) | (maxs << 12));
}
__device__ __forceinline__ uint4 packLowRecord(u32 tri_idx, u32 mn0, u32 mx0, u32 mn1, u32 mx1) {
	return make_uint4(tri_idx, mn0 | (mx0 << 20), mn1 | (mx1 << 20), (mx0 >> 12) | ((mx1 >> 12) << 8));
}

// raster.glsl:142-168: the spans clipped to the 8-pixel column starting at startx, as
// (first x, count) per row; then the pixel mask (bit y * 8 + x), fragment count, centroid sums
__device__ __forceinline__ u32 rowsToBits(u32 mins, u32 maxs, int startx, int &num_frags) {
	u32 bits = 0;
	num_frags = 0;
#pragma unroll
	for(int r = 0; r < 4; r++) {
		int mn = max((int)((mins >> (5 * r)) & 31) - startx, 0);
		int mx = min((int)((maxs >> (5 * r)) & 31) - startx, 7);
		int c = max(mx - mn + 1, 0);
		bits |= ((1u << c) - 1u) << ((mn & 7) + 8 * r);
		num_frags += c;
	}
	return bits;
}
__device__ __forceinline__ void rowsCentroid(u32 mins, u32 maxs, int startx, int &num_frags, int &csum_x, int &csum_y) {
	num_frags = 0, csum_x = 0, csum_y = 0;
#pragma unroll
	for(int r = 0; r < 4; r++) {
		int mn = max((int)((mins >> (5 * r)) & 31) - startx, 0);
		int mx = min((int)((maxs >> (5 * r)) & 31) - startx, 7);
		int c = max(mx - mn + 1, 0);
		num_frags += c;
		csum_x += (mn * 2 + c) * c;
		csum_y += (2 * r + 1) * c;
	}
}

// 32 x 32 bit-matrix transpose across the warp: lane j gives row j, lane p receives column p
__device__ __forceinline__ u32 transpose32(u32 x) {
	const u32 lane = laneId();
#pragma unroll
	for(int j = 16; j >= 1; j >>= 1) {
		const u32 m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
		u32 y = __shfl_xor_sync(0xffffffffu, x, j);
		x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
	}
	return x;
}

__device__ __forceinline__ void unpackHighRecord(uint2 r, u32 &tri, u32 &mins, u32 &maxs) {
	tri = r.x & 0xffffffu;
	mins = (r.x >> 24) | ((r.y & 0xfffu) << 8), maxs = r.y >> 12;
}
__device__ __forceinline__ void unpackLowRecord(uint4 r, bool lower, u32 &tri, u32 &mins, u32 &maxs) {
	tri = r.x;
	u32 w = lower ? r.z : r.y, hi = lower ? (r.w >> 8) : r.w;
	mins = w & 0xfffffu, maxs = (w >> 20) | ((hi & 0xffu) << 12);
}


} // namespace lucid
