// raster.cu -- per-bin rasterisation, block depth sort, sample shading and exact blending.
//
// Replaces data/shaders/raster_low.glsl, raster_high.glsl, shared/raster.glsl and
// shared/shading.glsl.  Same results, different decomposition:
//   * a persistent 256-thread CTA takes one bin at a time.  Phase A walks the bin's triangles
//     once and appends a record to the list of every 8x4 half-block (HIGH) or 8x8 block (LOW) the
//     triangle covers; the lists live in an L2-resident scratch slice owned by the CTA.  The
//     reference builds per-row lists first and filters them per block column
//     (raster_low.glsl:22-32,58-63,81-106; raster_high.glsl:54-144).
//   * phase B hands half-blocks / blocks to warps dynamically: depth keys, a register-tile
//     bitonic sort, then shading with lane = pixel.
//   * depth ties are broken by triangle index instead of an atomic arrival slot, so the output is
//     deterministic.
//   * instead of regrouping 32 shaded samples per round through shared-memory atomics and
//     shuffles (raster.glsl:358-396), a 32x32 bit transpose of the triangles' pixel masks gives
//     every pixel lane its own sample list for up to 256 samples at a time.
#include "common.cuh"

namespace lucid {

constexpr int RASTER_THREADS = 256;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int SEGMENT_SIZE = 256;
constexpr int SAMPLE_BUF = SEGMENT_SIZE + 32;
constexpr int MAX_BLOCK_TRIS = 256;		 // raster_low.glsl:17
constexpr int MAX_HBLOCK_TRIS = 4096;	 // raster_high.glsl:27

__device__ __forceinline__ const int *cntc(const Params &p, int which) {
	return p.counts + (size_t)which * p.bin_count;
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

// Filter definition (the reference leaves this to the Vulkan sampler; same arithmetic as
// oracle/lucid_oracle.cpp sampleTexture): repeat addressing, bilinear within a level, linear between
// the two nearest levels, isotropic lod.  Coordinates are wrapped once in floating point, so the
// 2x2 footprint leaves the level by at most one texel (compares, no integer remainder); lod takes
// log2 piecewise linearly from the exponent / mantissa bits; texels are filtered on the 0..255
// scale (byte -> float by a permute into the mantissa of 2^23) and scaled by 1/255 once.
__device__ __forceinline__ float4 texelBytes(u32 t) {
	const float magic = 8388608.0f; // 0x4b000000: float(2^23 + b) - 2^23 == float(b), exactly
	return make_float4(__uint_as_float(__byte_perm(t, 0x4b000000u, 0x7650)) - magic,
					   __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7651)) - magic,
					   __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7652)) - magic,
					   __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7653)) - magic);
}
// uf, vf in [0, 1]; result on the 0..255 scale
__device__ __forceinline__ float4 bilinear(const u32 *level_base, int w, int h, float uf, float vf) {
	float fx = __fmaf_rn(uf, float(w), -0.5f), fy = __fmaf_rn(vf, float(h), -0.5f);
	float x0f = floorf(fx), y0f = floorf(fy);
	float ax = fx - x0f, ay = fy - y0f;
	int x0 = f2i(x0f), y0 = f2i(y0f); // in [-1, size - 1]
	int x1 = x0 + 1, y1 = y0 + 1;
	if(x0 < 0)
		x0 += w;
	if(x1 >= w)
		x1 -= w;
	if(y0 < 0)
		y0 += h;
	if(y1 >= h)
		y1 -= h;
	const u32 *row0 = level_base + y0 * w, *row1 = level_base + y1 * w;
	u32 t00 = __ldg(row0 + x0), t10 = __ldg(row0 + x1), t01 = __ldg(row1 + x0), t11 = __ldg(row1 + x1);
	float4 c00 = texelBytes(t00), c10 = texelBytes(t10), c01 = texelBytes(t01), c11 = texelBytes(t11);
	float4 o;
#define LERP2(c)                                                                                   \
	{                                                                                              \
		float top = __fmaf_rn(c10.c - c00.c, ax, c00.c);                                           \
		float bot = __fmaf_rn(c11.c - c01.c, ax, c01.c);                                           \
		o.c = __fmaf_rn(bot - top, ay, top);                                                       \
	}
	LERP2(x) LERP2(y) LERP2(z) LERP2(w)
#undef LERP2
	return o;
}
__device__ __forceinline__ float4 sampleTexture(const Params &p, int slot, float u, float v, float dudx, float dvdx,
												float dudy, float dvdy) {
	const u32 *data = reinterpret_cast<const u32 *>(p.tex_data[slot]);
	if(data == nullptr)
		return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	const int wi = p.tex_width[slot], hi = p.tex_height[slot];
	float w0 = float(wi), h0 = float(hi);
	float ax = dudx * w0, ay = dvdx * h0, bx = dudy * w0, by = dvdy * h0;
	float rho2 = fmaxf(__fmaf_rn(ax, ax, ay * ay), __fmaf_rn(bx, bx, by * by));
	int levels = p.tex_levels[slot];
	float lod = 0.0f;
	if(rho2 > 1.0f)
		lod = float((int)(__float_as_uint(rho2) - 0x3f800000u)) * (0.5f / 8388608.0f);
	lod = clampf(lod, 0.0f, float(levels - 1));
	float l0f = floorf(lod);
	int l0 = f2i(l0f), l1 = min(l0 + 1, levels - 1);
	float a = lod - l0f;
	const float uf = u - floorf(u), vf = v - floorf(v);
	const float s = 1.0f / 255.0f;
	float4 c0 = bilinear(data + p.tex_level_offset[slot][l0], max(1, wi >> l0), max(1, hi >> l0), uf, vf);
	if(a == 0.0f || l1 == l0)
		return make_float4(c0.x * s, c0.y * s, c0.z * s, c0.w * s);
	float4 c1 = bilinear(data + p.tex_level_offset[slot][l1], max(1, wi >> l1), max(1, hi >> l1), uf, vf);
	return make_float4(__fmaf_rn(c1.x - c0.x, a, c0.x) * s, __fmaf_rn(c1.y - c0.y, a, c0.y) * s,
					   __fmaf_rn(c1.z - c0.z, a, c0.z) * s, __fmaf_rn(c1.w - c0.w, a, c0.w) * s);
}

__device__ __noinline__ u32 shadeSample(const Params &p, const LucidConfig &cfg, int ipx, int ipy, u32 tri_idx,
						   float &out_depth) {
	float px = float(ipx), py = float(ipy);
	const uint4 *rec = reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx);
	uint4 dq = __ldg(rec), misc = __ldg(rec + 1), b0q = __ldg(rec + 2), b1q = __ldg(rec + 3);
	// the attribute loads are issued together with the record (their addresses depend on tri_idx
	// only); whether they are used is decided by the instance flags once the record has arrived
	const u32 second = tri_idx & 1, quad_idx = tri_idx >> 1;
	uint4 attr_c = make_uint4(0, 0, 0, 0), attr_n = attr_c;
	if(p.vertex_colors)
		attr_c = __ldg(p.quad_colors + quad_idx);
	if(p.vertex_normals)
		attr_n = __ldg(p.quad_normals + quad_idx);
	float dx = __uint_as_float(dq.x), dy = __uint_as_float(dq.y), dz = __uint_as_float(dq.z);
	u32 flags = dq.w & 0xffffu, instance_id = dq.w >> 16;
	float e0x = __uint_as_float(b0q.x), e0y = __uint_as_float(b0q.y), e0z = __uint_as_float(b0q.z);
	float e1x = __uint_as_float(b1q.x), e1y = __uint_as_float(b1q.y), e1z = __uint_as_float(b1q.z);

	// the sample depth orders the blending: one rounding per operation, as in the reference; everything
	// below it is colour (fused multiply-adds, see the colour contract in common.cuh)
	float inv_ray_pos = dx * px + (dy * py + dz);
	out_depth = inv_ray_pos;
	if(misc.w != 0)
		return misc.z; // attribute-free triangle: colour was evaluated once in quad setup
	float ray_pos = rcp(inv_ray_pos);
	float e0 = __fmaf_rn(e0x, px, __fmaf_rn(e0y, py, e0z));
	float e1 = __fmaf_rn(e1x, px, __fmaf_rn(e1y, py, e1z));
	float b0 = e0 * ray_pos, b1 = e1 * ray_pos;

	float bdx0 = 0, bdx1 = 0, bdy0 = 0, bdy1 = 0;
	const bool textured = (flags & LUCID_INST_HAS_ALBEDO_TEXTURE) != 0;
	if(textured) {
		float ray_posx = rcp(inv_ray_pos + dx);
		float ray_posy = rcp(inv_ray_pos + dy);
		bdx0 = __fmaf_rn(e0 + e0x, ray_posx, -b0), bdx1 = __fmaf_rn(e1 + e1x, ray_posx, -b1);
		bdy0 = __fmaf_rn(e0 + e0y, ray_posy, -b0), bdy1 = __fmaf_rn(e1 + e1y, ray_posy, -b1);
	}
	b0 -= __uint_as_float(b0q.w), b1 -= __uint_as_float(b1q.w);

	float4 color = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	if(flags & LUCID_INST_HAS_COLOR)
		color = decodeRGBA8(misc.y);

	if(textured) {
		uint4 q0 = __ldg(p.quad_uv + (size_t)quad_idx * 2), q1 = __ldg(p.quad_uv + (size_t)quad_idx * 2 + 1);
		float t0x = __uint_as_float(q0.x), t0y = __uint_as_float(q0.y);
		float t1x = __uint_as_float(second == 0 ? q0.z : q1.x), t1y = __uint_as_float(second == 0 ? q0.w : q1.y);
		float t2x = __uint_as_float(second == 0 ? q1.x : q1.z), t2y = __uint_as_float(second == 0 ? q1.y : q1.w);
		float u = __fmaf_rn(b0, t1x, __fmaf_rn(b1, t2x, t0x)), v = __fmaf_rn(b0, t1y, __fmaf_rn(b1, t2y, t0y));
		float dudx = __fmaf_rn(bdx0, t1x, bdx1 * t2x), dvdx = __fmaf_rn(bdx0, t1y, bdx1 * t2y);
		float dudy = __fmaf_rn(bdy0, t1x, bdy1 * t2x), dvdy = __fmaf_rn(bdy0, t1y, bdy1 * t2y);
		if(flags & LUCID_INST_HAS_UV_RECT) {
			float4 r = __ldg(p.inst_uv_rects + instance_id);
			u = __fmaf_rn(r.z, fractf(u), r.x), v = __fmaf_rn(r.w, fractf(v), r.y);
			dudx *= r.z, dvdx *= r.w, dudy *= r.z, dvdy *= r.w;
		}
		const bool tex_opaque = (flags & LUCID_INST_TEX_OPAQUE) != 0;
		float4 tc = sampleTexture(p, tex_opaque ? 0 : 1, u, v, dudx, dvdx, dudy, dvdy);
		if(tex_opaque)
			tc.w = 1.0f;
		color.x *= tc.x, color.y *= tc.y, color.z *= tc.z, color.w *= tc.w;
	}
	if(flags & LUCID_INST_HAS_VERTEX_COLORS) {
		uint4 c = attr_c;
		float4 c0 = decodeRGBA8(c.x), c1 = decodeRGBA8(second ? c.z : c.y), c2 = decodeRGBA8(second ? c.w : c.z);
		float w0 = 1.0f - b0 - b1;
		color.x *= __fmaf_rn(w0, c0.x, __fmaf_rn(b0, c1.x, b1 * c2.x));
		color.y *= __fmaf_rn(w0, c0.y, __fmaf_rn(b0, c1.y, b1 * c2.y));
		color.z *= __fmaf_rn(w0, c0.z, __fmaf_rn(b0, c1.z, b1 * c2.z));
		color.w *= __fmaf_rn(w0, c0.w, __fmaf_rn(b0, c1.w, b1 * c2.w));
	}
	if(color.w == 0.0f)
		return 0;

	F3 normal;
	if(flags & LUCID_INST_HAS_VERTEX_NORMALS) {
		uint4 n = attr_n;
		F3 n0 = decodeNormalUint(n.x);
		F3 n1 = decodeNormalUint(second ? n.z : n.y) - n0, n2 = decodeNormalUint(second ? n.w : n.z) - n0;
		normal = mk3(__fmaf_rn(b0, n1.x, __fmaf_rn(b1, n2.x, n0.x)), __fmaf_rn(b0, n1.y, __fmaf_rn(b1, n2.y, n0.y)),
					 __fmaf_rn(b0, n1.z, __fmaf_rn(b1, n2.z, n0.z)));
	} else {
		normal = decodeNormalUint(misc.x);
	}
	return shadeFinal(globalColourTables(), lightTerms(cfg.lighting), color, normal);
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
__device__ __noinline__ void warpSortShared(u32 *keys, int n) {
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
__device__ __noinline__ void warpSortLarge(u32 *gkeys, int n, u32 *skeys) {
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
// final order -- and therefore the image -- independent of them.  Runs of equal depth are rare and
// short: the lane that finds the start of a run sorts it by insertion.
template <typename TriOf> __device__ void warpFixDepthTies(u32 *keys, int n, int slot_bits, TriOf triOf) {
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

// a block's record list
struct RecList {
	const unsigned char *base;
	int startx; // first x of the block's column inside the bin
	bool wide;	// LOW: 16-byte records
	bool lower; // LOW: take the lower half's spans
};
__device__ __forceinline__ void recAt(const RecList &l, u32 pos, u32 &mins, u32 &maxs, u32 &tri) {
	if(l.wide)
		unpackLowRecord(__ldg(reinterpret_cast<const uint4 *>(l.base) + pos), l.lower, tri, mins, maxs);
	else
		unpackHighRecord(__ldg(reinterpret_cast<const uint2 *>(l.base) + pos), tri, mins, maxs);
}

// ------------------------------------------------------------------------------------------------
// shading of one 8x4 half-block from its depth-sorted list

constexpr int CHUNK_SAMPLES = 256;
// RB_CHUNK64 (experiment build, off): a chunk takes up to 64 list entries instead of 32, as two
// sub-chunks that share one shading pass and one reduce pass.  Counted on the checker's lists
// (profiles/r1k_item_statistics.txt) that raises the busy lanes of the shading rounds from 66 % to 77 %
// and of the reduce loop from 46 % to 54 % on the 1M-triangle scene; measured, it is bit-exact and 10 %
// slower there (0.138 vs 0.126 ms): the second sub-chunk's scans and spills cost more than they return.
#ifdef RB_CHUNK64
constexpr int CHUNK_ENTRIES = 64;
#else
constexpr int CHUNK_ENTRIES = 32;
#endif

struct WarpScratch {
	float4 *stage;	 // CHUNK_ENTRIES: depth plane + constant colour of the chunk's triangles
	uint2 *results;	 // CHUNK_SAMPLES: (colour, depth) per sample
	uint2 *chunk;	 // CHUNK_ENTRIES: (pixel mask, first sample) of the chunk's triangles
	u32 *samples;	 // SAMPLE_BUF: pixel | tri << 8
	u32 *mask;		 // 32 (segment path)
	u32 *keys;		 // CAP
};
constexpr int WARP_SCRATCH_FIXED = CHUNK_ENTRIES * 16 + CHUNK_SAMPLES * 8 + CHUNK_ENTRIES * 8 + SAMPLE_BUF * 4 + 32 * 4;
__device__ __forceinline__ WarpScratch warpScratch(unsigned char *base) {
	WarpScratch ws;
	ws.stage = reinterpret_cast<float4 *>(base);
	ws.results = reinterpret_cast<uint2 *>(base + CHUNK_ENTRIES * 16);
	ws.chunk = ws.results + CHUNK_SAMPLES;
	ws.samples = reinterpret_cast<u32 *>(ws.chunk + CHUNK_ENTRIES);
	ws.mask = ws.samples + SAMPLE_BUF;
	ws.keys = ws.mask + 32;
	return ws;
}

__device__ __forceinline__ void writePixel(const Params &p, const LucidConfig &cfg, Reducer &red, int hb_x, int hb_y,
										   u32 px_frags, bool additive, bool vis_errors) {
	const int lane = laneId();
	// finishReduceSamples (shading.glsl:297-314)
	if(red.c2 != 0)
		reducerBlend(red, red.c2, additive);
	if(red.c1 != 0)
		reducerBlend(red, red.c1, additive);
	if(red.c0 != 0)
		reducerBlend(red, red.c0, additive);
	float fr = saturatef(__fmaf_rn(red.trans, cfg.background_color.x, red.r));
	float fg = saturatef(__fmaf_rn(red.trans, cfg.background_color.y, red.g));
	float fb = saturatef(__fmaf_rn(red.trans, cfg.background_color.z, red.b));
	int gx = hb_x + (lane & 7), gy = hb_y + (lane >> 3);
	if(gx < p.width && gy < p.height) {
		// rgba8 unorm store: round to nearest
		u32 out = f2u(fr * 255.0f + 0.5f) | (f2u(fg * 255.0f + 0.5f) << 8) | (f2u(fb * 255.0f + 0.5f) << 16) |
				  0xff000000u;
		p.image[(size_t)gy * p.image_pitch + gx] = out;
		if(p.frag_counts)
			p.frag_counts[(size_t)gy * p.width + gx] = px_frags;
	}
	if(vis_errors) {
		u32 inv = red.invalid;
#pragma unroll
		for(int o = 16; o > 0; o >>= 1)
			inv += __shfl_xor_sync(0xffffffffu, inv, o);
		if(lane == 0 && inv)
			atomicAdd(&p.info->stats[2], inv);
	}
}

// Lane = pixel.  The sorted list is consumed in chunks of up to 32 triangles (at most 256
// samples).  A 32x32 bit transpose of the chunk's pixel masks tells every pixel lane which of the
// chunk's triangles cover it, in list order; the lane then walks its own samples through the
// 3-entry window.  Triangles with a per-triangle constant colour are shaded by the pixel lane
// itself (depth plane from shared memory); other chunks shade all samples first, one sample per
// lane, and the pixel lanes pick the results up.  The reference's per-segment saturate
// (raster.glsl:394-395) cannot change the stored pixel -- the accumulators never decrease and the
// final value is saturated -- so segments only matter for the alpha-threshold early out, which
// keeps the segment-accurate path below.  Once every pixel of the half-block has zero
// transmittance, later samples add exactly +0 and are skipped.
// Per entry the loop needs the record (spans, triangle) and the aux word the key pass left behind
// (depth plane, constant colour): both addresses come from the sorted key alone, so the loads of
// the next chunk are issued before the current one is shaded.
struct ChunkEntry {
	u32 mins, maxs, tri;
	uint4 aux; // depth plane xyz, w = constant colour or AUX_VARYING
};
constexpr u32 AUX_VARYING = 0x00ffffffu; // alpha 0 with colour bits set: never produced by shadeConstant

__device__ __forceinline__ ChunkEntry loadEntry(const RecList &list, const uint4 *aux, const u32 *keys, int i, int count,
												u32 pos_mask) {
	ChunkEntry e;
	e.mins = 0xfffffu, e.maxs = 0, e.tri = 0; // empty spans (31 > 0 in every row)
	e.aux = make_uint4(0, 0, 0, 0);
	if(i < count) {
		u32 pos = keys[i] & pos_mask;
		recAt(list, pos, e.mins, e.maxs, e.tri);
		if(aux)
			e.aux = __ldcg(aux + pos);
	}
	return e;
}

__device__ __forceinline__ void shadeHalfBlock(const Params &p, const LucidConfig &cfg, const WarpScratch &ws,
											   const u32 *keys, const uint4 *aux, int count, u32 pos_mask, int hb_x,
											   int hb_y, const RecList &list) {
	const int lane = laneId();
	const bool additive = (p.opts & LUCID_OPT_ADDITIVE_BLENDING) != 0;
	const bool vis_errors = (p.opts & LUCID_OPT_VISUALIZE_ERRORS) != 0;
	const float fpx = float(hb_x + (lane & 7)), fpy = float(hb_y + (lane >> 3));
	Reducer red;
	reducerInit(red);
	u32 px_frags = 0;
	bool dead = false;

	ChunkEntry ahead = loadEntry(list, aux, keys, lane, count, pos_mask);
	int ahead_at = 0; // list index the `ahead` entries were loaded for
	for(int next = 0; next < count;) {
		if(ahead_at != next)
			ahead = loadEntry(list, aux, keys, next + lane, count, pos_mask);
		const ChunkEntry cur = ahead;
		const int i = next + lane;
		int nf;
		u32 bits = rowsToBits(cur.mins, cur.maxs, list.startx, nf);
		int incl = nf;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, incl, o);
			if(lane >= o)
				incl += t;
		}
		const bool in_chunk = i < count && incl <= CHUNK_SAMPLES;
		const int taken = __popc(__ballot_sync(0xffffffffu, in_chunk)); // a prefix, never empty
		const int total = __shfl_sync(0xffffffffu, incl, taken - 1);
		const int off = incl - nf;
		if(!in_chunk)
			bits = 0;
		next += taken;
		ahead_at = next;
		ahead = loadEntry(list, aux, keys, next + lane, count, pos_mask);

		u32 tm = transpose32(bits);
		px_frags += __popc(tm);
		if(dead)
			continue;
		// a pixel whose transmittance has reached zero takes exactly +0 from every later sample:
		// its samples are neither shaded nor reduced (the opaque cull of shading.glsl:31-32, exact part)
		const u32 dead_px = vis_errors ? 0u : __ballot_sync(0xffffffffu, red.trans == 0.0f);
		if((dead_px >> lane) & 1u)
			tm = 0;

		if(__all_sync(0xffffffffu, !in_chunk || cur.aux.w != AUX_VARYING)) {
			ws.stage[lane] = make_float4(__uint_as_float(cur.aux.x), __uint_as_float(cur.aux.y),
										 __uint_as_float(cur.aux.z), __uint_as_float(cur.aux.w));
			__syncwarp();
			while(tm) {
				int j = __ffs(tm) - 1;
				tm &= tm - 1;
				float4 s = ws.stage[j];
				float depth = s.x * fpx + (s.y * fpy + s.z);
				reducerPush(red, __float_as_uint(s.w), depth, additive, vis_errors);
			}
		} else {
			// samples of live pixels only: offsets are recomputed over the masked pixel sets
			u32 live = bits & ~dead_px;
			int live_off = off, live_total = total;
			if(dead_px != 0) {
				const int nl = __popc(live);
				int li = nl;
#pragma unroll
				for(int o = 1; o < 32; o <<= 1) {
					int t = __shfl_up_sync(0xffffffffu, li, o);
					if(lane >= o)
						li += t;
				}
				live_off = li - nl;
				live_total = __shfl_sync(0xffffffffu, li, 31);
			}
			if(in_chunk) {
				ws.chunk[lane] = make_uint2(live, (u32)live_off);
				u32 dst = (u32)live_off, word = cur.tri << 8, b = live;
				while(b) {
					u32 pid = __ffs(b) - 1;
					b &= b - 1;
					ws.samples[dst++] = pid | word;
				}
			}
			__syncwarp();
			for(int r0 = 0; r0 < live_total; r0 += 32) {
				int idx = r0 + lane;
				if(idx < live_total) {
					u32 val = ws.samples[idx], pid = val & 31u;
					float depth;
					u32 color = shadeSample(p, cfg, hb_x + (int)(pid & 7), hb_y + (int)(pid >> 3), val >> 8, depth);
					ws.results[idx] = make_uint2(color, __float_as_uint(depth));
				}
			}
			__syncwarp();
			while(tm) {
				int j = __ffs(tm) - 1;
				tm &= tm - 1;
				uint2 c = ws.chunk[j];
				uint2 res = ws.results[c.y + __popc(c.x & laneMaskLt())];
				reducerPush(red, res.x, __uint_as_float(res.y), additive, vis_errors);
			}
		}
		__syncwarp();
		if(!vis_errors && __all_sync(0xffffffffu, red.trans == 0.0f)) {
			dead = true;
			if(!p.frag_counts)
				break;
		}
	}
	writePixel(p, cfg, red, hb_x, hb_y, px_frags, additive, vis_errors);
}

#ifdef RB_CHUNK64
// shadeHalfBlock with chunks of up to 64 entries.  Sub-chunk A is taken exactly as in shadeHalfBlock;
// when all 32 of its entries fit, sub-chunk B continues the same sample budget (256 per chunk).  A
// constant-colour chunk only takes a B that is constant-colour too, so a chunk stays on one path.
// The per-pixel order of pushes is the list order in both variants: the pixels are identical.
__device__ __forceinline__ void shadeHalfBlock64(const Params &p, const LucidConfig &cfg, const WarpScratch &ws,
												 const u32 *keys, const uint4 *aux, int count, u32 pos_mask, int hb_x,
												 int hb_y, const RecList &list) {
	const int lane = laneId();
	const bool additive = (p.opts & LUCID_OPT_ADDITIVE_BLENDING) != 0;
	const bool vis_errors = (p.opts & LUCID_OPT_VISUALIZE_ERRORS) != 0;
	const float fpx = float(hb_x + (lane & 7)), fpy = float(hb_y + (lane >> 3));
	Reducer red;
	reducerInit(red);
	u32 px_frags = 0;
	bool dead = false;

	auto inclusiveScan = [&](int v) {
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, v, o);
			if(lane >= o)
				v += t;
		}
		return v;
	};

	ChunkEntry ahead = loadEntry(list, aux, keys, lane, count, pos_mask);
	for(int next = 0; next < count;) {
		// ---- sub-chunk A
		const ChunkEntry cur = ahead;
		int nf;
		u32 bits_a = rowsToBits(cur.mins, cur.maxs, list.startx, nf);
		const int incl_a = inclusiveScan(nf);
		const bool in_a = next + lane < count && incl_a <= CHUNK_SAMPLES;
		const int taken_a = __popc(__ballot_sync(0xffffffffu, in_a)); // a prefix, never empty
		int total = __shfl_sync(0xffffffffu, incl_a, taken_a - 1);
		if(!in_a)
			bits_a = 0;
		const bool const_a = __all_sync(0xffffffffu, !in_a || cur.aux.w != AUX_VARYING);
		next += taken_a;
		ahead = loadEntry(list, aux, keys, next + lane, count, pos_mask);
		u32 tm_a = transpose32(bits_a);
		px_frags += __popc(tm_a);
		const u32 dead_px = (vis_errors || dead) ? 0u : __ballot_sync(0xffffffffu, red.trans == 0.0f);

		// what the paths need of A goes to shared memory now, so B can reuse the registers
		int live_total = 0;
		if(!dead) {
			if(const_a) {
				ws.stage[lane] = make_float4(__uint_as_float(cur.aux.x), __uint_as_float(cur.aux.y),
											 __uint_as_float(cur.aux.z), __uint_as_float(cur.aux.w));
			} else {
				const u32 live = bits_a & ~dead_px;
				const int nl = __popc(live), li = inclusiveScan(nl);
				live_total = __shfl_sync(0xffffffffu, li, 31);
				if(in_a) {
					ws.chunk[lane] = make_uint2(live, (u32)(li - nl));
					u32 dst = (u32)(li - nl), word = cur.tri << 8, b = live;
					while(b) {
						u32 pid = __ffs(b) - 1;
						b &= b - 1;
						ws.samples[dst++] = pid | word;
					}
				}
			}
		}

		// ---- sub-chunk B: only after a full A, within the same 256 samples
		u32 tm_b = 0;
		if(taken_a == 32 && next < count) {
			const ChunkEntry cb = ahead;
			int nf_b;
			u32 bits_b = rowsToBits(cb.mins, cb.maxs, list.startx, nf_b);
			const int incl_b = total + inclusiveScan(nf_b);
			bool in_b = next + lane < count && incl_b <= CHUNK_SAMPLES;
			const bool const_b = __all_sync(0xffffffffu, !in_b || cb.aux.w != AUX_VARYING);
			if(const_a && !const_b)
				in_b = false; // a constant-colour chunk stays constant-colour; B starts the next chunk
			const int taken_b = __popc(__ballot_sync(0xffffffffu, in_b)); // a prefix, possibly empty
			if(taken_b > 0) {
				total = __shfl_sync(0xffffffffu, incl_b, taken_b - 1);
				if(!in_b)
					bits_b = 0;
				next += taken_b;
				ahead = loadEntry(list, aux, keys, next + lane, count, pos_mask);
				tm_b = transpose32(bits_b);
				px_frags += __popc(tm_b);
				if(!dead) {
					if(const_a) {
						ws.stage[32 + lane] = make_float4(__uint_as_float(cb.aux.x), __uint_as_float(cb.aux.y),
														  __uint_as_float(cb.aux.z), __uint_as_float(cb.aux.w));
					} else {
						const u32 live = bits_b & ~dead_px;
						const int nl = __popc(live), li = live_total + inclusiveScan(nl);
						if(in_b) {
							ws.chunk[32 + lane] = make_uint2(live, (u32)(li - nl));
							u32 dst = (u32)(li - nl), word = cb.tri << 8, b = live;
							while(b) {
								u32 pid = __ffs(b) - 1;
								b &= b - 1;
								ws.samples[dst++] = pid | word;
							}
						}
						live_total = __shfl_sync(0xffffffffu, li, 31);
					}
				}
			}
		}
		if(dead)
			continue;
		if((dead_px >> lane) & 1u)
			tm_a = 0, tm_b = 0;
		__syncwarp();

		if(const_a) {
#pragma unroll
			for(int half = 0; half < 2; half++) {
				u32 tm = half == 0 ? tm_a : tm_b;
				while(tm) {
					int j = __ffs(tm) - 1;
					tm &= tm - 1;
					float4 s = ws.stage[half * 32 + j];
					float depth = s.x * fpx + (s.y * fpy + s.z);
					reducerPush(red, __float_as_uint(s.w), depth, additive, vis_errors);
				}
			}
		} else {
			for(int r0 = 0; r0 < live_total; r0 += 32) {
				int idx = r0 + lane;
				if(idx < live_total) {
					u32 val = ws.samples[idx], pid = val & 31u;
					float depth;
					u32 color = shadeSample(p, cfg, hb_x + (int)(pid & 7), hb_y + (int)(pid >> 3), val >> 8, depth);
					ws.results[idx] = make_uint2(color, __float_as_uint(depth));
				}
			}
			__syncwarp();
#pragma unroll
			for(int half = 0; half < 2; half++) {
				u32 tm = half == 0 ? tm_a : tm_b;
				while(tm) {
					int j = __ffs(tm) - 1;
					tm &= tm - 1;
					uint2 c = ws.chunk[half * 32 + j];
					uint2 res = ws.results[c.y + __popc(c.x & laneMaskLt())];
					reducerPush(red, res.x, __uint_as_float(res.y), additive, vis_errors);
				}
			}
		}
		__syncwarp();
		if(!vis_errors && __all_sync(0xffffffffu, red.trans == 0.0f)) {
			dead = true;
			if(!p.frag_counts)
				break;
		}
	}
	writePixel(p, cfg, red, hb_x, hb_y, px_frags, additive, vis_errors);
}
#endif

// Segment-accurate variant (raster.glsl:292-396) for ALPHA_THRESHOLD: samples are expanded and
// consumed in the reference's 256-sample segments, so the early-out decisions fall on the same
// sample boundaries.
__device__ __noinline__ void shadeHalfBlockSegments(const Params &p, const LucidConfig &cfg, const WarpScratch &ws,
													const u32 *keys, int count, u32 pos_mask, int hb_x, int hb_y,
													const RecList &list) {
	const int lane = laneId();
	Reducer red;
	reducerInit(red);
	u32 px_frags = 0;
	int next = 0;		// next list entry to expand
	u32 seg_start = 0;	// sample offset of the current segment
	u32 off = 0;		// sample offset of entry `next`
	int carried = 0;	// samples spilled past the previous segment (< 32)
	bool stop = false;

	while(next < count || carried > 0) {
		// move the spill of the previous segment to the front (raster.glsl:302-304)
		if(carried > 0) {
			u32 v = ws.samples[SEGMENT_SIZE + lane];
			__syncwarp();
			ws.samples[lane] = v;
		}
		__syncwarp();
		const u32 seg_end = seg_start + SEGMENT_SIZE;
		while(next < count && off < seg_end) {
			int i = next + lane;
			ChunkEntry e = loadEntry(list, nullptr, keys, i, count, pos_mask);
			const u32 tri_idx = e.tri;
			int nf;
			u32 bits = rowsToBits(e.mins, e.maxs, list.startx, nf);
			int incl = nf;
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, incl, o);
				if(lane >= o)
					incl += t;
			}
			u32 my_off = off + (u32)(incl - nf);
			bool in_seg = i < count && my_off < seg_end;
			int taken = __popc(__ballot_sync(0xffffffffu, in_seg)); // a prefix of the lanes
			if(in_seg) {
				u32 dst = my_off - seg_start, word = tri_idx << 8;
				while(bits) {
					u32 pid = __ffs(bits) - 1;
					bits &= bits - 1;
					ws.samples[dst++] = pid | word;
				}
			}
			u32 consumed = __shfl_sync(0xffffffffu, (u32)incl, max(taken - 1, 0));
			if(taken > 0)
				off += consumed;
			next += taken;
			if(taken < 32)
				break;
		}
		__syncwarp();
		u32 avail = off - seg_start; // samples buffered for this segment (may exceed 256 by < 32)
		int nseg = (int)min(avail, (u32)SEGMENT_SIZE);
		carried = (int)(avail - (u32)nseg);

		for(int r0 = 0; r0 < nseg; r0 += 32) {
			int idx = r0 + lane;
			bool active = idx < nseg;
			u32 val = active ? ws.samples[idx] : 0u;
			ws.mask[lane] = 0;
			__syncwarp();
			u32 color = 0;
			float depth = 0.0f;
			if(active && !stop) {
				u32 pid = val & 31u;
				color = shadeSample(p, cfg, hb_x + (int)(pid & 7), hb_y + (int)(pid >> 3), val >> 8, depth);
			}
			if(active)
				atomicOr(&ws.mask[val & 31u], 1u << lane);
			__syncwarp();
			u32 pm = ws.mask[lane];
			px_frags += __popc(pm);
			if(stop)
				pm = 0;
			while(__any_sync(0xffffffffu, pm != 0)) {
				int bit = pm ? __ffs(pm) - 1 : 0;
				u32 c = __shfl_sync(0xffffffffu, color, bit);
				float d = __shfl_sync(0xffffffffu, depth, bit);
				if(pm) {
					pm &= pm - 1;
					reducerPush(red, c, d, false, false);
				}
			}
			__syncwarp();
		}
		red.r = saturatef(red.r), red.g = saturatef(red.g), red.b = saturatef(red.b);
		if(nseg == SEGMENT_SIZE && __all_sync(0xffffffffu, red.trans < (1.0f / 128.0f)))
			stop = true;
		seg_start += SEGMENT_SIZE;
	}
	writePixel(p, cfg, red, hb_x, hb_y, px_frags, false, false);
}

__device__ __forceinline__ void shadeHalfBlockAny(const Params &p, const LucidConfig &cfg, const WarpScratch &ws,
												 const u32 *keys, const uint4 *aux, int count, u32 pos_mask, int hb_x,
												 int hb_y, const RecList &list) {
	const bool alpha_thr = (p.opts & (LUCID_OPT_ALPHA_THRESHOLD | LUCID_OPT_ADDITIVE_BLENDING |
									   LUCID_OPT_VISUALIZE_ERRORS)) == LUCID_OPT_ALPHA_THRESHOLD;
	if(alpha_thr)
		shadeHalfBlockSegments(p, cfg, ws, keys, count, pos_mask, hb_x, hb_y, list);
	else
#ifdef RB_CHUNK64
		shadeHalfBlock64(p, cfg, ws, keys, aux, count, pos_mask, hb_x, hb_y, list);
#else
		shadeHalfBlock(p, cfg, ws, keys, aux, count, pos_mask, hb_x, hb_y, list);
#endif
}

// ------------------------------------------------------------------------------------------------
// bins

// triangle of entry t of the bin's sequence T: quads list (two triangles per quad) then tris list
__device__ __forceinline__ bool binTriangle(const Params &p, int t, int n_q, int q_off, int t_off, u32 &tri_idx) {
	if(t < n_q * 2) {
		u32 w = __ldg(p.bin_quads + q_off + (t >> 1));
		tri_idx = (w & 0x0fffffffu) * 2 + (t & 1);
		return ((w >> (30 + (t & 1))) & 1) == 0;
	}
	tri_idx = __ldg(p.bin_tris + t_off + (t - n_q * 2));
	return true;
}

struct BinInfo {
	int bin_id, n_q, q_off, n_t, t_off, n_T, pos_x, pos_y;
};
__device__ __forceinline__ BinInfo loadBin(const Params &p, int bin_id) {
	BinInfo b;
	b.bin_id = bin_id;
	b.n_q = cntc(p, LUCID_CNT_QUAD_COUNTS)[bin_id], b.q_off = cntc(p, LUCID_CNT_QUAD_OFFSETS)[bin_id];
	b.n_t = cntc(p, LUCID_CNT_TRI_COUNTS)[bin_id], b.t_off = cntc(p, LUCID_CNT_TRI_OFFSETS)[bin_id];
	b.n_T = b.n_q * 2 + b.n_t;
	int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
	b.pos_x = bin_x * BIN_SIZE, b.pos_y = bin_y * BIN_SIZE;
	return b;
}

// Phase A: every triangle of the bin is walked once over the 4-row (HIGH) or 8-row (LOW) groups
// its y range touches.  The scanline state is advanced row by row from the triangle's first group,
// exactly like the reference's incremental loop (raster_low.glsl:39-64, raster_high.glsl:54-90), so
// the spans truncated from it are bit-identical.  Walking is cheap but its trip count differs per
// triangle, so each warp only *queues* (triangle, group, scan state) items while walking and
// evaluates the spans (rasterBinStep, the expensive part) 32 queued items at a time with every
// lane busy.  emit(active, tri, group, mins0, maxs0, mins1, maxs1, bx) is called by all lanes.
constexpr int PHASE_A_RING = 64; // items; 32 bytes each, in the warp's (idle) phase-B scratch

template <bool HIGH, typename Emit>
__device__ __forceinline__ void binPhaseA(const Params &p, const BinInfo &b, uint4 *ring, Emit emit) {
	constexpr int shift = HIGH ? 2 : 3, rows_per_group = HIGH ? 4 : 8;
	const int lane = laneId(), warp = threadIdx.x >> 5;
	int q_head = 0, q_tail = 0; // warp-uniform
	auto drain = [&](int index, bool active) {
		uint4 a = ring[(index & (PHASE_A_RING - 1)) * 2], s = ring[(index & (PHASE_A_RING - 1)) * 2 + 1];
		RowScan rs;
		rs.scan[0] = __uint_as_float(a.x), rs.scan[1] = __uint_as_float(a.y), rs.scan[2] = __uint_as_float(a.z);
		rs.step[0] = __uint_as_float(s.x), rs.step[1] = __uint_as_float(s.y), rs.step[2] = __uint_as_float(s.z);
		rs.xneg = (a.w >> 27) & 7u;
		u32 mn0, mx0, bx0, mn1 = 0, mx1 = 0, bx1 = 0;
		rasterBinStep(rs, mn0, mx0, bx0);
		if(!HIGH)
			rasterBinStep(rs, mn1, mx1, bx1);
		emit(active, a.w & 0xffffffu, (int)((a.w >> 24) & 7u), mn0, mx0, mn1, mx1, active ? (bx0 | bx1) : 0u);
	};
	// the triangle's list word and scanline record are two dependent loads: those of the warp's next
	// 32 triangles are issued before the current ones are walked
	struct Fetched {
		u32 tri_idx;
		bool ok;
		uint4 s0, s1;
	};
	auto fetch = [&](int base) {
		Fetched f;
		const int t = base + lane;
		f.tri_idx = 0;
		f.ok = t < b.n_T && binTriangle(p, t, b.n_q, b.q_off, b.t_off, f.tri_idx);
		f.s0 = f.s1 = make_uint4(0, 0, 0, 0);
		if(f.ok) {
			const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_scan + f.tri_idx);
			f.s0 = __ldg(src), f.s1 = __ldg(src + 1);
		}
		return f;
	};
	Fetched ahead = fetch(warp * 32);
	for(int base = warp * 32; base < b.n_T; base += RASTER_THREADS) {
		const Fetched cur = ahead;
		if(base + RASTER_THREADS < b.n_T)
			ahead = fetch(base + RASTER_THREADS);
		const u32 tri_idx = cur.tri_idx;
		const bool ok = cur.ok;
		int n_g = 0, min_g = 0;
		float scan0 = 0, scan1 = 0, scan2 = 0, step0 = 0, step1 = 0, step2 = 0;
		u32 xneg = 0;
		if(ok) {
			const uint4 s0 = cur.s0, s1 = cur.s1;
			int ymin = (int)(s0.w & 0xffff) - b.pos_y, ymax = (int)(s0.w >> 16) - b.pos_y;
			min_g = min(max(ymin, 0), BIN_SIZE - 1) >> shift;
			n_g = (min(max(ymax, 0), BIN_SIZE - 1) >> shift) - min_g + 1;
			step0 = __uint_as_float(s1.x), step1 = __uint_as_float(s1.y), step2 = __uint_as_float(s1.z);
			xneg = s1.w & 7u;
			float start_x = float(b.pos_x), start_y = float(b.pos_y + min_g * rows_per_group);
			scan0 = __uint_as_float(s0.x) + (step0 * start_y - start_x);
			scan1 = __uint_as_float(s0.y) + (step1 * start_y - start_x);
			scan2 = __uint_as_float(s0.z) + (step2 * start_y - start_x);
		}
		const int max_ng = __reduce_max_sync(0xffffffffu, n_g);
		for(int k = 0; k < max_ng; k++) {
			const bool has = k < n_g;
			const u32 m = __ballot_sync(0xffffffffu, has);
			if(has) {
				int pos = (q_tail + __popc(m & laneMaskLt())) & (PHASE_A_RING - 1);
				ring[pos * 2] = make_uint4(__float_as_uint(scan0), __float_as_uint(scan1), __float_as_uint(scan2),
										   tri_idx | ((u32)(min_g + k) << 24) | (xneg << 27));
				ring[pos * 2 + 1] = make_uint4(__float_as_uint(step0), __float_as_uint(step1), __float_as_uint(step2), 0u);
				if(k + 1 < n_g) { // the state at the triangle's next group: one addition per pixel row
#pragma unroll
					for(int r = 0; r < rows_per_group; r++)
						scan0 += step0, scan1 += step1, scan2 += step2;
				}
			}
			q_tail += __popc(m);
			__syncwarp();
			if(q_tail - q_head >= 32) {
				drain(q_head + lane, true);
				q_head += 32;
				__syncwarp();
			}
		}
	}
	if(q_tail > q_head)
		drain(q_head + lane, lane < q_tail - q_head);
	__syncwarp();
}

// ------------------------------------------------------------------------------------------------
// stage 1: k_raster_bins -- block lists of every non-empty bin (generateRowTris + generateBlocks /
// computeRBlockGroups of the reference, raster_low.glsl:39-106, raster_high.glsl:54-144)

constexpr int HB_LIST_CAP = MAX_HBLOCK_TRIS; // records per half-block list (HIGH)
// work items of stage 2 are queued by size class (entries of the list), heaviest class first
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

__device__ __forceinline__ unsigned char *binLists(const Params &p, int bin_id) {
	return reinterpret_cast<unsigned char *>(p.block_lists) + (size_t)bin_id * BIN_LIST_BYTES;
}

struct BinShared {
	int count[32]; // entries per half-block (HIGH) / block (LOW)
	int holes[32]; // HIGH: columns inside a record's [first, last] range without coverage
	int bin_index, status;
};

// a persistent 256-thread CTA takes one bin at a time: HIGH bins first (they take longest)
__global__ void __launch_bounds__(RASTER_THREADS) k_raster_bins(const __grid_constant__ Params p, u32 background) {
	__shared__ BinShared sh;
	__shared__ __align__(16) uint4 s_ring[RASTER_WARPS][PHASE_A_RING * 2];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdlEntry();
	if(p.info->temp[1] != 0)
		return; // the bin lists did not fit their buffers (k_bin_scan): no list is valid, the frame is painted red
	const int n_high = p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH];
	const int n_low = p.info->bin_level_counts[LUCID_BIN_LEVEL_LOW];
	while(true) {
		__syncthreads();
		if(tid == 0)
			sh.bin_index = (int)atomicAdd(&p.work_counters[0], 1u);
		__syncthreads();
		const int idx = sh.bin_index;
		if(idx >= n_high + n_low)
			break;
		const long long t_bin = clock64();
		bool high = idx < n_high;
		const int bin_id = high ? cntc(p, LUCID_CNT_HIGH_BINS)[idx] : cntc(p, LUCID_CNT_LOW_BINS)[idx - n_high];
		const BinInfo b = loadBin(p, bin_id);
		unsigned char *lists = binLists(p, bin_id);

		if(!high) {
			if(tid < 32)
				sh.count[tid] = 0;
			if(tid == 0)
				sh.status = 0;
			__syncthreads();
			uint4 *recs = reinterpret_cast<uint4 *>(lists);
			binPhaseA<false>(p, b, s_ring[warp], [&](bool, u32 tri_idx, int g, u32 mn0, u32 mx0, u32 mn1, u32 mx1, u32 bx) {
				while(bx) {
					int c = __ffs(bx) - 1;
					bx &= bx - 1;
					int slot = atomicAdd(&sh.count[g * 4 + c], 1);
					if(slot < MAX_BLOCK_TRIS)
						recs[(g * 4 + c) * MAX_BLOCK_TRIS + slot] =
							packLowRecord(tri_idx, mn0, mx0, mn1, mx1);
				}
			});
			__syncthreads();
			if(tid < 16 && sh.count[tid] > MAX_BLOCK_TRIS)
				sh.status = 1;
			__syncthreads();
			if(sh.status != 0) {
				// too many triangles for one block: the bin is redone by the HIGH path
				// (raster_low.glsl:101-105,230-237); promoteBins (k_raster_blocks) appends it to the HIGH list
				if(tid == 0)
					p.bin_flags[bin_id] |= 1u;
				high = true;
			}
			__syncthreads();
		}
		if(high) {
			if(tid < 32)
				sh.count[tid] = 0, sh.holes[tid] = 0;
			if(tid == 0)
				sh.status = 0;
			__syncthreads();
			uint2 *recs = reinterpret_cast<uint2 *>(lists);
			binPhaseA<true>(p, b, s_ring[warp], [&](bool, u32 tri_idx, int g, u32 mn, u32 mx, u32, u32, u32 bx) {
				if(bx == 0)
					return;
				const int lo = __ffs(bx) - 1, hi = 31 - __clz(bx);
				u32 holes = ((2u << hi) - (1u << lo)) & ~bx;
				while(bx) {
					int c = __ffs(bx) - 1;
					bx &= bx - 1;
					int slot = atomicAdd(&sh.count[g * 4 + c], 1);
					if(slot < HB_LIST_CAP)
						recs[(g * 4 + c) * HB_LIST_CAP + slot] = packHighRecord(tri_idx, mn, mx);
				}
				while(holes) {
					int c = __ffs(holes) - 1;
					holes &= holes - 1;
					atomicAdd(&sh.holes[g * 4 + c], 1);
				}
			});
			__syncthreads();
			if(tid < 32) {
				// the reference's limit is on the estimated count (first..last column, holes
				// included): more than 4096 paints the bin red (raster_high.glsl:80-83,140-141).
				// 16384 records in one half-block row imply more than 4096 in one of its
				// half-blocks, so that limit is covered too.
				bool over = __any_sync(0xffffffffu, sh.count[tid] + sh.holes[tid] > MAX_HBLOCK_TRIS);
				if(tid == 0 && over) {
					p.bin_flags[bin_id] |= 2u;
					sh.status = 2;
				}
			}
			__syncthreads();
			if(sh.status != 0)
				continue; // finishBins (k_raster_blocks) paints the bin
		}

		// publish the non-empty blocks as work items of stage 2; empty ones only get the background
		const int n_blocks = high ? 32 : 16;
		if(tid < 32) {
			const int c = tid < n_blocks ? sh.count[tid] : 0;
			if(tid < n_blocks)
				p.block_counts[bin_id * 32 + tid] = c;
			// one queue per size class, consumed from the heaviest class down: the kernel ends with the
			// shortest items, so its tail is a few microseconds instead of one long list
			const u32 item = ((u32)bin_id << 6) | (high ? 32u : 0u) | (u32)tid;
			const int cls = itemClass(c);
#pragma unroll
			for(int k = 0; k < ITEM_CLASSES; k++) {
				const u32 m = __ballot_sync(0xffffffffu, c > 0 && cls == k);
				if(m == 0)
					continue;
				u32 base = 0;
				if(tid == __ffs(m) - 1)
					base = atomicAdd(&p.work_counters[3 + k], (u32)__popc(m));
				base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
				if((m >> tid) & 1)
					p.block_items[(size_t)k * p.block_items_cap + base + __popc(m & laneMaskLt())] = make_uint2(item, (u32)c);
			}
		}
		if(tid == 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(p.bin_cost) + bin_id,
					  (unsigned long long)(clock64() - t_bin) * RASTER_WARPS);
		const int rows = high ? 4 : 8;
		for(int blk = warp; blk < n_blocks; blk += RASTER_WARPS) {
			if(sh.count[blk] != 0)
				continue;
			const int bx0 = b.pos_x + (blk & 3) * 8, by0 = b.pos_y + (blk >> 2) * rows;
			for(int i = lane; i < 8 * rows; i += 32) {
				int gx = bx0 + (i & 7), gy = by0 + (i >> 3);
				if(gx < p.width && gy < p.height) {
					p.image[(size_t)gy * p.image_pitch + gx] = background;
					if(p.frag_counts)
						p.frag_counts[(size_t)gy * p.width + gx] = 0;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------
// stage 2: k_raster_blocks -- depth sort and shading; a work item is one 8x8 block of a LOW bin or
// one 8x4 half-block of a HIGH bin and belongs to one warp (generateBlocks / generateRBlocks sort +
// unpackSamples + shadeAndReduceSamples, raster_low.glsl:107-281, raster_high.glsl:146-348)

#ifndef RB_MIN_CTAS
#define RB_MIN_CTAS 5
#endif
#ifndef RB_PREFETCH_MAX
#define RB_PREFETCH_MAX 0 // items up to this many entries request their successor early
#endif
#ifndef RB_KEY_UNROLL
#define RB_KEY_UNROLL 2
#endif
constexpr int BLOCK_WARPS = 4;
constexpr int KEY_UNROLL = RB_KEY_UNROLL;
constexpr int WARP_SCRATCH_BYTES = WARP_SCRATCH_FIXED + SMEM_KEYS * 4;

// ------------------------------------------------------------------------------------------------
// frame bookkeeping done by the first CTAs of k_raster_blocks before they take work items (two
// more launches on a frame of a few hundred microseconds would cost more than the work itself):
// background for empty bins (the reference leaves them to the application's clear,
// lucid_app.cpp:606-619), red for bins over the reference's limits (raster_high.glsl:313-317), and
// the level bookkeeping of promoted bins: appended to the HIGH list in bin order
// (raster_low.glsl:230-237,294-298).  Everything read here was written by k_raster_bins.
__device__ __forceinline__ void finishBins(const Params &p, u32 background, u32 *s_mask) {
	const int lane = laneId(), warp = threadIdx.x >> 5;
	// bin lists over their capacity (k_bin_scan set temp[1], dispatch and k_raster_bins did nothing): every
	// owned bin is painted red, like a bin over the reference's own limits, and the host is told
	const bool list_overflow = p.info->temp[1] != 0;
	if(list_overflow && blockIdx.x == 0 && threadIdx.x == 0 && p.host_status)
		*p.host_status = 1u;
	// 32 bins per CTA and round (one round on a B200: 740 CTAs cover 23 680 bins)
	for(int first = blockIdx.x * 32; first < p.bin_count; first += gridDim.x * 32) {
		if(warp == 0) {
			const int b = first + lane;
			u32 kind = 0; // 1 background, 2 red
			if(b < p.bin_count && ownsBin(p, b)) {
				const bool empty = cntc(p, LUCID_CNT_TRI_COUNTS)[b] + cntc(p, LUCID_CNT_QUAD_COUNTS)[b] * 2 == 0;
				kind = (list_overflow || (p.bin_flags[b] & 2u)) ? 2u : empty ? 1u : 0u;
			}
			const u32 fill = __ballot_sync(0xffffffffu, kind != 0), red = __ballot_sync(0xffffffffu, kind == 2);
			if(lane == 0)
				s_mask[0] = fill, s_mask[1] = red;
		}
		__syncthreads();
		const u32 red = s_mask[1];
		u32 fill = s_mask[0];
		__syncthreads();
		for(int n = 0; fill; n++) {
			const int j = __ffs(fill) - 1;
			fill &= fill - 1;
			if((n & (BLOCK_WARPS - 1)) != warp)
				continue;
			const int b = first + j, by = b / p.bin_count_x, bx = b - by * p.bin_count_x;
			const u32 value = ((red >> j) & 1u) ? 0x000000ffu : background;
			const int gx = bx * BIN_SIZE + lane;
			if(gx < p.width)
#pragma unroll 4
				for(int y = 0; y < BIN_SIZE; y++) {
					const int gy = by * BIN_SIZE + y;
					if(gy >= p.height)
						break;
					p.image[(size_t)gy * p.image_pitch + gx] = value;
					if(p.frag_counts)
						p.frag_counts[(size_t)gy * p.width + gx] = 0;
				}
		}
	}
}

__device__ __forceinline__ void promoteBins(const Params &p, int *s_warp) {
	const int threads = BLOCK_WARPS * 32;
	const int n_low = p.info->bin_level_counts[LUCID_BIN_LEVEL_LOW];
	const int *low = cntc(p, LUCID_CNT_LOW_BINS);
	int *high = p.counts + (size_t)LUCID_CNT_HIGH_BINS * p.bin_count;
	const int n_high = p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH];
	const int per = (n_low + threads - 1) / threads;
	const int i0 = min((int)threadIdx.x * per, n_low), i1 = min(i0 + per, n_low);
	int mine = 0;
	for(int i = i0; i < i1; i++)
		mine += (p.bin_flags[low[i]] & 1u) ? 1 : 0;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int incl = mine;
	for(int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, incl, o);
		if(lane >= o)
			incl += t;
	}
	if(lane == 31)
		s_warp[warp] = incl;
	__syncthreads();
	int before = 0, total = 0;
	for(int w = 0; w < BLOCK_WARPS; w++) {
		before += w < warp ? s_warp[w] : 0;
		total += s_warp[w];
	}
	if(total == 0)
		return;
	int pos = n_high + before + incl - mine;
	for(int i = i0; i < i1; i++)
		if(p.bin_flags[low[i]] & 1u)
			high[pos++] = low[i];
	if(threadIdx.x == 0) {
		const int all = n_high + total;
		p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH] = all;
		u32 nd = (u32)min(all, p.max_dispatches / 2);
		if(nd > p.info->bin_level_dispatches[LUCID_BIN_LEVEL_HIGH][0])
			p.info->bin_level_dispatches[LUCID_BIN_LEVEL_HIGH][0] = nd;
	}
}

#ifdef RB_PHASE_CLOCKS
// experiment builds only (tools/variants.py): per-phase clock sums and per-warp finish times
__device__ unsigned long long g_phase[16];
__device__ unsigned long long g_warp_end[148 * RB_MIN_CTAS * 4];
__device__ __forceinline__ unsigned long long gtimer() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
#define PHASE_MARK(k)                                                                              \
	{                                                                                              \
		long long now_ = clock64();                                                                \
		ph[k] += (unsigned long long)(now_ - t_mark);                                              \
		t_mark = now_;                                                                             \
	}
#else
#define PHASE_MARK(k)
#endif

__global__ void __launch_bounds__(BLOCK_WARPS * 32, RB_MIN_CTAS)
	k_raster_blocks(const __grid_constant__ Params p, const __grid_constant__ LucidConfig cfg, u32 background) {
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ int s_misc[BLOCK_WARPS];
	const int lane = laneId(), warp = threadIdx.x >> 5;
	pdlEntry();
	finishBins(p, background, reinterpret_cast<u32 *>(s_misc));
	if(blockIdx.x == gridDim.x - 1) {
		__syncthreads();
		promoteBins(p, s_misc);
	}
	const WarpScratch ws = warpScratch(smem + (size_t)warp * WARP_SCRATCH_BYTES);
	u32 *large_keys = p.large_keys + (size_t)(blockIdx.x * BLOCK_WARPS + warp) * MAX_HBLOCK_TRIS;
	uint4 *aux = p.block_aux + (size_t)(blockIdx.x * BLOCK_WARPS + warp) * MAX_HBLOCK_TRIS;
	u32 class_end[ITEM_CLASSES]; // exclusive end of every class in ticket order
	{
		u32 acc = 0;
#pragma unroll
		for(int k = 0; k < ITEM_CLASSES; k++)
			class_end[k] = acc += p.work_counters[3 + k];
	}
	const u32 n_items = class_end[ITEM_CLASSES - 1];
	u32 frag_acc = 0, hbt_acc = 0;
	// Work fetch: the queue index of the next item is requested when the current item's keys are
	// built and its entry when they are sorted, so both round trips overlap the sort and the
	// shading instead of being waited for; a warp never holds more than one item ahead, which keeps
	// the dynamic balance.
	auto fetchIndex = [&]() { return lane == 0 ? atomicAdd(&p.work_counters[1], 1u) : 0u; };
	auto fetchEntry = [&](u32 i) {
		uint2 e = make_uint2(0, 0);
		if(lane == 0 && i < n_items) {
			int k = 0;
			u32 first = 0;
#pragma unroll
			for(int c = 0; c < ITEM_CLASSES - 1; c++)
				if(i >= class_end[c])
					k = c + 1, first = class_end[c];
			e = __ldcg(p.block_items + (size_t)k * p.block_items_cap + (i - first));
		}
		return e;
	};
#ifdef RB_PHASE_CLOCKS
	unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	long long t_mark = clock64();
	const unsigned long long t_begin = gtimer();
#endif
	uint2 next_entry = fetchEntry(fetchIndex());
	while(true) {
		const uint2 entry = next_entry;
		const u32 item = __shfl_sync(0xffffffffu, entry.x, 0);
		const int count = (int)__shfl_sync(0xffffffffu, entry.y, 0);
		PHASE_MARK(0) // work fetch
		if(count == 0)
			break;
		const long long t_item = clock64();
#ifdef RB_PHASE_CLOCKS
		ph[5] += 1, ph[6] += (unsigned long long)count;
#endif
		const int bin_id = (int)(item >> 6), sub = (int)(item & 31u);
		const bool high = (item & 32u) != 0;
		const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
		const int pos_x = bin_x * BIN_SIZE, pos_y = bin_y * BIN_SIZE;
		const int cx8 = (sub & 3) * 8, ry = sub >> 2;
		RecList list;
		list.wide = !high, list.lower = false, list.startx = cx8;
		list.base = binLists(p, bin_id) + (high ? (size_t)sub * HB_LIST_CAP * 8 : (size_t)sub * MAX_BLOCK_TRIS * 16);
		const bool large = count > SMEM_KEYS;
		u32 *keys = large ? large_keys : ws.keys;
		u32 *tie_tris = reinterpret_cast<u32 *>(ws.stage); // SMEM_KEYS words, idle until shading starts

		// depth keys from the centroid of the covered pixels (raster.glsl:142-176); four entries
		// per lane are in flight.  The pass leaves (depth plane, constant colour) per entry in the
		// warp's aux array for the shading loop.
		auto loadRec = [&](int i) {
			uint4 r4 = make_uint4(0, 0, 0, 0);
			if(i < count) {
				if(high) {
					uint2 r = __ldg(reinterpret_cast<const uint2 *>(list.base) + i);
					r4.x = r.x, r4.y = r.y;
				} else {
					r4 = __ldg(reinterpret_cast<const uint4 *>(list.base) + i);
				}
			}
			return r4;
		};
		uint4 rec_next[KEY_UNROLL];
#pragma unroll
		for(int u = 0; u < KEY_UNROLL; u++)
			rec_next[u] = loadRec(u * 32 + lane);
		for(int i0 = 0; i0 < count; i0 += 32 * KEY_UNROLL) {
			// records are one round trip, the triangles' sectors a second, dependent one: the records of
			// the next iteration are requested together with this iteration's sectors
			uint4 rec[KEY_UNROLL], dq[KEY_UNROLL], misc[KEY_UNROLL];
#pragma unroll
			for(int u = 0; u < KEY_UNROLL; u++) {
				rec[u] = rec_next[u];
				const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_shade + (rec[u].x & 0xffffffu));
				dq[u] = __ldg(src), misc[u] = __ldg(src + 1);
			}
			if(i0 + 32 * KEY_UNROLL < count) {
#pragma unroll
				for(int u = 0; u < KEY_UNROLL; u++)
					rec_next[u] = loadRec(i0 + 32 * KEY_UNROLL + u * 32 + lane);
			}
#pragma unroll
			for(int u = 0; u < KEY_UNROLL; u++) {
				const int i = i0 + u * 32 + lane;
				if(i >= count)
					continue;
				u32 tri_idx, mins, maxs, depth;
				if(high) {
					unpackHighRecord(make_uint2(rec[u].x, rec[u].y), tri_idx, mins, maxs);
					int nf, cx, cy;
					rowsCentroid(mins, maxs, cx8, nf, cx, cy);
					float scale = __fdiv_rn(0.5f, float(nf));
					float cpx = float(cx) * scale + (float(cx8) + float(pos_x));
					float cpy = float(cy) * scale + (float(ry * 4) + float(pos_y));
					depth = blockDepth(dq[u], cpx, cpy, float(0x7fffe)) << 14;
					frag_acc += (u32)nf;
				} else {
					int nf0, cx0, cy0, nf1, cx1, cy1;
					unpackLowRecord(rec[u], false, tri_idx, mins, maxs);
					rowsCentroid(mins, maxs, cx8, nf0, cx0, cy0);
					unpackLowRecord(rec[u], true, tri_idx, mins, maxs);
					rowsCentroid(mins, maxs, cx8, nf1, cx1, cy1);
					// both halves use row offsets 1,3,5,7 for the centroid, exactly as the reference does
					float cx = float(cx0) + float(cx1), cy = float(cy0) + float(cy1);
					float scale = __fdiv_rn(0.5f, float(nf0 + nf1));
					float cpx = cx * scale + float(pos_x + cx8), cpy = cy * scale + float(pos_y + ry * 8);
					depth = blockDepth(dq[u], cpx, cpy, float(0x3ffffe)) << 10;
					frag_acc += (u32)(nf0 + nf1);
				}
				keys[i] = (u32)i | depth;
				if(!large)
					tie_tris[i] = tri_idx;
				__stcg(aux + i, make_uint4(dq[u].x, dq[u].y, dq[u].z, misc[u].w != 0 ? misc[u].z : AUX_VARYING));
			}
		}
		__syncwarp();
		PHASE_MARK(1) // key pass
		const bool light_item = count <= RB_PREFETCH_MAX;
		u32 next_index = 0;
		if(light_item)
			next_index = fetchIndex();
		// stats: LOW counts the block's triangles once per half-block (raster_low.glsl:272-275),
		// HIGH the exact half-block list (raster_high.glsl:309-310)
		hbt_acc += lane == 0 ? (u32)count * (high ? 1u : 2u) : 0u;
		const int slot_bits = high ? 14 : 10;
		if(high || count > 3) { // LOW blocks with <= 3 triangles rely on the window alone (raster_low.glsl:144)
			if(large) {
				warpSortLarge(keys, count, ws.keys);
				warpFixDepthTies(keys, count, slot_bits, [&](u32 pos) {
					return __ldg(reinterpret_cast<const uint2 *>(list.base) + pos).x & 0xffffffu;
				});
			} else {
				warpSortShared(keys, count);
				warpFixDepthTies(keys, count, slot_bits, [&](u32 pos) { return tie_tris[pos]; });
			}
		}
		if(light_item)
			next_entry = fetchEntry(next_index);
		PHASE_MARK(2) // sort + ties
		const u32 pos_mask = (1u << slot_bits) - 1u;
		const int halves = high ? 1 : 2, hb_y = pos_y + ry * (high ? 4 : 8);
		for(int half = 0; half < halves; half++) {
			list.lower = half != 0;
			shadeHalfBlockAny(p, cfg, ws, keys, aux, count, pos_mask, pos_x + cx8, hb_y + half * 4, list);
			__syncwarp();
		}
		PHASE_MARK(3) // shading
		if(lane == 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(p.bin_cost) + bin_id, (unsigned long long)(clock64() - t_item));
		if(!light_item) // a long item takes its successor only when it is done (dynamic balance)
			next_entry = fetchEntry(fetchIndex());
	}
#ifdef RB_PHASE_CLOCKS
	if(lane == 0) {
		for(int k = 0; k < 8; k++)
			atomicAdd(&g_phase[k], ph[k]);
		g_warp_end[blockIdx.x * BLOCK_WARPS + warp] = gtimer() - t_begin;
		atomicMin(&g_phase[8], t_begin);
		atomicMax(&g_phase[9], gtimer());
	}
#endif
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		frag_acc += __shfl_xor_sync(0xffffffffu, frag_acc, o);
		hbt_acc += __shfl_xor_sync(0xffffffffu, hbt_acc, o);
	}
	if(lane == 0) {
		if(frag_acc)
			atomicAdd(&p.info->stats[0], frag_acc);
		if(hbt_acc)
			atomicAdd(&p.info->stats[1], hbt_acc);
	}
}

// ------------------------------------------------------------------------------------------------
// composite of the bin-row split: the pixels of the owned bins go from this device's image to the
// gathering device's image (a peer mapping) as full 128-byte bin rows -- the raster kernels' own
// stores are 32-byte half-block rows, which make four times as many NVLink packets and kept the
// gathering device's ingress busy for 0.4 ms after an 8-GPU 4K frame
__global__ void __launch_bounds__(256) k_composite_bins(const Params p, u32 *dst, int dst_pitch) {
	pdlEntry();
	const int lane = laneId(), warp = threadIdx.x >> 5;
	for(int b = p.bin_begin + (int)blockIdx.x; b < p.bin_end; b += gridDim.x) {
		const int by = b / p.bin_count_x, bx = b - by * p.bin_count_x;
		const int gx = bx * BIN_SIZE + lane;
		if(gx >= p.width)
			continue;
		for(int y = warp; y < BIN_SIZE; y += 8) {
			const int gy = by * BIN_SIZE + y;
			if(gy < p.height)
				dst[(size_t)gy * dst_pitch + gx] = p.image[(size_t)gy * p.image_pitch + gx];
		}
	}
}
void launchCompositeBins(const Params &p, u32 *dst, int dst_pitch, cudaStream_t stream, int num_sms) {
	launchPDL(k_composite_bins, num_sms * 8, 256, 0, stream, p, dst, dst_pitch);
}

#ifdef RB_PHASE_CLOCKS
extern "C" int lucid_debug_phase_clocks(unsigned long long *dst, unsigned long long *warp_end, int reset) {
	if(reset) {
		unsigned long long z[16] = {0};
		z[8] = ~0ull;
		return (int)cudaMemcpyToSymbol(g_phase, z, sizeof(z));
	}
	cudaMemcpyFromSymbol(dst, g_phase, sizeof(g_phase));
	return (int)cudaMemcpyFromSymbol(warp_end, g_warp_end, sizeof(g_warp_end));
}
#endif

static int rasterBinsGrid(int num_sms) { return num_sms * 4; }
static int rasterBlocksGrid(int num_sms) { return num_sms * RB_MIN_CTAS; }
size_t rasterLargeKeysCount(int num_sms) { return (size_t)rasterBlocksGrid(num_sms) * BLOCK_WARPS * MAX_HBLOCK_TRIS; }

void launchRaster(const Params &p, const LucidConfig &cfg, cudaStream_t stream, cudaEvent_t *ev, int num_sms) {
	static std::once_flag configured[64]; // function attributes are per device
	const int blocks_smem = BLOCK_WARPS * WARP_SCRATCH_BYTES;
	oncePerDevice(configured, [=] { cudaFuncSetAttribute(k_raster_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, blocks_smem); });
	const LucidVec4 &bg = cfg.background_color;
	auto q = [](float v) { return (u32)(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f + 0.5f); };
	const u32 bg8 = q(bg.x) | (q(bg.y) << 8) | (q(bg.z) << 16) | 0xff000000u;
	launchPDL(k_raster_bins, rasterBinsGrid(num_sms), RASTER_THREADS, 0, stream, p, bg8);
	if(ev)
		cudaEventRecord(ev[0], stream);
	launchPDL(k_raster_blocks, rasterBlocksGrid(num_sms), BLOCK_WARPS * 32, (size_t)blocks_smem, stream, p, cfg, bg8);
	if(ev) {
		cudaEventRecord(ev[1], stream);
		cudaEventRecord(ev[2], stream); // the finish stage is part of k_raster_blocks now
	}
}

} // namespace lucid
