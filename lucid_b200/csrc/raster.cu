// raster.cu -- per-bin rasterisation, block depth sort, sample shading and exact blending.
//
// Replaces data/shaders/raster_low.glsl, raster_high.glsl, shared/raster.glsl and
// shared/shading.glsl.  Same results, different decomposition:
//   * a work item is one block row of a LOW bin (8 pixel rows, four 8x8 blocks) or one half-block
//     row of a HIGH bin (4 pixel rows, four 8x4 half-blocks); a 128-thread CTA takes one item and
//     each of its four warps owns one block column.  The reference runs a whole bin per work
//     group and bounces row lists and half-block lists through global scratch
//     (raster_low.glsl:22-32,58-63); here LOW keeps everything in shared memory and HIGH keeps
//     only its row records in an L2-resident scratch slice.
//   * sort keys break depth ties by the triangle's position in the bin's (sorted) list instead
//     of an atomic arrival slot, so the output is deterministic.
//   * samples are expanded and consumed as a stream with the reference's 256-sample segments
//     (raster.glsl:71-72,292-396), which keeps the per-segment saturation points identical.
#include "common.cuh"

namespace lucid {

constexpr int RASTER_THREADS = 128;
constexpr int RASTER_WARPS = RASTER_THREADS / 32;
constexpr int SEGMENT_SIZE = 256;
constexpr int SAMPLE_BUF = SEGMENT_SIZE + 32;
constexpr int MAX_BLOCK_TRIS = 256;		 // raster_low.glsl:17
constexpr int MAX_HBLOCK_TRIS = 4096;	 // raster_high.glsl:27
constexpr int MAX_HBLOCK_ROW_TRIS = 16384; // raster_high.glsl:30
constexpr int LOW_MAX_TRIS = 1024;

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

// raster.glsl:142-168 -- one 8-wide column of four rows: pixel mask, fragment count, centroid sums
__device__ __forceinline__ u32 halfPixelMask(u32 mins, u32 maxs, int startx, int &num_frags,
											 int &csum_x, int &csum_y) {
	u32 bits = 0;
	num_frags = 0, csum_x = 0, csum_y = 0;
#pragma unroll
	for(int r = 0; r < 4; r++) {
		int mn = max((int)((mins >> (5 * r)) & 31) - startx, 0);
		int mx = min((int)((maxs >> (5 * r)) & 31) - startx, 7);
		int c = max(mx - mn + 1, 0);
		num_frags += c;
		csum_x += (mn * 2 + c) * c;
		csum_y += (2 * r + 1) * c;
		bits |= ((1u << c) - 1u) << (mn + 8 * r);
	}
	return bits;
}

// raster.glsl:170-176
__device__ __forceinline__ u32 blockDepth(const Params &p, u32 tri_idx, float cx, float cy, float range) {
	uint4 d = __ldg(reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx));
	float ray_pos = __uint_as_float(d.x) * cx + (__uint_as_float(d.y) * cy + __uint_as_float(d.z));
	float depth = range * saturatef(rsqrt_rn(ray_pos + 1.0f));
	return f2u(depth);
}

// ------------------------------------------------------------------------------------------------
// shading (shading.glsl:64-184)

__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }

__device__ __forceinline__ float4 texelFetch(const Params &p, int slot, int level, int x, int y) {
	int w = max(1, p.tex_width[slot] >> level), h = max(1, p.tex_height[slot] >> level);
	x = ((x % w) + w) % w;
	y = ((y % h) + h) % h;
	uchar4 t = __ldg(p.tex_data[slot] + p.tex_level_offset[slot][level] + (size_t)y * w + x);
	const float s = 1.0f / 255.0f;
	return make_float4(float(t.x) * s, float(t.y) * s, float(t.z) * s, float(t.w) * s);
}
__device__ float4 bilinear(const Params &p, int slot, int level, float u, float v) {
	int w = max(1, p.tex_width[slot] >> level), h = max(1, p.tex_height[slot] >> level);
	float fx = u * float(w) - 0.5f, fy = v * float(h) - 0.5f;
	float x0f = floorf(fx), y0f = floorf(fy);
	float ax = fx - x0f, ay = fy - y0f;
	int x0 = f2i(x0f), y0 = f2i(y0f);
	float4 c00 = texelFetch(p, slot, level, x0, y0), c10 = texelFetch(p, slot, level, x0 + 1, y0);
	float4 c01 = texelFetch(p, slot, level, x0, y0 + 1), c11 = texelFetch(p, slot, level, x0 + 1, y0 + 1);
	float4 o;
#define LERP2(c)                                                                                   \
	{                                                                                              \
		float top = c00.c + (c10.c - c00.c) * ax;                                                  \
		float bot = c01.c + (c11.c - c01.c) * ax;                                                  \
		o.c = top + (bot - top) * ay;                                                              \
	}
	LERP2(x) LERP2(y) LERP2(z) LERP2(w)
#undef LERP2
	return o;
}
// Filter definition (the reference leaves this to the Vulkan sampler): repeat addressing,
// bilinear within a level, linear between the two nearest levels, isotropic lod.
__device__ float4 sampleTexture(const Params &p, int slot, float u, float v, float dudx, float dvdx,
								float dudy, float dvdy) {
	if(p.tex_data[slot] == nullptr)
		return make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	float w0 = float(p.tex_width[slot]), h0 = float(p.tex_height[slot]);
	float ax = dudx * w0, ay = dvdx * h0, bx = dudy * w0, by = dvdy * h0;
	float rho2 = fmaxf(ax * ax + ay * ay, bx * bx + by * by);
	int levels = p.tex_levels[slot];
	float lod = 0.0f;
	if(rho2 > 1.0f)
		lod = 0.5f * log2_poly(rho2);
	lod = clampf(lod, 0.0f, float(levels - 1));
	float l0f = floorf(lod);
	int l0 = f2i(l0f), l1 = min(l0 + 1, levels - 1);
	float a = lod - l0f;
	float4 c0 = bilinear(p, slot, l0, u, v);
	if(a == 0.0f || l1 == l0)
		return c0;
	float4 c1 = bilinear(p, slot, l1, u, v);
	return make_float4(c0.x + (c1.x - c0.x) * a, c0.y + (c1.y - c0.y) * a, c0.z + (c1.z - c0.z) * a,
					   c0.w + (c1.w - c0.w) * a);
}

__device__ u32 shadeSample(const Params &p, const LucidConfig &cfg, int ipx, int ipy, u32 tri_idx,
						   float &out_depth) {
	float px = float(ipx), py = float(ipy);
	const uint4 *rec = reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx);
	uint4 dq = __ldg(rec), b0q = __ldg(rec + 1), b1q = __ldg(rec + 2), misc = __ldg(rec + 3);
	float dx = __uint_as_float(dq.x), dy = __uint_as_float(dq.y), dz = __uint_as_float(dq.z);
	u32 flags = dq.w & 0xffffu, instance_id = dq.w >> 16;
	float e0x = __uint_as_float(b0q.x), e0y = __uint_as_float(b0q.y), e0z = __uint_as_float(b0q.z);
	float e1x = __uint_as_float(b1q.x), e1y = __uint_as_float(b1q.y), e1z = __uint_as_float(b1q.z);

	float inv_ray_pos = dx * px + (dy * py + dz);
	out_depth = inv_ray_pos;
	if(misc.w != 0)
		return misc.z; // attribute-free triangle: colour was evaluated once in quad setup
	float ray_pos = rcp(inv_ray_pos);
	float e0 = e0x * px + (e0y * py + e0z);
	float e1 = e1x * px + (e1y * py + e1z);
	float b0 = e0 * ray_pos, b1 = e1 * ray_pos;

	float bdx0 = 0, bdx1 = 0, bdy0 = 0, bdy1 = 0;
	const bool textured = (flags & LUCID_INST_HAS_ALBEDO_TEXTURE) != 0;
	if(textured) {
		float ray_posx = rcp(inv_ray_pos + dx);
		float ray_posy = rcp(inv_ray_pos + dy);
		bdx0 = (e0 + e0x) * ray_posx - b0, bdx1 = (e1 + e1x) * ray_posx - b1;
		bdy0 = (e0 + e0y) * ray_posy - b0, bdy1 = (e1 + e1y) * ray_posy - b1;
	}
	b0 -= __uint_as_float(b0q.w), b1 -= __uint_as_float(b1q.w);

	float4 color = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	if(flags & LUCID_INST_HAS_COLOR)
		color = decodeRGBA8(misc.y);

	const u32 second = tri_idx & 1, quad_idx = tri_idx >> 1;
	if(textured) {
		uint4 q0 = __ldg(p.quad_uv + (size_t)quad_idx * 2), q1 = __ldg(p.quad_uv + (size_t)quad_idx * 2 + 1);
		float t0x = __uint_as_float(q0.x), t0y = __uint_as_float(q0.y);
		float t1x = __uint_as_float(second == 0 ? q0.z : q1.x), t1y = __uint_as_float(second == 0 ? q0.w : q1.y);
		float t2x = __uint_as_float(second == 0 ? q1.x : q1.z), t2y = __uint_as_float(second == 0 ? q1.y : q1.w);
		float u = b0 * t1x + (b1 * t2x + t0x), v = b0 * t1y + (b1 * t2y + t0y);
		float dudx = bdx0 * t1x + bdx1 * t2x, dvdx = bdx0 * t1y + bdx1 * t2y;
		float dudy = bdy0 * t1x + bdy1 * t2x, dvdy = bdy0 * t1y + bdy1 * t2y;
		if(flags & LUCID_INST_HAS_UV_RECT) {
			float4 r = __ldg(p.inst_uv_rects + instance_id);
			u = r.z * fractf(u) + r.x, v = r.w * fractf(v) + r.y;
			dudx *= r.z, dvdx *= r.w, dudy *= r.z, dvdy *= r.w;
		}
		float4 tc;
		if(flags & LUCID_INST_TEX_OPAQUE) {
			tc = sampleTexture(p, 0, u, v, dudx, dvdx, dudy, dvdy);
			tc.w = 1.0f;
		} else {
			tc = sampleTexture(p, 1, u, v, dudx, dvdx, dudy, dvdy);
		}
		color.x *= tc.x, color.y *= tc.y, color.z *= tc.z, color.w *= tc.w;
	}
	if(flags & LUCID_INST_HAS_VERTEX_COLORS) {
		uint4 c = __ldg(p.quad_colors + quad_idx);
		float4 c0 = decodeRGBA8(c.x), c1 = decodeRGBA8(second ? c.z : c.y), c2 = decodeRGBA8(second ? c.w : c.z);
		float w0 = 1.0f - b0 - b1;
		color.x *= w0 * c0.x + (b0 * c1.x + b1 * c2.x);
		color.y *= w0 * c0.y + (b0 * c1.y + b1 * c2.y);
		color.z *= w0 * c0.z + (b0 * c1.z + b1 * c2.z);
		color.w *= w0 * c0.w + (b0 * c1.w + b1 * c2.w);
	}
	if(color.w == 0.0f)
		return 0;

	F3 normal;
	if(flags & LUCID_INST_HAS_VERTEX_NORMALS) {
		uint4 n = __ldg(p.quad_normals + quad_idx);
		F3 n0 = decodeNormalUint(n.x);
		F3 n1 = decodeNormalUint(second ? n.z : n.y) - n0, n2 = decodeNormalUint(second ? n.w : n.z) - n0;
		normal = mk3(b0 * n1.x + (b1 * n2.x + n0.x), b0 * n1.y + (b1 * n2.y + n0.y),
					 b0 * n1.z + (b1 * n2.z + n0.z));
	} else {
		normal = decodeNormalUint(misc.x);
	}
	return shadeFinal(cfg.lighting, color, normal);
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
		s.r += cc.x * cc.w, s.g += cc.y * cc.w, s.b += cc.z * cc.w;
	} else {
		s.r += cc.x * cc.w * s.trans, s.g += cc.y * cc.w * s.trans, s.b += cc.z * cc.w * s.trans;
		s.trans *= 1.0f - cc.w;
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

__device__ __forceinline__ u32 bitonic32(u32 v) {
	const u32 lane = laneId();
#pragma unroll
	for(int k = 2; k <= 32; k <<= 1) {
#pragma unroll
		for(int j = k >> 1; j > 0; j >>= 1) {
			u32 o = __shfl_xor_sync(0xffffffffu, v, j);
			bool up = (lane & k) == 0 || k == 32;
			bool lower = (lane & j) == 0;
			v = (lower == up) ? min(v, o) : max(v, o);
		}
	}
	return v;
}

// the last five steps of a bitonic merge (partner distance 16..1) stay in registers
__device__ __forceinline__ u32 bitonicMerge32(u32 v, bool ascending) {
	const u32 lane = laneId();
#pragma unroll
	for(int j = 16; j > 0; j >>= 1) {
		u32 o = __shfl_xor_sync(0xffffffffu, v, j);
		bool lower = (lane & j) == 0;
		v = (lower == ascending) ? min(v, o) : max(v, o);
	}
	return v;
}

// keys[0..n) ascending; the array must have room for n rounded up to a power of two (>= 64)
__device__ void warpSortShared(u32 *keys, int n) {
	const int lane = laneId();
	if(n <= 32) {
		u32 v = lane < n ? keys[lane] : 0xffffffffu;
		v = bitonic32(v);
		if(lane < n)
			keys[lane] = v;
		__syncwarp();
		return;
	}
	int padded = 64;
	while(padded < n)
		padded <<= 1;
	for(int i = n + lane; i < padded; i += 32)
		keys[i] = 0xffffffffu;
	__syncwarp();
	// runs of 32 sorted in registers, alternating direction
	for(int base = 0; base < padded; base += 32) {
		u32 v = bitonic32(keys[base + lane]);
		if(base & 32)
			v = __shfl_sync(0xffffffffu, v, 31 - lane);
		keys[base + lane] = v;
	}
	__syncwarp();
	for(int k = 64; k <= padded; k <<= 1) {
		for(int j = k >> 1; j >= 32; j >>= 1) {
			for(int i = lane; i < (padded >> 1); i += 32) {
				int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
				bool ascending = (lo & k) == 0;
				u32 a = keys[lo], b = keys[hi];
				if((a > b) == ascending)
					keys[lo] = b, keys[hi] = a;
			}
			__syncwarp();
		}
		for(int base = 0; base < padded; base += 32)
			keys[base + lane] = bitonicMerge32(keys[base + lane], (base & k) == 0);
		__syncwarp();
	}
}

// Entries with equal quantised depth are ordered by triangle index.  The low key bits only make
// keys unique (they are list positions that depend on atomic arrival order); this pass makes the
// final order -- and therefore the image -- independent of them.  Ties are rare and short.
template <typename TriOf>
__device__ void warpFixDepthTies(u32 *keys, int n, int slot_bits, TriOf triOf) {
	const int lane = laneId();
	const u32 slot_mask = (1u << slot_bits) - 1u;
	bool tie = false;
	for(int i = lane; i + 1 < n; i += 32)
		tie = tie || (keys[i] >> slot_bits) == (keys[i + 1] >> slot_bits);
	if(!__any_sync(0xffffffffu, tie))
		return;
	while(true) {
		bool swapped = false;
#pragma unroll
		for(int parity = 0; parity < 2; parity++) {
			for(int i = 2 * lane + parity; i + 1 < n; i += 64) {
				u32 a = keys[i], b = keys[i + 1];
				if((a >> slot_bits) == (b >> slot_bits) && triOf(a & slot_mask) > triOf(b & slot_mask)) {
					keys[i] = b, keys[i + 1] = a;
					swapped = true;
				}
			}
			__syncwarp();
		}
		if(!__any_sync(0xffffffffu, swapped))
			break;
	}
}

// ------------------------------------------------------------------------------------------------
// the work item

template <bool HIGH> struct Geo {
	static constexpr int rows_per_group = HIGH ? 4 : 8;
	static constexpr int group_shift = HIGH ? 2 : 3;
	static constexpr int halves = HIGH ? 1 : 2;
	static constexpr int slot_bits = HIGH ? 14 : 10;
};

struct WarpScratch {
	u32 *keys;	  // CAP entries
	u32 *samples; // SAMPLE_BUF entries
	u32 *mask;	  // 32 entries
};

// Streams the sorted triangle list of one 8x4 half-block: expands triangles into samples in
// 256-sample segments, shades 32 samples per round and feeds every pixel's samples to its lane.
// getRow(slot) returns (mins, maxs, tri_idx) of the list entry for this half.
template <typename GetRow>
__device__ void shadeHalfBlock(const Params &p, const LucidConfig &cfg, const WarpScratch &ws,
							   int count, int slot_mask, int startx, int hb_x, int hb_y,
							   GetRow getRow, u32 &out_frags) {
	const int lane = laneId();
	const bool additive = (p.opts & LUCID_OPT_ADDITIVE_BLENDING) != 0;
	const bool vis_errors = (p.opts & LUCID_OPT_VISUALIZE_ERRORS) != 0;
	const bool alpha_thr = (p.opts & LUCID_OPT_ALPHA_THRESHOLD) != 0 && !additive && !vis_errors;
	Reducer red;
	reducerInit(red);
	u32 px_frags = 0, total_frags = 0;

	int next = 0;		 // next list entry to expand
	u32 seg_start = 0;	 // sample offset of the current segment
	u32 off = 0;		 // sample offset of entry `next`
	int carried = 0;	 // samples spilled past the previous segment (< 32)
	bool stop = false;

	while(next < count || carried > 0) {
		// move the spill of the previous segment to the front (raster.glsl:302-304)
		if(carried > 0) {
			u32 v = ws.samples[SEGMENT_SIZE + lane];
			__syncwarp();
			ws.samples[lane] = v;
		}
		__syncwarp();
		// expand entries whose first sample lies inside this segment
		const u32 seg_end = seg_start + SEGMENT_SIZE;
		while(next < count && off < seg_end) {
			int i = next + lane;
			u32 mins = 0, maxs = 0, tri_idx = 0;
			u32 bits = 0;
			int nf = 0;
			if(i < count) {
				u32 slot = ws.keys[i] & slot_mask;
				getRow(slot, mins, maxs, tri_idx);
				int cx, cy;
				bits = halfPixelMask(mins, maxs, startx, nf, cx, cy);
			}
			int incl = nf;
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, incl, o);
				if(lane >= o)
					incl += t;
			}
			u32 my_off = off + (u32)(incl - nf);
			bool in_seg = i < count && my_off < seg_end;
			u32 in_mask = __ballot_sync(0xffffffffu, in_seg);
			int taken = __popc(in_mask); // a prefix of the lanes
			if(in_seg) {
				u32 dst = my_off - seg_start;
				u32 word = tri_idx << 8;
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
		total_frags += (u32)nseg;

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
					reducerPush(red, c, d, additive, vis_errors);
				}
			}
			__syncwarp();
		}
		// end of segment: saturate (raster.glsl:394-395), optional early out
		red.r = saturatef(red.r), red.g = saturatef(red.g), red.b = saturatef(red.b);
		if(alpha_thr && nseg == SEGMENT_SIZE && __all_sync(0xffffffffu, red.trans < (1.0f / 128.0f)))
			stop = true;
		seg_start += SEGMENT_SIZE;
	}

	// finishReduceSamples (shading.glsl:297-314)
	if(red.c2 != 0)
		reducerBlend(red, red.c2, additive);
	if(red.c1 != 0)
		reducerBlend(red, red.c1, additive);
	if(red.c0 != 0)
		reducerBlend(red, red.c0, additive);
	float fr = saturatef(red.r + red.trans * cfg.background_color.x);
	float fg = saturatef(red.g + red.trans * cfg.background_color.y);
	float fb = saturatef(red.b + red.trans * cfg.background_color.z);
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
	out_frags = total_frags;
}

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

// scanline state of a triangle at the first pixel row of group `g` of the bin (groups of
// 8 rows for LOW, 4 for HIGH).  The walk starts at the triangle's first group inside the bin and
// adds the step row by row, exactly like the reference's incremental loop, so the values -- and
// the spans truncated from them -- are bit-identical.
template <int ROWS_PER_GROUP>
__device__ __forceinline__ bool rowScanAt(const Params &p, u32 tri_idx, int pos_x, int pos_y, int g, RowScan &rs) {
	constexpr int shift = ROWS_PER_GROUP == 8 ? 3 : 2;
	const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_scan + tri_idx);
	uint4 s0 = __ldg(src);
	int ymin = (int)(s0.w & 0xffff) - pos_y, ymax = (int)(s0.w >> 16) - pos_y;
	int min_g = min(max(ymin, 0), BIN_SIZE - 1) >> shift, max_g = min(max(ymax, 0), BIN_SIZE - 1) >> shift;
	if(g < min_g || g > max_g)
		return false;
	uint4 s1 = __ldg(src + 1);
	rs.step[0] = __uint_as_float(s1.x), rs.step[1] = __uint_as_float(s1.y), rs.step[2] = __uint_as_float(s1.z);
	rs.xneg = s1.w & 7u;
	float start_x = float(pos_x), start_y = float(pos_y + min_g * ROWS_PER_GROUP);
	rs.scan[0] = __uint_as_float(s0.x) + (rs.step[0] * start_y - start_x);
	rs.scan[1] = __uint_as_float(s0.y) + (rs.step[1] * start_y - start_x);
	rs.scan[2] = __uint_as_float(s0.z) + (rs.step[2] * start_y - start_x);
	for(int k = (g - min_g) * ROWS_PER_GROUP; k > 0; k--)
		rs.scan[0] += rs.step[0], rs.scan[1] += rs.step[1], rs.scan[2] += rs.step[2];
	return true;
}
__device__ __forceinline__ bool touchesGroup(const Params &p, u32 tri_idx, int pos_y, int g, int shift) {
	u32 y_aabb = __ldg(reinterpret_cast<const u32 *>(p.tri_scan + tri_idx) + 3);
	int ymin = (int)(y_aabb & 0xffff) - pos_y, ymax = (int)(y_aabb >> 16) - pos_y;
	int min_g = min(max(ymin, 0), BIN_SIZE - 1) >> shift, max_g = min(max(ymax, 0), BIN_SIZE - 1) >> shift;
	return g >= min_g && g <= max_g;
}

// LOW: one block row of a bin with fewer than 1024 triangles (raster_low.glsl)
__device__ void rasterLowItem(const Params &p, const LucidConfig &cfg, int bin_id, int by,
							  unsigned char *smem) {
	uint4 *s_rows = reinterpret_cast<uint4 *>(smem);					   // LOW_MAX_TRIS, indexed by t
	u32 *s_tri = reinterpret_cast<u32 *>(smem + LOW_MAX_TRIS * 16);		   // LOW_MAX_TRIS
	unsigned short *s_queue = reinterpret_cast<unsigned short *>(smem + LOW_MAX_TRIS * 20); // LOW_MAX_TRIS
	u32 *s_warp = reinterpret_cast<u32 *>(smem + LOW_MAX_TRIS * 22);
	__shared__ int s_qcount;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	WarpScratch ws;
	ws.keys = s_warp + warp * (MAX_BLOCK_TRIS + SAMPLE_BUF + 32);
	ws.samples = ws.keys + MAX_BLOCK_TRIS;
	ws.mask = ws.samples + SAMPLE_BUF;

	const int n_q = cntc(p, LUCID_CNT_QUAD_COUNTS)[bin_id], q_off = cntc(p, LUCID_CNT_QUAD_OFFSETS)[bin_id];
	const int n_t = cntc(p, LUCID_CNT_TRI_COUNTS)[bin_id], t_off = cntc(p, LUCID_CNT_TRI_OFFSETS)[bin_id];
	const int n_T = n_q * 2 + n_t; // < 1024
	const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
	const int pos_x = bin_x * BIN_SIZE, pos_y = bin_y * BIN_SIZE;
	if(tid == 0)
		s_qcount = 0;
	__syncthreads();

	// phase A1: which triangles of the bin reach this block row (y range only)
	for(int t0 = 0; t0 < n_T; t0 += RASTER_THREADS) {
		int t = t0 + tid;
		u32 tri_idx = 0;
		bool pass = t < n_T && binTriangle(p, t, n_q, q_off, t_off, tri_idx);
		pass = pass && touchesGroup(p, tri_idx, pos_y, by, 3);
		if(t < n_T)
			s_tri[t] = tri_idx;
		u32 m = __ballot_sync(0xffffffffu, pass);
		int base = 0;
		if(lane == 0 && m)
			base = atomicAdd(&s_qcount, __popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if(pass)
			s_queue[base + __popc(m & laneMaskLt())] = (unsigned short)t;
	}
	__syncthreads();
	// phase A2: spans of the queued triangles, all lanes busy (raster_low.glsl:39-64)
	const int n_queue = s_qcount;
	for(int q = tid; q < n_queue; q += RASTER_THREADS) {
		int t = s_queue[q];
		RowScan rs;
		uint4 rec = make_uint4(0, 0, 0, 0);
		if(rowScanAt<8>(p, s_tri[t], pos_x, pos_y, by, rs)) {
			u32 mn0, mx0, bx0, mn1, mx1, bx1;
			rasterBinStep(rs, mn0, mx0, bx0);
			rasterBinStep(rs, mn1, mx1, bx1);
			u32 bx = bx0 | bx1;
			if(bx != 0)
				rec = make_uint4(mn0 | (bx << 24), mn1, mx0, mx1);
		}
		s_rows[t] = rec;
	}
	__syncthreads();

	// phase B: warp = block column (raster_low.glsl:81-194)
	const int bx = warp;
	int count = 0;
	for(int q0 = 0; q0 < n_queue; q0 += 32) {
		int q = q0 + lane;
		int t = q < n_queue ? s_queue[q] : 0;
		bool has = q < n_queue && ((s_rows[t].x >> (24 + bx)) & 1);
		u32 m = __ballot_sync(0xffffffffu, has);
		int pos = count + __popc(m & laneMaskLt());
		if(has && pos < MAX_BLOCK_TRIS)
			ws.keys[pos] = (u32)t;
		count += __popc(m);
	}
	__syncwarp();
	if(count > MAX_BLOCK_TRIS) {
		// too many triangles for one block: the whole bin is redone by the HIGH path
		// (raster_low.glsl:101-105,230-237)
		if(lane == 0)
			atomicOr(&p.bin_flags[bin_id], 1u);
		return;
	}
	const int startx = bx * 8;
	u32 frag_acc = 0;
	for(int i = lane; i < count; i += 32) {
		u32 t = ws.keys[i];
		uint4 rec = s_rows[t];
		int nf0, cx0, cy0, nf1, cx1, cy1;
		halfPixelMask(rec.x, rec.z, startx, nf0, cx0, cy0);
		halfPixelMask(rec.y, rec.w, startx, nf1, cx1, cy1);
		// both halves use row offsets 1,3,5,7 for the centroid, exactly as the reference does
		float cx = float(cx0) + float(cx1), cy = float(cy0) + float(cy1);
		float scale = __fdiv_rn(0.5f, float(nf0 + nf1));
		float cpx = cx * scale + float(pos_x + bx * 8), cpy = cy * scale + float(pos_y + by * 8);
		u32 depth = blockDepth(p, s_tri[t], cpx, cpy, float(0x3ffffe));
		ws.keys[i] = t | (depth << 10);
		frag_acc += (u32)nf0 | ((u32)nf1 << 16);
	}
	__syncwarp();
	if(count > 3) { // blocks with <= 3 triangles rely on the window alone (raster_low.glsl:144)
		warpSortShared(ws.keys, count);
		warpFixDepthTies(ws.keys, count, 10, [&](u32 t) { return s_tri[t]; });
	}
	__syncwarp();

	auto getRow0 = [&](u32 slot, u32 &mins, u32 &maxs, u32 &tri) {
		uint4 r = s_rows[slot];
		mins = r.x, maxs = r.z, tri = s_tri[slot];
	};
	auto getRow1 = [&](u32 slot, u32 &mins, u32 &maxs, u32 &tri) {
		uint4 r = s_rows[slot];
		mins = r.y, maxs = r.w, tri = s_tri[slot];
	};
	u32 f0, f1;
	shadeHalfBlock(p, cfg, ws, count, 0x3ff, startx, pos_x + bx * 8, pos_y + by * 8, getRow0, f0);
	shadeHalfBlock(p, cfg, ws, count, 0x3ff, startx, pos_x + bx * 8, pos_y + by * 8 + 4, getRow1, f1);
#pragma unroll
	for(int o = 16; o > 0; o >>= 1)
		frag_acc += __shfl_xor_sync(0xffffffffu, frag_acc, o);
	if(lane == 0) {
		// stats: fragments, and the block's triangle count once per half-block (raster_low.glsl:272-275)
		atomicAdd(&p.bin_stats[bin_id * 4 + 0], (frag_acc & 0xffffu) + (frag_acc >> 16));
		atomicAdd(&p.bin_stats[bin_id * 4 + 1], (u32)count * 2u);
	}
}

__global__ void __launch_bounds__(RASTER_THREADS) k_raster_low(const Params p,
															   const __grid_constant__ LucidConfig cfg) {
	extern __shared__ __align__(16) unsigned char smem[];
	const int n_low = p.info->bin_level_counts[LUCID_BIN_LEVEL_LOW];
	const int item = blockIdx.x;
	if(item >= n_low * 4)
		return;
	const int bin_id = cntc(p, LUCID_CNT_LOW_BINS)[item >> 2];
	rasterLowItem(p, cfg, bin_id, item & 3, smem);
}

// appends promoted LOW bins to the HIGH list in bin order (raster_low.glsl:230-237,294-298)
__global__ void __launch_bounds__(1024) k_promote(const Params p) {
	__shared__ int s_warp[33];
	const int n_low = p.info->bin_level_counts[LUCID_BIN_LEVEL_LOW];
	const int *low = cntc(p, LUCID_CNT_LOW_BINS);
	int *high = p.counts + (size_t)LUCID_CNT_HIGH_BINS * p.bin_count;
	const int n_high = p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH];
	const int per = (n_low + 1023) / 1024;
	const int i0 = min((int)threadIdx.x * per, n_low), i1 = min(i0 + per, n_low);
	int mine = 0;
	for(int i = i0; i < i1; i++)
		mine += (p.bin_flags[low[i]] & 1u) ? 1 : 0;
	// exclusive scan over threads
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
	if(warp == 0) {
		int w = s_warp[lane], wi = w;
		for(int o = 1; o < 32; o <<= 1) {
			int t = __shfl_up_sync(0xffffffffu, wi, o);
			if(lane >= o)
				wi += t;
		}
		s_warp[lane] = wi - w;
		if(lane == 31)
			s_warp[32] = wi;
	}
	__syncthreads();
	int pos = n_high + s_warp[warp] + incl - mine;
	for(int i = i0; i < i1; i++)
		if(p.bin_flags[low[i]] & 1u)
			high[pos++] = low[i];
	if(threadIdx.x == 0 && s_warp[32] > 0) {
		int total = n_high + s_warp[32];
		p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH] = total;
		u32 nd = (u32)min(total, p.max_dispatches / 2);
		if(nd > p.info->bin_level_dispatches[LUCID_BIN_LEVEL_HIGH][0])
			p.info->bin_level_dispatches[LUCID_BIN_LEVEL_HIGH][0] = nd;
	}
}

// HIGH: one half-block row (4 pixel rows) of a dense bin (raster_high.glsl)
template <int CAP>
__device__ void rasterHighItem(const Params &p, const LucidConfig &cfg, int item, uint4 *scratch,
							   unsigned char *smem) {
	constexpr int ROW_CAP = CAP * 4 < MAX_HBLOCK_ROW_TRIS ? CAP * 4 : MAX_HBLOCK_ROW_TRIS;
	constexpr int QUEUE = RASTER_THREADS * 2;
	unsigned char *s_bx = smem;									  // ROW_CAP bytes
	u32 *s_warp = reinterpret_cast<u32 *>(smem + ROW_CAP);
	__shared__ u32 s_queue[QUEUE];
	__shared__ int s_qtail, s_row_count;
	__shared__ int s_est[4], s_exact[4], s_status;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	WarpScratch ws;
	ws.keys = s_warp + warp * (CAP + SAMPLE_BUF + 32);
	ws.samples = ws.keys + CAP;
	ws.mask = ws.samples + SAMPLE_BUF;

	const int bin_id = cntc(p, LUCID_CNT_HIGH_BINS)[item >> 3], rby = item & 7;
	const int n_q = cntc(p, LUCID_CNT_QUAD_COUNTS)[bin_id], q_off = cntc(p, LUCID_CNT_QUAD_OFFSETS)[bin_id];
	const int n_t = cntc(p, LUCID_CNT_TRI_COUNTS)[bin_id], t_off = cntc(p, LUCID_CNT_TRI_OFFSETS)[bin_id];
	const int n_T = n_q * 2 + n_t;
	const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
	const int pos_x = bin_x * BIN_SIZE, pos_y = bin_y * BIN_SIZE;

	if(tid < 4)
		s_est[tid] = 0, s_exact[tid] = 0;
	if(tid == 0) {
		s_status = (p.bin_flags[bin_id] & 2u) ? 2 : 0;
		s_qtail = 0, s_row_count = 0;
	}
	__syncthreads();
	if(s_status != 0)
		return; // another item already found the bin over a limit

	// phase A (raster_high.glsl:54-106).  A1 queues the triangles whose y range reaches this
	// half-block row; A2 evaluates 128 queued triangles at a time with every lane busy and appends
	// the non-empty ones to the row list.  List order does not matter (see warpFixDepthTies).
	int est[4] = {0, 0, 0, 0}, exact[4] = {0, 0, 0, 0};
	int qhead = 0;
	auto evaluate = [&](int qi, bool active) {
		u32 mn = 0, mx = 0, bx = 0, tri_idx = 0;
		if(active) {
			tri_idx = s_queue[qi & (QUEUE - 1)];
			RowScan rs;
			if(rowScanAt<4>(p, tri_idx, pos_x, pos_y, rby, rs))
				rasterBinStep(rs, mn, mx, bx);
		}
		bool has = bx != 0;
		u32 m = __ballot_sync(0xffffffffu, has);
		int base = 0;
		if(lane == 0 && m)
			base = atomicAdd(&s_row_count, __popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if(has) {
			int slot = base + __popc(m & laneMaskLt());
			if(slot < ROW_CAP) {
				scratch[slot] = make_uint4(mn, mx, tri_idx, bx);
				s_bx[slot] = (unsigned char)bx;
			}
			// estimated (first..last column, holes included) and exact per-half-block counts
			int lo = __ffs(bx) - 1, hi = 31 - __clz(bx);
#pragma unroll
			for(int c = 0; c < 4; c++) {
				est[c] += (c >= lo && c <= hi) ? 1 : 0;
				exact[c] += (bx >> c) & 1;
			}
		}
	};
	for(int t0 = 0; t0 < n_T; t0 += RASTER_THREADS) {
		int t = t0 + tid;
		u32 tri_idx = 0;
		bool pass = t < n_T && binTriangle(p, t, n_q, q_off, t_off, tri_idx);
		pass = pass && touchesGroup(p, tri_idx, pos_y, rby, 2);
		u32 m = __ballot_sync(0xffffffffu, pass);
		int base = 0;
		if(lane == 0 && m)
			base = atomicAdd(&s_qtail, __popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if(pass)
			s_queue[(base + __popc(m & laneMaskLt())) & (QUEUE - 1)] = tri_idx;
		__syncthreads();
		if(s_qtail - qhead >= RASTER_THREADS) {
			evaluate(qhead + tid, true);
			qhead += RASTER_THREADS;
		}
		__syncthreads();
	}
	{
		int rest = s_qtail - qhead;
		evaluate(qhead + tid, tid < rest);
	}
#pragma unroll
	for(int c = 0; c < 4; c++) {
		int e = est[c], x = exact[c];
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) {
			e += __shfl_xor_sync(0xffffffffu, e, o);
			x += __shfl_xor_sync(0xffffffffu, x, o);
		}
		if(lane == 0) {
			atomicAdd(&s_est[c], e);
			atomicAdd(&s_exact[c], x);
		}
	}
	__syncthreads();
	const int row_count = s_row_count;
	if(tid == 0) {
		int max_est = max(max(s_est[0], s_est[1]), max(s_est[2], s_est[3]));
		int max_exact = max(max(s_exact[0], s_exact[1]), max(s_exact[2], s_exact[3]));
		if(row_count > MAX_HBLOCK_ROW_TRIS || max_est > MAX_HBLOCK_TRIS) {
			// over the reference's limits: the bin is painted red (raster_high.glsl:80-83,140-141)
			atomicOr(&p.bin_flags[bin_id], 2u);
			s_status = 2;
		} else if(row_count > ROW_CAP || max_exact > CAP) {
			// does not fit this kernel's shared memory: hand the item to the large variant
			int idx = atomicAdd(&p.work_counters[3], 1u);
			p.deferred_items[idx] = item;
			s_status = 1;
		}
	}
	__syncthreads();
	if(s_status != 0)
		return;

	// phase B: warp = half-block column (raster_high.glsl:146-273)
	const int hbx = warp;
	int count = 0;
	for(int s0 = 0; s0 < row_count; s0 += 32) {
		int sl = s0 + lane;
		bool has = sl < row_count && ((s_bx[sl] >> hbx) & 1);
		u32 m = __ballot_sync(0xffffffffu, has);
		if(has)
			ws.keys[count + __popc(m & laneMaskLt())] = (u32)sl;
		count += __popc(m);
	}
	__syncwarp();
	const int startx = hbx * 8;
	u32 fsum = 0;
	for(int i = lane; i < count; i += 32) {
		u32 slot = ws.keys[i];
		uint4 rec = scratch[slot];
		int nf, cx, cy;
		halfPixelMask(rec.x, rec.y, startx, nf, cx, cy);
		float scale = __fdiv_rn(0.5f, float(nf));
		float cpx = float(cx) * scale + (float(hbx * 8) + float(pos_x));
		float cpy = float(cy) * scale + (float(rby * 4) + float(pos_y));
		u32 depth = blockDepth(p, rec.z, cpx, cpy, float(0x7fffe));
		ws.keys[i] = slot | (depth << 14);
		fsum += (u32)nf;
	}
	__syncwarp();
	warpSortShared(ws.keys, count);
	warpFixDepthTies(ws.keys, count, 14, [&](u32 slot) { return scratch[slot].z; });
	__syncwarp();
	auto getRow = [&](u32 slot, u32 &mins, u32 &maxs, u32 &tri) {
		uint4 r = scratch[slot];
		mins = r.x, maxs = r.y, tri = r.z;
	};
	u32 frags;
	shadeHalfBlock(p, cfg, ws, count, 0x3fff, startx, pos_x + hbx * 8, pos_y + rby * 4, getRow, frags);
#pragma unroll
	for(int o = 16; o > 0; o >>= 1)
		fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
	if(lane == 0) {
		// exact per-half-block counts (raster_high.glsl:309-310)
		atomicAdd(&p.bin_stats[bin_id * 4 + 2], fsum);
		atomicAdd(&p.bin_stats[bin_id * 4 + 3], (u32)count);
	}
}

template <int CAP, bool DEFERRED>
__global__ void __launch_bounds__(RASTER_THREADS) k_raster_high(const Params p,
																const __grid_constant__ LucidConfig cfg) {
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ int s_item;
	constexpr int ROW_CAP = CAP * 4 < MAX_HBLOCK_ROW_TRIS ? CAP * 4 : MAX_HBLOCK_ROW_TRIS;
	uint4 *scratch = p.high_scratch + (size_t)blockIdx.x * ROW_CAP;
	const int n_items = DEFERRED ? (int)p.work_counters[3] : p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH] * 8;
	while(true) {
		__syncthreads();
		if(threadIdx.x == 0)
			s_item = (int)atomicAdd(&p.work_counters[DEFERRED ? 2 : 1], 1u);
		__syncthreads();
		int idx = s_item;
		if(idx >= n_items)
			break;
		int item = DEFERRED ? p.deferred_items[idx] : idx;
		rasterHighItem<CAP>(p, cfg, item, scratch, smem);
	}
}

// background for bins no kernel writes, red for bins over the reference's limits, and the
// statistics (shading.glsl:38-53); LOW results of promoted bins are not counted
__global__ void __launch_bounds__(256) k_raster_finish(const Params p, u32 background, int fill_empty) {
	const int *qc = cntc(p, LUCID_CNT_QUAD_COUNTS), *tc = cntc(p, LUCID_CNT_TRI_COUNTS);
	u32 frags = 0, hbt = 0;
	for(int b = blockIdx.x; b < p.bin_count; b += gridDim.x) {
		int by = b / p.bin_count_x, bx = b - by * p.bin_count_x;
		u32 flags = p.bin_flags[b];
		int num_tris = tc[b] + qc[b] * 2;
		bool empty = num_tris == 0;
		bool error = (flags & 2u) != 0;
		if(by < p.row_begin || by >= p.row_end)
			continue;
		if((empty && fill_empty) || error) {
			u32 value = error ? 0x000000ffu : background;
			for(int i = threadIdx.x; i < BIN_SIZE * BIN_SIZE; i += blockDim.x) {
				int gx = bx * BIN_SIZE + (i & 31), gy = by * BIN_SIZE + (i >> 5);
				if(gx < p.width && gy < p.height) {
					p.image[(size_t)gy * p.image_pitch + gx] = value;
					if(p.frag_counts)
						p.frag_counts[(size_t)gy * p.width + gx] = 0;
				}
			}
		}
		if(threadIdx.x == 0 && !empty && !error) {
			bool high = num_tris >= 1024 || (flags & 1u);
			frags += p.bin_stats[b * 4 + (high ? 2 : 0)];
			hbt += p.bin_stats[b * 4 + (high ? 3 : 1)];
		}
	}
	if(threadIdx.x == 0) {
		if(frags)
			atomicAdd(&p.info->stats[0], frags);
		if(hbt)
			atomicAdd(&p.info->stats[1], hbt);
	}
}

constexpr int lowSmemBytes() { return LOW_MAX_TRIS * 22 + RASTER_WARPS * (MAX_BLOCK_TRIS + SAMPLE_BUF + 32) * 4; }
template <int CAP> constexpr int highSmemBytes() {
	return (CAP * 4 < MAX_HBLOCK_ROW_TRIS ? CAP * 4 : MAX_HBLOCK_ROW_TRIS) + RASTER_WARPS * (CAP + SAMPLE_BUF + 32) * 4;
}

int rasterHighGridSmall(int num_sms) { return num_sms * 8; }
int rasterHighGridLarge(int num_sms) { return num_sms * 2; }

void launchRaster(const Params &p, const LucidConfig &cfg, cudaStream_t stream, cudaEvent_t *ev,
				  int num_sms) {
	static bool configured = false;
	if(!configured) {
		cudaFuncSetAttribute(k_raster_low, cudaFuncAttributeMaxDynamicSharedMemorySize, lowSmemBytes());
		cudaFuncSetAttribute(k_raster_high<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
							 highSmemBytes<1024>());
		cudaFuncSetAttribute(k_raster_high<4096, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
							 highSmemBytes<4096>());
		configured = true;
	}
	k_raster_low<<<p.bin_count * 4, RASTER_THREADS, lowSmemBytes(), stream>>>(p, cfg);
	k_promote<<<1, 1024, 0, stream>>>(p);
	if(ev)
		cudaEventRecord(ev[0], stream);
	k_raster_high<1024, false><<<rasterHighGridSmall(num_sms), RASTER_THREADS, highSmemBytes<1024>(), stream>>>(p, cfg);
	k_raster_high<4096, true><<<rasterHighGridLarge(num_sms), RASTER_THREADS, highSmemBytes<4096>(), stream>>>(p, cfg);
	if(ev)
		cudaEventRecord(ev[1], stream);
	const LucidVec4 &bg = cfg.background_color;
	auto q = [](float v) { return (u32)(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f + 0.5f); };
	u32 bg8 = q(bg.x) | (q(bg.y) << 8) | (q(bg.z) << 16) | 0xff000000u;
	k_raster_finish<<<num_sms * 2, 256, 0, stream>>>(p, bg8, 1);
	if(ev)
		cudaEventRecord(ev[2], stream);
}

} // namespace lucid
