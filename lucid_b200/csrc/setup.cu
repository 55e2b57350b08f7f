// setup.cu -- quad/triangle setup for sm_100a.
//
// Replaces data/shaders/quad_setup.glsl (reference file:line cited per function).  What changes
// against the reference's structure: the work is split into k_quad_cull (one thread per input
// quad: culling, bin AABB, compaction) and k_tri_setup (one thread per visible triangle: the
// equations), so the expensive part runs barrier-free at full occupancy over compacted slots; the
// two global atomicAdd compactions (quad_setup.glsl:415-421) become a single-pass decoupled
// look-back scan over CTAs taken in ticket order, so visible-quad slots are a pure function of the
// input order (deterministic); per-sample triangle fields are written as one 64-byte record.
#include "common.cuh"

namespace lucid {

constexpr int SETUP_THREADS = 256;
constexpr int SETUP_PARTS = LUCID_MAX_INSTANCE_QUADS / SETUP_THREADS;

// positions are read from a 16-byte padded copy made once per lucid_set_geometry: one 128-bit load
// per vertex instead of three scalar ones on the tightly packed float3 array (quad_setup.glsl:64-66)
__device__ __forceinline__ F3 vertexLoad(const Params &p, u32 vi) {
	float4 v = __ldg(p.positions4 + vi);
	return mk3(v.x, v.y, v.z);
}
__global__ void __launch_bounds__(256) k_pad_positions(const float *positions, float4 *positions4, int num_verts) {
	for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_verts; i += gridDim.x * blockDim.x)
		positions4[i] = make_float4(positions[(size_t)i * 3], positions[(size_t)i * 3 + 1], positions[(size_t)i * 3 + 2], 1.0f);
}
void launchPadPositions(const float *positions, float4 *positions4, int num_verts, cudaStream_t stream) {
	k_pad_positions<<<148 * 4, 256, 0, stream>>>(positions, positions4, num_verts);
}

__device__ __forceinline__ u32 vertexClipMask(float4 v) {
	return (v.x < -v.w ? 0x01u : 0u) | (v.x > v.w ? 0x02u : 0u) | (v.y < -v.w ? 0x04u : 0u) |
		   (v.y > v.w ? 0x08u : 0u) | (v.z < -v.w ? 0x10u : 0u) | (v.z > v.w ? 0x20u : 0u);
}

__device__ __forceinline__ float &comp(float4 &v, int i) { return (&v.x)[i]; }

// Bounding boxes of the instances' vertices (bin-row split, LUCID_RENDER_CULL_INSTANCES): one CTA per instance,
// once per instance list.  k_quad_cull projects the box and drops the instance when it cannot reach an owned row.
__global__ void __launch_bounds__(256) k_instance_boxes(const Params p, float4 *boxes) {
	__shared__ float s_red[8][6];
	const LucidInstanceData inst = p.instances[blockIdx.x];
	const uint4 *ib = reinterpret_cast<const uint4 *>(reinterpret_cast<const u32 *>(p.quad_indices) + inst.index_offset);
	const float inf = __int_as_float(0x7f800000);
	float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
	for(int q = threadIdx.x; q < inst.num_quads; q += blockDim.x) {
		const uint4 vi = __ldg(ib + q);
		const u32 idx[4] = {vi.x, vi.y, vi.z, vi.w};
#pragma unroll
		for(int k = 0; k < 4; k++) {
			const u32 v = idx[k] + (u32)inst.vertex_offset;
			if(v >= (u32)p.num_verts)
				continue;
			const float4 pos = __ldg(p.positions4 + v);
			lo[0] = fminf(lo[0], pos.x), lo[1] = fminf(lo[1], pos.y), lo[2] = fminf(lo[2], pos.z);
			hi[0] = fmaxf(hi[0], pos.x), hi[1] = fmaxf(hi[1], pos.y), hi[2] = fmaxf(hi[2], pos.z);
		}
	}
#pragma unroll
	for(int k = 0; k < 3; k++)
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) {
			lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
			hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
		}
	if((threadIdx.x & 31) == 0)
		for(int k = 0; k < 3; k++)
			s_red[threadIdx.x >> 5][k] = lo[k], s_red[threadIdx.x >> 5][3 + k] = hi[k];
	__syncthreads();
	if(threadIdx.x == 0) {
		for(int w = 1; w < 8; w++)
			for(int k = 0; k < 3; k++)
				lo[k] = fminf(lo[k], s_red[w][k]), hi[k] = fmaxf(hi[k], s_red[w][3 + k]);
		boxes[blockIdx.x * 2] = make_float4(lo[0], lo[1], lo[2], 0.0f);
		boxes[blockIdx.x * 2 + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
	}
}
void launchInstanceBoxes(const Params &p, float4 *boxes, cudaStream_t stream) {
	if(p.num_instances > 0)
		k_instance_boxes<<<p.num_instances, 256, 0, stream>>>(p, boxes);
}

// true when the box certainly projects outside the owned bin rows: all eight corners are in front of the camera
// (w above a small positive bound) and their screen y range, widened by two pixels, misses [row_begin, row_end)
__device__ __forceinline__ bool boxOutsideOwnedRows(const Params &p, const LucidConfig &cfg, float4 lo, float4 hi) {
	if(!(lo.x <= hi.x))
		return false; // empty or invalid box: let the quads decide
	const LucidVec4 *m = cfg.view_proj_matrix;
	float ymin = __int_as_float(0x7f800000), ymax = -ymin;
#pragma unroll
	for(int c = 0; c < 8; c++) {
		const float x = (c & 1) ? hi.x : lo.x, y = (c & 2) ? hi.y : lo.y, z = (c & 4) ? hi.z : lo.z;
		const float cy = m[0].y * x + m[1].y * y + m[2].y * z + m[3].y;
		const float cw = m[0].w * x + m[1].w * y + m[2].w * z + m[3].w;
		if(!(cw > 1e-3f))
			return false;
		const float sy = (cy / cw + 1.0f) * (float(p.height) * 0.5f);
		ymin = fminf(ymin, sy), ymax = fmaxf(ymax, sy);
	}
	return ymax + 2.0f < float(p.row_begin * BIN_SIZE) || ymin - 2.0f >= float(p.row_end * BIN_SIZE);
}

// true when all eight corners lie outside one clip plane by more than the rounding of the transform: every vertex in
// the box then has that plane's bit in its clip mask, so processInputQuad would reject each of its quads
// (and_mask != 0, quad_setup.glsl:176-180)
__device__ __forceinline__ bool boxOutsideFrustum(const LucidConfig &cfg, float4 lo, float4 hi) {
	if(!(lo.x <= hi.x))
		return false;
	const LucidVec4 *m = cfg.view_proj_matrix;
	u32 all = 0x3fu;
#pragma unroll
	for(int c = 0; c < 8; c++) {
		const float x = (c & 1) ? hi.x : lo.x, y = (c & 2) ? hi.y : lo.y, z = (c & 4) ? hi.z : lo.z;
		float v[4], mag[4];
#pragma unroll
		for(int k = 0; k < 4; k++) {
			const float a = (&m[0].x)[k] * x, b = (&m[1].x)[k] * y, d = (&m[2].x)[k] * z, e = (&m[3].x)[k];
			v[k] = a + b + d + e;
			mag[k] = fabsf(a) + fabsf(b) + fabsf(d) + fabsf(e);
		}
		u32 mask = 0;
#pragma unroll
		for(int k = 0; k < 3; k++) {
			const float tol = 1e-5f * (mag[k] + mag[3]);
			mask |= (v[k] + v[3] < -tol ? 1u : 0u) << (2 * k);
			mask |= (v[k] - v[3] > tol ? 2u : 0u) << (2 * k);
		}
		all &= mask;
	}
	return all != 0;
}

// Instance culling (LUCID_RENDER_CULL_INSTANCES): the instances whose box can reach the frustum and the owned bin rows,
// in input order, for k_quad_cull to work on -- one CTA, an ordered compaction over tiles of 1024 instances.  The
// quads of the others only count as input.
__global__ void __launch_bounds__(1024) k_instance_select(const Params p, const __grid_constant__ LucidConfig cfg) {
	__shared__ int s_warp[32];
	__shared__ int s_base;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdlEntry();
	if(tid == 0)
		s_base = 0;
	__syncthreads();
	u32 skipped_quads = 0;
	for(int first = 0; first < p.num_instances; first += 1024) {
		const int i = first + tid;
		bool keep = false;
		if(i < p.num_instances) {
			const float4 lo = __ldg(p.inst_boxes + i * 2), hi = __ldg(p.inst_boxes + i * 2 + 1);
			keep = !(boxOutsideOwnedRows(p, cfg, lo, hi) || boxOutsideFrustum(cfg, lo, hi));
			if(!keep)
				skipped_quads += (u32)p.instances[i].num_quads;
		}
		const u32 bal = __ballot_sync(0xffffffffu, keep);
		if(lane == 0)
			s_warp[warp] = __popc(bal);
		__syncthreads();
		int before = s_base, total = 0;
		for(int w = 0; w < 32; w++) {
			before += w < warp ? s_warp[w] : 0;
			total += s_warp[w];
		}
		if(keep)
			p.active_instances[1 + before + __popc(bal & laneMaskLt())] = (u32)i;
		__syncthreads();
		if(tid == 0)
			s_base += total;
		__syncthreads();
	}
#pragma unroll
	for(int o = 16; o > 0; o >>= 1)
		skipped_quads += __shfl_xor_sync(0xffffffffu, skipped_quads, o);
	if(lane == 0 && skipped_quads)
		atomicAdd(&p.info->num_input_quads, skipped_quads);
	if(tid == 0)
		p.active_instances[0] = (u32)s_base;
}

// quad_setup.glsl:77-128 -- screen AABB of a triangle that crosses the near plane (Blinn 1996)
__device__ float4 clippedAABB(float4 v0, float4 v1, float4 v2, float w0, float w1, float w2,
							  u32 clipmask) {
	float4 aabb = make_float4(1.0f, 1.0f, -1.0f, -1.0f);
	float4 v[3] = {v0, v1, v2};
	float iw[3] = {w0, w1, w2};
	int any_vis = 0;
	u32 or_mask = clipmask | (clipmask >> 8) | (clipmask >> 16);
#pragma unroll
	for(int i = 0; i < 3; i++) {
		u32 cm = clipmask >> (i * 8);
		if((cm & 0x3) == 0) {
			any_vis |= 0x1;
			if(v[i].x - aabb.x * v[i].w < 0.0f)
				aabb.x = v[i].x * iw[i];
			if(v[i].x - aabb.z * v[i].w > 0.0f)
				aabb.z = v[i].x * iw[i];
		}
		if((cm & 0xc) == 0) {
			any_vis |= 0x10;
			if(v[i].y - aabb.y * v[i].w < 0.0f)
				aabb.y = v[i].y * iw[i];
			if(v[i].y - aabb.w * v[i].w > 0.0f)
				aabb.w = v[i].y * iw[i];
		}
	}
	if((any_vis & 0x0f) == 0) {
		aabb.x = -1.0f, aabb.z = 1.0f;
	} else if((or_mask & 0x3) != 0) {
#pragma unroll
		for(int i = 0; i < 3; i++) {
			u32 cm = clipmask >> (i * 8);
			if((cm & 0x1) != 0 && v[i].x - aabb.x * v[i].w < 0.0f)
				aabb.x = -1.0f;
			if((cm & 0x2) != 0 && v[i].x - aabb.z * v[i].w > 0.0f)
				aabb.z = 1.0f;
		}
	}
	if((any_vis & 0xf0) == 0) {
		aabb.y = -1.0f, aabb.w = 1.0f;
	} else if((or_mask & 0xc) != 0) {
#pragma unroll
		for(int i = 0; i < 3; i++) {
			u32 cm = clipmask >> (i * 8);
			if((cm & 0x4) != 0 && v[i].y - aabb.y * v[i].w < 0.0f)
				aabb.y = -1.0f;
			if((cm & 0x8) != 0 && v[i].y - aabb.w * v[i].w > 0.0f)
				aabb.w = 1.0f;
		}
	}
	return aabb;
}

__device__ __forceinline__ float4 plainAABB(float4 a, float4 b, float4 c) {
	return make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y),
					   fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y));
}

struct QuadResult {
	int status;	   // -1 visible, -2 visible but outside the owned bin rows, else rejection type
	int size_type; // 0 small, 1 large
	u32 enc_aabb, y_aabb0, y_aabb1;
};

// quad_setup.glsl:136-254
__device__ QuadResult processInputQuad(const Params &p, const LucidConfig &cfg, uint4 vi) {
	QuadResult out;
	out.status = -1, out.size_type = 0, out.enc_aabb = 0, out.y_aabb0 = 0, out.y_aabb1 = 0;
	u32 v0 = vi.x, v1 = vi.y, v2 = vi.z, v3 = vi.w;
	bool cull0 = v0 == v1 || v1 == v2 || v2 == v0;
	bool cull1 = v0 == v2 || v2 == v3 || v3 == v0;
	// an index outside the vertex buffer rejects the quad instead of reading out of bounds (the reference
	// trusts its indices; valid inputs never take this branch)
	const u32 nv = (u32)p.num_verts;
	if((cull0 && cull1) || v0 >= nv || v1 >= nv || v2 >= nv || v3 >= nv) {
		out.status = LUCID_REJECTION_OTHER;
		return out;
	}
	F3 vws[4] = {vertexLoad(p, v0), vertexLoad(p, v1), vertexLoad(p, v2), vertexLoad(p, v3)};
	cull0 = cull0 || same3(vws[0], vws[1]) || same3(vws[1], vws[2]) || same3(vws[2], vws[0]);
	cull1 = cull1 || same3(vws[0], vws[2]) || same3(vws[2], vws[3]) || same3(vws[3], vws[0]);

	if(cfg.enable_backface_culling != 0) {
		F3 org = xyz(cfg.frustum.ws_origin0);
		F3 p0 = vws[0] - org, p1 = vws[1] - org, p2 = vws[2] - org, p3 = vws[3] - org;
		F3 nrm0 = cross3(p2, p1 - p2);
		F3 nrm1 = cross3(p3, p2 - p3);
		float volume0 = dot3(p0, nrm0), volume1 = dot3(p0, nrm1);
		cull0 = cull0 || volume0 <= 0.0f;
		cull1 = cull1 || volume1 <= 0.0f;
		if(cull0 && cull1) {
			out.status = LUCID_REJECTION_BACKFACE;
			return out;
		}
	}
	u32 cull_flags = (cull0 ? 1u : 0u) | (cull1 ? 2u : 0u);

	float4 vndc[4];
	const LucidVec4 *m = cfg.view_proj_matrix;
#pragma unroll
	for(int i = 0; i < 4; i++) {
		F3 q = vws[i];
		vndc[i].x = m[0].x * q.x + m[1].x * q.y + m[2].x * q.z + m[3].x;
		vndc[i].y = m[0].y * q.x + m[1].y * q.y + m[2].y * q.z + m[3].y;
		vndc[i].z = m[0].z * q.x + m[1].z * q.y + m[2].z * q.z + m[3].z;
		vndc[i].w = m[0].w * q.x + m[1].w * q.y + m[2].w * q.z + m[3].w;
	}
	u32 clipmask = vertexClipMask(vndc[0]) | (vertexClipMask(vndc[1]) << 8) |
				   (vertexClipMask(vndc[2]) << 16) | (vertexClipMask(vndc[3]) << 24);
	u32 and_mask = clipmask & (clipmask >> 8) & (clipmask >> 16) & (clipmask >> 24) & 0xffu;
	u32 or_mask = clipmask | (clipmask >> 8) | (clipmask >> 16) | (clipmask >> 24);
	if(and_mask != 0) {
		out.status = LUCID_REJECTION_FRUSTUM;
		return out;
	}

	float4 aabb0 = make_float4(0, 0, 0, 0), aabb1 = aabb0;
	float iw[4] = {rcp(vndc[0].w), rcp(vndc[1].w), rcp(vndc[2].w), rcp(vndc[3].w)};
	bool near_far = (or_mask & 0x30) != 0;
	if(near_far) {
		aabb0 = clippedAABB(vndc[0], vndc[1], vndc[2], iw[0], iw[1], iw[2], clipmask);
		aabb1 = clippedAABB(vndc[0], vndc[2], vndc[3], iw[0], iw[2], iw[3],
							(clipmask & 0xffu) | ((clipmask & 0xffff0000u) >> 8));
	}
#pragma unroll
	for(int i = 0; i < 4; i++) {
		vndc[i].x *= iw[i];
		vndc[i].y *= iw[i];
		vndc[i].z *= iw[i];
	}
	if(!near_far) {
		aabb0 = plainAABB(vndc[0], vndc[1], vndc[2]);
		aabb1 = plainAABB(vndc[0], vndc[2], vndc[3]);
	}

	float sx = float(p.width) * 0.5f, sy = float(p.height) * 0.5f;
	float mx = float(p.width - 1), my = float(p.height - 1);
	aabb0 = make_float4((aabb0.x + 1.0f) * sx, (aabb0.y + 1.0f) * sy, (aabb0.z + 1.0f) * sx,
						(aabb0.w + 1.0f) * sy);
	aabb1 = make_float4((aabb1.x + 1.0f) * sx, (aabb1.y + 1.0f) * sy, (aabb1.z + 1.0f) * sx,
						(aabb1.w + 1.0f) * sy);
	float4 aabb = make_float4(fminf(aabb0.x, aabb1.x), fminf(aabb0.y, aabb1.y),
							  fmaxf(aabb0.z, aabb1.z), fmaxf(aabb0.w, aabb1.w));

	if(ceilf(aabb.x - 0.5001f) == floorf(aabb.z + 0.5001f) ||
	   ceilf(aabb.y - 0.5001f) == floorf(aabb.w + 0.5001f)) {
		out.status = LUCID_REJECTION_BETWEEN_SAMPLES;
		return out;
	}
#define ADJ(a)                                                                                     \
	a = make_float4(clampf(a.x + 0.49f, 0.0f, mx), clampf(a.y + 0.49f, 0.0f, my),                  \
					clampf(a.z - 0.49f, 0.0f, mx), clampf(a.w - 0.49f, 0.0f, my))
	ADJ(aabb0);
	ADJ(aabb1);
	ADJ(aabb);
#undef ADJ
	u32 b0 = f2u(aabb.x) >> BIN_SHIFT, b1 = f2u(aabb.y) >> BIN_SHIFT;
	u32 b2 = f2u(aabb.z) >> BIN_SHIFT, b3 = f2u(aabb.w) >> BIN_SHIFT;
	out.enc_aabb = (b0 & 0x7fu) | ((b1 & 0x7fu) << 7) | ((b2 & 0x7fu) << 14) |
				   ((b3 & 0x7fu) << 21) | (cull_flags << 30);
	u32 bsx = b2 - b0 + 1u, bsy = b3 - b1 + 1u;
	out.size_type = bsx * bsy <= 4u ? 0 : 1;
	out.y_aabb0 = f2u(aabb0.y) | (f2u(aabb0.w) << 16);
	out.y_aabb1 = f2u(aabb1.y) | (f2u(aabb1.w) << 16);
	// bin-row split across devices: a quad is kept only where its bin rows intersect the owned
	// range; the small/large decision above used the unclamped AABB
	if((int)b3 < p.row_begin || (int)b1 >= p.row_end)
		out.status = -2;
	return out;
}

// quad_setup.glsl:274-340
__device__ void storeTri(const Params &p, const LucidConfig &cfg, u32 tri_idx, u32 flags_id,
						 u32 inst_color, F3 tri0, F3 tri1, F3 tri2, u32 y_aabb, F3 ray_dir0) {
	F3 normal = cross3(tri0 - tri2, tri1 - tri0);
	float multiplier = rcp(__fsqrt_rn(dot3(normal, normal)));
	normal = normal * multiplier;
	u32 enc_normal = (flags_id & LUCID_INST_HAS_VERTEX_NORMALS) ? 0u : encodeNormalUint(normal);

	F3 edge0 = (tri0 - tri2) * multiplier;
	F3 edge1 = (tri1 - tri0) * multiplier;
	float plane_dist = dot3(normal, tri0);
	F3 nrm_tri0 = cross3(tri0, normal);
	float param0 = dot3(edge0, nrm_tri0);
	float param1 = dot3(edge1, nrm_tri0);
	edge0 = cross3(normal, edge0);
	edge1 = cross3(normal, edge1);

	F3 dirx = xyz(cfg.frustum.ws_dirx), diry = xyz(cfg.frustum.ws_diry);
	F3 dir0 = xyz(cfg.frustum.ws_dir0);
	edge0 = mk3(dot3(edge0, dirx), dot3(edge0, diry), dot3(edge0, ray_dir0));
	edge1 = mk3(dot3(edge1, dirx), dot3(edge1, diry), dot3(edge1, ray_dir0));
	F3 pnormal = normal * rcp(plane_dist);
	F3 depth_eq = mk3(dot3(pnormal, dirx), dot3(pnormal, diry), dot3(pnormal, ray_dir0));

	TriShade sh;
	sh.depth = make_uint4(__float_as_uint(depth_eq.x), __float_as_uint(depth_eq.y),
						  __float_as_uint(depth_eq.z), flags_id);
	sh.bary0 = make_uint4(__float_as_uint(edge0.x), __float_as_uint(edge0.y),
						  __float_as_uint(edge0.z), __float_as_uint(param0));
	sh.bary1 = make_uint4(__float_as_uint(edge1.x), __float_as_uint(edge1.y),
						  __float_as_uint(edge1.z), __float_as_uint(param1));
	const bool constant = (flags_id & INST_VARYING_MASK) == 0;
	sh.misc = make_uint4(enc_normal, inst_color,
						 constant ? shadeConstant(cfg.lighting, flags_id, inst_color, enc_normal) : 0u,
						 constant ? 1u : 0u);
	uint4 *dst = reinterpret_cast<uint4 *>(p.tri_shade + tri_idx);
	dst[0] = sh.depth, dst[1] = sh.misc, dst[2] = sh.bary0, dst[3] = sh.bary1;

	F3 nrm0 = cross3(tri2, tri1 - tri2);
	F3 nrm1 = cross3(tri0, tri2 - tri0);
	F3 nrm2 = cross3(tri1, tri0 - tri1);
	float volume = dot3(tri0, nrm0);
	if(volume < 0.0f)
		nrm0 = -nrm0, nrm1 = -nrm1, nrm2 = -nrm2;
	F3 e0 = mk3(dot3(nrm0, dirx), dot3(nrm0, diry), dot3(nrm0, dir0));
	F3 e1 = mk3(dot3(nrm1, dirx), dot3(nrm1, diry), dot3(nrm1, dir0));
	F3 e2 = mk3(dot3(nrm2, dirx), dot3(nrm2, diry), dot3(nrm2, dir0));
	float ix0 = rcp(e0.x), ix1 = rcp(e1.x), ix2 = rcp(e2.x);
	F3 scan_base = -mk3(e0.z * ix0, e1.z * ix1, e2.z * ix2);
	F3 scan_step = -mk3(e0.y * ix0, e1.y * ix1, e2.y * ix2);
	u32 x_signs = (e0.x < 0.0f ? 1u : 0u) | (e1.x < 0.0f ? 2u : 0u) | (e2.x < 0.0f ? 4u : 0u);
	u32 y_signs = (e0.y < 0.0f ? 8u : 0u) | (e1.y < 0.0f ? 16u : 0u) | (e2.y < 0.0f ? 32u : 0u);
	F3 scan = mk3(scan_step.x * 0.5f + scan_base.x - (-0.5f), scan_step.y * 0.5f + scan_base.y - (-0.5f),
				  scan_step.z * 0.5f + scan_base.z - (-0.5f));
	uint4 *sdst = reinterpret_cast<uint4 *>(p.tri_scan + tri_idx);
	sdst[0] = make_uint4(__float_as_uint(scan.x), __float_as_uint(scan.y), __float_as_uint(scan.z),
						 y_aabb);
	sdst[1] = make_uint4(__float_as_uint(scan_step.x), __float_as_uint(scan_step.y),
						 __float_as_uint(scan_step.z), x_signs | y_signs);
}

// look-back word: [63:62] status, [61:31] small count, [30:0] large count
constexpr u64 LB_AGGREGATE = 1ull << 62, LB_PREFIX = 2ull << 62, LB_STATUS = 3ull << 62;
__device__ __forceinline__ u64 lbPack(u64 status, u32 small, u32 large) {
	return status | ((u64)small << 31) | (u64)large;
}
__device__ __forceinline__ u64 lbLoad(const u64 *ptr) {
	return *reinterpret_cast<const volatile u64 *>(ptr);
}
__device__ __forceinline__ void lbStore(u64 *ptr, u64 v) {
	*reinterpret_cast<volatile u64 *>(ptr) = v;
}

// k_quad_cull: culling, bin AABB, small/large split and the ordered compaction into visible-quad
// slots (quad_setup.glsl:136-254,374-489).  One 256-thread CTA per instance; a thread handles the
// quads tid, tid + 256, ... of the instance, so the loads of up to four quads are in flight
// together and one look-back step covers 1024 quads.  The per-triangle records are left to
// k_tri_setup, which runs over the compacted slots without any barrier.
#ifndef SETUP_CULL_MIN_CTAS
#define SETUP_CULL_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_CULL_MIN_CTAS)
	k_quad_cull(const Params p, const __grid_constant__ LucidConfig cfg) {
	__shared__ u32 s_vid;
	__shared__ int s_counts[SETUP_PARTS * (SETUP_THREADS / 32)][2]; // per (part, warp): small, large
	__shared__ int s_base[2];
	__shared__ u32 s_rejected[LUCID_REJECTION_TYPE_COUNT];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdlEntry();
	PhaseTimer timer = timerStart(p); // setup_timers: 0 init & finish, 1 process input quads, 2 store tri data, 3 store quad data
	// CTAs are ordered by ticket, not by blockIdx, so every predecessor in the look-back chain is
	// guaranteed to be running or finished.
	if(tid == 0)
		s_vid = atomicAdd(p.setup_ticket, 1u);
	if(tid < LUCID_REJECTION_TYPE_COUNT)
		s_rejected[tid] = 0;
	__syncthreads();
	// instance culling: the chain only runs over the instances k_instance_select kept
	const u32 vid = s_vid;
	if(p.active_instances != nullptr && vid >= __ldcg(p.active_instances))
		return;
	const u32 inst_id = p.active_instances != nullptr ? __ldcg(p.active_instances + 1 + vid) : vid;
	const LucidInstanceData inst = p.instances[inst_id];
	if(tid == 0)
		atomicAdd(&p.info->num_input_quads, inst.num_quads);
	constexpr bool skip = false;

	timerMark(timer, p.info->setup_timers, 0);
	uint4 vi[SETUP_PARTS];
	QuadResult res[SETUP_PARTS];
	// one 128-bit load per quad: the index buffer is 4 x u32 per quad (quad_setup.glsl:405-409)
	const uint4 *ib = reinterpret_cast<const uint4 *>(reinterpret_cast<const u32 *>(p.quad_indices) + inst.index_offset);
#pragma unroll
	for(int k = 0; k < SETUP_PARTS; k++) {
		const int local_quad = k * SETUP_THREADS + tid;
		vi[k] = make_uint4(0, 0, 0, 0);
		if(local_quad < inst.num_quads && !skip) {
			vi[k] = __ldg(ib + local_quad);
			vi[k].x += inst.vertex_offset, vi[k].y += inst.vertex_offset;
			vi[k].z += inst.vertex_offset, vi[k].w += inst.vertex_offset;
		}
	}
	int before_small[SETUP_PARTS], before_large[SETUP_PARTS];
#pragma unroll
	for(int k = 0; k < SETUP_PARTS; k++) {
		const int local_quad = k * SETUP_THREADS + tid;
		res[k].status = -3, res[k].size_type = 0, res[k].enc_aabb = 0, res[k].y_aabb0 = res[k].y_aabb1 = 0;
		if(local_quad < inst.num_quads)
			res[k].status = -2;
		if(local_quad < inst.num_quads && !skip)
			res[k] = processInputQuad(p, cfg, vi[k]);
		const bool vis = res[k].status == -1;
		const u32 bs = __ballot_sync(0xffffffffu, vis && res[k].size_type == 0);
		const u32 bl = __ballot_sync(0xffffffffu, vis && res[k].size_type == 1);
		// rejection counters: one ballot per rejection type
#pragma unroll
		for(int t = 0; t < LUCID_REJECTION_TYPE_COUNT; t++) {
			const u32 m = __ballot_sync(0xffffffffu, res[k].status == t);
			if(m != 0 && lane == 0)
				atomicAdd(&s_rejected[t], __popc(m));
		}
		if(lane == 0)
			s_counts[k * (SETUP_THREADS / 32) + warp][0] = __popc(bs), s_counts[k * (SETUP_THREADS / 32) + warp][1] = __popc(bl);
		before_small[k] = __popc(bs & laneMaskLt()), before_large[k] = __popc(bl & laneMaskLt());
	}
	timerMark(timer, p.info->setup_timers, 1);
	__syncthreads();

	// warp 0: exclusive scan of the 32 (part, warp) counts in input order, then the decoupled
	// look-back over the CTAs' (small, large) totals
	if(warp == 0) {
		static_assert(SETUP_PARTS * (SETUP_THREADS / 32) == 32, "one count pair per lane");
		int cs = s_counts[lane][0], cl = s_counts[lane][1];
		int is = cs, il = cl;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			int ts = __shfl_up_sync(0xffffffffu, is, o), tl = __shfl_up_sync(0xffffffffu, il, o);
			if(lane >= o)
				is += ts, il += tl;
		}
		s_counts[lane][0] = is - cs, s_counts[lane][1] = il - cl;
		const u32 total_small = __shfl_sync(0xffffffffu, is, 31), total_large = __shfl_sync(0xffffffffu, il, 31);

		u64 *lb = p.setup_lookback;
		u32 ex_small = 0, ex_large = 0;
		if(vid == 0) {
			if(lane == 0)
				lbStore(lb, lbPack(LB_PREFIX, total_small, total_large));
		} else {
			if(lane == 0)
				lbStore(lb + vid, lbPack(LB_AGGREGATE, total_small, total_large));
			int base = (int)vid - 1;
			while(true) {
				int idx = base - lane;
				u64 st = idx >= 0 ? lbLoad(lb + idx) : lbPack(LB_PREFIX, 0, 0);
				while(__any_sync(0xffffffffu, (st & LB_STATUS) == 0)) {
					if((st & LB_STATUS) == 0)
						st = lbLoad(lb + idx);
				}
				u32 pm = __ballot_sync(0xffffffffu, (st & LB_STATUS) == LB_PREFIX);
				int first = pm ? __ffs(pm) - 1 : 32;
				u32 ps = lane <= first ? (u32)((st >> 31) & 0x7fffffffu) : 0u;
				u32 pl = lane <= first ? (u32)(st & 0x7fffffffu) : 0u;
#pragma unroll
				for(int o = 16; o > 0; o >>= 1) {
					ps += __shfl_xor_sync(0xffffffffu, ps, o);
					pl += __shfl_xor_sync(0xffffffffu, pl, o);
				}
				ex_small += ps, ex_large += pl;
				if(pm)
					break;
				base -= 32;
			}
			if(lane == 0)
				lbStore(lb + vid, lbPack(LB_PREFIX, ex_small + total_small, ex_large + total_large));
		}
		if(lane == 0)
			s_base[0] = ex_small, s_base[1] = ex_large;
	}
	__syncthreads();

	// slot assignment; quads past MAX_VISIBLE_QUADS (in input order) are dropped and counted
	const int mvq = p.max_visible_quads;
	int n_small = 0, n_large = 0, n_dropped = 0;
#pragma unroll
	for(int k = 0; k < SETUP_PARTS; k++) {
		const bool vis = res[k].status == -1;
		if(!vis)
			continue;
		const bool is_small = res[k].size_type == 0;
		const int w = k * (SETUP_THREADS / 32) + warp;
		int gs = s_base[0] + s_counts[w][0] + before_small[k], gl = s_base[1] + s_counts[w][1] + before_large[k];
		if(gs + gl < mvq) {
			const int slot = is_small ? gs : (mvq - 1) - gl;
			p.quad_aabbs[slot] = res[k].enc_aabb;
			p.quad_verts[slot] = vi[k];
			p.quad_setup_info[slot] = make_uint4(res[k].y_aabb0, res[k].y_aabb1, inst_id, 0u);
			n_small += is_small ? 1 : 0, n_large += is_small ? 0 : 1;
		} else {
			n_dropped++;
		}
	}
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		n_small += __shfl_xor_sync(0xffffffffu, n_small, o);
		n_large += __shfl_xor_sync(0xffffffffu, n_large, o);
		n_dropped += __shfl_xor_sync(0xffffffffu, n_dropped, o);
	}
	if(lane == 0) {
		if(n_small)
			atomicAdd(&p.info->num_visible_quads[0], n_small);
		if(n_large)
			atomicAdd(&p.info->num_visible_quads[1], n_large);
		if(n_dropped)
			atomicAdd(&p.info->temp[0], n_dropped);
	}
	if(tid < LUCID_REJECTION_TYPE_COUNT && s_rejected[tid] != 0)
		atomicAdd(&p.info->num_rejected_quads[tid], s_rejected[tid]);
	timerMark(timer, p.info->setup_timers, 0);
}

// k_tri_setup: one thread per triangle of a visible quad (storeTri / storeQuad,
// quad_setup.glsl:256-340): plane, barycentric and scanline equations, attribute repack
#ifndef SETUP_TRI_MIN_CTAS
#define SETUP_TRI_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(SETUP_THREADS, SETUP_TRI_MIN_CTAS) k_tri_setup(const Params p, const __grid_constant__ LucidConfig cfg) {
	pdlEntry();
	const int n_small = p.info->num_visible_quads[0], n_large = p.info->num_visible_quads[1];
	const int n_tris = (n_small + n_large) * 2;
	const F3 dir0 = xyz(cfg.frustum.ws_dir0), dirx = xyz(cfg.frustum.ws_dirx);
	const F3 diry = xyz(cfg.frustum.ws_diry), origin = xyz(cfg.frustum.ws_origin0);
	const F3 ray_dir0 = dir0 + (dirx + diry) * 0.5f;
	PhaseTimer timer = timerStart(p);
	for(int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tris; t += gridDim.x * blockDim.x) {
		const int q = t >> 1, second = t & 1;
		const int slot = q < n_small ? q : (p.max_visible_quads - 1) - (q - n_small);
		const u32 enc_aabb = p.quad_aabbs[slot];
		const uint4 v = p.quad_verts[slot];
		const uint4 qi = p.quad_setup_info[slot];
		const u32 inst_id = qi.z;
		const u32 flags = __ldg(&p.instances[inst_id].flags);
		if(((enc_aabb >> (30 + second)) & 1) == 0) {
			u32 i1 = second ? v.z : v.y, i2 = second ? v.w : v.z;
			F3 t0 = vertexLoad(p, v.x) - origin;
			F3 t1 = vertexLoad(p, i1) - origin;
			F3 t2 = vertexLoad(p, i2) - origin;
			storeTri(p, cfg, (u32)slot * 2 + second, flags | (inst_id << 16), __ldg(p.inst_colors + inst_id), t0, t1,
					 t2, second ? qi.y : qi.x, ray_dir0);
		}
		timerMark(timer, p.info->setup_timers, 2);
		if(second)
			continue;
		// quad records (quad_setup.glsl:256-272, 342-354)
		if((flags & LUCID_INST_HAS_VERTEX_COLORS) && p.vertex_colors)
			p.quad_colors[slot] = make_uint4(__ldg(p.vertex_colors + v.x), __ldg(p.vertex_colors + v.y),
											 __ldg(p.vertex_colors + v.z), __ldg(p.vertex_colors + v.w));
		if((flags & LUCID_INST_HAS_VERTEX_NORMALS) && p.vertex_normals)
			p.quad_normals[slot] = make_uint4(__ldg(p.vertex_normals + v.x), __ldg(p.vertex_normals + v.y),
											  __ldg(p.vertex_normals + v.z), __ldg(p.vertex_normals + v.w));
		if((flags & LUCID_INST_HAS_ALBEDO_TEXTURE) && p.vertex_uvs) {
			float2 t0 = __ldg(p.vertex_uvs + v.x), t1 = __ldg(p.vertex_uvs + v.y);
			float2 t2 = __ldg(p.vertex_uvs + v.z), t3 = __ldg(p.vertex_uvs + v.w);
			p.quad_uv[(size_t)slot * 2 + 0] = make_uint4(__float_as_uint(t0.x), __float_as_uint(t0.y),
														 __float_as_uint(t1.x - t0.x), __float_as_uint(t1.y - t0.y));
			p.quad_uv[(size_t)slot * 2 + 1] =
				make_uint4(__float_as_uint(t2.x - t0.x), __float_as_uint(t2.y - t0.y), __float_as_uint(t3.x - t0.x),
						   __float_as_uint(t3.y - t0.y));
		}
		timerMark(timer, p.info->setup_timers, 3);
	}
}

// Start of a frame (the clears of setupInputData, lucid_renderer.cpp:437): zeroes LucidInfo, the
// first six per-bin counter arrays, the setup look-back state and the row costs.  A kernel instead
// of cudaMemsetAsync keeps the copy engines free for the image read-back of the previous frame and
// the instance upload of the next one, which run concurrently on their own streams.
__global__ void __launch_bounds__(256) k_frame_begin(const Params p) {
	const int stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
	pdlEntry();
	u32 *info = reinterpret_cast<u32 *>(p.info);
	const int n_clear = (int)LUCID_INFO_U32_SIZE + p.bin_count * 6; // lucid_renderer.cpp:437
	for(int i = first; i < n_clear; i += stride)
		info[i] = 0;
	for(int i = first; i < p.num_setup_ctas; i += stride)
		p.setup_lookback[i] = 0;
	for(int i = first; i < p.bin_count; i += stride)
		p.bin_cost[i] = 0;
	if(first == 0)
		*p.setup_ticket = 0, p.tie_runs[0] = 0;
}

// End of a frame: LucidInfo and the per-bin arrays go to the pinned read-back buffer
// (m_info download, lucid_renderer.cpp:341-346) as posted PCIe writes
__global__ void __launch_bounds__(256) k_info_out(const Params p, u32 *host_info, int num_words) {
	const u32 *info = reinterpret_cast<const u32 *>(p.info);
	pdlEntry();
	for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_words; i += gridDim.x * blockDim.x)
		host_info[i] = info[i];
}

void launchFrameBegin(const Params &p, cudaStream_t stream) { launchPDL(k_frame_begin, 64, 256, 0, stream, p); }
void launchInfoOut(const Params &p, u32 *host_info, int num_words, cudaStream_t stream) {
	launchPDL(k_info_out, 32, 256, 0, stream, p, host_info, num_words);
}

void launchQuadSetup(const Params &p, const LucidConfig &cfg, cudaStream_t stream) {
	if(p.num_setup_ctas == 0)
		return;
	if(p.active_instances != nullptr)
		launchPDL(k_instance_select, 1, 1024, 0, stream, p, cfg);
	launchPDL(k_quad_cull, p.num_setup_ctas, SETUP_THREADS, 0, stream, p, cfg);
	launchPDL(k_tri_setup, 148 * 2 * SETUP_TRI_MIN_CTAS, SETUP_THREADS, 0, stream, p, cfg);
}

} // namespace lucid
