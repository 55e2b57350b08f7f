// comparators.cu -- SURVEY 8 f4: what the frame the renderer has just drawn would look like under the techniques
// LucidRaster is compared with (docs/readme.md:7-8) -- hardware alpha blending in submission order, as the
// reference's SimpleRenderer does it with the fixed-function pipeline (src/simple_renderer.cpp:69-132,134-196), and
// two approximate order-independent techniques.  No graphics API is involved: the kernels re-reduce the SAMPLES of
// the exact frame -- the sorted-entry stream k_block_sort left in memory holds every (triangle, half-block) pair with
// its pixel mask, depth plane and constant colour -- with another per-pixel rule.  The result answers "how wrong
// would the cheaper technique be on this frame", with the same coverage, the same sample colours and the same
// depths as the exact image next to it; it says nothing about how fast raster-operation hardware would be.
//
// Every mode starts with the reference's opaque phase (renderPhase(opaque = true): depth test `less` + depth write,
// no blending): the nearest sample of an INST_IS_OPAQUE instance gives the pixel's base colour and the depth zo the
// other samples are tested against (first submitted wins among equal depths).  Transparent samples that are
// strictly nearer than zo are then visited in SUBMISSION order (instance, quad of the instance, triangle):
//   LUCID_COMPARE_HW_BLEND  src_alpha / one_minus_src_alpha (src_alpha / one under ADDITIVE_BLENDING) on an 8-bit
//                           unorm target: every blend reads the target's bytes and rounds back to bytes
//   LUCID_COMPARE_WBOIT     weighted blended OIT (McGuire & Bavoil 2013, weight of eq. 7 on the ray position)
//   LUCID_COMPARE_MLAB4     multi-layer alpha blending (Salvi & Vaidyanathan 2014), four layers
// The CPU checker has the same three reductions (oracle/lucid_oracle.cpp comparePixel); images compare bit for bit.
//
// Submission order is not kept by the pipeline (visible-quad slots are compacted per size class), so
// k_submission_order recovers it: a warp per visible quad looks the quad's vertex indices up in its instance's
// index list.  k_compare: warp per work item, lane = pixel; the item's entries are sorted by submission order
// (64-bit keys, bitonic network in shared memory, lists over 1024 entries in an L2-resident scratch) and walked
// twice, once for the opaque phase and once for the transparent one.
#include "raster_common.cuh"

namespace lucid {

constexpr int CMP_WARPS = 4;		 // warps per CTA of k_compare
constexpr int CMP_SMEM_KEYS = 1024;	 // keys per warp sorted in shared memory
constexpr int CMP_CTAS_PER_SM = 2;
constexpr int CMP_SCRATCH_KEYS = MAX_HBLOCK_TRIS; // per warp, for longer lists

__global__ void __launch_bounds__(256) k_compare_fill(u32 *dst, size_t n, u32 value) {
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		dst[i] = value;
}

// order[slot] = instance << 10 | quad of the instance (the first one with the slot's vertex indices: quads of one
// instance with identical indices are identical, so their mutual order cannot change an image)
__global__ void __launch_bounds__(256) k_submission_order(const Params p, u32 *order) {
	const int lane = laneId();
	const int num_warps = (int)((gridDim.x * blockDim.x) >> 5), first = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int n_small = p.info->num_visible_quads[0], n_large = p.info->num_visible_quads[1];
	for(int q = first; q < n_small + n_large; q += num_warps) {
		const int slot = q < n_small ? q : (p.max_visible_quads - 1) - (q - n_small);
		const u32 inst_id = p.quad_setup_info[slot].z;
		const uint4 v = p.quad_verts[slot];
		const LucidInstanceData inst = p.instances[inst_id];
		const uint4 *ib = reinterpret_cast<const uint4 *>(reinterpret_cast<const u32 *>(p.quad_indices) + inst.index_offset);
		const u32 vo = (u32)inst.vertex_offset;
		u32 found = 0;
		for(int base = 0; base < inst.num_quads; base += 32) {
			const int l = base + lane;
			bool match = false;
			if(l < inst.num_quads) {
				const uint4 t = __ldg(ib + l);
				match = t.x + vo == v.x && t.y + vo == v.y && t.z + vo == v.z && t.w + vo == v.w;
			}
			const u32 m = __ballot_sync(0xffffffffu, match);
			if(m != 0) {
				found = (u32)(base + __ffs(m) - 1);
				break;
			}
		}
		if(lane == 0)
			order[slot] = (inst_id << 10) | found;
	}
}

// ascending bitonic sort of `padded` (a power of two) 64-bit keys by one warp, in shared memory or -- GLOBAL -- in an
// L2-resident array (accesses bypass L1, as in warpSortLarge)
template <bool GLOBAL> __device__ __forceinline__ u64 keyLoad(const u64 *k) {
	return GLOBAL ? __ldcg(reinterpret_cast<const unsigned long long *>(k)) : *k;
}
template <bool GLOBAL> __device__ __forceinline__ void keyStore(u64 *k, u64 v) {
	if(GLOBAL)
		__stcg(reinterpret_cast<unsigned long long *>(k), (unsigned long long)v);
	else
		*k = v;
}
template <bool GLOBAL> __device__ __noinline__ void warpSort64(u64 *keys, int padded) {
	const int lane = laneId();
	for(int k = 2; k <= padded; k <<= 1)
		for(int j = k >> 1; j > 0; j >>= 1) {
			for(int i = lane; i < (padded >> 1); i += 32) {
				const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
				const bool up = (lo & k) == 0;
				const u64 a = keyLoad<GLOBAL>(keys + lo), b = keyLoad<GLOBAL>(keys + hi);
				if((a > b) == up)
					keyStore<GLOBAL>(keys + lo, b), keyStore<GLOBAL>(keys + hi, a);
			}
			__syncwarp();
		}
}

// the RGBA8 sample of entry (tri, aux) at a pixel: the constant that came with the stream, or shadeSample
__device__ __forceinline__ u32 compareShade(const Params &p, const ColourTables &tab, const LightTerms &lt, u32 tri_idx, uint4 aux,
											float px, float py) {
	if(aux.w != AUX_VARYING)
		return aux.w;
	uint4 e[STAGE_WORDS];
#pragma unroll
	for(int i = 0; i < STAGE_WORDS; i++)
		e[i] = make_uint4(0, 0, 0, 0);
	stageEntry(p, lt, tri_idx, e);
	return shadeStaged(p, tab, lt, e, __uint_as_float(aux.x), __uint_as_float(aux.y), __uint_as_float(aux.z), px, py);
}

__device__ __forceinline__ u32 quant8(float v) { return f2u(saturatef(v) * 255.0f + 0.5f); }

// key: submission order of the triangle (27 bits) << 13 | opaque << 12 | entry of the item (12 bits)
constexpr u64 KEY_OPAQUE = 1ull << 12;
constexpr u32 KEY_ENTRY_MASK = 0xfffu;

__global__ void __launch_bounds__(CMP_WARPS * 32, CMP_CTAS_PER_SM)
	k_compare(const __grid_constant__ Params p, const __grid_constant__ LucidConfig cfg, const int mode, const u32 *order,
			  u64 *scratch, u32 *ticket, u32 *out_image, const u32 bg8) {
	extern __shared__ __align__(16) unsigned char smem[];
	__shared__ float2 s_s2l[LUCID_S2L_SIZE], s_l2s[LUCID_L2S_SIZE];
	const int lane = laneId(), warp = threadIdx.x >> 5;
	for(int i = threadIdx.x; i < LUCID_S2L_SIZE; i += CMP_WARPS * 32)
		s_s2l[i] = reinterpret_cast<const float2 *>(d_s2l_words)[i];
	for(int i = threadIdx.x; i < LUCID_L2S_SIZE; i += CMP_WARPS * 32)
		s_l2s[i] = reinterpret_cast<const float2 *>(d_l2s_words)[i];
	__syncthreads();
	ColourTables tab;
	tab.s2l = s_s2l, tab.l2s = s_l2s;
	const LightTerms lt = lightTerms(cfg.lighting);
	const bool additive = (p.opts & LUCID_OPT_ADDITIVE_BLENDING) != 0;
	u64 *skeys = reinterpret_cast<u64 *>(smem) + (size_t)warp * CMP_SMEM_KEYS;
	u64 *gkeys = scratch + ((size_t)blockIdx.x * CMP_WARPS + warp) * CMP_SCRATCH_KEYS;
	const float neg_inf = __int_as_float(0xff800000);

	u32 class_end[ITEM_CLASSES];
	{
		u32 acc = 0;
#pragma unroll
		for(int k = 0; k < ITEM_CLASSES; k++)
			class_end[k] = acc += p.work_counters[WC_CLASS + k];
	}
	const u32 n_items = class_end[ITEM_CLASSES - 1];

	while(true) {
		uint4 entry = make_uint4(0, 0, 0, 0);
		if(lane == 0) {
			const u32 i = atomicAdd(ticket, 1u);
			if(i < n_items)
				entry = fetchWorkItem(p, i, class_end);
		}
		const u32 item = __shfl_sync(0xffffffffu, entry.x, 0);
		const int count = (int)__shfl_sync(0xffffffffu, entry.y, 0);
		const u32 offset = __shfl_sync(0xffffffffu, entry.z, 0);
		if(count == 0)
			break;
		const int bin_id = (int)(item >> 6), sub = (int)(item & 31u);
		const bool high = (item & 32u) != 0;
		const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
		const uint4 *src_rec = p.sorted_rec + offset, *src_aux = p.sorted_aux + offset;

		// the item's entries in submission order
		int padded = 32;
		while(padded < count)
			padded <<= 1;
		const bool in_global = padded > CMP_SMEM_KEYS;
		u64 *keys = in_global ? gkeys : skeys;
		for(int e = lane; e < padded; e += 32) {
			u64 key = ~0ull;
			if(e < count) {
				const u32 tri = __ldcg(src_rec + e).x;
				const u32 o = __ldg(order + (tri >> 1));
				const u32 flags = __ldg(&p.instances[o >> 10].flags);
				key = ((u64)((o << 1) | (tri & 1u)) << 13) | ((flags & LUCID_INST_IS_OPAQUE) ? KEY_OPAQUE : 0ull) | (u64)e;
			}
			if(in_global)
				keyStore<true>(keys + e, key);
			else
				keys[e] = key;
		}
		__syncwarp();
		if(in_global)
			warpSort64<true>(keys, padded);
		else
			warpSort64<false>(keys, padded);

		const int hb_x = bin_x * BIN_SIZE + (sub & 3) * 8;
		const int y0 = bin_y * BIN_SIZE + (sub >> 2) * (high ? 4 : 8);
		const int halves = high ? 1 : 2;
		for(int half = 0; half < halves; half++) {
			const int hb_y = y0 + half * 4;
			const float fpx = float(hb_x + (lane & 7)), fpy = float(hb_y + (lane >> 3));

			// opaque phase: nearest opaque sample, the first submitted among equal depths
			float zo = neg_inf;
			int win = -1;
			for(int i = 0; i < count; i++) {
				const u64 key = in_global ? keyLoad<true>(keys + i) : keys[i];
				if(!(key & KEY_OPAQUE))
					continue;
				const int e = (int)((u32)key & KEY_ENTRY_MASK);
				const uint4 rec = __ldcg(src_rec + e);
				const u32 bits = half ? rec.z : rec.y;
				if(!((bits >> lane) & 1u))
					continue;
				const uint4 aux = __ldcg(src_aux + e);
				const float depth = __uint_as_float(aux.x) * fpx + (__uint_as_float(aux.y) * fpy + __uint_as_float(aux.z));
				if(depth > zo)
					zo = depth, win = e;
			}
			u32 base8 = bg8;
			if(win >= 0)
				base8 = compareShade(p, tab, lt, __ldcg(src_rec + win).x, __ldcg(src_aux + win), fpx, fpy);

			// transparent phase
			u32 dst8 = base8;									 // HW_BLEND
			float acc_r = 0.0f, acc_g = 0.0f, acc_b = 0.0f, acc_a = 0.0f, trans = 1.0f; // WBOIT
			float lr0 = 0.0f, lr1 = 0.0f, lr2 = 0.0f, lr3 = 0.0f, lg0 = 0.0f, lg1 = 0.0f, lg2 = 0.0f, lg3 = 0.0f; // MLAB4
			float lb0 = 0.0f, lb1 = 0.0f, lb2 = 0.0f, lb3 = 0.0f, lt0 = 1.0f, lt1 = 1.0f, lt2 = 1.0f, lt3 = 1.0f;
			float ld0 = neg_inf, ld1 = neg_inf, ld2 = neg_inf, ld3 = neg_inf;
			for(int i = 0; i < count; i++) {
				const u64 key = in_global ? keyLoad<true>(keys + i) : keys[i];
				if(key & KEY_OPAQUE)
					continue;
				const int e = (int)((u32)key & KEY_ENTRY_MASK);
				const uint4 rec = __ldcg(src_rec + e);
				const u32 bits = half ? rec.z : rec.y;
				if(!((bits >> lane) & 1u))
					continue;
				const uint4 aux = __ldcg(src_aux + e);
				const float depth = __uint_as_float(aux.x) * fpx + (__uint_as_float(aux.y) * fpy + __uint_as_float(aux.z));
				if(!(depth > zo))
					continue;
				const u32 color = compareShade(p, tab, lt, rec.x, aux, fpx, fpy);
				if(color == 0)
					continue;
				const float4 c = decodeRGBA8(color);
				if(mode == LUCID_COMPARE_HW_BLEND) {
					const float4 d = decodeRGBA8(dst8);
					float r, g, b;
					if(additive) {
						r = __fmaf_rn(c.x, c.w, d.x), g = __fmaf_rn(c.y, c.w, d.y), b = __fmaf_rn(c.z, c.w, d.z);
					} else {
						const float keep = 1.0f - c.w;
						r = __fmaf_rn(c.x, c.w, d.x * keep), g = __fmaf_rn(c.y, c.w, d.y * keep), b = __fmaf_rn(c.z, c.w, d.z * keep);
					}
					dst8 = quant8(r) | (quant8(g) << 8) | (quant8(b) << 16);
				} else if(mode == LUCID_COMPARE_WBOIT) {
					const float z = rcp(depth);
					const float z5 = z * 0.2f, t2 = z5 * z5;
					const float z200 = z * 0.005f, s2 = z200 * z200, s6 = (s2 * s2) * s2;
					const float den = (1e-5f + t2) + s6;
					const float wz = fminf(fmaxf(__fdiv_rn(10.0f, den), 1e-2f), 3e3f);
					const float w = c.w * wz, aw = c.w * w;
					acc_r = __fmaf_rn(c.x, aw, acc_r), acc_g = __fmaf_rn(c.y, aw, acc_g), acc_b = __fmaf_rn(c.z, aw, acc_b);
					acc_a = acc_a + aw;
					trans = __fmaf_rn(-c.w, trans, trans);
				} else {
					float fr = c.x * c.w, fg = c.y * c.w, fb = c.z * c.w, ft = 1.0f - c.w, fd = depth;
#define CMP_LAYER(R, G, B, T, D)                                                                                       \
	if(fd > D) {                                                                                                       \
		float t_;                                                                                                      \
		t_ = fr, fr = R, R = t_;                                                                                       \
		t_ = fg, fg = G, G = t_;                                                                                       \
		t_ = fb, fb = B, B = t_;                                                                                       \
		t_ = ft, ft = T, T = t_;                                                                                       \
		t_ = fd, fd = D, D = t_;                                                                                       \
	}
					CMP_LAYER(lr0, lg0, lb0, lt0, ld0)
					CMP_LAYER(lr1, lg1, lb1, lt1, ld1)
					CMP_LAYER(lr2, lg2, lb2, lt2, ld2)
					CMP_LAYER(lr3, lg3, lb3, lt3, ld3)
#undef CMP_LAYER
					lr3 = __fmaf_rn(fr, lt3, lr3), lg3 = __fmaf_rn(fg, lt3, lg3), lb3 = __fmaf_rn(fb, lt3, lb3);
					lt3 = lt3 * ft;
				}
			}

			u32 out = dst8;
			if(mode != LUCID_COMPARE_HW_BLEND) {
				const float4 base = decodeRGBA8(base8);
				float r, g, b;
				if(mode == LUCID_COMPARE_WBOIT) {
					const float den = fmaxf(acc_a, 1e-5f), show = 1.0f - trans;
					r = __fmaf_rn(__fdiv_rn(acc_r, den), show, base.x * trans);
					g = __fmaf_rn(__fdiv_rn(acc_g, den), show, base.y * trans);
					b = __fmaf_rn(__fdiv_rn(acc_b, den), show, base.z * trans);
				} else {
					float t = 1.0f;
					r = g = b = 0.0f;
					r = __fmaf_rn(lr0, t, r), g = __fmaf_rn(lg0, t, g), b = __fmaf_rn(lb0, t, b), t = t * lt0;
					r = __fmaf_rn(lr1, t, r), g = __fmaf_rn(lg1, t, g), b = __fmaf_rn(lb1, t, b), t = t * lt1;
					r = __fmaf_rn(lr2, t, r), g = __fmaf_rn(lg2, t, g), b = __fmaf_rn(lb2, t, b), t = t * lt2;
					r = __fmaf_rn(lr3, t, r), g = __fmaf_rn(lg3, t, g), b = __fmaf_rn(lb3, t, b), t = t * lt3;
					r = __fmaf_rn(base.x, t, r), g = __fmaf_rn(base.y, t, g), b = __fmaf_rn(base.z, t, b);
				}
				out = quant8(r) | (quant8(g) << 8) | (quant8(b) << 16);
			}
			const int gx = hb_x + (lane & 7), gy = hb_y + (lane >> 3);
			if(gx < p.width && gy < p.height)
				out_image[(size_t)gy * p.width + gx] = out | 0xff000000u;
		}
		__syncwarp(); // the keys are rewritten for the next item
	}
}

size_t compareScratchKeys(int num_sms) { return (size_t)num_sms * CMP_CTAS_PER_SM * CMP_WARPS * CMP_SCRATCH_KEYS; }

// order: max_visible_quads words; scratch: compareScratchKeys() 64-bit words; ticket: one word; out_image: width x height
void launchCompare(const Params &p, const LucidConfig &cfg, int mode, u32 *order, u64 *scratch, u32 *ticket, u32 *out_image,
				   cudaStream_t stream, int num_sms) {
	static std::once_flag configured[64];
	constexpr int smem = CMP_WARPS * CMP_SMEM_KEYS * 8;
	oncePerDevice(configured, [] { cudaFuncSetAttribute(k_compare, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
	const LucidVec4 &bg = cfg.background_color;
	auto q = [](float v) { return (u32)(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f + 0.5f); };
	const u32 bg8 = q(bg.x) | (q(bg.y) << 8) | (q(bg.z) << 16) | 0xff000000u;
	cudaMemsetAsync(ticket, 0, 4, stream);
	k_compare_fill<<<num_sms * 4, 256, 0, stream>>>(out_image, (size_t)p.width * p.height, bg8);
	k_submission_order<<<num_sms * 8, 256, 0, stream>>>(p, order);
	k_compare<<<num_sms * CMP_CTAS_PER_SM, CMP_WARPS * 32, smem, stream>>>(p, cfg, mode, order, scratch, ticket, out_image, bg8);
}

} // namespace lucid
