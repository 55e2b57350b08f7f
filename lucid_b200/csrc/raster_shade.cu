// raster_shade.cu -- stage 3 of the raster pipeline: sample shading and the exact per-pixel blend
// (unpackSamples + shadeAndReduceSamples + finishReduceSamples of the reference, raster.glsl:292-396,
// shading.glsl:107-314) over the depth-sorted entry stream k_block_sort left behind.
//
// One warp per work item (an 8x4 half-block of a HIGH bin, or the two halves of an 8x8 block of a LOW bin),
// lane = pixel.  The item's entries are consumed in chunks of up to 32 entries / 256 samples:
//   * the entry stream -- two contiguous planes of 16-byte words -- is staged in a shared-memory ring by 1-D bulk
//     async copies (cp.async.bulk + mbarrier complete_tx), four 16-entry blocks deep, issued by one lane; the
//     LUCID_SHADE_STREAM=ldg build of the same kernel reads it with coalesced 128-bit loads one chunk ahead;
//   * a 32x32 bit transpose of the chunk's pixel masks tells every pixel lane which entries cover it, in depth
//     order;
//   * chunks of constant-colour triangles are consumed by the pixel lanes directly (depth plane and colour come
//     with the stream); otherwise every entry's shading inputs are gathered once into a 128-byte stage, the
//     chunk's samples are shaded one per lane from the stage, and the pixel lanes pick the colours up;
//   * 3-entry window and blend per pixel; a pixel whose transmittance reached zero takes exactly +0 from every
//     later sample, so its samples are neither shaded nor reduced, and a half-block stops when all are there.
// The reference's per-segment saturate (raster.glsl:394-395) cannot change the stored pixel (the accumulators never
// decrease and the final value is saturated); segments only matter for the ALPHA_THRESHOLD early out, which has
// its own segment-accurate instantiation of the kernel.
#include "raster_common.cuh"

namespace lucid {

#ifndef SHADE_MIN_CTAS
#define SHADE_MIN_CTAS 7
#endif
constexpr int CHUNK_SAMPLES = 256;
constexpr int CHUNK_ENTRIES = 32;
constexpr int RING_BLOCK = 16;	// entries per bulk copy
constexpr int RING_BLOCKS = 4;	// blocks in the ring
constexpr int RING_ENTRIES = RING_BLOCK * RING_BLOCKS;

// per-warp shared memory
struct WarpMem {
	uint4 *stage;	// CHUNK_ENTRIES * STAGE_WORDS: per-entry shading inputs
	uint4 *plane;	// CHUNK_ENTRIES (ldg build) or the ring's second plane: depth plane xyz, constant colour
	uint4 *ring_rec; // RING_ENTRIES (tma build): triangle, pixel masks
	u32 *results;	// CHUNK_SAMPLES: colour per sample (the segment build keeps its sample words here too)
	unsigned short *samples; // CHUNK_SAMPLES: pixel | entry << 5
	uint2 *chunk;	// CHUNK_ENTRIES: (live pixel mask, first sample) of the chunk's entries
	unsigned long long *bars; // RING_BLOCKS mbarriers (tma build)
};
// + 64: the segment build spills 32 samples past the buffer
template <bool TMA>
constexpr int WARP_MEM_BYTES = CHUNK_ENTRIES * STAGE_WORDS * 16 + (TMA ? 2 * RING_ENTRIES * 16 + RING_BLOCKS * 8 : CHUNK_ENTRIES * 16) +
							   CHUNK_SAMPLES * 4 + CHUNK_SAMPLES * 2 + 64 + CHUNK_ENTRIES * 8;
template <bool TMA> __device__ __forceinline__ WarpMem warpMem(unsigned char *base) {
	WarpMem w;
	w.stage = reinterpret_cast<uint4 *>(base);
	base += CHUNK_ENTRIES * STAGE_WORDS * 16;
	if(TMA) {
		w.ring_rec = reinterpret_cast<uint4 *>(base);
		w.plane = w.ring_rec + RING_ENTRIES;
		base += 2 * RING_ENTRIES * 16;
		w.bars = reinterpret_cast<unsigned long long *>(base);
		base += RING_BLOCKS * 8;
	} else {
		w.ring_rec = nullptr, w.bars = nullptr;
		w.plane = reinterpret_cast<uint4 *>(base);
		base += CHUNK_ENTRIES * 16;
	}
	w.results = reinterpret_cast<u32 *>(base);
	w.samples = reinterpret_cast<unsigned short *>(base + CHUNK_SAMPLES * 4);
	w.chunk = reinterpret_cast<uint2 *>(base + CHUNK_SAMPLES * 4 + CHUNK_SAMPLES * 2 + 64);
	return w;
}

// ---- mbarrier / bulk-copy primitives (sm_90+ PTX; SASS: SYNCS.*, UBLKCP) -----------------------------------------
__device__ __forceinline__ u32 smemAddr(const void *ptr) { return (u32)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbarInit(unsigned long long *bar, u32 count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long *bar, u32 bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long *bar, u32 parity) {
	asm volatile("{\n"
				 ".reg .pred done;\n"
				 "WAIT_%=:\n"
				 "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
				 "@!done bra WAIT_%=;\n"
				 "}" ::"r"(smemAddr(bar)),
				 "r"(parity)
				 : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void *dst_smem, const void *src_gmem, u32 bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst_smem)),
				 "l"(src_gmem), "r"(bytes), "r"(smemAddr(bar))
				 : "memory");
}

// The item's entry stream through the shared-memory ring (tma build).  Blocks of RING_BLOCK entries carry a
// sequence number that runs on across the warp's items, so slot = seq & 3 and mbarrier phase = (seq >> 2) & 1.
struct RingStream {
	const uint4 *src_rec, *src_aux;
	int count, nblocks;
	u32 seq_base;
	int issued, waited; // blocks
};
__device__ __forceinline__ void ringIssue(const WarpMem &wm, RingStream &rs, int upto_block) {
	while(rs.issued < rs.nblocks && rs.issued < upto_block) {
		if(laneId() == 0) {
			const int b = rs.issued;
			const u32 slot = (rs.seq_base + (u32)b) & (RING_BLOCKS - 1);
			const u32 n = (u32)min(RING_BLOCK, rs.count - b * RING_BLOCK);
			mbarExpectTx(wm.bars + slot, n * 32u);
			bulkCopyG2S(wm.ring_rec + slot * RING_BLOCK, rs.src_rec + b * RING_BLOCK, n * 16u, wm.bars + slot);
			bulkCopyG2S(wm.plane + slot * RING_BLOCK, rs.src_aux + b * RING_BLOCK, n * 16u, wm.bars + slot);
		}
		rs.issued++;
	}
}
__device__ __forceinline__ void ringWait(const WarpMem &wm, RingStream &rs, int upto_block) {
	while(rs.waited < upto_block) {
		const u32 seq = rs.seq_base + (u32)rs.waited;
		mbarWait(wm.bars + (seq & (RING_BLOCKS - 1)), (seq / RING_BLOCKS) & 1u);
		rs.waited++;
	}
}
// blocks take consecutive slots, so entry e of the item sits at (first slot * RING_BLOCK + e) mod RING_ENTRIES
__device__ __forceinline__ int ringIndex(const RingStream &rs, int e) {
	return (int)((rs.seq_base * RING_BLOCK + (u32)e) & (RING_ENTRIES - 1));
}

// ------------------------------------------------------------------------------------------------

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

__device__ __forceinline__ int warpInclusiveScan(int v) {
	const int lane = laneId();
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, v, o);
		if(lane >= o)
			v += t;
	}
	return v;
}

struct ItemCtx {
	const uint4 *src_rec, *src_aux; // the item's slice of the sorted stream
	int count;
	int hb_x, hb_y; // top-left pixel of the half-block
	bool lower;		// LOW items: the block's lower half (second pixel mask)
	const float *opaque_depth; // PREPASS: the half-block's 32 nearest-opaque depths
};

// one 8x4 half-block from its sorted entries.  PREPASS (LUCID_OPT_OPAQUE_PREPASS): a sample behind the nearest opaque
// sample of its pixel does not exist -- it is not counted, shaded or reduced; the pixel lane decides that from the
// entry's depth plane before the chunk's samples are laid out.
template <bool TMA, bool PREPASS>
__device__ __forceinline__ void shadeHalfBlock(const Params &p, const LucidConfig &cfg, const ColourTables &tab, const LightTerms &lt,
											   const WarpMem &wm, const ItemCtx &it, u32 &seq_base) {
	const int lane = laneId();
	const bool additive = (p.opts & LUCID_OPT_ADDITIVE_BLENDING) != 0;
	const bool vis_errors = (p.opts & LUCID_OPT_VISUALIZE_ERRORS) != 0;
	const float fpx = float(it.hb_x + (lane & 7)), fpy = float(it.hb_y + (lane >> 3));
	const int count = it.count;
	Reducer red;
	reducerInit(red);
	u32 px_frags = 0;
	bool dead = false;
	const float zo = PREPASS ? __ldcg(it.opaque_depth + lane) : 0.0f;
	PhaseTimer timer = timerStart(p); // raster_timers: 2 unpack samples, 3 shade and reduce, 4 finish reduce

	RingStream rs;
	uint4 ahead_rec = make_uint4(0, 0, 0, 0), ahead_aux = make_uint4(0, 0, 0, 0);
	auto loadAhead = [&](int e) {
		ahead_rec = make_uint4(0, 0, 0, 0), ahead_aux = make_uint4(0, 0, 0, 0);
		if(e < count)
			ahead_rec = __ldcg(it.src_rec + e), ahead_aux = __ldcg(it.src_aux + e);
	};
	if(TMA) {
		rs.src_rec = it.src_rec, rs.src_aux = it.src_aux, rs.count = count;
		rs.nblocks = (count + RING_BLOCK - 1) / RING_BLOCK;
		rs.seq_base = seq_base, rs.issued = 0, rs.waited = 0;
		ringIssue(wm, rs, RING_BLOCKS);
	} else {
		loadAhead(lane);
	}

	for(int next = 0; next < count;) {
		const int base = next, e = base + lane;
		uint4 rec, aux;
		if(TMA) {
			// blocks below the chunk's first entry are consumed: their slots take the blocks four ahead
			ringIssue(wm, rs, base / RING_BLOCK + RING_BLOCKS);
			ringWait(wm, rs, (min(base + CHUNK_ENTRIES, count) - 1) / RING_BLOCK + 1);
			rec = make_uint4(0, 0, 0, 0), aux = make_uint4(0, 0, 0, 0);
			if(e < count)
				rec = wm.ring_rec[ringIndex(rs, e)], aux = wm.plane[ringIndex(rs, e)];
		} else {
			rec = ahead_rec, aux = ahead_aux;
		}
		u32 bits = e < count ? (it.lower ? rec.z : rec.y) : 0u;
		const int nf = __popc(bits);
		const int incl = warpInclusiveScan(nf);
		const bool in_chunk = e < count && incl <= CHUNK_SAMPLES;
		const int taken = __popc(__ballot_sync(0xffffffffu, in_chunk)); // a prefix, never empty
		const int total = __shfl_sync(0xffffffffu, incl, taken - 1);
		const int off = incl - nf;
		if(!in_chunk)
			bits = 0;
		next += taken;
		if(!TMA)
			loadAhead(next + lane);

		u32 tm = transpose32(bits);
		// the chunk's planes by chunk-local entry index
		const u32 plane_base = TMA ? rs.seq_base * RING_BLOCK + (u32)base : 0u;
		auto planeAt = [&](int j) { return wm.plane[TMA ? ((plane_base + (u32)j) & (RING_ENTRIES - 1)) : (u32)j]; };
		if(PREPASS) {
			// the pixel's surviving samples of this chunk; every one of them counts, also at a pixel that is
			// already opaque
			if(!TMA) {
				wm.plane[lane] = aux;
				__syncwarp();
			}
			u32 keep = 0;
			for(u32 b = tm; b; b &= b - 1) {
				const int j = __ffs(b) - 1;
				const uint4 s = planeAt(j);
				const float depth = __uint_as_float(s.x) * fpx + (__uint_as_float(s.y) * fpy + __uint_as_float(s.z));
				if(!(depth < zo))
					keep |= 1u << j;
			}
			tm = keep;
		}
		px_frags += __popc(tm);
		timerMark(timer, p.info->raster_timers, 2);
		if(dead)
			continue;
		// a pixel whose transmittance has reached zero takes exactly +0 from every later sample:
		// its samples are neither shaded nor reduced (the opaque cull of shading.glsl:31-32, exact part)
		const u32 dead_px = vis_errors ? 0u : __ballot_sync(0xffffffffu, red.trans == 0.0f);
		if((dead_px >> lane) & 1u)
			tm = 0;
		if(PREPASS) {
			// back to the entries' view: the pixels of entry `lane` that are still to be shaded
			bits = transpose32(tm);
			if(__all_sync(0xffffffffu, bits == 0)) {
				__syncwarp();
				continue;
			}
		}

		if(__all_sync(0xffffffffu, !in_chunk || aux.w != AUX_VARYING)) {
			// constant-colour triangles only: the pixel lane evaluates its own depths
			if(!TMA && !PREPASS) {
				wm.plane[lane] = aux;
				__syncwarp();
			}
			while(tm) {
				const int j = __ffs(tm) - 1;
				tm &= tm - 1;
				const uint4 s = planeAt(j);
				const float depth = __uint_as_float(s.x) * fpx + (__uint_as_float(s.y) * fpy + __uint_as_float(s.z));
				reducerPush(red, s.w, depth, additive, vis_errors);
			}
		} else {
			// samples of live pixels only: offsets are recomputed over the masked pixel sets
			const u32 live = PREPASS ? bits : bits & ~dead_px;
			int live_off = off, live_total = total;
			if(PREPASS || dead_px != 0) {
				const int nl = __popc(live);
				const int li = warpInclusiveScan(nl);
				live_off = li - nl;
				live_total = __shfl_sync(0xffffffffu, li, 31);
			}
			if(in_chunk) {
				wm.chunk[lane] = make_uint2(live, (u32)live_off);
				u32 dst = (u32)live_off, b = live;
				const u32 word = (u32)lane << 5;
				while(b) {
					const u32 pid = __ffs(b) - 1;
					b &= b - 1;
					wm.samples[dst++] = (unsigned short)(pid | word);
				}
				// everything shadeSample reads of this entry's triangle, once
				if(aux.w == AUX_VARYING && live != 0)
					stageEntry(p, lt, rec.x, wm.stage + lane * STAGE_WORDS);
			}
			if(!TMA && !PREPASS)
				wm.plane[lane] = aux;
			__syncwarp();
			for(int r0 = 0; r0 < live_total; r0 += 32) {
				const int idx = r0 + lane;
				if(idx < live_total) {
					const u32 sw = wm.samples[idx], pid = sw & 31u, j = sw >> 5;
					const uint4 s = planeAt((int)j);
					u32 color = s.w;
					if(s.w == AUX_VARYING)
						color = shadeStaged(p, tab, lt, wm.stage + j * STAGE_WORDS, __uint_as_float(s.x), __uint_as_float(s.y),
											__uint_as_float(s.z), float(it.hb_x + (int)(pid & 7)), float(it.hb_y + (int)(pid >> 3)));
					wm.results[idx] = color;
				}
			}
			__syncwarp();
			while(tm) {
				const int j = __ffs(tm) - 1;
				tm &= tm - 1;
				const uint2 c = wm.chunk[j];
				const u32 color = wm.results[c.y + __popc(c.x & laneMaskLt())];
				const uint4 s = planeAt(j);
				const float depth = __uint_as_float(s.x) * fpx + (__uint_as_float(s.y) * fpy + __uint_as_float(s.z));
				reducerPush(red, color, depth, additive, vis_errors);
			}
		}
		__syncwarp();
		timerMark(timer, p.info->raster_timers, 3);
		if(!vis_errors && __all_sync(0xffffffffu, red.trans == 0.0f)) {
			dead = true;
			if(!p.frag_counts && !PREPASS) // the pre-pass counts the surviving samples of the rest of the list too
				break;
		}
	}
	if(PREPASS) {
		u32 n = px_frags;
#pragma unroll
		for(int o = 16; o > 0; o >>= 1)
			n += __shfl_xor_sync(0xffffffffu, n, o);
		if(lane == 0 && n)
			atomicAdd(&p.info->stats[0], n);
	}
	if(TMA) {
		// copies still in flight land before the ring is used again
		ringWait(wm, rs, rs.issued);
		seq_base = rs.seq_base + (u32)rs.issued;
		__syncwarp();
	}
	writePixel(p, cfg, red, it.hb_x, it.hb_y, px_frags, additive, vis_errors);
	timerMark(timer, p.info->raster_timers, 4);
}

// shadeSample straight from global memory (segment-accurate build only)
__device__ __forceinline__ u32 shadeSampleGlobal(const Params &p, const ColourTables &tab, const LightTerms &lt, int ipx, int ipy,
												 u32 tri_idx, float &out_depth) {
	const float px = float(ipx), py = float(ipy);
	const uint4 *rec = reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx);
	const uint4 dq = __ldg(rec), misc = __ldg(rec + 1);
	const float dx = __uint_as_float(dq.x), dy = __uint_as_float(dq.y), dz = __uint_as_float(dq.z);
	out_depth = dx * px + (dy * py + dz);
	if(misc.w != 0)
		return misc.z; // attribute-free triangle: colour was evaluated once in quad setup
	uint4 e[STAGE_WORDS];
#pragma unroll
	for(int i = 0; i < STAGE_WORDS; i++)
		e[i] = make_uint4(0, 0, 0, 0);
	stageEntry(p, lt, tri_idx, e);
	return shadeStaged(p, tab, lt, e, dx, dy, dz, px, py);
}

// Segment-accurate variant (raster.glsl:292-396) for ALPHA_THRESHOLD: samples are expanded and
// consumed in the reference's 256-sample segments, so the early-out decisions fall on the same
// sample boundaries.
__device__ __forceinline__ void shadeHalfBlockSegments(const Params &p, const LucidConfig &cfg, const ColourTables &tab,
													   const LightTerms &lt, const WarpMem &wm, const ItemCtx &it) {
	const int lane = laneId();
	const int count = it.count;
	u32 *samples = wm.results; // SEGMENT_SIZE + 32 words: pixel | tri << 8 (results + samples + pad are contiguous)
	u32 *mask = reinterpret_cast<u32 *>(wm.chunk);
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
			u32 v = samples[SEGMENT_SIZE + lane];
			__syncwarp();
			samples[lane] = v;
		}
		__syncwarp();
		const u32 seg_end = seg_start + SEGMENT_SIZE;
		while(next < count && off < seg_end) {
			const int i = next + lane;
			uint4 rec = make_uint4(0, 0, 0, 0);
			if(i < count)
				rec = __ldcg(it.src_rec + i);
			const u32 tri_idx = rec.x;
			u32 bits = it.lower ? rec.z : rec.y;
			const int nf = __popc(bits);
			const int incl = warpInclusiveScan(nf);
			u32 my_off = off + (u32)(incl - nf);
			bool in_seg = i < count && my_off < seg_end;
			int taken = __popc(__ballot_sync(0xffffffffu, in_seg)); // a prefix of the lanes
			if(in_seg) {
				u32 dst = my_off - seg_start, word = tri_idx << 8;
				while(bits) {
					u32 pid = __ffs(bits) - 1;
					bits &= bits - 1;
					samples[dst++] = pid | word;
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
			u32 val = active ? samples[idx] : 0u;
			mask[lane] = 0;
			__syncwarp();
			u32 color = 0;
			float depth = 0.0f;
			if(active && !stop) {
				u32 pid = val & 31u;
				color = shadeSampleGlobal(p, tab, lt, it.hb_x + (int)(pid & 7), it.hb_y + (int)(pid >> 3), val >> 8, depth);
			}
			if(active)
				atomicOr(&mask[val & 31u], 1u << lane);
			__syncwarp();
			u32 pm = mask[lane];
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
	writePixel(p, cfg, red, it.hb_x, it.hb_y, px_frags, false, false);
}

// TMA: entry stream through the bulk-copy ring; SEGMENTS: the ALPHA_THRESHOLD build; PREPASS: LUCID_OPT_OPAQUE_PREPASS
template <bool TMA, bool SEGMENTS, bool PREPASS = false>
__global__ void __launch_bounds__(BLOCK_WARPS * 32, SHADE_MIN_CTAS)
	k_block_shade(const __grid_constant__ Params p, const __grid_constant__ LucidConfig cfg) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ float2 s_s2l[LUCID_S2L_SIZE], s_l2s[LUCID_L2S_SIZE];
	const int lane = laneId(), warp = threadIdx.x >> 5;
	// everything that does not depend on the previous kernel comes before the grid dependency wait: the colour
	// tables, the frame's light terms, the barriers of the ring
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	for(int i = threadIdx.x; i < LUCID_S2L_SIZE; i += BLOCK_WARPS * 32)
		s_s2l[i] = reinterpret_cast<const float2 *>(d_s2l_words)[i];
	for(int i = threadIdx.x; i < LUCID_L2S_SIZE; i += BLOCK_WARPS * 32)
		s_l2s[i] = reinterpret_cast<const float2 *>(d_l2s_words)[i];
	ColourTables tab;
	tab.s2l = s_s2l, tab.l2s = s_l2s;
	const LightTerms lt = lightTerms(cfg.lighting);
	const WarpMem wm = warpMem<TMA>(smem + (size_t)warp * WARP_MEM_BYTES<TMA>);
	if(TMA && lane == 0) {
		for(int i = 0; i < RING_BLOCKS; i++)
			mbarInit(wm.bars + i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	asm volatile("griddepcontrol.wait;" ::: "memory");

	u32 class_end[ITEM_CLASSES]; // exclusive end of every class in ticket order
	{
		u32 acc = 0;
#pragma unroll
		for(int k = 0; k < ITEM_CLASSES; k++)
			class_end[k] = acc += p.work_counters[WC_CLASS + k];
	}
	const u32 n_items = class_end[ITEM_CLASSES - 1];
	u32 seq_base = 0;
	// the queue entry of the next item is requested while the current one is shaded
	auto fetch = [&]() {
		uint4 e = make_uint4(0, 0, 0, 0);
		if(lane == 0) {
			const u32 i = atomicAdd(&p.work_counters[WC_SHADE], 1u);
			if(i < n_items)
				e = fetchWorkItem(p, i, class_end);
		}
		return e;
	};
	uint4 next_entry = fetch();
	while(true) {
		const u32 item = __shfl_sync(0xffffffffu, next_entry.x, 0);
		const int count = (int)__shfl_sync(0xffffffffu, next_entry.y, 0);
		const u32 offset = __shfl_sync(0xffffffffu, next_entry.z, 0);
		if(count == 0)
			break;
		next_entry = fetch();
		const long long t_item = clock64();
		const int bin_id = (int)(item >> 6), sub = (int)(item & 31u);
		const bool high = (item & 32u) != 0;
		const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
		ItemCtx it;
		it.src_rec = p.sorted_rec + offset, it.src_aux = p.sorted_aux + offset, it.count = count;
		it.hb_x = bin_x * BIN_SIZE + (sub & 3) * 8;
		const int y0 = bin_y * BIN_SIZE + (sub >> 2) * (high ? 4 : 8);
		const int halves = high ? 1 : 2;
		for(int half = 0; half < halves; half++) {
			it.lower = half != 0, it.hb_y = y0 + half * 4;
			if(PREPASS) {
				const int hbi = high ? sub : (((sub >> 2) * 2 + half) * 4 + (sub & 3));
				it.opaque_depth = p.opaque_depth + ((size_t)bin_id * 32 + hbi) * 32;
			}
			if(SEGMENTS)
				shadeHalfBlockSegments(p, cfg, tab, lt, wm, it);
			else
				shadeHalfBlock<TMA, PREPASS>(p, cfg, tab, lt, wm, it, seq_base);
			__syncwarp();
		}
		if(lane == 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(p.bin_cost) + bin_id, (unsigned long long)(clock64() - t_item));
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

void launchRasterBins(const Params &p, u32 background, cudaStream_t stream, int num_sms);
void launchBlockSort(const Params &p, u32 background, cudaStream_t stream, int num_sms);

bool shadeStreamTma() {
	static const bool tma = [] {
		// default: coalesced 128-bit loads one chunk ahead.  LUCID_SHADE_STREAM=tma selects the bulk-copy ring, measured
		// 9 % (10M-triangle scene) to 23 % (hairball) slower in the shading kernel: the ring costs 1.5 KB of shared
		// memory per warp (6 instead of 7 CTAs per SM) and the stream was never the latency that matters
		// (profiles/README.md, r2c)
		const char *v = getenv("LUCID_SHADE_STREAM");
		return v && v[0] == 't';
	}();
	return tma;
}

void launchRaster(const Params &p, const LucidConfig &cfg, cudaStream_t stream, cudaEvent_t *ev, int num_sms) {
	static std::once_flag configured[64]; // function attributes are per device
	constexpr int smem_tma = BLOCK_WARPS * WARP_MEM_BYTES<true>, smem_ldg = BLOCK_WARPS * WARP_MEM_BYTES<false>;
	oncePerDevice(configured, [] {
		cudaFuncSetAttribute(k_block_shade<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tma);
		cudaFuncSetAttribute(k_block_shade<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ldg);
		cudaFuncSetAttribute(k_block_shade<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ldg);
		cudaFuncSetAttribute(k_block_shade<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_ldg);
	});
	const LucidVec4 &bg = cfg.background_color;
	auto q = [](float v) { return (u32)(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f + 0.5f); };
	const u32 bg8 = q(bg.x) | (q(bg.y) << 8) | (q(bg.z) << 16) | 0xff000000u;
	launchRasterBins(p, bg8, stream, num_sms);
	if(ev)
		cudaEventRecord(ev[0], stream);
	launchBlockSort(p, bg8, stream, num_sms);
	if(ev)
		cudaEventRecord(ev[1], stream);
	const bool segments = (p.opts & (LUCID_OPT_ALPHA_THRESHOLD | LUCID_OPT_ADDITIVE_BLENDING | LUCID_OPT_VISUALIZE_ERRORS)) ==
						  LUCID_OPT_ALPHA_THRESHOLD;
	// (fewer persistent CTAs per SM, to leave room for the kernels of a neighbouring frame in flight, was measured:
	// 4 or 5 instead of 7 CTAs change nothing with three frames in flight and cost 10-20 % with one, profiles/r2u_*)
	const int grid = num_sms * SHADE_MIN_CTAS;
	if(segments)
		launchPDL((k_block_shade<false, true>), grid, BLOCK_WARPS * 32, (size_t)smem_ldg, stream, p, cfg);
	else if(opaquePrepass(p))
		launchPDL((k_block_shade<false, false, true>), grid, BLOCK_WARPS * 32, (size_t)smem_ldg, stream, p, cfg);
	else if(shadeStreamTma())
		launchPDL((k_block_shade<true, false>), grid, BLOCK_WARPS * 32, (size_t)smem_tma, stream, p, cfg);
	else
		launchPDL((k_block_shade<false, false>), grid, BLOCK_WARPS * 32, (size_t)smem_ldg, stream, p, cfg);
	if(ev)
		cudaEventRecord(ev[2], stream);
}

} // namespace lucid
