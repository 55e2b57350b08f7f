// raster_sort.cu -- stage 2 of the raster pipeline: depth sort of every block list (generateBlocks /
// generateRBlocks sort, raster_low.glsl:107-160, raster_high.glsl:146-260) and the frame's bookkeeping.
// A work item is one 8x8 block of a LOW bin or one 8x4 half-block of a HIGH bin and belongs to one warp.  The
// kernel leaves the item's entries in depth order in the sorted-entry stream: two contiguous planes of 16-byte
// words, (triangle, pixel mask of the upper / only half, pixel mask of the lower half, -) and (depth plane xyz,
// constant colour or AUX_VARYING), which is all k_block_shade reads of a list.
#include "raster_common.cuh"

namespace lucid {

#ifndef SORT_MIN_CTAS
#define SORT_MIN_CTAS 8
#endif
#ifndef RB_KEY_UNROLL
#define RB_KEY_UNROLL 2
#endif
constexpr int KEY_UNROLL = RB_KEY_UNROLL;

// ------------------------------------------------------------------------------------------------
// frame bookkeeping done by the first CTAs of k_block_sort before they take work items (two
// more launches on a frame of a few hundred microseconds would cost more than the work itself):
// background for empty bins (the reference leaves them to the application's clear,
// lucid_app.cpp:606-619), red for bins over the reference's limits (raster_high.glsl:313-317), and
// the level bookkeeping of promoted bins: appended to the HIGH list in bin order
// (raster_low.glsl:230-237,294-298).  Everything read here was written by k_raster_bins.
__device__ __forceinline__ void finishBins(const Params &p, u32 background, u32 *s_mask) {
	const int lane = laneId(), warp = threadIdx.x >> 5;
	// bin lists over their capacity (k_bin_scan set temp[1], dispatch and k_raster_bins did nothing): every
	// owned bin is painted red, like a bin over the reference's own limits, and the host is told
	const u32 overflow = p.info->temp[1]; // bit 0: bin lists, bit 1: sorted-entry stream (those bins carry flag 2)
	const bool list_overflow = (overflow & 1u) != 0;
	if(overflow != 0 && blockIdx.x == 0 && threadIdx.x == 0 && p.host_status)
		*p.host_status = overflow;
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


// PREPASS: LUCID_OPT_OPAQUE_PREPASS -- before the keys, the warp works out the depth of the nearest INST_IS_OPAQUE
// sample of every pixel of the item (lane = pixel, one pass over the opaque entries) and leaves it in
// p.opaque_depth for k_block_shade; an entry that lies behind that depth at every pixel of its half-block(s) --
// decided conservatively from the depth plane at the tile corners -- gets the padding key, sorts to the end and is
// cut off the item, so it is neither sorted nor streamed nor shaded.
template <bool PREPASS>
__global__ void __launch_bounds__(BLOCK_WARPS * 32, SORT_MIN_CTAS) k_block_sort(const __grid_constant__ Params p, u32 background) {
	__shared__ __align__(16) u32 s_keys[BLOCK_WARPS][SMEM_KEYS];
	__shared__ int s_misc[BLOCK_WARPS];
	const int lane = laneId(), warp = threadIdx.x >> 5;
	pdlEntry();
	finishBins(p, background, reinterpret_cast<u32 *>(s_misc));
	if(blockIdx.x == gridDim.x - 1) {
		__syncthreads();
		promoteBins(p, s_misc);
	}
	u32 *large_keys = p.large_keys + (size_t)(blockIdx.x * BLOCK_WARPS + warp) * MAX_HBLOCK_TRIS;
	u32 class_end[ITEM_CLASSES]; // exclusive end of every class in ticket order
	{
		u32 acc = 0;
#pragma unroll
		for(int k = 0; k < ITEM_CLASSES; k++)
			class_end[k] = acc += p.work_counters[WC_CLASS + k];
	}
	const u32 n_items = class_end[ITEM_CLASSES - 1];
	u32 frag_acc = 0, hbt_acc = 0;
	while(true) {
		u32 index = lane == 0 ? atomicAdd(&p.work_counters[WC_SORT], 1u) : 0u;
		index = __shfl_sync(0xffffffffu, index, 0);
		if(index >= n_items)
			break;
		const uint4 entry = fetchWorkItem(p, index, class_end);
		const long long t_item = clock64();
		PhaseTimer timer = timerStart(p); // raster_timers 1: generate blocks (keys, sort, sorted entries)
		const u32 item = entry.x;
		const int count = (int)entry.y;
		const int bin_id = (int)(item >> 6), sub = (int)(item & 31u);
		const bool high = (item & 32u) != 0;
		const int bin_y = bin_id / p.bin_count_x, bin_x = bin_id - bin_y * p.bin_count_x;
		const int pos_x = bin_x * BIN_SIZE, pos_y = bin_y * BIN_SIZE;
		const int cx8 = (sub & 3) * 8, ry = sub >> 2;
		const unsigned char *list = blockList(p, bin_id, sub, high);
		const bool large = count > SMEM_KEYS;
		u32 *keys = large ? large_keys : s_keys[warp];
		// lists of up to 512 entries sort inside the first half of the warp's key array (warpSortShared pads to a
		// power of two): the second half keeps the entries' triangle indices, so that ordering depth ties -- far
		// geometry collapses onto few key values -- reads shared memory instead of one dependent global load per
		// comparison (block sort of the 10M-triangle scene 0.550 -> 0.521 ms, hairball 1.142 -> 1.087 ms)
		const bool tris_in_smem = count <= SMEM_KEYS / 2;

		// depth keys from the centroid of the covered pixels (raster.glsl:142-176); the records are one round
		// trip, the triangles' depth planes a second, dependent one: the records of the next iteration are
		// requested together with this iteration's planes
		auto loadRec = [&](int i) {
			uint4 r4 = make_uint4(0, 0, 0, 0);
			if(i < count) {
				if(high) {
					uint2 r = __ldg(reinterpret_cast<const uint2 *>(list) + i);
					r4.x = r.x, r4.y = r.y;
				} else {
					r4 = __ldg(reinterpret_cast<const uint4 *>(list) + i);
				}
			}
			return r4;
		};
		// ---- opaque pre-pass: nearest opaque sample depth per pixel (sample depths are inverse ray positions:
		// larger = nearer); zfar[h] = the farthest of them over the pixels of half h (-inf: some pixel has no opaque cover)
		float zfar[2] = {-INFINITY, -INFINITY};
		const float tile_x = float(pos_x + cx8), tile_y = float(pos_y + ry * (high ? 4 : 8));
		if(PREPASS) {
			float zo[2] = {-INFINITY, -INFINITY};
			const float fpx = float(pos_x + cx8 + (lane & 7)), fpy = float(pos_y + ry * (high ? 4 : 8) + (lane >> 3));
			for(int i0 = 0; i0 < count; i0 += 32) {
				const uint4 rec = loadRec(i0 + lane);
				uint4 dq = make_uint4(0, 0, 0, 0);
				if(i0 + lane < count)
					dq = __ldg(reinterpret_cast<const uint4 *>(p.tri_shade + (rec.x & 0xffffffu)));
				const bool opaque = (dq.w & LUCID_INST_IS_OPAQUE) != 0;
				u32 m0 = 0, m1 = 0;
				if(opaque) {
					u32 tri_idx, mins, maxs;
					int nf;
					if(high) {
						unpackHighRecord(make_uint2(rec.x, rec.y), tri_idx, mins, maxs);
						m0 = rowsToBits(mins, maxs, cx8, nf);
					} else {
						unpackLowRecord(rec, false, tri_idx, mins, maxs);
						m0 = rowsToBits(mins, maxs, cx8, nf);
						unpackLowRecord(rec, true, tri_idx, mins, maxs);
						m1 = rowsToBits(mins, maxs, cx8, nf);
					}
				}
				for(u32 todo = __ballot_sync(0xffffffffu, opaque); todo; todo &= todo - 1) {
					const int j = __ffs(todo) - 1;
					const float dx = __uint_as_float(__shfl_sync(0xffffffffu, dq.x, j));
					const float dy = __uint_as_float(__shfl_sync(0xffffffffu, dq.y, j));
					const float dz = __uint_as_float(__shfl_sync(0xffffffffu, dq.z, j));
					const u32 c0 = __shfl_sync(0xffffffffu, m0, j), c1 = __shfl_sync(0xffffffffu, m1, j);
					if((c0 >> lane) & 1u)
						zo[0] = fmaxf(zo[0], dx * fpx + (dy * fpy + dz));
					if((c1 >> lane) & 1u)
						zo[1] = fmaxf(zo[1], dx * fpx + (dy * (fpy + 4.0f) + dz));
				}
			}
			const int halves = high ? 1 : 2;
			for(int h = 0; h < halves; h++) {
				const int hbi = high ? sub : ((ry * 2 + h) * 4 + (sub & 3));
				p.opaque_depth[((size_t)bin_id * 32 + hbi) * 32 + lane] = zo[h];
				float z = zo[h];
#pragma unroll
				for(int o = 16; o > 0; o >>= 1)
					z = fminf(z, __shfl_xor_sync(0xffffffffu, z, o));
				zfar[h] = z;
			}
		}
		const bool cull_any = PREPASS && (zfar[0] > -INFINITY || zfar[1] > -INFINITY);
		// the nearest depth a triangle's plane takes on a half-block: every rounding of the evaluation is monotone in
		// x and in y, so the largest value over the 8x4 pixels is taken at a corner
		auto nearestOnTile = [&](uint4 d, float y0) {
			const float dx = __uint_as_float(d.x), dy = __uint_as_float(d.y), dz = __uint_as_float(d.z);
			const float a = dx * tile_x, b = dx * (tile_x + 7.0f), c = dy * y0 + dz, e = dy * (y0 + 3.0f) + dz;
			return fmaxf(fmaxf(a + c, a + e), fmaxf(b + c, b + e));
		};
		int n_culled = 0;

		uint4 rec_next[KEY_UNROLL];
#pragma unroll
		for(int u = 0; u < KEY_UNROLL; u++)
			rec_next[u] = loadRec(u * 32 + lane);
		for(int i0 = 0; i0 < count; i0 += 32 * KEY_UNROLL) {
			uint4 rec[KEY_UNROLL], dq[KEY_UNROLL];
#pragma unroll
			for(int u = 0; u < KEY_UNROLL; u++) {
				rec[u] = rec_next[u];
				dq[u] = __ldg(reinterpret_cast<const uint4 *>(p.tri_shade + (rec[u].x & 0xffffffu)));
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
					if(nf == 0 && p.debug_records)
						debugRecord(p, LUCID_DEBUG_EMPTY_COVERAGE, item, 0, 0, 0, 0);
					if(cull_any && nearestOnTile(dq[u], tile_y) < zfar[0])
						depth = 0xffffffffu;
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
					if(nf0 + nf1 == 0 && p.debug_records)
						debugRecord(p, LUCID_DEBUG_EMPTY_COVERAGE, item, 0, 0, 0, 0);
					if(PREPASS && count <= 3)
						depth = 0; // unsorted lists keep their order when hidden entries are cut out of them
					if(cull_any && (nf0 == 0 || nearestOnTile(dq[u], tile_y) < zfar[0]) &&
					   (nf1 == 0 || nearestOnTile(dq[u], tile_y + 4.0f) < zfar[1]))
						depth = 0xffffffffu;
				}
				if(PREPASS && depth == 0xffffffffu)
					n_culled++;
				keys[i] = (u32)i | depth;
				if(tris_in_smem)
					keys[SMEM_KEYS / 2 + i] = high ? (rec[u].x & 0xffffffu) : rec[u].x; // what the tie order looks up
			}
		}
		__syncwarp();
		// stats: LOW counts the block's triangles once per half-block (raster_low.glsl:272-275),
		// HIGH the exact half-block list (raster_high.glsl:309-310)
		hbt_acc += lane == 0 ? (u32)count * (high ? 1u : 2u) : 0u;
		const int slot_bits = high ? 14 : 10;
		int kept = count; // entries that go into the stream
		if(PREPASS && cull_any) {
#pragma unroll
			for(int o = 16; o > 0; o >>= 1)
				n_culled += __shfl_xor_sync(0xffffffffu, n_culled, o);
			kept = count - n_culled;
		}
		if(high || count > 3) { // LOW blocks with <= 3 triangles rely on the window alone (raster_low.glsl:144)
			if(large)
				warpSortLarge(keys, count, s_keys[warp]);
			else
				warpSortShared(keys, count);
			if(p.debug_records) {
				// "making sure that tris are properly ordered" (raster_low.glsl:154-163, raster_high.glsl:196-203)
				if(p.debug_inject && index == 0 && count > 1 && lane == 0) {
					volatile u32 *vk = keys;
					const u32 t = vk[0];
					vk[0] = vk[1], vk[1] = t;
				}
				__syncwarp();
				for(int i = lane; i < count; i += 32) {
					const u32 value = large ? __ldcg(keys + i) : keys[i];
					const u32 prev_value = i == 0 ? 0u : large ? __ldcg(keys + i - 1) : keys[i - 1];
					if(value <= prev_value && i > 0)
						debugRecord(p, LUCID_DEBUG_UNSORTED, item, (u32)i, (u32)count, prev_value, value);
				}
				__syncwarp();
				if(p.debug_inject && index == 0 && count > 1 && lane == 0) {
					volatile u32 *vk = keys;
					const u32 t = vk[0];
					vk[0] = vk[1], vk[1] = t;
				}
				__syncwarp();
			}
			// depth ties by triangle index: the triangle of a list position is looked up in the list itself
			// a long run of ties is left to k_tie_runs: where it starts in the sorted-entry stream, and its length
			auto onLongRun = [&](int first, int len) {
				const u32 q = atomicAdd(p.tie_runs, 1u);
				if(q < TIE_RUN_QUEUE)
					reinterpret_cast<uint2 *>(p.tie_runs + 2)[q] = make_uint2(entry.z + (u32)first, (u32)len);
			};
			if(tris_in_smem)
				warpFixDepthTies(keys, kept, slot_bits, TIE_RUN_LONG_SMEM, [&](u32 pos) { return keys[SMEM_KEYS / 2 + pos]; }, onLongRun);
			else if(high)
				warpFixDepthTies(keys, kept, slot_bits, TIE_RUN_LONG,
								 [&](u32 pos) { return __ldg(reinterpret_cast<const uint2 *>(list) + pos).x & 0xffffffu; }, onLongRun);
			else
				warpFixDepthTies(keys, kept, slot_bits, TIE_RUN_LONG, [&](u32 pos) { return __ldg(reinterpret_cast<const uint4 *>(list) + pos).x; },
								 onLongRun);
		} else if(PREPASS && kept < count) {
			warpSortShared(keys, count); // keys are list positions: the kept entries move to the front in list order
		}
		if(PREPASS && kept < count && lane == 0)
			workItemSlot(p, index, class_end)->y = (u32)kept; // k_block_shade takes the item's length from here
		// the entries in sorted order: (triangle, pixel masks) and (depth plane, constant colour).  (Four entries per lane
		// and trip, to have the two dependent loads of several entries in flight, was measured: 0.52 -> 0.66 ms.)
		const u32 pos_mask = (1u << slot_bits) - 1u;
		uint4 *out_rec = p.sorted_rec + entry.z, *out_aux = p.sorted_aux + entry.z;
		for(int i = lane; i < kept; i += 32) {
			const u32 pos = (large ? __ldcg(keys + i) : keys[i]) & pos_mask;
			u32 tri_idx, mins, maxs, mask0, mask1 = 0;
			int nf;
			if(high) {
				unpackHighRecord(__ldg(reinterpret_cast<const uint2 *>(list) + pos), tri_idx, mins, maxs);
				mask0 = rowsToBits(mins, maxs, cx8, nf);
			} else {
				const uint4 r = __ldg(reinterpret_cast<const uint4 *>(list) + pos);
				unpackLowRecord(r, false, tri_idx, mins, maxs);
				mask0 = rowsToBits(mins, maxs, cx8, nf);
				unpackLowRecord(r, true, tri_idx, mins, maxs);
				mask1 = rowsToBits(mins, maxs, cx8, nf);
			}
			const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_shade + tri_idx);
			const uint4 dq = __ldg(src), misc = __ldg(src + 1);
			__stcg(out_rec + i, make_uint4(tri_idx, mask0, mask1, 0u));
			__stcg(out_aux + i, make_uint4(dq.x, dq.y, dq.z, misc.w != 0 ? misc.z : AUX_VARYING));
		}
		__syncwarp();
		timerMark(timer, p.info->raster_timers, 1);
		if(lane == 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(p.bin_cost) + bin_id, (unsigned long long)(clock64() - t_item));
	}
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		frag_acc += __shfl_xor_sync(0xffffffffu, frag_acc, o);
		hbt_acc += __shfl_xor_sync(0xffffffffu, hbt_acc, o);
	}
	if(lane == 0) {
		if(frag_acc && !PREPASS) // with the pre-pass stats[0] counts the surviving samples: k_block_shade adds them
			atomicAdd(&p.info->stats[0], frag_acc);
		if(hbt_acc)
			atomicAdd(&p.info->stats[1], hbt_acc);
	}
}

// Long runs of depth ties (k_block_sort's queue): a warp takes a run, copies its stream entries aside, keeps their
// triangle indices in shared memory, and moves every entry to the place its triangle index has among the run's --
// the order the short runs get by insertion.  Runs are rare (coplanar stacks, very distant geometry), so the kernel
// usually finds an empty queue; it exists so that such a frame costs milliseconds, not seconds.
constexpr int TIE_WARPS = 4;
__global__ void __launch_bounds__(TIE_WARPS * 32) k_tie_runs(const __grid_constant__ Params p) {
	extern __shared__ u32 s_tris[]; // TIE_WARPS x MAX_HBLOCK_TRIS
	const int lane = laneId(), warp = threadIdx.x >> 5;
	pdlEntry();
	const u32 n_runs = min(p.tie_runs[0], (u32)TIE_RUN_QUEUE);
	u32 *tris = s_tris + warp * MAX_HBLOCK_TRIS;
	uint4 *scratch = p.tie_scratch + (size_t)(blockIdx.x * TIE_WARPS + warp) * (2 * MAX_HBLOCK_TRIS);
	for(u32 run = blockIdx.x * TIE_WARPS + warp; run < n_runs; run += gridDim.x * TIE_WARPS) {
		const uint2 r = reinterpret_cast<const uint2 *>(p.tie_runs + 2)[run];
		const int len = min((int)r.y, MAX_HBLOCK_TRIS);
		uint4 *rec = p.sorted_rec + r.x, *aux = p.sorted_aux + r.x;
		for(int e = lane; e < len; e += 32) {
			const uint4 a = __ldcg(rec + e), b = __ldcg(aux + e);
			scratch[e] = a, scratch[MAX_HBLOCK_TRIS + e] = b;
			tris[e] = a.x;
		}
		__syncwarp();
		for(int e = lane; e < len; e += 32) {
			const u32 te = tris[e];
			int rank = 0;
			for(int m = 0; m < len; m++) {
				const u32 tm = tris[m];
				rank += (tm < te || (tm == te && m < e)) ? 1 : 0;
			}
			__stcg(rec + rank, scratch[e]);
			__stcg(aux + rank, scratch[MAX_HBLOCK_TRIS + e]);
		}
		__syncwarp();
	}
}

static int blockSortGrid(int num_sms) { return num_sms * SORT_MIN_CTAS; }
size_t rasterLargeKeysCount(int num_sms) { return (size_t)blockSortGrid(num_sms) * BLOCK_WARPS * MAX_HBLOCK_TRIS; }
static int tieRunsGrid(int num_sms) { return num_sms; }
size_t rasterTieScratchCount(int num_sms) { return (size_t)tieRunsGrid(num_sms) * TIE_WARPS * 2 * MAX_HBLOCK_TRIS; }

void launchBlockSort(const Params &p, u32 background, cudaStream_t stream, int num_sms) {
	if(opaquePrepass(p))
		launchPDL(k_block_sort<true>, blockSortGrid(num_sms), BLOCK_WARPS * 32, 0, stream, p, background);
	else
		launchPDL(k_block_sort<false>, blockSortGrid(num_sms), BLOCK_WARPS * 32, 0, stream, p, background);
	static std::once_flag configured[64];
	constexpr int tie_smem = TIE_WARPS * MAX_HBLOCK_TRIS * (int)sizeof(u32);
	oncePerDevice(configured, [] { cudaFuncSetAttribute(k_tie_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, tie_smem); });
	launchPDL(k_tie_runs, tieRunsGrid(num_sms), TIE_WARPS * 32, (size_t)tie_smem, stream, p);
}

} // namespace lucid
