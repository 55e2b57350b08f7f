// raster_bins.cu -- stage 1 of the raster pipeline: the block lists of every non-empty bin
// (generateRowTris + generateBlocks / computeRBlockGroups of the reference, raster_low.glsl:39-106,
// raster_high.glsl:54-144).  See raster_common.cuh for the pipeline.
#include "raster_common.cuh"

namespace lucid {

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

template <bool HIGH, int THREADS, typename Emit>
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
	for(int base = warp * 32; base < b.n_T; base += THREADS) {
		const Fetched cur = ahead;
		if(base + THREADS < b.n_T)
			ahead = fetch(base + THREADS);
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

struct BinShared {
	int count[32]; // entries per half-block (HIGH) / block (LOW)
	int holes[32]; // HIGH: columns inside a record's [first, last] range without coverage
	int fill[32];  // COMPACT: entries written so far by the filling pass
	u32 start[32]; // COMPACT: the list's first 8-byte unit in the pool
	int bin_index, status;
};

// a persistent CTA takes one bin at a time: HIGH bins first (they take longest).  256 threads when there are bins for
// every CTA slot of the device; a device that owns few bins (its share of a bin-range split) takes them with 512-thread
// CTAs, so that the heaviest bin -- the kernel's critical path -- is walked by 16 warps instead of 8
// COMPACT (LUCID_CREATE_COMPACT_LISTS): the walk below only counts; the bin then takes exactly its entries from the list
// pool with one atomic and a second walk fills the lists (slots by the same shared-memory atomics, so a list may come
// out in another order than the counting pass met it: the block sort orders it anyway).
template <int THREADS, bool COMPACT>
__global__ void __launch_bounds__(THREADS) k_raster_bins(const __grid_constant__ Params p, u32 background) {
	constexpr int WARPS = THREADS / 32;
	__shared__ BinShared sh;
	extern __shared__ __align__(16) uint4 s_ring_all[]; // WARPS rings of PHASE_A_RING items (32 bytes each)
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	pdlEntry();
	if(p.info->temp[1] != 0)
		return; // the bin lists did not fit their buffers (k_bin_scan): no list is valid, the frame is painted red
	const int n_high = p.info->bin_level_counts[LUCID_BIN_LEVEL_HIGH];
	const int n_low = p.info->bin_level_counts[LUCID_BIN_LEVEL_LOW];
	while(true) {
		__syncthreads();
		if(tid == 0)
			sh.bin_index = (int)atomicAdd(&p.work_counters[WC_BINS], 1u);
		__syncthreads();
		const int idx = sh.bin_index;
		if(idx >= n_high + n_low)
			break;
		const long long t_bin = clock64();
		PhaseTimer timer = timerStart(p); // raster_timers 0: generate rows (the span walk and the per-block lists)
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
			binPhaseA<false, THREADS>(p, b, s_ring_all + warp * (PHASE_A_RING * 2), [&](bool, u32 tri_idx, int g, u32 mn0, u32 mx0, u32 mn1, u32 mx1, u32 bx) {
				while(bx) {
					int c = __ffs(bx) - 1;
					bx &= bx - 1;
					int slot = atomicAdd(&sh.count[g * 4 + c], 1);
					if(!COMPACT && slot < MAX_BLOCK_TRIS)
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
				// (raster_low.glsl:101-105,230-237); promoteBins (k_block_sort) appends it to the HIGH list
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
			binPhaseA<true, THREADS>(p, b, s_ring_all + warp * (PHASE_A_RING * 2), [&](bool, u32 tri_idx, int g, u32 mn, u32 mx, u32, u32, u32 bx) {
				if(bx == 0)
					return;
				const int lo = __ffs(bx) - 1, hi = 31 - __clz(bx);
				u32 holes = ((2u << hi) - (1u << lo)) & ~bx;
				while(bx) {
					int c = __ffs(bx) - 1;
					bx &= bx - 1;
					int slot = atomicAdd(&sh.count[g * 4 + c], 1);
					if(!COMPACT && slot < HB_LIST_CAP)
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
				continue; // finishBins (k_block_sort) paints the bin
		}

		if(COMPACT) {
			// exactly the bin's entries from the pool (8-byte units: one per HIGH record, two per LOW record), the
			// lists one after the other; then the walk again, this time writing
			const int n_lists = high ? 32 : 16, units = high ? 1 : 2;
			if(tid < 32) {
				const int c = tid < n_lists ? sh.count[tid] * units : 0;
				int incl = c;
#pragma unroll
				for(int o = 1; o < 32; o <<= 1) {
					int t = __shfl_up_sync(0xffffffffu, incl, o);
					if(tid >= o)
						incl += t;
				}
				// an even number of units per bin: every bin starts on 16 bytes (the LOW records are 16-byte stores)
				const int total = (__shfl_sync(0xffffffffu, incl, 31) + 1) & ~1;
				u32 base = 0;
				if(tid == 0 && total > 0)
					base = atomicAdd(&p.work_counters[WC_LIST_POOL], (u32)total);
				base = __shfl_sync(0xffffffffu, base, 0);
				if((unsigned long long)base + (u32)total > p.list_pool_units) {
					if(tid == 0) { // the pool is full: a red bin and LUCID_E_LIMIT, like a full sorted-entry stream
						p.bin_flags[bin_id] |= 2u;
						p.info->temp[1] |= 2u;
						sh.status = 2;
					}
				} else {
					sh.start[tid] = base + (u32)(incl - c);
					sh.fill[tid] = 0;
					p.list_offsets[bin_id * 32 + tid] = base + (u32)(incl - c);
				}
			}
			__syncthreads();
			if(sh.status != 0)
				continue; // finishBins (k_block_sort) paints the bin
			unsigned char *pool = reinterpret_cast<unsigned char *>(p.block_lists);
			if(high)
				binPhaseA<true, THREADS>(p, b, s_ring_all + warp * (PHASE_A_RING * 2), [&](bool, u32 tri_idx, int g, u32 mn, u32 mx, u32, u32, u32 bx) {
					while(bx) {
						int c = __ffs(bx) - 1;
						bx &= bx - 1;
						const int l = g * 4 + c, slot = atomicAdd(&sh.fill[l], 1);
						if(slot < sh.count[l])
							reinterpret_cast<uint2 *>(pool + (size_t)sh.start[l] * 8)[slot] = packHighRecord(tri_idx, mn, mx);
					}
				});
			else
				binPhaseA<false, THREADS>(p, b, s_ring_all + warp * (PHASE_A_RING * 2), [&](bool, u32 tri_idx, int g, u32 mn0, u32 mx0, u32 mn1, u32 mx1, u32 bx) {
					while(bx) {
						int c = __ffs(bx) - 1;
						bx &= bx - 1;
						const int l = g * 4 + c, slot = atomicAdd(&sh.fill[l], 1);
						if(slot < sh.count[l])
							reinterpret_cast<uint4 *>(pool + (size_t)sh.start[l] * 8)[slot] = packLowRecord(tri_idx, mn0, mx0, mn1, mx1);
					}
				});
			__syncthreads();
		}

		// publish the non-empty blocks as work items of the block stages; empty ones only get the background.
		// The bin takes one slice of the sorted-entry stream for all its lists (one atomic), a block's slice
		// starts at the running sum of the counts before it.
		const int n_blocks = high ? 32 : 16;
		if(tid < 32) {
			const int c = tid < n_blocks ? sh.count[tid] : 0;
			if(tid < n_blocks)
				p.block_counts[bin_id * 32 + tid] = c;
			int incl = c;
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, incl, o);
				if(tid >= o)
					incl += t;
			}
			const int total = __shfl_sync(0xffffffffu, incl, 31);
			u32 base = 0;
			if(tid == 0 && total > 0)
				base = atomicAdd(&p.work_counters[WC_STREAM], (u32)total);
			base = __shfl_sync(0xffffffffu, base, 0);
			if((unsigned long long)base + (u32)total > p.stream_capacity) {
				// the sorted-entry stream is full (lucid_create: max_block_entries): the bin is painted red like
				// a bin over the reference's own limits and the frame reports LUCID_E_LIMIT
				if(tid == 0) {
					p.bin_flags[bin_id] |= 2u;
					p.info->temp[1] |= 2u;
					sh.status = 2;
				}
			} else {
				// one queue per size class, consumed from the heaviest class down: the kernels end with the
				// shortest items, so their tails are a few microseconds instead of one long list
				const u32 item = ((u32)bin_id << 6) | (high ? 32u : 0u) | (u32)tid;
				const int cls = itemClass(c);
#pragma unroll
				for(int k = 0; k < ITEM_CLASSES; k++) {
					const u32 m = __ballot_sync(0xffffffffu, c > 0 && cls == k);
					if(m == 0)
						continue;
					u32 qbase = 0;
					if(tid == __ffs(m) - 1)
						qbase = atomicAdd(&p.work_counters[WC_CLASS + k], (u32)__popc(m));
					qbase = __shfl_sync(0xffffffffu, qbase, __ffs(m) - 1);
					if((m >> tid) & 1)
						p.block_items[(size_t)k * p.block_items_cap + qbase + __popc(m & laneMaskLt())] =
							make_uint4(item, (u32)c, base + (u32)(incl - c), 0u);
				}
			}
		}
		__syncthreads();
		if(sh.status != 0)
			continue; // finishBins (k_block_sort) paints the bin
		timerMark(timer, p.info->raster_timers, 0);
		if(tid == 0)
			atomicAdd(reinterpret_cast<unsigned long long *>(p.bin_cost) + bin_id,
					  (unsigned long long)(clock64() - t_bin) * WARPS);
		const int rows = high ? 4 : 8;
		for(int blk = warp; blk < n_blocks; blk += WARPS) {
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


void launchRasterBins(const Params &p, u32 background, cudaStream_t stream, int num_sms) {
	// LUCID_RASTER_BINS_THREADS=256|512 overrides the choice (measurements).  Measured on the ranks of an 8-way split of
	// the 10M-triangle frame (profiles/r2w_*): with under a thousand heavy bins 512 threads take the list stage from
	// 0.145-0.158 to 0.106-0.109 ms, with many light bins 256 are faster (0.031 against 0.043 ms); 1024 never win
	static const int forced = [] {
		const char *v = getenv("LUCID_RASTER_BINS_THREADS");
		return v ? atoi(v) : 0;
	}();
	const int owned = p.bin_end - p.bin_begin;
	int threads = owned >= num_sms * 6 ? 256 : 512;
	if(forced == 256 || forced == 512)
		threads = forced;
	constexpr size_t ring_bytes_per_warp = PHASE_A_RING * 2 * sizeof(uint4);
	auto launch = [&](auto kernel, int ctas_per_sm, int nthreads) {
		launchPDL(kernel, num_sms * ctas_per_sm, nthreads, (size_t)(nthreads / 32) * ring_bytes_per_warp, stream, p, background);
	};
	if(p.compact_lists) {
		if(threads == 256)
			launch(k_raster_bins<256, true>, 4, 256);
		else
			launch(k_raster_bins<512, true>, 2, 512);
	} else {
		if(threads == 256)
			launch(k_raster_bins<256, false>, 4, 256);
		else
			launch(k_raster_bins<512, false>, 2, 512);
	}
}

} // namespace lucid
