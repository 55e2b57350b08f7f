// binning.cu -- bin counting, offsets + categorisation, dispatch and list canonicalisation.
//
// Replaces data/shaders/bin_counter.glsl, bin_categorizer.glsl and bin_dispatcher.glsl.
// Like the reference, a work group (CTA) counts its share of the small quads into a private
// shared-memory histogram and later claims one slice of every bin it touches
// (bin_counter.glsl:185-190, bin_dispatcher.glsl:205-209); unlike it, the share is a fixed
// contiguous chunk of the visible quads, so no batch list has to be kept for the replay: the
// dispatcher simply recounts its chunk.  Large triangles add +1/-1 to a per-row difference array
// that the scan kernel integrates; one CTA turns the counters into offsets and the LOW/HIGH bin
// lists.
// The order inside a bin's list is therefore arbitrary (as in the reference); the raster stage
// breaks depth-key ties by triangle index, so the image does not depend on it.
#include "common.cuh"

namespace lucid {

constexpr int BIN_THREADS = 256;

__device__ __forceinline__ int *cnt(const Params &p, int which) {
	return p.counts + (size_t)which * p.bin_count;
}

// scanline.glsl:28-52 -- edge functions evaluated at each bin's trivial-reject corner
struct BinScan {
	float mn[3], mx[3], step[3];
};
__device__ __forceinline__ BinScan loadBinScan(const TriScan &t, int &min_by, int &max_by) {
	BinScan s;
	const float inf = __int_as_float(0x7f800000);
	float scan[3] = {__uint_as_float(t.s0.x), __uint_as_float(t.s0.y), __uint_as_float(t.s0.z)};
	float step[3] = {__uint_as_float(t.s1.x), __uint_as_float(t.s1.y), __uint_as_float(t.s1.z)};
	min_by = (int)(t.s0.w & 0xffff) >> BIN_SHIFT;
	max_by = (int)(t.s0.w >> 16) >> BIN_SHIFT;
	u32 signs = t.s1.w;
	const float offset = float(BIN_SIZE) - 0.989f;
	float start_x = 0.99f, start_y = float(min_by * BIN_SIZE) - 0.01f;
#pragma unroll
	for(int i = 0; i < 3; i++) {
		bool xneg = (signs >> i) & 1, yneg = (signs >> (3 + i)) & 1;
		float yoff = yneg ? 0.0f : offset, xoff = xneg ? 0.0f : offset;
		float v = scan[i] + (step[i] * (yoff + start_y) - (xoff + start_x));
		s.mn[i] = xneg ? -inf : v;
		s.mx[i] = xneg ? v : inf;
		s.step[i] = step[i] * float(BIN_SIZE);
	}
	return s;
}
// bin_counter.glsl:53-62
__device__ __forceinline__ void binScanStep(BinScan &s, int &bmin, int &bmax) {
	float xmin = fmaxf(fmaxf(s.mn[0], s.mn[1]), s.mn[2]);
	float xmax = fminf(fminf(s.mx[0], s.mx[1]), s.mx[2]);
#pragma unroll
	for(int i = 0; i < 3; i++) {
		s.mn[i] += s.step[i];
		s.mx[i] += s.step[i];
	}
	bmin = f2i(xmin + 1.0f) >> BIN_SHIFT;
	bmax = f2i(xmax) >> BIN_SHIFT;
}

// ------------------------------------------------------------------------------------------------
// counting (bin_counter.glsl:64-134)

// contiguous chunk of the visible small quads owned by this CTA (same split in count and dispatch)
struct SmallChunk {
	int begin, end;
};
__device__ __forceinline__ SmallChunk smallChunk(int n_small) {
	const int per = ((n_small + (int)gridDim.x - 1) / (int)gridDim.x + BIN_THREADS - 1) / BIN_THREADS * BIN_THREADS;
	SmallChunk c;
	c.begin = min((int)blockIdx.x * per, n_small), c.end = min(c.begin + per, n_small);
	return c;
}

__global__ void __launch_bounds__(BIN_THREADS) k_bin_count(const Params p) {
	extern __shared__ int s_hist[]; // one counter per bin
	pdlEntry();
	const int n_small = p.info->num_visible_quads[0], n_large = p.info->num_visible_quads[1];
	int *quad_counts = cnt(p, LUCID_CNT_QUAD_COUNTS);
	int *tri_diff = cnt(p, LUCID_CNT_TRI_COUNTS); // per-row difference array until the scan kernel
	const int bcx = p.bin_count_x;
	const int stride = gridDim.x * blockDim.x;
	const int first = blockIdx.x * blockDim.x + threadIdx.x;
	PhaseTimer timer = timerStart(p); // bin_dispatcher_timers: 0 count small quads, 1 count large tris, 2 / 3 dispatch

	// small quads: every bin of the (<= 4 bins) AABB, conservative.  A CTA counts a contiguous chunk
	// of the visible quads into a private shared-memory histogram and adds its non-zero bins to the
	// global counters once: visible quads keep their input order, so a chunk is spatially coherent
	// and touches few bins, while quad-by-quad global atomics pile up on the scene's hot bins (the
	// reference keeps per-work-group histograms for the same reason, bin_counter.glsl:185-190)
	for(int b = threadIdx.x; b < p.bin_count; b += BIN_THREADS)
		s_hist[b] = 0;
	__syncthreads();
	const SmallChunk chunk = smallChunk(n_small);
	for(int base = chunk.begin; base < chunk.end; base += BIN_THREADS) {
		int q = base + threadIdx.x;
		bool valid = q < chunk.end;
		u32 enc = valid ? p.quad_aabbs[q] : 0u;
		int bsx = enc & 0x7f, bsy = (enc >> 7) & 0x7f, bex = (enc >> 14) & 0x7f, bey = (enc >> 21) & 0x7f;
		int w = bex - bsx + 1, n = valid ? w * (bey - bsy + 1) : 0;
#pragma unroll
		for(int k = 0; k < 4; k++) {
			int by = bsy + k / max(w, 1), bx = bsx + k % max(w, 1);
			bool ok = k < n && ownsBin(p, by * bcx + bx);
			if(ok)
				atomicAdd(&s_hist[by * bcx + bx], 1); // shared-memory atomics resolve same-bin lanes in hardware
		}
	}
	__syncthreads();
	for(int b = threadIdx.x; b < p.bin_count; b += BIN_THREADS) {
		const int c = s_hist[b];
		if(c != 0)
			atomicAdd(quad_counts + b, c);
	}
	timerMark(timer, p.info->bin_dispatcher_timers, 0);
	// large triangles: one thread per triangle, +1 at the first bin of each row span and -1 just
	// after the last (bin_counter.glsl:123-133)
	for(int i = first; i < n_large * 2; i += stride) {
		int quad_idx = (p.max_visible_quads - 1) - (i >> 1), second = i & 1;
		u32 enc = p.quad_aabbs[quad_idx];
		if((enc >> (30 + second)) & 1)
			continue;
		int bsx = enc & 0x7f, bex = (enc >> 14) & 0x7f, bsy, bey;
		TriScan t;
		const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_scan + ((u32)quad_idx * 2 + second));
		t.s0 = src[0], t.s1 = src[1];
		BinScan s = loadBinScan(t, bsy, bey);
		for(int by = bsy; by <= bey; by++) {
			int bmin, bmax;
			binScanStep(s, bmin, bmax);
			bmin = max(bmin, bsx), bmax = min(bmax, bex);
			clipToOwnedBins(p, by, bmin, bmax);
			if(bmax >= bmin) {
				atomicAdd(tri_diff + by * bcx + bmin, 1);
				if(bmax + 1 < bcx)
					atomicAdd(tri_diff + by * bcx + bmax + 1, -1);
			}
		}
	}
	timerMark(timer, p.info->bin_dispatcher_timers, 1);
}

// ------------------------------------------------------------------------------------------------
// offsets + categories (bin_categorizer.glsl:22-89), one CTA

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int MAX_SCAN_TILES = 128 * 128 / SCAN_THREADS; // 7-bit bin coordinates

// The CTA walks the bins in tiles of 1024 consecutive bins (thread = bin, so every load and store
// is coalesced; the values of all tiles are requested up front): one fused block scan per tile over
// (quad count, triangle count, LOW | HIGH << 16 membership) with the totals carried from tile to
// tile, so offsets and level lists come out in bin order.
__global__ void __launch_bounds__(SCAN_THREADS) k_bin_scan(const Params p) {
	__shared__ int s_warp[3][SCAN_WARPS + 1];
	pdlEntry();
	const int bc = p.bin_count, bcx = p.bin_count_x, bcy = p.bin_count_y;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int *qc = cnt(p, LUCID_CNT_QUAD_COUNTS), *qo = cnt(p, LUCID_CNT_QUAD_OFFSETS);
	int *qt = cnt(p, LUCID_CNT_QUAD_OFFSETS_TEMP);
	int *tc = cnt(p, LUCID_CNT_TRI_COUNTS), *to = cnt(p, LUCID_CNT_TRI_OFFSETS);
	int *tt = cnt(p, LUCID_CNT_TRI_OFFSETS_TEMP);
	int *low = cnt(p, LUCID_CNT_LOW_BINS), *high = cnt(p, LUCID_CNT_HIGH_BINS);

	// integrate the large-triangle difference arrays row by row (one warp per bin row; the row's
	// chunks of 32 bins are loaded together, then scanned in order)
	for(int by = warp; by < bcy; by += SCAN_WARPS) {
		int v[4];
#pragma unroll
		for(int c = 0; c < 4; c++) {
			const int bx = c * 32 + lane;
			v[c] = bx < bcx ? tc[by * bcx + bx] : 0;
		}
		int carry = 0;
#pragma unroll
		for(int c = 0; c < 4; c++) {
			if(c * 32 >= bcx)
				break;
			int x = v[c];
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				int t = __shfl_up_sync(0xffffffffu, x, o);
				if(lane >= o)
					x += t;
			}
			x += carry;
			const int bx = c * 32 + lane;
			if(bx < bcx)
				tc[by * bcx + bx] = x;
			carry = __shfl_sync(0xffffffffu, x, 31);
		}
	}
	__syncthreads();

	const int tiles = (bc + SCAN_THREADS - 1) / SCAN_THREADS;
	int q_of[MAX_SCAN_TILES], t_of[MAX_SCAN_TILES];
#pragma unroll
	for(int i = 0; i < MAX_SCAN_TILES; i++) {
		const int b = i * SCAN_THREADS + (int)threadIdx.x;
		q_of[i] = 0, t_of[i] = 0;
		if(i < tiles && b < bc)
			q_of[i] = qc[b], t_of[i] = tc[b];
	}
	int base_q = 0, base_t = 0, base_low = 0, base_high = 0;
#pragma unroll
	for(int i = 0; i < MAX_SCAN_TILES; i++) {
		if(i >= tiles)
			break;
		const int b = i * SCAN_THREADS + (int)threadIdx.x;
		const int q = q_of[i], t = t_of[i];
		const int num_tris = t + q * 2;
		const bool valid = b < bc;
		const bool is_low = valid && num_tris != 0 && num_tris < 1024, is_high = valid && num_tris >= 1024;
		int xq = q, xt = t, xl = (is_low ? 1 : 0) | (is_high ? 1 << 16 : 0);
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			int a0 = __shfl_up_sync(0xffffffffu, xq, o), a1 = __shfl_up_sync(0xffffffffu, xt, o);
			int a2 = __shfl_up_sync(0xffffffffu, xl, o);
			if(lane >= o)
				xq += a0, xt += a1, xl += a2;
		}
		if(lane == 31)
			s_warp[0][warp] = xq, s_warp[1][warp] = xt, s_warp[2][warp] = xl;
		__syncthreads();
		if(warp < 3) {
			int w = s_warp[warp][lane], wi = w;
#pragma unroll
			for(int o = 1; o < 32; o <<= 1) {
				int a = __shfl_up_sync(0xffffffffu, wi, o);
				if(lane >= o)
					wi += a;
			}
			s_warp[warp][lane] = wi - w;
			if(lane == 31)
				s_warp[warp][SCAN_WARPS] = wi;
		}
		__syncthreads();
		const int off_q = base_q + s_warp[0][warp] + xq - q, off_t = base_t + s_warp[1][warp] + xt - t;
		const int lh = s_warp[2][warp] + xl - ((is_low ? 1 : 0) | (is_high ? 1 << 16 : 0));
		if(valid) {
			qo[b] = off_q, qt[b] = off_q, to[b] = off_t, tt[b] = off_t;
			if(is_low)
				low[base_low + (lh & 0xffff)] = b;
			if(is_high)
				high[base_high + (lh >> 16)] = b;
			p.bin_flags[b] = 0;
		}
		base_q += s_warp[0][SCAN_WARPS], base_t += s_warp[1][SCAN_WARPS];
		base_low += s_warp[2][SCAN_WARPS] & 0xffff, base_high += s_warp[2][SCAN_WARPS] >> 16;
		__syncthreads();
	}
	const int tot_q = base_q, tot_t = base_t, tot_low = base_low, tot_high = base_high;
	const int tot_empty = bc - tot_low - tot_high;
	if(threadIdx.x == 0) {
		LucidInfo *info = p.info;
		info->bin_level_counts[LUCID_BIN_LEVEL_EMPTY] = tot_empty;
		info->bin_level_counts[LUCID_BIN_LEVEL_MICRO] = 0;
		info->bin_level_counts[LUCID_BIN_LEVEL_LOW] = tot_low;
		info->bin_level_counts[LUCID_BIN_LEVEL_MEDIUM] = 0;
		info->bin_level_counts[LUCID_BIN_LEVEL_HIGH] = tot_high;
		for(int l = 0; l < LUCID_BIN_LEVELS_COUNT; l++) {
			int md = p.max_dispatches >> (l == LUCID_BIN_LEVEL_HIGH ? 1 : 0);
			info->bin_level_dispatches[l][0] = min(info->bin_level_counts[l], md);
			info->bin_level_dispatches[l][1] = 1;
			info->bin_level_dispatches[l][2] = 1;
		}
		int n_small = info->num_visible_quads[0], n_large = info->num_visible_quads[1];
		info->num_counted_quads[0] = n_small, info->num_counted_quads[1] = n_large;
		// quad_setup.glsl:471-485 (kept for getStats-style consumers; nothing is launched from it)
		int num_tasks = (n_small + 4095) / 4096 + (n_large + 511) / 512;
		info->num_binning_dispatches[0] = (u32)min(max(num_tasks, 4), p.max_dispatches);
		info->num_binning_dispatches[1] = 1, info->num_binning_dispatches[2] = 1;
		// list capacity check: the reference sizes both lists at 2 * MAX_VISIBLE_QUADS and never checks
		if((u32)tot_q > p.bin_list_capacity || (u32)tot_t > p.bin_list_capacity)
			info->temp[1] = 1;
		for(int i = 0; i < WORK_COUNTERS; i++)
			p.work_counters[i] = 0;
	}
}

// ------------------------------------------------------------------------------------------------
// dispatch (bin_dispatcher.glsl:66-114)

// Large triangles are dispatched through a per-warp queue of span segments: every lane walks the
// bin rows of its own triangle (the incremental scan of bin_counter.glsl:53-62 is serial per
// triangle) but only *queues* (triangle, first bin, up to 8 bins) segments; 32 queued segments are
// then written by 32 lanes, each issuing all of its position claims before the first dependent
// store.  A thread that claimed and stored bin by bin spent one atomic round trip per (triangle,
// bin) on the critical path, and a tall or wide triangle held its warp for hundreds of them.  The
// reference balances the same work with per-row segment scans (bin_dispatcher.glsl:123-196).
constexpr int DISPATCH_RING = 64;  // segments per warp
constexpr int SEGMENT_BINS = 8;

__device__ __forceinline__ void drainSegments(const Params &p, int *tri_cursor, const uint2 *ring, int index, bool active) {
	const uint2 seg = ring[index & (DISPATCH_RING - 1)];
	const int cell = (int)(seg.y & 0xffffu), n = active ? (int)(seg.y >> 16) : 0;
	int pos[SEGMENT_BINS];
#pragma unroll
	for(int k = 0; k < SEGMENT_BINS; k++)
		if(k < n)
			pos[k] = atomicAdd(tri_cursor + cell + k, 1);
#pragma unroll
	for(int k = 0; k < SEGMENT_BINS; k++)
		if(k < n)
			p.bin_tris[pos[k]] = seg.x;
}

__global__ void __launch_bounds__(BIN_THREADS) k_bin_dispatch(const Params p) {
	__shared__ uint2 s_ring[BIN_THREADS / 32][DISPATCH_RING];
	extern __shared__ int s_hist[]; // one counter / cursor per bin
	pdlEntry();
	if(p.info->temp[1] != 0)
		return; // lists would overflow their buffers
	const int n_small = p.info->num_visible_quads[0], n_large = p.info->num_visible_quads[1];
	int *quad_cursor = cnt(p, LUCID_CNT_QUAD_OFFSETS_TEMP);
	int *tri_cursor = cnt(p, LUCID_CNT_TRI_OFFSETS_TEMP);
	const int bcx = p.bin_count_x;
	const int stride = gridDim.x * blockDim.x;
	const int lane = laneId();
	PhaseTimer timer = timerStart(p);

	// small quads: the CTA recounts its chunk into the shared-memory histogram, claims one range per
	// non-zero bin with a single global atomic (all claims of a thread are in flight together), and
	// hands out the positions inside its ranges with shared-memory atomics
	for(int b = threadIdx.x; b < p.bin_count; b += BIN_THREADS)
		s_hist[b] = 0;
	__syncthreads();
	const SmallChunk chunk = smallChunk(n_small);
	for(int pass = 0; pass < 2; pass++) {
		for(int base = chunk.begin; base < chunk.end; base += BIN_THREADS) {
			int q = base + threadIdx.x;
			bool valid = q < chunk.end;
			u32 enc = valid ? p.quad_aabbs[q] : 0u;
			u32 word = (u32)q | (enc & 0xf0000000u);
			int bsx = enc & 0x7f, bsy = (enc >> 7) & 0x7f, bex = (enc >> 14) & 0x7f, bey = (enc >> 21) & 0x7f;
			int w = bex - bsx + 1, n = valid ? w * (bey - bsy + 1) : 0;
#pragma unroll
			for(int k = 0; k < 4; k++) {
				int by = bsy + k / max(w, 1), bx = bsx + k % max(w, 1);
				bool ok = k < n && ownsBin(p, by * bcx + bx);
				if(ok) {
					const int pos = atomicAdd(&s_hist[by * bcx + bx], 1);
					if(pass == 1)
						p.bin_quads[pos] = word;
				}
			}
		}
		__syncthreads();
		if(pass == 0) {
#pragma unroll 4
			for(int b = threadIdx.x; b < p.bin_count; b += BIN_THREADS) {
				const int c = s_hist[b];
				if(c != 0)
					s_hist[b] = atomicAdd(quad_cursor + b, c);
			}
			__syncthreads();
		}
	}

	timerMark(timer, p.info->bin_dispatcher_timers, 2);
	uint2 *ring = s_ring[threadIdx.x >> 5];
	int q_head = 0, q_tail = 0; // warp-uniform
	// all lanes: queue the span [bmin, bmax] of bin row `row_cell / bcx` in segments, draining when 32 wait
	auto pushSpan = [&](u32 tri_idx, int row_cell, int bmin, int bmax) {
		while(true) {
			const bool has = bmin <= bmax;
			const u32 m = __ballot_sync(0xffffffffu, has);
			if(m == 0)
				break;
			if(has) {
				const int n = min(bmax - bmin + 1, SEGMENT_BINS);
				ring[(q_tail + __popc(m & laneMaskLt())) & (DISPATCH_RING - 1)] =
					make_uint2(tri_idx, (u32)(row_cell + bmin) | ((u32)n << 16));
				bmin += n;
			}
			q_tail += __popc(m);
			__syncwarp();
			if(q_tail - q_head >= 32) {
				drainSegments(p, tri_cursor, ring, q_head + lane, true);
				q_head += 32;
				__syncwarp();
			}
		}
	};
	for(int base = blockIdx.x * blockDim.x; base < n_large * 2; base += stride) {
		int i = base + threadIdx.x;
		bool valid = i < n_large * 2;
		int quad_idx = (p.max_visible_quads - 1) - (i >> 1), second = i & 1;
		u32 enc = valid ? p.quad_aabbs[quad_idx] : 0u;
		valid = valid && !((enc >> (30 + second)) & 1);
		u32 tri_idx = (u32)quad_idx * 2 + second;
		int bsx = enc & 0x7f, bex = (enc >> 14) & 0x7f, bsy = 0, bey = -1;
		BinScan s;
#pragma unroll
		for(int e = 0; e < 3; e++)
			s.mn[e] = s.mx[e] = s.step[e] = 0.0f;
		if(valid) {
			TriScan t;
			const uint4 *src = reinterpret_cast<const uint4 *>(p.tri_scan + tri_idx);
			t.s0 = src[0], t.s1 = src[1];
			s = loadBinScan(t, bsy, bey);
		}
		int rows = valid ? bey - bsy + 1 : 0;
		// A triangle with many segments (a wall across the screen) would keep the warp in the row
		// loop with one busy lane: the warp takes it over afterwards, one lane per bin row.
		const bool big = rows * ((bex - bsx) / SEGMENT_BINS + 1) > 48;
		u32 big_mask = __ballot_sync(0xffffffffu, big);
		const int own_rows = big ? 0 : rows;
		const int max_rows = __reduce_max_sync(0xffffffffu, own_rows);
		for(int r = 0; r < max_rows; r++) {
			int bmin = 0, bmax = -1;
			if(r < own_rows) {
				binScanStep(s, bmin, bmax);
				bmin = max(bmin, bsx), bmax = min(bmax, bex);
				clipToOwnedBins(p, bsy + r, bmin, bmax);
			}
			pushSpan(tri_idx, (bsy + r) * bcx, bmin, bmax);
		}
		while(big_mask) {
			const int src_lane = __ffs(big_mask) - 1;
			big_mask &= big_mask - 1;
			BinScan bs;
#pragma unroll
			for(int e = 0; e < 3; e++) {
				bs.mn[e] = __shfl_sync(0xffffffffu, s.mn[e], src_lane);
				bs.mx[e] = __shfl_sync(0xffffffffu, s.mx[e], src_lane);
				bs.step[e] = __shfl_sync(0xffffffffu, s.step[e], src_lane);
			}
			const int b_rows = __shfl_sync(0xffffffffu, rows, src_lane), b_bsy = __shfl_sync(0xffffffffu, bsy, src_lane);
			const int b_bsx = __shfl_sync(0xffffffffu, bsx, src_lane), b_bex = __shfl_sync(0xffffffffu, bex, src_lane);
			const u32 b_tri = __shfl_sync(0xffffffffu, tri_idx, src_lane);
			for(int r0 = 0; r0 < b_rows; r0 += 32) {
				// the row scan is a chain of additions: lane l repeats the first r0 + l of them, so its
				// state is bit for bit the one the serial loop has at that row
				BinScan mine = bs;
				for(int k = 0; k < lane; k++)
#pragma unroll
					for(int e = 0; e < 3; e++)
						mine.mn[e] += mine.step[e], mine.mx[e] += mine.step[e];
				const int r = r0 + lane;
				int bmin = 0, bmax = -1;
				if(r < b_rows) {
					binScanStep(mine, bmin, bmax);
					bmin = max(bmin, b_bsx), bmax = min(bmax, b_bex);
					clipToOwnedBins(p, b_bsy + r, bmin, bmax);
				}
				pushSpan(b_tri, (b_bsy + r) * bcx, bmin, bmax);
				for(int k = 0; k < 32; k++)
#pragma unroll
					for(int e = 0; e < 3; e++)
						bs.mn[e] += bs.step[e], bs.mx[e] += bs.step[e];
			}
		}
	}
	if(q_tail > q_head)
		drainSegments(p, tri_cursor, ring, q_head + lane, lane < q_tail - q_head);
	timerMark(timer, p.info->bin_dispatcher_timers, 3);
}

void launchBinning(const Params &p, cudaStream_t stream, cudaEvent_t *ev) {
	int grid = 148 * 4;
	const size_t hist_bytes = (size_t)p.bin_count * sizeof(int);
	static std::once_flag configured[64]; // function attributes are per device
	oncePerDevice(configured, [] { // up to 128 x 128 bins (7-bit bin coordinates)
		cudaFuncSetAttribute(k_bin_count, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * (int)sizeof(int));
		cudaFuncSetAttribute(k_bin_dispatch, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * (int)sizeof(int));
	});
	launchPDL(k_bin_count, grid, BIN_THREADS, hist_bytes, stream, p);
	if(ev)
		cudaEventRecord(ev[0], stream);
	launchPDL(k_bin_scan, 1, SCAN_THREADS, 0, stream, p);
	if(ev)
		cudaEventRecord(ev[1], stream);
	launchPDL(k_bin_dispatch, grid, BIN_THREADS, hist_bytes, stream, p);
	if(ev)
		cudaEventRecord(ev[2], stream);
}

} // namespace lucid
