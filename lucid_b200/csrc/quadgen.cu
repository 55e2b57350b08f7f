// quadgen.cu -- triangle -> quad pairing on the GPU: replaces the reference's offline CPU pass
// triNeighbours + quadNodes + genQuads (src/quad_generator.cpp:18-201, called from Scene::generateQuads,
// src/scene.cpp:237-242).  The renderer's primitive is the quad (two triangles sharing edge v0-v2,
// quad_setup.glsl:77-128), so a triangle mesh paired into quads halves the setup and binning work.
//
// Same pairing graph as the reference, bit for bit (the per-element rules are in quadgen_rules.h); the independent
// set is chosen by a round-synchronous greedy instead of the reference's sequential heap loop:
//   k_qg_edges       every directed edge goes into an open-addressing hash table keyed (v0, v1); the lowest
//                    (triangle, edge) owning it wins by atomicMin -- HashMap::emplace in input order
//   k_qg_neighbours  nb[t][i] = owner of the reversed edge
//   k_qg_masks       which (t, i) create a node; a block scan over the per-triangle counts numbers them in (t, i)
//                    order, the order the reference's loop creates them in
//   k_qg_nodes       node vertices / triangles / squareness, and tri_quads by the overwrite rule
//   k_qg_conflicts   <= 4 conflicting nodes per node
//   k_qg_select      ONE cooperative persistent kernel runs all rounds: live degree and key of every live node, grid
//                    barrier, a node whose key beats all its live neighbours is selected, grid barrier, neighbours of
//                    the selected leave; until no node is live.  Plain grid-stride loops: skipping finished chunks of
//                    nodes (a flag per chunk, or per-CTA lists of live chunks) was measured slower -- 241 and 187 ms
//                    against 137 ms for 10 M triangles -- because the flag load / the barrier per chunk serialises the
//                    loads a grid-stride loop keeps in flight
//   k_qg_emit_flags / k_qg_emit   quads in the order of their first triangle, unpaired triangles as (a, b, c, c)
// HBM-bound integer work: 12 B per triangle in, 16 B per quad out, ~150 B per triangle of intermediate traffic.
#include "../../include/lucid_quadgen.h"
#include "quadgen_rules.h"

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <string>

namespace cg = cooperative_groups;
using namespace lucid_qg;

namespace {

constexpr int QG_THREADS = 256;
constexpr int QG_AUGMENT_ROUNDS = 4; // fixed: a round without a path does nothing; the count is in the result
constexpr uint64_t EMPTY_KEY = ~0ull;

__device__ __forceinline__ uint32_t edgeSlot(uint64_t key, uint32_t mask) {
	key ^= key >> 33, key *= 0xff51afd7ed558ccdull, key ^= key >> 33, key *= 0xc4ceb9fe1a85ec53ull, key ^= key >> 33;
	return (uint32_t)key & mask;
}

__global__ void k_qg_edges(const int *tris, int nt, unsigned long long *keys, uint32_t *vals, uint32_t mask) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if(e >= nt * 3)
		return;
	const int t = e / 3, j = e - t * 3;
	const uint64_t key = ((uint64_t)(uint32_t)tris[t * 3 + j] << 32) | (uint32_t)tris[t * 3 + (j == 2 ? 0 : j + 1)];
	uint32_t slot = edgeSlot(key, mask);
	while(true) {
		const unsigned long long prev = atomicCAS(keys + slot, EMPTY_KEY, (unsigned long long)key);
		if(prev == EMPTY_KEY || prev == key) {
			atomicMin(vals + slot, (uint32_t)t * 4u + (uint32_t)j);
			return;
		}
		slot = (slot + 1) & mask;
	}
}

__global__ void k_qg_neighbours(const int *tris, int nt, const unsigned long long *keys, const uint32_t *vals, uint32_t mask, int *nb) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if(e >= nt * 3)
		return;
	const int t = e / 3, j = e - t * 3;
	const uint64_t key = ((uint64_t)(uint32_t)tris[t * 3 + (j == 2 ? 0 : j + 1)] << 32) | (uint32_t)tris[t * 3 + j];
	uint32_t slot = edgeSlot(key, mask);
	int out = -1;
	while(true) {
		const unsigned long long k = keys[slot];
		if(k == EMPTY_KEY)
			break;
		if(k == key) {
			const int owner = (int)(vals[slot] >> 2);
			if(owner != t)
				out = owner;
			break;
		}
		slot = (slot + 1) & mask;
	}
	nb[e] = out;
}

__global__ void k_qg_masks(const int *tris, const int *nb, int nt, unsigned char *mask, int *count) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt)
		return;
	const int m = createMask(tris, nb, t);
	mask[t] = (unsigned char)m;
	count[t] = popc3(m);
}

// ---- exclusive scan of n ints in place (three passes; the middle one is a single block) ----------------------------
constexpr int SCAN_ITEMS = 8, SCAN_TILE = QG_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int blockExclusiveScan(int v, int &total) {
	__shared__ int s_warp[QG_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int incl = v;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, incl, o);
		if(lane >= o)
			incl += t;
	}
	__syncthreads();
	if(lane == 31)
		s_warp[warp] = incl;
	__syncthreads();
	int before = 0, all = 0;
#pragma unroll
	for(int w = 0; w < QG_THREADS / 32; w++) {
		before += w < warp ? s_warp[w] : 0;
		all += s_warp[w];
	}
	total = all;
	return before + incl - v;
}
__global__ void k_qg_scan_tiles(int *data, int n, int *tile_sums, bool write) {
	const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	int v[SCAN_ITEMS], sum = 0;
#pragma unroll
	for(int k = 0; k < SCAN_ITEMS; k++) {
		v[k] = base + k < n ? data[base + k] : 0;
		sum += v[k];
	}
	int total;
	int run = blockExclusiveScan(sum, total);
	if(!write) {
		if(threadIdx.x == 0)
			tile_sums[blockIdx.x] = total;
		return;
	}
	run += tile_sums[blockIdx.x];
#pragma unroll
	for(int k = 0; k < SCAN_ITEMS; k++) {
		if(base + k < n)
			data[base + k] = run;
		run += v[k];
	}
}
// tile sums -> exclusive offsets, total behind them (n_tiles <= SCAN_TILE * 64)
__global__ void k_qg_scan_sums(int *tile_sums, int n_tiles) {
	__shared__ int s_carry;
	if(threadIdx.x == 0)
		s_carry = 0;
	__syncthreads();
	for(int first = 0; first < n_tiles; first += SCAN_TILE) {
		const int base = first + threadIdx.x * SCAN_ITEMS;
		int v[SCAN_ITEMS], sum = 0;
#pragma unroll
		for(int k = 0; k < SCAN_ITEMS; k++) {
			v[k] = base + k < n_tiles ? tile_sums[base + k] : 0;
			sum += v[k];
		}
		int total;
		int run = blockExclusiveScan(sum, total) + s_carry;
#pragma unroll
		for(int k = 0; k < SCAN_ITEMS; k++) {
			if(base + k < n_tiles)
				tile_sums[base + k] = run;
			run += v[k];
		}
		__syncthreads();
		if(threadIdx.x == 0)
			s_carry += total;
		__syncthreads();
	}
	if(threadIdx.x == 0)
		tile_sums[n_tiles] = s_carry;
}

__global__ void k_qg_nodes(const float *pos, const int *tris, const int *nb, const int *base, const unsigned char *mask, int nt,
						   int *tri_quads, int2 *node_tris, int4 *node_verts, float *node_sq) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= nt)
		return;
	for(int i = 0; i < 3; i++) {
		tri_quads[t * 3 + i] = finalTriQuad(nb, base, mask, t, i);
		if((mask[t] >> i) & 1) {
			const int q = nodeId(base, mask, t, i), u = nb[t * 3 + i];
			int v[4] = {tris[t * 3 + i], oppositeVert(tris + t * 3, tris + u * 3), tris[t * 3 + (i + 1) % 3], tris[t * 3 + (i + 2) % 3]};
			node_tris[q] = make_int2(t, u);
			node_verts[q] = make_int4(v[0], v[1], v[2], v[3]);
			node_sq[q] = squareness(pos, v);
		}
	}
}

__global__ void k_qg_conflicts(const int *tri_quads, const int2 *node_tris, int nq, int4 *conflicts, unsigned char *state) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if(q >= nq)
		return;
	int c[4];
	const int2 ab = node_tris[q];
	nodeConflicts(tri_quads, q, ab.x, ab.y, c);
	conflicts[q] = make_int4(c[0], c[1], c[2], c[3]);
	// a node in conflict with itself (its triangles are joined along two edges) can never be used
	state[q] = (c[0] == q || c[1] == q || c[2] == q || c[3] == q) ? 1 : 0;
}

// state: 0 live, 1 removed, 2 selected, 3 selected in the current round
// sync: [0..2] live counters in rotation, [3] rounds, [8..10] lowest live score in rotation
__global__ void __launch_bounds__(QG_THREADS) k_qg_select(const int4 *conflicts, const float *node_sq, int nq, float weight,
													   unsigned char *state, unsigned long long *key, int *sync) {
	cg::grid_group grid = cg::this_grid();
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	unsigned *min_slots = reinterpret_cast<unsigned *>(sync) + 8;
	for(int round = 0;; round++) {
		int *live_now = sync + round % 3;
		unsigned *min_now = min_slots + round % 3;
		// keys of the live nodes from their live degree
		int live = 0;
		unsigned lowest = 0xffffffffu;
		for(int q = tid; q < nq; q += stride) {
			if(state[q] != 0)
				continue;
			live++;
			const int4 c = conflicts[q];
			const int deg = (c.x >= 0 && state[c.x] == 0) + (c.y >= 0 && state[c.y] == 0) + (c.z >= 0 && state[c.z] == 0) +
							(c.w >= 0 && state[c.w] == 0);
			const unsigned long long k = nodeKey(q, deg, node_sq[q], weight);
			key[q] = k;
			lowest = min(lowest, (unsigned)(k >> 32));
		}
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) {
			live += __shfl_xor_sync(0xffffffffu, live, o);
			lowest = min(lowest, __shfl_xor_sync(0xffffffffu, lowest, o));
		}
		if((threadIdx.x & 31) == 0 && live) {
			atomicAdd(live_now, live);
			atomicMin(min_now, lowest);
		}
		if(tid == 0)
			sync[(round + 1) % 3] = 0, min_slots[(round + 1) % 3] = 0xffffffffu;
		grid.sync();
		if(*(volatile int *)live_now == 0) {
			if(tid == 0)
				sync[3] = round;
			return;
		}
		const unsigned limit = windowLimit(*(volatile unsigned *)min_now);
		// a live node inside the score window whose key beats every live neighbour's is selected (neighbours selected
		// in this round count as live: 3)
		for(int q = tid; q < nq; q += stride) {
			if(state[q] != 0)
				continue;
			const unsigned long long kq = key[q];
			if((unsigned)(kq >> 32) > limit)
				continue;
			const int4 c = conflicts[q];
			const int cs[4] = {c.x, c.y, c.z, c.w};
			bool best = true;
#pragma unroll
			for(int k = 0; k < 4; k++) {
				const int n = cs[k];
				if(n < 0)
					continue;
				const unsigned char sn = *(volatile unsigned char *)(state + n);
				if(sn != 0 && sn != 3)
					continue;
				const unsigned long long kn = key[n];
				if(kn < kq || (kn == kq && n < q))
					best = false;
			}
			if(best)
				state[q] = 3;
		}
		grid.sync();
		for(int q = tid; q < nq; q += stride) {
			if(state[q] != 3)
				continue;
			const int4 c = conflicts[q];
			const int cs[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
			for(int k = 0; k < 4; k++)
				if(cs[k] >= 0 && state[cs[k]] == 0)
					state[cs[k]] = 1;
			state[q] = 2;
		}
		grid.sync();
	}
}

// ---- augmentation: an unpaired triangle a next to a paired one (b, c) whose mate c has another unpaired neighbour d
// re-pairs as (a, b) + (c, d): one quad more.  Synchronous rounds; a path is applied when its first triangle holds the
// lowest claim on all four triangles, so the outcome does not depend on scheduling.
__device__ __forceinline__ bool selfConflict(const int4 *conflicts, int q) {
	const int4 c = conflicts[q];
	return c.x == q || c.y == q || c.z == q || c.w == q;
}
__global__ void k_qg_mates(const int2 *node_tris, const unsigned char *state, int nq, int *mate) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if(q < nq && state[q] == 2) {
		const int2 ab = node_tris[q];
		mate[ab.x] = q, mate[ab.y] = q;
	}
}
__global__ void k_qg_aug_propose(const int *tri_quads, const int2 *node_tris, const int4 *conflicts, const int *mate, int nt,
								 int4 *proposal, int *claim) {
	const int a = blockIdx.x * blockDim.x + threadIdx.x;
	if(a >= nt)
		return;
	int4 prop = make_int4(-1, -1, -1, -1);
	if(mate[a] < 0) {
		for(int i = 0; i < 3 && prop.x < 0; i++) {
			const int q1 = tri_quads[a * 3 + i];
			if(q1 < 0 || selfConflict(conflicts, q1))
				continue;
			const int2 t1 = node_tris[q1];
			if(t1.x != a && t1.y != a)
				continue;
			const int b = t1.x == a ? t1.y : t1.x;
			const int m = mate[b];
			if(m < 0 || firstIndex(tri_quads + b * 3, q1) < 0)
				continue;
			const int2 tm = node_tris[m];
			const int c = tm.x == b ? tm.y : tm.x;
			for(int j = 0; j < 3; j++) {
				const int q2 = tri_quads[c * 3 + j];
				if(q2 < 0 || q2 == m || selfConflict(conflicts, q2))
					continue;
				const int2 t2 = node_tris[q2];
				if(t2.x != c && t2.y != c)
					continue;
				const int d = t2.x == c ? t2.y : t2.x;
				if(d == a || d == b || mate[d] >= 0 || firstIndex(tri_quads + d * 3, q2) < 0)
					continue;
				prop = make_int4(q1, m, q2, d);
				atomicMin(claim + a, a), atomicMin(claim + b, a), atomicMin(claim + c, a), atomicMin(claim + d, a);
				break;
			}
		}
	}
	proposal[a] = prop;
}
__global__ void k_qg_aug_apply(const int2 *node_tris, const int4 *proposal, const int *claim, int nt,
							   unsigned char *state, int *mate, int *applied) {
	const int a = blockIdx.x * blockDim.x + threadIdx.x;
	bool done = false;
	if(a < nt) {
		const int4 prop = proposal[a];
		if(prop.x >= 0) {
			const int2 t1 = node_tris[prop.x], tm = node_tris[prop.y];
			const int b = t1.x == a ? t1.y : t1.x, c = tm.x == b ? tm.y : tm.x, d = prop.w;
			if(claim[a] == a && claim[b] == a && claim[c] == a && claim[d] == a) {
				state[prop.y] = 1, state[prop.x] = 2, state[prop.z] = 2;
				mate[a] = prop.x, mate[b] = prop.x, mate[c] = prop.z, mate[d] = prop.z;
				done = true;
			}
		}
	}
	const unsigned m = __ballot_sync(0xffffffffu, done);
	if((threadIdx.x & 31) == 0 && m)
		atomicAdd(applied, __popc(m));
}

__global__ void k_qg_mark(const int2 *node_tris, const unsigned char *state, int nq, unsigned char *second_tri) {
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if(q < nq && state[q] == 2)
		second_tri[node_tris[q].y] = 1;
}
__global__ void k_qg_emit_flags(const unsigned char *second_tri, int nt, int *emit) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if(t < nt)
		emit[t] = second_tri[t] ? 0 : 1;
}
__global__ void k_qg_emit(const int *tris, const int *tri_quads, const int4 *node_verts, const unsigned char *state,
						  const unsigned char *second_tri, const int *pos_of, int nt, int4 *out, int *num_degenerate) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	bool degenerate = false;
	if(t < nt && !second_tri[t]) {
		int sel = -1;
		for(int i = 0; i < 3 && sel < 0; i++) {
			const int q = tri_quads[t * 3 + i];
			if(q >= 0 && state[q] == 2)
				sel = q;
		}
		degenerate = sel < 0;
		out[pos_of[t]] = degenerate ? make_int4(tris[t * 3], tris[t * 3 + 1], tris[t * 3 + 2], tris[t * 3 + 2]) : node_verts[sel];
	}
	const unsigned m = __ballot_sync(0xffffffffu, degenerate);
	if((threadIdx.x & 31) == 0 && m)
		atomicAdd(num_degenerate, __popc(m));
}

thread_local std::string g_error;
int failQ(int code, const std::string &what) {
	g_error = what;
	return code;
}

struct Buffers {
	void *ptrs[32];
	int n = 0;
	~Buffers() {
		for(int i = 0; i < n; i++)
			cudaFree(ptrs[i]);
	}
	template <class T> cudaError_t alloc(T **p, size_t count) {
		cudaError_t e = cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
		if(e == cudaSuccess)
			ptrs[n++] = *p;
		return e;
	}
};

void exclusiveScan(int *data, int n, int *tile_sums, cudaStream_t s) {
	const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
	k_qg_scan_tiles<<<tiles, QG_THREADS, 0, s>>>(data, n, tile_sums, false);
	k_qg_scan_sums<<<1, QG_THREADS, 0, s>>>(tile_sums, tiles);
	k_qg_scan_tiles<<<tiles, QG_THREADS, 0, s>>>(data, n, tile_sums, true);
}

} // namespace

extern "C" {

const char *lucid_quadgen_last_error(void) { return g_error.c_str(); }

int lucid_quadgen(const float *positions, int32_t num_verts, const int32_t *tris, int32_t num_tris, float square_weight,
				  int32_t device, int32_t *out_quads, LucidQuadgenResult *result, LucidQuadgenGraph *graph) {
	if(!positions || !tris || !out_quads || !result || num_verts <= 0 || num_tris < 0)
		return failQ(-1, "lucid_quadgen: bad argument");
	if(num_tris > (1 << 28))
		return failQ(-3, "lucid_quadgen: more than 2^28 triangles");
	memset(result, 0, sizeof(*result));
	if(num_tris == 0)
		return 0;
	for(int64_t k = 0; k < (int64_t)num_tris * 3; k++)
		if(tris[k] < 0 || tris[k] >= num_verts)
			return failQ(-1, "lucid_quadgen: vertex index out of range in triangle " + std::to_string(k / 3));
#define CUQ(call)                                                                                                      \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if(e_ != cudaSuccess)                                                                                          \
			return failQ(-2, std::string("lucid_quadgen: ") + #call + ": " + cudaGetErrorString(e_));                  \
	} while(0)
	int dev_count = 0;
	CUQ(cudaGetDeviceCount(&dev_count));
	if(device < 0 || device >= dev_count)
		return failQ(-1, "lucid_quadgen: bad device ordinal");
	CUQ(cudaSetDevice(device));
	const int nt = num_tris, ne = nt * 3;
	uint32_t table = 1024;
	while(table < (uint32_t)ne * 2u)
		table <<= 1;
	const int max_nodes = ne / 2 + 1;
	Buffers b;
	float *d_pos;
	int *d_tris, *d_nb, *d_count, *d_tile, *d_tq, *d_emit, *d_sync;
	unsigned long long *d_keys, *d_nkey;
	uint32_t *d_vals;
	unsigned char *d_mask, *d_state, *d_second;
	int2 *d_ntris;
	int4 *d_nverts, *d_conf, *d_out, *d_prop;
	int *d_mate, *d_claim;
	float *d_sq;
	CUQ(b.alloc(&d_pos, (size_t)num_verts * 3));
	CUQ(b.alloc(&d_tris, (size_t)ne));
	CUQ(b.alloc(&d_nb, (size_t)ne));
	CUQ(b.alloc(&d_keys, (size_t)table));
	CUQ(b.alloc(&d_vals, (size_t)table));
	CUQ(b.alloc(&d_mask, (size_t)nt));
	CUQ(b.alloc(&d_count, (size_t)nt + 1));
	CUQ(b.alloc(&d_tile, (size_t)(nt / SCAN_TILE + 2)));
	CUQ(b.alloc(&d_tq, (size_t)ne));
	CUQ(b.alloc(&d_ntris, (size_t)max_nodes));
	CUQ(b.alloc(&d_nverts, (size_t)max_nodes));
	CUQ(b.alloc(&d_sq, (size_t)max_nodes));
	CUQ(b.alloc(&d_conf, (size_t)max_nodes));
	CUQ(b.alloc(&d_state, (size_t)max_nodes));
	CUQ(b.alloc(&d_nkey, (size_t)max_nodes));
	CUQ(b.alloc(&d_second, (size_t)nt));
	CUQ(b.alloc(&d_emit, (size_t)nt + 1));
	CUQ(b.alloc(&d_out, (size_t)nt));
	CUQ(b.alloc(&d_sync, (size_t)16));
	CUQ(b.alloc(&d_mate, (size_t)nt));
	CUQ(b.alloc(&d_claim, (size_t)nt));
	CUQ(b.alloc(&d_prop, (size_t)nt));
	cudaStream_t s = 0;
	cudaEvent_t ev0, ev1;
	CUQ(cudaEventCreate(&ev0));
	CUQ(cudaEventCreate(&ev1));
	struct EvGuard {
		cudaEvent_t a, b;
		~EvGuard() { cudaEventDestroy(a), cudaEventDestroy(b); }
	} guard{ev0, ev1};
	CUQ(cudaMemcpyAsync(d_pos, positions, (size_t)num_verts * 12, cudaMemcpyHostToDevice, s));
	CUQ(cudaMemcpyAsync(d_tris, tris, (size_t)ne * 4, cudaMemcpyHostToDevice, s));
	CUQ(cudaEventRecord(ev0, s));
	CUQ(cudaMemsetAsync(d_keys, 0xff, (size_t)table * 8, s));
	CUQ(cudaMemsetAsync(d_vals, 0xff, (size_t)table * 4, s));
	CUQ(cudaMemsetAsync(d_second, 0, (size_t)nt, s));
	CUQ(cudaMemsetAsync(d_sync, 0, 8 * 4, s));
	CUQ(cudaMemsetAsync(d_sync + 8, 0xff, 8 * 4, s));
	const int gb_e = (ne + QG_THREADS - 1) / QG_THREADS, gb_t = (nt + QG_THREADS - 1) / QG_THREADS;
	k_qg_edges<<<gb_e, QG_THREADS, 0, s>>>(d_tris, nt, d_keys, d_vals, table - 1);
	k_qg_neighbours<<<gb_e, QG_THREADS, 0, s>>>(d_tris, nt, d_keys, d_vals, table - 1, d_nb);
	k_qg_masks<<<gb_t, QG_THREADS, 0, s>>>(d_tris, d_nb, nt, d_mask, d_count);
	exclusiveScan(d_count, nt, d_tile, s);
	int nq = 0;
	{
		// node count = scanned offset of the last triangle + its own count
		int last_base = 0;
		unsigned char last_mask = 0;
		CUQ(cudaMemcpyAsync(&last_base, d_count + nt - 1, 4, cudaMemcpyDeviceToHost, s));
		CUQ(cudaMemcpyAsync(&last_mask, d_mask + nt - 1, 1, cudaMemcpyDeviceToHost, s));
		CUQ(cudaStreamSynchronize(s));
		nq = last_base + popc3(last_mask);
	}
	k_qg_nodes<<<gb_t, QG_THREADS, 0, s>>>(d_pos, d_tris, d_nb, d_count, d_mask, nt, d_tq, d_ntris, d_nverts, d_sq);
	int rounds = 0;
	if(nq > 0) {
		const int gb_q = (nq + QG_THREADS - 1) / QG_THREADS;
		k_qg_conflicts<<<gb_q, QG_THREADS, 0, s>>>(d_tq, d_ntris, nq, d_conf, d_state);
		int per_sm = 0, sms = 0;
		CUQ(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_qg_select, QG_THREADS, 0));
		CUQ(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
		int grid = std::max(1, std::min(per_sm, 4)) * sms; // a multiple of the SM count, all CTAs co-resident
		grid = std::min(grid, gb_q);
		const int4 *a_conf = d_conf;
		const float *a_sq = d_sq;
		int a_nq = nq;
		void *args[] = {&a_conf, &a_sq, &a_nq, &square_weight, &d_state, &d_nkey, &d_sync};
		CUQ(cudaLaunchCooperativeKernel((void *)k_qg_select, dim3(grid), dim3(QG_THREADS), args, 0, s));
		CUQ(cudaMemsetAsync(d_mate, 0xff, (size_t)nt * 4, s));
		k_qg_mates<<<gb_q, QG_THREADS, 0, s>>>(d_ntris, d_state, nq, d_mate);
		for(int round = 0; round < QG_AUGMENT_ROUNDS; round++) {
			CUQ(cudaMemsetAsync(d_claim, 0x7f, (size_t)nt * 4, s));
			k_qg_aug_propose<<<gb_t, QG_THREADS, 0, s>>>(d_tq, d_ntris, d_conf, d_mate, nt, d_prop, d_claim);
			k_qg_aug_apply<<<gb_t, QG_THREADS, 0, s>>>(d_ntris, d_prop, d_claim, nt, d_state, d_mate, d_sync + 5);
		}
		k_qg_mark<<<gb_q, QG_THREADS, 0, s>>>(d_ntris, d_state, nq, d_second);
	}
	k_qg_emit_flags<<<gb_t, QG_THREADS, 0, s>>>(d_second, nt, d_emit);
	exclusiveScan(d_emit, nt, d_tile, s);
	int *d_ndeg = d_sync + 4;
	k_qg_emit<<<gb_t, QG_THREADS, 0, s>>>(d_tris, d_tq, d_nverts, d_state, d_second, d_emit, nt, d_out, d_ndeg);
	CUQ(cudaEventRecord(ev1, s));
	CUQ(cudaGetLastError());
	int last_pos = 0, sync_host[8];
	unsigned char last_second = 0;
	CUQ(cudaMemcpyAsync(&last_pos, d_emit + nt - 1, 4, cudaMemcpyDeviceToHost, s));
	CUQ(cudaMemcpyAsync(&last_second, d_second + nt - 1, 1, cudaMemcpyDeviceToHost, s));
	CUQ(cudaMemcpyAsync(sync_host, d_sync, 32, cudaMemcpyDeviceToHost, s));
	CUQ(cudaStreamSynchronize(s));
	const int num_quads = last_pos + (last_second ? 0 : 1);
	rounds = sync_host[3];
	CUQ(cudaMemcpy(out_quads, d_out, (size_t)num_quads * 16, cudaMemcpyDeviceToHost));
	float ms = 0.0f;
	CUQ(cudaEventElapsedTime(&ms, ev0, ev1));
	result->num_quads = num_quads, result->num_degenerate = sync_host[4], result->num_nodes = nq, result->rounds = rounds;
	result->num_augmented = sync_host[5];
	result->device_ms = ms;
	if(graph) {
		if(graph->neighbours)
			CUQ(cudaMemcpy(graph->neighbours, d_nb, (size_t)ne * 4, cudaMemcpyDeviceToHost));
		if(graph->tri_quads)
			CUQ(cudaMemcpy(graph->tri_quads, d_tq, (size_t)ne * 4, cudaMemcpyDeviceToHost));
		if(graph->node_tris)
			CUQ(cudaMemcpy(graph->node_tris, d_ntris, (size_t)nq * 8, cudaMemcpyDeviceToHost));
		if(graph->node_verts)
			CUQ(cudaMemcpy(graph->node_verts, d_nverts, (size_t)nq * 16, cudaMemcpyDeviceToHost));
		if(graph->node_conflicts)
			CUQ(cudaMemcpy(graph->node_conflicts, d_conf, (size_t)nq * 16, cudaMemcpyDeviceToHost));
		if(graph->squareness)
			CUQ(cudaMemcpy(graph->squareness, d_sq, (size_t)nq * 4, cudaMemcpyDeviceToHost));
		if(graph->selected)
			CUQ(cudaMemcpy(graph->selected, d_state, (size_t)nq, cudaMemcpyDeviceToHost));
	}
#undef CUQ
	return 0;
}
}
