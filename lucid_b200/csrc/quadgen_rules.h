// quadgen_rules.h -- the per-element rules of the triangle -> quad pairing (quadgen.cu), as host/device functions so
// that tests/cpp/test_quadgen_rules.cpp can run the very same code over a mesh on the CPU.
//
// The reference builds its pairing graph with sequential loops whose results depend on the order of the iterations
// (src/quad_generator.cpp:56-118: a node is created at (triangle, edge) unless an earlier iteration already wrote that
// slot; a later iteration may overwrite a slot).  Below, the outcome of those loops is written as closed rules that
// read only the two triangles involved, so every (triangle, edge) and every node can be evaluated independently:
//
//   * nb[t][i]        the triangle that owns the reversed directed edge (the lowest one if several do), or -1
//   * creates(t, i)   with u = nb[t][i]: u lists t as a neighbour, u has a vertex outside t, and the slot was not
//                     written before the loop reached (t, i).  Only u < t can have written it (a creation at (u, j')
//                     writes slot firstIndex(nb[t], u) of t), and it did exactly when t has a vertex outside u.
//   * node id         rank of (t, i) among the creating pairs in (t, i) order
//   * tri_quads[t][i] the id written last: creations at (u, j') with nb[u][j'] == t write slot firstIndex(nb[t], u) of
//                     t; for u < t they come before (t, i) (which then does not create), for u > t after it
//   * conflicts       the other nodes on the node's two triangles, in the order the reference's loop meets them
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#ifdef __CUDACC__
#define QG_HD __host__ __device__ __forceinline__
#else
#define QG_HD inline
#endif

namespace lucid_qg {

QG_HD int firstIndex(const int *v3, int value) {
	return v3[0] == value ? 0 : v3[1] == value ? 1 : v3[2] == value ? 2 : -1;
}
// first vertex of triangle b that is not a vertex of triangle a, or -1 (quad_generator.cpp:76-83)
QG_HD int oppositeVert(const int *a, const int *b) {
	for(int k = 0; k < 3; k++)
		if(b[k] != a[0] && b[k] != a[1] && b[k] != a[2])
			return b[k];
	return -1;
}
QG_HD bool creates(const int *tris, const int *nb, int t, int i) {
	const int u = nb[t * 3 + i];
	if(u < 0)
		return false;
	if(firstIndex(nb + u * 3, t) < 0 || oppositeVert(tris + t * 3, tris + u * 3) < 0)
		return false;
	if(u > t || i != firstIndex(nb + t * 3, u))
		return true;
	return oppositeVert(tris + u * 3, tris + t * 3) < 0;
}
QG_HD int createMask(const int *tris, const int *nb, int t) {
	return (creates(tris, nb, t, 0) ? 1 : 0) | (creates(tris, nb, t, 1) ? 2 : 0) | (creates(tris, nb, t, 2) ? 4 : 0);
}
QG_HD int popc3(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }
// base[t]: number of nodes created by triangles before t; mask[t]: createMask
QG_HD int nodeId(const int *base, const unsigned char *mask, int t, int i) { return base[t] + popc3(mask[t] & ((1 << i) - 1)); }
QG_HD int finalTriQuad(const int *nb, const int *base, const unsigned char *mask, int t, int i) {
	int q = ((mask[t] >> i) & 1) ? nodeId(base, mask, t, i) : -1;
	const int u = nb[t * 3 + i];
	if(u >= 0 && i == firstIndex(nb + t * 3, u))
		for(int j = 2; j >= 0; j--)
			if(nb[u * 3 + j] == t && ((mask[u] >> j) & 1))
				return nodeId(base, mask, u, j);
	return q;
}
QG_HD void addConflict(int *c4, int idx) { // quad_generator.h:18-26
	if(c4[0] == idx || c4[1] == idx || c4[2] == idx || c4[3] == idx)
		return;
	for(int k = 0; k < 4; k++)
		if(c4[k] == -1) {
			c4[k] = idx;
			return;
		}
}
// conflicts of node q on triangles (a, b) = (creator, neighbour); tq = tri_quads
QG_HD void nodeConflicts(const int *tq, int q, int a, int b, int *c4) {
	c4[0] = c4[1] = c4[2] = c4[3] = -1;
	const int lo = a < b ? a : b, hi = a < b ? b : a;
	for(int pass = 0; pass < 2; pass++) {
		const int *s = tq + (pass == 0 ? lo : hi) * 3;
		for(int i = 0; i < 3; i++) {
			const int q0 = s[i], q1 = s[i == 2 ? 0 : i + 1];
			if(q0 < 0 || q1 < 0)
				continue;
			if(q0 == q)
				addConflict(c4, q1);
			if(q1 == q)
				addConflict(c4, q0);
		}
		if(lo == hi)
			break;
	}
}
#ifdef __CUDA_ARCH__
#define QG_SQRT(x) __fsqrt_rn(x)
#define QG_DIV(a, b) __fdiv_rn(a, b)
#define QG_MUL(a, b) __fmul_rn(a, b)
#define QG_ADD(a, b) __fadd_rn(a, b)
#else
#define QG_SQRT(x) sqrtf(x)
#define QG_DIV(a, b) ((a) / (b))
#define QG_MUL(a, b) ((a) * (b))
#define QG_ADD(a, b) ((a) + (b))
#endif
// quad_generator.cpp:8-16; one rounding per operation in the reference's association (no contraction)
QG_HD float squareness(const float *pos, const int *v4) {
	float e[4][3];
	for(int i = 0; i < 4; i++) {
		const int j = (i + 1) & 3;
		const float dx = QG_ADD(pos[v4[j] * 3 + 0], -pos[v4[i] * 3 + 0]);
		const float dy = QG_ADD(pos[v4[j] * 3 + 1], -pos[v4[i] * 3 + 1]);
		const float dz = QG_ADD(pos[v4[j] * 3 + 2], -pos[v4[i] * 3 + 2]);
		const float len = QG_SQRT(QG_ADD(QG_ADD(QG_MUL(dx, dx), QG_MUL(dy, dy)), QG_MUL(dz, dz)));
		e[i][0] = QG_DIV(dx, len), e[i][1] = QG_DIV(dy, len), e[i][2] = QG_DIV(dz, len);
	}
	float out = 0.0f;
	for(int i = 0; i < 4; i++) {
		const int j = (i + 1) & 3;
		const float d = QG_ADD(QG_ADD(QG_MUL(e[i][0], e[j][0]), QG_MUL(e[i][1], e[j][1])), QG_MUL(e[i][2], e[j][2]));
		out = QG_ADD(out, fabsf(d));
	}
	return QG_MUL(QG_ADD(4.0f, -out), 0.25f);
}
QG_HD uint32_t hash32(uint32_t x) { // lowbias32
	x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
	return x;
}
// order-preserving integer image of a score; NaN (a quad with a zero-length edge) sorts last on every machine
QG_HD uint32_t sortableScore(float score) {
	if(score != score)
		return 0xffffffffu;
	uint32_t u;
#ifdef __CUDA_ARCH__
	u = __float_as_uint(score);
#else
	memcpy(&u, &score, 4);
#endif
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
QG_HD float scoreOf(uint32_t sortable) {
	const uint32_t u = (sortable & 0x80000000u) ? (sortable & 0x7fffffffu) : ~sortable;
	float f;
#ifdef __CUDA_ARCH__
	f = __uint_as_float(u);
#else
	memcpy(&f, &u, 4);
#endif
	return f;
}
// the selection key of a live node: score = live degree - squareness * weight (quad_generator.cpp:141-144) as an
// order-preserving integer, then a hash of the node id (ties between equal scores must not follow the node order,
// or a regular mesh becomes one long dependency chain), then the id itself (compared by the caller)
QG_HD uint64_t nodeKey(int q, int live_degree, float sq, float weight) {
	const float score = QG_ADD(float(live_degree), -QG_MUL(sq, weight));
	return ((uint64_t)sortableScore(score) << 32) | hash32((uint32_t)q);
}
// A round only takes nodes whose score lies within this window above the lowest live score: the reference's heap
// always takes the globally lowest score, and the closer the rounds follow that order the closer the number of quads
// (a window of 2 ends within 0.2 % of the reference on irregular meshes, an unbounded one 3 % above).  The lowest
// node itself is always inside the window, so every round selects at least one node.
#define QG_SELECT_WINDOW 2.0f
QG_HD uint32_t windowLimit(uint32_t min_sortable) {
	if(min_sortable == 0xffffffffu)
		return min_sortable;
	const uint32_t lim = sortableScore(QG_ADD(scoreOf(min_sortable), QG_SELECT_WINDOW));
	return lim > min_sortable ? lim : min_sortable;
}

} // namespace lucid_qg
