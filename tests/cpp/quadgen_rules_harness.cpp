// quadgen_rules_harness.cpp -- runs the per-element rules of lucid_b200/csrc/quadgen_rules.h (the code the CUDA kernels
// execute) over a whole mesh on the CPU, one element after the other, so that tests/test_quadgen.py can hold them
// against the restatement of the reference's sequential loops without a GPU.  Test infrastructure.
#include "../../lucid_b200/csrc/quadgen_rules.h"
#include <math.h>
#include <stdint.h>
#include <unordered_map>
#include <vector>
using namespace lucid_qg;

extern "C" int quadgen_rules_graph(const float *pos, const int32_t *tris, int nt, int32_t *nb, int32_t *tri_quads, int32_t *node_tris,
								   int32_t *node_verts, int32_t *node_conflicts, float *sq) {
	// neighbours: lowest (triangle, edge) owning a directed edge wins (what the kernel's atomicMin leaves behind)
	std::unordered_map<uint64_t, uint32_t> owner;
	auto key = [](int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
	for(int t = 0; t < nt; t++)
		for(int j = 0; j < 3; j++) {
			auto k = key(tris[t * 3 + j], tris[t * 3 + (j + 1) % 3]);
			auto it = owner.find(k);
			const uint32_t v = (uint32_t)t * 4 + j;
			if(it == owner.end() || v < it->second)
				owner[k] = v;
		}
	for(int t = 0; t < nt; t++)
		for(int j = 0; j < 3; j++) {
			auto it = owner.find(key(tris[t * 3 + (j + 1) % 3], tris[t * 3 + j]));
			nb[t * 3 + j] = (it != owner.end() && (int)(it->second >> 2) != t) ? (int)(it->second >> 2) : -1;
		}
	std::vector<unsigned char> mask(nt);
	std::vector<int> base(nt + 1, 0);
	for(int t = 0; t < nt; t++)
		mask[t] = (unsigned char)createMask(tris, nb, t);
	for(int t = 0; t < nt; t++)
		base[t + 1] = base[t] + popc3(mask[t]);
	const int nq = base[nt];
	for(int t = 0; t < nt; t++)
		for(int i = 0; i < 3; i++) {
			tri_quads[t * 3 + i] = finalTriQuad(nb, base.data(), mask.data(), t, i);
			if((mask[t] >> i) & 1) {
				const int q = nodeId(base.data(), mask.data(), t, i), u = nb[t * 3 + i];
				node_tris[q * 2] = t, node_tris[q * 2 + 1] = u;
				int *v = node_verts + q * 4;
				v[0] = tris[t * 3 + i], v[1] = oppositeVert(tris + t * 3, tris + u * 3), v[2] = tris[t * 3 + (i + 1) % 3],
				v[3] = tris[t * 3 + (i + 2) % 3];
				sq[q] = squareness(pos, v);
			}
		}
	for(int q = 0; q < nq; q++)
		nodeConflicts(tri_quads, q, node_tris[q * 2], node_tris[q * 2 + 1], node_conflicts + q * 4);
	return nq;
}
