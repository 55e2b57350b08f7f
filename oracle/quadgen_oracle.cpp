// quadgen_oracle.cpp -- CPU restatement of the reference's triangle -> quad pairing.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lucid_b200/) may include, link or call this file; only
// tests/, __graft_entry__.smoke() and bench.py's CPU legs load it.
//
// PARITY: pinned.  The reference's own src/quad_generator.cpp compiles here (oracle/build_ref_quadgen.py ->
// oracle/_ref/libref_quadgen.so, sources read where they lie); tests/test_quadgen.py compares this restatement with it
// output for output (neighbours, quad nodes, squareness bits, conflicts, final quads) on seeded meshes, live when
// /root/reference is mounted and through tests/golden/ref_quadgen.json otherwise.
//
// What is restated (file:line under /root/reference):
//   squareness            src/quad_generator.cpp:8-16   (normalize = v / sqrt(dot), libfwk math_base.h:747-762)
//   triNeighbours         src/quad_generator.cpp:18-45  (hash map of directed edges, the first triangle wins)
//   quadNodes             src/quad_generator.cpp:56-118 (node per mutually adjacent pair, creation order, conflicts)
//   genQuads              src/quad_generator.cpp:121-201 (greedy independent set over a min-heap of degree -
//                         squareness * weight with live degree updates; emission in triangle order)
//   fwk::Heap             libfwk/include/fwk/heap.h:11-105 (its sift rules decide the order among equal scores)
// mode 0 is that algorithm.  mode 1 is the round-synchronous greedy the CUDA path computes (lucid_b200/csrc/
// quadgen.cu): same graph, same scores, but every round selects all nodes that beat their live neighbours at once,
// followed by four rounds of re-pairing along unpaired - paired - paired - unpaired paths.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

typedef std::array<int, 3> Tri;

struct Node {
	int tris[2];
	int verts[4];
	int conflicts[4] = {-1, -1, -1, -1};
	float squareness;
	void addConflict(int idx) { // quad_generator.h:18-26
		for(int c : conflicts)
			if(c == idx)
				return;
		for(int &c : conflicts)
			if(c == -1) {
				c = idx;
				break;
			}
	}
	int degree() const {
		int d = 0;
		for(int c : conflicts)
			d += c != -1;
		return d;
	}
};

float squarenessOf(const float *pos, const int v[4]) { // quad_generator.cpp:8-16
	float e[4][3];
	for(int i = 0; i < 4; i++) {
		const int j = (i + 1) & 3;
		float d[3];
		for(int k = 0; k < 3; k++)
			d[k] = pos[v[j] * 3 + k] - pos[v[i] * 3 + k];
		const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
		for(int k = 0; k < 3; k++)
			e[i][k] = d[k] / len;
	}
	float out = 0.0f;
	for(int i = 0; i < 4; i++) {
		const int j = (i + 1) & 3;
		out += std::fabs(e[i][0] * e[j][0] + e[i][1] * e[j][1] + e[i][2] * e[j][2]);
	}
	return (4.0f - out) * 0.25f;
}

struct Graph {
	std::vector<Tri> nb, tri_quads;
	std::vector<Node> nodes;
};

Graph buildGraph(const float *pos, const Tri *tris, int nt) {
	Graph g;
	g.nb.assign(nt, Tri{{-1, -1, -1}});
	g.tri_quads.assign(nt, Tri{{-1, -1, -1}});
	// triNeighbours: emplace keeps the first triangle that owns a directed edge
	std::unordered_map<uint64_t, int> edge_tri;
	edge_tri.reserve((size_t)nt * 4);
	auto key = [](int a, int b) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; };
	for(int i = 0; i < nt; i++)
		for(int j = 0; j < 3; j++)
			edge_tri.emplace(key(tris[i][j], tris[i][j == 2 ? 0 : j + 1]), i);
	for(int i = 0; i < nt; i++)
		for(int j = 0; j < 3; j++) {
			auto it = edge_tri.find(key(tris[i][j == 2 ? 0 : j + 1], tris[i][j]));
			if(it != edge_tri.end() && it->second != i)
				g.nb[i][j] = it->second;
		}
	// quadNodes
	for(int idx0 = 0; idx0 < nt; idx0++) {
		const Tri &tri0 = tris[idx0];
		for(int i = 0; i < 3; i++) {
			const int idx1 = g.nb[idx0][i];
			if(idx1 == -1 || g.tri_quads[idx0][i] != -1)
				continue;
			int j = -1;
			for(int k = 0; k < 3 && j == -1; k++)
				if(g.nb[idx1][k] == idx0)
					j = k;
			if(j == -1)
				continue;
			int opposite = -1;
			for(int ov : tris[idx1])
				if(ov != tri0[0] && ov != tri0[1] && ov != tri0[2]) {
					opposite = ov;
					break;
				}
			if(opposite == -1)
				continue;
			const int q = (int)g.nodes.size();
			Node n;
			n.tris[0] = idx0, n.tris[1] = idx1;
			n.verts[0] = tri0[i], n.verts[1] = opposite, n.verts[2] = tri0[(i + 1) % 3], n.verts[3] = tri0[(i + 2) % 3];
			n.squareness = squarenessOf(pos, n.verts);
			g.nodes.push_back(n);
			g.tri_quads[idx0][i] = q;
			g.tri_quads[idx1][j] = q;
		}
	}
	for(const Tri &tq : g.tri_quads)
		for(int i = 0; i < 3; i++) {
			const int q0 = tq[i], q1 = tq[i == 2 ? 0 : i + 1];
			if(q0 != -1 && q1 != -1) {
				g.nodes[q0].addConflict(q1);
				g.nodes[q1].addConflict(q0);
			}
		}
	return g;
}

// fwk::Heap<float> (libfwk/include/fwk/heap.h): parent of pos is pos / 2, children 2 pos and 2 pos + 1 -- slot 0 is
// its own left child.  Restated literally; the order among equal scores follows from it.
struct Heap {
	std::vector<std::pair<float, int>> heap;
	std::vector<int> indices;
	int size = 0;
	explicit Heap(int n) : heap(n), indices(n, -1) {}
	bool less(int a, int b) const { return heap[a].first < heap[b].first; }
	void updateIndex(int pos) { indices[heap[pos].second] = pos; }
	void heapify(int pos) {
		const int l = pos * 2, r = pos * 2 + 1;
		int smallest = l < size && less(l, pos) ? l : pos;
		if(r < size && less(r, smallest))
			smallest = r;
		if(smallest != pos) {
			std::swap(heap[pos], heap[smallest]);
			updateIndex(pos);
			updateIndex(smallest);
			heapify(smallest);
		}
	}
	std::pair<float, int> extractMin() {
		auto mn = heap[0];
		indices[mn.second] = -1;
		if(size > 1) {
			heap[0] = heap[--size];
			updateIndex(0);
			heapify(0);
		} else
			size = 0;
		return mn;
	}
	void update(int key_idx, float value) {
		int pos = indices[key_idx];
		if(pos == -1) {
			pos = size++;
			heap[pos].first = value;
		}
		if(value > heap[pos].first) {
			heap[pos].first = value;
			heapify(pos);
			return;
		}
		while(pos > 0 && value < heap[pos / 2].first) {
			heap[pos] = heap[pos / 2];
			updateIndex(pos);
			pos = pos / 2;
		}
		heap[pos] = {value, key_idx};
		updateIndex(pos);
	}
};

inline uint32_t hash32(uint32_t x) { // lowbias32
	x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
	return x;
}
inline uint32_t sortableFloat(float f) {
	if(f != f)
		return 0xffffffffu; // NaN (zero-length edge) sorts last
	uint32_t u;
	memcpy(&u, &f, 4);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float unsortableFloat(uint32_t s) {
	const uint32_t u = (s & 0x80000000u) ? (s & 0x7fffffffu) : ~s;
	float f;
	memcpy(&f, &u, 4);
	return f;
}

// visited: 0 live, 1 removed, 2 selected; visited_tris: the second triangle of every selected node
void selectReference(const Graph &g, float square_weight, std::vector<uint8_t> &visited, std::vector<uint8_t> &visited_tris) {
	const int nq = (int)g.nodes.size();
	std::vector<int> degree(nq);
	for(int q = 0; q < nq; q++)
		degree[q] = g.nodes[q].degree();
	auto score = [&](int q) { return float(degree[q]) - g.nodes[q].squareness * square_weight; };
	Heap heap(nq);
	for(int q = 0; q < nq; q++)
		heap.update(q, score(q));
	while(heap.size > 0) {
		const int q = heap.extractMin().second;
		if(visited[q])
			continue;
		visited[q] = 2;
		visited_tris[g.nodes[q].tris[1]] = 1;
		for(int n : g.nodes[q].conflicts)
			if(n != -1) {
				visited[n] = 1;
				for(int n2 : g.nodes[n].conflicts)
					if(n2 != -1 && !visited[n2]) {
						degree[n2]--;
						heap.update(n2, score(n2));
					}
			}
	}
}

int selectRounds(const Graph &g, float square_weight, std::vector<uint8_t> &visited) {
	const int nq = (int)g.nodes.size();
	// a node in conflict with itself (two triangles joined along two edges) can never be used
	for(int q = 0; q < nq; q++)
		for(int c : g.nodes[q].conflicts)
			if(c == q)
				visited[q] = 1;
	std::vector<uint64_t> key(nq);
	int rounds = 0;
	while(true) {
		int live = 0;
		for(int q = 0; q < nq; q++) {
			if(visited[q])
				continue;
			live++;
			int deg = 0;
			for(int c : g.nodes[q].conflicts)
				deg += c != -1 && !visited[c];
			const float score = float(deg) - g.nodes[q].squareness * square_weight;
			key[q] = ((uint64_t)sortableFloat(score) << 32) | hash32((uint32_t)q);
		}
		if(live == 0)
			break;
		rounds++;
		std::vector<int> chosen;
		// only nodes within the score window above the lowest live score (quadgen_rules.h: QG_SELECT_WINDOW)
		uint32_t lowest = 0xffffffffu;
		for(int q = 0; q < nq; q++)
			if(!visited[q])
				lowest = std::min(lowest, (uint32_t)(key[q] >> 32));
		uint32_t limit = lowest;
		if(lowest != 0xffffffffu)
			limit = std::max(lowest, sortableFloat(unsortableFloat(lowest) + 2.0f));
		for(int q = 0; q < nq; q++) {
			if(visited[q] || (uint32_t)(key[q] >> 32) > limit)
				continue;
			bool best = true;
			for(int c : g.nodes[q].conflicts)
				if(c != -1 && !visited[c] && (key[c] < key[q] || (key[c] == key[q] && c < q)))
					best = false;
			if(best)
				chosen.push_back(q);
		}
		for(int q : chosen)
			visited[q] = 2;
		for(int q : chosen)
			for(int c : g.nodes[q].conflicts)
				if(c != -1 && visited[c] == 0)
					visited[c] = 1;
	}
	return rounds;
}

// the re-pairing rounds of lucid_b200/csrc/quadgen.cu (k_qg_aug_propose / k_qg_aug_apply), one element after the other
int augmentRounds(const Graph &g, int nt, std::vector<uint8_t> &visited, int max_rounds) {
	const int nq = (int)g.nodes.size();
	auto selfConflict = [&](int q) {
		for(int c : g.nodes[q].conflicts)
			if(c == q)
				return true;
		return false;
	};
	auto listed = [&](int t, int q) { return g.tri_quads[t][0] == q || g.tri_quads[t][1] == q || g.tri_quads[t][2] == q; };
	std::vector<int> mate(nt, -1);
	for(int q = 0; q < nq; q++)
		if(visited[q] == 2)
			mate[g.nodes[q].tris[0]] = q, mate[g.nodes[q].tris[1]] = q;
	int applied = 0;
	struct Prop {
		int q1, m, q2, b, c, d;
	};
	for(int round = 0; round < max_rounds; round++) {
		std::vector<Prop> prop(nt, Prop{-1, -1, -1, -1, -1, -1});
		std::vector<int> claim(nt, 0x7f7f7f7f);
		for(int a = 0; a < nt; a++) {
			if(mate[a] >= 0)
				continue;
			for(int i = 0; i < 3 && prop[a].q1 < 0; i++) {
				const int q1 = g.tri_quads[a][i];
				if(q1 < 0 || selfConflict(q1))
					continue;
				const Node &n1 = g.nodes[q1];
				if(n1.tris[0] != a && n1.tris[1] != a)
					continue;
				const int b = n1.tris[0] == a ? n1.tris[1] : n1.tris[0];
				const int m = mate[b];
				if(m < 0 || !listed(b, q1))
					continue;
				const int c = g.nodes[m].tris[0] == b ? g.nodes[m].tris[1] : g.nodes[m].tris[0];
				for(int j = 0; j < 3; j++) {
					const int q2 = g.tri_quads[c][j];
					if(q2 < 0 || q2 == m || selfConflict(q2))
						continue;
					const Node &n2 = g.nodes[q2];
					if(n2.tris[0] != c && n2.tris[1] != c)
						continue;
					const int d = n2.tris[0] == c ? n2.tris[1] : n2.tris[0];
					if(d == a || d == b || mate[d] >= 0 || !listed(d, q2))
						continue;
					prop[a] = Prop{q1, m, q2, b, c, d};
					for(int t : {a, b, c, d})
						claim[t] = std::min(claim[t], a);
					break;
				}
			}
		}
		for(int a = 0; a < nt; a++) {
			const Prop &p = prop[a];
			if(p.q1 < 0 || claim[a] != a || claim[p.b] != a || claim[p.c] != a || claim[p.d] != a)
				continue;
			visited[p.m] = 1, visited[p.q1] = 2, visited[p.q2] = 2;
			mate[a] = mate[p.b] = p.q1, mate[p.c] = mate[p.d] = p.q2;
			applied++;
		}
	}
	return applied;
}

} // namespace

extern "C" {

// out_quads: room for 4 * num_tris ints; counts (8 ints): [quads, degenerate, nodes, rounds, augmented]
// optional graph products (may be null): neighbours 3T, tri_quads 3T, node_tris 2N', node_verts 4N', node_conflicts 4N',
// squareness N' (N' <= 3T / 2 rounded up; pass room for 2T nodes), selected N' bytes
int quadgen_oracle(const float *positions, int num_verts, const int32_t *tris_in, int num_tris, float square_weight, int mode,
				   int32_t *out_quads, int32_t *counts, int32_t *neighbours, int32_t *tri_quads, int32_t *node_tris,
				   int32_t *node_verts, int32_t *node_conflicts, float *squareness, uint8_t *selected) {
	(void)num_verts;
	const Tri *tris = reinterpret_cast<const Tri *>(tris_in);
	Graph g = buildGraph(positions, tris, num_tris);
	const int nq = (int)g.nodes.size();
	std::vector<uint8_t> visited(nq, 0), visited_tris(num_tris, 0);
	int rounds = 0, augmented = 0;
	if(mode == 0) {
		selectReference(g, square_weight, visited, visited_tris);
	} else {
		rounds = selectRounds(g, square_weight, visited);
		augmented = augmentRounds(g, num_tris, visited, 16);
		for(int q = 0; q < nq; q++)
			if(visited[q] == 2)
				visited_tris[g.nodes[q].tris[1]] = 1;
	}
	// emission in triangle order (quad_generator.cpp:176-198)
	int nout = 0, ndeg = 0;
	for(int t = 0; t < num_tris; t++) {
		if(visited_tris[t])
			continue;
		int sel = -1;
		for(int q : g.tri_quads[t])
			if(q != -1 && visited[q] == 2) {
				sel = q;
				break;
			}
		int32_t *o = out_quads + (size_t)nout * 4;
		if(sel == -1) {
			o[0] = tris[t][0], o[1] = tris[t][1], o[2] = tris[t][2], o[3] = tris[t][2];
			ndeg++;
		} else {
			for(int k = 0; k < 4; k++)
				o[k] = g.nodes[sel].verts[k];
		}
		nout++;
	}
	counts[0] = nout, counts[1] = ndeg, counts[2] = nq, counts[3] = rounds, counts[4] = augmented;
	if(neighbours)
		memcpy(neighbours, g.nb.data(), (size_t)num_tris * 12);
	if(tri_quads)
		memcpy(tri_quads, g.tri_quads.data(), (size_t)num_tris * 12);
	for(int q = 0; q < nq; q++) {
		if(node_tris)
			node_tris[q * 2] = g.nodes[q].tris[0], node_tris[q * 2 + 1] = g.nodes[q].tris[1];
		if(node_verts)
			memcpy(node_verts + q * 4, g.nodes[q].verts, 16);
		if(node_conflicts)
			memcpy(node_conflicts + q * 4, g.nodes[q].conflicts, 16);
		if(squareness)
			squareness[q] = g.nodes[q].squareness;
		if(selected)
			selected[q] = visited[q];
	}
	return 0;
}
}
