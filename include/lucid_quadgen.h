/* lucid_quadgen.h -- triangle -> quad pairing on the GPU.
 *
 * Replaces the offline CPU pass the reference runs per mesh when a scene is converted or loaded:
 *   Scene::generateQuads      src/scene.cpp:237-247          (the caller: one pairing per mesh, counts degenerate quads)
 *   triNeighbours             src/quad_generator.cpp:18-45   (src/quad_generator.h:38)
 *   quadNodes                 src/quad_generator.cpp:56-118  (src/quad_generator.h:39-40)
 *   genQuads                  src/quad_generator.cpp:121-201 (src/quad_generator.h:41-43)
 *
 * Contract against the reference:
 *   - the pairing graph (neighbours, quad nodes in creation order, node vertices, squareness bits, conflicts,
 *     tri_quads) is the reference's, word for word;
 *   - the set of nodes that become quads is a maximal independent set of that graph chosen by the same score
 *     (live degree - squareness * weight, lowest first), but by synchronous rounds instead of the reference's
 *     sequential heap loop: every round takes all nodes that beat their live neighbours; then four rounds of
 *     re-pairing along paths unpaired - paired - paired - unpaired win back what the coarser greedy lost.  The result
 *     is a valid pairing (every triangle in exactly one output quad, each quad a reference node or a degenerate
 *     (a, b, c, c)); the quad COUNT stays within 2 % of the reference's (tests/test_quadgen.py states the bound per
 *     mesh; on regular meshes it is equal);
 *   - output order is the reference's: quads in the order of their first triangle (quad_generator.cpp:176-198).
 * Plain pointers and sizes; all pointers are HOST memory, copies are made inside the call.
 */
#ifndef LUCID_QUADGEN_H
#define LUCID_QUADGEN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LucidQuadgenResult {
	int32_t num_quads;		/* entries written to out_quads */
	int32_t num_degenerate; /* of them (a, b, c, c): Mesh::num_degenerate_quads, src/scene.cpp:242-245 */
	int32_t num_nodes;		/* candidate quads (size of the reference's quad_nodes vector) */
	int32_t rounds;			/* selection rounds the device ran */
	float device_ms;		/* kernels only (CUDA events), without the host copies */
	int32_t num_augmented;	/* re-pairings (a, b) + (c, d) applied after the selection: quads gained */
	int32_t reserved[2];
} LucidQuadgenResult;

/* Optional read-back of the pairing graph (any pointer may be null).  Sizes in elements: T = num_tris,
 * N = result->num_nodes (never more than 3 T / 2 + 1; pass room for that).  Layouts are the reference's. */
typedef struct LucidQuadgenGraph {
	int32_t *neighbours;	 /* 3 T   triNeighbours()                     */
	int32_t *tri_quads;		 /* 3 T   second product of quadNodes()       */
	int32_t *node_tris;		 /* 2 N   QuadNode::tris                      */
	int32_t *node_verts;	 /* 4 N   QuadNode::verts                     */
	int32_t *node_conflicts; /* 4 N   QuadNode::conflicts                 */
	float *squareness;		 /* N     QuadNode::squareness                */
	uint8_t *selected;		 /* N     2 = became a quad, 1 = dropped      */
} LucidQuadgenGraph;

/* positions: 3 floats per vertex; tris: 3 vertex indices per triangle (a mesh's Scene::Mesh::tris);
 * square_weight: the reference passes 4.0 (src/scene_setup.cpp:186) or the input scene's quad_squareness;
 * out_quads: room for 4 * num_tris ints.  Returns 0, or -1 bad argument / -2 CUDA error / -3 size limit
 * (text in lucid_quadgen_last_error()). */
int lucid_quadgen(const float *positions, int32_t num_verts, const int32_t *tris, int32_t num_tris, float square_weight,
				  int32_t device, int32_t *out_quads, LucidQuadgenResult *result, LucidQuadgenGraph *graph);

const char *lucid_quadgen_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
