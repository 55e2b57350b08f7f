/* lucid_host.h -- host-side input preparation for the render call (no CUDA involved).
 *
 * Mirrors what the reference does on the CPU before it records its compute dispatches:
 *   LucidRenderer::uploadInstances   src/lucid_renderer.cpp:352-429
 *   LucidRenderer::setupInputData    src/lucid_renderer.cpp:431-451
 *   FrustumInfo / SceneLighting      src/shading.cpp:26-74
 *   fwk::perspective / lookAt        libfwk/src/math/matrix4.cpp:204-226
 *   fwk::Camera / OrbitingCamera     libfwk/src/gfx/camera.cpp:18-66, orbiting_camera.cpp:31-43
 *   fwk::Frustum::cornerRays         libfwk/src/math/frustum.cpp:17-62
 */
#ifndef LUCID_HOST_H
#define LUCID_HOST_H

#include "lucid_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fwk::Camera + CameraParams (perspective only) */
typedef struct LucidCamera {
	float pos[3], target[3], target_up[3];
	float fov_rad; /* vertical */
	float z_near, z_far;
	int viewport_width, viewport_height;
} LucidCamera;

/* SceneDrawCall, src/lucid_base.h:48-54 (bbox and triangle ranges are not used by this path) */
typedef struct LucidDrawCall {
	int32_t material_id;
	int32_t num_quads, quad_offset;
	uint32_t opts; /* DrawCallOpts bits == LUCID_INST_* */
} LucidDrawCall;

/* the parts of SceneMaterial that uploadInstances reads */
typedef struct LucidMaterial {
	float diffuse[3];
	float opacity;
	float uv_rect[4]; /* albedo map uv_rect: min.x min.y size.x size.y */
} LucidMaterial;

/* OrbitingCamera(center, distance, rot_horiz, rot_vert).toCamera(params); the defaults of the
 * reference application are fov 60 degrees, depth 1/16 .. 1024 (src/lucid_app.cpp:95) */
void lucid_host_orbit_camera(const float center[3], float distance, float rot_horiz, float rot_vert,
							 float fov_rad, float z_near, float z_far, int width, int height,
							 LucidCamera *out);

/* SceneLighting::makeDefault() converted to shader::Lighting */
void lucid_host_default_lighting(LucidLighting *out);

/* setupInputData: frustum, view_proj = proj * view, lighting, background, flags.
 * num_instances / instance_packet_size are taken from the CURRENT call (the reference reads the
 * previous frame's values, lucid_renderer.cpp:444-446 -- a known bug). */
void lucid_host_make_config(const LucidCamera *camera, const LucidLighting *lighting,
							const float background_rgba[4], int backface_culling, int num_instances,
							int max_dispatches, LucidConfig *out);

/* view and projection matrices, column major (16 floats each), for tests */
void lucid_host_camera_matrices(const LucidCamera *camera, float view[16], float proj[16]);

/* uploadInstances: one instance per <= 1024-quad slice of each draw call, RGBA8 colour
 * (truncating x255) and uv_rect per instance.  Returns the number of instances written
 * (clamped to capacity and LUCID_MAX_INSTANCES), or -1 on bad arguments. */
int lucid_host_build_instances(const LucidDrawCall *dcs, int num_dcs, const LucidMaterial *materials,
							   int num_materials, LucidInstanceData *out_instances,
							   uint32_t *out_colors, float *out_uv_rects, int capacity);

/* instance_packet_size = clamp(num_instances / (max_dispatches / 2), 1, 2) */
int lucid_host_packet_size(int num_instances, int max_dispatches);

/* Spatially coherent instances (the goal of meshPartition, src/meshlet.cpp:68-222): the order in which
 * the num_quads quads (4 vertex indices each) of one draw call should be listed so that the 1024-quad instances
 * uploadInstances cuts (src/lucid_renderer.cpp:352-429) have small bounding boxes -- ascending 30-bit Morton code of
 * the quad centroids inside their bounding box, equal codes in input order.  out_order[i] = index of the quad that
 * goes to place i.  Returns 0, or -1 on bad arguments (null pointers, an index at or above num_verts). */
int lucid_host_cluster_order(const float *positions, int32_t num_verts, const uint32_t *quads, int32_t num_quads,
							 int32_t *out_order);

#ifdef __cplusplus
}
#endif
#endif
