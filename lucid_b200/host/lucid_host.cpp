// Host-side input preparation for the render call; see include/lucid_host.h for the reference
// functions each entry point stands in for.  Plain C++17, no CUDA, no third-party math library.

#include "../../include/lucid_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct F3 {
	float x, y, z;
};
F3 f3(const float *p) { return F3{p[0], p[1], p[2]}; }
F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
F3 operator-(F3 a) { return F3{-a.x, -a.y, -a.z}; }
float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
F3 cross(F3 a, F3 b) {
	return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
F3 normalize(F3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
LucidVec4 v4(F3 a, float w) { return LucidVec4{a.x, a.y, a.z, w}; }

// camera basis exactly as fwk::Camera builds it: up() = cross(cross(forward, target_up), forward),
// lookAt(): side = normalize(cross(front, up)), rows (side, cross(side, front), -front)
struct Basis {
	F3 eye, side, up, front;
};
Basis cameraBasis(const LucidCamera &cam) {
	Basis b;
	b.eye = f3(cam.pos);
	b.front = normalize(f3(cam.target) - b.eye);
	F3 right = cross(b.front, f3(cam.target_up));
	F3 up = cross(right, b.front);
	b.side = normalize(cross(b.front, up));
	b.up = cross(b.side, b.front);
	return b;
}

// column-major 4x4
struct M4 {
	float c[4][4];
};
M4 viewMatrix(const Basis &b) {
	M4 m;
	F3 rows[3] = {b.side, b.up, -b.front};
	for(int r = 0; r < 3; r++) {
		m.c[0][r] = rows[r].x, m.c[1][r] = rows[r].y, m.c[2][r] = rows[r].z;
		m.c[3][r] = -dot(rows[r], b.eye);
	}
	m.c[0][3] = m.c[1][3] = m.c[2][3] = 0.0f, m.c[3][3] = 1.0f;
	return m;
}
M4 projMatrix(const LucidCamera &cam) {
	M4 m;
	memset(&m, 0, sizeof(m));
	float aspect = float(cam.viewport_width) / float(cam.viewport_height);
	float ctg = 1.0f / std::tan(0.5f * cam.fov_rad);
	float z_diff = cam.z_far - cam.z_near;
	m.c[0][0] = ctg / aspect;
	m.c[1][1] = -ctg; // y is flipped: view-space up lands on screen row 0
	m.c[2][2] = -(cam.z_far + cam.z_near) / z_diff;
	m.c[2][3] = -1.0f;
	m.c[3][2] = -(2.0f * cam.z_near * cam.z_far) / z_diff;
	return m;
}
M4 mul(const M4 &a, const M4 &b) {
	M4 o;
	for(int c = 0; c < 4; c++)
		for(int r = 0; r < 4; r++) {
			float s = 0.0f;
			for(int k = 0; k < 4; k++)
				s += a.c[k][r] * b.c[c][k];
			o.c[c][r] = s;
		}
	return o;
}

} // namespace

extern "C" {

void lucid_host_orbit_camera(const float center[3], float distance, float rot_horiz, float rot_vert,
							 float fov_rad, float z_near, float z_far, int width, int height,
							 LucidCamera *out) {
	float sh = std::sin(rot_horiz), ch = std::cos(rot_horiz);
	F3 forward{-sh, 0.0f, ch};
	F3 right{ch, 0.0f, sh};
	float sv = std::sin(rot_vert), cv = std::cos(rot_vert);
	// Rodrigues rotation of forward around right by rot_vert
	forward = forward * cv + cross(right, forward) * sv + right * (dot(right, forward) * (1.0f - cv));
	F3 c = f3(center);
	F3 pos = c - forward * distance;
	F3 up = cross(forward, right);
	out->pos[0] = pos.x, out->pos[1] = pos.y, out->pos[2] = pos.z;
	out->target[0] = c.x, out->target[1] = c.y, out->target[2] = c.z;
	out->target_up[0] = up.x, out->target_up[1] = up.y, out->target_up[2] = up.z;
	out->fov_rad = fov_rad, out->z_near = z_near, out->z_far = z_far;
	out->viewport_width = width, out->viewport_height = height;
}

void lucid_host_default_lighting(LucidLighting *out) {
	memset(out, 0, sizeof(*out));
	out->sun_dir = LucidVec4{0.842121f, -0.300567f, -0.447763f, 0.0f};
	out->sun_color = LucidVec4{0.8f, 0.8f, 0.8f, 1.0f};
	out->sun_power = 2.5f;
	out->ambient_color = LucidVec4{0.8f, 0.8f, 0.6f, 1.0f};
	out->ambient_power = 0.4f;
}

void lucid_host_camera_matrices(const LucidCamera *camera, float view[16], float proj[16]) {
	M4 v = viewMatrix(cameraBasis(*camera)), p = projMatrix(*camera);
	memcpy(view, &v, sizeof(v));
	memcpy(proj, &p, sizeof(p));
}

static uint32_t spreadBits3(uint32_t v) { // 10 bits -> every bit followed by two zero bits
	v &= 0x3ffu;
	v = (v | (v << 16)) & 0x030000ffu;
	v = (v | (v << 8)) & 0x0300f00fu;
	v = (v | (v << 4)) & 0x030c30c3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

int lucid_host_cluster_order(const float *positions, int32_t num_verts, const uint32_t *quads, int32_t num_quads,
							 int32_t *out_order) {
	if(num_quads < 0 || num_verts < 0 || (num_quads > 0 && (!positions || !quads || !out_order)))
		return -1;
	if(num_quads == 0)
		return 0;
	std::vector<float> cen((size_t)num_quads * 3);
	float lo[3], hi[3];
	for(int32_t q = 0; q < num_quads; q++) {
		double sum[3] = {0.0, 0.0, 0.0};
		for(int k = 0; k < 4; k++) {
			const uint32_t vi = quads[(size_t)q * 4 + k];
			if(vi >= (uint32_t)num_verts)
				return -1;
			for(int a = 0; a < 3; a++)
				sum[a] += (double)positions[(size_t)vi * 3 + a];
		}
		for(int a = 0; a < 3; a++) {
			const float c = (float)(sum[a] / 4.0);
			cen[(size_t)q * 3 + a] = c;
			lo[a] = q == 0 ? c : std::min(lo[a], c);
			hi[a] = q == 0 ? c : std::max(hi[a], c);
		}
	}
	std::vector<uint32_t> keys((size_t)num_quads);
	for(int32_t q = 0; q < num_quads; q++) {
		uint32_t key = 0;
		for(int a = 0; a < 3; a++) {
			const float extent = std::max(hi[a] - lo[a], 1e-30f);
			float t = (cen[(size_t)q * 3 + a] - lo[a]) / extent * 1024.0f;
			t = std::min(std::max(t, 0.0f), 1023.0f);
			key |= spreadBits3((uint32_t)t) << a;
		}
		keys[q] = key;
	}
	for(int32_t q = 0; q < num_quads; q++)
		out_order[q] = q;
	std::stable_sort(out_order, out_order + num_quads, [&](int32_t a, int32_t b) { return keys[a] < keys[b]; });
	return 0;
}

int lucid_host_packet_size(int num_instances, int max_dispatches) {
	int half = std::max(1, max_dispatches / 2);
	return std::min(std::max(num_instances / half, 1), 2);
}

void lucid_host_make_config(const LucidCamera *camera, const LucidLighting *lighting,
							const float background_rgba[4], int backface_culling, int num_instances,
							int max_dispatches, LucidConfig *out) {
	memset(out, 0, sizeof(*out));
	Basis b = cameraBasis(*camera);
	M4 view = viewMatrix(b), proj = projMatrix(*camera);
	M4 vp = mul(proj, view);
	memcpy(out->view_proj_matrix, &vp, sizeof(vp));

	// Corner rays of the view-space frustum: intersections of the side planes taken from the
	// projection matrix (left/up, down/left, right/down, up/right), unit length, through the eye.
	float aspect = float(camera->viewport_width) / float(camera->viewport_height);
	float ctg = 1.0f / std::tan(0.5f * camera->fov_rad);
	const float sx[4] = {-1.0f, -1.0f, 1.0f, 1.0f};
	const float sy[4] = {1.0f, -1.0f, -1.0f, 1.0f};
	F3 dirs[4];
	for(int i = 0; i < 4; i++) {
		F3 d = normalize(F3{sx[i] * aspect, sy[i], -ctg});
		// view -> world: the rotation part of the inverse view matrix
		dirs[i] = b.side * d.x + b.up * d.y + (-b.front) * d.z;
		out->frustum.ws_dirs[i] = v4(dirs[i], 0.0f);
		out->frustum.ws_origins[i] = v4(b.eye, 0.0f);
	}
	out->frustum.ws_origin0 = v4(b.eye, 1.0f);
	out->frustum.ws_dir0 = v4(dirs[0], 0.0f);
	out->frustum.ws_dirx = v4((dirs[3] - dirs[0]) * (1.0f / float(camera->viewport_width)), 0.0f);
	out->frustum.ws_diry = v4((dirs[1] - dirs[0]) * (1.0f / float(camera->viewport_height)), 0.0f);

	out->lighting = *lighting;
	out->background_color = LucidVec4{background_rgba[0], background_rgba[1], background_rgba[2],
									  background_rgba[3]};
	out->enable_backface_culling = backface_culling ? 1u : 0u;
	out->num_instances = num_instances;
	out->instance_packet_size = lucid_host_packet_size(num_instances, max_dispatches);
}

static uint32_t toU8(float v) {
	// IColor(FColor): clamp(c * 255, 0, 255) then truncation to u8
	float s = std::min(std::max(v * 255.0f, 0.0f), 255.0f);
	return (uint32_t)s;
}

int lucid_host_build_instances(const LucidDrawCall *dcs, int num_dcs, const LucidMaterial *materials,
							   int num_materials, LucidInstanceData *out_instances,
							   uint32_t *out_colors, float *out_uv_rects, int capacity) {
	if(!dcs || !materials || !out_instances || !out_colors || !out_uv_rects)
		return -1;
	capacity = std::min(capacity, (int)LUCID_MAX_INSTANCES);
	int n = 0;
	for(int d = 0; d < num_dcs; d++) {
		const LucidDrawCall &dc = dcs[d];
		if(dc.num_quads <= 0)
			continue;
		if(dc.material_id < 0 || dc.material_id >= num_materials)
			return -1;
		const LucidMaterial &mat = materials[dc.material_id];
		uint32_t color = toU8(mat.diffuse[0]) | (toU8(mat.diffuse[1]) << 8) |
						 (toU8(mat.diffuse[2]) << 16) | (toU8(mat.opacity) << 24);
		uint32_t opts = dc.opts;
		if(color != 0xffffffffu)
			opts |= LUCID_INST_HAS_COLOR;
		for(int i = 0; i < dc.num_quads; i += LUCID_MAX_INSTANCE_QUADS) {
			if(n >= capacity)
				return n; // the reference truncates the instance list the same way
			LucidInstanceData &inst = out_instances[n];
			inst.index_offset = dc.quad_offset * 4 + i * 4;
			inst.vertex_offset = 0;
			inst.num_quads = std::min((int)LUCID_MAX_INSTANCE_QUADS, dc.num_quads - i);
			inst.flags = opts & 0xffffu;
			out_colors[n] = color;
			memcpy(out_uv_rects + (size_t)n * 4, mat.uv_rect, 4 * sizeof(float));
			n++;
		}
	}
	return n;
}
}
