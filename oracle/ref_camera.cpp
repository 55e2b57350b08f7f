// ref_camera.cpp -- prints the camera-derived part of LucidConfig computed by the REFERENCE's own
// math library (libfwk, compiled from /root/reference where it lies; see build_ref.sh).
//
// TEST INFRASTRUCTURE ONLY.  Used by tests/test_host.py to pin lucid_host_make_config() and to
// generate tests/golden/ref_camera.json.  The sequence below follows FrustumInfo::FrustumInfo
// (src/shading.cpp:46-62) and LucidRenderer::setupInputData (src/lucid_renderer.cpp:439-441);
// everything it calls -- OrbitingCamera::toCamera, Camera::viewMatrix/projectionMatrix/matrix,
// Frustum(proj).cornerRays(), inverseOrZero, mulPoint, mulNormal -- is the reference's code.
//
// usage: ref_camera orbit cx cy cz dist rot_h rot_v fov_deg znear zfar width height
//        ref_camera lookat px py pz tx ty tz ux uy uz fov_deg znear zfar width height
//        ref_camera color r g b opacity   -> u32(IColor(FColor(diffuse, opacity))), the material colour of
//                                            LucidRenderer::uploadInstances (src/lucid_renderer.cpp:364-365)
#include <fwk/gfx/camera.h>
#include <fwk/gfx/color.h>
#include <fwk/gfx/orbiting_camera.h>
#include <fwk/math/frustum.h>
#include <fwk/math/matrix4.h>
#include <fwk/math/ray.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace fwk;

static void print3(const char *name, const float3 &v) { printf("%s %.9g %.9g %.9g\n", name, v.x, v.y, v.z); }

int main(int argc, char **argv) {
	if(argc < 2)
		return 1;
	CameraParams params;
	Camera cam;
	int a = 2;
	auto f = [&]() { return (float)atof(argv[a++]); };
	if(!strcmp(argv[1], "color") && argc == 6) {
		float r = f(), g = f(), b = f(), opacity = f();
		printf("color %u\n", u32(IColor(FColor(float3(r, g, b), opacity))));
		return 0;
	}
	if(!strcmp(argv[1], "orbit") && argc == 13) {
		float cx = f(), cy = f(), cz = f(); // sequenced reads (argument evaluation order is unspecified)
		float3 center(cx, cy, cz);
		float dist = f();
		float rot_h = f();
		float rot_v = f();
		params.fov_in_radians = degToRad(f());
		float zn = f();
		float zf = f();
		params.depth = {zn, zf};
		int w = atoi(argv[a++]), h = atoi(argv[a++]);
		params.viewport = IRect(0, 0, w, h);
		cam = OrbitingCamera(center, dist, rot_h, rot_v).toCamera(params);
	} else if(!strcmp(argv[1], "lookat") && argc == 16) {
		float v[9];
		for(int i = 0; i < 9; i++)
			v[i] = f();
		float3 pos(v[0], v[1], v[2]), target(v[3], v[4], v[5]), up(v[6], v[7], v[8]);
		params.fov_in_radians = degToRad(f());
		float zn = f();
		float zf = f();
		params.depth = {zn, zf};
		int w = atoi(argv[a++]), h = atoi(argv[a++]);
		params.viewport = IRect(0, 0, w, h);
		cam = Camera(pos, target, up, params);
	} else {
		return 2;
	}

	auto iview = inverseOrZero(cam.viewMatrix());
	auto rays = Frustum(cam.projectionMatrix()).cornerRays();
	float3 origins[4], dirs[4];
	for(int i = 0; i < 4; i++) {
		origins[i] = mulPoint(iview, rays[i].origin());
		dirs[i] = mulNormal(iview, rays[i].dir());
	}
	float3 dirx = (dirs[3] - dirs[0]) * (1.0f / cam.params().viewport.width());
	float3 diry = (dirs[1] - dirs[0]) * (1.0f / cam.params().viewport.height());
	for(int i = 0; i < 4; i++) {
		char name[32];
		snprintf(name, sizeof(name), "origin%d", i);
		print3(name, origins[i]);
		snprintf(name, sizeof(name), "dir%d", i);
		print3(name, dirs[i]);
	}
	print3("dirx", dirx);
	print3("diry", diry);
	Matrix4 vp = cam.matrix();
	printf("view_proj");
	for(int c = 0; c < 4; c++)
		for(int r = 0; r < 4; r++)
			printf(" %.9g", vp[c][r]);
	printf("\n");
	print3("pos", cam.pos());
	return 0;
}
