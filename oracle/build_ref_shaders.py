#!/usr/bin/env python
"""Compiles functions of the reference's compute shaders -- from the GLSL text where it lies under
/root/reference/data/shaders -- into oracle/_ref/libref_shaders.so, so the CPU checker can be pinned
against the reference's own source for the arithmetic that decides coverage (quad culling and bin
AABBs, triangle plane / barycentric / scanline equations, scanline parameters, the 4-row span step,
half-block centroids and block depth keys).

    python oracle/build_ref_shaders.py [--reference /root/reference] [--keep-source]

Nothing of the reference is copied into the repository: the extracted text only exists in the
git-ignored oracle/_ref/ (generated C++ next to the .so).  The translation is purely textual and is
listed in oracle/glsl_shim.h; the wrappers at the end of the generated file are ours.
Test infrastructure (tests/golden/make_ref_shader_golden.py turns its outputs into committed vectors).
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# (file under data/shaders, function names in the order they are emitted)
FUNCTIONS = [
    ("shared/funcs.glsl", ["encodeNormalUint", "encodeAABB28", "decodeAABB28", "decodeNormalUint", "decodeRGBA8", "encodeRGBA8",
                           "linearToSRGB", "SRGBToLinear", "finalShading"]),
    ("quad_setup.glsl", ["vertexLoad", "vertexClipMask", "computeClippedAABB", "computeAABB", "processInputQuad", "storeQuad", "storeTri", "addVisibleTri"]),
    ("shared/scanline.glsl", ["loadScanlineParamsRow", "loadScanlineParamsBin"]),
    ("bin_counter.glsl", ["scanlineStep", "countSmallQuadBins", "loadScanlineParamsBin", "countLargeTriBins"]),
    ("bin_dispatcher.glsl", ["dispatchQuad", "dispatchLargeTriSimple"]),
    ("shared/raster.glsl", ["rasterBinStep", "rasterHalfBlockCentroid", "rasterHalfBlockBits", "rasterBlockDepth"]),
    ("shared/shading.glsl", ["getTriangleParams", "getTriangleVertexColors", "getTriangleVertexNormals",
                             "getTriangleVertexTexCoords", "shadeSample"]),
]
# #define lines taken over from the reference (name -> file)
DEFINES = {
    "shared/definitions.glsl": ["REJECTION_TYPE_COUNT", "REJECTION_TYPE_OTHER", "REJECTION_TYPE_BACKFACE",
                                "REJECTION_TYPE_FRUSTUM", "REJECTION_TYPE_BETWEEN_SAMPLES", "INST_HAS_VERTEX_NORMALS",
                                "INST_HAS_VERTEX_COLORS", "INST_TEX_OPAQUE", "INST_HAS_UV_RECT", "INST_HAS_ALBEDO_TEXTURE",
                                "INST_HAS_COLOR",
                                "STORAGE_TRI_BARY_OFFSET", "STORAGE_TRI_SCAN_OFFSET", "STORAGE_TRI_DEPTH_OFFSET",
                                "STORAGE_QUAD_COLOR_OFFSET", "STORAGE_QUAD_NORMAL_OFFSET", "STORAGE_QUAD_TEXTURE_OFFSET"],
    "shared/funcs.glsl": ["SATURATE"],
    "quad_setup.glsl": ["MAX_INSTANCE_QUADS", "LSIZE"],
    "shared/raster.glsl": ["BIN_MASK", "HBLOCK_WIDTH", "HBLOCK_WIDTH_SHIFT", "HBLOCK_COLS", "HBLOCK_COLS_SHIFT",
                           "HBLOCK_COLS_MASK"],
}


def extract_function(text: str, name: str) -> str:
    m = re.search(r"^[A-Za-z_]\w*(?:\s+\w+)*\s+" + re.escape(name) + r"\s*\(", text, re.M)
    if not m:
        raise SystemExit(f"function {name} not found")
    i = text.index("{", m.end())
    depth, j = 0, i
    while True:
        c = text[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return text[m.start():j + 1]


def extract_struct(text: str, name: str) -> str:
    m = re.search(r"^struct\s+" + re.escape(name) + r"\s*\{", text, re.M)
    j = text.index("};", m.end())
    return text[m.start():j + 2]


def extract_define(text: str, name: str) -> str:
    m = re.search(r"^#define\s+" + re.escape(name) + r"\b.*$", text, re.M)
    if not m:
        raise SystemExit(f"#define {name} not found")
    return m.group(0)


def translate(code: str) -> str:
    """GLSL -> C++ on the text (see glsl_shim.h for the list)."""
    # comments may contain anything: drop them first
    code = re.sub(r"//[^\n]*", "", code)
    # float literals are 32-bit
    code = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)(?![\w.])", r"\1f", code)
    # out / inout parameters are references
    code = re.sub(r"\b(?:in\s+out|inout|out)\s+(\w+)\s+(\w+)", r"\1 &\2", code)
    # the one swizzle that is written to
    code = re.sub(r"(\w+(?:\[\d+\])?)\.xyz\s*\*=\s*([^;]+);", r"\1.mul_xyz(\2);", code)
    # swizzles that are read
    code = re.sub(r"\.(xyz|xzw|xy|zw)\b(?!\s*\()", r".\1()", code)
    # colour names of components; the one colour swizzle that is assigned to
    code = re.sub(r"(\w+)\.rgb\s*=(?!=)\s*([^;]+);", r"\1.set_xyz(\2);", code)
    code = re.sub(r"\.rgb\b", ".xyz()", code)
    code = re.sub(r"\.([rgba])\b", lambda m: "." + "xyzw"["rgba".index(m.group(1))], code)
    # conversions of floats to integers saturate
    code = re.sub(r"(?<![\w.])int\(", "glsl_int(", code)
    code = re.sub(r"(?<![\w.])uint\(", "glsl_uint(", code)
    # uvec4(uvec3, <int-typed>) needs the explicit component type
    code = code.replace("uvec4(floatBitsToUint(depth_eq), instance_flags_id)", "make_uvec4(floatBitsToUint(depth_eq), instance_flags_id)")
    code = code.replace("uvec4(floatBitsToUint(scan), y_aabb)", "make_uvec4(floatBitsToUint(scan), y_aabb)")
    code = code.replace("uvec4(floatBitsToUint(scan_step), x_signs | y_signs)", "make_uvec4(floatBitsToUint(scan_step), x_signs | y_signs)")
    # ivec4(<uint scalar>) broadcast
    code = re.sub(r"ivec4\((tri_mins|tri_maxs)\)", r"ivec4(glsl_int(\1))", code)
    # GLSL array initialisers with a trailing comma and `{...}` are fine in C++; bool vector ctor as well
    return code


WRAPPER = r'''
// ---- C ABI (ours): one emulated invocation per call ----------------------------------------------
extern "C" {

// LucidConfig as laid out in include/lucid_abi.h (352 bytes): frustum 12 x vec4, view_proj 4 x vec4,
// lighting 4 x vec4, background vec4, enable_backface_culling at word 84
static void loadConfig(const float *cfg352) {
	// structures.glsl Frustum: ws_origins[4], ws_dirs[4], ws_origin0, ws_dir0, ws_dirx, ws_diry (vec4 each)
	const float *f = cfg352;
	u_config.frustum.ws_origin0 = vec4(f[32], f[33], f[34], f[35]);
	u_config.frustum.ws_dir0 = vec4(f[36], f[37], f[38], f[39]);
	u_config.frustum.ws_dirx = vec4(f[40], f[41], f[42], f[43]);
	u_config.frustum.ws_diry = vec4(f[44], f[45], f[46], f[47]);
	for(int c = 0; c < 4; c++)
		u_config.view_proj_matrix.col[c] = vec4(f[48 + c * 4], f[49 + c * 4], f[50 + c * 4], f[51 + c * 4]);
	// structures.glsl Lighting: ambient_color, sun_color, sun_dir, sun_power, ambient_power (64 bytes)
	u_config.lighting.ambient_color = vec4(f[64], f[65], f[66], f[67]);
	u_config.lighting.sun_color = vec4(f[68], f[69], f[70], f[71]);
	u_config.lighting.sun_dir = vec4(f[72], f[73], f[74], f[75]);
	u_config.lighting.sun_power = f[76], u_config.lighting.ambient_power = f[77];
	u_config.background_color = vec4(f[80], f[81], f[82], f[83]);
	u_config.enable_backface_culling = ((const int *)cfg352)[84];
}

// quad_setup.glsl: processInputQuad for one quad given by four positions.
// out[0] status (-1 visible, else the rejection type), out[1] size type, out[2] enc_aabb,
// out[3], out[4] the two triangles' y ranges
void ref_process_quad(const float *cfg352, int width, int height, const float *pos12, const uint32_t *idx4,
					  uint32_t *out) {
	loadConfig(cfg352);
	VIEWPORT_SIZE_X = width, VIEWPORT_SIZE_Y = height;
	for(int i = 0; i < 12; i++)
		g_verts[i] = pos12[i];
	for(uint &r : s_rejected_quads)
		r = 0;
	s_num_visible[0] = s_num_visible[1] = 0;
	processInputQuad(0, idx4[0], idx4[1], idx4[2], idx4[3], 0);
	out[0] = 0xffffffffu, out[1] = 0, out[2] = out[3] = out[4] = 0;
	for(int t = 0; t < REJECTION_TYPE_COUNT; t++)
		if(s_rejected_quads[t])
			out[0] = (uint32_t)t;
	if(out[0] != 0xffffffffu)
		return;
	const int large = s_num_visible[1] != 0;
	const int slot = large ? LSIZE - 1 : 0;
	out[1] = (uint32_t)large, out[2] = s_quad_aabbs[slot];
	out[3] = s_tri_y_aabbs[slot].x, out[4] = s_tri_y_aabbs[slot].y;
}

// quad_setup.glsl: storeTri for one camera-relative triangle; s_ray_dir0 as main() computes it
// (dir0 + (dirx + diry) * 0.5).  out: bary 2 x uvec4, scan 2 x uvec4, depth uvec4, normal (21 words)
void ref_store_tri(const float *cfg352, const float *tri9, uint32_t flags_id, uint32_t y_aabb, uint32_t *out) {
	loadConfig(cfg352);
	s_ray_dir0 = u_config.frustum.ws_dir0.xyz() + (u_config.frustum.ws_dirx.xyz() + u_config.frustum.ws_diry.xyz()) * 0.5f;
	vec3 t0(tri9[0], tri9[1], tri9[2]), t1(tri9[3], tri9[4], tri9[5]), t2(tri9[6], tri9[7], tri9[8]);
	g_normals_storage[0] = 0;
	storeTri(0, flags_id, t0, t1, t2, y_aabb);
	const uvec4 *src[5] = {&g_uvec4_storage[STORAGE_TRI_BARY_OFFSET], &g_uvec4_storage[STORAGE_TRI_BARY_OFFSET + 1],
						   &g_uvec4_storage[STORAGE_TRI_SCAN_OFFSET], &g_uvec4_storage[STORAGE_TRI_SCAN_OFFSET + 1],
						   &g_uvec4_storage[STORAGE_TRI_DEPTH_OFFSET]};
	for(int i = 0; i < 5; i++)
		out[i * 4 + 0] = src[i]->x, out[i * 4 + 1] = src[i]->y, out[i * 4 + 2] = src[i]->z, out[i * 4 + 3] = src[i]->w;
	out[20] = g_normals_storage[0];
}

// scanline.glsl + raster.glsl: loadScanlineParamsRow at `start`, then `steps` calls of rasterBinStep.
// out: 3 words (min_bits, max_bits, bx_mask) per step
void ref_raster_rows(const uint32_t *scan8, float start_x, float start_y, int steps, uint32_t *out) {
	uvec4 v0(scan8[0], scan8[1], scan8[2], scan8[3]), v1(scan8[4], scan8[5], scan8[6], scan8[7]);
	ScanlineParams p = loadScanlineParamsRow(v0, v1, vec2(start_x, start_y));
	for(int s = 0; s < steps; s++) {
		uvec3 r = rasterBinStep(p);
		out[s * 3 + 0] = r.x, out[s * 3 + 1] = r.y, out[s * 3 + 2] = r.z;
	}
}

// scanline.glsl + bin_counter.glsl: loadScanlineParamsBin, then one scanlineStep per bin row.
// out[0] min_by, out[1] max_by, then (bmin, bmax) per row
void ref_bin_rows(const uint32_t *scan8, int32_t *out) {
	uvec4 v0(scan8[0], scan8[1], scan8[2], scan8[3]), v1(scan8[4], scan8[5], scan8[6], scan8[7]);
	int min_by, max_by;
	ScanlineParams p = loadScanlineParamsBin(v0, v1, min_by, max_by);
	out[0] = min_by, out[1] = max_by;
	for(int by = min_by, i = 0; by <= max_by && i < 128; by++, i++) {
		int bmin, bmax;
		scanlineStep(p, bmin, bmax);
		out[2 + i * 2] = bmin, out[3 + i * 2] = bmax;
	}
}

// raster.glsl: centroid sums, fragment count, packed (xmin, count) bits of one half-block column and the
// block depth key of a depth plane at a centroid.  out: cx bits, cy bits, num_frags, bits, depth
void ref_half_block(uint32_t mins, uint32_t maxs, int startx, const float *depth_eq3, float cpx, float cpy,
					float depth_range, uint32_t *out) {
	uint nf = 0, nf2 = 0;
	vec2 c = rasterHalfBlockCentroid(mins, maxs, startx, nf);
	out[0] = floatBitsToUint(c.x), out[1] = floatBitsToUint(c.y), out[2] = nf;
	out[3] = rasterHalfBlockBits(mins, maxs, startx, nf2);
	g_uvec4_storage[STORAGE_TRI_DEPTH_OFFSET] = uvec4(floatBitsToUint(depth_eq3[0]), floatBitsToUint(depth_eq3[1]),
													  floatBitsToUint(depth_eq3[2]), 0u);
	out[4] = rasterBlockDepth(vec2(cpx, cpy), 0, depth_range);
}

// shading.glsl: one pixel's reduction over n (colour, depth bits) samples in stream order, fed in rounds
// of 32 like shadeAndReduceSamples does (raster.glsl:358-396); out: r, g, b, a as binary32 bits
void ref_reduce_pixel(const float *cfg352, const uint32_t *samples, int n, uint32_t *out) {
	loadConfig(cfg352);
	ReductionContext ctx;
	initReduceSamples(ctx);
	for(int i = 0; i < n; i += 32) {
		const int m = n - i < 32 ? n - i : 32;
		for(int j = 0; j < m; j++)
			g_lane_samples[j] = uvec2(samples[2 * (i + j)], samples[2 * (i + j) + 1]);
		reduceSample(ctx, ctx.out_color, g_lane_samples[0], m == 32 ? 0xffffffffu : (1u << m) - 1u);
	}
	vec4 r = finishReduceSamples(ctx);
	out[0] = floatBitsToUint(r.x), out[1] = floatBitsToUint(r.y), out[2] = floatBitsToUint(r.z), out[3] = floatBitsToUint(r.w);
}
// shading.glsl: shadeSample of triangle `second` of one quad.  rec21: the triangle's record (bary 2x4, scan 2x4,
// depth 4, normal); attrs16: the quad's colours, normals, uv0, uv1; the texture fetch returns tex_preset.
// out: colour, depth bits, 6 textureGrad argument bits (coord, dx, dy), texture slot, fetch-happened flag
void ref_shade_sample(const float *cfg352, const uint32_t *rec21, const uint32_t *attrs16, uint32_t inst_color,
					  const float *uv_rect4, const float *tex_preset4, int px, int py, int second, uint32_t *out) {
	loadConfig(cfg352);
	const uint tri = (uint)second;
	for(int i = 0; i < 2; i++)
		g_uvec4_storage[STORAGE_TRI_BARY_OFFSET + tri * 2 + i] = uvec4(rec21[i * 4], rec21[i * 4 + 1], rec21[i * 4 + 2], rec21[i * 4 + 3]);
	g_uvec4_storage[STORAGE_TRI_DEPTH_OFFSET + tri] = uvec4(rec21[16], rec21[17], rec21[18], rec21[19]);
	g_normals_storage[tri] = rec21[20];
	g_uvec4_storage[STORAGE_QUAD_COLOR_OFFSET] = uvec4(attrs16[0], attrs16[1], attrs16[2], attrs16[3]);
	g_uvec4_storage[STORAGE_QUAD_NORMAL_OFFSET] = uvec4(attrs16[4], attrs16[5], attrs16[6], attrs16[7]);
	g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET] = uvec4(attrs16[8], attrs16[9], attrs16[10], attrs16[11]);
	g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET + 1] = uvec4(attrs16[12], attrs16[13], attrs16[14], attrs16[15]);
	const uint instance_id = rec21[19] >> 16;
	g_instance_colors[instance_id] = inst_color;
	g_instance_uv_rects[instance_id] = vec4(uv_rect4[0], uv_rect4[1], uv_rect4[2], uv_rect4[3]);
	g_tex_preset = vec4(tex_preset4[0], tex_preset4[1], tex_preset4[2], tex_preset4[3]);
	for(float &a : g_tex_args)
		a = 0.0f;
	float depth = 0.0f;
	out[0] = shadeSample(ivec2(px, py), tri, depth);
	out[1] = floatBitsToUint(depth);
	for(int i = 0; i < 8; i++)
		out[2 + i] = floatBitsToUint(g_tex_args[i]);
}
// quad_setup.glsl: storeQuad for a quad with vertices 0..3; out: colours 4, normals 4, uv0 4, uv1 4
void ref_store_quad(uint32_t flags, const uint32_t *colors4, const uint32_t *normals4, const float *uvs8, uint32_t *out) {
	for(int i = 0; i < 4; i++)
		g_colors[i] = colors4[i], g_normals[i] = normals4[i], g_tex_coords[i] = vec2(uvs8[i * 2], uvs8[i * 2 + 1]);
	const uvec4 zero(0u, 0u, 0u, 0u);
	g_uvec4_storage[STORAGE_QUAD_COLOR_OFFSET] = g_uvec4_storage[STORAGE_QUAD_NORMAL_OFFSET] = zero;
	g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET] = g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET + 1] = zero;
	storeQuad(0, flags, 0, 1, 2, 3);
	const uvec4 *src[4] = {&g_uvec4_storage[STORAGE_QUAD_COLOR_OFFSET], &g_uvec4_storage[STORAGE_QUAD_NORMAL_OFFSET],
						   &g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET], &g_uvec4_storage[STORAGE_QUAD_TEXTURE_OFFSET + 1]};
	for(int i = 0; i < 4; i++)
		out[i * 4 + 0] = src[i]->x, out[i * 4 + 1] = src[i]->y, out[i * 4 + 2] = src[i]->z, out[i * 4 + 3] = src[i]->w;
}
// Scene level: quad setup -> bin counting -> bin dispatch of one frame through the reference's per-invocation
// functions, invocations run one after the other (processInputQuad, addVisibleTri/storeTri, countSmallQuadBins,
// countLargeTriBins, dispatchQuad, dispatchLargeTriSimple).  What the shaders' main() functions add around them --
// slot assignment and the prefix sums of the categoriser -- is done here in the canonical order of the checker:
// visible small quads take slots 0.. in input order, large quads MAX_VISIBLE_QUADS-1.. downwards.
// out_counts: quad counts then triangle counts per bin; out_lists: per-bin quad lists then per-bin triangle lists,
// each bin's segment sorted; out_n: visible small, visible large, list lengths; returns 0, or 1 on overflow.
// quad_flags_id (may be null): per input quad, instance flags | instance id << 16; colors / normals / uvs (may be
// null): the vertex attribute buffers storeQuad repacks.
int ref_bin_scene(const float *cfg352, int width, int height, const float *positions, int num_verts,
				  const uint32_t *indices, int num_quads, const uint32_t *quad_flags_id, const uint32_t *colors,
				  const uint32_t *normals, const float *uvs, int32_t *out_counts, uint32_t *out_lists, uint32_t *out_n) {
	loadConfig(cfg352);
	VIEWPORT_SIZE_X = width, VIEWPORT_SIZE_Y = height;
	BIN_COUNT_X = (width + BIN_SIZE - 1) / BIN_SIZE;
	const int bcy = (height + BIN_SIZE - 1) / BIN_SIZE, bc = BIN_COUNT_X * bcy;
	if(num_verts > 524288 || bc > 128 * 128)
		return 1;
	for(int i = 0; i < num_verts * 3; i++)
		g_verts[i] = positions[i];
	for(int i = 0; i < num_verts; i++) {
		g_colors[i] = colors ? colors[i] : 0u, g_normals[i] = normals ? normals[i] : 0u;
		g_tex_coords[i] = uvs ? vec2(uvs[i * 2], uvs[i * 2 + 1]) : vec2(0.0f, 0.0f);
	}
	s_ray_dir0 = u_config.frustum.ws_dir0.xyz() + (u_config.frustum.ws_dirx.xyz() + u_config.frustum.ws_diry.xyz()) * 0.5f;
	int n_small = 0, n_large = 0;
	for(int q = 0; q < num_quads; q++) {
		for(uint &r : s_rejected_quads)
			r = 0;
		s_num_visible[0] = s_num_visible[1] = 0;
		processInputQuad((uint)q, indices[q * 4], indices[q * 4 + 1], indices[q * 4 + 2], indices[q * 4 + 3], 0);
		if(s_num_visible[0] + s_num_visible[1] == 0)
			continue;
		const bool large = s_num_visible[1] != 0;
		const int src = large ? LSIZE - 1 : 0;
		if(n_small + n_large + 1 >= MAX_VISIBLE_QUADS)
			return 1;
		const int slot = large ? (MAX_VISIBLE_QUADS - 1) - n_large++ : n_small++;
		g_quad_aabbs[slot] = s_quad_aabbs[src];
		const uint flags_id = quad_flags_id ? quad_flags_id[q] : 0u;
		storeQuad((uint)slot, flags_id & 0xffffu, s_quad_indices[src].x, s_quad_indices[src].y, s_quad_indices[src].z,
				  s_quad_indices[src].w); // addVisibleQuad, quad_setup.glsl:342-353
		addVisibleTri(slot, src, flags_id, 0);
		addVisibleTri(slot, src, flags_id, 1);
	}
	out_n[0] = (uint32_t)n_small, out_n[1] = (uint32_t)n_large;
	// counting
	for(int b = 0; b < bc; b++)
		s_bins[b] = 0;
	for(int q = 0; q < n_small; q++)
		countSmallQuadBins((uint)q);
	for(int b = 0; b < bc; b++)
		out_counts[b] = s_bins[b], s_bins[b] = 0;
	for(int k = 0; k < n_large; k++)
		for(int second = 0; second < 2; second++)
			countLargeTriBins((MAX_VISIBLE_QUADS - 1) - k, second);
	for(int by = 0; by < bcy; by++) // accumulateLargeTriCountsAcrossRows, its sequential form
		for(int bx = 0, accum = 0; bx < BIN_COUNT_X; bx++) {
			accum += s_bins[bx + by * BIN_COUNT_X];
			out_counts[bc + bx + by * BIN_COUNT_X] = accum;
		}
	// dispatch: exclusive prefix sums as the categoriser leaves them in the *_OFFSETS_TEMP arrays
	int total_q = 0, total_t = 0;
	for(int b = 0; b < bc; b++)
		total_q += out_counts[b], total_t += out_counts[bc + b];
	if(total_q > MAX_VISIBLE_QUADS * 8 || total_t > MAX_VISIBLE_QUADS * 64)
		return 1;
	for(int b = 0, off = 0; b < bc; b++)
		s_bins[b] = off, off += out_counts[b];
	for(int q = 0; q < n_small; q++)
		dispatchQuad(q);
	for(int b = 0, off = 0; b < bc; b++)
		s_bins[b] = off, off += out_counts[bc + b];
	for(int k = 0; k < n_large; k++)
		for(int second = 0; second < 2; second++)
			dispatchLargeTriSimple(k, second, n_large);
	out_n[2] = (uint32_t)total_q, out_n[3] = (uint32_t)total_t;
	for(int i = 0; i < total_q; i++)
		out_lists[i] = g_bin_quads[i];
	for(int i = 0; i < total_t; i++)
		out_lists[total_q + i] = g_bin_tris[i];
	return 0;
}
// Scene level, raster coverage: after ref_bin_scene has run on the same inputs (its buffers are kept), every
// triangle of every bin's lists is walked the way generateRowTris does -- raster_low.glsl:39-64 for LOW bins
// (y range clamped to the bin, loadScanlineParamsRow at the first 8-row block row, two rasterBinStep per block
// row), raster_high.glsl:54-90 for HIGH bins (first 4-row half-block row, one step per row) -- and the pixels
// of rasterHalfBlockBits are counted.  The two walks accumulate the scanline state from different rows, so a
// span can differ in its last pixel between them: which walk a bin gets is part of the result.  A bin is LOW
// with fewer than 1024 triangles (bin_categorizer.glsl:69-79) unless one of its 8x8 blocks collects more than
// 256 (raster_low.glsl:101-105: promoted to HIGH).  out: fragments per pixel; returns total fragments.
static int walkBin(int b, bool high, int q_off, int t_off, int n_q, int n_t, int width, int height, uint32_t *out,
				   uint64_t *total) {
	const ivec2 bin_pos((b % BIN_COUNT_X) * BIN_SIZE, (b / BIN_COUNT_X) * BIN_SIZE);
	const int shift = high ? 2 : 3, steps = high ? 1 : 2;
	int block_tris[16] = {0};
	for(int i = 0; i < n_q * 2 + n_t; i++) {
		uint tri_idx;
		if(i < n_q * 2) {
			const uint w = g_bin_quads[q_off + (i >> 1)];
			if((w >> (30 + (i & 1))) & 1)
				continue;
			tri_idx = (w & 0x0fffffffu) * 2 + (i & 1);
		} else {
			tri_idx = g_bin_tris[t_off + (i - n_q * 2)];
		}
		const uint scan_offset = STORAGE_TRI_SCAN_OFFSET + tri_idx * 2;
		const uvec4 val0 = g_uvec4_storage[scan_offset + 0], val1 = g_uvec4_storage[scan_offset + 1];
		const int min_by = clamp(glsl_int(val0.w & 0xffff) - bin_pos.y, 0, BIN_MASK) >> shift;
		const int max_by = clamp(glsl_int(val0.w >> 16) - bin_pos.y, 0, BIN_MASK) >> shift;
		ScanlineParams scan = loadScanlineParamsRow(val0, val1, vec2(float(bin_pos.x), float(bin_pos.y + (min_by << shift))));
		for(int by = min_by; by <= max_by; by++) {
			uint block_mask = 0;
			for(int s = 0; s < steps; s++) {
				const uvec3 bits = rasterBinStep(scan);
				block_mask |= bits.z;
				if(!out)
					continue;
				for(int hbx = 0; hbx < HBLOCK_COLS; hbx++) {
					if(!((bits.z >> hbx) & 1))
						continue;
					uint nf = 0;
					const uint packed = rasterHalfBlockBits(bits.x, bits.y, hbx * 8, nf);
					*total += nf;
					for(int r = 0; r < 4; r++) {
						const int xmin = (packed >> (7 * r)) & 7, cnt = (packed >> (7 * r + 3)) & 15;
						const int gy = bin_pos.y + ((by * steps + s) << 2) + r;
						for(int x = xmin; x < xmin + cnt; x++) {
							const int gx = bin_pos.x + hbx * 8 + x;
							if(gx < width && gy < height)
								out[gy * width + gx]++;
						}
					}
				}
			}
			if(!high)
				for(int bx = 0; bx < 4; bx++)
					if((block_mask >> bx) & 1)
						block_tris[by * 4 + bx]++;
		}
	}
	int most = 0;
	for(int c : block_tris)
		most = c > most ? c : most;
	return most;
}
uint64_t ref_frag_counts(int width, int height, const int32_t *counts, uint32_t *out, uint8_t *out_high) {
	const int bcy = (height + BIN_SIZE - 1) / BIN_SIZE, bc = BIN_COUNT_X * bcy;
	for(int i = 0; i < width * height; i++)
		out[i] = 0;
	uint64_t total = 0;
	int q_off = 0, t_off = 0;
	for(int b = 0; b < bc; b++) {
		const int n_q = counts[b], n_t = counts[bc + b];
		bool high = n_q * 2 + n_t >= 1024;
		if(!high && walkBin(b, false, q_off, t_off, n_q, n_t, width, height, nullptr, nullptr) > 256)
			high = true; // promoted
		out_high[b] = high;
		walkBin(b, high, q_off, t_off, n_q, n_t, width, height, out, &total);
		q_off += n_q, t_off += n_t;
	}
	return total;
}
// Scene level, image: after ref_bin_scene (buffers kept).  The flow of raster_low / raster_high restated around
// the reference's functions: per bin (level as in ref_frag_counts), per 8x8 block (LOW) or 8x4 half-block (HIGH),
// the triangles whose 4-row spans touch the column, in list order; depth key from rasterHalfBlockCentroid (LOW: the
// sum over both halves, raster_low.glsl:119-131) and rasterBlockDepth, 22 / 18 bits; stable sort by key (ties keep
// list order -- the reference's tie order is the arrival order of atomics), not at all for LOW blocks of <= 3
// (raster_low.glsl:144); then per half-block, per entry, per pixel of rasterHalfBlockBits: shadeSample, reduceSample;
// finishReduceSamples; rgba8 store rounds to nearest.  Bins without triangles keep `background8`.
struct BlockEntry {
	uint tri, key;
	uint mins[2], maxs[2];
};
static void shadeList(BlockEntry *list, int n, int halves, int startx, ivec2 block_pos, int width, int height, uint32_t *image) {
	for(int half = 0; half < halves; half++) {
		ReductionContext ctx[32];
		for(auto &c : ctx)
			initReduceSamples(c);
		for(int i = 0; i < n; i++) {
			uint nf = 0;
			const uint packed = rasterHalfBlockBits(list[i].mins[half], list[i].maxs[half], startx, nf);
			for(int r = 0; r < 4; r++) {
				const int xmin = (packed >> (7 * r)) & 7, cnt = (packed >> (7 * r + 3)) & 15;
				for(int x = xmin; x < xmin + cnt; x++) {
					float depth = 0.0f;
					const uint color = shadeSample(ivec2(block_pos.x + x, block_pos.y + half * 4 + r), list[i].tri, depth);
					g_lane_samples[0] = uvec2(color, floatBitsToUint(depth));
					reduceSample(ctx[r * 8 + x], ctx[r * 8 + x].out_color, g_lane_samples[0], 1u);
				}
			}
		}
		for(int p = 0; p < 32; p++) {
			const int gx = block_pos.x + (p & 7), gy = block_pos.y + half * 4 + (p >> 3);
			if(gx >= width || gy >= height)
				continue;
			const vec4 c = finishReduceSamples(ctx[p]);
			image[gy * width + gx] = glsl_uint(c.x * 255.0f + 0.5f) | (glsl_uint(c.y * 255.0f + 0.5f) << 8) |
									 (glsl_uint(c.z * 255.0f + 0.5f) << 16) | 0xff000000u;
		}
	}
}
int ref_shade_image(const float *cfg352, int width, int height, const int32_t *counts, const uint8_t *is_high,
					const uint32_t *inst_colors, const float *inst_uv_rects, int num_instances, uint32_t background8,
					uint32_t *image) {
	loadConfig(cfg352);
	const int bcy = (height + BIN_SIZE - 1) / BIN_SIZE, bc = BIN_COUNT_X * bcy;
	for(int i = 0; i < num_instances && i < 65536; i++) {
		g_instance_colors[i] = inst_colors[i];
		g_instance_uv_rects[i] = vec4(inst_uv_rects[i * 4], inst_uv_rects[i * 4 + 1], inst_uv_rects[i * 4 + 2], inst_uv_rects[i * 4 + 3]);
	}
	for(int i = 0; i < width * height; i++)
		image[i] = background8;
	static BlockEntry lists[16][4096];
	int q_off = 0, t_off = 0;
	for(int b = 0; b < bc; b++) {
		const int n_q = counts[b], n_t = counts[bc + b];
		const bool high = is_high[b] != 0;
		const ivec2 bin_pos((b % BIN_COUNT_X) * BIN_SIZE, (b / BIN_COUNT_X) * BIN_SIZE);
		const int shift = high ? 2 : 3, steps = high ? 1 : 2, n_rows = BIN_SIZE >> shift;
		if(n_q * 2 + n_t == 0)
			continue;
		// triangle list of the bin in the canonical order: quads by slot, then large triangles ascending
		static uint tris[65536];
		int n_tris = 0;
		for(int i = 0; i < n_q * 2; i++) {
			const uint w = g_bin_quads[q_off + (i >> 1)];
			if(!((w >> (30 + (i & 1))) & 1))
				tris[n_tris++] = (w & 0x0fffffffu) * 2 + (i & 1);
		}
		const int first_large = n_tris;
		for(int i = 0; i < n_t; i++)
			tris[n_tris++] = g_bin_tris[t_off + i];
		for(int i = first_large + 1; i < n_tris; i++) // insertion sort of the (short) large part
			for(int j = i; j > first_large && tris[j - 1] > tris[j]; j--) {
				uint t = tris[j];
				tris[j] = tris[j - 1], tris[j - 1] = t;
			}
		for(int row = 0; row < n_rows; row++) {
			int n[4] = {0, 0, 0, 0};
			for(int i = 0; i < n_tris; i++) {
				const uint scan_offset = STORAGE_TRI_SCAN_OFFSET + tris[i] * 2;
				const uvec4 val0 = g_uvec4_storage[scan_offset + 0], val1 = g_uvec4_storage[scan_offset + 1];
				const int min_by = clamp(glsl_int(val0.w & 0xffff) - bin_pos.y, 0, BIN_MASK) >> shift;
				const int max_by = clamp(glsl_int(val0.w >> 16) - bin_pos.y, 0, BIN_MASK) >> shift;
				if(row < min_by || row > max_by)
					continue;
				ScanlineParams scan = loadScanlineParamsRow(val0, val1, vec2(float(bin_pos.x), float(bin_pos.y + (min_by << shift))));
				uvec3 bits[2] = {uvec3(0u, 0u, 0u), uvec3(0u, 0u, 0u)};
				for(int by = min_by; by <= row; by++) // the incremental state, advanced from the triangle's first row
					for(int s = 0; s < steps; s++)
						bits[s] = rasterBinStep(scan);
				const uint bx_mask = bits[0].z | (steps == 2 ? bits[1].z : 0u);
				for(int bx = 0; bx < 4; bx++) {
					if(!((bx_mask >> bx) & 1) || n[bx] >= 4096)
						continue;
					BlockEntry &e = lists[bx][n[bx]++];
					e.tri = tris[i];
					e.mins[0] = bits[0].x, e.maxs[0] = bits[0].y, e.mins[1] = bits[1].x, e.maxs[1] = bits[1].y;
				}
			}
			for(int bx = 0; bx < 4; bx++) {
				const int cnt = n[bx], startx = bx * 8;
				const ivec2 block_pos(bin_pos.x + startx, bin_pos.y + (row << shift));
				for(int i = 0; i < cnt; i++) {
					BlockEntry &e = lists[bx][i];
					uint nf0 = 0, nf1 = 0;
					vec2 cpos = rasterHalfBlockCentroid(e.mins[0], e.maxs[0], startx, nf0);
					if(!high)
						cpos = cpos + rasterHalfBlockCentroid(e.mins[1], e.maxs[1], startx, nf1);
					const vec2 at = cpos * (0.5f / float(nf0 + nf1)) + vec2(float(block_pos.x), float(block_pos.y));
					e.key = rasterBlockDepth(at, e.tri, high ? float(0x7fffe) : float(0x3ffffe));
				}
				if(high || cnt > 3) // stable insertion sort by depth key
					for(int i = 1; i < cnt; i++)
						for(int j = i; j > 0 && lists[bx][j - 1].key > lists[bx][j].key; j--) {
							BlockEntry t = lists[bx][j];
							lists[bx][j] = lists[bx][j - 1], lists[bx][j - 1] = t;
						}
				shadeList(lists[bx], cnt, high ? 1 : 2, startx, block_pos, width, height, image);
			}
		}
		q_off += n_q, t_off += n_t;
	}
	return 0;
}
uint32_t ref_encode_rgba8(const float *rgba) { return encodeRGBA8(vec4(rgba[0], rgba[1], rgba[2], rgba[3])); }

} // extern "C"
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--keep-source", action="store_true", help="leave the generated C++ in oracle/_ref/ for inspection")
    args = ap.parse_args()
    shaders = os.path.join(args.reference, "data", "shaders")
    if not os.path.isdir(shaders):
        raise SystemExit(f"{shaders} not found: the reference tree is only mounted in the build container")
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)

    parts = ['#include "../glsl_shim.h"', "using namespace glsl;", ""]
    parts += ["// ---- specialisation constants (definitions.glsl CONSTANT(...)) as variables",
              "static int VIEWPORT_SIZE_X = 1280, VIEWPORT_SIZE_Y = 720;",
              "static const int BIN_SIZE = 32, BIN_SHIFT = 5, MAX_VISIBLE_QUADS = 32768;",
              "static int BIN_COUNT_X = 40;", ""]
    parts.append("// ---- #define lines of the reference")
    for rel, names in DEFINES.items():
        text = open(os.path.join(shaders, rel), encoding="latin-1").read()
        for n in names:
            parts.append(translate(extract_define(text, n)))
    parts += ["", "// ---- buffers, shared variables and the uniform block the functions refer to (ours)",
              "struct Frustum { vec4 ws_origin0, ws_dir0, ws_dirx, ws_diry; };",
              "struct Lighting { vec4 ambient_color, sun_color, sun_dir; float sun_power, ambient_power; };",
              "struct Config { Frustum frustum; mat4 view_proj_matrix; Lighting lighting; vec4 background_color; int enable_backface_culling; };",
              "static uint g_colors[524288], g_normals[524288];",
              "static vec2 g_tex_coords[524288];",
              "static uint g_instance_colors[65536];",
              "static vec4 g_instance_uv_rects[65536];",
              "// the Vulkan sampler is not part of the source: textureGrad records its arguments and returns a preset colour",
              "static int opaque_texture = 0, transparent_texture = 1;",
              "static vec4 g_tex_preset; static float g_tex_args[8];",
              "static vec4 textureGrad(int which, vec2 c, vec2 dx, vec2 dy) {",
              "	g_tex_args[0] = c.x, g_tex_args[1] = c.y, g_tex_args[2] = dx.x, g_tex_args[3] = dx.y, g_tex_args[4] = dy.x, g_tex_args[5] = dy.y;",
              "	g_tex_args[6] = float(which), g_tex_args[7] = 1.0f;",
              "	return g_tex_preset;",
              "}",
              "static Config u_config;",
              "static float g_verts[3 * 524288];",
              "static uint g_quad_aabbs[MAX_VISIBLE_QUADS];",
              "static uint g_bin_quads[MAX_VISIBLE_QUADS * 8], g_bin_tris[MAX_VISIBLE_QUADS * 64];",
              "static int s_bins[128 * 128];",
              "static uvec4 g_uvec4_storage[MAX_VISIBLE_QUADS * 14];",
              "static uint g_normals_storage[MAX_VISIBLE_QUADS * 2];",
              "static vec3 s_ray_dir0;",
              "static uint s_rejected_quads[REJECTION_TYPE_COUNT];",
              "static uint s_num_visible[2];",
              "static uint s_quad_aabbs[LSIZE];",
              "static uvec2 s_tri_y_aabbs[LSIZE];",
              "static uvec4 s_quad_indices[LSIZE];", ""]
    parts.append("// ---- reference text (translated)")
    scan = open(os.path.join(shaders, "shared/scanline.glsl"), encoding="latin-1").read()
    for rel, names in FUNCTIONS:
        text = open(os.path.join(shaders, rel), encoding="latin-1").read()
        if rel == "shared/scanline.glsl":
            # `min` / `max` members of ScanlineParams shadow the functions only inside GLSL's rules
            parts.append(translate(extract_struct(scan, "ScanlineParams")))
        for n in names:
            parts.append(f"// {rel}: {n}")
            parts.append(translate(extract_function(text, n)))
            parts.append("")
    # shared/shading.glsl: the per-pixel reduction (3-entry insertion window, blending, final colour), from
    # its RC_* defines to the end of finishReduceSamples, for one emulated lane: the subgroup operations
    # become "this lane" / a lookup in the 32 samples of the round
    shading = open(os.path.join(shaders, "shared/shading.glsl"), encoding="latin-1").read()
    a = shading.index("#define RC_COLOR_SIZE 3")
    fin = extract_function(shading, "finishReduceSamples")
    b = shading.index(fin) + len(fin)
    parts += ["// shared/shading.glsl: ReductionContext, swap, initReduceSamples, reduceSample, finishReduceSamples",
              "#define SUBGROUP_SIZE 32", "#define HALFGROUP_SIZE 32",
              "static uvec2 g_lane_samples[32];",
              "#define subgroupAny(x) (x)", "#define subgroupShuffle(v, lane) g_lane_samples[lane]",
              translate(shading[a:b]), ""]
    parts.append(WRAPPER)
    src = os.path.join(out_dir, "ref_shader_funcs.cpp")
    with open(src, "w") as f:
        f.write("\n".join(parts))
    so = os.path.join(out_dir, "libref_shaders.so")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3",
           "-Wno-div-by-zero", "-o", so, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stderr[:6000])
        raise SystemExit("g++ failed on the translated reference functions")
    if not args.keep_source:
        os.remove(src)
    print(so)


if __name__ == "__main__":
    main()
