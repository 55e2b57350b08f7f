// lucid_renderer.hpp -- C++ host facade with the reference's LucidRenderer surface
// (src/lucid_renderer.h:21-96, src/lucid_base.h:27-96) on top of the C ABI (lucid_b200.h).
//
// A caller of the reference constructs the renderer with options and a view size, hands it a
// RenderContext every frame and reads statistics back:
//     exConstruct(device, compiler, opts, view_size)   src/lucid_renderer.cpp:186-317
//     render(const Context &)                          src/lucid_renderer.cpp:319-350
//     getStats() / verifyInfo()                        src/lucid_renderer.cpp:709-836 / 629-707
// The same calls exist here with the Vulkan objects removed: vertex/index spans are plain
// pointers (host or CUDA device memory), the two atlas image views are RGBA8 mip chains, and the
// swap-chain image is a caller-supplied RGBA8 buffer.  Everything else (draw calls, materials,
// RenderConfig, lighting, camera, option flags, statistics groups) keeps the reference's names and
// meaning.  There is no CPU fallback: exConstruct fails when no CUDA device is usable.
#pragma once

#include "lucid_b200.h"
#include "lucid_host.h"

#include <string>
#include <vector>

namespace lucid_b200 {

// DEFINE_ENUM(LucidRenderOpt, ...) src/lucid_renderer.h:10-11; EnumFlags bit i = 1 << i
enum class LucidRenderOpt {
	debug_quad_setup,
	debug_bin_counter,
	debug_bin_dispatcher,
	debug_raster,
	timers,
	additive_blending,
	visualize_errors,
	alpha_threshold
};
struct LucidRenderOpts {
	unsigned bits = 0;
	LucidRenderOpts() = default;
	LucidRenderOpts(LucidRenderOpt o) : bits(1u << int(o)) {}
	LucidRenderOpts operator|(LucidRenderOpts rhs) const {
		LucidRenderOpts out;
		out.bits = bits | rhs.bits;
		return out;
	}
	bool operator&(LucidRenderOpt o) const { return (bits >> int(o)) & 1u; }
};
inline LucidRenderOpts operator|(LucidRenderOpt a, LucidRenderOpt b) {
	return LucidRenderOpts(a) | LucidRenderOpts(b);
}

// DEFINE_ENUM(DrawCallOpt, ...) src/lucid_base.h:44-46; the bits are LUCID_INST_* of lucid_abi.h
enum class DrawCallOpt {
	has_vertex_colors,
	has_vertex_tex_coords,
	has_vertex_normals,
	is_opaque,
	tex_opaque,
	has_uv_rect,
	has_albedo_tex,
	has_normal_tex,
	has_pbr_tex,
	has_inst_color
};
inline unsigned flag(DrawCallOpt o) { return 1u << int(o); }

struct int2 {
	int x = 0, y = 0;
};

// Expected-style result (libfwk Ex<>): falsy on error, check() aborts like the reference's .check()
struct Ex {
	int code = 0;
	std::string message;
	explicit operator bool() const { return code == 0; }
	void check() const;
};

// src/lucid_base.h:27-36 (sampler_setup: the filter is fixed, see DESIGN.md "Texture filter")
struct RenderConfig {
	float scene_opacity = 1.0f;
	unsigned char background_color[4] = {0, 30, 30, 255};
	bool backface_culling = false;
	bool additive_blending = false; // informational: blending mode is a construction option
};

// src/lucid_base.h:48-54
struct SceneDrawCall {
	int material_id = -1;
	int num_quads = 0, quad_offset = 0;
	unsigned opts = 0;
};

// the fields of SceneMaterial that uploadInstances reads (src/lucid_renderer.cpp:368-399)
struct SceneMaterial {
	float diffuse[3] = {1.0f, 1.0f, 1.0f};
	float opacity = 1.0f;
	float uv_rect[4] = {0.0f, 0.0f, 1.0f, 1.0f};
};

// src/lucid_base.h:62-70: positions float3 tightly packed, colors RGBA8, tex coords float2,
// normals 10-10-10.  Pointers are host memory or CUDA device memory (RenderContext::memory).
struct VertexArray {
	const float *positions = nullptr;
	const uint32_t *colors = nullptr;
	const float *tex_coords = nullptr;
	const uint32_t *normals = nullptr;
	int num_verts = 0;
};

struct AtlasTexture {
	const uint8_t *rgba8_mips = nullptr; // level l is max(1,w>>l) x max(1,h>>l), tightly packed
	int width = 0, height = 0, levels = 0;
};

// src/lucid_base.h:72-84
struct RenderContext {
	RenderConfig config;
	VertexArray verts;
	const uint32_t *quads_ib = nullptr; // 4 indices per quad
	int num_quads = 0;
	int memory = LUCID_MEM_HOST; // where verts / quads_ib live
	std::vector<SceneDrawCall> dcs;
	std::vector<SceneMaterial> materials;
	const AtlasTexture *opaque_tex = nullptr, *trans_tex = nullptr;
	LucidLighting lighting;
	LucidCamera camera;
	// the image the reference acquires from the swap chain (src/lucid_renderer.cpp:542-544)
	void *out_image = nullptr;
	size_t out_pitch_bytes = 0;
	int out_memory = LUCID_MEM_NONE;
};

struct StatsRow {
	std::string label, value, tooltip;
};
struct StatsGroup {
	std::vector<StatsRow> rows;
	std::string title;
	int label_width = 100;
};

class LucidRenderer {
  public:
	using Opt = LucidRenderOpt;
	using Opts = LucidRenderOpts;
	using Context = RenderContext;

	// 7-bit bin coordinates allow 4096; the reference declares 2560x2048 and never enforces it
	static constexpr int max_width = 4096, max_height = 4096;
	static constexpr int max_instances = LUCID_MAX_INSTANCES;
	static constexpr int max_instance_quads = LUCID_MAX_INSTANCE_QUADS;

	LucidRenderer() = default;
	~LucidRenderer();
	LucidRenderer(const LucidRenderer &) = delete;
	LucidRenderer &operator=(const LucidRenderer &) = delete;

	// device: CUDA ordinal; max_visible_quads 0 -> the reference's 4793490; bin_rows: owned bin
	// rows [x, y) for the multi-GPU split, {0,0} -> the whole view
	Ex exConstruct(Opts, int2 view_size, int device = 0, int max_visible_quads = 0,
				   int2 bin_rows = {});
	Ex render(const Context &);

	// verifyInfo prints what the reference prints and returns the number of invalid offsets
	int verifyInfo();
	std::vector<StatsGroup> getStats() const;
	const LucidInfo &lastInfo() const { return m_info; }
	const LucidConfig &lastConfig() const { return m_config; } // what setupInputData made of the last context
	Ex stageTimes(float ms[8]);

	Opts opts() const { return m_opts; }
	int binSize() const { return LUCID_BIN_SIZE; }
	int blockSize() const { return LUCID_BLOCK_SIZE; }
	int subgroupSize() const { return 32; }
	int maxVisibleQuads() const { return m_max_visible_quads; }
	int maxSceneQuads() const { return m_max_visible_quads * 5 / 2; }
	lucid_renderer *handle() const { return m_handle; }

  private:
	Ex error(int code, const char *what) const;

	lucid_renderer *m_handle = nullptr;
	Opts m_opts;
	int2 m_size;
	int m_bin_count = 0, m_max_visible_quads = 0, m_max_dispatches = 256;
	int m_num_instances = 0;
	// geometry / textures are (re)registered only when the context's pointers change
	VertexArray m_verts;
	const uint32_t *m_quads_ib = nullptr;
	int m_num_quads = 0;
	const AtlasTexture *m_tex[2] = {nullptr, nullptr};
	std::vector<LucidInstanceData> m_instances;
	std::vector<uint32_t> m_instance_colors;
	std::vector<float> m_instance_uv_rects;
	std::vector<uint32_t> m_last_info;
	LucidInfo m_info;
	LucidConfig m_config;
	bool m_last_info_updated = false;
};

// The comparison renderer of the reference's application (src/simple_renderer.h:19-60, SimpleRenderer::render,
// src/simple_renderer.cpp:134-196): the same RenderContext through the fixed-function pipeline -- an opaque phase with
// depth write, then alpha blending in submission order -- instead of the exact blend.  Here it is a second reduction
// of the samples the LucidRenderer it is bound to produces for the context (lucid_compare_render):
// render() draws the context with that renderer and writes the comparator's image to ctx.out_image (host memory).
// Besides the reference's hardware blending, the two approximate order-independent techniques the reference is
// measured against (docs/readme.md:7-8) can be selected.
class SimpleRenderer {
  public:
	enum class Technique { hw_blend = LUCID_COMPARE_HW_BLEND, wboit = LUCID_COMPARE_WBOIT, mlab4 = LUCID_COMPARE_MLAB4 };

	Ex exConstruct(LucidRenderer &sample_source, Technique = Technique::hw_blend);
	// wireframe (VPolygonMode::line in the reference) is not available: there is no line rasteriser on this path
	Ex render(const RenderContext &ctx, bool wireframe = false);
	float lastKernelMs() const { return m_kernel_ms; } // device time of the comparator's kernels

  private:
	LucidRenderer *m_source = nullptr;
	Technique m_technique = Technique::hw_blend;
	float m_kernel_ms = 0.0f;
};

} // namespace lucid_b200
