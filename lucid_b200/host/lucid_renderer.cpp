// C++ facade with the reference's LucidRenderer surface over the C ABI; see
// include/lucid_renderer.hpp for the reference functions each method stands in for.

#include "../../include/lucid_renderer.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lucid_b200 {

namespace {

// "12 345 678"-style grouping is what the reference's formatLarge prints in its stats tables
std::string formatLarge(long long v) {
	char digits[32];
	snprintf(digits, sizeof(digits), "%lld", v < 0 ? -v : v);
	std::string out;
	int n = (int)strlen(digits);
	for(int i = 0; i < n; i++) {
		out += digits[i];
		int left = n - 1 - i;
		if(left > 0 && left % 3 == 0)
			out += ' ';
	}
	return v < 0 ? "-" + out : out;
}
std::string fmt(const char *f, double a, double b = 0.0) {
	char buf[160];
	snprintf(buf, sizeof(buf), f, a, b);
	return buf;
}
std::string percentage(int value, int total) {
	char buf[64];
	snprintf(buf, sizeof(buf), "%d (%.0f %%)", value, total ? value * 100.0 / total : 0.0);
	return buf;
}

// shader timers -> percentages of their sum (lucid_renderer.cpp:582-596)
std::vector<StatsRow> timerRows(const uint32_t *timers, std::initializer_list<const char *> names) {
	std::vector<StatsRow> rows;
	double sum = 0.0;
	for(size_t i = 0; i < names.size(); i++)
		sum += timers[i];
	if(sum == 0.0)
		return rows;
	size_t i = 0;
	for(const char *name : names)
		rows.push_back({name, fmt("%.2f %%", timers[i++] * 100.0 / sum), ""});
	return rows;
}

} // namespace

void Ex::check() const {
	if(code != 0) {
		fprintf(stderr, "lucid_b200: %s (code %d)\n", message.c_str(), code);
		abort();
	}
}

LucidRenderer::~LucidRenderer() {
	if(m_handle)
		lucid_destroy(m_handle);
}

Ex LucidRenderer::error(int code, const char *what) const {
	Ex e;
	e.code = code;
	e.message = std::string(what) + ": " + lucid_last_error(m_handle);
	return e;
}

Ex LucidRenderer::exConstruct(Opts opts, int2 view_size, int device, int max_visible_quads,
							  int2 bin_rows) {
	if(m_handle) {
		lucid_destroy(m_handle);
		m_handle = nullptr;
	}
	LucidCreateInfo ci;
	memset(&ci, 0, sizeof(ci));
	ci.width = view_size.x, ci.height = view_size.y;
	ci.opts = opts.bits;
	ci.max_visible_quads = max_visible_quads;
	ci.device = device;
	ci.bin_row_begin = bin_rows.x, ci.bin_row_end = bin_rows.y;
	int rc = lucid_create(&ci, &m_handle);
	if(rc != LUCID_OK)
		return error(rc, "LucidRenderer::exConstruct");
	m_opts = opts;
	m_size = view_size;
	m_bin_count = lucid_bin_count(m_handle);
	m_max_visible_quads = max_visible_quads > 0 ? max_visible_quads : 4793490;
	m_verts = VertexArray();
	m_quads_ib = nullptr, m_num_quads = 0;
	m_tex[0] = m_tex[1] = nullptr;
	m_last_info.assign(LUCID_INFO_U32_SIZE + (size_t)m_bin_count * LUCID_COUNTS_PER_BIN, 0u);
	memset(&m_info, 0, sizeof(m_info));
	m_last_info_updated = false;
	return {};
}

Ex LucidRenderer::render(const Context &ctx) {
	if(!m_handle) {
		Ex e;
		e.code = LUCID_E_STATE, e.message = "LucidRenderer::render: exConstruct has not succeeded";
		return e;
	}
	// scene buffers are borrowed / uploaded once, like the reference borrows Scene's VBs and IBs
	const VertexArray &v = ctx.verts;
	if(v.positions != m_verts.positions || v.colors != m_verts.colors || v.tex_coords != m_verts.tex_coords ||
	   v.normals != m_verts.normals || v.num_verts != m_verts.num_verts || ctx.quads_ib != m_quads_ib ||
	   ctx.num_quads != m_num_quads) {
		int rc = lucid_set_geometry(m_handle, v.positions, v.num_verts, v.colors, v.tex_coords, v.normals,
									ctx.quads_ib, ctx.num_quads, ctx.memory);
		if(rc != LUCID_OK)
			return error(rc, "LucidRenderer::render (geometry)");
		m_verts = v, m_quads_ib = ctx.quads_ib, m_num_quads = ctx.num_quads;
	}
	const AtlasTexture *tex[2] = {ctx.opaque_tex, ctx.trans_tex};
	for(int slot = 0; slot < 2; slot++) {
		if(tex[slot] == m_tex[slot] || !tex[slot])
			continue;
		int rc = lucid_set_texture(m_handle, slot, tex[slot]->rgba8_mips, tex[slot]->width, tex[slot]->height,
								   tex[slot]->levels);
		if(rc != LUCID_OK)
			return error(rc, "LucidRenderer::render (texture)");
		m_tex[slot] = tex[slot];
	}

	// uploadInstances: LucidApp::drawScene scales material opacity by scene_opacity first
	// (src/lucid_app.cpp:646-650)
	std::vector<LucidDrawCall> dcs(ctx.dcs.size());
	for(size_t i = 0; i < dcs.size(); i++) {
		dcs[i].material_id = ctx.dcs[i].material_id;
		dcs[i].num_quads = ctx.dcs[i].num_quads, dcs[i].quad_offset = ctx.dcs[i].quad_offset;
		dcs[i].opts = ctx.dcs[i].opts;
	}
	std::vector<LucidMaterial> mats(ctx.materials.size());
	for(size_t i = 0; i < mats.size(); i++) {
		memcpy(mats[i].diffuse, ctx.materials[i].diffuse, sizeof(float) * 3);
		mats[i].opacity = ctx.materials[i].opacity * ctx.config.scene_opacity;
		memcpy(mats[i].uv_rect, ctx.materials[i].uv_rect, sizeof(float) * 4);
	}
	size_t cap = 1;
	for(auto &dc : dcs)
		cap += (size_t)(std::max(dc.num_quads, 0) + LUCID_MAX_INSTANCE_QUADS - 1) / LUCID_MAX_INSTANCE_QUADS;
	cap = std::min(cap, (size_t)LUCID_MAX_INSTANCES);
	m_instances.resize(cap), m_instance_colors.resize(cap), m_instance_uv_rects.resize(cap * 4);
	static const LucidMaterial no_material = {{1.0f, 1.0f, 1.0f}, 1.0f, {0.0f, 0.0f, 1.0f, 1.0f}};
	int n = lucid_host_build_instances(dcs.data(), (int)dcs.size(), mats.empty() ? &no_material : mats.data(),
									   (int)mats.size(), m_instances.data(), m_instance_colors.data(),
									   m_instance_uv_rects.data(), (int)cap);
	if(n < 0) {
		Ex e;
		e.code = LUCID_E_INVALID, e.message = "LucidRenderer::render: draw call refers to a missing material";
		return e;
	}
	m_num_instances = n;

	// setupInputData
	LucidConfig config;
	const unsigned char *bg = ctx.config.background_color;
	const float s = 1.0f / 255.0f;
	float background[4] = {bg[0] * s, bg[1] * s, bg[2] * s, bg[3] * s};
	lucid_host_make_config(&ctx.camera, &ctx.lighting, background, ctx.config.backface_culling ? 1 : 0, n,
						   m_max_dispatches, &config);

	m_config = config;
	int rc = lucid_render(m_handle, &config, m_instances.data(), m_instance_colors.data(),
						  m_instance_uv_rects.data(), n, ctx.out_image, ctx.out_pitch_bytes, ctx.out_memory, 0);
	if(rc != LUCID_OK)
		return error(rc, "LucidRenderer::render");
	rc = lucid_read_info(m_handle, m_last_info.data(), m_last_info.size());
	if(rc != LUCID_OK)
		return error(rc, "LucidRenderer::render (info)");
	memcpy(&m_info, m_last_info.data(), sizeof(m_info));
	m_last_info_updated = true;
	return {};
}

Ex SimpleRenderer::exConstruct(LucidRenderer &sample_source, Technique technique) {
	Ex e;
	if(!sample_source.handle()) {
		e.code = LUCID_E_STATE, e.message = "SimpleRenderer::exConstruct: the LucidRenderer has not been constructed";
		return e;
	}
	m_source = &sample_source, m_technique = technique;
	return e;
}

Ex SimpleRenderer::render(const RenderContext &ctx, bool wireframe) {
	Ex e;
	if(!m_source) {
		e.code = LUCID_E_STATE, e.message = "SimpleRenderer::render: exConstruct has not succeeded";
		return e;
	}
	if(wireframe || !ctx.out_image || ctx.out_memory != LUCID_MEM_HOST) {
		e.code = LUCID_E_INVALID;
		e.message = wireframe ? "SimpleRenderer::render: no wireframe mode on this path" :
								"SimpleRenderer::render: the comparator's image goes to host memory (out_image, LUCID_MEM_HOST)";
		return e;
	}
	// the exact frame of the context stays in the LucidRenderer's own image; its samples are reduced a second time
	RenderContext frame = ctx;
	frame.out_image = nullptr, frame.out_pitch_bytes = 0, frame.out_memory = LUCID_MEM_NONE;
	e = m_source->render(frame);
	if(!e)
		return e;
	int rc = lucid_compare_render(m_source->handle(), int(m_technique), &m_source->lastConfig(), ctx.out_image,
								  ctx.out_pitch_bytes, &m_kernel_ms);
	if(rc != LUCID_OK) {
		e.code = rc;
		e.message = std::string("SimpleRenderer::render: ") + lucid_last_error(m_source->handle());
	}
	return e;
}

Ex LucidRenderer::stageTimes(float ms[8]) {
	int rc = lucid_stage_times(m_handle, ms);
	if(rc != LUCID_OK)
		return error(rc, "LucidRenderer::stageTimes");
	return {};
}

int LucidRenderer::verifyInfo() {
	if(!m_last_info_updated)
		return 0;
	m_last_info_updated = false;
	const int *c = reinterpret_cast<const int *>(m_last_info.data() + LUCID_INFO_U32_SIZE);
	const int bc = m_bin_count;
	int total = 0;
	const char *kind[2] = {"quad", "tri"};
	for(int k = 0; k < 2; k++) {
		const int *counts = c + (size_t)bc * (k * 3), *offsets = counts + bc, *temps = offsets + bc;
		int printed[2] = {0, 0};
		for(int i = 0; i < bc; i++) {
			if(i > 0 && offsets[i] != offsets[i - 1] + counts[i - 1]) {
				total++;
				if(printed[0]++ < 32)
					printf("Invalid bin %s offset [%d]: %d != %d (prev_offset:%d + prev_count:%d)\n", kind[k], i,
						   offsets[i], offsets[i - 1] + counts[i - 1], offsets[i - 1], counts[i - 1]);
			}
			if(temps[i] != offsets[i] + counts[i]) {
				total++;
				if(printed[1]++ < 32)
					printf("Invalid temp bin %s offset [%d]: %d != %d (offset:%d + count:%d)\n", kind[k], i, temps[i],
						   offsets[i] + counts[i], offsets[i], counts[i]);
			}
		}
	}
	return total;
}

std::vector<StatsGroup> LucidRenderer::getStats() const {
	std::vector<StatsGroup> out;
	if(m_last_info.empty() || !m_handle)
		return out;
	LucidInfo info = m_info;
	const int bc = m_bin_count;
	const int *counters = reinterpret_cast<const int *>(m_last_info.data() + LUCID_INFO_U32_SIZE);
	const int *quad_counts = counters, *tri_counts = counters + (size_t)bc * 3;
	long long num_bin_quads = 0, num_bin_tris = 0;
	int max_quads = 0, max_tris = 0;
	for(int i = 0; i < bc; i++) {
		num_bin_quads += quad_counts[i], num_bin_tris += tri_counts[i];
		max_quads = std::max(max_quads, quad_counts[i]), max_tris = std::max(max_tris, tri_counts[i]);
	}
	int level_sum = 0;
	for(int l = 0; l < LUCID_BIN_LEVELS_COUNT; l++)
		level_sum += info.bin_level_counts[l];
	int num_promoted = level_sum - bc; // promoted bins are processed twice
	int visible = info.num_visible_quads[0] + info.num_visible_quads[1];
	double input = info.num_input_quads ? info.num_input_quads : 1;
	uint32_t rejected =
		info.num_rejected_quads[0] + info.num_rejected_quads[1] + info.num_rejected_quads[2] + info.num_rejected_quads[3];

	auto setup_timers =
		timerRows(info.setup_timers, {"init & finish", "process input quads", "store tri data", "store quad data"});
	auto dispatcher_timers = timerRows(info.bin_dispatcher_timers, {"count small quads", "count large tris",
																	  "dispatch small quads", "dispatch large tris"});
	auto raster_timers = timerRows(info.raster_timers, {"generate rows", "generate blocks", "unpack samples",
														 "shade and reduce", "finish reduce"});
	if(!setup_timers.empty())
		out.push_back({setup_timers, "quad_setup timers", 130});
	if(!dispatcher_timers.empty())
		out.push_back({dispatcher_timers, "bin_dispatcher timers", 130});
	if(!raster_timers.empty())
		out.push_back({raster_timers, "raster_low & raster_high timers", 130});

	std::vector<StatsRow> level_rows = {
		{"empty bins", percentage(info.bin_level_counts[LUCID_BIN_LEVEL_EMPTY], bc), ""},
		{"micro level bins", percentage(info.bin_level_counts[LUCID_BIN_LEVEL_MICRO], bc), ""},
		{"low level bins", percentage(info.bin_level_counts[LUCID_BIN_LEVEL_LOW], bc), ""},
		{"high level bins", percentage(info.bin_level_counts[LUCID_BIN_LEVEL_HIGH], bc), ""},
		{"promoted bins", percentage(num_promoted, bc), ""},
	};
	uint32_t fragments = info.stats[0], hblocks = info.stats[1], invalid = info.stats[2];
	double pixels = double(m_size.x) * m_size.y;
	std::vector<StatsRow> basic = {
		{"input instances", formatLarge(m_num_instances), ""},
		{"input quads", formatLarge(info.num_input_quads), ""},
		{"visible quads", formatLarge(visible) + fmt(" (%.2f %%)", visible / input * 100.0),
		 formatLarge(info.num_visible_quads[0]) + " small; " + formatLarge(info.num_visible_quads[1]) + " large"},
		{"rejected quads", formatLarge(rejected) + fmt(" (%.2f %%)", rejected / input * 100.0),
		 "backface: " + formatLarge(info.num_rejected_quads[1]) + "\nfrustum: " +
			 formatLarge(info.num_rejected_quads[2]) + "\nbetween-samples: " + formatLarge(info.num_rejected_quads[3])},
		{"bin quads", formatLarge(num_bin_quads), "Total per-bin quads"},
		{"bin tris", formatLarge(num_bin_tris), "Total per-bin tris"},
		{"max small quads / bin", formatLarge(max_quads), ""},
		{"max large tris / bin", formatLarge(max_tris), ""},
		{"half-block-tris", formatLarge(hblocks), ""},
		{"fragments", formatLarge(fragments),
		 fmt("%.3f avg fragments / pixel\n%.3f avg fragments / half-block-tri", fragments / pixels,
			 hblocks ? double(fragments) / hblocks : 0.0)},
	};
	if(m_opts & Opt::visualize_errors)
		basic.push_back({"invalid pixels", formatLarge(invalid),
						 fmt("%.3f %% total pixels invalid", invalid * 100.0 / pixels)});
	int last = -1;
	for(int i = 0; i < 64; i++)
		if(info.temp[i] != 0)
			last = i;
	if(last >= 0) {
		std::string temps;
		for(int i = 0; i <= last; i++)
			temps += (i ? ", " : "") + std::to_string(info.temp[i]);
		basic.push_back({"temps", temps, "[0] quads dropped past max visible quads, [1] bin list overflow"});
	}
	out.push_back({level_rows, "Bins categorized by quad density levels:", 130});
	out.push_back({basic, "", 130});
	out.push_back({{{"max visible quads", formatLarge(m_max_visible_quads), ""},
					{"max_dispatches", formatLarge(m_max_dispatches), ""}},
				   "LucidRenderer limits",
				   130});
	return out;
}

} // namespace lucid_b200
