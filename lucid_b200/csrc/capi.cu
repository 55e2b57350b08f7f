// capi.cu -- the C ABI declared in include/lucid_b200.h: device memory, streams, per-frame
// uploads, kernel sequencing and read-backs.  There is no CPU fallback: every entry point that
// needs the GPU fails with LUCID_E_CUDA when no device is usable.
#include "../../include/lucid_b200.h"
#include "common.cuh"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace lucid;

namespace lucid {
size_t compareScratchKeys(int num_sms);
void launchCompare(const Params &p, const LucidConfig &cfg, int mode, u32 *order, u64 *scratch, u32 *ticket, u32 *out_image,
				   cudaStream_t stream, int num_sms);
} // namespace lucid

static thread_local std::string g_create_error;

// Programmatic dependent launch per frame.  Overlapping the launch of kernel n+1 with the tail of kernel n pays on
// short frames (1M-triangle 1080p frame: 0.246 against 0.263 ms) and costs on long ones (10M-triangle 4K frame: 2.71
// against 2.51 ms, profiles/r3b_*), so a handle decides from the time its last finished frame took.
// LUCID_PDL=on|off (or LUCID_NO_PDL=1) fixes the choice for measurements.
static thread_local bool g_frame_pdl = true;
bool lucid::pdlEnabled() { return g_frame_pdl; }
static int pdlMode() { // 0 auto, 1 on, 2 off
	static const int mode = [] {
		if(getenv("LUCID_NO_PDL"))
			return 2;
		const char *v = getenv("LUCID_PDL");
		return !v ? 0 : v[1] == 'n' ? 1 : v[1] == 'f' ? 2 : 0;
	}();
	return mode;
}
constexpr float PDL_MAX_FRAME_MS = 1.0f;

static thread_local const char *g_failed_kernel = nullptr;
static thread_local cudaError_t g_failed_err = cudaSuccess;
void lucid::noteLaunchFailure(const char *kernel_name, cudaError_t err) {
	if(!g_failed_kernel)
		g_failed_kernel = kernel_name, g_failed_err = err;
}

struct lucid_renderer {
	LucidCreateInfo ci;
	Params p;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	int num_sms = 148;
	std::string error;

	// owned device allocations
	std::vector<void *> owned;
	void *geom_owned[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // [5] padded positions
	cudaMipmappedArray_t tex_array[2] = {nullptr, nullptr};
	cudaTextureObject_t tex_object[2] = {0, 0};
	// two renderer-owned images: while frame n is copied to the host on copy_stream, frame n+1
	// renders into the other one (the reference double-buffers its per-frame data the same way,
	// lucid_renderer.h:78-81)
	u32 *images[2] = {nullptr, nullptr};
	u32 *image = nullptr; // the renderer-owned image the last frame rendered into; null after a frame that
						  // rendered into a caller's LUCID_MEM_DEVICE image
	int image_index = 0;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t render_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
	bool copy_pending[2] = {false, false};
	static constexpr int NUM_STAGING = 3;
	cudaEvent_t upload_done[NUM_STAGING] = {nullptr, nullptr, nullptr};
	bool upload_pending[NUM_STAGING] = {false, false, false};
	u32 *frag_counts = nullptr;
	void *info_dev = nullptr;
	size_t info_words = 0;
	// per-frame instance data: a ring of device blocks filled by a copy on upload_stream at submission
	// time (instances, then uv rects, then colours, tightly packed), so the frame's kernels never read
	// host memory -- zero-copy reads queue behind the posted writes of an image read-back on PCIe
	unsigned char *d_inst_ring[NUM_STAGING] = {nullptr, nullptr, nullptr};
	// frame hand-over flags of the bin-row split (lucid_signal / lucid_wait_flags) and the one-shot gate of the
	// next frame's raster kernels
	u32 *d_sync = nullptr;
	const u32 *gate_flag = nullptr;
	u32 gate_value = 0;
	// LUCID_RENDER_CULL_INSTANCES: boxes of the instance list they were computed for
	float4 *d_inst_boxes = nullptr;
	u32 *d_active_instances = nullptr;
	std::vector<LucidInstanceData> boxed_instances;
	bool boxes_valid = false;
	cudaStream_t upload_stream = nullptr;
	cudaEvent_t upload_ready[NUM_STAGING] = {nullptr, nullptr, nullptr};

	// pinned host staging
	unsigned char *h_instances = nullptr; // instances + colors + uv rects
	u32 *h_info = nullptr;
	// one status word per staging slot, written by the frame's kernels only when the bin lists overflowed
	u32 *h_status = nullptr;
	bool info_valid = false;
	bool has_geometry = false;
	int num_quads = 0, num_verts = 0;

	// per-stage events of the last TIMING_RING frames
	static constexpr int TIMING_RING = 64;
	cudaEvent_t ev[TIMING_RING][8];
	bool ev_staged[TIMING_RING] = {};
	float last_frame_ms = 0.0f; // the latest frame known to be finished (programmatic dependent launch: on or off)
	// lucid_compare_render (comparators.cu), allocated on first use: submission order per visible-quad slot, sort
	// scratch for lists over 1024 entries, the work-item ticket, the comparator's image
	u32 *cmp_order = nullptr;
	u64 *cmp_scratch = nullptr;
	u32 *cmp_ticket = nullptr;
	u32 *cmp_image = nullptr;
	long long frame_counter = 0;
	bool pending = false;
};

namespace {

constexpr size_t STAGING_BYTES = (size_t)LUCID_MAX_INSTANCES * (16 + 4 + 16);

int fail(lucid_renderer *r, int code, const std::string &msg) {
	if(r)
		r->error = msg;
	else
		g_create_error = msg;
	return code;
}
int failCuda(lucid_renderer *r, cudaError_t e, const char *what) {
	return fail(r, LUCID_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if(e_ != cudaSuccess)                                                                      \
			return failCuda(r, e_, #call);                                                         \
	} while(0)

template <class T> cudaError_t devAlloc(lucid_renderer *r, T **ptr, size_t count) {
	void *v = nullptr;
	cudaError_t e = cudaMalloc(&v, count * sizeof(T));
	if(e == cudaSuccess) {
		r->owned.push_back(v);
		*ptr = (T *)v;
	}
	return e;
}

void freeAll(lucid_renderer *r) {
	for(void *v : r->owned)
		cudaFree(v);
	for(void *v : r->geom_owned)
		if(v)
			cudaFree(v);
	for(int i = 0; i < 2; i++) {
		if(r->tex_object[i])
			cudaDestroyTextureObject(r->tex_object[i]);
		if(r->tex_array[i])
			cudaFreeMipmappedArray(r->tex_array[i]);
	}
	if(r->h_instances)
		cudaFreeHost(r->h_instances);
	if(r->h_info)
		cudaFreeHost(r->h_info);
	if(r->h_status)
		cudaFreeHost(r->h_status);
	for(auto &set : r->ev)
		for(auto &e : set)
			if(e)
				cudaEventDestroy(e);
	for(int i = 0; i < 2; i++) {
		if(r->render_done[i])
			cudaEventDestroy(r->render_done[i]);
		if(r->copy_done[i])
			cudaEventDestroy(r->copy_done[i]);
	}
	for(cudaEvent_t e : r->upload_done)
		if(e)
			cudaEventDestroy(e);
	if(r->copy_stream)
		cudaStreamDestroy(r->copy_stream);
	if(r->upload_stream)
		cudaStreamDestroy(r->upload_stream);
	for(cudaEvent_t e : r->upload_ready)
		if(e)
			cudaEventDestroy(e);
	if(r->own_stream && r->stream)
		cudaStreamDestroy(r->stream);
}

} // namespace

extern "C" {

const char *lucid_last_error(const lucid_renderer *r) {
	return r ? r->error.c_str() : g_create_error.c_str();
}

int lucid_create(const LucidCreateInfo *info, lucid_renderer **out) {
	lucid_renderer *r = nullptr;
	if(!info || !out)
		return fail(nullptr, LUCID_E_INVALID, "lucid_create: null argument");
	*out = nullptr;
	if(info->width <= 0 || info->height <= 0)
		return fail(nullptr, LUCID_E_INVALID, "lucid_create: bad view size");
	// 7-bit bin coordinates (funcs.glsl:52-59): at most 128 bins of 32 pixels per axis
	if(info->width > 4096 || info->height > 4096)
		return fail(nullptr, LUCID_E_LIMIT, "lucid_create: view size above 4096 (7-bit bin coordinates)");
	int dev_count = 0;
	cudaError_t e = cudaGetDeviceCount(&dev_count);
	if(e != cudaSuccess || dev_count == 0)
		return fail(nullptr, LUCID_E_CUDA,
					std::string("lucid_create: no CUDA device (") + cudaGetErrorString(e) + ")");
	if(info->device < 0 || info->device >= dev_count)
		return fail(nullptr, LUCID_E_INVALID, "lucid_create: bad device ordinal");

	r = new lucid_renderer();
	memset(r->ev, 0, sizeof(r->ev));
	r->ci = *info;
	lucid_renderer *keep = r;
	auto bail = [&](int code) {
		g_create_error = keep->error;
		freeAll(keep);
		delete keep;
		return code;
	};
#define CUC(call)                                                                                  \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if(e_ != cudaSuccess) {                                                                    \
			failCuda(r, e_, #call);                                                                \
			return bail(LUCID_E_CUDA);                                                             \
		}                                                                                          \
	} while(0)
	CUC(cudaSetDevice(info->device));
	cudaDeviceProp prop;
	CUC(cudaGetDeviceProperties(&prop, info->device));
	r->num_sms = prop.multiProcessorCount;
	if(info->stream) {
		r->stream = (cudaStream_t)info->stream;
	} else {
		CUC(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
		r->own_stream = true;
	}
	for(auto &set : r->ev)
		for(auto &ev : set)
			CUC(cudaEventCreate(&ev));
	CUC(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking));
	CUC(cudaStreamCreateWithFlags(&r->upload_stream, cudaStreamNonBlocking));
	for(int i = 0; i < 2; i++) {
		CUC(cudaEventCreateWithFlags(&r->render_done[i], cudaEventDisableTiming));
		CUC(cudaEventCreateWithFlags(&r->copy_done[i], cudaEventDisableTiming));
	}

	Params &p = r->p;
	memset(&p, 0, sizeof(p));
	p.width = info->width, p.height = info->height;
	p.bin_count_x = (info->width + BIN_SIZE - 1) / BIN_SIZE;
	p.bin_count_y = (info->height + BIN_SIZE - 1) / BIN_SIZE;
	p.bin_count = p.bin_count_x * p.bin_count_y;
	p.max_visible_quads = info->max_visible_quads > 0 ? info->max_visible_quads : 4793490;
	p.max_dispatches = info->max_dispatches > 0 ? info->max_dispatches : 256;
	if(p.max_dispatches > LUCID_INFO_MAX_DISPATCHES)
		p.max_dispatches = LUCID_INFO_MAX_DISPATCHES;
	// 24-bit triangle index in sample words and 28-bit quad index in bin lists
	if(p.max_visible_quads > (1 << 23)) {
		fail(r, LUCID_E_LIMIT, "lucid_create: max_visible_quads above 2^23 (24-bit triangle index)");
		return bail(LUCID_E_LIMIT);
	}
	p.opts = info->opts;
	p.row_begin = 0, p.row_end = p.bin_count_y;
	if(info->bin_row_end > info->bin_row_begin) {
		p.row_begin = info->bin_row_begin < 0 ? 0 : info->bin_row_begin;
		p.row_end = info->bin_row_end > p.bin_count_y ? p.bin_count_y : info->bin_row_end;
	}

	const size_t mvq = (size_t)p.max_visible_quads;
	CUC(devAlloc(r, &p.quad_aabbs, mvq));
	CUC(devAlloc(r, &p.quad_verts, mvq));
	CUC(devAlloc(r, &p.quad_setup_info, mvq));
	CUC(devAlloc(r, &p.tri_scan, mvq * 2));
	CUC(devAlloc(r, &p.tri_shade, mvq * 2));
	CUC(devAlloc(r, &p.quad_colors, mvq));
	CUC(devAlloc(r, &p.quad_normals, mvq));
	CUC(devAlloc(r, &p.quad_uv, mvq * 2));
	p.bin_list_capacity = (u32)(mvq * 2);
	CUC(devAlloc(r, &p.bin_quads, (size_t)p.bin_list_capacity));
	CUC(devAlloc(r, &p.bin_tris, (size_t)p.bin_list_capacity));
	r->info_words = LUCID_INFO_U32_SIZE + (size_t)p.bin_count * LUCID_COUNTS_PER_BIN;
	u32 *info_dev = nullptr;
	CUC(devAlloc(r, &info_dev, r->info_words));
	r->info_dev = info_dev;
	p.info = reinterpret_cast<LucidInfo *>(info_dev);
	p.counts = reinterpret_cast<int *>(info_dev + LUCID_INFO_U32_SIZE);
	CUC(devAlloc(r, &p.setup_lookback, (size_t)LUCID_MAX_INSTANCES * 4));
	CUC(devAlloc(r, &p.setup_ticket, 4));
	CUC(devAlloc(r, &p.bin_flags, (size_t)p.bin_count));
	CUC(devAlloc(r, &p.work_counters, (size_t)WORK_COUNTERS));
	CUC(devAlloc(r, &p.bin_cost, (size_t)p.bin_count));
	p.bin_begin = p.row_begin * p.bin_count_x, p.bin_end = p.row_end * p.bin_count_x;
	CUC(devAlloc(r, &p.block_counts, (size_t)p.bin_count * 32));
	p.block_items_cap = (u32)p.bin_count * 32u;
	CUC(devAlloc(r, &p.block_items, (size_t)p.block_items_cap * 5)); // one region per size class (ITEM_CLASSES)
	CUC(devAlloc(r, &p.large_keys, rasterLargeKeysCount(r->num_sms)));
	CUC(devAlloc(r, &p.tie_runs, (size_t)2 + 2 * TIE_RUN_QUEUE));
	CUC(devAlloc(r, &p.tie_scratch, rasterTieScratchCount(r->num_sms)));
	// sorted-entry stream: one entry per (triangle, half-block or block) pair of the frame
	{
		const unsigned long long def = std::max<unsigned long long>(16ull * mvq, 1ull << 22);
		const unsigned long long want = info->max_block_entries > 0 ? (unsigned long long)info->max_block_entries : def;
		p.stream_capacity = (u32)std::min<unsigned long long>(want, 0x7fffffffull);
	}
	p.compact_lists = (info->flags & LUCID_CREATE_COMPACT_LISTS) != 0;
	p.list_offsets = nullptr, p.list_pool_units = 0;
	if(p.compact_lists) {
		// a stream entry is one list record: 16 bytes in a LOW list, 8 in a HIGH list
		p.list_pool_units = (u32)std::min<unsigned long long>(2ull * p.stream_capacity, 0xfffffff0ull);
		CUC(devAlloc(r, &p.block_lists, (size_t)p.list_pool_units / 2 + 1));
		CUC(devAlloc(r, &p.list_offsets, (size_t)p.bin_count * 32));
	} else {
		CUC(devAlloc(r, &p.block_lists, (size_t)p.bin_count * (BIN_LIST_BYTES / sizeof(uint4))));
	}
	CUC(devAlloc(r, &p.sorted_rec, (size_t)p.stream_capacity));
	CUC(devAlloc(r, &p.sorted_aux, (size_t)p.stream_capacity));
	if(info->opts & LUCID_OPT_OPAQUE_PREPASS)
		CUC(devAlloc(r, &p.opaque_depth, (size_t)p.bin_count * BIN_SIZE * BIN_SIZE));
	CUC(devAlloc(r, &r->images[0], (size_t)p.width * p.height));
	CUC(devAlloc(r, &r->images[1], (size_t)p.width * p.height));
	r->image = r->images[0];
	CUC(devAlloc(r, &r->frag_counts, (size_t)p.width * p.height));
	CUC(devAlloc(r, &r->d_inst_boxes, (size_t)LUCID_MAX_INSTANCES * 2));
	CUC(devAlloc(r, &r->d_active_instances, (size_t)LUCID_MAX_INSTANCES + 1));
	CUC(devAlloc(r, &r->d_sync, (size_t)LUCID_SYNC_FLAGS));
	CUC(cudaMemsetAsync(r->d_sync, 0, LUCID_SYNC_FLAGS * 4, r->stream));
	p.debug_records = nullptr;
	if(p.opts & LUCID_OPT_DEBUG_RASTER) {
		const size_t words = 2 + (size_t)LUCID_DEBUG_MAX_RECORDS * LUCID_DEBUG_RECORD_WORDS;
		CUC(devAlloc(r, &p.debug_records, words));
		CUC(cudaMemsetAsync(p.debug_records, 0, words * 4, r->stream));
		const u32 cap = LUCID_DEBUG_MAX_RECORDS;
		CUC(cudaMemcpyAsync(p.debug_records, &cap, 4, cudaMemcpyHostToDevice, r->stream));
		// test hook: the sort of the first work item's list is undone by one swap, so the UNSORTED check has something to find
		const char *inject = getenv("LUCID_DEBUG_RASTER_INJECT");
		p.debug_inject = inject && inject[0] == '1';
	}
	for(int i = 0; i < lucid_renderer::NUM_STAGING; i++) {
		CUC(devAlloc(r, &r->d_inst_ring[i], STAGING_BYTES));
		CUC(cudaEventCreateWithFlags(&r->upload_ready[i], cudaEventDisableTiming));
	}
	CUC(cudaMallocHost((void **)&r->h_instances, lucid_renderer::NUM_STAGING * STAGING_BYTES));
	for(int i = 0; i < lucid_renderer::NUM_STAGING; i++)
		CUC(cudaEventCreateWithFlags(&r->upload_done[i], cudaEventDisableTiming));
	CUC(cudaMallocHost((void **)&r->h_info, r->info_words * 4));
	CUC(cudaMallocHost((void **)&r->h_status, lucid_renderer::NUM_STAGING * 4));
	memset(r->h_status, 0, lucid_renderer::NUM_STAGING * 4);
	CUC(cudaMemsetAsync(r->info_dev, 0, r->info_words * 4, r->stream));
	CUC(cudaMemsetAsync(r->images[0], 0, (size_t)p.width * p.height * 4, r->stream));
	CUC(cudaMemsetAsync(r->images[1], 0, (size_t)p.width * p.height * 4, r->stream));
	CUC(cudaStreamSynchronize(r->stream));
#undef CUC
	*out = r;
	return LUCID_OK;
}

void lucid_destroy(lucid_renderer *r) {
	if(!r)
		return;
	cudaSetDevice(r->ci.device);
	cudaStreamSynchronize(r->stream);
	freeAll(r);
	delete r;
}

int lucid_bin_count(const lucid_renderer *r) { return r ? r->p.bin_count : LUCID_E_INVALID; }

int lucid_set_bin_rows(lucid_renderer *r, int32_t begin, int32_t end) {
	if(!r)
		return LUCID_E_INVALID;
	if(begin < 0 || end > r->p.bin_count_y || begin >= end)
		return fail(r, LUCID_E_INVALID, "lucid_set_bin_rows: bad range");
	r->p.row_begin = begin, r->p.row_end = end;
	r->p.bin_begin = begin * r->p.bin_count_x, r->p.bin_end = end * r->p.bin_count_x;
	return LUCID_OK;
}

int lucid_set_bin_range(lucid_renderer *r, int32_t begin, int32_t end) {
	if(!r)
		return LUCID_E_INVALID;
	if(begin < 0 || end > r->p.bin_count || begin >= end)
		return fail(r, LUCID_E_INVALID, "lucid_set_bin_range: bad range");
	r->p.bin_begin = begin, r->p.bin_end = end;
	r->p.row_begin = begin / r->p.bin_count_x, r->p.row_end = (end - 1) / r->p.bin_count_x + 1;
	return LUCID_OK;
}

int lucid_read_bin_costs(lucid_renderer *r, uint64_t *dst, int32_t num_bins) {
	if(!r || !dst || num_bins < 0)
		return LUCID_E_INVALID;
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	const size_t n = (size_t)std::min(num_bins, r->p.bin_count);
	CU(cudaMemcpy(dst, r->p.bin_cost, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_read_row_costs(lucid_renderer *r, uint64_t *dst, int32_t num_rows) {
	if(!r || !dst || num_rows < 0)
		return LUCID_E_INVALID;
	std::vector<uint64_t> bins((size_t)r->p.bin_count);
	int rc = lucid_read_bin_costs(r, bins.data(), r->p.bin_count);
	if(rc)
		return rc;
	for(int by = 0; by < std::min(num_rows, r->p.bin_count_y); by++) {
		uint64_t sum = 0;
		for(int bx = 0; bx < r->p.bin_count_x; bx++)
			sum += bins[(size_t)by * r->p.bin_count_x + bx];
		dst[by] = sum;
	}
	return LUCID_OK;
}

int lucid_set_geometry(lucid_renderer *r, const float *positions, int32_t num_verts,
					   const uint32_t *colors, const float *uvs, const uint32_t *normals,
					   const uint32_t *quad_indices, int32_t num_quads, int32_t memory) {
	if(!r)
		return LUCID_E_INVALID;
	if(!positions || !quad_indices || num_verts <= 0 || num_quads < 0)
		return fail(r, LUCID_E_INVALID, "lucid_set_geometry: positions and quad indices are required");
	CU(cudaSetDevice(r->ci.device));
	CU(cudaStreamSynchronize(r->stream));
	for(auto &v : r->geom_owned)
		if(v) {
			cudaFree(v);
			v = nullptr;
		}
	Params &p = r->p;
	if(memory == LUCID_MEM_DEVICE) {
		p.positions = positions, p.quad_indices = (const uint4 *)quad_indices;
		p.vertex_colors = colors, p.vertex_uvs = (const float2 *)uvs, p.vertex_normals = normals;
	} else if(memory == LUCID_MEM_HOST) {
		const void *src[5] = {positions, quad_indices, colors, uvs, normals};
		size_t bytes[5] = {(size_t)num_verts * 12, (size_t)num_quads * 16, (size_t)num_verts * 4,
						   (size_t)num_verts * 8, (size_t)num_verts * 4};
		for(int i = 0; i < 5; i++) {
			if(!src[i] || bytes[i] == 0)
				continue;
			CU(cudaMalloc(&r->geom_owned[i], bytes[i]));
			CU(cudaMemcpyAsync(r->geom_owned[i], src[i], bytes[i], cudaMemcpyHostToDevice, r->stream));
		}
		CU(cudaStreamSynchronize(r->stream));
		p.positions = (const float *)r->geom_owned[0];
		p.quad_indices = (const uint4 *)r->geom_owned[1];
		p.vertex_colors = (const u32 *)r->geom_owned[2];
		p.vertex_uvs = (const float2 *)r->geom_owned[3];
		p.vertex_normals = (const u32 *)r->geom_owned[4];
	} else {
		return fail(r, LUCID_E_INVALID, "lucid_set_geometry: bad memory kind");
	}
	if(((uintptr_t)p.quad_indices & 15) != 0)
		return fail(r, LUCID_E_INVALID, "lucid_set_geometry: quad indices must be 16-byte aligned");
	CU(cudaMalloc(&r->geom_owned[5], (size_t)num_verts * 16));
	p.positions4 = (const float4 *)r->geom_owned[5];
	launchPadPositions(p.positions, (float4 *)r->geom_owned[5], num_verts, r->stream);
	CU(cudaStreamSynchronize(r->stream));
	r->num_quads = num_quads, r->num_verts = num_verts;
	p.num_verts = num_verts;
	r->boxes_valid = false;
	r->has_geometry = true;
	return LUCID_OK;
}

int lucid_set_texture(lucid_renderer *r, int32_t slot, const uint8_t *data, int32_t width,
					  int32_t height, int32_t levels) {
	if(!r)
		return LUCID_E_INVALID;
	if(slot < 0 || slot > 1 || !data || width <= 0 || height <= 0 || levels <= 0 || levels > 16)
		return fail(r, LUCID_E_INVALID, "lucid_set_texture: bad argument");
	CU(cudaSetDevice(r->ci.device));
	CU(cudaStreamSynchronize(r->stream));
	// the chain goes into a mipmapped array sampled by the texture unit (raster_common.cuh sampleTexture)
	if(r->tex_object[slot]) {
		cudaDestroyTextureObject(r->tex_object[slot]);
		r->tex_object[slot] = 0;
	}
	if(r->tex_array[slot]) {
		cudaFreeMipmappedArray(r->tex_array[slot]);
		r->tex_array[slot] = nullptr;
	}
	r->p.tex_object[slot] = 0;
	if((width & (width - 1)) != 0 || (height & (height - 1)) != 0)
		return fail(r, LUCID_E_INVALID, "lucid_set_texture: width and height must be powers of two (the filter arithmetic of the "
										"texture unit is only pinned for those)");
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
	CU(cudaMallocMipmappedArray(&r->tex_array[slot], &desc, make_cudaExtent((size_t)width, (size_t)height, 0), (unsigned)levels));
	size_t off = 0;
	for(int l = 0; l < levels; l++) {
		const int lw = std::max(1, width >> l), lh = std::max(1, height >> l);
		cudaArray_t level;
		CU(cudaGetMipmappedArrayLevel(&level, r->tex_array[slot], (unsigned)l));
		CU(cudaMemcpy2DToArray(level, 0, 0, data + off, (size_t)lw * 4, (size_t)lw * 4, (size_t)lh, cudaMemcpyHostToDevice));
		off += (size_t)lw * lh * 4;
	}
	cudaResourceDesc res{};
	res.resType = cudaResourceTypeMipmappedArray;
	res.res.mipmap.mipmap = r->tex_array[slot];
	cudaTextureDesc td{};
	td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
	td.filterMode = cudaFilterModeLinear, td.mipmapFilterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat, td.normalizedCoords = 1;
	td.maxAnisotropy = 1, td.minMipmapLevelClamp = 0.0f, td.maxMipmapLevelClamp = float(levels - 1);
	CU(cudaCreateTextureObject(&r->tex_object[slot], &res, &td, nullptr));
	r->p.tex_object[slot] = (unsigned long long)r->tex_object[slot];
	r->p.tex_width[slot] = width, r->p.tex_height[slot] = height, r->p.tex_levels[slot] = levels;
	return LUCID_OK;
}

// test hook: n samples (u, v, lod) of the texture in `slot` straight through the texture unit
__global__ void k_debug_sample_texture(cudaTextureObject_t tex, const float *uvl, float4 *out, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n)
		out[i] = tex2DLod<float4>(tex, uvl[i * 3], uvl[i * 3 + 1], uvl[i * 3 + 2]);
}
int lucid_debug_sample_texture(lucid_renderer *r, int32_t slot, const float *uvl, int32_t n, float *out_rgba) {
	if(!r || slot < 0 || slot > 1 || !uvl || !out_rgba || n < 0)
		return LUCID_E_INVALID;
	if(!r->tex_object[slot])
		return fail(r, LUCID_E_STATE, "lucid_debug_sample_texture: no texture in the slot");
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	float *d_in = nullptr;
	float4 *d_out = nullptr;
	CU(cudaMalloc(&d_in, (size_t)n * 12 + 16));
	CU(cudaMalloc(&d_out, (size_t)n * 16 + 16));
	CU(cudaMemcpy(d_in, uvl, (size_t)n * 12, cudaMemcpyHostToDevice));
	if(n > 0)
		k_debug_sample_texture<<<(n + 255) / 256, 256, 0, r->stream>>>(r->tex_object[slot], d_in, d_out, n);
	CU(cudaStreamSynchronize(r->stream));
	CU(cudaMemcpy(out_rgba, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost));
	cudaFree(d_in), cudaFree(d_out);
	return LUCID_OK;
}

int lucid_wait(lucid_renderer *r) {
	if(!r)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	CU(cudaStreamSynchronize(r->stream));
	CU(cudaStreamSynchronize(r->copy_stream));
	CU(cudaStreamSynchronize(r->upload_stream));
	r->pending = false;
	r->copy_pending[0] = r->copy_pending[1] = false;
	for(bool &b : r->upload_pending)
		b = false;
	CU(cudaGetLastError());
	bool overflow = false, timed_out = false;
	for(int i = 0; i < lucid_renderer::NUM_STAGING; i++) {
		overflow = overflow || (r->h_status[i] & 3u) != 0;
		timed_out = timed_out || (r->h_status[i] & 4u) != 0;
		r->h_status[i] = 0;
	}
	if(timed_out)
		return fail(r, LUCID_E_STATE, "a device-side wait for a peer's frame flag gave up after 5 s (lucid_wait_flags / frame gate)");
	if(overflow)
		return fail(r, LUCID_E_LIMIT,
					"a frame exceeded the renderer's list storage (per-bin lists of 2 * max_visible_quads entries, or the "
					"sorted-entry stream of max_block_entries): the affected bins were painted red; create the renderer "
					"with a larger max_visible_quads / max_block_entries");
	return LUCID_OK;
}

int lucid_render(lucid_renderer *r, const LucidConfig *config, const LucidInstanceData *instances,
				 const uint32_t *instance_colors, const float *instance_uv_rects,
				 int32_t num_instances, void *out_rgba8, size_t pitch_bytes, int32_t out_memory,
				 uint32_t flags) {
	if(!r)
		return LUCID_E_INVALID;
	if(!config || num_instances < 0 || (num_instances > 0 && (!instances || !instance_colors)))
		return fail(r, LUCID_E_INVALID, "lucid_render: null argument");
	if(!r->has_geometry)
		return fail(r, LUCID_E_STATE, "lucid_render: lucid_set_geometry has not been called");
	if(num_instances > LUCID_MAX_INSTANCES)
		num_instances = LUCID_MAX_INSTANCES; // lucid_renderer.cpp:403-410 truncates the same way
	Params &p = r->p;
	for(int i = 0; i < num_instances; i++) {
		const LucidInstanceData &in = instances[i];
		if(in.num_quads < 0 || in.num_quads > LUCID_MAX_INSTANCE_QUADS || (in.index_offset & 3) != 0 ||
		   in.index_offset < 0 || (int64_t)in.index_offset / 4 + in.num_quads > r->num_quads ||
		   in.vertex_offset < 0 || in.vertex_offset >= r->num_verts)
			return fail(r, LUCID_E_INVALID, "lucid_render: instance " + std::to_string(i) + " out of range");
	}
	if(out_memory != LUCID_MEM_NONE && (!out_rgba8 || pitch_bytes < (size_t)p.width * 4 || (pitch_bytes & 3)))
		return fail(r, LUCID_E_INVALID, "lucid_render: bad output image");
	static const bool host_profile = getenv("LUCID_PROFILE_HOST") != nullptr;
	static double t_acc[6] = {0, 0, 0, 0, 0, 0};
	static long t_frames = 0;
	auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t0 = host_profile ? now() : 0.0;
	CU(cudaSetDevice(r->ci.device));
	cudaStream_t st = r->stream;
	// three pinned staging blocks: the host fills one while the frames that read the other two may
	// still be queued (a block is free again when the frame that read it has completed: its event
	// is recorded at the end of the frame, so no event sits between two kernels of a frame)
	const int sb = (int)(r->frame_counter % lucid_renderer::NUM_STAGING);
	if(r->upload_pending[sb]) {
		CU(cudaEventSynchronize(r->upload_done[sb]));
		r->upload_pending[sb] = false;
	}

	// per-frame uploads (uploadInstances / setupInputData): one pinned staging block, copied to the
	// ring's device block on the upload stream right away -- in asynchronous use the host runs ahead of
	// the GPU, so the data is resident long before the frame's first kernel
	unsigned char *h = r->h_instances + (size_t)sb * STAGING_BYTES;
	size_t n = (size_t)num_instances;
	memcpy(h, instances, n * 16);
	float *h_uv = reinterpret_cast<float *>(h + n * 16);
	if(instance_uv_rects) {
		memcpy(h_uv, instance_uv_rects, n * 16);
	} else {
		for(size_t i = 0; i < n; i++)
			h_uv[i * 4] = 0.0f, h_uv[i * 4 + 1] = 0.0f, h_uv[i * 4 + 2] = 1.0f, h_uv[i * 4 + 3] = 1.0f;
	}
	memcpy(h + n * 32, instance_colors, n * 4);
	unsigned char *d = r->d_inst_ring[sb];
	p.instances = reinterpret_cast<const LucidInstanceData *>(d);
	p.inst_uv_rects = reinterpret_cast<const float4 *>(d + n * 16);
	p.inst_colors = reinterpret_cast<const u32 *>(d + n * 32);
	if(n > 0)
		CU(cudaMemcpyAsync(d, h, n * 36, cudaMemcpyHostToDevice, r->upload_stream));
	CU(cudaEventRecord(r->upload_ready[sb], r->upload_stream));
	CU(cudaStreamWaitEvent(st, r->upload_ready[sb], 0));
	LucidConfig cfg = *config;
	cfg.num_instances = num_instances; // taken from this call, not the previous frame
	p.num_instances = num_instances;
	p.num_setup_ctas = num_instances; // one k_quad_cull CTA per instance
	p.host_status = r->h_status + sb;
	g_failed_kernel = nullptr;
	if(out_memory == LUCID_MEM_DEVICE) {
		p.image = (u32 *)out_rgba8;
		p.image_pitch = (int)(pitch_bytes / 4);
		r->image = nullptr; // no renderer-owned image holds this frame
	} else {
		// frames that stay on the device always use image 0 (the one lucid_image_pointer and the
		// IPC export name); frames copied to the host alternate
		const int b = out_memory == LUCID_MEM_HOST ? (r->image_index ^ 1) : 0;
		if(r->copy_pending[b]) // the image is still being copied out by an earlier frame
			CU(cudaStreamWaitEvent(st, r->copy_done[b], 0));
		r->image_index = b;
		r->image = r->images[b];
		p.image = r->image;
		p.image_pitch = p.width;
	}
	p.frag_counts = (flags & LUCID_RENDER_FRAG_COUNTS) ? r->frag_counts : nullptr;
	// instance boxes for the early instance cull: recomputed only when the instance list changes
	p.inst_boxes = nullptr;
	p.active_instances = nullptr;
	if((flags & LUCID_RENDER_CULL_INSTANCES) && num_instances > 0) {
		if(!r->boxes_valid || r->boxed_instances.size() != n ||
		   memcmp(r->boxed_instances.data(), instances, n * sizeof(LucidInstanceData)) != 0) {
			r->boxed_instances.assign(instances, instances + n);
			launchInstanceBoxes(p, r->d_inst_boxes, st); // after the upload wait above; before the frame's first kernel
			r->boxes_valid = true;
		}
		p.inst_boxes = r->d_inst_boxes;
		p.active_instances = r->d_active_instances;
	}

	// the time of the latest finished frame, without waiting for anything
	for(int back = 1; back <= 4 && back <= (int)std::min<long long>(r->frame_counter, 4); back++) {
		cudaEvent_t *pe = r->ev[(r->frame_counter - back) % lucid_renderer::TIMING_RING];
		if(cudaEventQuery(pe[7]) == cudaSuccess) {
			float ms = 0.0f;
			if(cudaEventElapsedTime(&ms, pe[0], pe[7]) == cudaSuccess)
				r->last_frame_ms = ms;
			break;
		}
	}
	cudaGetLastError(); // cudaErrorNotReady of the query is not an error of this call
	g_frame_pdl = pdlMode() == 1 || (pdlMode() == 0 && !(flags & LUCID_RENDER_NO_DEPENDENT_LAUNCH) && r->last_frame_ms < PDL_MAX_FRAME_MS);

	double t1 = host_profile ? now() : 0.0;
	const int ring = (int)(r->frame_counter % lucid_renderer::TIMING_RING);
	cudaEvent_t *ev = r->ev[ring];
	// per-stage events sit between the kernels and keep each launch from overlapping the tail of
	// its predecessor (programmatic dependent launch): frames that do not ask for them only get
	// the frame's first and last event
	const bool stage_events = !(flags & LUCID_RENDER_NO_STAGE_TIMES);
	r->ev_staged[ring] = stage_events;
	CU(cudaEventRecord(ev[0], st));
	// per-frame clears (setupInputData): LucidInfo and the first 6 per-bin arrays (lucid_renderer.cpp:437)
	launchFrameBegin(p, st);
	if(p.debug_records)
		CU(cudaMemsetAsync(p.debug_records + 1, 0, 4, st));
	launchQuadSetup(p, cfg, st);
	if(stage_events)
		CU(cudaEventRecord(ev[1], st));
	double t2 = host_profile ? now() : 0.0;
	launchBinning(p, st, stage_events ? &ev[2] : nullptr); // ev[2] count, ev[3] scan, ev[4] dispatch
	double t3 = host_profile ? now() : 0.0;
	// a shared image may only be stored into once the gathering device has released it (lucid_set_frame_gate)
	if(r->gate_flag) {
		launchWaitFlags(r->gate_flag, 1, r->gate_value, p.host_status, 5000000000ull, st);
		r->gate_flag = nullptr;
	}
	// ev[5] block lists, ev[6] block sort (+ frame bookkeeping), ev[7] shading = end of frame
	launchRaster(p, cfg, st, stage_events ? &ev[5] : nullptr, r->num_sms);
	if(!stage_events)
		CU(cudaEventRecord(ev[7], st));
	CU(cudaEventRecord(r->upload_done[sb], st));
	r->upload_pending[sb] = true;
	if(g_failed_kernel) {
		const std::string what = std::string("launch of ") + g_failed_kernel;
		g_failed_kernel = nullptr;
		cudaGetLastError();
		return failCuda(r, g_failed_err, what.c_str());
	}
	CU(cudaGetLastError());
	r->frame_counter++;
	double t4 = host_profile ? now() : 0.0;

	if(!(flags & LUCID_RENDER_SKIP_INFO)) {
		launchInfoOut(p, r->h_info, (int)r->info_words, st);
		r->info_valid = true;
	}
	if(out_memory == LUCID_MEM_HOST) {
		const int b = r->image_index;
		CU(cudaEventRecord(r->render_done[b], st));
		CU(cudaStreamWaitEvent(r->copy_stream, r->render_done[b], 0));
		if((flags & LUCID_RENDER_OWNED_BINS_ONLY) && (p.bin_begin > 0 || p.bin_end < p.bin_count)) {
			// the owned bins [bin_begin, bin_end) in row-major order: part of a first bin row, whole rows, part of a last one
			auto copyBins = [&](int by0, int by1, int bx0, int bx1) -> cudaError_t { // bin rows [by0, by1), bin columns [bx0, bx1)
				const int x0 = bx0 * BIN_SIZE, x1 = std::min(p.width, bx1 * BIN_SIZE);
				const int y0 = by0 * BIN_SIZE, y1 = std::min(p.height, by1 * BIN_SIZE);
				if(x1 <= x0 || y1 <= y0)
					return cudaSuccess;
				return cudaMemcpy2DAsync((char *)out_rgba8 + (size_t)y0 * pitch_bytes + (size_t)x0 * 4, pitch_bytes,
										 r->image + (size_t)y0 * p.width + x0, (size_t)p.width * 4, (size_t)(x1 - x0) * 4,
										 (size_t)(y1 - y0), cudaMemcpyDeviceToHost, r->copy_stream);
			};
			const int nbx = p.bin_count_x, last = p.bin_end - 1;
			const int by0 = p.bin_begin / nbx, bx0 = p.bin_begin % nbx, by1 = last / nbx, bx1 = last % nbx + 1;
			if(p.bin_end > p.bin_begin) {
				if(by0 == by1) {
					CU(copyBins(by0, by0 + 1, bx0, bx1));
				} else {
					CU(copyBins(by0, by0 + 1, bx0, nbx));
					CU(copyBins(by0 + 1, by1, 0, nbx));
					CU(copyBins(by1, by1 + 1, 0, bx1));
				}
			}
		} else if(pitch_bytes == (size_t)p.width * 4)
			CU(cudaMemcpyAsync(out_rgba8, r->image, pitch_bytes * p.height, cudaMemcpyDeviceToHost, r->copy_stream));
		else
			CU(cudaMemcpy2DAsync(out_rgba8, pitch_bytes, r->image, (size_t)p.width * 4, (size_t)p.width * 4,
								 p.height, cudaMemcpyDeviceToHost, r->copy_stream));
		CU(cudaEventRecord(r->copy_done[b], r->copy_stream));
		r->copy_pending[b] = true;
	}
	r->pending = true;
	if(host_profile) {
		double t5 = now();
		t_acc[0] += t1 - t0, t_acc[1] += t2 - t1, t_acc[2] += t3 - t2, t_acc[3] += t4 - t3, t_acc[4] += t5 - t4;
		if(++t_frames % 100 == 0) {
			fprintf(stderr, "lucid_render host us/frame: uploads %.1f setup %.1f binning %.1f raster %.1f readback %.1f\n",
					t_acc[0] / 100, t_acc[1] / 100, t_acc[2] / 100, t_acc[3] / 100, t_acc[4] / 100);
			for(double &t : t_acc)
				t = 0;
		}
	}
	if(!(flags & LUCID_RENDER_ASYNC))
		return lucid_wait(r);
	return LUCID_OK;
}

int lucid_read_info(lucid_renderer *r, uint32_t *dst, size_t num_words) {
	if(!r || !dst)
		return LUCID_E_INVALID;
	if(!r->info_valid)
		return fail(r, LUCID_E_STATE, "lucid_read_info: no frame has copied its info back");
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	memcpy(dst, r->h_info, std::min(num_words, r->info_words) * 4);
	return LUCID_OK;
}

int lucid_stage_times_at(lucid_renderer *r, int32_t frames_back, float ms[8]) {
	if(!r || !ms)
		return LUCID_E_INVALID;
	if(frames_back < 0 || frames_back >= lucid_renderer::TIMING_RING || frames_back >= r->frame_counter)
		return fail(r, LUCID_E_STATE, "lucid_stage_times: no such frame in the timing history");
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	const int ring = (int)((r->frame_counter - 1 - frames_back) % lucid_renderer::TIMING_RING);
	cudaEvent_t *ev = r->ev[ring];
	for(int i = 0; i < 7; i++) {
		ms[i] = 0.0f; // frames rendered with LUCID_RENDER_NO_STAGE_TIMES only have the frame time
		if(r->ev_staged[ring])
			CU(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
	}
	CU(cudaEventElapsedTime(&ms[7], ev[0], ev[7]));
	return LUCID_OK;
}

int lucid_stage_times(lucid_renderer *r, float ms[8]) { return lucid_stage_times_at(r, 0, ms); }

static int slotRange(lucid_renderer *r, int which, int count, size_t &first_slot) {
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	if(which < 0 || which > 1 || count < 0 || count > r->p.max_visible_quads)
		return fail(r, LUCID_E_INVALID, "bad slot range");
	first_slot = which == 0 ? 0 : (size_t)r->p.max_visible_quads - count;
	return LUCID_OK;
}

int lucid_read_quad_aabbs(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t count) {
	if(!r || !dst)
		return LUCID_E_INVALID;
	size_t first;
	int rc = slotRange(r, which, count, first);
	if(rc)
		return rc;
	std::vector<u32> tmp(count);
	CU(cudaMemcpy(tmp.data(), r->p.quad_aabbs + first, (size_t)count * 4, cudaMemcpyDeviceToHost));
	for(int i = 0; i < count; i++)
		dst[i] = which == 0 ? tmp[i] : tmp[count - 1 - i];
	return LUCID_OK;
}

int lucid_read_tri_records(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t num_quads) {
	if(!r || !dst)
		return LUCID_E_INVALID;
	size_t first;
	int rc = slotRange(r, which, num_quads, first);
	if(rc)
		return rc;
	size_t nt = (size_t)num_quads * 2;
	std::vector<TriScan> scan(nt);
	std::vector<TriShade> shade(nt);
	std::vector<u32> aabbs(num_quads);
	CU(cudaMemcpy(scan.data(), r->p.tri_scan + first * 2, nt * sizeof(TriScan), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(shade.data(), r->p.tri_shade + first * 2, nt * sizeof(TriShade), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(aabbs.data(), r->p.quad_aabbs + first, (size_t)num_quads * 4, cudaMemcpyDeviceToHost));
	for(int q = 0; q < num_quads; q++) {
		int src_q = which == 0 ? q : num_quads - 1 - q;
		for(int s = 0; s < 2; s++) {
			u32 *o = dst + ((size_t)q * 2 + s) * 21;
			if((aabbs[src_q] >> (30 + s)) & 1) { // culled triangles have no record
				memset(o, 0, 21 * 4);
				continue;
			}
			const TriScan &sc = scan[(size_t)src_q * 2 + s];
			const TriShade &sh = shade[(size_t)src_q * 2 + s];
			memcpy(o + 0, &sh.bary0, 16), memcpy(o + 4, &sh.bary1, 16);
			memcpy(o + 8, &sc.s0, 16), memcpy(o + 12, &sc.s1, 16);
			memcpy(o + 16, &sh.depth, 16);
			o[20] = sh.misc.x;
		}
	}
	return LUCID_OK;
}

int lucid_read_quad_attrs(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t num_quads) {
	if(!r || !dst)
		return LUCID_E_INVALID;
	size_t first;
	int rc = slotRange(r, which, num_quads, first);
	if(rc)
		return rc;
	std::vector<uint4> col(num_quads), nrm(num_quads), uv((size_t)num_quads * 2);
	CU(cudaMemcpy(col.data(), r->p.quad_colors + first, (size_t)num_quads * 16, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(nrm.data(), r->p.quad_normals + first, (size_t)num_quads * 16, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(uv.data(), r->p.quad_uv + first * 2, (size_t)num_quads * 32, cudaMemcpyDeviceToHost));
	for(int q = 0; q < num_quads; q++) {
		int s = which == 0 ? q : num_quads - 1 - q;
		memcpy(dst + (size_t)q * 16 + 0, &col[s], 16);
		memcpy(dst + (size_t)q * 16 + 4, &nrm[s], 16);
		memcpy(dst + (size_t)q * 16 + 8, &uv[(size_t)s * 2], 32);
	}
	return LUCID_OK;
}

int lucid_read_bin_lists(lucid_renderer *r, uint32_t *bin_quads, size_t nq, uint32_t *bin_tris, size_t nt) {
	if(!r)
		return LUCID_E_INVALID;
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	if(nq > r->p.bin_list_capacity || nt > r->p.bin_list_capacity)
		return fail(r, LUCID_E_INVALID, "lucid_read_bin_lists: count above capacity");
	if(bin_quads && nq)
		CU(cudaMemcpy(bin_quads, r->p.bin_quads, nq * 4, cudaMemcpyDeviceToHost));
	if(bin_tris && nt)
		CU(cudaMemcpy(bin_tris, r->p.bin_tris, nt * 4, cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_read_frag_counts(lucid_renderer *r, uint32_t *dst) {
	if(!r || !dst)
		return LUCID_E_INVALID;
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	CU(cudaMemcpy(dst, r->frag_counts, (size_t)r->p.width * r->p.height * 4, cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_read_image(lucid_renderer *r, void *dst, size_t pitch_bytes) {
	if(!r || !dst || pitch_bytes < (size_t)r->p.width * 4)
		return LUCID_E_INVALID;
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	if(!r->image)
		return fail(r, LUCID_E_STATE, "lucid_read_image: the last frame was rendered into the caller's device image");
	CU(cudaMemcpy2D(dst, pitch_bytes, r->image, (size_t)r->p.width * 4, (size_t)r->p.width * 4,
					r->p.height, cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_read_debug_records(lucid_renderer *r, uint32_t *dst, int32_t max_records, int32_t *num_records) {
	if(!r || !num_records || max_records < 0 || (max_records > 0 && !dst))
		return LUCID_E_INVALID;
	if(!r->p.debug_records)
		return fail(r, LUCID_E_STATE, "lucid_read_debug_records: the renderer was created without LUCID_OPT_DEBUG_RASTER");
	int rc = lucid_wait(r);
	if(rc)
		return rc;
	u32 n = 0;
	CU(cudaMemcpy(&n, r->p.debug_records + 1, 4, cudaMemcpyDeviceToHost));
	*num_records = (int32_t)std::min<u32>(n, 0x7fffffffu);
	const size_t stored = std::min(std::min((size_t)n, (size_t)max_records), (size_t)LUCID_DEBUG_MAX_RECORDS);
	if(stored)
		CU(cudaMemcpy(dst, r->p.debug_records + 2, stored * LUCID_DEBUG_RECORD_WORDS * 4, cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_compare_render(lucid_renderer *r, int32_t mode, const LucidConfig *config, void *out_rgba8, size_t pitch_bytes,
						 float *kernel_ms) {
	if(!r)
		return LUCID_E_INVALID;
	if(!config || !out_rgba8 || mode < 0 || mode >= LUCID_COMPARE_MODE_COUNT || pitch_bytes < (size_t)r->p.width * 4)
		return fail(r, LUCID_E_INVALID, "lucid_compare_render: bad argument");
	if(r->frame_counter == 0)
		return fail(r, LUCID_E_STATE, "lucid_compare_render: no frame has been rendered (the comparators re-reduce the last frame's samples)");
	if(r->p.opts & LUCID_OPT_OPAQUE_PREPASS)
		return fail(r, LUCID_E_STATE, "lucid_compare_render: LUCID_OPT_OPAQUE_PREPASS drops samples before they reach the entry stream");
	if(mode != LUCID_COMPARE_HW_BLEND && (r->p.opts & LUCID_OPT_ADDITIVE_BLENDING))
		return fail(r, LUCID_E_INVALID, "lucid_compare_render: the approximate-OIT comparators are defined for the normal blend only");
	int rc = lucid_wait(r); // a frame that overflowed its storage has no complete entry stream either
	if(rc)
		return rc;
	CU(cudaSetDevice(r->ci.device));
	if(!r->cmp_image) { // the image is allocated last: a partly failed attempt is repeated as a whole (the rest is freed with the handle)
		CU(devAlloc(r, &r->cmp_order, (size_t)r->p.max_visible_quads));
		CU(devAlloc(r, &r->cmp_scratch, compareScratchKeys(r->num_sms)));
		CU(devAlloc(r, &r->cmp_ticket, 4));
		CU(devAlloc(r, &r->cmp_image, (size_t)r->p.width * r->p.height));
	}
	cudaEvent_t t0 = nullptr, t1 = nullptr;
	CU(cudaEventCreate(&t0));
	if(cudaError_t ee = cudaEventCreate(&t1); ee != cudaSuccess) {
		cudaEventDestroy(t0);
		return failCuda(r, ee, "lucid_compare_render: cudaEventCreate");
	}
	cudaEventRecord(t0, r->stream);
	launchCompare(r->p, *config, mode, r->cmp_order, r->cmp_scratch, r->cmp_ticket, r->cmp_image, r->stream, r->num_sms);
	cudaEventRecord(t1, r->stream);
	cudaError_t e = cudaStreamSynchronize(r->stream);
	if(e == cudaSuccess)
		e = cudaGetLastError();
	float ms = 0.0f;
	if(e == cudaSuccess)
		e = cudaEventElapsedTime(&ms, t0, t1);
	cudaEventDestroy(t0);
	cudaEventDestroy(t1);
	if(e != cudaSuccess)
		return failCuda(r, e, "lucid_compare_render");
	if(kernel_ms)
		*kernel_ms = ms;
	CU(cudaMemcpy2D(out_rgba8, pitch_bytes, r->cmp_image, (size_t)r->p.width * 4, (size_t)r->p.width * 4, r->p.height,
					cudaMemcpyDeviceToHost));
	return LUCID_OK;
}

int lucid_composite_to(lucid_renderer *r, void *dst_rgba8_device, size_t pitch_bytes) {
	if(!r || !dst_rgba8_device || pitch_bytes < (size_t)r->p.width * 4 || (pitch_bytes & 3))
		return LUCID_E_INVALID;
	if(r->frame_counter == 0 || !r->image)
		return fail(r, LUCID_E_STATE, "lucid_composite_to: no frame has been rendered into the renderer's own image");
	CU(cudaSetDevice(r->ci.device));
	Params p = r->p;
	p.image = r->image, p.image_pitch = p.width;
	launchCompositeBins(p, (u32 *)dst_rgba8_device, (int)(pitch_bytes / 4), r->stream, r->num_sms);
	CU(cudaGetLastError());
	r->pending = true;
	return LUCID_OK;
}

int lucid_image_pointer(lucid_renderer *r, void **device_ptr, size_t *pitch_bytes) {
	if(!r || !device_ptr || !pitch_bytes)
		return LUCID_E_INVALID;
	// always image 0: the one frames that stay on the device (LUCID_MEM_NONE) render into, whatever the
	// previous frame's output was
	*device_ptr = r->images[0];
	*pitch_bytes = (size_t)r->p.width * 4;
	return LUCID_OK;
}

int lucid_ipc_export_image(lucid_renderer *r, void *handle64) {
	if(!r || !handle64)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, r->images[0])); // the image LUCID_MEM_NONE frames render into
	static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
	memcpy(handle64, &h, 64);
	return LUCID_OK;
}

int lucid_sync_pointer(lucid_renderer *r, uint32_t **device_flags) {
	if(!r || !device_flags)
		return LUCID_E_INVALID;
	*device_flags = r->d_sync;
	return LUCID_OK;
}

int lucid_ipc_export_sync(lucid_renderer *r, void *handle64) {
	if(!r || !handle64)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, r->d_sync));
	memcpy(handle64, &h, 64);
	return LUCID_OK;
}

int lucid_signal(lucid_renderer *r, uint32_t *flag, uint32_t value) {
	if(!r || !flag)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	launchSignal(flag, value, r->stream);
	CU(cudaGetLastError());
	r->pending = true;
	return LUCID_OK;
}

int lucid_wait_flags(lucid_renderer *r, const uint32_t *flags, int32_t count, uint32_t value) {
	if(!r || !flags || count < 0 || count > LUCID_SYNC_FLAGS)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	launchWaitFlags(flags, count, value, r->h_status, 5000000000ull, r->stream);
	CU(cudaGetLastError());
	r->pending = true;
	return LUCID_OK;
}

int lucid_set_frame_gate(lucid_renderer *r, const uint32_t *flag, uint32_t value) {
	if(!r)
		return LUCID_E_INVALID;
	r->gate_flag = flag, r->gate_value = value;
	return LUCID_OK;
}

int lucid_ipc_open_image(lucid_renderer *r, const void *handle64, void **device_ptr) {
	if(!r || !handle64 || !device_ptr)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	CU(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return LUCID_OK;
}

int lucid_ipc_close_image(lucid_renderer *r, void *device_ptr) {
	if(!r || !device_ptr)
		return LUCID_E_INVALID;
	CU(cudaSetDevice(r->ci.device));
	CU(cudaIpcCloseMemHandle(device_ptr));
	return LUCID_OK;
}
}
