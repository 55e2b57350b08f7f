/* lucid_abi.h -- buffer layouts shared between host, CUDA kernels and the CPU oracle.
 *
 * These are the byte layouts the reference shares between C++ and GLSL through
 * src/shader_structs.h:6-8 -> data/shaders/shared/structures.glsl:9-114 and
 * data/shaders/shared/definitions.glsl:42-137.  A caller that fills these structs the way
 * LucidRenderer::setupInputData / uploadInstances do (src/lucid_renderer.cpp:352-451) can hand
 * them to lucid_render() unchanged.  Only layouts are restated here; the static_asserts pin the
 * offsets measured from the reference headers.
 */
#ifndef LUCID_ABI_H
#define LUCID_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUCID_BIN_SIZE 32
#define LUCID_BIN_SHIFT 5
#define LUCID_BLOCK_SIZE 8
#define LUCID_MAX_INSTANCE_QUADS 1024 /* src/lucid_renderer.h:29 */
#define LUCID_MAX_INSTANCES (64 * 1024) /* src/lucid_renderer.h:28 */

#define LUCID_BIN_LEVELS_COUNT 5
#define LUCID_REJECTION_TYPE_COUNT 4
#define LUCID_TIMERS_COUNT 8
#define LUCID_STATS_COUNT 4
#define LUCID_INFO_MAX_DISPATCHES 256

/* bin density levels, definitions.glsl:79-83 */
enum {
	LUCID_BIN_LEVEL_EMPTY = 0,
	LUCID_BIN_LEVEL_MICRO = 1,
	LUCID_BIN_LEVEL_LOW = 2,
	LUCID_BIN_LEVEL_MEDIUM = 3,
	LUCID_BIN_LEVEL_HIGH = 4
};

/* instance flags == DrawCallOpts bits, definitions.glsl:85-94 / src/lucid_base.h:44-46 */
enum {
	LUCID_INST_HAS_VERTEX_COLORS = 0x001,
	LUCID_INST_HAS_VERTEX_TEX_COORDS = 0x002,
	LUCID_INST_HAS_VERTEX_NORMALS = 0x004,
	LUCID_INST_IS_OPAQUE = 0x008,
	LUCID_INST_TEX_OPAQUE = 0x010,
	LUCID_INST_HAS_UV_RECT = 0x020,
	LUCID_INST_HAS_ALBEDO_TEXTURE = 0x040,
	LUCID_INST_HAS_NORMAL_TEXTURE = 0x080,
	LUCID_INST_HAS_PBR_TEXTURE = 0x100,
	LUCID_INST_HAS_COLOR = 0x200
};

/* rejection reasons, definitions.glsl:96-100 */
enum {
	LUCID_REJECTION_OTHER = 0,
	LUCID_REJECTION_BACKFACE = 1,
	LUCID_REJECTION_FRUSTUM = 2,
	LUCID_REJECTION_BETWEEN_SAMPLES = 3
};

/* LucidRenderOpt bit positions, src/lucid_renderer.h:10-11 (EnumFlags: bit i = 1 << i) */
enum {
	LUCID_OPT_DEBUG_QUAD_SETUP = 1 << 0,
	LUCID_OPT_DEBUG_BIN_COUNTER = 1 << 1,
	LUCID_OPT_DEBUG_BIN_DISPATCHER = 1 << 2,
	LUCID_OPT_DEBUG_RASTER = 1 << 3,
	LUCID_OPT_TIMERS = 1 << 4,
	LUCID_OPT_ADDITIVE_BLENDING = 1 << 5,
	LUCID_OPT_VISUALIZE_ERRORS = 1 << 6,
	LUCID_OPT_ALPHA_THRESHOLD = 1 << 7,
	/* Extensions (not LucidRenderOpt bits of the reference).
	 * OPAQUE_PREPASS -- the TODO of shared/shading.glsl:31-32 as an option: at every pixel the nearest sample of an
	 * INST_IS_OPAQUE instance (the caller's promise that every sample of the instance has alpha 1, the reference's
	 * DrawCallOpt::is_opaque, src/scene.cpp:496) hides all samples behind it: they are dropped before sorting and
	 * shading.  The image does not change; per-pixel fragment counts and stats[0] count the surviving samples only, so
	 * it is outside the parity mode.  Ignored together with ADDITIVE_BLENDING or ALPHA_THRESHOLD. */
	LUCID_OPT_OPAQUE_PREPASS = 1 << 8
};

/* lucid_compare_render (lucid_b200.h): the per-pixel rule the last frame's samples are reduced with instead of the
 * exact front-to-back blend -- the renderers the reference is compared with (docs/readme.md:7-8) */
enum {
	LUCID_COMPARE_HW_BLEND = 0, /* SimpleRenderer (src/simple_renderer.cpp:69-132): opaque phase with depth write, then
								   alpha blending in submission order on an 8-bit target */
	LUCID_COMPARE_WBOIT = 1,	/* weighted blended OIT (McGuire & Bavoil 2013, weight of eq. 7) */
	LUCID_COMPARE_MLAB4 = 2,	/* multi-layer alpha blending with four layers (Salvi & Vaidyanathan 2014) */
	LUCID_COMPARE_MODE_COUNT = 3
};

/* LUCID_OPT_DEBUG_RASTER (the reference's raster_low_debug / raster_high_debug pipelines, src/lucid_renderer.cpp:
 * 147-158,263-296): the block stage checks what the `DEBUG_ENABLED` code of the shaders checks and writes a record per
 * violation (lucid_read_debug_records):
 *   LUCID_DEBUG_EMPTY_COVERAGE    a block-list entry without a covered pixel (raster_low.glsl:125-126,
 *                                 raster_high.glsl:184-185): values 0, 0, 0, 0 -- "bx_mask is invalid"
 *   LUCID_DEBUG_UNSORTED          sort keys not strictly increasing after the sort (raster_low.glsl:154-163,
 *                                 raster_high.glsl:196-203): values i, tri_count, prev_value, value
 * A record is {check id, thread of the CTA, work item (bin << 6 | HIGH << 5 | block), four values}; the reference's
 * records carry line id, local index and work-group index in those places. */
enum { LUCID_DEBUG_EMPTY_COVERAGE = 1, LUCID_DEBUG_UNSORTED = 2 };
#define LUCID_DEBUG_RECORD_WORDS 7
#define LUCID_DEBUG_MAX_RECORDS 65536

typedef struct LucidVec4 {
	float x, y, z, w;
} LucidVec4;

/* structures.glsl:9-14 */
typedef struct LucidInstanceData {
	int32_t index_offset;  /* first index of the instance in the quad index buffer (quad_offset * 4) */
	int32_t vertex_offset; /* added to every index; the reference always passes 0 */
	int32_t num_quads;	 /* <= LUCID_MAX_INSTANCE_QUADS */
	uint32_t flags;		   /* LUCID_INST_* */
} LucidInstanceData;

/* structures.glsl:16-21 (std140: padded to 64 bytes) */
typedef struct LucidLighting {
	LucidVec4 ambient_color;
	LucidVec4 sun_color;
	LucidVec4 sun_dir;
	float sun_power, ambient_power;
	float _pad[2];
} LucidLighting;

/* structures.glsl:24-28; all vectors in world space */
typedef struct LucidFrustum {
	LucidVec4 ws_origins[4], ws_dirs[4];
	LucidVec4 ws_origin0, ws_dir0;
	LucidVec4 ws_dirx, ws_diry;
} LucidFrustum;

/* structures.glsl:104-114 */
typedef struct LucidConfig {
	LucidFrustum frustum;
	LucidVec4 view_proj_matrix[4]; /* column major: view_proj_matrix[c] is column c */
	LucidLighting lighting;
	LucidVec4 background_color;
	uint32_t enable_backface_culling;
	int32_t num_instances;
	int32_t instance_packet_size;
	uint32_t _pad;
} LucidConfig;

/* structures.glsl:68-101.  The device copy is followed by 10 * bin_count ints (g_counts,
 * definitions.glsl:117-129): [0] quad counts [1] quad offsets [2] quad offsets temp (= ends)
 * [3..5] same for tris [6] micro list (unused) [7] LOW bin list [8] medium (unused) [9] HIGH list */
typedef struct LucidInfo {
	int32_t num_input_quads;
	int32_t num_visible_quads[2]; /* [0] small (bin area <= 4), [1] large */
	int32_t num_counted_quads[2];
	int32_t bin_level_counts[LUCID_BIN_LEVELS_COUNT];
	uint32_t a_small_bins, a_high_bins;
	uint32_t a_setup_work_groups;
	uint32_t a_dummy_counter;
	uint32_t num_binning_dispatches[3];
	uint32_t bin_level_dispatches[LUCID_BIN_LEVELS_COUNT][3];
	uint32_t num_rejected_quads[LUCID_REJECTION_TYPE_COUNT];
	uint32_t setup_timers[LUCID_TIMERS_COUNT];
	uint32_t raster_timers[LUCID_TIMERS_COUNT];
	uint32_t bin_dispatcher_timers[LUCID_TIMERS_COUNT];
	uint32_t stats[LUCID_STATS_COUNT]; /* [0] fragments [1] half-block-tris [2] invalid pixels */
	int32_t dispatcher_first_batch[2][LUCID_INFO_MAX_DISPATCHES];
	int32_t dispatcher_num_batches[2][LUCID_INFO_MAX_DISPATCHES];
	int32_t temp[64];
} LucidInfo;

#define LUCID_INFO_U32_SIZE (sizeof(LucidInfo) / 4)
#define LUCID_COUNTS_PER_BIN 10

/* indices of the per-bin arrays inside g_counts */
enum {
	LUCID_CNT_QUAD_COUNTS = 0,
	LUCID_CNT_QUAD_OFFSETS = 1,
	LUCID_CNT_QUAD_OFFSETS_TEMP = 2,
	LUCID_CNT_TRI_COUNTS = 3,
	LUCID_CNT_TRI_OFFSETS = 4,
	LUCID_CNT_TRI_OFFSETS_TEMP = 5,
	LUCID_CNT_MICRO_BINS = 6,
	LUCID_CNT_LOW_BINS = 7,
	LUCID_CNT_MEDIUM_BINS = 8,
	LUCID_CNT_HIGH_BINS = 9
};

#ifdef __cplusplus
}
static_assert(sizeof(LucidInstanceData) == 16, "InstanceData");
static_assert(sizeof(LucidLighting) == 64, "Lighting");
static_assert(sizeof(LucidFrustum) == 192, "Frustum");
static_assert(sizeof(LucidConfig) == 352, "LucidConfig");
static_assert(offsetof(LucidConfig, view_proj_matrix) == 192, "view_proj");
static_assert(offsetof(LucidConfig, lighting) == 256, "lighting");
static_assert(offsetof(LucidConfig, background_color) == 320, "background");
static_assert(offsetof(LucidConfig, enable_backface_culling) == 336, "backface");
static_assert(offsetof(LucidConfig, num_instances) == 340, "num_instances");
static_assert(offsetof(LucidConfig, instance_packet_size) == 344, "packet");
static_assert(sizeof(LucidInfo) == 4608, "LucidInfo");
static_assert(offsetof(LucidInfo, num_visible_quads) == 4, "num_visible_quads");
static_assert(offsetof(LucidInfo, num_counted_quads) == 12, "num_counted_quads");
static_assert(offsetof(LucidInfo, bin_level_counts) == 20, "bin_level_counts");
static_assert(offsetof(LucidInfo, a_small_bins) == 40, "a_small_bins");
static_assert(offsetof(LucidInfo, a_setup_work_groups) == 48, "a_setup_work_groups");
static_assert(offsetof(LucidInfo, num_binning_dispatches) == 56, "num_binning_dispatches");
static_assert(offsetof(LucidInfo, bin_level_dispatches) == 68, "bin_level_dispatches");
static_assert(offsetof(LucidInfo, num_rejected_quads) == 128, "num_rejected_quads");
static_assert(offsetof(LucidInfo, setup_timers) == 144, "setup_timers");
static_assert(offsetof(LucidInfo, raster_timers) == 176, "raster_timers");
static_assert(offsetof(LucidInfo, bin_dispatcher_timers) == 208, "bin_dispatcher_timers");
static_assert(offsetof(LucidInfo, stats) == 240, "stats");
static_assert(offsetof(LucidInfo, dispatcher_first_batch) == 256, "dispatcher_first_batch");
static_assert(offsetof(LucidInfo, dispatcher_num_batches) == 2304, "dispatcher_num_batches");
static_assert(offsetof(LucidInfo, temp) == 4352, "temp");
#endif

#endif
