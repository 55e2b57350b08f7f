/* lucid_b200.h -- C ABI of the B200 exact-OIT rasteriser (drop-in for LucidRenderer's hot path).
 *
 * Each entry point names the reference interface it stands in for (file:line under the
 * nadult/lucid tree).  Plain pointers and sizes only; the buffer layouts are in lucid_abi.h.
 * All calls on one handle must come from one host thread at a time (the reference is single
 * threaded, single queue); handles are independent, so one handle per GPU scales out.
 * Every function returns 0 on success or a negative LUCID_E_* code; lucid_last_error() gives text.
 */
#ifndef LUCID_B200_H
#define LUCID_B200_H

#include "lucid_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lucid_renderer lucid_renderer;

enum {
	LUCID_OK = 0,
	LUCID_E_INVALID = -1,  /* bad argument */
	LUCID_E_CUDA = -2,	   /* CUDA runtime error (no device, out of memory, launch failure) */
	LUCID_E_LIMIT = -3,	   /* a documented limit was exceeded (instances, bin coordinates) */
	LUCID_E_STATE = -4	   /* call sequence error (e.g. render before set_geometry) */
};

enum { LUCID_MEM_HOST = 0, LUCID_MEM_DEVICE = 1, LUCID_MEM_NONE = 2 };

enum {
	LUCID_RENDER_ASYNC = 1,		  /* return after enqueueing; pair with lucid_wait().  A host image is then
									 copied out on a second stream while the next frame renders, so its
									 buffer should be pinned and must not be reused before lucid_wait() */
	LUCID_RENDER_SKIP_INFO = 2,	  /* do not copy LucidInfo back this frame */
	LUCID_RENDER_FRAG_COUNTS = 4, /* also write the per-pixel fragment-count image (parity tests) */
	LUCID_RENDER_NO_STAGE_TIMES = 8, /* record only the frame's first and last timing event: without events
									 between them the kernels of a frame overlap their launches
									 (lucid_stage_times then reports the frame time only) */
	LUCID_RENDER_CULL_INSTANCES = 16 /* an instance whose bounding box lies outside one frustum plane or projects
									 outside the owned bin rows is dropped before its quads are loaded (the box
									 is computed once per instance list and cached; the kept instances are
									 compacted in input order by k_instance_select).  Conservative, so the
									 visible quads with their slots, the per-bin lists and the pixels are
									 unchanged; num_rejected_quads then only covers the processed instances */
	,
	LUCID_RENDER_OWNED_BINS_ONLY = 32 /* LUCID_MEM_HOST with a bin-row split: only the pixels of the owned bins are
									 copied to out_rgba8 (which has the layout of the whole image) -- every device
									 of a split delivers its own strip over its own PCIe link, e.g. into host
									 memory shared by the processes, instead of one device gathering the frame */
	,
	LUCID_RENDER_NO_DEPENDENT_LAUNCH = 64 /* launch the frame's kernels without programmatic dependent launch.  By
									 default a handle uses it while its frames take under a millisecond (the
									 launch of kernel n+1 overlaps the tail of kernel n) and not on longer frames,
									 where it costs more than it saves.  Callers that keep several handles busy on
									 one device (frames in flight) should pass this flag: the CTAs of a kernel
									 launched early wait on the SMs and take the place of the other handles'
									 kernels (10M-triangle 4K frame, two handles: 363 with, 424 frames/s without) */
};

/* LucidRenderer::exConstruct(device, compiler, opts, view_size), src/lucid_renderer.cpp:186-317.
 * Zero fields select the reference's defaults. */
typedef struct LucidCreateInfo {
	int32_t width, height;		  /* view_size; <= 4096 (7-bit bin coordinates, funcs.glsl:52-59) */
	uint32_t opts;				  /* LUCID_OPT_* (LucidRenderOpt flags) */
	int32_t max_visible_quads;	  /* 0 -> 4793490 (= min(2^30/224, VRAM_MB*1024), lucid_renderer.cpp:204) */
	int32_t max_dispatches;		  /* 0 -> 256 (lucid_renderer.cpp:203) */
	int32_t device;				  /* CUDA device ordinal */
	void *stream;				  /* cudaStream_t to run on; NULL -> the renderer creates one */
	int32_t bin_row_begin, bin_row_end; /* owned bin rows [begin,end) for the multi-GPU split; 0,0 -> all */
	uint32_t max_block_entries;	  /* capacity of the sorted-entry stream between the block sort and the shading
									 kernel: one 32-byte entry per (triangle, 8x4 half-block or 8x8 block) pair of a
									 frame; 0 -> max(16 * max_visible_quads, 2^22).  A frame that needs more paints the
									 bins that did not fit red and reports LUCID_E_LIMIT (the reference bounds the same
									 lists per work group: raster_low.glsl:22-32, raster_high.glsl:35-44) */
	uint32_t flags;				  /* LUCID_CREATE_* */
} LucidCreateInfo;

enum {
	/* Block lists (the per-bin lists between k_raster_bins and k_block_sort) in a pool sized by max_block_entries
	 * (16 bytes per entry) instead of a fixed 1 MiB slot per bin (8 GiB at 3840x2160): every bin is walked twice --
	 * a counting pass, one allocation of exactly its entries, a filling pass -- so the list stage takes about twice
	 * as long.  For handles that have to be small: several per device, devices with less memory. */
	LUCID_CREATE_COMPACT_LISTS = 1
};

int lucid_create(const LucidCreateInfo *info, lucid_renderer **out);
void lucid_destroy(lucid_renderer *r);
const char *lucid_last_error(const lucid_renderer *r); /* r may be NULL: error of the failed create */

/* RenderContext::verts / quads_ib (src/lucid_base.h:62-84): positions float[3*nv] tightly packed,
 * colors RGBA8, tex coords float[2*nv], normals 10-10-10 (scene.cpp:338-343), quad indices
 * u32[4*nq].  colors / uvs / normals may be NULL.  LUCID_MEM_HOST: copied now; LUCID_MEM_DEVICE:
 * borrowed, must stay valid while rendering (as the reference borrows the Scene's buffers). */
int lucid_set_geometry(lucid_renderer *r, const float *positions, int32_t num_verts,
					   const uint32_t *colors, const float *uvs, const uint32_t *normals,
					   const uint32_t *quad_indices, int32_t num_quads, int32_t memory);

/* opaque_tex / trans_tex (lucid_base.h:82, bound at lucid_renderer.cpp:550-552): slot 0 opaque,
 * slot 1 transparent; tightly packed RGBA8 mip chain, level l is max(1,w>>l) x max(1,h>>l).
 * Sampler: repeat, bilinear, mip-linear, no anisotropy (DESIGN.md "Texture filter"). */
int lucid_set_texture(lucid_renderer *r, int32_t slot, const uint8_t *rgba8_mips, int32_t width,
					  int32_t height, int32_t levels);

/* change the owned bin rows between frames (load balancing of the bin-row split) */
int lucid_set_bin_rows(lucid_renderer *r, int32_t begin, int32_t end);
/* per bin row, the warp cycles the raster kernels spent on it in the last frame: the weights from which
 * the caller picks the next frame's row boundaries ("choose boundaries from the previous
 * frame's per-row fragment counts"; measured cycles also cover list building and sorting) */
int lucid_read_row_costs(lucid_renderer *r, uint64_t *dst, int32_t num_rows);
/* finer than rows: own the bins [begin, end) in row-major order (a bin row may be shared between two
 * devices when single rows are too heavy to balance), and the per-bin costs to choose the ranges from.
 * Counts and lists of owned bins are identical to the full frame's; other bins stay empty. */
int lucid_set_bin_range(lucid_renderer *r, int32_t begin, int32_t end);
int lucid_read_bin_costs(lucid_renderer *r, uint64_t *dst, int32_t num_bins);

/* LucidRenderer::render(const Context&), src/lucid_renderer.cpp:319-350: config as filled by
 * setupInputData, instances / colours / uv rects as filled by uploadInstances (host pointers).
 * out_rgba8: RGBA8 image (the reference writes the swap-chain image), row pitch in bytes;
 * out_memory says where it lives; LUCID_MEM_NONE keeps the image in the renderer's own buffer.
 * A device pointer may be peer memory: the raster kernels then store across NVLink directly. */
int lucid_render(lucid_renderer *r, const LucidConfig *config, const LucidInstanceData *instances,
				 const uint32_t *instance_colors, const float *instance_uv_rects,
				 int32_t num_instances, void *out_rgba8, size_t pitch_bytes, int32_t out_memory,
				 uint32_t flags);
int lucid_wait(lucid_renderer *r);

/* m_last_info (lucid_renderer.cpp:341-346): LucidInfo followed by 10*bin_count ints */
int lucid_read_info(lucid_renderer *r, uint32_t *dst, size_t num_words);
int lucid_bin_count(const lucid_renderer *r);

/* per-stage GPU milliseconds of the last frame (PERF_GPU_SCOPE replacement, lucid_renderer.cpp:323,
 * 433,459,486,557,571): [0] quad setup [1] bin count [2] bin offsets + categories [3] bin dispatch
 * [4] raster: block lists of all LOW and HIGH bins (generate rows / generate blocks)
 * [5] raster: block sort + shading of all LOW and HIGH blocks [6] finish [7] whole frame.
 * LOW and HIGH bins run in the same two kernels, so their time is not reported separately.
 * lucid_stage_times_at: the frame `frames_back` frames ago (0 = last; 64 frames are kept). */
int lucid_stage_times(lucid_renderer *r, float ms[8]);
int lucid_stage_times_at(lucid_renderer *r, int32_t frames_back, float ms[8]);

/* ---- inspection of intermediate buffers (what tests compare with the CPU checker) ---- */
/* which: 0 small quads (slot i), 1 large quads (slot MVQ-1-i) */
int lucid_read_quad_aabbs(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t count);
/* 21 words per triangle in the reference's field order: bary0 bary1 scan0 scan1 depth normal */
int lucid_read_tri_records(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t num_quads);
/* 16 words per quad: colors normals uv0 uv1 */
int lucid_read_quad_attrs(lucid_renderer *r, int32_t which, uint32_t *dst, int32_t num_quads);
int lucid_read_bin_lists(lucid_renderer *r, uint32_t *bin_quads, size_t num_bin_quads,
						 uint32_t *bin_tris, size_t num_bin_tris);
int lucid_read_frag_counts(lucid_renderer *r, uint32_t *dst);
int lucid_read_image(lucid_renderer *r, void *dst_rgba8, size_t pitch_bytes);

/* ---- multi-GPU composite over NVLink: share one GPU's image with the other ranks ---- */
/* device pointer and pitch (bytes) of the renderer-owned image */
int lucid_image_pointer(lucid_renderer *r, void **device_ptr, size_t *pitch_bytes);
/* bin-row split, second way to composite: after a frame rendered into the renderer's own image, copy
 * the pixels of the owned bins to dst (device or peer pointer) as whole 128-byte bin rows, asynchronously
 * on the render stream.  Fewer, larger NVLink packets than the raster kernels' direct stores. */
int lucid_composite_to(lucid_renderer *r, void *dst_rgba8_device, size_t pitch_bytes);
/* 64-byte cudaIpcMemHandle_t of the renderer-owned image, to be sent to peer processes */
int lucid_ipc_export_image(lucid_renderer *r, void *handle64);
/* LUCID_OPT_DEBUG_RASTER: the records of the last frame (include/lucid_abi.h), up to max_records of them into dst
 * (LUCID_DEBUG_RECORD_WORDS words each); *num_records = how many the frame produced (may exceed what was stored).
 * Replaces printDebugData / shaderDebugDownloadResults, src/lucid_renderer.cpp:112-117. */
int lucid_read_debug_records(lucid_renderer *r, uint32_t *dst, int32_t max_records, int32_t *num_records);

/* Test hook: n samples (u, v, lod triples, host memory) of the texture in `slot`, fetched by the texture unit
 * exactly as the shading kernel fetches them (tex2DLod); out_rgba receives 4 floats per sample.  Pins the CPU
 * checker's restatement of the unit's filter arithmetic against the hardware (tests/test_gpu_parity.py). */
int lucid_debug_sample_texture(lucid_renderer *r, int32_t slot, const float *uvl, int32_t n, float *out_rgba);

/* ---- comparators (SimpleRenderer::render, src/simple_renderer.cpp:134-196, and the techniques of
 * docs/readme.md:7-8) ------------------------------------------------------------------------------------------
 * Re-reduces the SAMPLES of the frame this handle rendered last -- same coverage, same sample colours and depths --
 * with the per-pixel rule of a cheaper technique (LUCID_COMPARE_*, lucid_abi.h) and copies the RGBA8 image to host
 * memory: what hardware alpha blending in submission order, weighted blended OIT or 4-layer MLAB would have shown
 * where lucid_render shows the exact blend.  `config` is the frame's config (lighting, background).  Synchronous;
 * no lucid_render may be issued on the handle in between.  kernel_ms (may be NULL): device time of the comparator's
 * kernels -- the time of a CUDA restatement, not of raster-operation hardware.  LUCID_E_STATE without a frame or on
 * a LUCID_OPT_OPAQUE_PREPASS renderer; the approximate-OIT modes reject LUCID_OPT_ADDITIVE_BLENDING renderers. */
int lucid_compare_render(lucid_renderer *r, int32_t mode, const LucidConfig *config, void *out_rgba8, size_t pitch_bytes,
						 float *kernel_ms);

/* ---- frame hand-over of the bin-row split over NVLink, without a collective -------------------------------
 * Every renderer owns a block of LUCID_SYNC_FLAGS 32-bit flags in device memory (zero at creation).  The
 * gathering rank exports its block, the other ranks map it (lucid_ipc_open_image opens any handle exported by
 * this library) and, after rendering their strip of frame k into the shared image, store k into "their" flag:
 *     lucid_signal(r, peer_flags + rank, k);            // stream-ordered, system-scope release
 * The gathering rank waits for all of them on the device and then releases the image:
 *     lucid_wait_flags(r, flags + 1, world - 1, k);     // one warp, system-scope acquire, gives up after 5 s
 *     ... consume the image (read-back, present) ...
 *     lucid_signal(r, flags + LUCID_SYNC_RELEASED, k);
 * and a rank may only store into the shared image again when frame k - 1 has been released:
 *     lucid_set_frame_gate(r, peer_flags + LUCID_SYNC_RELEASED, k - 1);  // applies to the next lucid_render only:
 *                                                        // its raster kernels start after the flag reached the value
 * Setup and binning of the next frame overlap the wait.  A wait that times out makes the next lucid_wait()
 * return LUCID_E_STATE. */
#define LUCID_SYNC_FLAGS 64
#define LUCID_SYNC_RELEASED 32
int lucid_sync_pointer(lucid_renderer *r, uint32_t **device_flags);
int lucid_ipc_export_sync(lucid_renderer *r, void *handle64);
int lucid_signal(lucid_renderer *r, uint32_t *flag, uint32_t value);
int lucid_wait_flags(lucid_renderer *r, const uint32_t *flags, int32_t count, uint32_t value);
int lucid_set_frame_gate(lucid_renderer *r, const uint32_t *flag, uint32_t value);

/* maps a peer's image into this process; pass the result as out_rgba8 with LUCID_MEM_DEVICE */
int lucid_ipc_open_image(lucid_renderer *r, const void *handle64, void **device_ptr);
int lucid_ipc_close_image(lucid_renderer *r, void *device_ptr);

#ifdef __cplusplus
}
#endif
#endif
