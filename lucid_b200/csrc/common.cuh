// common.cuh -- parameter block, storage layout and the floating-point contract of the sm_100a
// kernels.  Everything is compiled with -fmad=false: each binary32 operation is rounded on its
// own, in the association written in the source, so the kernels and the CPU checker agree bit for
// bit on coverage, bin lists and fragment counts (DESIGN.md "Floating-point contract").
#pragma once

#include "../../include/lucid_abi.h"
#include "../../include/lucid_colour_tables.h"

#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>

namespace lucid {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr int BIN_SIZE = LUCID_BIN_SIZE;
constexpr int BIN_SHIFT = LUCID_BIN_SHIFT;

// ---- storage layout in HBM -----------------------------------------------------------------
// Visible quads occupy slots [0, n_small) (small: bin AABB area <= 4) and (MVQ - n_large, MVQ)
// (large, allocated downwards), as in quad_setup.glsl:437-444; tri_idx = slot * 2 + second_tri.
//   quad_aabbs[slot]   u32      28-bit bin AABB + 2 cull bits (quad_setup.glsl:241-242)
//   tri_scan[tri]      2x16 B   scanline record (scan.xyz, ymin|ymax<<16) (step.xyz, sign bits)
//   tri_shade[tri]     4x16 B   depth plane (xyz, flags|instance<<16), (flat normal 10-10-10,
//                               instance RGBA8, constant shaded RGBA8, 1 if the constant is valid),
//                               bary edge 0, bary edge 1 -- the first 32-byte sector is all the
//                               block sort and constant-colour triangles ever read
//   quad_colors/normals[slot] 16 B, quad_uv[slot] 2x16 B   optional vertex attributes
// The reference keeps the same fields in uvec4_storage / normals_storage
// (definitions.glsl:132-137); here the per-sample fields of one triangle share one 64-byte line
// and the instance colour is folded in, which removes a dependent load from the shading loop.
struct TriScan {
	uint4 s0, s1;
};
struct TriShade {
	uint4 depth, misc, bary0, bary1;
};

struct Params {
	int width, height;
	int bin_count_x, bin_count_y, bin_count;
	int max_visible_quads;
	// multi-GPU split: this device owns the bins [bin_begin, bin_end) in row-major order (whole bin rows
	// in the plain bin-row split; a row may be shared when rows are too coarse to balance);
	// [row_begin, row_end) are the bin rows that range touches
	int row_begin, row_end, bin_begin, bin_end;
	u32 opts;
	int max_dispatches;
	int num_instances;
	int num_setup_ctas;
	int num_verts; // vertices of the caller's geometry: indices at or above it reject the quad

	// caller geometry (device pointers)
	const float *positions;
	const float4 *positions4; // 16-byte padded copy of positions (renderer-owned)
	const uint4 *quad_indices;
	const u32 *vertex_colors;
	const float2 *vertex_uvs;
	const u32 *vertex_normals;
	// per-frame instance data
	const LucidInstanceData *instances;
	const u32 *inst_colors;
	const float4 *inst_uv_rects;
	const float4 *inst_boxes; // per instance (min, max) of its vertices, or null: instance culling (LUCID_RENDER_CULL_INSTANCES)
	u32 *active_instances;	  // with inst_boxes: [0] number of instances k_quad_cull works on, [1..] their ids in input order

	// renderer-owned storage
	u32 *quad_aabbs;
	uint4 *quad_verts;		// per visible slot: the quad's vertex indices (k_quad_cull -> k_tri_setup)
	uint4 *quad_setup_info; // per visible slot: y range of both triangles, instance id
	TriScan *tri_scan;
	TriShade *tri_shade;
	uint4 *quad_colors;
	uint4 *quad_normals;
	uint4 *quad_uv;
	u32 *bin_quads;
	u32 *bin_tris;
	u32 bin_list_capacity;
	LucidInfo *info;
	int *counts; // 10 * bin_count, directly after info in the same allocation
	u64 *setup_lookback;
	u32 *setup_ticket;
	u32 *sort_scratch;

	// raster state
	u32 *image;			  // RGBA8, pitch in pixels
	int image_pitch;
	u32 *frag_counts;	  // optional per-pixel fragment counts (debug / parity), may be null
	u32 *bin_flags;		  // per bin: bit 0 promoted LOW->HIGH, bit 1 over the reference's HIGH limits (red)
	u32 *work_counters;	  // WC_*: bins / sort items / shade items taken, items per size class, stream entries handed out
	u32 *host_status;	  // pinned host word of this frame: set non-zero when the bin lists overflowed (the frame is red)
	u64 *bin_cost;		  // per bin: warp cycles the raster kernels spent on it this frame (split balancing)
	uint4 *block_lists;	  // per bin BIN_LIST_BYTES: 32 half-block lists (HIGH) or 16 block lists (LOW); compact_lists: a pool
	u32 *list_offsets;	  // compact_lists: per bin 32 list starts in the pool, in 8-byte units
	u32 list_pool_units;  // compact_lists: capacity of the pool in 8-byte units
	bool compact_lists;	  // LUCID_CREATE_COMPACT_LISTS
	int *block_counts;	  // 32 per bin: entries of each list
	uint4 *block_items;	  // work items of the block stages (item, entries, stream offset, -): one region of block_items_cap per size class
	u32 block_items_cap;
	u32 *large_keys;	  // per k_block_sort warp: sort keys of lists too long for shared memory
	// long runs of equal depth keys that k_block_sort left in arrival order for k_tie_runs: [0] count, then
	// (first stream entry, length) pairs; tie_scratch: per k_tie_runs warp room for one run's entries (2 x 4096 uint4)
	u32 *tie_runs;
	uint4 *tie_scratch;
	// the sorted-entry stream k_block_sort writes and k_block_shade reads: per work item a slice of both planes
	uint4 *sorted_rec;	  // (triangle, pixel mask of the upper / only half, pixel mask of the lower half, -)
	uint4 *sorted_aux;	  // (depth plane xyz, constant colour or AUX_VARYING)
	u32 stream_capacity;  // entries
	// LUCID_OPT_OPAQUE_PREPASS: per pixel the depth of the nearest INST_IS_OPAQUE sample (-inf: none), tiled like the
	// work items -- [bin][half-block 0..31 = (4-row group, 8-pixel column)][pixel 0..31]; k_block_sort writes,
	// k_block_shade reads; null without the option
	float *opaque_depth;
	// LUCID_OPT_DEBUG_RASTER: [0] capacity in records, [1] records written (may run past the capacity), then records of
	// LUCID_DEBUG_RECORD_WORDS words (the reference's `_debug` shader variants write fwk's ShaderDebugRecord)
	u32 *debug_records;
	bool debug_inject; // test hook of the debug variant (LUCID_DEBUG_RASTER_INJECT=1 at lucid_create)
	// textures: the two atlases as CUDA mipmapped arrays behind texture objects (RGBA8 unorm, normalised coordinates,
	// wrap, linear + mip-linear); 0 = no texture in the slot
	int tex_width[2], tex_height[2], tex_levels[2];
	unsigned long long tex_object[2];
};

// ---- floating-point contract ------------------------------------------------------------------
__device__ __forceinline__ int f2i(float x) { return __float2int_rz(x); }	  // saturating, NaN -> 0
__device__ __forceinline__ u32 f2u(float x) { return __float2uint_rz(x); } // saturating, NaN -> 0
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
	return fminf(fmaxf(x, lo), hi);
}
__device__ __forceinline__ float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
// 1/x: __frcp_rn is correctly rounded, i.e. the same binary32 value as the IEEE quotient 1.0f / x
__device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float rsqrt_rn(float x) { return __frcp_rn(__fsqrt_rn(x)); }
// small unsigned integer (< 2^23) to float without the conversion pipe: OR it into the mantissa of
// 2^23 and subtract 2^23 -- exact, so identical to float(v)
__device__ __forceinline__ float smallUintToFloat(u32 v, float bias = 0.0f) {
	return __uint_as_float(v | 0x4b000000u) - (8388608.0f + bias);
}

struct F3 {
	float x, y, z;
};
__device__ __forceinline__ F3 mk3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator-(F3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ bool same3(F3 a, F3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross3(F3 a, F3 b) {
	return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ F3 xyz(const LucidVec4 &v) { return mk3(v.x, v.y, v.z); }

__device__ __forceinline__ u32 encodeNormalUint(F3 n) {
	u32 x = f2u(512.0f + n.x * 511.0f) & 0x3ffu;
	u32 y = f2u(512.0f + n.y * 511.0f) & 0x3ffu;
	u32 z = f2u(512.0f + n.z * 511.0f) & 0x3ffu;
	return x | (y << 10) | (z << 20);
}
__device__ __forceinline__ F3 decodeNormalUint(u32 n) {
	const float s = 1.0f / 511.0f;
	return mk3(smallUintToFloat((n >> 0) & 0x3ffu, 512.0f) * s, smallUintToFloat((n >> 10) & 0x3ffu, 512.0f) * s,
			   smallUintToFloat((n >> 20) & 0x3ffu, 512.0f) * s);
}
__device__ __forceinline__ float4 decodeRGBA8(u32 c) {
	const float s = 1.0f / 255.0f, magic = 8388608.0f;
	return make_float4((__uint_as_float(__byte_perm(c, 0x4b000000u, 0x7650)) - magic) * s,
					   (__uint_as_float(__byte_perm(c, 0x4b000000u, 0x7651)) - magic) * s,
					   (__uint_as_float(__byte_perm(c, 0x4b000000u, 0x7652)) - magic) * s,
					   (__uint_as_float(__byte_perm(c, 0x4b000000u, 0x7653)) - magic) * s);
}
__device__ __forceinline__ u32 encodeRGBA8(float4 c) {
	return f2u(c.x * 255.0f) | (f2u(c.y * 255.0f) << 8) | (f2u(c.z * 255.0f) << 16) |
		   (f2u(c.w * 255.0f) << 24);
}
// ---- colour contract ------------------------------------------------------
// Coverage, depth keys and sample depths are evaluated one rounding per operation (the translation
// units are compiled with -fmad=false).  Colour has a 1/255 budget: its multiply-adds are fused
// (explicit __fmaf_rn, mirrored by fmaf in oracle/lucid_oracle.cpp shadeSampleFast) and the two pow()
// of finalShading (funcs.glsl:261-271) are interpolations in two tables of (value, slope) pairs,
// include/lucid_colour_tables.h -- the same words the CPU checker reads.
static __device__ const u32 d_s2l_words[LUCID_S2L_SIZE * 2] = {LUCID_S2L_WORDS};
static __device__ const u32 d_l2s_words[LUCID_L2S_SIZE * 2] = {LUCID_L2S_WORDS};
struct ColourTables {
	const float2 *s2l, *l2s; // global (L1-cached) or a shared-memory copy
};
__device__ __forceinline__ ColourTables globalColourTables() {
	ColourTables t;
	t.s2l = reinterpret_cast<const float2 *>(d_s2l_words), t.l2s = reinterpret_cast<const float2 *>(d_l2s_words);
	return t;
}
// sRGB decoding of c (256 intervals over c * 255: decoded bytes hit the nodes)
__device__ __forceinline__ float tabSRGBToLinear(const ColourTables &tab, float c) {
	const float t = fminf(fmaxf(c * 255.0f, 0.0f), 255.0f);
	const float fi = truncf(t);
	const float2 e = tab.s2l[f2i(fi)];
	return __fmaf_rn(t - fi, e.y, e.x);
}
// sRGB encoding of x <= 1: linear branch below 0.0031308, else 32 intervals per octave
__device__ __forceinline__ float tabLinearToSRGB(const ColourTables &tab, float x) {
	const u32 bits = __float_as_uint(x);
	const u32 i = min((bits - LUCID_L2S_FIRST_BITS) >> LUCID_L2S_SHIFT, (u32)(LUCID_L2S_SIZE - 1)); // clamp: unused below 2^-9
	const float x0 = __uint_as_float(bits & ~((1u << LUCID_L2S_SHIFT) - 1u));
	const float2 e = tab.l2s[i];
	const float hi = __fmaf_rn(x - x0, e.y, e.x);
	return x < 0.0031308f ? 12.92f * x : hi;
}
__device__ __forceinline__ float finalShadeFast(const ColourTables &tab, float c, float light) {
	const float x = fminf(tabSRGBToLinear(tab, c) * light, 1.0f);
	return saturatef(tabLinearToSRGB(tab, x));
}
// per-frame light terms of finalShading: ambient and sun colour times their powers
struct LightTerms {
	float amb[3], sun[3], msun[3];
};
__device__ __forceinline__ LightTerms lightTerms(const LucidLighting &L) {
	LightTerms t;
	t.amb[0] = L.ambient_color.x * L.ambient_power, t.amb[1] = L.ambient_color.y * L.ambient_power;
	t.amb[2] = L.ambient_color.z * L.ambient_power;
	t.sun[0] = L.sun_color.x * L.sun_power, t.sun[1] = L.sun_color.y * L.sun_power, t.sun[2] = L.sun_color.z * L.sun_power;
	t.msun[0] = -L.sun_dir.x, t.msun[1] = -L.sun_dir.y, t.msun[2] = -L.sun_dir.z;
	return t;
}
// Lambert term of shadeSample (shading.glsl:181-183) + finalShading, then the RGBA8 sample
__device__ __forceinline__ u32 shadeFinal(const ColourTables &tab, const LightTerms &lt, float4 color, F3 normal) {
	const float ndl = __fmaf_rn(lt.msun[0], normal.x, __fmaf_rn(lt.msun[1], normal.y, lt.msun[2] * normal.z));
	const float light_value = fmaxf(0.0f, __fmaf_rn(ndl, 0.7f, 0.3f));
	color.x = finalShadeFast(tab, color.x, __fmaf_rn(lt.sun[0], light_value, lt.amb[0]));
	color.y = finalShadeFast(tab, color.y, __fmaf_rn(lt.sun[1], light_value, lt.amb[1]));
	color.z = finalShadeFast(tab, color.z, __fmaf_rn(lt.sun[2], light_value, lt.amb[2]));
	return encodeRGBA8(color);
}
// A triangle without vertex colours, vertex normals or a texture shades to the same RGBA8 value
// at every sample (instance colour x flat-normal lighting), so setup evaluates shadeSample's
// colour once per triangle and the raster loop only computes depth for it.
constexpr u32 INST_VARYING_MASK =
	LUCID_INST_HAS_VERTEX_COLORS | LUCID_INST_HAS_VERTEX_NORMALS | LUCID_INST_HAS_ALBEDO_TEXTURE;
__device__ __forceinline__ u32 shadeConstant(const LucidLighting &L, u32 flags, u32 inst_color, u32 enc_normal) {
	float4 color = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
	if(flags & LUCID_INST_HAS_COLOR)
		color = decodeRGBA8(inst_color);
	if(color.w == 0.0f)
		return 0;
	return shadeFinal(globalColourTables(), lightTerms(L), color, decodeNormalUint(enc_normal));
}

__device__ __forceinline__ u32 laneId() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 laneMaskLt() {
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

// ---- LUCID_OPT_TIMERS: the reference's `_timers` shader variants (shared/timers.glsl:9-39) -------------------------
// A warp's leader adds the clock ticks (>> 4, like UPDATE_TIMER) it spent since the previous mark to one of the eight
// slots of LucidInfo.setup_timers / bin_dispatcher_timers / raster_timers; getStats() turns them into the same
// percentage rows (src/lucid_renderer.cpp:582-596,754-762).  Off: one predicate test per mark.
struct PhaseTimer {
	long long t0;
	bool on;
};
__device__ __forceinline__ PhaseTimer timerStart(const Params &p) {
	PhaseTimer t;
	t.on = (p.opts & LUCID_OPT_TIMERS) != 0;
	t.t0 = t.on ? clock64() : 0;
	return t;
}
__device__ __forceinline__ void timerMark(PhaseTimer &t, u32 *slots, int idx) {
	if(t.on) {
		const long long now = clock64();
		if((threadIdx.x & 31) == 0)
			atomicAdd(slots + idx, (u32)((unsigned long long)(now - t.t0) >> 4));
		t.t0 = now;
	}
}

// DEBUG_RECORD of the reference's `_debug` shader variants (libfwk shader_debug: line, thread, work group, four
// values), here: check id, lane, work item, four values.  Only called under LUCID_OPT_DEBUG_RASTER.
__device__ __forceinline__ void debugRecord(const Params &p, u32 check_id, u32 item, u32 v0, u32 v1, u32 v2, u32 v3) {
	if(!p.debug_records)
		return;
	const u32 n = atomicAdd(p.debug_records + 1, 1u);
	if(n >= p.debug_records[0])
		return;
	u32 *dst = p.debug_records + 2 + (size_t)n * LUCID_DEBUG_RECORD_WORDS;
	dst[0] = check_id, dst[1] = threadIdx.x, dst[2] = item, dst[3] = v0, dst[4] = v1, dst[5] = v2, dst[6] = v3;
}

// ---- programmatic dependent launch -----------------------------------------------------------
// The frame is a chain of short kernels (tens of microseconds at 1080p), so the gap between two
// launches matters.  Every kernel is launched with programmatic stream serialisation and starts
// with pdlEntry(): it lets its own successor be scheduled as soon as all of this grid's CTAs are
// running, then waits until the predecessor grid has completed and its writes are visible.  The
// launch latency and CTA ramp-up of kernel n+1 overlap the tail of kernel n.
__device__ __forceinline__ void pdlEntry() {
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}
// bins of row `by` owned by this device, as an inclusive column range clipped to [bmin, bmax]
__device__ __forceinline__ void clipToOwnedBins(const Params &p, int by, int &bmin, int &bmax) {
	bmin = max(bmin, p.bin_begin - by * p.bin_count_x);
	bmax = min(bmax, p.bin_end - 1 - by * p.bin_count_x);
}
__device__ __forceinline__ bool ownsBin(const Params &p, int bin) { return bin >= p.bin_begin && bin < p.bin_end; }
bool pdlEnabled(); // capi.cu: off with LUCID_NO_PDL=1 (A/B timing)
// first failed launch of the calling thread's current frame (capi.cu reads and clears it): a bad launch
// configuration is reported with the kernel's name instead of surfacing at some later call
void noteLaunchFailure(const char *kernel_name, cudaError_t err);
#define launchPDL(kernel, ...) launchPDLNamed(#kernel, kernel, __VA_ARGS__)
template <typename... KArgs, typename... Args>
inline void launchPDLNamed(const char *name, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream,
						   Args &&...args) {
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3((unsigned)block);
	cfg.dynamicSmemBytes = smem, cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr, cfg.numAttrs = pdlEnabled() ? 1 : 0;
	const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
	if(err != cudaSuccess)
		noteLaunchFailure(name, err);
}
// per-device one-time function attributes (handles on different devices / threads may launch concurrently)
template <typename F> inline void oncePerDevice(std::once_flag (&flags)[64], F &&f) {
	int dev = 0;
	cudaGetDevice(&dev);
	std::call_once(flags[dev & 63], f);
}

// kernel launchers (each in its own translation unit)
void launchFrameBegin(const Params &p, cudaStream_t stream);
void launchInfoOut(const Params &p, u32 *host_info, int num_words, cudaStream_t stream);
void launchQuadSetup(const Params &p, const LucidConfig &cfg, cudaStream_t stream);
void launchInstanceBoxes(const Params &p, float4 *boxes, cudaStream_t stream);
void launchPadPositions(const float *positions, float4 *positions4, int num_verts, cudaStream_t stream);
void launchBinning(const Params &p, cudaStream_t stream, cudaEvent_t *stage_events);
void launchRaster(const Params &p, const LucidConfig &cfg, cudaStream_t stream,
				  cudaEvent_t *stage_events, int num_sms);
void launchSignal(u32 *flag, u32 value, cudaStream_t stream);
void launchWaitFlags(const u32 *flags, int count, u32 value, u32 *status, unsigned long long timeout_ns, cudaStream_t stream);
void launchCompositeBins(const Params &p, u32 *dst, int dst_pitch, cudaStream_t stream, int num_sms);
size_t rasterLargeKeysCount(int num_sms);
size_t rasterTieScratchCount(int num_sms); // uint4 words of k_tie_runs' scratch
constexpr int TIE_RUN_QUEUE = 65536;	  // long runs of depth ties one frame can hand to k_tie_runs
constexpr int WORK_COUNTERS = 12; // raster_common.cuh WC_*

// 32 half-block lists of up to 4096 8-byte records (raster_high.glsl:27); a LOW bin uses the first
// 16 x 256 16-byte records
constexpr size_t BIN_LIST_BYTES = (size_t)32 * 4096 * 8;

} // namespace lucid
