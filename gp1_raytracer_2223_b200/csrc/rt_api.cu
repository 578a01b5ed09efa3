// C-ABI implementation (include/rt_b200.h): context, scene replication to every device,
// kernel launch, row-strip split over the context's GPUs with the gather fused into the
// kernel's stores (peer-mapped frame buffer on device 0), and the copy out to the host
// surface.  There is deliberately no host fallback: every entry point fails loudly when
// CUDA is not there.
#include "rt_kernel.cuh"
#include "rt_aux_kernels.cuh"
#include "rt_pick.h"
#include "rt_wave_params.h"
#include "rt_bvh_build.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace
{
	thread_local std::string g_create_error;

	struct HostMesh
	{
		std::vector<float4> triangles;   // 3 per triangle
		std::vector<float4> nodes;       // 2 per BVH node, threaded (rt::BvhLink); empty when the mesh came without nodes
		std::vector<int32_t> split;      // rt::wave::kSplitStride words: the tree cut into subtrees (RT_KERNEL_WAVEFRONT); empty: none
		std::vector<uint8_t> root_map;   // per node: subtree number + 1 for the subtrees' roots, else 0
		bool split_built = false;        // split / root_map belong to `nodes` (built when a launch first wants them: build_split)
		float aabb_min[3] = { 0, 0, 0 };
		float aabb_max[3] = { 0, 0, 0 };
		int32_t cull_mode = RT_CULL_BACK_FACE;
		int32_t material = 0;
		bool uploaded = false;
		// device-side UpdateTransforms (rt_upload_mesh_source / rt_transform_mesh)
		std::vector<float> src_positions, src_normals;
		std::vector<int32_t> src_indices;
		bool has_source = false, source_dirty = false, has_transform = false, transform_dirty = false;
		float transform[16] = {};
		// device-side BuildBVH (rt_set_mesh_device_bvh): every rt_transform_mesh is one UpdateTransforms call of the
		// reference, build included; the builds are history-dependent, so none may be skipped or repeated
		bool device_bvh = false;
		std::vector<std::array<float, 16>> pending_builds;
	};

	// Offsets (in floats) of the SoA arrays inside the per-device float arena.
	struct ArenaLayout
	{
		static constexpr int sphere = 0;                                   // ox, oy, oz, r
		static constexpr int plane = sphere + 4 * rt::kMaxSpheres;         // ox, oy, oz, nx, ny, nz
		static constexpr int light = plane + 6 * rt::kMaxPlanes;           // ox, oy, oz, r, g, b, intensity
		static constexpr int total = light + 7 * rt::kMaxLights;
	};

	// Byte offsets inside the per-device static block (and its pinned host mirror): one copy per upload.
	struct StaticBlock
	{
		static constexpr size_t arena = 0;
		static constexpr size_t materials = (arena + sizeof(float) * ArenaLayout::total + 15) & ~size_t(15);
		static constexpr size_t light_type = materials + sizeof(float4) * 2 * rt::kMaxMaterials;
		static constexpr size_t bytes = light_type + sizeof(int32_t) * rt::kMaxLights;
		static constexpr size_t total = (bytes + rt::kMaxSpheres + rt::kMaxPlanes + 15) & ~size_t(15);
	};

	constexpr unsigned int kQueueRing = 256;

	struct DeviceState
	{
		int device = -1;
		cudaStream_t stream = nullptr;
		uint8_t* d_static = nullptr;       // StaticBlock: SoA arena, materials, light types, material indices
		float4* d_mesh = nullptr;          // mesh table (3 * kMaxMeshes) | triangle stream | BVH nodes, one block
		size_t mesh_capacity = 0;          // in float4
		size_t triangle_offset = 0, node_offset = 0;   // in float4, inside d_mesh
		cudaEvent_t ev_upload = nullptr;   // last scene copy on this device (pinned source may be reused after it)
		cudaEvent_t ev_foreign = nullptr;  // last pixel-kernel launch on a caller-supplied stream: scene writes on `stream` wait for it
		bool foreign_pending = false;
		struct MeshSourceDevice
		{
			float* positions = nullptr; float* normals = nullptr; int32_t* indices = nullptr;
			// device-side BuildBVH: the other half of the indices / normals ping-pong pair, scratch and results
			float* normals_alt = nullptr; int32_t* indices_alt = nullptr;
			void* build_block = nullptr;
			rt::BuildParams build{};           // pointers into build_block, filled once per source upload
			bool built = false;                // build results are valid (emit_mesh_kernel may copy them)
		};
		std::vector<MeshSourceDevice> sources;   // untransformed meshes (rt_upload_mesh_source), by mesh id
		long long build_shared_limit = -1, subtree_shared_limit = -1;   // dynamic shared memory the build kernels may use here; -1 = not asked yet
		uint32_t* d_frame = nullptr;
		size_t frame_capacity = 0;         // in pixels
		unsigned long long* d_counters = nullptr;
		cudaEvent_t ev_begin = nullptr, ev_kernel = nullptr, ev_done = nullptr;
		cudaStream_t copy_stream = nullptr;                 // device-to-host copies that overlap the next band's kernel
		cudaEvent_t ev_band[16] = {};
		unsigned int* d_band_done = nullptr;                // kMaxBands counters for the single-launch progressive present
		unsigned int* h_wave_jobs = nullptr;                // mapped pinned: {view jobs, shadow jobs} of the last wavefront frame (a hint for the next one)
		unsigned int* d_wave_jobs = nullptr;                // ... as the device sees it
		unsigned int* h_flags = nullptr;                    // 64 words of mapped pinned memory: the band watcher's messages to the host
		unsigned int* d_flags = nullptr;                    // the same words as the device sees them
		uint32_t watch_tag = 0;                             // value the watcher stores for the current frame
		struct PendingPresent                               // a watched frame whose band copies the host still has to issue
		{
			bool active = false;
			int bands = 0, next = 0, strips_per_band = 0, strip_first = 0, strip_step = 1, total_strips = 0, W = 0, H = 0;
			void* target = nullptr; int32_t pitch_bytes = 0;
			unsigned long long spins = 0;
		} pending;
		uint8_t* d_band_table = nullptr;                    // strip -> band of the current schedule (make_band_schedule)
		uint8_t* h_band_table = nullptr;                    // its pinned source
		int band_table_strips = 0, band_table_bands = 0;
		unsigned int* d_queues = nullptr;                   // kQueueRing work queues {next, finished} of the persistent kernel (self re-arming)
		unsigned int queue_cursor = 0;
		// Cost feedback for the persistent kernel's tile order on small shares (see prepare_cell_order)
		struct CellSchedule
		{
			static constexpr int kSlots = 4, kMaxCells = 2048;
			// the launch geometry the costs belong to; anything else starts over
			int n_strips = 0, grid_x = 0, strip_first = -1, strip_step = -1, row_begin = -1, row_end = -1, cells_x = 0, cell_h_log2 = 0, n_cells = 0;
			int mesh_path = -1, lighting_mode = -1, shadows = -1;
			unsigned long long measured_scene = 0;          // rt_context::scene_version at the last measurement
			cudaStream_t stream = nullptr;
			unsigned int* d_cost = nullptr;                 // kSlots x kMaxCells, one slice per launch in flight
			unsigned int* h_cost = nullptr;                 // pinned read-back, same shape
			uint16_t* d_order = nullptr;                    // kSlots x kMaxCells order tables
			uint16_t* h_order = nullptr;                    // pinned sources of their uploads
			cudaEvent_t ev_cost[kSlots] = {}, ev_order[kSlots] = {};
			cudaEvent_t ev_last = nullptr;                  // after the newest launch that read a table / wrote costs
			bool cost_pending[kSlots] = {}, order_used[kSlots] = {};
			int cost_cursor = 0, order_cursor = 0;
			int orders_made = 0, launches_since_measured = 0;
			const uint16_t* current = nullptr;              // the table the next launch walks (NULL: none learned yet)
		} cells;
		// RT_KERNEL_WAVEFRONT: the split tables of all meshes and the per-pixel scratch of the five launches
		int32_t* d_split = nullptr;
		uint8_t* d_root_map = nullptr;
		unsigned long long split_version = 0;               // of the split tables in d_split / d_root_map (rt_context::split_version)
		cudaEvent_t ev_split = nullptr;                     // after the last copy out of the pinned mirrors
		size_t root_map_capacity = 0;
		void* d_wave = nullptr;
		size_t wave_pixels = 0, wave_view_tasks = 0, wave_shadow_tasks = 0, wave_meshes = 0, wave_lights = 0;
		rt::wave::WaveParams wave{};
		int sm_count = 0;
		rt::SceneDevice view{};
	};
}

struct rt_context
{
	std::vector<DeviceState> devs;
	std::string error;
	bool peer_stores = false;           // every device can store into device 0's frame buffer
	int32_t mesh_path = RT_MESH_PATH_AUTO;
	int32_t kernel_variant = RT_KERNEL_AUTO;

	// pinned host mirror of the static block (SoA, as uploaded) and of the mesh block
	int32_t* h_split = nullptr;         // pinned mirror of the split tables (kMaxMeshes * kSplitStride words)
	unsigned long long split_version = 1, h_split_version = 0;     // of the meshes' split tables / of what the pinned mirrors hold
	uint8_t* h_root_map = nullptr;      // pinned mirror of the node -> subtree maps of all meshes, back to back
	size_t h_root_map_capacity = 0;
	uint8_t* h_static = nullptr;
	float* arena = nullptr;             // views into h_static
	float4* materials = nullptr;
	int32_t* light_type = nullptr;
	uint8_t* bytes = nullptr;
	float4* h_mesh = nullptr;
	size_t h_mesh_capacity = 0;         // in float4
	int32_t n_spheres = 0, n_planes = 0, n_lights = 0, n_materials = 0;
	std::vector<HostMesh> meshes;
	bool static_dirty = true, mesh_dirty = true;   // mirrors changed since the last push (pushed lazily, once per frame)

	cudaEvent_t ev_gather = nullptr, ev_d2h = nullptr;   // on device 0
	rt_timing timing{};
	int32_t last_width = 0, last_height = 0;

	unsigned long long scene_version = 1; // bumped by every scene copy / device-side transform: what the measured tile costs belong to
	int32_t* h_build_status = nullptr;  // pinned, one word per mesh: status of the last device-side BVH build (device 0)

	// host surfaces the CALLER asked us to pin (rt_register_surface): the registration lives until
	// rt_unregister_surface / rt_destroy, and the caller must keep the memory mapped that long
	std::vector<std::pair<void*, size_t>> registered;
	void* staging = nullptr;            // pinned bounce buffer for surfaces CUDA does not know as pinned
	size_t staging_bytes = 0;
};

namespace
{
	int fail(rt_context* ctx, int code, const char* fmt, ...)
	{
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		if (ctx) ctx->error = buf; else g_create_error = buf;
		return code;
	}

#define RT_CUDA(ctx, call)                                                                      \
	do {                                                                                        \
		const cudaError_t e_ = (call);                                                          \
		if (e_ != cudaSuccess)                                                                  \
			return fail((ctx), RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

	inline float bits_as_float(int32_t v) { float f; memcpy(&f, &v, 4); return f; }
	inline int32_t float_bits(float f) { int32_t v; memcpy(&v, &f, 4); return v; }

	// Bitwise comparison of an uploaded array with its slice of the pinned mirror: an upload that changes nothing (the
	// drop-in re-sends the whole scene every frame) must not dirty the block, wait for copies or bump the scene version.
	template <typename T>
	inline bool same_bits(const T* mirror, const T* src, int n) { return n <= 0 || memcmp(mirror, src, sizeof(T) * (size_t)n) == 0; }

	using rt::KernelFn;
	using rt::pick_kernel;
	using rt::pick_kernel_x2;
	using rt::pick_kernel_persistent;

	// One wave of the persistent kernel: SMs x CTAs resident per SM (asked of the runtime once per kernel).
	int persistent_ctas_per_sm(KernelFn fn, size_t dynamic_smem)
	{
		static std::map<std::pair<KernelFn, size_t>, int> cache;
		static std::mutex lock;
		std::lock_guard<std::mutex> guard(lock);
		const auto key = std::make_pair(fn, dynamic_smem);
		auto it = cache.find(key);
		if (it != cache.end()) return it->second;
		int n = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, rt::kPersistentThreads, dynamic_smem) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
		cache[key] = n;
		return n;
	}

	// The mesh body a frame runs (the reference's `#ifdef BVH`), or -1 with the error set.
	int resolve_mesh_path(rt_context* ctx, int requested)
	{
		bool all_have_nodes = true;
		for (const HostMesh& m : ctx->meshes) if (!m.triangles.empty() && m.nodes.empty()) all_have_nodes = false;
		if (requested == RT_MESH_PATH_SLAB_LINEAR) return RT_MESH_PATH_SLAB_LINEAR;
		if (requested == RT_MESH_PATH_BVH)
		{
			if (!all_have_nodes) { fail(ctx, RT_ERR_BAD_STATE, "RT_MESH_PATH_BVH was requested but a mesh was uploaded without BVH nodes"); return -1; }
			return RT_MESH_PATH_BVH;
		}
		return (all_have_nodes && !ctx->meshes.empty()) ? RT_MESH_PATH_BVH : RT_MESH_PATH_SLAB_LINEAR;
	}

	int flush_uploads(rt_context* ctx);
	int run_device_transforms(rt_context* ctx, bool block_rewritten);

	int validate_frame(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame)
	{
		if (!camera || !frame) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "camera and frame must not be NULL");
		if (frame->width <= 0 || frame->height <= 0 || frame->width > 65536 || frame->height > 65536)
			return fail(ctx, RT_ERR_INVALID_ARGUMENT, "frame size %d x %d out of range", frame->width, frame->height);
		if (frame->lighting_mode < 0 || frame->lighting_mode > 3)
			return fail(ctx, RT_ERR_INVALID_ARGUMENT, "lighting_mode %d out of range", frame->lighting_mode);
		if (frame->r_shift > 24 || frame->g_shift > 24 || frame->b_shift > 24)
			return fail(ctx, RT_ERR_INVALID_ARGUMENT, "pixel format shifts out of range");
		for (const HostMesh& m : ctx->meshes)
			if (!m.uploaded) return fail(ctx, RT_ERR_BAD_STATE, "rt_set_mesh_count announced a mesh that was never uploaded");
		for (const HostMesh& m : ctx->meshes)
			if (m.has_source && !m.has_transform) return fail(ctx, RT_ERR_BAD_STATE, "a mesh uploaded with rt_upload_mesh_source has no transform yet (rt_transform_mesh)");
		return flush_uploads(ctx);
	}

	rt::FrameParams make_params(const rt_camera* c, const rt_frame_desc* f)
	{
		rt::FrameParams p{};
		p.cam_ox = c->origin[0]; p.cam_oy = c->origin[1]; p.cam_oz = c->origin[2]; p.fov = c->fov;
		p.right_x = c->right[0]; p.right_y = c->right[1]; p.right_z = c->right[2];
		p.up_x = c->up[0]; p.up_y = c->up[1]; p.up_z = c->up[2];
		p.fwd_x = c->forward[0]; p.fwd_y = c->forward[1]; p.fwd_z = c->forward[2];
		p.aspect = f->aspect_ratio;
		p.width = f->width; p.height = f->height;
		p.lighting_mode = f->lighting_mode; p.shadows = f->shadows_enabled ? 1 : 0;
		p.r_shift = f->r_shift; p.g_shift = f->g_shift; p.b_shift = f->b_shift; p.alpha_mask = f->alpha_mask;
		return p;
	}

	// 32-bit pattern fill (driver memset through the runtime: cudaMemset2D would need a pitch; D32 is exact)
	inline cudaError_t cuda_fill32(uint32_t* dst, uint32_t value, size_t count, cudaStream_t stream)
	{
		if ((value & 0xffu) * 0x01010101u == value) return cudaMemsetAsync(dst, (int)(value & 0xffu), count * sizeof(uint32_t), stream);
		rt::fill32_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 148 * 8), 256, 0, stream>>>(dst, value, count);
		return cudaGetLastError();
	}

	// Byte offset of the signal word behind a frame of `pixels` pixels (same formula on every rank).
	inline size_t signal_offset(size_t pixels) { return (pixels * sizeof(uint32_t) + 255) & ~size_t(255); }

	int ensure_frame(rt_context* ctx, DeviceState& d, size_t pixels)
	{
		if (d.frame_capacity >= pixels) return RT_OK;
		RT_CUDA(ctx, cudaSetDevice(d.device));
		if (d.d_frame) { RT_CUDA(ctx, cudaStreamSynchronize(d.stream)); RT_CUDA(ctx, cudaFree(d.d_frame)); d.d_frame = nullptr; }
		// + one 256-byte trailer: the completion signal word of rt_frame_signal / rt_frame_wait
		RT_CUDA(ctx, cudaMalloc(&d.d_frame, signal_offset(pixels) + 256));
		RT_CUDA(ctx, cudaMemset(d.d_frame, 0, signal_offset(pixels) + 256));
		d.frame_capacity = pixels;
		return RT_OK;
	}

	void refresh_view(rt_context* ctx, DeviceState& d)
	{
		rt::SceneDevice& v = d.view;
		const float* a = reinterpret_cast<const float*>(d.d_static + StaticBlock::arena);
		v.sphere_ox = a + ArenaLayout::sphere; v.sphere_oy = v.sphere_ox + rt::kMaxSpheres;
		v.sphere_oz = v.sphere_oy + rt::kMaxSpheres; v.sphere_r = v.sphere_oz + rt::kMaxSpheres;
		v.plane_ox = a + ArenaLayout::plane; v.plane_oy = v.plane_ox + rt::kMaxPlanes; v.plane_oz = v.plane_oy + rt::kMaxPlanes;
		v.plane_nx = v.plane_oz + rt::kMaxPlanes; v.plane_ny = v.plane_nx + rt::kMaxPlanes; v.plane_nz = v.plane_ny + rt::kMaxPlanes;
		v.light_ox = a + ArenaLayout::light; v.light_oy = v.light_ox + rt::kMaxLights; v.light_oz = v.light_oy + rt::kMaxLights;
		v.light_r = v.light_oz + rt::kMaxLights; v.light_g = v.light_r + rt::kMaxLights; v.light_b = v.light_g + rt::kMaxLights;
		v.light_intensity = v.light_b + rt::kMaxLights;
		v.sphere_mat = d.d_static + StaticBlock::bytes; v.plane_mat = v.sphere_mat + rt::kMaxSpheres;
		v.light_type = reinterpret_cast<const int32_t*>(d.d_static + StaticBlock::light_type);
		v.materials = reinterpret_cast<const float4*>(d.d_static + StaticBlock::materials);
		v.mesh_table = d.d_mesh;
		v.triangles = d.d_mesh + d.triangle_offset;
		v.bvh_nodes = d.d_mesh + d.node_offset;
		v.n_spheres = ctx->n_spheres; v.n_planes = ctx->n_planes; v.n_lights = ctx->n_lights;
		v.n_materials = ctx->n_materials; v.n_meshes = (int32_t)ctx->meshes.size();
		v.k_neg0 = make_float2(-0.f, -0.f); v.k_one = make_float2(1.f, 1.f); v.k_mone = make_float2(-1.f, -1.f);
	}

	// The pinned mirrors are about to be rewritten: wait until every device has consumed them.
	int wait_uploads(rt_context* ctx)
	{
		for (DeviceState& d : ctx->devs) RT_CUDA(ctx, cudaEventSynchronize(d.ev_upload));
		return RT_OK;
	}

	// Push the pinned mirror of the small static arrays to every device: one asynchronous copy each,
	// ordered before any later launch on the device's stream (ev_upload orders foreign streams).
	int push_static(rt_context* ctx)
	{
		for (DeviceState& d : ctx->devs)
		{
			RT_CUDA(ctx, cudaSetDevice(d.device));
			RT_CUDA(ctx, cudaMemcpyAsync(d.d_static, ctx->h_static, StaticBlock::total, cudaMemcpyHostToDevice, d.stream));
			RT_CUDA(ctx, cudaEventRecord(d.ev_upload, d.stream));
			refresh_view(ctx, d);
		}
		return RT_OK;
	}

	// Rebuild the mesh block (table | triangle stream | BVH nodes) in the pinned mirror and push it to
	// every device with one asynchronous copy.
	int push_meshes(rt_context* ctx)
	{
		size_t n_tris = 0, n_nodes = 0;
		for (const HostMesh& hm : ctx->meshes) { n_tris += hm.triangles.size(); n_nodes += hm.nodes.size(); }
		const size_t table = 3 * (size_t)rt::kMaxMeshes;
		const size_t tri_off = table, node_off = tri_off + n_tris + 6;     // + two padding records: the loops read ahead
		const size_t total = node_off + n_nodes;
		int rc = wait_uploads(ctx);
		if (rc != RT_OK) return rc;
		if (total > ctx->h_mesh_capacity)
		{
			if (ctx->h_mesh) cudaFreeHost(ctx->h_mesh);
			ctx->h_mesh = nullptr; ctx->h_mesh_capacity = 0;
			const size_t cap = std::max<size_t>(total + total / 2, 8 * 1024);
			RT_CUDA(ctx, cudaHostAlloc(&ctx->h_mesh, cap * sizeof(float4), cudaHostAllocPortable));
			ctx->h_mesh_capacity = cap;
		}
		float4* h = ctx->h_mesh;
		for (size_t i = 0; i < table; ++i) h[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		int32_t first = 0, first_node = 0;
		float4* tris = h + tri_off;
		float4* nodes = h + node_off;
		for (size_t m = 0; m < ctx->meshes.size(); ++m)
		{
			const HostMesh& hm = ctx->meshes[m];
			const int32_t count = (int32_t)(hm.triangles.size() / 3), node_count = (int32_t)(hm.nodes.size() / 2);
			h[3 * m + 0] = make_float4(hm.aabb_min[0], hm.aabb_max[0], hm.aabb_min[1], hm.aabb_max[1]);
			h[3 * m + 1] = make_float4(hm.aabb_min[2], hm.aabb_max[2], bits_as_float(first), bits_as_float(count));
			h[3 * m + 2] = make_float4(bits_as_float(hm.cull_mode), bits_as_float(hm.material), bits_as_float(first_node), bits_as_float(node_count));
			if (count) memcpy(tris + 3 * (size_t)first, hm.triangles.data(), hm.triangles.size() * sizeof(float4));
			if (node_count) memcpy(nodes + 2 * (size_t)first_node, hm.nodes.data(), hm.nodes.size() * sizeof(float4));
			first += count; first_node += node_count;
		}
		for (int k = 0; k < 6; ++k) tris[n_tris + k] = make_float4(0.f, 0.f, 0.f, 0.f);
		for (DeviceState& d : ctx->devs)
		{
			RT_CUDA(ctx, cudaSetDevice(d.device));
			if (total > d.mesh_capacity)
			{
				RT_CUDA(ctx, cudaDeviceSynchronize());
				if (d.d_mesh) RT_CUDA(ctx, cudaFree(d.d_mesh));
				d.d_mesh = nullptr;
				const size_t cap = std::max<size_t>(total + total / 2, 8 * 1024);
				RT_CUDA(ctx, cudaMalloc(&d.d_mesh, cap * sizeof(float4)));
				d.mesh_capacity = cap;
			}
			d.triangle_offset = tri_off; d.node_offset = node_off;
			RT_CUDA(ctx, cudaMemcpyAsync(d.d_mesh, h, total * sizeof(float4), cudaMemcpyHostToDevice, d.stream));
			RT_CUDA(ctx, cudaEventRecord(d.ev_upload, d.stream));
			refresh_view(ctx, d);
		}
		return RT_OK;
	}

	// Dynamic shared memory the two build kernels may use on this device (opt-in limit minus their static part);
	// raises the kernels' limits the first time.  0 when an attribute cannot be set: the build then works in global memory.
	template <typename Kernel>
	long long raise_shared_limit(Kernel kernel, int device)
	{
		int optin = 0;
		cudaFuncAttributes attr{};
		if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess ||
		    cudaFuncGetAttributes(&attr, kernel) != cudaSuccess) { cudaGetLastError(); return 0; }
		const long long room = (long long)optin - (long long)attr.sharedSizeBytes - 1024;
		if (room <= 0) return 0;
		if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room) != cudaSuccess) { cudaGetLastError(); return 0; }
		return room;
	}

	void query_build_shared_limits(DeviceState& d)
	{
		if (d.build_shared_limit >= 0) return;
		d.build_shared_limit = 0; d.subtree_shared_limit = 0;
		if (getenv("RT_B200_BUILD_GLOBAL")) return;       // measurement: force the global-memory work arrays
		d.build_shared_limit = raise_shared_limit(rt::update_transforms_bvh_kernel, d.device);
		d.subtree_shared_limit = raise_shared_limit(rt::build_subtrees_kernel, d.device);
	}

	// Carves the scratch / result arrays of update_transforms_bvh_kernel out of one allocation.
	int allocate_build_block(rt_context* ctx, DeviceState::MeshSourceDevice& sd, int32_t V, int32_t T)
	{
		const size_t N = (size_t)std::max(2 * T - 1, 1);
		size_t offset = 0;
		auto take = [&](size_t bytes) { const size_t at = offset; offset = (offset + bytes + 15) & ~size_t(15); return at; };
		const size_t Tn = (size_t)std::max(T, 1);
		const size_t o_tpos = take(12 * (size_t)std::max(V, 1)), o_cen = take(12 * Tn), o_min = take(12 * Tn), o_max = take(12 * Tn), o_tn = take(12 * Tn),
		             o_order = take(4 * Tn), o_tmp = take(4 * Tn), o_rb = take(4 * Tn), o_fr = take(4 * Tn), o_bl = take(4 * Tn),
		             o_first = take(4 * N), o_count = take(4 * N), o_escape = take(4 * N), o_box = take(24 * N),
		             o_qa = take(4 * Tn), o_qb = take(4 * Tn), o_tris = take(48 * Tn), o_nodes = take(32 * N), o_info = take(sizeof(int32_t) * rt::kInfoWords);
		RT_CUDA(ctx, cudaMalloc(&sd.build_block, offset));
		RT_CUDA(ctx, cudaMemset(sd.build_block, 0, offset));
		char* base = (char*)sd.build_block;
		rt::BuildParams& b = sd.build;
		b = rt::BuildParams{};
		b.positions = sd.positions; b.vertex_count = V; b.triangle_count = T;
		b.tpos = (float*)(base + o_tpos); b.centroid = (float*)(base + o_cen); b.tri_min = (unsigned int*)(base + o_min); b.tri_max = (unsigned int*)(base + o_max);
		b.tnormal = (float*)(base + o_tn); b.order = (int32_t*)(base + o_order); b.order_tmp = (int32_t*)(base + o_tmp);
		b.rights_before = (int32_t*)(base + o_rb); b.front_right = (int32_t*)(base + o_fr); b.back_left = (int32_t*)(base + o_bl);
		b.node_first = (int32_t*)(base + o_first); b.node_count = (int32_t*)(base + o_count); b.node_escape = (int32_t*)(base + o_escape);
		b.node_box = (float*)(base + o_box); b.queue_a = (int32_t*)(base + o_qa); b.queue_b = (int32_t*)(base + o_qb);
		b.result_triangles = (float4*)(base + o_tris); b.result_nodes = (float4*)(base + o_nodes); b.result_info = (int32_t*)(base + o_info);
		return RT_OK;
	}

	// Meshes that are transformed on the device: (re)send their untransformed source when it changed, then
	//  * without device BVH: run transform_mesh_kernel when the transform changed or the mesh block was just
	//    rewritten from the host mirror (idempotent);
	//  * with device BVH: run update_transforms_bvh_kernel once per rt_transform_mesh call since the last frame, in
	//    call order (each build starts from the triangle order the previous one left; each writes its result into
	//    the block as well); when only the block was rewritten, emit_mesh_kernel copies the last build back.
	int run_device_transforms(rt_context* ctx, bool block_rewritten)
	{
		for (size_t m = 0; m < ctx->meshes.size(); ++m)
		{
			HostMesh& hm = ctx->meshes[m];
			if (!hm.has_source) continue;
			const int32_t T = (int32_t)(hm.src_indices.size() / 3), V = (int32_t)(hm.src_positions.size() / 3);
			int32_t first = 0, first_node = 0;
			for (size_t k = 0; k < m; ++k) { first += (int32_t)(ctx->meshes[k].triangles.size() / 3); first_node += (int32_t)(ctx->meshes[k].nodes.size() / 2); }
			for (DeviceState& d : ctx->devs)
			{
				RT_CUDA(ctx, cudaSetDevice(d.device));
				if (d.sources.size() < ctx->meshes.size()) d.sources.resize(ctx->meshes.size());
				DeviceState::MeshSourceDevice& sd = d.sources[m];
				if (hm.source_dirty)
				{
					RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
					cudaFree(sd.positions); cudaFree(sd.normals); cudaFree(sd.indices); cudaFree(sd.normals_alt); cudaFree(sd.indices_alt); cudaFree(sd.build_block);
					sd = DeviceState::MeshSourceDevice{};
					RT_CUDA(ctx, cudaMalloc(&sd.positions, std::max<size_t>(hm.src_positions.size(), 1) * sizeof(float)));
					RT_CUDA(ctx, cudaMalloc(&sd.normals, std::max<size_t>(hm.src_normals.size(), 1) * sizeof(float)));
					RT_CUDA(ctx, cudaMalloc(&sd.indices, std::max<size_t>(hm.src_indices.size(), 1) * sizeof(int32_t)));
					// one-off upload (pageable source: the copy is complete when the call returns)
					RT_CUDA(ctx, cudaMemcpy(sd.positions, hm.src_positions.data(), hm.src_positions.size() * sizeof(float), cudaMemcpyHostToDevice));
					RT_CUDA(ctx, cudaMemcpy(sd.normals, hm.src_normals.data(), hm.src_normals.size() * sizeof(float), cudaMemcpyHostToDevice));
					RT_CUDA(ctx, cudaMemcpy(sd.indices, hm.src_indices.data(), hm.src_indices.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
					if (hm.device_bvh)
					{
						RT_CUDA(ctx, cudaMalloc(&sd.normals_alt, std::max<size_t>(hm.src_normals.size(), 1) * sizeof(float)));
						RT_CUDA(ctx, cudaMalloc(&sd.indices_alt, std::max<size_t>(hm.src_indices.size(), 1) * sizeof(int32_t)));
						const int rc = allocate_build_block(ctx, sd, V, T);
						if (rc != RT_OK) return rc;
					}
				}
				if (hm.device_bvh)
				{
					for (const std::array<float, 16>& transform : hm.pending_builds)
					{
						rt::BuildParams& b = sd.build;
						memcpy(b.m, transform.data(), sizeof b.m);
						b.indices_in = sd.indices; b.normals_in = sd.normals; b.indices_out = sd.indices_alt; b.normals_out = sd.normals_alt;
						// the build also writes the mesh's slices of the scene block (the last one of the frame is what stays)
						b.scene_triangles = T > 0 ? d.d_mesh + d.triangle_offset + 3 * (size_t)first : nullptr;     // an empty mesh owns no slices
						b.scene_nodes = T > 0 ? d.d_mesh + d.node_offset + 2 * (size_t)first_node : nullptr;
						b.scene_table = T > 0 ? d.d_mesh + 3 * m : nullptr;
						query_build_shared_limits(d);
						size_t work_bytes = sizeof(float) * rt::kBuildWorkWordsPerTriangle * (size_t)std::max(T, 1);
						b.work_in_shared = (long long)work_bytes <= d.build_shared_limit ? 1 : 0;
						if (!b.work_in_shared) work_bytes = 0;
						// subtrees are at most the whole mesh; what does not fit works in the global scratch
						const long long fit = d.subtree_shared_limit / (long long)(sizeof(int32_t) * rt::kSubtreeWordsPerTriangle);
						b.subtree_shared_triangles = (int32_t)std::min<long long>(fit, std::min(T, 4096));
						const size_t subtree_bytes = sizeof(int32_t) * rt::kSubtreeWordsPerTriangle * (size_t)b.subtree_shared_triangles;
						rt::update_transforms_bvh_kernel<<<1, rt::kBuildThreads, work_bytes, d.stream>>>(b);
						RT_CUDA(ctx, cudaGetLastError());
						rt::build_subtrees_kernel<<<rt::kSubtreeCtas, rt::kSubtreeThreads, subtree_bytes, d.stream>>>(b);
						ctx->timing.kernel_launches++;
						RT_CUDA(ctx, cudaGetLastError());
						std::swap(sd.indices, sd.indices_alt); std::swap(sd.normals, sd.normals_alt);     // the order the build left
						sd.built = true;
						if (&d == &ctx->devs[0] && T > rt::BvhLink::kMaxLeafTriangles) RT_CUDA(ctx, cudaMemcpyAsync(ctx->h_build_status + m, b.result_info + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, d.stream));
						ctx->timing.kernel_launches++;
					}
					if (!hm.pending_builds.empty()) RT_CUDA(ctx, cudaEventRecord(d.ev_upload, d.stream));
					if (sd.built && hm.pending_builds.empty() && block_rewritten && T > 0)
					{
						rt::EmitParams e{};
						e.result_triangles = sd.build.result_triangles; e.result_nodes = sd.build.result_nodes; e.result_info = sd.build.result_info;
						e.triangle_count = T;
						e.triangles = d.d_mesh + d.triangle_offset + 3 * (size_t)first;
						e.nodes = d.d_mesh + d.node_offset + 2 * (size_t)first_node;
						e.table = d.d_mesh + 3 * m;
						rt::emit_mesh_kernel<<<1, 256, 0, d.stream>>>(e);
						RT_CUDA(ctx, cudaGetLastError());
						RT_CUDA(ctx, cudaEventRecord(d.ev_upload, d.stream));
					}
				}
				else if (hm.has_transform && (hm.transform_dirty || hm.source_dirty || block_rewritten) && T > 0)
				{
					rt::TransformParams tp{};
					memcpy(tp.m, hm.transform, sizeof tp.m);
					tp.positions = sd.positions; tp.indices = sd.indices; tp.normals = sd.normals;
					tp.triangle_count = T;
					tp.triangles = d.d_mesh + d.triangle_offset + 3 * (size_t)first;
					tp.table = d.d_mesh + 3 * m;
					tp.first_triangle = first;
					rt::transform_mesh_kernel<<<1, 256, 0, d.stream>>>(tp);
					RT_CUDA(ctx, cudaGetLastError());
					RT_CUDA(ctx, cudaEventRecord(d.ev_upload, d.stream));
				}
			}
			// a leaf wider than the link's triangle field can only come out of a mesh that large: only then is the
			// status word worth a host synchronisation
			const bool check_status = hm.device_bvh && !hm.pending_builds.empty() && T > rt::BvhLink::kMaxLeafTriangles;
			hm.source_dirty = false; hm.transform_dirty = false; hm.pending_builds.clear();
			if (check_status)
			{
				RT_CUDA(ctx, cudaStreamSynchronize(ctx->devs[0].stream));
				if (ctx->h_build_status[m] != 0)
					return fail(ctx, RT_ERR_CAPACITY, "the device-side BVH build of mesh %d produced a leaf wider than %d triangles", (int)m, rt::BvhLink::kMaxLeafTriangles);
			}
		}
		return RT_OK;
	}

	// Scene uploads only touch the pinned mirrors; the device copies happen here, once, right before the
	// next launch (one copy per dirty block and device instead of one per rt_upload_* call).
	int flush_uploads(rt_context* ctx)
	{
		int rc = RT_OK;
		bool writes = ctx->static_dirty || ctx->mesh_dirty;
		for (const HostMesh& hm : ctx->meshes) writes = writes || hm.transform_dirty || hm.source_dirty || !hm.pending_builds.empty();
		if (writes)
		{
			// A pixel kernel launched on a caller-supplied stream may still be reading d_static / d_mesh: every scene
			// write below rides d.stream, so d.stream queues up behind that launch first (launch() recorded ev_foreign).
			for (DeviceState& d : ctx->devs)
			{
				if (!d.foreign_pending) continue;
				RT_CUDA(ctx, cudaSetDevice(d.device));
				RT_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.ev_foreign, 0));
				d.foreign_pending = false;
			}
		}
		if (ctx->static_dirty) { if ((rc = push_static(ctx)) != RT_OK) return rc; ctx->static_dirty = false; ctx->scene_version++; }
		bool pushed = false;
		if (ctx->mesh_dirty) { if ((rc = push_meshes(ctx)) != RT_OK) return rc; ctx->mesh_dirty = false; pushed = true; ctx->scene_version++; }
		for (const HostMesh& hm : ctx->meshes) if (hm.transform_dirty || !hm.pending_builds.empty()) { ctx->scene_version++; break; }
		return run_device_transforms(ctx, pushed);
	}

	inline float std_min_f(float a, float b) { return (b < a) ? b : a; }
	inline float std_max_f(float a, float b) { return (a < b) ? b : a; }

	// Launch one device's share.  `dst` must be addressable from device d.
	// ---- expensive tiles first ---------------------------------------------------------------------------------------
	// A launch ends when its last warp tile does.  Tiles that see a mesh take up to four times as long as the others
	// (a deep traversal for the view ray and for every shadow ray), and in the plain top-to-bottom order the queue runs
	// empty while many of them are still in flight: measured on the 4K bunny frame, 29 us of a 775 us launch and 52 us
	// of a 132 us launch (one eighth of the frame) pass between "queue empty" and "last warp out".  Handing out the
	// tiles inside the screen rectangle of the meshes' boxes first leaves the cheap ones for the end.  Scheduling
	// only: which tile a queue position means; pixels do not depend on it.  Not under a progressive present, whose
	// copies follow the bands top to bottom.
	void tiles_to_render_first(rt_context* ctx, rt::FrameParams& p, int n_strips, int resident_warps)
	{
		static const bool off = getenv("RT_B200_PLAIN_ORDER") != nullptr;
		p.first_tiles = 0;
		if (off || p.band_done || ctx->meshes.empty() || p.grid_x < 2 || n_strips < 2) return;
		// worth its decode only when the drain is a large part of the launch: fewer than 32 warp tiles per resident warp
		// (measured on the 4K bunny frame: the whole frame +0.6 %, a half +-0, a quarter -2.3 %, an eighth -8.5 %)
		if ((long long)p.grid_x * n_strips * rt::kSignalsPerTile >= 32ll * resident_warps) return;
		const float fov = p.fov, aspect = p.aspect;
		if (!(fov > 0.f) || !(aspect > 0.f)) return;
		// screen bounds (pixels) of every mesh box corner; a corner at or behind the camera plane gives up
		float x_lo = 1e30f, x_hi = -1e30f, y_lo = 1e30f, y_hi = -1e30f;
		bool any = false;
		for (const HostMesh& hm : ctx->meshes)
		{
			if (hm.triangles.empty()) continue;
			float corners[8][3];
			if (hm.has_source)
			{
				if (!hm.has_transform || hm.src_positions.empty()) return;
				float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
				for (size_t v = 0; v + 2 < hm.src_positions.size(); v += 3)
					for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], hm.src_positions[v + a]); hi[a] = std::max(hi[a], hm.src_positions[v + a]); }
				const float* m = hm.transform;
				for (int c = 0; c < 8; ++c)
				{
					const float x = (c & 1) ? hi[0] : lo[0], y = (c & 2) ? hi[1] : lo[1], z = (c & 4) ? hi[2] : lo[2];
					for (int a = 0; a < 3; ++a) corners[c][a] = x * m[a] + y * m[4 + a] + z * m[8 + a] + m[12 + a];
				}
			}
			else
			{
				for (int c = 0; c < 8; ++c)
					for (int a = 0; a < 3; ++a) corners[c][a] = ((c >> a) & 1) ? hm.aabb_max[a] : hm.aabb_min[a];
			}
			for (int c = 0; c < 8; ++c)
			{
				const float dx = corners[c][0] - p.cam_ox, dy = corners[c][1] - p.cam_oy, dz = corners[c][2] - p.cam_oz;
				const float f = dx * p.fwd_x + dy * p.fwd_y + dz * p.fwd_z;
				if (!(f > 1e-3f)) return;
				const float cx = (dx * p.right_x + dy * p.right_y + dz * p.right_z) / f, cy = (dx * p.up_x + dy * p.up_y + dz * p.up_z) / f;
				// inverse of Renderer.cpp:107-108
				const float px = (cx / (aspect * fov) + 1.f) * 0.5f * (float)p.width, py = (1.f - cy / fov) * 0.5f * (float)p.height;
				if (!(px == px) || !(py == py)) return;
				x_lo = std::min(x_lo, px); x_hi = std::max(x_hi, px); y_lo = std::min(y_lo, py); y_hi = std::max(y_hi, py);
				any = true;
			}
		}
		if (!any) return;
		const int total_strips = (p.row_end - p.row_begin + rt::kBlockH - 1) / rt::kBlockH;
		int x0 = std::max(0, (int)std::floor(x_lo / rt::kBlockW)), x1 = std::min(p.grid_x, (int)std::floor(x_hi / rt::kBlockW) + 1);
		const int s0 = std::max(0, (int)std::floor((y_lo - (float)p.row_begin) / rt::kBlockH)), s1 = std::min(total_strips, (int)std::floor((y_hi - (float)p.row_begin) / rt::kBlockH) + 1);
		if (x1 <= x0 || s1 <= s0) return;
		// the magic-number divisions of the decode need divisors of at least 2
		if (x1 - x0 < 2) { if (x1 < p.grid_x) ++x1; else --x0; }
		if (p.grid_x - (x1 - x0) == 1) { x0 = 0; x1 = p.grid_x; }
		// frame strips [s0, s1) -> this launch's strips k with k * step + first inside
		const int k0 = std::max(0, (s0 - p.strip_first + p.strip_step - 1) / p.strip_step);
		const int k1 = std::min(n_strips, s1 > p.strip_first ? (s1 - 1 - p.strip_first) / p.strip_step + 1 : 0);
		if (k1 <= k0) return;
		const int w = x1 - x0;
		const long long first = (long long)w * (k1 - k0), all = (long long)p.grid_x * n_strips;
		if (first >= all || first * 8 < all / 8) return;         // everything, or too little to matter
		p.first_x0 = x0; p.first_x1 = x1; p.first_k0 = k0; p.first_k1 = k1; p.first_tiles = (int32_t)first;
		p.first_w_magic = (uint32_t)((1ull << 32) / (unsigned)w) + 1u;
		p.rest_w_magic = p.grid_x > w ? (uint32_t)((1ull << 32) / (unsigned)(p.grid_x - w)) + 1u : 0u;
	}

	// ---- ... and, from the second launch of the same geometry on, the cells that WERE expensive first -----------------
	// The rectangle above is a guess made before anything is known; it misses, for instance, the floor in front of the
	// bunny, whose shadow rays climb through the mesh's box.  The kernel therefore also measures: every warp tile adds
	// its SM clocks to the counter of its cell (4 or 8 tile columns x 2^h strips, at most 2048 cells per launch), the
	// counters are read back asynchronously, and the next launch of the same geometry on the same stream walks the
	// cells in descending cost (longest processing time first).  Nothing here waits: a read-back that has not
	// finished is simply not used yet.  Same conditions as the rectangle (small shares, no progressive present).
	int ensure_cell_buffers(rt_context* ctx, DeviceState& d)
	{
		DeviceState::CellSchedule& cs = d.cells;
		if (cs.d_cost) return RT_OK;
		constexpr size_t n = (size_t)DeviceState::CellSchedule::kSlots * DeviceState::CellSchedule::kMaxCells;
		RT_CUDA(ctx, cudaMalloc(&cs.d_cost, n * sizeof(unsigned int)));
		RT_CUDA(ctx, cudaMalloc(&cs.d_order, n * sizeof(uint16_t)));
		RT_CUDA(ctx, cudaHostAlloc(&cs.h_cost, n * sizeof(unsigned int), cudaHostAllocPortable));
		RT_CUDA(ctx, cudaHostAlloc(&cs.h_order, n * sizeof(uint16_t), cudaHostAllocPortable));
		for (int i = 0; i < DeviceState::CellSchedule::kSlots; ++i)
		{
			RT_CUDA(ctx, cudaEventCreateWithFlags(&cs.ev_cost[i], cudaEventDisableTiming));
			RT_CUDA(ctx, cudaEventCreateWithFlags(&cs.ev_order[i], cudaEventDisableTiming));
		}
		RT_CUDA(ctx, cudaEventCreateWithFlags(&cs.ev_last, cudaEventDisableTiming));
		return RT_OK;
	}

	// Before a persistent launch.  Sets p.cell_* (order and / or cost slice); *cost_slot = the slice to read back after
	// the launch, or -1.
	int prepare_cell_order(rt_context* ctx, DeviceState& d, rt::FrameParams& p, cudaStream_t stream, int n_strips, int resident_warps, int mesh_path, int* cost_slot)
	{
		static const bool off = getenv("RT_B200_PLAIN_ORDER") != nullptr || getenv("RT_B200_NO_COST_FEEDBACK") != nullptr;
		*cost_slot = -1;
		p.cell_order = nullptr; p.cell_cost = nullptr; p.total_items = 0; p.cells_x = 1; p.cell_w_log2 = 3; p.cell_h_log2 = 0; p.cells_x_magic = 0;
		if (off || p.band_done || p.grid_x < 2 || n_strips < 2) return RT_OK;
		// measured on the 4K bunny frame against the plain order: whole frame -1.6 %, a half -4 %, a quarter -5 %, an
		// eighth -12 % (RT_B200_ORDER_GATE = n: only launches with fewer than n warp tiles per resident warp)
		static const long long gate = [] { const char* e = getenv("RT_B200_ORDER_GATE"); const int v = e ? atoi(e) : 0; return (long long)(v > 0 ? v : (1 << 20)); }();
		if ((long long)p.grid_x * n_strips * rt::kSignalsPerTile >= gate * resident_warps) return RT_OK;
		constexpr int kSlots = DeviceState::CellSchedule::kSlots, kMaxCells = DeviceState::CellSchedule::kMaxCells;
		// small shares want fine cells (4 tile columns: an eighth of the 4K frame 0.148 -> 0.142 ms, a quarter 0.238 -> 0.229),
		// large ones lose a little locality to them (a half 0.419 -> 0.430): 8 columns from 16 warp tiles per warp on
		const int w_log2 = ((long long)p.grid_x * n_strips * rt::kSignalsPerTile < 16ll * resident_warps) ? 2 : 3;
		const int cells_x = (p.grid_x + (1 << w_log2) - 1) >> w_log2;
		if (cells_x < 2 || cells_x > kMaxCells / 2) return RT_OK;
		int h = 0;
		while ((((n_strips + (1 << h) - 1) >> h) * cells_x) > kMaxCells) ++h;
		const int cells_y = (n_strips + (1 << h) - 1) >> h, n_cells = cells_x * cells_y;
		if (n_cells < 4) return RT_OK;
		int rc = ensure_cell_buffers(ctx, d);
		if (rc != RT_OK) return rc;
		DeviceState::CellSchedule& cs = d.cells;
		if (cs.n_strips != n_strips || cs.grid_x != p.grid_x || cs.strip_first != p.strip_first || cs.strip_step != p.strip_step ||
		    cs.row_begin != p.row_begin || cs.row_end != p.row_end || cs.stream != stream ||
		    cs.mesh_path != mesh_path || cs.lighting_mode != p.lighting_mode || cs.shadows != p.shadows)
		{
			// tables and counters of the old stream may still be in use there: the new stream queues up behind it
			if (cs.stream && cs.stream != stream) RT_CUDA(ctx, cudaStreamWaitEvent(stream, cs.ev_last, 0));
			cs.current = nullptr; cs.orders_made = 0; cs.launches_since_measured = 0;
			for (bool& b : cs.cost_pending) b = false;           // late read-backs of the old geometry are ignored
			cs.n_strips = n_strips; cs.grid_x = p.grid_x; cs.strip_first = p.strip_first; cs.strip_step = p.strip_step;
			cs.row_begin = p.row_begin; cs.row_end = p.row_end; cs.stream = stream;
			cs.mesh_path = mesh_path; cs.lighting_mode = p.lighting_mode; cs.shadows = p.shadows;
			cs.cells_x = cells_x; cs.cell_h_log2 = h; cs.n_cells = n_cells;
		}
		p.cells_x = cells_x; p.cell_w_log2 = w_log2; p.cell_h_log2 = h; p.cells_x_magic = (uint32_t)((1ull << 32) / (unsigned)cells_x) + 1u;
		// newest finished read-back -> a new order table
		int fresh = -1;
		for (int i = 0; i < kSlots; ++i)
		{
			const int slot = (cs.cost_cursor + kSlots - 1 - i) % kSlots;
			if (!cs.cost_pending[slot] || cudaEventQuery(cs.ev_cost[slot]) != cudaSuccess) continue;
			if (fresh < 0) fresh = slot;
			cs.cost_pending[slot] = false;
		}
		cudaGetLastError();      // cudaErrorNotReady of the queries above
		if (fresh >= 0)
		{
			const int oslot = cs.order_cursor % kSlots;
			if (!cs.order_used[oslot] || cudaEventQuery(cs.ev_order[oslot]) == cudaSuccess)
			{
				const unsigned int* cost = cs.h_cost + (size_t)fresh * kMaxCells;
				uint16_t* order = cs.h_order + (size_t)oslot * kMaxCells;
				static thread_local int idx[kMaxCells];
				for (int c = 0; c < n_cells; ++c) idx[c] = c;
				std::stable_sort(idx, idx + n_cells, [&](int a, int b) { return cost[a] > cost[b]; });
				for (int c = 0; c < n_cells; ++c) order[c] = (uint16_t)idx[c];
				RT_CUDA(ctx, cudaMemcpyAsync(cs.d_order + (size_t)oslot * kMaxCells, order, sizeof(uint16_t) * (size_t)n_cells, cudaMemcpyHostToDevice, stream));
				RT_CUDA(ctx, cudaEventRecord(cs.ev_order[oslot], stream));
				cs.order_used[oslot] = true; cs.order_cursor++; cs.orders_made++;
				cs.current = cs.d_order + (size_t)oslot * kMaxCells;
			}
			else cudaGetLastError();
		}
		if (cs.current)
		{
			p.cell_order = cs.current;
			p.total_items = (n_cells << (w_log2 + h)) * rt::kSignalsPerTile;
			p.first_tiles = 0;                                // the measured order replaces the guessed rectangle
		}
		// Measuring costs three stream operations around the kernel (clear, read back, and the upload of the order it
		// leads to), which is most of what a good order wins on a 150 us launch: measure the first launches of a geometry
		// (under the guessed order, then under the first measured one), after that only every 256th (camera and scene
		// move on; a stale order is still a valid one).
		const int cslot = cs.cost_cursor % kSlots;
		bool in_flight = false;
		for (bool b : cs.cost_pending) in_flight = in_flight || b;
		// ... or every 16th while the scene keeps changing (uploads, new poses)
		const bool measure = !in_flight && (cs.orders_made < 2 || cs.launches_since_measured >= 256 ||
		                                    (cs.measured_scene != ctx->scene_version && cs.launches_since_measured >= 16));
		cs.launches_since_measured++;
		if (measure && !cs.cost_pending[cslot])
		{
			cs.launches_since_measured = 0; cs.measured_scene = ctx->scene_version;
			RT_CUDA(ctx, cudaMemsetAsync(cs.d_cost + (size_t)cslot * kMaxCells, 0, sizeof(unsigned int) * (size_t)n_cells, stream));
			p.cell_cost = cs.d_cost + (size_t)cslot * kMaxCells;
			*cost_slot = cslot;
		}
		return RT_OK;
	}

	int finish_cell_order(rt_context* ctx, DeviceState& d, const rt::FrameParams& p, cudaStream_t stream, int cost_slot)
	{
		DeviceState::CellSchedule& cs = d.cells;
		if (!p.cell_cost && !p.cell_order) return RT_OK;
		RT_CUDA(ctx, cudaEventRecord(cs.ev_last, stream));
		if (cost_slot < 0) return RT_OK;
		constexpr int kMaxCells = DeviceState::CellSchedule::kMaxCells;
		RT_CUDA(ctx, cudaMemcpyAsync(cs.h_cost + (size_t)cost_slot * kMaxCells, cs.d_cost + (size_t)cost_slot * kMaxCells,
		                             sizeof(unsigned int) * (size_t)cs.n_cells, cudaMemcpyDeviceToHost, stream));
		RT_CUDA(ctx, cudaEventRecord(cs.ev_cost[cost_slot], stream));
		cs.cost_pending[cost_slot] = true; cs.cost_cursor++;
		return RT_OK;
	}

	void build_split(HostMesh& hm);

	// The split tables of RT_KERNEL_WAVEFRONT on device `d`, current with the meshes: built per mesh when first wanted
	// (build_split), staged once per change in the pinned mirrors, copied on the launch's stream.
	int ensure_splits(rt_context* ctx, DeviceState& d, cudaStream_t stream)
	{
		bool rebuilt = false;
		for (HostMesh& hm : ctx->meshes)
			if (!hm.split_built) { build_split(hm); hm.split_built = true; rebuilt = true; }
		if (rebuilt) ++ctx->split_version;
		if (d.split_version == ctx->split_version) return RT_OK;
		size_t map_bytes = 16;
		for (const HostMesh& hm : ctx->meshes) map_bytes += hm.root_map.size();
		const size_t split_words = ctx->meshes.size() * (size_t)rt::wave::kSplitStride;
		if (ctx->h_split_version != ctx->split_version)
		{
			// nobody may still be reading the mirrors' old contents
			for (DeviceState& o : ctx->devs) if (o.split_version) RT_CUDA(ctx, cudaEventSynchronize(o.ev_split));
			if (map_bytes > ctx->h_root_map_capacity)
			{
				if (ctx->h_root_map) cudaFreeHost(ctx->h_root_map);
				ctx->h_root_map = nullptr; ctx->h_root_map_capacity = 0;
				RT_CUDA(ctx, cudaHostAlloc(&ctx->h_root_map, map_bytes * 2, cudaHostAllocPortable));
				ctx->h_root_map_capacity = map_bytes * 2;
			}
			memset(ctx->h_split, 0, split_words * sizeof(int32_t));
			size_t at = 0;
			for (size_t m = 0; m < ctx->meshes.size(); ++m)
			{
				const HostMesh& hm = ctx->meshes[m];
				if (hm.split.empty()) continue;
				int32_t* block = ctx->h_split + m * rt::wave::kSplitStride;
				memcpy(block, hm.split.data(), sizeof(int32_t) * rt::wave::kSplitStride);
				block[1] = (int32_t)at;
				memcpy(ctx->h_root_map + at, hm.root_map.data(), hm.root_map.size());
				at += hm.root_map.size();
			}
			ctx->h_split_version = ctx->split_version;
		}
		if (map_bytes > d.root_map_capacity)
		{
			RT_CUDA(ctx, cudaDeviceSynchronize());
			if (d.d_root_map) RT_CUDA(ctx, cudaFree(d.d_root_map));
			d.d_root_map = nullptr; d.root_map_capacity = 0;
			RT_CUDA(ctx, cudaMalloc(&d.d_root_map, map_bytes * 2));
			d.root_map_capacity = map_bytes * 2;
		}
		RT_CUDA(ctx, cudaMemcpyAsync(d.d_root_map, ctx->h_root_map, map_bytes, cudaMemcpyHostToDevice, stream));
		if (split_words) RT_CUDA(ctx, cudaMemcpyAsync(d.d_split, ctx->h_split, split_words * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
		RT_CUDA(ctx, cudaEventRecord(d.ev_split, stream));
		d.split_version = ctx->split_version;
		return RT_OK;
	}

	// `watched` (optional): p.host_flags asks for the band watcher (rt_kernel.cuh, watch_bands); set to true when the launch
	// has one - only the persistent kernel can - else the caller orders its copies with stream waits.
	int launch(rt_context* ctx, DeviceState& d, rt::FrameParams p, cudaStream_t stream, int n_strips, bool* watched = nullptr)
	{
		if (watched) *watched = false;
		if (n_strips <= 0) return RT_OK;
		const int path = resolve_mesh_path(ctx, ctx->mesh_path);
		if (path < 0) return RT_ERR_BAD_STATE;
		p.vector_store = (p.width % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dst) & 15u) == 0);
		const dim3 grid((unsigned)((p.width + rt::kBlockW - 1) / rt::kBlockW), (unsigned)n_strips, 1);
		if (grid.y > 65535u) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "frame too tall for one launch");
		if (stream != d.stream) RT_CUDA(ctx, cudaStreamWaitEvent(stream, d.ev_upload, 0));   // scene copies ride d.stream
		p.grid_x = (int32_t)grid.x; p.n_strips = n_strips;
		p.grid_x_magic = (uint32_t)((1ull << 32) / grid.x) + 1u;
		// the persistent kernel decodes tile -> (strip, column) with one multiply; exact while tile * grid_x < 2^32
		const bool decodable = grid.x > 1u && (unsigned long long)grid.x * grid.y * grid.x < (1ull << 32);
		int variant = ctx->kernel_variant;
		const long long tiles = (long long)grid.x * (long long)grid.y;
		// RT_KERNEL_WAVEFRONT needs the BVH body, host-uploaded trees (their split tables) and few enough lights for one mask word.
		// AUTO: rays as the unit of work pay when the frame alone cannot fill the machine (fewer than ~4 warp tiles per
		// resident warp) while its pixels are expensive (deep trees walked without pruning): measured on
		// Scene_W4_OptionalScene, see DESIGN.md
		static const long long wave_min_nodes = [] { const char* e = getenv("RT_B200_WAVE_MIN_NODES"); return e ? atoll(e) : 1024ll; }();
		static const long long wave_max_tiles_per_sm = [] { const char* e = getenv("RT_B200_WAVE_MAX_TILES_PER_SM"); return e ? atoll(e) : 320ll; }();
		long long mesh_nodes = 0, wave_subtrees = 0;
		for (const HostMesh& hm : ctx->meshes) if (!hm.triangles.empty()) mesh_nodes += (long long)(hm.nodes.size() / 2);
		bool wave_possible = path == RT_MESH_PATH_BVH && ctx->n_lights <= 16 && tiles * rt::kSignalsPerTile < (1ll << 23);
		const bool wave_by_choice = variant == RT_KERNEL_AUTO && mesh_nodes >= wave_min_nodes && tiles * rt::kSignalsPerTile <= wave_max_tiles_per_sm * d.sm_count;
		if (wave_possible && (variant == RT_KERNEL_WAVEFRONT || wave_by_choice))
		{
			const int src = ensure_splits(ctx, d, stream);       // (only now: the tables cost host time per mesh upload)
			if (src != RT_OK) return src;
			for (const HostMesh& hm : ctx->meshes)
			{
				if (hm.triangles.empty()) continue;
				if (hm.split.empty()) wave_possible = false; else wave_subtrees += hm.split[0];
			}
		}
		else wave_possible = false;
		// the job lists are sized for the worst case (every tile reaches every subtree for every light): keep them below 64 MB
		if (tiles * rt::kSignalsPerTile * std::max(wave_subtrees, 1ll) * std::max(ctx->n_lights, 1) > (8ll << 20)) wave_possible = false;
		if (tiles * rt::kThreads * (long long)std::max<size_t>(ctx->meshes.size(), 1) * std::max(ctx->n_lights, 1) > (16ll << 20)) wave_possible = false;     // the per-ray subtree masks: <= 128 MB
		if (variant == RT_KERNEL_WAVEFRONT && !wave_possible) variant = RT_KERNEL_AUTO;
		if (wave_by_choice && wave_possible) variant = RT_KERNEL_WAVEFRONT;
		KernelFn persistent = nullptr;
		int wave = 0;
		if (variant == RT_KERNEL_AUTO || variant == RT_KERNEL_PERSISTENT)
		{
			persistent = pick_kernel_persistent(p.lighting_mode, p.shadows, path == RT_MESH_PATH_BVH);
			wave = d.sm_count * persistent_ctas_per_sm(persistent, rt::dynamic_smem_bytes(rt::kPersistentThreads, ctx->n_materials));
			// AUTO: persistent warps pay off once the frame is many waves deep; small frames keep one CTA per tile.
			// Device-only launches can walk their tiles in measured-cost order (prepare_cell_order), which already pays at
			// 6 warp tiles per resident warp (an eighth of the 4K bunny frame: 0.148 ms against 0.162 ms tiled)
			if (variant == RT_KERNEL_AUTO)
				variant = (tiles * rt::kSignalsPerTile >= (p.band_done ? 8ll : 4ll) * wave * (rt::kPersistentThreads / 32)) ? RT_KERNEL_PERSISTENT : RT_KERNEL_SCALAR;
			if (!decodable) variant = RT_KERNEL_SCALAR;
		}
		if (variant != RT_KERNEL_PERSISTENT) p.host_flags = nullptr;
		if (variant == RT_KERNEL_WAVEFRONT)
		{
			const size_t pixels = (size_t)tiles * rt::kThreads;
			// small frames wait for their longest jobs: those kernels take the subtrees in parts (rt_wave_params.h); measured
			// crossover (tools/wave_sweep.sh): ~110 jobs per SM
			static const long long parts_below_per_sm = [] { const char* e = getenv("RT_B200_WAVE_PARTS_BELOW"); return e ? atoll(e) : 110ll; }();
			const long long parts_below = parts_below_per_sm * d.sm_count;
			const long long warp_tiles = tiles * rt::kSignalsPerTile;
			// the job counts of the wavefront frame before this one (whatever has arrived: a hint, the frame is the same
			// either way); before the first frame: a guess from the number of tiles
			const long long view_jobs_hint = d.h_wave_jobs[0] ? (long long)d.h_wave_jobs[0] : 2 * warp_tiles;
			const long long shadow_jobs_hint = d.h_wave_jobs[1] ? (long long)d.h_wave_jobs[1] : 4 * warp_tiles * std::max(ctx->n_lights, 1);
			d.wave.view_parts = view_jobs_hint < parts_below ? 1u : 0u;
			d.wave.shadow_parts = shadow_jobs_hint < parts_below ? 1u : 0u;
			d.wave.setup_per_light = warp_tiles <= parts_below ? 1u : 0u;
			// a job per (warp tile, subtree) and, for shadow rays, per light: sized for the worst case, so the lists cannot overflow
			const size_t view_jobs = (size_t)warp_tiles * (size_t)wave_subtrees;
			const size_t shadow_jobs = view_jobs * (size_t)std::max(ctx->n_lights, 1);
			if (pixels > d.wave_pixels || view_jobs > d.wave_view_tasks || shadow_jobs > d.wave_shadow_tasks || ctx->meshes.size() != d.wave_meshes || (size_t)ctx->n_lights != d.wave_lights)
			{
				RT_CUDA(ctx, cudaDeviceSynchronize());
				if (d.d_wave) RT_CUDA(ctx, cudaFree(d.d_wave));
				d.d_wave = nullptr;
				d.wave_pixels = d.wave_view_tasks = d.wave_shadow_tasks = 0;
				const size_t n_m = std::max<size_t>(ctx->meshes.size(), 1), n_l = (size_t)std::max(ctx->n_lights, 1);
				const size_t bytes = pixels * (8 + 16 + 4 + 8 * n_m + 8 * n_m * n_l) + (view_jobs + shadow_jobs) * 8 + 256;
				RT_CUDA(ctx, cudaMalloc(&d.d_wave, bytes));
				char* base = (char*)d.d_wave;
				d.wave.shadow_origin = (float4*)base; base += pixels * 16;
				d.wave.hit_key = (unsigned long long*)base; base += pixels * 8;
				d.wave.view_jobs = (uint2*)base; base += view_jobs * 8;
				d.wave.shadow_jobs = (uint2*)base; base += shadow_jobs * 8;
				d.wave.view_alive = (unsigned long long*)base; base += pixels * 8 * n_m;
				d.wave.shadow_alive = (unsigned long long*)base; base += pixels * 8 * n_m * n_l;
				d.wave.occluded = (unsigned int*)base; base += pixels * 4;
				d.wave.counters = (unsigned int*)(((uintptr_t)base + 15) & ~(uintptr_t)15);
				d.wave_pixels = pixels; d.wave_view_tasks = view_jobs; d.wave_shadow_tasks = shadow_jobs; d.wave_meshes = ctx->meshes.size(); d.wave_lights = (size_t)ctx->n_lights;
			}
			d.wave.view_capacity = (unsigned int)d.wave_view_tasks; d.wave.shadow_capacity = (unsigned int)d.wave_shadow_tasks;
			d.wave.jobs_report = d.d_wave_jobs;
			d.wave.split = d.d_split;
			d.wave.root_map = d.d_root_map;
			RT_CUDA(ctx, rt::wave_launch(d.view, p, d.wave, grid, d.sm_count, stream));
			ctx->timing.kernel_launches += rt::wave_launch_count(p.shadows && ctx->n_lights > 0) - 1;
		}
		else if (variant == RT_KERNEL_PACKED) pick_kernel_x2(p.lighting_mode, p.shadows, path == RT_MESH_PATH_BVH)<<<grid, rt::pick_threads_x2(), 0, stream>>>(d.view, p);
		else if (variant == RT_KERNEL_SCALAR) pick_kernel(p.lighting_mode, p.shadows, path == RT_MESH_PATH_BVH)<<<grid, rt::kThreads, rt::dynamic_smem_bytes(rt::kThreads, ctx->n_materials), stream>>>(d.view, p);
		else
		{
			// persistent warps: one wave of CTAs, never more than the launch has tiles; every launch takes the
			// next queue of the ring (a queue is all zeros again when its kernel ends, and launches that could
			// overlap - other streams - are far fewer than the ring is long)
			const KernelFn fn = persistent;
			tiles_to_render_first(ctx, p, n_strips, wave * (rt::kPersistentThreads / 32));
			int cost_slot = -1;
			const int prc = prepare_cell_order(ctx, d, p, stream, n_strips, wave * (rt::kPersistentThreads / 32), path, &cost_slot);
			if (prc != RT_OK) return prc;
			p.queue = d.d_queues + 2 * (d.queue_cursor++ % kQueueRing);
			const long long ctas_of_work = (tiles * rt::kSignalsPerTile + rt::kPersistentThreads / 32 - 1) / (rt::kPersistentThreads / 32);
			// band watcher: CTA 0 of the wave tells the host when a band is complete instead of rendering
			static const bool no_watcher = getenv("RT_B200_NO_BAND_WATCHER") != nullptr;
			const bool watcher = watched && p.host_flags && p.band_done && !p.band_local && !p.band_table && !no_watcher && wave > 8;
			if (!watcher) p.host_flags = nullptr;
			if (watched) *watched = watcher;
			fn<<<(unsigned)std::min<long long>(wave, ctas_of_work + (watcher ? 1 : 0)), rt::kPersistentThreads, rt::dynamic_smem_bytes(rt::kPersistentThreads, ctx->n_materials), stream>>>(d.view, p);
			RT_CUDA(ctx, cudaGetLastError());
			const int frc = finish_cell_order(ctx, d, p, stream, cost_slot);
			if (frc != RT_OK) return frc;
		}
		RT_CUDA(ctx, cudaGetLastError());
		if (stream != d.stream)
		{
			RT_CUDA(ctx, cudaEventRecord(d.ev_foreign, stream));
			d.foreign_pending = true;
		}
		ctx->timing.kernel_launches++;
		return RT_OK;
	}

	// Render the whole frame into device 0's frame buffer, split over all devices.
	int render_to_device0(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame)
	{
		int rc = validate_frame(ctx, camera, frame);
		if (rc != RT_OK) return rc;
		const size_t pixels = (size_t)frame->width * (size_t)frame->height;
		const int n = (int)ctx->devs.size();
		DeviceState& d0 = ctx->devs[0];
		if ((rc = ensure_frame(ctx, d0, pixels)) != RT_OK) return rc;
		ctx->timing = rt_timing{};
		ctx->last_width = frame->width; ctx->last_height = frame->height;

		const int total_strips = (frame->height + rt::kBlockH - 1) / rt::kBlockH;
		rt::FrameParams base = make_params(camera, frame);

		if (n == 1 || ctx->peer_stores)
		{
			// 8-row strips dealt round-robin; every device stores straight into device 0's frame
			// (for k > 0 those are peer stores over NVLink: render and gather are one kernel).
			for (int k = 0; k < n; ++k)
			{
				DeviceState& d = ctx->devs[k];
				RT_CUDA(ctx, cudaSetDevice(d.device));
				rt::FrameParams p = base;
				p.row_begin = 0; p.row_end = frame->height;
				p.strip_first = k; p.strip_step = n;
				p.dst_full_frame = 1; p.dst = d0.d_frame;
				const int strips = (total_strips - k + n - 1) / n;
				RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
				if ((rc = launch(ctx, d, p, d.stream, strips)) != RT_OK) return rc;
				RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
			}
			RT_CUDA(ctx, cudaSetDevice(d0.device));
			for (int k = 1; k < n; ++k) RT_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->devs[k].ev_kernel, 0));
			RT_CUDA(ctx, cudaEventRecord(ctx->ev_gather, d0.stream));
		}
		else
		{
			// No peer mapping: contiguous bands rendered locally, then one peer copy per band.
			const int band_strips = (total_strips + n - 1) / n;
			for (int k = 0; k < n; ++k)
			{
				DeviceState& d = ctx->devs[k];
				const int row_begin = std::min(frame->height, k * band_strips * rt::kBlockH);
				const int row_end = std::min(frame->height, (k + 1) * band_strips * rt::kBlockH);
				const int rows = row_end - row_begin;
				RT_CUDA(ctx, cudaSetDevice(d.device));
				RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
				if (rows > 0)
				{
					rt::FrameParams p = base;
					p.row_begin = row_begin; p.row_end = row_end; p.strip_first = 0; p.strip_step = 1;
					if (k == 0) { p.dst_full_frame = 1; p.dst = d0.d_frame; }
					else
					{
						if ((rc = ensure_frame(ctx, d, (size_t)rows * frame->width)) != RT_OK) return rc;
						p.dst_full_frame = 0; p.dst = d.d_frame;
					}
					if ((rc = launch(ctx, d, p, d.stream, (rows + rt::kBlockH - 1) / rt::kBlockH)) != RT_OK) return rc;
				}
				RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
				if (k > 0 && rows > 0)
					RT_CUDA(ctx, cudaMemcpyPeerAsync(d0.d_frame + (size_t)row_begin * frame->width, d0.device, d.d_frame, d.device,
					                                 (size_t)rows * frame->width * sizeof(uint32_t), d.stream));
				RT_CUDA(ctx, cudaEventRecord(d.ev_done, d.stream));
			}
			RT_CUDA(ctx, cudaSetDevice(d0.device));
			for (int k = 1; k < n; ++k) RT_CUDA(ctx, cudaStreamWaitEvent(d0.stream, ctx->devs[k].ev_done, 0));
			RT_CUDA(ctx, cudaEventRecord(ctx->ev_gather, d0.stream));
		}
		return RT_OK;
	}

	int collect_timing(rt_context* ctx, bool with_d2h)
	{
		DeviceState& d0 = ctx->devs[0];
		float kernel_ms = 0.f;
		for (DeviceState& d : ctx->devs)
		{
			RT_CUDA(ctx, cudaSetDevice(d.device));
			RT_CUDA(ctx, cudaEventSynchronize(d.ev_kernel));
			float ms = 0.f;
			RT_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_kernel));
			kernel_ms = std::max(kernel_ms, ms);
		}
		RT_CUDA(ctx, cudaSetDevice(d0.device));
		RT_CUDA(ctx, cudaEventSynchronize(ctx->ev_gather));
		float to_gather = 0.f, to_end = 0.f;
		RT_CUDA(ctx, cudaEventElapsedTime(&to_gather, d0.ev_begin, ctx->ev_gather));
		ctx->timing.kernel_ms = kernel_ms;
		ctx->timing.gather_ms = std::max(0.f, to_gather - kernel_ms);
		if (with_d2h)
		{
			RT_CUDA(ctx, cudaEventSynchronize(ctx->ev_d2h));
			RT_CUDA(ctx, cudaEventElapsedTime(&to_end, d0.ev_begin, ctx->ev_d2h));
			ctx->timing.d2h_ms = to_end - to_gather;
			ctx->timing.total_ms = to_end;
		}
		else
		{
			ctx->timing.d2h_ms = 0.f;
			ctx->timing.total_ms = to_gather;
		}
		return RT_OK;
	}

	// Make `host` a legal target for an asynchronous device-to-host copy.  Returns the pointer to copy into through
	// `target`: host itself when CUDA already knows the range as pinned (cudaHostAlloc / cudaHostRegister by the caller,
	// rt_register_surface, managed memory), else the library's own pinned bounce buffer.  The library never pins caller
	// memory behind the caller's back: a registration outliving the buffer (free + a new mapping at the same address)
	// would send later copies through stale pages.
	int prepare_host(rt_context* ctx, void* host, size_t bytes, void** target)
	{
		static const bool force_bounce = getenv("RT_B200_FORCE_STAGING") != nullptr;     // tests: take the bounce-buffer path
		if (!force_bounce)
		{
			// both ends of the span must be pinned (a surface larger than its registration falls back to the bounce buffer)
			cudaPointerAttributes first{}, last{};
			const cudaError_t e0 = cudaPointerGetAttributes(&first, host);
			const cudaError_t e1 = cudaPointerGetAttributes(&last, (char*)host + (bytes ? bytes - 1 : 0));
			cudaGetLastError();
			auto pinned = [](cudaError_t e, const cudaPointerAttributes& a) { return e == cudaSuccess && (a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged); };
			if (pinned(e0, first) && pinned(e1, last)) { *target = host; return RT_OK; }
		}
		if (ctx->staging_bytes < bytes)
		{
			if (ctx->staging) cudaFreeHost(ctx->staging);
			ctx->staging = nullptr; ctx->staging_bytes = 0;
			RT_CUDA(ctx, cudaHostAlloc(&ctx->staging, bytes, cudaHostAllocPortable));
			ctx->staging_bytes = bytes;
		}
		*target = ctx->staging;
		return RT_OK;
	}

	int download(rt_context* ctx, uint32_t* host_dst, int32_t pitch_bytes, bool record_event)
	{
		DeviceState& d0 = ctx->devs[0];
		const int W = ctx->last_width, H = ctx->last_height;
		if (W <= 0 || H <= 0) return fail(ctx, RT_ERR_BAD_STATE, "no frame has been rendered yet");
		if (!host_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "host_dst must not be NULL");
		if (pitch_bytes < 4 * W || (pitch_bytes & 3)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "pitch_bytes %d too small or unaligned for width %d", pitch_bytes, W);
		const size_t span = (size_t)pitch_bytes * (size_t)(H - 1) + (size_t)W * 4u;
		void* target = nullptr;
		int rc = prepare_host(ctx, host_dst, span, &target);
		if (rc != RT_OK) return rc;
		RT_CUDA(ctx, cudaSetDevice(d0.device));
		if (pitch_bytes == 4 * W)
			RT_CUDA(ctx, cudaMemcpyAsync(target, d0.d_frame, (size_t)W * H * 4u, cudaMemcpyDeviceToHost, d0.stream));
		else
			RT_CUDA(ctx, cudaMemcpy2DAsync(target, (size_t)pitch_bytes, d0.d_frame, (size_t)W * 4u, (size_t)W * 4u, (size_t)H, cudaMemcpyDeviceToHost, d0.stream));
		if (record_event) RT_CUDA(ctx, cudaEventRecord(ctx->ev_d2h, d0.stream));
		RT_CUDA(ctx, cudaStreamSynchronize(d0.stream));
		if (target != host_dst)
		{
			if (pitch_bytes == 4 * W) memcpy(host_dst, target, (size_t)W * H * 4u);
			else for (int y = 0; y < H; ++y) memcpy((char*)host_dst + (size_t)y * pitch_bytes, (char*)target + (size_t)y * pitch_bytes, (size_t)W * 4u);
		}
		return RT_OK;
	}
}

namespace
{
	// Turns the reference's BVHNode array into the threaded layout of rt::BvhLink.  Every index is
	// validated: a malformed tree is an error, never a hang or an out-of-bounds read on the GPU.
	int thread_bvh(rt_context* ctx, const rt_mesh_desc* mesh, std::vector<float4>& out)
	{
		out.clear();
		const int32_t n = mesh->bvh_node_count;
		if (!mesh->bvh_nodes || n <= 0 || mesh->triangle_count == 0) return RT_OK;
		if (n > rt::BvhLink::kMaxNodes) return fail(ctx, RT_ERR_CAPACITY, "%d BVH nodes exceed the capacity of %d", n, rt::BvhLink::kMaxNodes);
		if (mesh->triangle_count > rt::BvhLink::kFirstMask) return fail(ctx, RT_ERR_CAPACITY, "%d triangles exceed what a leaf link can address (%d)", mesh->triangle_count, rt::BvhLink::kFirstMask);
		out.assign(2 * (size_t)n, make_float4(0.f, 0.f, 0.f, 0.f));
		std::vector<char> seen((size_t)n, 0);
		struct Item { int32_t node, escape; };
		std::vector<Item> stack;
		stack.push_back({ 0, -1 });
		int64_t covered = 0;
		while (!stack.empty())
		{
			const Item it = stack.back();
			stack.pop_back();
			if (it.node < 0 || it.node >= n) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "BVH child index %d outside [0, %d)", it.node, n);
			if (seen[it.node]) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "BVH node %d is reachable twice", it.node);
			seen[it.node] = 1;
			const rt_bvh_node& nd = mesh->bvh_nodes[it.node];
			int32_t first, leaf_tris = 0;
			if (nd.idx_count > 0)
			{
				if (nd.idx_count % 3 || nd.first_idx % 3 || (uint64_t)nd.first_idx + nd.idx_count > 3ull * (uint64_t)mesh->triangle_count)
					return fail(ctx, RT_ERR_INVALID_ARGUMENT, "BVH leaf %d owns indices [%u, %u) outside the mesh", it.node, nd.first_idx, nd.first_idx + nd.idx_count);
				leaf_tris = (int32_t)(nd.idx_count / 3);
				if (leaf_tris > rt::BvhLink::kMaxLeafTriangles) return fail(ctx, RT_ERR_CAPACITY, "BVH leaf with %d triangles exceeds the capacity of %d", leaf_tris, rt::BvhLink::kMaxLeafTriangles);
				first = (int32_t)(nd.first_idx / 3);
				covered += leaf_tris;
			}
			else
			{
				first = (int32_t)nd.left_node;
				// recursion order of Utils.h:285-286: left, then left + 1; left's subtree is followed by left + 1
				stack.push_back({ first + 1, it.escape });
				stack.push_back({ first, first + 1 });
			}
			out[2 * (size_t)it.node + 0] = make_float4(nd.min_aabb[0], nd.max_aabb[0], nd.min_aabb[1], nd.max_aabb[1]);
			out[2 * (size_t)it.node + 1] = make_float4(nd.min_aabb[2], nd.max_aabb[2],
			                                            bits_as_float(leaf_tris > 0 ? rt::BvhLink::leaf(first, leaf_tris) : rt::BvhLink::inner(first)),
			                                            bits_as_float(rt::BvhLink::miss(it.escape)));
		}
		if (covered != mesh->triangle_count) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "BVH leaves cover %lld of %d triangles", (long long)covered, mesh->triangle_count);

		return RT_OK;
	}
}

namespace
{
	// RT_KERNEL_WAVEFRONT: cut a host-uploaded tree into at most kMaxSubtrees subtrees of comparable size and every subtree
	// into at most kFine parts (rt_wave_params.h); a part is walked from its root until the walk leaves through the root's
	// escape link.  Reads the threaded records (thread_bvh validated them); built when a launch first wants the tables
	// (ensure_splits): a mesh that is uploaded anew every frame and never rendered this way does not pay for them, and
	// one that is (an animated Scene_W4_OptionalScene) pays two linear passes over its nodes - this runs between the
	// caller's Render() and the first kernel of the frame.
	void build_split(HostMesh& hm)
	{
		using namespace rt::wave;
		hm.split.clear();
		hm.root_map.clear();
		const int32_t n = (int32_t)(hm.nodes.size() / 2);
		if (n <= 0 || hm.device_bvh) return;
		auto hit_of = [&](int32_t node) { return float_bits(hm.nodes[2 * (size_t)node + 1].z); };
		auto miss_of = [&](int32_t node) { return float_bits(hm.nodes[2 * (size_t)node + 1].w); };
		auto left_of = [&](int32_t node) { return hit_of(node) / rt::BvhLink::kNodeBytes; };
		auto is_leaf = [&](int32_t node) { return rt::BvhLink::is_leaf(hit_of(node)); };
		// Children sit behind their parent in the reference's array (nodesUsed only grows, DataTypes.h:372-373): one pass
		// down marks what the root reaches, one pass up adds up what every subtree spans - nodes below it (count, lowest
		// and highest index) and triangles (count, first, last).  A tree in another order gets no tables (the wavefront
		// form is then not offered for the scene).
		thread_local std::vector<int32_t> scratch;
		scratch.assign(6 * (size_t)n, 0);
		int32_t* size = scratch.data(), * node_lo = size + n, * node_hi = node_lo + n, * tri_lo = node_hi + n, * tri_hi = tri_lo + n, * tri_sum = tri_hi + n;
		size[0] = 1;
		for (int32_t i = 0; i < n; ++i)
		{
			if (!size[i] || is_leaf(i)) continue;
			const int32_t left = left_of(i);
			if (left <= i || left + 1 >= n) return;
			size[left] = size[left + 1] = 1;
		}
		for (int32_t i = n; i-- > 0;)
		{
			if (!size[i]) continue;
			const int hit = hit_of(i);
			if (rt::BvhLink::is_leaf(hit))
			{
				tri_lo[i] = rt::BvhLink::leaf_first(hit);
				tri_sum[i] = rt::BvhLink::leaf_count(hit);
				tri_hi[i] = tri_lo[i] + tri_sum[i] - 1;
				node_lo[i] = INT32_MAX; node_hi[i] = -1;
				continue;
			}
			const int32_t l = hit / rt::BvhLink::kNodeBytes, r = l + 1;
			size[i] = 1 + size[l] + size[r];
			node_lo[i] = std::min(l, std::min(node_lo[l], node_lo[r]));
			node_hi[i] = std::max(r, std::max(node_hi[l], node_hi[r]));
			tri_lo[i] = std::min(tri_lo[l], tri_lo[r]);
			tri_hi[i] = std::max(tri_hi[l], tri_hi[r]);
			tri_sum[i] = tri_sum[l] + tri_sum[r];
		}
		// cut(root, pieces, depth): start from `root` and keep replacing the largest inner piece by its two children (each
		// remembers the nodes between `root` and itself, as far as a part's record has room) until there are `pieces` or
		// nothing is left to cut
		struct Piece { int32_t node, depth, above[kFineAncestors]; };
		auto cut = [&](int32_t root, int pieces, int depth, Piece* out)
		{
			int count = 1;
			out[0] = Piece{ root, 0, {} };
			while (count < pieces)
			{
				int pick = -1;
				for (int e = 0; e < count; ++e)
				{
					if (is_leaf(out[e].node) || out[e].depth >= depth) continue;
					if (pick < 0 || size[out[e].node] > size[out[pick].node]) pick = e;
				}
				if (pick < 0) break;
				Piece child = out[pick];
				if (child.depth < kFineAncestors) child.above[child.depth] = child.node;
				++child.depth;
				child.node = left_of(out[pick].node);
				for (int e = count; e > pick + 1; --e) out[e] = out[e - 1];
				out[pick] = child;
				++child.node;
				out[pick + 1] = child;
				++count;
			}
			return count;
		};
		static const int cap = [] { const char* e = getenv("RT_B200_WAVE_SUBTREES"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= kMaxSubtrees) ? v : kMaxSubtrees; }();
		static const int fine_cap = [] { const char* e = getenv("RT_B200_WAVE_PARTS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= kFine) ? v : kFine; }();
		Piece subtrees[kMaxSubtrees], parts[kFine];
		const int n_subtrees = cut(0, n >= 64 ? cap : 1, INT32_MAX, subtrees);        // a shallow tree is one job
		hm.split.assign((size_t)kSplitStride, 0);
		hm.split[0] = n_subtrees;
		hm.root_map.assign((size_t)n, 0);
		for (int e = 0; e < n_subtrees; ++e) hm.root_map[(size_t)subtrees[e].node] = (uint8_t)(e + 1);
		for (int e = 0; e < n_subtrees; ++e)
		{
			const int n_parts = cut(subtrees[e].node, fine_cap, kFineAncestors, parts);
			for (int f = 0; f <= kFine; ++f)
			{
				if (f < kFine && f >= n_parts) continue;
				const Piece piece = f < kFine ? parts[f] : Piece{ subtrees[e].node, 0, {} };
				const int32_t node = piece.node;
				int32_t* rec = hm.split.data() + kSplitHeader + ((size_t)e * (kFine + 1) + (size_t)f) * kSplitWords;
				rec[0] = node * rt::BvhLink::kNodeBytes;
				rec[1] = miss_of(node);
				const int32_t below = size[node] - 1, tris = tri_sum[node];
				const bool nodes_contiguous = below == 0 || (node_lo[node] == left_of(node) && node_hi[node] - node_lo[node] + 1 == below);
				const bool tris_contiguous = tri_hi[node] - tri_lo[node] + 1 == tris;
				rec[2] = below ? node_lo[node] * rt::BvhLink::kNodeBytes : 0;
				rec[3] = below;
				rec[4] = tri_lo[node];
				rec[5] = tris;
				rec[6] = kPartPresent | ((nodes_contiguous && tris_contiguous && (size_t)(1 + below) * 32 + (size_t)tris * 48 <= (size_t)kStageBytes) ? kPartStageable : 0);
				rec[7] = piece.depth;
				for (int k = 0; k < piece.depth; ++k) rec[8 + k] = piece.above[k] * rt::BvhLink::kNodeBytes;
			}
		}
	}
}

namespace
{
	// cuStreamWaitValue32 through the runtime's driver entry point lookup: no link-time dependency on
	// libcuda, so the library still loads on a box without a driver.
	typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
	WaitValue32Fn wait_value32()
	{
		static const WaitValue32Fn fn = []() -> WaitValue32Fn {
			if (const char* e = getenv("RT_B200_NO_STREAM_WAIT")) if (atoi(e)) return nullptr;
			void* p = nullptr;
			cudaDriverEntryPointQueryResult st = cudaDriverEntryPointSymbolNotFound;
			if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
			return reinterpret_cast<WaitValue32Fn>(p);
		}();
		return fn;
	}

	// ---- the bands of a progressive present -------------------------------------------------------------------------
	// The copy stream runs  wait(band b complete) -> copy(band b)  band after band.  Measured on one B200 (4K frame,
	// tools/share_e2e.py with RT_B200_PIPELINE_BANDS): with equal bands the frame is in host memory 0.149 / 0.103 / 0.086 /
	// 0.16 ms after the kernel for 4 / 8 / 16 / 32 bands - the last band's bytes against ~9 us of fixed cost per band.
	// Bands of unequal size (small at both ends, RT_B200_BAND_PROFILE=1) were tried and lost at N = 1 (0.826 vs 0.792 ms
	// kernel + copies with 12 vs 16 bands) and won little on an eighth of the frame (0.172 vs 0.182 ms): equal bands stay
	// the default, the table-driven form stays for experiments.
	struct BandSchedule
	{
		int bands = 0;
		int first[65] = {};      // band b = strips [first[b], first[b + 1])
	};
	BandSchedule make_band_schedule(int total_strips, int wanted)
	{
		BandSchedule s;
		const int n = std::max(1, std::min(std::min(wanted, 64), total_strips));
		static const bool equal = getenv("RT_B200_BAND_PROFILE") == nullptr;          // default: equal bands
		double weight[64], sum = 0.0;
		for (int b = 0; b < n; ++b)
		{
			// a raised sine over the band index: ends ~1/8 of the middle
			const double x = std::sin(3.14159265358979323846 * (b + 0.5) / n);
			weight[b] = equal ? 1.0 : 0.12 + x * x;
			sum += weight[b];
		}
		double acc = 0.0;
		s.first[0] = 0;
		int made = 0;
		for (int b = 0; b < n; ++b)
		{
			acc += weight[b];
			int end = (b == n - 1) ? total_strips : (int)std::lround(acc / sum * total_strips);
			end = std::max(end, s.first[made] + 1);
			end = std::min(end, total_strips);
			if (end <= s.first[made]) continue;
			s.first[++made] = end;
			if (end == total_strips) break;
		}
		if (s.first[made] != total_strips) s.first[made] = total_strips;
		s.bands = made;
		return s;
	}

	// Makes the device's strip -> band table the one of `s` (a copy on the device's stream, ordered before the kernel).
	int use_band_schedule(rt_context* ctx, DeviceState& d, const BandSchedule& s, int total_strips, rt::FrameParams& p)
	{
		if (total_strips > 8192 || s.bands > 64) { p.band_table = nullptr; return RT_OK; }      // callers then fall back to equal bands
		if (d.band_table_strips != total_strips || d.band_table_bands != s.bands)
		{
			RT_CUDA(ctx, cudaStreamSynchronize(d.stream));                // the pinned source may still be in flight
			for (int b = 0; b < s.bands; ++b)
				for (int k = s.first[b]; k < s.first[b + 1]; ++k) d.h_band_table[k] = (uint8_t)b;
			RT_CUDA(ctx, cudaMemcpyAsync(d.d_band_table, d.h_band_table, (size_t)total_strips, cudaMemcpyHostToDevice, d.stream));
			d.band_table_strips = total_strips; d.band_table_bands = s.bands;
		}
		p.band_table = d.d_band_table;
		return RT_OK;
	}

	// One device: rt_render overlaps the present copy with the rendering.  The frame is cut into bands
	// of strips.  Preferred form ("progressive present"): ONE kernel launch; every CTA bumps its band's
	// counter when its pixels are in memory, and the copy stream waits on each counter with
	// cuStreamWaitValue32 before it sends that band to the host surface - no launch boundaries, no
	// tail per band.  Fallback (no stream memory operations): one launch per band + events.
	int render_pipelined(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame, uint32_t* host_dst, int32_t pitch_bytes)
	{
		int rc = validate_frame(ctx, camera, frame);
		if (rc != RT_OK) return rc;
		const int W = frame->width, H = frame->height;
		if (pitch_bytes < 4 * W || (pitch_bytes & 3)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "pitch_bytes %d too small or unaligned for width %d", pitch_bytes, W);
		DeviceState& d = ctx->devs[0];
		if ((rc = ensure_frame(ctx, d, (size_t)W * (size_t)H)) != RT_OK) return rc;
		const size_t span = (size_t)pitch_bytes * (size_t)(H - 1) + (size_t)W * 4u;
		void* target = nullptr;
		if ((rc = prepare_host(ctx, host_dst, span, &target)) != RT_OK) return rc;
		ctx->timing = rt_timing{};
		ctx->last_width = W; ctx->last_height = H;

		const int n_dev = (int)ctx->devs.size();
		const WaitValue32Fn wait = (n_dev == 1 || ctx->peer_stores) ? wait_value32() : nullptr;
		static const int requested = [] { const char* e = getenv("RT_B200_PIPELINE_BANDS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 64) ? v : 0; }();
		const int total_strips = (H + rt::kBlockH - 1) / rt::kBlockH;
		const int grid_x = (W + rt::kBlockW - 1) / rt::kBlockW;
		// a band should hold enough CTAs to be worth a copy of its own
		const int max_bands = std::max(1, (int)(((long long)total_strips * grid_x) / 2048));
		const int wanted = requested ? requested : (wait ? 16 : 4);
		const int bands = std::min(std::min(std::min(wanted, wait ? 64 : 16), max_bands), total_strips);
		const int strips_per_band = (total_strips + bands - 1) / bands;
		// one device: bands of unequal size (make_band_schedule); the multi-device gather flow keeps equal bands
		// (its per-GPU share arithmetic in signal_band_done assumes them)
		const BandSchedule schedule = make_band_schedule(total_strips, bands);

		RT_CUDA(ctx, cudaSetDevice(d.device));
		rt::FrameParams base = make_params(camera, frame);
		static const bool profile = getenv("RT_B200_BAND_PROFILE") != nullptr;
		if (profile && wait && n_dev == 1 && (rc = use_band_schedule(ctx, d, schedule, total_strips, base)) != RT_OK) return rc;
		const bool scheduled = base.band_table != nullptr;
		auto copy_band = [&](int r0, int r1) -> int
		{
			char* dst = (char*)target + (size_t)r0 * (size_t)pitch_bytes;
			const uint32_t* src = d.d_frame + (size_t)r0 * (size_t)W;
			if (pitch_bytes == 4 * W)
				RT_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)(r1 - r0) * W * 4u, cudaMemcpyDeviceToHost, d.copy_stream));
			else
				RT_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)pitch_bytes, src, (size_t)W * 4u, (size_t)W * 4u, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, d.copy_stream));
			return RT_OK;
		};

		if (wait)
		{
			RT_CUDA(ctx, cudaMemsetAsync(d.d_band_done, 0, sizeof(unsigned int) * 64, d.stream));
			RT_CUDA(ctx, cudaEventRecord(d.ev_band[0], d.stream));
			RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_band[0], 0));      // counters are zero before anyone polls them
			// every device renders its strips (dealt round-robin) straight into device 0's frame and bumps
			// device 0's band counters; with one device this is the plain single-launch case
			for (int k = 0; k < n_dev; ++k)
			{
				DeviceState& dk = ctx->devs[k];
				RT_CUDA(ctx, cudaSetDevice(dk.device));
				if (k > 0) RT_CUDA(ctx, cudaStreamWaitEvent(dk.stream, d.ev_band[0], 0));   // ... nor bumps them
				RT_CUDA(ctx, cudaEventRecord(dk.ev_begin, dk.stream));
				rt::FrameParams p = base;
				p.row_begin = 0; p.row_end = H; p.strip_first = k; p.strip_step = n_dev; p.dst_full_frame = 1; p.dst = d.d_frame;
				p.band_done = d.d_band_done; p.strips_per_band = strips_per_band;
				p.band_local = n_dev > 1 ? dk.d_band_done + 64 : nullptr;      // second half of each device's own array
				if ((rc = launch(ctx, dk, p, dk.stream, (total_strips - k + n_dev - 1) / n_dev)) != RT_OK) return rc;
				RT_CUDA(ctx, cudaEventRecord(dk.ev_kernel, dk.stream));
			}
			RT_CUDA(ctx, cudaSetDevice(d.device));
			for (int k = 1; k < n_dev; ++k) RT_CUDA(ctx, cudaStreamWaitEvent(d.stream, ctx->devs[k].ev_kernel, 0));
			for (int b = 0; b < (scheduled ? schedule.bands : bands); ++b)
			{
				const int s0 = scheduled ? schedule.first[b] : b * strips_per_band, s1 = scheduled ? schedule.first[b + 1] : std::min(total_strips, (b + 1) * strips_per_band);
				if (s1 <= s0) break;
				// the last band is complete when the kernel is: an event wait reacts faster than a polled counter
				if (s1 == total_strips && n_dev == 1) RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_kernel, 0));
				else
				{
					const CUresult cr = wait((CUstream)d.copy_stream, (CUdeviceptr)(uintptr_t)(d.d_band_done + b), (cuuint32_t)((s1 - s0) * grid_x * rt::kSignalsPerTile), CU_STREAM_WAIT_VALUE_GEQ);
					if (cr != CUDA_SUCCESS) return fail(ctx, RT_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)cr);
				}
				if ((rc = copy_band(s0 * rt::kBlockH, std::min(H, s1 * rt::kBlockH))) != RT_OK) return rc;
			}
		}
		else
		{
			RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
			for (int b = 0; b < bands; ++b)
			{
				const int r0 = std::min(H, b * strips_per_band * rt::kBlockH), r1 = std::min(H, (b + 1) * strips_per_band * rt::kBlockH);
				if (r1 <= r0) break;
				rt::FrameParams p = base;
				p.row_begin = r0; p.row_end = r1; p.strip_first = 0; p.strip_step = 1; p.dst_full_frame = 1; p.dst = d.d_frame;
				if ((rc = launch(ctx, d, p, d.stream, (r1 - r0 + rt::kBlockH - 1) / rt::kBlockH)) != RT_OK) return rc;
				RT_CUDA(ctx, cudaEventRecord(d.ev_band[b], d.stream));
				RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_band[b], 0));
				if ((rc = copy_band(r0, r1)) != RT_OK) return rc;
			}
			RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
		}
		RT_CUDA(ctx, cudaEventRecord(ctx->ev_gather, d.stream));
		RT_CUDA(ctx, cudaEventRecord(ctx->ev_d2h, d.copy_stream));
		RT_CUDA(ctx, cudaStreamSynchronize(d.copy_stream));
		RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
		if (target != host_dst)
		{
			if (pitch_bytes == 4 * W) memcpy(host_dst, target, (size_t)W * H * 4u);
			else for (int y = 0; y < H; ++y) memcpy((char*)host_dst + (size_t)y * pitch_bytes, (char*)target + (size_t)y * pitch_bytes, (size_t)W * 4u);
		}
		return collect_timing(ctx, true);
	}

	// ---- direct present: every device copies the strips it rendered straight into the host surface ----------------
	// The gather-to-device-0 flow moves the whole frame through ONE PCIe link.  When the destination is host memory
	// anyway, each device presents its own strips over its own link: device k renders strips k, k + n, ... into its
	// own frame buffer (full-frame layout) and its copy stream follows the kernel band by band (the single-device
	// counters of signal_band_done; one strided 2-D copy per band).  No peer traffic, no collective.

	// Strips first, first + step, ... (count of them) of device d's frame buffer -> host target.
	int copy_strips_to_host(rt_context* ctx, DeviceState& d, int W, int H, int first, int step, int count, void* target, int32_t pitch_bytes)
	{
		if (count <= 0) return RT_OK;
		const int total_strips = (H + rt::kBlockH - 1) / rt::kBlockH;
		const int last = first + (count - 1) * step;
		int full = count;
		const int tail_rows = H - (total_strips - 1) * rt::kBlockH;          // rows of the frame's last strip
		if (last == total_strips - 1 && tail_rows != rt::kBlockH) --full;
		const size_t row_bytes = (size_t)W * 4u;
		if (pitch_bytes == 4 * W && step == 1)
		{
			// one device's consecutive strips in a tightly packed surface: one contiguous range (incl. a short last strip)
			const size_t rows = (size_t)full * rt::kBlockH + (full < count ? (size_t)tail_rows : 0u);
			RT_CUDA(ctx, cudaMemcpyAsync((char*)target + (size_t)first * rt::kBlockH * row_bytes, d.d_frame + (size_t)first * rt::kBlockH * W, rows * row_bytes, cudaMemcpyDeviceToHost, d.copy_stream));
			return RT_OK;
		}
		if (pitch_bytes == 4 * W)
		{
			if (full > 0)
				RT_CUDA(ctx, cudaMemcpy2DAsync((char*)target + (size_t)first * rt::kBlockH * row_bytes, (size_t)step * rt::kBlockH * row_bytes,
				                               d.d_frame + (size_t)first * rt::kBlockH * W, (size_t)step * rt::kBlockH * row_bytes,
				                               (size_t)rt::kBlockH * row_bytes, (size_t)full, cudaMemcpyDeviceToHost, d.copy_stream));
		}
		else
		{
			for (int j = 0; j < full; ++j)
			{
				const size_t r0 = (size_t)(first + j * step) * rt::kBlockH;
				RT_CUDA(ctx, cudaMemcpy2DAsync((char*)target + r0 * (size_t)pitch_bytes, (size_t)pitch_bytes, d.d_frame + r0 * W, row_bytes, row_bytes,
				                               (size_t)rt::kBlockH, cudaMemcpyDeviceToHost, d.copy_stream));
			}
		}
		if (full < count)
		{
			const size_t r0 = (size_t)last * rt::kBlockH;
			RT_CUDA(ctx, cudaMemcpy2DAsync((char*)target + r0 * (size_t)pitch_bytes, (size_t)pitch_bytes, d.d_frame + r0 * W, row_bytes, row_bytes,
			                               (size_t)tail_rows, cudaMemcpyDeviceToHost, d.copy_stream));
		}
		return RT_OK;
	}

	// The host side of the band watcher: issue the copies of a device's pending frame for every band whose flag has
	// arrived.  Returns through `done` whether all of the device's bands have been issued.  `block` = spin until the next
	// band is ready (single device) instead of returning when it is not.
	int service_present(rt_context* ctx, DeviceState& d, bool block, bool* done)
	{
		DeviceState::PendingPresent& pp = d.pending;
		*done = !pp.active;
		if (!pp.active) return RT_OK;
		bool current = false;            // cudaSetDevice only when there is something to issue: polling the flags needs no device
		while (pp.next < pp.bands)
		{
			const int b = pp.next;
			const int s0 = b * pp.strips_per_band, s1 = std::min(pp.total_strips, (b + 1) * pp.strips_per_band);
			if (s1 <= s0) { pp.next = pp.bands; break; }
			if (__atomic_load_n(d.h_flags + b, __ATOMIC_ACQUIRE) != d.watch_tag)
			{
				// a kernel that ended without the flag (launch failure, fault) must not hang the caller
				if ((++pp.spins & 0xfffu) == 0)
				{
					if (!current) { RT_CUDA(ctx, cudaSetDevice(d.device)); current = true; }
					const cudaError_t q = cudaStreamQuery(d.stream);
					if (q != cudaErrorNotReady)
					{
						if (q != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "the pixel kernel failed: %s", cudaGetErrorString(q));
						if (__atomic_load_n(d.h_flags + b, __ATOMIC_ACQUIRE) != d.watch_tag) return fail(ctx, RT_ERR_CUDA, "the pixel kernel ended without reporting band %d", b);
					}
				}
				if (!block) return RT_OK;
#if defined(__x86_64__)
				__builtin_ia32_pause();
#endif
				continue;
			}
			const int first_mine = s0 + ((pp.strip_first - s0) % pp.strip_step + pp.strip_step) % pp.strip_step;
			const int mine = first_mine < s1 ? (s1 - 1 - first_mine) / pp.strip_step + 1 : 0;
			if (mine > 0)
			{
				if (!current) { RT_CUDA(ctx, cudaSetDevice(d.device)); current = true; }
				const int rc = copy_strips_to_host(ctx, d, pp.W, pp.H, first_mine, pp.strip_step, mine, pp.target, pp.pitch_bytes);
				if (rc != RT_OK) return rc;
			}
			++pp.next;
		}
		pp.active = false;
		*done = true;
		if (!current) RT_CUDA(ctx, cudaSetDevice(d.device));
		RT_CUDA(ctx, cudaEventRecord(d.ev_done, d.copy_stream));
		return RT_OK;
	}

	// Enqueues kernel + copies of one device's share; the caller synchronises d.copy_stream (ev_done is its last event).
	// With a band watcher in the launch the copies are NOT enqueued here: d.pending says what the host has to issue as the
	// bands arrive (service_present).
	int enqueue_direct_present(rt_context* ctx, DeviceState& d, const rt::FrameParams& base, int strip_first, int strip_step, void* target, int32_t pitch_bytes)
	{
		const int W = base.width, H = base.height;
		int rc = ensure_frame(ctx, d, (size_t)W * (size_t)H);
		if (rc != RT_OK) return rc;
		const int total_strips = (H + rt::kBlockH - 1) / rt::kBlockH;
		const int grid_x = (W + rt::kBlockW - 1) / rt::kBlockW;
		const int my_strips = strip_first < total_strips ? (total_strips - strip_first + strip_step - 1) / strip_step : 0;
		const WaitValue32Fn wait = wait_value32();
		static const int requested = [] { const char* e = getenv("RT_B200_PIPELINE_BANDS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 64) ? v : 0; }();
		// a band should hold enough of THIS device's CTAs to be worth a copy of its own
		const int max_bands = std::max(1, (int)(((long long)my_strips * grid_x) / 512));
		// 32 bands when the host issues the copies as the watcher reports them (a finer pipeline: the frame's last copy is a
		// thirty-second of it; 24-32 measured best, 48 and up pay per copy), 16 when the copy stream waits on the counters itself
		static const bool host_issues = getenv("RT_B200_NO_BAND_WATCHER") == nullptr;
		const int bands = std::max(1, std::min(std::min(requested ? requested : (host_issues ? 32 : 16), max_bands), total_strips));
		const int strips_per_band = (total_strips + bands - 1) / bands;
		const BandSchedule schedule = make_band_schedule(total_strips, bands);

		RT_CUDA(ctx, cudaSetDevice(d.device));
		rt::FrameParams p = base;
		p.row_begin = 0; p.row_end = H; p.strip_first = strip_first; p.strip_step = strip_step; p.dst_full_frame = 1; p.dst = d.d_frame;
		static const bool profile = getenv("RT_B200_BAND_PROFILE") != nullptr;
		if (profile && wait && my_strips > 0 && (rc = use_band_schedule(ctx, d, schedule, total_strips, p)) != RT_OK) return rc;
		const bool scheduled = p.band_table != nullptr;
		if (wait && my_strips > 0)
		{
			RT_CUDA(ctx, cudaMemsetAsync(d.d_band_done, 0, sizeof(unsigned int) * 64, d.stream));
			RT_CUDA(ctx, cudaEventRecord(d.ev_band[0], d.stream));
			RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_band[0], 0));      // counters are zero before anyone polls them
			RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
			p.band_done = d.d_band_done; p.strips_per_band = strips_per_band; p.band_local = nullptr;
			p.host_flags = d.d_flags; p.watch_tag = ++d.watch_tag; p.watch_bands = bands;
			bool watched = false;
			if ((rc = launch(ctx, d, p, d.stream, my_strips, &watched)) != RT_OK) return rc;
			RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
			if (watched)
			{
				DeviceState::PendingPresent& pp = d.pending;
				pp.active = true; pp.spins = 0; pp.bands = bands; pp.next = 0; pp.strips_per_band = strips_per_band; pp.strip_first = strip_first; pp.strip_step = strip_step;
				pp.total_strips = total_strips; pp.W = W; pp.H = H; pp.target = target; pp.pitch_bytes = pitch_bytes;
				return RT_OK;
			}
			for (int b = 0; b < (scheduled ? schedule.bands : bands); ++b)
			{
				const int s0 = scheduled ? schedule.first[b] : b * strips_per_band, s1 = scheduled ? schedule.first[b + 1] : std::min(total_strips, (b + 1) * strips_per_band);
				if (s1 <= s0) break;
				// this device's strips of the band: those congruent to strip_first modulo strip_step (as signal_band_done counts them)
				const int first_mine = s0 + ((strip_first - s0) % strip_step + strip_step) % strip_step;
				const int mine = first_mine < s1 ? (s1 - 1 - first_mine) / strip_step + 1 : 0;
				if (mine == 0) continue;
				if (s1 == total_strips) RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_kernel, 0));      // the last band: see render_pipelined
				else
				{
					const CUresult cr = wait((CUstream)d.copy_stream, (CUdeviceptr)(uintptr_t)(d.d_band_done + b), (cuuint32_t)(mine * grid_x * rt::kSignalsPerTile), CU_STREAM_WAIT_VALUE_GEQ);
					if (cr != CUDA_SUCCESS) return fail(ctx, RT_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)cr);
				}
				if ((rc = copy_strips_to_host(ctx, d, W, H, first_mine, strip_step, mine, target, pitch_bytes)) != RT_OK) return rc;
			}
		}
		else
		{
			RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
			if ((rc = launch(ctx, d, p, d.stream, my_strips)) != RT_OK) return rc;
			RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
			RT_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_kernel, 0));
			if ((rc = copy_strips_to_host(ctx, d, W, H, strip_first, strip_step, my_strips, target, pitch_bytes)) != RT_OK) return rc;
		}
		RT_CUDA(ctx, cudaEventRecord(d.ev_done, d.copy_stream));
		return RT_OK;
	}

	// Bounce-buffer case of prepare_host: only the strips this call produced may be copied back.
	void unbounce_strips(void* host_dst, const void* target, int W, int H, int32_t pitch_bytes, int strip_first, int strip_step)
	{
		const int total_strips = (H + rt::kBlockH - 1) / rt::kBlockH;
		for (int s = strip_first; s < total_strips; s += strip_step)
			for (int y = s * rt::kBlockH; y < std::min(H, (s + 1) * rt::kBlockH); ++y)
				memcpy((char*)host_dst + (size_t)y * pitch_bytes, (const char*)target + (size_t)y * pitch_bytes, (size_t)W * 4u);
	}

	// Devices [0, n_present) of the context present strips strip_first + k * strip_step_per_device ... : in-process
	// (all devices, device k takes strips k, k + n, ...) or one process per GPU (one device, its rank / world pair).
	int render_direct(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame, uint32_t* host_dst, int32_t pitch_bytes,
	                  int n_present, int strip_first, int strip_step)
	{
		// RT_B200_HOST_TIMING=1 (measurement only): where this thread's time goes, per frame, on stderr
		static const bool host_timing = getenv("RT_B200_HOST_TIMING") != nullptr;
		using Clock = std::chrono::steady_clock;
		Clock::time_point mark[6];
		auto stamp = [&](int i) { if (host_timing) mark[i] = Clock::now(); };
		stamp(0);
		int rc = validate_frame(ctx, camera, frame);
		if (rc != RT_OK) return rc;
		const int W = frame->width, H = frame->height;
		if (pitch_bytes < 4 * W || (pitch_bytes & 3)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "pitch_bytes %d too small or unaligned for width %d", pitch_bytes, W);
		const size_t span = (size_t)pitch_bytes * (size_t)(H - 1) + (size_t)W * 4u;
		void* target = nullptr;
		if ((rc = prepare_host(ctx, host_dst, span, &target)) != RT_OK) return rc;
		stamp(1);
		ctx->timing = rt_timing{};
		ctx->last_width = W; ctx->last_height = H;
		const rt::FrameParams base = make_params(camera, frame);
		for (int k = 0; k < n_present; ++k)
			if ((rc = enqueue_direct_present(ctx, ctx->devs[k], base, strip_first + k, strip_step, target, pitch_bytes)) != RT_OK) return rc;
		stamp(2);
		// watched launches: this thread issues every device's band copies as the bands arrive
		for (bool all_done = false; !all_done;)
		{
			all_done = true;
			for (int k = 0; k < n_present; ++k)
			{
				bool done = true;
				if ((rc = service_present(ctx, ctx->devs[k], n_present == 1, &done)) != RT_OK) return rc;
				all_done = all_done && done;
			}
		}
		stamp(3);
		float kernel_ms = 0.f, total_ms = 0.f;
		for (int k = 0; k < n_present; ++k)
		{
			DeviceState& d = ctx->devs[k];
			RT_CUDA(ctx, cudaSetDevice(d.device));
			RT_CUDA(ctx, cudaStreamSynchronize(d.copy_stream));
			RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
			stamp(4);
			float k_ms = 0.f, t_ms = 0.f;
			RT_CUDA(ctx, cudaEventElapsedTime(&k_ms, d.ev_begin, d.ev_kernel));
			RT_CUDA(ctx, cudaEventElapsedTime(&t_ms, d.ev_begin, d.ev_done));
			kernel_ms = std::max(kernel_ms, k_ms); total_ms = std::max(total_ms, t_ms);
		}
		RT_CUDA(ctx, cudaSetDevice(ctx->devs[0].device));
		if (target != host_dst)
			for (int k = 0; k < n_present; ++k) unbounce_strips(host_dst, target, W, H, pitch_bytes, strip_first + k, strip_step);
		ctx->timing.kernel_ms = kernel_ms; ctx->timing.gather_ms = 0.f; ctx->timing.d2h_ms = std::max(0.f, total_ms - kernel_ms); ctx->timing.total_ms = total_ms;
		stamp(5);
		if (host_timing)
		{
			auto us = [&](int a, int b) { return std::chrono::duration<double, std::micro>(mark[b] - mark[a]).count(); };
			fprintf(stderr, "rt_render host: prepare %.1f us, enqueue %.1f us, issue copies as bands arrive %.1f us, wait for the streams %.1f us, timing + return %.1f us; device: kernel %.1f us, kernel + copies %.1f us\n",
			        us(0, 1), us(1, 2), us(2, 3), us(3, 4), us(4, 5), kernel_ms * 1e3, total_ms * 1e3);
		}
		return RT_OK;
	}
}

extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

const char* rt_last_error(const rt_context* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int rt_device_count(const rt_context* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int rt_create(const int32_t* device_ids, int32_t n_devices, rt_context** out_ctx)
{
	if (!out_ctx) return fail(nullptr, RT_ERR_INVALID_ARGUMENT, "out_ctx must not be NULL");
	*out_ctx = nullptr;
	int available = 0;
	if (cudaGetDeviceCount(&available) != cudaSuccess || available <= 0)
	{
		cudaGetLastError();
		return fail(nullptr, RT_ERR_NO_DEVICE, "no CUDA device is visible: this library has no CPU path");
	}
	std::vector<int> ids;
	if (!device_ids || n_devices <= 0)
	{
		int cur = 0;
		if (cudaGetDevice(&cur) != cudaSuccess) return fail(nullptr, RT_ERR_CUDA, "cudaGetDevice failed");
		ids.push_back(cur);
	}
	else
	{
		for (int i = 0; i < n_devices; ++i)
		{
			if (device_ids[i] < 0 || device_ids[i] >= available) return fail(nullptr, RT_ERR_INVALID_ARGUMENT, "device id %d not in [0, %d)", device_ids[i], available);
			if (std::find(ids.begin(), ids.end(), device_ids[i]) != ids.end()) return fail(nullptr, RT_ERR_INVALID_ARGUMENT, "device id %d listed twice", device_ids[i]);
			ids.push_back(device_ids[i]);
		}
	}

	rt_context* ctx = new rt_context();
	int previous = 0;
	cudaGetDevice(&previous);
	auto bail = [&](int code) { g_create_error = ctx->error; rt_destroy(ctx); cudaSetDevice(previous); return code; };
#define RT_CREATE(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(ctx, RT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(RT_ERR_CUDA); } } while (0)

	for (int id : ids)
	{
		cudaDeviceProp prop{};
		RT_CREATE(cudaGetDeviceProperties(&prop, id));
		if (prop.major < 10)
		{
			fail(ctx, RT_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library carries sm_100a code only", id, prop.name, prop.major, prop.minor);
			return bail(RT_ERR_NO_DEVICE);
		}
		DeviceState d;
		d.device = id;
		RT_CREATE(cudaSetDevice(id));
		RT_CREATE(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
		RT_CREATE(cudaMalloc(&d.d_static, StaticBlock::total));
		RT_CREATE(cudaEventCreateWithFlags(&d.ev_upload, cudaEventDisableTiming));
		RT_CREATE(cudaEventRecord(d.ev_upload, d.stream));
		RT_CREATE(cudaEventCreateWithFlags(&d.ev_foreign, cudaEventDisableTiming));
		RT_CREATE(cudaEventCreateWithFlags(&d.ev_split, cudaEventDisableTiming));
		RT_CREATE(cudaMalloc(&d.d_counters, sizeof(unsigned long long) * RT_COUNTER_SLOTS));
		RT_CREATE(cudaEventCreate(&d.ev_begin));
		RT_CREATE(cudaEventCreate(&d.ev_kernel));
		RT_CREATE(cudaEventCreate(&d.ev_done));
		RT_CREATE(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
		for (cudaEvent_t& e : d.ev_band) RT_CREATE(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		RT_CREATE(cudaMalloc(&d.d_band_done, sizeof(unsigned int) * 128));
		RT_CREATE(cudaHostAlloc(&d.h_flags, sizeof(unsigned int) * 64, cudaHostAllocMapped | cudaHostAllocPortable));
		RT_CREATE(cudaHostAlloc(&d.h_wave_jobs, sizeof(unsigned int) * 4, cudaHostAllocMapped | cudaHostAllocPortable));
		d.h_wave_jobs[0] = d.h_wave_jobs[1] = d.h_wave_jobs[2] = d.h_wave_jobs[3] = 0u;
		RT_CREATE(cudaHostGetDevicePointer((void**)&d.d_wave_jobs, d.h_wave_jobs, 0));
		memset(d.h_flags, 0, sizeof(unsigned int) * 64);
		RT_CREATE(cudaHostGetDevicePointer((void**)&d.d_flags, d.h_flags, 0));
		RT_CREATE(cudaMalloc(&d.d_band_table, 8192));
		RT_CREATE(cudaHostAlloc(&d.h_band_table, 8192, cudaHostAllocPortable));
		RT_CREATE(cudaMemset(d.d_band_done, 0, sizeof(unsigned int) * 128));
		RT_CREATE(cudaMalloc(&d.d_split, sizeof(int32_t) * rt::kMaxMeshes * rt::wave::kSplitStride));
		RT_CREATE(cudaMemset(d.d_split, 0, sizeof(int32_t) * rt::kMaxMeshes * rt::wave::kSplitStride));
		RT_CREATE(cudaMalloc(&d.d_queues, sizeof(unsigned int) * 2 * kQueueRing));
		RT_CREATE(cudaMemset(d.d_queues, 0, sizeof(unsigned int) * 2 * kQueueRing));
		d.sm_count = prop.multiProcessorCount;
		ctx->devs.push_back(d);
	}
	RT_CREATE(cudaSetDevice(ids[0]));
	RT_CREATE(cudaHostAlloc(&ctx->h_build_status, sizeof(int32_t) * rt::kMaxMeshes, cudaHostAllocPortable));
	memset(ctx->h_build_status, 0, sizeof(int32_t) * rt::kMaxMeshes);
	RT_CREATE(cudaHostAlloc(&ctx->h_split, sizeof(int32_t) * rt::kMaxMeshes * rt::wave::kSplitStride, cudaHostAllocPortable));
	RT_CREATE(cudaHostAlloc(&ctx->h_static, StaticBlock::total, cudaHostAllocPortable));
	memset(ctx->h_static, 0, StaticBlock::total);
	ctx->arena = reinterpret_cast<float*>(ctx->h_static + StaticBlock::arena);
	ctx->materials = reinterpret_cast<float4*>(ctx->h_static + StaticBlock::materials);
	ctx->light_type = reinterpret_cast<int32_t*>(ctx->h_static + StaticBlock::light_type);
	ctx->bytes = ctx->h_static + StaticBlock::bytes;
	RT_CREATE(cudaEventCreate(&ctx->ev_gather));
	RT_CREATE(cudaEventCreate(&ctx->ev_d2h));

	// Peer-map device 0's memory into every other device so their kernels can store the
	// finished pixels straight into the gathered frame.
	ctx->peer_stores = ids.size() > 1;
	for (size_t k = 1; k < ids.size(); ++k)
	{
		int can = 0;
		RT_CREATE(cudaDeviceCanAccessPeer(&can, ids[k], ids[0]));
		if (!can) { ctx->peer_stores = false; continue; }
		RT_CREATE(cudaSetDevice(ids[k]));
		const cudaError_t e = cudaDeviceEnablePeerAccess(ids[0], 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ctx->peer_stores = false;
		cudaGetLastError();
	}
	const int rc = flush_uploads(ctx);
	if (rc != RT_OK) return bail(rc);
	cudaSetDevice(previous);
	*out_ctx = ctx;
	return RT_OK;
#undef RT_CREATE
}

int rt_destroy(rt_context* ctx)
{
	if (!ctx) return RT_OK;
	for (DeviceState& d : ctx->devs)
	{
		cudaSetDevice(d.device);
		if (d.stream) cudaStreamSynchronize(d.stream);
		cudaFree(d.d_static); cudaFree(d.d_mesh); cudaFree(d.d_frame); cudaFree(d.d_counters); cudaFree(d.d_band_done); cudaFree(d.d_queues);
		cudaFree(d.d_split); cudaFree(d.d_wave); cudaFree(d.d_root_map); cudaFree(d.d_band_table); cudaFreeHost(d.h_band_table); cudaFreeHost(d.h_flags); cudaFreeHost(d.h_wave_jobs);
		if (d.ev_upload) cudaEventDestroy(d.ev_upload);
		if (d.ev_foreign) cudaEventDestroy(d.ev_foreign);
		if (d.ev_split) cudaEventDestroy(d.ev_split);
		cudaFree(d.cells.d_cost); cudaFree(d.cells.d_order); cudaFreeHost(d.cells.h_cost); cudaFreeHost(d.cells.h_order);
		for (cudaEvent_t e : d.cells.ev_cost) if (e) cudaEventDestroy(e);
		for (cudaEvent_t e : d.cells.ev_order) if (e) cudaEventDestroy(e);
		if (d.cells.ev_last) cudaEventDestroy(d.cells.ev_last);
		for (auto& sd : d.sources) { cudaFree(sd.positions); cudaFree(sd.normals); cudaFree(sd.indices); cudaFree(sd.normals_alt); cudaFree(sd.indices_alt); cudaFree(sd.build_block); }
		if (d.ev_begin) cudaEventDestroy(d.ev_begin);
		if (d.ev_kernel) cudaEventDestroy(d.ev_kernel);
		if (d.ev_done) cudaEventDestroy(d.ev_done);
		for (cudaEvent_t e : d.ev_band) if (e) cudaEventDestroy(e);
		if (d.copy_stream) { cudaStreamSynchronize(d.copy_stream); cudaStreamDestroy(d.copy_stream); }
		if (d.stream) cudaStreamDestroy(d.stream);
	}
	if (ctx->ev_gather) cudaEventDestroy(ctx->ev_gather);
	if (ctx->ev_d2h) cudaEventDestroy(ctx->ev_d2h);
	for (auto& r : ctx->registered) cudaHostUnregister(r.first);
	if (ctx->staging) cudaFreeHost(ctx->staging);
	if (ctx->h_split) cudaFreeHost(ctx->h_split);
	if (ctx->h_root_map) cudaFreeHost(ctx->h_root_map);
	if (ctx->h_static) cudaFreeHost(ctx->h_static);
	if (ctx->h_mesh) cudaFreeHost(ctx->h_mesh);
	if (ctx->h_build_status) cudaFreeHost(ctx->h_build_status);
	cudaGetLastError();
	delete ctx;
	return RT_OK;
}

int rt_upload_spheres(rt_context* ctx, const rt_spheres_soa* s)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!s || s->count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad sphere array");
	if (s->count > rt::kMaxSpheres) return fail(ctx, RT_ERR_CAPACITY, "%d spheres exceed the capacity of %d", s->count, rt::kMaxSpheres);
	if (s->count > 0 && (!s->origin_x || !s->origin_y || !s->origin_z || !s->radius || !s->material_index))
		return fail(ctx, RT_ERR_INVALID_ARGUMENT, "sphere SoA has a NULL array");
	float* a = ctx->arena + ArenaLayout::sphere;
	if (s->count == ctx->n_spheres && same_bits(a, s->origin_x, s->count) && same_bits(a + rt::kMaxSpheres, s->origin_y, s->count) &&
	    same_bits(a + 2 * rt::kMaxSpheres, s->origin_z, s->count) && same_bits(a + 3 * rt::kMaxSpheres, s->radius, s->count) &&
	    same_bits(ctx->bytes, s->material_index, s->count)) return RT_OK;
	if (int w = wait_uploads(ctx)) return w;
	for (int i = 0; i < s->count; ++i)
	{
		a[i] = s->origin_x[i]; a[rt::kMaxSpheres + i] = s->origin_y[i]; a[2 * rt::kMaxSpheres + i] = s->origin_z[i];
		a[3 * rt::kMaxSpheres + i] = s->radius[i];
		ctx->bytes[i] = s->material_index[i];
	}
	ctx->n_spheres = s->count;
	ctx->static_dirty = true;
	return RT_OK;
}

int rt_upload_planes(rt_context* ctx, const rt_planes_soa* p)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!p || p->count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad plane array");
	if (p->count > rt::kMaxPlanes) return fail(ctx, RT_ERR_CAPACITY, "%d planes exceed the capacity of %d", p->count, rt::kMaxPlanes);
	if (p->count > 0 && (!p->origin_x || !p->origin_y || !p->origin_z || !p->normal_x || !p->normal_y || !p->normal_z || !p->material_index))
		return fail(ctx, RT_ERR_INVALID_ARGUMENT, "plane SoA has a NULL array");
	float* a = ctx->arena + ArenaLayout::plane;
	const int M = rt::kMaxPlanes;
	if (p->count == ctx->n_planes && same_bits(a, p->origin_x, p->count) && same_bits(a + M, p->origin_y, p->count) && same_bits(a + 2 * M, p->origin_z, p->count) &&
	    same_bits(a + 3 * M, p->normal_x, p->count) && same_bits(a + 4 * M, p->normal_y, p->count) && same_bits(a + 5 * M, p->normal_z, p->count) &&
	    same_bits(ctx->bytes + rt::kMaxSpheres, p->material_index, p->count)) return RT_OK;
	if (int w = wait_uploads(ctx)) return w;
	for (int i = 0; i < p->count; ++i)
	{
		a[i] = p->origin_x[i]; a[M + i] = p->origin_y[i]; a[2 * M + i] = p->origin_z[i];
		a[3 * M + i] = p->normal_x[i]; a[4 * M + i] = p->normal_y[i]; a[5 * M + i] = p->normal_z[i];
		ctx->bytes[rt::kMaxSpheres + i] = p->material_index[i];
	}
	ctx->n_planes = p->count;
	ctx->static_dirty = true;
	return RT_OK;
}

int rt_upload_lights(rt_context* ctx, const rt_lights_soa* l)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!l || l->count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad light array");
	if (l->count > rt::kMaxLights) return fail(ctx, RT_ERR_CAPACITY, "%d lights exceed the capacity of %d", l->count, rt::kMaxLights);
	if (l->count > 0 && (!l->origin_x || !l->origin_y || !l->origin_z || !l->color_r || !l->color_g || !l->color_b || !l->intensity || !l->type))
		return fail(ctx, RT_ERR_INVALID_ARGUMENT, "light SoA has a NULL array");
	float* a = ctx->arena + ArenaLayout::light;
	const int M = rt::kMaxLights;
	if (l->count == ctx->n_lights && same_bits(a, l->origin_x, l->count) && same_bits(a + M, l->origin_y, l->count) && same_bits(a + 2 * M, l->origin_z, l->count) &&
	    same_bits(a + 3 * M, l->color_r, l->count) && same_bits(a + 4 * M, l->color_g, l->count) && same_bits(a + 5 * M, l->color_b, l->count) &&
	    same_bits(a + 6 * M, l->intensity, l->count) && same_bits(ctx->light_type, l->type, l->count)) return RT_OK;
	if (int w = wait_uploads(ctx)) return w;
	for (int i = 0; i < l->count; ++i)
	{
		a[i] = l->origin_x[i]; a[M + i] = l->origin_y[i]; a[2 * M + i] = l->origin_z[i];
		a[3 * M + i] = l->color_r[i]; a[4 * M + i] = l->color_g[i]; a[5 * M + i] = l->color_b[i];
		a[6 * M + i] = l->intensity[i];
		ctx->light_type[i] = l->type[i];
	}
	ctx->n_lights = l->count;
	ctx->static_dirty = true;
	return RT_OK;
}

int rt_upload_materials(rt_context* ctx, const rt_material_desc* materials, int32_t count)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (count < 0 || (count > 0 && !materials)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad material array");
	if (count > rt::kMaxMaterials) return fail(ctx, RT_ERR_CAPACITY, "%d materials exceed the capacity of %d", count, rt::kMaxMaterials);
	for (int i = 0; i < count; ++i)
		if (materials[i].tag < RT_MATERIAL_SOLID_COLOR || materials[i].tag > RT_MATERIAL_COOK_TORRENCE) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "material %d has unknown tag %d", i, materials[i].tag);
	bool unchanged = count == ctx->n_materials;
	for (int i = 0; unchanged && i < count; ++i)
	{
		const rt_material_desc& m = materials[i];
		const float4 m0 = make_float4(bits_as_float(m.tag), m.color[0], m.color[1], m.color[2]), m1 = make_float4(m.p0, m.p1, m.p2, 0.f);
		unchanged = memcmp(&ctx->materials[2 * i], &m0, sizeof m0) == 0 && memcmp(&ctx->materials[2 * i + 1], &m1, sizeof m1) == 0;
	}
	if (unchanged) return RT_OK;
	if (int w = wait_uploads(ctx)) return w;
	for (int i = 0; i < count; ++i)
	{
		const rt_material_desc& m = materials[i];
		ctx->materials[2 * i] = make_float4(bits_as_float(m.tag), m.color[0], m.color[1], m.color[2]);
		ctx->materials[2 * i + 1] = make_float4(m.p0, m.p1, m.p2, 0.f);
	}
	ctx->n_materials = count;
	ctx->static_dirty = true;
	return RT_OK;
}

int rt_set_mesh_count(rt_context* ctx, int32_t mesh_count)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "negative mesh count");
	if (mesh_count > rt::kMaxMeshes) return fail(ctx, RT_ERR_CAPACITY, "%d meshes exceed the capacity of %d", mesh_count, rt::kMaxMeshes);
	if ((size_t)mesh_count == ctx->meshes.size()) return RT_OK;      // called every frame by the drop-in: nothing changed
	ctx->meshes.resize((size_t)mesh_count);
	ctx->mesh_dirty = true;
	return RT_OK;
}

int rt_set_mesh_path(rt_context* ctx, int32_t mesh_path)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_path < RT_MESH_PATH_AUTO || mesh_path > RT_MESH_PATH_BVH) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "unknown mesh path %d", mesh_path);
	ctx->mesh_path = mesh_path;
	return RT_OK;
}

int rt_upload_mesh_source(rt_context* ctx, int32_t mesh_id, const rt_mesh_source* src)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_id < 0 || mesh_id >= (int32_t)ctx->meshes.size()) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh id %d outside the announced count %d", mesh_id, (int)ctx->meshes.size());
	if (!src || src->triangle_count < 0 || src->vertex_count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad mesh source descriptor");
	if (src->triangle_count > 0 && (!src->positions || !src->indices || !src->normals)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh source has a NULL array");
	if (src->cull_mode < RT_CULL_FRONT_FACE || src->cull_mode > RT_CULL_NONE) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "unknown cull mode %d", src->cull_mode);
	for (int i = 0; i < 3 * src->triangle_count; ++i)
		if (src->indices[i] < 0 || src->indices[i] >= src->vertex_count) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "index %d of mesh %d is out of range", i, mesh_id);
	HostMesh& hm = ctx->meshes[(size_t)mesh_id];
	hm.src_positions.assign(src->positions, src->positions + 3 * (size_t)src->vertex_count);
	hm.src_indices.assign(src->indices, src->indices + 3 * (size_t)src->triangle_count);
	hm.src_normals.assign(src->normals, src->normals + 3 * (size_t)src->triangle_count);
	// the stream slice is sized now and filled by transform_mesh_kernel; no BVH -> slab + linear body
	// (rt_set_mesh_device_bvh reserves the node slice and switches to update_transforms_bvh_kernel)
	hm.triangles.assign(3 * (size_t)src->triangle_count, make_float4(0.f, 0.f, 0.f, 0.f));
	hm.nodes.clear();
	hm.split.clear(); hm.root_map.clear();
	hm.device_bvh = false; hm.pending_builds.clear();
	for (int k = 0; k < 3; ++k) { hm.aabb_min[k] = 0.f; hm.aabb_max[k] = 0.f; }
	hm.cull_mode = src->cull_mode;
	hm.material = src->material_index;
	hm.uploaded = true;
	hm.has_source = true; hm.source_dirty = true; hm.has_transform = false;
	ctx->mesh_dirty = true;
	return RT_OK;
}

int rt_transform_mesh(rt_context* ctx, int32_t mesh_id, const float* transform)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_id < 0 || mesh_id >= (int32_t)ctx->meshes.size()) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh id %d outside the announced count %d", mesh_id, (int)ctx->meshes.size());
	if (!transform) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "transform must not be NULL");
	HostMesh& hm = ctx->meshes[(size_t)mesh_id];
	if (!hm.has_source) return fail(ctx, RT_ERR_BAD_STATE, "mesh %d was not uploaded with rt_upload_mesh_source", mesh_id);
	memcpy(hm.transform, transform, sizeof hm.transform);
	hm.has_transform = true; hm.transform_dirty = true;
	if (hm.device_bvh)
	{
		std::array<float, 16> t;
		memcpy(t.data(), transform, sizeof hm.transform);
		hm.pending_builds.push_back(t);
	}
	return RT_OK;
}

int rt_set_mesh_device_bvh(rt_context* ctx, int32_t mesh_id, int32_t enable)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_id < 0 || mesh_id >= (int32_t)ctx->meshes.size()) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh id %d outside the announced count %d", mesh_id, (int)ctx->meshes.size());
	HostMesh& hm = ctx->meshes[(size_t)mesh_id];
	if (!hm.has_source) return fail(ctx, RT_ERR_BAD_STATE, "mesh %d was not uploaded with rt_upload_mesh_source", mesh_id);
	if (hm.has_transform) return fail(ctx, RT_ERR_BAD_STATE, "rt_set_mesh_device_bvh must be called before the first rt_transform_mesh of mesh %d", mesh_id);
	const size_t T = hm.src_indices.size() / 3;
	if (enable && T > (size_t)rt::BvhLink::kFirstMask) return fail(ctx, RT_ERR_CAPACITY, "%zu triangles exceed what the BVH node links can address", T);
	hm.device_bvh = enable != 0;
	// node slice: 2T - 1 nodes at most, written by emit_mesh_kernel
	if (hm.device_bvh && T > 0) hm.nodes.assign(2 * (2 * T - 1), make_float4(0.f, 0.f, 0.f, 0.f)); else hm.nodes.clear();
	hm.source_dirty = true;      // (re)allocate the device-side build buffers
	ctx->mesh_dirty = true;
	return RT_OK;
}

int rt_set_kernel_variant(rt_context* ctx, int32_t variant)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (variant < RT_KERNEL_AUTO || variant > RT_KERNEL_WAVEFRONT) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "unknown kernel variant %d", variant);
	ctx->kernel_variant = variant;
	return RT_OK;
}

int rt_upload_mesh(rt_context* ctx, int32_t mesh_id, const rt_mesh_desc* mesh)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_id < 0 || mesh_id >= (int32_t)ctx->meshes.size()) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh id %d outside the announced count %d", mesh_id, (int)ctx->meshes.size());
	if (!mesh || mesh->triangle_count < 0 || mesh->vertex_count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad mesh descriptor");
	if (mesh->triangle_count > 0 && (!mesh->positions || !mesh->indices || !mesh->normals)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh has a NULL array");
	if (mesh->cull_mode < RT_CULL_FRONT_FACE || mesh->cull_mode > RT_CULL_NONE) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "unknown cull mode %d", mesh->cull_mode);
	for (int i = 0; i < 3 * mesh->triangle_count; ++i)
		if (mesh->indices[i] < 0 || mesh->indices[i] >= mesh->vertex_count) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "index %d of mesh %d is out of range", i, mesh_id);

	HostMesh& hm = ctx->meshes[(size_t)mesh_id];
	if (hm.has_source)
	{
		// the id held a device-transformed mesh: forget its build state and give its device buffers back
		hm.device_bvh = false; hm.pending_builds.clear(); hm.transform_dirty = false; hm.source_dirty = false;
		hm.src_positions.clear(); hm.src_normals.clear(); hm.src_indices.clear();
		for (DeviceState& d : ctx->devs)
		{
			if ((size_t)mesh_id >= d.sources.size()) continue;
			RT_CUDA(ctx, cudaSetDevice(d.device));
			RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
			DeviceState::MeshSourceDevice& sd = d.sources[(size_t)mesh_id];
			cudaFree(sd.positions); cudaFree(sd.normals); cudaFree(sd.indices); cudaFree(sd.normals_alt); cudaFree(sd.indices_alt); cudaFree(sd.build_block);
			sd = DeviceState::MeshSourceDevice{};
		}
	}
	hm.has_source = false; hm.has_transform = false;
	hm.triangles.resize(3 * (size_t)mesh->triangle_count);
	// Bounds of the indexed vertices, started like a BVH root box (reference
	// source/DataTypes.h:310-321 with MaxVector / MinVector of source/Vector3.cpp:13-14).
	float bmin[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, bmax[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
	for (int t = 0; t < mesh->triangle_count; ++t)
	{
		const float* v0 = mesh->positions + 3 * (size_t)mesh->indices[3 * t];
		const float* v1 = mesh->positions + 3 * (size_t)mesh->indices[3 * t + 1];
		const float* v2 = mesh->positions + 3 * (size_t)mesh->indices[3 * t + 2];
		const float* n = mesh->normals + 3 * (size_t)t;
		// e1 = v1 - v0, e2 = v2 - v0 (reference source/Utils.h:143-144): single IEEE subtractions
		hm.triangles[3 * t + 0] = make_float4(v0[0], v0[1], v0[2], n[0]);
		hm.triangles[3 * t + 1] = make_float4(v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2], n[1]);
		hm.triangles[3 * t + 2] = make_float4(v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2], n[2]);
		const float* vs[3] = { v0, v1, v2 };
		for (const float* v : vs)
			for (int k = 0; k < 3; ++k) { bmin[k] = std_min_f(bmin[k], v[k]); bmax[k] = std_max_f(bmax[k], v[k]); }
	}
	for (int k = 0; k < 3; ++k)
	{
		hm.aabb_min[k] = mesh->aabb_min ? mesh->aabb_min[k] : bmin[k];
		hm.aabb_max[k] = mesh->aabb_max ? mesh->aabb_max[k] : bmax[k];
	}
	hm.cull_mode = mesh->cull_mode;
	hm.material = mesh->material_index;
	if (mesh->bvh_node_count < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "negative BVH node count");
	hm.split.clear(); hm.root_map.clear(); hm.split_built = false;
	const int brc = thread_bvh(ctx, mesh, hm.nodes);
	if (brc != RT_OK) { hm.uploaded = false; return brc; }
	hm.uploaded = true;
	ctx->mesh_dirty = true;
	return RT_OK;
}

int rt_render_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	int rc = render_to_device0(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	RT_CUDA(ctx, cudaSetDevice(ctx->devs[0].device));
	RT_CUDA(ctx, cudaStreamSynchronize(ctx->devs[0].stream));
	return collect_timing(ctx, false);
}

int rt_render(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame, uint32_t* host_dst, int32_t pitch_bytes)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!host_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "host_dst must not be NULL");
	// several devices and a host destination: every device presents its own strips over its own PCIe link
	// (RT_B200_PRESENT=gather keeps the flow through device 0's frame buffer, e.g. to compare)
	static const bool gather_first = [] { const char* e = getenv("RT_B200_PRESENT"); return e && strcmp(e, "gather") == 0; }();
	if (ctx->devs.size() > 1 && !gather_first) return render_direct(ctx, camera, frame, host_dst, pitch_bytes, (int)ctx->devs.size(), 0, (int)ctx->devs.size());
	// one device: the same flow with a single share - kernel with band counters, band copies issued by this thread as the
	// kernel's band watcher reports them (RT_B200_SINGLE_PRESENT=stream keeps round 1's copy stream with polled waits)
	static const bool stream_present = [] { const char* e = getenv("RT_B200_SINGLE_PRESENT"); return e && strcmp(e, "stream") == 0; }();
	if (ctx->devs.size() == 1 && !stream_present) return render_direct(ctx, camera, frame, host_dst, pitch_bytes, 1, 0, 1);
	if (ctx->devs.size() == 1 || (ctx->peer_stores && wait_value32())) return render_pipelined(ctx, camera, frame, host_dst, pitch_bytes);
	int rc = render_to_device0(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if ((rc = download(ctx, host_dst, pitch_bytes, true)) != RT_OK) return rc;
	return collect_timing(ctx, true);
}

int rt_download_frame(rt_context* ctx, uint32_t* host_dst, int32_t pitch_bytes)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	return download(ctx, host_dst, pitch_bytes, false);
}

int rt_render_rows_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                          int32_t row_begin, int32_t row_count, void* device_dst, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	int rc = validate_frame(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if (row_begin < 0 || row_count < 0 || row_begin + row_count > frame->height) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "rows [%d, %d) outside the frame", row_begin, row_begin + row_count);
	if (!device_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "device_dst must not be NULL");
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	rt::FrameParams p = make_params(camera, frame);
	p.row_begin = row_begin; p.row_end = row_begin + row_count;
	p.strip_first = 0; p.strip_step = 1; p.dst_full_frame = 0; p.dst = (uint32_t*)device_dst;
	ctx->timing = rt_timing{};
	rc = launch(ctx, d, p, stream, (row_count + rt::kBlockH - 1) / rt::kBlockH);
	if (rc != RT_OK) return rc;
	if (!cuda_stream) RT_CUDA(ctx, cudaStreamSynchronize(stream));
	return RT_OK;
}

int rt_render_strips_device(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                            int32_t strip_first, int32_t strip_step, void* device_dst, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	int rc = validate_frame(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if (strip_step <= 0 || strip_first < 0 || strip_first >= strip_step) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "strip_first %d / strip_step %d is not a rank / world pair", strip_first, strip_step);
	if (!device_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "device_dst must not be NULL");
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	const int total_strips = (frame->height + rt::kBlockH - 1) / rt::kBlockH;
	rt::FrameParams p = make_params(camera, frame);
	p.row_begin = 0; p.row_end = frame->height;
	p.strip_first = strip_first; p.strip_step = strip_step; p.dst_full_frame = 0; p.dst = (uint32_t*)device_dst;
	ctx->timing = rt_timing{};
#ifdef RT_DRAIN_PROBE
	// experiment build: timestamps of the launch (see render_kernel_persistent), printed after a synchronisation
	{
		const unsigned long long init[3] = { ~0ull, ~0ull, 0ull };
		RT_CUDA(ctx, cudaMemcpyAsync(d.d_counters, init, sizeof init, cudaMemcpyHostToDevice, stream));
		RT_CUDA(ctx, cudaStreamSynchronize(stream));
		p.counters = d.d_counters;
	}
#endif
	rc = launch(ctx, d, p, stream, (total_strips - strip_first + strip_step - 1) / strip_step);
	if (rc != RT_OK) return rc;
#ifdef RT_DRAIN_PROBE
	{
		unsigned long long t[3] = {};
		RT_CUDA(ctx, cudaStreamSynchronize(stream));
		RT_CUDA(ctx, cudaMemcpy(t, d.d_counters, sizeof t, cudaMemcpyDeviceToHost));
		fprintf(stderr, "drain probe: strips %d/%d  queue empty after %.1f us, last warp out after %.1f us (drain %.1f us)\n", strip_first, strip_step,
		        (double)(t[1] - t[0]) * 1e-3, (double)(t[2] - t[0]) * 1e-3, (double)(t[2] - t[1]) * 1e-3);
	}
#endif
	if (!cuda_stream) RT_CUDA(ctx, cudaStreamSynchronize(stream));
	return RT_OK;
}

int rt_unstripe_device(rt_context* ctx, const void* device_src, void* device_dst, int32_t width, int32_t height,
                       int32_t world, int32_t strips_per_rank, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!device_src || !device_dst || width <= 0 || height <= 0 || world <= 0 || strips_per_rank <= 0)
		return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad unstripe arguments");
	const int total_strips = (height + rt::kBlockH - 1) / rt::kBlockH;
	if ((long long)world * strips_per_rank < total_strips) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "%d x %d strips do not cover %d rows", world, strips_per_rank, height);
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	const int vec = (width % 4 == 0) && ((reinterpret_cast<uintptr_t>(device_src) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(device_dst) & 15u) == 0);
	const long long units = vec ? (long long)width / 4 * height : (long long)width * height;
	const int threads = 256;
	const unsigned blocks = (unsigned)std::min<long long>((units + threads - 1) / threads, 148LL * 16);
	rt::unstripe_kernel<<<blocks, threads, 0, stream>>>((const uint32_t*)device_src, (uint32_t*)device_dst, width, height, world, strips_per_rank, vec);
	RT_CUDA(ctx, cudaGetLastError());
	ctx->timing.kernel_launches++;
	if (!cuda_stream) RT_CUDA(ctx, cudaStreamSynchronize(stream));
	return RT_OK;
}

int rt_frame_export(rt_context* ctx, int32_t width, int32_t height, void* out_handle)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!out_handle || width <= 0 || height <= 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad frame export arguments");
	static_assert(sizeof(cudaIpcMemHandle_t) <= RT_IPC_HANDLE_BYTES, "IPC handle does not fit");
	DeviceState& d = ctx->devs[0];
	int rc = ensure_frame(ctx, d, (size_t)width * (size_t)height);
	if (rc != RT_OK) return rc;
	RT_CUDA(ctx, cudaSetDevice(d.device));
	RT_CUDA(ctx, cudaMemset((char*)d.d_frame + signal_offset((size_t)width * (size_t)height), 0, 256));
	cudaIpcMemHandle_t h;
	RT_CUDA(ctx, cudaIpcGetMemHandle(&h, d.d_frame));
	memset(out_handle, 0, RT_IPC_HANDLE_BYTES);
	memcpy(out_handle, &h, sizeof h);
	ctx->last_width = width; ctx->last_height = height;
	return RT_OK;
}

int rt_frame_import(rt_context* ctx, const void* handle, void** out_device_ptr)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!handle || !out_device_ptr) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad frame import arguments");
	RT_CUDA(ctx, cudaSetDevice(ctx->devs[0].device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	RT_CUDA(ctx, cudaIpcOpenMemHandle(out_device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return RT_OK;
}

int rt_frame_release(rt_context* ctx, void* device_ptr)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!device_ptr) return RT_OK;
	RT_CUDA(ctx, cudaSetDevice(ctx->devs[0].device));
	RT_CUDA(ctx, cudaIpcCloseMemHandle(device_ptr));
	return RT_OK;
}

int rt_render_strips_to_frame(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                              int32_t strip_first, int32_t strip_step, void* frame_device_ptr, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	int rc = validate_frame(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if (strip_step <= 0 || strip_first < 0 || strip_first >= strip_step) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "strip_first %d / strip_step %d is not a rank / world pair", strip_first, strip_step);
	DeviceState& d = ctx->devs[0];
	if (!frame_device_ptr)
	{
		if ((rc = ensure_frame(ctx, d, (size_t)frame->width * (size_t)frame->height)) != RT_OK) return rc;
		frame_device_ptr = d.d_frame;
		ctx->last_width = frame->width; ctx->last_height = frame->height;
	}
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	const int total_strips = (frame->height + rt::kBlockH - 1) / rt::kBlockH;
	rt::FrameParams p = make_params(camera, frame);
	p.row_begin = 0; p.row_end = frame->height;
	p.strip_first = strip_first; p.strip_step = strip_step; p.dst_full_frame = 1; p.dst = (uint32_t*)frame_device_ptr;
	ctx->timing = rt_timing{};
	rc = launch(ctx, d, p, stream, (total_strips - strip_first + strip_step - 1) / strip_step);
	if (rc != RT_OK) return rc;
	if (!cuda_stream) RT_CUDA(ctx, cudaStreamSynchronize(stream));
	return RT_OK;
}

int rt_render_strips_to_host(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                             int32_t strip_first, int32_t strip_step, uint32_t* host_dst, int32_t pitch_bytes)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!host_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "host_dst must not be NULL");
	if (strip_step <= 0 || strip_first < 0 || strip_first >= strip_step) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "strip_first %d / strip_step %d is not a rank / world pair", strip_first, strip_step);
	return render_direct(ctx, camera, frame, host_dst, pitch_bytes, 1, strip_first, strip_step);
}

int rt_render_strips_to_frame_banded(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame,
                                     int32_t strip_first, int32_t strip_step, void* frame_device_ptr,
                                     int32_t bands, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	int rc = validate_frame(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if (strip_step <= 0 || strip_first < 0 || strip_first >= strip_step) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "strip_first %d / strip_step %d is not a rank / world pair", strip_first, strip_step);
	if (bands < 1 || bands > RT_MAX_PRESENT_BANDS) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bands %d outside [1, %d]", bands, RT_MAX_PRESENT_BANDS);
	DeviceState& d = ctx->devs[0];
	const size_t pixels = (size_t)frame->width * (size_t)frame->height;
	if (!frame_device_ptr)
	{
		if ((rc = ensure_frame(ctx, d, pixels)) != RT_OK) return rc;
		frame_device_ptr = d.d_frame;
		ctx->last_width = frame->width; ctx->last_height = frame->height;
	}
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	const int total_strips = (frame->height + rt::kBlockH - 1) / rt::kBlockH;
	rt::FrameParams p = make_params(camera, frame);
	p.row_begin = 0; p.row_end = frame->height;
	p.strip_first = strip_first; p.strip_step = strip_step; p.dst_full_frame = 1; p.dst = (uint32_t*)frame_device_ptr;
	// trailer word 0 is the whole-frame signal (rt_frame_signal), words 1.. are the band counters
	p.band_done = reinterpret_cast<unsigned int*>((char*)frame_device_ptr + signal_offset(pixels)) + 1;
	p.strips_per_band = (total_strips + bands - 1) / bands;
	p.band_local = d.d_band_done + 64;
	ctx->timing = rt_timing{};
	rc = launch(ctx, d, p, stream, (total_strips - strip_first + strip_step - 1) / strip_step);
	if (rc != RT_OK) return rc;
	if (!cuda_stream) RT_CUDA(ctx, cudaStreamSynchronize(stream));
	return RT_OK;
}

int rt_frame_present(rt_context* ctx, uint32_t* host_dst, int32_t pitch_bytes, int32_t bands, uint32_t frame_number)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	DeviceState& d = ctx->devs[0];
	const int W = ctx->last_width, H = ctx->last_height;
	if (!d.d_frame || W <= 0 || H <= 0) return fail(ctx, RT_ERR_BAD_STATE, "rt_frame_present needs an exported frame");
	if (!host_dst) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "host_dst must not be NULL");
	if (pitch_bytes < 4 * W || (pitch_bytes & 3)) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "pitch_bytes %d too small or unaligned for width %d", pitch_bytes, W);
	if (bands < 1 || bands > RT_MAX_PRESENT_BANDS) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bands %d outside [1, %d]", bands, RT_MAX_PRESENT_BANDS);
	const WaitValue32Fn wait = wait_value32();
	if (!wait) return fail(ctx, RT_ERR_BAD_STATE, "stream memory operations (cuStreamWaitValue32) are not available");
	const size_t span = (size_t)pitch_bytes * (size_t)(H - 1) + (size_t)W * 4u;
	void* target = nullptr;
	int rc = prepare_host(ctx, host_dst, span, &target);
	if (rc != RT_OK) return rc;
	RT_CUDA(ctx, cudaSetDevice(d.device));
	const int total_strips = (H + rt::kBlockH - 1) / rt::kBlockH, grid_x = (W + rt::kBlockW - 1) / rt::kBlockW;
	const int strips_per_band = (total_strips + bands - 1) / bands;
	unsigned int* counters = reinterpret_cast<unsigned int*>((char*)d.d_frame + signal_offset((size_t)W * (size_t)H)) + 1;
	for (int b = 0; b < bands; ++b)
	{
		const int s0 = b * strips_per_band, s1 = std::min(total_strips, (b + 1) * strips_per_band);
		if (s1 <= s0) break;
		const cuuint32_t expected = (cuuint32_t)((uint64_t)frame_number * (uint64_t)((s1 - s0) * grid_x * rt::kSignalsPerTile));
		const CUresult cr = wait((CUstream)d.copy_stream, (CUdeviceptr)(uintptr_t)(counters + b), expected, CU_STREAM_WAIT_VALUE_GEQ);
		if (cr != CUDA_SUCCESS) return fail(ctx, RT_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)cr);
		const int r0 = s0 * rt::kBlockH, r1 = std::min(H, s1 * rt::kBlockH);
		char* dst = (char*)target + (size_t)r0 * (size_t)pitch_bytes;
		const uint32_t* src = d.d_frame + (size_t)r0 * (size_t)W;
		if (pitch_bytes == 4 * W)
			RT_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)(r1 - r0) * W * 4u, cudaMemcpyDeviceToHost, d.copy_stream));
		else
			RT_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)pitch_bytes, src, (size_t)W * 4u, (size_t)W * 4u, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, d.copy_stream));
	}
	RT_CUDA(ctx, cudaStreamSynchronize(d.copy_stream));
	if (target != host_dst)
	{
		if (pitch_bytes == 4 * W) memcpy(host_dst, target, (size_t)W * H * 4u);
		else for (int y = 0; y < H; ++y) memcpy((char*)host_dst + (size_t)y * pitch_bytes, (char*)target + (size_t)y * pitch_bytes, (size_t)W * 4u);
	}
	return RT_OK;
}

int rt_frame_signal(rt_context* ctx, void* frame_device_ptr, int32_t width, int32_t height, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!frame_device_ptr || width <= 0 || height <= 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad frame signal arguments");
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	unsigned int* word = reinterpret_cast<unsigned int*>((char*)frame_device_ptr + signal_offset((size_t)width * (size_t)height));
	rt::frame_signal_kernel<<<1, 1, 0, stream>>>(word);
	RT_CUDA(ctx, cudaGetLastError());
	ctx->timing.kernel_launches++;
	return RT_OK;
}

int rt_frame_wait(rt_context* ctx, uint32_t expected, void* cuda_stream)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	DeviceState& d = ctx->devs[0];
	if (!d.d_frame || ctx->last_width <= 0) return fail(ctx, RT_ERR_BAD_STATE, "rt_frame_wait needs an exported frame");
	const WaitValue32Fn wait = wait_value32();
	if (!wait) return fail(ctx, RT_ERR_BAD_STATE, "stream memory operations (cuStreamWaitValue32) are not available");
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d.stream;
	char* word = (char*)d.d_frame + signal_offset((size_t)ctx->last_width * (size_t)ctx->last_height);
	const CUresult cr = wait((CUstream)stream, (CUdeviceptr)(uintptr_t)word, (cuuint32_t)expected, CU_STREAM_WAIT_VALUE_GEQ);
	if (cr != CUDA_SUCCESS) return fail(ctx, RT_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)cr);
	return RT_OK;
}

int rt_register_surface(rt_context* ctx, void* host_ptr, size_t bytes)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!host_ptr || bytes == 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "bad surface");
	for (auto& r : ctx->registered) if (r.first == host_ptr) return fail(ctx, RT_ERR_BAD_STATE, "surface %p is already registered", host_ptr);
	RT_CUDA(ctx, cudaSetDevice(ctx->devs[0].device));
	RT_CUDA(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
	ctx->registered.emplace_back(host_ptr, bytes);
	return RT_OK;
}

int rt_unregister_surface(rt_context* ctx, void* host_ptr)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	for (size_t i = 0; i < ctx->registered.size(); ++i)
	{
		if (ctx->registered[i].first != host_ptr) continue;
		// no copy of ours may still target it: every rt_render* that writes host memory is blocking, so only a sync for safety
		for (DeviceState& d : ctx->devs) { cudaSetDevice(d.device); cudaStreamSynchronize(d.copy_stream); cudaStreamSynchronize(d.stream); }
		cudaSetDevice(ctx->devs[0].device);
		RT_CUDA(ctx, cudaHostUnregister(host_ptr));
		ctx->registered.erase(ctx->registered.begin() + (long)i);
		return RT_OK;
	}
	return fail(ctx, RT_ERR_BAD_STATE, "surface %p was not registered with rt_register_surface", host_ptr);
}

int rt_clear_frame(rt_context* ctx, uint32_t pixel)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	DeviceState& d = ctx->devs[0];
	if (!d.d_frame || ctx->last_width <= 0 || ctx->last_height <= 0) return fail(ctx, RT_ERR_BAD_STATE, "no frame buffer yet (render or export a frame first)");
	RT_CUDA(ctx, cudaSetDevice(d.device));
	const size_t pixels = (size_t)ctx->last_width * (size_t)ctx->last_height;
	RT_CUDA(ctx, cuda_fill32(d.d_frame, pixel, pixels, d.stream));
	RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
	return RT_OK;
}

int rt_host_arrive_and_wait(volatile int64_t* words, int32_t stride_words, int32_t rank, int32_t world, int64_t frame, double timeout_seconds)
{
	if (!words || stride_words <= 0 || world <= 0 || rank < 0 || rank >= world) return RT_ERR_INVALID_ARGUMENT;
	__atomic_store_n(const_cast<int64_t*>(words) + (size_t)rank * stride_words, frame, __ATOMIC_RELEASE);
	const auto t0 = std::chrono::steady_clock::now();
	for (unsigned long long spins = 0;; ++spins)
	{
		bool all = true;
		for (int r = 0; r < world && all; ++r)
			all = __atomic_load_n(const_cast<int64_t*>(words) + (size_t)r * stride_words, __ATOMIC_ACQUIRE) >= frame;
		if (all) return RT_OK;
#if defined(__x86_64__)
		__builtin_ia32_pause();
#endif
		if ((spins & 0xffffu) == 0xffffu && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_seconds) return RT_ERR_BAD_STATE;
	}
}

int rt_get_timing(const rt_context* ctx, rt_timing* out_timing)
{
	if (!ctx || !out_timing) return RT_ERR_INVALID_ARGUMENT;
	*out_timing = ctx->timing;
	return RT_OK;
}

int rt_read_mesh_build(rt_context* ctx, int32_t mesh_id, int32_t* indices, float* normals, rt_built_node* nodes, int32_t node_capacity, int32_t* out_node_count)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (mesh_id < 0 || mesh_id >= (int32_t)ctx->meshes.size()) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "mesh id %d outside the announced count %d", mesh_id, (int)ctx->meshes.size());
	HostMesh& hm = ctx->meshes[(size_t)mesh_id];
	if (!hm.device_bvh) return fail(ctx, RT_ERR_BAD_STATE, "mesh %d is not built on the device (rt_set_mesh_device_bvh)", mesh_id);
	if (!hm.has_transform) return fail(ctx, RT_ERR_BAD_STATE, "mesh %d has not been transformed yet (rt_transform_mesh)", mesh_id);
	if (nodes && node_capacity < 0) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "negative node capacity");
	int rc = flush_uploads(ctx);       // runs the builds that are still pending, in call order
	if (rc != RT_OK) return rc;
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
	const DeviceState::MeshSourceDevice& sd = d.sources[(size_t)mesh_id];
	const size_t T = hm.src_indices.size() / 3;
	if (indices && T) RT_CUDA(ctx, cudaMemcpy(indices, sd.indices, 3 * T * sizeof(int32_t), cudaMemcpyDeviceToHost));
	if (normals && T) RT_CUDA(ctx, cudaMemcpy(normals, sd.normals, 3 * T * sizeof(float), cudaMemcpyDeviceToHost));
	int32_t info[8] = {};
	RT_CUDA(ctx, cudaMemcpy(info, sd.build.result_info, sizeof info, cudaMemcpyDeviceToHost));
	if (out_node_count) *out_node_count = info[0];
	if (nodes)
	{
		if (info[0] > node_capacity) return fail(ctx, RT_ERR_CAPACITY, "the build has %d nodes, the caller's array holds %d", info[0], node_capacity);
		std::vector<float4> records(2 * (size_t)info[0]);
		if (info[0]) RT_CUDA(ctx, cudaMemcpy(records.data(), sd.build.result_nodes, records.size() * sizeof(float4), cudaMemcpyDeviceToHost));
		for (int32_t n = 0; n < info[0]; ++n)
		{
			const float4 a = records[2 * (size_t)n], b = records[2 * (size_t)n + 1];
			int32_t hit, miss;
			memcpy(&hit, &b.z, sizeof hit); memcpy(&miss, &b.w, sizeof miss);
			rt_built_node& out = nodes[n];
			out.min_aabb[0] = a.x; out.max_aabb[0] = a.y; out.min_aabb[1] = a.z; out.max_aabb[1] = a.w; out.min_aabb[2] = b.x; out.max_aabb[2] = b.y;
			out.first = rt::BvhLink::is_leaf(hit) ? rt::BvhLink::leaf_first(hit) : hit / rt::BvhLink::kNodeBytes;
			out.triangle_count = rt::BvhLink::is_leaf(hit) ? rt::BvhLink::leaf_count(hit) : 0;
			out.escape = miss < 0 ? -1 : miss / rt::BvhLink::kNodeBytes;
		}
	}
	return RT_OK;
}

int rt_count_frame(rt_context* ctx, const rt_camera* camera, const rt_frame_desc* frame, int32_t mesh_path, rt_counters* out_counters)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!out_counters) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "out_counters must not be NULL");
	int rc = validate_frame(ctx, camera, frame);
	if (rc != RT_OK) return rc;
	if (mesh_path < RT_MESH_PATH_AUTO || mesh_path > RT_MESH_PATH_BVH) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "unknown mesh path %d", mesh_path);
	const int path = resolve_mesh_path(ctx, mesh_path);
	if (path < 0) return RT_ERR_BAD_STATE;
	DeviceState& d = ctx->devs[0];
	const size_t pixels = (size_t)frame->width * (size_t)frame->height;
	if ((rc = ensure_frame(ctx, d, pixels)) != RT_OK) return rc;
	RT_CUDA(ctx, cudaSetDevice(d.device));
	RT_CUDA(ctx, cudaMemsetAsync(d.d_counters, 0, sizeof(unsigned long long) * RT_COUNTER_SLOTS, d.stream));
	rt::FrameParams p = make_params(camera, frame);
	p.row_begin = 0; p.row_end = frame->height; p.strip_first = 0; p.strip_step = 1;
	p.dst_full_frame = 1; p.dst = d.d_frame; p.counters = d.d_counters;
	p.vector_store = (p.width % 4 == 0);
	const dim3 grid((unsigned)((p.width + rt::kBlockW - 1) / rt::kBlockW), (unsigned)((p.height + rt::kBlockH - 1) / rt::kBlockH), 1);
	rt::pick_kernel_count(path == RT_MESH_PATH_BVH)<<<grid, rt::kThreads, rt::dynamic_smem_bytes(rt::kThreads, ctx->n_materials), d.stream>>>(d.view, p);
	RT_CUDA(ctx, cudaGetLastError());
	RT_CUDA(ctx, cudaMemcpyAsync(out_counters->slot, d.d_counters, sizeof(unsigned long long) * RT_COUNTER_SLOTS, cudaMemcpyDeviceToHost, d.stream));
	RT_CUDA(ctx, cudaStreamSynchronize(d.stream));
	ctx->last_width = frame->width; ctx->last_height = frame->height;
	return RT_OK;
}

int rt_measure_fp32_peak(rt_context* ctx, int32_t use_fma, double* out_tflops, float* out_ms)
{
	if (!ctx) return RT_ERR_INVALID_ARGUMENT;
	if (!out_tflops) return fail(ctx, RT_ERR_INVALID_ARGUMENT, "out_tflops must not be NULL");
	DeviceState& d = ctx->devs[0];
	RT_CUDA(ctx, cudaSetDevice(d.device));
	cudaDeviceProp prop{};
	RT_CUDA(ctx, cudaGetDeviceProperties(&prop, d.device));
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iterations = 1 << 15;
	float* sink = reinterpret_cast<float*>(d.d_counters);
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep)   // first repetition warms the clocks up
	{
		RT_CUDA(ctx, cudaEventRecord(d.ev_begin, d.stream));
		if (use_fma) rt::fp32_peak_kernel<true><<<blocks, threads, 0, d.stream>>>(sink, 0.999f, 1e-3f, iterations);
		else rt::fp32_peak_kernel<false><<<blocks, threads, 0, d.stream>>>(sink, 0.999f, 1e-3f, iterations);
		RT_CUDA(ctx, cudaGetLastError());
		RT_CUDA(ctx, cudaEventRecord(d.ev_kernel, d.stream));
		RT_CUDA(ctx, cudaEventSynchronize(d.ev_kernel));
		float ms = 0.f;
		RT_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_begin, d.ev_kernel));
		if (rep > 0) best = std::min(best, ms);
	}
	const double flops = 2.0 * 8.0 * (double)iterations * (double)blocks * (double)threads;
	*out_tflops = flops / ((double)best * 1e-3) / 1e12;
	if (out_ms) *out_ms = best;
	return RT_OK;
}

} // extern "C"
