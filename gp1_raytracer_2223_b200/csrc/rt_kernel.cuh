// The pixel kernel: one thread per pixel, warps own 8x4 pixel tiles, a 256-thread CTA owns a
// 32x8 pixel strip segment.  Replaces Renderer::RenderPixel (reference
// source/Renderer.cpp:100-182) and everything it calls:
//   Scene::GetClosestHit / DoesHit        source/Scene.cpp:29-96
//   GeometryUtils::HitTest_* / SlabTest   source/Utils.h:15-216, 290-327 (both bodies of HitTest_TriangleMesh:
//                                         the #else slab + linear loop and the shipped BVH walk, Utils.h:221-288)
//   LightUtils                            source/Utils.h:341-369
//   Material::Shade x4, BRDF::*           source/Material.h:34-129, source/BRDFs.h:14-99
//   ColorRGB::MaxToOne + pack             source/ColorRGB.h:12-17, source/Renderer.cpp:176-181
//
// Data movement: spheres, planes, lights, materials and the mesh table are staged from the
// SoA upload buffers into shared memory once per CTA; triangles are streamed as three float4
// per triangle (v0|nx, e1|ny, e2|nz) with warp-uniform 128-bit loads that live in L1/L2; the
// only HBM traffic is the 4 B/pixel result, written with 128-bit stores.  This file also holds the small
// kernels around it: unstripe (multi-process gather tail), the FP32 peak probe, the frame completion
// signal and the device-side TriangleMesh::UpdateTransforms.
#pragma once

#include "rt_device.cuh"
#include "../../include/rt_b200.h"

namespace rt
{
	constexpr int kMaxSpheres = 64;
	constexpr int kMaxPlanes = 64;
	constexpr int kMaxLights = 16;
	constexpr int kMaxMaterials = 256;   // materialIndex is an unsigned char in the reference
	constexpr int kMaxMeshes = 32;

	constexpr int kTileW = 8;            // pixels per warp tile
	constexpr int kTileH = 4;
#ifndef RT_BLOCK_W
#define RT_BLOCK_W 32
#endif
	constexpr int kBlockW = RT_BLOCK_W;  // pixels per CTA strip segment
	constexpr int kBlockH = 8;
	constexpr int kThreads = kBlockW * kBlockH;
	constexpr int kWarpsX = kBlockW / kTileW;

	// Device views of the uploaded SoA buffers (one copy per GPU).
	struct SceneDevice
	{
		const float* sphere_ox; const float* sphere_oy; const float* sphere_oz; const float* sphere_r;
		const uint8_t* sphere_mat;
		const float* plane_ox; const float* plane_oy; const float* plane_oz;
		const float* plane_nx; const float* plane_ny; const float* plane_nz;
		const uint8_t* plane_mat;
		const float* light_ox; const float* light_oy; const float* light_oz;
		const float* light_r; const float* light_g; const float* light_b;
		const float* light_intensity;
		const int32_t* light_type;
		const float4* materials;       // 2 float4 per material: {tag bits, r, g, b} {p0, p1, p2, -}
		const float4* mesh_table;      // 3 float4 per mesh: {min.x, max.x, min.y, max.y}, {min.z, max.z, first triangle, triangle count}, {cull, material, first BVH node, BVH node count}
		const float4* triangles;       // 3 float4 per triangle: {v0 xyz, n.x} {e1 xyz, n.y} {e2 xyz, n.z}
		const float4* bvh_nodes;       // 2 float4 per node: {min.x, max.x, min.y, max.y}, {min.z, max.z, hit, miss}; see BvhLink
		int32_t n_spheres, n_planes, n_lights, n_materials, n_meshes;
		// (-0,-0), (1,1), (-1,-1): third operands of the packed FFMA2 arithmetic (struct Pk).  They are
		// kernel parameters on purpose - see rt_kernel_x2.cuh on why literals would break exactness.
		float2 k_neg0, k_one, k_mone;
	};

	struct FrameParams
	{
		float cam_ox, cam_oy, cam_oz, fov;
		float right_x, right_y, right_z;
		float up_x, up_y, up_z;
		float fwd_x, fwd_y, fwd_z;
		float aspect;
		int32_t width, height;
		int32_t row_begin, row_end;      // rows of the frame this launch may touch
		int32_t strip_first, strip_step; // CTA row k renders 8-row strip (k * strip_step + strip_first) counted from row_begin
		int32_t dst_full_frame;          // 1: dst addresses the whole frame (row = py); 0: dst is this launch's packed band
		int32_t lighting_mode, shadows;
		uint32_t r_shift, g_shift, b_shift, alpha_mask;
		int32_t vector_store;            // 1 when width % 4 == 0 and dst is 16-byte aligned
		uint32_t* dst;
		unsigned long long* counters;    // counters build only
		unsigned int* band_done;         // progressive present: band_done[b] counts the finished CTAs of band b (NULL = off)
		int32_t strips_per_band;         // a band = this many consecutive 8-row strips of the frame ...
		const uint8_t* band_table;       // ... or, when set, band_table[strip]: bands of unequal size (single-GPU counters only)
		// Band watcher (persistent kernel): CTA 0 does not render; one of its threads waits for each band's counter and then
		// stores watch_tag into host_flags[band] (mapped pinned memory) - the host thread blocked in rt_render polls those
		// words and issues the band's copy itself (no stream memory operations, see watch_bands)
		unsigned int* host_flags;        // NULL = off
		uint32_t watch_tag;
		int32_t watch_bands;
		unsigned int* band_local;        // multi-GPU: this GPU's own per-band counters (see signal_band_done)
		int32_t grid_x, n_strips;        // the launch's tile grid: 32-pixel columns x 8-row strips (set by launch())
		unsigned int* queue;             // persistent kernel: {next work item, finished warps}, both zero between launches
		uint32_t grid_x_magic;           // floor(2^32 / grid_x) + 1: tile / grid_x == __umulhi(tile, magic) while tile * grid_x < 2^32
		// persistent kernel, scheduling only: the tiles of the rectangle [first_x0, first_x1) x [first_k0, first_k1)
		// (tile columns x strips of this launch) are handed out before all others (first_tiles == 0: plain order)
		int32_t first_x0, first_x1, first_k0, first_k1;
		int32_t first_tiles;             // (first_x1 - first_x0) * (first_k1 - first_k0)
		uint32_t first_w_magic;          // floor(2^32 / (first_x1 - first_x0)) + 1
		uint32_t rest_w_magic;           // floor(2^32 / (grid_x - (first_x1 - first_x0))) + 1 (unused when the rectangle spans the width)
		// ... or cell by cell: the tile grid is cut into cells of (1 << cell_w_log2) columns x (1 << cell_h_log2) strips, and the queue
		// walks them in the order of cell_order (most expensive cell of the previous launch first).  Cells on the right
		// and bottom edge are padded: queue positions that fall outside the grid render nothing.
		const uint16_t* cell_order;      // cell_order[r] = r-th cell to render (NULL = off)
		unsigned int* cell_cost;         // += SM clocks / 16 of every warp tile, per cell (NULL = off)
		int32_t cells_x, cell_w_log2, cell_h_log2;
		uint32_t cells_x_magic;          // floor(2^32 / cells_x) + 1
		int32_t total_items;             // queue positions of the launch when cells pad it (0 = grid_x * n_strips * kSignalsPerTile)
	};

	struct Ray
	{
		V3 o, d, inv;
		float tmin, tmax;
		bool nan_safe;   // no component of inv is infinite: the slab products cannot be 0 * inf = NaN
	};

	struct Hit
	{
		V3 origin, normal;
		float t;
		int material;
		bool did;
	};

	struct SharedScene
	{
		float4 sphere[kMaxSpheres];          // {ox, oy, oz, radius}
		float4 plane_o[kMaxPlanes];          // {ox, oy, oz, material bits}
		float4 plane_n[kMaxPlanes];          // {nx, ny, nz, (plane origin - camera origin) . n: HitTest_Plane's numerator for view rays}
		float4 light_a[kMaxLights];          // {ox, oy, oz, intensity}
		float4 light_b[kMaxLights];          // {r, g, b, type bits}
		float4 mesh[3 * kMaxMeshes];
		float4 sphere_view[kMaxSpheres];     // {c - camera origin, |c - camera origin|^2}: HitTest_Sphere's ray-independent part for view rays
		// the planes once more, two per record, laid out for the packed plane loops (plane_pair_*): planes 2j and 2j + 1 as
		// {ox0, ox1, oy0, oy1} {oz0, oz1, nx0, nx1} {ny0, ny1, nz0, nz1}; an odd count is padded with a copy of the last
		// plane (same test, same result: an any-hit cannot change, and a closest hit keeps the first of two equal t)
		float4 plane_pair[3 * (kMaxPlanes / 2)];
		float2 plane_pair_view[kMaxPlanes / 2];   // {num0, num1} of HitTest_Plane for view rays, as plane_n[].w
		// per mesh: the addresses of its first triangle record and its first node record (two 64-bit pointers as bits),
		// so that a ray entering a mesh does no pointer arithmetic
		float4 mesh_ptr[kMaxMeshes];
		uint8_t sphere_mat[kMaxSpheres];
		// LAST, and only its first 2 * n_materials records exist in the scalar kernels' dynamic shared memory
		// (staged_scene_bytes): a full table is 8 KB per CTA for scenes that use a handful of materials
		float4 material[2 * kMaxMaterials];
	};
	inline __host__ __device__ size_t staged_scene_bytes(int n_materials) { return offsetof(SharedScene, material) + sizeof(float4) * 2 * (size_t)n_materials; }


	// The staged scene seen from the pixel code: ONE 32-bit shared-memory address plus compile-time field offsets, read
	// with explicit ld.shared.  (Passing `SharedScene&` around makes nvcc rebuild the block's shared-window base - S2R
	// SR_CgaCtaId, MOV, LEA - in every iteration of every loop that touches the scene.)  The handle is created after the
	// staging barrier and made opaque there, so no load can be scheduled above the barrier.
	struct Staged
	{
		unsigned int base;
		static __device__ __forceinline__ float4 ld4(unsigned int a) { float4 v; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
		static __device__ __forceinline__ float2 ld2(unsigned int a) { float2 v; asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
		static __device__ __forceinline__ int ld8(unsigned int a) { int v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#define RT_STAGED_F4(name, field) __device__ __forceinline__ float4 name(int i) const { return ld4(base + (unsigned int)offsetof(SharedScene, field) + 16u * (unsigned int)i); }
		RT_STAGED_F4(sphere, sphere) RT_STAGED_F4(sphere_view, sphere_view) RT_STAGED_F4(plane_o, plane_o) RT_STAGED_F4(plane_n, plane_n)
		RT_STAGED_F4(light_a, light_a) RT_STAGED_F4(light_b, light_b) RT_STAGED_F4(mesh, mesh) RT_STAGED_F4(material, material) RT_STAGED_F4(plane_pair, plane_pair) RT_STAGED_F4(mesh_ptr, mesh_ptr)
#undef RT_STAGED_F4
		__device__ __forceinline__ float2 plane_pair_view(int j) const { return ld2(base + (unsigned int)offsetof(SharedScene, plane_pair_view) + 8u * (unsigned int)j); }
		__device__ __forceinline__ int sphere_mat(int i) const { return ld8(base + (unsigned int)offsetof(SharedScene, sphere_mat) + (unsigned int)i); }
	};
	// call after the __syncthreads() that ends stage_scene
	__device__ __forceinline__ Staged staged_handle(const SharedScene& sc)
	{
		Staged s;
		s.base = (unsigned int)__cvta_generic_to_shared(&sc);
		asm volatile("" : "+r"(s.base) :: "memory");
		return s;
	}

	template <bool COUNT>
	struct Counters
	{
		__device__ __forceinline__ void hit(int) {}
		__device__ __forceinline__ void flush(unsigned long long*) {}
		__device__ __forceinline__ void flush_partial(unsigned long long*) {}
	};
	template <>
	struct Counters<true>
	{
		unsigned int c[RT_COUNTER_SLOTS];
		__device__ Counters() { for (int i = 0; i < RT_COUNTER_SLOTS; ++i) c[i] = 0; }
		__device__ __forceinline__ void hit(int slot) { c[slot]++; }
		// one thread's counts (the out-of-line literal walk: divergent, so no warp reduction)
		__device__ void flush_partial(unsigned long long* out)
		{
			for (int i = 0; i < RT_COUNTER_SLOTS; ++i) if (c[i]) atomicAdd(&out[i], (unsigned long long)c[i]);
		}
		__device__ void flush(unsigned long long* out)
		{
			for (int i = 0; i < RT_COUNTER_SLOTS; ++i)
			{
				const unsigned int s = __reduce_add_sync(0xffffffffu, c[i]);
				if ((threadIdx.x & 31) == 0 && s) atomicAdd(&out[i], (unsigned long long)s);
			}
		}
	};

	// Packed FP32 arithmetic (Blackwell FFMA2): two independent IEEE multiplies / adds / subtracts per issue
	// slot, each individually rounded:  a*b = fma(a, b, -0),  a+b = fma(a, 1, b),  a-b = fma(b, -1, a).
	// The constants are run-time values (SceneDevice::k_*), never literals (rt_kernel_x2.cuh explains why).
	struct Pk
	{
		float2 neg0, one, mone;
		__device__ __forceinline__ float2 mul(float2 a, float2 b) const { return __ffma2_rn(a, b, neg0); }
		__device__ __forceinline__ float2 add(float2 a, float2 b) const { return __ffma2_rn(a, one, b); }
		__device__ __forceinline__ float2 sub(float2 a, float2 b) const { return __ffma2_rn(b, mone, a); }
	};
	__device__ __forceinline__ Pk make_pk(const SceneDevice& dev) { Pk k; k.neg0 = dev.k_neg0; k.one = dev.k_one; k.mone = dev.k_mone; return k; }
	__device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }

	__device__ __forceinline__ Ray make_ray(V3 o, V3 d, float tmin, float tmax)
	{
		Ray r;
		r.o = o; r.d = d;
		r.inv = v3(rcp(d.x), rcp(d.y), rcp(d.z));   // DataTypes.h:550-563
		r.tmin = tmin; r.tmax = tmax;
		r.nan_safe = (fabsf(r.inv.x) < INFINITY) && (fabsf(r.inv.y) < INFINITY) && (fabsf(r.inv.z) < INFINITY);
		return r;
	}

	// HitTest_Sphere, Utils.h:52-71.  Returns t through `t_out`.  `ov` = centre - ray origin and `ov2` = ov . ov
	// are passed in: every view ray of a frame shares them (staged once per CTA), shadow rays compute them here.
	template <bool SHADOW, bool COUNT>
	__device__ __forceinline__ bool hit_sphere_from(const V3 ov, const float ov2, const float radius, const Ray& ray, float& t_out, Counters<COUNT>& cnt)
	{
		const float p = dot(ray.d, ov);
		const float perp = sub(ov2, mul(p, p));
		const float r2 = mul(radius, radius);
		if (r2 < perp) { cnt.hit(SHADOW ? RT_CNT_SPHERE_S_DISC : RT_CNT_SPHERE_P_DISC); return false; }
		const float t = sub(p, root(sub(r2, perp)));
		if (t < ray.tmin || t > ray.tmax) { cnt.hit(SHADOW ? RT_CNT_SPHERE_S_TREJ : RT_CNT_SPHERE_P_TREJ); return false; }
		cnt.hit(SHADOW ? RT_CNT_SPHERE_S_HIT : RT_CNT_SPHERE_P_HIT);
		t_out = t;
		return true;
	}

	template <bool SHADOW, bool COUNT>
	__device__ __forceinline__ bool hit_sphere(const float4 s, const Ray& ray, float& t_out, Counters<COUNT>& cnt)
	{
		const V3 ov = v3(s) - ray.o;
		return hit_sphere_from<SHADOW>(ov, dot(ov, ov), s.w, ray, t_out, cnt);
	}

	// Can t = RN(num / den) satisfy tmin <= t < limit (tmin = 1e-4 > 0)?  Returns false only when that is impossible,
	// so the IEEE division is skipped for the planes a ray cannot reach:
	//   opposite signs                              -> t <= -0 < tmin
	//   |num| > |den| * limit * (1 + 1e-6)          -> |num / den| > limit * (1 + 7e-7) in exact arithmetic (the two
	//                                                  roundings of the bound cost < 2.4e-7), and rounding is monotonic:
	//                                                  t >= limit.  Not used when the bound is so small that it may
	//                                                  have lost precision to underflow.
	// Zero and NaN operands fall through to the exact test (0 / x, x / 0 and NaN all fail its range check there).
	// `limit_up` = limit * 1.000001f, computed once per ray by the caller.
	__device__ __forceinline__ bool plane_may_hit(float num, float den, float limit_up)
	{
		const float bound = mul(fabsf(den), limit_up);
		if ((__float_as_int(num) ^ __float_as_int(den)) < 0) return false;
		if (bound > 1e-30f && fabsf(num) > bound) return false;
		return true;
	}

	// HitTest_Plane, Utils.h:82-98.  `num` = (plane origin - ray origin) . n is passed in for the same reason.
	template <bool SHADOW, bool COUNT>
	__device__ __forceinline__ bool hit_plane_from(const float num, const float4 pn, const Ray& ray, float& t_out, Counters<COUNT>& cnt)
	{
		const float den = dot(ray.d, v3(pn));
		cnt.hit(SHADOW ? RT_CNT_PLANE_S_TEST : RT_CNT_PLANE_P_TEST);
		if (!plane_may_hit(num, den, mul(ray.tmax, 1.000001f))) return false;
		const float t = quo(num, den);
		if (t >= ray.tmin && t < ray.tmax) { t_out = t; return true; }
		return false;
	}

	template <bool SHADOW, bool COUNT>
	__device__ __forceinline__ bool hit_plane(const float4 po, const float4 pn, const Ray& ray, float& t_out, Counters<COUNT>& cnt)
	{
		return hit_plane_from<SHADOW>(dot(v3(po) - ray.o, v3(pn)), pn, ray, t_out, cnt);
	}

	// HitTest_Plane for two planes at once (Utils.h:82-98): the numerators (plane origin - ray origin) . n and the
	// denominators d . n of planes 2j and 2j + 1 in the lanes of packed FFMA2s, every operation individually rounded
	// in the reference's order (Pk) - 13 instructions per pair of shadow-ray tests instead of 26, 5 instead of 10 for
	// view rays, whose numerators are staged per CTA.
	__device__ __forceinline__ float2 plane_pair_den(const Pk& K, const float4 b, const float4 c, const V3 d)
	{
		const float2 nx = make_float2(b.z, b.w), ny = make_float2(c.x, c.y), nz = make_float2(c.z, c.w);
		return K.add(K.add(K.mul(splat(d.x), nx), K.mul(splat(d.y), ny)), K.mul(splat(d.z), nz));
	}
	__device__ __forceinline__ float2 plane_pair_num(const Pk& K, const float4 a, const float4 b, const float4 c, const V3 o)
	{
		const float2 nx = make_float2(b.z, b.w), ny = make_float2(c.x, c.y), nz = make_float2(c.z, c.w);
		const float2 dx = K.sub(make_float2(a.x, a.y), splat(o.x)), dy = K.sub(make_float2(a.z, a.w), splat(o.y)), dz = K.sub(make_float2(b.x, b.y), splat(o.z));
		return K.add(K.add(K.mul(dx, nx), K.mul(dy, ny)), K.mul(dz, nz));
	}

	// Any plane between a shadow ray's origin and its light?  (Scene::DoesHit's plane loop, Scene.cpp:79-85)
	template <bool COUNT>
	__device__ __forceinline__ bool planes_any(const Pk& K, const Staged sc, int n_planes, const Ray& ray, Counters<COUNT>& cnt)
	{
		const float limit_up = mul(ray.tmax, 1.000001f);
#pragma unroll 2      // two pairs per trip: measured best on the 4K frames (1: +3 %, 3: +1 %)
		for (int j = 0; 2 * j < n_planes; ++j)
		{
			const float4 a = sc.plane_pair(3 * j), b = sc.plane_pair(3 * j + 1), c = sc.plane_pair(3 * j + 2);
			const float2 num = plane_pair_num(K, a, b, c, ray.o), den = plane_pair_den(K, b, c, ray.d);
			cnt.hit(RT_CNT_PLANE_S_TEST);
			if (2 * j + 1 < n_planes) cnt.hit(RT_CNT_PLANE_S_TEST);
			if (plane_may_hit(num.x, den.x, limit_up) || plane_may_hit(num.y, den.y, limit_up))
			{
				// rare: the exact test for both planes of the pair (the filter only ever skips planes whose exact test fails)
				const float t0 = quo(num.x, den.x), t1 = quo(num.y, den.y);
				if ((t0 >= ray.tmin && t0 < ray.tmax) || (t1 >= ray.tmin && t1 < ray.tmax)) return true;
			}
		}
		return false;
	}

	// Closest plane in front of a view ray (Scene::GetClosestHit's plane loop, Scene.cpp:45-53): planes in order, strict
	// '<' so the first one wins ties.  `best_t` comes in as the closest sphere's t.  Only planes that can still beat
	// best_t reach the division (the counters build keeps the reference's own range so that its counts are the reference's).
	template <bool COUNT>
	__device__ __forceinline__ void planes_closest(const Pk& K, const Staged sc, int n_planes, const Ray& ray, float& best_t, int& best_plane, Counters<COUNT>& cnt)
	{
#pragma unroll 1
		for (int j = 0; 2 * j < n_planes; ++j)
		{
			const float4 b = sc.plane_pair(3 * j + 1), c = sc.plane_pair(3 * j + 2);
			const float2 num = sc.plane_pair_view(j), den = plane_pair_den(K, b, c, ray.d);
			cnt.hit(RT_CNT_PLANE_P_TEST);
			if (2 * j + 1 < n_planes) cnt.hit(RT_CNT_PLANE_P_TEST);
			const float limit_up = mul(COUNT ? ray.tmax : best_t, 1.000001f);
			if (plane_may_hit(num.x, den.x, limit_up))
			{
				const float t = quo(num.x, den.x);
				if (t >= ray.tmin && t < ray.tmax) { cnt.hit(RT_CNT_PLANE_P_HIT); if (t < best_t) { best_t = t; best_plane = 2 * j; } }
			}
			// (the second plane is filtered against the limit of before the first one: still a valid limit, only looser)
			if ((!COUNT || 2 * j + 1 < n_planes) && plane_may_hit(num.y, den.y, limit_up))
			{
				const float t = quo(num.y, den.y);
				if (t >= ray.tmin && t < ray.tmax) { cnt.hit(RT_CNT_PLANE_P_HIT); if (t < best_t) { best_t = t; best_plane = 2 * j + 1; } }
			}
		}
	}

	// SlabTest_TriangleMesh / SlabTest_BVH, Utils.h:194-216, 221-243, on a box stored as
	// b0 = {min.x, max.x, min.y, max.y}, b1 = {min.z, max.z, -, -}: the (tx1, tx2), (ty1, ty2), (tz1, tz2)
	// pairs are six packed operations instead of twelve scalar ones.
	//
	// FAST = true replaces the std::min / std::max ternaries by FMNMX (fminf / fmaxf).  The two only
	// differ when an operand is NaN (and in the sign of a zero result, which the final comparisons
	// cannot see).  A NaN can only come from 0 * inf, i.e. from an infinite 1/dir component, so rays
	// whose inverse direction is finite (Ray::nan_safe, all but axis-parallel rays) take the fast
	// form with the same boolean result; the others take the literal one.
	template <bool FAST>
	__device__ __forceinline__ bool slab_test(const Pk&, const float4 b0, const float4 b1, const Ray& ray)
	{
		// (b - o) * inv per pair: b - o = fma(o, -1, b) and x * inv = fma(x, inv, +0), each rounded once like the FSUB / FMUL
		// they replace.  The constants are literals here (unlike Pk's): ptxas can only contract a multiply INTO a
		// following add, and these products feed nothing but min / max and comparisons.  The products are a * b + (+0),
		// not a * b + (-0): the two differ only in the sign of a zero product, which those cannot see either.
		const float2 zero = make_float2(0.f, 0.f), mone = make_float2(-1.f, -1.f);
		const float2 tx = __ffma2_rn(__ffma2_rn(splat(ray.o.x), mone, make_float2(b0.x, b0.y)), splat(ray.inv.x), zero);
		const float2 ty = __ffma2_rn(__ffma2_rn(splat(ray.o.y), mone, make_float2(b0.z, b0.w)), splat(ray.inv.y), zero);
		const float2 tz = __ffma2_rn(__ffma2_rn(splat(ray.o.z), mone, make_float2(b1.x, b1.y)), splat(ray.inv.z), zero);
		if (FAST)
		{
			const float t_min = fmaxf(fmaxf(fminf(tx.x, tx.y), fminf(ty.x, ty.y)), fminf(tz.x, tz.y));
			const float t_max = fminf(fminf(fmaxf(tx.x, tx.y), fmaxf(ty.x, ty.y)), fmaxf(tz.x, tz.y));
			return t_max > 0 && t_max >= t_min;
		}
		float t_min = std_min(tx.x, tx.y);
		float t_max = std_max(tx.x, tx.y);
		t_min = std_max(t_min, std_min(ty.x, ty.y));
		t_max = std_min(t_max, std_max(ty.x, ty.y));
		t_min = std_max(t_min, std_min(tz.x, tz.y));
		t_max = std_min(t_max, std_max(tz.x, tz.y));
		return t_max > 0 && t_max >= t_min;
	}

	// One triangle of the stream: {v0 xyz, n.x} {e1 xyz, n.y} {e2 xyz, n.z}, e1 = v1 - v0 and
	// e2 = v2 - v0 precomputed at upload (reference source/Utils.h:143-144: the same two IEEE
	// subtractions wherever they run).
	struct Tri
	{
		float4 a0, a1, a2;
	};
	__device__ __forceinline__ Tri load_tri(const float4* p)
	{
		Tri t;
		t.a0 = __ldg(p); t.a1 = __ldg(p + 1); t.a2 = __ldg(p + 2);
		return t;
	}

	// The two early outs at the top of HitTest_Triangle (Utils.h:111-139) folded into one
	// comparison per cull mode.  CULL is the mode that applies to this ray kind (already
	// inverted for shadow rays, Utils.h:114-127).  Same truth table as the reference, NaN included:
	//   front-face culling: reject if |c| < eps or c < 0   <=>  pass iff !(c < eps)
	//   back-face culling:  reject if |c| < eps or c > 0   <=>  pass iff !(c > -eps)
	//   no culling:         reject if |c| < eps
	template <int CULL>
	__device__ __forceinline__ bool cull_pass(float c)
	{
		if (CULL == RT_CULL_FRONT_FACE) return !(c < FLT_EPSILON);
		if (CULL == RT_CULL_BACK_FACE) return !(c > -FLT_EPSILON);
		return !(fabsf(c) < FLT_EPSILON);
	}

	// Moeller-Trumbore body of HitTest_Triangle, Utils.h:143-160, after the cull test passed.
	// `s` = ray origin - v0 is passed in so that rays sharing an origin share it.
	template <bool SHADOW, bool COUNT>
	__device__ __forceinline__ bool triangle_body(const Tri& T, V3 o, V3 d, V3 s, float tmin, float tmax, float& t_out, Counters<COUNT>& cnt)
	{
		constexpr int base = SHADOW ? RT_CNT_TRI_S_CULLED : RT_CNT_TRI_P_CULLED;
		const V3 e1 = v3(T.a1), e2 = v3(T.a2);
		const V3 h = cross(d, e2);
		const float a = dot(e1, h);
		if (fabsf(a) < FLT_EPSILON) { cnt.hit(base + 1); return false; }

		const float f = rcp(a);
		const float u = mul(f, dot(s, h));
		if (u < 0.f || u > 1.f) { cnt.hit(base + 2); return false; }

		const V3 q = cross(s, e1);
		const float v = mul(f, dot(d, q));
		if (v < 0.f || add(u, v) > 1.f) { cnt.hit(base + 3); return false; }

		const float t = mul(f, dot(e2, q));
		if (t < tmin || t >= tmax) { cnt.hit(base + 4); return false; }
		cnt.hit(base + 5);
		t_out = t;
		return true;
	}

	// Closest hit over one mesh's triangles (HitTest_TriangleMesh, Utils.h:300-324, #else branch):
	// every triangle in upload order, strict '<' so the first triangle wins ties.  The stream is
	// software-pipelined one triangle ahead (the buffer carries one padding record at its end).
	template <int CULL, bool COUNT>
	__device__ __forceinline__ void closest_one(const Tri& T, int i, const Ray& ray, float& best_t, int& best_tri, Counters<COUNT>& cnt)
	{
		const float c = dot(v3(T.a0.w, T.a1.w, T.a2.w), ray.d);
		if (cull_pass<CULL>(c))
		{
			float t;
			if (triangle_body<false>(T, ray.o, ray.d, ray.o - v3(T.a0), ray.tmin, ray.tmax, t, cnt))
			{
				if (t < best_t) { best_t = t; best_tri = i; }
			}
		}
		else cnt.hit(RT_CNT_TRI_P_CULLED);
	}

	template <int CULL, bool COUNT>
	__device__ __forceinline__ void mesh_closest(const float4* tri, int count, const Ray& ray, float& best_t, int& best_tri, Counters<COUNT>& cnt)
	{
		// two records in flight, ping-pong (the buffer carries two padding records at its end)
		Tri A = load_tri(tri);
		for (int i = 0; i < count; i += 2)
		{
			const Tri B = load_tri(tri + 3 * (i + 1));
			closest_one<CULL>(A, i, ray, best_t, best_tri, cnt);
			A = load_tri(tri + 3 * (i + 2));
			if (i + 1 < count) closest_one<CULL>(B, i + 1, ray, best_t, best_tri, cnt);
		}
	}

	template <int CULL, bool COUNT>
	__device__ __forceinline__ bool shadow_one(const Tri& T, const Ray& ray, Counters<COUNT>& cnt)
	{
		const float c = dot(v3(T.a0.w, T.a1.w, T.a2.w), ray.d);
		if (!cull_pass<CULL>(c)) { cnt.hit(RT_CNT_TRI_S_CULLED); return false; }
		float t;
		return triangle_body<true>(T, ray.o, ray.d, ray.o - v3(T.a0), ray.tmin, ray.tmax, t, cnt);
	}

	// Any hit over one mesh's triangles (HitTest_TriangleMesh with ignoreHitRecord, Utils.h:300-324):
	// upload order, stop at the first hit.
	template <int CULL, bool COUNT>
	__device__ __forceinline__ bool mesh_any(const float4* tri, int count, const Ray& ray, Counters<COUNT>& cnt)
	{
		Tri A = load_tri(tri);
		for (int i = 0; i < count; i += 2)
		{
			const Tri B = load_tri(tri + 3 * (i + 1));
			if (shadow_one<CULL>(A, ray, cnt)) return true;
			A = load_tri(tri + 3 * (i + 2));
			if (i + 1 < count && shadow_one<CULL>(B, ray, cnt)) return true;
		}
		return false;
	}

	// ---- the reference's shipped mesh path: IntersectionTest_BVH, Utils.h:246-288 ------------------
	//
	// The nodes are the reference's own (TriangleMesh::BuildBVH output, uploaded as they are), so the
	// box tests, the set of triangles a ray meets and their order are the reference's by
	// construction.  The recursion (left child, then left + 1) is unrolled at upload time into a
	// threaded tree: every node carries the index of the node that follows its subtree in that
	// depth-first order ("escape"), so the walk needs no stack:
	//     hit box:  inner -> first child;  leaf -> test its triangles, then escape
	//     miss box: escape
	// Leaves own contiguous triangle ranges in ascending order, so triangles are met in upload
	// order, exactly as the recursion meets them (strict '<' tie-breaking is preserved).
	struct BvhLink
	{
		// The two link words of a node record {min.x, max.x, min.y, max.y} {min.z, max.z, hit, miss}:
		//   hit   where the walk goes when the ray meets the box: an inner node's left child as a BYTE offset from the
		//         mesh's first node record (>= 0, ready to be added to the base address), or, with the sign bit set, a
		//         leaf's triangles: kLeafBit | count << kFirstBits | first triangle (relative to the mesh)
		//   miss  the escape link as a byte offset: the node that follows this subtree in the reference's depth-first
		//         order; -1 ends the walk
		// A walk step is therefore: test the box, take `hit` or `miss`, add it to the base - no index arithmetic, no
		// decode except on leaves.
		static constexpr int kNodeBytes = 32;
		static constexpr int kFirstBits = 20;
		static constexpr int kFirstMask = (1 << kFirstBits) - 1;
		static constexpr int kMaxLeafTriangles = (1 << (31 - kFirstBits)) - 1;
		static constexpr int kMaxNodes = (1 << 25) - 1;                     // byte offsets stay positive ints
		static constexpr int kEnd = -1;
		static constexpr unsigned int kLeafBit = 0x80000000u;
		__host__ __device__ static int inner(int left_child) { return left_child * kNodeBytes; }
		__host__ __device__ static int leaf(int first_triangle, int count) { return (int)(kLeafBit | ((unsigned int)count << kFirstBits) | (unsigned int)first_triangle); }
		__host__ __device__ static int miss(int escape_node) { return escape_node < 0 ? kEnd : escape_node * kNodeBytes; }
		__host__ __device__ static bool is_leaf(int hit) { return hit < 0; }
		__host__ __device__ static int leaf_first(int hit) { return hit & kFirstMask; }
		__host__ __device__ static int leaf_count(int hit) { return (int)(((unsigned int)hit & ~kLeafBit) >> kFirstBits); }
	};
	__device__ __forceinline__ const float4* node_at(const float4* nodes, int offset)
	{
		return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(nodes) + (unsigned int)offset);
	}

	// One leaf of the walk: its triangles in ascending order (the order the recursion meets them).  The cull mode is a
	// warp-uniform switch here, not a template parameter of the whole walk: the node loop does not depend on it.
	template <bool COUNT>
	__device__ __forceinline__ void leaf_closest(int cull, const float4* tri, int first, int count, const Ray& ray, float& best_t, int& best_tri, Counters<COUNT>& cnt)
	{
		if (cull == RT_CULL_BACK_FACE) { for (int k = 0; k < count; ++k) closest_one<RT_CULL_BACK_FACE>(load_tri(tri + 3 * (first + k)), first + k, ray, best_t, best_tri, cnt); }
		else if (cull == RT_CULL_FRONT_FACE) { for (int k = 0; k < count; ++k) closest_one<RT_CULL_FRONT_FACE>(load_tri(tri + 3 * (first + k)), first + k, ray, best_t, best_tri, cnt); }
		else { for (int k = 0; k < count; ++k) closest_one<RT_CULL_NONE>(load_tri(tri + 3 * (first + k)), first + k, ray, best_t, best_tri, cnt); }
	}

	// `cull` is the mode that applies to shadow rays (already inverted, Utils.h:114-127)
	template <bool COUNT>
	__device__ __forceinline__ bool leaf_any(int cull, const float4* tri, int first, int count, const Ray& ray, Counters<COUNT>& cnt)
	{
		if (cull == RT_CULL_BACK_FACE) { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_BACK_FACE>(load_tri(tri + 3 * (first + k)), ray, cnt)) return true; }
		else if (cull == RT_CULL_FRONT_FACE) { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_FRONT_FACE>(load_tri(tri + 3 * (first + k)), ray, cnt)) return true; }
		else { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_NONE>(load_tri(tri + 3 * (first + k)), ray, cnt)) return true; }
		return false;
	}

	// The walk.  ANY = DoesHit's form (stop at the first triangle hit: the reference only leaves the leaf and keeps
	// walking, which cannot change the boolean), else GetClosestHit's (strict '<': the first triangle in walk order
	// wins ties).  FAST: see slab_test.
	template <bool ANY, bool FAST, bool COUNT>
	__device__ __forceinline__ bool bvh_walk(const Pk& K, int cull, const float4* nodes, const float4* tri, const Ray& ray, float& best_t, int& best_tri, Counters<COUNT>& cnt)
	{
		// the 64-bit base stays in a register pair (made opaque: ptxas would otherwise rebuild it from the constant bank
		// and the mesh table in every iteration - five instructions instead of one)
		unsigned long long base = reinterpret_cast<unsigned long long>(nodes);
		asm volatile("" : "+l"(base));
		int at = 0;
		do
		{
			const float4* rec = reinterpret_cast<const float4*>(base + (unsigned int)at);
			const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1);
			cnt.hit(ANY ? RT_CNT_BVH_S_NODE : RT_CNT_BVH_P_NODE);
			const bool inside = slab_test<FAST>(K, n0, n1, ray);
			const int hit = __float_as_int(n1.z), miss = __float_as_int(n1.w);
			at = inside ? hit : miss;
			if (inside && BvhLink::is_leaf(hit))
			{
				const int first = BvhLink::leaf_first(hit), count = BvhLink::leaf_count(hit);
				if (ANY) { if (leaf_any(cull, tri, first, count, ray, cnt)) return true; }
				else leaf_closest(cull, tri, first, count, ray, best_t, best_tri, cnt);
				at = miss;
			}
		} while (at >= 0);
		return false;
	}

	// Rays with an infinite 1 / dir component (axis-parallel: a handful per frame at most) take the literal
	// std::min / std::max form of the slab test: a second inline copy of the walk that is practically never executed
	// (it costs code size, not registers or instruction-cache footprint).
	template <bool COUNT>
	__device__ __forceinline__ void bvh_closest_any_cull(const Pk& K, int cull, const float4* nodes, const float4* tri, const Ray& ray, float& best_t, int& best_tri, Counters<COUNT>& cnt)
	{
		if (ray.nan_safe) bvh_walk<false, true>(K, cull, nodes, tri, ray, best_t, best_tri, cnt);
		else bvh_walk<false, false>(K, cull, nodes, tri, ray, best_t, best_tri, cnt);
	}

	// Utils.h:114-127: shadow rays see the opposite cull mode
	template <bool COUNT>
	__device__ __forceinline__ bool bvh_any_any_cull(const Pk& K, int cull, const float4* nodes, const float4* tri, const Ray& ray, Counters<COUNT>& cnt)
	{
		const int shadow_cull = cull == RT_CULL_BACK_FACE ? RT_CULL_FRONT_FACE : (cull == RT_CULL_FRONT_FACE ? RT_CULL_BACK_FACE : RT_CULL_NONE);
		float t = FLT_MAX; int tri_id = -1;
		if (ray.nan_safe) return bvh_walk<true, true>(K, shadow_cull, nodes, tri, ray, t, tri_id, cnt);
		return bvh_walk<true, false>(K, shadow_cull, nodes, tri, ray, t, tri_id, cnt);
	}

	__device__ __forceinline__ const float4* pointer_from_bits(float lo, float hi)
	{
		return reinterpret_cast<const float4*>(((unsigned long long)(unsigned int)__float_as_int(hi) << 32) | (unsigned long long)(unsigned int)__float_as_int(lo));
	}

	// Scene::GetClosestHit, Scene.cpp:29-66: spheres, planes, meshes in order; strict '<' keeps
	// the first primitive on ties.  The reference's shared scratch HitRecord never changes the
	// outcome (its stale t is always >= the running closest t), so a plain running minimum is
	// the same function.
	template <bool BVH, bool COUNT>
	__device__ __forceinline__ Hit closest_hit(const Staged sc, const SceneDevice& dev, const Ray& ray, Counters<COUNT>& cnt, unsigned long long* dev_counters)
	{
		const Pk K = make_pk(dev);
		Hit best;
		best.t = FLT_MAX; best.did = false; best.material = 0;
		best.origin = v3(0.f, 0.f, 0.f); best.normal = v3(0.f, 0.f, 0.f);

		int best_sphere = -1;
#pragma unroll 1
		for (int i = 0; i < dev.n_spheres; ++i)
		{
			float t;
			const float4 sv = sc.sphere_view(i);
			if (hit_sphere_from<false>(v3(sv), sv.w, sc.sphere(i).w, ray, t, cnt))
			{
				if (t < best.t) { best.t = t; best_sphere = i; cnt.hit(RT_CNT_SPHERE_P_CLOSEST); }
			}
		}
		if (best_sphere >= 0)
		{
			const float4 s = sc.sphere(best_sphere);
			best.did = true;
			best.material = sc.sphere_mat(best_sphere);
			best.origin = ray.o + ray.d * best.t;          // Utils.h:67
			best.normal = best.origin - v3(s);             // Utils.h:68
			normalize(best.normal);                        // Scene.cpp:40
		}

		int best_plane = -1;
		planes_closest(K, sc, dev.n_planes, ray, best.t, best_plane, cnt);
		if (best_plane >= 0)
		{
			best.did = true;
			best.material = __float_as_int(sc.plane_o(best_plane).w);
			best.normal = v3(sc.plane_n(best_plane));          // Utils.h:91
			best.origin = ray.o + ray.d * best.t;              // Utils.h:92
		}

#pragma unroll 1
		for (int m = 0; m < dev.n_meshes; ++m)
		{
			const float4 info = sc.mesh(3 * m + 2), ptrs = sc.mesh_ptr(m);
			const int cull = __float_as_int(info.x);
			const float4* tri = pointer_from_bits(ptrs.x, ptrs.y);
			int best_tri = -1;
			if (BVH)
			{
				if (__float_as_int(info.w) == 0) continue;           // no nodes: an empty mesh
				bvh_closest_any_cull(K, cull, pointer_from_bits(ptrs.z, ptrs.w), tri, ray, best.t, best_tri, cnt);
			}
			else
			{
				const float4 b0 = sc.mesh(3 * m), b1 = sc.mesh(3 * m + 1);
				const int count = __float_as_int(b1.w);
				cnt.hit(RT_CNT_SLAB_P_TEST);
				if (!(ray.nan_safe ? slab_test<true>(K, b0, b1, ray) : slab_test<false>(K, b0, b1, ray))) continue;
				cnt.hit(RT_CNT_SLAB_P_PASS);
				if (cull == RT_CULL_BACK_FACE) mesh_closest<RT_CULL_BACK_FACE>(tri, count, ray, best.t, best_tri, cnt);
				else if (cull == RT_CULL_FRONT_FACE) mesh_closest<RT_CULL_FRONT_FACE>(tri, count, ray, best.t, best_tri, cnt);
				else mesh_closest<RT_CULL_NONE>(tri, count, ray, best.t, best_tri, cnt);
			}
			if (best_tri >= 0)
			{
				const Tri T = load_tri(tri + 3 * best_tri);
				best.did = true;
				best.material = __float_as_int(info.y);
				best.normal = v3(T.a0.w, T.a1.w, T.a2.w);      // Utils.h:178: the stored face normal
				best.origin = ray.o + ray.d * best.t;          // Utils.h:162
			}
		}
		return best;
	}

	// Scene::DoesHit, Scene.cpp:68-96: any-hit in the same order.
	template <bool BVH, bool COUNT>
	__device__ __forceinline__ bool does_hit(const Staged sc, const SceneDevice& dev, const Ray& ray, Counters<COUNT>& cnt, unsigned long long* dev_counters)
	{
		const Pk K = make_pk(dev);
		float t;
#pragma unroll 1
		for (int i = 0; i < dev.n_spheres; ++i)
			if (hit_sphere<true>(sc.sphere(i), ray, t, cnt)) return true;
		if (planes_any(K, sc, dev.n_planes, ray, cnt)) return true;
#pragma unroll 1
		for (int m = 0; m < dev.n_meshes; ++m)
		{
			const float4 info = sc.mesh(3 * m + 2), ptrs = sc.mesh_ptr(m);
			const int cull = __float_as_int(info.x);
			const float4* tri = pointer_from_bits(ptrs.x, ptrs.y);
			// Utils.h:114-127: shadow rays see the opposite cull mode
			bool hit;
			if (BVH)
			{
				if (__float_as_int(info.w) == 0) continue;
				hit = bvh_any_any_cull(K, cull, pointer_from_bits(ptrs.z, ptrs.w), tri, ray, cnt);
			}
			else
			{
				const float4 b0 = sc.mesh(3 * m), b1 = sc.mesh(3 * m + 1);
				const int count = __float_as_int(b1.w);
				cnt.hit(RT_CNT_SLAB_S_TEST);
				if (!(ray.nan_safe ? slab_test<true>(K, b0, b1, ray) : slab_test<false>(K, b0, b1, ray))) continue;
				cnt.hit(RT_CNT_SLAB_S_PASS);
				if (cull == RT_CULL_BACK_FACE) hit = mesh_any<RT_CULL_FRONT_FACE>(tri, count, ray, cnt);
				else if (cull == RT_CULL_FRONT_FACE) hit = mesh_any<RT_CULL_BACK_FACE>(tri, count, ray, cnt);
				else hit = mesh_any<RT_CULL_NONE>(tri, count, ray, cnt);
			}
			if (hit) return true;
		}
		return false;
	}

	constexpr float kPi = 3.14159265358979323846f;   // MathHelpers.h:7

	// BRDF::GeometryFunction_SchlickGGX, BRDFs.h:78-86.
	__device__ __forceinline__ float schlick_ggx(V3 n, V3 v, float roughness)
	{
		const float a = mul(roughness, roughness);
		const float a1 = add(a, 1.f);
		const float k = quo(mul(a1, a1), 8.f);
		const float c = std_max(dot(n, v), 0.f);
		return quo(c, add(mul(c, sub(1.f, k)), k));
	}

	// Material::Shade as a tagged-union switch (Material.h:41-44, 60-63, 83-87, 107-123).
	// `view`: anything with load(2) = {-viewDir, -}: only the Phong and Cook-Torrance branches fetch it
	struct ViewInRegisters
	{
		V3 v;
		__device__ __forceinline__ float4 load(int) const { return make_float4(v.x, v.y, v.z, 0.f); }
	};
	template <bool COUNT, class View>
	__device__ __forceinline__ V3 shade(const float4 m0, const float4 m1, V3 n, V3 l, const View& view, Counters<COUNT>& cnt)
	{
		const int tag = __float_as_int(m0.x);
		const V3 color = v3(m0.y, m0.z, m0.w);
		if (tag == RT_MATERIAL_SOLID_COLOR)
		{
			cnt.hit(RT_CNT_SHADE_SOLID);
			return color;
		}
		if (tag == RT_MATERIAL_LAMBERT || tag == RT_MATERIAL_LAMBERT_PHONG)
		{
			// BRDF::Lambert(kd, cd) = (cd * kd) / PI, BRDFs.h:14-17: already evaluated by stage_scene
			V3 out = color;
			if (tag == RT_MATERIAL_LAMBERT) { cnt.hit(RT_CNT_SHADE_LAMBERT); return out; }
			cnt.hit(RT_CNT_SHADE_PHONG);
			const V3 v = v3(view.load(2));
			// BRDF::Phong, BRDFs.h:33-40
			const float nl = std_max(dot(n, l), 0.f);
			const V3 reflect = l - n * mul(2.f, nl);
			const float cosa = std_max(dot(reflect, v), 0.f);
			const float spec = mul(m1.y, power(cosa, m1.z));
			return v3(add(out.x, spec), add(out.y, spec), add(out.z, spec));
		}
		if (tag == RT_MATERIAL_COOK_TORRENCE)
		{
			cnt.hit(RT_CNT_SHADE_COOK_TORRENCE);
			const V3 v = v3(view.load(2));
			const float metal = m1.x, rough = m1.y;
			V3 h = v + l;
			normalize(h);
			const bool dielectric = (metal == 0.f);
			const V3 f0 = dielectric ? v3(0.04f, 0.04f, 0.04f) : color;
			// FresnelFunction_Schlick, BRDFs.h:49-53
			const float pw = power(sub(1.f, std_max(dot(h, v), 0.f)), 5.f);
			const V3 F = v3(add(f0.x, mul(sub(1.f, f0.x), pw)), add(f0.y, mul(sub(1.f, f0.y), pw)), add(f0.z, mul(sub(1.f, f0.z), pw)));
			// NormalDistribution_GGX, BRDFs.h:62-68
			const float a = mul(rough, rough);
			const float a2 = mul(a, a);
			const float nh = std_max(dot(n, h), 0.f);
			const float inner = add(mul(mul(nh, nh), sub(mul(a, a), 1.f)), 1.f);
			const float D = quo(a2, mul(kPi, mul(inner, inner)));
			// GeometryFunction_Smith, BRDFs.h:96-99
			const float G = mul(schlick_ggx(n, v, rough), schlick_ggx(n, l, rough));
			const float denom = mul(mul(4.f, std_max(dot(v, n), 0.0001f)), std_max(dot(l, n), 0.0001f));
			const V3 spec = v3(quo(mul(mul(F.x, D), G), denom), quo(mul(mul(F.y, D), G), denom), quo(mul(mul(F.z, D), G), denom));
			const V3 kd = dielectric ? v3(sub(1.f, F.x), sub(1.f, F.y), sub(1.f, F.z)) : v3(0.f, 0.f, 0.f);
			const V3 diff = v3(quo(mul(color.x, kd.x), kPi), quo(mul(color.y, kd.y), kPi), quo(mul(color.z, kd.z), kPi));
			return diff + spec;
		}
		return v3(0.f, 0.f, 0.f);
	}

	// LightUtils::GetRadiance, Utils.h:355-369 (note: measured from the un-offset hit origin).
	__device__ __forceinline__ V3 radiance(const float4 la, const float4 lb, V3 target)
	{
		const int type = __float_as_int(lb.w);
		const V3 color = v3(lb);
		if (type == RT_LIGHT_POINT)
		{
			const V3 d = v3(la) - target;
			return color * quo(la.w, dot(d, d));
		}
		if (type == RT_LIGHT_DIRECTIONAL) return color * la.w;
		return v3(0.f, 0.f, 0.f);
	}

	// Ray generation, Renderer.cpp:104-114
	__device__ __forceinline__ Ray view_ray(const FrameParams& p, int px, int py)
	{
		// Renderer.cpp:107-108 (the two expressions really do associate differently)
		const float cx = mul(mul(sub(mul(2.f, quo(add((float)px, 0.5f), (float)p.width)), 1.f), p.aspect), p.fov);
		const float cy = mul(sub(1.f, quo(mul(2.f, add((float)py, 0.5f)), (float)p.height)), p.fov);

		// Matrix::TransformVector(cx, cy, 1), Matrix.cpp:35-42; x * 1.f is exact
		Ray view;
		view.o = v3(p.cam_ox, p.cam_oy, p.cam_oz);
		view.d = v3(add(add(mul(p.right_x, cx), mul(p.up_x, cy)), p.fwd_x),
		            add(add(mul(p.right_y, cx), mul(p.up_y, cy)), p.fwd_y),
		            add(add(mul(p.right_z, cx), mul(p.up_z, cy)), p.fwd_z));
		normalize_and_invert(view.d, view.inv, view.nan_safe);        // Renderer.cpp:111-113, DataTypes.h:550-563
		view.tmin = 0.0001f; view.tmax = FLT_MAX;
		return view;
	}

	// The shadow ray of one light, Renderer.cpp:128-136; its direction is also the `l` of the shading
	// (GetDirectionToLight, Utils.h:341-353: light.origin - p for both light types)
	__device__ __forceinline__ Ray shadow_ray_to(const float4 la, int ltype, const V3 origin_offset)
	{
		Ray r;
		r.o = origin_offset;
		r.d = (ltype == RT_LIGHT_POINT || ltype == RT_LIGHT_DIRECTIONAL) ? (v3(la) - origin_offset) : v3(0.f, 0.f, 0.f);
		r.tmax = normalize_and_invert(r.d, r.inv, r.nan_safe);
		r.tmin = 0.0001f;
		return r;
	}

	// One unshadowed light's contribution by lighting mode, Renderer.cpp:143-171
	template <bool COUNT, class View>
	__device__ __forceinline__ V3 add_light(int mode, V3 color, const Staged sc, const float4 la, const float4 lb, const V3 l, const V3 hit_origin, const V3 hit_normal,
	                                        int material, const View& view, Counters<COUNT>& cnt)
	{
		if (mode == RT_LIGHTING_COMBINED)
		{
			const float oa = std_max(dot(hit_normal, l), 0.f);
			const V3 e = radiance(la, lb, hit_origin);
			const V3 brdf = shade(sc.material(2 * material), sc.material(2 * material + 1), hit_normal, l, view, cnt);
			// observedArea * radiance * brdf == (radiance * oa) * brdf, Renderer.cpp:152
			return color + v3(mul(mul(e.x, oa), brdf.x), mul(mul(e.y, oa), brdf.y), mul(mul(e.z, oa), brdf.z));
		}
		if (mode == RT_LIGHTING_OBSERVED_AREA)
		{
			const float oa = std_max(dot(hit_normal, l), 0.f);
			return color + v3(oa, oa, oa);
		}
		if (mode == RT_LIGHTING_RADIANCE) return color + radiance(la, lb, hit_origin);
		if (mode == RT_LIGHTING_BRDF) return color + shade(sc.material(2 * material), sc.material(2 * material + 1), hit_normal, l, view, cnt);
		return color;
	}

	// ColorRGB::MaxToOne (ColorRGB.h:12-17), static_cast<uint8_t>(c * 255) (truncation; the x86 reference goes through
	// cvttss2si and keeps the low byte) and SDL_MapRGB, Renderer.cpp:176-181
	__device__ __forceinline__ uint32_t pack_pixel(const FrameParams& p, V3 color)
	{
		const float max_value = std_max(color.x, std_max(color.y, color.z));
		if (max_value > 1.f) { color.x = quo(color.x, max_value); color.y = quo(color.y, max_value); color.z = quo(color.z, max_value); }
		const uint32_t R = (uint32_t)__float2int_rz(mul(color.x, 255.f)) & 0xffu;
		const uint32_t G = (uint32_t)__float2int_rz(mul(color.y, 255.f)) & 0xffu;
		const uint32_t B = (uint32_t)__float2int_rz(mul(color.z, 255.f)) & 0xffu;
		return (R << p.r_shift) | (G << p.g_shift) | (B << p.b_shift) | p.alpha_mask;
	}

	// What only the shading of a pixel needs - hit point, normal, material, view direction - waits in shared memory
	// while the shadow rays are traced, instead of occupying ten registers (or, as ptxas would have it, local memory
	// inside the light loop).  One slot of three float4 per thread, [record][thread] so that a warp's 128-bit
	// accesses are conflict-free.  Volatile: the loads stay where they are written, after the traversal.
	template <int THREADS>
	struct Parked
	{
		float4 slot[4][THREADS];           // record 3 = the persistent kernel's tile coordinates (two words)
	};
	struct ParkedRef
	{
		unsigned int at;         // shared address of this thread's first record
		unsigned int stride;     // bytes between records
		__device__ __forceinline__ void store_words(unsigned int a, unsigned int b) const { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(at + 3u * stride), "r"(a), "r"(b) : "memory"); }
		__device__ __forceinline__ uint2 load_words() const { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(at + 3u * stride) : "memory"); return v; }
		__device__ __forceinline__ void store(int k, float4 v) const
		{
			asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(at + (unsigned int)k * stride), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
		}
		__device__ __forceinline__ float4 load(int k) const
		{
			float4 v;
			asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(at + (unsigned int)k * stride) : "memory");
			return v;
		}
	};
	template <int THREADS>
	__device__ __forceinline__ ParkedRef parked_ref(const Parked<THREADS>& storage)
	{
		ParkedRef r;
		r.at = (unsigned int)__cvta_generic_to_shared(&storage.slot[0][threadIdx.x]);
		r.stride = (unsigned int)(sizeof(float4) * THREADS);
		return r;
	}

	template <int MODE, int SHADOWS, bool BVH, bool COUNT>
	__device__ __forceinline__ uint32_t render_pixel(const Staged sc, const ParkedRef park, const SceneDevice& dev, const FrameParams& p,
	                                                  int px, int py, Counters<COUNT>& cnt)
	{
		const int mode = (MODE >= 0) ? MODE : p.lighting_mode;
		const bool shadows = (SHADOWS >= 0) ? (SHADOWS != 0) : (p.shadows != 0);
		cnt.hit(RT_CNT_PIXELS);

		const Ray view = view_ray(p, px, py);

		float shadow_factor = 1.f;
		V3 color = v3(0.f, 0.f, 0.f);
		V3 origin_offset;
		bool did;
		{
			const Hit hit = closest_hit<BVH>(sc, dev, view, cnt, p.counters);
			did = hit.did;
			origin_offset = hit.origin + hit.normal * 0.0001f;          // Renderer.cpp:126
			park.store(0, make_float4(hit.origin.x, hit.origin.y, hit.origin.z, __int_as_float(hit.material)));
			park.store(1, make_float4(hit.normal.x, hit.normal.y, hit.normal.z, 0.f));
			park.store(2, make_float4(-view.d.x, -view.d.y, -view.d.z, 0.f));      // Renderer.cpp:150: Shade(hit, l, -viewDir)
		}
		if (did)
		{
			cnt.hit(RT_CNT_HIT_PIXELS);
#pragma unroll 1
			for (int li = 0; li < dev.n_lights; ++li)
			{
				cnt.hit(RT_CNT_LIGHT_ITERATIONS);
				const float4 la = sc.light_a(li);
				const int ltype = __float_as_int(sc.light_b(li).w);
				// GetDirectionToLight, Utils.h:341-353: light.origin - p for both light types
				const Ray shadow_ray = shadow_ray_to(la, ltype, origin_offset);

				if (shadows)
				{
					cnt.hit(RT_CNT_SHADOW_RAYS);
					if (does_hit<BVH>(sc, dev, shadow_ray, cnt, p.counters))
					{
						cnt.hit(RT_CNT_OCCLUDED);
						shadow_factor = mul(shadow_factor, 0.95f);     // Renderer.cpp:139-140
						continue;
					}
				}
				cnt.hit(RT_CNT_LIT);

				const V3 l = shadow_ray.d;
				const float4 h0 = park.load(0), h1 = park.load(1);
				const V3 hit_origin = v3(h0), hit_normal = v3(h1);
				const int material = __float_as_int(h0.w);
				const float4 lb = sc.light_b(li);
				color = add_light(mode, color, sc, la, lb, l, hit_origin, hit_normal, material, park, cnt);
			}
			color = color * shadow_factor;                                  // Renderer.cpp:173
		}

		return pack_pixel(p, color);
	}

	// `cam` is the origin every view ray of the frame starts from: the parts of HitTest_Sphere / HitTest_Plane
	// that depend on the ray origin only (Utils.h:54-55, 86) are evaluated here once per CTA for those rays,
	// with the same operations in the same order as the per-ray code.
	template <int THREADS>
	__device__ __forceinline__ void stage_scene(SharedScene& sc, const SceneDevice& dev, const V3 cam)
	{
		const int tid = threadIdx.x;
		for (int i = tid; i < dev.n_spheres; i += THREADS)
		{
			const float4 s = make_float4(dev.sphere_ox[i], dev.sphere_oy[i], dev.sphere_oz[i], dev.sphere_r[i]);
			const V3 ov = v3(s) - cam;
			sc.sphere[i] = s;
			sc.sphere_view[i] = make_float4(ov.x, ov.y, ov.z, dot(ov, ov));
			sc.sphere_mat[i] = dev.sphere_mat[i];
		}
		for (int i = tid; i < dev.n_planes; i += THREADS)
		{
			const float4 po = make_float4(dev.plane_ox[i], dev.plane_oy[i], dev.plane_oz[i], __int_as_float((int)dev.plane_mat[i]));
			const V3 n = v3(dev.plane_nx[i], dev.plane_ny[i], dev.plane_nz[i]);
			sc.plane_o[i] = po;
			sc.plane_n[i] = make_float4(n.x, n.y, n.z, dot(v3(po) - cam, n));
		}
		for (int j = tid; j < (dev.n_planes + 1) / 2; j += THREADS)
		{
			float o[2][3], n[2][3], num[2];
			for (int h = 0; h < 2; ++h)
			{
				const int i = min(2 * j + h, dev.n_planes - 1);
				o[h][0] = dev.plane_ox[i]; o[h][1] = dev.plane_oy[i]; o[h][2] = dev.plane_oz[i];
				n[h][0] = dev.plane_nx[i]; n[h][1] = dev.plane_ny[i]; n[h][2] = dev.plane_nz[i];
				num[h] = dot(v3(o[h][0], o[h][1], o[h][2]) - cam, v3(n[h][0], n[h][1], n[h][2]));
			}
			sc.plane_pair[3 * j + 0] = make_float4(o[0][0], o[1][0], o[0][1], o[1][1]);
			sc.plane_pair[3 * j + 1] = make_float4(o[0][2], o[1][2], n[0][0], n[1][0]);
			sc.plane_pair[3 * j + 2] = make_float4(n[0][1], n[1][1], n[0][2], n[1][2]);
			sc.plane_pair_view[j] = make_float2(num[0], num[1]);
		}
		for (int i = tid; i < dev.n_lights; i += THREADS)
		{
			sc.light_a[i] = make_float4(dev.light_ox[i], dev.light_oy[i], dev.light_oz[i], dev.light_intensity[i]);
			sc.light_b[i] = make_float4(dev.light_r[i], dev.light_g[i], dev.light_b[i], __int_as_float(dev.light_type[i]));
		}
		for (int i = tid; i < 3 * dev.n_meshes; i += THREADS) sc.mesh[i] = dev.mesh_table[i];
		for (int m = tid; m < dev.n_meshes; m += THREADS)
		{
			const float4 b1 = dev.mesh_table[3 * m + 1], info = dev.mesh_table[3 * m + 2];
			const unsigned long long tri = reinterpret_cast<unsigned long long>(dev.triangles + 3 * (size_t)__float_as_int(b1.z));
			const unsigned long long nodes = reinterpret_cast<unsigned long long>(dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z));
			sc.mesh_ptr[m] = make_float4(__int_as_float((int)(unsigned int)tri), __int_as_float((int)(unsigned int)(tri >> 32)),
			                             __int_as_float((int)(unsigned int)nodes), __int_as_float((int)(unsigned int)(nodes >> 32)));
		}
		for (int i = tid; i < dev.n_materials; i += THREADS)
		{
			float4 m0 = dev.materials[2 * i];
			const float4 m1 = dev.materials[2 * i + 1];
			const int tag = __float_as_int(m0.x);
			if (tag == RT_MATERIAL_LAMBERT || tag == RT_MATERIAL_LAMBERT_PHONG)
			{
				// BRDF::Lambert(kd, cd) = (cd * kd) / PI (BRDFs.h:14-17) depends on the material only:
				// evaluate it once per CTA with the same three operations per channel
				m0.y = quo(mul(m0.y, m1.x), 3.14159265358979323846f);
				m0.z = quo(mul(m0.z, m1.x), 3.14159265358979323846f);
				m0.w = quo(mul(m0.w, m1.x), 3.14159265358979323846f);
			}
			sc.material[2 * i] = m0;
			sc.material[2 * i + 1] = m1;
		}
	}

	// Progressive present: tell the copy stream (cuStreamWaitValue32 on band_done[b]) that pixels are in
	// memory.  Counters run in warp tiles (8x4 pixels): a 32x8 CTA tile is kSignalsPerTile of them.  The
	// caller has synchronised the threads whose stores it reports (CTA barrier / __syncwarp) and calls this
	// from ONE thread with the strip index `k` of the launch and the number of warp tiles finished.
	//  * One GPU (band_local == NULL): a release-ordered reduction on band_done[b].  A release
	//    (MEMBAR.ALL.GPU) is enough; __threadfence() would also invalidate the SM's L1 (CCTL.IVALL) and evict
	//    the BVH nodes and triangles the other resident warps are streaming.
	//  * Several GPUs: band_done lives on GPU 0 (peer memory) and a system-scope release per tile is far too
	//    expensive.  Tiles count on their own GPU (band_local[b], gpu scope); the one that completes this GPU's
	//    share of a band forwards the whole share with ONE system-scope release and re-arms the local counter.
	constexpr int kSignalsPerTile = kThreads / 32;
	__device__ __forceinline__ void signal_band_done(const FrameParams& p, int k, unsigned int units)
	{
		const int strip = k * p.strip_step + p.strip_first;      // position in the frame, whoever renders it
		const int band = p.band_table ? (int)__ldg(p.band_table + strip) : strip / p.strips_per_band;
		if (!p.band_local)
		{
			asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p.band_done + band), "r"(units) : "memory");
			return;
		}
		// strips of this band that belong to this GPU: those congruent to strip_first modulo strip_step
		const int total_strips = (p.row_end - p.row_begin + kBlockH - 1) / kBlockH;
		const int s0 = band * p.strips_per_band, s1 = min(total_strips, s0 + p.strips_per_band);
		const int first_mine = s0 + ((p.strip_first - s0) % p.strip_step + p.strip_step) % p.strip_step;
		const int mine = first_mine < s1 ? (s1 - 1 - first_mine) / p.strip_step + 1 : 0;
		const unsigned int share = (unsigned int)mine * (unsigned int)p.grid_x * (unsigned int)kSignalsPerTile;
		unsigned int old;
		asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p.band_local + band), "r"(units) : "memory");
		if (old + units == share)
		{
			p.band_local[band] = 0u;
			asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p.band_done + band), "r"(share) : "memory");
		}
	}

	// CTA-tile form: the whole 32x8 tile of this CTA is in memory.
	__device__ __forceinline__ void signal_band_done(const FrameParams& p)
	{
		__syncthreads();
		if (threadIdx.x == 0) signal_band_done(p, (int)blockIdx.y, (unsigned int)kSignalsPerTile);
	}

	template <int MODE, int SHADOWS, bool BVH, bool COUNT>
	__global__ void __launch_bounds__(kThreads)
	render_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p)
	{
		// dynamic shared memory: Parked<kThreads> | SharedScene cut after its last material (dynamic_smem_bytes)
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		Parked<kThreads>& parked = *reinterpret_cast<Parked<kThreads>*>(dynamic_smem);
		SharedScene& storage = *reinterpret_cast<SharedScene*>(dynamic_smem + sizeof(Parked<kThreads>));
		stage_scene<kThreads>(storage, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));
		__syncthreads();
		const Staged sc = staged_handle(storage);
		const ParkedRef park = parked_ref(parked);

		const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
		const int tx = lane & (kTileW - 1), ty = lane >> 3;
		const int wx = warp % kWarpsX, wy = warp / kWarpsX;
		const int px = blockIdx.x * kBlockW + wx * kTileW + tx;
		const int local_y = wy * kTileH + ty;
		const int py = p.row_begin + ((int)blockIdx.y * p.strip_step + p.strip_first) * kBlockH + local_y;
		const bool valid = (px < p.width) && (py < p.row_end);

		Counters<COUNT> cnt;
		uint32_t pixel = 0;
		if (valid) pixel = render_pixel<MODE, SHADOWS, BVH, COUNT>(sc, park, dev, p, px, py, cnt);
		if (COUNT) cnt.flush(p.counters);

		const int dst_row = p.dst_full_frame ? py : ((int)blockIdx.y * kBlockH + local_y);
		uint32_t* row = p.dst + (size_t)dst_row * (size_t)p.width;
		if (p.vector_store)
		{
			// four neighbouring pixels of a tile row -> one 128-bit store
			const uint32_t p1 = __shfl_down_sync(0xffffffffu, pixel, 1);
			const uint32_t p2 = __shfl_down_sync(0xffffffffu, pixel, 2);
			const uint32_t p3 = __shfl_down_sync(0xffffffffu, pixel, 3);
			if (valid && (tx & 3) == 0) *reinterpret_cast<uint4*>(row + px) = make_uint4(pixel, p1, p2, p3);
		}
		else if (valid)
		{
			row[px] = pixel;
		}

		if (p.band_done) signal_band_done(p);
	}

	// The shipped form of the pixel kernel: persistent warps.  The grid is one wave (SMs x resident CTAs);
	// the scene is staged into shared memory once per CTA, then every warp pulls 8x4 warp tiles off a
	// device-wide queue (ids run through the launch's tiles in the order the tiled kernel's CTAs would be
	// scheduled: neighbouring warps work on neighbouring pixels, bands finish in order) until the queue is
	// empty.  Against the tiled form (render_kernel, one CTA per 32x8 tile) this removes the per-tile
	// staging + barrier and the idle warp slots of a CTA waiting for its slowest warp.  The next id is
	// fetched before the current tile is rendered, so the atomic's latency is never exposed.  The queue
	// re-arms itself: the last warp to see it empty zeroes both words for the next launch.
	__device__ __forceinline__ int next_work_item(unsigned int* queue, int lane)
	{
		unsigned int id = 0;
		if (lane == 0) id = atomicAdd(queue, 1u);
		return (int)__shfl_sync(0xffffffffu, id, 0);
	}

	struct TileCoords { int k, px, py, local_y, cell; bool valid; };
	__device__ __forceinline__ TileCoords decode_work_item(const FrameParams& p, int item, int lane)
	{
		TileCoords c;
		const unsigned int tile = (unsigned int)item / kSignalsPerTile, warp = (unsigned int)item % kSignalsPerTile;
		int bx;
		if (p.cell_order)
		{
			const int shift = p.cell_w_log2 + p.cell_h_log2;
			const unsigned int rank = tile >> shift, within = tile & ((1u << shift) - 1u);
			const unsigned int cell = __ldg(p.cell_order + rank);
			const int cy = (int)__umulhi(cell, p.cells_x_magic), cx = (int)cell - cy * p.cells_x;
			c.k = (cy << p.cell_h_log2) + (int)(within >> p.cell_w_log2);
			bx = (cx << p.cell_w_log2) + (int)(within & ((1u << p.cell_w_log2) - 1u));
			if (bx >= p.grid_x || c.k >= p.n_strips) { bx = p.grid_x; c.k = p.n_strips; }      // padding: lands outside the frame below
		}
		else if (p.first_tiles == 0)
		{
			c.k = (int)__umulhi(tile, p.grid_x_magic);
			bx = (int)tile - c.k * p.grid_x;
		}
		else if ((int)tile < p.first_tiles)
		{
			// inside the rectangle, row by row
			const int w = p.first_x1 - p.first_x0;
			const int row = (int)__umulhi(tile, p.first_w_magic);
			c.k = p.first_k0 + row;
			bx = p.first_x0 + ((int)tile - row * w);
		}
		else
		{
			// everything else in the plain order: full rows above, the two side pieces of the rectangle's rows, full rows below
			const int w = p.first_x1 - p.first_x0;
			unsigned int j = tile - (unsigned int)p.first_tiles;
			const unsigned int above = (unsigned int)(p.first_k0 * p.grid_x);
			const unsigned int beside = (unsigned int)((p.first_k1 - p.first_k0) * (p.grid_x - w));
			if (j < above)
			{
				c.k = (int)__umulhi(j, p.grid_x_magic);
				bx = (int)j - c.k * p.grid_x;
			}
			else if (j - above < beside)
			{
				j -= above;
				const int row = (int)__umulhi(j, p.rest_w_magic);
				const int col = (int)j - row * (p.grid_x - w);
				c.k = p.first_k0 + row;
				bx = col < p.first_x0 ? col : col + w;
			}
			else
			{
				j -= above + beside;
				const int row = (int)__umulhi(j, p.grid_x_magic);
				c.k = p.first_k1 + row;
				bx = (int)j - row * p.grid_x;
			}
		}
		const int wx = warp % kWarpsX, wy = warp / kWarpsX;
		c.px = bx * kBlockW + wx * kTileW + (lane & (kTileW - 1));
		c.local_y = wy * kTileH + (lane >> 3);
		c.py = p.row_begin + (c.k * p.strip_step + p.strip_first) * kBlockH + c.local_y;
		c.valid = (c.px < p.width) && (c.py < p.row_end) && (bx < p.grid_x) && (c.k < p.n_strips);
		c.cell = (c.k >> p.cell_h_log2) * p.cells_x + (bx >> p.cell_w_log2);
		return c;
	}

	// ---- band watcher ------------------------------------------------------------------------------------------------
	// A copy stream that waits on every band's counter with cuStreamWaitValue32 reacts late when it reaches the wait
	// before the band is complete (the wait is polled by the front end: tens of microseconds), which is the normal
	// case whenever rendering a band and copying it take about equally long.  Instead ONE thread of the persistent grid
	// watches the counters (an acquire load every ~100 ns, from L2) and tells the host, which is blocked in rt_render
	// anyway, through a word of mapped pinned memory; the host issues that band's copy immediately.
	__device__ __forceinline__ unsigned int load_acquire_gpu(const unsigned int* p)
	{
		unsigned int v;
		asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
		return v;
	}

	__device__ __forceinline__ void watch_bands(const FrameParams& p)
	{
		if (threadIdx.x != 0) return;
		const int total_strips = (p.row_end - p.row_begin + kBlockH - 1) / kBlockH;
		for (int b = 0; b < p.watch_bands; ++b)
		{
			const int s0 = b * p.strips_per_band, s1 = min(total_strips, s0 + p.strips_per_band);
			if (s1 <= s0) break;
			// this launch's strips of the band: those congruent to strip_first modulo strip_step (as signal_band_done counts them)
			const int first_mine = s0 + ((p.strip_first - s0) % p.strip_step + p.strip_step) % p.strip_step;
			const int mine = first_mine < s1 ? (s1 - 1 - first_mine) / p.strip_step + 1 : 0;
			if (mine > 0)
			{
				const unsigned int expected = (unsigned int)mine * (unsigned int)p.grid_x * (unsigned int)kSignalsPerTile;
				while (load_acquire_gpu(p.band_done + b) < expected) __nanosleep(100);
			}
			// the band's pixels are in device memory (the acquire above pairs with the tiles' releases); the host only
			// needs to learn that: a system-scope release store into its mapped word
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p.host_flags + b), "r"(p.watch_tag) : "memory");
		}
	}

	// 128-thread CTAs, 8 per SM: 64 registers, 32 warps per SM.  (Round 1 ran 9 CTAs at 56 registers; with the round-2
	// loops the 56-register build spills into the per-light code and is 7 % slower on the 4K bunny frame: 0.761 vs 0.708 ms.)
#ifndef RT_PERSISTENT_MIN_CTAS
#define RT_PERSISTENT_MIN_CTAS 8
#endif
#ifndef RT_PERSISTENT_THREADS
#define RT_PERSISTENT_THREADS 128
#endif
	constexpr int kPersistentThreads = RT_PERSISTENT_THREADS;   // CTA size is free here: warps are the workers
	// dynamic shared memory of a launch of the tiled (kThreads) / persistent (kPersistentThreads) kernel
	inline size_t dynamic_smem_bytes(int threads, int n_materials) { return sizeof(float4) * 4 * (size_t)threads + staged_scene_bytes(n_materials); }
	template <int MODE, int SHADOWS, bool BVH>
	__global__ void __launch_bounds__(kPersistentThreads, RT_PERSISTENT_MIN_CTAS)
	render_kernel_persistent(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p)
	{
		if (p.host_flags && blockIdx.x == 0)       // this CTA is the frame's band watcher
		{
			watch_bands(p);
			return;
		}
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		Parked<kPersistentThreads>& parked = *reinterpret_cast<Parked<kPersistentThreads>*>(dynamic_smem);
		SharedScene& storage = *reinterpret_cast<SharedScene*>(dynamic_smem + sizeof(Parked<kPersistentThreads>));
		__shared__ unsigned int tile_clock[kPersistentThreads / 32];
		stage_scene<kPersistentThreads>(storage, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));
		__syncthreads();
		const Staged sc = staged_handle(storage);
		const ParkedRef park = parked_ref(parked);

		const int lane = threadIdx.x & 31;
		const int total = p.total_items ? p.total_items : p.grid_x * p.n_strips * kSignalsPerTile;
		Counters<false> cnt;

#ifdef RT_DRAIN_PROBE
		// experiment: p.counters = {first warp start, first warp that found the queue empty, last warp out} in globaltimer ns
		unsigned long long probe_t;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(probe_t));
		if (p.counters && lane == 0) atomicMin(p.counters, probe_t);
#endif
		int item = next_work_item(p.queue, lane);
		while (item < total)
		{
#ifdef RT_PERSIST_PREFETCH
			const int upcoming = next_work_item(p.queue, lane);
#endif
			uint32_t pixel = 0;
			if (p.cell_cost && lane == 0) tile_clock[threadIdx.x >> 5] = (unsigned int)clock();      // parked in shared memory: no register across the traversals
			// The tile's coordinates wait in shared memory while the pixel is rendered, as two packed words ({py, px} and
			// {valid, strip, row in strip}): cheaper than decoding the work item a second time, and no registers across the
			// traversals either.
			{
				const TileCoords c = decode_work_item(p, item, lane);
				park.store_words(((unsigned int)c.py << 16) | ((unsigned int)c.px & 0xffffu), (c.valid ? 0x80000000u : 0u) | ((unsigned int)c.k << 3) | (unsigned int)c.local_y);
				if (c.valid) pixel = render_pixel<MODE, SHADOWS, BVH, false>(sc, park, dev, p, c.px, c.py, cnt);
			}
			TileCoords c;
			{
				const uint2 words = park.load_words();
				const unsigned int where = words.x, strip_row = words.y;
				c.px = (int)(where & 0xffffu); c.py = (int)(where >> 16);
				c.k = (int)((strip_row & 0x7fffffffu) >> 3); c.local_y = (int)(strip_row & 7u); c.valid = (strip_row >> 31) != 0u;
			}
			const int dst_row = p.dst_full_frame ? c.py : (c.k * kBlockH + c.local_y);
			uint32_t* row = p.dst + (size_t)dst_row * (size_t)p.width;
			if (p.vector_store)
			{
				const uint32_t p1 = __shfl_down_sync(0xffffffffu, pixel, 1);
				const uint32_t p2 = __shfl_down_sync(0xffffffffu, pixel, 2);
				const uint32_t p3 = __shfl_down_sync(0xffffffffu, pixel, 3);
				if (c.valid && (lane & 3) == 0) *reinterpret_cast<uint4*>(row + c.px) = make_uint4(pixel, p1, p2, p3);
			}
			else if (c.valid)
			{
				row[c.px] = pixel;
			}
			if (p.band_done)
			{
				__syncwarp();
				if (lane == 0) signal_band_done(p, c.k, 1u);
			}
			if (p.cell_cost && lane == 0 && c.k < p.n_strips) atomicAdd(p.cell_cost + decode_work_item(p, item, lane).cell, ((unsigned int)clock() - tile_clock[threadIdx.x >> 5]) >> 4);
#ifdef RT_PERSIST_PREFETCH
			item = upcoming;
#else
			item = next_work_item(p.queue, lane);
#endif
		}

#ifdef RT_DRAIN_PROBE
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(probe_t));
		if (p.counters && lane == 0) { atomicMin(p.counters + 1, probe_t); atomicMax(p.counters + 2, probe_t); }
#endif
		// every warp of the grid arrives here exactly once, after its last fetch
		if (lane == 0)
		{
			const unsigned int warps = (gridDim.x - (p.host_flags ? 1u : 0u)) * (unsigned int)(kPersistentThreads / 32);
			if (atomicAdd(p.queue + 1, 1u) + 1u == warps)
			{
				p.queue[0] = 0u;
				p.queue[1] = 0u;
			}
		}
	}

}
