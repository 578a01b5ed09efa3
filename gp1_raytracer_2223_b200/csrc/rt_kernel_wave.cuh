// Rays, not pixels: the pixel path cut into five launches so that a frame's unit of work is ONE RAY IN ONE SUBTREE of a
// mesh's BVH instead of one pixel (view ray + every shadow ray, each a full walk, back to back on one thread).
//
// Why.  The reference's walk (IntersectionTest_BVH, source/Utils.h:246-288) neither orders children nor prunes by
// distance: a ray through a 3 082-triangle mesh visits hundreds of nodes, each step a dependent fetch.  With one
// thread per pixel a small frame has too few threads to hide that chain and the frame takes as long as its slowest
// pixel: Scene_W4_OptionalScene (source/Scene.cpp:439-474) at 320x240 ran 0.70 ms on 21 % of the warp slots
// (profiles/r02/r02_base_optional320_tiled_mix.csv), four serial walks per pixel.
//
// How.  Closest hit is a minimum over primitives, any hit a logical OR: both can be evaluated subtree by subtree in any
// order as long as every box and triangle test is the reference's and ties resolve the way its loop order resolves
// them.  So:
//   K1 primary      one thread per pixel: ray generation, spheres, planes -> hit_key[pixel]; then the TOP of every mesh's
//                   tree (the nodes above its <= 64 subtrees) is walked: one view job per (warp tile, subtree) whose root
//                   some ray of the tile reaches
//   K2 view walk    one warp per view job: the tile's 32 rays walk that ONE subtree (each after the boxes of the
//                   subtree's ancestors, which the recursion would have tested on its way down) and merge what they find
//                   with atomicMin on hit_key
//   K3 shadow setup one thread per pixel: hit record from the key, shadow-ray origin, spheres and planes against every
//                   light's shadow ray -> occluded[pixel]; shadow jobs (tile, light, subtree) like K1's
//   K4 shadow walk  one warp per shadow job: any-hit walk of one subtree -> atomicOr on occluded
//   K5 shade        one thread per pixel: Renderer.cpp:120-181 with the occlusion bits in place of Scene::DoesHit
// hit_key = t's bit pattern (t > 0, so unsigned order is numeric order) above a primitive number that grows in the
// reference's test order - spheres, planes, then the meshes' triangles in upload order (Scene.cpp:29-66; inside a mesh
// the walk meets leaves in ascending triangle order) - so the 64-bit minimum IS "smallest t, first tested wins ties"
// (strict '<' at Scene.cpp:37,48,58, Utils.h:275-278).  Every arithmetic expression is the same device function the
// one-kernel path calls; the frames are bit-identical (tests: every golden frame through RT_KERNEL_WAVEFRONT).
//
// Cost: ~28 bytes of scratch per pixel and five launches, which is why RT_KERNEL_AUTO only picks this form when the
// frame is small against the machine and the meshes are deep (rt_api.cu, launch()).
#pragma once

#include "rt_kernel.cuh"
#include "rt_wave_params.h"

namespace rt
{
namespace wave
{
	// Programmatic dependent launch (wave_launch): the next kernel of the chain may become resident once every CTA of this
	// one has said so; it must not touch anything this kernel's predecessors wrote before wait_for_previous_launch().
	// Without the launch attribute both are no-ops.
	__device__ __forceinline__ void let_next_launch_begin() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
	__device__ __forceinline__ void wait_for_previous_launch() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

	__device__ __forceinline__ unsigned long long make_key(float t, unsigned int primitive) { return ((unsigned long long)__float_as_uint(t) << 32) | primitive; }

	// pixel of thread `tid` of CTA (bx, k) of the launch: the tiled kernel's mapping (render_kernel)
	struct Where { int px, py, local_y; bool valid; };
	__device__ __forceinline__ Where where_am_i(const FrameParams& p, int bx, int k, int tid)
	{
		Where w;
		const int lane = tid & 31, warp = tid >> 5;
		const int wx = warp % kWarpsX, wy = warp / kWarpsX;
		w.px = bx * kBlockW + wx * kTileW + (lane & (kTileW - 1));
		w.local_y = wy * kTileH + (lane >> 3);
		w.py = p.row_begin + (k * p.strip_step + p.strip_first) * kBlockH + w.local_y;
		w.valid = (w.px < p.width) && (w.py < p.row_end);
		return w;
	}

	// The top of a mesh's tree: the walk of source/Utils.h:246-288 from the root, except that a subtree root is not
	// entered - its number goes into `alive` and the walk continues at its escape link (the job that owns the subtree
	// tests its box).  Returns the subtrees this ray reaches, one bit each.
	template <bool FAST>
	__device__ __forceinline__ unsigned long long walk_top(const float4* nodes, const uint8_t* root_of, const Ray& ray)
	{
		Pk K{};
		unsigned long long alive = 0ull;
		int at = 0;
		while (at >= 0)
		{
			const float4* rec = node_at(nodes, at);
			const float4 n1 = __ldg(rec + 1);
			const int s = (int)__ldg(root_of + (at >> 5));
			if (s) { alive |= 1ull << (s - 1); at = __float_as_int(n1.w); continue; }
			const bool inside = slab_test<FAST>(K, __ldg(rec), n1, ray);
			at = inside ? __float_as_int(n1.z) : __float_as_int(n1.w);         // top nodes are inner nodes: `hit` is the left child
		}
		return alive;
	}

	// One job per subtree that some ray of the warp reaches.  `extra` = the job word's bits above the subtree number.
	__device__ __forceinline__ void emit_jobs(unsigned long long alive, unsigned int tile, unsigned int extra, uint2* jobs, unsigned int* counter, unsigned int capacity, unsigned int* overflow)
	{
		// OR over the warp (two halves of 32 bits)
		const unsigned int lo = __reduce_or_sync(0xffffffffu, (unsigned int)alive), hi = __reduce_or_sync(0xffffffffu, (unsigned int)(alive >> 32));
		const unsigned int n = (unsigned int)(__popc(lo) + __popc(hi));
		if (n == 0) return;
		const unsigned int lane = threadIdx.x & 31;
		unsigned int first = 0;
		if (lane == 0) first = atomicAdd(counter, n);
		first = __shfl_sync(0xffffffffu, first, 0);
		if (first + n > capacity) { if (lane == 0) atomicExch(overflow, 1u); return; }
		// lane k writes the k-th and (k + 32)-th set bit
		const unsigned long long all = ((unsigned long long)hi << 32) | lo;
		for (unsigned int k = lane; k < n; k += 32)
		{
			unsigned long long rest = all;
			for (unsigned int j = 0; j < k; ++j) rest &= rest - 1;          // drop the k lowest set bits
			const unsigned int s = (unsigned int)(__ffsll((long long)rest) - 1);
			jobs[first + k] = make_uint2(tile, extra | s);
		}
	}

	// The walk of ONE subtree: from its root until the walk leaves it through the root's escape link `end`.  Only rays
	// that reached the root in walk_top (the recursion tested the boxes of the root's ancestors there) call this.
	template <bool ANY, bool FAST>
	__device__ __forceinline__ bool walk_subtree(int cull, const float4* nodes, const float4* tri, const int32_t* entry, const Ray& ray, float& best_t, int& best_tri)
	{
		Counters<false> cnt;
		Pk K{};
		int at = __ldg(entry);
		const int end = __ldg(entry + 1);
		while (at != end)
		{
			const float4* rec = node_at(nodes, at);
			const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1);
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
		}
		return false;
	}

	// ---- the same walk out of shared memory ----------------------------------------------------------------------------
	// A job's time is a chain of dependent fetches - node, node, leaf, triangle, ... - and a frame's walk kernels take as
	// long as their longest job (measured, tools/wave_jobs.py: the jobs of Scene_W4_OptionalScene at 320x240 add up to
	// 2 us of the machine, the longest one alone takes 57 us, nearly all of it waiting for L2).  So the warp first copies
	// the subtree - some tens of node records and triangles, contiguous in memory (rt_wave_params.h) - into its own piece of
	// shared memory with independent, coalesced loads, rewriting the links as offsets into the copy; the walk then waits
	// for shared memory instead.  Same boxes, same triangles, same order, same arithmetic.
	// (cp.async: global -> shared without a register in between, so the copy is in flight while the warp sets its rays up)
	__device__ __forceinline__ void copy16_async(float4* dst_shared, const float4* src)
	{
		asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned int)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
	}
	// layout of a warp's piece of shared memory: kFineAncestors node records (the boxes above the part), the part's root,
	// its descendants, its triangles
	constexpr int kStagedAncestorWords = 2 * kFineAncestors;
	__device__ __forceinline__ int stage_begin(float4* region, const float4* nodes, const float4* tri, const int32_t* entry, unsigned int lane)
	{
		const int root = __ldg(entry), below = __ldg(entry + 2), n_below = __ldg(entry + 3), tri_first = __ldg(entry + 4), n_tri = __ldg(entry + 5), n_above = __ldg(entry + 7);
		if ((int)lane < 2 * n_above) copy16_async(region + lane, node_at(nodes, __ldg(entry + 8 + (lane >> 1))) + (lane & 1));
		float4* staged_nodes = region + kStagedAncestorWords;
		const float4* root_rec = node_at(nodes, root);
		const float4* below_rec = node_at(nodes, below);
		const int node_words = 2 * (1 + n_below);
		for (int j = (int)lane; j < node_words; j += 32) copy16_async(staged_nodes + j, j < 2 ? root_rec + j : below_rec + (j - 2));
		float4* staged_tri = staged_nodes + node_words;
		const float4* src = tri + 3 * (size_t)tri_first;
		for (int j = (int)lane; j < 3 * n_tri; j += 32) copy16_async(staged_tri + j, src + j);
		asm volatile("cp.async.commit_group;" ::: "memory");
		return node_words;
	}
	// nobody needs the copy after all: it still has to land before the next one is started into the same memory
	__device__ __forceinline__ void stage_abandon()
	{
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();
	}
	// waits for the copy and rewrites the links as offsets into it: the root is record 0, its descendants follow; a
	// leaf's first triangle counts from the part's first triangle
	__device__ __forceinline__ void stage_finish(float4* region, int node_words, const int32_t* entry, unsigned int lane)
	{
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncwarp();                       // a thread only waits for its own copies
		const int end = __ldg(entry + 1), below = __ldg(entry + 2), tri_first = __ldg(entry + 4);
		float4* staged_nodes = region + kStagedAncestorWords;
		for (int j = 2 * (int)lane + 1; j < node_words; j += 64)             // every lane fixes what it copied or what a neighbour copied: hence the barrier below
		{
			int2* links = reinterpret_cast<int2*>(&staged_nodes[j].z);
			int2 l = *links;
			l.x = BvhLink::is_leaf(l.x) ? l.x - tri_first : l.x - below + BvhLink::kNodeBytes;
			l.y = l.y == end ? BvhLink::kEnd : l.y - below + BvhLink::kNodeBytes;
			*links = l;
		}
		__syncwarp();
	}

	// the boxes between the subtree's root and the part, out of the copy
	template <bool FAST>
	__device__ __forceinline__ bool enters_staged(const float4* region, int n_above, const Ray& ray)
	{
		Pk K{};
		bool in = true;
		for (int k = 0; k < n_above && in; ++k) in = slab_test<FAST>(K, region[2 * k], region[2 * k + 1], ray);
		return in;
	}

	__device__ __forceinline__ Tri staged_tri_at(const float4* tri, int i)
	{
		Tri t;
		t.a0 = tri[3 * i]; t.a1 = tri[3 * i + 1]; t.a2 = tri[3 * i + 2];
		return t;
	}

	// walk_subtree over the copy; best_tri counts from the subtree's first triangle
	template <bool ANY, bool FAST>
	__device__ __forceinline__ bool walk_staged(int cull, const float4* region, int node_words, const Ray& ray, float& best_t, int& best_tri)
	{
		Counters<false> cnt;
		Pk K{};
		const float4* staged_nodes = region + kStagedAncestorWords;
		const float4* tri = staged_nodes + node_words;
		int at = 0;
		do
		{
			const float4* rec = node_at(staged_nodes, at);
			const float4 n0 = rec[0], n1 = rec[1];
			const bool inside = slab_test<FAST>(K, n0, n1, ray);
			const int hit = __float_as_int(n1.z), miss = __float_as_int(n1.w);
			at = inside ? hit : miss;
			if (inside && BvhLink::is_leaf(hit))
			{
				const int first = BvhLink::leaf_first(hit), count = BvhLink::leaf_count(hit);
				if (ANY)
				{
					if (cull == RT_CULL_BACK_FACE) { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_BACK_FACE>(staged_tri_at(tri, first + k), ray, cnt)) return true; }
					else if (cull == RT_CULL_FRONT_FACE) { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_FRONT_FACE>(staged_tri_at(tri, first + k), ray, cnt)) return true; }
					else { for (int k = 0; k < count; ++k) if (shadow_one<RT_CULL_NONE>(staged_tri_at(tri, first + k), ray, cnt)) return true; }
				}
				else
				{
					if (cull == RT_CULL_BACK_FACE) { for (int k = 0; k < count; ++k) closest_one<RT_CULL_BACK_FACE>(staged_tri_at(tri, first + k), first + k, ray, best_t, best_tri, cnt); }
					else if (cull == RT_CULL_FRONT_FACE) { for (int k = 0; k < count; ++k) closest_one<RT_CULL_FRONT_FACE>(staged_tri_at(tri, first + k), first + k, ray, best_t, best_tri, cnt); }
					else { for (int k = 0; k < count; ++k) closest_one<RT_CULL_NONE>(staged_tri_at(tri, first + k), first + k, ray, best_t, best_tri, cnt); }
				}
				at = miss;
			}
		} while (at >= 0);
		return false;
	}

	// The boxes between a subtree's root and one of its parts: all must be met for the recursion to arrive at the part
	template <bool FAST>
	__device__ __forceinline__ bool enters_part(const float4* nodes, const int32_t* entry, const Ray& ray)
	{
		Pk K{};
		const int n = __ldg(entry + 7);
		bool in = true;
		for (int k = 0; k < n && in; ++k)
		{
			const float4* rec = node_at(nodes, __ldg(entry + 8 + k));
			in = slab_test<FAST>(K, __ldg(rec), __ldg(rec + 1), ray);
		}
		return in;
	}

	// What a view ray hits before any mesh is looked at: spheres, then planes (Scene.cpp:29-52), as a key
	__device__ __forceinline__ unsigned long long primary_key(const Staged sc, const SceneDevice& dev, const Pk& K, const Ray& ray)
	{
		Counters<false> cnt;
		float best_t = FLT_MAX;
		int best_sphere = -1, best_plane = -1;
#pragma unroll 1
		for (int i = 0; i < dev.n_spheres; ++i)
		{
			float t;
			const float4 sv = sc.sphere_view(i);
			if (hit_sphere_from<false>(v3(sv), sv.w, sc.sphere(i).w, ray, t, cnt) && t < best_t) { best_t = t; best_sphere = i; }
		}
		planes_closest(K, sc, dev.n_planes, ray, best_t, best_plane, cnt);
		if (best_plane >= 0) return make_key(best_t, kPlaneBase + (unsigned int)best_plane);
		if (best_sphere >= 0) return make_key(best_t, (unsigned int)best_sphere);
		return kNoHit;
	}

	// ---- K1 ----------------------------------------------------------------------------------------------------------
	__global__ void __launch_bounds__(kThreads)
	primary_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p, const __grid_constant__ WaveParams w)
	{
		let_next_launch_begin();
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		SharedScene& storage = *reinterpret_cast<SharedScene*>(dynamic_smem);
		stage_scene<kThreads>(storage, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));
		__syncthreads();
		const Staged sc = staged_handle(storage);

		const Where me = where_am_i(p, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
		const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
		const unsigned int pixel = cta * kThreads + threadIdx.x;
		const Pk K = make_pk(dev);

		Ray ray{};
		unsigned long long key = kNoHit;
		if (me.valid)
		{
			ray = view_ray(p, me.px, me.py);
			key = primary_key(sc, dev, K, ray);
		}
		w.hit_key[pixel] = key;
		w.occluded[pixel] = 0u;            // K3 and K4 only ever set bits

		// view jobs: (warp tile, mesh, subtree) for every subtree some ray of the tile reaches
		const unsigned int tile = cta * kSignalsPerTile + (threadIdx.x >> 5);
#pragma unroll 1
		for (int m = 0; m < dev.n_meshes; ++m)
		{
			const float4 info = sc.mesh(3 * m + 2);
			if (__float_as_int(info.w) == 0) continue;
			const int32_t* split = w.split + (size_t)m * kSplitStride;
			const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
			unsigned long long alive = 0ull;
			if (me.valid) alive = ray.nan_safe ? walk_top<true>(nodes, w.root_map + __ldg(split + 1), ray) : walk_top<false>(nodes, w.root_map + __ldg(split + 1), ray);
			w.view_alive[(size_t)pixel * dev.n_meshes + m] = alive;
			emit_jobs(alive, tile, (unsigned int)m << 8, w.view_jobs, w.counters, w.view_capacity, w.counters + 4);
		}
	}

	// One unit of the view walk: the rays of warp tile `tile` in subtree `s` of mesh `m` - the whole subtree (part ==
	// kFine) or one of its parts.  `region`: the warp's piece of shared memory (PARTS only).
	template <bool PARTS>
	__device__ __forceinline__ void view_unit(const SceneDevice& dev, const FrameParams& p, const WaveParams& w, float4* region, unsigned int lane,
	                                          unsigned int tile, unsigned int m, unsigned int s, unsigned int part)
	{
		const int32_t* entry = w.split + (size_t)m * kSplitStride + kSplitHeader + (s * (kFine + 1) + part) * kSplitWords;
		const int flags = __ldg(entry + 6);
		if (flags & kPartPresent)
		{
			const unsigned int cta = tile / kSignalsPerTile;
			const int tid = (int)((tile % kSignalsPerTile) * 32u + lane);
			const Where me = where_am_i(p, (int)(cta % (unsigned int)p.grid_x), (int)(cta / (unsigned int)p.grid_x), tid);
			const unsigned int pixel = cta * kThreads + (unsigned int)tid;
			const float4 b1 = __ldg(dev.mesh_table + 3 * m + 1), info = __ldg(dev.mesh_table + 3 * m + 2);
			const int first_tri = __float_as_int(b1.z);
			const float4* tri = dev.triangles + 3 * (size_t)first_tri;
			const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
			const int cull = __float_as_int(info.x);
			// the copy starts before anything is known about the rays: it is in flight while their masks arrive
			const unsigned long long subtrees_reached = w.view_alive[(size_t)pixel * dev.n_meshes + m];
			const bool staged = PARTS && (flags & kPartStageable) != 0;      // whole subtrees are walked where they are: few of a tile's rays see much of one
			int node_words = 0;
			if (staged) node_words = stage_begin(region, nodes, tri, entry, lane);
			// rays that never reach the subtree, or do not get from its root to this part, sit the unit out
			bool reaches = me.valid && ((subtrees_reached >> s) & 1ull);
			if (__ballot_sync(0xffffffffu, reaches) == 0u) { if (staged) stage_abandon(); }
			else
			{
				Ray ray{};
				if (reaches) ray = view_ray(p, me.px, me.py);
				if (staged)
				{
					stage_finish(region, node_words, entry, lane);
					if (reaches) reaches = ray.nan_safe ? enters_staged<true>(region, __ldg(entry + 7), ray) : enters_staged<false>(region, __ldg(entry + 7), ray);
				}
				else if (reaches) reaches = ray.nan_safe ? enters_part<true>(nodes, entry, ray) : enters_part<false>(nodes, entry, ray);
				if (reaches)
				{
					const float best_t = __uint_as_float((unsigned int)(w.hit_key[pixel] >> 32));       // a stale read only makes the bound looser
					// (the bound only decides which candidates are worth an atomic: the triangle tests never see it)
					float t = FLT_MAX;
					int best_tri = -1;
					if (staged)
					{
						if (ray.nan_safe) walk_staged<false, true>(cull, region, node_words, ray, t, best_tri);
						else walk_staged<false, false>(cull, region, node_words, ray, t, best_tri);
						best_tri += __ldg(entry + 4);
					}
					else if (ray.nan_safe) walk_subtree<false, true>(cull, nodes, tri, entry, ray, t, best_tri);
					else walk_subtree<false, false>(cull, nodes, tri, entry, ray, t, best_tri);
					if (t < FLT_MAX && t <= best_t)
						atomicMin(w.hit_key + pixel, make_key(t, kTriangleBase + (unsigned int)(first_tri + best_tri)));
				}
				__syncwarp();                // the next unit overwrites the copy
			}
		}
	}

	// ---- K2 ----------------------------------------------------------------------------------------------------------
	template <bool PARTS>
	__global__ void __launch_bounds__(kWalkWarps * 32, 8)
	view_walk_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p, const __grid_constant__ WaveParams w)
	{
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		float4* region = reinterpret_cast<float4*>(dynamic_smem + (PARTS ? (threadIdx.x >> 5) * kRegionBytes : 0u));       // (no shared memory without PARTS)
		// jobs are handed out by a counter: their costs differ by orders of magnitude
		const unsigned int lane = threadIdx.x & 31;
		let_next_launch_begin();
		wait_for_previous_launch();
		const unsigned int n_jobs = min(w.counters[0], w.view_capacity);
		// whole subtrees or their parts (rt_wave_params.h; the host chooses, wave_launch)
		constexpr unsigned int shift = PARTS ? kFineShift : 0;
		const unsigned int n_units = n_jobs << shift;
		// units are handed out by a counter, one at a time: their costs differ by orders of magnitude, and a warp that
		// reserved units ahead would sit on them while it works through a long one
		for (;;)
		{
			unsigned int unit = 0;
			if (lane == 0) unit = atomicAdd(w.counters + 2, 1u);
			unit = __shfl_sync(0xffffffffu, unit, 0);
			if (unit >= n_units) break;
			const long long unit_began = w.job_cycles ? clock64() : 0ll;
			const uint2 word = w.view_jobs[unit >> shift];
			const unsigned int tile = word.x, m = (word.y >> 8) & 0xffu, s = word.y & 0xffu;
			view_unit<PARTS>(dev, p, w, region, lane, tile, m, s, shift ? unit % kFine : (unsigned int)kFine);
			if (w.job_cycles) { __syncwarp(); if (lane == 0) w.job_cycles[unit] = (unsigned int)(clock64() - unit_began); }
		}
	}

	// The hit record of a key, exactly as Scene::GetClosestHit leaves it (Scene.cpp:35-63, Utils.h:67-68, 91-92, 162, 178)
	__device__ __forceinline__ Hit hit_of_key(const Staged sc, const SceneDevice& dev, const Ray& ray, unsigned long long key)
	{
		Hit h;
		h.did = key != kNoHit;
		h.t = __uint_as_float((unsigned int)(key >> 32));
		h.material = 0; h.origin = v3(0.f, 0.f, 0.f); h.normal = v3(0.f, 0.f, 0.f);
		if (!h.did) return h;
		const unsigned int primitive = (unsigned int)key;
		h.origin = ray.o + ray.d * h.t;
		if (primitive < kPlaneBase)
		{
			h.material = sc.sphere_mat((int)primitive);
			h.normal = h.origin - v3(sc.sphere((int)primitive));
			normalize(h.normal);
		}
		else if (primitive < kTriangleBase)
		{
			const int i = (int)(primitive - kPlaneBase);
			h.material = __float_as_int(sc.plane_o(i).w);
			h.normal = v3(sc.plane_n(i));
		}
		else
		{
			const int global_tri = (int)(primitive - kTriangleBase);
			const Tri T = load_tri(dev.triangles + 3 * (size_t)global_tri);
			h.normal = v3(T.a0.w, T.a1.w, T.a2.w);
			// the mesh that owns the triangle (meshes are stored back to back in upload order)
			int m = 0;
			for (int k = 0; k < dev.n_meshes; ++k)
			{
				const float4 b1 = sc.mesh(3 * k + 1);
				const int first = __float_as_int(b1.z), count = __float_as_int(b1.w);
				if (global_tri >= first && global_tri < first + count) m = k;
			}
			h.material = __float_as_int(sc.mesh(3 * m + 2).y);
		}
		return h;
	}

	// Scene::DoesHit's spheres and planes (Scene.cpp:71-89) for one shadow ray
	__device__ __forceinline__ bool blocked_before_meshes(const Staged sc, const SceneDevice& dev, const Pk& K, const Ray& ray)
	{
		Counters<false> cnt;
		float t;
#pragma unroll 1
		for (int i = 0; i < dev.n_spheres; ++i)
			if (hit_sphere<true>(sc.sphere(i), ray, t, cnt)) return true;
		return planes_any(K, sc, dev.n_planes, ray, cnt);
	}

	// ---- K3 ----------------------------------------------------------------------------------------------------------
	// gridDim.z = 1 or the number of lights: in a small frame a pixel's shadow rays are set up by one thread each (the top
	// of the tree is a chain of dependent steps, and the frame has too few pixels to hide three of them back to back)
	__global__ void __launch_bounds__(kThreads)
	shadow_setup_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p, const __grid_constant__ WaveParams w)
	{
		let_next_launch_begin();
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		SharedScene& storage = *reinterpret_cast<SharedScene*>(dynamic_smem);
		stage_scene<kThreads>(storage, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));         // (the scene tables are not a kernel's output)
		__syncthreads();
		wait_for_previous_launch();
		const Staged sc = staged_handle(storage);

		const Where me = where_am_i(p, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
		const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
		const unsigned int pixel = cta * kThreads + threadIdx.x;
		const unsigned int tile = cta * kSignalsPerTile + (threadIdx.x >> 5);
		const Pk K = make_pk(dev);

		bool did = false;
		V3 origin_offset = v3(0.f, 0.f, 0.f);
		if (me.valid)
		{
			const Ray view = view_ray(p, me.px, me.py);
			const Hit hit = hit_of_key(sc, dev, view, w.hit_key[pixel]);
			did = hit.did;
			origin_offset = hit.origin + hit.normal * 0.0001f;          // Renderer.cpp:126
		}
		if (blockIdx.z == 0) w.shadow_origin[pixel] = make_float4(origin_offset.x, origin_offset.y, origin_offset.z, did ? 1.f : 0.f);

#pragma unroll 1
		for (int li = (int)blockIdx.z; li < dev.n_lights; li += (int)gridDim.z)
		{
			bool open = did;
			Ray ray{};
			if (open)
			{
				ray = shadow_ray_to(sc.light_a(li), __float_as_int(sc.light_b(li).w), origin_offset);
				if (blocked_before_meshes(sc, dev, K, ray)) { open = false; atomicOr(w.occluded + pixel, 1u << li); }
			}
#pragma unroll 1
			for (int m = 0; m < dev.n_meshes; ++m)
			{
				const float4 info = sc.mesh(3 * m + 2);
				if (__float_as_int(info.w) == 0) continue;
				const int32_t* split = w.split + (size_t)m * kSplitStride;
				const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
				unsigned long long alive = 0ull;
				if (open) alive = ray.nan_safe ? walk_top<true>(nodes, w.root_map + __ldg(split + 1), ray) : walk_top<false>(nodes, w.root_map + __ldg(split + 1), ray);
				w.shadow_alive[((size_t)pixel * dev.n_lights + li) * dev.n_meshes + m] = alive;
				emit_jobs(alive, tile, ((unsigned int)li << 16) | ((unsigned int)m << 8), w.shadow_jobs, w.counters + 1, w.shadow_capacity, w.counters + 4);
			}
		}
	}

	// One unit of the shadow walk: the shadow rays towards light `li` of warp tile `tile` in subtree `s` of mesh `m`
	template <bool PARTS>
	__device__ __forceinline__ void shadow_unit(const SceneDevice& dev, const FrameParams& p, const WaveParams& w, float4* region, unsigned int lane,
	                                            unsigned int tile, unsigned int li, unsigned int m, unsigned int s, unsigned int part)
	{
		const int32_t* entry = w.split + (size_t)m * kSplitStride + kSplitHeader + (s * (kFine + 1) + part) * kSplitWords;
		const int flags = __ldg(entry + 6);
		if (flags & kPartPresent)
		{
			const unsigned int cta = tile / kSignalsPerTile;
			const unsigned int pixel = cta * kThreads + (tile % kSignalsPerTile) * 32u + lane;
			const unsigned int bit = 1u << li;
			const float4 b1 = __ldg(dev.mesh_table + 3 * m + 1), info = __ldg(dev.mesh_table + 3 * m + 2);
			const float4* tri = dev.triangles + 3 * (size_t)__float_as_int(b1.z);
			const float4* nodes = dev.bvh_nodes + 2 * (size_t)__float_as_int(info.z);
			const int cull = __float_as_int(info.x);
			// Utils.h:114-127: shadow rays see the opposite cull mode
			const int shadow_cull = cull == RT_CULL_BACK_FACE ? RT_CULL_FRONT_FACE : (cull == RT_CULL_FRONT_FACE ? RT_CULL_BACK_FACE : RT_CULL_NONE);
			// rays that never reach the subtree, do not get from its root to this part, or are already known to be in shadow
			// (racy read: an optimisation only) sit the unit out
			// the copy starts before anything is known about the rays: it is in flight while their masks arrive
			const float4 so = w.shadow_origin[pixel];
			const unsigned long long subtrees_reached = w.shadow_alive[((size_t)pixel * dev.n_lights + li) * dev.n_meshes + m];
			const unsigned int known_occluded = w.occluded[pixel];
			const bool staged = PARTS && (flags & kPartStageable) != 0;      // whole subtrees are walked where they are: few of a tile's rays see much of one
			int node_words = 0;
			if (staged) node_words = stage_begin(region, nodes, tri, entry, lane);
			bool reaches = ((subtrees_reached >> s) & 1ull) && !(known_occluded & bit);
			if (__ballot_sync(0xffffffffu, reaches) == 0u) { if (staged) stage_abandon(); }
			else
			{
				Ray ray{};
				if (reaches)
				{
					const float4 la = make_float4(__ldg(dev.light_ox + li), __ldg(dev.light_oy + li), __ldg(dev.light_oz + li), 0.f);
					ray = shadow_ray_to(la, __ldg(dev.light_type + li), v3(so));
				}
				if (staged)
				{
					stage_finish(region, node_words, entry, lane);
					if (reaches) reaches = ray.nan_safe ? enters_staged<true>(region, __ldg(entry + 7), ray) : enters_staged<false>(region, __ldg(entry + 7), ray);
				}
				else if (reaches) reaches = ray.nan_safe ? enters_part<true>(nodes, entry, ray) : enters_part<false>(nodes, entry, ray);
				if (reaches)
				{
					float t = FLT_MAX; int tri_id = -1;
					bool blocked;
					if (staged) blocked = ray.nan_safe ? walk_staged<true, true>(shadow_cull, region, node_words, ray, t, tri_id) : walk_staged<true, false>(shadow_cull, region, node_words, ray, t, tri_id);
					else blocked = ray.nan_safe ? walk_subtree<true, true>(shadow_cull, nodes, tri, entry, ray, t, tri_id) : walk_subtree<true, false>(shadow_cull, nodes, tri, entry, ray, t, tri_id);
					if (blocked) atomicOr(w.occluded + pixel, bit);
				}
				__syncwarp();                // the next unit overwrites the copy
			}
		}
	}

	// ---- K4 ----------------------------------------------------------------------------------------------------------
	template <bool PARTS>
	__global__ void __launch_bounds__(kWalkWarps * 32, 8)
	shadow_walk_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p, const __grid_constant__ WaveParams w)
	{
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		float4* region = reinterpret_cast<float4*>(dynamic_smem + (PARTS ? (threadIdx.x >> 5) * kRegionBytes : 0u));       // (no shared memory without PARTS)
		const unsigned int lane = threadIdx.x & 31;
		let_next_launch_begin();
		wait_for_previous_launch();
		const unsigned int n_jobs = min(w.counters[1], w.shadow_capacity);
		// whole subtrees or their parts (rt_wave_params.h; the host chooses, wave_launch)
		constexpr unsigned int shift = PARTS ? kFineShift : 0;
		const unsigned int n_units = n_jobs << shift;
		// units are handed out by a counter, one at a time: their costs differ by orders of magnitude, and a warp that
		// reserved units ahead would sit on them while it works through a long one
		for (;;)
		{
			unsigned int unit = 0;
			if (lane == 0) unit = atomicAdd(w.counters + 3, 1u);
			unit = __shfl_sync(0xffffffffu, unit, 0);
			if (unit >= n_units) break;
			const long long unit_began = w.job_cycles ? clock64() : 0ll;
			const uint2 word = w.shadow_jobs[unit >> shift];
			const unsigned int tile = word.x, li = (word.y >> 16) & 0xffu, m = (word.y >> 8) & 0xffu, s = word.y & 0xffu;
			shadow_unit<PARTS>(dev, p, w, region, lane, tile, li, m, s, shift ? unit % kFine : (unsigned int)kFine);
			if (w.job_cycles) { __syncwarp(); if (lane == 0) w.job_cycles[(size_t)w.view_capacity * kFine + unit] = (unsigned int)(clock64() - unit_began); }
		}
	}

	// Renderer.cpp:120-181 for one pixel, given what it hit and which lights are blocked
	__device__ __forceinline__ uint32_t shade_pixel(int mode, const Staged sc, const SceneDevice& dev, const FrameParams& p, const Ray& view, unsigned long long key, unsigned int occluded)
	{
		Counters<false> cnt;
		const Hit hit = hit_of_key(sc, dev, view, key);
		float shadow_factor = 1.f;
		V3 color = v3(0.f, 0.f, 0.f);
		if (hit.did)
		{
			const V3 origin_offset = hit.origin + hit.normal * 0.0001f;     // Renderer.cpp:126
			const ViewInRegisters view_neg{ neg(view.d) };                  // Renderer.cpp:150
#pragma unroll 1
			for (int li = 0; li < dev.n_lights; ++li)
			{
				if ((occluded >> li) & 1u) { shadow_factor = mul(shadow_factor, 0.95f); continue; }     // Renderer.cpp:137-141
				const float4 la = sc.light_a(li), lb = sc.light_b(li);
				const Ray to_light = shadow_ray_to(la, __float_as_int(lb.w), origin_offset);
				color = add_light(mode, color, sc, la, lb, to_light.d, hit.origin, hit.normal, hit.material, view_neg, cnt);
			}
			color = color * shadow_factor;                                  // Renderer.cpp:173
		}
		return pack_pixel(p, color);
	}
	// the pixel of every lane of a warp tile into the frame; `k` = the tile's strip of the launch (all lanes call)
	__device__ __forceinline__ void store_pixel(const FrameParams& p, const Where& me, int k, uint32_t pixel)
	{
		const int dst_row = p.dst_full_frame ? me.py : (k * kBlockH + me.local_y);
		uint32_t* row = p.dst + (size_t)dst_row * (size_t)p.width;
		if (p.vector_store)
		{
			const uint32_t p1 = __shfl_down_sync(0xffffffffu, pixel, 1);
			const uint32_t p2 = __shfl_down_sync(0xffffffffu, pixel, 2);
			const uint32_t p3 = __shfl_down_sync(0xffffffffu, pixel, 3);
			if (me.valid && (threadIdx.x & 3) == 0) *reinterpret_cast<uint4*>(row + me.px) = make_uint4(pixel, p1, p2, p3);
		}
		else if (me.valid)
		{
			row[me.px] = pixel;
		}
	}

	// ---- K5 ----------------------------------------------------------------------------------------------------------
	template <int MODE>
	__global__ void __launch_bounds__(kThreads)
	shade_kernel(const __grid_constant__ SceneDevice dev, const __grid_constant__ FrameParams p, const __grid_constant__ WaveParams w)
	{
		extern __shared__ __align__(16) unsigned char dynamic_smem[];
		SharedScene& storage = *reinterpret_cast<SharedScene*>(dynamic_smem);
		stage_scene<kThreads>(storage, dev, v3(p.cam_ox, p.cam_oy, p.cam_oz));
		__syncthreads();
		wait_for_previous_launch();
		const Staged sc = staged_handle(storage);

		const Where me = where_am_i(p, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x);
		const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
		const unsigned int pixel_index = cta * kThreads + threadIdx.x;
		uint32_t pixel = 0;
		if (me.valid) pixel = shade_pixel(MODE, sc, dev, p, view_ray(p, me.px, me.py), w.hit_key[pixel_index], w.occluded[pixel_index]);
		store_pixel(p, me, (int)blockIdx.y, pixel);
		if (p.band_done) signal_band_done(p);
		// what the host sizes the next frame's walk kernels by (rt_api.cu, launch())
		if (w.jobs_report && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { w.jobs_report[0] = w.counters[0]; w.jobs_report[1] = w.counters[1]; }
	}
}
}
