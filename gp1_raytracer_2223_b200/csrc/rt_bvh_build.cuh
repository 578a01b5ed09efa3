// TriangleMesh::UpdateTransforms as the reference ships it - vertex / normal transform AND BuildBVH - on the
// device (reference source/DataTypes.h:210-236, 294-483 with BVH and USE_BINS defined,
// DataTypes.h:8-9).  One CTA of 32 warps per mesh.
//
// What has to be reproduced, and how:
//  * the tree: per node the binned SAH split of FindBestSplitPlane (DataTypes.h:398-483: centroid bounds, 8 bins,
//    7 planes, 3 axes, first strictly smaller cost wins), the no-split test of Subdivide (DataTypes.h:333-336)
//    and the leaf rule idxCount <= 8 (DataTypes.h:327) - every float operation in the reference's order
//    (rt_device.cuh arithmetic), minima / maxima in any order (exact and commutative; the one observable
//    difference is which of -0.0f / +0.0f a bound keeps when both occur);
//  * the triangle order: Subdivide partitions indices / normals IN PLACE with a two-pointer sweep
//    (DataTypes.h:344-363) and the next UpdateTransforms starts from the order the last one left, so the mesh
//    state lives on the device between calls (indices / normals ping-pong buffers).  The sweep's result has a
//    closed form (partition_range below), so it runs as scans and scatters instead of a serial loop;
//  * the walk order of IntersectionTest_BVH (Utils.h:246-288: left child, then left + 1): children are a pair,
//    every node gets the "escape" link of rt::BvhLink when it is created (left -> right sibling, right ->
//    parent's escape).  Node NUMBERS are free (the walk never compares them): pairs are handed out by an atomic
//    counter instead of the reference's depth-first nodesUsed++, which is what lets independent subtrees be
//    built concurrently.
//
// Work distribution, two launches per UpdateTransforms call:
//  * update_transforms_bvh_kernel, one CTA of 32 warps: transform, per-triangle data, root, then the tree level by
//    level while a level has at most 8 nodes - the CTA splits into 1 / 2 / 4 / 8 teams of warps (named barriers),
//    one node per team (one warp per node once the nodes are small).  The first level with more than 8 nodes (at
//    most 16) is handed over;
//  * build_subtrees_kernel, 16 CTAs of 16 warps: every CTA takes handed-over nodes off a ticket counter and builds
//    their subtrees to the bottom the same way (teams, then one warp per node), position-indexed scratch in its
//    shared memory.  The last CTA to finish writes what the build leaves behind: indices / normals for the next
//    call, the triangle stream and the node records of the pixel kernel, into the results and into the scene block.
// A build is instruction-issue bound per SM (about 1.4 k warp instructions per inner node), which is why the wide
// part of the tree goes to several SMs instead of more warps of one.
#pragma once

#include "rt_kernel.cuh"

namespace rt
{
	struct BuildParams
	{
		float m[16];                    // Matrix::data[0..3] of finalTransform, row by row
		const float* positions;         // 3 per vertex, untransformed
		int32_t vertex_count;
		int32_t triangle_count;
		const int32_t* indices_in;      // order before this call
		const float* normals_in;        // untransformed face normals, same order
		int32_t* indices_out;           // order after this call (the other half of the ping-pong pair)
		float* normals_out;
		// scratch, all sized by the host (T = triangles, N = 2T - 1 nodes at most)
		float* tpos;                    // 3V   transformedPositions
		float* centroid;                // 3T   per triangle slot
		unsigned int* tri_min;          // 3T   float_key of the min / max over the slot's three vertices
		unsigned int* tri_max;          // 3T
		float* tnormal;                 // 3T   transformedNormals per slot
		int32_t* order;                 // T    order[k] = slot that sits at position k
		int32_t* order_tmp;             // T    partition output before it is copied back
		int32_t* rights_before;         // T    partition: right-hand elements in front of position k, inside its node
		int32_t* front_right;           // T    partition: position of the node's k-th right-hand element from the front
		int32_t* back_left;             // T    partition: distance from the node's end of its k-th left-hand element from the back
		int32_t* node_first;            // N    first triangle position
		int32_t* node_count;            // N    triangles
		int32_t* node_escape;           // N
		float* node_box;                // 6N   min xyz, max xyz
		int32_t* queue_a;               // T    nodes of the level being split
		int32_t* queue_b;               // T    nodes created for the next level
		// results (persistent per mesh: copied into the scene's mesh block by emit_mesh_kernel)
		float4* result_triangles;       // 3T   {v0|nx}{e1|ny}{e2|nz} in the new order
		float4* result_nodes;           // 2N   device node records (rt::BvhLink)
		int32_t* result_info;           // kInfoWords: [0] nodes used, [1] status (0 ok, 1 leaf too large for BvhLink), [2..7] root box bits, hand-over
		// the mesh's slices of the scene's mesh block: written at the end of the build as well (NULL: not)
		float4* scene_triangles;
		float4* scene_nodes;
		float4* scene_table;
		int32_t work_in_shared;         // top kernel: the per-triangle work arrays (14 words per triangle) live in dynamic shared memory
		int32_t subtree_shared_triangles;   // subtree kernel: subtrees up to this size keep their position-indexed scratch in shared memory
	};

	// result_info words beyond the results proper: the hand-over between the two kernels
	enum : int { kInfoNodesUsed = 0, kInfoStatus = 1, kInfoRootBox = 2, kInfoSubtrees = 8, kInfoTicket = 9, kInfoFinished = 10,
	             kInfoSubtreeList = 16, kInfoWords = 32 };

	// The arrays every pass of the build reads and writes.  Shared memory when the mesh fits (latency of a pass is a few
	// dependent accesses: ~30 cycles each there, an L2 round trip each in global memory), else the host's scratch.
	struct BuildWork
	{
		float* centroid;                // 3T   per triangle slot
		unsigned int* tri_min;          // 3T   float_key of the triangle's box
		unsigned int* tri_max;          // 3T
		int32_t* order;                 // T    order[k] = slot that sits at position k
		int32_t* order_tmp;             // T
		int32_t* rights_before;         // T
		int32_t* front_right;           // T
		int32_t* back_left;             // T
	};
	constexpr int kBuildWorkWordsPerTriangle = 14;

	constexpr int kBuildThreads = 1024;         // top kernel
	constexpr int kBuildWarps = kBuildThreads / 32;
	constexpr int kSubtreeThreads = 512;        // subtree kernel
	constexpr int kSubtreeCtas = 16;            // a handed-over level has at most 2 * kMaxTeams nodes
	constexpr int kMaxTeams = 8;                // teams of more than one warp: named barriers 1..8
	constexpr int kWarpNodeTriangles = 96;      // levels whose nodes are all this small go to one warp per node
	constexpr int kAtomicBinTriangles = 64;     // nodes up to this size bin with plain shared-memory atomics
	constexpr int kSubtreeWordsPerTriangle = 5; // order, order_tmp, rights_before, front_right, back_left

	// order-preserving map float -> unsigned for integer min / max (REDUX, shared-memory atomics); no NaN reaches it
	__device__ __forceinline__ unsigned int float_key(float f)
	{
		const unsigned int b = __float_as_uint(f);
		return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
	}
	__device__ __forceinline__ float key_float(unsigned int k)
	{
		return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
	}

	// AABB::Area, DataTypes.h:75-79
	__device__ __forceinline__ float box_area(const float mn[3], const float mx[3])
	{
		const float ex = sub(mx[0], mn[0]), ey = sub(mx[1], mn[1]), ez = sub(mx[2], mn[2]);
		return add(add(mul(ex, ey), mul(ey, ez)), mul(ez, ex));
	}

	// Per team (and per warp once teams are single warps): reduction slots, bins of the three axes, the split that
	// was chosen.
	struct TeamScratch
	{
		unsigned int centroid_lo[3], centroid_hi[3];
		unsigned int bin_count[3][8];
		unsigned int bin_lo[3][8][3];           // float_key of the bin box
		unsigned int bin_hi[3][8][3];
		unsigned int child_lo[2][3], child_hi[2][3];
		int warp_total[kBuildWarps];
		float best, split_pos;
		int axis, pair;
	};

	struct Team
	{
		int threads;       // 32 * warps
		int warps;
		int tid;           // thread within the team
		int warp;          // warp within the team
		int barrier;       // named barrier of the team (teams of more than one warp)
		TeamScratch* s;
	};

	__device__ __forceinline__ void team_sync(const Team& t)
	{
		if (t.threads == 32) __syncwarp();
		else asm volatile("bar.sync %0, %1;" :: "r"(t.barrier), "r"(t.threads) : "memory");
	}

	// The device record of a finished node (SceneDevice::bvh_nodes): box, first child / first triangle, link.
	__device__ __forceinline__ void write_node_record(const BuildParams& p, int node, bool leaf, int first_or_child, int leaf_triangles)
	{
		const float* b = p.node_box + 6 * node;
		if (leaf && leaf_triangles > BvhLink::kMaxLeafTriangles) atomicExch(p.result_info + 1, 1);
		const int hit = leaf ? BvhLink::leaf(first_or_child, leaf_triangles) : BvhLink::inner(first_or_child);
		p.result_nodes[2 * node] = make_float4(b[0], b[3], b[1], b[4]);
		p.result_nodes[2 * node + 1] = make_float4(b[2], b[5], __int_as_float(hit), __int_as_float(BvhLink::miss(p.node_escape[node])));
	}

	// The in-place partition of Subdivide (DataTypes.h:344-363) over positions [first, first + count):
	//     i = first, j = last;  while (i <= j)  left(i) ? ++i : (swap(i, j), --j)
	// Each element is examined exactly once, always at position i: elements come from the front while the last
	// one examined was left-hand and from the back while it was right-hand.  Hence, with k a position relative
	// to `first`, R(k) the right-hand elements in front of k and L(k) the left-hand elements behind k:
	//   * an element is reached from the front iff  k + B(k) < count,  B(k) = 0 if R(k) == 0, else 1 + the
	//     distance from the end of the R(k)-th left-hand element counted from the back (count if there is none);
	//     then a left-hand element stays where it is and a right-hand one ends at  count - 1 - B(k);
	//   * otherwise it is reached from the back: a left-hand element fills the hole of the (L(k) + 1)-th
	//     right-hand element counted from the front, a right-hand one moves down by one position.
	// (Checked against the literal loop for every flag pattern up to 14 elements and random ones up to 4000.)
	// Returns the left count; `order` holds the result.  Also reorders when everything is right-hand, like the loop.
	__device__ __forceinline__ int partition_range(const BuildWork& p, const Team& t, int first, int count, int axis, float split_pos)
	{
		TeamScratch& s = *t.s;
		const int lane = threadIdx.x & 31;
		int running = 0;
		for (int base = 0; base < count; base += t.threads)
		{
			const int k = base + t.tid;
			const bool active = k < count;
			const bool right = active && !(p.centroid[3 * p.order[first + k] + axis] < split_pos);
			const unsigned int ballot = __ballot_sync(0xffffffffu, right);
			int before = running + __popc(ballot & ((1u << lane) - 1u));
			if (t.threads > 32)
			{
				if (lane == 0) s.warp_total[t.warp] = __popc(ballot);
				team_sync(t);
				for (int w = 0; w < t.warps; ++w)
				{
					const int c = s.warp_total[w];
					if (w < t.warp) before += c;
					running += c;
				}
				team_sync(t);
			}
			else running += __popc(ballot);
			if (active)
			{
				p.rights_before[first + k] = before;
				if (right) p.front_right[first + before] = k;
			}
		}
		const int n_left = count - running;
		if (n_left == count) return n_left;                      // nothing moves
		for (int k = t.tid; k < count; k += t.threads)
		{
			const bool left = p.centroid[3 * p.order[first + k] + axis] < split_pos;
			if (left) p.back_left[first + n_left - (k - p.rights_before[first + k]) - 1] = count - 1 - k;
		}
		team_sync(t);
		for (int k = t.tid; k < count; k += t.threads)
		{
			const int slot = p.order[first + k];
			const bool left = p.centroid[3 * slot + axis] < split_pos;
			const int r = p.rights_before[first + k];
			const int from_back = (r == 0) ? 0 : (r > n_left ? count : p.back_left[first + r - 1] + 1);
			int final_position;
			if (k + from_back < count) final_position = left ? k : count - 1 - from_back;
			else
			{
				const int lefts_behind = n_left - (k - r) - (left ? 1 : 0);
				final_position = left ? p.front_right[first + lefts_behind] : k - 1;
			}
			p.order_tmp[first + final_position] = slot;
		}
		team_sync(t);
		for (int k = t.tid; k < count; k += t.threads) p.order[first + k] = p.order_tmp[first + k];
		team_sync(t);
		return n_left;
	}

	// Subdivide (DataTypes.h:323-389) of one node by one team.  Children that need splitting go to the next level's queue.
	__device__ __forceinline__ void subdivide_node(const BuildParams& p, const BuildWork& w, int node, const Team& t, int* nodes_used, int* next, int* next_count, int* next_big)
	{
		TeamScratch& s = *t.s;
		const int lane = threadIdx.x & 31;
		const int first = p.node_first[node], count = p.node_count[node];
		const unsigned int idx_count = 3u * (unsigned int)count;
		const unsigned int key_max = float_key(FLT_MAX), key_min = float_key(FLT_MIN);

		// every reduction starts like the reference's boxes: +FLT_MAX / +FLT_MIN (Vector3.cpp:13-14)
		if (t.tid < 3) { s.centroid_lo[t.tid] = key_max; s.centroid_hi[t.tid] = key_min; }
		if (t.tid < 6) { (&s.child_lo[0][0])[t.tid] = key_max; (&s.child_hi[0][0])[t.tid] = key_min; }
		for (int k = t.tid; k < 24; k += t.threads) (&s.bin_count[0][0])[k] = 0u;
		for (int k = t.tid; k < 72; k += t.threads) { (&s.bin_lo[0][0][0])[k] = key_max; (&s.bin_hi[0][0][0])[k] = key_min; }
		team_sync(t);

		// ---- FindBestSplitPlane, DataTypes.h:398-483; centroid bounds of the three axes in one pass ----
		{
			unsigned int lo[3] = { key_max, key_max, key_max }, hi[3] = { key_min, key_min, key_min };
			for (int k = t.tid; k < count; k += t.threads)
			{
				const int slot = w.order[first + k];
				for (int a = 0; a < 3; ++a)
				{
					const unsigned int c = float_key(w.centroid[3 * slot + a]);
					lo[a] = min(lo[a], c);
					hi[a] = max(hi[a], c);
				}
			}
			for (int a = 0; a < 3; ++a)
			{
				lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
				hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
			}
			if (lane == 0)
				for (int a = 0; a < 3; ++a) { atomicMin(&s.centroid_lo[a], lo[a]); atomicMax(&s.centroid_hi[a], hi[a]); }
		}
		team_sync(t);
		float lo[3], diff[3], scale[3];
		bool use_axis[3];
		for (int a = 0; a < 3; ++a)
		{
			lo[a] = key_float(s.centroid_lo[a]);
			diff[a] = sub(key_float(s.centroid_hi[a]), lo[a]);
			use_axis[a] = !(fabsf(diff[a]) < FLT_EPSILON);                             // DataTypes.h:421-422
			scale[a] = quo(8.f, diff[a]);
		}

		// bins of the three axes, DataTypes.h:425-441: lanes that share a bin reduce among themselves (REDUX), one of
		// them adds the group to the team's bins
		for (int base = 0; base < count; base += t.threads)
		{
			const int k = base + t.tid;
			const bool active = k < count;
			int slot = 0;
			float c[3] = { 0.f, 0.f, 0.f };
			unsigned int mn[3] = { key_max, key_max, key_max }, mx[3] = { key_min, key_min, key_min };
			if (active)
			{
				slot = w.order[first + k];
				for (int a = 0; a < 3; ++a)
				{
					c[a] = w.centroid[3 * slot + a];
					mn[a] = w.tri_min[3 * slot + a];
					mx[a] = w.tri_max[3 * slot + a];
				}
			}
			for (int a = 0; a < 3; ++a)
			{
				if (!use_axis[a]) continue;                                              // uniform over the team
				const int bin = active ? min(7, __float2int_rz(mul(sub(c[a], lo[a]), scale[a]))) : -1;
				if (count <= kAtomicBinTriangles)                                        // uniform over the team
				{
					if (active)
					{
						atomicAdd(&s.bin_count[a][bin], 3u);
						for (int d = 0; d < 3; ++d) { atomicMin(&s.bin_lo[a][bin][d], mn[d]); atomicMax(&s.bin_hi[a][bin][d], mx[d]); }
					}
					continue;
				}
				const unsigned int peers = __match_any_sync(0xffffffffu, bin);
				if (active)
				{
					unsigned int gmn[3], gmx[3];
					for (int d = 0; d < 3; ++d) { gmn[d] = __reduce_min_sync(peers, mn[d]); gmx[d] = __reduce_max_sync(peers, mx[d]); }
					if (lane == __ffs(peers) - 1)
					{
						atomicAdd(&s.bin_count[a][bin], 3u * (unsigned int)__popc(peers));
						for (int d = 0; d < 3; ++d) { atomicMin(&s.bin_lo[a][bin][d], gmn[d]); atomicMax(&s.bin_hi[a][bin][d], gmx[d]); }
					}
				}
			}
		}
		team_sync(t);

		// the 7 planes of the 3 axes, DataTypes.h:443-479.  Lane 8a + b of the team's first warp loads bin b of axis a;
		// an inclusive scan up the 8 lanes gives "bins 0..b" (leftBox / leftCount of plane b), one down the lanes gives
		// "bins b..7", whose neighbour b + 1 is rightBox / rightCount of plane b.  Boxes stay float_keys until the areas.
		// The reference walks axis 0..2, plane 0..6 and keeps the first strictly smaller cost: the lowest lane holding
		// the minimum.
		if (t.warp == 0)
		{
			const int a = lane >> 3, b = lane & 7;
			const bool holds_bin = lane < 24;
			unsigned int llo[3] = { key_max, key_max, key_max }, lhi[3] = { key_min, key_min, key_min };
			int lcount = 0;
			if (holds_bin)
			{
				lcount = (int)s.bin_count[a][b];
				for (int d = 0; d < 3; ++d) { llo[d] = s.bin_lo[a][b][d]; lhi[d] = s.bin_hi[a][b][d]; }
			}
			unsigned int rlo[3] = { llo[0], llo[1], llo[2] }, rhi[3] = { lhi[0], lhi[1], lhi[2] };
			int rcount = lcount;
			for (int off = 1; off < 8; off <<= 1)
			{
				const int uc = __shfl_up_sync(0xffffffffu, lcount, off, 8), dc = __shfl_down_sync(0xffffffffu, rcount, off, 8);
				unsigned int ulo[3], uhi[3], dlo[3], dhi[3];
				for (int d = 0; d < 3; ++d)
				{
					ulo[d] = __shfl_up_sync(0xffffffffu, llo[d], off, 8); uhi[d] = __shfl_up_sync(0xffffffffu, lhi[d], off, 8);
					dlo[d] = __shfl_down_sync(0xffffffffu, rlo[d], off, 8); dhi[d] = __shfl_down_sync(0xffffffffu, rhi[d], off, 8);
				}
				if (b >= off) { lcount += uc; for (int d = 0; d < 3; ++d) { llo[d] = min(llo[d], ulo[d]); lhi[d] = max(lhi[d], uhi[d]); } }
				if (b + off < 8) { rcount += dc; for (int d = 0; d < 3; ++d) { rlo[d] = min(rlo[d], dlo[d]); rhi[d] = max(rhi[d], dhi[d]); } }
			}
			// plane b: right side = bins b + 1..7 = the lane above
			rcount = __shfl_down_sync(0xffffffffu, rcount, 1, 8);
			for (int d = 0; d < 3; ++d) { rlo[d] = __shfl_down_sync(0xffffffffu, rlo[d], 1, 8); rhi[d] = __shfl_down_sync(0xffffffffu, rhi[d], 1, 8); }
			float cost = 0.f;
			bool candidate = false;
			if (holds_bin && b < 7 && (a == 0 ? use_axis[0] : (a == 1 ? use_axis[1] : use_axis[2])))
			{
				float lmn[3], lmx[3], rmn[3], rmx[3];
				for (int d = 0; d < 3; ++d) { lmn[d] = key_float(llo[d]); lmx[d] = key_float(lhi[d]); rmn[d] = key_float(rlo[d]); rmx[d] = key_float(rhi[d]); }
				// DataTypes.h:472: leftCount[i] * leftArea[i] + rightCount[i] * rightArea[i]
				cost = add(mul((float)lcount, box_area(lmn, lmx)), mul((float)rcount, box_area(rmn, rmx)));
				candidate = cost < FLT_MAX;                                              // bestCost starts at FLT_MAX; false for NaN
			}
			float m = candidate ? cost : INFINITY;
			for (int sh = 16; sh > 0; sh >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, sh));
			const unsigned int winners = __ballot_sync(0xffffffffu, candidate && cost == m);
			if (lane == 0)
			{
				if (winners)
				{
					const int win = __ffs(winners) - 1, wa = win >> 3, wp = win & 7;
					const float wlo = wa == 0 ? lo[0] : (wa == 1 ? lo[1] : lo[2]), wdiff = wa == 0 ? diff[0] : (wa == 1 ? diff[1] : diff[2]);
					s.axis = wa;
					s.split_pos = add(wlo, mul(quo(wdiff, 8.f), (float)(wp + 1)));            // DataTypes.h:469, 476
					s.best = m;
				}
				else { s.axis = 0; s.split_pos = 0.f; s.best = FLT_MAX; }
			}
		}
		team_sync(t);
		const float best = s.best, split_pos = s.split_pos;
		const int axis = s.axis;

		// ---- Subdivide, DataTypes.h:333-336: keep the node as a leaf if splitting is not cheaper ----
		const float no_split = mul((float)idx_count, box_area(p.node_box + 6 * node, p.node_box + 6 * node + 3));
		bool leaf = best >= no_split;
		int left_count = 0;
		if (!leaf)
		{
			left_count = partition_range(w, t, first, count, axis, split_pos);          // DataTypes.h:344-363
			leaf = (left_count == 0 || left_count == count);                           // DataTypes.h:366-369
		}
		if (leaf)
		{
			if (t.tid == 0) write_node_record(p, node, true, first, count);
			team_sync(t);                                                                // scratch is reused by the next node
			return;
		}

		// ---- children, DataTypes.h:371-388: both boxes in one pass (UpdateNodeBounds, DataTypes.h:310-321) ----
		if (t.tid == 0) s.pair = atomicAdd(nodes_used, 2);
		for (int base = 0; base < count; base += t.threads)
		{
			const int k = base + t.tid;
			const bool active = k < count;
			const int side = (k < left_count) ? 0 : 1;
			const unsigned int in_left = __ballot_sync(0xffffffffu, active && side == 0), in_right = __ballot_sync(0xffffffffu, active && side == 1);
			if (active)
			{
				const int slot = w.order[first + k];
				const unsigned int peers = side ? in_right : in_left;
				unsigned int gmn[3], gmx[3];
				for (int d = 0; d < 3; ++d)
				{
					gmn[d] = __reduce_min_sync(peers, w.tri_min[3 * slot + d]);
					gmx[d] = __reduce_max_sync(peers, w.tri_max[3 * slot + d]);
				}
				if (lane == __ffs(peers) - 1)
					for (int d = 0; d < 3; ++d) { atomicMin(&s.child_lo[side][d], gmn[d]); atomicMax(&s.child_hi[side][d], gmx[d]); }
			}
		}
		team_sync(t);
		if (t.tid == 0)
		{
			const int pair = s.pair;
			const int child[2] = { pair, pair + 1 };
			const int child_first[2] = { first, first + left_count };
			const int child_count[2] = { left_count, count - left_count };
			for (int c = 0; c < 2; ++c)
			{
				p.node_first[child[c]] = child_first[c];
				p.node_count[child[c]] = child_count[c];
				p.node_escape[child[c]] = (c == 0) ? child[1] : p.node_escape[node];
				for (int d = 0; d < 3; ++d) { p.node_box[6 * child[c] + d] = key_float(s.child_lo[c][d]); p.node_box[6 * child[c] + 3 + d] = key_float(s.child_hi[c][d]); }
				if (3 * child_count[c] <= 8) write_node_record(p, child[c], true, child_first[c], child_count[c]);   // DataTypes.h:327
				else { next[atomicAdd(next_count, 1)] = child[c]; atomicMax(next_big, child_count[c]); }
			}
			write_node_record(p, node, false, child[0], 0);
		}
		team_sync(t);
	}

	// Level state of one CTA: node counts and the largest node of the current and the next level.
	struct LevelState
	{
		int count[2];
		int big[2];
	};

	// Builds level after level from the nodes in `cur` (level.count[0] of them, the largest level.big[0] triangles).
	// HAND_OVER: stops in front of the first level with more than kMaxTeams nodes and returns its size (the nodes are
	// in *handed); 0 when the tree was finished.
	template <bool HAND_OVER>
	__device__ __forceinline__ int run_levels(const BuildParams& p, const BuildWork& w, TeamScratch* scratch, LevelState& level, int* nodes_used,
	                                          int* cur, int* nxt, int** handed)
	{
		const int tid = threadIdx.x, warp = tid >> 5, n_warps = blockDim.x >> 5;
		int parity = 0;
		while (true)
		{
			const int n = level.count[parity], big = level.big[parity];
			if (n == 0) return 0;
			if (HAND_OVER && n > kMaxTeams) { *handed = cur; return n; }
			Team t;
			int n_teams;
			if (n > kMaxTeams || big <= kWarpNodeTriangles) { n_teams = n_warps; t.warps = 1; }
			else
			{
				// a team is as many warps as the level's largest node has triangles per lane, no more: warps that would
				// only run the passes with every lane idle cost issue slots the others need (the rest of the CTA sits out)
				n_teams = n <= 1 ? 1 : (n <= 2 ? 2 : (n <= 4 ? 4 : 8));
				t.warps = n_warps / n_teams;
				while (t.warps > 1 && 32 * (t.warps / 2) >= big) t.warps /= 2;
			}
			t.threads = 32 * t.warps;
			const int team = warp / t.warps;
			t.warp = warp - team * t.warps;
			t.tid = tid - team * t.threads;
			t.barrier = 1 + team;
			t.s = &scratch[team < n_teams ? team : 0];
			if (team < n_teams)
				for (int q = team; q < n; q += n_teams)
					subdivide_node(p, w, cur[q], t, nodes_used, nxt, &level.count[parity ^ 1], &level.big[parity ^ 1]);
			__syncthreads();
			if (tid == 0) { level.count[parity] = 0; level.big[parity] = 0; }
			int* swap = cur; cur = nxt; nxt = swap;
			parity ^= 1;
			__syncthreads();
		}
	}

	__global__ void __launch_bounds__(kBuildThreads)
	update_transforms_bvh_kernel(const __grid_constant__ BuildParams p)
	{
		extern __shared__ __align__(16) float dynamic_shared[];
		__shared__ TeamScratch scratch[kBuildWarps];
		__shared__ LevelState level;
		__shared__ int nodes_used;
		const int tid = threadIdx.x, lane = tid & 31;
		const int T = p.triangle_count;
		BuildWork w;
		if (p.work_in_shared)
		{
			w.centroid = dynamic_shared; w.tri_min = reinterpret_cast<unsigned int*>(w.centroid + 3 * T); w.tri_max = w.tri_min + 3 * T;
			w.order = reinterpret_cast<int32_t*>(w.tri_max + 3 * T); w.order_tmp = w.order + T; w.rights_before = w.order_tmp + T;
			w.front_right = w.rights_before + T; w.back_left = w.front_right + T;
		}
		else
		{
			w.centroid = p.centroid; w.tri_min = p.tri_min; w.tri_max = p.tri_max; w.order = p.order; w.order_tmp = p.order_tmp;
			w.rights_before = p.rights_before; w.front_right = p.front_right; w.back_left = p.back_left;
		}

		// transformedPositions, DataTypes.h:216-222 (Matrix::TransformPoint, Matrix.cpp:49-56)
		for (int v = tid; v < p.vertex_count; v += kBuildThreads)
		{
			const V3 q = transform_point(p.m, p.positions[3 * v], p.positions[3 * v + 1], p.positions[3 * v + 2]);
			p.tpos[3 * v] = q.x; p.tpos[3 * v + 1] = q.y; p.tpos[3 * v + 2] = q.z;
		}
		if (tid == 0)
		{
			nodes_used = 1; level.count[0] = 0; level.count[1] = 0; level.big[0] = 0; level.big[1] = 0;
			p.result_info[kInfoStatus] = 0; p.result_info[kInfoSubtrees] = 0; p.result_info[kInfoTicket] = 0; p.result_info[kInfoFinished] = 0;
		}
		__syncthreads();

		// per triangle slot: transformedNormals (DataTypes.h:224-230), centroid (DataTypes.h:349), vertex min / max.
		// The slot-indexed arrays also go to global memory: the subtree kernel reads them there.
		for (int t = tid; t < T; t += kBuildThreads)
		{
			V3 v[3];
			for (int k = 0; k < 3; ++k)
			{
				const int vi = p.indices_in[3 * t + k];
				v[k] = v3(p.tpos[3 * vi], p.tpos[3 * vi + 1], p.tpos[3 * vi + 2]);
			}
			const V3 c = ((v[0] + v[1]) + v[2]) * 0.3333f;
			const float cc[3] = { c.x, c.y, c.z };
			const unsigned int mn[3] = { float_key(std_min(std_min(v[0].x, v[1].x), v[2].x)), float_key(std_min(std_min(v[0].y, v[1].y), v[2].y)),
			                             float_key(std_min(std_min(v[0].z, v[1].z), v[2].z)) };
			const unsigned int mx[3] = { float_key(std_max(std_max(v[0].x, v[1].x), v[2].x)), float_key(std_max(std_max(v[0].y, v[1].y), v[2].y)),
			                             float_key(std_max(std_max(v[0].z, v[1].z), v[2].z)) };
			for (int d = 0; d < 3; ++d)
			{
				w.centroid[3 * t + d] = cc[d]; w.tri_min[3 * t + d] = mn[d]; w.tri_max[3 * t + d] = mx[d];
				if (p.work_in_shared) { p.centroid[3 * t + d] = cc[d]; p.tri_min[3 * t + d] = mn[d]; p.tri_max[3 * t + d] = mx[d]; }
			}
			const float nx = p.normals_in[3 * t], ny = p.normals_in[3 * t + 1], nz = p.normals_in[3 * t + 2];
			V3 n = v3(add(add(mul(p.m[0], nx), mul(p.m[4], ny)), mul(p.m[8], nz)),
			          add(add(mul(p.m[1], nx), mul(p.m[5], ny)), mul(p.m[9], nz)),
			          add(add(mul(p.m[2], nx), mul(p.m[6], ny)), mul(p.m[10], nz)));
			normalize(n);
			p.tnormal[3 * t] = n.x; p.tnormal[3 * t + 1] = n.y; p.tnormal[3 * t + 2] = n.z;
			w.order[t] = t;
		}
		__syncthreads();

		// BuildBVH, DataTypes.h:294-308: the root owns everything; its box (UpdateNodeBounds) by the whole CTA
		{
			TeamScratch& s = scratch[0];
			const unsigned int key_max = float_key(FLT_MAX), key_min = float_key(FLT_MIN);
			if (tid < 3) { s.child_lo[0][tid] = key_max; s.child_hi[0][tid] = key_min; }
			__syncthreads();
			unsigned int mn[3] = { key_max, key_max, key_max }, mx[3] = { key_min, key_min, key_min };
			for (int t = tid; t < T; t += kBuildThreads)
				for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], w.tri_min[3 * t + d]); mx[d] = max(mx[d], w.tri_max[3 * t + d]); }
			for (int d = 0; d < 3; ++d) { mn[d] = __reduce_min_sync(0xffffffffu, mn[d]); mx[d] = __reduce_max_sync(0xffffffffu, mx[d]); }
			if (lane == 0)
				for (int d = 0; d < 3; ++d) { atomicMin(&s.child_lo[0][d], mn[d]); atomicMax(&s.child_hi[0][d], mx[d]); }
			__syncthreads();
			if (tid == 0)
			{
				p.node_first[0] = 0; p.node_count[0] = T; p.node_escape[0] = -1;
				for (int d = 0; d < 3; ++d)
				{
					const float lo = key_float(s.child_lo[0][d]), hi = key_float(s.child_hi[0][d]);
					p.node_box[d] = lo; p.node_box[3 + d] = hi;
					p.result_info[kInfoRootBox + d] = __float_as_int(lo); p.result_info[kInfoRootBox + 3 + d] = __float_as_int(hi);
				}
				if (3 * T <= 8) write_node_record(p, 0, true, 0, T);
				else { p.queue_a[0] = 0; level.count[0] = 1; level.big[0] = T; }
			}
			__syncthreads();
		}

		int* handed = nullptr;
		const int n_handed = run_levels<true>(p, w, scratch, level, &nodes_used, p.queue_a, p.queue_b, &handed);

		// what the subtree kernel needs: the order so far, the node counter, the handed-over nodes
		if (p.work_in_shared)
			for (int k = tid; k < T; k += kBuildThreads) p.order[k] = w.order[k];
		if (tid < n_handed) p.result_info[kInfoSubtreeList + tid] = handed[tid];
		if (tid == 0) { p.result_info[kInfoNodesUsed] = nodes_used; p.result_info[kInfoSubtrees] = n_handed; }
	}

	// The order the build leaves behind: indices / normals for the next call, the triangle stream for the pixel kernel
	// ({v0|nx}{e1|ny}{e2|nz}, e1 = v1 - v0, e2 = v2 - v0: Utils.h:143-144), node records and the mesh table's rows.
	__device__ __forceinline__ void write_build_results(const BuildParams& p)
	{
		const int tid = threadIdx.x, n_threads = blockDim.x;
		const int T = p.triangle_count;
		for (int k = tid; k < T; k += n_threads)
		{
			const int slot = p.order[k];
			V3 v[3];
			for (int c = 0; c < 3; ++c)
			{
				const int vi = p.indices_in[3 * slot + c];
				p.indices_out[3 * k + c] = vi;
				p.normals_out[3 * k + c] = p.normals_in[3 * slot + c];
				v[c] = v3(p.tpos[3 * vi], p.tpos[3 * vi + 1], p.tpos[3 * vi + 2]);
			}
			const V3 e1 = v[1] - v[0], e2 = v[2] - v[0];
			const float4 r0 = make_float4(v[0].x, v[0].y, v[0].z, p.tnormal[3 * slot]);
			const float4 r1 = make_float4(e1.x, e1.y, e1.z, p.tnormal[3 * slot + 1]);
			const float4 r2 = make_float4(e2.x, e2.y, e2.z, p.tnormal[3 * slot + 2]);
			p.result_triangles[3 * k + 0] = r0; p.result_triangles[3 * k + 1] = r1; p.result_triangles[3 * k + 2] = r2;
			if (p.scene_triangles) { p.scene_triangles[3 * k + 0] = r0; p.scene_triangles[3 * k + 1] = r1; p.scene_triangles[3 * k + 2] = r2; }
		}
		if (p.scene_nodes)
		{
			const int n_nodes = p.result_info[kInfoNodesUsed];
			for (int i = tid; i < 2 * n_nodes; i += n_threads) p.scene_nodes[i] = p.result_nodes[i];
			if (tid == 0)
			{
				const float4 keep1 = p.scene_table[1], keep2 = p.scene_table[2];
				p.scene_table[0] = make_float4(p.node_box[0], p.node_box[3], p.node_box[1], p.node_box[4]);
				p.scene_table[1] = make_float4(p.node_box[2], p.node_box[5], keep1.z, keep1.w);
				p.scene_table[2] = make_float4(keep2.x, keep2.y, keep2.z, __int_as_float(n_nodes));
			}
		}
	}

	__global__ void __launch_bounds__(kSubtreeThreads)
	build_subtrees_kernel(const __grid_constant__ BuildParams p)
	{
		extern __shared__ __align__(16) float dynamic_shared[];
		__shared__ TeamScratch scratch[kSubtreeThreads / 32];
		__shared__ LevelState level;
		__shared__ int ticket, last_cta;
		const int tid = threadIdx.x;
		const int n_subtrees = p.result_info[kInfoSubtrees];
		int* nodes_used = p.result_info + kInfoNodesUsed;

		while (true)
		{
			if (tid == 0) ticket = atomicAdd(p.result_info + kInfoTicket, 1);
			__syncthreads();
			const int mine = ticket;
			if (mine >= n_subtrees) break;                                  // uniform
			const int root = p.result_info[kInfoSubtreeList + mine];
			const int first = p.node_first[root], count = p.node_count[root];

			// slot-indexed arrays are read-only here; position-indexed scratch of [first, first + count) in shared memory
			BuildWork w;
			w.centroid = p.centroid; w.tri_min = p.tri_min; w.tri_max = p.tri_max;
			const bool in_shared = count <= p.subtree_shared_triangles;
			if (in_shared)
			{
				int32_t* base = reinterpret_cast<int32_t*>(dynamic_shared) - first;
				w.order = base; w.order_tmp = base + count; w.rights_before = base + 2 * count; w.front_right = base + 3 * count; w.back_left = base + 4 * count;
				for (int k = tid; k < count; k += kSubtreeThreads) w.order[first + k] = p.order[first + k];
			}
			else
			{
				w.order = p.order; w.order_tmp = p.order_tmp; w.rights_before = p.rights_before; w.front_right = p.front_right; w.back_left = p.back_left;
			}
			int* cur = p.queue_a + first;                                   // a level of this subtree has fewer than `count` nodes
			int* nxt = p.queue_b + first;
			if (tid == 0) { cur[0] = root; level.count[0] = 1; level.big[0] = count; level.count[1] = 0; level.big[1] = 0; }
			__syncthreads();
			int* unused = nullptr;
			run_levels<false>(p, w, scratch, level, nodes_used, cur, nxt, &unused);
			if (in_shared)
				for (int k = tid; k < count; k += kSubtreeThreads) p.order[first + k] = w.order[first + k];
			__syncthreads();                                                // before `ticket` and the scratch are reused
		}

		// the last CTA to get here writes the results: everything the others wrote is visible to it
		__threadfence();
		__syncthreads();
		if (tid == 0) last_cta = (atomicAdd(p.result_info + kInfoFinished, 1) == (int)gridDim.x - 1) ? 1 : 0;
		__syncthreads();
		if (!last_cta) return;
		__threadfence();
		write_build_results(p);
	}

	// Copies a mesh's last build into the scene's mesh block: triangle stream slice, node slice, and the mesh
	// table's box / node rows.  Runs when the block was rewritten from the host mirror and no new build is due
	// (a rebuild there would advance the triangle order a second time).
	struct EmitParams
	{
		const float4* result_triangles;
		const float4* result_nodes;
		const int32_t* result_info;
		int32_t triangle_count;
		float4* triangles;       // this mesh's slice of the stream
		float4* nodes;           // this mesh's slice of the node array
		float4* table;           // this mesh's 3 rows of the mesh table
	};

	__global__ void __launch_bounds__(256)
	emit_mesh_kernel(const __grid_constant__ EmitParams p)
	{
		const int n_nodes = p.result_info[0];
		for (int i = threadIdx.x; i < 3 * p.triangle_count; i += blockDim.x) p.triangles[i] = p.result_triangles[i];
		for (int i = threadIdx.x; i < 2 * n_nodes; i += blockDim.x) p.nodes[i] = p.result_nodes[i];
		if (threadIdx.x == 0)
		{
			const float4 keep1 = p.table[1], keep2 = p.table[2];
			p.table[0] = make_float4(__int_as_float(p.result_info[2]), __int_as_float(p.result_info[5]), __int_as_float(p.result_info[3]), __int_as_float(p.result_info[6]));
			p.table[1] = make_float4(__int_as_float(p.result_info[4]), __int_as_float(p.result_info[7]), keep1.z, keep1.w);
			p.table[2] = make_float4(keep2.x, keep2.y, keep2.z, __int_as_float(n_nodes));
		}
	}
}
