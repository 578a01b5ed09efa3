// TriangleMesh::UpdateTransforms as the reference ships it - vertex / normal transform AND BuildBVH - on the
// device (reference source/DataTypes.h:210-236, 294-483 with BVH and USE_BINS defined,
// DataTypes.h:8-9).  One CTA per mesh.
//
// What has to be reproduced, and how:
//  * the tree: per node the binned SAH split of FindBestSplitPlane (DataTypes.h:398-483: centroid bounds, 8 bins,
//    7 planes, 3 axes, first strictly smaller cost wins), the no-split test of Subdivide (DataTypes.h:333-336)
//    and the leaf rule idxCount <= 8 (DataTypes.h:327) - every float operation in the reference's order
//    (rt_device.cuh arithmetic), minima / maxima in any order (exact and commutative);
//  * the triangle order: Subdivide partitions indices / normals IN PLACE with a two-pointer sweep
//    (DataTypes.h:344-363) and the next UpdateTransforms starts from the order the last one left, so the mesh
//    state lives on the device between calls (indices / normals ping-pong buffers) and the sweep is run
//    literally (one lane per node over a precomputed predicate);
//  * the walk order of IntersectionTest_BVH (Utils.h:246-288: left child, then left + 1): children are a pair,
//    every node gets the "escape" link of rt::BvhLink when it is created (left -> right sibling, right ->
//    parent's escape).  Node NUMBERS are free (the walk never compares them): pairs are handed out by an atomic
//    counter instead of the reference's depth-first nodesUsed++, which is what lets a whole tree level be built
//    in parallel, one warp per node.
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
		float* tri_min;                 // 3T   min / max over the slot's three vertices
		float* tri_max;                 // 3T
		float* tnormal;                 // 3T   transformedNormals per slot
		int32_t* order;                 // T    order[k] = slot that sits at position k
		uint8_t* left_flag;             // T    partition predicate per position
		int32_t* node_first;            // N    first triangle position
		int32_t* node_count;            // N    triangles
		int32_t* node_escape;           // N
		float* node_box;                // 6N   min xyz, max xyz
		int32_t* queue_a;               // T    nodes of the level being split
		int32_t* queue_b;               // T    nodes created for the next level
		// results (persistent per mesh: copied into the scene's mesh block by emit_mesh_kernel)
		float4* result_triangles;       // 3T   {v0|nx}{e1|ny}{e2|nz} in the new order
		float4* result_nodes;           // 2N   device node records (rt::BvhLink)
		int32_t* result_info;           // [0] nodes used, [1] status (0 ok, 1 leaf too large for BvhLink), [2..7] root box bits
	};

	constexpr int kBuildThreads = 512;
	constexpr int kBuildWarps = kBuildThreads / 32;

	// order-preserving map float -> unsigned for shared-memory atomicMin / atomicMax (no NaN can reach it)
	__device__ __forceinline__ unsigned int float_key(float f)
	{
		const unsigned int b = __float_as_uint(f);
		return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
	}
	__device__ __forceinline__ float key_float(unsigned int k)
	{
		return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
	}

	__device__ __forceinline__ float warp_min(float v)
	{
		for (int s = 16; s > 0; s >>= 1) v = std_min(v, __shfl_xor_sync(0xffffffffu, v, s));
		return v;
	}
	__device__ __forceinline__ float warp_max(float v)
	{
		for (int s = 16; s > 0; s >>= 1) v = std_max(v, __shfl_xor_sync(0xffffffffu, v, s));
		return v;
	}

	// AABB::Area, DataTypes.h:75-79
	__device__ __forceinline__ float box_area(const float mn[3], const float mx[3])
	{
		const float ex = sub(mx[0], mn[0]), ey = sub(mx[1], mn[1]), ez = sub(mx[2], mn[2]);
		return add(add(mul(ex, ey), mul(ey, ez)), mul(ez, ex));
	}

	// UpdateNodeBounds, DataTypes.h:310-321, over positions [first, first + count) of `order`: one warp.
	// Starts from +FLT_MAX / +FLT_MIN like the reference (Vector3.cpp:13-14).  Every lane returns the box.
	__device__ __forceinline__ void warp_range_bounds(const BuildParams& p, int first, int count, int lane, float mn[3], float mx[3])
	{
		for (int k = 0; k < 3; ++k) { mn[k] = FLT_MAX; mx[k] = FLT_MIN; }
		for (int t = lane; t < count; t += 32)
		{
			const int slot = p.order[first + t];
			for (int k = 0; k < 3; ++k)
			{
				mn[k] = std_min(mn[k], p.tri_min[3 * slot + k]);
				mx[k] = std_max(mx[k], p.tri_max[3 * slot + k]);
			}
		}
		for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
	}

	// The device record of a finished node (SceneDevice::bvh_nodes): box, first child / first triangle, link.
	__device__ __forceinline__ void write_node_record(const BuildParams& p, int node, bool leaf, int first_or_child, int leaf_triangles)
	{
		const float* b = p.node_box + 6 * node;
		if (leaf && leaf_triangles > BvhLink::kMaxLeafTriangles) atomicExch(p.result_info + 1, 1);
		const int link = (p.node_escape[node] + 1) | ((leaf ? leaf_triangles : 0) << BvhLink::kEscapeBits);
		p.result_nodes[2 * node] = make_float4(b[0], b[3], b[1], b[4]);
		p.result_nodes[2 * node + 1] = make_float4(b[2], b[5], __int_as_float(first_or_child), __int_as_float(link));
	}

	struct BinScratch
	{
		unsigned int count[8];
		unsigned int lo[8][3];      // float_key of the bin box minimum
		unsigned int hi[8][3];
	};

	// Subdivide (DataTypes.h:323-389) of one node by one warp.  Children are appended to `next`.
	__device__ __forceinline__ void subdivide_node(const BuildParams& p, int node, int lane, BinScratch& bins,
	                                               int* nodes_used, int* next, int* next_count)
	{
		const int first = p.node_first[node], count = p.node_count[node];
		const unsigned int idx_count = 3u * (unsigned int)count;

		// ---- FindBestSplitPlane, DataTypes.h:398-483 ----
		float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
		for (int t = lane; t < count; t += 32)
		{
			const int slot = p.order[first + t];
			for (int k = 0; k < 3; ++k)
			{
				const float c = p.centroid[3 * slot + k];
				lo[k] = std_min(lo[k], c);
				hi[k] = std_max(hi[k], c);
			}
		}
		for (int k = 0; k < 3; ++k) { lo[k] = warp_min(lo[k]); hi[k] = warp_max(hi[k]); }

		float best = FLT_MAX, split_pos = 0.f;
		int axis = 0;
		for (int a = 0; a < 3; ++a)
		{
			const float diff = sub(hi[a], lo[a]);
			if (fabsf(diff) < FLT_EPSILON) continue;                                   // DataTypes.h:421-422 (warp-uniform)

			// bins, DataTypes.h:425-441: counts and boxes through shared-memory atomics
			__syncwarp();
			if (lane < 8)
			{
				bins.count[lane] = 0u;
				for (int k = 0; k < 3; ++k) { bins.lo[lane][k] = float_key(FLT_MAX); bins.hi[lane][k] = float_key(FLT_MIN); }
			}
			__syncwarp();
			const float scale = quo(8.f, diff);
			for (int t = lane; t < count; t += 32)
			{
				const int slot = p.order[first + t];
				int bin = __float2int_rz(mul(sub(p.centroid[3 * slot + a], lo[a]), scale));
				bin = min(7, bin);
				atomicAdd(&bins.count[bin], 3u);
				for (int k = 0; k < 3; ++k)
				{
					atomicMin(&bins.lo[bin][k], float_key(p.tri_min[3 * slot + k]));
					atomicMax(&bins.hi[bin][k], float_key(p.tri_max[3 * slot + k]));
				}
			}
			__syncwarp();

			// the 7 planes, DataTypes.h:443-466: lane i owns plane i (left = bins 0..i, right = bins i+1..7)
			float cost = 0.f;
			bool candidate = false;
			if (lane < 7)
			{
				float lmn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, lmx[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
				float rmn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, rmx[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
				int left_count = 0, right_count = 0;
				for (int b = 0; b < 8; ++b)
				{
					const bool is_left = b <= lane;
					const int c = (int)bins.count[b];
					if (is_left) left_count += c; else right_count += c;
					for (int k = 0; k < 3; ++k)
					{
						const float bl = key_float(bins.lo[b][k]), bh = key_float(bins.hi[b][k]);
						if (is_left) { lmn[k] = std_min(lmn[k], bl); lmx[k] = std_max(lmx[k], bh); }
						else { rmn[k] = std_min(rmn[k], bl); rmx[k] = std_max(rmx[k], bh); }
					}
				}
				// DataTypes.h:472: leftCount[i] * leftArea[i] + rightCount[i] * rightArea[i]
				cost = add(mul((float)left_count, box_area(lmn, lmx)), mul((float)right_count, box_area(rmn, rmx)));
				candidate = cost < best;                                              // false for NaN (an empty side: 0 * inf)
			}
			// "first strictly smaller cost wins" over planes 0..6 == the lowest plane that holds the minimum of the
			// costs below the running best
			float m = candidate ? cost : INFINITY;
			for (int s = 4; s > 0; s >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s, 8));
			m = __shfl_sync(0xffffffffu, m, 0);
			const unsigned int winners = __ballot_sync(0xffffffffu, candidate && cost == m);
			if (winners)
			{
				const int plane = __ffs(winners) - 1;
				axis = a;
				split_pos = add(lo[a], mul(quo(diff, 8.f), (float)(plane + 1)));           // DataTypes.h:469, 476
				best = m;
			}
		}

		// ---- Subdivide, DataTypes.h:333-336: keep the node as a leaf if splitting is not cheaper ----
		const float no_split = mul((float)idx_count, box_area(p.node_box + 6 * node, p.node_box + 6 * node + 3));
		bool leaf = best >= no_split;

		int left_count = 0;
		if (!leaf)
		{
			// ---- the in-place partition, DataTypes.h:344-363: predicate in parallel, sweep by one lane ----
			for (int t = lane; t < count; t += 32)
				p.left_flag[first + t] = p.centroid[3 * p.order[first + t] + axis] < split_pos ? 1 : 0;
			__syncwarp();
			if (lane == 0)
			{
				int i = first, j = first + count - 1;
				while (i <= j)
				{
					if (p.left_flag[i]) ++i;
					else
					{
						const int oi = p.order[i], oj = p.order[j];
						p.order[i] = oj; p.order[j] = oi;
						p.left_flag[i] = p.left_flag[j];
						--j;
					}
				}
				left_count = i - first;
			}
			left_count = __shfl_sync(0xffffffffu, left_count, 0);
			__syncwarp();
			leaf = (left_count == 0 || left_count == count);                           // DataTypes.h:366-369
		}

		if (leaf)
		{
			if (lane == 0) write_node_record(p, node, true, first, count);
			return;
		}

		// ---- children, DataTypes.h:371-388 ----
		int pair = 0;
		if (lane == 0) pair = atomicAdd(nodes_used, 2);
		pair = __shfl_sync(0xffffffffu, pair, 0);
		const int child[2] = { pair, pair + 1 };
		const int child_first[2] = { first, first + left_count };
		const int child_count[2] = { left_count, count - left_count };
		for (int c = 0; c < 2; ++c)
		{
			float mn[3], mx[3];
			warp_range_bounds(p, child_first[c], child_count[c], lane, mn, mx);
			if (lane == 0)
			{
				p.node_first[child[c]] = child_first[c];
				p.node_count[child[c]] = child_count[c];
				p.node_escape[child[c]] = (c == 0) ? child[1] : p.node_escape[node];
				for (int k = 0; k < 3; ++k) { p.node_box[6 * child[c] + k] = mn[k]; p.node_box[6 * child[c] + 3 + k] = mx[k]; }
				if (3 * child_count[c] <= 8) write_node_record(p, child[c], true, child_first[c], child_count[c]);   // DataTypes.h:327
				else next[atomicAdd(next_count, 1)] = child[c];
			}
		}
		if (lane == 0) write_node_record(p, node, false, child[0], 0);
		__syncwarp();
	}

	__global__ void __launch_bounds__(kBuildThreads)
	update_transforms_bvh_kernel(const __grid_constant__ BuildParams p)
	{
		__shared__ BinScratch bins[kBuildWarps];
		__shared__ int nodes_used, level_count[2];
		const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
		const int T = p.triangle_count;

		// transformedPositions, DataTypes.h:216-222 (Matrix::TransformPoint, Matrix.cpp:49-56)
		for (int v = tid; v < p.vertex_count; v += kBuildThreads)
		{
			const V3 q = transform_point(p.m, p.positions[3 * v], p.positions[3 * v + 1], p.positions[3 * v + 2]);
			p.tpos[3 * v] = q.x; p.tpos[3 * v + 1] = q.y; p.tpos[3 * v + 2] = q.z;
		}
		if (tid == 0) { nodes_used = 1; level_count[0] = 0; level_count[1] = 0; p.result_info[1] = 0; }
		__syncthreads();

		// per triangle slot: transformedNormals (DataTypes.h:224-230), centroid (DataTypes.h:349), vertex min / max
		for (int t = tid; t < T; t += kBuildThreads)
		{
			V3 v[3];
			for (int k = 0; k < 3; ++k)
			{
				const int vi = p.indices_in[3 * t + k];
				v[k] = v3(p.tpos[3 * vi], p.tpos[3 * vi + 1], p.tpos[3 * vi + 2]);
			}
			const V3 c = ((v[0] + v[1]) + v[2]) * 0.3333f;
			p.centroid[3 * t] = c.x; p.centroid[3 * t + 1] = c.y; p.centroid[3 * t + 2] = c.z;
			p.tri_min[3 * t] = std_min(std_min(v[0].x, v[1].x), v[2].x); p.tri_max[3 * t] = std_max(std_max(v[0].x, v[1].x), v[2].x);
			p.tri_min[3 * t + 1] = std_min(std_min(v[0].y, v[1].y), v[2].y); p.tri_max[3 * t + 1] = std_max(std_max(v[0].y, v[1].y), v[2].y);
			p.tri_min[3 * t + 2] = std_min(std_min(v[0].z, v[1].z), v[2].z); p.tri_max[3 * t + 2] = std_max(std_max(v[0].z, v[1].z), v[2].z);
			const float nx = p.normals_in[3 * t], ny = p.normals_in[3 * t + 1], nz = p.normals_in[3 * t + 2];
			V3 n = v3(add(add(mul(p.m[0], nx), mul(p.m[4], ny)), mul(p.m[8], nz)),
			          add(add(mul(p.m[1], nx), mul(p.m[5], ny)), mul(p.m[9], nz)),
			          add(add(mul(p.m[2], nx), mul(p.m[6], ny)), mul(p.m[10], nz)));
			normalize(n);
			p.tnormal[3 * t] = n.x; p.tnormal[3 * t + 1] = n.y; p.tnormal[3 * t + 2] = n.z;
			p.order[t] = t;
		}
		__syncthreads();

		// BuildBVH, DataTypes.h:294-308: the root owns everything
		if (warp == 0)
		{
			float mn[3], mx[3];
			warp_range_bounds(p, 0, T, lane, mn, mx);
			if (lane == 0)
			{
				p.node_first[0] = 0; p.node_count[0] = T; p.node_escape[0] = -1;
				for (int k = 0; k < 3; ++k)
				{
					p.node_box[k] = mn[k]; p.node_box[3 + k] = mx[k];
					p.result_info[2 + k] = __float_as_int(mn[k]); p.result_info[5 + k] = __float_as_int(mx[k]);
				}
				if (3 * T <= 8) write_node_record(p, 0, true, 0, T);
				else { p.queue_a[0] = 0; level_count[0] = 1; }
			}
		}
		__syncthreads();

		// one tree level per round, one warp per node
		int* cur = p.queue_a;
		int* nxt = p.queue_b;
		int parity = 0;
		while (true)
		{
			const int n = level_count[parity];
			if (n == 0) break;
			for (int q = warp; q < n; q += kBuildWarps)
				subdivide_node(p, cur[q], lane, bins[warp], &nodes_used, nxt, &level_count[parity ^ 1]);
			__syncthreads();
			if (tid == 0) level_count[parity] = 0;
			int* swap = cur; cur = nxt; nxt = swap;
			parity ^= 1;
			__syncthreads();
		}

		// the order the build leaves behind: indices / normals for the next call, the triangle stream for the
		// pixel kernel ({v0|nx}{e1|ny}{e2|nz}, e1 = v1 - v0, e2 = v2 - v0: Utils.h:143-144)
		for (int k = tid; k < T; k += kBuildThreads)
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
			p.result_triangles[3 * k + 0] = make_float4(v[0].x, v[0].y, v[0].z, p.tnormal[3 * slot]);
			p.result_triangles[3 * k + 1] = make_float4(e1.x, e1.y, e1.z, p.tnormal[3 * slot + 1]);
			p.result_triangles[3 * k + 2] = make_float4(e2.x, e2.y, e2.z, p.tnormal[3 * slot + 2]);
		}
		if (tid == 0) p.result_info[0] = nodes_used;
	}

	// Copies a mesh's last build into the scene's mesh block: triangle stream slice, node slice, and the mesh
	// table's box / node rows.  Runs after every build and whenever the block was rewritten from the host mirror
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
