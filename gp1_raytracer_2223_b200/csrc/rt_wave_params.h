// What the host (rt_api.cu) and the kernels of RT_KERNEL_WAVEFRONT (rt_kernel_wave.cuh) share: the layout of the
// per-mesh split tables and the scratch buffers of one frame.
#pragma once

#include "rt_kernel.cuh"

namespace rt
{
namespace wave
{
	// Two levels.  A mesh's tree is cut into at most kMaxSubtrees subtrees: one bit each in the per-ray masks walk_top
	// fills, one job per (warp tile, subtree) some ray of the tile reaches.  Every subtree is cut again into at most kFine
	// parts, and a walk kernel's unit of work is one job's rays in ONE part: a frame's walk kernels take as long as their
	// longest unit (tools/wave_jobs.py), so the parts are what bounds a small frame's time.  A ray that reaches the
	// subtree's root enters a part after the boxes between that root and the part's root (at most kFineAncestors), which
	// the recursion of Utils.h:246-288 would have tested on its way down.
	constexpr int kMaxSubtrees = 64;
	constexpr int kFineShift = 2, kFine = 1 << kFineShift;
	constexpr int kFineAncestors = 3;
	// Cutting finer only pays while the machine would wait for the longest unit: every unit sets its rays up again, and
	// the shared memory the parts are copied into is taken from the L1 that whole-subtree walks live on.  So a walk
	// kernel with many jobs takes whole subtrees as its units, from global memory (it is bound by the sum of its work, not
	// by the longest piece).  The host chooses per launch (WaveParams::view_parts / shadow_parts) from the job counts of
	// the frame before; choosing per job by the number of rays in it measured worse.
	// per mesh: [0] = subtree count, [1] = offset of the mesh's node -> subtree map in `root_map`, then per subtree
	// kFine + 1 records of kSplitWords ints, the subtree's parts and then (record kFine) the subtree as a whole:
	//   [0] root, [1] end      byte offsets of node records (rt::BvhLink): the walk of the part starts at `root` and is over
	//                          when it arrives at `end`, the root's escape link
	//   [2] first descendant (byte offset), [3] descendants, [4] first triangle, [5] triangles
	//   [6] kPartPresent | kPartStageable: the walk kernels may copy the part into shared memory when the root's
	//       descendants are the contiguous records [2], [2] + [3] (the reference allocates a node's whole left subtree
	//       before its right one, DataTypes.h:372-388), its triangles the contiguous range [4], [4] + [5], and all of it
	//       fits kStageBytes
	//   [7] ancestors, [8..] their byte offsets: the nodes between the subtree's root (inclusive) and the part's root
	//       (exclusive)
	constexpr int kSplitWords = 12;
	constexpr int kSplitHeader = 2;
	constexpr int kSplitStride = kSplitHeader + kMaxSubtrees * (kFine + 1) * kSplitWords;
	constexpr int kPartPresent = 1, kPartStageable = 2;
	// the walk kernels: kWalkWarps warps per CTA, each with kStageBytes of shared memory for the part it is walking
	constexpr int kWalkWarps = 4;
	constexpr int kStageBytes = 6144;                                 // root + descendants + triangles of a stageable part
	constexpr int kRegionBytes = kStageBytes + 32 * kFineAncestors;   // + the boxes above it

	constexpr unsigned long long kNoHit = ((unsigned long long)0x7f7fffffu << 32) | 0xffffffffull;     // t = FLT_MAX, no primitive
	constexpr unsigned int kPlaneBase = (unsigned int)kMaxSpheres, kTriangleBase = (unsigned int)(kMaxSpheres + kMaxPlanes);

	struct WaveParams
	{
		unsigned long long* hit_key;       // per pixel of the launch (CTA-major: cta * kThreads + thread)
		float4* shadow_origin;             // per pixel: {origin + normal * 1e-4, 1 if the view ray hit anything}
		unsigned int* occluded;            // per pixel: bit li = light li's shadow ray is blocked
		unsigned long long* view_alive;    // per (pixel, mesh): the subtrees the view ray reaches (walk_top)
		unsigned long long* shadow_alive;  // per (pixel, light, mesh): the same for the shadow rays
		uint2* view_jobs;                  // {warp tile, mesh << 8 | subtree}: some ray of the tile reaches that subtree's root
		uint2* shadow_jobs;                // {warp tile, light << 16 | mesh << 8 | subtree}
		unsigned int* counters;            // [0] view jobs, [1] shadow jobs, [2] / [3] next job of the view / shadow walk, [4] overflow flag
		const int32_t* split;              // kMaxMeshes * kSplitStride
		const uint8_t* root_map;           // per node of every mesh: subtree number + 1 if the node is a subtree root, else 0
		unsigned int view_capacity, shadow_capacity;
		unsigned int view_parts, shadow_parts;      // 1: that walk kernel takes the subtrees' parts as its units (the host's choice)
		unsigned int* jobs_report;         // mapped host memory: the shade kernel leaves {view jobs, shadow jobs} there
		unsigned int setup_per_light;      // 1: shadow setup runs one thread per (pixel, light)
		unsigned int* job_cycles;          // measurement only (RT_B200_WAVE_JOB_CLOCKS): clocks per unit of the view walk, then of the shadow walk; normally null
	};
}
}
