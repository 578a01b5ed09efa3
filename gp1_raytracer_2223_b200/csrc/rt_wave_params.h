// What the host (rt_api.cu) and the kernels of RT_KERNEL_WAVEFRONT (rt_kernel_wave.cuh) share: the layout of the
// per-mesh split tables and the scratch buffers of one frame.
#pragma once

#include "rt_kernel.cuh"

namespace rt
{
namespace wave
{
	constexpr int kMaxSubtrees = 64;
	constexpr int kMaxAncestors = 16;
	// per mesh: [0] = subtree count, [1] = offset of the mesh's node -> subtree map in `root_map`, then kMaxSubtrees records
	// of kSplitWords ints {root, end, ancestor count, ancestors...} (byte offsets of node records, rt::BvhLink)
	constexpr int kSplitWords = 3 + kMaxAncestors;
	constexpr int kSplitHeader = 2;
	constexpr int kSplitStride = kSplitHeader + kMaxSubtrees * kSplitWords;

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
	};
}
}
