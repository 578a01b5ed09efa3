// Kernel selection across translation units: the pixel kernel's template instantiations are compiled in their own
// .cu files (rt_pick_tiled.cu, rt_pick_persistent.cu, rt_pick_x2.cu) so that ptxas works on them in parallel; each
// exports one function returning the instantiation for a lighting mode / shadows / mesh body.
#pragma once

#include "rt_kernel.cuh"
#include "rt_wave_params.h"

namespace rt
{
	using KernelFn = void (*)(const SceneDevice, const FrameParams);

	KernelFn pick_kernel(int mode, int shadows, bool bvh);               // render_kernel<M, S, BVH, false>, kThreads per CTA
	KernelFn pick_kernel_count(bool bvh);                                // render_kernel<-1, -1, BVH, true>: the counters build
	KernelFn pick_kernel_x2(int mode, int shadows, bool bvh);            // x2::render_kernel_x2, pick_threads_x2() per CTA
	int pick_threads_x2();
	int pick_block_w_x2();
	KernelFn pick_kernel_persistent(int mode, int shadows, bool bvh);    // render_kernel_persistent, kPersistentThreads per CTA

	// RT_KERNEL_WAVEFRONT (rt_kernel_wave.cuh): all five launches of one frame share, BVH body only
	cudaError_t wave_launch(const SceneDevice& dev, const FrameParams& p, const wave::WaveParams& w, dim3 grid, int sm_count, cudaStream_t stream);
	int wave_launch_count(int shadows);
}
