// The ray-parallel ("wavefront") form of the pixel path: five launches (rt_kernel_wave.cuh).
#include "rt_pick.h"
#include "rt_kernel_wave.cuh"

namespace rt
{
	cudaError_t wave_launch(const SceneDevice& dev, const FrameParams& p, const wave::WaveParams& w, dim3 grid, int sm_count, cudaStream_t stream)
	{
		const size_t smem = staged_scene_bytes(dev.n_materials);
		cudaError_t e = cudaMemsetAsync(w.counters, 0, 8 * sizeof(unsigned int), stream);
		if (e != cudaSuccess) return e;
		const unsigned int walkers = (unsigned int)sm_count * 8u;         // 8 CTAs of 8 warps per SM: one wave of walkers, grid-stride over the jobs
		wave::primary_kernel<<<grid, kThreads, smem, stream>>>(dev, p, w);
		wave::view_walk_kernel<<<walkers, 256, 0, stream>>>(dev, p, w);
		if (p.shadows) wave::shadow_setup_kernel<1><<<grid, kThreads, smem, stream>>>(dev, p, w);
		else wave::shadow_setup_kernel<0><<<grid, kThreads, smem, stream>>>(dev, p, w);
		if (p.shadows) wave::shadow_walk_kernel<<<walkers, 256, 0, stream>>>(dev, p, w);
		switch (p.lighting_mode)
		{
		case RT_LIGHTING_OBSERVED_AREA: wave::shade_kernel<RT_LIGHTING_OBSERVED_AREA><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		case RT_LIGHTING_RADIANCE: wave::shade_kernel<RT_LIGHTING_RADIANCE><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		case RT_LIGHTING_BRDF: wave::shade_kernel<RT_LIGHTING_BRDF><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		default: wave::shade_kernel<RT_LIGHTING_COMBINED><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		}
		return cudaGetLastError();
	}
	int wave_launch_count(int shadows) { return shadows ? 5 : 4; }
}
