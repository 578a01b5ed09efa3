// The ray-parallel ("wavefront") form of the pixel path: five launches (rt_kernel_wave.cuh).
#include "rt_pick.h"
#include "rt_kernel_wave.cuh"

#include <cstdio>
#include <cstdlib>

namespace rt
{
	cudaError_t wave_launch(const SceneDevice& dev, const FrameParams& p, const wave::WaveParams& w, dim3 grid, int sm_count, cudaStream_t stream)
	{
		const size_t smem = staged_scene_bytes(dev.n_materials);
		cudaError_t e = cudaMemsetAsync(w.counters, 0, 8 * sizeof(unsigned int), stream);
		if (e != cudaSuccess) return e;
		// RT_B200_WAVE_TIMING=1 (measurement only): events between the launches, printed to stderr after a synchronisation
		static const bool timing = getenv("RT_B200_WAVE_TIMING") != nullptr;
		cudaEvent_t ev[6] = {};
		int n_ev = 0;
		auto mark = [&]() { if (timing) { cudaEventCreate(&ev[n_ev]); cudaEventRecord(ev[n_ev], stream); ++n_ev; } };
		mark();
		const unsigned int walkers = (unsigned int)sm_count * 8u;         // 8 CTAs of 8 warps per SM: one wave of walkers, grid-stride over the jobs
		wave::primary_kernel<<<grid, kThreads, smem, stream>>>(dev, p, w);
		mark();
		wave::view_walk_kernel<<<walkers, 256, 0, stream>>>(dev, p, w);
		mark();
		if (p.shadows) wave::shadow_setup_kernel<1><<<grid, kThreads, smem, stream>>>(dev, p, w);
		else wave::shadow_setup_kernel<0><<<grid, kThreads, smem, stream>>>(dev, p, w);
		mark();
		if (p.shadows) wave::shadow_walk_kernel<<<walkers, 256, 0, stream>>>(dev, p, w);
		mark();
		switch (p.lighting_mode)
		{
		case RT_LIGHTING_OBSERVED_AREA: wave::shade_kernel<RT_LIGHTING_OBSERVED_AREA><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		case RT_LIGHTING_RADIANCE: wave::shade_kernel<RT_LIGHTING_RADIANCE><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		case RT_LIGHTING_BRDF: wave::shade_kernel<RT_LIGHTING_BRDF><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		default: wave::shade_kernel<RT_LIGHTING_COMBINED><<<grid, kThreads, smem, stream>>>(dev, p, w); break;
		}
		mark();
		if (timing)
		{
			cudaStreamSynchronize(stream);
			unsigned int counters[8] = {};
			cudaMemcpy(counters, w.counters, sizeof counters, cudaMemcpyDeviceToHost);
			float ms[5] = {};
			for (int i = 0; i + 1 < n_ev; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
			fprintf(stderr, "wave: primary %.1f us, view walk %.1f us (%u jobs), shadow setup %.1f us, shadow walk %.1f us (%u jobs), shade %.1f us\n",
			        ms[0] * 1e3f, ms[1] * 1e3f, counters[0], ms[2] * 1e3f, ms[3] * 1e3f, counters[1], ms[4] * 1e3f);
			for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
		}
		return cudaGetLastError();
	}
	int wave_launch_count(int shadows) { return shadows ? 5 : 4; }
}
