// The ray-parallel ("wavefront") form of the pixel path: five launches (rt_kernel_wave.cuh).
#include "rt_pick.h"
#include "rt_kernel_wave.cuh"

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

namespace rt
{
	namespace
	{
		// A launch that may begin before the launch ahead of it in the stream has finished (programmatic dependent launch):
		// its CTAs become resident as that kernel's drain away, run their prologue - staging the scene - and then wait in
		// wait_for_previous_launch() until everything the previous kernel wrote is visible.  A small frame is five short
		// kernels; this takes the launch latency and the prologues out of the gaps between them.
		template <class... Params, class... Args>
		cudaError_t launch_chained(bool chained, void (*kernel)(Params...), dim3 grid, unsigned int threads, size_t smem, cudaStream_t stream, const Args&... args)
		{
			cudaLaunchConfig_t config = {};
			config.gridDim = grid; config.blockDim = dim3(threads); config.dynamicSmemBytes = smem; config.stream = stream;
			cudaLaunchAttribute attribute[1];
			attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
			attribute[0].val.programmaticStreamSerializationAllowed = 1;
			config.attrs = attribute; config.numAttrs = chained ? 1u : 0u;
			return cudaLaunchKernelEx(&config, kernel, args...);
		}
	}

	cudaError_t wave_launch(const SceneDevice& dev, const FrameParams& p, const wave::WaveParams& w_in, dim3 grid, int sm_count, cudaStream_t stream)
	{
		wave::WaveParams w = w_in;
		w.job_cycles = nullptr;
		// RT_B200_WAVE_JOB_CLOCKS=1 (measurement only, with RT_B200_WAVE_TIMING): the clocks every job took
		static const bool job_clocks = getenv("RT_B200_WAVE_JOB_CLOCKS") != nullptr && getenv("RT_B200_WAVE_TIMING") != nullptr;
		if (job_clocks)
		{
			cudaMalloc(&w.job_cycles, sizeof(unsigned int) * wave::kFine * ((size_t)w.view_capacity + w.shadow_capacity));
			cudaMemsetAsync(w.job_cycles, 0, sizeof(unsigned int) * wave::kFine * ((size_t)w.view_capacity + w.shadow_capacity), stream);
		}
		const size_t smem = staged_scene_bytes(dev.n_materials);
		cudaError_t e = cudaMemsetAsync(w.counters, 0, 8 * sizeof(unsigned int), stream);
		if (e != cudaSuccess) return e;
		// RT_B200_WAVE_TIMING=1 (measurement only): events between the launches, printed to stderr after a synchronisation
		static const bool timing = getenv("RT_B200_WAVE_TIMING") != nullptr;
		cudaEvent_t ev[6] = {};
		int n_ev = 0;
		auto mark = [&]() { if (timing) { cudaEventCreate(&ev[n_ev]); cudaEventRecord(ev[n_ev], stream); ++n_ev; } };
		mark();
		// the walkers: one wave of CTAs of kWalkWarps warps, units come from a counter; with PARTS every warp has room for
		// the part it is walking
		constexpr size_t parts_smem = (size_t)wave::kWalkWarps * wave::kRegionBytes;
		static const unsigned int walkers_per_sm = [&]
		{
			int n[4] = {};
			cudaFuncSetAttribute(wave::view_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)parts_smem);
			cudaFuncSetAttribute(wave::shadow_walk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)parts_smem);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n[0], wave::view_walk_kernel<true>, wave::kWalkWarps * 32, parts_smem);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n[1], wave::shadow_walk_kernel<true>, wave::kWalkWarps * 32, parts_smem);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n[2], wave::view_walk_kernel<false>, wave::kWalkWarps * 32, 0);
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n[3], wave::shadow_walk_kernel<false>, wave::kWalkWarps * 32, 0);
			return (unsigned int)std::max(1, *std::min_element(n, n + 4));
		}();
		const unsigned int walkers = (unsigned int)sm_count * walkers_per_sm;
		static const bool chained = getenv("RT_B200_WAVE_NO_CHAIN") == nullptr;
		const dim3 walk_grid(walkers);
		const unsigned int walk_threads = wave::kWalkWarps * 32;
		wave::primary_kernel<<<grid, kThreads, smem, stream>>>(dev, p, w);
		mark();
		if (w.view_parts) e = launch_chained(chained, wave::view_walk_kernel<true>, walk_grid, walk_threads, parts_smem, stream, dev, p, w);
		else e = launch_chained(chained, wave::view_walk_kernel<false>, walk_grid, walk_threads, 0, stream, dev, p, w);
		if (e != cudaSuccess) return e;
		mark();
		if (p.shadows && dev.n_lights > 0)
		{
			e = launch_chained(chained, wave::shadow_setup_kernel, dim3(grid.x, grid.y, w.setup_per_light ? (unsigned int)dev.n_lights : 1u), kThreads, smem, stream, dev, p, w);
			if (e != cudaSuccess) return e;
		}
		mark();
		if (p.shadows && dev.n_lights > 0)
		{
			if (w.shadow_parts) e = launch_chained(chained, wave::shadow_walk_kernel<true>, walk_grid, walk_threads, parts_smem, stream, dev, p, w);
			else e = launch_chained(chained, wave::shadow_walk_kernel<false>, walk_grid, walk_threads, 0, stream, dev, p, w);
			if (e != cudaSuccess) return e;
		}
		mark();
		switch (p.lighting_mode)
		{
		case RT_LIGHTING_OBSERVED_AREA: e = launch_chained(chained, wave::shade_kernel<RT_LIGHTING_OBSERVED_AREA>, grid, kThreads, smem, stream, dev, p, w); break;
		case RT_LIGHTING_RADIANCE: e = launch_chained(chained, wave::shade_kernel<RT_LIGHTING_RADIANCE>, grid, kThreads, smem, stream, dev, p, w); break;
		case RT_LIGHTING_BRDF: e = launch_chained(chained, wave::shade_kernel<RT_LIGHTING_BRDF>, grid, kThreads, smem, stream, dev, p, w); break;
		default: e = launch_chained(chained, wave::shade_kernel<RT_LIGHTING_COMBINED>, grid, kThreads, smem, stream, dev, p, w); break;
		}
		if (e != cudaSuccess) return e;
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
			if (w.job_cycles)
			{
				auto report = [&](const char* name, size_t first, size_t n)
				{
					if (n == 0) return;
					std::vector<unsigned int> c(n);
					cudaMemcpy(c.data(), w.job_cycles + first, n * sizeof(unsigned int), cudaMemcpyDeviceToHost);
					c.erase(std::remove(c.begin(), c.end(), 0u), c.end());       // units that never ran (whole-subtree units use the first n)
					n = c.size();
					if (n == 0) return;
					std::sort(c.begin(), c.end());
					double sum = 0; for (unsigned int v : c) sum += v;
					fprintf(stderr, "wave: %s units: clocks mean %.0f, p50 %u, p90 %u, p99 %u, max %u; sum / (%d SMs x 64 warps) = %.0f clocks\n",
					        name, sum / (double)n, c[n / 2], c[n * 9 / 10], c[n * 99 / 100], c[n - 1], sm_count, sum / (sm_count * 64.0));
				};
				report("view", 0, (size_t)(w.view_parts ? wave::kFine : 1) * std::min(counters[0], w.view_capacity));
				report("shadow", (size_t)wave::kFine * w.view_capacity, (size_t)(w.shadow_parts ? wave::kFine : 1) * std::min(counters[1], w.shadow_capacity));
			}
		}
		if (w.job_cycles) cudaFree(w.job_cycles);
		return cudaGetLastError();
	}
	int wave_launch_count(int shadows) { return shadows ? 5 : 3; }
}
