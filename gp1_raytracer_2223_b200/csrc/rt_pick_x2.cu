// Instantiations of the packed two-pixels-per-thread kernel (RT_KERNEL_PACKED).
#include "rt_pick.h"
#include "rt_kernel_x2.cuh"

namespace rt
{
#ifndef RT_EXPERIMENT_BLOCK
	static_assert(x2::kBlockW == kBlockW, "both kernels must cut the frame into the same CTA grid");
#endif
	KernelFn pick_kernel_x2(int mode, int shadows, bool bvh)
	{
#define RT_ROW(M) { { x2::render_kernel_x2<M, 0, false>, x2::render_kernel_x2<M, 1, false> }, { x2::render_kernel_x2<M, 0, true>, x2::render_kernel_x2<M, 1, true> } }
		static const KernelFn table[4][2][2] = {
			RT_ROW(RT_LIGHTING_OBSERVED_AREA), RT_ROW(RT_LIGHTING_RADIANCE), RT_ROW(RT_LIGHTING_BRDF), RT_ROW(RT_LIGHTING_COMBINED),
		};
#undef RT_ROW
		return table[mode][bvh ? 1 : 0][shadows ? 1 : 0];
	}
	int pick_threads_x2() { return x2::kThreads; }
	int pick_block_w_x2() { return x2::kBlockW; }
}
