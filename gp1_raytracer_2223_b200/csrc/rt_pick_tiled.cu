// Instantiations of the tiled pixel kernel (one CTA per 32x8 tile) and of the counters build.
#include "rt_pick.h"

namespace rt
{
	KernelFn pick_kernel(int mode, int shadows, bool bvh)
	{
#define RT_ROW(M) { { render_kernel<M, 0, false, false>, render_kernel<M, 1, false, false> }, { render_kernel<M, 0, true, false>, render_kernel<M, 1, true, false> } }
		static const KernelFn table[4][2][2] = {
			RT_ROW(RT_LIGHTING_OBSERVED_AREA), RT_ROW(RT_LIGHTING_RADIANCE), RT_ROW(RT_LIGHTING_BRDF), RT_ROW(RT_LIGHTING_COMBINED),
		};
#undef RT_ROW
		return table[mode][bvh ? 1 : 0][shadows ? 1 : 0];
	}

	KernelFn pick_kernel_count(bool bvh)
	{
		return bvh ? render_kernel<-1, -1, true, true> : render_kernel<-1, -1, false, true>;
	}
}
