// Instantiations of the persistent-warp pixel kernel (the shipped form).
#include "rt_pick.h"

namespace rt
{
	KernelFn pick_kernel_persistent(int mode, int shadows, bool bvh)
	{
#define RT_ROW(M) { { render_kernel_persistent<M, 0, false>, render_kernel_persistent<M, 1, false> }, { render_kernel_persistent<M, 0, true>, render_kernel_persistent<M, 1, true> } }
		static const KernelFn table[4][2][2] = {
			RT_ROW(RT_LIGHTING_OBSERVED_AREA), RT_ROW(RT_LIGHTING_RADIANCE), RT_ROW(RT_LIGHTING_BRDF), RT_ROW(RT_LIGHTING_COMBINED),
		};
#undef RT_ROW
		return table[mode][bvh ? 1 : 0][shadows ? 1 : 0];
	}
}
