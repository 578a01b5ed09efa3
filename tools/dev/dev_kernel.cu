// Development aid: ONE instantiation of the pixel kernel, for quick SASS inspection:
//   tools/dev/sass.sh [extra nvcc flags]   -> /tmp/dev.sass
#include "../../gp1_raytracer_2223_b200/csrc/rt_kernel.cuh"
void* dev_kernel_address() { return (void*)rt::render_kernel_persistent<3, 1, true>; }
