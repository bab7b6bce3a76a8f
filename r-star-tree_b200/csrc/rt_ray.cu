// rt_ray.cu -- instantiations and launcher of the ray traversal kernels (f32, 3-D).
#include "rt_traverse_ray.cuh"

namespace rtb
{
namespace
{
template <int W, bool BOXKEY, uint32_t RMODE>
cudaError_t launch_ray_mode(const RayLaunch& p)
{
  auto kernel = ray_traverse_kernel<W, BOXKEY, RMODE>;
  const size_t smem = traverse_smem_bytes(p.tree.num_levels, W);
  cudaError_t err = cudaSuccess;
  if (smem > 48 * 1024)
  {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess)
      return err;
  }
  RayLaunch q = p;
  int grid = 0;
  if ((err = plan_work((const void*)kernel, smem, p.n, p.grid, p.stream, q.work, grid)) != cudaSuccess)
    return err;
  kernel<<<grid, TRAVERSE_THREADS, smem, p.stream>>>(q);
  return cudaGetLastError();
}

template <int W, bool BOXKEY>
cudaError_t launch_ray_layout(const RayLaunch& p)
{
  switch (p.mode)
  {
  case RT_RAY_ALL_HITS:
    switch (p.pass)
    {
    case PASS_COUNT: return launch_ray_mode<W, BOXKEY, RAY_COUNT>(p);
    case PASS_COUNT_STATS: return launch_ray_mode<W, BOXKEY, RAY_COUNT_STATS>(p);
    case PASS_WRITE: return launch_ray_mode<W, BOXKEY, RAY_WRITE>(p);
    default: return cudaErrorInvalidValue;
    }
  case RT_RAY_CLOSEST:
    return p.pass == PASS_COUNT_STATS ? launch_ray_mode<W, BOXKEY, RAY_CLOSEST_STATS>(p)
                                      : launch_ray_mode<W, BOXKEY, RAY_CLOSEST>(p);
  case RT_RAY_ANY:
    return p.pass == PASS_COUNT_STATS ? launch_ray_mode<W, BOXKEY, RAY_ANY_STATS>(p)
                                      : launch_ray_mode<W, BOXKEY, RAY_ANY>(p);
  default: return cudaErrorInvalidValue;
  }
}
} // namespace

cudaError_t launch_ray_kernel(const RayLaunch& p)
{
  const bool box = p.key_kind == RT_KEY_BOX;
  if (p.width == 8)
    return box ? launch_ray_layout<8, true>(p) : launch_ray_layout<8, false>(p);
  if (p.width == 16)
    return box ? launch_ray_layout<16, true>(p) : launch_ray_layout<16, false>(p);
  return cudaErrorInvalidValue;
}
} // namespace rtb
