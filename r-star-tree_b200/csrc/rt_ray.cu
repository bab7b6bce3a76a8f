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
  const size_t smem = (size_t)(TRAVERSE_THREADS / 32) * (32u * p.tree.num_levels + HIT_BUF) * sizeof(uint32_t);
  cudaError_t err = cudaSuccess;
  if (smem > 48 * 1024)
  {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess)
      return err;
  }
  int grid = p.grid;
  if (grid <= 0)
  {
    int device = 0, sms = 0, per_sm = 0;
    if ((err = cudaGetDevice(&device)) != cudaSuccess)
      return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
      return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRAVERSE_THREADS, smem))
        != cudaSuccess)
      return err;
    grid = sms * (per_sm < 1 ? 1 : per_sm);
  }
  const unsigned long long chunks = (p.n + QUERY_CHUNK - 1) / QUERY_CHUNK;
  const unsigned long long needed = (chunks + TRAVERSE_THREADS / 32 - 1) / (TRAVERSE_THREADS / 32);
  if ((unsigned long long)grid > needed)
    grid = (int)(needed > 0 ? needed : 1);
  kernel<<<grid, TRAVERSE_THREADS, smem, p.stream>>>(p);
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
  case RT_RAY_CLOSEST: return launch_ray_mode<W, BOXKEY, RAY_CLOSEST>(p);
  case RT_RAY_ANY: return launch_ray_mode<W, BOXKEY, RAY_ANY>(p);
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
