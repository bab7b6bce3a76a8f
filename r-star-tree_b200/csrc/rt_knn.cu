// rt_knn.cu -- instantiations and launcher of the kNN traversal kernel (float32, 2-D / 3-D, W = 8).
#include "rt_traverse_knn.cuh"

namespace rtb
{
namespace
{
template <int D, int W, bool BOXKEY>
cudaError_t launch_knn_kind(const KnnLaunch& p)
{
  auto kernel = knn_traverse_kernel<D, W, BOXKEY>;
  const size_t smem = (size_t)(TRAVERSE_THREADS / 32) * 2 * (32u / W + 32u * p.tree.num_levels) * sizeof(uint32_t);
  cudaError_t err = cudaSuccess;
  if (smem > 48 * 1024)
  {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess)
      return err;
  }
  KnnLaunch q = p;
  int grid = 0;
  if ((err = plan_work((const void*)kernel, smem, p.n, p.grid, p.stream, q.work, grid)) != cudaSuccess)
    return err;
  kernel<<<grid, TRAVERSE_THREADS, smem, p.stream>>>(q);
  return cudaGetLastError();
}
} // namespace

bool knn_supported(uint32_t scalar, uint32_t dim, uint32_t width)
{
  return scalar == RT_F32 && (dim == 2 || dim == 3) && width == 8;
}

cudaError_t launch_knn_kernel(const KnnLaunch& p, uint32_t dim, uint32_t width, uint32_t key_kind)
{
  if (width != 8)
    return cudaErrorInvalidValue;
  const bool box = key_kind == RT_KEY_BOX;
  if (dim == 2)
    return box ? launch_knn_kind<2, 8, true>(p) : launch_knn_kind<2, 8, false>(p);
  if (dim == 3)
    return box ? launch_knn_kind<3, 8, true>(p) : launch_knn_kind<3, 8, false>(p);
  return cudaErrorInvalidValue;
}
} // namespace rtb
