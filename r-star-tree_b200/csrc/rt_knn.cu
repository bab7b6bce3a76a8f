// rt_knn.cu -- instantiations and launcher of the kNN traversal kernel: every float32 (dim, width) pair
// the box kernels are instantiated for (rt_box_f32.cu), point and box keys.
#include "rt_traverse_knn.cuh"

namespace rtb
{
namespace
{
template <int D, int W, bool BOXKEY, bool COLLECT = false>
cudaError_t launch_knn_kind(const KnnLaunch& p)
{
  if constexpr (!COLLECT)
  {
    if (p.collect)
      return launch_knn_kind<D, W, BOXKEY, true>(p);
  }
  auto kernel = knn_traverse_kernel<D, W, BOXKEY, COLLECT>;
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

typedef cudaError_t (*knn_launch_fn)(const KnnLaunch&);

#define RTB_KNN_CASE(D, W)                                                                     \
  if (dim == D && width == W)                                                                  \
    return box ? &launch_knn_kind<D, W, true> : &launch_knn_kind<D, W, false>;

knn_launch_fn find_knn_kernel(uint32_t dim, uint32_t width, bool box)
{
  RTB_KNN_CASE(1, 8)
  RTB_KNN_CASE(2, 8)
  RTB_KNN_CASE(3, 8)
  RTB_KNN_CASE(4, 8)
  RTB_KNN_CASE(8, 8)
  RTB_KNN_CASE(2, 16)
  RTB_KNN_CASE(3, 16)
  RTB_KNN_CASE(8, 16)
  return nullptr;
}
} // namespace

bool knn_supported(uint32_t scalar, uint32_t dim, uint32_t width)
{
  return scalar == RT_F32 && find_knn_kernel(dim, width, false) != nullptr;
}

cudaError_t launch_knn_kernel(const KnnLaunch& p, uint32_t dim, uint32_t width, uint32_t key_kind)
{
  const knn_launch_fn launch = find_knn_kernel(dim, width, key_kind == RT_KEY_BOX);
  return launch ? launch(p) : cudaErrorInvalidValue;
}
} // namespace rtb
