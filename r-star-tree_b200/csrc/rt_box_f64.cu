// rt_box_f64.cu -- float64 instantiations of the box traversal kernel.
#include "rt_traverse_box.cuh"

namespace rtb
{
box_launch_fn find_box_kernel_f64(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode)
{
  RTB_BOX_CASE(double, 1, 8)
  RTB_BOX_CASE(double, 2, 8)
  RTB_BOX_CASE(double, 3, 8)
  return nullptr;
}
} // namespace rtb
