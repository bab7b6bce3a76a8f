// rt_box_i32.cu -- int32 instantiations of the box traversal kernel.
#include "rt_traverse_box.cuh"

namespace rtb
{
box_launch_fn find_box_kernel_i32(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode)
{
  RTB_BOX_CASE(int32_t, 1, 8)
  RTB_BOX_CASE(int32_t, 2, 8)
  RTB_BOX_CASE(int32_t, 3, 8)
  return nullptr;
}
} // namespace rtb
