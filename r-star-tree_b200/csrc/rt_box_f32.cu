// rt_box_f32.cu -- float32 instantiations of the box traversal kernel.
#include "rt_traverse_box.cuh"

namespace rtb
{
box_launch_fn find_box_kernel_f32(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode)
{
  RTB_BOX_CASE(float, 1, 8)
  RTB_BOX_CASE(float, 2, 8)
  RTB_BOX_CASE(float, 3, 8)
  RTB_BOX_CASE(float, 4, 8)
  RTB_BOX_CASE(float, 8, 8)
  RTB_BOX_CASE(float, 2, 16)
  RTB_BOX_CASE(float, 3, 16)
  RTB_BOX_CASE(float, 8, 16)
  return nullptr;
}
} // namespace rtb
