// tests/shim/fixed_random_device.h -- force-included (-include) when the reference's example programs are
// compiled for tests/test_reference_unit_tests.py: they seed std::mt19937 from std::random_device, and a
// byte-for-byte comparison of their output under two header sets needs the same seed in both builds.
// Test infrastructure only.
#pragma once
#include <random>
namespace std
{
struct rtb_fixed_random_device
{
  typedef unsigned int result_type;
  result_type operator()() { return 20240611u; }
};
} // namespace std
#define random_device rtb_fixed_random_device
