// tests/shim/gtest/gtest.h -- a few dozen lines standing in for GoogleTest (not in this image), so that the
// reference's OWN unit tests (/root/reference/test/rtree.cpp, compiled where they lie, never copied) can be
// built against the host library of this repo: tests/test_reference_unit_tests.py.
// Test infrastructure only.  Supports what that file uses: TEST, ASSERT_* / EXPECT_* with a trailing
// `<< message`, testing::InitGoogleTest, RUN_ALL_TESTS.  An ASSERT_* that fails returns from the test body.
#pragma once
#include <cstdio>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing
{
struct TestCase
{
  const char* suite;
  const char* name;
  void (*body)();
};
inline std::vector<TestCase>& registry()
{
  static std::vector<TestCase> all;
  return all;
}
inline int& failures()
{
  static int n = 0;
  return n;
}
struct Registrar
{
  Registrar(const char* suite, const char* name, void (*body)()) { registry().push_back({ suite, name, body }); }
};
// collects the `<< message` of a failed check and reports when it goes out of scope
struct Failure
{
  std::ostringstream text;
  Failure(const char* file, int line, const std::string& what)
  {
    ++failures();
    text << file << ":" << line << ": Failure\n  " << what << "\n  ";
  }
  ~Failure() { std::cerr << text.str() << std::endl; }
  template <typename T>
  Failure& operator<<(const T& v)
  {
    text << v;
    return *this;
  }
};
// `return Void() = Failure(..) << ...;` lets an ASSERT leave a void function after streaming
struct Void
{
  void operator=(const Failure&) const {}
};
inline void InitGoogleTest(int*, char**) {}
inline int run_all()
{
  int failed_tests = 0;
  for (const TestCase& t : registry())
  {
    const int before = failures();
    std::printf("[ RUN      ] %s.%s\n", t.suite, t.name);
    std::fflush(stdout);
    t.body();
    const bool ok = failures() == before;
    std::printf("[ %s ] %s.%s\n", ok ? "      OK" : " FAILED ", t.suite, t.name);
    failed_tests += ok ? 0 : 1;
  }
  std::printf("[==========] %zu tests ran, %d failed\n", registry().size(), failed_tests);
  return failed_tests == 0 ? 0 : 1;
}
} // namespace testing

#define RUN_ALL_TESTS() ::testing::run_all()
#define TEST(suite, name)                                                                   \
  static void suite##_##name##_body();                                                      \
  static ::testing::Registrar suite##_##name##_registrar(#suite, #name, &suite##_##name##_body); \
  static void suite##_##name##_body()

#define SHIM_CHECK_(cond, what, on_fail) \
  if (cond)                              \
    ;                                    \
  else                                   \
    on_fail ::testing::Failure(__FILE__, __LINE__, what)
#define SHIM_FATAL_ return ::testing::Void() =
#define SHIM_NONFATAL_ (void)

#define ASSERT_TRUE(c) SHIM_CHECK_((c), "expected true: " #c, SHIM_FATAL_)
#define ASSERT_FALSE(c) SHIM_CHECK_(!(c), "expected false: " #c, SHIM_FATAL_)
#define ASSERT_EQ(a, b) SHIM_CHECK_((a) == (b), "expected " #a " == " #b, SHIM_FATAL_)
#define ASSERT_NE(a, b) SHIM_CHECK_((a) != (b), "expected " #a " != " #b, SHIM_FATAL_)
#define ASSERT_LT(a, b) SHIM_CHECK_((a) < (b), "expected " #a " < " #b, SHIM_FATAL_)
#define ASSERT_LE(a, b) SHIM_CHECK_((a) <= (b), "expected " #a " <= " #b, SHIM_FATAL_)
#define ASSERT_GT(a, b) SHIM_CHECK_((a) > (b), "expected " #a " > " #b, SHIM_FATAL_)
#define ASSERT_GE(a, b) SHIM_CHECK_((a) >= (b), "expected " #a " >= " #b, SHIM_FATAL_)
#define EXPECT_TRUE(c) SHIM_CHECK_((c), "expected true: " #c, SHIM_NONFATAL_)
#define EXPECT_FALSE(c) SHIM_CHECK_(!(c), "expected false: " #c, SHIM_NONFATAL_)
#define EXPECT_EQ(a, b) SHIM_CHECK_((a) == (b), "expected " #a " == " #b, SHIM_NONFATAL_)
#define EXPECT_NE(a, b) SHIM_CHECK_((a) != (b), "expected " #a " != " #b, SHIM_NONFATAL_)
#define EXPECT_LT(a, b) SHIM_CHECK_((a) < (b), "expected " #a " < " #b, SHIM_NONFATAL_)
#define EXPECT_LE(a, b) SHIM_CHECK_((a) <= (b), "expected " #a " <= " #b, SHIM_NONFATAL_)
#define EXPECT_GT(a, b) SHIM_CHECK_((a) > (b), "expected " #a " > " #b, SHIM_NONFATAL_)
#define EXPECT_GE(a, b) SHIM_CHECK_((a) >= (b), "expected " #a " >= " #b, SHIM_NONFATAL_)
