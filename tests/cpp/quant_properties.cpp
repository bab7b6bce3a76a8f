// tests/cpp/quant_properties.cpp -- the quantiser of r-star-tree_b200/csrc/rt_quant.cuh, compiled for the HOST
// from the very header the kernels include, and checked for the properties DESIGN.md 3.1 argues from:
//   1. q is monotone non-decreasing, for every grid the builder can produce and every float (incl. +-inf,
//      denormals, values far outside the grid); NaN lower bounds go to the bottom, NaN upper bounds to the top
//   2. conservative filter: closed overlap of a bound with a query (reference example/sample/sample.cpp:48-54)
//      implies quant_pass on their quantised forms; point-in-box (aabb.hpp:206-210) implies quant_pass
//   3. quant_strictly_inside implies the reference's is_inside -- qlo < p < qhi on every axis whose query bounds are
//      numbers -- so such a candidate needs no float look-up
//   4. an empty slot (a = 0 in every half) never passes
//   5. the packed 8-D leaf: 0x80018000 - (G | s) is G | n in both halves, i.e. the record the traversal derives
//      equals the record quant_slot would have stored for the point
// The device intrinsics the header uses are given their host meaning below (-ffp-contract=off: a*b and a-b round
// once, like __fmul_rn / __fsub_rn).  Prints "ok <count>" per property; exits non-zero on the first violation.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }

#include "rt_quant.cuh"

using namespace rtb;

static float bits_float(uint32_t b)
{
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}

#define REQUIRE(cond, ...)                          \
  do                                                \
  {                                                 \
    if (!(cond))                                    \
    {                                               \
      std::printf("VIOLATION %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                     \
      std::printf("\n");                            \
      std::exit(1);                                 \
    }                                               \
  } while (0)

struct Axis
{
  float lo, scale;
};

// the grid of one axis as quant_grid_kernel (rt_quant.cu) derives it from the extent [l, h] of the keys
static Axis make_axis(float l, float h)
{
  const float extent = h - l;
  const bool usable = extent > 0.0f && extent < std::numeric_limits<float>::infinity();
  return { usable ? l : 0.0f, usable ? 32766.0f / extent : 0.0f };
}

static std::mt19937_64 rng(20241020);

// floats that hurt: around the grid's ends, cell boundaries, huge, tiny, infinite
static float nasty(const Axis& g, float l, float h)
{
  std::uniform_int_distribution<int> pick(0, 11);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  std::uniform_int_distribution<int> cell(0, 32767), ulps(-3, 3);
  float v;
  switch (pick(rng))
  {
  case 0: v = l; break;
  case 1: v = h; break;
  case 2: v = l + (h - l) * u(rng); break;
  case 3: v = l - (h - l) * u(rng) * 2.f; break;
  case 4: v = h + (h - l) * u(rng) * 2.f; break;
  case 5: v = g.scale > 0 ? g.lo + (float)cell(rng) / g.scale : l; break; // a cell boundary
  case 6: v = std::numeric_limits<float>::infinity(); break;
  case 7: v = -std::numeric_limits<float>::infinity(); break;
  case 8: v = std::numeric_limits<float>::denorm_min() * (float)cell(rng); break;
  case 9: v = (u(rng) - 0.5f) * 3.0e38f; break;
  case 10: v = 0.0f; break;
  default: v = l + (h - l) * u(rng); break;
  }
  // and a few ulps beside it
  if (std::isfinite(v))
  {
    uint32_t b;
    std::memcpy(&b, &v, 4);
    const int d = ulps(rng);
    if ((b & 0x7FFFFFFFu) > 8u && (b & 0x7F800000u) != 0x7F800000u)
      b += (uint32_t)d;
    float w = bits_float(b);
    if (std::isfinite(w))
      v = w;
  }
  return v;
}

static void random_extent(float& l, float& h, bool ordinary)
{
  std::uniform_int_distribution<int> kind(0, 7);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  static const int plain[4] = { 0, 1, 2, 7 };
  switch (ordinary ? plain[kind(rng) & 3] : kind(rng))
  {
  case 0: l = 0.f, h = 1.f; break;                                   // the unit cube of config 2
  case 1: l = -0.1f - u(rng), h = 1.1f + u(rng); break;
  case 2: l = u(rng) * 1e12f, h = l + 1e12f * (1.f + u(rng)); break;   // 10^12 ranges
  case 3: l = 1.0f + u(rng), h = bits_float([&] { uint32_t b; std::memcpy(&b, &l, 4); return b + 5u; }()); break; // few ulps
  case 4: l = 0.f, h = std::numeric_limits<float>::denorm_min() * 1000.f; break;
  case 5: l = -3.0e38f, h = 3.0e38f; break;                         // extent overflows: scale 0
  case 6: l = h = u(rng); break;                                     // zero extent: scale 0
  default: l = -u(rng) * 100.f, h = u(rng) * 100.f + 1e-3f; break;
  }
}

template <int D>
static void run(long rounds)
{
  long monotone = 0, conservative = 0, strict = 0, empties = 0, packed = 0;
  for (long r = 0; r < rounds; ++r)
  {
    float gl[D], gh[D], glo[8], gscale[8];
    Axis ax[D];
    for (int d = 0; d < D; ++d)
    {
      random_extent(gl[d], gh[d], (r & 1) != 0); // every other grid has ordinary extents on all axes
      ax[d] = make_axis(gl[d], gh[d]);
      glo[d] = ax[d].lo, gscale[d] = ax[d].scale;
    }
    for (int rep = 0; rep < 64; ++rep)
    {
      // 1. monotone, per axis
      for (int d = 0; d < D; ++d)
      {
        float a = nasty(ax[d], gl[d], gh[d]), b = nasty(ax[d], gl[d], gh[d]);
        if (a > b)
          std::swap(a, b);
        const uint32_t qa = quant(a, glo[d], gscale[d]), qb = quant(b, glo[d], gscale[d]);
        REQUIRE(qa >= 1 && qa <= 32767 && qb >= 1 && qb <= 32767, "range: q(%a) = %u, q(%a) = %u", a, qa, b, qb);
        REQUIRE(qa <= qb, "not monotone: q(%a) = %u > q(%a) = %u (lo %a scale %a)", a, qa, b, qb, glo[d], gscale[d]);
        REQUIRE(quant_upper(b, glo[d], gscale[d]) == qb, "quant_upper differs on a number");
        ++monotone;
      }
      const float nan = std::numeric_limits<float>::quiet_NaN();
      REQUIRE(quant(nan, glo[0], gscale[0]) == 1u && quant_upper(nan, glo[0], gscale[0]) == 32767u, "NaN mapping");

      // 2. / 3. a node bound, a point and a query box
      float lo[D], hi[D], p[D], qlo[D], qhi[D];
      for (int d = 0; d < D; ++d)
      {
        lo[d] = nasty(ax[d], gl[d], gh[d]), hi[d] = nasty(ax[d], gl[d], gh[d]);
        if (lo[d] > hi[d])
          std::swap(lo[d], hi[d]);
        p[d] = nasty(ax[d], gl[d], gh[d]);
        qlo[d] = nasty(ax[d], gl[d], gh[d]), qhi[d] = nasty(ax[d], gl[d], gh[d]);
        if ((rng() & 7) != 0 && qlo[d] > qhi[d]) // mostly well-formed, sometimes inverted
          std::swap(qlo[d], qhi[d]);
        if ((rng() & 63) == 0)
          qlo[d] = nan;
        if ((rng() & 63) == 1)
          qhi[d] = nan;
        // make overlaps likely: now and then the query takes a corner of the bound, or the point itself
        if ((rng() & 3) == 0)
          qlo[d] = hi[d];
        if ((rng() & 3) == 1)
          qhi[d] = lo[d];
        if ((rng() & 3) == 2)
          qlo[d] = p[d];
      }
      // every other round the query is built around the point (a few cells, a few ulps, or a tenth of the extent
      // to each side), so that inside / strictly-inside cases are common in every dimension
      if ((rep & 1) != 0)
      {
        std::uniform_real_distribution<float> u(0.f, 1.f);
        const int how = (int)(rng() % 3);
        for (int d = 0; d < D; ++d)
        {
          if (!std::isfinite(p[d]) || rng() % 10 != 0)
            p[d] = gl[d] + (gh[d] - gl[d]) * u(rng); // mostly a key inside the grid, as real keys are
          const float cell = gscale[d] > 0 ? 1.0f / gscale[d] : std::fabs(p[d]) * 1e-6f;
          const float w = how == 0 ? cell * (0.25f + 4.f * u(rng))
                                   : (how == 1 ? std::fabs(p[d]) * 2.4e-7f * (float)(rng() % 4) : (gh[d] - gl[d]) * 0.1f * u(rng));
          qlo[d] = p[d] - w, qhi[d] = p[d] + w;
          lo[d] = p[d] - w * u(rng), hi[d] = p[d] + w * u(rng);
        }
      }
      uint32_t a[D], ap[D], b[D];
      quant_slot<D>(lo, hi, glo, gscale, a);
      quant_slot<D>(p, p, glo, gscale, ap);
      quant_query<D>(qlo, qhi, glo, gscale, b);
      bool overlap = true, inside = true;
      for (int d = 0; d < D; ++d)
      {
        overlap = overlap && !(hi[d] < qlo[d] || lo[d] > qhi[d]);
        inside = inside && !(qlo[d] > p[d]) && !(p[d] > qhi[d]);
      }
      if (overlap)
      {
        REQUIRE(quant_pass<D>(a, b), "closed overlap but the quantised test rejects");
        ++conservative;
      }
      if (inside)
      {
        REQUIRE(quant_pass<D>(ap, b), "point in box but the quantised test rejects");
        ++conservative;
      }
      if (quant_strictly_inside<D>(ap, b))
      {
        REQUIRE(quant_pass<D>(ap, b), "strictly inside without passing");
        // (a NaN query bound constrains nothing, in the reference's comparisons and on the grid alike: the
        // claim is the reference's test, and qlo < p < qhi on every axis whose bounds are numbers)
        REQUIRE(inside, "strictly inside on the grid but the reference's is_inside fails");
        for (int d = 0; d < D; ++d)
          REQUIRE((qlo[d] != qlo[d] || qlo[d] < p[d]) && (qhi[d] != qhi[d] || p[d] < qhi[d]),
                  "strictly inside on the grid but not qlo < p < qhi on axis %d: %a %a %a", d, qlo[d], p[d], qhi[d]);
        ++strict;
      }
      // 4. the empty slot
      uint32_t empty[D];
      for (int j = 0; j < D; ++j)
        empty[j] = QUANT_GUARD;
      REQUIRE(!quant_pass<D>(empty, b), "an empty slot passes");
      ++empties;
      // 5. packed leaves (D even: the q-words are the first D/2 words)
      if (D % 2 == 0)
      {
        for (int j = 0; j < D / 2; ++j)
        {
          const uint32_t derived = 0x80018000u - ap[j];
          REQUIRE(derived == ap[D / 2 + j], "packed leaf: derived %08x, stored %08x", derived, ap[D / 2 + j]);
          ++packed;
        }
      }
    }
  }
  std::printf("dim %d: ok monotone %ld, conservative %ld, strictly-inside %ld, empty %ld, packed %ld\n", D, monotone,
              conservative, strict, empties, packed);
  REQUIRE(conservative > 10 * rounds && strict > rounds / 100, "the generators did not produce enough overlaps / strictly-inside cases");
}

int main(int argc, char** argv)
{
  const long rounds = argc > 1 ? std::atol(argv[1]) : 20000;
  run<2>(rounds);
  run<3>(rounds);
  run<8>(rounds / 2);
  return 0;
}
