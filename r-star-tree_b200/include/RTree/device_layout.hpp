// RTree/device_layout.hpp -- host side (header-only C++17) of rtree-b200.
//
// flatten_device(): the step the reference leaves to the user
// (reference README.md:265-274, "upload the buffers yourself").  It takes a
// flatten_result_t -- ours or the reference's, the function is duck-typed on the
// field names of reference include/RTree/rtree.hpp:841-870 -- and re-lays it out for
// the GPU kernels:
//
//   * nodes renumbered breadth-first, so a level is one contiguous range, siblings
//     are adjacent and the top of the tree is a small prefix of the buffer;
//   * one fixed-width block per node, W = MAX_ENTRIES rounded up to 8 or 16 slots; a
//     slot's record is  lo_0 .. lo_{D-1} link hi_0 .. hi_{D-1}  (32-bit words), stored as a
//     structure of arrays of 16-byte vectors: row k holds words 4k..4k+3 of all W slots,
//     so a warp lane reads "its" child with one coalesced 128-bit load per row;
//   * leaves of point-keyed trees keep one corner only:  p_0 .. p_{D-1} id;
//   * a child link is the child block's byte offset / 16, so the kernels never multiply;
//   * empty slots hold an inverted box and RT_INVALID_ID, so lanes need no size test.
// The exact layout is specified in include/rtree_b200.h (rt_tree_desc).
//
// The result is a POD descriptor (rt_tree_desc, include/rtree_b200.h) plus one
// contiguous 256-byte aligned host buffer; hand desc() to rt_upload().
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <rtree_b200.h>

namespace eh
{
namespace rtree
{

// Only the customisation point is needed (min_point / max_point / DIM / scalar_type), and only
// by declaration, so this header works with this library's RTree/geometry.hpp AND with the
// reference's own RTree/geometry_traits.hpp.
template <typename GeometryType>
struct geometry_traits;

/// what a leaf slot reports as the "id" of a hit
enum class id_source
{
  data_index,  ///< index into flatten_result_t::data (always possible)
  mapped_value ///< static_cast<uint32_t>(data[index]); integral mapped types only
};

namespace detail
{
template <typename Geometry, typename Key>
struct key_is_box : std::is_same<Geometry, Key>
{
};

template <typename S>
struct scalar_code;
template <>
struct scalar_code<float>
{
  constexpr static std::uint32_t value = RT_F32;
};
template <>
struct scalar_code<double>
{
  constexpr static std::uint32_t value = RT_F64;
};
template <>
struct scalar_code<std::int32_t>
{
  constexpr static std::uint32_t value = RT_I32;
};

/// 256-byte aligned, zero-initialised byte buffer
class aligned_bytes
{
  unsigned char* _p = nullptr;
  std::size_t _n = 0;

public:
  aligned_bytes() = default;
  explicit aligned_bytes(std::size_t n)
      : _n(n)
  {
    const std::size_t rounded = (n + 255) / 256 * 256;
    _p = static_cast<unsigned char*>(std::aligned_alloc(256, rounded ? rounded : 256));
    if (_p == nullptr)
      throw std::bad_alloc();
    std::memset(_p, 0, rounded ? rounded : 256);
  }
  aligned_bytes(aligned_bytes&& rhs) noexcept
      : _p(rhs._p)
      , _n(rhs._n)
  {
    rhs._p = nullptr;
    rhs._n = 0;
  }
  aligned_bytes& operator=(aligned_bytes&& rhs) noexcept
  {
    if (this != &rhs)
    {
      std::free(_p);
      _p = rhs._p;
      _n = rhs._n;
      rhs._p = nullptr;
      rhs._n = 0;
    }
    return *this;
  }
  aligned_bytes(aligned_bytes const&) = delete;
  aligned_bytes& operator=(aligned_bytes const&) = delete;
  ~aligned_bytes() { std::free(_p); }
  unsigned char* data() { return _p; }
  unsigned char const* data() const { return _p; }
  std::size_t size() const { return _n; }
};

template <typename Mapped>
inline typename std::enable_if<std::is_integral<Mapped>::value && sizeof(Mapped) <= 4, std::uint32_t>::type
mapped_as_id(Mapped const& v)
{
  return static_cast<std::uint32_t>(v);
}
template <typename Mapped>
inline typename std::enable_if<!(std::is_integral<Mapped>::value && sizeof(Mapped) <= 4), std::uint32_t>::type
mapped_as_id(Mapped const&)
{
  throw std::invalid_argument("id_source::mapped_value needs an integral mapped type of <= 32 bits");
}
} // namespace detail

/// block width for a node capacity: 8 or 16 slots
constexpr inline std::uint32_t device_block_width(std::uint32_t max_entries)
{
  return max_entries <= 8 ? 8u : 16u;
}

/// The tree image flatten_device() produces.  Movable, not copyable.
struct device_tree_t
{
  std::uint32_t dim = 0;
  std::uint32_t scalar = 0;
  std::uint32_t width = 0;
  std::uint32_t max_entries = 0;
  std::uint32_t key_kind = 0;
  std::uint32_t leaf_level = 0;
  std::uint32_t num_values = 0;
  std::uint32_t internal_stride = 0;
  std::uint32_t leaf_stride = 0;
  std::uint64_t leaf_offset = 0;
  /// first node of every level, plus the total: [leaf_level + 2]
  std::vector<std::uint32_t> level_node_begin;
  /// new (breadth-first) index -> index in flatten_result_t::nodes
  std::vector<std::uint32_t> bfs_to_dfs;
  detail::aligned_bytes blocks;

  std::uint32_t num_levels() const { return leaf_level + 1; }
  std::uint32_t num_nodes() const { return level_node_begin.empty() ? 0 : level_node_begin.back(); }

  /// descriptor for rt_upload(); borrows this object's buffers
  rt_tree_desc desc() const
  {
    rt_tree_desc d;
    std::memset(&d, 0, sizeof d);
    d.abi_version = RT_ABI_VERSION;
    d.dim = dim;
    d.scalar = scalar;
    d.width = width;
    d.max_entries = max_entries;
    d.key_kind = key_kind;
    d.leaf_level = leaf_level;
    d.num_levels = num_levels();
    d.num_nodes = num_nodes();
    d.num_values = num_values;
    d.internal_stride = internal_stride;
    d.leaf_stride = leaf_stride;
    d.leaf_offset = leaf_offset;
    d.blocks_bytes = blocks.size();
    d.blocks = blocks.data();
    d.level_node_begin = level_node_begin.data();
    d.blocks_memory = RT_MEM_HOST;
    return d;
  }
};

/// Re-lay a flatten_result_t (field names of reference rtree.hpp:841-870) as device
/// blocks.  MaxEntries / KeyIsBox describe the tree that was flattened.
template <std::uint32_t MaxEntries, bool KeyIsBox, typename FlattenResult>
device_tree_t flatten_device(FlattenResult const& flat, id_source ids = id_source::data_index)
{
  using geometry_type = typename std::decay<decltype(flat.children_bound[0])>::type;
  using gtraits = geometry_traits<geometry_type>;
  using scalar = typename gtraits::scalar_type;
  constexpr std::uint32_t D = static_cast<std::uint32_t>(gtraits::DIM);
  constexpr std::uint32_t W = device_block_width(MaxEntries);
  static_assert(MaxEntries <= 16, "device blocks are at most 16 slots wide");
  static_assert(D >= 1 && D <= 8, "device blocks support 1..8 dimensions");

  if (flat.nodes.empty())
    throw std::invalid_argument("flatten_device: empty flatten_result_t (no root node)");
  if (flat.children.size() >= 0xFFFFFFFFull || flat.data.size() >= 0xFFFFFFFFull)
    throw std::length_error("flatten_device: more than 2^32-2 entries");

  device_tree_t out;
  out.dim = D;
  out.scalar = detail::scalar_code<scalar>::value;
  out.width = W;
  out.max_entries = MaxEntries;
  out.key_kind = KeyIsBox ? RT_KEY_BOX : RT_KEY_POINT;
  out.leaf_level = flat.leaf_level;
  out.num_values = static_cast<std::uint32_t>(flat.data.size());
  const std::uint32_t internal_words = rt_record_words(D, out.scalar, 1);
  const std::uint32_t leaf_words = rt_record_words(D, out.scalar, KeyIsBox ? 1 : 0);
  out.internal_stride = rt_block_bytes(W, internal_words);
  out.leaf_stride = rt_block_bytes(W, leaf_words);

  // ---- breadth-first renumbering (levels are implicit in the flat form: a node is a
  //      leaf iff it is leaf_level steps below node 0, reference test/rtree.cpp:383)
  const std::uint32_t n_nodes = static_cast<std::uint32_t>(flat.nodes.size());
  std::vector<std::uint32_t>& order = out.bfs_to_dfs;
  order.reserve(n_nodes);
  order.push_back(flat.root);
  out.level_node_begin.push_back(0);
  for (std::uint32_t level = 0; level < flat.leaf_level; ++level)
  {
    const std::uint32_t begin = out.level_node_begin[level];
    const std::uint32_t end = static_cast<std::uint32_t>(order.size());
    out.level_node_begin.push_back(end);
    for (std::uint32_t b = begin; b < end; ++b)
    {
      auto const& node = flat.nodes[order[b]];
      if (node.size > MaxEntries)
        throw std::invalid_argument("flatten_device: node wider than MAX_ENTRIES");
      for (std::uint32_t c = 0; c < node.size; ++c)
        order.push_back(flat.children[node.offset + c]);
    }
  }
  out.level_node_begin.push_back(static_cast<std::uint32_t>(order.size()));
  if (order.size() != n_nodes)
    throw std::invalid_argument("flatten_device: nodes unreachable from the root");

  const std::uint32_t leaf_begin = out.level_node_begin[flat.leaf_level];
  const std::uint32_t n_leaves = n_nodes - leaf_begin;
  const std::uint64_t internal_bytes = std::uint64_t(leaf_begin) * out.internal_stride;
  out.leaf_offset = (internal_bytes + 127) / 128 * 128;
  out.blocks = detail::aligned_bytes(out.leaf_offset + std::uint64_t(n_leaves) * out.leaf_stride);

  if ((out.leaf_offset + std::uint64_t(n_leaves) * out.leaf_stride) / 16 >= RT_INVALID_ID)
    throw std::length_error("flatten_device: tree image larger than 64 GB");

  const scalar empty_lo = std::numeric_limits<scalar>::has_infinity
                              ? std::numeric_limits<scalar>::infinity()
                              : std::numeric_limits<scalar>::max();
  const scalar empty_hi = std::numeric_limits<scalar>::has_infinity
                              ? -std::numeric_limits<scalar>::infinity()
                              : std::numeric_limits<scalar>::lowest();
  constexpr std::uint32_t SW = sizeof(scalar) / 4; // words per scalar

  // word j of slot c of a block with `words`-word records (see include/rtree_b200.h)
  auto word_at = [](unsigned char* block, std::uint32_t words, std::uint32_t c, std::uint32_t j) -> std::uint32_t*
  {
    static const std::uint32_t tail_bytes[4] = { 0u, 4u, 8u, 16u };
    const std::uint32_t full = words / 4;
    if (j < 4 * full)
      return reinterpret_cast<std::uint32_t*>(block + (j / 4) * W * 16 + c * 16 + (j % 4) * 4);
    return reinterpret_cast<std::uint32_t*>(block + full * W * 16 + c * tail_bytes[words % 4] + (j - 4 * full) * 4);
  };
  auto put_scalar = [&](unsigned char* block, std::uint32_t words, std::uint32_t c, std::uint32_t j, scalar v)
  {
    std::uint32_t raw[SW];
    std::memcpy(raw, &v, sizeof v);
    for (std::uint32_t k = 0; k < SW; ++k)
      *word_at(block, words, c, j + k) = raw[k];
  };

  // children of consecutive breadth-first nodes are consecutive breadth-first nodes
  std::uint32_t next_child = 1;
  unsigned char* base = out.blocks.data();
  for (std::uint32_t b = 0; b < n_nodes; ++b)
  {
    auto const& node = flat.nodes[order[b]];
    const bool is_leaf = b >= leaf_begin;
    unsigned char* block = is_leaf ? base + out.leaf_offset + std::uint64_t(b - leaf_begin) * out.leaf_stride
                                   : base + std::uint64_t(b) * out.internal_stride;
    const bool one_corner = is_leaf && !KeyIsBox;
    const std::uint32_t words = is_leaf ? leaf_words : internal_words;
    if (node.size > MaxEntries)
      throw std::invalid_argument("flatten_device: node wider than MAX_ENTRIES");
    for (std::uint32_t c = 0; c < W; ++c)
    {
      const bool used = c < node.size;
      std::uint32_t link = RT_INVALID_ID;
      for (std::uint32_t d = 0; d < D; ++d)
      {
        scalar lo = empty_lo, hi = empty_hi;
        if (used)
        {
          geometry_type const& g = flat.children_bound[node.offset + c];
          lo = gtraits::min_point(g, static_cast<int>(d));
          hi = gtraits::max_point(g, static_cast<int>(d));
        }
        // record: lo corner, link (one word, plus an alignment word for 64-bit scalars), hi corner
        put_scalar(block, words, c, d * SW, lo);
        if (!one_corner)
          put_scalar(block, words, c, (D + 1 + d) * SW, hi);
      }
      if (used)
      {
        if (is_leaf)
        {
          const std::uint32_t data_index = flat.children[node.offset + c];
          link = (ids == id_source::mapped_value) ? detail::mapped_as_id(flat.data[data_index]) : data_index;
        }
        else
        {
          const std::uint32_t child = next_child++;
          const std::uint64_t child_offset
              = child >= leaf_begin ? out.leaf_offset + std::uint64_t(child - leaf_begin) * out.leaf_stride
                                    : std::uint64_t(child) * out.internal_stride;
          link = static_cast<std::uint32_t>(child_offset / 16);
        }
      }
      *word_at(block, words, c, D * SW) = link;
    }
  }
  return out;
}

// ---- on-disk form of the image --------------------------------------------------------------
// A tree built once on the CPU (minutes for 10^8 keys) can be saved and reloaded without the
// insert loop.  Little-endian, one file:
//   header (128 bytes)  magic "RTB200IM", format version, the rt_tree_desc scalars, section sizes,
//                       FNV-1a-64 checksum of everything after the header
//   level_node_begin    (num_levels + 1) x u32
//   bfs_to_dfs          num_nodes x u32
//   padding to a multiple of 256 bytes, then the blocks exactly as rt_upload() takes them
namespace detail
{
struct image_file_header
{
  char magic[8];
  std::uint32_t format_version;
  std::uint32_t header_bytes;
  std::uint32_t abi_version, dim, scalar, width, max_entries, key_kind, leaf_level, num_levels, num_nodes,
      num_values, internal_stride, leaf_stride;
  std::uint64_t leaf_offset;
  std::uint64_t blocks_bytes;
  std::uint64_t blocks_file_offset;
  std::uint64_t checksum;
  unsigned char reserved[128 - 8 - 4 * 14 - 8 * 4];
};
static_assert(sizeof(image_file_header) == 128, "image_file_header layout");

constexpr std::uint32_t IMAGE_FORMAT_VERSION = 1;

inline std::uint64_t fnv1a64(const void* data, std::size_t n, std::uint64_t h = 0xcbf29ce484222325ull)
{
  // 8 bytes at a time (a byte-wise FNV over gigabytes would dominate the load)
  const unsigned char* p = static_cast<const unsigned char*>(data);
  std::size_t i = 0;
  for (; i + 8 <= n; i += 8)
  {
    std::uint64_t w;
    std::memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x100000001b3ull;
  }
  for (; i < n; ++i)
    h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

struct file_closer
{
  std::FILE* f;
  ~file_closer()
  {
    if (f)
      std::fclose(f);
  }
};
} // namespace detail

/// write `image` to `path`; throws std::runtime_error on I/O failure
inline void save_device_tree(device_tree_t const& image, std::string const& path)
{
  detail::image_file_header h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "RTB200IM", 8);
  h.format_version = detail::IMAGE_FORMAT_VERSION;
  h.header_bytes = sizeof h;
  const rt_tree_desc d = image.desc();
  h.abi_version = d.abi_version, h.dim = d.dim, h.scalar = d.scalar, h.width = d.width;
  h.max_entries = d.max_entries, h.key_kind = d.key_kind, h.leaf_level = d.leaf_level;
  h.num_levels = d.num_levels, h.num_nodes = d.num_nodes, h.num_values = d.num_values;
  h.internal_stride = d.internal_stride, h.leaf_stride = d.leaf_stride;
  h.leaf_offset = d.leaf_offset;
  h.blocks_bytes = d.blocks_bytes;
  if (image.level_node_begin.size() != std::size_t(h.num_levels) + 1 || image.bfs_to_dfs.size() != h.num_nodes)
    throw std::runtime_error("save_device_tree: inconsistent image");
  const std::uint64_t tables = 4ull * (image.level_node_begin.size() + image.bfs_to_dfs.size());
  h.blocks_file_offset = (sizeof h + tables + 255) / 256 * 256;
  std::vector<unsigned char> pad(std::size_t(h.blocks_file_offset - sizeof h - tables), 0);
  std::uint64_t sum = detail::fnv1a64(image.level_node_begin.data(), 4 * image.level_node_begin.size());
  sum = detail::fnv1a64(image.bfs_to_dfs.data(), 4 * image.bfs_to_dfs.size(), sum);
  sum = detail::fnv1a64(pad.data(), pad.size(), sum);
  sum = detail::fnv1a64(image.blocks.data(), image.blocks.size(), sum);
  h.checksum = sum;

  detail::file_closer file { std::fopen(path.c_str(), "wb") };
  if (file.f == nullptr)
    throw std::runtime_error("save_device_tree: cannot open " + path);
  auto put = [&](const void* p, std::size_t n)
  {
    if (n && std::fwrite(p, 1, n, file.f) != n)
      throw std::runtime_error("save_device_tree: short write to " + path);
  };
  put(&h, sizeof h);
  put(image.level_node_begin.data(), 4 * image.level_node_begin.size());
  put(image.bfs_to_dfs.data(), 4 * image.bfs_to_dfs.size());
  put(pad.data(), pad.size());
  put(image.blocks.data(), image.blocks.size());
  if (std::fflush(file.f) != 0)
    throw std::runtime_error("save_device_tree: flush failed for " + path);
}

/// read an image written by save_device_tree(); throws std::runtime_error if the file is not a
/// complete, uncorrupted image of a format this header understands
inline device_tree_t load_device_tree(std::string const& path)
{
  detail::file_closer file { std::fopen(path.c_str(), "rb") };
  if (file.f == nullptr)
    throw std::runtime_error("load_device_tree: cannot open " + path);
  auto get = [&](void* p, std::size_t n, const char* what)
  {
    if (n && std::fread(p, 1, n, file.f) != n)
      throw std::runtime_error(std::string("load_device_tree: file truncated in ") + what);
  };
  detail::image_file_header h;
  get(&h, sizeof h, "header");
  if (std::memcmp(h.magic, "RTB200IM", 8) != 0)
    throw std::runtime_error("load_device_tree: not an rtree-b200 image (bad magic)");
  if (h.format_version != detail::IMAGE_FORMAT_VERSION || h.header_bytes != sizeof h)
    throw std::runtime_error("load_device_tree: unsupported image format version");
  if (h.abi_version != RT_ABI_VERSION)
    throw std::runtime_error("load_device_tree: image written for another ABI version");
  if (h.num_levels != h.leaf_level + 1 || h.num_levels > 64 || h.num_nodes == 0 || h.dim < 1 || h.dim > 8
      || h.blocks_bytes > (std::uint64_t(1) << 40))
    throw std::runtime_error("load_device_tree: implausible header");
  const std::uint64_t tables = 4ull * (std::uint64_t(h.num_levels) + 1 + h.num_nodes);
  if (h.blocks_file_offset != (sizeof h + tables + 255) / 256 * 256)
    throw std::runtime_error("load_device_tree: implausible header");

  device_tree_t out;
  out.dim = h.dim, out.scalar = h.scalar, out.width = h.width, out.max_entries = h.max_entries;
  out.key_kind = h.key_kind, out.leaf_level = h.leaf_level, out.num_values = h.num_values;
  out.internal_stride = h.internal_stride, out.leaf_stride = h.leaf_stride, out.leaf_offset = h.leaf_offset;
  out.level_node_begin.resize(std::size_t(h.num_levels) + 1);
  out.bfs_to_dfs.resize(h.num_nodes);
  get(out.level_node_begin.data(), 4 * out.level_node_begin.size(), "level table");
  get(out.bfs_to_dfs.data(), 4 * out.bfs_to_dfs.size(), "node table");
  std::vector<unsigned char> pad(std::size_t(h.blocks_file_offset - sizeof h - tables), 0);
  get(pad.data(), pad.size(), "padding");
  out.blocks = detail::aligned_bytes(h.blocks_bytes);
  get(out.blocks.data(), out.blocks.size(), "blocks");
  unsigned char extra;
  if (std::fread(&extra, 1, 1, file.f) != 0)
    throw std::runtime_error("load_device_tree: trailing bytes after the blocks");
  std::uint64_t sum = detail::fnv1a64(out.level_node_begin.data(), 4 * out.level_node_begin.size());
  sum = detail::fnv1a64(out.bfs_to_dfs.data(), 4 * out.bfs_to_dfs.size(), sum);
  sum = detail::fnv1a64(pad.data(), pad.size(), sum);
  sum = detail::fnv1a64(out.blocks.data(), out.blocks.size(), sum);
  if (sum != h.checksum)
    throw std::runtime_error("load_device_tree: checksum mismatch (corrupted image)");
  if (out.level_node_begin.back() != h.num_nodes)
    throw std::runtime_error("load_device_tree: level table inconsistent with the header");
  return out;
}

} // namespace rtree
} // namespace eh
