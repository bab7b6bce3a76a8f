// pcie_probe.cu -- what bounds the end-to-end numbers at 4 / 8 GPUs: the host side of the D2H copies.
//
// Every GPU returns its own CSR (763 MB per 16 M queries) to pinned host memory; round 1 measured
// about 72 GB/s of D2H for ALL GPUs of a box together, whatever their number.  This probe asks
// whether the host allocation changes that ceiling.  One process, one thread and two streams per
// GPU; G GPUs copy at once, for G = 1, 2, 4, ... up to the GPUs present:
//
//   hostalloc     cudaHostAlloc(default)                       what the engine uses (HostBuf)
//   hostalloc-wc  cudaHostAlloc(write-combined)                no CPU-cache snoop on the DMA writes
//   thp-register  mmap + MADV_HUGEPAGE + touch + cudaHostRegister   2 MB pages behind the IOMMU
//   hugetlb-reg   mmap(MAP_HUGETLB) + cudaHostRegister         (only if the box has hugetlb pages)
//   malloc-4k-reg posix_memalign + MADV_NOHUGEPAGE + cudaHostRegister   4 KB pages, for contrast
//
// and, per variant, D2H alone with one and with two concurrent copies per GPU, H2D alone, and both
// directions together (the shape of a pipelined rt_query_boxes call: 1 byte in per 2 bytes out).
//
//   nvcc -O2 -std=c++17 -o /tmp/pcie_probe scripts/pcie_probe.cu && /tmp/pcie_probe
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <sys/mman.h>

#define CK(x)                                                                                  \
  do                                                                                           \
  {                                                                                            \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess)                                                                     \
    {                                                                                          \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                            \
      std::exit(1);                                                                            \
    }                                                                                          \
  } while (0)

static const size_t OUT_BYTES = 768ull << 20, IN_BYTES = 384ull << 20;

struct HostBlock
{
  void* p = nullptr;
  size_t bytes = 0;
  int how = 0; // 0 cudaHostAlloc, 1 mmap + register, 2 posix_memalign + register
};

static bool host_alloc(const std::string& kind, size_t bytes, HostBlock& b)
{
  b.bytes = bytes;
  if (kind == "hostalloc" || kind == "hostalloc-wc")
  {
    b.how = 0;
    return cudaHostAlloc(&b.p, bytes, kind == "hostalloc" ? cudaHostAllocPortable
                                                          : (cudaHostAllocPortable | cudaHostAllocWriteCombined))
           == cudaSuccess;
  }
  if (kind == "thp-register" || kind == "hugetlb-reg")
  {
    b.how = 1;
    const int flags = MAP_PRIVATE | MAP_ANONYMOUS | (kind == "hugetlb-reg" ? MAP_HUGETLB : 0);
    b.p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, flags, -1, 0);
    if (b.p == MAP_FAILED)
    {
      b.p = nullptr;
      return false;
    }
    if (kind == "thp-register")
      madvise(b.p, bytes, MADV_HUGEPAGE);
    std::memset(b.p, 1, bytes);
    if (cudaHostRegister(b.p, bytes, cudaHostRegisterPortable) != cudaSuccess)
    {
      cudaGetLastError();
      munmap(b.p, bytes);
      b.p = nullptr;
      return false;
    }
    return true;
  }
  b.how = 2;
  if (posix_memalign(&b.p, 4096, bytes) != 0)
    return false;
  madvise(b.p, bytes, MADV_NOHUGEPAGE);
  std::memset(b.p, 1, bytes);
  if (cudaHostRegister(b.p, bytes, cudaHostRegisterPortable) != cudaSuccess)
  {
    cudaGetLastError();
    std::free(b.p);
    b.p = nullptr;
    return false;
  }
  return true;
}

static void host_free(HostBlock& b)
{
  if (!b.p)
    return;
  if (b.how == 0)
    cudaFreeHost(b.p);
  else
  {
    cudaHostUnregister(b.p);
    if (b.how == 1)
      munmap(b.p, b.bytes);
    else
      std::free(b.p);
  }
  b.p = nullptr;
}

struct Gpu
{
  int dev;
  void *d_out, *d_in;
  cudaStream_t s[3];
  HostBlock h_out, h_in;
};

// what: 0 = D2H one copy, 1 = D2H as two halves on two streams, 2 = H2D, 3 = both directions
static double round_ms(std::vector<Gpu>& g, int G, int what, int rounds)
{
  for (int i = 0; i < G; ++i)
  {
    CK(cudaSetDevice(g[i].dev));
    CK(cudaDeviceSynchronize());
  }
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int i = 0; i < G; ++i)
  {
    th.emplace_back([&, i]() {
      Gpu& x = g[i];
      CK(cudaSetDevice(x.dev));
      for (int r = 0; r < rounds; ++r)
      {
        if (what == 0 || what == 3)
          CK(cudaMemcpyAsync(x.h_out.p, x.d_out, OUT_BYTES, cudaMemcpyDeviceToHost, x.s[0]));
        if (what == 1)
        {
          CK(cudaMemcpyAsync(x.h_out.p, x.d_out, OUT_BYTES / 2, cudaMemcpyDeviceToHost, x.s[0]));
          CK(cudaMemcpyAsync((char*)x.h_out.p + OUT_BYTES / 2, (char*)x.d_out + OUT_BYTES / 2, OUT_BYTES / 2,
                             cudaMemcpyDeviceToHost, x.s[1]));
        }
        if (what == 2 || what == 3)
          CK(cudaMemcpyAsync(x.d_in, x.h_in.p, IN_BYTES, cudaMemcpyHostToDevice, x.s[2]));
      }
      CK(cudaDeviceSynchronize());
    });
  }
  for (auto& t : th)
    t.join();
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return ms / rounds;
}

int main()
{
  int n_dev = 0;
  CK(cudaGetDeviceCount(&n_dev));
  std::vector<Gpu> g(n_dev);
  for (int i = 0; i < n_dev; ++i)
  {
    g[i].dev = i;
    CK(cudaSetDevice(i));
    CK(cudaMalloc(&g[i].d_out, OUT_BYTES));
    CK(cudaMalloc(&g[i].d_in, IN_BYTES));
    CK(cudaMemset(g[i].d_out, 1, OUT_BYTES));
    for (auto& s : g[i].s)
      CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  }
  // hugetlb pages, if the box lets root reserve them (ignored when it does not)
  if (FILE* f = std::fopen("/proc/sys/vm/nr_hugepages", "w"))
  {
    std::fprintf(f, "%d\n", (int)((OUT_BYTES + IN_BYTES) / (2 << 20)) * n_dev + 64);
    std::fclose(f);
  }
  const char* kinds[] = { "hostalloc", "hostalloc-wc", "thp-register", "hugetlb-reg", "malloc-4k-reg" };
  const char* names[] = { "D2H", "D2H two streams", "H2D", "H2D + D2H together" };
  std::printf("%d GPU(s); %zu MB out + %zu MB in per GPU and round\n", n_dev, OUT_BYTES >> 20, IN_BYTES >> 20);
  for (const char* kind : kinds)
  {
    bool ok = true;
    for (int i = 0; i < n_dev && ok; ++i)
    {
      CK(cudaSetDevice(i));
      ok = host_alloc(kind, OUT_BYTES, g[i].h_out) && host_alloc(kind, IN_BYTES, g[i].h_in);
    }
    if (!ok)
    {
      std::printf("%-14s: allocation not available on this box\n", kind);
      for (int i = 0; i < n_dev; ++i)
      {
        host_free(g[i].h_out);
        host_free(g[i].h_in);
      }
      continue;
    }
    for (int G = 1; G <= n_dev; G *= 2)
    {
      for (int what = 0; what < 4; ++what)
      {
        round_ms(g, G, what, 2);
        const double ms = round_ms(g, G, what, 5);
        const double gb = ((what == 2 ? 0.0 : (double)OUT_BYTES) + (what >= 2 ? (double)IN_BYTES : 0.0)) / 1e9;
        std::printf("%-14s %d GPU(s) %-20s %7.2f ms per round -> %6.1f GB/s per GPU, %6.1f GB/s aggregate\n", kind, G,
                    names[what], ms, gb / (ms * 1e-3), gb * G / (ms * 1e-3));
        std::fflush(stdout);
      }
    }
    for (int i = 0; i < n_dev; ++i)
    {
      host_free(g[i].h_out);
      host_free(g[i].h_in);
    }
  }
  return 0;
}
