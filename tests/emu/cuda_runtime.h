// Host stand-in for <cuda_runtime.h> (TEST INFRASTRUCTURE): lets g++ compile the warp-level device headers of
// open_duck_playground_b200/csrc for a CPU emulation in which one warp is 32 host threads and every warp intrinsic is an
// exchange through a shared buffer between two barriers.  Valid for code whose *_sync intrinsics are reached by all 32
// lanes in the same order (true for the collision routines this is used for).  Never part of the product path.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define ODUCK_WARP_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)

struct alignas(16) float4 { float x, y, z, w; };

namespace warp_emu {
inline std::barrier<>* bar = nullptr;          // one warp at a time
inline uint32_t xchg[32];
inline thread_local int lane = 0;
inline void sync() { bar->arrive_and_wait(); }
// tools/emu_shuffle_count.py: warp exchanges (= SHFL / VOTE / REDUX instructions of the real kernel) per phase of the substep
inline long long exchanges[16] = {0};
inline int phase = 15;
#define ODUCK_PHASE_MARK(bit) { if (warp_emu::lane == 0) warp_emu::phase = (bit); }
template <typename F>
inline uint32_t exchange(uint32_t mine, F pick) {   // every lane publishes `mine`, then reads what `pick` selects
  if (lane == 0) ++exchanges[phase & 15];
  xchg[lane] = mine;
  sync();
  uint32_t r = pick(xchg);
  sync();
  return r;
}
inline uint32_t bits(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
inline float flt(uint32_t u) { float v; std::memcpy(&v, &u, 4); return v; }
}  // namespace warp_emu

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::sync(); }
inline int __shfl_sync(unsigned, int v, int src) { return (int)warp_emu::exchange((uint32_t)v, [&](const uint32_t* x) { return x[src & 31]; }); }
inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return warp_emu::exchange(v, [&](const uint32_t* x) { return x[src & 31]; }); }
inline float __shfl_sync(unsigned, float v, int src) { return warp_emu::flt(warp_emu::exchange(warp_emu::bits(v), [&](const uint32_t* x) { return x[src & 31]; })); }
inline int __shfl_up_sync(unsigned, int v, int d) { return (int)warp_emu::exchange((uint32_t)v, [&](const uint32_t* x) { return warp_emu::lane >= d ? x[warp_emu::lane - d] : (uint32_t)v; }); }
inline float __shfl_up_sync(unsigned, float v, int d) { return warp_emu::flt(warp_emu::exchange(warp_emu::bits(v), [&](const uint32_t* x) { return warp_emu::lane >= d ? x[warp_emu::lane - d] : warp_emu::bits(v); })); }
inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)warp_emu::exchange((uint32_t)v, [&](const uint32_t* x) { return x[(warp_emu::lane ^ m) & 31]; }); }
inline float __shfl_xor_sync(unsigned, float v, int m) { return warp_emu::flt(warp_emu::exchange(warp_emu::bits(v), [&](const uint32_t* x) { return x[(warp_emu::lane ^ m) & 31]; })); }
inline unsigned __ballot_sync(unsigned, bool p) {
  return warp_emu::exchange(p ? 1u : 0u, [&](const uint32_t* x) { unsigned b = 0; for (int i = 0; i < 32; i++) b |= (x[i] & 1u) << i; return b; });
}
inline float emu_wmaxf(float v) {   // redux.sync.max.f32
  return warp_emu::flt(warp_emu::exchange(warp_emu::bits(v), [&](const uint32_t* x) { float m = warp_emu::flt(x[0]); for (int i = 1; i < 32; i++) m = std::fmax(m, warp_emu::flt(x[i])); return warp_emu::bits(m); }));
}
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __int_as_float(int i) { float v; std::memcpy(&v, &i, 4); return v; }
inline int __float_as_int(float v) { int i; std::memcpy(&i, &v, 4); return i; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline void __sincosf(float a, float* s, float* c) { *s = std::sin(a); *c = std::cos(a); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float a) { return std::exp(a); }
using std::max;
using std::min;
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline void __syncthreads() {}      // one warp per emulated CTA (the kernels' CTA barriers only align instruction streams)
inline void sincosf(float a, float* s, float* c) { *s = std::sin(a); *c = std::cos(a); }

// ---- what a whole translation unit of the library (csrc/oduck_cuda.cu) needs: launch geometry, runtime API, kernel launch
#include <cstdlib>
#include <thread>
#include <vector>
#define __align__(n) alignas(n)
struct uint3 { unsigned x, y, z; };
inline thread_local uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
inline uint3 gridDim = {1, 1, 1}, blockDim = {32, 1, 1};
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < h; r++) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline float __uint_as_float(unsigned u) { return warp_emu::flt(u); }
template <class T> inline T __ldg(const T* p) { return *p; }

namespace warp_emu {
inline unsigned char* smem = nullptr;
// kernel<<<grid, 32, smem_bytes>>>(p): the blocks run one after the other, each as 32 threads (WPB = 1: one warp per block)
template <class K, class P>
inline void launch(K kernel, int grid, size_t smem_bytes, const P& p) {
  std::vector<unsigned char> mem(smem_bytes + 64);
  smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mem.data()) + 63) & ~(uintptr_t)63);
  gridDim.x = (unsigned)grid;
  std::barrier<> b(32);
  bar = &b;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++)
    th.emplace_back([&, l]() {
      lane = l;
      threadIdx.x = (unsigned)l;
      for (int blk = 0; blk < grid; blk++) {
        blockIdx.x = (unsigned)blk;
        kernel(p);
        sync();                       // the next block reuses the shared-memory block
      }
    });
  for (auto& t : th) t.join();
}
}  // namespace warp_emu
