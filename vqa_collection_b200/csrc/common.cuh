// Shared host/device helpers for the vqa_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vqa_b200.h"

namespace vqa {

// ---- error reporting (thread local, returned through vqa_last_error) -------
char* error_buffer();
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int launch_count();
void reset_launch_count();

#define VQA_CUDA_CHECK(expr)                                                         \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return ::vqa::fail(VQA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,  \
                         cudaGetErrorString(_e));                                    \
  } while (0)

#define VQA_LAUNCH_CHECK()                                                           \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return ::vqa::fail(VQA_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__,        \
                         __LINE__, cudaGetErrorString(_e));                          \
    ::vqa::count_launch();                                                           \
  } while (0)

#define VQA_REQUIRE(cond, ...)                                                       \
  do {                                                                               \
    if (!(cond)) return ::vqa::fail(VQA_ERR_INVALID, __VA_ARGS__);                   \
  } while (0)

int require_sm100();          // VQA_OK or VQA_ERR_UNSUPPORTED
int sm_count();
int current_device();         // cudaGetDevice, -1 when there is none

// cudaFuncSetAttribute / occupancy answers are per DEVICE, not per process: one of these per call site remembers which
// devices have been set up (a second device in one process — the reference's `decoder_device`, main.py:88 — would
// otherwise launch with the 48 KB default and fail).  Devices >= 64 are simply set up on every call.
struct DeviceOnce {
  unsigned long long done = 0;
  bool need(int dev) const { return dev < 0 || dev >= 64 || !((done >> dev) & 1ull); }
  void mark(int dev) { if (dev >= 0 && dev < 64) done |= 1ull << dev; }
};
// a small per-device integer cache (-1 = not computed yet)
struct DeviceInt {
  int v[64];
  DeviceInt() { for (int i = 0; i < 64; ++i) v[i] = -1; }
  int& at(int dev) { return v[(dev >= 0 && dev < 64) ? dev : 0]; }
};

inline size_t elem_size(int dtype) { return dtype == VQA_BF16 ? 2 : 4; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch (PDL) -----------------------------------------
// The forward path is a chain of short kernels on one stream.  Each kernel calls griddep_launch()
// once its CTA is set up (the NEXT kernel's CTAs may then be scheduled as SMs free up and run their
// own prologue: barrier init, TMEM allocation, tensor-map prefetch) and griddep_wait() before it
// touches global memory (returns when the PREVIOUS kernel has completed and its writes are
// visible).  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();            // VQA_B200_NO_PDL=1 turns the launch attribute off (A/B timing)
// Consumers of a tensor that the previous kernel swept in ascending order walk it in DESCENDING order, so that they
// start on the part that is still in the 126 MB L2 (VQA_B200_NO_L2_ORDER=1 restores ascending order for A/B timing).
bool l2_order_enabled();

// launch `kernel` so that it may start while the previous kernel of the stream is still draining
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- device-side element helpers -------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float to_f(float x) { return x; }
  static __device__ __forceinline__ float from_f(float x) { return x; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
};

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for f32)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- internal launchers shared between api.cu and the kernels --------------
// token-table form of the question encoder (gru_pair.cu): gi_table fp16 [ntoken_rows, 3H]; zero_after_counter = bytes
// behind the GRU's 256-byte counter block that the call clears together with the counters
struct GruTokenTable { const void* gi_table; const int64_t* tokens; int ntoken_rows; size_t zero_after_counter; };
// steps [t, t_end) of the fp32-class (VQA_F16X2) GRU, GEMM + gate update fused (gemm_tc.cu)
struct GruStepSplit {
  int B, T, t, t_end, H, ntoken_rows;
  const int64_t* tokens; const float* gi_table; const float* b_hh;
  const void* wh_packed_planes;             // [2][3H][H] fp16, gate-interleaved 96-row blocks: [r|z|n] x 32 units (engine.pack_gru(units=32))
  void* h_planes;                           // [2 buffers][2 planes][B][H] fp16: step t reads buffer (t-1)&1, writes t&1
  void* h_planes_last;                      // where step T-1 writes its plane pair instead (or NULL)
  float* h;                                 // f32 [B,H], updated in place
  float* h_out_last;                        // where step T-1 writes the f32 state instead (or NULL)
  int* counter;                             // >= ceil(B/128) ints (one launch for all steps) or NULL
};
struct GruTrainSave { float *R, *Z, *N, *HN, *Hs; };      // f32 [T,B,H] each; Hs slot t = state after step t
int linear_simt(const vqa_linear_args& a, cudaStream_t s);
int linear_tc(const vqa_linear_args& a, cudaStream_t s);
int linear_tc_part_width();
int linear_tc_tile_count(const vqa_linear_args& a);
int linear_tc_tiles_n(const vqa_linear_args& a);

}  // namespace vqa
