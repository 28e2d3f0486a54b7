// Probe: tcgen05.mma with an MN-major (n-contiguous) B operand loaded by TMA as [k rows][64 n] boxes.
// D[128,128] = A[128,64] (K-major) x Bm[64,128] (MN-major).  Build: nvcc -arch=sm_100a -I vqa_collection_b200/csrc
// Run on a B200; prints max abs error vs a host reference for a few descriptor variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc_common.cuh"

using namespace vqa::tc;

namespace vqa { int fail(int c, const char* f, ...) { printf("fail %s\n", f); return c; } void count_launch(int) {} int sm_count() { return 148; } }

__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t sa = base, sb = base + 16384, bar = base + 16384 + 16384, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_ptr = (volatile uint32_t*)(bp + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 128);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768);
    tma_load_2d(sa, &tmA, bar, 0, 0);
    tma_load_2d(sb, &tmB, bar, 0, 0);          // n 0..63,   k rows 0..63
    tma_load_2d(sb + 8192, &tmB, bar, 64, 0);  // n 64..127, k rows 0..63
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    // idesc: D f32, A/B bf16, A K-major, B MN-major (bit 16), N=128, M=128
    const uint32_t idesc = make_idesc_bf16(128, 128) | (1u << 16);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = make_sw128_kmajor_desc(sa) + 2 * k;
      uint64_t bd;
      if (variant == 0) bd = make_sw128_mn_desc(sb + k * 2048, 8192, 1024);        // LBO = MN-atom stride, SBO = k-group stride
      else if (variant == 1) bd = make_sw128_mn_desc(sb + k * 2048, 1024, 8192);   // swapped
      else bd = make_sw128_mn_desc(sb + k * 2048, 8192, 1024) ;
      umma_bf16(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  return (int)((EncodeTiledFn)f)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  std::vector<__nv_bfloat16> A(128 * 64), B(64 * 128);
  std::vector<float> Af(128 * 64), Bf(64 * 128), ref(128 * 128, 0.f), out(128 * 128);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
  for (int i = 0; i < 64 * 128; ++i) { float v = (rand() % 13 - 6) / 4.f; B[i] = __float2bfloat16(v); Bf[i] = __bfloat162float(B[i]); }
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += Af[m * 64 + k] * Bf[k * 128 + n]; ref[m * 128 + n] = s; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  printf("map rc %d %d\n", make_map(&tmA, dA, 128, 64, 64, 128), make_map(&tmB, dB, 64, 128, 128, 64));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0, out.size() * 4);
    probe<<<1, 128, 40000>>>(tmA, tmB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs(out[i] - ref[i]));
    printf("variant %d: %s max abs err %.4f  (D[0,0]=%.3f ref %.3f; D[5,70]=%.3f ref %.3f)\n", variant, cudaGetErrorString(e), err, out[0], ref[0], out[5 * 128 + 70], ref[5 * 128 + 70]);
  }
  return 0;
}
