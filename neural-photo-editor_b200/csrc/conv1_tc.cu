// conv1_tc.cu -- enc_conv1 (reference IAN_simple.py:73-83 / IAN.py:71-80: 3 -> 128 channels, 5x5, stride 2, pad 2,
// bias, LeakyRectify(0.2)) on the tensor cores.
//
// K = 3*25 = 75 is too thin for a TMA-fed tap GEMM (the input has 3 channels, not a multiple of 64), so the im2col
// tile is built by threads: 8 producer warps stage the 11 x 68 x 3 float32 input patch of a 4-row x 32-column output
// tile in shared memory, expand it to the 128 x 80 (K padded) operand, split every value into bf16 hi|lo and write it
// straight into the 128B-swizzled K-major layout tcgen05 reads (the same layout TMA would produce), then
// fence.proxy.async + mbarrier hand it to the MMA warp.  Weights (128 x 80, hi|lo) are TMA-loaded once per CTA.
// 15 MMAs per tile (5 K-slices x 3 passes), two TMEM accumulator buffers, epilogue = bias + LReLU + re-split + NHWC
// stores into the a1 activation planes.  Roles: warp 0 weight TMA, warp 1 MMA, warps 2-9 epilogue, warps 10-25 im2col
// (16 producer warps, four threads per operand row: the expansion is the kernel's critical path -- round 1 ran it on
// 8 warps at IPC 1.2 with 70 % of the issue slots empty, i.e. latency-bound, so the fix is more warps in flight).
#include <cstdio>
#include <cstring>

#include "edge.h"
#include "tc_ptx.cuh"

namespace ian {

struct Conv1Maps {
  CUtensorMap b;   // weights (K=128 padded, 128 cout, 2 planes)
};

namespace {

using namespace tc;

constexpr int kProducers = 512;
constexpr int kThreads = 320 + kProducers;
constexpr int kChunkPlane = 128 * 64 * 2;          // one 128-row x 64-k bf16 plane: 16 KB
constexpr int kAStage = 2 * 2 * kChunkPlane;       // 2 K chunks x (hi|lo): 64 KB
constexpr int kBBytes = 2 * 2 * kChunkPlane;       // weights: 2 K chunks x (hi|lo) x 128 rows: 64 KB
constexpr int kPatchRows = 11, kPatchCols = 68;
constexpr int kPatchFloats = 3 * kPatchRows * kPatchCols;
constexpr int kSmemBytes = 1024 + 2 * kAStage + kBBytes + 2 * kPatchFloats * 4 + 256 + 512;   // + barriers + staged bias
constexpr int kTilesPerImage = 8;                  // 32 output rows / 4

// one 16-byte unit (8 consecutive k) of row m: values -> bf16 hi|lo -> swizzled position in both planes
template <int K0>
__device__ __forceinline__ void put_unit(uint8_t* stage, const float* patch, int m, int r, int c) {
  __align__(16) __nv_bfloat162 hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const int k = K0 + 2 * e + d;
      if (k < 75) {
        const int ch = k / 25, i = (k % 25) / 5, j = k % 5;
        v[d] = patch[(ch * kPatchRows + 2 * r + i) * kPatchCols + 2 * c + j];
      } else {
        v[d] = 0.f;
      }
    }
    hi[e] = __floats2bfloat162_rn(v[0], v[1]);
    const float2 hf = __bfloat1622float2(hi[e]);
    lo[e] = __floats2bfloat162_rn(v[0] - hf.x, v[1] - hf.y);
  }
  constexpr int chunk = K0 / 64, unit = (K0 % 64) / 8;
  uint8_t* base = stage + chunk * (2 * kChunkPlane) + m * 128 + ((unit ^ (m & 7)) << 4);
  *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(base + kChunkPlane) = *reinterpret_cast<const uint4*>(lo);
}

// the 10 sixteen-byte units (K = 80) of an operand row are split 3 | 3 | 2 | 2 over the row's four producer threads
__device__ __forceinline__ void build_quarter(int quarter, uint8_t* stage, const float* patch, int m, int r, int c) {
  if (quarter == 0) {
    put_unit<0>(stage, patch, m, r, c); put_unit<8>(stage, patch, m, r, c); put_unit<16>(stage, patch, m, r, c);
  } else if (quarter == 1) {
    put_unit<24>(stage, patch, m, r, c); put_unit<32>(stage, patch, m, r, c); put_unit<40>(stage, patch, m, r, c);
  } else if (quarter == 2) {
    put_unit<48>(stage, patch, m, r, c); put_unit<56>(stage, patch, m, r, c);
  } else {
    put_unit<64>(stage, patch, m, r, c); put_unit<72>(stage, patch, m, r, c);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv1_tc_kernel(const __grid_constant__ Conv1Maps maps, const float* __restrict__ x, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ out, const long long out_plane, const int n_img) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base, b_base = a_base + 2 * kAStage, p_base = b_base + kBBytes;
  const uint32_t bar_base = p_base + 2 * kPatchFloats * 4;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (4 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (6 + b); };
  const uint32_t b_bar = bar_base + 64u, tmem_slot = bar_base + 72u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_al + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_al + (bar_base + 256u - smem_base));   // 128 floats, 16-byte aligned
  if (threadIdx.x < 128) bias_s[threadIdx.x] = __ldg(bias + threadIdx.x);   // the epilogue reads it 2048 times per thread

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_img * kTilesPerImage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), kProducers / 32);
      mbar_init(empty_bar(s), 1);
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {                                    // weights, once
      mbar_expect_tx(b_bar, kBBytes);
      tma_load_3d(&maps.b, b_bar, b_base, 0, 0, 0);
      tma_load_3d(&maps.b, b_bar, b_base + 2 * kChunkPlane, 64, 0, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow; one elected lane issues) =====================
    {
      constexpr uint32_t idesc = make_idesc_bf16_m128(128);
      mbar_wait(b_bar, 0);
      uint32_t t = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++t) {
        const uint32_t s = t & 1u, use = t >> 1;
        const uint32_t acc_main = tmem_base + s * 256, acc_cross = acc_main + 128;
        mbar_wait(tempty_bar(s), (use & 1u) ^ 1u);
        mbar_wait(full_bar(s), use & 1u);
        tc_fence_after();
        const uint32_t sa = a_base + s * kAStage;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {              // K = 80: chunk 0 slices 0..3, chunk 1 slice 0
            const int chunk = ks >> 2, kk = ks & 3;
            const uint64_t ko = (uint64_t)(kk * 2);
            const uint64_t a_hi = make_sw128_desc(sa + chunk * 2 * kChunkPlane) + ko;
            const uint64_t a_lo = make_sw128_desc(sa + chunk * 2 * kChunkPlane + kChunkPlane) + ko;
            const uint64_t b_hi = make_sw128_desc(b_base + chunk * 2 * kChunkPlane) + ko;
            const uint64_t b_lo = make_sw128_desc(b_base + chunk * 2 * kChunkPlane + kChunkPlane) + ko;
            const uint32_t acc = ks > 0 ? 1u : 0u;
            umma_bf16(acc_main, a_hi, b_hi, idesc, acc);
            umma_bf16(acc_cross, a_lo, b_hi, idesc, acc);
            umma_bf16(acc_cross, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(s));
        }
        __syncwarp();
      }
    }
  } else if (warp < 10) {
    // ===================== epilogue: bias + LeakyRectify(0.2) + hi|lo re-split + NHWC store =====================
    const int ew = warp - 2, lg = warp & 3, half = ew >> 2;
    const int m = lg * 32 + lane;
    uint32_t t = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++t) {
      const int n = w / kTilesPerImage, p0 = (w % kTilesPerImage) * 4;
      const uint32_t s = t & 1u, use = t >> 1;
      const long long pix = ((long long)n * 32 + p0 + (m >> 5)) * 32 + (m & 31);
      const uint32_t lane_addr = tmem_base + s * 256 + ((uint32_t)(lg * 32) << 16);
      mbar_wait(tfull_bar(s), use & 1u);
      tc_fence_after();
      // software-pipelined drain: the TMEM loads of chunk k+1 are in flight while chunk k is converted and stored
      uint32_t vm[2][16], vc[2][16];
      __syncwarp();
      tmem_ld16(lane_addr + half * 64, vm[0]);
      tmem_ld16(lane_addr + 128 + half * 64, vc[0]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cb = half * 64 + 16 * k;
        tmem_ld_wait();
        if (k < 3) {
          __syncwarp();
          tmem_ld16(lane_addr + cb + 16, vm[(k + 1) & 1]);
          tmem_ld16(lane_addr + 128 + cb + 16, vc[(k + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(s));
        }
        const uint32_t* am = vm[k & 1];
        const uint32_t* ac = vc[k & 1];
        __align__(16) __nv_bfloat162 hi[8], lo[8];
        __align__(16) float bv[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(bv)[j] = reinterpret_cast<const float4*>(bias_s + cb)[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v0 = __uint_as_float(am[2 * j]) + __uint_as_float(ac[2 * j]) + bv[2 * j];
          float v1 = __uint_as_float(am[2 * j + 1]) + __uint_as_float(ac[2 * j + 1]) + bv[2 * j + 1];
          v0 = fmaf(0.4f, fabsf(v0), 0.6f * v0);
          v1 = fmaf(0.4f, fabsf(v1), 0.6f * v1);
          hi[j] = __floats2bfloat162_rn(v0, v1);
          const float2 hf = __bfloat1622float2(hi[j]);
          lo[j] = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
        }
        uint4* oh = reinterpret_cast<uint4*>(out + pix * 128 + cb);
        uint4* ol = reinterpret_cast<uint4*>(out + out_plane + pix * 128 + cb);
        oh[0] = reinterpret_cast<const uint4*>(hi)[0]; oh[1] = reinterpret_cast<const uint4*>(hi)[1];
        ol[0] = reinterpret_cast<const uint4*>(lo)[0]; ol[1] = reinterpret_cast<const uint4*>(lo)[1];
      }
    }
  } else {
    // ===================== im2col producers (warps 10..25) =====================
    const int pt = threadIdx.x - 320;                   // 0..511
    // quarter-major: the 32 lanes of a warp build the SAME unit group of 32 consecutive rows (no divergence, and the
    // patch reads of a warp walk consecutive columns)
    const int quarter = pt >> 7, m = pt & 127;
    const int r = m >> 5, c = m & 31;
    // the patch of tile t+1 is fetched into registers while tile t is being expanded (global latency hidden)
    constexpr int kPer = (kPatchFloats + kProducers - 1) / kProducers;    // 5 floats per thread
    float pre[kPer];
    auto fetch = [&](int w) {
      const int n = w / kTilesPerImage, p0 = (w % kTilesPerImage) * 4;
#pragma unroll
      for (int e = 0; e < kPer; ++e) {
        const int i = pt + e * kProducers;
        float v = 0.f;
        if (i < kPatchFloats) {
          const int ch = i / (kPatchRows * kPatchCols), rem = i % (kPatchRows * kPatchCols);
          const int iy = 2 * p0 - 2 + rem / kPatchCols, ix = rem % kPatchCols - 2;   // rows 2*p0-2..2*p0+8, cols -2..65
          if (iy >= 0 && iy < 64 && ix >= 0 && ix < 64) v = __ldg(x + (((long long)n * 3 + ch) * 64 + iy) * 64 + ix);
        }
        pre[e] = v;
      }
    };
    if ((int)blockIdx.x < total) fetch(blockIdx.x);
    uint32_t t = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++t) {
      const uint32_t s = t & 1u, use = t >> 1;
      float* patch = reinterpret_cast<float*>(smem_al + (p_base - smem_base)) + s * kPatchFloats;
      uint8_t* stage = smem_al + (a_base - smem_base) + s * kAStage;
      mbar_wait(empty_bar(s), (use & 1u) ^ 1u);         // the MMAs that read this stage have retired
#pragma unroll
      for (int e = 0; e < kPer; ++e)
        if (pt + e * kProducers < kPatchFloats) patch[pt + e * kProducers] = pre[e];
      asm volatile("bar.sync 2, 512;" ::: "memory");    // patch complete (producer warps only)
      if (w + (int)gridDim.x < total) fetch(w + gridDim.x);
      build_quarter(quarter, stage, patch, m, r, c);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

Conv1Maps* conv1_build_maps(const __nv_bfloat16* wt, long long wt_plane, char* err, int errlen) {
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  Conv1Maps* m = new Conv1Maps();
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[3] = {128, 128, 2};
  cuuint64_t strides[2] = {128 * 2, (cuuint64_t)wt_plane * 2};
  cuuint32_t box[3] = {64, 128, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&m->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)wt, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(conv1 B) failed: %d", (int)r); delete m; return nullptr; }
  return m;
}

void conv1_free_maps(Conv1Maps* m) { delete m; }

int launch_conv1_tc(const Conv1Maps* maps, const float* x, const float* bias, __nv_bfloat16* out, long long plane, int n,
                    cudaStream_t st) {
  static DeviceOnce attr_set;
  const int dev = cur_device();
  if (!attr_set.is_done(dev)) {
    if (cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return -1;
    attr_set.set_done(dev);
  }
  const int num_sms = tc_num_sms();
  const int total = n * kTilesPerImage;
  const int grid = total < num_sms ? total : num_sms;
  conv1_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(*maps, x, bias, out, plane, n);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ian
