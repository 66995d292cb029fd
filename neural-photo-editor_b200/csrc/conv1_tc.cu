// conv1_tc.cu -- enc_conv1 (reference IAN_simple.py:73-83 / IAN.py:71-80: 3 -> 128 channels, 5x5, stride 2, pad 2,
// bias, LeakyRectify(0.2)) on the tensor cores.
//
// K = 3*25 = 75 is too thin for a TMA-fed tap GEMM (the input has 3 channels, not a multiple of 64), so the im2col
// tile is built by threads: 16 producer warps stage the 11 x 67 x 3 float32 input patch of a 4-row x 32-column output
// tile in shared memory, expand it to the 128 x 80 (K padded) operand, split every value into bf16 hi|lo and write it
// straight into the 128B-swizzled K-major layout tcgen05 reads (the same layout TMA would produce), then
// fence.proxy.async + mbarrier hand it to the MMA warp.  Weights (128 x 80, hi|lo) are TMA-loaded once per CTA.
// 15 MMAs per tile (5 K-slices x 3 passes), two TMEM accumulator buffers, epilogue = bias + LReLU + re-split.
// Roles: warp 0 weight TMA, warp 1 MMA, warps 2-9 epilogue, warps 10-25 im2col (four threads per operand row).
//
// Operand layout: K = 80 is one 64-wide chunk (hi plane, lo plane) plus a 16-wide tail.  The tail's hi AND lo slices share
// ONE 128-byte-row plane (hi at K-slice position 0, lo at position 1: a K = 16 slice is just a +32 B start offset in the
// descriptor), so a stage is 48 KB instead of 64 KB -- which pays for the output staging below.
//
// Output: the 128 pixels x 128 channels of a tile are 32 KB CONTIGUOUS per plane in the NHWC activation.  Round 2's
// first form stored them from registers, 16 B per lane at a 256 B stride: ncu showed 16 of 32 bytes used per sector, the
// L1 store path 65 % busy and every role (producers' LDS/STS included) queueing behind it ("stall_mio").  Now the
// epilogue writes the tile into four 128B-swizzled staging tiles in shared memory (conflict-free 16-byte stores) and one
// thread per half issues two TMA stores: no global store instruction is left in the kernel.
#include <cstdio>
#include <cstring>

#include "edge.h"
#include "tc_ptx.cuh"

namespace ian {

struct Conv1Maps {
  CUtensorMap b;   // weights: 3 blocks of [128 cout][64 k] (hi k<64 | lo k<64 | hi k 64..79, lo k 64..79, zeros)
};
struct Conv1OutMap {
  CUtensorMap out; // a1 activation [2 planes][n*1024 pixels][128 ch] bf16, box 64 ch x 128 pixels, 128B swizzle
};

namespace {

using namespace tc;

constexpr int kProducers = 512;
constexpr int kThreads = 320 + kProducers;
constexpr int kChunkPlane = 128 * 64 * 2;          // one 128-row x 64-k bf16 plane: 16 KB
constexpr int kAStage = 3 * kChunkPlane;           // k 0..63 hi | k 0..63 lo | tail plane (hi, lo slices): 48 KB
constexpr int kBBytes = 3 * kChunkPlane;           // weights, same three blocks x 128 rows: 48 KB
constexpr int kOutBytes = 4 * kChunkPlane;         // output staging: (channel half) x (hi|lo) tiles of 128 pixels x 64 ch
constexpr int kPatchRows = 11, kPatchCols = 67;    // input rows 2*p0-2 .. 2*p0+8, columns -2 .. 64
constexpr int kPatchFloats = 3 * kPatchRows * kPatchCols;
constexpr int kPatchBytes = (kPatchFloats * 4 + 15) / 16 * 16;
// no alignment slack: the kernel has no static shared memory, so the dynamic window starts 1024-aligned (the kernel traps
// if it ever does not); 227 KB per CTA minus the 1 KB the toolchain reserves per block on sm_100
constexpr int kSmemBytes = 2 * kAStage + kBBytes + kOutBytes + 2 * kPatchBytes + 128 + 512;   // + barriers + staged bias
static_assert(kSmemBytes <= 232448 - 1024, "conv1_tc: shared memory budget");
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
  constexpr int unit = (K0 % 64) / 8;
  if (K0 < 64) {
    uint8_t* base = stage + m * 128 + ((unit ^ (m & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + kChunkPlane) = *reinterpret_cast<const uint4*>(lo);
  } else {                                             // tail plane: hi in K-slice position 0 (units 0,1), lo in position 1 (units 2,3)
    uint8_t* row = stage + 2 * kChunkPlane + m * 128;
    *reinterpret_cast<uint4*>(row + ((unit ^ (m & 7)) << 4)) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(row + (((unit + 2) ^ (m & 7)) << 4)) = *reinterpret_cast<const uint4*>(lo);
  }
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
conv1_tc_kernel(const __grid_constant__ Conv1Maps maps, const __grid_constant__ Conv1OutMap omap, const float* __restrict__ x,
                const float* __restrict__ bias, const int n_img) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();                     // the 128B-swizzled tiles need 1024-byte alignment (see kSmemBytes)
  uint8_t* smem_al = smem_raw;
  const uint32_t a_base = smem_base, b_base = a_base + 2 * kAStage, o_base = b_base + kBBytes, p_base = o_base + kOutBytes;
  const uint32_t bar_base = p_base + 2 * kPatchBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (4 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (6 + b); };
  const uint32_t b_bar = bar_base + 64u, tmem_slot = bar_base + 72u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_al + (tmem_slot - smem_base));
  float* bias_s = reinterpret_cast<float*>(smem_al + (bar_base + 128u - smem_base));   // 128 floats, 16-byte aligned
  if (threadIdx.x < 128) bias_s[threadIdx.x] = __ldg(bias + threadIdx.x);   // the epilogue reads it 2048 times per thread

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_img * kTilesPerImage;

  if (threadIdx.x == 0) {
    pdl_trigger();                                      // tapgemm.h: PDL
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
    if (lane == 0) {                                    // weights, once: three 16 KB blocks in one box (constants: no pdl_wait)
      mbar_expect_tx(b_bar, kBBytes);
      tma_load_3d(&maps.b, b_bar, b_base, 0, 0, 0);
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
          for (int ks = 0; ks < 5; ++ks) {              // K = 80: slices 0..3 of the 64-wide chunk, then the tail
            const bool tail = ks == 4;
            const uint64_t ko = (uint64_t)((ks & 3) * 2);   // a K = 16 slice is 32 B = 2 descriptor address units
            const uint64_t a_hi = tail ? make_sw128_desc(sa + 2 * kChunkPlane) : make_sw128_desc(sa) + ko;
            const uint64_t a_lo = tail ? make_sw128_desc(sa + 2 * kChunkPlane) + 2 : make_sw128_desc(sa + kChunkPlane) + ko;
            const uint64_t b_hi = tail ? make_sw128_desc(b_base + 2 * kChunkPlane) : make_sw128_desc(b_base) + ko;
            const uint64_t b_lo = tail ? make_sw128_desc(b_base + 2 * kChunkPlane) + 2 : make_sw128_desc(b_base + kChunkPlane) + ko;
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
    // ============ epilogue: bias + LeakyRectify(0.2) + hi|lo re-split -> swizzled staging tiles -> TMA store ============
    const int ew = warp - 2, lg = warp & 3, half = ew >> 2;
    const int m = lg * 32 + lane;
    const bool issuer = (ew & 3) == 0 && lane == 0;     // one thread per channel half issues that half's two TMA stores
    const uint32_t o_hi = o_base + (uint32_t)(half * 2) * kChunkPlane, o_lo = o_hi + kChunkPlane;
    const uint32_t row_hi = o_hi + (uint32_t)m * 128u, row_lo = o_lo + (uint32_t)m * 128u;
    const uint32_t mx = (uint32_t)(m & 7);
    const int group_bar = 3 + half;                      // named barrier of the half's four warps
    pdl_wait();                                          // a1 may still be read by the previous step's enc_conv2 only transitively; be exact
    uint32_t t = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++t) {
      const int n = w / kTilesPerImage, p0 = (w % kTilesPerImage) * 4;
      const uint32_t s = t & 1u, use = t >> 1;
      const uint32_t lane_addr = tmem_base + s * 256 + ((uint32_t)(lg * 32) << 16);
      mbar_wait(tfull_bar(s), use & 1u);
      tc_fence_after();
      // software-pipelined drain: the TMEM loads of chunk k+1 are in flight while chunk k is converted and staged
      uint32_t vm[2][16], vc[2][16];
      __syncwarp();
      tmem_ld16(lane_addr + half * 64, vm[0]);
      tmem_ld16(lane_addr + 128 + half * 64, vc[0]);
      if (t > 0) {                                      // the previous tile's TMA stores have read the staging tiles
        if (issuer) bulk_wait_group_read0();
        asm volatile("bar.sync %0, 128;" ::"r"(group_bar) : "memory");
      }
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
        // channels 16k .. 16k+15 of this half = 16-byte units 2k, 2k+1 of the pixel's 128-byte row (128B swizzle: unit ^ row%8;
        // the 32 lanes of a warp are 32 consecutive rows -> 8 bank groups x 4 lanes: conflict-free)
        const uint32_t u0 = ((uint32_t)(2 * k) ^ mx) << 4, u1 = ((uint32_t)(2 * k + 1) ^ mx) << 4;
        st_shared_v4(row_hi + u0, reinterpret_cast<const uint4*>(hi)[0]);
        st_shared_v4(row_hi + u1, reinterpret_cast<const uint4*>(hi)[1]);
        st_shared_v4(row_lo + u0, reinterpret_cast<const uint4*>(lo)[0]);
        st_shared_v4(row_lo + u1, reinterpret_cast<const uint4*>(lo)[1]);
      }
      fence_proxy_async_smem();                          // generic-proxy writes -> visible to the TMA store
      asm volatile("bar.sync %0, 128;" ::"r"(group_bar) : "memory");
      if (issuer) {
        const int pix0 = (n * 32 + p0) * 32;             // the tile's 128 pixels are consecutive in NHWC
        tma_store_3d(&omap.out, o_hi, half * 64, pix0, 0);
        tma_store_3d(&omap.out, o_lo, half * 64, pix0, 1);
        bulk_commit_group();
      }
    }
    if (issuer) bulk_wait_group0();                      // every byte has left before the CTA retires its shared memory
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
          const int iy = 2 * p0 - 2 + rem / kPatchCols, ix = rem % kPatchCols - 2;   // rows 2*p0-2..2*p0+8, cols -2..64
          if (iy >= 0 && iy < 64 && ix >= 0 && ix < 64) v = __ldg(x + (((long long)n * 3 + ch) * 64 + iy) * 64 + ix);
        }
        pre[e] = v;
      }
    };
    pdl_wait();                                         // x may be the output of an earlier kernel of the caller's stream
    if ((int)blockIdx.x < total) fetch(blockIdx.x);
    uint32_t t = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++t) {
      const uint32_t s = t & 1u, use = t >> 1;
      float* patch = reinterpret_cast<float*>(smem_al + (p_base - smem_base) + s * kPatchBytes);
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

Conv1Maps* conv1_build_maps(const __nv_bfloat16* wt, char* err, int errlen) {
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  Conv1Maps* m = new Conv1Maps();
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[3] = {64, 128, 3};
  cuuint64_t strides[2] = {64 * 2, 128 * 64 * 2};
  cuuint32_t box[3] = {64, 128, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&m->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)wt, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(conv1 B) failed: %d", (int)r); delete m; return nullptr; }
  return m;
}

void conv1_free_maps(Conv1Maps* m) { delete m; }

// the a1 activation of one plan: [2 planes][n*1024 pixels][128 channels] bf16, stored in 64-channel x 128-pixel boxes
Conv1OutMap* conv1_build_out_map(__nv_bfloat16* out, long long plane, int n, char* err, int errlen) {
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  Conv1OutMap* m = new Conv1OutMap();
  memset(m, 0, sizeof(*m));
  cuuint64_t dims[3] = {128, (cuuint64_t)n * 1024, 2};
  cuuint64_t strides[2] = {128 * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&m->out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)out, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(conv1 out) failed: %d", (int)r); delete m; return nullptr; }
  return m;
}

void conv1_free_out_map(Conv1OutMap* m) { delete m; }

int launch_conv1_tc(const Conv1Maps* maps, const Conv1OutMap* omap, const float* x, const float* bias, int n, cudaStream_t st) {
  static DeviceOnce attr_set;
  const int dev = cur_device();
  if (!attr_set.is_done(dev)) {
    if (cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return -1;
    attr_set.set_done(dev);
  }
  const int num_sms = tc_num_sms();
  const int total = n * kTilesPerImage;
  const int grid = total < num_sms ? total : num_sms;
  if (launch_pdl(conv1_tc_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, *maps, *omap, x, bias, n) != cudaSuccess) return -1;
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ian
