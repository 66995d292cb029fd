// head_tc.cu -- the RGB-Beta head of IAN.py / IANv1.py (reference IAN.py:183-207, layers.py:207-258, 397-408) without the
// HBM tap table.
//
// The head applies three 128 -> 2 channel MDC convolutions (scales [2,3,4]: 33 distinct dilated tap offsets, halo 4) to
// the 64x64x128 feature map, then an autoregressive 2->2 / 4->2 channel pair of MDC convolutions with sigmoids and the
// Beta mean 2a/(a+b)-1.  Round 1 ran the first part as a dense 128 -> 33*6 GEMM into a channel-major table in HBM
// (3.2 MB per image) that a gather kernel read back: ~6x the algorithmic bytes, 0.56 + 0.26 ms at batch 512.
//
// head_tc_kernel: a work item is (image, conv k in {R, G_a, B_a}).  The CTA streams the image's 32 M tiles (2 rows x 64
// pixels) through ONE dense GEMM each,  T[pixel][t*2+f] = sum_c h[pixel][c] * Wk[t][f][c]  (N = 66 -> 80, K = 128: no
// shifts, the feature map is staged exactly once per conv), drains T to shared memory, and every epilogue thread adds
// the taps that land on ITS 16 output pixels,  ha[p][f] += T[p + off_t][t*2+f],  into registers (taps are sorted by row
// offset, so a tile contributes to an output row through at most two contiguous column ranges).  No atomics, fixed order.
// Roles as in decout_tc.cu: warp 0 TMA (weights per item, A ring), warp 1 tcgen05 issue (float32 = 3-pass bf16 split,
// main|cross accumulators; or single pass), warps 2-9 epilogue; two TMEM buffers so tile i+1 multiplies while tile i
// is gathered.
//
// head_rgb_kernel: one CTA per image keeps R and G in shared memory and runs the autoregressive part (sigmoid R; G from
// 33 taps of R; B from 33 taps of [R,G]; Beta means) in one pass -- three kernels and two HBM round trips before.
#include <cstdio>
#include <cstring>

#include "edge.h"
#include "tc_ptx.cuh"

namespace ian {

struct HeadMaps {
  CUtensorMap a, a1;   // feature map planes (C=128, W=64, H=64, N, planes): box {64 ch, 64, 2, 1, planes}
  CUtensorMap b, b1;   // weights (K=128, 3*80 rows, planes): box {64, 80, planes}
};

struct HeadTaps {       // taps sorted by dy; tap j reads T column pair j
  int dy_start[10];     // taps with row offset dy = -4 + i are [dy_start[i], dy_start[i+1])
  int dx[33];
};

namespace {

using namespace tc;

constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int BN = 80;                        // 33 taps x 2 filters = 66, padded to a legal UMMA N
constexpr int kNT = 33;
constexpr int kAStages = 3;
constexpr int kTLd = 67;                      // T row pitch in floats (odd: conflict-free column access)
constexpr int kTBytes = 128 * kTLd * 4;
constexpr int kTilesPerImage = 32;

template <int PASSES> struct HeadCfg {
  static constexpr int kPlanes = PASSES == 3 ? 2 : 1;
  static constexpr int kAStage = 128 * 64 * 2 * kPlanes;      // one K chunk of an A tile
  static constexpr int kBChunk = BN * 64 * 2 * kPlanes;       // one K chunk of the conv's weights
  static constexpr int kSmemBytes = 1024 + kAStages * kAStage + 2 * kBChunk + kTBytes + 256;
};

template <int PASSES>
__global__ void __launch_bounds__(kThreads, 1)
head_tc_kernel(const __grid_constant__ HeadMaps maps, const __grid_constant__ HeadTaps taps, float* __restrict__ ha /*[n][6][4096]*/,
               const int n_img) {
  using Cfg = HeadCfg<PASSES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + kAStages * Cfg::kAStage;
  const uint32_t t_base = b_base + 2 * Cfg::kBChunk;
  const uint32_t bar_base = t_base + kTBytes;
  float* Ts = reinterpret_cast<float*>(smem_al + (t_base - smem_base));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kAStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kAStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kAStages + 2 + b); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * kAStages + 4);
  const uint32_t bempty_bar = bar_base + 8u * (2 * kAStages + 5);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kAStages + 6);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n_img * 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiThreads / 32); }
    mbar_init(bfull_bar, 1);
    mbar_init(bempty_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t i = 0, it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int n = w / 3, k = w % 3;
        mbar_wait(bempty_bar, (it & 1u) ^ 1u);           // the previous item's MMAs have finished reading the weights
        mbar_expect_tx(bfull_bar, 2 * Cfg::kBChunk);
        tma_load_3d(PASSES == 3 ? &maps.b : &maps.b1, bfull_bar, b_base, 0, k * BN, 0);
        tma_load_3d(PASSES == 3 ? &maps.b : &maps.b1, bfull_bar, b_base + Cfg::kBChunk, 64, k * BN, 0);
        for (int mt = 0; mt < kTilesPerImage; ++mt)
          for (int c = 0; c < 2; ++c, ++i) {
            const int s = i % kAStages;
            mbar_wait(empty_bar(s), ((i / kAStages) & 1u) ^ 1u);
            mbar_expect_tx(full_bar(s), Cfg::kAStage);
            tma_load_5d(PASSES == 3 ? &maps.a : &maps.a1, full_bar(s), a_base + s * Cfg::kAStage, c * 64, 0, 2 * mt, n, 0);
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_m128(BN);
      uint32_t i = 0, t = 0, it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        mbar_wait(bfull_bar, it & 1u);
        tc_fence_after();
        for (int mt = 0; mt < kTilesPerImage; ++mt, ++t) {
          const uint32_t buf = t & 1u, use = t >> 1;
          const uint32_t acc_main = tmem_base + buf * 256, acc_cross = acc_main + 128;
          mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);
          tc_fence_after();
          for (int c = 0; c < 2; ++c, ++i) {
            const int s = i % kAStages;
            mbar_wait(full_bar(s), (i / kAStages) & 1u);
            tc_fence_after();
            const uint32_t sa = a_base + s * Cfg::kAStage, sb = b_base + c * Cfg::kBChunk;
            const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + 128 * 64 * 2);
            const uint64_t b_hi = make_sw128_desc(sb), b_lo = make_sw128_desc(sb + BN * 64 * 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);
              const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
              umma_bf16(acc_main, a_hi + ko, b_hi + ko, idesc, acc);
              if (PASSES == 3) {
                umma_bf16(acc_cross, a_lo + ko, b_hi + ko, idesc, acc);
                umma_bf16(acc_cross, a_hi + ko, b_lo + ko, idesc, 1u);
              }
            }
            umma_commit(empty_bar(s));
          }
          umma_commit(tfull_bar(buf));
        }
        umma_commit(bempty_bar);                         // weights of this item no longer read once these MMAs retire
      }
    }
  } else {
    // ===================== epilogue: TMEM -> smem T tile -> tap gather into registers =====================
    const int et = threadIdx.x - 64;                    // 0..255
    const int ew = warp - 2;
    const int lg = warp & 3;                            // TMEM lane group
    const int half = ew >> 2;                           // T columns [0,48) or [48,80)
    const int row = lg * 32 + lane;                     // T tile row = pixel (pr*64 + q) of input rows 2mt + pr
    const int q = et & 63, r0 = et >> 6;                // this thread's output pixels: column q, rows r0 + 4*i, i = 0..15
    uint32_t t = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int n = w / 3, k = w % 3;
      float acc[16][2];
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
      for (int mt = 0; mt < kTilesPerImage; ++mt, ++t) {
        const uint32_t buf = t & 1u, use = t >> 1;
        const uint32_t lane_addr = tmem_base + buf * 256 + ((uint32_t)(lg * 32) << 16);
        mbar_wait(tfull_bar(buf), use & 1u);
        tc_fence_after();
        const int c_begin = half ? 48 : 0, c_end = half ? 80 : 48;
#pragma unroll 1
        for (int cb = c_begin; cb < c_end; cb += 16) {
          uint32_t vm[16], vc[16];
          __syncwarp();
          tmem_ld16(lane_addr + cb, vm);
          if (PASSES == 3) tmem_ld16(lane_addr + 128 + cb, vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb + j < 2 * kNT)
              Ts[row * kTLd + cb + j] = PASSES == 3 ? __uint_as_float(vm[j]) + __uint_as_float(vc[j]) : __uint_as_float(vm[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));     // TMEM buffer free: the next tile's MMAs may start
        asm volatile("bar.sync 1, 256;" ::: "memory");   // T tile complete (epilogue warps only)

        // ha[p][f] += T[(p + dy, q + dx)][j*2 + f] for the taps whose input row p + dy lies in this tile (rows 2mt, 2mt+1)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int p = r0 + 4 * i;
#pragma unroll
          for (int pr = 0; pr < 2; ++pr) {
            const int dy = 2 * mt + pr - p;              // the row offset that reaches input row 2mt + pr from p
            if (dy < -4 || dy > 4) continue;
            const int j0 = taps.dy_start[dy + 4], j1 = taps.dy_start[dy + 5];
            for (int j = j0; j < j1; ++j) {
              const int qq = q + taps.dx[j];
              if (qq < 0 || qq > 63) continue;           // image border (rows outside the image are simply never a tile)
              const float* tp = Ts + (pr * 64 + qq) * kTLd + 2 * j;
              acc[i][0] += tp[0];
              acc[i][1] += tp[1];
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // T tile consumed: may be overwritten
      }
      float* o0 = ha + ((long long)n * 6 + 2 * k) * 4096 + q;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        o0[(r0 + 4 * i) * 64] = acc[i][0];
        o0[4096 + (r0 + 4 * i) * 64] = acc[i][1];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- autoregressive part: one CTA per image, R and G kept in shared memory ---------------------------------------
//   R = sig(ha[0:2]);  G = sig(ha[2:4] + MDC_Gb(R));  B = sig(ha[4:6] + MDC_Bb([R,G]));  out_c = 2 a/(a+b+1e-8) - 1
// taps: [33][2] (dy,dx) in the ORIGINAL tap order of wgb [33][2 out][2 in] / wbb [33][2 out][4 in]
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int kRgbThreads = 512;

__global__ void __launch_bounds__(kRgbThreads) head_rgb_kernel(const float* __restrict__ ha /*[n][6][4096]*/, const int* __restrict__ taps,
                                                               const float* __restrict__ wgb, const float* __restrict__ wbb, int ntaps,
                                                               float* __restrict__ xhat, float* __restrict__ rg /*(n,4096,4)*/,
                                                               float* __restrict__ bsave /*nullable (n,4096,2)*/) {
  extern __shared__ float sm[];
  float2* R = reinterpret_cast<float2*>(sm);             // [4096]
  float2* G = R + 4096;                                  // [4096]
  float* wg = reinterpret_cast<float*>(G + 4096);        // [33*4]
  float* wb = wg + 33 * 4;                               // [33*8]
  int* tp = reinterpret_cast<int*>(wb + 33 * 8);         // [33*2]
  const long long img = blockIdx.x;
  const float* h = ha + img * 6 * 4096;
  for (int i = threadIdx.x; i < ntaps * 4; i += kRgbThreads) wg[i] = wgb[i];
  for (int i = threadIdx.x; i < ntaps * 8; i += kRgbThreads) wb[i] = wbb[i];
  for (int i = threadIdx.x; i < ntaps * 2; i += kRgbThreads) tp[i] = taps[i];
  for (int i = threadIdx.x; i < 4096; i += kRgbThreads) R[i] = make_float2(sigmoid_f(h[i]), sigmoid_f(h[4096 + i]));
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += kRgbThreads) {
    const int p = i >> 6, q = i & 63;
    float g0 = h[2 * 4096 + i], g1 = h[3 * 4096 + i];
    for (int t = 0; t < ntaps; ++t) {
      const int pp = p + tp[2 * t], qq = q + tp[2 * t + 1];
      if (pp < 0 || pp > 63 || qq < 0 || qq > 63) continue;
      const float2 r = R[pp * 64 + qq];
      const float4 w = *reinterpret_cast<const float4*>(wg + t * 4);     // [out0: in0,in1 | out1: in0,in1]
      g0 = fmaf(r.x, w.x, fmaf(r.y, w.y, g0));
      g1 = fmaf(r.x, w.z, fmaf(r.y, w.w, g1));
    }
    G[i] = make_float2(sigmoid_f(g0), sigmoid_f(g1));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += kRgbThreads) {
    const int p = i >> 6, q = i & 63;
    float b0 = h[4 * 4096 + i], b1 = h[5 * 4096 + i];
    for (int t = 0; t < ntaps; ++t) {
      const int pp = p + tp[2 * t], qq = q + tp[2 * t + 1];
      if (pp < 0 || pp > 63 || qq < 0 || qq > 63) continue;
      const float2 r = R[pp * 64 + qq], g = G[pp * 64 + qq];
      const float4 w0 = *reinterpret_cast<const float4*>(wb + t * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wb + t * 8 + 4);
      b0 = fmaf(r.x, w0.x, fmaf(r.y, w0.y, fmaf(g.x, w0.z, fmaf(g.y, w0.w, b0))));
      b1 = fmaf(r.x, w1.x, fmaf(r.y, w1.y, fmaf(g.x, w1.z, fmaf(g.y, w1.w, b1))));
    }
    const float B0 = sigmoid_f(b0), B1 = sigmoid_f(b1);
    const float2 r = R[i], g = G[i];
    *reinterpret_cast<float4*>(rg + (img * 4096 + i) * 4) = make_float4(r.x, r.y, g.x, g.y);
    if (bsave) *reinterpret_cast<float2*>(bsave + (img * 4096 + i) * 2) = make_float2(B0, B1);
    float* o = xhat + img * 3 * 4096 + i;
    o[0] = 2.f * (r.x / (r.x + r.y + 1e-8f)) - 1.f;      // beta_layer (layers.py:408)
    o[4096] = 2.f * (g.x / (g.x + g.y + 1e-8f)) - 1.f;
    o[8192] = 2.f * (B0 / (B0 + B1 + 1e-8f)) - 1.f;
  }
}

}  // namespace

HeadMaps* head_build_maps(const __nv_bfloat16* fh4, long long fh4_plane, int n_img, const __nv_bfloat16* wt, long long wt_plane,
                          char* err, int errlen) {
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  HeadMaps* m = new HeadMaps();
  memset(m, 0, sizeof(*m));
  for (int planes = 2; planes >= 1; --planes) {
    {
      cuuint64_t dims[5] = {128, 64, 64, (cuuint64_t)n_img, 2};
      cuuint64_t strides[4] = {128 * 2, 64 * 128 * 2, 64 * 64 * 128 * 2, (cuuint64_t)fh4_plane * 2};
      cuuint32_t box[5] = {64, 64, 2, 1, (cuuint32_t)planes};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(planes == 2 ? &m->a : &m->a1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)fh4, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(head A) failed: %d", (int)r); delete m; return nullptr; }
    }
    {
      cuuint64_t dims[3] = {128, 3 * BN, 2};
      cuuint64_t strides[2] = {128 * 2, (cuuint64_t)wt_plane * 2};
      cuuint32_t box[3] = {64, BN, (cuuint32_t)planes};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(planes == 2 ? &m->b : &m->b1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)wt, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(head B) failed: %d", (int)r); delete m; return nullptr; }
    }
  }
  return m;
}

void head_free_maps(HeadMaps* m) { delete m; }

int launch_head_tc(const HeadMaps* maps, const int* dy_start /*[10]*/, const int* dx /*[33]*/, int passes, float* ha, int n,
                   cudaStream_t st) {
  HeadTaps tp;
  for (int i = 0; i < 10; ++i) tp.dy_start[i] = dy_start[i];
  for (int i = 0; i < 33; ++i) tp.dx[i] = dx[i];
  static DeviceOnce attr3, attr1;
  const int dev = cur_device();
  const int total = n * 3;
  const int num_sms = tc_num_sms();
  const int grid = total < num_sms ? total : num_sms;
  if (passes == 3) {
    if (!attr3.is_done(dev)) {
      if (cudaFuncSetAttribute(head_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadCfg<3>::kSmemBytes) != cudaSuccess) return -1;
      attr3.set_done(dev);
    }
    head_tc_kernel<3><<<grid, kThreads, HeadCfg<3>::kSmemBytes, st>>>(*maps, tp, ha, n);
  } else {
    if (!attr1.is_done(dev)) {
      if (cudaFuncSetAttribute(head_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadCfg<1>::kSmemBytes) != cudaSuccess) return -1;
      attr1.set_done(dev);
    }
    head_tc_kernel<1><<<grid, kThreads, HeadCfg<1>::kSmemBytes, st>>>(*maps, tp, ha, n);
  }
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_head_rgb(const float* ha, const int* taps, const float* wgb, const float* wbb, int ntaps, float* xhat, float* rg,
                    float* bsave, int n, cudaStream_t st) {
  static DeviceOnce attr;
  const int dev = cur_device();
  const int smem = 2 * 4096 * 8 + 33 * (4 + 8) * 4 + 33 * 2 * 4 + 64;
  if (!attr.is_done(dev)) {
    if (cudaFuncSetAttribute(head_rgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    attr.set_done(dev);
  }
  if (ntaps != 33) return -1;
  head_rgb_kernel<<<n, kRgbThreads, smem, st>>>(ha, taps, wgb, wbb, ntaps, xhat, rg, bsave);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace ian
