// tapgemm_tc2.cu -- the shifted-tap GEMM (tapgemm.h) on CTA PAIRS: tcgen05.mma.cta_group::2, UMMA M = 256.
//
// Why a second kernel.  In float32 mode (3-pass bf16 split) the main|cross accumulators of a 128 x 256 tile fill all 512
// TMEM columns, so the one-CTA kernel (tapgemm_tc.cu) cannot start tile i+1 while the epilogue drains tile i, and a
// 128 x 128 tile -- which would double-buffer -- is operand-feed bound on one SM (M128 N128 K16 reads 8 KB of smem per
// 64 tensor clocks = the whole 128 B/clk).  A CTA pair computing a 256 x BN tile with ONE tcgen05.mma.cta_group::2 per
// K slice fixes both: each SM keeps only its own 128 rows x BN columns of accumulators (BN = 128: 2 x 128 columns per
// buffer -> two buffers fit -> the epilogue of tile i overlaps the MMAs of tile i+1), and each SM stages only HALF of
// the weight tile (the tensor cores of both SMs read both halves), so the feed is 6 KB per 64 clocks.  Work items are
// half as large per SM as before, which also quarters the wave-quantisation loss (512 pair-tiles over 74 pairs).
//
// Roles per CTA (both CTAs of a pair run the same code on their own 128 rows / their own half of the weight rows):
//   warp 0     : TMA producer: A box {64 ch, Wt, Ht, Nt, planes} of ITS m-tile, B box {64 ch, BN/2 rows, planes} of ITS
//                half of the n-tile, both signalling the LEADER's (cluster rank 0) full barrier
//                (cp.async.bulk.tensor ... .cta_group::2, mbarrier address mapped to rank 0)
//   warp 1     : TMEM alloc (cta_group::2, both CTAs); rank 0 only: one thread issues the MMAs for the pair and commits
//                with .multicast::cluster to the empty / accumulator-full barriers of BOTH CTAs
//   warps 2..9 : epilogue of this CTA's 128 rows; accumulator-empty arrives go to the leader's barrier (mapa)
// Whole-tile scheduling only (pair p takes work p, p + 74, ...): split-K, stream-K, the 16-wide head convs and the
// channel-major float32 output stay on the one-CTA kernel, which remains the path for small batches.
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "tapgemm.h"
#include "tc_ptx.cuh"

namespace ian {

struct Tc2Maps {
  CUtensorMap a[4];     // activation views, box = {64 ch, Wt, Ht, Nt, 2 planes}
  CUtensorMap b;        // weights, box = {64 ch, 64 rows (half of BN = 128), 2 planes}
  CUtensorMap a1[4];    // hi plane only (bf16 mode)
  CUtensorMap b1;       // bf16 mode: box = {64 ch, BN1/2 rows, 1 plane}
  int Wt, Ht, Nt, BN;   // BN: float32-split mode tile width (128: main|cross x 2 buffers = 512 TMEM columns)
  int BN1;              // bf16 mode tile width: 256 when Cout % 256 == 0 (one accumulator: 2 x 256 columns), else 128
  mutable int sk_choice[2];   // cached stream-K decision per mode (float32-split, bf16): -1 unknown, 0 whole tiles, 1 stream-K
};

namespace {

using namespace tc;

constexpr int BM = 128;                 // rows per CTA (UMMA M = 256 over the pair)
constexpr int BK = 64;
constexpr int kThreads = 320;           // producer, MMA, 8 epilogue warps
constexpr int kEpiWarps = 8;

template <int BN, int PASSES> struct Tc2Cfg {
  static constexpr int kPlanes = PASSES == 3 ? 2 : 1;
  static constexpr int kATileBytes = BM * BK * 2 * kPlanes;
  static constexpr int kBHalfBytes = (BN / 2) * BK * 2 * kPlanes;
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static constexpr int kStagesFit = (196 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kAccCols = (PASSES == 3 ? 2 : 1) * BN;      // main | cross
  static constexpr int kAccBufs = 2;
  static_assert(kAccBufs * kAccCols <= 512, "two accumulator buffers must fit in TMEM");
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiWarps * 1024;
};

// ---- cluster / cta_group::2 PTX ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {     // shared::cta addr -> shared::cluster addr of `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {           // arrive on a barrier of another CTA of the cluster
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t cta_mask) {     // same barrier offset in every CTA of the mask
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// instruction descriptor: kind::f16, A/B = BF16 K-major, D = F32, M = 256 (pair), N = n
__host__ __device__ constexpr uint32_t make_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int CH, int PASSES>
__device__ __forceinline__ void store_split2(__nv_bfloat16* dst, long long plane, const float (&v)[CH]) {
  __align__(16) __nv_bfloat162 hi[CH / 2], lo[CH / 2];
#pragma unroll
  for (int j = 0; j < CH / 2; ++j) {
    hi[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    if (PASSES == 3) {
      const float2 hf = __bfloat1622float2(hi[j]);
      lo[j] = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
  }
  static_assert(CH % 16 == 0, "an epilogue chunk is a whole number of 32-byte sectors per plane");
  const uint4* h4 = reinterpret_cast<const uint4*>(hi);
  const uint4* l4 = reinterpret_cast<const uint4*>(lo);
#pragma unroll
  for (int j = 0; j < CH / 16; ++j) {                    // dst is 32-byte aligned: channel offsets are multiples of CH >= 16
    tc::st_global_256(dst + 16 * j, h4[2 * j], h4[2 * j + 1]);
    if (PASSES == 3) tc::st_global_256(dst + plane + 16 * j, l4[2 * j], l4[2 * j + 1]);
  }
}

// work item of a PAIR: (phase | n-tile | pair of m-tiles); this CTA's m-tile is 2*mp + rank
struct PairWork {
  int phase, co0, it0, it1;
  int n0, p0, q0, mtile;
  int ks;                       // split-K: which K split of the tile (whole-tile schedule, g.ksplit > 1)
  // stream-K: 0 = whole tile; 1 = contributor (a later part of a tile: raw sums -> this CTA's workspace slot);
  // 2 = finisher (the first part of a tile cut by a pair boundary: adds the parts of pairs pair+1 .. sk_last, in order)
  int sk_role, sk_last;
};

// pair-tile index within a phase -> this CTA's tile
template <int BN>
__device__ __forceinline__ void decode_tile(const TapGemm& g, const Tc2Maps& maps, int phase, int r, int rank, PairWork& wi) {
  const int tiles_q = g.Wg / maps.Wt, tiles_p = g.Hg / maps.Ht;
  const int tiles_m = tiles_q * tiles_p * ((g.n_img + maps.Nt - 1) / maps.Nt);
  const int pairs_m = (tiles_m + 1) / 2;
  wi.phase = phase;
  const int mp = r % pairs_m;
  const int nt = r / pairs_m;
  int mt = 2 * mp + rank;
  wi.mtile = mt;
  if (mt >= tiles_m) {                                  // odd tile count: phantom tile, fully out of range (TMA zero fill)
    wi.n0 = g.n_img; wi.p0 = 0; wi.q0 = 0;
  } else {
    const int qb = mt % tiles_q; mt /= tiles_q;
    const int pb = mt % tiles_p; mt /= tiles_p;
    wi.n0 = mt * maps.Nt; wi.p0 = pb * maps.Ht; wi.q0 = qb * maps.Wt;
  }
  wi.co0 = nt * BN;
}

// Work iteration of one pair.  SK = false: whole pair-tiles, w = pair + i * npairs (longest phases first).
// SK = true (stream-K, same ordered atomic-free scheme as tapgemm_tc.cu's WorkIter): the launch is ONE linear space of T
// K steps over (phase | n-tile | m-pair | K step) and pair c owns steps [T*c/G, T*(c+1)/G): every SM pair gets the same
// tensor work however the tile count divides by 74 and however unequal the phases are.  A tile cut by a pair boundary is
// finished by the pair holding its FIRST K steps (reached at the END of its range) after the pairs holding the later
// steps (which they run FIRST) have published raw partial sums + a release flag.  Both CTAs of a pair walk the same
// schedule; CTA rank r exchanges partial sums with CTA rank r of the neighbouring pairs (same 128 rows of the tile).
template <int BN, bool SK>
struct PairIter {
  int w, total, stride, per_phase;                                  // !SK
  int T, G, cur, end;                                               // SK
  int pair, nchunk;
  __device__ __forceinline__ int iters_of(const TapGemm& g, int ph) const { return g.phase[ph].ntaps * nchunk; }
  __device__ __forceinline__ int boundary(int c) const { return (int)((long long)T * c / G); }
  __device__ __forceinline__ int owner(int gi) const {
    int c = (int)((long long)gi * G / T);
    while (c + 1 < G && boundary(c + 1) <= gi) ++c;
    while (c > 0 && boundary(c) > gi) --c;
    return c;
  }
  __device__ __forceinline__ void init(const TapGemm& g, const Tc2Maps& maps, int total_work, int pair_, int npairs) {
    pair = pair_;
    const int tiles_q = g.Wg / maps.Wt, tiles_p = g.Hg / maps.Ht;
    const int tiles_m = tiles_q * tiles_p * ((g.n_img + maps.Nt - 1) / maps.Nt);
    per_phase = ((tiles_m + 1) / 2) * (g.Cout / BN);
    nchunk = g.Cin / BK;
    if (!SK) { w = pair; total = total_work; stride = npairs; return; }
    T = total_work; G = npairs;
    cur = boundary(pair); end = boundary(pair + 1);
  }
  __device__ __forceinline__ bool next(const TapGemm& g, const Tc2Maps& maps, int rank, PairWork& wi) {
    if (!SK) {
      if (w >= total) return false;
      // (phase | K split | pair-tile): split-K layers (g.ksplit > 1, dense layers with few tiles and a deep K) store raw
      // sums to slab ks of g.ws like the one-CTA kernel; splitk_finalize adds the slabs in split order
      const int per_phase_k = per_phase * g.ksplit;
      const int r = w % per_phase_k;
      decode_tile<BN>(g, maps, w / per_phase_k, r % per_phase, rank, wi);
      const int ip = iters_of(g, wi.phase);
      wi.ks = r / per_phase;
      wi.it0 = (int)((long long)ip * wi.ks / g.ksplit);
      wi.it1 = (int)((long long)ip * (wi.ks + 1) / g.ksplit);
      wi.sk_role = 0; wi.sk_last = 0;
      w += stride;
      return true;
    }
    if (cur >= end) return false;
    int gi = cur, ph = 0, phase_start = 0;
    while (ph + 1 < g.nphase && gi >= phase_start + per_phase * iters_of(g, ph)) { phase_start += per_phase * iters_of(g, ph); ++ph; }
    const int ip = iters_of(g, ph);
    const int tile = (gi - phase_start) / ip, it = (gi - phase_start) % ip;
    int len = ip - it;
    if (len > end - gi) len = end - gi;
    decode_tile<BN>(g, maps, ph, tile, rank, wi);
    wi.ks = 0;
    wi.it0 = it; wi.it1 = it + len;
    wi.sk_role = it > 0 ? 1 : (len < ip ? 2 : 0);
    wi.sk_last = wi.sk_role == 2 ? owner(phase_start + tile * ip + ip - 1) : 0;
    cur = gi + len;
    return true;
  }
};

template <int BN, int PASSES, bool SK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tapgemm_tc2_kernel(const __grid_constant__ TapGemm g, const __grid_constant__ Tc2Maps maps, const int total_work) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  constexpr int kATileBytes = Cfg::kATileBytes;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in both CTAs of the pair, so every address below is pair-symmetric
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                    // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };             // one per CTA (multicast commit)
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };         // one per CTA (multicast commit)
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };    // used in the leader only
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  const uint32_t stage_smem = bar_base + 256u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_al + (tmem_slot - smem_base));
  float* stage_ptr = reinterpret_cast<float*>(smem_al + (stage_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    pdl_trigger();                                      // the next kernel of the chain may move in as this grid's CTAs retire
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);                        // the leader's arrive.expect_tx; bytes come from both CTAs' TMA
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 2 * kEpiWarps);          // epilogue warps of BOTH CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem2_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // peer barriers initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nchunk = g.Cin / BK;
  pdl_wait();                                           // prologue done; everything below reads / writes activations (tapgemm.h: PDL)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; whole warp in uniform control flow, one elected lane issues) ==========
    {
      uint32_t i = 0;
      PairIter<BN, SK> iter;
      iter.init(g, maps, total_work, pair, npairs);
      PairWork wi;
      while (iter.next(g, maps, (int)rank, wi)) {
        const Phase ph = g.phase[wi.phase];
        for (int it = wi.it0; it < wi.it1; ++it, ++i) {
          const int s = i % S;
          const uint32_t par = (i / S) & 1u;
          const Tap tap = g.taps[ph.tap_begin + it / nchunk];
          const int c0 = (it % nchunk) * BK;
          mbar_wait(empty_bar(s), par ^ 1u);            // the pair's MMAs that read this stage (in BOTH CTAs) have retired
          const uint32_t lead_full = mapa_u32(full_bar(s), 0);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);
            tma2_load_5d(PASSES == 3 ? &maps.a[tap.view] : &maps.a1[tap.view], lead_full, sa, c0, wi.q0 + tap.dw, wi.p0 + tap.dh,
                         wi.n0, 0);
            tma2_load_3d(PASSES == 3 ? &maps.b : &maps.b1, lead_full, sa + kATileBytes, c0,
                         tap.wtile * g.Cout + wi.co0 + (int)rank * (BN / 2), 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The WHOLE warp walks the schedule and waits on the barriers (uniform control flow: descriptors live in uniform
    // registers); one elected lane issues the MMAs and commits.  At N = 128 a UMMA executes in 64 clocks, so the issue cost
    // per instruction is on the critical path of this kernel.
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_m256(BN);
      uint32_t i = 0, t = 0;
      PairIter<BN, SK> iter;
      iter.init(g, maps, total_work, pair, npairs);
      PairWork wi;
      for (; iter.next(g, maps, 0, wi); ++t) {
        const uint32_t buf = t & 1u, use = t >> 1;
        const uint32_t acc_main = tmem_base + buf * Cfg::kAccCols, acc_cross = acc_main + BN;
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);    // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        for (int it = wi.it0; it < wi.it1; ++it, ++i) {
          const int s = i % S;
          const uint32_t par = (i / S) & 1u;
          mbar_wait(full_bar(s), par);                  // A and B halves of both CTAs have landed
          tc_fence_after();
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + BM * BK * 2);
          const uint64_t b_hi = make_sw128_desc(sa + kATileBytes), b_lo = make_sw128_desc(sa + kATileBytes + (BN / 2) * BK * 2);
          const uint32_t first = (it > wi.it0) ? 1u : 0u;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);    // 32 bytes per K=16 slice, in 16-byte units
              const uint32_t acc = (k > 0) ? 1u : first;
              umma2_bf16(acc_main, a_hi + ko, b_hi + ko, idesc, acc);
              if (PASSES == 3) {
                umma2_bf16(acc_cross, a_lo + ko, b_hi + ko, idesc, acc);
                umma2_bf16(acc_cross, a_hi + ko, b_lo + ko, idesc, 1u);
              }
            }
            umma2_commit_mc(empty_bar(s), 3);           // frees the stage in both CTAs when these MMAs retire
          }
          __syncwarp();
        }
        if (elect_one_sync()) umma2_commit_mc(tfull_bar(buf), 3);   // accumulators complete: both CTAs' epilogues may read
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs, own 128 rows) =====================
    constexpr int CH = 32;
    constexpr int COLS_PER_WARP = BN / 2;
    const int ew = warp - 2;
    const int lg = warp & 3;                            // TMEM lane group this warp may access
    const int half = ew >> 2;                           // which half of the tile's columns
    float* my_stage = stage_ptr + ew * 256;             // [0,128): scale, [128,256): shift of this warp's columns
    const float act_a = g.act == ACT_LRELU ? 0.6f : g.act == ACT_RELU ? 0.5f : 1.f;
    const float act_b = g.act == ACT_LRELU ? 0.4f : g.act == ACT_RELU ? 0.5f : 0.f;
    const int ml = lg * 32 + lane;                      // tile row
    const int wl = ml % maps.Wt;
    const int hl = (ml / maps.Wt) % maps.Ht;
    const int nl = ml / (maps.Wt * maps.Ht);
    const uint32_t lead_tempty[2] = {mapa_u32(tempty_bar(0), 0), mapa_u32(tempty_bar(1), 0)};
    uint32_t t = 0;
    const int my_cta = 2 * pair + (int)rank;            // workspace / flag slot of this CTA
    PairIter<BN, SK> iter;
    iter.init(g, maps, total_work, pair, npairs);
    PairWork wi;
    for (; iter.next(g, maps, (int)rank, wi); ++t) {
      const Phase ph = g.phase[wi.phase];
      const uint32_t buf = t & 1u, use = t >> 1;
      if (g.scale_pix_stride == 0) {                    // stage this tile's per-channel scale/shift while the MMAs run
        __syncwarp();
        const int cbase = wi.co0 + half * COLS_PER_WARP;
        for (int c = lane; c < COLS_PER_WARP; c += 32) {
          my_stage[c] = g.scale ? __ldg(g.scale + cbase + c) : 1.f;
          my_stage[128 + c] = g.shift ? __ldg(g.shift + cbase + c) : 0.f;
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(buf), use & 1u);
      tc_fence_after();
      if (SK && wi.sk_role == 2) {                      // finisher: the later parts were computed first; wait for them
        for (int k = pair + 1 + lane; k <= wi.sk_last; k += 32) {
          const int* fl = g.sk_flags + (2 * k + (int)rank) * kEpiWarps + ew;
          const long long t0 = clock64();
          int fv;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(fv) : "l"(fl) : "memory");
            if (clock64() - t0 > 4000000000LL) __trap();
          } while (fv != g.sk_epoch);
        }
        __syncwarp();
      }
      const int n = wi.n0 + nl, p = wi.p0 + hl, q = wi.q0 + wl;
      const bool valid = n < g.n_img;
      const int oh = p * g.osh + ph.oh0, ow = q * g.osw + ph.ow0;
      const long long pix = (long long)(n * g.Hout + oh) * g.Wout + ow;
      const uint32_t lane_addr = tmem_base + buf * Cfg::kAccCols + ((uint32_t)(lg * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < COLS_PER_WARP; cc += CH) {
        const int cb = half * COLS_PER_WARP + cc;
        const int co = wi.co0 + cb;
        float v[CH];
        __syncwarp();                                   // tcgen05.ld is .aligned: reconverge first
        if (PASSES == 3) {
          uint32_t vm[CH], vc[CH];
          tmem_ld<CH>(lane_addr + cb, vm);
          tmem_ld<CH>(lane_addr + BN + cb, vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(vm[j]) + __uint_as_float(vc[j]);
        } else {
          uint32_t vm[CH];
          tmem_ld<CH>(lane_addr + cb, vm);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(vm[j]);
        }
        if (cc + CH >= COLS_PER_WARP) {                  // last TMEM read of this work item: release the buffer to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(tempty_bar(buf)); else mbar_arrive_cluster(lead_tempty[buf]);
          }
        }
        if (SK && wi.sk_role == 1) {                     // contributor: raw partial sums -> this CTA's workspace slot
          float4* wp = reinterpret_cast<float4*>(g.sk_ws + (((long long)my_cta * kEpiWarps + ew) * 32 + lane) * COLS_PER_WARP + cc);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) __stcg(wp + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          continue;
        }
        if (SK && wi.sk_role == 2) {                     // finisher: add the later parts, in pair order
          for (int k = pair + 1; k <= wi.sk_last; ++k) {
            const float4* rp = reinterpret_cast<const float4*>(g.sk_ws + (((long long)(2 * k + (int)rank) * kEpiWarps + ew) * 32 + lane) * COLS_PER_WARP + cc);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) {
              const float4 a = __ldcg(rp + j);
              v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
            }
          }
        }
        if (!SK && g.ksplit > 1) {                       // this K split's slab; the finalize kernel adds them in order
          if (valid) {
            float4* wsp = reinterpret_cast<float4*>(g.ws + (long long)wi.ks * g.ws_slab + pix * g.Cout + co);
#pragma unroll
            for (int j = 0; j < CH / 4; ++j) __stcg(wsp + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
          continue;
        }
        if (!valid) continue;
        const long long off = pix * g.Cout + co;
        if (g.out_raw) store_split2<CH, PASSES>(g.out_raw + off, g.out_raw_plane, v);   // pre-BN value (MDBLOCK residual input)
        if (g.res && !g.res_after) {                                     // residual add before BatchNorm (MDBLOCK, layers.py:411-416)
          const uint4* rh = reinterpret_cast<const uint4*>(g.res + off);
          const uint4* rl = reinterpret_cast<const uint4*>(g.res + g.res_plane + off);
#pragma unroll
          for (int j8 = 0; j8 < CH / 8; ++j8) {
            const uint4 h4 = __ldg(rh + j8);
            const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h4);
            if (PASSES == 3) {
              const uint4 l4 = __ldg(rl + j8);
              const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l4);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += __bfloat162float(hb[j]) + __bfloat162float(lb[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += __bfloat162float(hb[j]);
            }
          }
        }
        if (g.act == ACT_MASK) {
          const int si = co + (oh * g.Wout + ow) * g.scale_pix_stride;
          const uint4* mk = reinterpret_cast<const uint4*>(g.mask + off);
#pragma unroll
          for (int j8 = 0; j8 < CH / 8; ++j8) {
            const uint4 m4 = __ldg(mk + j8);
            const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float sc = g.scale_pix_stride ? __ldg(g.scale + si + j8 * 8 + j) : my_stage[cc + j8 * 8 + j];
              v[j8 * 8 + j] = v[j8 * 8 + j] * sc * (__bfloat162float(mb[j]) > 0.f ? 1.f : g.mask_slope);
            }
          }
          if (g.res && g.res_after) {                    // gradient of the block's residual branch joins after the mask/scale
            const uint4* rh = reinterpret_cast<const uint4*>(g.res + off);
            const uint4* rl = reinterpret_cast<const uint4*>(g.res + g.res_plane + off);
#pragma unroll
            for (int j8 = 0; j8 < CH / 8; ++j8) {
              const uint4 h4 = __ldg(rh + j8);
              const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h4);
              if (PASSES == 3) {
                const uint4 l4 = __ldg(rl + j8);
                const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l4);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += __bfloat162float(hb[j]) + __bfloat162float(lb[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j8 * 8 + j] += __bfloat162float(hb[j]);
              }
            }
          }
        } else {
#pragma unroll
          for (int j4 = 0; j4 < CH / 4; ++j4) {
            const float4 sc = *reinterpret_cast<const float4*>(my_stage + cc + 4 * j4);
            const float4 sf = *reinterpret_cast<const float4*>(my_stage + 128 + cc + 4 * j4);
            v[4 * j4 + 0] = fmaf(v[4 * j4 + 0], sc.x, sf.x);
            v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], sc.y, sf.y);
            v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], sc.z, sf.z);
            v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], sc.w, sf.w);
          }
          if (g.act == ACT_ELU) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = v[j] > 0.f ? v[j] : expm1f(v[j]);
          } else if (g.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaf(act_b, fabsf(v[j]), act_a * v[j]);
          }
        }
        if (g.out) store_split2<CH, PASSES>(g.out + off, g.out_plane, v);
        if (g.out_f32) {
          float4* of = reinterpret_cast<float4*>(g.out_f32 + off);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) of[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      if (SK && wi.sk_role == 1) {                       // publish this warp's sub-block of partial sums
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          int* fl = g.sk_flags + my_cta * kEpiWarps + ew;
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(fl), "r"(g.sk_epoch) : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // the peer may still read this CTA's smem / signal its barriers
  if (warp == 1) tmem2_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace

Tc2Maps* tc2_build_maps(const TapGemm& g, char* err, int errlen) {
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  if (g.Cin % 64 || g.Cout % 128) { snprintf(err, errlen, "2-CTA path needs Cin%%64==0 and Cout%%128==0 (got %d,%d)", g.Cin, g.Cout); return nullptr; }
  Tc2Maps* m = new Tc2Maps();
  memset(m, 0, sizeof(*m));
  tile_shape(g.Hg, g.Wg, m->Wt, m->Ht, m->Nt);
  m->BN = 128;
  m->BN1 = (g.Cout % 256 == 0) ? 256 : 128;
  m->sk_choice[0] = m->sk_choice[1] = -1;
  if (g.Wg % m->Wt || g.Hg % m->Ht || m->Wt * m->Ht * m->Nt != BM) {
    snprintf(err, errlen, "M grid %dx%d does not tile into 128-row boxes", g.Hg, g.Wg);
    delete m; return nullptr;
  }
  bool used[4] = {false, false, false, false};
  int max_tile = 0;
  for (int p = 0; p < g.nphase; ++p)
    for (int t = 0; t < g.phase[p].ntaps; ++t) {
      const Tap& tp = g.taps[g.phase[p].tap_begin + t];
      used[tp.view] = true;
      if (tp.wtile > max_tile) max_tile = tp.wtile;
    }
  for (int v = 0; v < 4; ++v) {
    if (!used[v]) continue;
    const int vh = v >> 1, vw = v & 1;
    const cuuint64_t Hv = (g.Hin - vh + g.sh - 1) / g.sh, Wv = (g.Win - vw + g.sw - 1) / g.sw;
    cuuint64_t dims[5] = {(cuuint64_t)g.Cin, Wv, Hv, (cuuint64_t)g.n_img, 2};
    cuuint64_t strides[4] = {(cuuint64_t)g.sw * g.Cin * 2, (cuuint64_t)g.sh * g.Win * g.Cin * 2,
                             (cuuint64_t)g.Hin * g.Win * g.Cin * 2, (cuuint64_t)g.a_plane * 2};
    cuuint32_t box[5] = {BK, (cuuint32_t)m->Wt, (cuuint32_t)m->Ht, (cuuint32_t)m->Nt, 2};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = (void*)(g.a + ((long long)vh * g.Win + vw) * g.Cin);
    CUresult r = enc(&m->a[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(A view %d) failed: %d", v, (int)r); delete m; return nullptr; }
    box[4] = 1;
    r = enc(&m->a1[v], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(A1 view %d) failed: %d", v, (int)r); delete m; return nullptr; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.Cin, (cuuint64_t)(max_tile + 1) * g.Cout, 2};
    cuuint64_t strides[2] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.b_plane * 2};
    cuuint32_t box[3] = {BK, (cuuint32_t)(m->BN / 2), 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)g.b, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(B) failed: %d", (int)r); delete m; return nullptr; }
    box[1] = (cuuint32_t)(m->BN1 / 2);
    box[2] = 1;
    r = enc(&m->b1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)g.b, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(B1) failed: %d", (int)r); delete m; return nullptr; }
  }
  return m;
}

void tc2_free_maps(Tc2Maps* m) { delete m; }

// pair-tiles of this launch, and whether the pair kernel should take it (enough whole tiles to fill the 74 pairs twice)
static long long pair_tiles_bn(const TapGemm& g, const Tc2Maps* maps, int bn) {
  const int tiles_m = (g.Wg / maps->Wt) * (g.Hg / maps->Ht) * ((g.n_img + maps->Nt - 1) / maps->Nt);
  return (long long)((tiles_m + 1) / 2) * (g.Cout / bn) * g.nphase;
}
long long tc2_pair_tiles(const TapGemm& g, const Tc2Maps* maps) { return pair_tiles_bn(g, maps, maps->BN); }

template <int BN, int PASSES, bool SK>
static int launch_pair(const TapGemm& g, const Tc2Maps* maps, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, PASSES>;
  static DeviceOnce attr_set;
  const int dev = cur_device();
  if (!attr_set.is_done(dev)) {
    if (cudaFuncSetAttribute(tapgemm_tc2_kernel<BN, PASSES, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return -1;
    attr_set.set_done(dev);
  }
  const int tiles = (int)pair_tiles_bn(g, maps, BN) * (SK ? 1 : g.ksplit);
  const int pairs_hw = tc_num_sms() / 2;
  int total_work = tiles, pairs = tiles < pairs_hw ? tiles : pairs_hw;
  if (SK) {                                              // T K-steps, one pair per SM pair, every pair gets T/G of them
    const int per_phase = tiles / g.nphase;
    long long T = 0;
    for (int p = 0; p < g.nphase; ++p) T += (long long)per_phase * g.phase[p].ntaps * (g.Cin / BK);
    total_work = (int)T;
    pairs = pairs_hw;
  }
  if (launch_pdl(tapgemm_tc2_kernel<BN, PASSES, SK>, dim3(2 * pairs), dim3(kThreads), Cfg::kSmemBytes, st, g, *maps, total_work) != cudaSuccess)
    return -1;
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

static bool want_streamk_uncached(const TapGemm& g, const Tc2Maps* maps, int bn);
// stream-K pays one partial-sum round trip per pair boundary; it is used where the static whole-tile schedule would leave
// >= 8 % of the pair-time idle (makespan test on the host, cached per layer).
static bool want_streamk(const TapGemm& g, const Tc2Maps* maps, int bn) {
  if (!g.sk_ws) return false;
  int& cached = maps->sk_choice[g.passes == 1 ? 1 : 0];
  if (cached >= 0) return cached == 1;
  cached = want_streamk_uncached(g, maps, bn) ? 1 : 0;
  return cached == 1;
}
static bool want_streamk_uncached(const TapGemm& g, const Tc2Maps* maps, int bn) {
  const int tiles = (int)pair_tiles_bn(g, maps, bn);
  const int G = tc_num_sms() / 2;
  if (tiles < G / 2) return false;
  // batches that the host API replays as CUDA graphs (<= 32 images; stream-K's epoch is a kernel argument, so captured
  // launches run whole tiles) keep ONE schedule in both forms; and a tile needs a K loop worth cutting (the dense layers
  // have 2-16 K steps: a partial-sum round trip costs more than it balances)
  if (g.n_img <= 32) return false;
  for (int p = 0; p < g.nphase; ++p)
    if (g.phase[p].ntaps * (g.Cin / BK) < 24) return false;
  const int per_phase = tiles / g.nphase;
  long long T = 0, makespan = 0;
  std::vector<long long> load(G, 0);
  for (int w = 0; w < tiles; ++w) {
    const long long it = (long long)g.phase[w / per_phase].ntaps * (g.Cin / BK);
    load[w % G] += it;
    T += it;
  }
  for (long long v : load) makespan = v > makespan ? v : makespan;
  return makespan * G * 100 >= T * 108;
}

int launch_tapgemm_tc2(const TapGemm& g, const Tc2Maps* maps, cudaStream_t st) {
  if (g.ksplit < 1 || g.out_f32_t) return -1;
  if (g.ksplit > 1) {                                    // split-K: whole-tile schedule over (K split | pair-tile), float32 mode
    if (g.passes != 3 || !g.ws) return -1;
    return launch_pair<128, 3, false>(g, maps, st);
  }
  if (g.passes == 1) {
    if (maps->BN1 != 256) return -1;                      // Cout = 128 in bf16 mode: one-CTA kernel
    return want_streamk(g, maps, 256) ? launch_pair<256, 1, true>(g, maps, st) : launch_pair<256, 1, false>(g, maps, st);
  }
  return want_streamk(g, maps, 128) ? launch_pair<128, 3, true>(g, maps, st) : launch_pair<128, 3, false>(g, maps, st);
}

}  // namespace ian
