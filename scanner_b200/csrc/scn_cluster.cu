// scn_cluster.cu -- single-pass fused spectrum-sense kernel for N = 2^14, 2^15, 2^16 (thread-block clusters + DSMEM).
//
// One transform no longer fits one CTA's shared memory beyond 2^14 points (2^16 complex fp32 = 512 KB), so a CLUSTER of
// C = N / 16384 CTAs (1, 2 or 4 SMs) owns it: the transform lives in the cluster's distributed shared memory from the
// moment the raw samples are loaded to the moment dB and detections leave -- every sample crosses HBM exactly once,
// with no intermediate in global memory (the four-step path of scn_large.cu needed ~28 B/sample of HBM/L2 round trips).
//
// Decomposition (decimation in frequency): R = 4 C rows of M = 4096 points,
//     n = R n2 + n1  (n1 < R: row, n2 < M),        k = k2 + M k1  (k2 < M, k1 < R)
//     X[k2 + M k1] = sum_n1 W_R^(n1 k1) * W_N^(n1 k2) * A[n1][k2],    A[n1][k2] = sum_n2 W_M^(n2 k2) x[R n2 + n1]
//   phase 0  CTA c loads rows n1 = 4c .. 4c+3 (4 consecutive samples out of every R: 32-byte sectors for fp32 IQ),
//            [DC: int32 sums, exchanged over DSMEM -- the integer sum is exact in any order], convert + window
//            (utility.cpp:9-84, process.cpp:28-34), into its shared memory, one 4096-point row each;
//   phase 1  four 4096-point row FFTs, 64 threads per row, 64 x 64 with one in-place exchange (the scn_p64.cuh plan),
//            the inter-stage twiddle W_N^(n1 k2) = W_N^(n1 t) * W_N^(64 n1 q) folded in (first factor into the row FFT's
//            own twiddles, second one multiply per output);
//   exchange CTA c'' owns k2 in [c'' M/C, (c''+1) M/C): every thread stores its 64 outputs into the owner's shared
//            memory (st.async with mbarrier completion; 3/4 of them cross SMs when C = 4) as S[n1][k2 local] -- the
//            row buffers are dead by then and are reused as S;
//   phase 2  per k2 one radix-R DFT over n1 (R = 4, 8, 16) -> X[k2 + M k1], |X|^2, K-average, dB (utility.cpp:86-98),
//            detection (process.cpp:46-61).  For fixed k1 a warp holds 32 consecutive bins: coalesced stores, one
//            mask word per ballot.
//   epilogue hit records must be ordered over the WHOLE spectrum: per-(k1, CTA) hit counts are exchanged over DSMEM
//            and prefix-summed in shifted-bin order.
// Synchronisation inside the loop is mbarriers only (rows read / rows arrived / DC sums / hit-count table), see st_async
// below; cluster barriers only after mbarrier init and before exit.
// Same results contract as every other size (tests/test_gpu_large.py, tests/test_gpu_parity.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "scn_dispatch.h"
#include "scn_wpt.cuh"

#ifndef SCN_CL_PRE_F32
#define SCN_CL_PRE_F32 16      // fp32 IQ: how many of the 32 raw loads per thread are issued one buffer ahead
#endif

namespace scn {

constexpr int kClM = 4096;                       // row length
constexpr int kClRows = 4;                       // rows per CTA
constexpr int kClThreads = 256;                  // 64 per row
constexpr int kClRowElems = kClM + kClM / 64 + 4; // padded row (one float2 per 64) for the in-place 64 x 64 exchange, + 4
                                                 // so that rows r and r + 2 sit 16 banks apart (phase-0 stores of a lane
                                                 // pair go to rows {0,2} / {1,3} of the same column)
constexpr int kClMaskWords = 512;                // mask words a CTA owns: R * (M/C/32) = 512 for every C

struct ClusterSmem {
  float2 rows[kClRows * kClRowElems];            // phase 0/1: four padded rows; phase 2: S[R][M/C] (131 072 B fit)
  uint32_t mask[kClMaskWords];                   // [k1][word]
  uint32_t prefix[kClMaskWords];                 // hits before this word inside the (k1, this CTA) range
  uint32_t table[16 * 4];                        // [k1][cta] hit counts of the whole cluster
  uint32_t base[16];                             // rank of the first hit of (k1, this CTA)
  int32_t dcsum[2 * 4];                          // [cta][I, Q]
  int32_t red[2 * 8];                            // per-warp partial sums
  uint32_t total;
  uint64_t bar[4];                               // mbarriers: kBarReady, kBarRows, kBarDc, kBarTable
};
enum { kBarReady = 0, kBarRows = 1, kBarDc = 2, kBarTable = 3 };

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_shared(const void* p, uint32_t rank) {       // address of `p` in CTA `rank`'s smem
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
  return out;
}
// Remote stores are ASYNC stores that count their bytes on an mbarrier in the destination CTA (STAS): the receiver
// waits for its own mbarrier phase, so no release/acquire cluster barrier is needed around an exchange.  That matters:
// barrier.cluster.arrive.release compiles to MEMBAR.ALL.GPU (waits for every outstanding GLOBAL store of the CTA, i.e.
// the spectrum rows just written) and barrier.cluster.wait to CCTL.IVALL (drops the L1 lines holding window and
// twiddles) -- together 10-15 % of all stall samples of the first version of this kernel (profiles/r02m_ncu_cl16_*).
__device__ __forceinline__ void st_async(uint32_t addr, float2 v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(addr), "f"(v.x), "f"(v.y), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async(uint32_t addr, uint32_t v, uint32_t bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar) : "memory");
}
// arrive on a peer's mbarrier without a fence: orders nothing but "this warp got here" (used for "done READING")
__device__ __forceinline__ void mbar_arrive_peer_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

// twiddle tables (host: scn_api.cu, layout 3):
//   twA[(r-1)*64 + t]  = exp(-2 pi i t r / 4096), r = 1..63           (row FFT, as scn_p64.cuh)
//   twC[n1*64 + t]     = exp(-2 pi i n1 t / N)                        (thread constant, folded into the row FFT)
//   twB[n1*64 + q]     = exp(-2 pi i 64 n1 q / N)                     (per output)
constexpr int kClTwA = 63 * 64;
__host__ __device__ constexpr int cluster_twiddle_elems(int R) { return kClTwA + 2 * R * 64; }

template <int C, int KIND, bool DC, bool AVG>
__global__ void __launch_bounds__(kClThreads, 1) spectrum_sense_cluster_kernel(const KernelParams p) {
  constexpr int R = 4 * C, M = kClM, N = R * M, LOG2N = (C == 1 ? 14 : C == 2 ? 15 : 16);
  constexpr int KL = M / C;                        // k2 values this CTA owns
  constexpr int U = KL / kClThreads;               // k2 per thread in phase 2 (16 / 8 / 4)
  constexpr int WK = KL / 32;                      // mask words per k1 in this CTA
  constexpr bool kInt = KindTraits<KIND>::kInt;
  constexpr bool kDC = DC && kInt;
  constexpr int kBytes = KindTraits<KIND>::kBytes;
  static_assert(U * R == 64, "64 outputs per thread in phase 2");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (the warp index used for mask-word addressing is computed behind an asm so that the compiler cannot tie it to
  //  `lane == 0` facts: with plain tid >> 5 the address of a later, unpredicated read lost its "& ~3" -- a misaligned
  //  shared read caught by compute-sanitizer)
  uint32_t wslot;
  asm("shr.u32 %0, %1, 5;" : "=r"(wslot) : "r"(uint32_t(tid)));
  const int j = tid >> 6, t = tid & 63;            // phase 1: row j of this CTA, column t
  const uint32_t crank = (C == 1) ? 0u : cluster_rank();
  const uint32_t n1 = 4u * crank + uint32_t(j);    // global row index of this thread's row
  const uint32_t half = N / 2;
  const uint32_t K = AVG ? p.averaging : 1u;
  const uint32_t cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
  const float2* twA = p.twiddles;
  const float2* twC = p.twiddles + kClTwA;
  const float2* twB = twC + R * 64;
  float2* rows = sm.rows;

  // ---- raw samples of one buffer, as loaded.  This CTA's four rows are samples R n2 + 4 crank + {0..3}; a lane PAIR
  // (tid even / odd) takes one n2 = tid / 2 + 128 u, u < 32: the even lane rows 0,1, the odd lane rows 2,3, so one warp-
  // wide load covers 16 whole 32-byte groups (16 B per lane for fp32 IQ, 8 B for int16 IQ, 4 B for int8 IQ, 2 x 4 B for
  // split int16) and every 128-byte line is touched once.  The loads are issued one buffer AHEAD, at the start of
  // phase 2 (whose register need is small), so their HBM latency hides behind phase 2 and the epilogue; fp32 IQ keeps
  // only half of them in registers that long (K > 1: none, 64 accumulators), the rest loads in place.
  constexpr int kRawWords = KIND == SCN_KIND_FLOAT_COMPLEX ? 4 : KIND == SCN_KIND_BYTE_COMPLEX ? 1 : 2;
  // (int16 IQ with K > 1 -- 64 accumulators live across the row FFTs -- spilled with all 32 in registers: 16 is +7 %,
  //  profiles/r02zo_cluster_int_prefetch_depth_ab.txt; for int8 IQ and K = 1 all 32 measured best)
  constexpr int kPre = KIND == SCN_KIND_FLOAT_COMPLEX ? (AVG ? 0 : SCN_CL_PRE_F32)
                     : (KIND == SCN_KIND_SHORT_COMPLEX && AVG) ? 16 : 32;
  const uint32_t lhalf = uint32_t(tid) & 1u, n2base = uint32_t(tid) >> 1;
  uint32_t rawv[32][kRawWords];
  auto load_raw = [&](const uint8_t* buf, int u_begin, int u_end) {
#pragma unroll
    for (int u = 0; u < 32; u++) {
      if (u < u_begin || u >= u_end) continue;
      const size_t e = size_t(R) * (n2base + 128u * u) + 4u * crank + 2u * lhalf;   // first of this lane's 2 samples
      if constexpr (KIND == SCN_KIND_FLOAT_COMPLEX) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(buf) + e / 2);
        rawv[u][0] = a.x; rawv[u][1] = a.y; rawv[u][2] = a.z; rawv[u][3] = a.w;
      } else if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
        rawv[u][0] = __ldg(reinterpret_cast<const uint32_t*>(buf) + e / 2);
      } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(buf) + e / 2);
        rawv[u][0] = v.x; rawv[u][1] = v.y;
      } else {                                                               // split: I block, then Q block
        rawv[u][0] = __ldg(reinterpret_cast<const uint32_t*>(buf) + e / 2);
        rawv[u][1] = __ldg(reinterpret_cast<const uint32_t*>(buf + size_t(N) * 2) + e / 2);
      }
    }
  };
  // int32 sums of I and Q over this thread's 64 samples (utility.cpp:20-24,44-48,72-76), packed dot products
  auto raw_sums = [&](int& si, int& sq) {
    si = 0; sq = 0;
#pragma unroll
    for (int u = 0; u < 32; u++) {
      if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
        si = __dp4a(int(rawv[u][0]), 0x00010001, si); sq = __dp4a(int(rawv[u][0]), 0x01000100, sq);
      } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
#pragma unroll
        for (int x = 0; x < 2; x++) { si = __dp2a_lo(int(rawv[u][x]), 0x00000001, si); sq = __dp2a_lo(int(rawv[u][x]), 0x00000100, sq); }
      } else if constexpr (KIND == SCN_KIND_SHORT) {
        si = __dp2a_lo(int(rawv[u][0]), 0x00000101, si); sq = __dp2a_lo(int(rawv[u][1]), 0x00000101, sq);
      }
    }
  };
  // convert + scale + window (utility.cpp:27-30,52-55,80-83; process.cpp:28-34) into rows 2 lhalf and 2 lhalf + 1
  auto store_rows = [&](int dci, int dcq) {
    const float2* win = reinterpret_cast<const float2*>(p.window);
    float2* r0 = rows + (2u * lhalf) * kClRowElems;
#pragma unroll
    for (int u = 0; u < 32; u++) {
      const uint32_t n2 = n2base + 128u * u;
      const float2 w = __ldg(win + (size_t(R) * n2 + 4u * crank + 2u * lhalf) / 2);   // taps pre-scaled by 1/max (exact)
      const float ws[2] = {w.x, w.y};
#pragma unroll
      for (int x = 0; x < 2; x++) {
        float2 val;
        if constexpr (KIND == SCN_KIND_FLOAT_COMPLEX) {
          val = __fmul2_rn(make_float2(__uint_as_float(rawv[u][2 * x]), __uint_as_float(rawv[u][2 * x + 1])),
                           make_float2(ws[x], ws[x]));
        } else {
          int xi, xq;
          if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
            const uint32_t h = rawv[u][0] >> (16 * x);
            xi = int(static_cast<signed char>(h & 0xff));
            xq = int(static_cast<signed char>((h >> 8) & 0xff));
          } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
            xi = int(static_cast<short>(rawv[u][x] & 0xffff));
            xq = int(static_cast<short>(rawv[u][x] >> 16));
          } else {
            xi = int(static_cast<short>((rawv[u][0] >> (16 * x)) & 0xffff));
            xq = int(static_cast<short>((rawv[u][1] >> (16 * x)) & 0xffff));
          }
          val = make_float2(__fmul_rn(float(xi - dci), ws[x]), __fmul_rn(float(xq - dcq), ws[x]));
        }
        r0[x * kClRowElems + n2] = val;
      }
    }
  };
  auto buffer_ptr = [&](uint32_t s, uint32_t k) { return p.raw + (size_t(s) * K + k) * size_t(N) * kBytes; };

  // candidate bins of this thread (process.cpp:46-53), fixed for the whole launch: slot u R + k1 <-> bin
  // crank KL + tid + 256 u + M k1
  uint32_t cand_lo = 0, cand_hi = 0;
#pragma unroll
  for (int slot = 0; slot < 64; slot++) {
    const uint32_t bin = crank * KL + uint32_t(tid) + 256u * (slot / R) + uint32_t(M) * (slot % R), i = bin ^ half;
    const bool cand = !(bin < p.dc_ignore || (N - bin) < p.dc_ignore) && !(i < (half - p.use_window) || i > (half + p.use_window));
    if (slot < 32) cand_lo |= (cand ? 1u : 0u) << slot; else cand_hi |= (cand ? 1u : 0u) << (slot - 32);
  }

  uint32_t ph_ready = 0, ph_rows = 0, ph_dc = 0, ph_table = 0;      // mbarrier phase parities
  if constexpr (C > 1) {
    if (tid == 0) {
      mbar_init(&sm.bar[kBarReady], 8u * C);         // one arrival per warp of every CTA
      mbar_init(&sm.bar[kBarRows], 1);               // thread 0's expect_tx + the bytes of the other CTAs
      mbar_init(&sm.bar[kBarDc], 1);
      mbar_init(&sm.bar[kBarTable], 1);
    }
    cluster_arrive();                                // every CTA's mbarriers exist before anyone signals them
    cluster_wait();
  }
  float acc[AVG ? 64 : 1];
  if (cluster_id < p.n_spectra && kPre > 0) load_raw(buffer_ptr(cluster_id, 0), 0, kPre);
  for (uint32_t s = cluster_id; s < p.n_spectra; s += n_clusters) {
    for (uint32_t k = 0; k < K; k++) {
      // ---- phase 0: [DC], convert + window into this CTA's four rows -------------------------------------------------
      if constexpr (kPre < 32) load_raw(buffer_ptr(s, k), kPre, 32);
      int dci = 0, dcq = 0;
      if constexpr (kDC) {
        int si, sq;
        raw_sums(si, sq);
        si = __reduce_add_sync(0xffffffffu, si);
        sq = __reduce_add_sync(0xffffffffu, sq);
        if (lane == 0) { sm.red[2 * warp] = si; sm.red[2 * warp + 1] = sq; }
        __syncthreads();
        if (tid < C) {                                                     // this CTA's sums -> every CTA's table
          int ti = 0, tq = 0;
#pragma unroll
          for (int w = 0; w < 8; w++) { ti += sm.red[2 * w]; tq += sm.red[2 * w + 1]; }
          if constexpr (C == 1) {
            sm.dcsum[0] = ti; sm.dcsum[1] = tq;
          } else {
            if (tid == 0) mbar_expect_tx(&sm.bar[kBarDc], 8u * C);         // two ints from each CTA (this one included)
            const uint32_t rbar = map_shared(&sm.bar[kBarDc], uint32_t(tid));
            st_async(map_shared(&sm.dcsum[2 * crank], uint32_t(tid)), uint32_t(ti), rbar);
            st_async(map_shared(&sm.dcsum[2 * crank + 1], uint32_t(tid)), uint32_t(tq), rbar);
          }
        }
        if constexpr (C == 1) __syncthreads();
        else { mbar_wait(&sm.bar[kBarDc], ph_dc); ph_dc ^= 1u; }
        int ti = 0, tq = 0;
#pragma unroll
        for (int c = 0; c < C; c++) { ti += sm.dcsum[2 * c]; tq += sm.dcsum[2 * c + 1]; }
        dci = int(unsigned(ti) >> LOG2N);                                  // unsigned division by N (utility.cpp:25-26,49-50,77-78)
        dcq = int(unsigned(tq) >> LOG2N);
      }
      store_rows(dci, dcq);
      __syncthreads();

      // ---- phase 1: 4096-point FFT of row j (64 threads), in place ---------------------------------------------
      float2 v[64];
      float2* row = rows + j * kClRowElems;
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = row[t + 64 * r];
      __syncthreads();                                   // every column is in registers: the row may be scattered into
      dft64_inplace(v);
      {
        float2* base = row + 65 * t;
#pragma unroll
        for (int x = 0; x < 64; x++) base[dft64_out_index(x)] = v[x];
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = row[t + 65 * r];
      {
        // v[8a + b] *= W_N^(n1 t) * W_4096^(t (8a + b)); the thread constant rides on the w^(8a) factors
        const float2* tw = twA + t;
        const float2 cst = __ldg(twC + n1 * 64 + t);
        float2 wb[8];
#pragma unroll
        for (int b = 1; b < 8; b++) wb[b] = __ldg(tw + (b - 1) * 64);
        v[0] = cmul(v[0], cst);
#pragma unroll
        for (int b = 1; b < 8; b++) v[b] = cmul(v[b], cmul(cst, wb[b]));
#pragma unroll
        for (int a = 1; a < 8; a++) {
          const float2 wa = cmul(cst, __ldg(tw + (8 * a - 1) * 64));
          v[8 * a] = cmul(v[8 * a], wa);
#pragma unroll
          for (int b = 1; b < 8; b++) v[8 * a + b] = cmul(v[8 * a + b], cmul(wa, wb[b]));
        }
      }
      if constexpr (C > 1) {                             // this warp is done reading its row (the values were consumed
        __syncwarp();                                    // above): tell every CTA of the cluster, this one included
        if (lane < C) mbar_arrive_peer_relaxed(map_shared(&sm.bar[kBarReady], uint32_t(lane)));
      }
      dft64_inplace(v);                                  // slot x: A[n1][k2 = t + 64 q] * W_N^(n1 t), q = dft64_out_index(x)
      // ---- exchange: slot (t, q) -> S[n1][t + 64 (q mod 64/C)] in the CTA that owns k2 = t + 64 q ---------------------
      if constexpr (C > 1) {                             // every warp of every CTA of the cluster has finished reading
        mbar_wait(&sm.bar[kBarReady], ph_ready); ph_ready ^= 1u;
        if (tid == 0) mbar_expect_tx(&sm.bar[kBarRows], uint32_t(C - 1) * kClThreads * (64 / C) * uint32_t(sizeof(float2)));
      } else {
        __syncthreads();
      }
      {
        const float2* tb = twB + n1 * 64;
        float2* own = rows + size_t(n1) * KL + t;        // S[n1][t + 64 (q mod QPC)] of whichever CTA owns q
        constexpr int QPC = 64 / C;                      // q values per owner
#pragma unroll
        for (int c = 0; c < C; c++) {
          if (uint32_t(c) == crank) {                    // this CTA's own share: plain shared-memory stores
#pragma unroll
            for (int x = 0; x < 64; x++)
              if (dft64_out_index(x) / QPC == c)
                own[64 * (dft64_out_index(x) % QPC)] = cmul(v[x], __ldg(tb + dft64_out_index(x)));   // times W_N^(64 n1 q)
          } else {
            const uint32_t dst = map_shared(own, uint32_t(c)), rbar = map_shared(&sm.bar[kBarRows], uint32_t(c));
#pragma unroll
            for (int x = 0; x < 64; x++)
              if (dft64_out_index(x) / QPC == c)
                st_async(dst + uint32_t(sizeof(float2)) * 64u * uint32_t(dft64_out_index(x) % QPC),
                         cmul(v[x], __ldg(tb + dft64_out_index(x))), rbar);
          }
        }
      }
      __syncthreads();                                   // this CTA's own share
      if constexpr (C > 1) { mbar_wait(&sm.bar[kBarRows], ph_rows); ph_rows ^= 1u; }   // ... and everyone else's: S is complete
      {                                                  // next buffer's loads go in flight behind phase 2 + epilogue
        uint32_t ns = s, nk = k + 1;
        if (nk == K) { nk = 0; ns = s + n_clusters; }
        if (ns < p.n_spectra) {
          const uint8_t* nb = buffer_ptr(ns, nk);
          if constexpr (kPre > 0) load_raw(nb, 0, kPre);
          if constexpr (kPre < 32) {
            // the loads that have no registers to wait in (u >= kPre, i.e. the upper (32 - kPre)/32 of the buffer,
            // contiguous): one thread asks the copy engine to pull this CTA's 1/C of them into L2 -- no LSU work
            if (tid == 0) {
              constexpr uint32_t kBegin = uint32_t(N) * kBytes / 32u * kPre, kShare = (uint32_t(N) * kBytes - kBegin) / C;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nb + kBegin + crank * kShare), "r"(kShare) : "memory");
            }
          }
        }
      }

      // ---- phase 2: radix-R over n1 for k2 = crank KL + tid + 256 u; power; K-average ----------------------------------
      float pw[64];                                      // [u][k1]
#pragma unroll
      for (int u = 0; u < U; u++) {
        float2 x[kPts];
        const float2* src = rows + tid + 256 * u;
#pragma unroll
        for (int r = 0; r < R; r++) x[r] = src[size_t(r) * KL];
        if constexpr (R == 16) dft16(x);
        else if constexpr (R == 8) dft8<1, 0>(x);
        else dft4<1, 0>(x);
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
          const float2 sq2 = __fmul2_rn(x[k1], x[k1]);   // fl(re*re), fl(im*im): no FMA contraction
          float pv = __fadd_rn(sq2.x, sq2.y);
          if constexpr (AVG) pv = acc[u * R + k1] = (k == 0) ? pv : __fadd_rn(acc[u * R + k1], pv);
          pw[u * R + k1] = pv;
        }
      }
      if (k + 1 < K) {
        __syncthreads();                                 // S has been read: the next buffer's rows may be written
        continue;
      }

      // ---- epilogue: dB, spectrum out, detection; bins k = k2 + M k1 --------------------------------------------------
      uint32_t hit_lo = 0, hit_hi = 0;                   // bit (u R + k1)
      float* out = p.spectra ? p.spectra + size_t(s) * N + crank * KL + tid : nullptr;
#pragma unroll
      for (int u = 0; u < U; u++) {
#pragma unroll
        for (int k1 = 0; k1 < R; k1++) {
          const int slot = u * R + k1;
          const float pbar = AVG ? __fmul_rn(pw[slot], p.inv_averaging) : pw[slot];
          const float db = kDbPerLog2 * __log2f(pbar);
          pw[slot] = db;
          if (out) out[256 * u + M * k1] = db;
          if (db > p.threshold) {                        // strict >, NaN never hits (process.cpp:54)
            if (slot < 32) hit_lo |= 1u << slot; else hit_hi |= 1u << (slot - 32);
          }
        }
      }
      hit_lo &= cand_lo;
      hit_hi &= cand_hi;
      // mask words of this warp: slot (u, k1) <-> word k1 WK + 8 u + wslot.  Zero them, then one ballot per slot that
      // has a hit anywhere in the warp (a couple per spectrum).
      {
        const int s0 = lane, s1 = lane + 32;
        sm.mask[(s0 % R) * WK + 8 * (s0 / R) + wslot] = 0u;
        sm.mask[(s1 % R) * WK + 8 * (s1 / R) + wslot] = 0u;
      }
      __syncwarp();
      uint32_t rem_lo = __reduce_or_sync(0xffffffffu, hit_lo), rem_hi = __reduce_or_sync(0xffffffffu, hit_hi);
      while ((rem_lo | rem_hi) != 0u) {                  // warp-uniform
        int slot;
        if (rem_lo) { slot = __ffs(rem_lo) - 1; rem_lo &= rem_lo - 1; } else { slot = 32 + __ffs(rem_hi) - 1; rem_hi &= rem_hi - 1; }
        const uint32_t mine = slot < 32 ? (hit_lo >> slot) & 1u : (hit_hi >> (slot - 32)) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) sm.mask[(slot % R) * WK + 8 * (slot / R) + wslot] = bal;
      }
      __syncthreads();
      // per k1: exclusive prefix over this CTA's WK words, the (k1, CTA) count to every CTA, mask words to HBM
      for (int k1 = warp; k1 < R; k1 += 8) {
        uint32_t run = 0;
#pragma unroll
        for (int w0 = 0; w0 < WK; w0 += 32) {
          const uint32_t mw = sm.mask[k1 * WK + w0 + lane];
          const uint32_t c = __popc(mw);
          uint32_t inc = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += a;
          }
          sm.prefix[k1 * WK + w0 + lane] = run + inc - c;
          run += __shfl_sync(0xffffffffu, inc, 31);
          if (p.masks != nullptr) {
            const uint32_t word0 = ((crank * KL + uint32_t(M) * k1) ^ half) >> 5;      // first word of this range
            p.masks[size_t(s) * (N / 32) + word0 + w0 + lane] = mw;
          }
        }
        if constexpr (C == 1) {
          if (lane == 0) sm.table[k1 * 4] = run;
        } else if (lane < C) {
          st_async(map_shared(&sm.table[k1 * 4 + crank], uint32_t(lane)), run, map_shared(&sm.bar[kBarTable], uint32_t(lane)));
        }
      }
      if constexpr (C == 1) {
        __syncthreads();
      } else {
        if (tid == 0) mbar_expect_tx(&sm.bar[kBarTable], uint32_t(R * C) * 4u);   // R counts from each CTA (this one included)
        mbar_wait(&sm.bar[kBarTable], ph_table); ph_table ^= 1u;
      }
      // ranks: shifted-bin order is (k1 ^ R/2) major, owner CTA minor
      if (tid < R) {
        const uint32_t mine = uint32_t(tid) ^ uint32_t(R / 2);
        uint32_t before = 0, total = 0;
        for (int kk = 0; kk < R; kk++)
          for (int c = 0; c < C; c++) {
            const uint32_t cnt = sm.table[kk * 4 + c];
            const uint32_t key = uint32_t(kk) ^ uint32_t(R / 2);
            total += cnt;
            if (key < mine || (C > 1 && key == mine && uint32_t(c) < crank)) before += cnt;
          }
        sm.base[tid] = before;
        if (tid == 0) {
          sm.total = total;
          if (crank == 0 && p.counts != nullptr) p.counts[s] = total;
        }
      }
      __syncthreads();
      if (p.hits != nullptr && (hit_lo | hit_hi) != 0u) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
          for (int k1 = 0; k1 < R; k1++)
            if ((u * R + k1 < 32 ? hit_lo >> (u * R + k1) : hit_hi >> (u * R + k1 - 32)) & 1u) {
              const uint32_t word = uint32_t(k1 * WK + 8 * u) + wslot;
              const uint32_t rank = sm.base[k1] + sm.prefix[word] + __popc(sm.mask[word] & ((1u << lane) - 1u));
              if (rank < p.hit_cap) {
                scn_hit h;
                h.bin = (crank * KL + uint32_t(tid) + 256u * u + uint32_t(M) * k1) ^ half;
                h.power_db = pw[u * R + k1];
                p.hits[size_t(s) * p.hit_cap + rank] = h;
              }
            }
      }
      __syncthreads();                                   // mask / prefix / base / S are free for the next spectrum
    }
  }
  // no CTA may exit while a peer can still address its shared memory
  if constexpr (C > 1) {
    cluster_arrive();
    cluster_wait();
  }
}

template <int C, int KIND>
static const void* cluster_func(bool dc, bool avg) {
  if (dc) return avg ? reinterpret_cast<const void*>(&spectrum_sense_cluster_kernel<C, KIND, true, true>)
                     : reinterpret_cast<const void*>(&spectrum_sense_cluster_kernel<C, KIND, true, false>);
  return avg ? reinterpret_cast<const void*>(&spectrum_sense_cluster_kernel<C, KIND, false, true>)
             : reinterpret_cast<const void*>(&spectrum_sense_cluster_kernel<C, KIND, false, false>);
}

template <int C>
static const void* cluster_func_kind(int kind, bool dc, bool avg) {
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: return cluster_func<C, SCN_KIND_BYTE_COMPLEX>(dc, avg);
    case SCN_KIND_SHORT: return cluster_func<C, SCN_KIND_SHORT>(dc, avg);
    case SCN_KIND_SHORT_COMPLEX: return cluster_func<C, SCN_KIND_SHORT_COMPLEX>(dc, avg);
    case SCN_KIND_FLOAT_COMPLEX: return cluster_func<C, SCN_KIND_FLOAT_COMPLEX>(false, avg);
    default: return nullptr;
  }
}

bool variant_cluster(int kind, int log2n, bool dc, bool avg, KernelVariant* out) {
  const void* f = nullptr;
  int C = 0;
  switch (log2n) {
    case 14: C = 1; f = cluster_func_kind<1>(kind, dc, avg); break;
    case 15: C = 2; f = cluster_func_kind<2>(kind, dc, avg); break;
    case 16: C = 4; f = cluster_func_kind<4>(kind, dc, avg); break;
    default: return false;
  }
  if (!f) return false;
  out->func = f;
  out->threads = kClThreads;
  out->smem_bytes = sizeof(ClusterSmem);
  out->transforms_per_cta = 1;
  out->name = C == 1 ? "spectrum_sense_cluster<1 CTA, 4 x 4096><N=2^14>"
            : C == 2 ? "spectrum_sense_cluster<2-CTA cluster, DSMEM, 8 x 4096><N=2^15>"
                     : "spectrum_sense_cluster<4-CTA cluster, DSMEM, 16 x 4096><N=2^16>";
  out->twiddle_layout = 3;
  out->cluster = C;
  return true;
}

}  // namespace scn
