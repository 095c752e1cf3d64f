// radix_sort.cu — LSD one-sweep radix sort for sm_100a (no thrust / cub).
//
// Replaces the cub::DeviceRadixSort / DeviceMergeSort call sites of the reference:
//   cpp/src/sort/sorted_order_radix.cu:57-180  (SortPairs over (key, row index))
//   cpp/src/sort/sort_radix.cu:59-161          (SortKeys fast path of cudf::sort)
//   cpp/src/sort/sort_column_impl.cuh:35-97    (nullable single column; comparator semantics)
//   cpp/src/sort/sort_impl.cuh:31-96           (table dispatch, defaults, multi-column lexicographic)
//
// Design (one GPU, N rows, key type of W bytes => W passes of 8 bits):
//   1. histogram kernel: one streaming read of the keys, W x 256 digit counts (smem atomics with a
//      warp match_all shortcut for constant digits).
//   2. plan kernel (1 CTA): exclusive scan of each histogram -> global digit bases; passes whose
//      digit is constant over all keys are marked trivial and skipped on the device (no host sync);
//      ping-pong buffer roles are chosen so that the last executed pass writes the output buffer.
//   3. one-sweep pass kernel per digit: dynamic tile ids, warp-level MATCH.ANY ranking (stable),
//      per-digit decoupled look-back over 32-bit {flag,count} words, keys and row indices staged
//      through shared memory in tile-sorted order and written as coalesced per-digit runs.
//      Pass 1 generates row indices on the fly; the last pass of sorted_order writes indices only.
//   4. nullable column: warp-ballot/popc compaction splits valid rows (twiddled key, row index) from
//      null rows (row index in input order), then 3. runs on the valid part only.
// Keys are twiddled to unsigned order-preserving bits on first load (device_utils.cuh) and inverted
// for descending order, which keeps the sort stable like cub's Descending variants.
#include "common.cuh"
#include "device_utils.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b2 {
namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX      = 256;
constexpr uint32_t FLAG_AGG  = 1u << 30;
constexpr uint32_t FLAG_INCL = 2u << 30;
constexpr uint32_t VAL_MASK  = (1u << 30) - 1;

struct pass_plan {
  int32_t trivial;  // 1: every key has the same digit -> pass skipped
  int32_t key_src;  // 0 raw input, 1 buffer A, 2 buffer B
  int32_t key_dst;  // 1 / 2 (unused when last && pairs)
  int32_t idx_src;  // -1 implicit iota, 0 buffer X (output), 1 buffer Y (temp)
  int32_t idx_dst;  // 0 / 1
  int32_t last;     // 1: last executed pass
  int32_t hybrid;   // 1: partial LSD (top digits only) + segment fix-up follows; this pass keeps its keys and does not untwiddle
  int32_t pad;
};

struct sort_ctl {
  pass_plan plan[8];
  uint32_t base[2][8][RADIX];  // [portion parity][pass][digit] global start offset of digit
  int32_t any_pass;            // number of executed passes
  uint32_t nan_count;          // FLOAT keys only
  // hybrid sort (see segment_fix_kernel): the executed passes cover only the digits >= fix_shift / 8; rows whose keys agree
  // on those bits form short segments that the fix-up orders by the remaining low bits
  int32_t hybrid;              // decided by the plan kernel
  int32_t fix_shift;           // segment prefix = key >> fix_shift
  int32_t fix_key_buf;         // key buffer (1 / 2) written by the last executed pass
  int32_t fix_idx_buf;         // payload / row-id buffer (0 / 1 / 2) written by the last executed pass
  uint32_t overflow;           // set by the fix-up when a segment is too long for it: the host reruns the full LSD sort
  // two-phase histogram of the hybrid plan: phase 1 counts the top four digits only; the low digits are counted (phase 2) only
  // when the plan cannot stop above them
  int32_t need_low;            // 1: phase 1 could not decide, the low digits' histograms are required
  int32_t fix_fast;            // hybrid plan: which segment_fix_kernel instantiation runs (0: plain walks, 1: branch-free first neighbours)
  unsigned long long vary;     // OR of (key ^ first key) over the input
};

template <typename UK>
__device__ __forceinline__ UK twiddle_rt(UK bits, int kind, UK desc_mask)
{
  UK k;
  if (kind == (int)key_kind::SIGNED) k = twiddle_in<UK, key_kind::SIGNED>(bits);
  else if (kind == (int)key_kind::FLOAT) {
    if constexpr (sizeof(UK) >= 4) k = twiddle_in<UK, key_kind::FLOAT>(bits);
    else k = bits;
  } else k = bits;
  return k ^ desc_mask;
}
// inverse for integer kinds (float keys never take the keys-only path)
template <typename UK>
__device__ __forceinline__ UK untwiddle_rt(UK k, int kind, UK desc_mask)
{
  k ^= desc_mask;
  if (kind == (int)key_kind::SIGNED) k ^= (UK(1) << (sizeof(UK) * 8 - 1));
  return k;
}

// ------------------------------------------------------------------------------------------------
// 1. histogram
// ------------------------------------------------------------------------------------------------
// MIX (64-bit keys, join partitioning): keys are mix64(raw) and only the digits of passes 6 and 7 are counted.
// pass_mask: bit p set = count digit p (the hybrid plan first looks at the top digits only; a partition pass needs one).
// gate: when not null the kernel returns at once unless *gate != 0 (second-phase histogram of the low digits).
// vary_out: when not null receives the OR over all keys of (key ^ first key): a digit is constant over the input iff its
// byte of that word is zero (lets the plan know which uncounted digits are trivial).
template <typename UK, bool MIX = false>
__global__ void __launch_bounds__(512) histogram_kernel(const UK* __restrict__ keys, int64_t n, int raw, int kind,
                                                       UK desc_mask, uint32_t* __restrict__ ghist,
                                                       uint32_t* __restrict__ nan_count, uint32_t pass_mask = 0xffu,
                                                       const int32_t* __restrict__ gate = nullptr,
                                                       unsigned long long* __restrict__ vary_out = nullptr)
{
  constexpr int NP = sizeof(UK);
  if (gate != nullptr && *gate == 0) return;
  __shared__ uint32_t sh[NP][RADIX];
  for (int i = threadIdx.x; i < NP * RADIX; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  uint32_t nans = 0;
  constexpr int VEC = 16 / sizeof(UK);
  // head (unaligned prefix), vector body, tail
  const uintptr_t addr = reinterpret_cast<uintptr_t>(keys);
  int64_t head = ((16 - (addr & 15)) & 15) / sizeof(UK);
  if (head > n) head = n;
  const int64_t nvec = (n - head) / VEC;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;

  // reference key for the vary word: the (transformed) first key of the column
  UK k_first;
  {
    const UK r0 = keys[0];
    if constexpr (MIX) k_first = (UK)mix64((uint64_t)r0);
    else k_first = raw ? twiddle_rt<UK>(r0, kind, desc_mask) : r0;
  }
  uint64_t vary_acc = 0;
  auto account = [&](UK rawbits, bool active) {
    UK k;
    if constexpr (MIX) k = (UK)mix64((uint64_t)rawbits);
    else k = raw ? twiddle_rt<UK>(rawbits, kind, desc_mask) : rawbits;
    if (active && raw && kind == (int)key_kind::FLOAT && (UK)(k ^ desc_mask) == (UK)~UK(0)) nans++;
    if (active) vary_acc |= (uint64_t)(k ^ k_first);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!((pass_mask >> p) & 1u)) continue;
      unsigned d = (unsigned)(k >> (p * 8)) & 255u;
      unsigned amask = __ballot_sync(0xffffffffu, active);
      if (amask == 0) continue;
      int pred = 0;
      // constant-digit shortcut: one add per warp instead of 32 same-address atomics
      if (amask == 0xffffffffu) __match_all_sync(0xffffffffu, d, &pred);
      if (pred) {
        if (lane_id() == 0) atomicAdd(&sh[p][d], 32u);
      } else if (active) {
        atomicAdd(&sh[p][d], 1u);
      }
    }
  };

  // fast path: whole warps of full 16-byte vectors. Digits that are constant across the warp (the
  // common case for the high bytes of real data) are detected with two REDUX ops on key ^ key(lane 0)
  // and counted by one lane, so that skewed inputs do not serialise on same-address atomics.
  auto account_fast = [&](UK rawbits) {
    UK k;
    if constexpr (MIX) k = (UK)mix64((uint64_t)rawbits);
    else k = raw ? twiddle_rt<UK>(rawbits, kind, desc_mask) : rawbits;
    if (raw && kind == (int)key_kind::FLOAT && (UK)(k ^ desc_mask) == (UK)~UK(0)) nans++;
    // bits in which some lane differs from the column's first key: a digit whose byte is zero here is the same in all 32
    // lanes (counted once per warp), and the OR over all warps tells which digits are constant over the whole input
    const uint64_t d64 = (uint64_t)(k ^ k_first);
    uint32_t vary_lo = __reduce_or_sync(0xffffffffu, (uint32_t)d64);
    uint32_t vary_hi = sizeof(UK) > 4 ? __reduce_or_sync(0xffffffffu, (uint32_t)(d64 >> 32)) : 0u;
    const uint64_t vary = ((uint64_t)vary_hi << 32) | vary_lo;
    vary_acc |= vary;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!((pass_mask >> p) & 1u)) continue;
      const unsigned d = (unsigned)(k >> (p * 8)) & 255u;
      if (((vary >> (p * 8)) & 255u) == 0) {
        if (lane_id() == 0) atomicAdd(&sh[p][d], 32u);
      } else {
        atomicAdd(&sh[p][d], 1u);
      }
    }
  };
  const int4* vkeys = reinterpret_cast<const int4*>(keys + head);
  int64_t v = tid;
  for (; (v | 31) < nvec; v += nthreads) {
    int4 q = ld_nc_v4(vkeys + v);
    UK tmp[VEC];
    memcpy(tmp, &q, 16);
#pragma unroll
    for (int j = 0; j < VEC; ++j) account_fast(tmp[j]);
  }
  // remaining (< 32) vectors of the last partial warp-row, plus head / tail scalars: block 0, warp 0
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int64_t vrem0 = nvec / 32 * 32;
    {
      const int64_t vv = vrem0 + threadIdx.x;
      const bool act = vv < nvec;
      int4 q = act ? ld_nc_v4(vkeys + vv) : make_int4(0, 0, 0, 0);
      UK tmp[VEC];
      memcpy(tmp, &q, 16);
#pragma unroll
      for (int j = 0; j < VEC; ++j) account(tmp[j], act);
    }
    int64_t tail_start = head + nvec * VEC;
    int64_t nscalar = head + (n - tail_start);
    for (int64_t b = 0; b < nscalar; b += 32) {
      int64_t j = b + threadIdx.x;
      bool act = j < nscalar;
      int64_t e = j < head ? j : tail_start + (j - head);
      UK rb = act ? keys[e] : UK(0);
      account(rb, act);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NP * RADIX; i += blockDim.x) {
    uint32_t c = (&sh[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
  if (nan_count) {
    nans = warp_sum(nans);
    if (lane_id() == 0 && nans) atomicAdd(nan_count, nans);
  }
  if (vary_out) {
    const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)vary_acc);
    const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(vary_acc >> 32));
    if (lane_id() == 0 && (lo | hi)) atomicOr(vary_out, ((unsigned long long)hi << 32) | lo);
  }
}

// ------------------------------------------------------------------------------------------------
// 2. plan
// ------------------------------------------------------------------------------------------------
// mode_pairs: 1 sorted_order (indices), 0 keys-only.  raw: 1 keys come from the user's column
// (implicit indices), 0 keys already twiddled in buffer A with explicit indices in idx buffer
// `pre_idx_buf`.
// hyb_allowed: the caller can run the segment fix-up (and rerun without it on overflow), so the plan may stop the LSD
// passes early. The number of top digits that has to be sorted is chosen from the digit histograms: if the digits
// were independent, rows would share a given combination of the chosen digits with probability prod_p sum_d (c_pd / n)^2,
// i.e. a segment would hold about n * prod rows. The estimate only selects the plan; the fix-up verifies it.
constexpr double HYB_MAX_EXPECTED_SEGMENT = 4.0;
constexpr int HYB_MIN_SAVED_PASSES = 2;

// phase 0: every digit was counted. phase 1 (64-bit keys, hybrid allowed): only digits 4..7 were counted; the low digits are
// known to be constant or not from ctl->vary; if the hybrid plan can stop within the top digits it is final, otherwise
// ctl->need_low is raised, the gated second histogram counts digits 0..3 and phase 2 plans with everything (phase 2 returns
// at once when phase 1 was final).
__global__ void plan_kernel(const uint32_t* __restrict__ ghist, int npass, uint32_t n, int raw, int pre_idx_buf,
                            sort_ctl* ctl, int first_pass, int last_pass, int hyb_allowed, int phase = 0)
{
  __shared__ uint32_t warp_tot[8];
  __shared__ int triv[8];
  __shared__ double sq[RADIX];
  __shared__ double coll[8];  // sum_d (c_d / n)^2 of each pass
  const int d = threadIdx.x;  // 256 threads
  if (phase == 2 && ctl->need_low == 0) return;
  const int counted_from = phase == 1 ? 4 : 0;  // digits below were not counted
  const unsigned long long vary = ctl->vary;
  for (int p = 0; p < npass; ++p) {
    uint32_t c = ghist[p * RADIX + d];
    if (p < counted_from) c = ((vary >> (8 * p)) & 0xffull) == 0 ? (d == 0 ? n : 0u) : 0u;  // all-or-nothing stand-in: constant digit or unknown
    if (d == 0) triv[p] = 0;
    {
      const double f = (double)c / (double)n;
      sq[d] = f * f;
    }
    __syncthreads();
    if (d == 0) {
      double t = 0;
      for (int j = 0; j < RADIX; ++j) t += sq[j];
      coll[p] = t;
    }
    if (c == n || p < first_pass || p > last_pass) triv[p] = 1;  // passes outside [first,last] are skipped
    uint32_t inc = warp_inclusive_sum(c);
    if ((d & 31) == 31) warp_tot[d >> 5] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (d >> 5); ++w) woff += warp_tot[w];
    ctl->base[0][p][d] = woff + inc - c;
    __syncthreads();
  }
  if (d == 0) {
    int nexec = 0;
    for (int p = 0; p < npass; ++p) nexec += triv[p] ? 0 : 1;
    int hybrid = 0, low = 0;
    double seg_est = 0.0;    // hybrid plan: expected rows per segment
    bool undecided = false;  // phase 1: the decision would need an uncounted digit
    if (hyb_allowed && nexec > HYB_MIN_SAVED_PASSES) {
      double e = (double)n;
      int k = 0;
      low = npass;
      for (int p = npass - 1; p >= 0 && (k == 0 || e > HYB_MAX_EXPECTED_SEGMENT); --p) {  // at least one pass
        if (triv[p]) continue;
        if (p < counted_from) { undecided = true; break; }
        e *= coll[p];
        ++k;
        low = p;
      }
      if (!undecided && e <= HYB_MAX_EXPECTED_SEGMENT && nexec - k >= HYB_MIN_SAVED_PASSES) {
        hybrid = 1;
        seg_est = e;
        for (int p = 0; p < low; ++p) triv[p] = 1;
        nexec = k;
      }
    }
    if (phase == 1) {
      // final only when the hybrid plan stops within the counted digits, or when no uncounted digit has to be sorted at all
      bool low_needed = false;
      for (int p = 0; p < counted_from; ++p) low_needed = low_needed || !triv[p];
      if (!hybrid && low_needed) {
        ctl->need_low = 1;
        return;
      }
      ctl->need_low = 0;
    }
    ctl->hybrid    = hybrid;
    // expected rows per segment decides the fix-up flavour: mostly single-row segments (1e9 uniform keys: 0.23) are fastest with the
    // plain walks (5.6 vs 6.9 ms), ~2-row segments (a rank's shard of the sharded sort: 1.9) with the branch-free form (8.8 vs 9.5 ms)
    ctl->fix_fast  = (hybrid && seg_est > 0.75) ? 1 : 0;
    ctl->fix_shift = low * RADIX_BITS;
    ctl->overflow  = 0;
    // idx buffers: 0 = output, 1 = temp. the last executed pass must write 0.
    // key buffers: 1 = A, 2 = B (keys-only: 1 = output, 2 = temp; last executed pass must write 1)
    int k = 0;
    int key_cur = raw ? 0 : 1;
    int idx_cur = raw ? -1 : pre_idx_buf;
    for (int p = 0; p < npass; ++p) {
      pass_plan pl{};
      pl.trivial = triv[p];
      if (!triv[p]) {
        int remaining_after = nexec - 1 - k;  // passes after this one
        pl.key_src = key_cur;
        pl.idx_src = idx_cur;
        // A pass never writes the buffer it reads. Raw input: ping-pong so that the final pass lands in
        // key buffer 1 (the keys-only output) / idx buffer 0 (the output column). Pre-compacted input
        // (keys in A, row ids in idx buffer 1): keys alternate A/B (their final home is irrelevant,
        // pairs mode never returns keys); row ids go to 0 whenever an even number of passes remains,
        // else to a non-zero buffer other than the source (third buffer only for the first pass).
        // Hybrid (raw input only): the last pass must land in the TEMP buffers (key 2 / idx 1), because the fix-up writes
        // the output buffers (key 1 / idx 0) from them.
        const int par = (remaining_after + hybrid) % 2;
        pl.key_dst = raw ? (par == 0 ? 1 : 2) : (key_cur == 1 ? 2 : 1);
        pl.idx_dst = par == 0 ? 0 : (idx_cur == 1 ? 2 : 1);
        pl.last    = remaining_after == 0;
        pl.hybrid  = hybrid;
        if (pl.last) {
          ctl->fix_key_buf = pl.key_dst;
          ctl->fix_idx_buf = pl.idx_dst;
        }
        key_cur = pl.key_dst;
        idx_cur = pl.idx_dst;
        ++k;
      }
      ctl->plan[p] = pl;
    }
    ctl->any_pass = nexec;
  }
}

__global__ void set_fix_fast_kernel(sort_ctl* ctl, int v) { ctl->fix_fast = v; }

// ------------------------------------------------------------------------------------------------
// 3. one-sweep pass
// ------------------------------------------------------------------------------------------------
struct pass_args {
  const void* key_bufs[3];  // [0] raw input column data (already offset), [1], [2]
  int32_t* idx_bufs[3];
  sort_ctl* ctl;
  uint32_t* status;         // [tiles][256] flagged count / prefix rows for this (pass, portion)
  uint32_t* tile_counter;   // for this (pass, portion)
  int64_t portion_start;    // element offset of this portion
  uint32_t portion_n;       // elements in this portion
  int32_t pass;
  int32_t portion_parity;   // which base[] copy to read; the other is written for the next portion
  int32_t has_next_portion;
  int32_t kind;             // key_kind for raw loads / keys-only untwiddle
  int32_t pairs;            // 1: (key, idx) ; 0: keys only
  uint64_t desc_mask;
  int32_t keep_keys;        // pairs mode: also write the keys in the last executed pass (partial sorts)
  const void* val_in;       // CARRY kernels: the caller's payload column (first pass source); last member on purpose
  // RANGE kernels (sharded sort: the range partition IS the exchange): digit = number of splitters <= key; the rows of digit d
  // are written to range_key_dst[d] / range_val_dst[d] (local or PEER memory) at their rank inside this GPU's digit-d run
  const void* range_splitters;       // device array of (range_parts - 1) twiddled keys, ascending
  void* const* range_key_dst;        // device array of range_parts pointers
  void* const* range_val_dst;        // same for the carried payload (CARRY) or null
  int32_t range_parts;
  int32_t range_hash;                // 1: no splitters, bucket = range_hash_bucket(key) (hash partition: the sharded join's shuffle)
  // MIX kernels, estimated bases (radix_partition_mix_carry_est): no histogram ran; digit d owns rows [d * est_cap, (d + 1) * est_cap) of the
  // output, rows beyond that range are dropped and ctl->overflow is raised; the last tile leaves every digit's end offset in
  // ctl->base[portion_parity ^ 1][pass]
  uint32_t est_cap;
};

// Hash bucket of the fused partition pass. The additive constant decorrelates it from the local radix join, which partitions
// by the top bits of mix64(key) itself: rows that share an exchange bucket still spread over all of the join's partitions.
__host__ __device__ __forceinline__ unsigned range_hash_bucket(uint64_t twiddled_key, int parts)
{
  const uint64_t h = mix64(twiddled_key + 0x9E3779B97F4A7C15ull);
  return (unsigned)(((h >> 32) * (uint64_t)parts) >> 32);
}

// Block-wide barrier of a warp-specialised kernel: the ranking warps and the look-back warps reach it from different branches, which
// `__syncthreads()` only tolerates in practice; a named barrier with the explicit thread count (`bar.sync 2, n`) is the
// PTX-conformant form for sm_70+ (barrier 0 is what `__syncthreads()` lowers to, and compute-sanitizer synccheck holds it to the
// C++ rule that every thread reaches the same call site).
__device__ __forceinline__ void cta_barrier(int nthreads)
{
#ifdef B2_EMU
  (void)nthreads;
  __syncthreads();
#else
  asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ void ranker_barrier(int nthreads)
{
#ifdef B2_EMU
  ::emu::named_barrier(1, nthreads);
#else
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint32_t* p)
{
#ifdef B2_EMU
  uint4 v;
  memcpy(&v, p, sizeof(v));
  return v;
#else
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
#endif
}

// One CTA = THREADS ranking threads (NWARPS warps holding IPT keys per thread) + LBW look-back warps.
// Decoupled look-back without fences: a tile's digit counts (later: inclusive prefixes) are published
// as rows of 256 words written in 16-byte groups, ONE lane per group, with a 2-bit state in the top
// bits of the group's first word (0 nothing, 1 aggregate, 2 inclusive; the other three words carry
// plain 32-bit values).  A 16-byte aligned store is observed all-or-nothing by a 16-byte load, so a
// reader lane checks one flag per group and needs no memory fence (measured: the release/acquire
// variant with a separate state word was 1.6x slower; polling 256 individually flagged words spent
// 35 % of all issued instructions in the polling loops).  Each look-back lane owns one group
// (DPL = 4 digits), fetches LBT predecessor rows per round with independent 128-bit loads and folds
// them with warp-uniform decisions (REDUX and/or over the ready / inclusive bits).
constexpr int LBW = 2;   // look-back warps per CTA
constexpr int DPL = 4;   // digits per look-back lane (one 128-bit load per predecessor tile)
constexpr int LBT = 8;   // predecessor tiles fetched per round

// VT = uint32_t: payload = 32-bit row ids (generated in the first pass) — the shipped path.
// VT = uint64_t / CARRY: payload = the caller's 8-byte (or 4-byte with VT = uint32_t) values column, loaded coalesced in
// the first pass and carried through every pass, so that sort_by_key needs neither row ids nor a gather
// (EXPERIMENTAL in round 1: opt-in with B2_SORT_CARRY=1, not yet run on hardware).
// MIX: raw 64-bit keys are replaced by mix64(key) on load (hash-join partitioning: the first executed pass reads the
// packed key column itself, so the mixed keys are never materialised unsorted).
// SAFE (default): a __syncwarp between the followers' read of the peer bitmap and the leader's clear — race-free under
// independent thread scheduling (compute-sanitizer racecheck flags the form without it); measured cost: none (7.86 vs 7.85 ms).
#ifdef B2_EMU
constexpr bool EMU_BUILD = true;
#else
constexpr bool EMU_BUILD = false;
#endif
// RMW (default): the leader advances the warp's running digit offset with one ATOMS.ADD (returning the old value) instead of
// LDS + STS — one shared-memory operation less per key in a kernel bound by shared-memory wavefronts (7.49 vs 7.85 ms per pass).
// BULK: full, 16-byte aligned key tiles arrive in shared memory through ONE bulk async copy (cp.async.bulk, the 1-D TMA
// path: UBLKCP in SASS) signalled by an mbarrier, and the ranking warps pick their keys up from there instead of issuing
// IPT global loads each (B2_SORT_CFG=12; other tiles take the ordinary loads).
template <typename UK, int THREADS, int IPT, int MINB, typename VT = uint32_t, bool CARRY = false, bool MIX = false, bool SAFE = true,
          bool RMW = true, bool BULK = false, bool RANGE = false>
__global__ void __launch_bounds__(THREADS + 32 * LBW, MINB) onesweep_kernel(pass_args a)
{
  constexpr int TILE   = THREADS * IPT;
  constexpr int NWARPS = THREADS / 32;
  static_assert(THREADS >= RADIX, "need one ranking thread per digit");
  static_assert(!RANGE || (!MIX && !BULK), "the range-partition pass uses the plain key load");

  const pass_plan pl = a.ctl->plan[a.pass];
  if (pl.trivial) return;

  B2_DYNAMIC_SMEM(smem_raw);
  constexpr int STAGE_W = sizeof(UK) > sizeof(VT) ? sizeof(UK) : sizeof(VT);  // staging holds keys, then the payload
  UK* s_keys          = reinterpret_cast<UK*>(smem_raw);
  VT* s_vals          = reinterpret_cast<VT*>(smem_raw);
  uint32_t* s_whist   = reinterpret_cast<uint32_t*>(smem_raw + (size_t)STAGE_W * TILE);  // [NWARPS][256]
  uint32_t* s_bm      = s_whist + NWARPS * RADIX;  // [NWARPS][256] per-warp digit -> lane bitmaps (ranking)
  uint32_t* s_off     = s_bm + NWARPS * RADIX;     // [256] global offset of digit - tile-local start
  uint32_t* s_cnt     = s_off + RADIX;             // [256] {tile count of digit, tile-local start} pairs
  uint32_t* s_misc    = s_cnt + 2 * RADIX;         // [16]
  // RANGE only: splitters, per-digit destination pointers, digit of every staged item
  UK* s_split         = reinterpret_cast<UK*>(s_misc + 16);                  // [256]
  void** s_kdst       = reinterpret_cast<void**>(s_split + RADIX);           // [256]
  void** s_vdst       = s_kdst + RADIX;                                      // [256]
  uint8_t* s_dig      = reinterpret_cast<uint8_t*>(s_vdst + RADIX);          // [TILE]

  const int tid  = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const bool ranker = warp < NWARPS;

  if (tid == 0) s_misc[0] = atomicAdd(a.tile_counter, 1u);
  if constexpr (BULK && !EMU_BUILD) {
    if (tid == 0) {
      const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_misc + 12);  // 8-byte aligned slot behind the scan scratch (s_misc[1..8])
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if (ranker) {
#pragma unroll
    for (int j = 0; j < RADIX / 32; ++j) {
      s_whist[warp * RADIX + j * 32 + lane] = 0;
      s_bm[warp * RADIX + j * 32 + lane]    = 0;
    }
  }
  if constexpr (RANGE) {
    for (int i = tid; i < RADIX; i += THREADS + 32 * LBW) {
      s_split[i] = (!a.range_hash && i < a.range_parts - 1) ? static_cast<const UK*>(a.range_splitters)[i] : ~UK(0);
      s_kdst[i]  = i < a.range_parts ? a.range_key_dst[i] : nullptr;
      s_vdst[i]  = (CARRY && i < a.range_parts) ? a.range_val_dst[i] : nullptr;
    }
  }
  __syncthreads();
  const uint32_t tile = s_misc[0];
  const uint32_t tile_base = tile * (uint32_t)TILE;  // within portion
  const uint32_t tile_n = min((uint32_t)TILE, a.portion_n - tile_base);
  const bool full = tile_n == (uint32_t)TILE;
  const uint32_t pad = (uint32_t)TILE - tile_n;      // padding items sit at the end of digit 255
  const int shift = a.pass * RADIX_BITS;
  const UK desc = (UK)a.desc_mask;

  if (!ranker) {
    // ================================ look-back warp ==============================================
    cta_barrier(THREADS + 32 * LBW);  // (S2) agg[tile][*] written by the rankers; s_cnt = {count, tile-local start}
    const int d0 = ((warp - NWARPS) * 32 + lane) * DPL;
    uint32_t cnt[DPL], loc[DPL], excl[DPL];
#pragma unroll
    for (int j = 0; j < DPL; j += 2) {
      const uint4 cs = *reinterpret_cast<const uint4*>(s_cnt + 2 * (d0 + j));  // {cnt, start, cnt, start}
      cnt[j] = cs.x; loc[j] = cs.y; cnt[j + 1] = cs.z; loc[j + 1] = cs.w;
      excl[j] = 0; excl[j + 1] = 0;
    }
    uint32_t* const my_row = a.status + (size_t)tile * RADIX + d0;
    auto publish = [&](uint32_t flag, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
#ifdef B2_EMU
      my_row[0] = flag | w0; my_row[1] = w1; my_row[2] = w2; my_row[3] = w3;
#else
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(my_row), "r"(flag | w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
#endif
    };
    if (tile == 0) {
      publish(FLAG_INCL, cnt[0], cnt[1], cnt[2], cnt[3]);  // first tile of the portion: counts are inclusive
    } else {
      publish(FLAG_AGG, cnt[0], cnt[1], cnt[2], cnt[3]);
      int64_t t = (int64_t)tile - 1;  // nearest predecessor not folded yet
      bool done = false;
      while (!done) {
        uint4 v[LBT];
#pragma unroll
        for (int r = 0; r < LBT; ++r) {
          const int64_t tt = t - r;
          v[r] = make_uint4(FLAG_INCL, 0, 0, 0);  // tiles before the first one: inclusive zero
          if (tt >= 0) v[r] = ld_volatile_v4(a.status + (size_t)tt * RADIX + d0);
        }
        uint32_t rdy = 0, inc = 0;
#pragma unroll
        for (int r = 0; r < LBT; ++r) {
          rdy |= ((v[r].x >> 30) != 0u ? 1u : 0u) << r;
          inc |= (v[r].x >> 31) << r;
        }
        const uint32_t rdy_all = __reduce_and_sync(0xffffffffu, rdy);
        const uint32_t inc_all = __reduce_and_sync(0xffffffffu, inc);
        const uint32_t inc_any = __reduce_or_sync(0xffffffffu, inc);
        const uint32_t usable  = rdy_all & ~(inc_any & ~inc_all);      // rows every lane sees in the same state
        const int n_ready   = __ffs(~usable) - 1;                       // leading usable rows (LBT if all)
        const int first_inc = (inc_all & usable) ? (__ffs(inc_all & usable) - 1) : LBT;
        const int take = min(n_ready, first_inc + 1);
        if (take == 0) { __nanosleep(40); continue; }
#pragma unroll
        for (int r = 0; r < LBT; ++r) {
          if (r < take) { excl[0] += v[r].x & VAL_MASK; excl[1] += v[r].y; excl[2] += v[r].z; excl[3] += v[r].w; }
        }
        done = first_inc < n_ready;
        t -= take;
      }
      publish(FLAG_INCL, excl[0] + cnt[0], excl[1] + cnt[1], excl[2] + cnt[2], excl[3] + cnt[3]);
    }
    const uint32_t* gb = &a.ctl->base[a.portion_parity][a.pass][d0];
    bool wants_end = a.has_next_portion != 0;
    if constexpr (MIX) wants_end = wants_end || a.est_cap != 0u;
    const bool last_of_portion = wants_end && tile_base + tile_n == a.portion_n;
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
      const uint32_t g = gb[j];
      s_off[d0 + j] = g + excl[j] - loc[j];
      if (last_of_portion) a.ctl->base[a.portion_parity ^ 1][a.pass][d0 + j] = g + excl[j] + cnt[j];
    }
    cta_barrier(THREADS + 32 * LBW);  // (S4) offsets ready
    return;
  }

  // ================================== ranking warps ================================================
  // ---- load (warp-striped: item i of lane l in warp w = w*32*IPT + i*32 + l) -------------------
  UK key[IPT];
  const uint32_t wbase = tile_base + warp * (32 * IPT) + lane;
  {
    const UK* src = static_cast<const UK*>(pl.key_src == 0 ? a.key_bufs[0] : (pl.key_src == 1 ? a.key_bufs[1] : a.key_bufs[2])) + a.portion_start;
    bool via_smem = false;
    if constexpr (BULK) via_smem = full && (reinterpret_cast<uintptr_t>(src + tile_base) & 15) == 0;  // uniform over the CTA
    if (via_smem) {
      if constexpr (BULK) {
        constexpr uint32_t BYTES = (uint32_t)(sizeof(UK) * TILE);
        static_assert(!BULK || BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
        if constexpr (EMU_BUILD) {  // emulator: the same data movement with ordinary loads
          for (int q = tid; q < TILE; q += THREADS) s_keys[q] = src[tile_base + q];
          ranker_barrier(THREADS);
        } else {
          const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(s_misc + 12);
          if (tid == 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_keys);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src + tile_base), "r"(BYTES), "r"(mbar)
                         : "memory");
          }
          asm volatile(
            "{\n .reg .pred p;\n BULK_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra BULK_DONE;\n bra BULK_WAIT;\n BULK_DONE:\n}" ::"r"(mbar)
            : "memory");
        }
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = s_keys[warp * (32 * IPT) + i * 32 + lane];
      }
    } else if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) key[i] = ld_stream(src + wbase + i * 32);
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        uint32_t e = wbase + i * 32;
        key[i] = e < a.portion_n ? ld_stream(src + e) : UK(0);
      }
    }
    if (pl.key_src == 0) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        if constexpr (MIX) key[i] = (UK)mix64((uint64_t)key[i]);
        else key[i] = twiddle_rt<UK>(key[i], a.kind, desc);
      }
    }
    if (!full) {
      // padding items take the maximum key so that they rank last in the tile
#pragma unroll
      for (int i = 0; i < IPT; ++i)
        if (wbase + i * 32 >= a.portion_n) key[i] = ~UK(0);
    }
  }

  // digit of a key: a byte of it, or (RANGE) the number of splitters <= key — padding items (all-ones key) land in the last bucket
  auto digit_of = [&](UK k) -> unsigned {
    if constexpr (RANGE) {
      if (a.range_hash) return range_hash_bucket((uint64_t)k, a.range_parts);
      int lo = 0, hi = a.range_parts - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_split[mid] <= k) lo = mid + 1;
        else hi = mid;
      }
      return (unsigned)lo;
    } else {
      return (unsigned)(k >> shift) & 255u;
    }
  };
  unsigned dg[RANGE ? IPT : 1];
  if constexpr (RANGE) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) dg[i] = (wbase + i * 32 >= a.portion_n) ? (unsigned)(RADIX - 1) : digit_of(key[i]);  // padding ranks last (digit 255)
  }
  auto digit_at = [&](int i) -> unsigned {
    if constexpr (RANGE) return dg[i];
    else return (unsigned)(key[i] >> shift) & 255u;
  };
  // ---- early counts: warp-private digit histogram by shared-memory atomics (no dependency chain) ----
  uint32_t* my_hist = s_whist + warp * RADIX;
#pragma unroll
  for (int i = 0; i < IPT; ++i) atomicAdd(&my_hist[digit_at(i)], 1u);
  __syncwarp();
  // number of distinct digits in this warp's 32*IPT keys: picks the ranking flavour below
  int distinct = 0;
#pragma unroll
  for (int j = 0; j < RADIX / 32; ++j) distinct += my_hist[j * 32 + lane] != 0u;
  distinct = __reduce_add_sync(0xffffffffu, distinct);
  ranker_barrier(THREADS);  // (S1)

  // ---- per digit: warp counts -> warp offsets; publish the aggregate ---------------------------
  uint32_t tstart = 0, count = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      uint32_t c = s_whist[w * RADIX + tid];
      s_whist[w * RADIX + tid] = count;  // exclusive offset of warp w inside digit `tid`
      count += c;
    }
    const uint32_t padded = count;
    if (tid == RADIX - 1) count -= pad;
    // exclusive scan of the padded counts over digits: 8 warps of 32 digits
    uint32_t inc = warp_inclusive_sum(padded);
    if (lane == 31) s_misc[1 + warp] = inc;
    tstart = inc - padded;
  }
  ranker_barrier(THREADS);  // (S1b)
  if (tid < RADIX) {
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w) woff += (w < warp) ? s_misc[1 + w] : 0u;
    tstart += woff;
    s_cnt[2 * tid]     = count;
    s_cnt[2 * tid + 1] = tstart;
    // fold the tile-local digit start into the warp offsets: position = s_whist[w][d] + rank in warp
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s_whist[w * RADIX + tid] += tstart;
  }
  cta_barrier(THREADS + 32 * LBW);  // (S2) releases the look-back warps; warp offsets final

  // ---- rank within warp (stable): MATCH.ANY peers + running per-warp digit offsets --------------
  // All MATCH ops are issued first (independent, pipelined); only the counter chain is serial.
  // Peer masks (lanes of the warp holding the same digit), measured cost per warp-item per SM
  // (scripts/ubench/rank_probe.cu): MATCH.ANY 60 cycles at ~30 distinct digits but 9 at <= 4;
  // 8 ballots 24; shared-memory atomicOr bitmap 17 (36 when all lanes collide).  Hence: bitmap for
  // the general case, MATCH.ANY when the warp holds only a handful of distinct digits.
  uint32_t pos[IPT];
  uint32_t* my_bm = s_bm + warp * RADIX;
  if (distinct > 4) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const unsigned d = digit_at(i);
      atomicOr(&my_bm[d], 1u << lane);
      __syncwarp();
      const unsigned peers = my_bm[d];
      const unsigned lt = __popc(peers & lanemask_lt());
      uint32_t prev = 0;
      // The emulator (tests/emu) runs the lanes of a warp one after the other between rendezvous points, so the
      // leader's clear below would be seen by the followers' read above. On the GPU the warp executes this
      // straight-line stretch converged (validated on hardware); SAFE is the formally race-free variant.
      if constexpr (SAFE || EMU_BUILD) __syncwarp();
      if (lt == 0) {
        if constexpr (RMW) {
          prev = atomicAdd(&my_hist[d], (uint32_t)__popc(peers));
        } else {
          prev = my_hist[d];
          my_hist[d] = prev + __popc(peers);
        }
        my_bm[d] = 0;
      }
      __syncwarp();
      prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
      pos[i] = prev + lt;
    }
  } else {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const unsigned d = digit_at(i);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const unsigned lt = __popc(peers & lanemask_lt());
      uint32_t prev = 0;
      if (lt == 0) {
        prev = my_hist[d];
        my_hist[d] = prev + __popc(peers);
      }
      __syncwarp();
      prev = __shfl_sync(0xffffffffu, prev, __ffs(peers) - 1);
      pos[i] = prev + lt;
    }
  }
  // tile-sorted staging of the keys
#pragma unroll
  for (int i = 0; i < IPT; ++i) s_keys[pos[i]] = key[i];
  if constexpr (RANGE) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) s_dig[pos[i]] = (uint8_t)dg[i];
  }
  // row ids are fetched only now (their registers replace the dead key registers); the loads
  // overlap with the key write-out below
  VT idx[IPT];
  if (a.pairs) {
    if (pl.idx_src < 0) {
      if constexpr (CARRY) {
        // first pass: the payload column is read at the rows' input positions (coalesced)
        const VT* vsrc = static_cast<const VT*>(a.val_in) + a.portion_start;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          const uint32_t e = wbase + i * 32;
          idx[i] = e < a.portion_n ? ld_stream(vsrc + e) : VT(0);
        }
      } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) idx[i] = (uint32_t)(a.portion_start + wbase + i * 32);
      }
    } else {
      const VT* isrc = reinterpret_cast<const VT*>(pl.idx_src == 0 ? a.idx_bufs[0] : (pl.idx_src == 1 ? a.idx_bufs[1] : a.idx_bufs[2])) + a.portion_start;
      if (full) {
#pragma unroll
        for (int i = 0; i < IPT; ++i) idx[i] = ld_stream(isrc + wbase + i * 32);
      } else {
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
          uint32_t e = wbase + i * 32;
          idx[i] = e < a.portion_n ? ld_stream(isrc + e) : VT(0);
        }
      }
    }
  }
  cta_barrier(THREADS + 32 * LBW);  // (S4) keys staged, scatter offsets ready

  const bool write_keys = !(a.pairs && pl.last) || a.keep_keys || pl.hybrid;
  UK* kdst = static_cast<UK*>(const_cast<void*>(pl.key_dst == 1 ? a.key_bufs[1] : a.key_bufs[2]));
  uint32_t dst[IPT];
  uint8_t dstd[RANGE ? IPT : 1];
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    const uint32_t q = j * THREADS + tid;
    UK k = s_keys[q];
    unsigned d;
    if constexpr (RANGE) d = s_dig[q];
    else d = (unsigned)(k >> shift) & 255u;
    dst[j] = s_off[d] + q;
    if constexpr (MIX) {
      if (a.est_cap != 0u && dst[j] >= (d + 1u) * a.est_cap) {  // digit d's reserved range is full: the caller falls back to exact bases
        if (q < tile_n) a.ctl->overflow = 1u;
        dst[j] = 0xFFFFFFFFu;
      }
    }
    if constexpr (RANGE) {
      dstd[j] = (uint8_t)d;
      if (q < tile_n) static_cast<UK*>(s_kdst[d])[dst[j]] = untwiddle_rt<UK>(k, a.kind, desc);  // the receiver sorts raw column values
    } else if (write_keys && q < tile_n && (!MIX || dst[j] != 0xFFFFFFFFu)) {
      if (!a.pairs && pl.last && !pl.hybrid) k = untwiddle_rt<UK>(k, a.kind, desc);
      kdst[dst[j]] = k;
    }
  }
  if (a.pairs) {
    ranker_barrier(THREADS);
#pragma unroll
    for (int i = 0; i < IPT; ++i) s_vals[pos[i]] = idx[i];
    ranker_barrier(THREADS);
    VT* idst = reinterpret_cast<VT*>(pl.idx_dst == 0 ? a.idx_bufs[0] : (pl.idx_dst == 1 ? a.idx_bufs[1] : a.idx_bufs[2]));
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      const uint32_t q = j * THREADS + tid;
      if constexpr (RANGE) {
        if (q < tile_n) static_cast<VT*>(s_vdst[dstd[j]])[dst[j]] = s_vals[q];
      } else {
        if (q < tile_n && (!MIX || dst[j] != 0xFFFFFFFFu)) idst[dst[j]] = s_vals[q];
      }
    }
  }
}

// all passes trivial: the sorted order is the input order
template <typename UK, bool MIX = false>
__global__ void finalize_kernel(pass_args a, int64_t n, int raw, int pre_idx_buf, int carry_bytes)
{
  if (a.ctl->any_pass != 0) return;
  if (carry_bytes) {  // carried payload and no executed pass: the output is the payload column itself
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      if (carry_bytes == 8) reinterpret_cast<uint64_t*>(a.idx_bufs[0])[i] = static_cast<const uint64_t*>(a.val_in)[i];
      else reinterpret_cast<uint32_t*>(a.idx_bufs[0])[i] = static_cast<const uint32_t*>(a.val_in)[i];
      if (a.keep_keys && raw) {  // partition passes keep the (mixed) keys next to the payload
        UK k = static_cast<const UK*>(a.key_bufs[0])[i];
        if constexpr (MIX) k = (UK)mix64((uint64_t)k);
        static_cast<UK*>(const_cast<void*>(a.key_bufs[1]))[i] = k;
      }
    }
    return;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (a.pairs) {
      if (raw) a.idx_bufs[0][i] = (int32_t)i;
      else if (pre_idx_buf != 0) a.idx_bufs[0][i] = a.idx_bufs[1][i];
      if (a.keep_keys && raw) {
        UK k = static_cast<const UK*>(a.key_bufs[0])[i];
        if constexpr (MIX) k = (UK)mix64((uint64_t)k);
        static_cast<UK*>(const_cast<void*>(a.key_bufs[1]))[i] = k;
      }
    } else {
      // keys-only (always raw): copy input to output
      static_cast<UK*>(const_cast<void*>(a.key_bufs[1]))[i] = static_cast<const UK*>(a.key_bufs[0])[i];
    }
  }
}

// Hybrid sort, second half.  The executed LSD passes ordered the rows by the key bits >= fix_shift (stable), so rows that
// agree on those bits ("segments") are contiguous and still in input order.  Each row finds its segment by walking
// outwards over the neighbouring keys in shared memory and takes its final place = segment start + number of segment
// rows that precede it in (key, input position) order.  With the plan kernel's choice of digits a segment holds a
// handful of rows (n / 2^32 = 0.23 on average for 1e9 uniform 64-bit keys after four passes), so this is one
// streaming pass: read key + payload, write payload (or the untwiddled key) a few places away.
// A walk stops after FIX_HALO rows.  A row whose walk was cut short keeps its place if every key it saw equals its own
// (a run of duplicates longer than the window: windows of neighbouring rows overlap, so if nobody objects the whole
// run is one value and already in order); otherwise it raises ctl->overflow and the host reruns the full LSD sort.
constexpr int FIX_THREADS = 256;
constexpr int FIX_IPT     = 8;
constexpr int FIX_TILE    = FIX_THREADS * FIX_IPT;
constexpr int FIX_HALO    = 64;

template <typename UK, typename VT, int FIX_FAST>
__global__ void __launch_bounds__(FIX_THREADS) segment_fix_kernel(pass_args a, int64_t n)
{
  if (!a.ctl->hybrid || a.ctl->fix_fast != (FIX_FAST ? 1 : 0)) return;
  __shared__ UK sk[FIX_TILE + 2 * FIX_HALO];
  const int shift = a.ctl->fix_shift;
  const UK* __restrict__ keys = static_cast<const UK*>(a.ctl->fix_key_buf == 1 ? a.key_bufs[1] : a.key_bufs[2]);
  const VT* __restrict__ vin  = reinterpret_cast<const VT*>(a.ctl->fix_idx_buf == 1 ? a.idx_bufs[1] : a.idx_bufs[2]);
  VT* __restrict__ vout = reinterpret_cast<VT*>(a.idx_bufs[0]);
  UK* __restrict__ kout = static_cast<UK*>(const_cast<void*>(a.key_bufs[1]));
  const UK desc = (UK)a.desc_mask;
  const int64_t ntiles = (n + FIX_TILE - 1) / FIX_TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * FIX_TILE;
    __syncthreads();  // the previous tile's walks are done
    for (int q = threadIdx.x; q < FIX_TILE + 2 * FIX_HALO; q += FIX_THREADS) {
      const int64_t g = base - FIX_HALO + q;
      sk[q] = (g >= 0 && g < n) ? ld_stream(keys + g) : UK(0);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < FIX_IPT; ++j) {
      const int i = j * FIX_THREADS + threadIdx.x;
      const int64_t gi = base + i;
      if (gi >= n) continue;
      VT v{};
      if (a.pairs) v = ld_stream(vin + gi);
      const UK k  = sk[FIX_HALO + i];
      const UK pf = k >> shift;
      // rows to the left / right that exist (array ends are segment ends)
      const int lmax = (int)(gi < FIX_HALO ? gi : (int64_t)FIX_HALO);
      const int rmax = (int)(n - 1 - gi < FIX_HALO ? n - 1 - gi : (int64_t)FIX_HALO);
      int left = 0, before = 0;
      bool all_equal = true, cut = false;
      // The first FIX_FAST neighbours on each side are examined without branches (every lane of the warp does the same
      // work: a divergent walk costs the warp its LONGEST segment, which tripled the kernel's time at ~2 rows per segment);
      // only rows whose segment reaches further continue with the loops below.
      bool in_l = true, in_r = true;  // FIX_FAST = 0: no unrolled part, both walks start at the row itself
#pragma unroll
      for (int s = 1; s <= FIX_FAST; ++s) {
        const UK ol = sk[FIX_HALO + i - s];   // inside the halo: FIX_HALO >= FIX_FAST
        const UK orr = sk[FIX_HALO + i + s];
        in_l = in_l && s <= lmax && (UK)(ol >> shift) == pf;
        in_r = in_r && s <= rmax && (UK)(orr >> shift) == pf;
        left += in_l ? 1 : 0;
        before += (in_l && ol <= k) ? 1 : 0;   // earlier rows win ties
        before += (in_r && orr < k) ? 1 : 0;
        all_equal = all_equal && (!in_l || ol == k) && (!in_r || orr == k);
      }
      int right = 0;
      if (in_l) {  // the segment extends further to the left
        for (;;) {
          if (left == lmax) { cut = lmax == FIX_HALO; break; }
          const UK o = sk[FIX_HALO + i - left - 1];
          if ((UK)(o >> shift) != pf) break;
          ++left;
          before += o <= k ? 1 : 0;
          all_equal = all_equal && o == k;
        }
      }
      if (in_r) {
        right = FIX_FAST;
        for (;;) {
          if (right == rmax) { cut = cut || rmax == FIX_HALO; break; }
          const UK o = sk[FIX_HALO + i + right + 1];
          if ((UK)(o >> shift) != pf) break;
          ++right;
          before += o < k ? 1 : 0;
          all_equal = all_equal && o == k;
        }
      }
      int64_t dst = gi - left + before;
      if (cut) {
        dst = gi;
        if (!all_equal) atomicOr(&a.ctl->overflow, 1u);
      }
      if (a.pairs) vout[dst] = v;
      else kout[dst] = untwiddle_rt<UK>(k, a.kind, desc);
    }
  }
}

// descending float keys: cub sorts the (nan_bias, value) tuple descending, i.e. NaNs come first in
// DESCENDING row order (sorted_order_radix.cu:41-50,121-131). The stable pass above leaves them
// ascending; reverse that prefix.
__global__ void reverse_nan_prefix_kernel(int32_t* idx, const uint32_t* nan_count)
{
  const uint32_t m = *nan_count;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m / 2; i += stride) {
    int32_t x = idx[i], y = idx[m - 1 - i];
    idx[i] = y;
    idx[m - 1 - i] = x;
  }
}

// ------------------------------------------------------------------------------------------------
// 4. null compaction (single nullable column)
// ------------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_ROWS    = 2048;  // rows per CTA = 64 mask words

__global__ void __launch_bounds__(CP_THREADS) valid_count_kernel(const uint32_t* __restrict__ mask, int64_t bit_offset,
                                                                  int64_t n, uint32_t* __restrict__ tile_valid)
{
  // each warp counts one tile of CP_ROWS rows
  const int64_t tile = (int64_t)blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t ntiles = (n + CP_ROWS - 1) / CP_ROWS;
  if (tile >= ntiles) return;
  const int64_t row0 = tile * CP_ROWS;
  const int64_t last_word = (bit_offset + n - 1) >> 5;
  uint32_t c = 0;
  for (int w = lane_id(); w < CP_ROWS / 32; w += 32) {
    int64_t r = row0 + (int64_t)w * 32;
    if (r < n) {
      uint32_t bits = load_mask_word_unaligned(mask, bit_offset + r, last_word);
      int64_t rem = n - r;
      if (rem < 32) bits &= (1u << rem) - 1u;
      c += __popc(bits);
    }
  }
  c = warp_sum(c);
  if (lane_id() == 0) tile_valid[tile] = c;
}

// single-CTA exclusive scan over tile counts (ntiles <= ~1M); also writes total
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t* __restrict__ tile_valid, int64_t ntiles,
                                                          uint32_t* __restrict__ total)
{
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < ntiles; b += 1024) {
    int64_t i = b + threadIdx.x;
    uint32_t v = i < ntiles ? tile_valid[i] : 0;
    uint32_t inc = warp_inclusive_sum(v);
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    uint32_t c = carry;
    if (i < ntiles) tile_valid[i] = c + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

template <typename UK>
__global__ void __launch_bounds__(CP_THREADS) compact_kernel(const UK* __restrict__ keys, const uint32_t* __restrict__ mask,
                                                             int64_t bit_offset, int64_t n, int kind, UK desc_mask,
                                                             const uint32_t* __restrict__ tile_valid_excl,
                                                             UK* __restrict__ out_keys, int32_t* __restrict__ out_valid_idx,
                                                             int32_t* __restrict__ out_null_idx)
{
  __shared__ uint32_t wcount[CP_ROWS / 32];  // valid count per 32-row word
  const int64_t tile = blockIdx.x;
  const int64_t row0 = tile * CP_ROWS;
  const int64_t last_word = (bit_offset + n - 1) >> 5;
  const uint32_t vbase = tile_valid_excl[tile];
  const uint32_t nbase = (uint32_t)row0 - vbase;
  // 64 words per tile; thread t<64 loads word t
  uint32_t bits = 0;
  if (threadIdx.x < CP_ROWS / 32) {
    int64_t r = row0 + (int64_t)threadIdx.x * 32;
    if (r < n) {
      bits = load_mask_word_unaligned(mask, bit_offset + r, last_word);
      int64_t rem = n - r;
      if (rem < 32) bits &= (1u << rem) - 1u;
    }
    wcount[threadIdx.x] = __popc(bits);
  }
  __syncthreads();
  // each warp handles words warp, warp+8, ... ; prefix of earlier words by summing smem (<=64 adds)
  for (int w = threadIdx.x >> 5; w < CP_ROWS / 32; w += CP_THREADS / 32) {
    int64_t r = row0 + (int64_t)w * 32 + lane_id();
    uint32_t before = 0;
    for (int j = lane_id(); j < w; j += 32) before += wcount[j];
    before = warp_sum(before);
    bool in = r < n;
    bool valid = false;
    if (in) {
      // recompute this word's bits (cheap, L1-resident)
      uint32_t wb = load_mask_word_unaligned(mask, bit_offset + row0 + (int64_t)w * 32, last_word);
      valid = (wb >> lane_id()) & 1u;
    }
    unsigned vb = __ballot_sync(0xffffffffu, valid);
    unsigned nb = __ballot_sync(0xffffffffu, in && !valid);
    if (valid) {
      uint32_t o = vbase + before + __popc(vb & lanemask_lt());
      out_keys[o]      = twiddle_rt<UK>(keys[r], kind, desc_mask);
      out_valid_idx[o] = (int32_t)r;
    } else if (in) {
      uint32_t o = nbase + ((uint32_t)w * 32 - before) + __popc(nb & lanemask_lt());
      out_null_idx[o] = (int32_t)r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
struct tile_cfg { int threads; int ipt; };


template <typename UK, int T, int I, typename VT = uint32_t>
size_t onesweep_smem(bool range = false)
{
  // every instantiation declares the RANGE arrays behind s_misc, only RANGE launches allocate them
  return (sizeof(UK) > sizeof(VT) ? sizeof(UK) : sizeof(VT)) * (size_t)T * I + sizeof(uint32_t) * (2 * (T / 32) * RADIX + 3 * RADIX + 16) +
         (range ? (sizeof(UK) + 2 * sizeof(void*)) * RADIX + (size_t)T * I : 0);
}

int64_t portion_limit()
{
  static int64_t lim = [] {
    const char* e = std::getenv("B2_SORT_PORTION");
    int64_t v = e ? std::atoll(e) : 0;
    if (v <= 0) v = (int64_t(1) << 30) - 16384;
    return v;
  }();
  return lim;
}

// B2_SORT_HYBRID=0 switches the partial-LSD + fix-up plan off; B2_SORT_HYBRID_MIN=<rows> moves its lower size limit
// (tests run it on small inputs).
int64_t hybrid_min_rows()
{
  static int64_t v = [] {
    const char* off = std::getenv("B2_SORT_HYBRID");
    if (off && std::atoi(off) == 0) return INT64_MAX;
    const char* e = std::getenv("B2_SORT_HYBRID_MIN");
    return e ? (int64_t)std::atoll(e) : (int64_t(1) << 16);
  }();
  return v;
}

// Sort `n` keys.
//  raw_keys != nullptr : keys are the user's raw column (twiddled on load, implicit row ids)
//  raw_keys == nullptr : keys are pre-twiddled in bufA with explicit row ids in idx buffer pre_idx_buf
//  pairs: idx_out receives the permutation ; keys-only: bufA is the OUTPUT, bufB the temp.
template <typename UK, int T, int I, int MINB, typename VT = uint32_t, bool CARRY = false, bool MIX = false, bool SAFE = true,
          bool RMW = true, bool BULK = false>
void run_radix_cfg(const UK* raw_keys, UK* bufA, UK* bufB, int32_t* idx_out, int32_t* idx_tmp, int32_t* idx_tmp2, int pre_idx_buf, int64_t n,
                   int kind, bool descending, bool pairs, cudaStream_t stream, int first_pass = 0, int last_pass = 7,
                   bool keep_keys = false, const void* val_in = nullptr, uint32_t* top_digit_base_out = nullptr)
{
  constexpr int NP = sizeof(UK);
  constexpr int TILE = T * I;
  const bool raw = raw_keys != nullptr;
  const UK desc_mask = descending ? ~UK(0) : UK(0);

  const int64_t plim = std::max<int64_t>(TILE, portion_limit() / TILE * TILE);
  const int64_t nportions = (n + plim - 1) / plim;
  const int64_t tiles_per_portion = (std::min(n, plim) + TILE - 1) / TILE;

  // control block + histograms + status words + tile counters in one zeroed allocation
  const size_t ctl_bytes   = (sizeof(sort_ctl) + 255) / 256 * 256;
  const size_t hist_bytes  = sizeof(uint32_t) * NP * RADIX;
  const size_t cnt_bytes   = (sizeof(uint32_t) * NP * nportions + 255) / 256 * 256;
  const size_t status_per  = sizeof(uint32_t) * RADIX * (size_t)tiles_per_portion;
  const size_t status_bytes = status_per * NP * nportions;
  dbuf work(ctl_bytes + hist_bytes + cnt_bytes + status_bytes, stream);
  auto* ctl       = reinterpret_cast<sort_ctl*>(work.ptr);
  auto* ghist     = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes);
  auto* counters  = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes + hist_bytes);
  auto* status    = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes + hist_bytes + cnt_bytes);

  static std::atomic<uint64_t> attr_done{0};  // per device: the opt-in to > 48 KB of dynamic shared memory is a per-context setting
  once_per_device(attr_done, [] {
    B2_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<UK, T, I, MINB, VT, CARRY, MIX, SAFE, RMW, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)onesweep_smem<UK, T, I, VT>()));
  });

  // Hybrid plan (64-bit raw keys, full sort): LSD passes over the top digits only, then segment_fix_kernel. The plan
  // kernel decides on the device; the host learns the outcome from one 4-byte read-back after the fix-up.
  bool try_hybrid = sizeof(UK) == 8 && !MIX && raw && first_pass == 0 && last_pass >= NP - 1 && !keep_keys && n >= hybrid_min_rows();

  for (;;) {
    // control block, histograms and tile counters are zeroed here; the look-back rows of a (pass, portion) right before its launch, so
    // that passes the plan skips cost nothing (1e9 rows: 163 MB per executed pass instead of 1.3 GB per sort)
    B2_CUDA_TRY(cudaMemsetAsync(work.ptr, 0, ctl_bytes + hist_bytes + cnt_bytes, stream));
    {
      int grid = (int)std::min<int64_t>((n + 512 * 16 - 1) / (512 * 16), NUM_SMS_B200 * 4);
      grid = std::max(grid, 1);
      const UK* hkeys = raw ? raw_keys : bufA;
      uint32_t* nanp = (raw && kind == (int)key_kind::FLOAT) ? &ctl->nan_count : nullptr;
      // digits that can be executed at all: [first_pass, last_pass]
      const uint32_t span_mask = ((last_pass >= NP - 1 ? (1u << NP) : (1u << (last_pass + 1))) - 1u) & ~((1u << std::max(first_pass, 0)) - 1u);
      if (try_hybrid) {
        // two-phase: the top four digits first (half the shared-memory atomics); the low four only if the plan needs them
        {
          prof_scope ps("histogram", stream);
          B2_LAUNCH((histogram_kernel<UK, MIX>), grid, 512, 0, stream, hkeys, n, raw ? 1 : 0, kind, desc_mask, ghist, nanp, 0xf0u,
                    (const int32_t*)nullptr, &ctl->vary);
        }
        B2_LAUNCH(plan_kernel, 1, RADIX, 0, stream, ghist, NP, (uint32_t)n, raw ? 1 : 0, pre_idx_buf, ctl, first_pass, last_pass, 1, 1);
        {
          prof_scope ps("histogram", stream);
          B2_LAUNCH((histogram_kernel<UK, MIX>), grid, 512, 0, stream, hkeys, n, raw ? 1 : 0, kind, desc_mask, ghist, (uint32_t*)nullptr, 0x0fu,
                    &ctl->need_low, (unsigned long long*)nullptr);
        }
        B2_LAUNCH(plan_kernel, 1, RADIX, 0, stream, ghist, NP, (uint32_t)n, raw ? 1 : 0, pre_idx_buf, ctl, first_pass, last_pass, 1, 2);
      } else {
        {
          prof_scope ps("histogram", stream);
          B2_LAUNCH((histogram_kernel<UK, MIX>), grid, 512, 0, stream, hkeys, n, raw ? 1 : 0, kind, desc_mask, ghist, nanp, span_mask,
                    (const int32_t*)nullptr, (unsigned long long*)nullptr);
        }
        B2_LAUNCH(plan_kernel, 1, RADIX, 0, stream, ghist, NP, (uint32_t)n, raw ? 1 : 0, pre_idx_buf, ctl, first_pass, last_pass, 0, 0);
      }
    }

    {
      static const int force_fix = [] {  // test hook: B2_SORT_FIX_FAST=0/1 overrides the plan's choice of the fix-up flavour
        const char* e = std::getenv("B2_SORT_FIX_FAST");
        return e ? (std::atoi(e) != 0 ? 1 : 0) : -1;
      }();
      if (force_fix >= 0 && try_hybrid) B2_LAUNCH(set_fix_fast_kernel, 1, 1, 0, stream, ctl, force_fix);
    }
    pass_args a{};
    a.key_bufs[0] = raw_keys;
    a.key_bufs[1] = bufA;
    a.key_bufs[2] = bufB;
    a.idx_bufs[0] = idx_out;
    a.idx_bufs[1] = idx_tmp;
    a.idx_bufs[2] = idx_tmp2;
    a.ctl = ctl;
    a.kind = kind;
    a.pairs = pairs ? 1 : 0;
    a.keep_keys = keep_keys ? 1 : 0;
    a.val_in = val_in;
    a.desc_mask = (uint64_t)desc_mask;
    // Large inputs: read the plan back (one small copy + sync, ~20 us against milliseconds of passes) and launch only the
    // passes it executes — a skipped pass would otherwise start one CTA per tile just to return (163 K CTAs at 1e9 rows).
    // Small inputs launch every digit's kernel and stay free of host synchronisation.
    uint32_t skip_mask = 0;
    bool plan_hybrid = true;
    static const int64_t readback_min = [] {
      const char* e = std::getenv("B2_SORT_PLAN_READBACK_MIN");  // test hook
      return e ? (int64_t)std::atoll(e) : (int64_t(1) << 22);
    }();
    if (n >= readback_min) {
      pass_plan hp[8];
      int32_t hflag = 0;
      B2_CUDA_TRY(cudaMemcpyAsync(hp, &ctl->plan[0], sizeof(pass_plan) * 8, cudaMemcpyDeviceToHost, stream));
      B2_CUDA_TRY(cudaMemcpyAsync(&hflag, &ctl->hybrid, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      B2_CUDA_TRY(cudaStreamSynchronize(stream));
      for (int p = 0; p < NP; ++p)
        if (hp[p].trivial) skip_mask |= 1u << p;
      plan_hybrid = hflag != 0;
    }
    for (int p = std::max(0, first_pass); p < NP && p <= last_pass; ++p) {
      if ((skip_mask >> p) & 1u) continue;
      for (int64_t q = 0; q < nportions; ++q) {
        const int64_t start = q * plim;
        const int64_t pn = std::min(plim, n - start);
        a.pass = p;
        a.portion_start = start;
        a.portion_n = (uint32_t)pn;
        a.portion_parity = (int)(q & 1);
        a.has_next_portion = q + 1 < nportions;
        a.status = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(status) + (size_t)(p * nportions + q) * status_per);
        a.tile_counter = counters + p * nportions + q;
        const int64_t ntiles = (pn + TILE - 1) / TILE;
        const size_t smem_bytes = onesweep_smem<UK, T, I, VT>();
        B2_CUDA_TRY(cudaMemsetAsync(a.status, 0, sizeof(uint32_t) * RADIX * (size_t)ntiles, stream));
        prof_scope ps("onesweep", stream);
        B2_LAUNCH((onesweep_kernel<UK, T, I, MINB, VT, CARRY, MIX, SAFE, RMW, BULK>), (unsigned)ntiles, T + 32 * LBW, smem_bytes, stream, a);
      }
    }
    if constexpr (sizeof(UK) == 8 && !MIX) {
      if (try_hybrid && plan_hybrid) {
        const int64_t ntiles = (n + FIX_TILE - 1) / FIX_TILE;
        const int grid = (int)std::min<int64_t>(ntiles, NUM_SMS_B200 * 8);
        prof_scope ps("segment_fix", stream);
        B2_LAUNCH((segment_fix_kernel<UK, VT, 0>), grid, FIX_THREADS, 0, stream, a, n);  // the plan's ctl->fix_fast picks one of the two
        B2_LAUNCH((segment_fix_kernel<UK, VT, 3>), grid, FIX_THREADS, 0, stream, a, n);
      }
    }
    {
      int grid = (int)std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 8);
      B2_LAUNCH((finalize_kernel<UK, MIX>), std::max(grid, 1), 256, 0, stream, a, n, raw ? 1 : 0, pre_idx_buf, CARRY ? (int)sizeof(VT) : 0);
    }
    if (!try_hybrid || !plan_hybrid) break;
    uint32_t overflow = 0;
    B2_CUDA_TRY(cudaMemcpyAsync(&overflow, &ctl->overflow, sizeof(overflow), cudaMemcpyDeviceToHost, stream));
    B2_CUDA_TRY(cudaStreamSynchronize(stream));
    if (!overflow) break;
    try_hybrid = false;  // a segment was longer than the fix-up window (digits not independent): full LSD sort of the untouched input
  }
  if (pairs && raw && kind == (int)key_kind::FLOAT && descending) {
    B2_LAUNCH(reverse_nan_prefix_kernel, NUM_SMS_B200 * 4, 256, 0, stream, idx_out, &ctl->nan_count);
  }
  if (top_digit_base_out)  // start offset of every value of the most significant digit (partition boundaries)
    B2_CUDA_TRY(cudaMemcpyAsync(top_digit_base_out, &ctl->base[0][NP - 1][0], sizeof(uint32_t) * RADIX, cudaMemcpyDeviceToDevice, stream));
}

int sort_cfg_env()
{
  static int v = [] {
    const char* e = std::getenv("B2_SORT_CFG");
    return e ? std::atoi(e) : 0;
  }();
  return v;
}

template <typename UK>
void run_radix(const UK* raw_keys, UK* bufA, UK* bufB, int32_t* idx_out, int32_t* idx_tmp, int pre_idx_buf, int64_t n,
               int kind, bool descending, bool pairs, cudaStream_t stream, int32_t* idx_tmp2 = nullptr)
{
#define B2_RUN(T, I, MINB) \
  run_radix_cfg<UK, T, I, MINB>(raw_keys, bufA, bufB, idx_out, idx_tmp, idx_tmp2, pre_idx_buf, n, kind, descending, pairs, stream)
  if constexpr (sizeof(UK) == 8) {
    switch (sort_cfg_env()) {  // tuning knob (B2_SORT_CFG); 0 is the shipped default
      case 1: B2_RUN(256, 16, 2); break;
      case 2: B2_RUN(256, 16, 3); break;
      case 3: B2_RUN(384, 12, 2); break;
      case 4: B2_RUN(512, 16, 1); break;
      case 5: B2_RUN(512, 12, 1); break;
      case 6: B2_RUN(640, 12, 1); break;
      case 7: B2_RUN(256, 20, 2); break;
      case 8: B2_RUN(320, 16, 2); break;
      case 9: B2_RUN(320, 12, 2); break;
      case 10:  // default shape, round-1 ranking: no extra __syncwarp (relies on warp convergence), offsets by LDS + STS
        run_radix_cfg<UK, 384, 16, 2, uint32_t, false, false, false, false>(raw_keys, bufA, bufB, idx_out, idx_tmp, idx_tmp2, pre_idx_buf, n,
                                                                            kind, descending, pairs, stream);
        break;
      case 11:  // default shape, race-free ranking, offsets by LDS + STS
        run_radix_cfg<UK, 384, 16, 2, uint32_t, false, false, true, false>(raw_keys, bufA, bufB, idx_out, idx_tmp, idx_tmp2, pre_idx_buf, n,
                                                                           kind, descending, pairs, stream);
        break;
      case 12:  // default shape and ranking, key tiles by one bulk async copy (TMA 1-D) + mbarrier
        run_radix_cfg<UK, 384, 16, 2, uint32_t, false, false, true, true, true>(raw_keys, bufA, bufB, idx_out, idx_tmp, idx_tmp2, pre_idx_buf,
                                                                                n, kind, descending, pairs, stream);
        break;
      default: B2_RUN(384, 16, 2); break;
    }
  } else {
    B2_RUN(512, 16, 1);
  }
#undef B2_RUN
}

}  // namespace

// Stable partial sort of n 64-bit keys by their two most significant bytes (passes 6 and 7 only):
// keys_out / idx_out receive the keys and their original positions grouped by the 16-bit prefix.
// Used by the hash join to make build and probe walk the table region by region (L2 locality).
void radix_partition_top16(const uint64_t* keys_in, int64_t n, uint64_t* keys_out, int32_t* idx_out, cudaStream_t stream)
{
  dbuf b(sizeof(uint64_t) * n, stream), it(sizeof(int32_t) * n, stream);
  run_radix_cfg<uint64_t, 384, 16, 2>(keys_in, keys_out, b.as<uint64_t>(), idx_out, it.as<int32_t>(), nullptr, 0, n,
                                      (int)key_kind::UNSIGNED, false, true, stream, 6, 7, true);
}

// Same, but `packed_keys` are the raw packed join keys: the kernels apply mix64 on load (histogram of the two
// needed digits only), so keys_out receives mix64(key) grouped by its top 16 bits. EXPERIMENTAL (radix_join.cu).
void radix_partition_top16_mix(const uint64_t* packed_keys, int64_t n, uint64_t* keys_out, int32_t* idx_out, cudaStream_t stream)
{
  dbuf b(sizeof(uint64_t) * n, stream), it(sizeof(int32_t) * n, stream);
  run_radix_cfg<uint64_t, 384, 16, 2, uint32_t, false, true>(packed_keys, keys_out, b.as<uint64_t>(), idx_out, it.as<int32_t>(), nullptr, 0,
                                                             n, (int)key_kind::UNSIGNED, false, true, stream, 6, 7, true);
}

// One stable partition pass by the top byte of mix64(key), carrying one 4- or 8-byte payload column next to the mixed
// keys (hash groupby: rows of a group meet in one of 256 partitions; mix64 is undone with unmix64). part_base[d] =
// first row of partition d.
void radix_partition_mix_carry(const uint64_t* keys, const void* vals, int val_bytes, int64_t n, uint64_t* mixed_keys_out, void* vals_out,
                               uint32_t* part_base, cudaStream_t stream)
{
  dbuf b(sizeof(uint64_t) * n, stream), vt((size_t)val_bytes * n, stream);
  if (val_bytes == 8)
    run_radix_cfg<uint64_t, 384, 16, 2, uint64_t, true, true>(keys, mixed_keys_out, b.as<uint64_t>(), static_cast<int32_t*>(vals_out),
                                                               vt.as<int32_t>(), nullptr, 0, n, (int)key_kind::UNSIGNED, false, true, stream,
                                                               7, 7, true, vals, part_base);
  else
    run_radix_cfg<uint64_t, 384, 16, 2, uint32_t, true, true>(keys, mixed_keys_out, b.as<uint64_t>(), static_cast<int32_t*>(vals_out),
                                                               vt.as<int32_t>(), nullptr, 0, n, (int)key_kind::UNSIGNED, false, true, stream,
                                                               7, 7, true, vals, part_base);
}

namespace {
// plan of the single executed pass `pass` with estimated bases: digit d starts at d * cap
__global__ void est_plan_kernel(sort_ctl* ctl, int pass, uint32_t cap)
{
  const int d = threadIdx.x;  // 256 threads
  ctl->base[0][pass][d] = (uint32_t)d * cap;
  if (d < 8) {
    pass_plan pl{};
    pl.trivial = d != pass;
    if (d == pass) {
      pl.key_src = 0;
      pl.idx_src = -1;
      pl.key_dst = 1;
      pl.idx_dst = 0;
      pl.last    = 1;
      pl.hybrid  = 0;
    }
    ctl->plan[d] = pl;
  }
  if (d == 0) {
    ctl->any_pass = 1;
    ctl->hybrid   = 0;
    ctl->overflow = 0;
  }
}

// counts[b] = sampled rows (every `stride`-th) whose mix64(key) has top byte b
__global__ void __launch_bounds__(256) est_sample_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t stride, unsigned int* __restrict__ counts)
{
  __shared__ unsigned int sh[RADIX];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t m = (n + stride - 1) / stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh[(unsigned)(mix64(keys[i * stride]) >> 56)], 1u);
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sh[threadIdx.x]);
}

template <typename VT>
bool est_pass_impl(const uint64_t* keys, const void* vals, int64_t n, uint32_t cap, uint64_t* mixed_keys_out, void* vals_out, uint32_t* part_base,
                   uint32_t* part_end, cudaStream_t stream)
{
  constexpr int T = 384, I = 16, TILE = T * I, PASS = 7;
  using UK = uint64_t;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  const size_t ctl_bytes = (sizeof(sort_ctl) + 255) / 256 * 256;
  const size_t status_bytes = sizeof(uint32_t) * RADIX * (size_t)ntiles;
  dbuf work(ctl_bytes + 256 + status_bytes, stream);
  auto* ctl = reinterpret_cast<sort_ctl*>(work.ptr);
  auto* counter = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes);
  auto* status = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes + 256);
  static std::atomic<uint64_t> attr_done{0};
  once_per_device(attr_done, [] {
    B2_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<UK, T, I, 2, VT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)onesweep_smem<UK, T, I, VT>()));
  });
  B2_CUDA_TRY(cudaMemsetAsync(work.ptr, 0, work.bytes, stream));
  B2_LAUNCH(est_plan_kernel, 1, RADIX, 0, stream, ctl, PASS, cap);
  pass_args a{};
  a.key_bufs[0] = keys;
  a.key_bufs[1] = mixed_keys_out;
  a.key_bufs[2] = mixed_keys_out;
  a.idx_bufs[0] = static_cast<int32_t*>(vals_out);
  a.idx_bufs[1] = static_cast<int32_t*>(vals_out);
  a.idx_bufs[2] = static_cast<int32_t*>(vals_out);
  a.ctl = ctl;
  a.kind = (int)key_kind::UNSIGNED;
  a.pairs = 1;
  a.keep_keys = 1;
  a.val_in = vals;
  a.desc_mask = 0;
  a.pass = PASS;
  a.portion_start = 0;
  a.portion_n = (uint32_t)n;
  a.portion_parity = 0;
  a.has_next_portion = 0;
  a.status = status;
  a.tile_counter = counter;
  a.est_cap = cap;
  {
    prof_scope ps("onesweep", stream);
    B2_LAUNCH((onesweep_kernel<UK, T, I, 2, VT, true, true>), (unsigned)ntiles, T + 32 * LBW, (onesweep_smem<UK, T, I, VT>()), stream, a);
  }
  B2_CUDA_TRY(cudaMemcpyAsync(part_base, &ctl->base[0][PASS][0], sizeof(uint32_t) * RADIX, cudaMemcpyDeviceToDevice, stream));
  B2_CUDA_TRY(cudaMemcpyAsync(part_end, &ctl->base[1][PASS][0], sizeof(uint32_t) * RADIX, cudaMemcpyDeviceToDevice, stream));
  uint32_t overflow = 0;
  B2_CUDA_TRY(cudaMemcpyAsync(&overflow, &ctl->overflow, sizeof(overflow), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  return overflow == 0;
}
}  // namespace

// Rows per partition to reserve for the histogram-free partition pass below, or 0 when a strided sample of the keys says the
// partitions are too uneven for it (few distinct keys, hot keys): max sampled share + 5 sigma of the sampling noise must fit.
uint32_t radix_partition_est_capacity(const uint64_t* keys, int64_t n, cudaStream_t stream)
{
  static const int64_t min_rows = [] {
    const char* e = std::getenv("B2_GROUPBY_EST_MIN");  // test hook
    return e ? (int64_t)std::atoll(e) : (int64_t(1) << 22);
  }();
  if (n < min_rows || n > portion_limit()) return 0;
  if (const char* e = std::getenv("B2_GROUPBY_EST_CAP")) return (uint32_t)std::max(1, std::atoi(e));  // test hook: forces the overflow fallback
  const int64_t want = int64_t(1) << 20;
  const int64_t stride = std::max<int64_t>(1, n / want);
  const int64_t m = (n + stride - 1) / stride;
  dbuf cnt(sizeof(unsigned int) * RADIX, stream);
  B2_CUDA_TRY(cudaMemsetAsync(cnt.ptr, 0, cnt.bytes, stream));
  B2_LAUNCH(est_sample_kernel, NUM_SMS_B200 * 4, 256, 0, stream, keys, n, stride, cnt.as<unsigned int>());
  unsigned int h[RADIX];
  B2_CUDA_TRY(cudaMemcpyAsync(h, cnt.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  unsigned int mx = 0;
  for (int d = 0; d < RADIX; ++d) mx = std::max(mx, h[d]);
  const double cap = (double)(n / RADIX) * 1.25 + 4096.0;
  const double worst = ((double)mx + 5.0 * std::sqrt((double)mx) + 1.0) * ((double)n / (double)m);
  if (worst > cap || cap * RADIX >= 4.0e9) return 0;
  return (uint32_t)cap;
}

// radix_partition_mix_carry without the histogram: the pass runs with estimated bases (partition d owns rows [d * cap, (d + 1) * cap) of
// the outputs, which hold 256 * cap rows) and reports the exact end of every partition in part_end[d]. Returns false when a partition
// overflowed its range (the outputs are then incomplete: run radix_partition_mix_carry).
bool radix_partition_mix_carry_est(const uint64_t* keys, const void* vals, int val_bytes, int64_t n, uint32_t cap, uint64_t* mixed_keys_out,
                                   void* vals_out, uint32_t* part_base, uint32_t* part_end, cudaStream_t stream)
{
  if (val_bytes == 8) return est_pass_impl<uint64_t>(keys, vals, n, cap, mixed_keys_out, vals_out, part_base, part_end, stream);
  return est_pass_impl<uint32_t>(keys, vals, n, cap, mixed_keys_out, vals_out, part_base, part_end, stream);
}

namespace {
// bucket b of a key = number of splitters <= key (twiddled order); counts[b] += rows of bucket b
template <typename UK>
__global__ void __launch_bounds__(512) range_count_kernel(const UK* __restrict__ keys, int64_t n, int kind, const UK* __restrict__ splitters, int P,
                                                          unsigned long long* __restrict__ counts)
{
  __shared__ UK sp[RADIX];
  __shared__ unsigned int cnt[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += blockDim.x) {
    sp[i]  = (splitters != nullptr && i < P - 1) ? splitters[i] : ~UK(0);
    cnt[i] = 0;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const UK k = twiddle_rt<UK>(ld_stream(keys + i), kind, UK(0));
    int lo = 0, hi = P - 1;
    if (splitters == nullptr) lo = hi = (int)range_hash_bucket((uint64_t)k, P);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sp[mid] <= k) lo = mid + 1;
      else hi = mid;
    }
    atomicAdd(&cnt[lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x)
    if (cnt[i]) atomicAdd(&counts[i], (unsigned long long)cnt[i]);
}
template <typename UK>
__global__ void twiddle_splitters_kernel(const UK* __restrict__ raw, int m, int kind, UK* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = twiddle_rt<UK>(raw[i], kind, UK(0));
}
}  // namespace

// Range partition of one null-free 8-byte integer-like key column (and optionally one null-free 4- / 8-byte payload column) straight
// into P destination buffers — local or PEER memory (sharded sort: the partition pass is the bucket exchange).
// Step 1 (range_partition_counts): rows per bucket, so that the ranks can agree on where each one writes.
// Step 2 (range_partition_scatter): ONE one-sweep pass (stable; per-(tile, bucket) runs of ~6144 / P rows, i.e. 6 KB NVLink
// writes at P = 8) whose digit is the bucket; key_dst[b] / val_dst[b] = address of this rank's first row of bucket b.
void range_partition_counts(const b2_column_view& keys, const void* splitters, int P, int64_t* out_counts, cudaStream_t stream)
{
  using UK = uint64_t;
  const int64_t n = keys.size;
  const int kind = is_signed_id(storage_type(keys.type_id)) ? (int)key_kind::SIGNED : (int)key_kind::UNSIGNED;
  for (int b = 0; b < P; ++b) out_counts[b] = 0;
  if (n == 0) return;
  dbuf cnt(sizeof(unsigned long long) * RADIX, stream), sp(sizeof(UK) * RADIX, stream);
  B2_CUDA_TRY(cudaMemsetAsync(cnt.ptr, 0, cnt.bytes, stream));
  const bool hashed = splitters == nullptr;  // no splitters: hash partition (range_hash_bucket)
  if (P > 1 && !hashed) B2_LAUNCH((twiddle_splitters_kernel<UK>), 1, RADIX, 0, stream, static_cast<const UK*>(splitters), P - 1, kind, sp.as<UK>());
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 511) / 512, NUM_SMS_B200 * 4));
  {
    prof_scope ps("range_count", stream);
    B2_LAUNCH((range_count_kernel<UK>), grid, 512, 0, stream, static_cast<const UK*>(keys.data) + keys.offset, n, kind,
              hashed ? static_cast<const UK*>(nullptr) : sp.as<UK>(), P, cnt.as<unsigned long long>());
  }
  unsigned long long h[RADIX];
  B2_CUDA_TRY(cudaMemcpyAsync(h, cnt.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_TRY(cudaStreamSynchronize(stream));
  for (int b = 0; b < P; ++b) out_counts[b] = (int64_t)h[b];
}

template <typename VT, bool CARRY>
static void range_scatter_impl(const b2_column_view& keys, const void* vals, const void* splitters, int P, void* const* key_dst, void* const* val_dst,
                               cudaStream_t stream)
{
  using UK = uint64_t;
  constexpr int T = 384, I = 16, TILE = T * I;
  const int64_t n = keys.size;
  const int kind = is_signed_id(storage_type(keys.type_id)) ? (int)key_kind::SIGNED : (int)key_kind::UNSIGNED;
  const int64_t plim = std::max<int64_t>(TILE, portion_limit() / TILE * TILE);
  const int64_t nportions = (n + plim - 1) / plim;
  const int64_t tiles_per_portion = (std::min(n, plim) + TILE - 1) / TILE;
  const size_t ctl_bytes = (sizeof(sort_ctl) + 255) / 256 * 256;
  const size_t cnt_bytes = (sizeof(uint32_t) * nportions + 255) / 256 * 256;
  const size_t status_per = sizeof(uint32_t) * RADIX * (size_t)tiles_per_portion;
  const size_t tab_bytes = sizeof(UK) * RADIX + 2 * sizeof(void*) * RADIX;
  dbuf work(ctl_bytes + cnt_bytes + status_per * nportions + tab_bytes, stream);
  B2_CUDA_TRY(cudaMemsetAsync(work.ptr, 0, work.bytes, stream));
  auto* ctl      = reinterpret_cast<sort_ctl*>(work.ptr);
  auto* counters = reinterpret_cast<uint32_t*>(static_cast<char*>(work.ptr) + ctl_bytes);
  auto* status   = static_cast<char*>(work.ptr) + ctl_bytes + cnt_bytes;
  auto* d_split  = reinterpret_cast<UK*>(status + status_per * nportions);
  auto* d_kdst   = reinterpret_cast<void**>(d_split + RADIX);
  auto* d_vdst   = d_kdst + RADIX;
  if (P > 1 && splitters != nullptr) B2_LAUNCH((twiddle_splitters_kernel<UK>), 1, RADIX, 0, stream, static_cast<const UK*>(splitters), P - 1, kind, d_split);
  B2_CUDA_TRY(cudaMemcpyAsync(d_kdst, key_dst, sizeof(void*) * P, cudaMemcpyHostToDevice, stream));
  if (CARRY) B2_CUDA_TRY(cudaMemcpyAsync(d_vdst, val_dst, sizeof(void*) * P, cudaMemcpyHostToDevice, stream));
  // the plan of the single pass 0: executed, raw keys, last pass (keys-only mode untwiddles what it writes); base = 0: positions
  // count from the start of each bucket's run
  pass_plan pl{};
  pl.trivial = 0; pl.key_src = 0; pl.key_dst = 1; pl.idx_src = -1; pl.idx_dst = 0; pl.last = 1; pl.hybrid = 0;
  B2_CUDA_TRY(cudaMemcpyAsync(&ctl->plan[0], &pl, sizeof(pl), cudaMemcpyHostToDevice, stream));

  static std::atomic<uint64_t> attr_done{0};
  once_per_device(attr_done, [] {
    B2_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<UK, T, I, 2, VT, CARRY, false, true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)onesweep_smem<UK, T, I, VT>(true)));
  });
  pass_args a{};
  a.key_bufs[0] = static_cast<const UK*>(keys.data) + keys.offset;
  a.ctl = ctl;
  a.kind = kind;
  a.pairs = CARRY ? 1 : 0;
  a.val_in = vals;
  a.desc_mask = 0;
  a.range_splitters = d_split;
  a.range_key_dst = d_kdst;
  a.range_val_dst = d_vdst;
  a.range_parts = P;
  a.range_hash = (P > 1 && splitters == nullptr) ? 1 : 0;
  a.pass = 0;
  for (int64_t q = 0; q < nportions; ++q) {
    const int64_t start = q * plim;
    const int64_t pn = std::min(plim, n - start);
    a.portion_start = start;
    a.portion_n = (uint32_t)pn;
    a.portion_parity = (int)(q & 1);
    a.has_next_portion = q + 1 < nportions;
    a.status = reinterpret_cast<uint32_t*>(status + (size_t)q * status_per);
    a.tile_counter = counters + q;
    const int64_t ntiles = (pn + TILE - 1) / TILE;
    prof_scope ps("range_scatter", stream);
    const size_t smem_bytes = onesweep_smem<UK, T, I, VT>(true);
    B2_LAUNCH((onesweep_kernel<UK, T, I, 2, VT, CARRY, false, true, true, false, true>), (unsigned)ntiles, T + 32 * LBW, smem_bytes, stream, a);
  }
}

void range_partition_scatter(const b2_column_view& keys, const b2_column_view* values, const void* splitters, int P, void* const* key_dst,
                             void* const* val_dst, cudaStream_t stream)
{
  if (keys.size == 0) return;
  if (values == nullptr) return range_scatter_impl<uint32_t, false>(keys, nullptr, splitters, P, key_dst, nullptr, stream);
  const int vw = type_width(values->type_id);
  const void* vin = static_cast<const char*>(values->data) + (size_t)values->offset * vw;
  if (vw == 8) range_scatter_impl<uint64_t, true>(keys, vin, splitters, P, key_dst, val_dst, stream);
  else range_scatter_impl<uint32_t, true>(keys, vin, splitters, P, key_dst, val_dst, stream);
}

namespace {

int kind_of(int32_t storage_id)
{
  if (is_float_id(storage_id)) return (int)key_kind::FLOAT;
  if (is_signed_id(storage_id)) return (int)key_kind::SIGNED;
  return (int)key_kind::UNSIGNED;  // unsigned ints and BOOL8
}

template <typename UK>
column_ptr sorted_order_single(const b2_column_view& col, bool ascending, bool nulls_before, cudaStream_t stream)
{
  const int64_t n = col.size;
  const int sid   = storage_type(col.type_id);
  const int kind  = kind_of(sid);
  auto out = make_column(B2_INT32, col.size, false, stream);
  const UK* data = static_cast<const UK*>(col.data) + col.offset;
  int32_t* out_idx = out->data.as<int32_t>();

  if (!has_nulls(col)) {
    dbuf a(sizeof(UK) * n, stream), b(sizeof(UK) > 1 ? sizeof(UK) * n : 0, stream), it(sizeof(int32_t) * n, stream);
    run_radix<UK>(data, a.as<UK>(), b.as<UK>(), out_idx, it.as<int32_t>(), 0, n, kind, !ascending, true, stream);
    return out;
  }
  // nullable: nulls first iff (null_order == BEFORE) xor descending (sort_column_impl.cuh:35-57)
  const bool nulls_first = nulls_before != !ascending;
  const int64_t n_null  = col.null_count;
  const int64_t n_valid = n - n_null;
  const int64_t ntiles = (n + CP_ROWS - 1) / CP_ROWS;
  dbuf tv(sizeof(uint32_t) * (ntiles + 1), stream);
  B2_LAUNCH(valid_count_kernel, (unsigned)((ntiles + 7) / 8), CP_THREADS, 0, stream, col.null_mask, (int64_t)col.offset, n,
            tv.as<uint32_t>());
  B2_LAUNCH(scan_tiles_kernel, 1, 1024, 0, stream, tv.as<uint32_t>(), ntiles, tv.as<uint32_t>() + ntiles);
  dbuf a(sizeof(UK) * std::max<int64_t>(n_valid, 1), stream), b(sizeof(UK) * std::max<int64_t>(n_valid, 1), stream);
  dbuf it(sizeof(int32_t) * std::max<int64_t>(n_valid, 1), stream), it2(sizeof(int32_t) * std::max<int64_t>(n_valid, 1), stream);
  int32_t* valid_out = out_idx + (nulls_first ? n_null : 0);
  int32_t* null_out  = out_idx + (nulls_first ? 0 : n_valid);
  const UK desc_mask = ascending ? UK(0) : ~UK(0);
  // compaction writes explicit row ids into the TEMP idx buffer (buffer 1); the plan then makes the
  // last executed pass land in buffer 0 = valid_out.
  B2_LAUNCH((compact_kernel<UK>), (unsigned)ntiles, CP_THREADS, 0, stream, data, col.null_mask, (int64_t)col.offset, n,
            kind, desc_mask, tv.as<uint32_t>(), a.as<UK>(), it.as<int32_t>(), null_out);
  if (n_valid > 0)
    run_radix<UK>(nullptr, a.as<UK>(), b.as<UK>(), valid_out, it.as<int32_t>(), 1, n_valid, kind, !ascending, true, stream,
                  it2.as<int32_t>());
  return out;
}

}  // namespace

bool is_radix_sortable(const b2_column_view& c) { return !has_nulls(c) && is_fixed_width(c.type_id); }

// cudf::detail::sorted_order(column_view) — cpp/src/sort/sort_column.cu:22-44
static column_ptr sorted_order_column(const b2_column_view& col, bool ascending, bool nulls_before, cudaStream_t stream)
{
  switch (type_width(col.type_id)) {
    case 1: return sorted_order_single<uint8_t>(col, ascending, nulls_before, stream);
    case 2: return sorted_order_single<uint16_t>(col, ascending, nulls_before, stream);
    case 4: return sorted_order_single<uint32_t>(col, ascending, nulls_before, stream);
    case 8: return sorted_order_single<uint64_t>(col, ascending, nulls_before, stream);
    default: B2_FAIL(B2_ERR_DATA_TYPE, "sorted_order: unsupported (non fixed-width) key type");
  }
}

// ---- multi-column lexicographic order (sort_impl.cuh:61-93): LSD over columns, last to first ----
namespace {
template <typename UK>
__global__ void gather_twiddle_kernel(const UK* __restrict__ keys, const uint32_t* __restrict__ mask, int64_t bit_offset,
                                      const int32_t* __restrict__ perm, int64_t n, int kind, UK desc_mask,
                                      UK* __restrict__ out_keys, uint8_t* __restrict__ out_null)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int32_t r = perm ? perm[i] : (int32_t)i;
    bool valid = mask == nullptr || bit_is_set(mask, bit_offset + r);
    out_keys[i] = valid ? twiddle_rt<UK>(keys[r], kind, desc_mask) : UK(0);
    if (out_null) out_null[i] = valid ? 0 : 1;
  }
}
// stable 2-way partition of perm by flag (0 first when zero_first) — used for the null "digit"
__global__ void flag_count_kernel(const uint8_t* __restrict__ flag, int64_t n, uint32_t* __restrict__ tile_ones)
{
  // one warp per 1024 elements
  const int64_t tile = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int64_t ntiles = (n + 1023) / 1024;
  if (tile >= ntiles) return;
  uint32_t c = 0;
  for (int j = 0; j < 32; ++j) {
    int64_t i = tile * 1024 + j * 32 + lane_id();
    c += (i < n && flag[i]) ? 1u : 0u;
  }
  c = warp_sum(c);
  if (lane_id() == 0) tile_ones[tile] = c;
}
__global__ void flag_partition_kernel(const uint8_t* __restrict__ flag, const int32_t* __restrict__ perm_in, int64_t n,
                                      const uint32_t* __restrict__ tile_ones_excl, const uint32_t* __restrict__ total_ones,
                                      int ones_first, int32_t* __restrict__ perm_out)
{
  const int64_t tile = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int64_t ntiles = (n + 1023) / 1024;
  if (tile >= ntiles) return;
  const uint32_t tot1 = *total_ones;
  const uint32_t tot0 = (uint32_t)n - tot1;
  uint32_t ones_before = tile_ones_excl[tile];
  uint32_t zeros_before = (uint32_t)(tile * 1024) - ones_before;
  for (int j = 0; j < 32; ++j) {
    int64_t i = tile * 1024 + j * 32 + lane_id();
    bool in = i < n;
    bool f = in && flag[i];
    unsigned b1 = __ballot_sync(0xffffffffu, f);
    unsigned b0 = __ballot_sync(0xffffffffu, in && !f);
    if (in) {
      uint32_t o = f ? (ones_first ? 0u : tot0) + ones_before + __popc(b1 & lanemask_lt())
                     : (ones_first ? tot1 : 0u) + zeros_before + __popc(b0 & lanemask_lt());
      perm_out[o] = perm_in[i];
    }
    ones_before += __popc(b1);
    zeros_before += __popc(b0);
  }
}
__global__ void compose_perm_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ order, int64_t n,
                                    int32_t* __restrict__ out)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = perm[order[i]];
}

template <typename UK>
void lex_step(const b2_column_view& col, bool ascending, bool nulls_before, int32_t*& perm, int32_t*& perm_alt,
              bool& have_perm, dbuf& ka, dbuf& kb, dbuf& kc, dbuf& ord, dbuf& otmp, dbuf& nullflag, dbuf& tiles, cudaStream_t stream)
{
  const int64_t n = col.size;
  const int kind  = kind_of(storage_type(col.type_id));
  const UK desc_mask = ascending ? UK(0) : ~UK(0);
  const bool nullable = has_nulls(col);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, NUM_SMS_B200 * 16);
  B2_LAUNCH((gather_twiddle_kernel<UK>), grid, 256, 0, stream, static_cast<const UK*>(col.data) + col.offset,
            nullable ? col.null_mask : nullptr, (int64_t)col.offset, have_perm ? perm : nullptr, n, kind, desc_mask,
            ka.as<UK>(), nullable ? nullflag.as<uint8_t>() : nullptr);
  // sort positions 0..n-1 by the gathered key: explicit ids = iota in buffer 1 (otmp), result in ord
  // (raw=false path wants explicit ids; generate iota cheaply through the same compose kernel is
  // not possible, so run with pre-twiddled keys + implicit ids by treating them as UNSIGNED raw.)
  run_radix<UK>(ka.as<UK>(), kb.as<UK>(), kc.as<UK>(), ord.as<int32_t>(), otmp.as<int32_t>(), 0, n,
                (int)key_kind::UNSIGNED, false, true, stream);
  // perm' = perm ∘ ord
  if (have_perm) {
    B2_LAUNCH(compose_perm_kernel, grid, 256, 0, stream, perm, ord.as<int32_t>(), n, perm_alt);
    std::swap(perm, perm_alt);
  } else {
    B2_CUDA_TRY(cudaMemcpyAsync(perm, ord.ptr, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, stream));
    have_perm = true;
  }
  if (nullable) {
    // null flag of the rows in the new order, then a stable partition (the null "digit")
    const bool nulls_first = nulls_before != !ascending;
    const int64_t ntiles = (n + 1023) / 1024;
    // flags in new order: gather nullflag (indexed by old position) through ord
    // reuse kb as byte scratch
    uint8_t* f2 = kb.as<uint8_t>();
    // f2[i] = nullflag[ord[i]]
    B2_LAUNCH((gather_twiddle_kernel<uint8_t>), grid, 256, 0, stream, nullflag.as<uint8_t>(), (const uint32_t*)nullptr,
              (int64_t)0, ord.as<int32_t>(), n, (int)key_kind::UNSIGNED, (uint8_t)0, f2, (uint8_t*)nullptr);
    B2_LAUNCH(flag_count_kernel, (unsigned)((ntiles + 7) / 8), 256, 0, stream, f2, n, tiles.as<uint32_t>());
    B2_LAUNCH(scan_tiles_kernel, 1, 1024, 0, stream, tiles.as<uint32_t>(), ntiles, tiles.as<uint32_t>() + ntiles);
    B2_LAUNCH(flag_partition_kernel, (unsigned)((ntiles + 7) / 8), 256, 0, stream, f2, perm, n, tiles.as<uint32_t>(),
              tiles.as<uint32_t>() + ntiles, nulls_first ? 1 : 0, perm_alt);
    std::swap(perm, perm_alt);
  }
}
}  // namespace

// cudf::detail::sorted_order(table_view) — cpp/src/sort/sort_impl.cuh:31-96
column_ptr sorted_order(const std::vector<b2_column_view>& keys, const std::vector<uint8_t>& order,
                        const std::vector<uint8_t>& null_prec, bool /*stable: every path here is stable*/,
                        cudaStream_t stream)
{
  if (keys.empty() || keys[0].size == 0) return make_column(B2_INT32, 0, false, stream);
  B2_EXPECTS(order.empty() || order.size() == keys.size(), B2_ERR_LOGIC,
             "Mismatch between number of columns and column order.");
  B2_EXPECTS(null_prec.empty() || null_prec.size() == keys.size(), B2_ERR_LOGIC,
             "Mismatch between number of columns and null_precedence size.");
  auto asc = [&](size_t i) { return order.empty() ? true : order[i] == B2_ASCENDING; };
  auto before = [&](size_t i) { return null_prec.empty() ? true : null_prec[i] == B2_NULL_BEFORE; };
  if (keys.size() == 1) return sorted_order_column(keys[0], asc(0), before(0), stream);

  const int64_t n = keys[0].size;
  auto out = make_column(B2_INT32, (int32_t)n, false, stream);
  dbuf alt(sizeof(int32_t) * n, stream), ka(8 * n, stream), kb(8 * n, stream), kc(8 * n, stream), ord(sizeof(int32_t) * n, stream),
    otmp(sizeof(int32_t) * n, stream), nullflag(n, stream), tiles(sizeof(uint32_t) * ((n + 1023) / 1024 + 1), stream);
  int32_t* perm = out->data.as<int32_t>();
  int32_t* perm_alt = alt.as<int32_t>();
  bool have_perm = false;
  for (size_t c = keys.size(); c-- > 0;) {
    const auto& col = keys[c];
    switch (type_width(col.type_id)) {
      case 1: lex_step<uint8_t>(col, asc(c), before(c), perm, perm_alt, have_perm, ka, kb, kc, ord, otmp, nullflag, tiles, stream); break;
      case 2: lex_step<uint16_t>(col, asc(c), before(c), perm, perm_alt, have_perm, ka, kb, kc, ord, otmp, nullflag, tiles, stream); break;
      case 4: lex_step<uint32_t>(col, asc(c), before(c), perm, perm_alt, have_perm, ka, kb, kc, ord, otmp, nullflag, tiles, stream); break;
      case 8: lex_step<uint64_t>(col, asc(c), before(c), perm, perm_alt, have_perm, ka, kb, kc, ord, otmp, nullflag, tiles, stream); break;
      default: B2_FAIL(B2_ERR_DATA_TYPE, "sorted_order: unsupported (non fixed-width) key type");
    }
  }
  if (perm != out->data.as<int32_t>()) {
    B2_CUDA_TRY(cudaMemcpyAsync(out->data.ptr, perm, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, stream));
    // `alt` now aliases the live result until the copy has run; it is freed stream-ordered after it.
  }
  return out;
}

// sort_by_key of ONE non-null 4- or 8-byte values column by ONE non-null fixed-width key column carries the payload
// through the passes instead of row ids: no random-access gather afterwards (B200 fetches 128 bytes per random 8-byte
// read: 26 ms per 1e9 rows). B2_SORT_CARRY=0 selects the row-id + gather path for every shape.
bool sort_carry_enabled()
{
  static bool v = [] {
    const char* e = std::getenv("B2_SORT_CARRY");
    return !e || std::atoi(e) != 0;
  }();
  return v;
}
bool sort_carry_applicable(const b2_column_view& keys, const b2_column_view& values, bool ascending)
{
  const int vw = type_width(values.type_id);
  const bool float_desc = is_float_id(storage_type(keys.type_id)) && !ascending;  // NaN-prefix reversal works on row ids only
  return sort_carry_enabled() && is_radix_sortable(keys) && !has_nulls(values) && (vw == 4 || vw == 8) && !float_desc &&
         keys.size == values.size && keys.size > 0;
}
column_ptr sort_by_key_carry(const b2_column_view& keys, const b2_column_view& values, bool ascending, cudaStream_t stream)
{
  const int64_t n = keys.size;
  const int kind = kind_of(storage_type(keys.type_id));
  auto out = make_column(values.type_id, (int32_t)n, false, stream);
  const int vw = type_width(values.type_id);
  dbuf vtmp((size_t)vw * n, stream);
  const void* vin = static_cast<const char*>(values.data) + (size_t)values.offset * vw;
  auto go = [&](auto ktag, auto vtag) {
    using UK = decltype(ktag);
    using VT = decltype(vtag);
    dbuf a(sizeof(UK) * n, stream), b(sizeof(UK) > 1 ? sizeof(UK) * n : 0, stream);
    if constexpr (sizeof(UK) == 8) {
      // B2_SORT_CFG on the payload-carrying kernel: 10 = round-1 ranking (no extra __syncwarp, offsets by LDS + STS), 11 = race-free
      // ranking with LDS + STS offsets; the default is race-free + one ATOMS.ADD per digit run (measured: 7.49 vs 7.85 ms per pass)
      switch (sort_cfg_env()) {
        case 10:
          run_radix_cfg<UK, 384, 16, 2, VT, true, false, false, false>(static_cast<const UK*>(keys.data) + keys.offset, a.as<UK>(), b.as<UK>(),
                                                                       out->data.as<int32_t>(), vtmp.as<int32_t>(), nullptr, 0, n, kind,
                                                                       !ascending, true, stream, 0, 7, false, vin);
          break;
        case 11:
          run_radix_cfg<UK, 384, 16, 2, VT, true, false, true, false>(static_cast<const UK*>(keys.data) + keys.offset, a.as<UK>(), b.as<UK>(),
                                                                      out->data.as<int32_t>(), vtmp.as<int32_t>(), nullptr, 0, n, kind,
                                                                      !ascending, true, stream, 0, 7, false, vin);
          break;
        default:
          run_radix_cfg<UK, 384, 16, 2, VT, true>(static_cast<const UK*>(keys.data) + keys.offset, a.as<UK>(), b.as<UK>(),
                                                  out->data.as<int32_t>(), vtmp.as<int32_t>(), nullptr, 0, n, kind, !ascending, true,
                                                  stream, 0, 7, false, vin);
      }
    } else
      run_radix_cfg<UK, 512, 16, 1, VT, true>(static_cast<const UK*>(keys.data) + keys.offset, a.as<UK>(), b.as<UK>(),
                                              out->data.as<int32_t>(), vtmp.as<int32_t>(), nullptr, 0, n, kind, !ascending, true,
                                              stream, 0, 7, false, vin);
  };
  auto by_key = [&](auto vtag) {
    switch (type_width(keys.type_id)) {
      case 1: go(uint8_t{}, vtag); break;
      case 2: go(uint16_t{}, vtag); break;
      case 4: go(uint32_t{}, vtag); break;
      default: go(uint64_t{}, vtag); break;
    }
  };
  if (vw == 8) by_key(uint64_t{}); else by_key(uint32_t{});
  return out;
}

// cudf::detail::sort_radix — cpp/src/sort/sort_radix.cu:151-161 (integer / chrono / bool keys;
// float columns go through sorted_order + gather so that NaN payloads and -0/+0 survive)
column_ptr sort_single_column(const b2_column_view& col, bool ascending, cudaStream_t stream)
{
  const int64_t n = col.size;
  auto out = make_column(col.type_id, col.size, false, stream);
  if (n == 0) return out;
  const int kind = kind_of(storage_type(col.type_id));
  B2_EXPECTS(kind != (int)key_kind::FLOAT, B2_ERR_LOGIC, "keys-only radix path is for integer-like keys");
  auto run = [&](auto tag) {
    using UK = decltype(tag);
    dbuf tmp(sizeof(UK) > 1 ? sizeof(UK) * n : 0, stream);
    run_radix<UK>(static_cast<const UK*>(col.data) + col.offset, out->data.as<UK>(), tmp.as<UK>(), nullptr, nullptr, 0, n, kind,
                  !ascending, false, stream);
  };
  switch (type_width(col.type_id)) {
    case 1: run(uint8_t{}); break;
    case 2: run(uint16_t{}); break;
    case 4: run(uint32_t{}); break;
    case 8: run(uint64_t{}); break;
    default: B2_FAIL(B2_ERR_DATA_TYPE, "sort: unsupported key type");
  }
  return out;
}

}  // namespace b2
