// passage_common.cuh — device-side definitions shared by the ahead-of-time build (nvcc, passage.cu) and the run-time
// specialisation of the throughput kernel (NVRTC, see specialise.inl).  No host headers in here when
// STRYKE_NVRTC is defined: NVRTC has no libc, so the fixed-width integer types and the few ABI constants the kernels
// need are defined locally.  The struct layouts below are the launch ABI between the host code and BOTH compilations.
#pragma once

#ifdef STRYKE_NVRTC
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#define STRYKE_KIND_APRIORI 0
#define STRYKE_KIND_KAPLAN 1
#define STRYKE_MAX_MOVES 64
#define STRYKE_MAX_PAIRS 128
#define STRYKE_UNIT_STATIC_W 4
#define STRYKE_UNIT_COEF_W 8
#define STRYKE_SPECIES_W 20
#define STRYKE_DAILY_W 5
#define STRYKE_FLAG_NO_PROMOTE 1u
typedef struct StrykeCellGroup {
  int64_t cell0;
  int32_t iterations, n_active_days, day0, species, active_off, reserved;
  uint64_t id_base;
} StrykeCellGroup;
#else
#include <stdint.h>
#include "../../include/stryke_b200.h"
#endif

namespace stryke {

// ------------------------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11), hand-written for the passage kernels.
// Counter = (fish index, cell id lo, cell id hi, block number), key = 64-bit run seed, so every draw of every fish
// is a pure function of (seed, scenario/species/iteration/day cell id, fish, slot) and independent of how cells
// are partitioned over GPUs.
// ------------------------------------------------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;

  __host__ __device__ static __forceinline__ void round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                        uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }

  // 10 rounds; key schedule k += (W0, W1) per round (warp-uniform, so it lives in uniform registers / immediates).
  __host__ __device__ static __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c0, c1, c2, c3, k0, k1);
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }

  // the same 10 rounds with the key schedule precomputed by the host (ks[2r], ks[2r+1]): when ks lives in the kernel
  // parameter bank the round keys are constant-bank operands of the XORs instead of 18 uniform adds per block
  __device__ static __forceinline__ void block_ks(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                  const uint32_t (&ks)[20], uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) round(c0, c1, c2, c3, ks[2 * r], ks[2 * r + 1]);
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
#ifndef STRYKE_NVRTC
  __host__ static inline void key_schedule(uint64_t seed, uint32_t ks[20]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { ks[2 * r] = k0; ks[2 * r + 1] = k1; k0 += W0; k1 += W1; }
  }
#endif
};

// ------------------------------------------------------------------------------------------------------------------
// device-side table records
// ------------------------------------------------------------------------------------------------------------------
struct __align__(4) NodeRec {   // 12 bytes
  float apriori32;              // np.float32(surv_dict[node]) (stryke.py:1405, :1575)
  int16_t edge_off;
  uint8_t deg;
  uint8_t kind;
  int8_t unit;
  uint8_t hasU;
  uint16_t pad;
};

enum : uint8_t { NEED_DICE = 1, NEED_RR = 2, NEED_DEPTH = 4, NEED_CAUSE = 8, NEED_ROUTE = 16, NEED_TRIVIAL = 32, NEED_UCHECK = 64 };
// NEED_UCHECK (throughput kernels): some node of the level has a 'U' in its name, i.e. the level can decide entrainment
// NEED_TRIVIAL (throughput kernel only): the level is one a-priori node with survival 1.0, <= 1 successor, no 'U'

// Native-RNG word budget.  A fish consumes Philox words in a fixed order — word 0: length (Box-Muller radius from the high
// 16 bits, angle from the low 16 bits); then per move, only where the level can need them (NEED_*): one word for the
// runner position r/R (high 16 bits) and the mortality-cause roll (low 16 bits), one for the survival dice, one for the
// routing draw (23 bits each); the barotrauma depth draw is made of spare bits of the first two (u13_depth below).  Word slot s lives in Philox block s >> 2 at position
// s & 3; the slots are assigned densely by the host, so the C2 facility needs 4 words = ONE block per fish.
struct __align__(4) MoveRec {   // 16 bytes
  uint8_t lex;                  // rank of "state_k" in lexicographic column order (stryke.py:3272)
  uint8_t need;                 // NEED_* : which draws this move can consume (static facility property)
  uint16_t slot_rc;             // word slot of the r/R + cause word
  uint16_t level_begin, level_end;  // reachable pairs of this move in pair_node[]
  uint16_t slot_route;          // word slot of the routing draw
  uint8_t max_deg;              // largest out-degree among the nodes of this level
  uint8_t pad;                  // throughput kernel: node of the level's first pair
  uint16_t slot_depth;          // unused (the depth draw takes spare bits of the r/R + cause and dice words)
  uint16_t slot_dice;           // word slot of the survival dice
};

struct InjectedDev {
  const int64_t* fish_off;
  const double *presence, *count_draw, *length_z, *dice, *rR, *depth, *cause, *route;
  const double *ghost_rR, *ghost_depth, *ghost_cause;
};
struct TraceDev {
  float* length_ft; double* length_ft64; uint8_t* state; float* rate; uint8_t* survival; double* strike; double* baro;
};

// Compact tally rows (plan_launch_compact): ONE row per cell that holds fish, nothing for the others.  A row is a slot of
// row_w = n_pairs + 3 words.  Narrow layout (n <= narrow_max, so that no counter can pass 65 535): 16-bit counters in pairs,
//   w0 = pop_size | mortality_barotrauma << 16,  w1 = num_entrained | num_survived << 16,
//   w2 = mortality_impingement | mortality_blade_strike << 16,  w[3 + p] = successes_p | count_p << 16.
// Wide layout (larger cells; TWO consecutive slots = 2 n_pairs + 6 words): 32-bit counters,
//   w0 = pop_size, w1..w5 = the five Daily columns, w[6 + 2p] = successes_p, w[7 + 2p] = count_p.
// Per aligned block of 32 cells: blk_row = slot of its first present cell, blk_mask = cells with fish, blk_wide = cells in
// the wide layout.  count_kernel zeroes the rows and writes pop_size; the passage kernels add with red.global.
struct RowsDev {
  uint32_t* rows;               // nullptr = dense output (KParams::daily / state_daily)
  uint32_t* blk_row; uint32_t* blk_mask; uint32_t* blk_wide;
  int row_w, narrow_max;
  unsigned int cap;             // slots the rows buffer holds
};

struct KParams {
  const NodeRec* nodes; const uint8_t* edge_dst; const MoveRec* moves; const uint8_t* pair_node;
  const double* unit_static;
  int n_nodes, n_units, n_moves, n_edges, n_pairs, start_node, barotrauma;
  const double* edge_cdf; const uint8_t* node_stay; const double* unit_coef;
  const double* species; int n_species;
  const int32_t* cell_day; const int32_t* cell_species; const uint64_t* cell_id;
  const int32_t* cell_n; const int64_t* tile_off; const int64_t* fish_off; int64_t n_cells;
  uint32_t* daily; uint32_t* state_daily;
  uint32_t seed_lo, seed_hi;
  int64_t n_fish_total;
  InjectedDev inj; TraceDev trace;
  uint32_t ks[20];                    // Philox key schedule of (seed_lo, seed_hi), throughput kernels
  unsigned long long* work_counter;   // throughput kernels: next chunk of tiles to hand out (zeroed before every launch)
  int chunk_tiles, chunk_min;         // largest and smallest chunk (tiles) of the guided schedule
  int rng_wide;                       // general fp64 kernel: wide native stream (53-bit uniforms, sequential cause rolls)
  uint32_t flags;                     // STRYKE_FLAG_*
  // implicit cell list (StrykeCells.groups): cell c of group g = cell0 + d * iterations + i  ->  day row, species, Philox id
  const StrykeCellGroup* groups; int n_groups; const int32_t* active_days; uint64_t id_days, id_add;
  int64_t cell_base;                  // index of this launch's first cell in the plan's cell list (launches over a range)
  RowsDev out;                        // compact tally rows
  int peak_units; const double* peak_params; float* peak_hours;      // peaking operations (StrykePeaking), hours [plan cells][units]
  unsigned long long* timeline;       // debug (STRYKE_B200_TIMELINE=1): per warp of the grid: start ns, end ns, chunks, tiles
  int pool_all;                       // pooled remainders: cells below this many fish (32 or 64) are pooled entirely
  const int64_t* blk_tile;            // pooled remainders: first tile of every aligned block of 32 cells [n_blocks + 1]
  const uint32_t* blk_S;              // pooled remainders (passage_fast.cuh): per block of 32 cells, shared tiles [0:8) | full tiles of
                                      // its owned cells [8:24) | mixed [31]
  int big_full;                       // pooled remainders: cells with more full tiles than this are "big" (pool_big); -1 = every cell
};

// Self-made bounds checks (compute-sanitizer is not available on the test pool): with -DSTRYKE_BOUNDS_CHECK every index
// into the staged tables is verified before use; a violation ORs bits 63+62 and a site code into the plan's error word
// (KParams::work_counter[-1]), which plan_status turns into an error.  Compiled out otherwise.
#ifdef STRYKE_BOUNDS_CHECK
#define STRYKE_BCHK(P, cond, site)                                                                              \
  do {                                                                                                          \
    if (!(cond)) atomicOr((P).work_counter - 1, (3ull << 62) | ((unsigned long long)(site) << 48));             \
  } while (0)
#else
#define STRYKE_BCHK(P, cond, site) do { } while (0)
#endif

// (day row, species row, Philox cell id) of cell c of the launch — from the explicit arrays or from the group table.
// With a group table the lookup is split: group_cursor finds the group of a BASE cell (a binary search and one 64-bit
// division, done once per thread block of count_kernel / per block of 32 cells of the passage kernel, with warp-uniform
// operands), cell_at then places the cells that follow it with 32-bit arithmetic.
struct CellInfo { int day, species; uint64_t id; };
struct GroupCursor {
  int64_t base, end;            // launch-relative cells [base, end) lie in this group from `base` on
  int iterations, day0, species, act_off, dd0, rem0;
  uint64_t id_base;
};
__device__ __forceinline__ GroupCursor group_cursor(const KParams& P, int64_t c) {
  GroupCursor G;
  const int64_t gc = c + P.cell_base;
  int lo = 0, hi = P.n_groups;                      // last group with cell0 <= gc
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.groups[mid].cell0 <= gc) lo = mid; else hi = mid; }
  const StrykeCellGroup g = P.groups[lo];
  const int64_t local = gc - g.cell0;
  G.base = c;
  G.end = c + ((int64_t)g.iterations * g.n_active_days - local);
  G.iterations = g.iterations; G.day0 = g.day0; G.species = g.species; G.act_off = g.active_off; G.id_base = g.id_base;
  const int64_t dd = local / g.iterations;
  G.dd0 = (int)dd; G.rem0 = (int)(local - dd * g.iterations);
  return G;
}
__device__ __forceinline__ CellInfo cell_at(const KParams& P, const GroupCursor& G, int64_t c) {
  CellInfo ci;
  if (P.groups == nullptr) {
    ci.day = P.cell_day[c]; ci.species = P.cell_species[c]; ci.id = P.cell_id[c];
    return ci;
  }
  GroupCursor H = G;
  if (c < G.base || c >= G.end || c - G.base > 0x3fffffffll) H = group_cursor(P, c);      // (rare: the cell lies in a later group)
  const unsigned int r = (unsigned int)H.rem0 + (unsigned int)(c - H.base);
  const unsigned int q = r / (unsigned int)H.iterations;
  const int it = (int)(r - q * (unsigned int)H.iterations);
  const int rel = P.active_days[H.act_off + H.dd0 + (int)q];
  ci.day = H.day0 + rel; ci.species = H.species;
  ci.id = (H.id_base + (uint64_t)it) * P.id_days + (uint64_t)rel + P.id_add;
  return ci;
}
__device__ __forceinline__ CellInfo cell_info(const KParams& P, int64_t c) {
  GroupCursor G{};
  if (P.groups != nullptr) G = group_cursor(P, c);
  return cell_at(P, G, c);
}

// Lane-held accumulators of one cell (lane p + 32 r owns pair p + 32 r, lanes 0..4 the Daily columns) -> global tallies.
// Dense output: red.global into daily[cell][5] / state_daily[cell][pair][2].  Compact output: into the cell's row.
// Called by all 32 lanes.
template <int PR>
__device__ __forceinline__ void tally_flush(const KParams& P, int64_t cell, int lane, int n_pairs, uint32_t (&acc_suc)[PR],
                                            uint32_t (&acc_cnt)[PR], uint32_t& acc_daily) {
  if (P.out.rows == nullptr) {
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const int p = r * 32 + lane;
      if (p < n_pairs && acc_cnt[r]) {
        uint32_t* dst = P.state_daily + ((size_t)cell * n_pairs + p) * 2;
        if (acc_suc[r]) atomicAdd(dst, acc_suc[r]);
        atomicAdd(dst + 1, acc_cnt[r]);
      }
    }
    if (lane < STRYKE_DAILY_W && acc_daily) atomicAdd(P.daily + (size_t)cell * STRYKE_DAILY_W + lane, acc_daily);
  } else {
    const int64_t b = cell >> 5;
    const int lc = (int)(cell & 31);
    const uint32_t m = P.out.blk_mask[b], wm = P.out.blk_wide[b], below = (1u << lc) - 1u;
    const unsigned int slot = P.out.blk_row[b] + __popc(m & below) + __popc(wm & below);
    const bool wide = (wm >> lc) & 1u;
    const uint32_t up = __shfl_down_sync(0xffffffffu, acc_daily, 1);
    if (slot + (wide ? 2u : 1u) <= P.out.cap) {
      uint32_t* row = P.out.rows + (size_t)slot * P.out.row_w;
      if (wide) {
#pragma unroll
        for (int r = 0; r < PR; ++r) {
          const int p = r * 32 + lane;
          if (p < n_pairs && acc_cnt[r]) {
            if (acc_suc[r]) atomicAdd(row + 6 + 2 * p, acc_suc[r]);
            atomicAdd(row + 7 + 2 * p, acc_cnt[r]);
          }
        }
        if (lane < STRYKE_DAILY_W && acc_daily) atomicAdd(row + 1 + lane, acc_daily);
      } else {
#pragma unroll
        for (int r = 0; r < PR; ++r) {
          const int p = r * 32 + lane;
          if (p < n_pairs && acc_cnt[r]) atomicAdd(row + 3 + p, acc_suc[r] | (acc_cnt[r] << 16));
        }
        const uint32_t w = acc_daily | (up << 16);
        if (lane == 0 && w) atomicAdd(row + 1, w);                       // num_entrained | num_survived << 16
        if (lane == 2 && w) atomicAdd(row + 2, w);                       // impingement | blade strike << 16
        if (lane == 4 && acc_daily) atomicAdd(row, acc_daily << 16);     // pop_size | barotrauma << 16
      }
    }
  }
#pragma unroll
  for (int r = 0; r < PR; ++r) { acc_suc[r] = 0; acc_cnt[r] = 0; }
  acc_daily = 0;
}

// Last index i in [lo, hi) with off[i] <= t (off is non-decreasing, off[lo] <= t), found by the whole warp: every step the 32
// lanes probe 32 evenly spaced entries at once, so 3 rounds of loads place a tile among 32 768 offsets (a lane-serial binary
// search pays 15 dependent memory latencies for the same).  All lanes call it with the same arguments and get the result.
__device__ __forceinline__ int64_t warp_search_last_le(const int64_t* __restrict__ off, int64_t lo, int64_t hi, int64_t t, int lane) {
  while (hi - lo > 1) {
    const int64_t span = hi - lo;
    const int64_t step = (span + 31) >> 5;                      // probes lo + (lane + 1) * step, clipped
    const int64_t idx = lo + (int64_t)(lane + 1) * step;
    const bool le = idx < hi && off[idx] <= t;
    const int k = __popc(__ballot_sync(0xffffffffu, le));        // probes 1..k hold offsets <= t (monotone)
    const int64_t nlo = lo + (int64_t)k * step;
    const int64_t nhi = lo + (int64_t)(k + 1) * step;
    lo = nlo;
    hi = nhi < hi ? nhi : hi;
  }
  return lo;
}

// Pooled remainders: a cell of n fish keeps pool_full(n) tiles of its own and sends pool_rem(n) fish to the shared tiles of
// its block.  Cells below pool_all fish (64 when the packed 6-/10-bit tally fields can hold 63 fish of one cell, else 32) are
// pooled entirely: entering and leaving a cell costs more than simulating its one or two tiles.
__host__ __device__ __forceinline__ int pool_rem(int n, int pool_all) { return n < pool_all ? n : (n & 31); }
__host__ __device__ __forceinline__ int pool_full(int n, int pool_all) { return n < pool_all ? 0 : (n >> 5); }
// Owned and big cells.  The tiles of a block of 32 cells are laid out as [shared tiles][full tiles of its OWNED cells][full
// tiles of its BIG cells]; the first two parts (the block's "atomic part") are simulated by ONE warp — the one whose chunk
// holds the block's first tile — so the compact rows of the owned cells are assembled in shared memory and written once,
// with plain stores, into rows nobody zeroed.  Big cells (more than big_full full tiles, or too many fish for the 16-bit
// row layout) are the ones whose tiles may be split between warps: their rows are zeroed by count_kernel and receive
// red.global adds, as every row did before.
__host__ __device__ __forceinline__ bool pool_big(int n, int pool_all, int big_full, int narrow_max) {
  return pool_full(n, pool_all) > big_full || n > narrow_max;
}
constexpr uint32_t POOL_S_MASK = 0xffu, POOL_OWNED_SHIFT = 8, POOL_OWNED_MASK = 0xffffu, POOL_MIXED = 0x80000000u;
constexpr int POOL_BIG_FULL_MAX = 255;      // 32 cells x 255 tiles fit POOL_OWNED_MASK
__host__ __device__ constexpr int pool_rwp(int n_pairs) { return (n_pairs + 3) | 1; }      // padded row (odd: no bank conflicts, lane <-> cell)

constexpr double P_ATM = 101325.0;    // scipy.constants.atm (stryke.py:1472)
constexpr double G_SI = 9.80665;      // scipy.constants.g   (stryke.py:1471)
constexpr double RHO = 998.2;         // stryke.py:1473
constexpr double FT2M = 0.3048;       // stryke.py:1497
constexpr double CM2FT = 0.0328084;   // stryke.py:3077
constexpr int SPW = 12;               // per-fish species words staged in shared memory
constexpr int USW = 6;                // per-unit static words staged: rack, intake_vel, fb_depth, tail_depth, c0, c1
constexpr int BLOCK = 256, WARPS = BLOCK / 32;

// single-MUFU approximations (no denormal / range fix-up code): arguments are known to be normal and in range
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_2pi(float turns) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(turns * 6.283185307f)); return y; }
__device__ __forceinline__ float exp_approx(float x) { return ex2_approx(x * 1.44269504f); }

// Uniforms of the native stream.  Every uniform is the CENTRE of its grid cell: (k + 1/2) / 2^b for the b random bits k,
// i.e. strictly inside (0, 1) and symmetric.  A floor grid k / 2^b would make P(u >= a) low by 2^-(b+1) on average — at
// 16 bits 7.6e-6, measurable on the mortality-cause counters at 1e9-1e10 fish — and put a point mass at 0; with the
// centred grid the error of P(u >= a) is zero-mean over thresholds a that vary from fish to fish (strike and barotrauma
// probabilities depend on the fish's length) and at most 2^-24 = 6e-8 for the fixed thresholds (a-priori survival,
// routing cdf), which the 23-bit uniforms serve.  Cost: none — the constant of the FADD changes from 1 to 1 - 2^-(b+1).
// All values are exact in fp32 AND fp64, so the fp64 kernels see the SAME uniforms as the fp32 ones: a native-RNG run
// differs between the precisions only where the arithmetic rounds across a threshold (tests compare them fish by fish).
template <typename Real> __device__ __forceinline__ Real u23(uint32_t x);    // bits [31:9]: dice, routing
template <typename Real> __device__ __forceinline__ Real u16hi(uint32_t w);  // bits [31:16]: r/R, Box-Muller radius
template <typename Real> __device__ __forceinline__ Real u16lo(uint32_t w);  // bits [15:0]: cause roll, Box-Muller angle
// 23 random mantissa bits under exponent 0: v in [1, 2); one ALU op + one FADD, no integer->float conversion (XU pipe)
__device__ __forceinline__ float mantissa_float(uint32_t x) { return __uint_as_float((x >> 9) | 0x3f800000u); }
template <> __device__ __forceinline__ float u23<float>(uint32_t x) { return mantissa_float(x) - 0.99999994039535522461f; }   // 1 - 2^-24
template <> __device__ __forceinline__ double u23<double>(uint32_t x) { return ((double)(x >> 9) + 0.5) * 1.1920928955078125e-07; }
// piece already moved to mantissa bits [22:7]: ((x & 0x007fff80) | 0x3f800000) as ONE three-input logic op
__device__ __forceinline__ float mantissa16(uint32_t x) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(0x007fff80u), "r"(0x3f800000u));
  return __uint_as_float(r);
}
#define STRYKE_C16 0.99999237060546875f        /* 1 - 2^-17 */
template <> __device__ __forceinline__ float u16hi<float>(uint32_t w) { return mantissa16(w >> 9) - STRYKE_C16; }
template <> __device__ __forceinline__ float u16lo<float>(uint32_t w) { return mantissa16(w << 7) - STRYKE_C16; }
template <> __device__ __forceinline__ double u16hi<double>(uint32_t w) { return ((double)(w >> 16) + 0.5) * (1.0 / 65536.0); }
template <> __device__ __forceinline__ double u16lo<double>(uint32_t w) { return ((double)(w & 0xffffu) + 0.5) * (1.0 / 65536.0); }

// Barotrauma levels take no word of their own for the depth draw: r/R then uses the top 12 bits of the r/R + cause word,
// and the depth uniform is built from 13 otherwise unused bits — bits [19:16] of that word and the 9 low bits of the dice
// word (the dice uniform reads bits [31:9]).  Centred like the others: (k + 1/2) / 2^12, (k + 1/2) / 2^13.
template <typename Real> __device__ __forceinline__ Real u12hi(uint32_t w);
template <typename Real> __device__ __forceinline__ Real u13_depth(uint32_t w_rc, uint32_t w_dice);
template <> __device__ __forceinline__ float u12hi<float>(uint32_t w) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(w >> 9), "r"(0x007ff800u), "r"(0x3f800000u));
  return __uint_as_float(r) - 0.9998779296875f;            // 1 - 2^-13
}
template <> __device__ __forceinline__ double u12hi<double>(uint32_t w) { return ((double)(w >> 20) + 0.5) * (1.0 / 4096.0); }
__device__ __forceinline__ uint32_t depth_bits(uint32_t w_rc, uint32_t w_dice) { return ((w_rc >> 7) & 0x1e00u) | (w_dice & 0x1ffu); }
template <> __device__ __forceinline__ float u13_depth<float>(uint32_t w_rc, uint32_t w_dice) {
  return __uint_as_float((depth_bits(w_rc, w_dice) << 10) | 0x3f800000u) - 0.99993896484375f;   // 1 - 2^-14
}
template <> __device__ __forceinline__ double u13_depth<double>(uint32_t w_rc, uint32_t w_dice) {
  return ((double)depth_bits(w_rc, w_dice) + 0.5) * (1.0 / 8192.0);
}

// Length draw: Box-Muller from ONE word — radius from the high half, angle from the low half (65 536 radii x 65 536
// angles; |z| <= sqrt(-2 ln 2^-17) = 4.85, i.e. 1.2e-6 of the normal mass is folded inwards.  Lengths are capped at 150 cm
// (stryke.py:3076) which is z = 3.8 for the reference's example species, and a strike probability is linear in the
// length, so the truncation moves no tally by more than 1e-8 relative; tests/test_gpu_high_power.py compares 1e10 fish
// against the fp64 kernel's wide stream (two 53-bit uniforms, |z| <= 8.6) and tests/test_gpu_native_rng.py the tail mass).
__device__ __forceinline__ float std_normal(float u1, float u2) {   // Box-Muller: MUFU lg2, sqrt, cos; u1, u2 in (0, 1)
  return sqrt_approx(-1.3862943611f * lg2_approx(u1)) * cos_2pi(u2);   // -2 ln u = -2 ln2 lg2 u
}

}  // namespace stryke
