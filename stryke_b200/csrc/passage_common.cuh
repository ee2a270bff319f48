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
  const uint32_t* blk_S;              // pooled remainders (passage_fast.cuh): per block of 32 cells, shared tiles [0:31) | mixed [31]
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

template <typename Real> __device__ __forceinline__ Real u_co(uint32_t x);   // [0,1)
template <typename Real> __device__ __forceinline__ Real u_oc(uint32_t x);   // (0,1]
// 23 random mantissa bits under exponent 0: v in [1, 2); one ALU op + one FADD, no integer->float conversion (XU pipe)
__device__ __forceinline__ float mantissa_float(uint32_t x) { return __uint_as_float((x >> 9) | 0x3f800000u); }
template <> __device__ __forceinline__ float u_co<float>(uint32_t x) { return mantissa_float(x) - 1.0f; }
template <> __device__ __forceinline__ float u_oc<float>(uint32_t x) { return 2.0f - mantissa_float(x); }
// fp64 kernels see the SAME uniforms as the fp32 ones (23 bits of the word): a native-RNG run then differs between the
// precisions only where the arithmetic rounds across a threshold (tests compare the two fish population by population)
template <> __device__ __forceinline__ double u_co<double>(uint32_t x) { return (double)(x >> 9) * 1.1920928955078125e-07; }
template <> __device__ __forceinline__ double u_oc<double>(uint32_t x) { return 1.0 - (double)(x >> 9) * 1.1920928955078125e-07; }

// 16-bit uniforms from the high / low half of a word (disjoint bits), multiples of 2^-16: exact in fp32 and fp64, so
// every kernel sees the same value.  hi_co, lo_co in [0, 1): k / 65536;  hi_oc in (0, 1]: 1 - k / 65536.
template <typename Real> __device__ __forceinline__ Real u16hi_co(uint32_t w);
template <typename Real> __device__ __forceinline__ Real u16hi_oc(uint32_t w);
template <typename Real> __device__ __forceinline__ Real u16lo_co(uint32_t w);
// piece already moved to mantissa bits [22:7]: ((x & 0x007fff80) | 0x3f800000) as ONE three-input logic op
__device__ __forceinline__ float mantissa16(uint32_t x) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(0x007fff80u), "r"(0x3f800000u));
  return __uint_as_float(r);
}
template <> __device__ __forceinline__ float u16hi_co<float>(uint32_t w) { return mantissa16(w >> 9) - 1.0f; }
template <> __device__ __forceinline__ float u16hi_oc<float>(uint32_t w) { return 2.0f - mantissa16(w >> 9); }
template <> __device__ __forceinline__ float u16lo_co<float>(uint32_t w) { return mantissa16(w << 7) - 1.0f; }
template <> __device__ __forceinline__ double u16hi_co<double>(uint32_t w) { return (double)(w >> 16) * (1.0 / 65536.0); }
template <> __device__ __forceinline__ double u16hi_oc<double>(uint32_t w) { return 1.0 - (double)(w >> 16) * (1.0 / 65536.0); }
template <> __device__ __forceinline__ double u16lo_co<double>(uint32_t w) { return (double)(w & 0xffffu) * (1.0 / 65536.0); }

// Barotrauma levels take no word of their own for the depth draw: r/R then uses the top 12 bits of the r/R + cause word,
// and the depth uniform is built from 13 otherwise unused bits — bits [19:16] of that word and the 9 low bits of the dice
// word (the dice uniform reads bits [31:9]).  Multiples of 2^-12 / 2^-13: identical in fp32 and fp64.
template <typename Real> __device__ __forceinline__ Real u12hi_co(uint32_t w);
template <typename Real> __device__ __forceinline__ Real u13_depth(uint32_t w_rc, uint32_t w_dice);
template <> __device__ __forceinline__ float u12hi_co<float>(uint32_t w) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(w >> 9), "r"(0x007ff800u), "r"(0x3f800000u));
  return __uint_as_float(r) - 1.0f;
}
template <> __device__ __forceinline__ double u12hi_co<double>(uint32_t w) { return (double)(w >> 20) * (1.0 / 4096.0); }
__device__ __forceinline__ uint32_t depth_bits(uint32_t w_rc, uint32_t w_dice) { return ((w_rc >> 7) & 0x1e00u) | (w_dice & 0x1ffu); }
template <> __device__ __forceinline__ float u13_depth<float>(uint32_t w_rc, uint32_t w_dice) {
  return __uint_as_float((depth_bits(w_rc, w_dice) << 10) | 0x3f800000u) - 1.0f;
}
template <> __device__ __forceinline__ double u13_depth<double>(uint32_t w_rc, uint32_t w_dice) {
  return (double)depth_bits(w_rc, w_dice) * (1.0 / 8192.0);
}

__device__ __forceinline__ float std_normal(float u1_oc, float u2_co) {   // Box-Muller: MUFU lg2, sqrt, cos
  return sqrt_approx(-1.3862943611f * lg2_approx(u1_oc)) * cos_2pi(u2_co);   // -2 ln u = -2 ln2 lg2 u
}

}  // namespace stryke
