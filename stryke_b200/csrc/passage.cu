// passage.cu — hand-written sm_100a kernels for stryke's individual-based Monte Carlo passage simulation and the
// C ABI declared in include/stryke_b200.h.
//
// One thread per simulated fish, one warp per tile of 32 fish of ONE (scenario, species, iteration, day) cell.
// A persistent grid gives every warp a contiguous range of tiles; the warp walks its cells, keeps the day's routing
// CDF + unit coefficients staged in a per-warp shared-memory slot, and accumulates the cell's tallies in lane-held
// registers (warp ballots / REDUX), flushing them with coalesced red.global adds when the cell changes.
//
// Six steps of the hot path (reference Stryke/stryke.py):
//   (1) count_kernel      presence roll + daily entrainment count  (:3025-3034, population_sim :2505-2618)
//   (2) passage_kernel    lognormal / normal length draw           (:3068-3090)
//   (3)                   routing by cumulative transition probability, searchsorted(cdf, u, 'right') (:2058)
//   (4)                   Franke blade strike (:846-1127), Pflugrath barotrauma (:1451-1517), impingement (:1416-1423)
//   (5)                   Bernoulli survival roll against the fp32-rounded rate (:1575, :3189, :3210)
//   (6)                   Daily + State_Daily tallies (:3272-3366)
//
// Template axes: Real = float (throughput mode, folded coefficients, MUFU intrinsics) or double (validation mode:
// the reference's operation order, no FMA contraction), INJECT = draws read from injected arrays (bit-exact replay
// of the reference's numpy uniforms, incl. np.vectorize's probe call on fish 0) or native Philox4x32-10.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include <fstream>
#include <iterator>
#include <map>
#include <mutex>
#include <sstream>

#include "../../include/stryke_b200.h"
#include "passage_common.cuh"
#include "passage_fast.cuh"

namespace stryke {

// fp64: the reference's operation order with every rounding kept (no FMA contraction)
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }

template <typename Real> struct UCW;
template <> struct UCW<float> { static constexpr int W = 4; };    // folded: fa, fb, fc|fl, has_flow
template <> struct UCW<double> { static constexpr int W = 8; };   // raw unit_coef row

// ---- blade strike + barotrauma, fp64 validation flavour (operation order of stryke.py:907-922, :1492-1510) --------
__device__ __forceinline__ double strike_surv64(int kind, const double* uc, double L, double rR) {
  const double lam = uc[0], N = uc[1], D = uc[2];
  double x;
  if (kind == STRYKE_KIND_KAPLAN) {
    const double a = atan(div_(uc[3], mul_(uc[4], rR)));
    x = add_(div_(cos(a), uc[5]), div_(sin(a), mul_(M_PI, rR)));
  } else {
    x = uc[3];
  }
  const double p = mul_(mul_(lam, div_(mul_(N, L), D)), x);
  return fmin(fmax(sub_(1.0, p), 0.0), 1.0);
}
__device__ __forceinline__ double baro_surv64(double depth_ft, double tail_ft, double b0, double b1) {
  const double p1 = add_(P_ATM, mul_(mul_(RHO, G_SI), mul_(depth_ft, FT2M)));
  const double p2 = add_(P_ATM, mul_(mul_(RHO, G_SI), mul_(tail_ft, FT2M)));
  const double e = exp(add_(b0, mul_(b1, log(div_(p1, p2)))));
  return sub_(1.0, div_(e, add_(1.0, e)));
}
// ---- fp32 throughput flavour: folded per-(day, unit) coefficients, MUFU intrinsics ---------------------------------
//   Kaplan: tan a = fa / rR  =>  cos a = rR * rsqrt(rR^2 + fa^2), sin a = fa * rsqrt(rR^2 + fa^2)
//           p = fc * L * rsqrt(rR^2 + fa^2) * (rR * fb + fa / (pi rR))
__device__ __forceinline__ float strike_surv32(int kind, const float* uc, float L, float rR) {
  float p;
  if (kind == STRYKE_KIND_KAPLAN) {
    const float fa = uc[0], fb = uc[1];
    const float inv = rsqrt_approx(fmaf(rR, rR, fa * fa));
    p = uc[2] * L * inv * fmaf(rR, fb, fa * 0.318309886f * rcp_approx(rR));
  } else {
    p = uc[2] * L;
  }
  p = uc[3] > 0.f ? p : INFINITY;   // no flow through the unit: p_strike = inf in the reference (Appendix E10)
  return __saturatef(1.f - p);
}
__device__ __forceinline__ float baro_surv32(float depth_ft, float c0, float c1, float b0l2e, float b1) {
  const float ratio = fmaf(c1, depth_ft, c0);                 // P1/P2
  const float e = ex2_approx(fmaf(b1, lg2_approx(ratio), b0l2e));   // exp(b0 + b1 ln ratio)
  return rcp_approx(1.f + e);                                       // 1 - e/(1+e)
}

// ------------------------------------------------------------------------------------------------------------------
// native RNG: word slot -> Philox block, warp-uniform refill points
// ------------------------------------------------------------------------------------------------------------------
struct FishRng {
  uint32_t c0, c1, c2, k0, k1;
  uint32_t w[4];
  int cur;
  __device__ __forceinline__ void init(uint32_t fish, uint64_t cell_id, uint32_t seed_lo, uint32_t seed_hi) {
    c0 = fish; c1 = (uint32_t)cell_id; c2 = (uint32_t)(cell_id >> 32); k0 = seed_lo; k1 = seed_hi; cur = -1;
  }
  __device__ __forceinline__ uint32_t word(int slot) {   // slot is warp-uniform
    const int b = slot >> 2;
    if (b != cur) { Philox::block(c0, c1, c2, (uint32_t)b, k0, k1, w); cur = b; }
    const int s = slot & 3;
    return (s & 2) ? ((s & 1) ? w[3] : w[2]) : ((s & 1) ? w[1] : w[0]);
  }
};
__device__ __forceinline__ double std_normal(double u1_oc, double u2_co) {
  return sqrt(-2.0 * log(u1_oc)) * cospi(2.0 * u2_co);
}

// ------------------------------------------------------------------------------------------------------------------
// (1) presence roll + daily entrainment count, one thread per cell
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {   // numpy-style 53-bit double in [0,1)
  return (double)((((uint64_t)hi >> 5) << 26) | ((uint64_t)lo >> 6)) * (1.0 / 9007199254740992.0);
}

template <bool INJECT>
__global__ void count_kernel(KParams P, const double* __restrict__ day_Q, int32_t* __restrict__ cell_n,
                             int max_daily_fish, unsigned long long* __restrict__ err) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cells) return;
  const double* sp = P.species + (int64_t)P.cell_species[c] * STRYKE_SPECIES_W;
  const int kind = (int)sp[12];
  const double shape = sp[13], loc = sp[14], scale = sp[15], max_rate = sp[16], occur = sp[17];
  uint32_t w[4], v[4];
  double presence;
  if (INJECT) {
    presence = P.inj.presence[c];
  } else {
    const uint64_t id = P.cell_id[c];
    Philox::block(0xFFFFFFFFu, (uint32_t)id, (uint32_t)(id >> 32), 0u, P.seed_lo, P.seed_hi, w);
    Philox::block(0xFFFFFFFFu, (uint32_t)id, (uint32_t)(id >> 32), 1u, P.seed_lo, P.seed_hi, v);
    presence = u53(w[0], w[1]);
  }
  int64_t n = 0;
  if (occur >= presence) {                                       // stryke.py:3027
    if (kind == STRYKE_ENT_FIXED) {
      n = (int64_t)sp[18];                                       // stryke.py:3029
    } else {
      const bool capped = isfinite(max_rate) && max_rate > 0.0;
      double rate;
      if (kind == STRYKE_ENT_PARETO) {                           // truncated_rvs, stryke.py:75-106, :2550-2559
        const double upper = capped ? max_rate * 10.0 : loc + fmax(1.0, scale * 1000.0);
        const double yu = (upper - loc) / scale;
        const double f_low = 0.0, f_high = yu >= 1.0 ? 1.0 - pow(yu, -shape) : 0.0;
        if (!isfinite(f_high) || f_high - f_low < 1e-12) {
          rate = upper;
        } else {
          const double q = INJECT ? P.inj.count_draw[c] : f_low + (f_high - f_low) * u53(w[2], w[3]);
          rate = loc + scale * pow(1.0 - q, -1.0 / shape);       // scipy pareto._ppf
        }
      } else if (kind == STRYKE_ENT_LOGNORMAL) {                 // scipy lognorm._rvs = exp(s*Z), stryke.py:2563
        double z;
        if (INJECT) z = P.inj.count_draw[c];
        else z = sqrt(-2.0 * log(1.0 - u53(w[2], w[3]))) * cospi(2.0 * u53(v[0], v[1]));
        rate = exp(shape * z) * scale + loc;
      } else {
        const double q = INJECT ? P.inj.count_draw[c] : u53(w[2], w[3]);
        if (kind == STRYKE_ENT_WEIBULL) {                        // weibull_min._ppf, stryke.py:2565
          rate = pow(-log1p(-q), 1.0 / shape) * scale + loc;
        } else {                                                 // genextreme._ppf, stryke.py:2561
          const double x = -log(-log(q));
          rate = (shape == 0.0 ? x : -expm1(-shape * x) / shape) * scale + loc;
        }
      }
      rate = fabs(rate);                                         // stryke.py:2567
      if (capped && rate > max_rate * 10.0) rate = max_rate * 10.0;   // stryke.py:2578-2582
      const double Mft3 = (60.0 * 60.0 * 24.0 * day_Q[P.cell_day[c]]) / 1000000.0;   // stryke.py:2585
      const double nf = rint(Mft3 * rate);                       // np.round: half to even
      if (!isfinite(nf) || nf < 0.0) { atomicMax(err, 0x8000000000000000ull | (unsigned long long)c); n = 0; }
      else n = (int64_t)nf;
    }
    if (n == 0) n = 1;                                           // stryke.py:3032-3033
    if (max_daily_fish > 0 && kind != STRYKE_ENT_FIXED && n > max_daily_fish) {    // stryke.py:2610-2616
      atomicMax(err, ((unsigned long long)n << 32) | (unsigned long long)(c & 0xFFFFFFFFll));
      n = 0;
    }
    if (n > 0x7FFFFFFFll) { atomicMax(err, 0x8000000000000000ull | (unsigned long long)c); n = 0; }
  }
  cell_n[c] = (int32_t)n;
}

// ------------------------------------------------------------------------------------------------------------------
// exclusive scans of tiles-per-cell and fish-per-cell (three small kernels; int64 offsets)
// ------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 1024, SCAN_PER = 4, SCAN_CH = SCAN_T * SCAN_PER;

__device__ __forceinline__ void block_scan2(int64_t& a, int64_t& b, int64_t* sh) {   // inclusive, 1024 threads
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t ta = __shfl_up_sync(0xffffffffu, a, o), tb = __shfl_up_sync(0xffffffffu, b, o);
    if (lane >= o) { a += ta; b += tb; }
  }
  if (lane == 31) { sh[wid] = a; sh[32 + wid] = b; }
  __syncthreads();
  if (wid == 0) {
    int64_t sa = sh[lane], sb = sh[32 + lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t ta = __shfl_up_sync(0xffffffffu, sa, o), tb = __shfl_up_sync(0xffffffffu, sb, o);
      if (lane >= o) { sa += ta; sb += tb; }
    }
    sh[lane] = sa; sh[32 + lane] = sb;
  }
  __syncthreads();
  if (wid > 0) { a += sh[wid - 1]; b += sh[32 + wid - 1]; }
  __syncthreads();
}

__global__ void __launch_bounds__(SCAN_T) scan_partial(const int32_t* __restrict__ cell_n, int64_t n_cells,
                                                       int64_t* __restrict__ part) {
  __shared__ int64_t sh[64];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + threadIdx.x * SCAN_PER;
  int64_t t = 0, f = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER; ++i)
    if (base + i < n_cells) { const int32_t n = cell_n[base + i]; f += n; t += (n + 31) >> 5; }
  block_scan2(t, f, sh);
  if (threadIdx.x == SCAN_T - 1) { part[2 * blockIdx.x] = t; part[2 * blockIdx.x + 1] = f; }
}
__global__ void __launch_bounds__(SCAN_T) scan_blocks(int64_t* __restrict__ part, int nb) {   // exclusive, in place
  __shared__ int64_t sh[64];
  const int per = (nb + SCAN_T - 1) / SCAN_T;
  const int lo = threadIdx.x * per, hi = min(nb, lo + per);
  int64_t t = 0, f = 0;
  for (int i = lo; i < hi; ++i) { t += part[2 * i]; f += part[2 * i + 1]; }
  int64_t it = t, jf = f;
  block_scan2(it, jf, sh);
  int64_t et = it - t, ef = jf - f;
  for (int i = lo; i < hi; ++i) {
    const int64_t a = part[2 * i], b = part[2 * i + 1];
    part[2 * i] = et; part[2 * i + 1] = ef; et += a; ef += b;
  }
}
__global__ void __launch_bounds__(SCAN_T) scan_final(const int32_t* __restrict__ cell_n, int64_t n_cells,
                                                     const int64_t* __restrict__ part, int64_t* __restrict__ tile_off,
                                                     int64_t* __restrict__ fish_off) {
  __shared__ int64_t sh[64];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + threadIdx.x * SCAN_PER;
  int32_t nn[SCAN_PER];
  int64_t t = 0, f = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER; ++i) {
    nn[i] = base + i < n_cells ? cell_n[base + i] : 0;
    f += nn[i]; t += (nn[i] + 31) >> 5;
  }
  int64_t it = t, jf = f;
  block_scan2(it, jf, sh);
  int64_t et = it - t + part[2 * blockIdx.x], ef = jf - f + part[2 * blockIdx.x + 1];
#pragma unroll
  for (int i = 0; i < SCAN_PER; ++i) {
    if (base + i < n_cells) { tile_off[base + i] = et; fish_off[base + i] = ef; }
    et += (nn[i] + 31) >> 5; ef += nn[i];
    if (base + i == n_cells - 1) { tile_off[n_cells] = et; fish_off[n_cells] = ef; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// (2)-(6) the passage kernel
// ------------------------------------------------------------------------------------------------------------------

struct SmemLayout {   // byte offsets into dynamic shared memory (all 16-byte aligned)
  int nodes, edge_dst, moves, pair_node, ustatic, species, warp0, warp_stride, cdf, stay, ucoef, total;
};

template <typename Real>
__host__ __device__ inline SmemLayout smem_layout(int n_nodes, int n_units, int n_moves, int n_edges, int n_pairs,
                                                  int n_species_staged) {
  auto al = [](int x) { return (x + 15) & ~15; };
  SmemLayout L;
  int o = 0;
  L.nodes = o; o = al(o + n_nodes * (int)sizeof(NodeRec));
  L.edge_dst = o; o = al(o + n_edges);
  L.moves = o; o = al(o + n_moves * (int)sizeof(MoveRec));
  L.pair_node = o; o = al(o + n_pairs);
  L.ustatic = o; o = al(o + n_units * USW * (int)sizeof(Real));
  L.species = o; o = al(o + n_species_staged * SPW * (int)sizeof(Real));
  L.warp0 = o;
  int w = 0;
  L.cdf = w; w = al(w + n_edges * (int)sizeof(Real));
  L.ucoef = w; w = al(w + n_units * UCW<Real>::W * (int)sizeof(Real));
  L.stay = w; w = al(w + n_nodes);
  L.warp_stride = w;
  L.total = o + w * WARPS;
  return L;
}

template <typename Real, bool INJECT, int PR>
__global__ void __launch_bounds__(BLOCK) passage_kernel(const KParams P, const int n_species_staged) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout SL = smem_layout<Real>(P.n_nodes, P.n_units, P.n_moves, P.n_edges, P.n_pairs, n_species_staged);
  NodeRec* s_nodes = reinterpret_cast<NodeRec*>(smem + SL.nodes);
  uint8_t* s_edge_dst = smem + SL.edge_dst;
  MoveRec* s_moves = reinterpret_cast<MoveRec*>(smem + SL.moves);
  uint8_t* s_pair = smem + SL.pair_node;
  Real* s_ustatic = reinterpret_cast<Real*>(smem + SL.ustatic);
  Real* s_species = reinterpret_cast<Real*>(smem + SL.species);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* wbase = smem + SL.warp0 + wid * SL.warp_stride;
  Real* s_cdf = reinterpret_cast<Real*>(wbase + SL.cdf);
  Real* s_ucoef = reinterpret_cast<Real*>(wbase + SL.ucoef);
  uint8_t* s_stay = wbase + SL.stay;

  // ---- block-level staging of the static facility tables ---------------------------------------------------------
  for (int i = threadIdx.x; i < P.n_nodes; i += BLOCK) s_nodes[i] = P.nodes[i];
  for (int i = threadIdx.x; i < P.n_edges; i += BLOCK) s_edge_dst[i] = P.edge_dst[i];
  for (int i = threadIdx.x; i < P.n_moves; i += BLOCK) s_moves[i] = P.moves[i];
  for (int i = threadIdx.x; i < P.n_pairs; i += BLOCK) s_pair[i] = P.pair_node[i];
  for (int u = threadIdx.x; u < P.n_units; u += BLOCK) {
    const double* us = P.unit_static + u * STRYKE_UNIT_STATIC_W;
    const double p2 = P_ATM + RHO * G_SI * (us[3] * FT2M);
    s_ustatic[u * USW + 0] = (Real)us[0]; s_ustatic[u * USW + 1] = (Real)us[1];
    s_ustatic[u * USW + 2] = (Real)us[2]; s_ustatic[u * USW + 3] = (Real)us[3];
    s_ustatic[u * USW + 4] = (Real)(P_ATM / p2); s_ustatic[u * USW + 5] = (Real)(RHO * G_SI * FT2M / p2);
  }
  for (int i = threadIdx.x; i < n_species_staged * SPW; i += BLOCK)
    s_species[i] = (Real)P.species[(i / SPW) * STRYKE_SPECIES_W + (i % SPW)];
  __syncthreads();

  // ---- this warp's contiguous tile range -------------------------------------------------------------------------
  const int64_t T = P.tile_off[P.n_cells];
  const int64_t G = (int64_t)gridDim.x * WARPS, gw = (int64_t)blockIdx.x * WARPS + wid;
  const int64_t t0 = (T / G) * gw + min(gw, T % G), t1 = t0 + T / G + (gw < T % G ? 1 : 0);
  if (t0 >= t1) return;
  int64_t c;
  {  // last cell with tile_off[c] <= t0
    int64_t lo = 0, hi = P.n_cells;
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (P.tile_off[mid] <= t0) lo = mid; else hi = mid; }
    c = lo;
  }
  int64_t c_end = P.tile_off[c + 1], c_begin = P.tile_off[c];
  int cur_day = -1;
  int64_t cur_cell = -1;

  // per-cell registers
  Real sp_ls = 0, sp_ll = 0, sp_lsc = 0, sp_lm = 0, sp_lsd = 0, sp_wr = 0, sp_uc = 0, sp_d1 = 0, sp_d2 = 0, sp_b0 = 0, sp_b1 = 0;
  int sp_mode = 0, n_c = 0;
  uint64_t cid = 0;
  int64_t fo = 0;
  uint32_t acc_suc[PR], acc_cnt[PR], acc_daily = 0;
#pragma unroll
  for (int r = 0; r < PR; ++r) { acc_suc[r] = 0; acc_cnt[r] = 0; }

  auto flush = [&](int64_t cell) {
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const int p = r * 32 + lane;
      if (p < P.n_pairs && acc_cnt[r]) {
        uint32_t* dst = P.state_daily + ((size_t)cell * P.n_pairs + p) * 2;
        if (acc_suc[r]) atomicAdd(dst, acc_suc[r]);
        atomicAdd(dst + 1, acc_cnt[r]);
      }
      acc_suc[r] = 0; acc_cnt[r] = 0;
    }
    if (lane < STRYKE_DAILY_W && acc_daily) atomicAdd(P.daily + (size_t)cell * STRYKE_DAILY_W + lane, acc_daily);
    acc_daily = 0;
  };

  for (int64_t t = t0; t < t1; ++t) {
    while (t >= c_end) { ++c; c_begin = c_end; c_end = P.tile_off[c + 1]; }
    if (c != cur_cell) {
      if (cur_cell >= 0) flush(cur_cell);
      cur_cell = c;
      n_c = P.cell_n[c];
      cid = P.cell_id[c];
      if (INJECT || P.trace.state || P.trace.length_ft64) fo = P.fish_off[c];
      const int s = P.cell_species[c];
      if (s < n_species_staged) {
        const Real* q = s_species + s * SPW;
        sp_ls = q[0]; sp_ll = q[1]; sp_lsc = q[2]; sp_lm = q[3]; sp_lsd = q[4]; sp_mode = (int)q[5];
        sp_wr = q[6]; sp_uc = q[7]; sp_d1 = q[8]; sp_d2 = q[9]; sp_b0 = q[10]; sp_b1 = q[11];
      } else {
        const double* q = P.species + (int64_t)s * STRYKE_SPECIES_W;
        sp_ls = (Real)q[0]; sp_ll = (Real)q[1]; sp_lsc = (Real)q[2]; sp_lm = (Real)q[3]; sp_lsd = (Real)q[4];
        sp_mode = (int)q[5]; sp_wr = (Real)q[6]; sp_uc = (Real)q[7]; sp_d1 = (Real)q[8]; sp_d2 = (Real)q[9];
        sp_b0 = (Real)q[10]; sp_b1 = (Real)q[11];
      }
      const int d = P.cell_day[c];
      if (d != cur_day) {   // stage the day's routing CDF, stay flags and unit coefficients in this warp's slot
        cur_day = d;
        __syncwarp();
        const double* cdf = P.edge_cdf + (size_t)d * P.n_edges;
        for (int i = lane; i < P.n_edges; i += 32) s_cdf[i] = (Real)cdf[i];
        const uint8_t* st = P.node_stay + (size_t)d * P.n_nodes;
        for (int i = lane; i < P.n_nodes; i += 32) s_stay[i] = st[i];
        const double* uc = P.unit_coef + (size_t)d * P.n_units * STRYKE_UNIT_COEF_W;
        if (sizeof(Real) == 8) {
          for (int i = lane; i < P.n_units * STRYKE_UNIT_COEF_W; i += 32) s_ucoef[i] = (Real)uc[i];
        } else {
          for (int u = lane; u < P.n_units; u += 32) {
            const double* r = uc + u * STRYKE_UNIT_COEF_W;
            const double lnd = r[0] * r[1] / r[2];
            // a Kaplan row keeps pi*ada*Ewd, 2*Qwd, 8*Qwd; other runners keep X in slot 3 (slot 4 is 0)
            const bool kap = r[4] != 0.0 || r[5] != 0.0;
            s_ucoef[u * 4 + 0] = kap ? (Real)(r[3] / r[4]) : (Real)0;
            s_ucoef[u * 4 + 1] = kap ? (Real)(1.0 / r[5]) : (Real)0;
            s_ucoef[u * 4 + 2] = kap ? (Real)lnd : (Real)(lnd * r[3]);
            s_ucoef[u * 4 + 3] = (Real)r[6];
          }
        }
        __syncwarp();
      }
    }
    const int64_t tile_in_cell = t - c_begin;
    const int fish = (int)(tile_in_cell * 32 + lane);
    const bool active = fish < n_c;
    const int64_t gi = fo + fish;                 // global fish index (injected / trace arrays)
    const int64_t NF = P.n_fish_total;

    FishRng rng;
    if (!INJECT) rng.init((uint32_t)fish, cid, P.seed_lo, P.seed_hi);

    // ---- (2) length draw ---------------------------------------------------------------------------------------
    Real z;
    if (INJECT) z = active ? (Real)P.inj.length_z[gi] : (Real)0;
    else { const uint32_t a = rng.word(0); z = std_normal(u16hi_oc<Real>(a), u16lo_co<Real>(a)); }
    Real L;
    if (sizeof(Real) == 8) {
      if (sp_mode == 0) {   // |Z-lognormal| cm, cap 150, -> ft  (scipy: exp(s*Z)*scale + loc; stryke.py:3070-3077)
        double v = fabs(add_(mul_(exp(mul_((double)sp_ls, (double)z)), (double)sp_lsc), (double)sp_ll));
        v = v > 150.0 ? 150.0 : v;
        L = (Real)mul_(v, CM2FT);
      } else {              // |N(mean, sd)| inches -> ft  (stryke.py:3090)
        L = (Real)div_(fabs(add_((double)sp_lm, mul_((double)sp_lsd, (double)z))), 12.0);
      }
    } else {
      if (sp_mode == 0) L = (Real)(fminf(fabsf(fmaf(exp_approx((float)sp_ls * (float)z), (float)sp_lsc, (float)sp_ll)), 150.f) * 0.0328084f);
      else L = (Real)(fabsf(fmaf((float)sp_lsd, (float)z, (float)sp_lm)) * (1.f / 12.f));
    }
    if (P.trace.length_ft64 && active && gi < NF) { P.trace.length_ft64[gi] = (double)L; if (P.trace.length_ft) P.trace.length_ft[gi] = (float)L; }

    int loc = P.start_node;
    bool alive = active;
    int best_rank = 255;
    bool ent_surv = false;
    uint32_t cz = 0;   // cause counters packed: imp | blade << 8 | baro << 16 (one fish visits <= 64 moves)
    const Real blocked_w = L * sp_wr;    // stryke.py:1418

    for (int k = 0; k < P.n_moves; ++k) {
      const MoveRec mv = s_moves[k];
      const NodeRec nd = s_nodes[loc];
      const bool prev_alive = alive;
      float rate32 = 0.f;
      // word slots of this move (MoveRec): r/R + cause share one word, depth, dice, route one each
      const int s_dep = mv.slot_depth, s_dice = mv.slot_dice;
      const int s_rt = mv.slot_route;
      uint32_t w_rc = 0, w_dep = 0;
      if (!INJECT) {   // warp-uniform draws for this move (refill points are identical for every lane)
        if (mv.need & (NEED_RR | NEED_CAUSE)) w_rc = rng.word(mv.slot_rc);
        if (mv.need & NEED_DEPTH) w_dep = rng.word(s_dep);
      }
      Real strike_t = (Real)NAN, baro_t = (Real)NAN;
      if (alive) {
        if (nd.kind == STRYKE_KIND_APRIORI) {
          rate32 = nd.apriori32;
        } else {
          // ---- (4) impingement, blade strike, barotrauma (node_surv_rate, stryke.py:1410-1572) --------------------
          const Real* us = s_ustatic + nd.unit * USW;
          const Real* uc = s_ucoef + nd.unit * UCW<Real>::W;
          const bool blocked = blocked_w > us[0];
          const Real imp = (blocked && sp_uc < us[1]) ? (Real)0 : (Real)1;
          Real strike = (Real)1, baro = (Real)1;
          if (!blocked) {
            Real rR = (Real)0.75;
            if (nd.kind == STRYKE_KIND_KAPLAN) {
              if (INJECT) rR = (Real)add_(0.3, mul_(sub_(1.0, 0.3), P.inj.rR[(int64_t)k * NF + gi]));
              else if (sizeof(Real) == 8) rR = (Real)(0.3 + 0.7 * (double)u16hi_co<Real>(w_rc));
              else rR = (Real)fmaf(0.7f, (float)u16hi_co<Real>(w_rc), 0.3f);
            }
            if (sizeof(Real) == 8) strike = (Real)strike_surv64(nd.kind, (const double*)uc, (double)L, (double)rR);
            else strike = (Real)strike_surv32(nd.kind, (const float*)uc, (float)L, (float)rR);
            strike_t = strike;
            if (P.barotrauma) {
              if (sizeof(Real) == 8) {
                const double fb = (double)us[2], lo = mul_(fb, (double)sp_d1), hi = mul_(fb, (double)sp_d2);
                const double ud = INJECT ? P.inj.depth[(int64_t)k * NF + gi] : (double)u_co<Real>(w_dep);
                baro = (Real)baro_surv64(add_(lo, mul_(sub_(hi, lo), ud)), (double)us[3], (double)sp_b0, (double)sp_b1);
              } else {
                const float ud = INJECT ? (float)P.inj.depth[(int64_t)k * NF + gi] : (float)u_co<Real>(w_dep);
                const float depth = (float)us[2] * fmaf((float)sp_d2 - (float)sp_d1, ud, (float)sp_d1);
                baro = (Real)baro_surv32(depth, (float)us[4], (float)us[5], (float)sp_b0 * 1.44269504f, (float)sp_b1);
              }
              baro_t = baro;
            }
          }
          const Real prob = (sizeof(Real) == 8) ? (Real)mul_(mul_((double)imp, (double)strike), (double)baro)
                                                : imp * strike * baro;
          rate32 = (float)prob;                         // np.float32(prob), stryke.py:1575
          // mortality-cause rolls, independent of the survival dice (stryke.py:1545-1560, Appendix E4)
          if (INJECT) {
            const double* cu = P.inj.cause + ((int64_t)k * 4) * NF + gi;
            if (cu[0] > (double)imp) cz += 1u;
            else if (cu[NF] > (double)strike) cz += 1u << 8;
            else if (cu[2 * NF] > (double)baro) cz += 1u << 16;
          } else {
            // one uniform splits [0,1) into the same three probabilities the reference's sequential rolls have:
            // P(imp) = 1-imp, P(blade) = imp (1-strike), P(baro) = imp strike (1-baro)
            const Real u = u16lo_co<Real>(w_rc);
            const Real a1 = imp, a2 = imp * strike, a3 = a2 * baro;
            if (u >= a1) cz += 1u;
            else if (u >= a2) cz += 1u << 8;
            else if (u >= a3) cz += 1u << 16;
          }
        }
      }
      // np.vectorize's probe call: fish 0 is evaluated once more with its own draws; only the cause counters keep
      // the side effect (stryke.py:3197)
      if (INJECT && P.inj.ghost_cause && fish == 0 && active && prev_alive && nd.kind != STRYKE_KIND_APRIORI) {
        const Real* us = s_ustatic + nd.unit * USW;
        const Real* uc = s_ucoef + nd.unit * UCW<Real>::W;
        const bool blocked = blocked_w > us[0];
        const double imp = (blocked && sp_uc < us[1]) ? 0.0 : 1.0;
        double strike = 1.0, baro = 1.0;
        const int64_t gk = (int64_t)k * P.n_cells + c;
        if (!blocked) {
          double rR = 0.75;
          if (nd.kind == STRYKE_KIND_KAPLAN) rR = add_(0.3, mul_(sub_(1.0, 0.3), P.inj.ghost_rR[gk]));
          if (sizeof(Real) == 8) strike = strike_surv64(nd.kind, (const double*)uc, (double)L, rR);
          else strike = (double)strike_surv32(nd.kind, (const float*)uc, (float)L, (float)rR);
          if (P.barotrauma) {
            const double fb = (double)us[2], lo = mul_(fb, (double)sp_d1), hi = mul_(fb, (double)sp_d2);
            baro = baro_surv64(add_(lo, mul_(sub_(hi, lo), P.inj.ghost_depth[gk])), (double)us[3], (double)sp_b0, (double)sp_b1);
          }
        }
        const double* cu = P.inj.ghost_cause + ((int64_t)k * 4) * P.n_cells + c;
        if (cu[0] > imp) cz += 1u;
        else if (cu[P.n_cells] > strike) cz += 1u << 8;
        else if (cu[2 * P.n_cells] > baro) cz += 1u << 16;
      }

      // ---- (5) Bernoulli survival roll: dice <= float64(float32(prob)) (stryke.py:3210) ----------------------------
      bool surv;
      if (INJECT) {
        surv = active && (P.inj.dice[(int64_t)k * NF + gi] <= (double)rate32);
      } else if (mv.need & NEED_DICE) {
        const uint32_t wd = rng.word(s_dice);
        surv = alive && (u_oc<Real>(wd) <= (Real)rate32);
      } else {
        surv = alive && rate32 >= 1.0f;    // every node of this level has survival 1.0: no draw needed
      }
      if (P.trace.state && active && gi < NF) {     // gi >= NF: caller-supplied capacity too small (plan_status reports it)
        const int64_t o = (int64_t)k * NF + gi;
        P.trace.state[o] = (uint8_t)loc; P.trace.survival[o] = surv ? 1 : 0; P.trace.rate[o] = rate32;
        if (P.trace.strike) { P.trace.strike[o] = (double)strike_t; P.trace.baro[o] = (double)baro_t; }
      }

      // ---- (6) State_Daily: among fish alive after move k-1, per node: successes and count (stryke.py:3302-3351) --
      for (int p = mv.level_begin; p < mv.level_end; ++p) {
        const int node = s_pair[p];
        const bool here = loc == node;
        const uint32_t bc = __ballot_sync(0xffffffffu, prev_alive && here);
        const uint32_t bs = __ballot_sync(0xffffffffu, prev_alive && surv && here);
#pragma unroll
        for (int r = 0; r < PR; ++r)
          if ((p >> 5) == r && (p & 31) == lane) { acc_cnt[r] += __popc(bc); acc_suc[r] += __popc(bs); }
      }
      // entrainment bookkeeping: first 'U' state in lexicographic column order (stryke.py:3272-3292)
      if (nd.hasU && active && (int)mv.lex < best_rank) { best_rank = mv.lex; ent_surv = surv; }

      // ---- (3) movement: searchsorted(cdf, u, side='right') over the node's successors (stryke.py:2058) -----------
      if (k < P.n_moves - 1) {
        Real u = (Real)0;
        if (!INJECT && (mv.need & NEED_ROUTE)) u = u_co<Real>(rng.word(s_rt));
        if (surv && nd.deg > 0 && !s_stay[loc]) {
          if (INJECT) u = (Real)P.inj.route[(int64_t)k * NF + gi];
          int j = 0;
          if (INJECT && sizeof(Real) == 4) {
            const double* cdf64 = P.edge_cdf + (size_t)cur_day * P.n_edges + nd.edge_off;
            const double ud = P.inj.route[(int64_t)k * NF + gi];
            for (int e = 0; e < nd.deg - 1; ++e) j += (cdf64[e] <= ud) ? 1 : 0;
          } else {
            const Real* cdf = s_cdf + nd.edge_off;
            for (int e = 0; e < nd.deg - 1; ++e) j += (cdf[e] <= u) ? 1 : 0;
          }
          loc = s_edge_dst[nd.edge_off + j];
        }
      }
      alive = surv;
    }

    // ---- Daily tallies of this tile (stryke.py:3281-3298, :3353-3366) ----------------------------------------------
    const bool ent = active && best_rank < 255;
    const uint32_t b_ent = __ballot_sync(0xffffffffu, ent);
    const uint32_t b_sur = __ballot_sync(0xffffffffu, ent && ent_surv);
    const uint32_t any_cz = __ballot_sync(0xffffffffu, cz != 0);
    uint32_t r_imp = 0, r_bld = 0, r_bar = 0;
    if (any_cz) {
      r_imp = __reduce_add_sync(0xffffffffu, cz & 0xffu);
      r_bld = __reduce_add_sync(0xffffffffu, (cz >> 8) & 0xffu);
      r_bar = __reduce_add_sync(0xffffffffu, (cz >> 16) & 0xffu);
    }
    acc_daily += lane == 0 ? __popc(b_ent) : lane == 1 ? __popc(b_sur) : lane == 2 ? r_imp : lane == 3 ? r_bld : lane == 4 ? r_bar : 0u;
  }
  if (cur_cell >= 0) flush(cur_cell);
}

// ------------------------------------------------------------------------------------------------------------------
// small utility kernels: per-fish probability functions, Philox KAT, issue-rate calibration
// ------------------------------------------------------------------------------------------------------------------
template <typename Real>
__global__ void strike_kernel(int kind, const double* __restrict__ uc8, int64_t n, const double* __restrict__ L,
                              const double* __restrict__ rR, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = (kind == STRYKE_KIND_KAPLAN) ? rR[i] : 0.75;
  if (sizeof(Real) == 8) {
    out[i] = strike_surv64(kind, uc8, L[i], r);
  } else {
    const double lnd = uc8[0] * uc8[1] / uc8[2];
    const bool kap = kind == STRYKE_KIND_KAPLAN;
    float f[4] = {kap ? (float)(uc8[3] / uc8[4]) : 0.f, kap ? (float)(1.0 / uc8[5]) : 0.f,
                  kap ? (float)lnd : (float)(lnd * uc8[3]), (float)uc8[6]};
    out[i] = (double)strike_surv32(kind, f, (float)L[i], (float)r);
  }
}
template <typename Real>
__global__ void baro_kernel(int64_t n, const double* __restrict__ depth, double tail, double b0, double b1,
                            double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (sizeof(Real) == 8) {
    out[i] = baro_surv64(depth[i], tail, b0, b1);
  } else {
    const double p2 = P_ATM + RHO * G_SI * (tail * FT2M);
    out[i] = (double)baro_surv32((float)depth[i], (float)(P_ATM / p2), (float)(RHO * G_SI * FT2M / p2),
                                 (float)b0 * 1.44269504f, (float)b1);
  }
}
__global__ void philox_kernel(int64_t n, const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key,
                              uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[4];
  Philox::block(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], w);
  out[4 * i] = w[0]; out[4 * i + 1] = w[1]; out[4 * i + 2] = w[2]; out[4 * i + 3] = w[3];
}
// Issue-rate calibration: 16 independent FFMA chains per thread, 128 FFMA per loop iteration (PTX, so the count is
// exact).  FFMA issues at one warp instruction per scheduler per cycle (both FP32 datapaths of an SM sub-partition),
// so the achieved FFMA lane-ops/s is the measured instruction-issue ceiling the passage kernel is compared with.
constexpr int CAL_ITERS = 1024, CAL_OPS = 128;
__global__ void __launch_bounds__(256) calibrate_kernel(uint32_t* out, uint32_t seed) {
  float x[16];
  const float m = 1.0f + 1e-7f * (float)(seed & 7), c = 1e-9f * (float)(threadIdx.x & 3);
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i);
#pragma unroll 1
  for (int it = 0; it < CAL_ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(m), "f"(c));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __float_as_uint(x[i]);
  if (s == 0x12345678u) out[0] = s;   // never true in practice; keeps the chains live
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static inline uint32_t __float_as_uint_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#define CK(expr)                                                                                           \
  do {                                                                                                     \
    cudaError_t e_ = (expr);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return fail(STRYKE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                      \
  } while (0)

// Device buffers come from the stream-ordered pool of the legacy default stream (cudaMallocAsync): after the first
// call the pool keeps its memory (release threshold = max), so the one-shot host-buffer entry point does not pay
// cudaMalloc/cudaFree for its ~30 tables on every call.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFreeAsync(p, 0); }
  cudaError_t alloc(size_t b) { bytes = b; return cudaMallocAsync(&p, b ? b : 16, 0); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
static int device_sm_count(int device, int* n_sm) {
  static int cache[64] = {0};
  if (device >= 0 && device < 64 && cache[device] > 0) { *n_sm = cache[device]; return 0; }
  int v = 0;
  cudaError_t e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return 1;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (device >= 0 && device < 64) cache[device] = v;
  *n_sm = v;
  return 0;
}
template <typename T>
static cudaError_t upload(DevBuf& d, const T* h, size_t n, cudaStream_t st = 0) {
  cudaError_t e = d.alloc(n * sizeof(T));
  if (e != cudaSuccess) return e;
  if (n == 0 || h == nullptr) return cudaSuccess;
  return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

}  // namespace stryke

using namespace stryke;

struct StrykePlan {
  int device = 0, precision = 32, n_sm = 0, blocks_per_sm = 0, max_daily_fish = 0;
  bool inject = false, fixed_counts = false;
  KParams K{};
  int n_species_staged = 0, pr = 1;
  size_t smem = 0;
  int64_t n_cells = 0;
  int nb_scan = 0;
  DevBuf nodes, edge_dst, moves, pair_node, ustatic, edge_cdf, node_stay, unit_coef, day_Q, species, cell_day,
      cell_species, cell_id, tile_off, fish_off, part, err;
  DevBuf i_fish_off, i_presence, i_count, i_z, i_dice, i_rR, i_depth, i_cause, i_route, i_grR, i_gdepth, i_gcause;
  DevBuf t_len32, t_len64, t_state, t_rate, t_surv, t_strike, t_baro;
  DevBuf fast_nodes;
  FastTables fast{};
  size_t fast_smem_bytes = 0;
  int fast_stride = 1;
  bool use_fast = false;
  cudaKernel_t jit_kernel = nullptr;
  std::string kernel_name = "passage_kernel (general)";
  int64_t trace_fish = 0;
  bool has_trace = false, trace_probs = false;
};

namespace jit {
// Layout of the packed per-lane tallies of the specialised kernel (LevelTally / FieldRec in passage_fast.cuh).
struct TallyPlan {
  int NW = 0, HOLD = 1, W_ES = 0, B_ES = 0, W_CAUSE = 0;
  std::vector<LevelTally> lv;
  std::vector<FieldRec> fields;
  std::vector<int> node_local;       // local index of every node inside its (non-trivial, multi-node) level, -1 = none
};
static TallyPlan plan_tally(const std::vector<MoveRec>& moves, const std::vector<uint8_t>& pairs, int n_nodes);
struct Dims { int n_nodes, n_units, n_species_staged, start_node; };
static std::string generate(const std::vector<MoveRec>& moves, const std::vector<uint8_t>& pairs, int stride, int pr, bool baro,
                            const TallyPlan& T, const Dims& D);
static const std::vector<char>* compile(const std::string& src, std::string* err, bool fresh = false);
static int load(int device, const std::string& src, const std::vector<char>& bin, cudaKernel_t* kern, std::string* err);
}  // namespace jit

static int validate(const StrykeFacility* f, const StrykeDayTables* d, const StrykeSpeciesTable* s,
                    const StrykeCells* c, const StrykeOpts* o) {
  if (!f || !d || !s || !c || !o) return fail(STRYKE_E_ARG, "null argument");
  if (f->n_nodes < 1 || f->n_moves < 1 || f->n_units < 0 || f->n_edges < 0 || f->n_pairs < 1)
    return fail(STRYKE_E_ARG, "empty facility tables");
  if (f->n_nodes > STRYKE_MAX_NODES || f->n_units > STRYKE_MAX_UNITS || f->n_edges > STRYKE_MAX_EDGES ||
      f->n_moves > STRYKE_MAX_MOVES || f->n_pairs > STRYKE_MAX_PAIRS)
    return fail(STRYKE_E_LIMIT, "facility exceeds compiled limits (nodes<=255, units<=64, edges<=1024, moves<=64, pairs<=128)");
  if (f->start_node < 0 || f->start_node >= f->n_nodes) return fail(STRYKE_E_ARG, "start_node out of range");
  if (d->n_days < 1 || s->n_species < 1 || c->n_cells < 0) return fail(STRYKE_E_ARG, "empty day/species/cell tables");
  if (o->precision != 32 && o->precision != 64) return fail(STRYKE_E_ARG, "precision must be 32 or 64");
  for (int i = 0; i < f->n_nodes; ++i) {
    if (f->node_kind[i] > STRYKE_KIND_PUMP) return fail(STRYKE_E_ARG, "unknown node kind");
    if (f->node_kind[i] != STRYKE_KIND_APRIORI && (f->node_unit[i] < 0 || f->node_unit[i] >= f->n_units))
      return fail(STRYKE_E_ARG, "turbine node without unit parameters (KeyError in the reference, stryke.py:1410)");
    const int deg = f->edge_off[i + 1] - f->edge_off[i];
    if (deg < 0 || deg > 255 || f->edge_off[i + 1] > f->n_edges) return fail(STRYKE_E_ARG, "bad edge_off");
  }
  for (int e = 0; e < f->n_edges; ++e)
    if (f->edge_dst[e] < 0 || f->edge_dst[e] >= f->n_nodes) return fail(STRYKE_E_ARG, "edge_dst out of range");
  for (int p = 0; p < f->n_pairs; ++p)
    if (f->pair_node[p] < 0 || f->pair_node[p] >= f->n_nodes) return fail(STRYKE_E_ARG, "pair_node out of range");
  if (f->level_off[0] != 0 || f->level_off[f->n_moves] != f->n_pairs) return fail(STRYKE_E_ARG, "bad level_off");
  for (int64_t i = 0; i < c->n_cells; ++i) {
    if (c->cell_day[i] < 0 || c->cell_day[i] >= d->n_days) return fail(STRYKE_E_ARG, "cell_day out of range");
    if (c->cell_species[i] < 0 || c->cell_species[i] >= s->n_species) return fail(STRYKE_E_ARG, "cell_species out of range");
  }
  return STRYKE_OK;
}

extern "C" {

const char* stryke_b200_last_error(void) { return g_err.c_str(); }
int stryke_b200_abi_version(void) { return STRYKE_B200_ABI_VERSION; }
uint64_t stryke_b200_launch_count(void) { return g_launches.load(); }
int stryke_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int stryke_b200_plan_destroy(StrykePlan* plan) {
  if (plan) {
    cudaSetDevice(plan->device);
    cudaDeviceSynchronize();      // launches on other streams may still read the tables
    delete plan;
  }
  return STRYKE_OK;
}

int stryke_b200_plan_create(const StrykeFacility* f, const StrykeDayTables* d, const StrykeSpeciesTable* s,
                            const StrykeCells* c, const StrykeOpts* o, StrykePlan** out) {
  if (!out) return fail(STRYKE_E_ARG, "null plan out-pointer");
  *out = nullptr;
  int rc = validate(f, d, s, c, o);
  if (rc) return rc;
  if (stryke_b200_device_count() <= 0)
    return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  CK(cudaSetDevice(o->device));
  StrykePlan* pl = new StrykePlan();
  struct Guard { StrykePlan* p; ~Guard() { if (p) delete p; } } guard{pl};
  pl->device = o->device; pl->precision = o->precision; pl->max_daily_fish = o->max_daily_fish;
  pl->inject = o->injected != nullptr;
  pl->n_cells = c->n_cells;
  if (device_sm_count(o->device, &pl->n_sm)) return fail(STRYKE_E_CUDA, "cannot query the SM count");

  // ---- static facility records -----------------------------------------------------------------------------------
  std::vector<NodeRec> nodes(f->n_nodes);
  for (int i = 0; i < f->n_nodes; ++i) {
    NodeRec& r = nodes[i];
    r.apriori32 = (float)f->node_apriori[i];
    r.edge_off = (int16_t)f->edge_off[i];
    r.deg = (uint8_t)(f->edge_off[i + 1] - f->edge_off[i]);
    r.kind = f->node_kind[i];
    r.unit = (int8_t)f->node_unit[i];
    r.hasU = f->node_has_U[i];
    r.pad = 0;
  }
  std::vector<uint8_t> edst(f->n_edges), pnode(f->n_pairs);
  for (int e = 0; e < f->n_edges; ++e) edst[e] = (uint8_t)f->edge_dst[e];
  for (int p = 0; p < f->n_pairs; ++p) pnode[p] = (uint8_t)f->pair_node[p];
  std::vector<MoveRec> moves(f->n_moves);
  // Philox word layout (shared by both passage kernels): block 0 words 0,1 = length draw.  Every level with per-node
  // draws owns a fresh block (rR, depth, cause, dice at positions 0..3).  A routing word takes a free position of the
  // block that is current at that point of the move loop, else opens a new block.
  int next_slot = 1;
  for (int k = 0; k < f->n_moves; ++k) {
    MoveRec& m = moves[k];
    m.lex = f->move_lex_rank[k];
    m.level_begin = (uint16_t)f->level_off[k];
    m.level_end = (uint16_t)f->level_off[k + 1];
    uint8_t need = 0;
    for (int p = m.level_begin; p < m.level_end; ++p) {
      const NodeRec& r = nodes[pnode[p]];
      if (r.kind != STRYKE_KIND_APRIORI) {
        need |= NEED_DICE | NEED_CAUSE;
        if (r.kind == STRYKE_KIND_KAPLAN) need |= NEED_RR;
        if (f->barotrauma) need |= NEED_DEPTH;
      } else if (!(r.apriori32 >= 1.0f)) {
        need |= NEED_DICE;
      }
      if (k < f->n_moves - 1 && r.deg > 1) need |= NEED_ROUTE;
    }
    m.need = need;
    // dense word slots in consumption order (slot 0 is the length word)
    m.slot_rc = m.slot_depth = m.slot_dice = m.slot_route = 0;
    if (need & (NEED_RR | NEED_CAUSE)) m.slot_rc = (uint16_t)next_slot++;
    if (need & NEED_DEPTH) m.slot_depth = (uint16_t)next_slot++;
    if (need & NEED_DICE) m.slot_dice = (uint16_t)next_slot++;
    if (need & NEED_ROUTE) m.slot_route = (uint16_t)next_slot++;
    int md = 0;
    for (int p = m.level_begin; p < m.level_end; ++p) md = std::max(md, (int)nodes[pnode[p]].deg);
    m.max_deg = (uint8_t)md;
    m.pad = 0;
  }
  CK(upload(pl->nodes, nodes.data(), nodes.size()));
  CK(upload(pl->edge_dst, edst.data(), edst.size()));
  CK(upload(pl->moves, moves.data(), moves.size()));
  CK(upload(pl->pair_node, pnode.data(), pnode.size()));
  CK(upload(pl->ustatic, f->unit_static, (size_t)f->n_units * STRYKE_UNIT_STATIC_W));
  CK(upload(pl->edge_cdf, d->edge_cdf, (size_t)d->n_days * f->n_edges));
  CK(upload(pl->node_stay, d->node_stay, (size_t)d->n_days * f->n_nodes));
  CK(upload(pl->unit_coef, d->unit_coef, (size_t)d->n_days * f->n_units * STRYKE_UNIT_COEF_W));
  CK(upload(pl->day_Q, d->curr_Q, (size_t)d->n_days));
  CK(upload(pl->species, s->params, (size_t)s->n_species * STRYKE_SPECIES_W));
  CK(upload(pl->cell_day, c->cell_day, (size_t)c->n_cells));
  CK(upload(pl->cell_species, c->cell_species, (size_t)c->n_cells));
  CK(upload(pl->cell_id, c->cell_id, (size_t)c->n_cells));
  CK(pl->tile_off.alloc((size_t)(c->n_cells + 1) * sizeof(int64_t)));
  CK(pl->fish_off.alloc((size_t)(c->n_cells + 1) * sizeof(int64_t)));
  pl->nb_scan = (int)((c->n_cells + SCAN_CH - 1) / SCAN_CH);
  CK(pl->part.alloc((size_t)(pl->nb_scan > 0 ? pl->nb_scan : 1) * 2 * sizeof(int64_t)));
  CK(pl->err.alloc(2 * sizeof(unsigned long long)));    // [0] error word of count_kernel, [1] chunk counter of the throughput kernels

  KParams& K = pl->K;
  K.nodes = pl->nodes.as<NodeRec>(); K.edge_dst = pl->edge_dst.as<uint8_t>(); K.moves = pl->moves.as<MoveRec>();
  K.pair_node = pl->pair_node.as<uint8_t>(); K.unit_static = pl->ustatic.as<double>();
  K.n_nodes = f->n_nodes; K.n_units = f->n_units; K.n_moves = f->n_moves; K.n_edges = f->n_edges;
  K.n_pairs = f->n_pairs; K.start_node = f->start_node; K.barotrauma = f->barotrauma;
  K.edge_cdf = pl->edge_cdf.as<double>(); K.node_stay = pl->node_stay.as<uint8_t>(); K.unit_coef = pl->unit_coef.as<double>();
  K.species = pl->species.as<double>(); K.n_species = s->n_species;
  K.cell_day = pl->cell_day.as<int32_t>(); K.cell_species = pl->cell_species.as<int32_t>(); K.cell_id = pl->cell_id.as<uint64_t>();
  K.tile_off = pl->tile_off.as<int64_t>(); K.fish_off = pl->fish_off.as<int64_t>(); K.n_cells = c->n_cells;
  K.seed_lo = (uint32_t)o->seed; K.seed_hi = (uint32_t)(o->seed >> 32);
  Philox::key_schedule(o->seed, K.ks);
  K.n_fish_total = 0;
  K.work_counter = pl->err.as<unsigned long long>() + 1;
  K.chunk_tiles = 128;
  if (const char* e = getenv("STRYKE_B200_CHUNK_TILES")) K.chunk_tiles = std::max(1, atoi(e));

  // ---- injected draws ----------------------------------------------------------------------------------------------
  int64_t NF = 0;
  if (o->injected) {
    const StrykeInjected* in = o->injected;
    NF = in->n_fish_total;
    const size_t M = (size_t)f->n_moves, C = (size_t)c->n_cells;
    if (!in->fish_off || !in->presence || !in->length_z || !in->dice || !in->rR || !in->depth || !in->cause || !in->route)
      return fail(STRYKE_E_ARG, "injected mode needs fish_off, presence, length_z, dice, rR, depth, cause, route");
    if (in->fish_off[C] != NF) return fail(STRYKE_E_ARG, "injected fish_off[n_cells] != n_fish_total");
    CK(upload(pl->i_fish_off, in->fish_off, C + 1));
    CK(upload(pl->i_presence, in->presence, C));
    CK(upload(pl->i_count, in->count_draw, in->count_draw ? C : 0));
    CK(upload(pl->i_z, in->length_z, (size_t)NF));
    CK(upload(pl->i_dice, in->dice, M * NF));
    CK(upload(pl->i_rR, in->rR, M * NF));
    CK(upload(pl->i_depth, in->depth, M * NF));
    CK(upload(pl->i_cause, in->cause, M * 4 * NF));
    CK(upload(pl->i_route, in->route, M * NF));
    K.inj.fish_off = pl->i_fish_off.as<int64_t>(); K.inj.presence = pl->i_presence.as<double>();
    K.inj.count_draw = pl->i_count.as<double>(); K.inj.length_z = pl->i_z.as<double>();
    K.inj.dice = pl->i_dice.as<double>(); K.inj.rR = pl->i_rR.as<double>(); K.inj.depth = pl->i_depth.as<double>();
    K.inj.cause = pl->i_cause.as<double>(); K.inj.route = pl->i_route.as<double>();
    if (in->ghost_cause) {
      if (!in->ghost_rR || !in->ghost_depth) return fail(STRYKE_E_ARG, "ghost_cause needs ghost_rR and ghost_depth");
      CK(upload(pl->i_grR, in->ghost_rR, M * C));
      CK(upload(pl->i_gdepth, in->ghost_depth, M * C));
      CK(upload(pl->i_gcause, in->ghost_cause, M * 4 * C));
      K.inj.ghost_rR = pl->i_grR.as<double>(); K.inj.ghost_depth = pl->i_gdepth.as<double>();
      K.inj.ghost_cause = pl->i_gcause.as<double>();
    }
    K.n_fish_total = NF;
  }
  // ---- optional per-fish trace -------------------------------------------------------------------------------------
  if (o->trace) {
    int64_t n_tr = NF;
    if (!o->injected && o->trace->n_fish > 0) {
      n_tr = o->trace->n_fish;      // population mode: count known from an earlier run with the same seed (checked in plan_status)
    } else if (!o->injected) {   // fish offsets must be known up front: only fixed-count (anadromous) cells
      n_tr = 0;
      for (int64_t i = 0; i < c->n_cells; ++i) {
        const double* sp = s->params + (size_t)c->cell_species[i] * STRYKE_SPECIES_W;
        if ((int)sp[12] != STRYKE_ENT_FIXED || !(sp[17] >= 1.0))
          return fail(STRYKE_E_ARG, "per-fish trace needs injected draws or fixed counts with occur_prob = 1");
        int64_t n = (int64_t)sp[18];
        n_tr += n == 0 ? 1 : n;
      }
    }
    const size_t M = (size_t)f->n_moves;
    pl->has_trace = true; pl->trace_fish = n_tr; pl->trace_probs = o->trace->strike != nullptr;
    CK(pl->t_len32.alloc((size_t)n_tr * 4)); CK(pl->t_len64.alloc((size_t)n_tr * 8));
    CK(pl->t_state.alloc(M * n_tr)); CK(pl->t_rate.alloc(M * n_tr * 4)); CK(pl->t_surv.alloc(M * n_tr));
    K.trace.length_ft = pl->t_len32.as<float>(); K.trace.length_ft64 = pl->t_len64.as<double>();
    K.trace.state = pl->t_state.as<uint8_t>(); K.trace.rate = pl->t_rate.as<float>(); K.trace.survival = pl->t_surv.as<uint8_t>();
    if (pl->trace_probs) {
      CK(pl->t_strike.alloc(M * n_tr * 8)); CK(pl->t_baro.alloc(M * n_tr * 8));
      K.trace.strike = pl->t_strike.as<double>(); K.trace.baro = pl->t_baro.as<double>();
    }
    K.n_fish_total = n_tr;
  }

  // ---- launch geometry -----------------------------------------------------------------------------------------------
  pl->n_species_staged = s->n_species <= 64 ? s->n_species : 0;
  pl->pr = f->n_pairs <= 32 ? 1 : f->n_pairs <= 64 ? 2 : 4;
  const SmemLayout SL = pl->precision == 64
      ? smem_layout<double>(f->n_nodes, f->n_units, f->n_moves, f->n_edges, f->n_pairs, pl->n_species_staged)
      : smem_layout<float>(f->n_nodes, f->n_units, f->n_moves, f->n_edges, f->n_pairs, pl->n_species_staged);
  pl->smem = (size_t)SL.total;
  if (pl->smem > 200 * 1024) return fail(STRYKE_E_LIMIT, "facility tables do not fit in shared memory");
  pl->blocks_per_sm = o->blocks_per_sm;
  // ---- throughput variant: fp32, native Philox, tallies only ---------------------------------------------------------
  pl->use_fast = pl->precision == 32 && !pl->inject && !pl->has_trace && o->kernel_variant != 1;
  if (pl->use_fast) {
    int stride = 1;
    for (int i = 0; i < f->n_nodes; ++i) stride = std::max(stride, (int)nodes[i].deg);
    pl->fast_stride = stride;
    pl->fast_smem_bytes = (size_t)fast_smem(f->n_nodes, f->n_units, stride, pl->n_species_staged).total;
    if (pl->fast_smem_bytes > 96 * 1024) pl->use_fast = false;   // very wide graphs: the general kernel's CSR layout
  }
  if (pl->use_fast) {
    for (int k = 0; k < f->n_moves; ++k) {
      MoveRec m = moves[k];
      const NodeRec& n0 = nodes[pnode[m.level_begin]];
      m.pad = pnode[m.level_begin];
      if (m.need == 0 && m.level_end - m.level_begin == 1 && !n0.hasU) m.need |= NEED_TRIVIAL;
      for (int p = m.level_begin; p < m.level_end; ++p)
        if (nodes[pnode[p]].hasU) m.need |= NEED_UCHECK;
      pl->fast.moves[k] = m;
    }
    for (int p = 0; p < f->n_pairs; ++p) pl->fast.pair_node[p] = pnode[p];
    const jit::TallyPlan tally = jit::plan_tally(std::vector<MoveRec>(pl->fast.moves, pl->fast.moves + f->n_moves), pnode, f->n_nodes);
    std::vector<uint2> fn(f->n_nodes);
    for (int i = 0; i < f->n_nodes; ++i) {
      fn[i].x = __float_as_uint_host(nodes[i].apriori32);
      fn[i].y = pack_node(nodes[i].kind, nodes[i].hasU, nodes[i].unit, tally.NW && tally.node_local[i] > 0 ? tally.node_local[i] : 0);
    }
    CK(upload(pl->fast_nodes, fn.data(), fn.size()));
    pl->kernel_name = "passage_fast_kernel (generic tables, ahead-of-time)";
    // ---- specialise for this facility when the run is large enough to amortise a ~1-3 s compile ---------------------
    double fish_hint = 0.0;
    for (int64_t i = 0; i < c->n_cells; ++i) {
      const double* sp = s->params + (size_t)c->cell_species[i] * STRYKE_SPECIES_W;
      fish_hint += (int)sp[12] == STRYKE_ENT_FIXED ? sp[18] : 50.0;
    }
    const char* env = getenv("STRYKE_B200_JIT");
    const bool want = o->kernel_variant == 3 || (o->kernel_variant == 0 && !(env && env[0] == '0') &&
                                                 ((env && env[0] == '1') || fish_hint >= 2.0e7));
    if (want) {
      std::vector<MoveRec> mv(pl->fast.moves, pl->fast.moves + f->n_moves);
      const std::string src = jit::generate(mv, pnode, pl->fast_stride, pl->pr, f->barotrauma != 0, tally,
                                                jit::Dims{f->n_nodes, f->n_units, pl->n_species_staged, f->start_node});
      std::string err;
      const std::vector<char>* bin = jit::compile(src, &err);
      cudaKernel_t k = nullptr;
      if (bin && jit::load(o->device, src, *bin, &k, &err) == 0) {
        pl->jit_kernel = k;
        pl->kernel_name = "passage_spec (tables specialised with NVRTC)";
      } else if (o->kernel_variant == 3) {
        return fail(STRYKE_E_CUDA, "kernel specialisation failed: " + err);
      }
    }
  } else {
    pl->kernel_name = "passage_kernel (general)";
  }
  if (o->kernel_variant == 3 && !pl->jit_kernel) return fail(STRYKE_E_ARG, "kernel_variant 3 needs fp32, native RNG, no trace");
  CK(cudaStreamSynchronize(0));      // uploads done: the caller may release its host tables
  guard.p = nullptr;
  *out = pl;
  return STRYKE_OK;
}

}  // extern "C"

template <typename Real, bool INJECT, int PR>
static int launch_passage(StrykePlan* pl, cudaStream_t st) {
  auto kern = passage_kernel<Real, INJECT, PR>;
  if (pl->smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
  int bps = pl->blocks_per_sm;
  if (bps <= 0) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, BLOCK, pl->smem));
    if (bps < 1) bps = 1;
  }
  kern<<<pl->n_sm * bps, BLOCK, pl->smem, st>>>(pl->K, pl->n_species_staged);
  g_launches++;
  CK(cudaGetLastError());
  return STRYKE_OK;
}
template <typename Real, bool INJECT>
static int launch_passage_pr(StrykePlan* pl, cudaStream_t st) {
  switch (pl->pr) {
    case 1: return launch_passage<Real, INJECT, 1>(pl, st);
    case 2: return launch_passage<Real, INJECT, 2>(pl, st);
    default: return launch_passage<Real, INJECT, 4>(pl, st);
  }
}

template <int PR, bool BARO, int MINB>
static int launch_fast(StrykePlan* pl, cudaStream_t st) {
  auto kern = passage_fast_kernel<PR, BARO, MINB>;
  if (pl->fast_smem_bytes > 48 * 1024)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->fast_smem_bytes));
  int bps = pl->blocks_per_sm;
  if (bps <= 0) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, BLOCK, pl->fast_smem_bytes));
    if (bps < 1) bps = 1;
  }
  kern<<<pl->n_sm * bps, BLOCK, pl->fast_smem_bytes, st>>>(pl->K, pl->fast, pl->fast_nodes.as<uint2>(), pl->fast_stride, pl->n_species_staged);
  g_launches++;
  CK(cudaGetLastError());
  return STRYKE_OK;
}
template <int MINB>
static int launch_fast_minb(StrykePlan* pl, cudaStream_t st) {
  const bool b = pl->K.barotrauma != 0;
  switch (pl->pr) {
    case 1: return b ? launch_fast<1, true, MINB>(pl, st) : launch_fast<1, false, MINB>(pl, st);
    case 2: return b ? launch_fast<2, true, MINB>(pl, st) : launch_fast<2, false, MINB>(pl, st);
    default: return b ? launch_fast<4, true, MINB>(pl, st) : launch_fast<4, false, MINB>(pl, st);
  }
}
// ------------------------------------------------------------------------------------------------------------------
// Run-time specialisation of the throughput kernel (NVRTC).  The hand-written body passage_fast_body<> is compiled
// once per facility shape with a table provider whose per-move records, pair list and row stride are compile-time
// constants generated from the plan: the move loop unrolls, every "does this level need a draw" test, Philox slot and
// pair index folds to an immediate.  libnvrtc is loaded lazily with dlopen; when it is not available the plan simply
// keeps the ahead-of-time generic instantiation (same results, tested).
// ------------------------------------------------------------------------------------------------------------------
namespace jit {
typedef struct _nvrtcProgram* nvrtcProgram;
struct Api {
  int (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*log_size)(nvrtcProgram, size_t*) = nullptr;
  int (*log)(nvrtcProgram, char*) = nullptr;
  int (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
  int (*cubin)(nvrtcProgram, char*) = nullptr;
  int (*destroy)(nvrtcProgram*) = nullptr;
  bool ok = false;
};
static Api& api() {
  static Api a = [] {
    Api x;
    void* h = nullptr;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return x;
    x.create = (decltype(x.create))dlsym(h, "nvrtcCreateProgram");
    x.compile = (decltype(x.compile))dlsym(h, "nvrtcCompileProgram");
    x.log_size = (decltype(x.log_size))dlsym(h, "nvrtcGetProgramLogSize");
    x.log = (decltype(x.log))dlsym(h, "nvrtcGetProgramLog");
    x.cubin_size = (decltype(x.cubin_size))dlsym(h, "nvrtcGetCUBINSize");
    x.cubin = (decltype(x.cubin))dlsym(h, "nvrtcGetCUBIN");
    x.destroy = (decltype(x.destroy))dlsym(h, "nvrtcDestroyProgram");
    x.ok = x.create && x.compile && x.log_size && x.log && x.cubin_size && x.cubin && x.destroy;
    return x;
  }();
  return a;
}
static std::string g_source_dir;
static std::mutex g_mu;
static std::map<std::string, std::vector<char>> g_cubins;        // generated source -> cubin
struct Loaded { cudaLibrary_t lib; cudaKernel_t kern; };
static std::map<std::pair<int, std::string>, Loaded> g_loaded;   // (device, source) -> loaded kernel

static bool read_file(const std::string& path, std::string* out) {
  std::ifstream f(path);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}


// NW stays 0 (ballot tallies) when the facility does not fit the scheme: move loop not unrolled (> 32 moves), too many
// words for the register file, a node whose local index differs between two levels, > 40 nodes in one level.
static TallyPlan plan_tally(const std::vector<MoveRec>& moves, const std::vector<uint8_t>& pairs, int n_nodes) {
  TallyPlan T;
  const int M = (int)moves.size();
  T.node_local.assign((size_t)std::max(n_nodes, 1), -1);
  if (M > 32) return T;
  std::vector<int> fill;             // 6-bit fields used in every word (5 per word)
  auto alloc = [&](int n) {          // n <= 5 contiguous fields inside one word -> (word, first bit)
    for (size_t w = 0; w < fill.size(); ++w)
      if (fill[w] + n <= 5) { const int b = 6 * fill[w]; fill[w] += n; return std::make_pair((int)w, b); }
    fill.push_back(n);
    return std::make_pair((int)fill.size() - 1, 0);
  };
  auto field = [&](int w, int sh, int bits, int target, int p0, int len) {
    T.fields.push_back(FieldRec{(uint8_t)w, (uint8_t)sh, (uint8_t)bits, (uint8_t)target, (uint8_t)p0, (uint8_t)len});
  };
  int n_cause = 0;
  T.lv.assign((size_t)M, LevelTally{});
  // passive level: nobody can die there (no dice <=> all nodes a-priori with survival >= 1), survivors = arrivals.
  // Consecutive passive ONE-node levels form a run: the same fish are counted at each of them, one field serves all.
  auto passive1 = [&](int k) { return !(moves[k].need & NEED_DICE) && moves[k].level_end - moves[k].level_begin == 1; };
  for (int k = 0; k < M; ++k) {
    const MoveRec& m = moves[k];
    const int n = m.level_end - m.level_begin;
    const bool same = !(m.need & NEED_DICE);
    if (m.need & NEED_CAUSE) ++n_cause;
    if (passive1(k)) {
      if (k > 0 && passive1(k - 1)) continue;                             // inside a run (kind 0)
      int run = 1;
      while (k + run < M && passive1(k + run)) ++run;                      // one pair per level, consecutive pair ids
      const auto a = alloc(1);
      T.lv[k] = LevelTally{1, (uint8_t)a.first, (uint8_t)a.second, 0, 0, 1, 1};
      field(a.first, a.second, 6, 2, m.level_begin, run);
    } else if (n <= 5) {
      const auto a = alloc(n);
      const auto b = same ? a : alloc(n);
      T.lv[k] = LevelTally{(uint8_t)(n == 1 ? 2 : 3), (uint8_t)a.first, (uint8_t)a.second, (uint8_t)b.first, (uint8_t)b.second, 1,
                           (uint8_t)same};
      for (int i = 0; i < n; ++i) {
        field(a.first, a.second + 6 * i, 6, same ? 2 : 0, m.level_begin + i, 1);
        if (!same) field(b.first, b.second + 6 * i, 6, 1, m.level_begin + i, 1);
      }
    } else {
      if (n > 40) return TallyPlan{};
      const int nwl = (n + 4) / 5;
      const int wc = (int)fill.size();
      for (int j = 0; j < (same ? 1 : 2) * nwl; ++j) fill.push_back(5);
      T.lv[k] = LevelTally{4, (uint8_t)wc, 0, (uint8_t)(wc + (same ? 0 : nwl)), 0, (uint8_t)nwl, (uint8_t)same};
      for (int i = 0; i < n; ++i) {
        field(wc + i / 5, 6 * (i % 5), 6, same ? 2 : 0, m.level_begin + i, 1);
        if (!same) field(wc + nwl + i / 5, 6 * (i % 5), 6, 1, m.level_begin + i, 1);
      }
    }
    if (n > 1)
      for (int i = 0; i < n; ++i) {
        const int v = pairs[m.level_begin + i];
        if (v >= (int)T.node_local.size()) T.node_local.resize((size_t)v + 1, -1);
        if (T.node_local[v] >= 0 && T.node_local[v] != i) return TallyPlan{};
        T.node_local[v] = i;
      }
  }
  const auto es = alloc(2);
  T.W_ES = es.first; T.B_ES = es.second;
  field(es.first, es.second, 6, 3, 0, 1);
  field(es.first, es.second + 6, 6, 3, 1, 1);
  T.W_CAUSE = (int)fill.size();
  fill.push_back(5);
  for (int j = 0; j < 3; ++j) field(T.W_CAUSE, 10 * j, 10, 3, 2 + j, 1);
  T.HOLD = std::min(63, 1023 / std::max(1, n_cause));
  // n_cause <= 31: the single-tile drain sums whole words, 32 lanes x n_cause must fit the 10-bit cause fields
  if ((int)fill.size() > 24 || T.fields.size() > 250 || T.HOLD < 4 || n_cause > 31) return TallyPlan{};
  T.NW = (int)fill.size();
  return T;
}

// resident 256-thread blocks per SM the specialised kernel is compiled for (register budget: 4 -> 64, 3 -> 80, 2 -> 128)
static int spec_min_blocks() {
  const char* e = getenv("STRYKE_B200_SPEC_MINB");
  const int v = e ? atoi(e) : 3;      // measured on B200: 3 (80 registers) beats 4 (64, rematerialises) and 2
  return v >= 1 && v <= 8 ? v : 3;
}

// The table provider + kernel entry for one plan.
static std::string generate(const std::vector<MoveRec>& moves, const std::vector<uint8_t>& pairs, int stride, int pr, bool baro,
                            const TallyPlan& T, const Dims& D) {
  std::ostringstream o;
  const int M = (int)moves.size(), P = (int)pairs.size();
  o << "#define STRYKE_NVRTC 1\n";
  if (const char* defs = getenv("STRYKE_B200_JIT_DEFINES")) {      // experiment switches: "A=1,B" -> #define A 1 / #define B
    std::string d(defs), tok;
    std::stringstream ss(d);
    while (std::getline(ss, tok, ',')) {
      if (tok.empty()) continue;
      const size_t eq = tok.find('=');
      o << "#define " << (eq == std::string::npos ? tok : tok.substr(0, eq) + " " + tok.substr(eq + 1)) << "\n";
    }
  }
  o << "#include \"passage_fast.cuh\"\nnamespace stryke {\nstruct SpecJit {\n"
    << "  static constexpr bool kStatic = true;\n"
    << "  static constexpr int M = " << M << ", STRIDE = " << stride << ", UNROLL_K = " << (M <= 32 ? M : 1) << ", UNROLL_P = 8;\n"
    << "  static constexpr int N_NODES = " << D.n_nodes << ", N_UNITS = " << D.n_units << ", N_PAIRS = " << P
    << ", N_SPECIES_STAGED = " << D.n_species_staged << ", START_NODE = " << D.start_node << ";\n"
    << "  static constexpr int NW = " << T.NW << ", NF = " << (T.NW ? (int)T.fields.size() : 0) << ", HOLD = " << T.HOLD
    << ", W_ES = " << T.W_ES << ", B_ES = " << T.B_ES << ", W_CAUSE = " << T.W_CAUSE << ";\n"
    << "  static __device__ constexpr MoveRec move(const FastTables&, int k) {\n    constexpr MoveRec T[" << M << "] = {";
  for (int k = 0; k < M; ++k) {
    const MoveRec& m = moves[k];
    o << (k ? ", " : "") << "{" << (int)m.lex << ", " << (int)m.need << ", " << m.slot_rc << ", " << m.level_begin << ", "
      << m.level_end << ", " << m.slot_route << ", " << (int)m.max_deg << ", " << (int)m.pad << ", " << m.slot_depth << ", "
      << m.slot_dice << "}";
  }
  o << "};\n    return T[k];\n  }\n  static __device__ constexpr int pair_node(const FastTables&, int p) {\n    constexpr uint8_t N[" << P << "] = {";
  for (int p = 0; p < P; ++p) o << (p ? ", " : "") << (int)pairs[p];
  o << "};\n    return (int)N[p];\n  }\n";
  if (T.NW) {
    o << "  static __device__ constexpr LevelTally tally(int k) {\n    constexpr LevelTally T[" << M << "] = {";
    for (int k = 0; k < M; ++k) {
      const LevelTally& l = T.lv[k];
      o << (k ? ", " : "") << "{" << (int)l.kind << ", " << (int)l.wc << ", " << (int)l.bc << ", " << (int)l.ws << ", " << (int)l.bs << ", "
        << (int)l.nwl << ", " << (int)l.same << "}";
    }
    o << "};\n    return T[k];\n  }\n  static __device__ constexpr FieldRec field(int f) {\n    constexpr FieldRec T[" << T.fields.size() << "] = {";
    for (size_t f = 0; f < T.fields.size(); ++f) {
      const FieldRec& r = T.fields[f];
      o << (f ? ", " : "") << "{" << (int)r.w << ", " << (int)r.sh << ", " << (int)r.bits << ", " << (int)r.target << ", " << (int)r.p0 << ", " << (int)r.len << "}";
    }
    o << "};\n    return T[f];\n  }\n";
    // per-lane views of the same layout for the single-tile drain
    const int n_lane = 32 * pr;
    std::vector<uint32_t> lcfg((size_t)n_lane, 0u), dcfg(32, 0u);
    for (const FieldRec& r : T.fields) {
      if (r.target == 3) {
        if (r.p0 < 32) dcfg[r.p0] = (uint32_t)r.w | ((uint32_t)r.sh << 5) | ((r.bits == 10 ? 1u : 0u) << 10) | (1u << 11);
        continue;
      }
      for (int p2 = r.p0; p2 < r.p0 + r.len && p2 < n_lane; ++p2) {
        if (r.target != 1) lcfg[p2] = (lcfg[p2] & ~0x3ffu) | (uint32_t)r.w | ((uint32_t)r.sh << 5);
        if (r.target != 0) lcfg[p2] = (lcfg[p2] & ~(0x3ffu << 10)) | ((uint32_t)r.w << 10) | ((uint32_t)r.sh << 15);
        lcfg[p2] |= 1u << 20;
      }
    }
    o << "  static __device__ uint32_t lane_cfg(int i) {\n    constexpr uint32_t T[" << n_lane << "] = {";
    for (int i = 0; i < n_lane; ++i) o << (i ? ", " : "") << lcfg[i] << "u";
    o << "};\n    return T[i];\n  }\n  static __device__ uint32_t daily_cfg(int i) {\n    constexpr uint32_t T[32] = {";
    for (int i = 0; i < 32; ++i) o << (i ? ", " : "") << dcfg[i] << "u";
    o << "};\n    return T[i];\n  }\n";
  } else {
    o << "  static __device__ constexpr LevelTally tally(int) { return LevelTally{}; }\n"
      << "  static __device__ constexpr FieldRec field(int) { return FieldRec{}; }\n"
      << "  static __device__ constexpr uint32_t lane_cfg(int) { return 0u; }\n"
      << "  static __device__ constexpr uint32_t daily_cfg(int) { return 0u; }\n";
  }
  o << "};\n}  // namespace stryke\n"
    << "extern \"C\" __global__ void __launch_bounds__(256, " << spec_min_blocks() << ") passage_spec(const stryke::KParams P, const stryke::FastTables FT,\n"
    << "    const uint2* __restrict__ g_nodes, const int stride, const int n_species_staged) {\n"
    << "  stryke::passage_fast_body<stryke::SpecJit, " << pr << ", " << (baro ? "true" : "false") << ">(P, FT, g_nodes, stride, n_species_staged);\n}\n";
  return o.str();
}

// nullptr + message on failure
static const std::vector<char>* compile(const std::string& src, std::string* err, bool fresh) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cubins.find(src);
  if (it != g_cubins.end() && !fresh) return &it->second;
  Api& a = api();
  if (!a.ok) { *err = "libnvrtc not available"; return nullptr; }
  std::string fast, common;
  if (g_source_dir.empty()) {      // default: csrc/ next to the shared library this code was loaded from
    Dl_info info;
    if (dladdr((const void*)&stryke_b200_abi_version, &info) && info.dli_fname) {
      std::string so(info.dli_fname);
      const size_t slash = so.rfind('/');
      g_source_dir = (slash == std::string::npos ? std::string(".") : so.substr(0, slash)) + "/csrc";
    }
  }
  if (g_source_dir.empty() || !read_file(g_source_dir + "/passage_fast.cuh", &fast) || !read_file(g_source_dir + "/passage_common.cuh", &common)) {
    *err = "kernel sources not found (stryke_b200_set_source_dir)";
    return nullptr;
  }
  // on-disk cache of compiled specialisations: key = FNV-1a of (generated source, both headers, target)
  uint64_t h = 1469598103934665603ull;
  const std::string* parts[3] = {&src, &fast, &common};
  for (const std::string* part : parts)
    for (unsigned char ch : *part) { h ^= ch; h *= 1099511628211ull; }
  std::string dir;
  if (const char* e = getenv("STRYKE_B200_CACHE_DIR")) dir = e;
  else dir = g_source_dir + "/../.jit_cache";      // inside the package directory
  char name[64];
  snprintf(name, sizeof name, "/sm_100a_%016llx.cubin", (unsigned long long)h);
  const std::string path = dir + name;
  if (!fresh) {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      std::vector<char> bin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
      if (bin.size() > 1024) return &(g_cubins[src] = std::move(bin));
    }
  }
  const char* headers[2] = {fast.c_str(), common.c_str()};
  const char* names[2] = {"passage_fast.cuh", "passage_common.cuh"};
  nvrtcProgram prog = nullptr;
  if (a.create(&prog, src.c_str(), "passage_spec.cu", 2, headers, names) != 0) { *err = "nvrtcCreateProgram failed"; return nullptr; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
  const int rc = a.compile(prog, 3, opts);
  if (rc != 0) {
    size_t n = 0;
    a.log_size(prog, &n);
    std::string log(n, '\0');
    if (n) a.log(prog, &log[0]);
    a.destroy(&prog);
    *err = "NVRTC compile failed: " + log.substr(0, 2000);
    return nullptr;
  }
  size_t n = 0;
  a.cubin_size(prog, &n);
  std::vector<char> bin(n);
  a.cubin(prog, bin.data());
  a.destroy(&prog);
  if (system(("mkdir -p '" + dir + "' 2>/dev/null").c_str()) == 0) {      // best effort; a read-only home is fine
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    std::ofstream f(tmp, std::ios::binary);
    if (f && f.write(bin.data(), (std::streamsize)bin.size())) { f.close(); rename(tmp.c_str(), path.c_str()); }
  }
  return &(g_cubins[src] = std::move(bin));
}

static int load(int device, const std::string& src, const std::vector<char>& bin, cudaKernel_t* kern, std::string* err) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_pair(device, src);
  auto it = g_loaded.find(key);
  if (it == g_loaded.end()) {
    Loaded l{};
    cudaError_t e = cudaLibraryLoadData(&l.lib, bin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&l.kern, l.lib, "passage_spec");
    if (e != cudaSuccess) { *err = std::string("loading the specialised kernel failed: ") + cudaGetErrorString(e); cudaGetLastError(); return 1; }
    it = g_loaded.emplace(key, l).first;
  }
  *kern = it->second.kern;
  return 0;
}
}  // namespace jit

static int launch_jit(StrykePlan* pl, cudaStream_t st) {
  const void* fn = (const void*)pl->jit_kernel;
  if (pl->fast_smem_bytes > 48 * 1024)
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->fast_smem_bytes));
  int bps = pl->blocks_per_sm;
  if (bps <= 0) bps = std::max(1, std::min(jit::spec_min_blocks(), (int)((200 * 1024) / std::max<size_t>(pl->fast_smem_bytes, 1))));
  const uint2* nodes = pl->fast_nodes.as<uint2>();
  void* args[] = {(void*)&pl->K, (void*)&pl->fast, (void*)&nodes, (void*)&pl->fast_stride, (void*)&pl->n_species_staged};
  CK(cudaLaunchKernel(fn, dim3(pl->n_sm * bps), dim3(BLOCK), args, pl->fast_smem_bytes, st));
  g_launches++;
  return STRYKE_OK;
}

static int launch_fast_any(StrykePlan* pl, cudaStream_t st) {
  if (pl->jit_kernel) return launch_jit(pl, st);
  return launch_fast_minb<4>(pl, st);
}

extern "C" {

int stryke_b200_plan_launch(StrykePlan* pl, int32_t* d_cell_n, uint32_t* d_daily, uint32_t* d_state_daily, void* stream) {
  if (!pl || !d_cell_n || !d_daily || !d_state_daily) return fail(STRYKE_E_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(pl->device));
  KParams& K = pl->K;
  K.cell_n = d_cell_n; K.daily = d_daily; K.state_daily = d_state_daily;
  if (pl->n_cells == 0) return STRYKE_OK;
  CK(cudaMemsetAsync(d_daily, 0, (size_t)pl->n_cells * STRYKE_DAILY_W * 4, st));
  CK(cudaMemsetAsync(d_state_daily, 0, (size_t)pl->n_cells * K.n_pairs * 2 * 4, st));
  CK(cudaMemsetAsync(pl->err.p, 0, 2 * sizeof(unsigned long long), st));
  const int cb = (int)((pl->n_cells + 255) / 256);
  if (pl->inject) count_kernel<true><<<cb, 256, 0, st>>>(K, pl->day_Q.as<double>(), d_cell_n, pl->max_daily_fish, pl->err.as<unsigned long long>());
  else count_kernel<false><<<cb, 256, 0, st>>>(K, pl->day_Q.as<double>(), d_cell_n, pl->max_daily_fish, pl->err.as<unsigned long long>());
  scan_partial<<<pl->nb_scan, SCAN_T, 0, st>>>(d_cell_n, pl->n_cells, pl->part.as<int64_t>());
  scan_blocks<<<1, SCAN_T, 0, st>>>(pl->part.as<int64_t>(), pl->nb_scan);
  scan_final<<<pl->nb_scan, SCAN_T, 0, st>>>(d_cell_n, pl->n_cells, pl->part.as<int64_t>(), pl->tile_off.as<int64_t>(), pl->fish_off.as<int64_t>());
  g_launches += 4;
  CK(cudaGetLastError());
  if (pl->use_fast) return launch_fast_any(pl, st);
  if (pl->precision == 64) return pl->inject ? launch_passage_pr<double, true>(pl, st) : launch_passage_pr<double, false>(pl, st);
  return pl->inject ? launch_passage_pr<float, true>(pl, st) : launch_passage_pr<float, false>(pl, st);
}

int stryke_b200_set_source_dir(const char* dir) {
  if (!dir) return fail(STRYKE_E_ARG, "null argument");
  std::lock_guard<std::mutex> lock(jit::g_mu);
  jit::g_source_dir = dir;
  return STRYKE_OK;
}

const char* stryke_b200_plan_kernel(const StrykePlan* pl) { return pl ? pl->kernel_name.c_str() : ""; }

// Compile (not load) the specialised kernel for a small built-in facility shape: needs libnvrtc, no GPU.
int stryke_b200_jit_selftest(void) {
  std::vector<MoveRec> mv = {{0, 32, 0, 0, 1, 0, 1, 0, 0, 0}, {1, 16, 0, 1, 2, 1, 5, 1, 0, 0}, {2, 75, 2, 2, 7, 0, 1, 2, 0, 3},
                             {3, 32, 0, 7, 8, 0, 1, 7, 0, 0}, {4, 32, 0, 8, 9, 0, 1, 8, 0, 0}, {5, 32, 0, 9, 10, 0, 0, 8, 0, 0}};
  std::vector<uint8_t> pairs = {0, 1, 2, 3, 4, 5, 6, 7, 8, 8};
  for (int baro = 0; baro < 2; ++baro) {
    std::string err;
    const jit::TallyPlan T = jit::plan_tally(mv, pairs, 9);
    if (!T.NW) return fail(STRYKE_E_CUDA, "packed tally layout failed for the self-test facility");
    if (!jit::compile(jit::generate(mv, pairs, 5, 1, baro != 0, T, jit::Dims{9, 3, 1, 0}), &err, /*fresh=*/true)) return fail(STRYKE_E_CUDA, err);
  }
  return STRYKE_OK;
}

int stryke_b200_plan_set_seed(StrykePlan* pl, uint64_t seed) {
  if (!pl) return fail(STRYKE_E_ARG, "null plan");
  pl->K.seed_lo = (uint32_t)seed; pl->K.seed_hi = (uint32_t)(seed >> 32);
  Philox::key_schedule(seed, pl->K.ks);
  return STRYKE_OK;
}

int stryke_b200_plan_status(StrykePlan* pl, void* stream, int64_t* n_fish_total) {
  if (!pl) return fail(STRYKE_E_ARG, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(pl->device));
  unsigned long long err = 0;
  int64_t nf = 0;
  CK(cudaMemcpyAsync(&err, pl->err.p, sizeof(err), cudaMemcpyDeviceToHost, st));
  if (pl->n_cells > 0) CK(cudaMemcpyAsync(&nf, pl->fish_off.as<int64_t>() + pl->n_cells, sizeof(nf), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_fish_total) *n_fish_total = nf;
  if (err & 0x8000000000000000ull)
    return fail(STRYKE_E_ARG, "Computed non-finite or negative population size in cell " + std::to_string(err & 0x7FFFFFFFFFFFFFFFull));
  if (err)
    return fail(STRYKE_E_MAX_FISH, "Computed population size n=" + std::to_string(err >> 32) + " exceeds STRYKE_MAX_DAILY_FISH=" +
                                       std::to_string(pl->max_daily_fish) + " (cell " + std::to_string(err & 0xFFFFFFFFull) + ")");
  if ((pl->inject || pl->has_trace) && nf != pl->K.n_fish_total)
    return fail(STRYKE_E_ARG, "fish count computed on the device (" + std::to_string(nf) + ") != injected/trace fish count (" +
                                  std::to_string(pl->K.n_fish_total) + ")");
  return STRYKE_OK;
}

int stryke_b200_plan_fetch_trace(StrykePlan* pl, const StrykeTrace* h, void* stream) {
  if (!pl || !h || !pl->has_trace) return fail(STRYKE_E_ARG, "plan has no trace");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(pl->device));
  const size_t n = (size_t)pl->trace_fish, M = (size_t)pl->K.n_moves;
  if (h->length_ft) CK(cudaMemcpyAsync(h->length_ft, pl->t_len32.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (h->length_ft64) CK(cudaMemcpyAsync(h->length_ft64, pl->t_len64.p, n * 8, cudaMemcpyDeviceToHost, st));
  if (h->state) CK(cudaMemcpyAsync(h->state, pl->t_state.p, M * n, cudaMemcpyDeviceToHost, st));
  if (h->rate) CK(cudaMemcpyAsync(h->rate, pl->t_rate.p, M * n * 4, cudaMemcpyDeviceToHost, st));
  if (h->survival) CK(cudaMemcpyAsync(h->survival, pl->t_surv.p, M * n, cudaMemcpyDeviceToHost, st));
  if (h->strike && pl->trace_probs) CK(cudaMemcpyAsync(h->strike, pl->t_strike.p, M * n * 8, cudaMemcpyDeviceToHost, st));
  if (h->baro && pl->trace_probs) CK(cudaMemcpyAsync(h->baro, pl->t_baro.p, M * n * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return STRYKE_OK;
}

int stryke_b200_run(const StrykeFacility* f, const StrykeDayTables* d, const StrykeSpeciesTable* s, const StrykeCells* c,
                    const StrykeOpts* o, StrykeTallies* out) {
  if (!out || !out->cell_n || !out->daily || !out->state_daily) return fail(STRYKE_E_ARG, "null tally buffers");
  StrykePlan* pl = nullptr;
  int rc = stryke_b200_plan_create(f, d, s, c, o, &pl);
  if (rc) return rc;
  struct Guard { StrykePlan* p; ~Guard() { stryke_b200_plan_destroy(p); } } guard{pl};
  DevBuf dn, dd, ds;
  const size_t C = (size_t)c->n_cells, P = (size_t)f->n_pairs;
  CK(dn.alloc(C * 4)); CK(dd.alloc(C * STRYKE_DAILY_W * 4)); CK(ds.alloc(C * P * 2 * 4));
  rc = stryke_b200_plan_launch(pl, dn.as<int32_t>(), dd.as<uint32_t>(), ds.as<uint32_t>(), nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out->cell_n, dn.p, C * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(out->daily, dd.p, C * STRYKE_DAILY_W * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(out->state_daily, ds.p, C * P * 2 * 4, cudaMemcpyDeviceToHost, 0));
  rc = stryke_b200_plan_status(pl, nullptr, nullptr);
  if (rc) return rc;
  if (o->trace) return stryke_b200_plan_fetch_trace(pl, o->trace, nullptr);
  return STRYKE_OK;
}

static int params15_to_coef(int kind, const double* p, double* uc) {
  // H, RPM, D, Q, ada, N, Qopt, Qper, lambda, iota, D1, D2, B, Q_p, gamma   (param_dict keys, stryke.py:2749-2787)
  const double H = p[0], RPM = p[1], D = p[2], Q = p[3], ada = p[4], N = p[5], Qopt = p[6], Qper = p[7], lam = p[8],
               iota = p[9], D1 = p[10], D2 = p[11], B = p[12], Q_p = p[13], gamma = p[14];
  const double g = 32.2, omega = RPM * ((2 * M_PI) / 60);
  const double Ewd = (g * H) / ((omega * D) * (omega * D)), Qwd = Q / (omega * (D * D * D));
  for (int i = 0; i < 8; ++i) uc[i] = 0.0;
  uc[0] = lam; uc[1] = N; uc[2] = D; uc[6] = Q > 0 ? 1.0 : 0.0;
  if (kind == STRYKE_KIND_KAPLAN) {
    uc[3] = M_PI * ada * Ewd; uc[4] = 2 * Qwd; uc[5] = 8 * Qwd;
  } else if (kind == STRYKE_KIND_PROPELLER) {
    const double rR = 0.75, Qwd_opt = Qopt / (omega * (D * D * D));
    const double beta = atan((M_PI / 8 * rR) / Qwd_opt);
    const double a = atan((M_PI * Ewd * ada) / (2 * Qwd * rR) + (M_PI / 8 * rR) / Qwd - tan(beta));
    uc[3] = (cos(a) / (8 * Qwd)) + sin(a) / (M_PI * rR);
  } else if (kind == STRYKE_KIND_FRANCIS) {
    const double r12 = D1 / D2;
    const double beta = atan((0.707 * (M_PI / 8)) / (iota * Qper * (r12 * r12 * r12)));
    const double alpha = atan((2 * M_PI * Ewd * ada) / Qwd * (B / D1) + (M_PI * (0.707 * 0.707)) / (2 * Qwd) * (B / D1) * ((D2 / D1) * (D2 / D1)) -
                              4 * 0.707 * tan(beta) * (B / D1) * (D1 / D2));
    uc[3] = (sin(alpha) * (B / D1)) / (2 * Qwd) + cos(alpha) / M_PI;
  } else if (kind == STRYKE_KIND_PUMP) {
    const double Qpwd = Q_p / (omega * (D * D * D)), r12 = D1 / D2;
    const double beta_p = atan((0.707 * M_PI / 8) / (Qpwd * (r12 * r12 * r12)));
    uc[0] = gamma; uc[2] = 0.707 * D2; uc[6] = 1.0;
    uc[3] = ((sin(beta_p) * (B / D1)) / (2 * Qpwd)) + (cos(beta_p) / M_PI);
  } else {
    return fail(STRYKE_E_ARG, "unknown runner kind");
  }
  return STRYKE_OK;
}

int stryke_b200_strike_survival(int kind, const double* params15, int64_t n, const double* L, const double* rR,
                                int precision, int device, double* out) {
  if (!params15 || !L || !out || n < 0) return fail(STRYKE_E_ARG, "null argument");
  if (kind == STRYKE_KIND_KAPLAN && !rR) return fail(STRYKE_E_ARG, "Kaplan needs the blade position rR");
  if (stryke_b200_device_count() <= 0) return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  double uc[8];
  int rc = params15_to_coef(kind, params15, uc);
  if (rc) return rc;
  if (n == 0) return STRYKE_OK;
  CK(cudaSetDevice(device));
  DevBuf dL, dR, dO, dU;
  CK(upload(dL, L, (size_t)n)); CK(upload(dR, rR, rR ? (size_t)n : 0)); CK(upload(dU, uc, 8)); CK(dO.alloc((size_t)n * 8));
  const int nb = (int)((n + 255) / 256);
  if (precision == 64) strike_kernel<double><<<nb, 256>>>(kind, dU.as<double>(), n, dL.as<double>(), dR.as<double>(), dO.as<double>());
  else strike_kernel<float><<<nb, 256>>>(kind, dU.as<double>(), n, dL.as<double>(), dR.as<double>(), dO.as<double>());
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dO.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return STRYKE_OK;
}

int stryke_b200_baro_survival(int64_t n, const double* depth, double tail, double b0, double b1, int precision,
                              int device, double* out) {
  if (!depth || !out || n < 0) return fail(STRYKE_E_ARG, "null argument");
  if (stryke_b200_device_count() <= 0) return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  if (n == 0) return STRYKE_OK;
  CK(cudaSetDevice(device));
  DevBuf dD, dO;
  CK(upload(dD, depth, (size_t)n)); CK(dO.alloc((size_t)n * 8));
  const int nb = (int)((n + 255) / 256);
  if (precision == 64) baro_kernel<double><<<nb, 256>>>(n, dD.as<double>(), tail, b0, b1, dO.as<double>());
  else baro_kernel<float><<<nb, 256>>>(n, dD.as<double>(), tail, b0, b1, dO.as<double>());
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dO.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return STRYKE_OK;
}

int stryke_b200_philox(int64_t n, const uint32_t* ctr4, const uint32_t* key2, int device, uint32_t* out4) {
  if (!ctr4 || !key2 || !out4 || n < 0) return fail(STRYKE_E_ARG, "null argument");
  if (stryke_b200_device_count() <= 0) return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  if (n == 0) return STRYKE_OK;
  CK(cudaSetDevice(device));
  DevBuf dC, dK, dO;
  CK(upload(dC, ctr4, (size_t)n * 4)); CK(upload(dK, key2, (size_t)n * 2)); CK(dO.alloc((size_t)n * 16));
  philox_kernel<<<(int)((n + 255) / 256), 256>>>(n, dC.as<uint32_t>(), dK.as<uint32_t>(), dO.as<uint32_t>());
  g_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpy(out4, dO.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
  return STRYKE_OK;
}

int stryke_b200_calibrate_issue(int device, void* stream, double* lane_ops_per_s, double* ms_out) {
  if (!lane_ops_per_s) return fail(STRYKE_E_ARG, "null argument");
  if (stryke_b200_device_count() <= 0) return fail(STRYKE_E_NOGPU, "no CUDA device visible: stryke_b200 has no CPU fallback");
  CK(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  int n_sm = 0;
  if (device_sm_count(device, &n_sm)) return fail(STRYKE_E_CUDA, "cannot query the SM count");
  DevBuf o;
  CK(o.alloc(16));
  const int grid = n_sm * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  calibrate_kernel<<<grid, 256, 0, st>>>(o.as<uint32_t>(), 12345u);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, st));
    calibrate_kernel<<<grid, 256, 0, st>>>(o.as<uint32_t>(), 12345u + rep);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  g_launches += 6;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  // per loop iteration and chain: IMAD (mul+add), LOP3 (xor/and/or fused) -> 2 lane-ops x 8 chains
  const double ops = (double)grid * 256.0 * (double)CAL_ITERS * (double)CAL_OPS;
  *lane_ops_per_s = ops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return STRYKE_OK;
}

}  // extern "C"

#include "bootstrap.cuh"
