// passage_kernels.cuh — device code of the general path (part of passage.cu): fp64 / fp32 probability functions in the
// reference's operation order, the count and scan kernels, the general passage kernel (injected or native draws,
// optional per-fish trace), the exported per-fish probability kernels, the Philox known-answer kernel and the
// issue-rate calibration kernel.  Included inside namespace stryke.
#pragma once

// fp64: the reference's operation order with every rounding kept (no FMA contraction)
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }

template <typename Real> struct UCW;
template <> struct UCW<float> { static constexpr int W = 5; };    // folded: fa, fb, fc|fl, has_flow, c0 of the day (0 = hydrostatic)
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
// friction-loss pressure mode (unit_coef slot 7 != 0): P1/P2 = c0 + c1 * depth with the day's c0 from the host (Darcy-Weisbach
// head loss of the unit's discharge, Stryke/barotrauma.py:43-137) and c1 = rho g 0.3048 / P2 of the unit
__device__ __forceinline__ double baro_surv64_ratio(double ratio, double b0, double b1) {
  const double e = exp(add_(b0, mul_(b1, log(ratio))));
  return sub_(1.0, div_(e, add_(1.0, e)));
}
__device__ __forceinline__ float baro_surv32(float depth_ft, float c0, float c1, float b0l2e, float b1) {
  const float ratio = fmaf(c1, depth_ft, c0);                 // P1/P2
  const float e = ex2_approx(fmaf(b1, lg2_approx(ratio), b0l2e));   // exp(b0 + b1 ln ratio)
  return rcp_approx(1.f + e);                                       // 1 - e/(1+e)
}

// ------------------------------------------------------------------------------------------------------------------
// native RNG: word slot -> Philox block, warp-uniform refill points
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {   // numpy-style 53-bit double in [0,1)
  return (double)((((uint64_t)hi >> 5) << 26) | ((uint64_t)lo >> 6)) * (1.0 / 9007199254740992.0);
}
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
  // wide stream (StrykeOpts.rng_mode = 1): one draw = two words (even slot) = a 53-bit uniform in [0, 1) like numpy's
  __device__ __forceinline__ double wide(int slot) { const uint32_t hi = word(slot), lo = word(slot + 1); return u53(hi, lo); }
};
// Wide stream layout: block 0 = the two uniforms of the length draw; move k owns blocks 1 + 4k .. 4 + 4k, draw j of the
// move sits at word slot 4 + 16 k + 2 j:  j = 0 r/R, 1 depth, 2..4 the reference's sequential cause rolls (stryke.py:1545-1560),
// 5 survival dice, 6 routing.
__device__ __forceinline__ int wide_slot(int k, int j) { return 4 + 16 * k + 2 * j; }
__device__ __forceinline__ double std_normal(double u1_oc, double u2_co) {
  return sqrt(-2.0 * log(u1_oc)) * cospi(2.0 * u2_co);
}

// ------------------------------------------------------------------------------------------------------------------
// (1) presence roll + daily entrainment count, one thread per cell
// ------------------------------------------------------------------------------------------------------------------

// Pooled remainders (throughput kernel, population mode: most present cells hold fewer than 32 fish).  Cells are taken
// in aligned blocks of 32; a cell of n fish keeps n >> 5 full tiles of its own, the n & 31 fish left over are laid end to
// end with those of the other cells of its block and cut into shared tiles.  A block whose cells differ in species or
// day ("mixed") keeps one tile per remainder instead (the day tables and species parameters are warp-uniform).  Returns
// the block's number of shared tiles (bit 31: mixed).  Called by all 32 lanes; lanes past the last cell pass in = false.
__device__ __forceinline__ uint32_t pool_block(bool in, int n, int sp, int dy, int pool_all, int big_full, int narrow_max) {
  const int sp0 = __shfl_sync(0xffffffffu, sp, 0), dy0 = __shfl_sync(0xffffffffu, dy, 0);
  const bool mixed = __ballot_sync(0xffffffffu, in && (sp != sp0 || dy != dy0)) != 0u;
  const int r = pool_rem(n, pool_all);
  const uint32_t tot = __reduce_add_sync(0xffffffffu, mixed ? (uint32_t)((r + 31) & ~31) : (uint32_t)r);
  const uint32_t owned = __reduce_add_sync(0xffffffffu, pool_big(n, pool_all, big_full, narrow_max) ? 0u : (uint32_t)pool_full(n, pool_all));
  return ((tot + 31u) >> 5) | (owned << POOL_OWNED_SHIFT) | (mixed ? POOL_MIXED : 0u);
}

// Per-species constants of the truncated Pareto draw (truncated_rvs, stryke.py:75-106): the upper truncation point, the
// CDF there and the exponent of the inverse CDF depend on the Population row only.  Computed once per plan, on the
// device (the same operations count_kernel would run per cell, so the counts do not change).
constexpr int STRYKE_PARETO_W = 4;      // -1/shape, upper, F(upper), 0
__global__ void pareto_constants_kernel(const double* __restrict__ species, int n_species, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_species) return;
  const double* sp = species + (int64_t)s * STRYKE_SPECIES_W;
  const double shape = sp[13], loc = sp[14], scale = sp[15], max_rate = sp[16];
  const bool capped = isfinite(max_rate) && max_rate > 0.0;
  const double upper = capped ? max_rate * 10.0 : loc + fmax(1.0, scale * 1000.0);
  const double yu = (upper - loc) / scale;
  out[s * STRYKE_PARETO_W + 0] = -1.0 / shape;
  out[s * STRYKE_PARETO_W + 1] = upper;
  out[s * STRYKE_PARETO_W + 2] = yu >= 1.0 ? 1.0 - pow(yu, -shape) : 0.0;
  out[s * STRYKE_PARETO_W + 3] = 0.0;
}

// ------------------------------------------------------------------------------------------------------------------
// (1) count_kernel: presence roll + daily entrainment count, one thread per cell — and, in the same pass, everything the
// passage kernel needs to find its work: the exclusive prefix sums of tiles, fish and row slots over the cells (a
// single-pass scan with decoupled look-back over the thread blocks), the per-block tables of the pooled remainders and of
// the compact rows, and the zeroed rows themselves.  One launch instead of count + three scan kernels + memsets.
// ------------------------------------------------------------------------------------------------------------------
constexpr int COUNT_T = 256, COUNT_K = 4, COUNT_CELLS = COUNT_T * COUNT_K;      // threads, blocks of 32 cells per warp, cells per thread block
#ifndef STRYKE_COUNT_MINB
#define STRYKE_COUNT_MINB 4
#endif
constexpr int COUNT_MINB = STRYKE_COUNT_MINB;      // resident thread blocks per SM the register budget is set for (63 registers, no spills)
// plan status words.  ERR_WORD: 0 = fine; top bits 11 = bounds check (debug build), 101 = cell_day / cell_species out of
// range, 100 = non-finite or negative count (low 32 bits: cell); else count << 32 | cell of a cell above max_daily_fish.
// ERR_WORK (chunk counter of the throughput kernels) and ERR_TICKET are cleared before every launch; ERR_SLOTS carries the
// slot count from one range of cells to the next, ERR_FISH_CUM the fish.
enum : int { ERR_WORD = 0, ERR_WORK = 1, ERR_TICKET = 2, ERR_FISH = 3, ERR_FISH_CUM = 4, ERR_SLOTS = 5, ERR_TILES = 6, ERR_N = 8 };
constexpr unsigned long long ERRBIT_RANGE = 0xA000000000000000ull, ERRBIT_NONFINITE = 0x8000000000000000ull;
// Decoupled look-back state: ONE 16-byte word per thread block of count_kernel, cleared before every launch.  State and the
// three running sums travel together in a single 128-bit store / load (as in CUB's single-pass scan), so no memory fence is
// needed between "value written" and "flag set":
//   x = state (0 empty, 1 aggregate of the block, 2 inclusive prefix) | slots << 2   (slots < 2^30)
//   y = tiles [31:0],  z = tiles [39:32] | fish [55:32] << 8,  w = fish [31:0]
struct ScanDev { uint4* status; };
__device__ __forceinline__ uint4 scan_pack(unsigned int state, long long tiles, long long fish, long long slots) {
  return make_uint4(state | ((unsigned int)slots << 2), (unsigned int)tiles,
                    ((unsigned int)((unsigned long long)tiles >> 32) & 0xffu) | ((unsigned int)((unsigned long long)fish >> 32) << 8),
                    (unsigned int)fish);
}
__device__ __forceinline__ void scan_unpack(const uint4 v, long long& tiles, long long& fish, long long& slots) {
  slots = (long long)(v.x >> 2);
  tiles = (long long)(((unsigned long long)(v.z & 0xffu) << 32) | v.y);
  fish = (long long)(((unsigned long long)(v.z >> 8) << 32) | v.w);
}
__device__ __forceinline__ void scan_store(uint4* p, const uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 scan_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cv.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
struct CountOut {
  int32_t* cell_n;
  int64_t* tile_off;              // [n_cells + 1] or nullptr (pooled plans use blk_tile)
  int64_t* fish_off;              // [n_cells + 1] or nullptr (only the injected / trace paths read it)
  int64_t* blk_tile;              // [n_blocks + 1] or nullptr
  uint32_t* blk_S;                // [n_blocks] or nullptr: shared tiles of the block | mixed << 31
  unsigned long long* err;        // ERR_* words
  long long slot0;                // compact rows: first slot of this launch, < 0 = continue after the previous launch
  unsigned long long* slots_after;    // optional: receives the slot count after this launch's cells
};

// One thread block = COUNT_CELLS consecutive cells: every warp takes COUNT_K aligned blocks of 32 cells, one after the
// other (phase 1: draw the counts, add up the warp's tiles / fish / slots), then ONE look-back per thread block, then the
// outputs that need the prefix (phase 2).  Four blocks per warp amortise the barrier and the look-back latency.
template <bool INJECT>
__global__ void __launch_bounds__(COUNT_T, COUNT_MINB) count_kernel(const KParams P, const double* __restrict__ day_Q, const CountOut O,
                                                        const int max_daily_fish, const double* __restrict__ pareto,
                                                        const ScanDev S, const int n_days, const int n_species) {
  __shared__ unsigned int s_bid;
  __shared__ GroupCursor s_cur;
  __shared__ long long s_w[COUNT_T / 32][3];
  __shared__ long long s_pre[3];
  __shared__ int s_n[COUNT_CELLS];
  __shared__ uint32_t s_info[COUNT_CELLS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    s_bid = atomicAdd(reinterpret_cast<unsigned int*>(O.err + ERR_TICKET), 1u);
    if (P.groups != nullptr) s_cur = group_cursor(P, (int64_t)s_bid * COUNT_CELLS);      // the group of the block's first cell
  }
  __syncthreads();
  const unsigned int bid = s_bid;      // thread blocks take their cells in ticket order: a block only ever waits for started ones
  const int64_t cta0 = (int64_t)bid * COUNT_CELLS;
  const uint32_t below = (1u << lane) - 1u;
  // ---- phase 1 ---------------------------------------------------------------------------------------------------------
  long long tw = 0, fw = 0, sw = 0;       // this warp's tiles, fish, row slots
#pragma unroll 1
  for (int k = 0; k < COUNT_K; ++k) {
  const int lc = (wid * COUNT_K + k) * 32 + lane;        // cell within the thread block
  const int64_t c = cta0 + lc;
  const bool in = c < P.n_cells;
  int species_c = 0, day_c = 0, n_c = 0;
  if (in) {
  const CellInfo ci = cell_at(P, s_cur, c);
  species_c = ci.species; day_c = ci.day;
  if (species_c < 0 || species_c >= n_species || day_c < 0 || day_c >= n_days) {      // (validated here, not on the host)
    atomicMax(O.err, ERRBIT_RANGE | (unsigned long long)(c & 0xFFFFFFFFll));
    species_c = 0; day_c = 0;
  } else {
  const double* sp = P.species + (int64_t)species_c * STRYKE_SPECIES_W;
  const int kind = (int)sp[12];
  const double shape = sp[13], loc = sp[14], scale = sp[15], max_rate = sp[16], occur = sp[17];
  uint32_t w[4], v[4] = {0u, 0u, 0u, 0u};
  bool idle = false;
  if (!INJECT && P.peak_units > 0) {      // peaking operations (stryke.py:2405-2431): coin + clipped lognormal hours per unit
    double tot = 0.0;
    for (int j = 0; j < P.peak_units; ++j) {
      const double* pp = P.peak_params + ((int64_t)species_c * P.peak_units + j) * 5;
      double h;
      if (isnan(pp[1])) {
        h = pp[4];
      } else {
        Philox::block(0xFFFFFFFEu - (uint32_t)j, (uint32_t)ci.id, (uint32_t)(ci.id >> 32), 0u, P.seed_lo, P.seed_hi, w);
        const double coin = ((double)w[0] + 0.5) * (1.0 / 4294967296.0);
        if (coin <= pp[0]) {
          h = 0.0;
        } else {
          const double z = sqrt(-2.0 * log(1.0 - u53(w[1], w[2]))) * cospi(2.0 * ((double)w[3] + 0.5) * (1.0 / 4294967296.0));
          h = exp(pp[1] * z) * pp[3] + pp[2];               // scipy lognorm._rvs
          h = h > 24.0 ? 24.0 : h < 0.0 ? 0.0 : h;
        }
      }
      if (P.peak_hours) P.peak_hours[(c + P.cell_base) * P.peak_units + j] = (float)h;
      tot += h;
    }
    idle = !(tot > 0.0);                                    // stryke.py:3025: nothing is drawn or written for the day
  }
  double presence;
  if (INJECT) {
    presence = P.inj.presence[c];
  } else {
    const uint64_t id = ci.id;
    Philox::block(0xFFFFFFFFu, (uint32_t)id, (uint32_t)(id >> 32), 0u, P.seed_lo, P.seed_hi, w);
    if (kind == STRYKE_ENT_LOGNORMAL) Philox::block(0xFFFFFFFFu, (uint32_t)id, (uint32_t)(id >> 32), 1u, P.seed_lo, P.seed_hi, v);
    presence = u53(w[0], w[1]);
  }
  int64_t n = 0;
  if (!idle && occur >= presence) {                              // stryke.py:3027
    if (kind == STRYKE_ENT_FIXED) {
      n = (int64_t)sp[18];                                       // stryke.py:3029
    } else {
      const bool capped = isfinite(max_rate) && max_rate > 0.0;
      double rate;
      if (kind == STRYKE_ENT_PARETO) {                           // truncated_rvs, stryke.py:75-106, :2550-2559
        const double* pc = pareto + species_c * STRYKE_PARETO_W;
        const double upper = pc[1];
        const double f_low = 0.0, f_high = pc[2];
        if (!isfinite(f_high) || f_high - f_low < 1e-12) {
          rate = upper;
        } else {
          const double q = INJECT ? P.inj.count_draw[c] : f_low + (f_high - f_low) * u53(w[2], w[3]);
          rate = loc + scale * pow(1.0 - q, pc[0]);              // scipy pareto._ppf
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
      const double Mft3 = (60.0 * 60.0 * 24.0 * day_Q[day_c]) / 1000000.0;   // stryke.py:2585
      const double nf = rint(Mft3 * rate);                       // np.round: half to even
      if (!isfinite(nf) || nf < 0.0 || nf > 2147483647.0) { atomicMax(O.err, ERRBIT_NONFINITE | (unsigned long long)(c & 0xFFFFFFFFll)); n = 0; }
      else n = (int64_t)nf;
    }
    if (n == 0 && !(P.flags & STRYKE_FLAG_NO_PROMOTE)) n = 1;    // stryke.py:3032-3033 (run(), not population_sim)
    if (n < 0 || n > 0x7FFFFFFFll) { atomicMax(O.err, ERRBIT_NONFINITE | (unsigned long long)(c & 0xFFFFFFFFll)); n = 0; }
    if (max_daily_fish > 0 && n > max_daily_fish) {              // stryke.py:2610-2616 (drawn counts), :3052 (fixed counts)
      atomicMax(O.err, ((unsigned long long)(n & 0x7FFFFFFFll) << 32) | (unsigned long long)(c & 0xFFFFFFFFll));
      n = 0;
    }
  }
  n_c = (int)n;
  }
  O.cell_n[c] = (int32_t)n_c;
  }
  s_n[lc] = n_c;
  const uint32_t mask = __ballot_sync(0xffffffffu, n_c > 0);
  const uint32_t wmask = P.out.rows ? __ballot_sync(0xffffffffu, n_c > P.out.narrow_max) : 0u;
  uint32_t info = 0;
  if (O.blk_S) info = pool_block(in, n_c, species_c, day_c, P.pool_all, P.big_full, P.out.rows ? P.out.narrow_max : 0x7fffffff);
  if (lane == 0) s_info[lc >> 5] = info;
  long long t = O.blk_S ? (long long)pool_full(n_c, P.pool_all) : (long long)((n_c + 31) >> 5);
  long long f = n_c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, o); f += __shfl_xor_sync(0xffffffffu, f, o); }
  tw += t + (long long)(info & POOL_S_MASK); fw += f; sw += P.out.rows ? __popc(mask) + __popc(wmask) : 0;
  }
  if (lane == 0) { s_w[wid][0] = tw; s_w[wid][1] = fw; s_w[wid][2] = sw; }
  __syncthreads();
  // ---- decoupled look-back over the thread blocks (warp 0) -----------------------------------------------------------
  if (wid == 0) {
    long long a[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < COUNT_T / 32; ++q) { a[0] += s_w[q][0]; a[1] += s_w[q][1]; a[2] += s_w[q][2]; }
    long long run[3] = {0, 0, 0};
    if (bid == 0) {
      run[2] = O.slot0 >= 0 ? O.slot0 : (long long)O.err[ERR_SLOTS];      // slots continue after the previous range of cells
      if (lane == 0) scan_store(S.status, scan_pack(2u, a[0], a[1], a[2] + run[2]));
    } else {
      if (lane == 0) scan_store(S.status + bid, scan_pack(1u, a[0], a[1], a[2]));
      long long look = (long long)bid - 1;
      for (;;) {
        const long long idx = look - lane;
        uint4 v = make_uint4(2u, 0u, 0u, 0u);                   // before the first block: an inclusive prefix of zero
        if (idx >= 0) {      // spin on the state word (a volatile 32-bit load), then take the whole 16-byte word in one load
          const volatile unsigned int* st = reinterpret_cast<const volatile unsigned int*>(S.status + idx);
          while ((*st & 3u) == 0u) { }
          v = scan_load(S.status + idx);
        }
        const uint32_t im = __ballot_sync(0xffffffffu, (v.x & 3u) == 2u);
        const int first = __ffs(im) - 1;                         // nearest block that already knows its inclusive prefix
        long long q[3] = {0, 0, 0};
        if (first < 0 || lane <= first) scan_unpack(v, q[0], q[1], q[2]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
          run[j] += q[j];
        }
        if (first >= 0) break;
        look -= 32;
      }
      if (lane == 0) scan_store(S.status + bid, scan_pack(2u, run[0] + a[0], run[1] + a[1], run[2] + a[2]));
    }
    if (lane == 0) {
      s_pre[0] = run[0]; s_pre[1] = run[1]; s_pre[2] = run[2];
      if (bid == gridDim.x - 1) {                                // the block of the last cells: totals and end markers
        O.err[ERR_TILES] = (unsigned long long)(run[0] + a[0]);
        O.err[ERR_FISH] = (unsigned long long)(run[1] + a[1]);
        O.err[ERR_FISH_CUM] += (unsigned long long)(run[1] + a[1]);
        O.err[ERR_SLOTS] = (unsigned long long)(run[2] + a[2]);
        if (O.slots_after) *O.slots_after = (unsigned long long)(run[2] + a[2]);
        if (O.tile_off) O.tile_off[P.n_cells] = run[0] + a[0];
        if (O.fish_off) O.fish_off[P.n_cells] = run[1] + a[1];
        if (O.blk_tile) O.blk_tile[(P.n_cells + 31) >> 5] = run[0] + a[0];
      }
    }
  }
  __syncthreads();
  // ---- phase 2: outputs ---------------------------------------------------------------------------------------------------
  long long bt = s_pre[0], bf = s_pre[1], bs = s_pre[2];
  for (int q = 0; q < wid; ++q) { bt += s_w[q][0]; bf += s_w[q][1]; bs += s_w[q][2]; }
#pragma unroll 1
  for (int k = 0; k < COUNT_K; ++k) {
    const int lc = (wid * COUNT_K + k) * 32 + lane;
    const int64_t c = cta0 + lc;
    const bool in = c < P.n_cells;
    const bool blk_in = (c & ~31ll) < P.n_cells;
    const int64_t blk = c >> 5;
    const int n_c = s_n[lc];
    const uint32_t info = s_info[lc >> 5];
    const uint32_t mask = __ballot_sync(0xffffffffu, n_c > 0);
    const uint32_t wmask = P.out.rows ? __ballot_sync(0xffffffffu, n_c > P.out.narrow_max) : 0u;
    const int slots_w = P.out.rows ? __popc(mask) + __popc(wmask) : 0;
    long long t = O.blk_S ? (long long)pool_full(n_c, P.pool_all) + (lane == 0 ? (long long)(info & POOL_S_MASK) : 0ll) : (long long)((n_c + 31) >> 5);
    long long f = n_c;
    const long long t_own = t, f_own = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long ta = __shfl_up_sync(0xffffffffu, t, o), fa = __shfl_up_sync(0xffffffffu, f, o);
      if (lane >= o) { t += ta; f += fa; }
    }
    if (in) {
      if (O.tile_off) O.tile_off[c] = bt + t - t_own;
      if (O.fish_off) O.fish_off[c] = bf + f - f_own;
    }
    if (blk_in && lane == 0) {
      if (O.blk_S) O.blk_S[blk] = info;
      if (O.blk_tile) O.blk_tile[blk] = bt;
    }
    if (P.out.rows && blk_in) {
      if (lane == 0) { P.out.blk_row[blk] = (uint32_t)bs; P.out.blk_mask[blk] = mask; P.out.blk_wide[blk] = wmask; }
      // Rows that receive red.global adds are zeroed here, and their cell's count goes into word 0: every row of a plan
      // without pooled remainders, the rows of the big cells of a pooled one (the owned cells' rows are written whole, once,
      // by the warp that simulates the block: passage_fast.cuh).
      const bool big = n_c > 0 && (!O.blk_S || pool_big(n_c, P.pool_all, P.big_full, P.out.narrow_max));
      const uint32_t bigm = __ballot_sync(0xffffffffu, big);
      const unsigned long long cap_w = (unsigned long long)P.out.cap * (unsigned long long)P.out.row_w;
      const unsigned long long slot = (unsigned long long)bs + __popc(mask & below) + __popc(wmask & below);
      if (bigm == mask) {                                    // all of them: one coalesced sweep over the block's rows
        const unsigned long long w0 = (unsigned long long)bs * (unsigned long long)P.out.row_w;
        const int nw = slots_w * P.out.row_w;
        for (int i = lane; i < nw; i += 32)
          if (w0 + i < cap_w) P.out.rows[w0 + i] = 0u;
        __syncwarp();
      } else if (big) {
        const unsigned long long w0 = slot * (unsigned long long)P.out.row_w;
        const int nw = (((wmask >> lane) & 1u) ? 2 : 1) * P.out.row_w;
        for (int i = 1; i < nw; ++i)
          if (w0 + i < cap_w) P.out.rows[w0 + i] = 0u;
      }
      if (big && slot + ((wmask >> lane) & 1u) < (unsigned long long)P.out.cap) P.out.rows[slot * P.out.row_w] = (uint32_t)n_c;
    }
    bt += __shfl_sync(0xffffffffu, t, 31); bf += __shfl_sync(0xffffffffu, f, 31); bs += slots_w;
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
  const bool wide = !INJECT && sizeof(Real) == 8 && P.rng_wide != 0;      // 53-bit uniforms, one draw per two words

  // per-cell registers
  Real sp_ls = 0, sp_ll = 0, sp_lsc = 0, sp_lm = 0, sp_lsd = 0, sp_wr = 0, sp_uc = 0, sp_d1 = 0, sp_d2 = 0, sp_b0 = 0, sp_b1 = 0;
  int sp_mode = 0, n_c = 0;
  uint64_t cid = 0;
  int64_t fo = 0;
  uint32_t acc_suc[PR], acc_cnt[PR], acc_daily = 0;
#pragma unroll
  for (int r = 0; r < PR; ++r) { acc_suc[r] = 0; acc_cnt[r] = 0; }

  auto flush = [&](int64_t cell) { tally_flush<PR>(P, cell, lane, P.n_pairs, acc_suc, acc_cnt, acc_daily); };

  for (int64_t t = t0; t < t1; ++t) {
    while (t >= c_end) { ++c; c_begin = c_end; c_end = P.tile_off[c + 1]; }
    if (c != cur_cell) {
      if (cur_cell >= 0) flush(cur_cell);
      cur_cell = c;
      n_c = P.cell_n[c];
      const CellInfo ci = cell_info(P, c);
      cid = ci.id;
      if (INJECT || P.trace.state || P.trace.length_ft64) fo = P.fish_off[c];
      const int s = ci.species;
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
      const int d = ci.day;
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
            s_ucoef[u * 5 + 0] = kap ? (Real)(r[3] / r[4]) : (Real)0;
            s_ucoef[u * 5 + 1] = kap ? (Real)(1.0 / r[5]) : (Real)0;
            s_ucoef[u * 5 + 2] = kap ? (Real)lnd : (Real)(lnd * r[3]);
            s_ucoef[u * 5 + 3] = (Real)r[6];
            s_ucoef[u * 5 + 4] = (Real)r[7];
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
    else if (wide) z = (Real)std_normal(1.0 - rng.wide(0), rng.wide(2));
    else { const uint32_t a = rng.word(0); z = std_normal(u16hi<Real>(a), u16lo<Real>(a)); }
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
      STRYKE_BCHK(P, loc >= 0 && loc < P.n_nodes, 11);
      const NodeRec nd = s_nodes[loc];
      const bool prev_alive = alive;
      float rate32 = 0.f;
      // word slots of this move (MoveRec): r/R + cause share one word, depth, dice, route one each
      const int s_dice = mv.slot_dice;
      const int s_rt = mv.slot_route;
      uint32_t w_rc = 0, w_dice = 0;
      if (!INJECT && !wide) {   // warp-uniform draws for this move (refill points are identical for every lane)
        if (mv.need & (NEED_RR | NEED_CAUSE)) w_rc = rng.word(mv.slot_rc);
        if (mv.need & NEED_DICE) w_dice = rng.word(s_dice);
      }
      const bool with_depth = (mv.need & NEED_DEPTH) != 0;     // barotrauma level: 12-bit r/R, depth from spare bits
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
              else if (wide) rR = (Real)add_(0.3, mul_(sub_(1.0, 0.3), rng.wide(wide_slot(k, 0))));
              else if (sizeof(Real) == 8) rR = (Real)(0.3 + 0.7 * (double)(with_depth ? u12hi<Real>(w_rc) : u16hi<Real>(w_rc)));
              else rR = (Real)fmaf(0.7f, (float)(with_depth ? u12hi<Real>(w_rc) : u16hi<Real>(w_rc)), 0.3f);
            }
            if (sizeof(Real) == 8) strike = (Real)strike_surv64(nd.kind, (const double*)uc, (double)L, (double)rR);
            else strike = (Real)strike_surv32(nd.kind, (const float*)uc, (float)L, (float)rR);
            strike_t = strike;
            if (P.barotrauma) {
              if (sizeof(Real) == 8) {
                const double fb = (double)us[2], lo = mul_(fb, (double)sp_d1), hi = mul_(fb, (double)sp_d2);
                const double ud = INJECT ? P.inj.depth[(int64_t)k * NF + gi] : wide ? rng.wide(wide_slot(k, 1)) : (double)u13_depth<Real>(w_rc, w_dice);
                const double dep = add_(lo, mul_(sub_(hi, lo), ud));
                if ((double)uc[7] != 0.0) baro = (Real)baro_surv64_ratio((double)uc[7] + (double)us[5] * dep, (double)sp_b0, (double)sp_b1);
                else baro = (Real)baro_surv64(dep, (double)us[3], (double)sp_b0, (double)sp_b1);
              } else {
                const float ud = INJECT ? (float)P.inj.depth[(int64_t)k * NF + gi] : (float)u13_depth<Real>(w_rc, w_dice);
                const float depth = (float)us[2] * fmaf((float)sp_d2 - (float)sp_d1, ud, (float)sp_d1);
                baro = (Real)baro_surv32(depth, (float)uc[4] != 0.f ? (float)uc[4] : (float)us[4], (float)us[5], (float)sp_b0 * 1.44269504f, (float)sp_b1);
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
          } else if (wide) {     // the reference's own sequence of rolls, each with a fresh uniform
            if (rng.wide(wide_slot(k, 2)) > (double)imp) cz += 1u;
            else if (rng.wide(wide_slot(k, 3)) > (double)strike) cz += 1u << 8;
            else if (rng.wide(wide_slot(k, 4)) > (double)baro) cz += 1u << 16;
          } else {
            // one uniform splits [0,1) into the same three probabilities the reference's sequential rolls have:
            // P(imp) = 1-imp, P(blade) = imp (1-strike), P(baro) = imp strike (1-baro)
            const Real u = u16lo<Real>(w_rc);
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
            const double dep = add_(lo, mul_(sub_(hi, lo), P.inj.ghost_depth[gk]));
            const double c0d = sizeof(Real) == 8 ? (double)uc[7] : (double)uc[4];
            baro = c0d != 0.0 ? baro_surv64_ratio(c0d + (double)us[5] * dep, (double)sp_b0, (double)sp_b1)
                              : baro_surv64(dep, (double)us[3], (double)sp_b0, (double)sp_b1);
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
      } else if (wide) {
        surv = alive && (rng.wide(wide_slot(k, 5)) <= (double)rate32);
      } else if (mv.need & NEED_DICE) {
        surv = alive && (u23<Real>(w_dice) <= (Real)rate32);
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
        if (!INJECT && wide) { if (surv && nd.deg > 1) u = (Real)rng.wide(wide_slot(k, 6)); }
        else if (!INJECT && (mv.need & NEED_ROUTE)) u = u23<Real>(rng.word(s_rt));
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
          STRYKE_BCHK(P, j >= 0 && j < (int)nd.deg && nd.edge_off + j < P.n_edges, 12);
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

