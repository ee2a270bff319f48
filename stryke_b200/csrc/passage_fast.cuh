// passage_fast.cuh — the throughput variant of the passage kernel: fp32 arithmetic, native Philox stream, tallies only.
//
// Same algorithm, same Philox word slots and same fp32 formulas as passage_kernel<float, false, PR> (which is
// validated bit-for-bit against the reference in injected mode), restructured for instruction issue, the roofline
// that binds this path:
//   * the warp index is made provably warp-uniform (shuffle), and the per-move records + the (move, node) pair list
//     live in the kernel parameter (constant) bank: the tile / cell / move / pair loops, every "does this level need
//     a draw" test and the Philox refill decisions run on the uniform datapath with no BSSY/BSYNC reconvergence code;
//   * levels that consist of one a-priori node with survival 1.0 and a single successor (river nodes, forebays,
//     tailraces: most levels of a real facility) take a ~10-instruction path: every live fish is there, survives and
//     moves on, so the tally is one ballot;
//   * the branchy per-fish code of node_surv_rate (stryke.py:1396-1572) is straight-line predicated arithmetic: one
//     unified strike formula (Kaplan and linear runners differ only in staged coefficients), selects instead of ifs;
//   * the day's routing table is staged per warp as fixed-stride rows (cdf padded with 2.0, successor ids padded with
//     the node itself; "fish stays today" rows point at themselves), so the categorical draw is a predicate-free
//     count of cdf[e] <= u and the move is one byte load;
//   * a level with per-node draws owns one Philox block (rR, depth, cause, dice at positions 0..3); the key schedule
//     comes straight from the constant bank (uniform adds).
//
// The body is a template over a *table provider*.  RuntimeTables (ahead-of-time build) reads the per-move records from
// the constant bank.  At plan creation the host can also generate a provider whose tables are compile-time constants of
// the plan's facility and compile this same body with NVRTC (specialise.inl): the move loop then unrolls
// completely and every need test, word slot and pair index folds away — measured 347 executed instructions per tile of
// 32 fish on the C2 facility against 660 for the generic instantiation.
// tests/test_gpu_native_rng.py requires identical tallies from the general kernel, the generic and the specialised
// instantiation of this body.
#pragma once
#include "passage_common.cuh"

namespace stryke {

struct FastTables {                    // by-value kernel parameter (constant bank)
  MoveRec moves[STRYKE_MAX_MOVES];
  uint8_t pair_node[STRYKE_MAX_PAIRS];
};

// FastNode.y: kind[0:3) hasU[3] unit[4:11); packed tallies: word offset of the node inside its level [24:27),
// 6 * (local index mod 5) [27:32)
__host__ __device__ inline uint32_t pack_node(int kind, int hasU, int unit, int local_index = 0) {
  return (uint32_t)kind | ((uint32_t)(hasU ? 1 : 0) << 3) | ((uint32_t)(unit < 0 ? 0 : unit) << 4) |
         ((uint32_t)(local_index / 5) << 24) | ((uint32_t)(6 * (local_index % 5)) << 27);
}

struct FastSmem {
  int nodes, ustatic, uc1, species, warp0, warp_stride, cdf, dst, ucoef, uc0, total;
  int pool_blk, pool_ef, pool_rank, pool_seg, pool_sd, pool_full;   // per-warp scratch of the pooled-remainder path (pool_nw > 0)
  int pool_row;                                    // ... and the rows of the block's owned cells being assembled (own)
  int pool_gcur;                                   // ... and the group cursor of the last block (cell groups)
  int pool_field;                                  // per block: where every (pair, arrivals | survivors) / Daily counter sits
};
// floats per row of the staged routing cdf: the stride-1 thresholds a draw is compared with, padded to whole float4s
__host__ __device__ constexpr int cdf_row(int stride) { return stride <= 5 ? 4 : ((stride - 1 + 3) & ~3); }
__host__ __device__ constexpr FastSmem fast_smem(int n_nodes, int n_units, int stride, int n_species_staged, int pool_nw = 0,
                                                 int n_pairs = 0, bool own = false) {
  FastSmem L{};
  int o = 0;
  L.nodes = o; o = (o + n_nodes * 8 + 15) & ~15;
  L.ustatic = o; o = (o + n_units * 16 + 15) & ~15;
  L.uc1 = o; o = (o + n_units * 4 + 15) & ~15;
  L.species = o; o = (o + n_species_staged * SPW * 4 + 15) & ~15;
  if (pool_nw > 0) { L.pool_field = o; o = (o + (2 * n_pairs + STRYKE_DAILY_W) * 4 + 15) & ~15; }
  L.warp0 = o;
  int w = 0;
  L.cdf = w; w = (w + n_nodes * cdf_row(stride) * 4 + 15) & ~15;
  L.ucoef = w; w = (w + n_units * 16 + 15) & ~15;
  L.uc0 = w; w = (w + n_units * 4 + 15) & ~15;                // barotrauma: P1/P2 offset of the day (friction-loss mode) or of the unit
  L.dst = w; w = (w + n_nodes * stride * 4 + 15) & ~15;     // int32 entries: no sub-word register packing in the hot loop
  if (pool_nw > 0) {
    L.pool_blk = w; w += 32 * 16;                            // uint4 per cell of the block: n, cell id (lo, hi), remainder offset
    L.pool_ef = w; w += 32 * 8;                              // int64 per cell: full tiles of the block before this cell
    L.pool_rank = w; w += 32 * 4;                            // lane of the rho-th cell that has a remainder
    L.pool_sd = w; w += 32 * 8;                              // int2 per cell: species row, day row
    L.pool_full = w; w += 32 * 4;                            // full tiles of the cell
    L.pool_seg = w; w = (w + 32 * pool_nw * 4 + 15) & ~15;         // per cell of the block: accumulated tally words
    if (own) { L.pool_row = w; w = (w + 32 * pool_rwp(n_pairs) * 4 + 15) & ~15; }      // per cell: its compact row (narrow layout)
    L.pool_gcur = w; w += 64;                                // GroupCursor
  }
  L.warp_stride = w;
  L.total = o + w * WARPS;
  return L;
}

// acc_a += a, acc_b += b on the lane whose id equals `p` (one ISETP + two predicated adds)
__device__ __forceinline__ void add_if_lane(uint32_t& acc_a, uint32_t& acc_b, int p, int lane_id, uint32_t a, uint32_t b) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, %3;\n\t@q add.u32 %0, %0, %4;\n\t@q add.u32 %1, %1, %5;\n\t}"
      : "+r"(acc_a), "+r"(acc_b) : "r"(p), "r"(lane_id), "r"(a), "r"(b));
}

// -1 when a <= b, else 0 (one FSET)
__device__ __forceinline__ int le_mask(float a, float b) {
  int r;
  asm("set.le.s32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}

// acc += a on the lane whose id equals `p`
__device__ __forceinline__ void add_if_lane1(uint32_t& acc, int p, int lane_id, uint32_t a) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %1, %2;\n\t@q add.u32 %0, %0, %3;\n\t}" : "+r"(acc) : "r"(p), "r"(lane_id), "r"(a));
}

// Packed per-lane tallies (specialised instantiation only; the layout is generated per facility by the host,
// jit::plan_tally in specialise.inl).  Every lane counts ITS OWN fish in 6-bit fields of a few registers ("tally words"):
// one field per (move, node) pair for arrivals and one for survivors, one per run of trivial levels, two for the
// Daily entrained / survived flags; the three mortality causes use 10-bit fields (a fish rolls a cause at every turbine
// level, stryke.py:1545-1560).  A tile adds at most 1 to a 6-bit field, so the words are only reduced across the warp
// (one REDUX per field, added to the lane that owns the pair) every HOLD tiles or when the cell changes — instead of
// two ballots + two popcounts + predicated adds per pair and tile.
struct LevelTally { uint8_t kind, wc, bc, ws, bs, nwl, same; };  // kind: 0 nothing (inside a passive run), 1 head of a passive
                                                              // run (field wc/bc), 2 one node, 3 <= 5 nodes (field base +
                                                              // 6 * local index, from the node record), 4 > 5 nodes
struct FieldRec { uint8_t w, sh, bits, target, p0, len; };    // target: 0 arrivals of pair p0, 1 survivors of pair p0,
                                                              // 2 both for pairs [p0, p0+len), 3 Daily column p0

// Table provider of the generic instantiation: per-move records and pair list from the constant bank at run time.
struct RuntimeTables {
  static constexpr bool kStatic = false;
  static constexpr int M = 1, STRIDE = 1, UNROLL_K = 1, UNROLL_P = 1;
  static constexpr int N_NODES = 0, N_UNITS = 0, N_PAIRS = 0, N_SPECIES_STAGED = 0, START_NODE = 0;   // run-time (KParams)
  static constexpr int NW = 0, NF = 0, HOLD = 1, W_ES = 0, B_ES = 0, W_CAUSE = 0;   // no packed tallies: ballots per pair
  static constexpr bool POOL = false;                 // pooled remainders need the packed tallies
  static constexpr bool OWN = false;                  // owned rows need the pooled remainders
  static __device__ __forceinline__ MoveRec move(const FastTables& FT, int k) { return FT.moves[k]; }
  static __device__ __forceinline__ int pair_node(const FastTables& FT, int p) { return (int)FT.pair_node[p]; }
  static __device__ __forceinline__ LevelTally tally(int) { return LevelTally{}; }
  static __device__ __forceinline__ FieldRec field(int) { return FieldRec{}; }
  static __device__ __forceinline__ uint32_t lane_cfg(int) { return 0u; }
  static __device__ __forceinline__ uint32_t daily_cfg(int) { return 0u; }
};

template <class TP, int PR, bool BARO>
__device__ __forceinline__ void passage_fast_body(const KParams& P, const FastTables& FT, const uint2* __restrict__ g_nodes,
                                                  const int stride_rt, const int n_species_staged) {
  // facility dimensions: compile-time constants in the specialised instantiation (the shared-memory layout folds to
  // immediates), kernel parameters otherwise
  const int stride = TP::kStatic ? TP::STRIDE : stride_rt;
  const int n_nodes = TP::kStatic ? TP::N_NODES : P.n_nodes, n_units = TP::kStatic ? TP::N_UNITS : P.n_units;
  const int n_pairs = TP::kStatic ? TP::N_PAIRS : P.n_pairs, start_node = TP::kStatic ? TP::START_NODE : P.start_node;
  const int n_sp = TP::kStatic ? TP::N_SPECIES_STAGED : n_species_staged;
  const int crow = cdf_row(stride);
  extern __shared__ __align__(16) unsigned char smem[];
  const FastSmem SL = fast_smem(n_nodes, n_units, stride, n_sp, TP::POOL ? TP::NW : 0, n_pairs, TP::OWN);
  uint2* s_nodes = reinterpret_cast<uint2*>(smem + SL.nodes);
  float4* s_ustatic = reinterpret_cast<float4*>(smem + SL.ustatic);     // rack, intake_vel, fb_depth, c0
  float* s_uc1 = reinterpret_cast<float*>(smem + SL.uc1);               // c1
  float* s_species = reinterpret_cast<float*>(smem + SL.species);
  const int lane = threadIdx.x & 31;
  const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  unsigned char* wbase = smem + SL.warp0 + wid * SL.warp_stride;
  float* s_cdf = reinterpret_cast<float*>(wbase + SL.cdf);              // [n_nodes][cdf_row(stride)], padded with 2.0
  float4* s_ucoef = reinterpret_cast<float4*>(wbase + SL.ucoef);        // fa, fb, fc, has_flow
  float* s_uc0 = reinterpret_cast<float*>(wbase + SL.uc0);              // c0: P1/P2 = c0 + c1 * depth
  int* s_dst = reinterpret_cast<int*>(wbase + SL.dst);                  // [n_nodes][stride], padded with the node itself

  for (int i = threadIdx.x; i < n_nodes; i += BLOCK) s_nodes[i] = g_nodes[i];
  for (int u = threadIdx.x; u < n_units; u += BLOCK) {
    const double* us = P.unit_static + u * STRYKE_UNIT_STATIC_W;
    const double p2 = P_ATM + RHO * G_SI * (us[3] * FT2M);
    s_ustatic[u] = make_float4((float)us[0], (float)us[1], (float)us[2], (float)(P_ATM / p2));
    s_uc1[u] = (float)(RHO * G_SI * FT2M / p2);
  }
  for (int i = threadIdx.x; i < n_sp * SPW; i += BLOCK)
    s_species[i] = (float)P.species[(i / SPW) * STRYKE_SPECIES_W + (i % SPW)];
  __syncthreads();

  int cur_day = -1, cur_species = -1;                 // what the staged day tables / the species registers hold
  float sp_ls = 0, sp_ll = 0, sp_lsc = 0, sp_lm = 0, sp_lsd = 0, sp_wr = 0, sp_uc = 0, sp_d1 = 0, sp_dd = 0, sp_b0 = 0, sp_b1 = 0;
  int sp_mode = 0;
  uint32_t acc_suc[PR], acc_cnt[PR], acc_daily = 0;
  int lane_r[PR];
#pragma unroll
  for (int r = 0; r < PR; ++r) { acc_suc[r] = 0; acc_cnt[r] = 0; lane_r[r] = lane + 32 * r; }

  constexpr bool PK = TP::NW > 0;                     // packed per-lane tallies (see LevelTally)
  uint32_t W[PK ? TP::NW : 1];
#pragma unroll
  for (int j = 0; j < (PK ? TP::NW : 1); ++j) W[j] = 0;
  int held = 0;                                       // tiles accumulated in W since the last drain
  auto drain = [&]() {                                // W (per-lane fields) -> warp sums -> the lanes that own the pairs
    if constexpr (PK) {
#pragma unroll
      for (int f = 0; f < TP::NF; ++f) {
        const FieldRec F = TP::field(f);
        const uint32_t v = __reduce_add_sync(0xffffffffu, (W[F.w] >> F.sh) & ((1u << F.bits) - 1u));
        if (F.target == 3) {
          add_if_lane1(acc_daily, F.p0, lane, v);
        } else {
#pragma unroll
          for (int r = 0; r < PR; ++r) {
            if (F.p0 + F.len <= 32 * r || F.p0 >= 32 * (r + 1)) continue;
            const bool mine = (unsigned)(lane_r[r] - (int)F.p0) < (unsigned)F.len;
            if (F.target != 1) acc_cnt[r] += mine ? v : 0u;
            if (F.target != 0) acc_suc[r] += mine ? v : 0u;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TP::NW; ++j) W[j] = 0;
      held = 0;
    }
  };

  // Single-tile drain (most cells of a population-mode run hold fewer than 32 fish): after ONE tile every field of a lane
  // is at most its per-tile maximum, so whole words can be summed (one REDUX per word instead of one per field) and
  // every lane picks the fields of its own pairs.  lane_cfg: arrivals word [0:5) shift [5:10), survivors word [10:15)
  // shift [15:20), pair exists [20]; daily_cfg: word [0:5) shift [5:10), 10-bit field [10], column exists [11].
  uint32_t lcfg[PR], dcfg = 0;
#pragma unroll
  for (int r = 0; r < PR; ++r) lcfg[r] = PK ? TP::lane_cfg(lane + 32 * r) : 0u;
  if (PK) dcfg = TP::daily_cfg(lane);
  auto drain_one_tile = [&]() {
    if constexpr (PK) {
      uint32_t S[TP::NW];
#pragma unroll
      for (int j = 0; j < TP::NW; ++j) S[j] = __reduce_add_sync(0xffffffffu, W[j]);
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        const uint32_t cf = lcfg[r];
        const uint32_t cw = cf & 31u, sw = (cf >> 10) & 31u;
        uint32_t vc = S[0], vs = S[0];
#pragma unroll
        for (int j = 1; j < TP::NW; ++j) { vc = cw == (uint32_t)j ? S[j] : vc; vs = sw == (uint32_t)j ? S[j] : vs; }
        const uint32_t m = (cf >> 20) & 1u ? 63u : 0u;
        acc_cnt[r] += (vc >> ((cf >> 5) & 31u)) & m;
        acc_suc[r] += (vs >> ((cf >> 15) & 31u)) & m;
      }
      {
        const uint32_t dw = dcfg & 31u;
        uint32_t vd = S[0];
#pragma unroll
        for (int j = 1; j < TP::NW; ++j) vd = dw == (uint32_t)j ? S[j] : vd;
        const uint32_t m = (dcfg >> 11) & 1u ? ((dcfg >> 10) & 1u ? 1023u : 63u) : 0u;
        acc_daily += (vd >> ((dcfg >> 5) & 31u)) & m;
      }
#pragma unroll
      for (int j = 0; j < TP::NW; ++j) W[j] = 0;
      held = 0;
    }
  };

  auto flush = [&](int64_t cell) { tally_flush<PR>(P, cell, lane, n_pairs, acc_suc, acc_cnt, acc_daily); };

  const int n_moves = TP::kStatic ? TP::M : P.n_moves;

  // species parameters and the day's tables of cell c (both kept across cells: neighbours share them)
  auto load_params = [&](const int s, const int d) {
    if (s == cur_species) {
      // cells are ordered group by group: the neighbours of a cell almost always belong to the same species
    } else if (s < n_sp) {
      cur_species = s;
      const float* q = s_species + s * SPW;
      sp_ls = q[0]; sp_ll = q[1]; sp_lsc = q[2]; sp_lm = q[3]; sp_lsd = q[4]; sp_mode = (int)q[5];
      sp_wr = q[6]; sp_uc = q[7]; sp_d1 = q[8]; sp_dd = q[9] - q[8]; sp_b0 = q[10] * 1.44269504f; sp_b1 = q[11];
    } else {
      cur_species = s;
      const double* q = P.species + (int64_t)s * STRYKE_SPECIES_W;
      sp_ls = (float)q[0]; sp_ll = (float)q[1]; sp_lsc = (float)q[2]; sp_lm = (float)q[3]; sp_lsd = (float)q[4];
      sp_mode = (int)q[5]; sp_wr = (float)q[6]; sp_uc = (float)q[7]; sp_d1 = (float)q[8];
      sp_dd = (float)q[9] - (float)q[8]; sp_b0 = (float)q[10] * 1.44269504f; sp_b1 = (float)q[11];
    }
    if (d != cur_day) {   // stage the day's routing rows and unit coefficients in this warp's slot
      cur_day = d;
      __syncwarp();
      const double* cdf = P.edge_cdf + (size_t)d * P.n_edges;
      const uint8_t* st = P.node_stay + (size_t)d * n_nodes;
      for (int i = lane; i < n_nodes * crow; i += 32) {
        const int v = i / crow, e = i - v * crow;
        const NodeRec nr = P.nodes[v];
        s_cdf[i] = (e < (int)nr.deg && !st[v]) ? (float)cdf[nr.edge_off + e] : 2.0f;
      }
      for (int i = lane; i < n_nodes * stride; i += 32) {
        const int v = i / stride, e = i - v * stride;
        const NodeRec nr = P.nodes[v];
        s_dst[i] = (e < (int)nr.deg && !st[v]) ? (int)P.edge_dst[nr.edge_off + e] : v;
      }
      const double* uc = P.unit_coef + (size_t)d * n_units * STRYKE_UNIT_COEF_W;
      for (int u = lane; u < n_units; u += 32) {
        const double* r = uc + u * STRYKE_UNIT_COEF_W;
        const double lnd = r[0] * r[1] / r[2];
        const bool kap = r[4] != 0.0 || r[5] != 0.0;
        // linear runners use the Kaplan formula with fa = 0, fb = 1, rR = 1, inv = 1:  p = fc * L
        s_ucoef[u] = make_float4(kap ? (float)(r[3] / r[4]) : 0.f, kap ? (float)(1.0 / r[5]) : 1.f,
                                 kap ? (float)lnd : (float)(lnd * r[3]), (float)r[6]);
        if (BARO) s_uc0[u] = r[7] != 0.0 ? (float)r[7] : s_ustatic[u].w;      // friction-loss mode: the day's offset
      }
      __syncwarp();
    }
  };

  // one tile: lane simulates fish `fish` of the cell with id (cid_lo, cid_hi); tallies go to W (packed) or acc_*
  auto simulate = [&](const int fish, const bool active, const uint32_t cid_lo, const uint32_t cid_hi) {
    // ---- Philox: counter (fish, cell id, block), key = seed (constant bank) -------------------------------------
    uint32_t w[4];
    int cur_b = 0;
    Philox::block_ks((uint32_t)fish, cid_lo, cid_hi, 0u, P.ks, w);
    // word slot s -> block s >> 2 (generated when first needed: the slots are consumed in increasing order and the
    // test is warp-uniform, compile-time in the specialised instantiation), position s & 3
    auto word = [&](int s) -> uint32_t {
      if ((s >> 2) != cur_b) {
        cur_b = s >> 2;
        Philox::block_ks((uint32_t)fish, cid_lo, cid_hi, (uint32_t)cur_b, P.ks, w);
      }
      const int i = s & 3;
      return (i & 2) ? ((i & 1) ? w[3] : w[2]) : ((i & 1) ? w[1] : w[0]);
    };

    // ---- (2) length: word 0, radius from the high half, angle from the low half --------------------------------------
    const float z = std_normal(u16hi<float>(w[0]), u16lo<float>(w[0]));
    float L;
    if (sp_mode == 0) L = fminf(fabsf(fmaf(exp_approx(sp_ls * z), sp_lsc, sp_ll)), 150.f) * 0.0328084f;
    else L = fabsf(fmaf(sp_lsd, z, sp_lm)) * (1.f / 12.f);
    const float blocked_w = L * sp_wr;

    int loc = start_node;
    bool alive = active;
    int best_rank = 255;
    bool ent_surv = false;
    uint32_t cz = 0;

#pragma unroll(TP::UNROLL_K)
    for (int k = 0; k < n_moves; ++k) {
      const MoveRec mv = TP::move(FT, k);                // constant bank (uniform index) or compile-time constant
      if (mv.need & NEED_TRIVIAL) {
        // one a-priori node, survival 1.0, at most one successor, no 'U' in its name: every live fish is here,
        // survives (stryke.py:3210 with rate 1.0) and moves to the only successor (or stays, stryke.py:1849-1855)
        if constexpr (PK) {
          const LevelTally LT = TP::tally(k);           // one field for a whole run of trivial levels
          if (LT.kind == 1) W[LT.wc] += alive ? (1u << LT.bc) : 0u;
        } else {
          const uint32_t pc = __popc(__ballot_sync(0xffffffffu, alive));
#pragma unroll
          for (int r = 0; r < PR; ++r) add_if_lane(acc_cnt[r], acc_suc[r], mv.level_begin, lane_r[r], pc, pc);
        }
        STRYKE_BCHK(P, (int)mv.pad < n_nodes, 1);
        if (k < n_moves - 1) loc = alive ? s_dst[(int)mv.pad * stride] : loc;
        continue;
      }
      // a level with ONE reachable node: every live fish is there, so its index is a compile-time constant in the
      // specialised instantiation (table offsets fold to immediates; dead lanes read a valid record they never use)
      const int at = (TP::kStatic && mv.level_end - mv.level_begin == 1) ? TP::pair_node(FT, mv.level_begin) : loc;
      STRYKE_BCHK(P, at >= 0 && at < n_nodes, 2);
      const uint2 nd = s_nodes[at];
      const uint32_t pk = nd.y;
      const int kind = pk & 7;
      const bool prev_alive = alive;
      float rate = alive ? __uint_as_float(nd.x) : 0.f;
      uint32_t w_dice = 0;
      if (mv.need & NEED_CAUSE) {                        // this level has turbine nodes (uniform)
        const uint32_t w_rc = word(mv.slot_rc);          // r/R in the high half, cause roll in the low half
        const float u_cau = u16lo<float>(w_rc);
        // barotrauma: 12-bit r/R, depth draw from spare bits of this word and of the dice word (passage_common.cuh)
        const float u_rr = BARO ? u12hi<float>(w_rc) : u16hi<float>(w_rc);
        if (BARO) w_dice = word(mv.slot_dice);
        const float u_dep = BARO ? u13_depth<float>(w_rc, w_dice) : 0.f;
        const bool turb = alive && kind != STRYKE_KIND_APRIORI;
        const int unit = turb ? (int)((pk >> 4) & 0x7f) : 0;
        STRYKE_BCHK(P, unit >= 0 && unit < (n_units > 0 ? n_units : 1), 3);
        const float4 us = s_ustatic[unit];
        const float4 uc = s_ucoef[unit];
        const bool blocked = blocked_w > us.x;                               // stryke.py:1419
        const float imp = (blocked && sp_uc < us.y) ? 0.f : 1.f;             // stryke.py:1420-1423
        const bool kap = kind == STRYKE_KIND_KAPLAN;
        const float rR = kap ? fmaf(0.7f, u_rr, 0.3f) : 1.f;
        const float inv = kap ? rsqrt_approx(fmaf(rR, rR, uc.x * uc.x)) : 1.f;   // exact 1 for the linear runners
        float p = uc.z * L * inv * fmaf(rR, uc.y, uc.x * 0.318309886f * rcp_approx(rR));
        p = uc.w > 0.f ? p : __int_as_float(0x7f800000);                     // no flow: p_strike = inf (Appendix E10)
        const float strike = blocked ? 1.f : __saturatef(1.f - p);
        float baro = 1.f;
        if (BARO) {
          const float depth = us.z * fmaf(sp_dd, u_dep, sp_d1);
          const float ratio = fmaf(s_uc1[unit], depth, s_uc0[unit]);
          const float e = ex2_approx(fmaf(sp_b1, lg2_approx(ratio), sp_b0));
          baro = blocked ? 1.f : rcp_approx(1.f + e);
        }
        const float a2 = imp * strike, a3 = a2 * baro;
        rate = turb ? a3 : rate;
        constexpr int CB = PK ? 10 : 8;                                      // width of the per-lane cause counters
        const uint32_t code = u_cau >= imp ? 1u : u_cau >= a2 ? (1u << CB) : u_cau >= a3 ? (1u << (2 * CB)) : 0u;
        cz += turb ? code : 0u;
      }
      bool surv;
      if (mv.need & NEED_DICE) surv = alive && (u23<float>((BARO && (mv.need & NEED_CAUSE)) ? w_dice : word(mv.slot_dice)) <= rate);
      else surv = alive;      // no dice <=> every node of the level is a-priori with survival >= 1 (host rule): u <= rate always

      // ---- (6) State_Daily: lane p owns pair p ---------------------------------------------------------------------
      if constexpr (PK) {
        const LevelTally LT = TP::tally(k);          // LT.same: nobody can die here, the survivors are the arrivals
        if (LT.kind == 1) {                          // head of a run of passive one-node levels
          W[LT.wc] += prev_alive ? (1u << LT.bc) : 0u;
        } else if (LT.kind == 2) {
          W[LT.wc] += prev_alive ? (1u << LT.bc) : 0u;
          W[LT.ws] += surv ? (1u << LT.bs) : 0u;
        } else if (LT.kind == 3) {
          const uint32_t i6 = pk >> 27;                                      // 6 * local index of this lane's node
          W[LT.wc] += prev_alive ? ((1u << LT.bc) << i6) : 0u;
          if (!LT.same) W[LT.ws] += surv ? ((1u << LT.bs) << i6) : 0u;
        } else if (LT.kind == 4) {
          const uint32_t one = 1u << (pk >> 27);
          const int wi = (int)((pk >> 24) & 7u);
#pragma unroll
          for (int j = 0; j < LT.nwl; ++j) {
            W[LT.wc + j] += (prev_alive && wi == j) ? one : 0u;
            if (!LT.same) W[LT.ws + j] += (surv && wi == j) ? one : 0u;
          }
        }
      } else {
        const int key_c = prev_alive ? loc : 0x1ff, key_s = surv ? loc : 0x1ff;
#pragma unroll(TP::UNROLL_P)
        for (int p = mv.level_begin; p < mv.level_end; ++p) {                // uniform bounds
          const int node = TP::pair_node(FT, p);
          const uint32_t pc = __popc(__ballot_sync(0xffffffffu, key_c == node));
          const uint32_t ps = __popc(__ballot_sync(0xffffffffu, key_s == node));
#pragma unroll
          for (int r = 0; r < PR; ++r) add_if_lane(acc_cnt[r], acc_suc[r], p, lane_r[r], pc, ps);
        }
      }
      if (mv.need & NEED_UCHECK) {                       // first 'U' state in lexicographic column order (Appendix E5)
        const bool newU = ((pk >> 3) & 1u) && active && (int)mv.lex < best_rank;
        best_rank = newU ? (int)mv.lex : best_rank;
        ent_surv = newU ? surv : ent_surv;
      }

      // ---- (3) movement --------------------------------------------------------------------------------------------
      if (k < n_moves - 1) {
        int j = 0;
        if (mv.need & NEED_ROUTE) {
          const float u = u23<float>(word(mv.slot_route));
          // searchsorted(cdf, u, 'right') = number of thresholds <= u; rows are whole float4s padded with 2.0
          const float4* rowq = reinterpret_cast<const float4*>(s_cdf) + at * (crow >> 2);
          const int groups = ((int)mv.max_deg - 1 + 3) >> 2;
#pragma unroll(TP::UNROLL_P)
          for (int g = 0; g < groups; ++g) {
            const float4 q = rowq[g];
            j -= le_mask(q.x, u) + le_mask(q.y, u) + le_mask(q.z, u) + le_mask(q.w, u);
          }
        }
        STRYKE_BCHK(P, j >= 0 && j < stride, 4);
        loc = surv ? s_dst[at * stride + j] : loc;
        STRYKE_BCHK(P, loc >= 0 && loc < n_nodes, 5);
      }
      alive = surv;
    }
    const bool ent = active && best_rank < 255;
    if constexpr (PK) {
      W[TP::W_ES] += ent ? (ent_surv ? (0x41u << TP::B_ES) : (1u << TP::B_ES)) : 0u;
      W[TP::W_CAUSE] += cz;          // (the caller counts the tiles held in W and drains them, see simulate_run)
    } else {
      const uint32_t n_ent = __popc(__ballot_sync(0xffffffffu, ent));
      const uint32_t n_sur = __popc(__ballot_sync(0xffffffffu, ent && ent_surv));
      add_if_lane1(acc_daily, 0, lane, n_ent);
      add_if_lane1(acc_daily, 1, lane, n_sur);
      if (__ballot_sync(0xffffffffu, cz != 0)) {
        add_if_lane1(acc_daily, 2, lane, __reduce_add_sync(0xffffffffu, cz & 0xffu));
        add_if_lane1(acc_daily, 3, lane, __reduce_add_sync(0xffffffffu, (cz >> 8) & 0xffu));
        add_if_lane1(acc_daily, 4, lane, __reduce_add_sync(0xffffffffu, (cz >> 16) & 0xffu));
      }
    }
  };

  // `run` consecutive tiles of one cell with nothing to decide between them: lane's fish `fish`, `fish + 32`, ...; a fish is
  // simulated for real while its index is below n_lim.  The trip counts are made provably warp-uniform (shuffle), so the
  // loop around the ~170-instruction body is a counter and a branch — no reconvergence barriers, no per-tile tests of
  // "end of cell? end of chunk? time to fetch?" (measured on C2 before: 27 of 203 executed instructions per tile and 18 %
  // of the stall samples were that bookkeeping).  The packed counters are drained every HOLD tiles.
  auto simulate_run = [&](int run, int fish, const int n_lim, const uint32_t cid_lo, const uint32_t cid_hi) {
    while (run > 0) {
      int r = run;
      if constexpr (PK) r = r < TP::HOLD - held ? r : TP::HOLD - held;
      r = __shfl_sync(0xffffffffu, r, 0);
#pragma unroll 1
      for (int i = 0; i < r; ++i, fish += 32) simulate(fish, fish < n_lim, cid_lo, cid_hi);
      run -= r;
      if constexpr (PK) { held += r; if (held == TP::HOLD) drain(); }
    }
  };

  const int64_t T = TP::POOL ? P.blk_tile[(P.n_cells + 31) >> 5] : P.tile_off[P.n_cells];
  // Work distribution.  The SM's warp arbiter does not share issue slots evenly, so equal static ranges finish unevenly and
  // leave the SM short of warps at the end; tiles are therefore handed out dynamically.
  // Guided self-scheduling: a warp takes (tiles not yet handed out) / (2 x warps of the grid) consecutive tiles from one
  // global counter, at most P.chunk_tiles and at least P.chunk_min: long chunks while there is plenty of work (one cell
  // search per chunk), short ones at the end so that all warps finish together.  The next chunk is fetched while the
  // current one is being simulated.
  // The next chunk is fetched when the current one has PREFETCH_TILES tiles left to hand out: late enough that its size is
  // decided on the work that is REALLY left (a chunk fetched a whole chunk ahead is one round too large, and the kernel's
  // tail is as long as the largest chunk: measured 1.0-1.4 chunk durations per launch), early enough that the atomic's
  // latency hides under the last tiles.
  constexpr int PREFETCH_TILES = 6;
  const int64_t G = (int64_t)gridDim.x * WARPS;
  unsigned long long tl_start = 0, tl_chunks = 0, tl_tiles = 0;
  if (P.timeline) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_start));
  int64_t nx0 = 0, nx1 = 0;                           // the chunk fetched ahead: tiles [nx0, nx1)
  long long my_done = 0;                              // tiles this warp has taken so far
  auto fetch_chunk = [&]() {
    unsigned long long start = 0, size = 0;
    if (lane == 0) {
      const long long seen = (long long)*reinterpret_cast<volatile unsigned long long*>(P.work_counter);
      long long sz = (T - seen) / (2 * G);
      // Warps do not run at one speed: the SM's arbiter favours the older thread blocks, the youngest third of the grid gets
      // about half the issue slots (measured: 372 .. 2104 tiles per warp in one launch).  A chunk sized for the average
      // warp keeps a slow one busy long after the queue is empty (the kernel's tail: ~100 us per launch), so the guided
      // size is scaled by this warp's own share of the work so far, my_done / (seen / G), between 1/4 and 3/2.
      long long cap = P.chunk_tiles;
      if (seen > 4 * G && T < 8192 * G) {      // (long launches: the tail is noise, keep the chunk starts few)
        float r = (float)my_done * (float)G / (float)seen;
        r = r < 0.25f ? 0.25f : r > 1.5f ? 1.5f : r;
        sz = (long long)((float)sz * r);
        if (r < 1.f) cap = (long long)((float)cap * r);      // a starved warp left alone at the end finishes its chunk at single-warp speed
      }
      sz = sz > cap ? cap : sz;
      sz = sz < (long long)P.chunk_min ? (long long)P.chunk_min : sz;
      size = (unsigned long long)sz;
      start = atomicAdd(P.work_counter, size);
      my_done += sz;
    }
    nx0 = (int64_t)__shfl_sync(0xffffffffu, start, 0);
    nx1 = nx0 + (int64_t)__shfl_sync(0xffffffffu, size, 0);
  };
  fetch_chunk();
  int64_t cur_cell = -1;
  bool cur_owned = false;                             // the current cell is an owned one (pool_big): its row is assembled in s_row
  bool pool_dirty = false;                            // the accumulators of the current block hold something
  constexpr int RWP = pool_rwp(TP::kStatic ? TP::N_PAIRS : 0);
  uint32_t* s_row = TP::OWN ? reinterpret_cast<uint32_t*>(wbase + SL.pool_row) : nullptr;      // [32 cells][RWP]
  auto leave_cell = [&]() {                           // packed counters -> owners -> global tallies of the current cell
    if (cur_cell >= 0) {
      if (PK && held) { if (held == 1) drain_one_tile(); else drain(); }
      if (TP::OWN && cur_owned) {                     // into the cell's row in shared memory (narrow layout, see tally_flush)
        uint32_t* row = s_row + (int)(cur_cell & 31) * RWP;
#pragma unroll
        for (int r = 0; r < PR; ++r) {
          const int p = r * 32 + lane;
          if (p < n_pairs) row[3 + p] += acc_suc[r] | (acc_cnt[r] << 16);
          acc_suc[r] = 0; acc_cnt[r] = 0;
        }
        const uint32_t up = __shfl_down_sync(0xffffffffu, acc_daily, 1);
        const uint32_t w = acc_daily | (up << 16);
        if (lane == 0) row[1] += w;                     // num_entrained | num_survived << 16
        if (lane == 2) row[2] += w;                     // impingement | blade strike << 16
        if (lane == 4) row[0] += acc_daily << 16;       // (pop_size) | barotrauma << 16
        acc_daily = 0;
        pool_dirty = true;
      } else {
        flush(cur_cell);
      }
      cur_cell = -1;
    }
  };

  if constexpr (!TP::POOL) {
  int64_t c = -1, c_begin = 0, c_end = 0;             // current cell and its tile range [c_begin, c_end)
  int n_c = 0;
  uint32_t cid_lo = 0, cid_hi = 0;
  for (;;) {
  const int64_t t0 = nx0;
  if (t0 >= T) break;
  const int64_t t1 = nx1 < T ? nx1 : T;
  tl_chunks += 1; tl_tiles += (unsigned long long)(t1 - t0);
  bool fetched = false;     // the next chunk is fetched near the END of this one (see PREFETCH_TILES)
  bool fresh = false;       // the tile about to be simulated is the first one this warp sees of cell c
  if (t0 >= c_end) {        // not a continuation: last cell with tile_off[c] <= t0 (cells without fish have empty ranges)
    c = warp_search_last_le(P.tile_off, c < 0 ? 0 : c, P.n_cells, t0, lane);
    c_begin = P.tile_off[c]; c_end = P.tile_off[c + 1];
    fresh = true;
  }
  // 32-bit bookkeeping inside the chunk: tiles left in the current cell, index of the next tile's first fish, tiles of
  // the chunk still to simulate
  const int64_t room = c_end - t0;
  int left = room < (int64_t)(1 << 30) ? (int)room : (1 << 30);
  int fish_base = (int)(t0 - c_begin) * 32;
  int todo = (int)(t1 - t0);
  while (todo > 0) {
    if (left == 0) {
      do { ++c; c_begin = c_end; c_end = P.tile_off[c + 1]; } while (c_end == c_begin);
      const int64_t rm = c_end - c_begin;
      left = rm < (int64_t)(1 << 30) ? (int)rm : (1 << 30);
      fish_base = 0;
      fresh = true;
    }
    if (fresh) {
      fresh = false;
      leave_cell();
      cur_cell = c;
      STRYKE_BCHK(P, c >= 0 && c < P.n_cells, 6);
      n_c = P.cell_n[c];
      const CellInfo ci = cell_info(P, c);
      cid_lo = (uint32_t)ci.id; cid_hi = (uint32_t)(ci.id >> 32);
      load_params(ci.species, ci.day);
    }
    int run = left < todo ? left : todo;               // up to the end of the cell or of the chunk ...
    if (!fetched) {                                    // ... or to the point where the next chunk is fetched
      if (todo <= PREFETCH_TILES) { fetch_chunk(); fetched = true; }
      else run = run < todo - PREFETCH_TILES ? run : todo - PREFETCH_TILES;
    }
    STRYKE_BCHK(P, fish_base >= 0 && fish_base < n_c, 7);     // a tile always holds at least one fish of its cell
    simulate_run(run, fish_base + lane, n_c, cid_lo, cid_hi);
    fish_base += run << 5;
    left -= run;
    todo -= run;
  }
  if (!fetched) fetch_chunk();
  }
  } else {
  // ---- pooled remainders (see pool_block): blocks of 32 cells; the tiles of block b are its shared tiles, then the
  // full tiles of its cells in order.  blk_tile[b] is the block's first tile.  A shared tile's per-lane counters are
  // added to per-cell accumulator words in shared memory (whole words: a cell has at most 31 pooled fish, no field can
  // carry); when the warp leaves the block the accumulators go to global memory, one lane per (cell, counter).
  uint4* s_blk = reinterpret_cast<uint4*>(wbase + SL.pool_blk);
  long long* s_ef = reinterpret_cast<long long*>(wbase + SL.pool_ef);
  int* s_rank = reinterpret_cast<int*>(wbase + SL.pool_rank);
  int2* s_sd = reinterpret_cast<int2*>(wbase + SL.pool_sd);
  int* s_full = reinterpret_cast<int*>(wbase + SL.pool_full);           // full tiles of every cell of the block
  uint32_t* s_acc = reinterpret_cast<uint32_t*>(wbase + SL.pool_seg);        // [32 cells][NW]
  uint32_t* s_field = reinterpret_cast<uint32_t*>(smem + SL.pool_field);     // [2 n_pairs + 5]: word [0:5) shift [5:10) 10-bit [10] exists [11]
  const int NI = 2 * n_pairs + STRYKE_DAILY_W;
  if (wid == 0) {
#pragma unroll
    for (int r = 0; r < PR; ++r) {
      const int p = r * 32 + lane;
      if (p < n_pairs) {
        const uint32_t ex = ((lcfg[r] >> 20) & 1u) << 11;
        s_field[2 * p] = ((lcfg[r] >> 10) & 0x3ffu) | ex;                   // survivors
        s_field[2 * p + 1] = (lcfg[r] & 0x3ffu) | ex;                       // arrivals
      }
    }
    if (lane < STRYKE_DAILY_W) s_field[2 * n_pairs + lane] = dcfg & 0xfffu;
  }
#pragma unroll
  for (int q = 0; q < TP::NW; ++q) s_acc[lane * TP::NW + q] = 0u;
  __syncthreads();
  const int64_t NB = (P.n_cells + 31) >> 5;
  auto blk_off = [&](int64_t b) -> int64_t { return P.blk_tile[b]; };
  const uint32_t below = (2u << lane) - 1u;           // bits 0 .. lane
  int64_t b = -1, b_begin = 0, b_end = 0;             // current block and its tile range
  int S_b = 0, nrem = 0;                              // its shared tiles, its cells with a remainder
  uint32_t cur_rmask = 0, cur_bigm = 0xffffffffu;     // bit j: cell j of the block has pooled fish / is big
  int64_t a_end = 0;                                  // end of the block's atomic part
  bool own = false;                                   // this warp is simulating the atomic part of block b
  const int nmax_big = P.out.rows != nullptr ? (int)P.out.narrow_max : 0x7fffffff;
  if constexpr (TP::OWN) {
    for (int i = lane; i < 32 * RWP; i += 32) s_row[i] = 0u;
  }
  static_assert(sizeof(GroupCursor) <= 64, "GroupCursor slot");
  GroupCursor* s_gcur = reinterpret_cast<GroupCursor*>(wbase + SL.pool_gcur);
  *s_gcur = GroupCursor{};                            // (base = end = 0: no group yet)
  __syncwarp();
  auto pool_flush = [&]() {                           // accumulators of block b -> global tallies (after leave_cell: an owned
    if (pool_dirty) {                                 // cell's last tiles go into its row first)
      pool_dirty = false;
      __syncwarp();
      if (P.out.rows != nullptr) {
        // compact rows, transposed: lane <-> cell of the block.  Every lane unpacks the packed fields of ITS cell's
        // accumulator words (word and shift of every pair are compile-time constants of the facility) into the row's 16-bit
        // pairs.  An owned cell's row is complete now (pooled fish here, full tiles in s_row): it is written once, with
        // plain stores.  A big cell's pooled fish are added to its row (red.global: other warps add its full tiles); cells
        // above narrow_max take the 32-bit layout.
        const uint32_t bm = P.out.blk_mask[b], bw = P.out.blk_wide[b];
        const uint32_t lower = (1u << lane) - 1u;
        const unsigned int slot = P.out.blk_row[b] + __popc(bm & lower) + __popc(bw & lower);
        const bool wide = (bw >> lane) & 1u;
        const bool big = (cur_bigm >> lane) & 1u;
        const bool owned = TP::OWN && ((bm >> lane) & 1u) && !big;
        if ((owned || ((cur_rmask >> lane) & 1u)) && slot + (wide ? 2u : 1u) <= P.out.cap) {
          uint32_t A[TP::NW > 0 ? TP::NW : 1];
#pragma unroll
          for (int q = 0; q < TP::NW; ++q) A[q] = s_acc[lane * TP::NW + q];
          uint32_t* row = P.out.rows + (size_t)slot * P.out.row_w;
          uint32_t dv[STRYKE_DAILY_W];
#pragma unroll
          for (int j = 0; j < STRYKE_DAILY_W; ++j) {
            const uint32_t cf = TP::daily_cfg(j);
            dv[j] = ((cf >> 11) & 1u) ? ((A[cf & 31u] >> ((cf >> 5) & 31u)) & (((cf >> 10) & 1u) ? 1023u : 63u)) : 0u;
          }
          if (owned) {
            const uint32_t* sr = s_row + lane * RWP;
            row[0] = s_blk[lane].x + (dv[4] << 16) + sr[0];
            row[1] = (dv[0] | (dv[1] << 16)) + sr[1];
            row[2] = (dv[2] | (dv[3] << 16)) + sr[2];
#pragma unroll
            for (int p = 0; p < (TP::kStatic ? TP::N_PAIRS : 0); ++p) {
              const uint32_t cf = TP::lane_cfg(p);
              uint32_t v = 0u;
              if ((cf >> 20) & 1u) {
                const uint32_t cnt = (A[cf & 31u] >> ((cf >> 5) & 31u)) & 63u, suc = (A[(cf >> 10) & 31u] >> ((cf >> 15) & 31u)) & 63u;
                v = suc | (cnt << 16);
              }
              row[3 + p] = v + sr[3 + p];
            }
          } else if (!wide) {
            const uint32_t w1 = dv[0] | (dv[1] << 16), w2 = dv[2] | (dv[3] << 16);
            if (w1) atomicAdd(row + 1, w1);
            if (w2) atomicAdd(row + 2, w2);
            if (dv[4]) atomicAdd(row, dv[4] << 16);
#pragma unroll
            for (int p = 0; p < (TP::kStatic ? TP::N_PAIRS : 0); ++p) {
              const uint32_t cf = TP::lane_cfg(p);
              if (!((cf >> 20) & 1u)) continue;
              const uint32_t cnt = (A[cf & 31u] >> ((cf >> 5) & 31u)) & 63u, suc = (A[(cf >> 10) & 31u] >> ((cf >> 15) & 31u)) & 63u;
              const uint32_t v = suc | (cnt << 16);
              if (v) atomicAdd(row + 3 + p, v);
            }
          } else {
#pragma unroll
            for (int j = 0; j < STRYKE_DAILY_W; ++j)
              if (dv[j]) atomicAdd(row + 1 + j, dv[j]);
#pragma unroll
            for (int p = 0; p < (TP::kStatic ? TP::N_PAIRS : 0); ++p) {
              const uint32_t cf = TP::lane_cfg(p);
              if (!((cf >> 20) & 1u)) continue;
              const uint32_t cnt = (A[cf & 31u] >> ((cf >> 5) & 31u)) & 63u, suc = (A[(cf >> 10) & 31u] >> ((cf >> 15) & 31u)) & 63u;
              if (suc) atomicAdd(row + 6 + 2 * p, suc);
              if (cnt) atomicAdd(row + 7 + 2 * p, cnt);
            }
          }
        }
        if constexpr (TP::OWN) {
          __syncwarp();
#pragma unroll
          for (int w = 0; w < RWP; ++w) s_row[lane * RWP + w] = 0u;
        }
      } else {
      for (int i0 = 0; i0 < NI; i0 += 32) {           // lane <-> counter i0 + lane, loop over the cells with a remainder
        const int i = i0 + lane;
        const uint32_t cf = i < NI ? s_field[i] : 0u;
        const int wq = (int)(cf & 31u);
        const uint32_t sh = (cf >> 5) & 31u;
        const uint32_t mask = ((cf >> 11) & 1u) ? (((cf >> 10) & 1u) ? 1023u : 63u) : 0u;
        const bool state = i < 2 * n_pairs;
        uint32_t* base = state ? P.state_daily + i : P.daily + (i - 2 * n_pairs);
        const uint32_t per_cell = state ? (uint32_t)(2 * n_pairs) : (uint32_t)STRYKE_DAILY_W;
        base += (size_t)(b << 5) * per_cell;
        for (int rho = 0; rho < nrem; ++rho) {
          const int jl = s_rank[rho];
          const uint32_t v = (s_acc[jl * TP::NW + wq] >> sh) & mask;
          STRYKE_BCHK(P, (b << 5) + jl < P.n_cells, 12);
          if (v) atomicAdd(base + (uint32_t)jl * per_cell, v);
        }
      }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < TP::NW; ++q) s_acc[lane * TP::NW + q] = 0u;
      __syncwarp();
    }
  };
  auto clamp30 = [](long long x) -> int { return x < (long long)(1 << 30) ? (int)x : (1 << 30); };
  // One loop, one decision per run: what is the next tile (a new chunk? next block? a shared tile? full tiles of which
  // cell?), then `run` tiles with nothing to decide between them.  (One loop rather than chunks x runs so that leave_cell
  // and pool_flush, with the drains and flushes inlined into them, exist once in the code: the kernel is I-cache bound
  // outside its inner loop.)  A chunk is [t0, t1) — except that the atomic part of a block (pool_big) goes, whole, to the
  // warp whose chunk holds the block's first tile: t_lim moves past t1 to the end of an atomic part this warp has started,
  // and an atomic part that began before t0 is skipped (its owner runs or has run past its own t1).
  int64_t t = 0, t_lim = 0;                           // next tile; first tile this warp does not simulate
  bool fetched = true;                                // nx0 / nx1 hold a chunk that has not been taken yet
  bool new_block = false;
  for (;;) {
    bool done = false, search = false;
    if (t >= t_lim) {                                 // the chunk is finished: take the next one
      if (!fetched) fetch_chunk();
      const int64_t t0 = nx0;
      if (t0 >= T) {
        done = true;
      } else {
        t = t0; t_lim = nx1 < T ? nx1 : T;
        tl_chunks += 1; tl_tiles += (unsigned long long)(t_lim - t0);
        fetched = false;                              // its successor is fetched near the END of this one (see PREFETCH_TILES)
        own = false;
        search = t0 < b_begin || t0 >= b_end;         // (else: still inside the block where the last chunk ended)
      }
    }
    leave_cell();
    if (done || search || t >= b_end) {
      pool_flush();
      if (done) break;
      if (search) {      // last block with blk_off <= t (blocks without fish have empty ranges; chunks are handed out in increasing order)
        b = warp_search_last_le(P.blk_tile, (b >= 0 && t >= b_end) ? b : 0, NB, t, lane);
        b_begin = blk_off(b); b_end = blk_off(b + 1);
      } else {
        do { ++b; b_begin = b_end; b_end = blk_off(b + 1); } while (b_end == b_begin);
      }
      new_block = true;
    }
    if (new_block) {
      new_block = false;
      own = false;
      STRYKE_BCHK(P, b >= 0 && b < NB, 8);
      const int64_t ci = (b << 5) + lane;
      const bool in = ci < P.n_cells;
      const int n = in ? P.cell_n[ci] : 0;
      CellInfo cinf{0, 0, 0ull};
      GroupCursor gcur{};
      if (P.groups != nullptr) {      // warp-uniform: the group of the block's first cell — nearly always the last block's group
        gcur = *s_gcur;
        const int64_t c0 = b << 5;
        if (c0 < gcur.base || c0 >= gcur.end || c0 - gcur.base > 0x3fffff00ll) {
          gcur = group_cursor(P, c0);
          __syncwarp();
          *s_gcur = gcur;
          __syncwarp();
        }
      }
      if (in) cinf = cell_at(P, gcur, ci);
      const uint64_t id = cinf.id;
      const uint32_t info = P.blk_S[b];
      S_b = (int)(info & POOL_S_MASK);
      const int owned_full = (int)((info >> POOL_OWNED_SHIFT) & POOL_OWNED_MASK);
      a_end = b_begin + S_b + owned_full;
      const int r = pool_rem(n, P.pool_all), f = pool_full(n, P.pool_all);
      const bool big = pool_big(n, P.pool_all, P.big_full, nmax_big);
      const int pr = (info & POOL_MIXED) ? ((r + 31) & ~31) : r;   // mixed block: every remainder is padded to whole tiles of its own
      int sr = pr, so = big ? 0 : f;                      // prefix sums: pooled fish, full tiles of the owned cells ...
      long long sb = big ? f : 0;                         // ... and of the big ones, which come after them
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tr = __shfl_up_sync(0xffffffffu, sr, o), to = __shfl_up_sync(0xffffffffu, so, o);
        const long long tb = __shfl_up_sync(0xffffffffu, sb, o);
        if (lane >= o) { sr += tr; so += to; sb += tb; }
      }
      const uint32_t rmask = __ballot_sync(0xffffffffu, r > 0);      // cells with pooled fish
      nrem = __popc(rmask);
      cur_rmask = rmask;
      cur_bigm = __ballot_sync(0xffffffffu, big);
      STRYKE_BCHK(P, __shfl_sync(0xffffffffu, so, 31) == owned_full, 14);
      __syncwarp();
      s_blk[lane] = make_uint4((uint32_t)n, (uint32_t)id, (uint32_t)(id >> 32), (uint32_t)(sr - pr));
      s_ef[lane] = big ? (long long)owned_full + (sb - f) : (long long)(so - f);      // first full tile of the cell within the block's full tiles
      s_full[lane] = f;
      s_sd[lane] = make_int2(cinf.species, cinf.day);
      if (r > 0) s_rank[__popc(rmask & (below >> 1))] = lane;
      __syncwarp();
    }
    if (t < a_end && !own) {
      if (t == b_begin) { own = true; t_lim = t_lim > a_end ? t_lim : a_end; }
      else { t = a_end; continue; }
    }
    if (t == b_begin && S_b > 0) {
      // The block's shared tiles (this warp owns them: they open the atomic part), all in one go.  Shared tile js holds
      // positions [32 js, 32 js + 32) of the block's concatenated remainders.  M: cells whose remainder starts in this tile
      // (bit = lane of its first fish); jl: lane (within the block) of the cell of this lane's fish.  The tile's per-lane
      // counters go, whole words, into the per-cell accumulators.
      const uint4 mine = s_blk[lane];
      int rel = pool_rem((int)mine.x, P.pool_all) != 0 ? (int)mine.w : 0x40000000;   // first pooled fish of MY cell, relative to the tile
#pragma unroll 1
      for (int js = 0; js < S_b; ++js, rel -= 32) {
        const uint32_t M = __reduce_or_sync(0xffffffffu, (unsigned)rel < 32u ? (1u << rel) : 0u);
        const int K0 = __popc(__ballot_sync(0xffffffffu, rel < 0));            // remainders that start before the tile
        const int rank = K0 + __popc(M & below) - 1;
        STRYKE_BCHK(P, rank >= 0 && rank < nrem, 9);
        const int jl = s_rank[rank];
        STRYKE_BCHK(P, jl >= 0 && jl < 32, 10);
        const uint4 cellrec = s_blk[jl];
        const int n = (int)cellrec.x;
        const int own_fish = pool_full(n, P.pool_all) << 5;        // fish of the cell's own full tiles come first
        const int fish = own_fish + ((js << 5) + lane - (int)cellrec.w);
        STRYKE_BCHK(P, fish >= own_fish && (lane != 0 || fish < n), 11);
        const int2 sd = s_sd[__shfl_sync(0xffffffffu, jl, 0)];
        load_params(sd.x, sd.y);
        simulate(fish, fish < n, cellrec.y, cellrec.z);
        uint32_t* row = s_acc + jl * TP::NW;
#pragma unroll
        for (int q = 0; q < TP::NW; ++q) {
          if (W[q]) atomicAdd(row + q, W[q]);
          W[q] = 0;
        }
      }
      held = 0;
      pool_dirty = true;
      t += S_b;
      if (!fetched && t_lim - t <= PREFETCH_TILES) { fetch_chunk(); fetched = true; }
      continue;
    }
    // full tile jj of the block: which cell?  Its tiles up to the end of the cell (or of what this warp may take) are one run.
    const long long jj = t - b_begin - S_b;
    const long long ef = s_ef[lane];
    const int f = s_full[lane];
    const int sel = __ffs(__ballot_sync(0xffffffffu, f > 0 && jj >= ef && jj < ef + f)) - 1;
    STRYKE_BCHK(P, sel >= 0 && sel < 32, 13);
    const long long ef_s = s_ef[sel];
    const uint4 cellrec = s_blk[sel];
    int run = __shfl_sync(0xffffffffu, clamp30(ef_s + (long long)s_full[sel] - jj), 0);
    const int room = clamp30(t_lim - t);
    run = run < room ? run : room;
    const int fish = __shfl_sync(0xffffffffu, (int)(jj - ef_s) * 32, 0) + lane;
    cur_cell = (b << 5) + sel;
    cur_owned = TP::OWN && P.out.rows != nullptr && !((cur_bigm >> sel) & 1u);
    STRYKE_BCHK(P, !cur_owned || (own && t + run <= a_end), 15);
    const int2 sd = s_sd[sel];
    load_params(sd.x, sd.y);
    t += run;
    if (!fetched && t_lim - t <= PREFETCH_TILES) { fetch_chunk(); fetched = true; }
    simulate_run(run, fish, 0x7fffffff, cellrec.y, cellrec.z);      // (full tiles: every lane has a fish)
  }
  }
  if constexpr (!TP::POOL) leave_cell();
  if (P.timeline && lane == 0) {
    unsigned long long tl_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_end));
    unsigned long long* rec = P.timeline + 4 * ((size_t)blockIdx.x * WARPS + wid);
    rec[0] = tl_start; rec[1] = tl_end; rec[2] = tl_chunks; rec[3] = tl_tiles;
  }
}

#ifndef STRYKE_NVRTC
// ahead-of-time instantiation: tables from the constant bank
template <int PR, bool BARO, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) passage_fast_kernel(const KParams P, const __grid_constant__ FastTables FT,
                                                                const uint2* __restrict__ g_nodes, const int stride,
                                                                const int n_species_staged) {
  passage_fast_body<RuntimeTables, PR, BARO>(P, FT, g_nodes, stride, n_species_staged);
}
#endif

}  // namespace stryke
