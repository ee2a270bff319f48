/* stryke_b200.h — C ABI of the B200-native passage-simulation engine.
 *
 * Drop-in boundary for the hot path of knebiolo/stryke's `simulation.run()`:
 * everything the reference does per (scenario, species, iteration, day) cell between the presence roll and the
 * `Daily` / `State_Daily` rows (reference: Stryke/stryke.py:3025-3383) happens behind
 * `stryke_b200_run` (host buffers in, host tallies out) or the plan API (device-resident tables, caller-owned
 * device tally buffers, caller's stream).  Plain pointers and sizes only; no exceptions cross the ABI; every entry
 * point returns 0 or a negative STRYKE_E_* code and `stryke_b200_last_error()` holds the message (thread-local).
 *
 * All file:line citations are relative to the reference tree (/root/reference).
 */
#ifndef STRYKE_B200_H
#define STRYKE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STRYKE_B200_ABI_VERSION 2

/* error codes */
#define STRYKE_OK            0
#define STRYKE_E_ARG        -1   /* invalid argument / table shape */
#define STRYKE_E_CUDA       -2   /* CUDA runtime error (message in last_error) */
#define STRYKE_E_LIMIT      -3   /* facility exceeds a compiled limit (nodes, units, edges, moves, pairs) */
#define STRYKE_E_MAX_FISH   -4   /* a cell's population exceeded max_daily_fish (stryke.py:2610-2616, :3052) */
#define STRYKE_E_NOGPU      -5   /* no CUDA device: there is no CPU fallback */

/* node kinds: `Surv_Fun` of the Nodes table (stryke.py:1399-1448) */
#define STRYKE_KIND_APRIORI   0
#define STRYKE_KIND_KAPLAN    1  /* stryke.py:846  */
#define STRYKE_KIND_PROPELLER 2  /* stryke.py:924  */
#define STRYKE_KIND_FRANCIS   3  /* stryke.py:1011 */
#define STRYKE_KIND_PUMP      4  /* stryke.py:1096 */

/* entrainment-count kinds: `dist` of the Population table (stryke.py:2550-2565) */
#define STRYKE_ENT_FIXED      0  /* anadromous mode, n = int(Fish)  (stryke.py:3028-3029) */
#define STRYKE_ENT_PARETO     1
#define STRYKE_ENT_LOGNORMAL  2
#define STRYKE_ENT_WEIBULL    3
#define STRYKE_ENT_EXTREME    4

/* compiled limits */
#define STRYKE_MAX_NODES  255
#define STRYKE_MAX_UNITS   64
#define STRYKE_MAX_EDGES 1024
#define STRYKE_MAX_MOVES   64
#define STRYKE_MAX_PAIRS  128

#define STRYKE_UNIT_STATIC_W 4   /* rack_spacing_ft, intake_vel, fb_depth_ft, tail_depth_ft */
#define STRYKE_UNIT_COEF_W   8   /* see StrykeDayTables.unit_coef */
#define STRYKE_SPECIES_W    20   /* see StrykeSpeciesTable.params */
#define STRYKE_DAILY_W       5
#define STRYKE_RANGE_ALIGN 1024   /* ranges of cells of plan_launch_compact start at multiples of this */   /* num_entrained, num_survived, mortality_impingement, _blade_strike, _barotrauma */

/* Static description of the facility DAG (reference: create_route stryke.py:1581-1655, Nodes/Edges/Unit tables). */
typedef struct StrykeFacility {
  int32_t n_nodes, n_units, n_moves, n_edges, n_pairs, start_node;
  int32_t barotrauma;            /* barotrauma_required (stryke.py:2672-2711) */
  const uint8_t* node_kind;      /* [n_nodes] STRYKE_KIND_*                                               */
  const double*  node_apriori;   /* [n_nodes] `Survival` for a-priori nodes (rounded to fp32 on device, :1575) */
  const int16_t* node_unit;      /* [n_nodes] unit index or -1                                            */
  const uint8_t* node_has_U;     /* [n_nodes] node NAME contains 'U' (entrainment rule, stryke.py:3280)   */
  const int32_t* edge_off;       /* [n_nodes+1] CSR offsets into edge_dst / edge_cdf, successors in ROUTING order */
  const int16_t* edge_dst;       /* [n_edges]                                                             */
  const uint8_t* move_lex_rank;  /* [n_moves] rank of "state_k" among sorted column names (stryke.py:3272) */
  const int32_t* level_off;      /* [n_moves+1] reachable (move, node) pairs of move k: pair_node[level_off[k] .. level_off[k+1]) */
  const int16_t* pair_node;      /* [n_pairs]                                                             */
  const double*  unit_static;    /* [n_units][STRYKE_UNIT_STATIC_W]                                       */
} StrykeFacility;

/* Per-(scenario, day) tables built on the host from the day's discharge
 * (reference: compute_unit_flow_allocations stryke.py:1657-1821, movement stryke.py:1857-2035, u_param_dict[u]['Q'] :2977).
 * unit_coef[d][u][.]: 0 lambda, 1 N, 2 D,
 *    Kaplan:              3 pi*ada*Ewd, 4 2*Qwd, 5 8*Qwd          (strike = lambda*(N*L/D)*(cos a/(8Qwd) + sin a/(pi rR)), a = atan(c3/(c4*rR)))
 *    Propeller/Francis/Pump: 3 X  with p_strike = lambda*(N*L/D)*X (Pump: lambda slot holds gamma, D slot holds 0.707*D2)
 *    6 = 1 if the unit has flow today (Q > 0)
 *    7 = barotrauma, friction-loss pressure mode: the day's offset c0 of P1/P2 = c0 + c1 * fish_depth_ft, with
 *        c1 = rho g 0.3048 / P2 and P2 = atm + rho g 0.3048 * unit_static[3] (Darcy-Weisbach head loss of the unit's discharge,
 *        Stryke/barotrauma.py:43-137 composed as Stryke/stryke_barotrauma.py:565-601); 0 = hydrostatic pressures (stryke.py:1496-1503) */
typedef struct StrykeDayTables {
  int32_t n_days;
  const double*  curr_Q;     /* [n_days]                                                                  */
  const double*  edge_cdf;   /* [n_days][n_edges] cdf = cumsum(p)/cumsum(p)[-1] inside each node's segment (np.random.choice) */
  const uint8_t* node_stay;  /* [n_days][n_nodes] 1 = a live fish stays put (stryke.py:1847-1855, :2018)   */
  const double*  unit_coef;  /* [n_days][n_units][STRYKE_UNIT_COEF_W]                                     */
} StrykeDayTables;

/* Population table rows (stryke.py:2858-2890).  params[s][.]:
 *  0 length shape s, 1 length location, 2 length scale (cm), 3 Length_mean (in), 4 Length_sd, 5 length mode (0 lognormal :3070, 1 normal :3090),
 *  6 width ratio (stryke.py:1260-1294), 7 U_crit, 8 d_1, 9 d_2 (vertical habitat, :1475-1484), 10 beta_0, 11 beta_1,
 *  12 entrainment kind STRYKE_ENT_*, 13 shape, 14 location, 15 scale, 16 max_ent_rate (NaN = none), 17 occur_prob, 18 Fish, 19 reserved */
typedef struct StrykeSpeciesTable {
  int32_t n_species;
  const double* params;      /* [n_species][STRYKE_SPECIES_W] */
} StrykeSpeciesTable;

/* One (scenario, species) group of an implicit cell list: the reference's loops `for i in range(iterations)` x
 * `for day in flow_df` (stryke.py:2897-2902) for one Population row.  Cell cell0 + d * iterations + i is iteration i on the
 * group's d-th day with tot_hours > 0; that day is row day0 + active_days[active_off + d] of StrykeDayTables and the
 * cell's Philox id is (id_base + i) * id_days + active_days[active_off + d] + id_add. */
typedef struct StrykeCellGroup {
  int64_t  cell0;
  int32_t  iterations, n_active_days, day0, species, active_off, reserved;
  uint64_t id_base;
} StrykeCellGroup;

/* The cells to simulate: one entry per (scenario, species, iteration, day) with tot_hours > 0 — either three explicit
 * arrays, or (cell_day == NULL) the group table above: a year of 1000 iterations x 20 species x 5 scenarios is then a few
 * KB instead of 16 bytes per cell crossing PCIe on every call. */
typedef struct StrykeCells {
  int64_t n_cells;
  const int32_t*  cell_day;      /* [n_cells] index into StrykeDayTables                                   */
  const int32_t*  cell_species;  /* [n_cells] index into StrykeSpeciesTable                                */
  const uint64_t* cell_id;       /* [n_cells] logical id mixed into the Philox counter (partition-independent) */
  int32_t n_groups;              /* implicit list: groups ordered by cell0, covering [0, n_cells)           */
  int32_t n_active;              /* length of active_days                                                   */
  const StrykeCellGroup* groups;
  const int32_t* active_days;
  uint64_t id_days, id_add;
} StrykeCells;

/* Injected draws (bit-exact replay of the reference's numpy uniforms).  All pointers are
 * HOST pointers for stryke_b200_run / plan_create.  Fish of cell c occupy [fish_off[c], fish_off[c+1]) on the fish axis. */
typedef struct StrykeInjected {
  int64_t n_fish_total;
  const int64_t* fish_off;      /* [n_cells+1]                                                             */
  const double* presence;       /* [n_cells]  np.random.uniform(0,1)            stryke.py:3026             */
  const double* count_draw;     /* [n_cells]  PCG64 draw of population_sim      stryke.py:2559-2565 (NaN if fixed) */
  const double* length_z;       /* [n_fish]   PCG64 standard normals            stryke.py:3070             */
  const double* dice;           /* [n_moves][n_fish]                            stryke.py:3189             */
  const double* rR;             /* [n_moves][n_fish] raw uniform                stryke.py:900              */
  const double* depth;          /* [n_moves][n_fish] raw uniform                stryke.py:1494             */
  const double* cause;          /* [n_moves][4][n_fish]                         stryke.py:1545-1560        */
  const double* route;          /* [n_moves][n_fish]                            stryke.py:2058             */
  const double* ghost_rR;       /* [n_moves][n_cells] np.vectorize probe call on fish 0 (stryke.py:3197); may be NULL */
  const double* ghost_depth;    /* [n_moves][n_cells]                                                      */
  const double* ghost_cause;    /* [n_moves][4][n_cells]                                                   */
} StrykeInjected;

/* Optional per-fish trace (host pointers, may be NULL): the reference's `fishes` table columns (stryke.py:3105-3236). */
typedef struct StrykeTrace {
  float*    length_ft;     /* [n_fish]           `population` (fp32, stryke.py:3113)  */
  double*   length_ft64;   /* [n_fish]                                               */
  uint8_t*  state;         /* [n_moves][n_fish]  node index of state_k               */
  float*    rate;          /* [n_moves][n_fish]  rates_k (fp32, stryke.py:1575)      */
  uint8_t*  survival;      /* [n_moves][n_fish]  survival_k                          */
  double*   strike;        /* [n_moves][n_fish]  blade-strike survival where evaluated, NaN elsewhere */
  double*   baro;          /* [n_moves][n_fish]  barotrauma survival where evaluated, NaN elsewhere   */
  int64_t   n_fish;        /* capacity of the arrays.  0: the engine derives the fish count itself, which needs injected
                            * draws or fixed counts with occur_prob = 1.  > 0 (population mode): the caller learnt the
                            * count from an earlier run with the same seed; a different count is an error              */
} StrykeTrace;

/* Peaking / pumped-storage operations (daily_hours, stryke.py:2405-2431): in every cell each unit flips a not-operating coin
 * (uniform <= Prob_Not_Op -> 0 hours) and otherwise draws lognormal hours, exp(shape Z) scale + location clipped to
 * [0, 24]; a cell whose units all sit idle is skipped altogether (no fish, no rows: stryke.py:3025).  Drawn on the device from
 * the cell's Philox stream (counter word 0 = 0xFFFFFFFE - unit).  params[s][u][.] for species row s (the Operating Scenarios
 * rows of its scenario) and unit u: 0 Prob_Not_Op, 1 shape, 2 location, 3 scale, 4 fixed hours (used when shape is NaN: a
 * run-of-river facility next to a peaking one). */
typedef struct StrykePeaking {
  int32_t n_units;
  const double* params;         /* [n_species][n_units][5]                                                          */
  float* hours;                 /* out, HOST, [n_cells][n_units] hours operated; may be NULL                         */
} StrykePeaking;

typedef struct StrykeOpts {
  int32_t  device;              /* CUDA ordinal                                                            */
  int32_t  precision;           /* 32 or 64 (bits of the per-fish arithmetic)                               */
  uint64_t seed;                /* Philox4x32-10 key (native RNG)                                          */
  int32_t  max_daily_fish;      /* STRYKE_MAX_DAILY_FISH (stryke.py:2610); <=0 disables the check          */
  int32_t  blocks_per_sm;       /* 0 = default                                                             */
  const StrykeInjected* injected;  /* NULL = native Philox                                                 */
  const StrykeTrace*    trace;     /* NULL = tallies only (needs fish_off: injected mode or fixed counts)   */
  int32_t  kernel_variant;      /* 0 = auto; 1 = general passage_kernel; 2 = throughput kernel, generic tables (ahead-of-time);
                                   3 = throughput kernel specialised for the facility with NVRTC (error if unavailable).
                                   auto: fp32 + native RNG + no trace -> throughput kernel, specialised when the run is large */
  int32_t  rng_mode;            /* native stream: 0 = compact word budget (16/23-bit centred uniforms, one cause uniform; all kernels);
                                   1 = wide: every draw a 53-bit uniform of its own, Box-Muller from two of them, the reference's
                                   sequential cause rolls (stryke.py:1545-1560) — fp64 general kernel only; the arbiter the
                                   throughput kernels are checked against at 1e10 fish (tests/test_gpu_high_power.py) */
  uint32_t flags;               /* STRYKE_FLAG_* */
  const StrykePeaking* peaking; /* NULL = every listed cell operates (run-of-river: the host lists only days with hours) */
} StrykeOpts;

#define STRYKE_FLAG_NO_PROMOTE 1u   /* keep a drawn count of 0 (population_sim itself, stryke.py:2596-2618; run() promotes it to 1, :3032) */

/* Host-side result buffers, caller-allocated. */
typedef struct StrykeTallies {
  int32_t*  cell_n;        /* [n_cells] pop_size (0 = species absent that day, zero row stryke.py:3392-3426) */
  uint32_t* daily;         /* [n_cells][STRYKE_DAILY_W]                                                     */
  uint32_t* state_daily;   /* [n_cells][n_pairs][2] successes, count (stryke.py:3302-3351)                   */
} StrykeTallies;

/* Compact tallies: one row per cell that holds fish (nothing for a species absent that day, stryke.py:3392-3426), 16-bit
 * counters wherever the cell's population allows.  A slot is row_w = n_pairs + 3 uint32 words.
 *   narrow cell (pop_size <= narrow_max), ONE slot:  w0 = pop_size | mortality_barotrauma << 16,
 *       w1 = num_entrained | num_survived << 16,  w2 = mortality_impingement | mortality_blade_strike << 16,
 *       w[3 + p] = successes_p | count_p << 16   (State_Daily of reachable pair p, stryke.py:3302-3351)
 *   wide cell, TWO slots:  w0 = pop_size, w1..w5 = num_entrained, num_survived, mortality_impingement, _blade_strike,
 *       _barotrauma, w[6 + 2p] = successes_p, w[7 + 2p] = count_p
 * Cells are taken in aligned blocks of 32: blk_row[b] = slot of the block's first present cell, bit j of blk_mask[b] =
 * cell 32 b + j holds fish, bit j of blk_wide[b] = it uses the wide layout.  The slot of cell 32 b + j is
 * blk_row[b] + popcount(blk_mask[b] & ((1 << j) - 1)) + popcount(blk_wide[b] & ((1 << j) - 1)).
 * HOST buffers for stryke_b200_run_compact (pinned memory lets the copies overlap the kernels). */
typedef struct StrykeCompact {
  int64_t   n_blocks;       /* (n_cells + 31) / 32                                                            */
  uint32_t* blk_row;        /* [n_blocks]                                                                     */
  uint32_t* blk_mask;       /* [n_blocks]                                                                     */
  uint32_t* blk_wide;       /* [n_blocks]                                                                     */
  uint32_t* rows;           /* [rows_cap][row_w]                                                              */
  int64_t   rows_cap;       /* slots the caller allocated                                                     */
  int64_t   n_slots;        /* out: slots used (if > rows_cap the call fails with STRYKE_E_ARG and says so)  */
  int64_t   n_fish;         /* out: fish simulated                                                            */
  int32_t   row_w;          /* out: n_pairs + 3                                                               */
  int32_t   narrow_max;     /* out: largest pop_size kept in the narrow layout                                */
  int32_t   n_slabs;        /* in: cut the cell list into this many launches so that the device->host copy of one slab's
                               rows runs under the next slab's kernels (0 = choose)                           */
} StrykeCompact;

typedef struct StrykePlan StrykePlan;

const char* stryke_b200_last_error(void);
int      stryke_b200_abi_version(void);
int      stryke_b200_device_count(void);
uint64_t stryke_b200_launch_count(void);     /* kernels launched by this library in this process */

/* One-shot, host buffers in and out: H2D of the tables, count + scan + passage kernels, D2H of the tallies.
 * Replaces the iteration/day loops of simulation.run() (stryke.py:2897-3383) for the given cells. */
int stryke_b200_run(const StrykeFacility* fac, const StrykeDayTables* days, const StrykeSpeciesTable* species,
                    const StrykeCells* cells, const StrykeOpts* opts, StrykeTallies* out);

/* One-shot with compact host results: the production form of stryke_b200_run for population-mode projects (millions of
 * cells of a few dozen fish).  out->cell_n of `tallies` is not needed: pop_size rides in the rows. */
int stryke_b200_run_compact(const StrykeFacility* fac, const StrykeDayTables* days, const StrykeSpeciesTable* species,
                            const StrykeCells* cells, const StrykeOpts* opts, StrykeCompact* out);

/* Plan API: tables stay resident in HBM; tallies live in caller-owned DEVICE buffers (e.g. torch tensors that are
 * then all-reduced with NCCL); launches are asynchronous on `stream` (a cudaStream_t, 0 = legacy default). */
int stryke_b200_plan_create(const StrykeFacility* fac, const StrykeDayTables* days, const StrykeSpeciesTable* species,
                            const StrykeCells* cells, const StrykeOpts* opts, StrykePlan** plan);
int stryke_b200_plan_launch(StrykePlan* plan, int32_t* d_cell_n, uint32_t* d_daily, uint32_t* d_state_daily,
                            void* stream);
/* Same hot path with compact rows (see StrykeCompact), for cells [cell_lo, cell_hi) of the plan (cell_lo a multiple of
 * STRYKE_RANGE_ALIGN; cell_hi = -1: to the end).  d_rows holds rows_cap slots; the per-block tables live in the plan
 * (stryke_b200_plan_compact_tables).  Slots continue from the previous range when first_slot < 0, else start at first_slot.
 * A range that starts at cell 0 starts a new pass: the status words (error, fish and slot totals) start over.
 * d_slots_after (device, may be NULL) receives the slot count after this range — what a caller that overlaps the copy of one
 * range with the kernels of the next needs to know.  Ranges of one pass go to ONE stream, in order.
 * The launch initialises every row it hands out (whole-row stores for the cells one warp simulates alone, zero-fill +
 * atomic adds for the few whose tiles are split between warps): the caller never clears d_rows. */
int stryke_b200_plan_launch_compact(StrykePlan* plan, int64_t cell_lo, int64_t cell_hi, int32_t* d_cell_n, uint32_t* d_rows,
                                    int64_t rows_cap, int64_t first_slot, uint64_t* d_slots_after, void* stream);
/* DEVICE pointers of the plan's per-block tables [n_blocks] and the row geometry. */
int stryke_b200_plan_compact_tables(StrykePlan* plan, uint32_t** d_blk_row, uint32_t** d_blk_mask, uint32_t** d_blk_wide,
                                    int32_t* row_w, int32_t* narrow_max);
/* Let the launches write the per-block tables straight into caller-owned DEVICE arrays of n_blocks words each (e.g. inside the
 * buffer a rank exchanges with its peers) instead of the plan's own; NULLs switch back. */
int stryke_b200_plan_set_compact_tables(StrykePlan* plan, uint32_t* d_blk_row, uint32_t* d_blk_mask, uint32_t* d_blk_wide);
/* Slots and fish counted so far by the launches on `stream` (syncs the stream). */
int stryke_b200_plan_compact_totals(StrykePlan* plan, void* stream, int64_t* n_slots, int64_t* n_fish);
const char* stryke_b200_plan_kernel(const StrykePlan* plan);      /* which passage kernel the plan will launch */
int stryke_b200_set_source_dir(const char* dir);                 /* directory holding passage_fast.cuh / passage_common.cuh (run-time specialisation) */
int stryke_b200_jit_selftest(void);                              /* compiles a specialised kernel with NVRTC; needs no GPU */
/* Measurement support (no reference counterpart): with enable != 0 every later plan_launch brackets its passage kernel
 * (not the count / scan kernels before it) with CUDA events on the launch stream; plan_kernel_ms waits for them, returns
 * the summed kernel time and the number of launches since the last call, and forgets them. */
int stryke_b200_plan_time_kernel(StrykePlan* plan, int enable);
int stryke_b200_plan_kernel_ms(StrykePlan* plan, double* total_ms, int32_t* n_launches);
int stryke_b200_plan_set_seed(StrykePlan* plan, uint64_t seed);   /* Philox key of the next launches */
int stryke_b200_plan_status(StrykePlan* plan, void* stream, int64_t* n_fish_total);   /* syncs the stream; reports STRYKE_E_MAX_FISH etc. */
int stryke_b200_plan_fetch_trace(StrykePlan* plan, const StrykeTrace* host_trace, void* stream);
int stryke_b200_plan_fetch_hours(StrykePlan* plan, float* host_hours, void* stream);   /* [n_cells][n_units] of StrykePeaking */
int stryke_b200_plan_destroy(StrykePlan* plan);

/* Per-fish probability functions on the device (Kaplan/Propeller/Francis/Pump stryke.py:846-1127 and the
 * Pflugrath barotrauma response stryke.py:1492-1510 + Stryke/barotrauma.py:139-156).  Host arrays in/out.
 * params: H, RPM, D, Q, ada, N, Qopt, Qper, lambda, iota, D1, D2, B, Q_p, gamma (15 doubles). */
int stryke_b200_strike_survival(int kind, const double* params15, int64_t n, const double* length_ft,
                                const double* rR, int precision, int device, double* out);
int stryke_b200_baro_survival(int64_t n, const double* fish_depth_ft, double tail_depth_ft, double beta_0,
                              double beta_1, int precision, int device, double* out);

/* Philox4x32-10 block function evaluated on the device (for known-answer tests). */
int stryke_b200_philox(int64_t n, const uint32_t* ctr4, const uint32_t* key2, int device, uint32_t* out4);

/* Issue-rate calibration micro-kernel: lane-ops/s of 16 independent FFMA chains per thread at full occupancy. */
int stryke_b200_calibrate_issue(int device, void* stream, double* lane_ops_per_s, double* ms);

/* Means of n_boot resamples (with replacement, len n) of data[n]: the loop of bootstrap_mean_ci (Stryke/stryke.py:155-159)
 * that summary() (stryke.py:3520-4101) runs a few hundred times.  Host pointers; indices from Philox4x32-10 keyed by seed. */
int stryke_b200_bootstrap_means(const double* data, int64_t n, int32_t n_boot, uint64_t seed, int32_t device, double* means);

#ifdef __cplusplus
}
#endif
#endif /* STRYKE_B200_H */
