// GEMM job engine for the Parrot decoder hot path.
//
// Every dense contraction on the path -- the per-step GRU gate / candidate
// products inside the scan, their reverse-time dgrads, and the batched
// readout / feedback / weight-gradient products outside it -- is expressed as a
// *job*: one 128-row output tile  D[128 x N] = sum_seg A_seg[128 x K_seg] * B_seg[N x K_seg]^T
// followed by a fused epilogue.  Operands are bf16 hi/lo planes (x = hi + lo to
// ~2^-17), both K-major; the tile is computed as hi*hi + lo*hi + hi*lo with
// fp32 accumulation in TMEM ("bf16x3", see DESIGN.md: single-pass bf16/tf32 fail
// the reference-parity gate through an 800-step recurrence).
//
// Weights are the M-side (A) operand so that the 128-lane TMEM datapath is full
// at batch 64 (SURVEY hard part 2); the batch / frame axis is the N side.
//
// kernel roles (192 threads, 1 CTA / SM):
//   warp 0      TMA producer: cp.async.bulk.tensor -> 128B-swizzled smem ring
//   warp 1      tcgen05.mma issuer (one elected lane), owns the TMEM allocation
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> fused epilogue -> global
//
// A SIMT twin (job_kernel_simt) executes the same job table with plain loads and
// fp32 FMAs through the same epilogues.  It exists for verification only: the
// tests compare it with the tensor-core path at BASELINE sizes where the CPU
// oracle is too slow.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace pb {

typedef __nv_bfloat16 bf16;

constexpr int TILE_M = 128;
constexpr int KB = 64;  // k elements per pipeline block (one 128-byte swizzle span)
constexpr int MAX_SEG = 6;
constexpr int NO_SLOT = -(1 << 30);
// 8 warps: warp 0 TMA, warp 1 MMA, warps 2-5 TMEM epilogue, warps 6-7 epilogue helpers.  Two warps per scheduler
// sub-partition leave the full 255 registers per thread: the epilogues keep every partial-tile / operand load of
// a work item in flight, and ONE spilled load result serialises those L2 round trips.  Measured on one box, same
// session (ms per training step at the base configuration): 12 warps (168-register cap, 430 B of spills) 138.9,
// 10 warps 143.9, 16 warps 140.7, 8 warps 120.2, 8 warps + MAX_KSPLIT 6 115.4.
// (PB_* are experiment knobs for variant builds, see tools/build_variants.sh)
#ifndef PB_ENGINE_THREADS
#define PB_ENGINE_THREADS 256
#endif
constexpr int ENGINE_THREADS = PB_ENGINE_THREADS;
constexpr int EPI_GROUP_THREADS = ENGINE_THREADS - 64; // warps 2.. share the post-reduction (split-K) epilogue
__device__ __forceinline__ void epi_group_sync() {
  asm volatile("bar.sync 1, %0;" ::"r"(EPI_GROUP_THREADS) : "memory");
}
constexpr int SMEM_BYTES = 200 * 1024;
#ifndef PB_MAX_KSPLIT
#define PB_MAX_KSPLIT 6   // 4: 121.5, 5: 119.9, 6: 115.4, 8: 120.2 ms/step (8 warps)
#endif
constexpr int MAX_KSPLIT = PB_MAX_KSPLIT;   // scratch stride per split group

enum Epi : int {
  EPI_PLAIN = 0,      // out = acc*scale (+bias) (+= out), optional hi/lo planes
  EPI_GATES = 1,      // forward scan: sigmoid -> z / r, r*h_prev planes
  EPI_CAND = 2,       // forward scan: tanh -> c, h_t, planes
  EPI_BWD_RH = 3,     // backward scan: d(r*h) -> gate pre-activation grads
  EPI_BWD_STATE = 4,  // backward scan: accumulate into a carried-state gradient
};

struct Seg {
  int a_map;   // tensor-map index of the A hi plane (lo plane = a_map + 1)
  int a_row;   // first row of the tile in A
  int a_k;     // first k column in A
  int b_map;   // tensor-map index of the B hi plane (lo = b_map + 1)
  int b_row;   // first row (sample) in B
  int b_k;     // first k column in B
  int b_slot;  // NO_SLOT: B map is 2-D; else 3-D and the slot coordinate is t + b_slot
  int nkb;     // number of 64-wide k blocks
  int a_nkb;   // > 0: A is a tile-contiguous weight pack with a_nkb k blocks per 128-row tile: tile
               // (row_tile, k_block) is rows [(row_tile*a_nkb + k_block)*128, +128) of a [.][64] matrix
  int b_slots; // slots of the B plane (copied from MapRaw at table build: no dependent global load in seg_valid)
};

struct PlainArgs {
  float* out;
  const float* bias;   // indexed by output row (feature); may be null
  bf16* hi;            // optional bf16 planes of the result, [sample][feature]
  bf16* lo;
  long long ldo;       // leading dimension of out (floats)
  long long ldp;       // leading dimension of the planes (elements)
  long long out_tstride;  // added per scan step t (scan-resident plain jobs)
  int n_pad;           // >0: sample n = q*n_pad + r, stored at row q*n_valid + r when r < n_valid
  int n_valid;
  int n_total;         // samples >= n_total are not stored
  int flags;           // 1 accumulate, 2 store transposed (out[row*ldo + n]), 4 planes indexed by padded n
  float scale;
  int rowbias_ld;      // > 0: `bias` is a [n_pad rows][rowbias_ld] matrix indexed by (sample % n_pad, feature)
};
enum { PF_ACC = 1, PF_TRANS = 2, PF_PLANE_PADDED = 4 };
enum { QF_FUSED_PRE = 256 };   // EPI_BWD_STATE job flag (Job::pa.flags): run the GRU pre-pass of step t - 1 in the finish   // EPI_BWD_STATE job flag (Job::pa.flags): run the GRU pre-pass of step t - 1 in the finish

struct Job {
  int nseg, epi, lag, layer;
  int row0;     // first output row (feature) of the tile
  int m_valid;  // rows of the tile that exist
  int n0;       // first sample column
  int aux;      // epilogue specific
  // split-K: `ksplit` jobs share one output tile (`group`); part `kpart` contracts its share of the
  // k blocks, parks the partial tile in scratch, and the last part to arrive sums all parts in part
  // order (deterministic) and runs the epilogue.
  int ksplit, kpart, group;
  // grouped persistent scans: the LAST `res_kb` k blocks of this part's range keep their weight tiles resident in
  // tensor memory (A operand of tcgen05.mma read from TMEM: no TMA traffic, no shared-memory operand reads);
  // hi plane of resident block i at TMEM column res_col + 64 i, lo plane 32 columns behind it
  int res_kb;
  int res_col, pad0_, pad1_, pad2_;
  Seg seg[MAX_SEG];
  PlainArgs pa;
};

// raw addresses behind each tensor map (SIMT twin + debugging)
struct MapRaw {
  const bf16* base;
  long long row_pitch;   // elements
  long long slot_pitch;  // elements (0 for 2-D)
  int rows, cols;        // extents per slot (logical, also for tiled packs)
  int slots;
  int box_rows;
  int tiled_nkb;         // > 0: tile-contiguous weight pack (see Seg::a_nkb)
  int pad_;
};

// ---- scan context (device resident; see scan.cu for the phase structure) ----
struct LayerBuf {
  float* h;            // [T+1][B][H]   state sequence, slot 0 = state entering the segment
  bf16 *h_hi, *h_lo;   // [T+1][Np][Hp]
  bf16 *rh_hi, *rh_lo; // [T][Np][Hp]   (reset * previous state), B operand of the candidate product
  float *z, *r, *c;    // [T][B][H]
  const float* base;   // [B][3H]  time-constant input: summed Fork biases (+ speaker), [cell | gates]
  const float* fb;     // [T][B][3H] teacher-forcing feedback term or null
  const float* pre;    // [T][B][3H] hoisted pre-activation terms (products whose inputs do not depend on this
                       // layer's own recurrence: teacher-forced feedback, lower layers, attention context), or null
  // backward
  float* dh;           // [T+1][B][H]  gradient wrt h slot s (accumulated)
  float* drh;          // [B][H] scratch: d(r*h) of the current step
  float* da;           // [T][B][3H]   pre-activation grads [cell | gates] (fp32, for bias / feedback grads)
  bf16 *da_hi, *da_lo; // [T][Np][3Hp] (cell at 0, gates at Hp)
};
struct ScanCtx {
  int T, B, Np, H, Hp, C, Cp, A, U;
  long long base_tstride;   // 0: L.base is [B][3H] ; layer_norm mode: [T][B][3H] (per-step pre-activation terms)
  LayerBuf L[3];
  float* dw;           // [T+1][B][C] gradient wrt w slot s
};

struct EngineParams {
  const Job* jobs;
  int njobs;
  const CUtensorMap* maps;
  const MapRaw* raws;
  const ScanCtx* ctx;
  int tick;       // scan tick; t = tick - job.lag (forward) or as given by dir
  int T;          // jobs whose t falls outside [0, T) are skipped
  int n_cols;     // UMMA N of every job in this launch (box rows of every B map used)
  int reverse;    // 0: t = tick - lag ; 1: t = (T - 1) - (tick - lag)
  int chunk_samples;  // > 0: EPI_PLAIN jobs of this table cover ONE chunk of `chunk_samples` consecutive samples; the
                      // launch's `tick` is the chunk event e, job j works on chunk e - j.lag (see chunk_window)
  float* split_scratch;        // [group][part][n_cols][128] partial tiles
  unsigned int* split_count;   // [group] arrival counters (zero between launches)
  int debug_flags;             // reserved for experiments
  int coop_epilogue;           // 1: all parts of a split tile share the final epilogue (needs <= 1 job per CTA)
  unsigned long long* timeline;  // debug: [cta][16] globaltimer stamps at pipeline milestones (or null)
  int tl_tick;                   // debug: only the launch / persistent tick with this value writes the timeline (-1: any)
  int cta0, ncta;                // the CTAs [cta0, cta0 + ncta) of the grid run this table (CTA rank r takes jobs r, r + ncta, ..)
};
__device__ __forceinline__ int cta_rank(const EngineParams& P) { return (int)blockIdx.x - P.cta0; }

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot)                                                                                 \
  do {                                                                                           \
    if (P.timeline && (P.tl_tick < 0 || P.tl_tick == tick))                                      \
      P.timeline[(size_t)blockIdx.x * 16 + (slot)] = gtime();                                    \
  } while (0)

__device__ __forceinline__ int job_time(const EngineParams& P, const Job& jb, int tick) {
  int t = tick - jb.lag;
  if (P.chunk_samples > 0 && jb.epi == EPI_PLAIN) return t;   // chunk index, see chunk_window
  if (P.reverse) t = (P.T - 1) - t;
  return t;
}
// Sample window [shift, limit) of a plain job.  Ordinary tables: the whole sample axis.  Chunk tables (products whose
// operands become available one chunk of time steps at a time inside the persistent scans): chunk ce = event - lag
// covers samples [ce * cs, (ce + 1) * cs) forward, [n_total - (ce + 1) * cs, n_total - ce * cs) in the reverse sweep,
// clipped to [0, n_total); the job's own sample tile n0 is relative to `shift`.
__device__ __forceinline__ bool chunk_window(const EngineParams& P, const Job& jb, int ce, int& shift, int& limit) {
  shift = 0; limit = jb.pa.n_total;
  if (P.chunk_samples <= 0) return true;
  if (ce < 0) return false;
  const int cs = P.chunk_samples, nt = jb.pa.n_total;   // (sample counts are far below 2^31 / chunks)
  if (ce > nt / cs) return false;
  if (!P.reverse) {
    if (ce * cs >= nt) return false;
    shift = ce * cs;
    limit = (ce + 1) * cs < nt ? (ce + 1) * cs : nt;
  } else {
    const int hi = nt - ce * cs;
    if (hi <= 0) return false;
    const int lo = hi - cs;
    shift = lo > 0 ? lo : 0;
    limit = hi;
  }
  return true;
}

// A segment whose B operand is slot-indexed is skipped when its slot falls outside the plane
// (merged backward jobs straddle three pipeline stages; at the ends of the sweep some are absent).
__device__ __forceinline__ bool seg_valid(const EngineParams& P, const Seg& sg, int t) {
  if (sg.b_slot == NO_SLOT) return true;
  const int s = t + sg.b_slot;
  return s >= 0 && s < sg.b_slots;
}
// total number of k blocks of the job at time t; 0 means "skip this job"
__device__ __forceinline__ int job_total_kb(const EngineParams& P, const Job& jb, int t) {
  if (jb.epi != EPI_BWD_STATE && (t < 0 || t >= P.T)) return 0;
  if (jb.epi == EPI_BWD_STATE && jb.lag != 0 && (t < 0 || t >= P.T)) return 0;   // single-time state jobs
  if (jb.epi == EPI_PLAIN) {
    int shift, limit;
    if (!chunk_window(P, jb, t, shift, limit) || jb.n0 + shift >= limit) return 0;
  }
  int n = 0;
  for (int s = 0; s < jb.nseg; ++s)
    if (seg_valid(P, jb.seg[s], t)) n += jb.seg[s].nkb;
  return n;
}

// k-block range [lo, hi) of the flattened (valid segment, block) sequence owned by this split part
__device__ __forceinline__ void job_kb_range(const Job& jb, int total_kb, int& lo, int& hi) {
  if (jb.ksplit <= 1) { lo = 0; hi = total_kb; return; }
  lo = (total_kb * jb.kpart) / jb.ksplit;          // (total_kb < 2^20, kpart < 8: 32-bit arithmetic is enough)
  hi = (total_kb * (jb.kpart + 1)) / jb.ksplit;
}

// ------------------------------------------------------------------ epilogues
// Everything an epilogue needs, snapshotted into registers ONCE per job.  The job table and the scan context
// live in global memory; reading them inside the store loops forces the compiler to reload pointers and
// scalars after every store (possible aliasing), which serialises the epilogue into dependent L2 round trips.
// Kept small (generic pointer slots): register pressure in the epilogues is what limits the scan (see top):
//   GATES      p0 base  p1 h   p2 z   p3 r    p4 rh_hi  p5 rh_lo
//   CAND       p0 base  p1 h   p2 z   p3 c    p4 h_hi   p5 h_lo
//   BWD_RH     p0 r     p1 h   p2 dh  p3 da   p4 da_hi  p5 da_lo
//   BWD_STATE  p0 dst (F = dstF, slot offset = slot_off)
//   PLAIN      p0 out   p1 bias p2 hi p3 lo ; l0 ldo, l1 ldp, l2 out_tstride ; n_pad n_valid n_total flags scale
struct EpiLocal {
  int epi, row0, m_valid, n0, slot_off;
  int B, H, Np, Hp, dstF;
  int n_pad, n_valid, n_total, flags;
  int rowbias_ld;
  float scale;
  void *p0, *p1, *p2, *p3, *p4, *p5;
  const float* pre;   // GATES / CAND: hoisted pre-activation terms [T][B][3H] (or null), added to base
  long long l0, l1, l2;
};
__device__ __forceinline__ EpiLocal make_epi_local(const Job& jb, const ScanCtx* ctx, int n_shift = 0,
                                                   int n_limit = 0) {
  EpiLocal E;
  E.epi = jb.epi; E.row0 = jb.row0; E.m_valid = jb.m_valid; E.n0 = jb.n0;
  E.slot_off = jb.pa.n_pad;
  E.B = E.H = E.Np = E.Hp = E.dstF = 0;
  E.n_pad = E.n_valid = E.n_total = E.flags = 0; E.scale = 1.0f; E.rowbias_ld = 0;
  E.p0 = E.p1 = E.p2 = E.p3 = E.p4 = E.p5 = nullptr;
  E.pre = nullptr;
  E.l0 = E.l1 = E.l2 = 0;
  if (jb.epi == EPI_PLAIN) {
    E.p0 = jb.pa.out; E.p1 = (void*)jb.pa.bias; E.p2 = jb.pa.hi; E.p3 = jb.pa.lo;
    E.l0 = jb.pa.ldo; E.l1 = jb.pa.ldp; E.l2 = jb.pa.out_tstride;
    E.n_pad = jb.pa.n_pad; E.n_valid = jb.pa.n_valid; E.n_total = jb.pa.n_total; E.flags = jb.pa.flags;
    E.scale = jb.pa.scale;
    E.rowbias_ld = jb.pa.rowbias_ld;
    E.n0 = jb.n0 + n_shift;
    if (n_limit > 0) E.n_total = n_limit;
    return E;
  }
  E.B = ctx->B; E.H = ctx->H; E.Np = ctx->Np; E.Hp = ctx->Hp;
  E.l0 = ctx->base_tstride;
  const LayerBuf& L = ctx->L[jb.layer];
  switch (jb.epi) {
    case EPI_GATES:
      E.p0 = (void*)L.base; E.p1 = L.h; E.p2 = L.z; E.p3 = L.r; E.p4 = L.rh_hi; E.p5 = L.rh_lo; E.pre = L.pre; break;
    case EPI_CAND:
      E.p0 = (void*)L.base; E.p1 = L.h; E.p2 = L.z; E.p3 = L.c; E.p4 = L.h_hi; E.p5 = L.h_lo; E.pre = L.pre; break;
    case EPI_BWD_RH:
      E.p0 = L.r; E.p1 = L.h; E.p2 = L.dh; E.p3 = L.da; E.p4 = L.da_hi; E.p5 = L.da_lo; break;
    case EPI_BWD_STATE:
      if (jb.aux == 3) { E.p0 = ctx->dw; E.dstF = ctx->C; }
      else {
        const LayerBuf& D = ctx->L[jb.aux];
        E.p0 = D.dh; E.dstF = ctx->H;
        if (jb.pa.flags & QF_FUSED_PRE) {
          E.flags = QF_FUSED_PRE;
          E.p1 = D.z; E.p2 = D.c; E.p3 = D.h; E.p4 = D.da_hi; E.p5 = D.da_lo; E.pre = D.da;
        }
      }
      break;
  }
  return E;
}

// Thread <-> output row (feature).  v[j] is the accumulator for sample n_base + j.
// sample index -> destination row: n = q * n_pad + r (block q of n_pad padded rows, row r) is stored at row
// q * n_valid + r when r < n_valid.
template <int W>
struct EpiOps { float a[W], b[W], c[W]; };

// operands of the plain epilogue that come from memory (old value for PF_ACC, row bias): requested for all W columns
// before the accumulator is consumed, so that a tile costs one L2 round trip per W columns instead of one per element
template <int W>
__device__ __forceinline__ void epi_plain_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  const bool acc = (E.flags & PF_ACC) != 0, rb = E.rowbias_ld > 0;
  if (!acc && !rb) return;
  const int f = E.row0 + row;
  const float* out = E.p0 ? (const float*)E.p0 + (long long)t * E.l2 : nullptr;
  const float* biasp = (const float*)E.p1;
  int n = E.n0 + n_base, q = 0, r = n;
  if (E.n_pad > 0) { q = n / E.n_pad; r = n - q * E.n_pad; }
#pragma unroll
  for (int j = 0; j < W; ++j, ++n) {
    o.a[j] = 0.0f; o.b[j] = 0.0f;
    const bool live = row < E.m_valid && j < ncols && n < E.n_total;
    long long srow = n;
    int rr = r;
    if (E.n_pad > 0) {
      srow = (long long)q * E.n_valid + rr;
      if (++r == E.n_pad) { r = 0; ++q; }
    }
    if (!live || (E.n_pad > 0 && rr >= E.n_valid)) continue;
    if (acc && out) o.a[j] = __ldcg((E.flags & PF_TRANS) ? out + (long long)f * E.l0 + srow : out + srow * E.l0 + f);
    if (rb) o.b[j] = __ldg(biasp + (long long)rr * E.rowbias_ld + f);
  }
}
template <int W>
__device__ __forceinline__ void epi_plain(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                          const float* v, const EpiOps<W>& o) {
  if (row >= E.m_valid) return;
  const int f = E.row0 + row;
  const float* biasp = (const float*)E.p1;
  const bool acc = (E.flags & PF_ACC) != 0, rb = E.rowbias_ld > 0;
  const float bias = (biasp && !rb) ? __ldg(biasp + f) : 0.0f;
  float* out = E.p0 ? (float*)E.p0 + (long long)t * E.l2 : nullptr;
  bf16* hi = (bf16*)E.p2;
  bf16* lo = (bf16*)E.p3;
  // sample index -> (block q of n_pad padded rows, row r inside it): one division per call, then increments
  int n = E.n0 + n_base, q = 0, r = n;
  if (E.n_pad > 0) { q = n / E.n_pad; r = n - q * E.n_pad; }
#pragma unroll
  for (int j = 0; j < W; ++j, ++n) {
    if (j >= ncols || n >= E.n_total) break;
    long long srow = n;
    if (E.n_pad > 0) {
      const int rr = r;
      srow = (long long)q * E.n_valid + rr;
      if (++r == E.n_pad) { r = 0; ++q; }
      if (rr >= E.n_valid) continue;
    }
    float y = v[j] * E.scale + bias;
    if (rb) y += o.b[j];
    if (out) {
      float* p = (E.flags & PF_TRANS) ? out + (long long)f * E.l0 + srow : out + srow * E.l0 + f;
      if (acc) y += o.a[j];
      *p = y;
    }
    if (hi) {
      const long long prow = (E.flags & PF_PLANE_PADDED) ? (long long)n : srow;
      bf16 h, l;
      split_bf16(y, h, l);
      hi[prow * E.l1 + f] = h;
      lo[prow * E.l1 + f] = l;
    }
  }
}

// The scan epilogues are split in two stages so that callers can issue the operand loads BEFORE they consume
// the accumulator (in-order issue: a load placed after the first use of an outstanding load cannot start).
// ---- forward scan, gates tile: rows [0, 2H) of [update | reset]   (model.py:659-662)
template <int W>
__device__ __forceinline__ void epi_gates_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  const int H = E.H, B = E.B, f = E.row0 + row;
  const bool is_z = f < H;
  const int fr = is_z ? f : f - H;
  const float* __restrict__ basep = (const float*)E.p0 + (long long)t * E.l0 + H + f;
  const float* __restrict__ prep = E.pre ? E.pre + (long long)t * B * 3 * H + H + f : nullptr;
  const float* __restrict__ hprev = (const float*)E.p1 + (long long)t * B * H + fr;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    const bool ok = row < E.m_valid && j < ncols && b < B;
    o.a[j] = !ok ? 0.0f : (prep ? __ldcg(prep + (long long)b * 3 * H) : __ldg(basep + (long long)b * 3 * H));
    o.b[j] = (ok && !is_z) ? hprev[(long long)b * H] : 0.0f;
  }
}
template <int W>
__device__ __forceinline__ void epi_gates_apply(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                                const float* v, const EpiOps<W>& o) {
  if (row >= E.m_valid) return;
  const int H = E.H, B = E.B, f = E.row0 + row;
  const bool is_z = f < H;
  const int fr = is_z ? f : f - H;
  float* __restrict__ outp = (float*)(is_z ? E.p2 : E.p3) + (long long)t * B * H + fr;
  bf16* __restrict__ phi = (bf16*)E.p4 + (long long)t * E.Np * E.Hp + fr;
  bf16* __restrict__ plo = (bf16*)E.p5 + (long long)t * E.Np * E.Hp + fr;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    if (j < ncols && b < B) {
      const float g = sigmoidf_fast(v[j] + o.a[j]);
      outp[(long long)b * H] = g;
      if (!is_z) {
        bf16 hh, ll;
        split_bf16(g * o.b[j], hh, ll);
        phi[(long long)b * E.Hp] = hh;
        plo[(long long)b * E.Hp] = ll;
      }
    }
  }
}

// ---- forward scan, candidate tile: rows [0, H)
template <int W>
__device__ __forceinline__ void epi_cand_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  const int H = E.H, B = E.B, f = E.row0 + row;
  const long long tb = (long long)t * B * H + f;
  const float* __restrict__ basep = (const float*)E.p0 + (long long)t * E.l0 + f;
  const float* __restrict__ prep = E.pre ? E.pre + (long long)t * B * 3 * H + f : nullptr;
  const float* __restrict__ zp = (const float*)E.p2 + tb;
  const float* __restrict__ hp_ = (const float*)E.p1 + tb;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    const bool ok = row < E.m_valid && j < ncols && b < B;
    o.a[j] = !ok ? 0.0f : (prep ? __ldcg(prep + (long long)b * 3 * H) : __ldg(basep + (long long)b * 3 * H));
    o.b[j] = ok ? zp[(long long)b * H] : 0.0f;
    o.c[j] = ok ? hp_[(long long)b * H] : 0.0f;
  }
}
template <int W>
__device__ __forceinline__ void epi_cand_apply(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                               const float* v, const EpiOps<W>& o) {
  if (row >= E.m_valid) return;
  const int H = E.H, B = E.B, f = E.row0 + row;
  const long long tb = (long long)t * B * H + f;
  float* __restrict__ hp_ = (float*)E.p1 + tb;          // slot t + 1 = + B*H
  float* __restrict__ cp = (float*)E.p3 + tb;
  bf16* __restrict__ phi = (bf16*)E.p4 + (long long)(t + 1) * E.Np * E.Hp + f;
  bf16* __restrict__ plo = (bf16*)E.p5 + (long long)(t + 1) * E.Np * E.Hp + f;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    if (j < ncols && b < B) {
      const float cc = tanhf_fast(v[j] + o.a[j]);
      const float hn = cc * o.b[j] + o.c[j] * (1.0f - o.b[j]);
      cp[(long long)b * H] = cc;
      hp_[(long long)b * H + (long long)B * H] = hn;  // slot t + 1
      bf16 hh, ll;
      split_bf16(hn, hh, ll);
      phi[(long long)b * E.Hp] = hh;
      plo[(long long)b * E.Hp] = ll;
    }
  }
}

// ---- backward scan: tile of d(r*h) = da_c * Ws^T, rows = state features [0, H).
// Finishes the GRU step backward for those features:
//   dr = drh * h_prev ; dh_prev += drh * r ; da_g[reset half] = dr r (1-r) -> planes + fp32
// (the update half of da_g and da_c come from the elementwise pre-pass, gru_bwd_pre_kernel)
template <int W>
__device__ __forceinline__ void epi_bwd_rh_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  const int H = E.H, B = E.B, f = E.row0 + row;
  const long long tb = (long long)t * B * H + f;
  const float* __restrict__ rp = (const float*)E.p0 + tb;
  const float* __restrict__ hpv = (const float*)E.p1 + tb;
  const float* __restrict__ dhp = (const float*)E.p2 + tb;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    const bool ok = row < E.m_valid && j < ncols && b < B;
    o.a[j] = ok ? rp[(long long)b * H] : 0.0f;
    o.b[j] = ok ? hpv[(long long)b * H] : 0.0f;
    o.c[j] = ok ? dhp[(long long)b * H] : 0.0f;
  }
}
template <int W>
__device__ __forceinline__ void epi_bwd_rh_apply(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                                 const float* v, const EpiOps<W>& o) {
  if (row >= E.m_valid) return;
  const int H = E.H, B = E.B, f = E.row0 + row;
  const long long tb = (long long)t * B * H + f;
  float* __restrict__ dhp = (float*)E.p2 + tb;    // slot t (state before the step)
  float* __restrict__ dap = (float*)E.p3 + (long long)t * B * 3 * H + 2 * H + f;
  bf16* __restrict__ phi = (bf16*)E.p4 + (long long)t * E.Np * (3 * E.Hp) + E.Hp + H + f;
  bf16* __restrict__ plo = (bf16*)E.p5 + (long long)t * E.Np * (3 * E.Hp) + E.Hp + H + f;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    if (j < ncols && b < B) {
      const float drh = v[j];
      const float dr = drh * o.b[j];
      dhp[(long long)b * H] = o.c[j] + drh * o.a[j];
      const float dag = dr * o.a[j] * (1.0f - o.a[j]);
      dap[(long long)b * 3 * H] = dag;
      bf16 hh, ll;
      split_bf16(dag, hh, ll);
      phi[(long long)b * 3 * E.Hp] = hh;
      plo[(long long)b * 3 * E.Hp] = ll;
    }
  }
}

// ---- backward scan: accumulate a dgrad tile into a carried gradient buffer (dh of a layer or dw), slot t + slot_off
template <int W>
__device__ __forceinline__ void epi_bwd_state_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  const int F = E.dstF, B = E.B;
  const float* __restrict__ dst = (const float*)E.p0 + (long long)(t + E.slot_off) * B * F + (E.row0 + row);
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    o.a[j] = (row < E.m_valid && j < ncols && b < B) ? dst[(long long)b * F] : 0.0f;
  }
}
template <int W>
__device__ __forceinline__ void epi_bwd_state_apply(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                                    const float* v, const EpiOps<W>& o) {
  if (row >= E.m_valid) return;
  const int F = E.dstF, B = E.B;
  float* __restrict__ dst = (float*)E.p0 + (long long)(t + E.slot_off) * B * F + (E.row0 + row);
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int b = n_base + j;
    if (j < ncols && b < B) dst[(long long)b * F] = o.a[j] + v[j];
  }
}

// W = number of columns held in v[] (compile time: the loops are fully unrolled over W)
template <int W>
__device__ __forceinline__ void epilogue_load(const EpiLocal& E, int t, int row, int n_base, int ncols, EpiOps<W>& o) {
  switch (E.epi) {
    case EPI_GATES: epi_gates_load<W>(E, t, row, n_base, ncols, o); break;
    case EPI_CAND: epi_cand_load<W>(E, t, row, n_base, ncols, o); break;
    case EPI_BWD_RH: epi_bwd_rh_load<W>(E, t, row, n_base, ncols, o); break;
    case EPI_BWD_STATE: epi_bwd_state_load<W>(E, t, row, n_base, ncols, o); break;
    default: epi_plain_load<W>(E, t, row, n_base, ncols, o); break;
  }
}
template <int W>
__device__ __forceinline__ void epilogue_apply(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                               const float* v, const EpiOps<W>& o) {
  switch (E.epi) {
    case EPI_PLAIN: epi_plain<W>(E, t, row, n_base, ncols, v, o); break;
    case EPI_GATES: epi_gates_apply<W>(E, t, row, n_base, ncols, v, o); break;
    case EPI_CAND: epi_cand_apply<W>(E, t, row, n_base, ncols, v, o); break;
    case EPI_BWD_RH: epi_bwd_rh_apply<W>(E, t, row, n_base, ncols, v, o); break;
    case EPI_BWD_STATE: epi_bwd_state_apply<W>(E, t, row, n_base, ncols, v, o); break;
  }
}
template <int W>
__device__ __forceinline__ void run_epilogue(const EpiLocal& E, int t, int row, int n_base, int ncols,
                                             const float* v) {
  EpiOps<W> o;
  epilogue_load<W>(E, t, row, n_base, ncols, o);
  epilogue_apply<W>(E, t, row, n_base, ncols, v, o);
}


// ------------------------------------------------------------------ quad epilogues (split-K finish)
// Thread <-> FOUR consecutive output rows (features) of ONE sample column: every global access of the finish is a
// 128-bit access (fp32 stashes [.][b][f], partial tiles [col][row], operand planes 4 x bf16 = 8 bytes), and all
// loads of a thread's items are issued together: ONE L2 round trip per finish instead of one per 4-column item.
// Data produced earlier in the same persistent launch is read with ld.global.cg (L2), never through L1.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// 4 floats -> 4 bf16 hi + 4 bf16 lo (same rounding as split_bf16), packed for one 8-byte store each
__device__ __forceinline__ void split4(float4 x, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
  hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
  lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}
struct QOps { float4 a, b, c, d, e; };

// can this job's split-K finish use the quad path?  (feature counts divisible by 4, aligned stashes)
__device__ __forceinline__ bool quad_ok(const EpiLocal& E) {
  if (E.epi == EPI_PLAIN) return false;
  if ((E.m_valid & 3) || (E.row0 & 3)) return false;
  if (E.epi == EPI_BWD_STATE) return (E.dstF & 3) == 0;
  return (E.H & 3) == 0 && (E.Hp & 3) == 0;
}
// memory operands of the quad epilogues: array 0 / 1 / 2 of (t, row quad rq, sample b); null = not needed
//   GATES      0: hoisted-or-base pre-activation term   1: h_prev (reset rows only)
//   CAND       0: hoisted-or-base term   1: z   2: h_prev
//   BWD_RH     0: r   1: h_prev   2: dh
//   BWD_STATE  0: destination
//   BWD_STATE + QF_FUSED_PRE (grouped backward scan, own-layer state dgrad): the finish owns the FINAL dh[slot t] of
//   its elements, so it also runs the elementwise GRU backward pre-pass of step t - 1 on them (gru_bwd_pre_rows):
//              1: z   2: c   3: h (slot t - 1)   4: dh (slot t - 1)      -- all of step t - 1
template <int DIR = 0>
__device__ __forceinline__ int q_narr(const EpiLocal& E) {
  const int epi = E.epi;
  if (DIR == 1) return epi == EPI_GATES ? 2 : 3;
  if (epi == EPI_BWD_STATE) return (E.flags & QF_FUSED_PRE) ? 5 : 1;
  if (DIR == 2) return 3;
  return epi == EPI_GATES ? 2 : 3;
}
template <int DIR = 0>
__device__ __forceinline__ const float* q_src(const EpiLocal& E, int t, int rq, int b, int arr) {
  const int H = E.H, B = E.B, f = E.row0 + 4 * rq;
  if (4 * rq >= E.m_valid || b >= B) return nullptr;
  const long long tb = ((long long)t * B + b) * H;
  int epi = E.epi;
  if (DIR == 1 && epi != EPI_GATES) epi = EPI_CAND;        // (compile-time pruning of the cases a kernel cannot see)
  if (DIR == 2 && epi != EPI_BWD_STATE) epi = EPI_BWD_RH;
  switch (epi) {
    case EPI_GATES:
      if (arr == 0)
        return (E.pre ? E.pre + ((long long)t * B + b) * 3 * H
                      : (const float*)E.p0 + (long long)t * E.l0 + (long long)b * 3 * H) + H + f;
      return (arr == 1 && f >= H) ? (const float*)E.p1 + tb + (f - H) : nullptr;
    case EPI_CAND:
      if (arr == 0)
        return (E.pre ? E.pre + ((long long)t * B + b) * 3 * H
                      : (const float*)E.p0 + (long long)t * E.l0 + (long long)b * 3 * H) + f;
      return (const float*)(arr == 1 ? E.p2 : E.p1) + tb + f;
    case EPI_BWD_RH:
      return (const float*)(arr == 0 ? E.p0 : (arr == 1 ? E.p1 : E.p2)) + tb + f;
    case EPI_BWD_STATE:
      if (arr == 0) return (const float*)E.p0 + ((long long)(t + E.slot_off) * B + b) * E.dstF + f;
      if (!(E.flags & QF_FUSED_PRE) || t < 1) return nullptr;
      return (const float*)(arr == 1 ? E.p1 : (arr == 2 ? E.p2 : (arr == 3 ? E.p3 : E.p0))) + tb - (long long)B * H + f;
    default: return nullptr;
  }
}
template <int DIR = 0>
__device__ __forceinline__ void q_load(const EpiLocal& E, int t, int rq, int b, QOps& o) {
  const float* p0 = q_src<DIR>(E, t, rq, b, 0);
  const float* p1 = q_src<DIR>(E, t, rq, b, 1);
  const float* p2 = q_src<DIR>(E, t, rq, b, 2);
  o.a = p0 ? ldcg4(p0) : f4zero();
  o.b = p1 ? ldcg4(p1) : f4zero();
  o.c = p2 ? ldcg4(p2) : f4zero();
  o.d = f4zero(); o.e = f4zero();
  if (DIR != 1 && E.epi == EPI_BWD_STATE && (E.flags & QF_FUSED_PRE)) {
    const float* p3 = q_src<DIR>(E, t, rq, b, 3);
    const float* p4 = q_src<DIR>(E, t, rq, b, 4);
    o.d = p3 ? ldcg4(p3) : f4zero();
    o.e = p4 ? ldcg4(p4) : f4zero();
  }
}

template <int DIR = 0>
__device__ __forceinline__ void q_apply(const EpiLocal& E, int t, int rq, int b, float4 v, const QOps& o) {
  const int H = E.H, B = E.B, f = E.row0 + 4 * rq;
  if (4 * rq >= E.m_valid || b >= B) return;
  int epi = E.epi;
  if (DIR == 1 && epi != EPI_GATES) epi = EPI_CAND;
  if (DIR == 2 && epi != EPI_BWD_STATE) epi = EPI_BWD_RH;
  switch (epi) {
    case EPI_GATES: {
      const bool is_z = f < H;
      const int fr = is_z ? f : f - H;
      float4 g;
      g.x = sigmoidf_fast(v.x + o.a.x); g.y = sigmoidf_fast(v.y + o.a.y);
      g.z = sigmoidf_fast(v.z + o.a.z); g.w = sigmoidf_fast(v.w + o.a.w);
      *reinterpret_cast<float4*>((float*)(is_z ? E.p2 : E.p3) + ((long long)t * B + b) * H + fr) = g;
      if (!is_z) {
        uint2 hh, ll;
        split4(make_float4(g.x * o.b.x, g.y * o.b.y, g.z * o.b.z, g.w * o.b.w), hh, ll);
        const long long po = ((long long)t * E.Np + b) * E.Hp + fr;
        *reinterpret_cast<uint2*>((bf16*)E.p4 + po) = hh;
        *reinterpret_cast<uint2*>((bf16*)E.p5 + po) = ll;
      }
    } break;
    case EPI_CAND: {
      const long long tb = ((long long)t * B + b) * H + f;
      float4 cc, hn;
      cc.x = tanhf_fast(v.x + o.a.x); cc.y = tanhf_fast(v.y + o.a.y);
      cc.z = tanhf_fast(v.z + o.a.z); cc.w = tanhf_fast(v.w + o.a.w);
      hn.x = cc.x * o.b.x + o.c.x * (1.0f - o.b.x); hn.y = cc.y * o.b.y + o.c.y * (1.0f - o.b.y);
      hn.z = cc.z * o.b.z + o.c.z * (1.0f - o.b.z); hn.w = cc.w * o.b.w + o.c.w * (1.0f - o.b.w);
      *reinterpret_cast<float4*>((float*)E.p3 + tb) = cc;
      *reinterpret_cast<float4*>((float*)E.p1 + tb + (long long)B * H) = hn;   // slot t + 1
      uint2 hh, ll;
      split4(hn, hh, ll);
      const long long po = ((long long)(t + 1) * E.Np + b) * E.Hp + f;
      *reinterpret_cast<uint2*>((bf16*)E.p4 + po) = hh;
      *reinterpret_cast<uint2*>((bf16*)E.p5 + po) = ll;
    } break;
    case EPI_BWD_RH: {
      const long long tb = ((long long)t * B + b) * H + f;
      float4 dhn, dag;
      dhn.x = o.c.x + v.x * o.a.x; dhn.y = o.c.y + v.y * o.a.y; dhn.z = o.c.z + v.z * o.a.z; dhn.w = o.c.w + v.w * o.a.w;
      dag.x = (v.x * o.b.x) * o.a.x * (1.0f - o.a.x); dag.y = (v.y * o.b.y) * o.a.y * (1.0f - o.a.y);
      dag.z = (v.z * o.b.z) * o.a.z * (1.0f - o.a.z); dag.w = (v.w * o.b.w) * o.a.w * (1.0f - o.a.w);
      *reinterpret_cast<float4*>((float*)E.p2 + tb) = dhn;
      *reinterpret_cast<float4*>((float*)E.p3 + ((long long)t * B + b) * 3 * H + 2 * H + f) = dag;
      uint2 hh, ll;
      split4(dag, hh, ll);
      const long long po = ((long long)t * E.Np + b) * (3 * E.Hp) + E.Hp + H + f;
      *reinterpret_cast<uint2*>((bf16*)E.p4 + po) = hh;
      *reinterpret_cast<uint2*>((bf16*)E.p5 + po) = ll;
    } break;
    case EPI_BWD_STATE: {
      const float4 dh1 = f4add(o.a, v);
      *reinterpret_cast<float4*>((float*)E.p0 + ((long long)(t + E.slot_off) * B + b) * E.dstF + f) = dh1;
      if ((E.flags & QF_FUSED_PRE) && t >= 1) {
        // GRU backward pre-pass of step t - 1 on these elements (same arithmetic as gru_bwd_pre_rows)
        const int tp = t - 1;
        float4 dh0, dac, dagz;
#define PB_PRE1(m)                                              \
        {                                                       \
          const float dh = dh1.m, z = o.b.m, cc = o.c.m;        \
          dh0.m = o.e.m + dh * (1.0f - z);                      \
          dac.m = (dh * z) * (1.0f - cc * cc);                  \
          dagz.m = (dh * (cc - o.d.m)) * z * (1.0f - z);        \
        }
        PB_PRE1(x) PB_PRE1(y) PB_PRE1(z) PB_PRE1(w)
#undef PB_PRE1
        *reinterpret_cast<float4*>((float*)E.p0 + ((long long)tp * B + b) * H + f) = dh0;
        float* dap = const_cast<float*>(E.pre) + ((long long)tp * B + b) * 3 * H;
        *reinterpret_cast<float4*>(dap + f) = dac;
        *reinterpret_cast<float4*>(dap + H + f) = dagz;
        const long long po = ((long long)tp * E.Np + b) * (3 * E.Hp);
        uint2 hh, ll;
        split4(dac, hh, ll);
        *reinterpret_cast<uint2*>((bf16*)E.p4 + po + f) = hh;
        *reinterpret_cast<uint2*>((bf16*)E.p5 + po + f) = ll;
        split4(dagz, hh, ll);
        *reinterpret_cast<uint2*>((bf16*)E.p4 + po + E.Hp + f) = hh;
        *reinterpret_cast<uint2*>((bf16*)E.p5 + po + E.Hp + f) = ll;
      }
    } break;
    default: break;
  }
}

// ---- operand staging: the finish operands of a part's columns are copied global -> shared with cp.async (no
// registers, no waiting) as soon as the phase starts; the finish reads them back from shared memory after the
// partial-tile exchange.  Slot (column cs of the part, array arr): 128 rows x 4 bytes; thread `lane` owns bytes
// [16 lane, 16 lane + 16) of every slot it fills and is the only reader of them (no CTA barrier needed).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int DIR = 0>
__device__ __forceinline__ void q_stage(const EpiLocal& E, int t, int lane, int ew, int nwarps, int c_lo, int c_hi,
                                        uint8_t* stg, int stg_cols) {
  const int narr = q_narr<DIR>(E);
  for (int c = c_lo + ew; c < c_hi; c += nwarps) {
    const int cs = c - c_lo;
    if (cs >= stg_cols) break;
#pragma unroll
    for (int arr = 0; arr < 5; ++arr) {
      if (arr < narr && (arr < 3 || DIR != 1)) {
        const float* src = q_src<DIR>(E, t, lane, c, arr);
        if (src) cp_async16(stg + ((size_t)(cs * narr + arr) * TILE_M + 4 * lane) * 4, src);
      }
    }
  }
}
template <int DIR = 0>
__device__ __forceinline__ void q_fetch(const EpiLocal& E, int t, int lane, int c, int c_lo, const uint8_t* stg,
                                        int stg_cols, QOps& o) {
  const int cs = c - c_lo;
  if (cs >= stg_cols) { q_load<DIR>(E, t, lane, c, o); return; }
  const int narr = q_narr<DIR>(E);
  const float4* s0 = reinterpret_cast<const float4*>(stg + ((size_t)(cs * narr) * TILE_M + 4 * lane) * 4);
  o.a = s0[0];
  o.b = narr > 1 ? s0[TILE_M / 4] : f4zero();
  o.c = narr > 2 ? s0[2 * (TILE_M / 4)] : f4zero();
  o.d = (DIR != 1 && narr > 3) ? s0[3 * (TILE_M / 4)] : f4zero();
  o.e = (DIR != 1 && narr > 4) ? s0[4 * (TILE_M / 4)] : f4zero();
}

// Finish of one split-K part: per batch a thread handles cb = min(QC, QS / ksplit) columns and keeps ksplit * cb <= QS
// partial-tile quads in flight (slot i <-> column i / ksplit, part i % ksplit), so the register footprint does not
// depend on how a table happens to be split.  Column sums run in part order (deterministic).
constexpr int QC = 4;    // columns per batch
constexpr int QS = 12;   // partial-tile slots per batch
__device__ __forceinline__ int q_cols_per_batch(int ksplit) { return ksplit <= 3 ? QC : (QS / ksplit < QC ? QS / ksplit : QC); }

// SENT: the partial tiles are exchanged without a counter.  The scratch starts filled with SPLIT_SENTINEL (a NaN bit
// pattern no partial sum takes); a reader polls its quads until every float differs from it and writes the sentinel
// back once the quad is consumed -- the next writer of the slot is a full tick (>= 2 group barriers) away.  Saves the
// writer's fence + atomic + the reader's poll of the counter + the dependent re-read: ~1.5 us per phase.
constexpr uint32_t SPLIT_SENTINEL = 0xFFFFFFFFu;
__device__ __forceinline__ bool quad_ready(const float4& x) {
  return __float_as_uint(x.x) != SPLIT_SENTINEL && __float_as_uint(x.y) != SPLIT_SENTINEL &&
         __float_as_uint(x.z) != SPLIT_SENTINEL && __float_as_uint(x.w) != SPLIT_SENTINEL;
}
template <int DIR = 0, bool SENT = false>
__device__ __forceinline__ void q_finish(const EpiLocal& E, int t, int lane, int ew, int nwarps, int c_lo, int c_hi,
                                         int ksplit, int n_cols, const float* base, const uint8_t* stg, int stg_cols) {
  const int cb = q_cols_per_batch(ksplit);
  const size_t part_stride = (size_t)n_cols * TILE_M;
  for (int k0 = 0; c_lo + ew + nwarps * k0 < c_hi; k0 += cb) {
    float4 x[QS];
    unsigned int spins = 0;
    bool again;
    do {
      again = false;
      int kk = 0, pp = 0;
      const float* colp = base + (size_t)(c_lo + ew + nwarps * k0) * TILE_M + 4 * lane;   // column kk, part 0
      const float* ptr = colp;
#pragma unroll
      for (int i = 0; i < QS; ++i) {
        const bool live = kk < cb && c_lo + ew + nwarps * (k0 + kk) < c_hi;
        x[i] = live ? ldcg4(ptr) : f4zero();
        ptr += part_stride;
        if (++pp == ksplit) { pp = 0; ++kk; colp += (size_t)nwarps * TILE_M; ptr = colp; }
      }
      if (SENT) {
#pragma unroll
        for (int i = 0; i < QS; ++i) again |= !quad_ready(x[i]);
        if (again && ++spins > (1u << 22)) pb_timeout(2);
      }
    } while (SENT && again);
    if (SENT) {
      // consumed: hand the slots back (sentinel) for the same phase of the next tick
      int kk = 0, pp = 0;
      float* colp = const_cast<float*>(base) + (size_t)(c_lo + ew + nwarps * k0) * TILE_M + 4 * lane;
      float* ptr = colp;
      const float s = __uint_as_float(SPLIT_SENTINEL);
#pragma unroll
      for (int i = 0; i < QS; ++i) {
        const bool live = kk < cb && c_lo + ew + nwarps * (k0 + kk) < c_hi;
        if (live) __stcg(reinterpret_cast<float4*>(ptr), make_float4(s, s, s, s));
        ptr += part_stride;
        if (++pp == ksplit) { pp = 0; ++kk; colp += (size_t)nwarps * TILE_M; ptr = colp; }
      }
    }
    if (k0 == 0) cp_async_wait_all();   // this thread's staged operands (requested at the start of the phase)
#pragma unroll 1
    for (int k = 0; k < cb; ++k) {      // (not unrolled: one copy of the epilogue code; x[] keeps static indices)
      const int c = c_lo + ew + nwarps * (k0 + k);
      if (c >= c_hi) break;
      float4 v = f4zero();
      int kk = 0, pp = 0;
#pragma unroll
      for (int i = 0; i < QS; ++i) {
        if (kk == k) v = f4add(v, x[i]);   // part order: deterministic
        if (++pp == ksplit) { pp = 0; ++kk; }
      }
      QOps o;
      q_fetch<DIR>(E, t, lane, c, c_lo, stg, stg_cols, o);
      q_apply<DIR>(E, t, lane, c, v, o);
    }
  }
}

// Per-CTA copy of the ONE job a CTA runs in a scan table of a persistent kernel (job index = CTA index, every tick)
// and of its epilogue context, kept in shared memory: the roles read them with LDS instead of chains of dependent
// global loads behind every grid barrier.
struct PhaseCache {
  Job job;
  EpiLocal epi;
  int valid;
  int pad_[3];
};

// grid barrier wait (defined with the persistent-kernel helpers below)
__device__ __forceinline__ void grid_wait_ext(const unsigned int* ctr, unsigned int target);

// ---------------------------------------------------------- tensor-core pipeline
// Shared-memory carve-up and the role-private running state.  The three role loops below are shared by
// the one-launch-per-phase kernel (job_kernel_tc) and the persistent scan kernels (kernels.cuh), which keep
// the pipeline state alive across phases and ticks.
struct Pipe {
  uint8_t* tiles;
  uint64_t *full_bar, *empty_bar, *tfull_bar, *tempty_bar;
  uint32_t tmem_base;
  volatile uint32_t* split_flag;
  int nstages, n_cols;
  uint32_t a_bytes, b_bytes, stage_bytes;
  int stage;       // producer / MMA: ring position
  uint32_t phase;  // producer / MMA: ring parity
  int it;          // MMA / epilogue: accumulator uses so far
  uint8_t* stg;    // epilogue: operand staging region of the quad finish (persistent kernels) or null
  int stg_bytes;
  uint8_t* cache_area;   // >= 1536 bytes behind the barriers (PhaseCache copies of the persistent kernels)
  uint32_t acc_stride;   // TMEM columns between the two accumulator buffers (256; grouped scans: the widest n_cols)
};

// returns the first byte after the pipeline's shared memory (1024-aligned ring + barriers)
__device__ __forceinline__ uint8_t* pipe_setup(Pipe& p, uint8_t* smem, int n_cols) {
  const int warp = threadIdx.x >> 5;
  p.n_cols = n_cols;
  p.a_bytes = TILE_M * KB * 2;  // 16 KB
  p.b_bytes = (uint32_t)n_cols * KB * 2;
  p.stage_bytes = 2 * p.a_bytes + 2 * p.b_bytes;
  p.nstages = (SMEM_BYTES - 2048) / (int)p.stage_bytes;
  if (p.nstages > 8) p.nstages = 8;
  p.tiles = smem;
  p.full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
  p.empty_bar = p.full_bar + 8;
  p.tfull_bar = p.empty_bar + 8;    // [2] accumulator ready
  p.tempty_bar = p.tfull_bar + 2;   // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p.tempty_bar + 2);
  p.split_flag = tmem_slot + 1;
  p.stage = 0; p.phase = 0; p.it = 0;
  p.acc_stride = 256u;
  p.stg = nullptr; p.stg_bytes = 0;
  p.cache_area = reinterpret_cast<uint8_t*>(p.full_bar) + 256;   // the ring leaves >= 2048 bytes here
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(&p.full_bar[s], 1);
      mbar_init(&p.empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&p.tfull_bar[s], 1);
      mbar_init(&p.tempty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  p.tmem_base = *tmem_slot;
  return smem + SMEM_BYTES;
}

__device__ __forceinline__ void pipe_teardown(Pipe& p) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) tmem_dealloc(p.tmem_base, 512);
}

// ------------------------------------------------ TMA producer (one thread: warp 0, lane 0)
// Every operand is a (hi, lo) pair of bf16 planes at a constant distance; the tensor map of a pair has one extra
// "plane" dimension of extent 2, so ONE cp.async.bulk.tensor brings the hi tile and, right behind it in shared memory,
// the lo tile (the single producer thread issues half as many instructions; each costs ~0.5 us of issue latency).
__device__ __forceinline__ void producer_load_a(Pipe& p, const Seg& sg, int kb, const CUtensorMap* ma, uint8_t* st,
                                                uint64_t* fb, uint64_t pol_keep) {
  if (sg.a_nkb > 0) {
    // tile-contiguous weight pack: one 16 KB contiguous block per plane, kept in L2 (re-read every step)
    const int trow = ((sg.a_row >> 7) * sg.a_nkb + (sg.a_k >> 6) + kb) << 7;
    if (pol_keep) tma_load_3d_hint(st, ma, fb, 0, trow, 0, pol_keep);
    else tma_load_3d(st, ma, fb, 0, trow, 0);   // no per-instruction hint: the stream's access-policy window decides
  } else {
    tma_load_3d(st, ma, fb, sg.a_k + kb * KB, sg.a_row, 0);
  }
}
__device__ __forceinline__ void producer_load_b(Pipe& p, const Seg& sg, int kb, int t, const CUtensorMap* mb,
                                                uint8_t* st, uint64_t* fb, int n_shift) {
  if (sg.b_slot == NO_SLOT) tma_load_3d(st + 2 * p.a_bytes, mb, fb, sg.b_k + kb * KB, sg.b_row + n_shift, 0);
  else tma_load_4d(st + 2 * p.a_bytes, mb, fb, sg.b_k + kb * KB, sg.b_row, t + sg.b_slot, 0);
}

// gridbar != nullptr (persistent kernels): the activation (B) operands of this phase are produced by other CTAs
// in the previous phase, so their loads must wait for grid barrier `target`; the WEIGHT (A) tiles do not depend
// on it.  The producer therefore fills the free ring slots with weight tiles first (expect_tx without arrive),
// waits for the barrier, and completes those slots with the activation tiles (arrive + expect_tx).
__device__ __forceinline__ void producer_run(Pipe& p, const EngineParams& P, int tick,
                                             const unsigned int* gridbar = nullptr, unsigned int target = 0,
                                             const PhaseCache* pc = nullptr) {
  // debug_flags bit 0: no per-instruction L2 hint on the weight tiles; bits 1-2: evict_last on a fraction of lines
  uint64_t pol_keep = 0;
  if (!(P.debug_flags & 1)) {
    const int fr = (P.debug_flags >> 1) & 3;
    pol_keep = fr == 0 ? l2_policy_evict_last() : (fr == 1 ? l2_policy_evict_last_075() : l2_policy_evict_last_050());
  }
  bool waited = (gridbar == nullptr) || target == 0;
  const uint32_t tx_bytes = 2 * p.a_bytes + 2 * p.b_bytes;   // <= stage_bytes (the ring is sized for the widest phase)
  for (int j = cta_rank(P); j < P.njobs; j += P.ncta) {
    const Job& jb = pc ? pc->job : P.jobs[j];
    const int t = job_time(P, jb, tick);
    const int total_kb = job_total_kb(P, jb, t);
    if (total_kb == 0) continue;
    int klo, khi;
    job_kb_range(jb, total_kb, klo, khi);
    int n_shift = 0;
    if (jb.epi == EPI_PLAIN && P.chunk_samples > 0) { int lim; chunk_window(P, jb, t, n_shift, lim); }
    if (!waited) {
      // descriptors of ALL operands (they are static; only the activation DATA waits for the barrier): fetched
      // together now instead of one L2 round trip in front of every first use after the barrier
      for (int s = 0; s < jb.nseg; ++s) {
        const Seg& sg = jb.seg[s];
        tma_prefetch_desc(P.maps + sg.a_map);
        tma_prefetch_desc(P.maps + sg.b_map);
      }
    }
    int early = 0;   // k blocks of this job whose weight tiles were issued before the barrier
    if (!waited) {
      // ---- pass 1: weight tiles of the first min(nstages, khi - klo) k blocks
      int kidx = 0, st_i = p.stage;
      uint32_t ph = p.phase;
      for (int s = 0; s < jb.nseg && early < p.nstages; ++s) {
        const Seg sg = jb.seg[s];
        if (!seg_valid(P, sg, t)) continue;
        for (int kb = 0; kb < sg.nkb && early < p.nstages; ++kb, ++kidx) {
          if (kidx < klo || kidx >= khi) continue;
          mbar_wait(&p.empty_bar[st_i], ph ^ 1);
          uint8_t* st = p.tiles + (size_t)st_i * p.stage_bytes;
          mbar_expect_tx_only(&p.full_bar[st_i], 2 * p.a_bytes);
          producer_load_a(p, sg, kb, P.maps + sg.a_map, st, &p.full_bar[st_i], pol_keep);
          ++early;
          if (++st_i == p.nstages) { st_i = 0; ph ^= 1; }
        }
      }
      grid_wait_ext(gridbar, target);
      TL(10);
      waited = true;
    }
    // ---- pass 2: everything else, in ring order
    int kidx = 0, done = 0;
    for (int s = 0; s < jb.nseg; ++s) {
      const Seg sg = jb.seg[s];
      if (!seg_valid(P, sg, t)) continue;
      const CUtensorMap* ma = P.maps + sg.a_map;
      const CUtensorMap* mb = P.maps + sg.b_map;
      for (int kb = 0; kb < sg.nkb; ++kb, ++kidx) {
        if (kidx < klo || kidx >= khi) continue;
        uint8_t* st = p.tiles + (size_t)p.stage * p.stage_bytes;
        uint64_t* fb = &p.full_bar[p.stage];
        if (done < early) {
          mbar_expect_tx(fb, 2 * p.b_bytes);           // arrive + the activation bytes
          producer_load_b(p, sg, kb, t, mb, st, fb, n_shift);
        } else {
          mbar_wait(&p.empty_bar[p.stage], p.phase ^ 1);
          mbar_expect_tx(fb, tx_bytes);
          producer_load_a(p, sg, kb, ma, st, fb, pol_keep);
          producer_load_b(p, sg, kb, t, mb, st, fb, n_shift);
        }
        ++done;
        if (++p.stage == p.nstages) { p.stage = 0; p.phase ^= 1; }
      }
    }
  }
  if (!waited) grid_wait_ext(gridbar, target);   // no job this phase: still observe the barrier
  TL(2);
}

// ------------------------------------------------ MMA issuer (warp 1; lane 0 issues)
__device__ __forceinline__ void mma_run(Pipe& p, const EngineParams& P, int tick, const PhaseCache* pc = nullptr) {
  const int lane = threadIdx.x & 31;
  const uint32_t idesc = umma_idesc_bf16(TILE_M, p.n_cols);
  for (int j = cta_rank(P); j < P.njobs; j += P.ncta) {
    const Job& jb = pc ? pc->job : P.jobs[j];
    const int t = job_time(P, jb, tick);
    const int all_kb = job_total_kb(P, jb, t);
    if (all_kb == 0) continue;
    int klo, khi;
    job_kb_range(jb, all_kb, klo, khi);
    const int total_kb = khi - klo;
    if (total_kb == 0) continue;  // empty split part: no accumulator is used, the epilogue contributes zeros
    const int buf = p.it & 1;
    const uint32_t use = (uint32_t)(p.it >> 1);
    mbar_wait(&p.tempty_bar[buf], (use & 1) ^ 1);
    tc_fence_after();
    const uint32_t tmem_d = p.tmem_base + (uint32_t)buf * p.acc_stride;
    for (int kbi = 0; kbi < total_kb; ++kbi) {
      mbar_wait(&p.full_bar[p.stage], p.phase);
      tc_fence_after();
      if (lane == 0 && kbi == 0) TL(11);
      if (lane == 0) {
        const uint8_t* st = p.tiles + (size_t)p.stage * p.stage_bytes;
        const uint64_t da_hi = umma_desc_sw128(st);
        const uint64_t da_lo = umma_desc_sw128(st + p.a_bytes);
        const uint64_t db_hi = umma_desc_sw128(st + 2 * p.a_bytes);
        const uint64_t db_lo = umma_desc_sw128(st + 2 * p.a_bytes + p.b_bytes);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);  // 32 bytes per k16 step
          umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, (kbi | k) != 0);
          umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, 1);
          umma_bf16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1);
        }
        umma_commit(&p.empty_bar[p.stage]);                        // frees the smem slot
        if (kbi == total_kb - 1) umma_commit(&p.tfull_bar[buf]);   // accumulator complete
      }
      __syncwarp();
      if (++p.stage == p.nstages) { p.stage = 0; p.phase ^= 1; }
    }
    ++p.it;
  }
  if (lane == 0) TL(3);
}

// ------------------------------------------------ scan phases of the persistent kernels: ONE cached job per CTA
// k-block order of a part's range [klo, khi): first the STREAMED blocks (weight tile through the TMA ring), then the
// RESIDENT blocks (weight tile in tensor memory since kernel start, only the activation tile is loaded).
__device__ __forceinline__ int job_full_kb(const Job& jb) {
  int n = 0;
  for (int s = 0; s < jb.nseg; ++s) n += jb.seg[s].nkb;
  return n;
}
__device__ __forceinline__ void kb_locate(const Job& jb, int kidx, int& seg, int& kb) {
  seg = 0; kb = kidx;
  for (int s = 0; s < jb.nseg; ++s) {
    if (kb < jb.seg[s].nkb) { seg = s; return; }
    kb -= jb.seg[s].nkb;
  }
}
// resident blocks of this part at time t (0 when a segment is out of range: the range no longer matches the tiles
// that were loaded at kernel start -- cannot happen in the grouped tables, every segment is valid for t in [0, T))
__device__ __forceinline__ int job_resident(const Job& jb, int total_kb, int n) {
  if (jb.res_kb <= 0 || total_kb != job_full_kb(jb)) return 0;
  return jb.res_kb < n ? jb.res_kb : n;
}

// weight tiles of the resident blocks: global (tile-contiguous pack, 128 B per row and plane) -> tensor memory.
// Called by all threads once per kernel and job; warps 2..5 do the work (one TMEM lane quarter each).
__device__ __forceinline__ void resident_preload(Pipe& p, const EngineParams& P, const PhaseCache* pc) {
  if (cta_rank(P) < 0 || cta_rank(P) >= P.njobs) return;
  const Job& jb = pc->job;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (jb.res_kb <= 0 || warp < 2 || warp >= 6) return;
  int klo, khi;
  job_kb_range(jb, job_full_kb(jb), klo, khi);
  const int n = khi - klo, nres = jb.res_kb < n ? jb.res_kb : n;
  const int q = warp & 3, row = q * 32 + lane;
  for (int i = 0; i < nres; ++i) {
    int seg, kb;
    kb_locate(jb, klo + (n - nres) + i, seg, kb);
    const Seg& sg = jb.seg[seg];
    const long long trow = ((long long)(sg.a_row >> 7) * sg.a_nkb + (sg.a_k >> 6) + kb) << 7;
#pragma unroll 1
    for (int w = 0; w < 2; ++w) {
      const uint4* src = reinterpret_cast<const uint4*>(P.raws[sg.a_map + w].base + (trow + row) * KB);
      uint32_t r[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = __ldg(src + c);
        r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
      }
      tmem_st_32x32(p.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jb.res_col + i * 64 + w * 32), r);
    }
  }
  tmem_st_wait();
}

__device__ __forceinline__ void producer_scan(Pipe& p, const EngineParams& P, int tick, const PhaseCache* pc,
                                              const unsigned int* gridbar, unsigned int target, bool prefetch) {
  const uint64_t pol_keep = (P.debug_flags & 1) ? 0 : l2_policy_evict_last();
  bool waited = (gridbar == nullptr) || target == 0;
  const uint32_t tx_bytes = 2 * p.a_bytes + 2 * p.b_bytes;
  if (cta_rank(P) < P.njobs) {
    const Job& jb = pc->job;
    const int t = job_time(P, jb, tick);
    const int total_kb = job_total_kb(P, jb, t);
    if (total_kb > 0) {
      int klo, khi;
      job_kb_range(jb, total_kb, klo, khi);
      const int n = khi - klo, nres = job_resident(jb, total_kb, n), nstream = n - nres;
      // flattened (valid segment, block) sequence of this part: seg / kb of block klo + i
      auto locate = [&](int i, int& sgi, int& kb) {
        int kidx = klo + i;
        sgi = 0; kb = 0;
        for (int s = 0; s < jb.nseg; ++s) {
          if (!seg_valid(P, jb.seg[s], t)) continue;
          if (kidx < jb.seg[s].nkb) { sgi = s; kb = kidx; return; }
          kidx -= jb.seg[s].nkb;
        }
      };
      int early = 0;
      if (!waited) {
        for (int s = 0; s < jb.nseg; ++s) {
          tma_prefetch_desc(P.maps + jb.seg[s].a_map);
          tma_prefetch_desc(P.maps + jb.seg[s].b_map);
        }
        if (prefetch) {
          // weight tiles of the first streamed blocks do not depend on the barrier: issue them now (expect_tx without
          // arrive), complete the slots with the activation tiles afterwards
          int st_i = p.stage;
          uint32_t ph = p.phase;
          for (; early < nstream && early < p.nstages; ++early) {
            int sgi, kb;
            locate(early, sgi, kb);
            const Seg& sg = jb.seg[sgi];
            mbar_wait(&p.empty_bar[st_i], ph ^ 1);
            uint8_t* st = p.tiles + (size_t)st_i * p.stage_bytes;
            mbar_expect_tx_only(&p.full_bar[st_i], 2 * p.a_bytes);
            producer_load_a(p, sg, kb, P.maps + sg.a_map, st, &p.full_bar[st_i], pol_keep);
            if (++st_i == p.nstages) { st_i = 0; ph ^= 1; }
          }
        }
        grid_wait_ext(gridbar, target);
        TL(10);
        waited = true;
      }
      for (int i = 0; i < n; ++i) {
        int sgi, kb;
        locate(i, sgi, kb);
        const Seg& sg = jb.seg[sgi];
        uint8_t* st = p.tiles + (size_t)p.stage * p.stage_bytes;
        uint64_t* fb = &p.full_bar[p.stage];
        if (i < early) {
          mbar_expect_tx(fb, 2 * p.b_bytes);
          producer_load_b(p, sg, kb, t, P.maps + sg.b_map, st, fb, 0);
        } else {
          mbar_wait(&p.empty_bar[p.stage], p.phase ^ 1);
          if (i < nstream) {
            mbar_expect_tx(fb, tx_bytes);
            producer_load_a(p, sg, kb, P.maps + sg.a_map, st, fb, pol_keep);
          } else {
            mbar_expect_tx(fb, 2 * p.b_bytes);
          }
          producer_load_b(p, sg, kb, t, P.maps + sg.b_map, st, fb, 0);
        }
        if (++p.stage == p.nstages) { p.stage = 0; p.phase ^= 1; }
      }
    }
  }
  if (!waited) grid_wait_ext(gridbar, target);
  TL(2);
}

__device__ __forceinline__ void mma_scan(Pipe& p, const EngineParams& P, int tick, const PhaseCache* pc) {
  const int lane = threadIdx.x & 31;
  if (cta_rank(P) >= P.njobs) return;
  const uint32_t idesc = umma_idesc_bf16(TILE_M, p.n_cols);
  const Job& jb = pc->job;
  const int t = job_time(P, jb, tick);
  const int all_kb = job_total_kb(P, jb, t);
  if (all_kb == 0) return;
  int klo, khi;
  job_kb_range(jb, all_kb, klo, khi);
  const int n = khi - klo;
  if (n == 0) return;
  const int nres = job_resident(jb, all_kb, n), nstream = n - nres;
  const int buf = p.it & 1;
  const uint32_t use = (uint32_t)(p.it >> 1);
  mbar_wait(&p.tempty_bar[buf], (use & 1) ^ 1);
  tc_fence_after();
  const uint32_t tmem_d = p.tmem_base + (uint32_t)buf * p.acc_stride;
  for (int kbi = 0; kbi < n; ++kbi) {
    mbar_wait(&p.full_bar[p.stage], p.phase);
    tc_fence_after();
    if (lane == 0 && kbi == 0) TL(11);
    if (lane == 0) {
      const uint8_t* st = p.tiles + (size_t)p.stage * p.stage_bytes;
      const uint64_t db_hi = umma_desc_sw128(st + 2 * p.a_bytes);
      const uint64_t db_lo = umma_desc_sw128(st + 2 * p.a_bytes + p.b_bytes);
      if (kbi < nstream) {
        const uint64_t da_hi = umma_desc_sw128(st);
        const uint64_t da_lo = umma_desc_sw128(st + p.a_bytes);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
          umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, (kbi | k) != 0);
          umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, 1);
          umma_bf16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1);
        }
      } else {
        const uint32_t ta_hi = p.tmem_base + (uint32_t)(jb.res_col + (kbi - nstream) * 64), ta_lo = ta_hi + 32;
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
          umma_bf16_ts(tmem_d, ta_hi + k * 8, db_hi + adv, idesc, (kbi | k) != 0);
          umma_bf16_ts(tmem_d, ta_lo + k * 8, db_hi + adv, idesc, 1);
          umma_bf16_ts(tmem_d, ta_hi + k * 8, db_lo + adv, idesc, 1);
        }
      }
      umma_commit(&p.empty_bar[p.stage]);
      if (kbi == n - 1) umma_commit(&p.tfull_bar[buf]);
    }
    __syncwarp();
    if (++p.stage == p.nstages) { p.stage = 0; p.phase ^= 1; }
  }
  ++p.it;
  if (lane == 0) TL(3);
}

#ifndef PB_SPLIT_BATCH
#define PB_SPLIT_BATCH 3   // partial tiles summed in batches of 3 parts: 111.5 vs 115.0 ms/step unbatched (8 warps)
#endif
#ifndef PB_SPLIT_ITEM_W
#define PB_SPLIT_ITEM_W 4
#endif
#ifndef PB_PART_W
#define PB_PART_W 16   // columns per tcgen05.ld when a split-K partial tile is parked
#endif
#ifndef PB_QMAX
#define PB_QMAX 4     // columns per warp whose loads are in flight together in the quad split-K finish
#endif
#ifndef PB_TMEM_W
#define PB_TMEM_W 8   // columns per tcgen05.ld chunk of the unsplit (TMEM) epilogue path
#endif
// Unsplit tile straight out of tensor memory, 8 columns per step.  The memory operands of step i + 1 (old value for
// PF_ACC, row bias) are requested BEFORE step i is applied, so a tile costs one exposed L2 round trip instead of one
// per step (measured: 31 us per 128 x 128 tile with 8-column steps whose loads were issued and consumed in place).
template <class LOAD, class APPLY>
__device__ __forceinline__ void tmem_tile_pipelined(uint32_t taddr, int n_cols, LOAD load, APPLY apply) {
  EpiOps<8> ops[2];
  load(0, ops[0]);
#pragma unroll 1
  for (int n0 = 0; n0 < n_cols; n0 += 16) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = n0 + 8 * h;
      if (c0 < n_cols) {
        float v[8];
        tmem_ld_32x8(taddr + c0, v);
        if (c0 + 8 < n_cols) load(c0 + 8, ops[h ^ 1]);
        tmem_ld_wait();
        apply(c0, v, ops[h]);
      }
    }
  }
}

// ---- lean plain epilogue (EPI_PLAIN, not transposed): chunk products of the persistent scans, readout / dgrad /
// hoisted products outside them.  The generic epi_plain keeps its context in a 150-byte struct that the 255-register
// kernels hold in local memory: every field use was a dependent local load, 2.8 k cycles per 8-column step, 38 us per
// 128 x 128 tile (measured; the MMAs of the same tile take 9 us).  Here the loop state is a cursor of pointers that
// advance by constant strides, and the memory operands of step i + 1 (old value for PF_ACC, row bias) are requested
// before step i is applied.
struct PlainK {
  long long ldo, ldp, rb_ld;
  const float* pb0;
  bf16 *hi, *lo;
  int n_pad, n_valid, n_total, flags;
  float scale, bias;
  bool rb, acc, has_out;
};
struct PlainCur { float* po; const float* pb; long long pl; int r, n; };
__device__ __forceinline__ void plain_setup(const EpiLocal& E, int t, int row, PlainK& k, PlainCur& c) {
  const int f = E.row0 + row;
  k.ldo = E.l0; k.ldp = E.l1; k.rb_ld = E.rowbias_ld;
  k.hi = (bf16*)E.p2; k.lo = (bf16*)E.p3;
  k.n_pad = E.n_pad > 0 ? E.n_pad : (1 << 30);
  k.n_valid = E.n_pad > 0 ? E.n_valid : (1 << 30);
  k.n_total = E.n_total; k.flags = E.flags; k.scale = E.scale;
  k.rb = E.rowbias_ld > 0; k.acc = (E.flags & PF_ACC) != 0; k.has_out = E.p0 != nullptr;
  const float* biasp = (const float*)E.p1;
  k.bias = (biasp && !k.rb && row < E.m_valid) ? __ldg(biasp + f) : 0.0f;
  k.pb0 = k.rb ? biasp + f : nullptr;
  const int n = E.n0;
  int q = 0, r = n;
  if (E.n_pad > 0) { q = n / E.n_pad; r = n - q * E.n_pad; }
  const long long srow = (long long)q * (E.n_pad > 0 ? E.n_valid : 0) + (r < k.n_valid ? r : k.n_valid) +
                         (E.n_pad > 0 ? 0 : 0);
  c.n = n; c.r = r;
  c.po = (float*)E.p0 + (long long)t * E.l2 + (E.n_pad > 0 ? srow : (long long)n) * k.ldo + f;
  c.pb = k.rb ? k.pb0 + (long long)r * k.rb_ld : nullptr;
  c.pl = ((E.flags & PF_PLANE_PADDED) ? (long long)n : (E.n_pad > 0 ? srow : (long long)n)) * k.ldp + f;
}
__device__ __forceinline__ void plain_next(const PlainK& k, PlainCur& c) {
  const bool valid = c.r < k.n_valid;
  if (valid) c.po += k.ldo;
  if (valid || (k.flags & PF_PLANE_PADDED)) c.pl += k.ldp;
  ++c.n;
  if (++c.r == k.n_pad) { c.r = 0; c.pb = k.pb0; }
  else if (k.rb) c.pb += k.rb_ld;
}
template <int W>
__device__ __forceinline__ void plain_load(const PlainK& k, PlainCur& c, bool row_ok, float* a, float* b) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const bool live = row_ok && c.n < k.n_total && c.r < k.n_valid;
    a[j] = (live && k.acc && k.has_out) ? __ldcg(c.po) : 0.0f;
    b[j] = (live && k.rb) ? __ldg(c.pb) : 0.0f;
    plain_next(k, c);
  }
}
template <int W>
__device__ __forceinline__ void plain_apply(const PlainK& k, PlainCur& c, bool row_ok, const float* v, const float* a,
                                            const float* b) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const bool live = row_ok && c.n < k.n_total && c.r < k.n_valid;
    if (live) {
      const float y = v[j] * k.scale + k.bias + b[j] + a[j];
      if (k.has_out) *c.po = y;
      if (k.hi) {
        bf16 h, l;
        split_bf16(y, h, l);
        k.hi[c.pl] = h;
        k.lo[c.pl] = l;
      }
    }
    plain_next(k, c);
  }
}
// one unsplit plain tile out of tensor memory (thread <-> output row)
__device__ __forceinline__ void plain_tile(const EpiLocal& E, int t, int row, uint32_t taddr, int n_cols) {
  PlainK k;
  PlainCur cl, ca;
  plain_setup(E, t, row, k, cl);
  ca = cl;
  const bool row_ok = row < E.m_valid;
  float a[2][8], b[2][8];
  plain_load<8>(k, cl, row_ok, a[0], b[0]);
#pragma unroll 1
  for (int n0 = 0; n0 < n_cols; n0 += 16) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = n0 + 8 * h;
      if (c0 < n_cols) {
        float v[8];
        tmem_ld_32x8(taddr + c0, v);
        if (c0 + 8 < n_cols) plain_load<8>(k, cl, row_ok, a[h ^ 1], b[h ^ 1]);
        tmem_ld_wait();
        plain_apply<8>(k, ca, row_ok, v, a[h], b[h]);
      }
    }
  }
}

// ------------------------------------------------ epilogue (warps 2..5 read TMEM; warps 6.. help after split-K)
// All epilogue warps run this loop.  Unsplit jobs: warps 2..5 read the accumulator (one TMEM lane quarter each) and
// apply the epilogue; the helpers skip.  Split jobs: warps 2..5 park the partial tile in scratch, then ALL ten
// warps share the post-reduction epilogue of this part's column slice as 4-column work items (the reduction
// reads global memory, so any warp can do it; the helpers hide load / MUFU latencies).
template <bool CACHED>
__device__ __forceinline__ void epilogue_run(Pipe& p, const EngineParams& P, int tick, const PhaseCache* pc = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = p.n_cols;
  const bool tmem_warp = warp < 6;
  const int q = warp & 3;  // TMEM lane quarter this warp may access
  const int row = q * 32 + lane;
  const int gtid = threadIdx.x - 64;   // 0..319 inside the epilogue group
  for (int j = cta_rank(P); j < P.njobs; j += P.ncta) {
    const Job& jb = CACHED ? pc->job : P.jobs[j];
    const int t = job_time(P, jb, tick);
    const int all_kb = job_total_kb(P, jb, t);
    if (all_kb == 0) continue;
    int klo, khi;
    job_kb_range(jb, all_kb, klo, khi);
    const bool have_acc = khi > klo;
    const int buf = p.it & 1;
    const uint32_t use = (uint32_t)(p.it >> 1);
    int cw_shift = 0, cw_limit = 0;
    if (!CACHED && jb.epi == EPI_PLAIN && P.chunk_samples > 0) chunk_window(P, jb, t, cw_shift, cw_limit);
    // CACHED: the epilogue context sits in shared memory (fields are read with LDS when needed: no registers held,
    // no dependent global loads); otherwise it is rebuilt from the job table / scan context in registers
    const EpiLocal Ereg = CACHED ? EpiLocal() : make_epi_local(jb, P.ctx, cw_shift, cw_limit);
    const EpiLocal& E = CACHED ? pc->epi : Ereg;
    const int ksplit = jb.ksplit, kpart = jb.kpart, group = jb.group;
    // Quad finish (cooperative split-K only): this part finishes columns [qc_lo, qc_hi) of the tile; warp ew takes
    // columns qc_lo + ew, + NEW, ...; lane = row quad.  The epilogue operands are requested NOW (cp.async into the
    // staging region), before the accumulator is complete: they only depend on the previous phase (already behind
    // the grid barrier), so their latency hides under the MMAs and the partial-tile exchange.
    const bool quad = ksplit > 1 && P.coop_epilogue && quad_ok(E);
    constexpr int NEW = EPI_GROUP_THREADS / 32;
    const int ew = warp - 2;
    const int qc_lo = (n_cols * kpart) / ksplit, qc_hi = (n_cols * (kpart + 1)) / ksplit;
    const int stg_cols = p.stg ? p.stg_bytes / (q_narr(E) * TILE_M * 4) : 0;
    if (quad) q_stage(E, t, lane, ew, NEW, qc_lo, qc_hi, p.stg, stg_cols);
    if (threadIdx.x == 64) TL(9);
    if (tmem_warp) {
      const uint32_t taddr = p.tmem_base + (uint32_t)buf * p.acc_stride + ((uint32_t)(q * 32) << 16);
      if (have_acc) {
        mbar_wait(&p.tfull_bar[buf], use & 1);
        tc_fence_after();
      }
      if (threadIdx.x == 64) TL(4);
      if (ksplit <= 1 && E.epi == EPI_PLAIN && !(E.flags & PF_TRANS)) {
        plain_tile(E, t, row, taddr, n_cols);
      } else if (ksplit <= 1) {
        tmem_tile_pipelined(
            taddr, n_cols, [&](int c0, EpiOps<8>& o) { epilogue_load<8>(E, t, row, c0, 8, o); },
            [&](int c0, const float* v, const EpiOps<8>& o) { epilogue_apply<8>(E, t, row, c0, 8, v, o); });
      } else {
        // park the partial tile: scratch[group][part][col][row]  (row fastest -> coalesced)
        float* part = P.split_scratch + ((size_t)group * MAX_KSPLIT + kpart) * (size_t)n_cols * TILE_M;
#if PB_PART_W == 32
        for (int n0 = 0; n0 < n_cols; n0 += 32) {
          float v[32];
          if (have_acc) {
            tmem_ld_32x32(taddr + n0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.0f;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (n0 + i < n_cols) __stcg(part + (size_t)(n0 + i) * TILE_M + row, v[i]);
        }
#else
        for (int n0 = 0; n0 < n_cols; n0 += 16) {
          float v[16];
          if (have_acc) {
            tmem_ld_32x16(taddr + n0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) __stcg(part + (size_t)(n0 + i) * TILE_M + row, v[i]);
        }
#endif
      }
      if (have_acc) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p.tempty_bar[buf]);
      }
      if (threadIdx.x == 64) TL(5);
    }
    if (ksplit > 1) {
      if (tmem_warp) __threadfence();
      epi_group_sync();   // partial tile written by warps 2..5
      const float* base = P.split_scratch + (size_t)group * MAX_KSPLIT * (size_t)n_cols * TILE_M;
      int c_lo = 0, c_hi = 0;
      if (P.coop_epilogue) {
        // Every part of the tile finishes a slice of its columns once all parts have arrived.  Safe because
        // the launch has at most one job per CTA and all CTAs are co-resident (grid <= SM count).
        if (warp == 2 && lane == 0) {
          atomicAdd(P.split_count + group, 1u);
          unsigned int seen, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.split_count + group) : "memory");
            if (++spins > (1u << 24)) pb_timeout(2);
          } while ((seen & 0xffffu) < (unsigned int)ksplit);
        }
        epi_group_sync();
        if (threadIdx.x == 64) TL(6);
        c_lo = (n_cols * kpart) / ksplit;
        c_hi = (n_cols * (kpart + 1)) / ksplit;
      } else {
        if (warp == 2 && lane == 0) {
          const unsigned int old = atomicAdd(P.split_count + group, 1u);
          const bool last = old == (unsigned int)(ksplit - 1);
          if (last) P.split_count[group] = 0u;   // ready for the next launch
          *p.split_flag = last ? 1u : 0u;
        }
        epi_group_sync();
        if (threadIdx.x == 64) TL(6);
        if (*p.split_flag) { c_lo = 0; c_hi = n_cols; }
      }
      if (quad) {
        q_finish(E, t, lane, ew, NEW, qc_lo, qc_hi, ksplit, n_cols, base, p.stg, stg_cols);
      } else {
      __threadfence();
      // work item = (row, group of 4 columns); consecutive threads take consecutive rows (coalesced).  All loads of
      // an item (partial tiles + epilogue operands) are issued before the first use.  (8-column items halve the
      // number of latency rounds but measured slower: 118.6 vs 115.4 ms/step even without spills.)
      constexpr int IW = PB_SPLIT_ITEM_W;   // columns per work item
      const int ngroups = (c_hi - c_lo + IW - 1) / IW;
      for (int e = gtid; e < ngroups * TILE_M; e += EPI_GROUP_THREADS) {
        const int r = e & (TILE_M - 1);
        const int n0 = c_lo + (e >> 7) * IW;
        const int nc = min(IW, c_hi - n0);
        float v[IW];
#pragma unroll
        for (int i = 0; i < IW; ++i) v[i] = 0.0f;
        EpiOps<IW> ops;
#if PB_SPLIT_BATCH
        // partial tiles in batches of PB_SPLIT_BATCH parts: bounds the registers of the loads in flight (a spilled
        // load result serialises the L2 round trips); the second batch only exists for tiles split more than that
        epilogue_load<IW>(E, t, r, n0, nc, ops);
#pragma unroll
        for (int p0 = 0; p0 < MAX_KSPLIT; p0 += PB_SPLIT_BATCH) {
          if (p0 < ksplit) {
            float x[PB_SPLIT_BATCH][IW];
#pragma unroll
            for (int q = 0; q < PB_SPLIT_BATCH; ++q) {
              const float* src = base + (size_t)(p0 + q) * n_cols * TILE_M;
#pragma unroll
              for (int i = 0; i < IW; ++i)
                x[q][i] = (p0 + q < ksplit && i < nc) ? __ldcg(src + (size_t)(n0 + i) * TILE_M + r) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < PB_SPLIT_BATCH; ++q)   // part order: deterministic
#pragma unroll
              for (int i = 0; i < IW; ++i) v[i] += x[q][i];
          }
        }
#else
        float x[MAX_KSPLIT][IW];
#pragma unroll
        for (int pp = 0; pp < MAX_KSPLIT; ++pp) {
          const float* src = base + (size_t)pp * n_cols * TILE_M;
#pragma unroll
          for (int i = 0; i < IW; ++i)
            x[pp][i] = (pp < ksplit && i < nc) ? __ldcg(src + (size_t)(n0 + i) * TILE_M + r) : 0.0f;
        }
        epilogue_load<IW>(E, t, r, n0, nc, ops);   // operand loads in flight together with the partial-tile loads
#pragma unroll
        for (int pp = 0; pp < MAX_KSPLIT; ++pp)   // part order: deterministic
#pragma unroll
          for (int i = 0; i < IW; ++i) v[i] += x[pp][i];
#endif
        epilogue_apply<IW>(E, t, r, n0, nc, v, ops);
      }
      }   // !quad
      if (threadIdx.x == 64) TL(7);
      epi_group_sync();
      if (P.coop_epilogue && warp == 2 && lane == 0) {
        // the last part to finish resets the arrival counter for the next use
        const unsigned int old = atomicAdd(P.split_count + group, 0x10000u);
        if ((old >> 16) == (unsigned int)(ksplit - 1)) P.split_count[group] = 0u;
      }
    }
    if (have_acc) ++p.it;   // accumulator buffers advance only when one was used (mirrors the MMA warp)
  }
}

// ------------------------------------------------ compact epilogues of the persistent kernels
// The generic epilogue_run above carries every job kind (five epilogues x split / unsplit x thread-per-row paths) and
// compiles to ~14 K instructions per inlined copy; three copies made the persistent scan kernels ~750 KB of SASS, far
// beyond the instruction caches, and every phase of every tick started on cold code.  The persistent kernels
// therefore use two specialised epilogues: scan phases (one cached job per CTA, always the quad split-K finish --
// an unsplit job is simply a 1-part split) and chunk phases (plain unsplit jobs only).
// DIR: 1 forward scan (GATES / CAND), 2 backward scan (BWD_RH / BWD_STATE).
template <int DIR>
__device__ __forceinline__ void epilogue_scan(Pipe& p, const EngineParams& P, int tick, const PhaseCache* pc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = p.n_cols;
  if (cta_rank(P) >= P.njobs) return;
  const Job& jb = pc->job;
  const int t = job_time(P, jb, tick);
  const int all_kb = job_total_kb(P, jb, t);
  if (all_kb == 0) return;
  int klo, khi;
  job_kb_range(jb, all_kb, klo, khi);
  const bool have_acc = khi > klo;
  const int buf = p.it & 1;
  const uint32_t use = (uint32_t)(p.it >> 1);
  // forward sweep: context copied into registers (31.0 vs 31.4 ms per forward scan); backward sweep: read from shared
  // memory field by field -- the copy costs 240 B of extra spills there and 3 ms (same-box A/B, tools/ab_variants.sh)
  EpiLocal Ereg;
  if (DIR == 1) Ereg = pc->epi;
  const EpiLocal& E = DIR == 1 ? Ereg : pc->epi;
  const int ksplit = jb.ksplit, kpart = jb.kpart, group = jb.group;
  constexpr int NEW = EPI_GROUP_THREADS / 32;
  const int ew = warp - 2;
  const int c_lo = (n_cols * kpart) / ksplit, c_hi = (n_cols * (kpart + 1)) / ksplit;
  const int stg_cols = p.stg_bytes / (q_narr<DIR>(E) * TILE_M * 4);
  // operands of this part's columns: global -> shared, asynchronously, as soon as the grid barrier is behind us
  q_stage<DIR>(E, t, lane, ew, NEW, c_lo, c_hi, p.stg, stg_cols);
  if (threadIdx.x == 64) TL(9);
  const float* base = P.split_scratch + (size_t)group * MAX_KSPLIT * (size_t)n_cols * TILE_M;
  if (warp < 6) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t taddr = p.tmem_base + (uint32_t)buf * p.acc_stride + ((uint32_t)(q * 32) << 16);
    if (have_acc) {
      mbar_wait(&p.tfull_bar[buf], use & 1);
      tc_fence_after();
    }
    if (threadIdx.x == 64) TL(4);
    // park the partial tile: scratch[group][part][col][row]  (row fastest -> coalesced); the stores themselves
    // publish it (sentinel exchange, see q_finish)
    float* part = const_cast<float*>(base) + (size_t)kpart * (size_t)n_cols * TILE_M;
#pragma unroll 1
    for (int n0 = 0; n0 < n_cols; n0 += 16) {
      float v[16];
      if (have_acc) {
        tmem_ld_32x16(taddr + n0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        // (a partial sum that happens to carry the sentinel's bits -- only a NaN can -- becomes the canonical NaN)
        const float y = __float_as_uint(v[i]) == SPLIT_SENTINEL ? __uint_as_float(0x7FFFFFFFu) : v[i];
        __stcg(part + (size_t)(n0 + i) * TILE_M + row, y);
      }
    }
    if (have_acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p.tempty_bar[buf]);
    }
    if (threadIdx.x == 64) TL(5);
  }
  q_finish<DIR, true>(E, t, lane, ew, NEW, c_lo, c_hi, ksplit, n_cols, base, p.stg, stg_cols);
  if (threadIdx.x == 64) TL(7);
  if (have_acc) ++p.it;
}

__device__ __forceinline__ void epilogue_chunk(Pipe& p, const EngineParams& P, int tick) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = p.n_cols;
  if (warp >= 6) return;   // plain unsplit jobs: the four TMEM warps do all the work
  const int q = warp & 3;
  const int row = q * 32 + lane;
#pragma unroll 1
  for (int j = cta_rank(P); j < P.njobs; j += P.ncta) {
    const Job& jb = P.jobs[j];
    const int t = job_time(P, jb, tick);
    const int all_kb = job_total_kb(P, jb, t);
    if (all_kb == 0) continue;
    const int buf = p.it & 1;
    const uint32_t use = (uint32_t)(p.it >> 1);
    int cw_shift = 0, cw_limit = 0;
    if (P.chunk_samples > 0) chunk_window(P, jb, t, cw_shift, cw_limit);
    const EpiLocal E = make_epi_local(jb, P.ctx, cw_shift, cw_limit);
    const uint32_t taddr = p.tmem_base + (uint32_t)buf * p.acc_stride + ((uint32_t)(q * 32) << 16);
    mbar_wait(&p.tfull_bar[buf], use & 1);
    tc_fence_after();
    plain_tile(E, t, row, taddr, n_cols);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p.tempty_bar[buf]);
    ++p.it;
  }
}

__device__ __forceinline__ uint8_t* align_smem(uint8_t* raw) {
  // 1024-byte alignment is required by the 128B swizzle atoms
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}

// ---------------------------------------------------------- one launch per phase
__global__ void __launch_bounds__(ENGINE_THREADS, 1) job_kernel_tc(const EngineParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tick = P.tick;
  if (threadIdx.x == 0) TL(0);
  Pipe p;
  pipe_setup(p, align_smem(smem_raw), P.n_cols);
  if (threadIdx.x == 0) TL(1);
  if (warp == 0) {
    if (lane == 0) producer_run(p, P, P.tick);
  } else if (warp == 1) {
    mma_run(p, P, P.tick);
  } else {
    epilogue_run<false>(p, P, P.tick);
  }
  pipe_teardown(p);
  if (threadIdx.x == 0) TL(8);
}

// ---------------------------------------------------------- grid barrier (persistent kernels)
// Monotonic arrival counter, zeroed by the host before the launch; barrier k completes when the counter
// reaches (k + 1) * gridDim.x.  All CTAs are co-resident (cooperative launch).
__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
  asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy stores -> later TMA (async proxy) reads
  __threadfence();
  atomicAdd(ctr, 1u);
}
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target);
__device__ __forceinline__ void grid_wait_ext(const unsigned int* ctr, unsigned int target) { grid_wait(ctr, target); }
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target) {
  unsigned int seen, spins = 0;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
    if (++spins > (1u << 26)) pb_timeout(1);
  } while (seen < target);
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// ------------------------------------------------------------------ SIMT twin
// 128 threads, thread <-> output row.  Operands are rebuilt as hi + lo in fp32,
// so the only difference from the tensor-core path is the dropped lo*lo term
// (~2^-34 relative) and the summation order.
__global__ void __launch_bounds__(128) job_kernel_simt(const EngineParams P) {
  __shared__ float As[TILE_M][KB + 1];
  __shared__ float Bs[32][KB + 1];
  const int tid = threadIdx.x;
  for (int j = cta_rank(P); j < P.njobs; j += P.ncta) {
    const Job& jb = P.jobs[j];
    const int t = job_time(P, jb, P.tick);
    if (job_total_kb(P, jb, t) == 0) continue;
    if (jb.ksplit > 1 && jb.kpart != 0) continue;   // the SIMT twin does not split: part 0 does the whole tile
    int cw_shift = 0, cw_limit = 0;
    if (jb.epi == EPI_PLAIN && P.chunk_samples > 0) chunk_window(P, jb, t, cw_shift, cw_limit);
    const EpiLocal E = make_epi_local(jb, P.ctx, cw_shift, cw_limit);
    for (int n0 = 0; n0 < P.n_cols; n0 += 32) {
      const int nc = (P.n_cols - n0 >= 32) ? 32 : (P.n_cols - n0);
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
      for (int s = 0; s < jb.nseg; ++s) {
        const Seg sg = jb.seg[s];
        if (!seg_valid(P, sg, t)) continue;
        const MapRaw ra = P.raws[sg.a_map], ral = P.raws[sg.a_map + 1];
        const MapRaw rb = P.raws[sg.b_map], rbl = P.raws[sg.b_map + 1];
        const long long bslot = (sg.b_slot == NO_SLOT) ? 0 : (long long)(t + sg.b_slot) * rb.slot_pitch;
        for (int kb = 0; kb < sg.nkb; ++kb) {
          __syncthreads();
          for (int e = tid; e < TILE_M * KB; e += 128) {
            const int r = e / KB, k = e % KB;
            const int gr = sg.a_row + r, gk = sg.a_k + kb * KB + k;
            float x = 0.0f;
            if (gr < ra.rows && gk < ra.cols) {
              long long o = (long long)gr * ra.row_pitch + gk;
              if (ra.tiled_nkb > 0)
                o = (((long long)(gr >> 7) * ra.tiled_nkb + (gk >> 6)) * TILE_M + (gr & 127)) * KB + (gk & 63);
              x = __bfloat162float(ra.base[o]) + __bfloat162float(ral.base[o]);
            }
            As[r][k] = x;
          }
          for (int e = tid; e < 32 * KB; e += 128) {
            const int r = e / KB, k = e % KB;
            const int gr = sg.b_row + (sg.b_slot == NO_SLOT ? cw_shift : 0) + n0 + r, gk = sg.b_k + kb * KB + k;
            float x = 0.0f;
            if (r < nc && gr < rb.rows && gk < rb.cols) {
              const long long o = bslot + (long long)gr * rb.row_pitch + gk;
              x = __bfloat162float(rb.base[o]) + __bfloat162float(rbl.base[o]);
            }
            Bs[r][k] = x;
          }
          __syncthreads();
          for (int k = 0; k < KB; ++k) {
            const float a = As[tid][k];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(a, Bs[i][k], acc[i]);
          }
        }
      }
      run_epilogue<32>(E, t, tid, n0, nc, acc);
    }
  }
}

}  // namespace pb
