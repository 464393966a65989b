// Host side of libparrot_b200.so: configuration, parameter inventory, workspace
// carving, TMA tensor maps, GEMM job tables, and the forward / backward /
// sampling / optimizer orchestration behind the C ABI of include/parrot_b200.h.
//
// Reference call sites replaced (all in /root/reference/model.py unless noted):
//   Parrot.compute_cost 552-824, step 651-724, sample_model_fun 827-1059,
//   initial_states 529-549, Encoder.apply 233-247; train.py:100-108 (optimizer).
#include "../../include/parrot_b200.h"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace pb;

// ------------------------------------------------------------------ utilities
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,      \
               cudaGetErrorString(e_));                                                       \
      throw std::runtime_error(buf_);                                                         \
    }                                                                                         \
  } while (0)
#define REQUIRE(cond, msg)                                                     \
  do {                                                                         \
    if (!(cond)) throw std::runtime_error(std::string("parrot_b200: ") + msg); \
  } while (0)
static bool debug_sync() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PARROT_DEBUG_SYNC");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}
static thread_local const char* g_ctx = "";   // what the host was launching (debug messages)
#define LAUNCH(kern, grid, block, smem, st, ...)                                                     \
  do {                                                                                               \
    kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                                    \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
    CK(cudaGetLastError());                                                                          \
    if (debug_sync()) {                                                                              \
      cudaError_t e2_ = cudaStreamSynchronize(st);                                                   \
      if (e2_ != cudaSuccess) {                                                                      \
        char b2_[512];                                                                               \
        snprintf(b2_, sizeof b2_, "kernel %s failed (%s) while running [%s]", #kern,                 \
                 cudaGetErrorString(e2_), g_ctx);                                                    \
        throw std::runtime_error(b2_);                                                               \
      }                                                                                              \
    }                                                                                                \
  } while (0)

template <typename F>
static int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

static inline int rup(int x, int m) { return (x + m - 1) / m * m; }
static inline long long rupl(long long x, long long m) { return (x + m - 1) / m * m; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int gs_blocks(long long n, int block = 256) {
  long long b = (n + block - 1) / block;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    REQUIRE(p && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// ------------------------------------------------------------------ dimensions
struct Dims {
  int H, R, D, A, C, U, B, T, E, S, K, Dtot, IN, NC;
  int Np, Hp, Cp, Rp, Dp, Dtp, Ap;
  bool enc, spk, weak, full, gmm, sampling, ln;
};
static Dims make_dims(const parrot_config& c) {
  Dims d;
  d.H = c.rnn_h_dim; d.R = c.readouts_dim; d.D = c.output_dim; d.A = c.attention_size;
  d.enc = c.encoder_type == 1;
  d.E = c.encoder_dim; d.IN = c.input_dim; d.NC = c.num_characters;
  d.C = d.enc ? 2 * c.encoder_dim : c.input_dim;
  d.U = c.text_len; d.B = c.batch_size; d.T = c.seq_len; d.S = c.speaker_dim; d.K = c.k_gmm;
  d.gmm = c.which_cost == 1;
  d.Dtot = d.gmm ? 2 * d.D * d.K + d.K : d.D;
  d.spk = c.use_speaker != 0;
  d.full = c.full_feedback != 0;
  d.weak = c.weak_feedback != 0 || d.full;
  d.sampling = c.sampling != 0;
  d.ln = c.layer_norm != 0;
  d.Np = rup(d.B, 16); d.Hp = rup(d.H, 64); d.Cp = rup(d.C, 64); d.Rp = rup(d.R, 64);
  d.Dp = rup(d.D, 64); d.Ap = 64;
  // planes of d(pred): the GMM blocks [mu | sigma | coeff] start at 64-aligned columns (TMA inner
  // coordinates must stay 16-byte aligned)
  d.Dtp = d.gmm ? 2 * rup(d.D * d.K, 64) + rup(d.K, 64) : rup(d.D, 64);
  return d;
}
static void check_cfg(const parrot_config& c) {
  REQUIRE(c.batch_size >= 1 && c.batch_size <= 256, "batch_size per device must be in [1, 256]");
  REQUIRE(c.seq_len >= 1 && c.text_len >= 1, "seq_len / text_len must be positive");
  REQUIRE(3 * c.attention_size <= 64 && c.attention_size <= 32, "attention_size must be <= 21");
  REQUIRE(c.which_cost == 0 || c.k_gmm <= 32, "k_gmm must be <= 32");
  REQUIRE(c.encoder_type == 0 || c.encoder_type == 1, "encoder_type must be 0 (None) or 1 (bidirectional)");
  REQUIRE(c.encoder_type == 1 || c.num_characters == c.input_dim, "model.py:224 num_characters == input_dim");
  REQUIRE(c.rnn_h_dim >= 64 && c.rnn_h_dim % 64 == 0, "rnn_h_dim must be a multiple of 64");
  REQUIRE(c.readouts_dim >= 16, "readouts_dim too small");
}

// ------------------------------------------------------------------ parameters
struct PInfo {
  std::string name;
  long long off;
  int rows, cols;  // cols == 0: vector of `rows`
  long long numel() const { return cols ? (long long)rows * cols : rows; }
};
struct PReg {
  std::vector<PInfo> v;
  std::map<std::string, int> idx;
  long long total = 0;
  void add(const std::string& n, int r, int c) {
    PInfo p{n, total, r, c};
    idx[n] = (int)v.size();
    v.push_back(p);
    total += rupl(p.numel(), 64);
  }
  void lin(const std::string& n, int i, int o) { add(n + ".W", i, o); add(n + ".b", o, 0); }
  void fork(const std::string& n, int i, const std::vector<std::string>& outs, const std::vector<int>& dims) {
    for (size_t k = 0; k < outs.size(); ++k) lin(n + "/fork_" + outs[k], i, dims[k]);
  }
  void gru(const std::string& n, int d) {
    add(n + ".state_to_state", d, d);
    add(n + ".state_to_gates", d, 2 * d);
    add(n + ".initial_state", d, 0);
  }
  // Linear bricks may be named without the ".W" suffix
  const PInfo& get(const std::string& n) const {
    auto it = idx.find(n);
    if (it == idx.end()) it = idx.find(n + ".W");
    if (it == idx.end()) throw std::runtime_error("parrot_b200: unknown parameter " + n);
    return v[it->second];
  }
  bool has(const std::string& n) const { return idx.count(n) != 0; }
};
// Same inventory and order as oracle.param_shapes (model.py:251-506); tests compare the two.
static PReg make_params(const parrot_config& c) {
  const Dims d = make_dims(c);
  PReg P;
  const std::string root = "/parrot";
  auto ln = [](int i) { return std::to_string(i); };
  if (d.enc) {
    P.add(root + "/encoder/embed_label.W", d.NC, d.IN);
    for (const char* dir : {"forward", "backward"}) {
      const std::string base = root + "/encoder/encoder/" + dir;
      P.gru(base + "/gatedrecurrent", d.E);
      P.fork(base + "/fork", d.IN, {"inputs", "gate_inputs"}, {d.E, 2 * d.E});
    }
  }
  for (int i = 1; i <= 3; ++i) P.gru(root + "/rnn" + ln(i), d.H);
  for (int i = 1; i <= 3; ++i) P.lin(root + "/h" + ln(i) + "_to_readout", d.H, d.R);
  P.fork(root + "/h1_to_h2", d.H, {"rnn2_inputs", "rnn2_gates"}, {d.H, 2 * d.H});
  P.fork(root + "/h1_to_h3", d.H, {"rnn3_inputs", "rnn3_gates"}, {d.H, 2 * d.H});
  P.fork(root + "/h2_to_h3", d.H, {"rnn3_inputs", "rnn3_gates"}, {d.H, 2 * d.H});
  if (!d.gmm) P.lin(root + "/readout_to_output", d.R, d.D);
  else P.fork(root + "/readout_to_output", d.R, {"gmm_mu", "gmm_sigma", "gmm_coeff"}, {d.D * d.K, d.D * d.K, d.K});
  for (int i = 1; i <= 3; ++i)
    P.fork(root + "/inp_to_h" + ln(i), d.C, {"rnn" + ln(i) + "_inputs", "rnn" + ln(i) + "_gates"}, {d.H, 2 * d.H});
  P.fork(root + "/h1_to_att", d.H, {"alpha", "beta", "kappa"}, {d.A, d.A, d.A});
  P.lin(root + "/att_to_readout", d.C, d.R);
  if (d.spk) {
    P.add(root + "/lookuptable.W", c.num_speakers, d.S);
    for (int i = 1; i <= 3; ++i)
      P.fork(root + "/speaker_to_h" + ln(i), d.S, {"rnn" + ln(i) + "_inputs", "rnn" + ln(i) + "_gates"},
             {d.H, 2 * d.H});
    P.lin(root + "/speaker_to_readout", d.S, d.R);
    if (!d.gmm) P.lin(root + "/speaker_to_output", d.S, d.D);
    else P.fork(root + "/speaker_to_output", d.S, {"gmm_mu", "gmm_sigma", "gmm_coeff"}, {d.D * d.K, d.D * d.K, d.K});
  }
  if (d.full) {
    P.fork(root + "/out_to_h2", d.D, {"rnn2_inputs", "rnn2_gates"}, {d.H, 2 * d.H});
    P.fork(root + "/out_to_h3", d.D, {"rnn3_inputs", "rnn3_gates"}, {d.H, 2 * d.H});
  }
  if (d.weak) P.fork(root + "/out_to_h1", d.D, {"rnn1_inputs", "rnn1_gates"}, {d.H, 2 * d.H});
  P.add(root + ".initial_w", d.C, 0);
  return P;
}

// ------------------------------------------------------------------ model
struct Plane {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  int rows = 0;       // rows per slot
  int pitch = 0;      // elements per row (multiple of 64)
  int slots = 1;
  long long rows_alloc = 0;  // total rows allocated (slots*rows rounded up to 128)
  int tiled_nkb = 0;         // > 0: tile-contiguous weight pack with this many k blocks per 128-row tile
};
struct WPack {
  std::string name;
  int in = 0, out = 0;
  Plane fwd, bwd;     // fwd: [out_p128][in_p64] (W^T), bwd: [in_p128][out_p64] (W)
  int fwd_map = -1, bwd_map = -1;
  bool need_bwd = true;
};
struct Table {
  int off = 0, count = 0, n_cols = 0;
  int chunk_samples = 0;   // > 0: chunk table (engine.cuh chunk_window)
  // split-K scratch of the table: tables that only ever run one after the other share region 0; tables of the
  // grouped persistent scans (several tables in flight at once) own a region behind it (uniq >= 0: float / counter
  // offsets relative to the end of the shared region)
  long long uniq_floats = -1;
  int uniq_groups = -1;
};
struct Buf {
  size_t off;
  size_t bytes;
};
struct WGrad {
  int xt_map, a_k, dyt_map, b_row, in, out;
  long long goff;
};

struct parrot_model {
  parrot_config cfg;
  Dims d;
  PReg P;
  float* params = nullptr;
  float* grads = nullptr;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0, ws_used = 0;
  bool dry = true;
  std::map<std::string, Buf> bufs;
  std::vector<CUtensorMap> maps;
  std::vector<MapRaw> raws;
  CUtensorMap* d_maps = nullptr;
  MapRaw* d_raws = nullptr;
  std::vector<Job> jobs;
  Job* d_jobs = nullptr;
  ScanCtx ctx;
  ScanCtx* d_ctx = nullptr;
  std::map<std::string, WPack> packs;
  std::map<std::string, Plane> planes;
  std::map<std::string, int> map_scan, map_plain, map_tA, map_tB;  // plane name -> hi map index
  std::map<std::string, Table> tables;
  std::vector<WGrad> wgrads;
  int max_groups = 0;
  size_t max_split_floats = 0;
  long long uniq_split_floats = 0;   // total of the per-table scratch regions (grouped scans)
  int uniq_split_groups = 0;
  int Tc = 0;              // chunk length of the chunk-lagged layer wavefront (0: not used)
  int grp_f[3] = {0, 0, 0};   // CTAs per layer group of the grouped persistent scans (forward / backward sweep)
  int grp_b[3] = {0, 0, 0};
  std::map<int, cudaGraphExec_t> samp_graphs;   // sampling loop as a CUDA graph, keyed by (speaker, injected noise)
  int att_slices = 0;      // K slices of the attention projection (persistent scan)
  unsigned long long* timeline = nullptr;
  unsigned long long* stamps = nullptr;   // debug: persistent forward scan per-barrier stamps
  int stamp_bars = 0;
  int tl_tick = -1;
  unsigned long long* stamps_bwd = nullptr;
  int sm_count = 0;
  bool persistent_ok = true;   // cleared when a cooperative launch is refused
  float* d_split_scratch = nullptr;
  unsigned int* d_split_count = nullptr;
  unsigned int* d_gridbar = nullptr;
  bool dirty = true;
  float last_start_flag = 1.0f;
  bool have_fwd = false;
  // optional per-launch timing (bench.py roofline): CUDA events around every tagged launch
  int profiling = 0;   // 0 off, 1 section events only (persistent kernels stay on), 2 per-launch events (per-phase launches)
  struct ProfRec { std::string key; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
  cudaEvent_t prof_begin(const std::string& key, cudaStream_t st) {
    if (!profiling) return nullptr;
    if (profiling < 2 && key.compare(0, 4, "sec_") != 0) return nullptr;
    ProfRec r;
    r.key = key;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    prof.push_back(r);
    return r.e1;
  }
  static void prof_end(cudaEvent_t e1, cudaStream_t st) { if (e1) cudaEventRecord(e1, st); }

  // ---- workspace ----
  void* alloc(const std::string& name, size_t bytes) {
    ws_used = (ws_used + 1023) & ~size_t(1023);
    const size_t off = ws_used;
    ws_used += bytes;
    if (!name.empty()) bufs[name] = Buf{off, bytes};
    if (dry) return nullptr;
    REQUIRE(ws_used <= ws_bytes, "workspace too small");
    return ws + off;
  }
  float* falloc(const std::string& name, long long n) { return (float*)alloc(name, (size_t)n * 4); }
  float* fbuf(const std::string& name) const {
    auto it = bufs.find(name);
    if (it == bufs.end()) throw std::runtime_error("parrot_b200: unknown buffer " + name);
    return (float*)(ws + it->second.off);
  }
  float* pp(const std::string& n) const { return params + P.get("/parrot" + n).off; }
  float* gp(const std::string& n) const { return grads + P.get("/parrot" + n).off; }

  Plane make_plane(const std::string& name, int rows, int cols, int slots) {
    Plane p;
    p.rows = rows;
    p.pitch = rup(cols, 64);
    p.slots = slots;
    p.rows_alloc = rupl((long long)rows * slots, 128);
    const size_t bytes = (size_t)p.rows_alloc * p.pitch * 2;
    p.hi = (bf16*)alloc(name + ".hi", bytes);
    p.lo = (bf16*)alloc(name + ".lo", bytes);
    planes[name] = p;
    return p;
  }
  // returns the index of the hi map; lo map follows
  int make_map(const Plane& p, int rank, int box_rows) {
    const int idx = (int)maps.size();
    for (int w = 0; w < 2; ++w) {
      CUtensorMap m;
      memset(&m, 0, sizeof m);
      MapRaw r;
      r.base = w ? p.lo : p.hi;
      r.row_pitch = p.pitch;
      r.box_rows = box_rows;
      r.cols = p.pitch;
      r.tiled_nkb = p.tiled_nkb;
      r.pad_ = 0;
      if (rank == 2) {
        r.slot_pitch = 0; r.rows = (int)p.rows_alloc; r.slots = 1;
      } else {
        r.slot_pitch = (long long)p.rows * p.pitch; r.rows = p.rows; r.slots = p.slots;
      }
      if (!dry && w == 0) {
        // dims: k, rows[, slots], plane (hi, lo): one instruction loads both planes of a tile (lo tile behind the hi tile)
        cuuint64_t dims[4] = {(cuuint64_t)p.pitch, (cuuint64_t)(rank == 2 ? p.rows_alloc : p.rows),
                              (cuuint64_t)p.slots, 2};
        const cuuint64_t plane_stride = (cuuint64_t)((const uint8_t*)p.lo - (const uint8_t*)p.hi);
        cuuint64_t strides[3] = {(cuuint64_t)p.pitch * 2, (cuuint64_t)p.rows * p.pitch * 2, plane_stride};
        if (p.tiled_nkb > 0) {
          // tile-contiguous pack seen as a [tiles*128][64] matrix: every 128 x 64 box is one contiguous 16 KB block
          dims[0] = 64;
          dims[1] = (cuuint64_t)(p.rows_alloc / 128) * p.tiled_nkb * 128;
          strides[0] = 128;
        }
        cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 2};
        cuuint32_t es[4] = {1, 1, 1, 1};
        int trank = rank + 1;
        if (rank == 2) { dims[2] = 2; strides[1] = plane_stride; box[2] = 2; }
        REQUIRE(plane_stride % 16 == 0 && plane_stride > 0, "operand planes must be 16-byte aligned and hi < lo");
        CUresult res = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)trank, (void*)r.base, dims,
                                    strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (res != CUDA_SUCCESS) {
          char b[256];
          snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) rank %d rows %d pitch %d box %d", (int)res,
                   trank, p.rows, p.pitch, box_rows);
          throw std::runtime_error(b);
        }
      }
      maps.push_back(m);
      raws.push_back(r);
    }
    return idx;
  }
  // Forward planes are allocated here, one after the other; finish_packs() then allocates all backward planes.
  // Each set is one contiguous region of the workspace, so that one L2 access-policy window covers it.
  std::vector<std::string> pack_order;
  size_t fwd_region_off = 0, fwd_region_bytes = 0, bwd_region_off = 0, bwd_region_bytes = 0;
  WPack& add_pack(const std::string& pname, bool need_bwd) {
    const PInfo& pi = P.get("/parrot" + pname);
    WPack w;
    w.name = pname; w.in = pi.rows; w.out = pi.cols; w.need_bwd = need_bwd;
    if (pack_order.empty()) fwd_region_off = ws_used;
    w.fwd = make_plane("pack.f" + pname, rup(w.out, 128), w.in, 1);
    w.fwd.tiled_nkb = w.fwd.pitch / 64;
    planes["pack.f" + pname] = w.fwd;
    w.fwd_map = make_map(w.fwd, 2, 128);
    fwd_region_bytes = ws_used - fwd_region_off;
    packs[pname] = w;
    pack_order.push_back(pname);
    return packs[pname];
  }
  void finish_packs() {
    bwd_region_off = ws_used;
    for (auto& pname : pack_order) {
      WPack& w = packs[pname];
      if (!w.need_bwd) continue;
      w.bwd = make_plane("pack.b" + pname, rup(w.in, 128), w.out, 1);
      w.bwd.tiled_nkb = w.bwd.pitch / 64;
      planes["pack.b" + pname] = w.bwd;
      w.bwd_map = make_map(w.bwd, 2, 128);
    }
    bwd_region_bytes = ws_used - bwd_region_off;
  }
};

static const int NT = 128;  // sample tile of the batched (outside-the-scan) products
// output-feature tile of the weight-gradient products: N = 256 is the only width at which a tcgen05.mma of this engine
// runs at the tensor-pipe floor behind a TMA ring (tools/tma_bw.cu: 136 cycles per MMA vs 110 at N = 128 for half the work)
#ifndef PB_WGRAD_NT
#define PB_WGRAD_NT 256
#endif
static const int WNT = PB_WGRAD_NT;
static const int COLSUM_CHUNKS = 64;
#ifndef PB_SPLIT_TARGET
#define PB_SPLIT_TARGET 148   // CTAs one scan phase is split over (experiment knob)
#endif
static const int SGEMM_SPLIT = 64;   // K chunks of the split SIMT GEMM (encoder weight gradients)

// ------------------------------------------------------------------ job builders
static Seg mkseg(int a_map, int a_row, int a_k, int b_map, int b_row, int b_k, int b_slot, int nkb) {
  Seg s;
  s.a_map = a_map; s.a_row = a_row; s.a_k = a_k; s.b_map = b_map; s.b_row = b_row; s.b_k = b_k;
  s.b_slot = b_slot; s.nkb = nkb; s.a_nkb = 0; s.b_slots = 1;
  return s;
}
static Job blank_job() {
  Job j;
  memset(&j, 0, sizeof j);
  j.pa.scale = 1.0f;
  return j;
}

static std::string LN(int l) { return std::to_string(l + 1); }

// forward scan jobs of one layer (gates or candidate): model.py:655-662, 692-722
// hoisted: only the products that depend on the layer's own recurrence stay in the scan (state_to_gates /
// state_to_state, and for layer 1 the attention context w_{t-1}); everything else reaches the epilogue through
// LayerBuf::pre (chunk-lagged wavefront, kernels.cuh).
static void build_fwd_layer_jobs(parrot_model& M, std::vector<Job>& out, int layer, bool gates, int lag,
                                 bool ln = false, bool hoisted = false) {
  const Dims& d = M.d;
  const int rows = gates ? 2 * d.H : d.H;
  const std::string l = LN(layer);
  const std::string fk = gates ? "_gates" : "_inputs";
  for (int mt = 0; mt < cdiv(rows, 128); ++mt) {
    Job j = blank_job();
    j.epi = gates ? EPI_GATES : EPI_CAND;
    j.lag = lag; j.layer = layer; j.row0 = mt * 128;
    j.m_valid = std::min(128, rows - mt * 128);
    int ns = 0;
    const int hmap = M.map_scan["h" + l];
    if (gates)
      j.seg[ns++] = mkseg(M.packs["/rnn" + l + ".state_to_gates"].fwd_map, mt * 128, 0, hmap, 0, 0, 0, d.Hp / 64);
    else
      j.seg[ns++] = mkseg(M.packs["/rnn" + l + ".state_to_state"].fwd_map, mt * 128, 0, M.map_scan["rh" + l], 0, 0,
                          0, d.Hp / 64);
    // attention context: layer 1 consumes w_{t-1} (slot t), layers 2,3 consume w_t (slot t+1)
    if (!hoisted || layer == 0)
      j.seg[ns++] = mkseg(M.packs["/inp_to_h" + l + "/fork_rnn" + l + fk].fwd_map, mt * 128, 0, M.map_scan["w"], 0, 0,
                          layer == 0 ? 0 : 1, d.Cp / 64);
    if (layer >= 1 && !ln && !hoisted)
      j.seg[ns++] = mkseg(M.packs["/h1_to_h" + l + "/fork_rnn" + l + fk].fwd_map, mt * 128, 0, M.map_scan["h1"], 0, 0,
                          1, d.Hp / 64);
    if (layer == 2 && !ln && !hoisted)
      j.seg[ns++] = mkseg(M.packs["/h2_to_h3/fork_rnn3" + fk].fwd_map, mt * 128, 0, M.map_scan["h2"], 0, 0, 1,
                          d.Hp / 64);
    if (!ln && !hoisted && ((layer == 0 && d.weak) || (layer > 0 && d.full)))
      j.seg[ns++] = mkseg(M.packs["/out_to_h" + l + "/fork_rnn" + l + fk].fwd_map, mt * 128, 0, M.map_scan["xin"], 0,
                          0, 0, d.Dp / 64);
    j.nseg = ns;
    out.push_back(j);
  }
}

// batched product over all frames: out[sample][f0 + row] = sum_seg ...
struct PlainSeg {
  int a_map, a_k, b_map, b_row0, b_k, nkb;
};
static void build_plain_jobs(std::vector<Job>& out, int Mrows, int f0, long long n_samples,
                             const std::vector<PlainSeg>& segs, const PlainArgs& pa) {
  // sample-tile major: the CTAs running at the same time share one activation tile (L2 hits) and walk the
  // (small, L2-resident) weight tiles; feature-tile major would re-stream the activation planes per feature tile
  for (int nt = 0; nt < cdiv(n_samples, NT); ++nt)
    for (int mt = 0; mt < cdiv(Mrows, 128); ++mt) {
      Job j = blank_job();
      j.epi = EPI_PLAIN;
      j.row0 = f0 + mt * 128;
      j.m_valid = std::min(128, Mrows - mt * 128);
      j.n0 = nt * NT;
      j.nseg = (int)segs.size();
      for (size_t s = 0; s < segs.size(); ++s)
        j.seg[s] = mkseg(segs[s].a_map, mt * 128, segs[s].a_k, segs[s].b_map, segs[s].b_row0 + nt * NT, segs[s].b_k,
                         NO_SLOT, segs[s].nkb);
      j.pa = pa;
      out.push_back(j);
    }
}

// tables assembled from several build_plain_jobs calls: keep all jobs of one sample tile adjacent
static void sort_by_sample_tile(std::vector<Job>& js) {
  std::stable_sort(js.begin(), js.end(), [](const Job& a, const Job& b) { return a.n0 < b.n0; });
}

static int job_kb(const Job& j) {
  int n = 0;
  for (int s = 0; s < j.nseg; ++s) n += j.seg[s].nkb;
  return n;
}
// Split the K dimension of the jobs of a scan table so that one launch spreads over `target_ctas` CTAs
// (each tile of a recurrent phase would otherwise stream its full K alone: per-SM bandwidth bound).
static std::vector<Job> split_jobs(const std::vector<Job>& js, int target_ctas, int max_split, int* groups) {
  long long total = 0;
  for (auto& j : js) total += job_kb(j);
  int per = (int)std::max<long long>(1, (total + target_ctas - 1) / target_ctas);
  std::vector<Job> out;
  for (;;) {
    long long parts = 0;
    for (auto& j : js) parts += std::min(max_split, std::max(1, (job_kb(j) + per - 1) / per));
    if (parts <= target_ctas || per > total) break;
    ++per;
  }
  // CTAs left over go to the least-split tiles: their post-reduction epilogue covers the widest column slice
  // (n_cols / parts) and therefore sets the length of the phase's tail
  std::vector<int> parts_of(js.size());
  long long used = 0;
  for (size_t i = 0; i < js.size(); ++i) {
    parts_of[i] = std::min(max_split, std::max(1, (job_kb(js[i]) + per - 1) / per));
    used += parts_of[i];
  }
  for (;;) {
    int lo = max_split + 1;
    for (size_t i = 0; i < js.size(); ++i)
      if (parts_of[i] < job_kb(js[i])) lo = std::min(lo, parts_of[i]);
    if (lo >= max_split) break;
    int n_lo = 0;
    for (size_t i = 0; i < js.size(); ++i)
      if (parts_of[i] == lo && parts_of[i] < job_kb(js[i])) ++n_lo;
    if (n_lo == 0 || used + n_lo > target_ctas) break;   // only whole levels: a partial upgrade leaves the tail as is
    for (size_t i = 0; i < js.size(); ++i)
      if (parts_of[i] == lo && parts_of[i] < job_kb(js[i])) { ++parts_of[i]; ++used; }
  }
  int g = 0;
  for (size_t ji = 0; ji < js.size(); ++ji) {
    const Job& j = js[ji];
    const int sp = parts_of[ji];
    for (int p = 0; p < sp; ++p) {
      Job c = j;
      c.ksplit = sp; c.kpart = p; c.group = g;
      out.push_back(c);
    }
    ++g;
  }
  *groups = g;
  return out;
}

static void push_table(parrot_model& M, const std::string& name, const std::vector<Job>& js_in, int n_cols,
                       int split_target = 0, int chunk_samples = 0, bool own_scratch = false) {
  std::vector<Job> js = js_in;
  for (auto& j : js)
    for (int s = 0; s < j.nseg; ++s) {
      j.seg[s].a_nkb = M.raws[j.seg[s].a_map].tiled_nkb;   // tile-contiguous packs
      j.seg[s].b_slots = M.raws[j.seg[s].b_map].slots;
    }
  Table t;
  t.off = (int)M.jobs.size();
  t.n_cols = n_cols;
  t.chunk_samples = chunk_samples;
  if (split_target > 0) {
    int groups = 0;
    std::vector<Job> sp = split_jobs(js, split_target, MAX_KSPLIT, &groups);
    // parts of one tile must run concurrently with distinct CTAs: order parts-major so that CTA i gets job i
    t.count = (int)sp.size();
    if (getenv("PARROT_DEBUG_SPLIT")) {
      std::map<std::pair<int, int>, int> hist;   // (k blocks of the tile, parts) -> tiles
      for (auto& j : sp)
        if (j.kpart == 0) ++hist[{job_kb(j), j.ksplit}];
      fprintf(stderr, "[split] %-10s %3d CTAs:", name.c_str(), t.count);
      for (auto& kv : hist) fprintf(stderr, "  %dx(kb=%d parts=%d)", kv.second, kv.first.first, kv.first.second);
      fprintf(stderr, "\n");
    }
    M.jobs.insert(M.jobs.end(), sp.begin(), sp.end());
    if (own_scratch) {
      t.uniq_floats = M.uniq_split_floats; t.uniq_groups = M.uniq_split_groups;
      M.uniq_split_floats += (long long)groups * MAX_KSPLIT * n_cols * TILE_M;
      M.uniq_split_groups += groups;
    } else {
      M.max_groups = std::max(M.max_groups, groups);
      M.max_split_floats = std::max(M.max_split_floats, (size_t)groups * MAX_KSPLIT * n_cols * TILE_M);
    }
  } else {
    t.count = (int)js.size();
    M.jobs.insert(M.jobs.end(), js.begin(), js.end());
  }
  M.tables[name] = t;
}

static void run_table(parrot_model& M, const std::string& name, int tick, int T, int reverse, cudaStream_t st) {
  const Table& t = M.tables.at(name);
  if (t.count == 0) return;
  static thread_local char ctxbuf[128];
  snprintf(ctxbuf, sizeof ctxbuf, "table %s tick %d njobs %d n_cols %d", name.c_str(), tick, t.count, t.n_cols);
  g_ctx = ctxbuf;
  EngineParams P;
  P.jobs = M.d_jobs + t.off; P.njobs = t.count; P.maps = M.d_maps; P.raws = M.d_raws; P.ctx = M.d_ctx;
  P.tick = tick; P.T = T; P.n_cols = t.n_cols; P.reverse = reverse;
  P.chunk_samples = t.chunk_samples;
  P.split_scratch = M.d_split_scratch; P.split_count = M.d_split_count;
  P.timeline = M.tl_tick >= 0 ? nullptr : M.timeline; P.tl_tick = -1;   // (tl_tick >= 0: persistent-tick debugging)
  P.coop_epilogue = (t.count <= 148 && M.sm_count >= 148) ? 1 : 0;
  { const char* e = getenv("PARROT_DEBUG_FLAGS"); P.debug_flags = e ? atoi(e) : 0; }
  cudaEvent_t pe = M.prof_begin(name, st);
  P.cta0 = 0;
  if (M.cfg.gemm_impl == 1) {
    const int grid = std::min(t.count, 148 * 8);
    P.ncta = grid;
    LAUNCH(job_kernel_simt, grid, 128, 0, st, P);
  } else {
    const int grid = std::min(t.count, 148);
    P.ncta = grid;
    LAUNCH(job_kernel_tc, grid, ENGINE_THREADS, SMEM_BYTES + 1024, st, P);
  }
  parrot_model::prof_end(pe, st);
}

// Grouped persistent scans: CTA r of a layer group runs job r of the group's two scan tables every step.  The weight
// tiles of up to GROUP_RES_KB of its k blocks stay resident in the CTA's tensor memory for the whole sweep (engine.cuh
// resident_preload / mma_scan): first table first, the second table gets what is left.
static void assign_resident(parrot_model& M, const std::string& ta, const std::string& tb) {
  if (!M.tables.count(ta) || !M.tables.count(tb)) return;
  if (const char* e = getenv("PARROT_NO_RESIDENT")) { if (e[0] && e[0] != '0') return; }
  const Table& A = M.tables.at(ta);
  const Table& B = M.tables.at(tb);
  auto part_kb = [](const Job& j) {
    const int total = job_kb(j);
    if (j.ksplit <= 1) return total;
    return (total * (j.kpart + 1)) / j.ksplit - (total * j.kpart) / j.ksplit;
  };
  auto tiled = [&](const Job& j) {
    for (int s = 0; s < j.nseg; ++s)
      if (j.seg[s].a_nkb <= 0) return false;
    return true;
  };
  for (int r = 0; r < std::max(A.count, B.count); ++r) {
    int left = GROUP_RES_KB, col = GROUP_RES_COL0;
    for (const Table* t : {&A, &B}) {
      if (r >= t->count) continue;
      Job& j = M.jobs[t->off + r];
      const int take = tiled(j) ? std::min(part_kb(j), left) : 0;
      j.res_kb = take; j.res_col = col;
      left -= take; col += 64 * take;
    }
  }
}

// ------------------------------------------------------------------ build
static void build(parrot_model& M) {
  const Dims& d = M.d;
  M.ws_used = 0;
  M.bufs.clear(); M.maps.clear(); M.raws.clear(); M.jobs.clear(); M.packs.clear(); M.planes.clear();
  M.tables.clear(); M.wgrads.clear(); M.max_groups = 0; M.max_split_floats = 0; M.pack_order.clear();
  M.uniq_split_floats = 0; M.uniq_split_groups = 0;
  const int T = d.T, B = d.B, H = d.H, Np = d.Np;
  const bool train = !d.sampling;

  // ---- weight packs
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l);
    M.add_pack("/rnn" + s + ".state_to_gates", train);
    M.add_pack("/rnn" + s + ".state_to_state", train);
    M.add_pack("/inp_to_h" + s + "/fork_rnn" + s + "_inputs", train);
    M.add_pack("/inp_to_h" + s + "/fork_rnn" + s + "_gates", train);
    M.add_pack("/h" + s + "_to_readout", train);
    if ((l == 0 && d.weak) || (l > 0 && d.full)) {
      M.add_pack("/out_to_h" + s + "/fork_rnn" + s + "_inputs", false);
      M.add_pack("/out_to_h" + s + "/fork_rnn" + s + "_gates", false);
    }
  }
  for (const char* nm : {"/h1_to_h2/fork_rnn2", "/h1_to_h3/fork_rnn3", "/h2_to_h3/fork_rnn3"}) {
    M.add_pack(std::string(nm) + "_inputs", train);
    M.add_pack(std::string(nm) + "_gates", train);
  }
  M.add_pack("/att_to_readout", train);
  std::vector<std::string> out_forks;
  std::vector<int> out_off, out_dim, out_poff;
  if (!d.gmm) {
    out_forks = {"/readout_to_output"}; out_off = {0}; out_dim = {d.D}; out_poff = {0};
  } else {
    out_poff = {0, rup(d.D * d.K, 64), 2 * rup(d.D * d.K, 64)};
    out_forks = {"/readout_to_output/fork_gmm_mu", "/readout_to_output/fork_gmm_sigma",
                 "/readout_to_output/fork_gmm_coeff"};
    out_off = {0, d.D * d.K, 2 * d.D * d.K};
    out_dim = {d.D * d.K, d.D * d.K, d.K};
  }
  for (auto& f : out_forks) M.add_pack(f, train);
  M.finish_packs();

  // ---- fp32 staging of small derived quantities
  M.falloc("att_wT", (long long)3 * d.A * H);
  M.falloc("att_b", 3 * d.A);
  for (int l = 0; l < 3; ++l) M.falloc("bias_l" + LN(l), 3 * H);
  M.falloc("bias_ro", d.R);
  M.falloc("bias_out", d.Dtot);
  for (int l = 0; l < 3; ++l) M.falloc("base" + LN(l), (long long)B * 3 * H);
  if (d.spk) {
    M.falloc("spk_emb", (long long)B * d.S);
    M.falloc("spk_ro", (long long)B * d.R);
    M.falloc("spk_out", (long long)B * d.Dtot);
    M.falloc("spk_dproj", (long long)B * 3 * H);
    M.falloc("spk_demb", (long long)B * d.S);
    M.falloc("spk_dro", (long long)B * d.R);
    M.falloc("spk_dout", (long long)B * d.Dtot);
  }
  // ---- carried state (model.py:534-546)
  for (int l = 0; l < 3; ++l) M.falloc("last_h" + LN(l), (long long)B * H);
  M.falloc("last_k", (long long)B * d.A);
  M.falloc("last_w", (long long)B * d.C);
  M.falloc("cost", 4);
  M.falloc("opt_stats", 4);

  // ---- encoder
  M.falloc("ctx", (long long)B * d.U * d.C);
  if (d.enc) {
    const int E = d.E;
    const long long LNn = (long long)B * d.U;
    M.falloc("enc_proj", (long long)d.NC * 2 * 3 * E);  // per character: [dir][xi(E) | xg(2E)]
    M.falloc("enc_dproj", (long long)d.NC * 2 * 3 * E);
    M.alloc("enc_lab", (size_t)LNn * 4);
    for (int dir = 0; dir < 2; ++dir) {
      const std::string s = std::to_string(dir);
      M.falloc("enc_xi" + s, LNn * E); M.falloc("enc_xg" + s, LNn * 2 * E);
      M.falloc("enc_z" + s, LNn * E); M.falloc("enc_r" + s, LNn * E); M.falloc("enc_c" + s, LNn * E);
      M.falloc("enc_sp" + s, LNn * E);
      if (train) {
        M.falloc("enc_dxi" + s, LNn * E); M.falloc("enc_dxg" + s, LNn * 2 * E);
        M.falloc("enc_ds0" + s, (long long)std::max(B, d.U) * E);
        M.falloc("enc_rs" + s, LNn * E);
      }
    }
    M.falloc("enc_out", LNn * 2 * E);
    if (train) M.falloc("enc_dout", LNn * 2 * E);
    M.falloc("enc_xcat", LNn * 3 * E);
  }

  // ---- scan state / stashes
  ScanCtx& X = M.ctx;
  memset(&X, 0, sizeof X);
  X.T = T; X.B = B; X.Np = Np; X.H = H; X.Hp = d.Hp; X.C = d.C; X.Cp = d.Cp; X.A = d.A; X.U = d.U;
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l);
    LayerBuf& L = X.L[l];
    L.h = M.falloc("h" + s, (long long)(T + 1) * B * H);
    Plane ph = M.make_plane("h" + s, Np, H, T + 1);
    L.h_hi = ph.hi; L.h_lo = ph.lo;
    Plane pr = M.make_plane("rh" + s, Np, H, T);
    L.rh_hi = pr.hi; L.rh_lo = pr.lo;
    L.z = M.falloc("z" + s, (long long)T * B * H);
    L.r = M.falloc("r" + s, (long long)T * B * H);
    L.c = M.falloc("c" + s, (long long)T * B * H);
    L.base = M.dry ? nullptr : M.fbuf("base" + s);
    L.fb = nullptr;
    M.map_scan["h" + s] = M.make_map(ph, 3, Np);
    M.map_scan["rh" + s] = M.make_map(pr, 3, Np);
    M.map_plain["h" + s] = M.make_map(ph, 2, NT);
    if (train) {
      L.dh = M.falloc("dh" + s, (long long)(T + 1) * B * H);
      L.da = M.falloc("da" + s, (long long)T * B * 3 * H);
      Plane pd = M.make_plane("da" + s, Np, 3 * d.Hp, T);
      L.da_hi = pd.hi; L.da_lo = pd.lo;
      M.map_scan["da" + s] = M.make_map(pd, 3, Np);
      if (!d.ln) {
        M.map_plain["da" + s] = M.make_map(pd, 2, NT);
        // hoisted pre-activation terms of the chunk-lagged wavefront (layer 1: only the teacher-forced feedback)
        if (l > 0 || d.weak) L.pre = M.falloc("pre" + s, (long long)T * B * 3 * H);
      }
    }
  }
  M.Tc = (train && !d.ln) ? (T >= 64 ? 16 : 8) : 0;
  if (const char* e = getenv("PARROT_TC")) { if (M.Tc > 0 && atoi(e) >= 4) M.Tc = atoi(e); }   // (>= 4: see scan_bwd_grouped)
  {
    // CTAs per layer group of the grouped persistent scans.  Layer 1 (+ attention) is the longest dependent chain and
    // needs >= B CTAs for the window stage; layers 2 / 3 also run the hoisted chunk products.
    auto parse = [](const char* name, int* out, int a, int b, int c) {
      out[0] = a; out[1] = b; out[2] = c;
      const char* e = getenv(name);
      int x, y, z;
      if (e && sscanf(e, "%d,%d,%d", &x, &y, &z) == 3 && x > 0 && y > 0 && z > 0 && x + y + z <= 148) {
        out[0] = x; out[1] = y; out[2] = z;
      }
    };
    parse("PARROT_GROUPS_F", M.grp_f, 64, 40, 44);
    parse("PARROT_GROUPS_B", M.grp_b, 64, 36, 48);   // (layer 3 carries the largest chunk dgrads)
  }
  M.att_slices = H / ATT_KS;
  M.falloc("att_hat_part", (long long)M.att_slices * B * 3 * d.A);
  M.falloc("w", (long long)(T + 1) * B * d.C);
  Plane pw = M.make_plane("w", Np, d.C, T + 1);
  M.map_scan["w"] = M.make_map(pw, 3, Np);
  M.map_plain["w"] = M.make_map(pw, 2, NT);
  M.falloc("kappa", (long long)(T + 1) * B * d.A);
  M.falloc("phi", (long long)T * B * d.U);
  M.falloc("ab", (long long)T * B * 2 * d.A);
  M.falloc("att_e", (long long)T * B * 3 * d.A);
  M.falloc("att_hat", (long long)B * 3 * d.A);
  if (d.weak) {
    Plane px = M.make_plane("xin", Np, d.D, d.sampling ? T + 1 : T);
    M.map_scan["xin"] = M.make_map(px, 3, Np);
    if (train && !d.ln) M.map_plain["xin"] = M.make_map(px, 2, NT);
  }
  if (train) {
    X.dw = M.falloc("dw", (long long)(T + 1) * B * d.C);
    M.falloc("dk_carry", (long long)B * d.A);
    M.falloc("datt", (long long)T * B * 3 * d.A);
    M.make_plane("datt", Np, d.Ap, T);
    M.falloc("dctx", (long long)B * d.U * d.C);
  }
  // ---- readout / emitter
  const int TR = d.sampling ? 1 : T;  // the sampler keeps one step of readouts
  M.falloc("ro", (long long)TR * B * d.R);
  Plane pro = M.make_plane("ro", Np, d.R, TR);
  M.map_plain["ro"] = M.make_map(pro, 2, d.sampling ? Np : NT);
  M.falloc("pred", (long long)TR * B * d.Dtot);
  M.falloc("cost_tb", (long long)TR * B);
  if (d.sampling) {
    // staged copies of the per-call inputs: the sampling loop is replayed as a CUDA graph whose nodes read these
    M.alloc("samp_in_labels", (size_t)B * d.U * 4);
    M.falloc("samp_in_lmask", (long long)B * d.U);
    M.alloc("samp_in_speaker", (size_t)B * 4);
    M.falloc("samp_in_unis", (long long)T * B);
    M.falloc("samp_in_normals", (long long)T * B * d.D);
    M.alloc("samp_in_seed", 16);
    M.falloc("samp_x", (long long)T * B * d.D);
    M.falloc("samp_pi", (long long)T * B * (d.gmm ? d.K : d.D));
  }
  if (train) {
    M.falloc("next_x", (long long)T * B * d.D);
    M.falloc("dpred", (long long)T * B * d.Dtot);
    Plane pdp = M.make_plane("dpred", Np, d.Dtp, T);
    M.map_plain["dpred"] = M.make_map(pdp, 2, NT);
    M.falloc("dread", (long long)T * B * d.R);
    Plane pdr = M.make_plane("dread", Np, d.R, T);
    M.map_plain["dread"] = M.make_map(pdr, 2, NT);
    // transposed operand planes for the weight gradients: [features][samples]
    auto tplane = [&](const std::string& nm, int feat_pitch, long long samples, bool as_a) {
      Plane p = M.make_plane("T." + nm, rup(feat_pitch, 128), (int)samples, 1);
      if (as_a) M.map_tA[nm] = M.make_map(p, 2, 128);
      else M.map_tB[nm] = M.make_map(p, 2, WNT);
    };
    for (int l = 0; l < 3; ++l) {
      tplane("h" + LN(l), d.Hp, (long long)(T + 1) * Np, true);
      tplane("rh" + LN(l), d.Hp, (long long)T * Np, true);
      tplane("da" + LN(l), 3 * d.Hp, (long long)T * Np, false);
    }
    tplane("w", d.Cp, (long long)(T + 1) * Np, true);
    if (d.weak) tplane("xin", d.Dp, (long long)T * Np, true);
    tplane("ro", d.Rp, (long long)T * Np, true);
    tplane("dread", d.Rp, (long long)T * Np, false);
    tplane("dpred", d.Dtp, (long long)T * Np, false);
    tplane("datt", d.Ap, (long long)T * Np, false);
    if (d.ln) {
      // layer_norm=True: pre-norm Fork outputs (kept for the backward), per-step pre-activation terms, and the
      // gradients wrt the pre-norm values with their operand planes
      for (int l = 0; l < 3; ++l) {
        M.falloc("preT" + LN(l), (long long)T * B * 3 * H);
        M.falloc("ropre" + LN(l), (long long)T * B * d.R);
        M.falloc("dro" + LN(l), (long long)T * B * d.R);
        Plane pr = M.make_plane("dro" + LN(l), Np, d.R, T);
        M.map_plain["dro" + LN(l)] = M.make_map(pr, 2, NT);
        tplane("dro" + LN(l), d.Rp, (long long)T * Np, false);
        if ((l == 0 && d.weak) || (l > 0 && d.full)) {
          M.falloc("qfb" + LN(l), (long long)T * B * 3 * H);
          M.falloc("dfb" + LN(l), (long long)T * B * 3 * H);
          M.make_plane("dfb" + LN(l), Np, 3 * d.Hp, T);
          tplane("dfb" + LN(l), 3 * d.Hp, (long long)T * Np, false);
        }
        if (d.spk) { M.falloc("spk_pre" + LN(l), (long long)B * 3 * H); }
      }
      if (d.spk) M.falloc("spk_tmp", (long long)B * 3 * H);
      for (const char* nm : {"12", "13", "23"}) {
        M.falloc(std::string("q") + nm, (long long)T * B * 3 * H);
        M.falloc(std::string("dq") + nm, (long long)T * B * 3 * H);
        Plane pq = M.make_plane(std::string("dq") + nm, Np, 3 * d.Hp, T);
        M.map_scan[std::string("dq") + nm] = M.make_map(pq, 3, Np);
        tplane(std::string("dq") + nm, 3 * d.Hp, (long long)T * Np, false);
      }
      if (d.weak) M.map_plain["xin"] = M.make_map(M.planes.at("xin"), 2, NT);
    }
    M.alloc("opt_scratch", 1024 * 8);
    M.falloc("bias_scratch", (long long)std::max(3 * H, d.R) + d.Dtot + 128);
    M.falloc("colsum_scratch", (long long)COLSUM_CHUNKS * 4096);
    if (d.enc) M.falloc("sgemm_scratch", (long long)SGEMM_SPLIT * d.E * 2 * d.E);
  }
  M.alloc("gemm_scratch", 1024 * 8);
  if (d.sampling && d.ln) {
    // sampler with layer_norm: one step of pre-activation terms / pre-norm Fork outputs
    for (int l = 0; l < 3; ++l) {
      M.falloc("preT" + LN(l), (long long)B * 3 * H);
      M.falloc("ropre" + LN(l), (long long)B * d.R);
      if ((l == 0 && d.weak) || (l > 0 && d.full)) M.falloc("qfb" + LN(l), (long long)B * 3 * H);
      if (d.spk) M.falloc("spk_pre" + LN(l), (long long)B * 3 * H);
    }
    for (const char* nm : {"q12", "q13", "q23"}) M.falloc(nm, (long long)B * 3 * H);
  }
  if (d.ln && !M.dry) {
    for (int l = 0; l < 3; ++l) X.L[l].base = M.fbuf("preT" + LN(l));
    X.base_tstride = train ? (long long)B * 3 * H : 0;
  }

  // ============================ job tables ============================
  if (train && !d.ln) {
    // chunk-lagged wavefront (kernels.cuh): layer l runs l * Tc steps behind layer 1
    const int Tc = M.Tc;
    std::vector<Job> A, Bj;
    for (int l = 0; l < 3; ++l) build_fwd_layer_jobs(M, A, l, true, l * Tc, false, true);
    for (int l = 0; l < 3; ++l) build_fwd_layer_jobs(M, Bj, l, false, l * Tc, false, true);
    push_table(M, "fwdA", A, Np, PB_SPLIT_TARGET);
    push_table(M, "fwdB", Bj, Np, PB_SPLIT_TARGET);
    // hoisted products -> pre_l[t][b][3H] (plain stores, biases stay in base_l).  (source plane, first plane row,
    // pack prefix, k blocks); "_inputs" packs fill features [0, H), "_gates" packs [H, 3H).
    struct HSrc { std::string plane; int row0; std::string pack; int nkb; };
    auto hoist = [&](std::vector<Job>& js, int layer, long long n_samples, const std::vector<HSrc>& src, int lag) {
      const std::string l = LN(layer);
      for (int part = 0; part < 2; ++part) {
        const int rows = part == 0 ? H : 2 * H, f0 = part == 0 ? 0 : H;
        PlainArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.scale = 1.0f;
        pa.out = M.dry ? nullptr : M.fbuf("pre" + l);
        pa.ldo = 3 * H; pa.n_pad = Np; pa.n_valid = B; pa.n_total = T * Np;
        // the time-constant term base_l[b][3H] (summed Fork biases + speaker) is folded into pre_l here, so that
        // the scan epilogues read ONE pre-activation array
        pa.bias = M.dry ? nullptr : M.fbuf("base" + l);
        pa.rowbias_ld = 3 * H;
        std::vector<PlainSeg> segs;
        for (auto& sc : src)
          segs.push_back({M.packs[sc.pack + "/fork_rnn" + l + (part == 0 ? "_inputs" : "_gates")].fwd_map, 0,
                          M.map_plain[sc.plane], sc.row0, 0, sc.nkb});
        const size_t first = js.size();
        build_plain_jobs(js, rows, f0, n_samples, segs, pa);
        for (size_t i = first; i < js.size(); ++i) js[i].lag = lag;
      }
    };
    if (d.weak) {   // layer 1: x_{t-1} . out_to_h1, all frames at once before the scan
      std::vector<Job> js;
      hoist(js, 0, (long long)T * Np, {{"xin", 0, "/out_to_h1", d.Dp / 64}}, 0);
      sort_by_sample_tile(js);
      push_table(M, "hoist1", js, NT);
    }
    {
      // one chunk of Tc steps: pre2 from (h1_t, w_t) of chunk e, pre3 from (h1_t, h2_t, w_t) of chunk e - 1
      std::vector<Job> js;
      std::vector<HSrc> s3 = {{"h1", Np, "/h1_to_h3", d.Hp / 64}, {"h2", Np, "/h2_to_h3", d.Hp / 64},
                              {"w", Np, "/inp_to_h3", d.Cp / 64}};
      std::vector<HSrc> s2 = {{"h1", Np, "/h1_to_h2", d.Hp / 64}, {"w", Np, "/inp_to_h2", d.Cp / 64}};
      if (d.full) {
        s3.push_back({"xin", 0, "/out_to_h3", d.Dp / 64});
        s2.push_back({"xin", 0, "/out_to_h2", d.Dp / 64});
      }
      hoist(js, 2, (long long)Tc * Np, s3, 1);   // the longer jobs first
      hoist(js, 1, (long long)Tc * Np, s2, 0);
      push_table(M, "chunkF", js, NT, 0, Tc * Np);
      // grouped persistent scan (kernels.cuh): one table set per layer group, every job at lag 0 (a group's tick IS
      // its layer's step), split over the CTAs of the group, own split-K scratch (the groups run concurrently)
      for (int l = 0; l < 3; ++l) {
        std::vector<Job> ga, gb;
        build_fwd_layer_jobs(M, ga, l, true, 0, false, true);
        build_fwd_layer_jobs(M, gb, l, false, 0, false, true);
        push_table(M, "gA" + LN(l), ga, Np, M.grp_f[l], 0, true);
        push_table(M, "gB" + LN(l), gb, Np, M.grp_f[l], 0, true);
        assign_resident(M, "gA" + LN(l), "gB" + LN(l));
      }
      std::vector<Job> c2, c3;
      hoist(c2, 1, (long long)Tc * Np, s2, 0);
      hoist(c3, 2, (long long)Tc * Np, s3, 0);
      sort_by_sample_tile(c2);
      sort_by_sample_tile(c3);
      push_table(M, "gC2", c2, NT, 0, Tc * Np);
      push_table(M, "gC3", c3, NT, 0, Tc * Np);
    }
  } else if (train) {
    // layer_norm: one layer at a time (lag 0); the normalised Fork outputs reach the epilogues through preT
    for (int l = 0; l < 3; ++l) {
      std::vector<Job> A, Bj;
      build_fwd_layer_jobs(M, A, l, true, 0, true);
      build_fwd_layer_jobs(M, Bj, l, false, 0, true);
      push_table(M, "lnA" + LN(l), A, Np, 148);
      push_table(M, "lnB" + LN(l), Bj, Np, 148);
    }
  } else {
    for (int l = 0; l < 3; ++l) {
      std::vector<Job> A, Bj;
      build_fwd_layer_jobs(M, A, l, true, 0, d.ln);
      build_fwd_layer_jobs(M, Bj, l, false, 0, d.ln);
      // split over K like the training scan phases: a sampling step is a chain of ~11 dependent launches, and an
      // unsplit tile contracts its whole K (up to 36 k-blocks) alone -- 24 us per launch on 16-24 of the 148 SMs
      push_table(M, "sampA" + LN(l), A, Np, 148);
      push_table(M, "sampB" + LN(l), Bj, Np, 148);
    }
  }
  auto planes_of = [&](const std::string& nm) -> Plane& { return M.planes.at(nm); };
  // readouts (model.py:739-753): [h1,h2,h3,w] . [W1;W2;W3;Wa]
  {
    std::vector<Job> js;
    PlainArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.scale = 1.0f;
    pa.out = M.dry ? nullptr : M.fbuf("ro");
    pa.bias = M.dry ? nullptr : M.fbuf("bias_ro");
    pa.hi = planes_of("ro").hi; pa.lo = planes_of("ro").lo;
    pa.ldo = d.R; pa.ldp = planes_of("ro").pitch;
    pa.n_pad = Np; pa.n_valid = B;
    pa.flags = PF_PLANE_PADDED | (d.spk ? PF_ACC : 0);
    if (train && d.ln) {
      // layer_norm (model.py:743-746): the three state readouts are normalised one by one, so they are
      // produced separately (pre-norm, with their own bias); the attention readout is added afterwards.
      pa.n_total = T * Np;
      std::vector<Job> pre;
      for (int l = 0; l < 3; ++l) {
        PlainArgs pl;
        memset(&pl, 0, sizeof pl);
        pl.scale = 1.0f;
        pl.out = M.dry ? nullptr : M.fbuf("ropre" + LN(l));
        pl.bias = M.dry ? nullptr : M.pp("/h" + LN(l) + "_to_readout.b");
        pl.ldo = d.R; pl.n_pad = Np; pl.n_valid = B; pl.n_total = T * Np;
        std::vector<PlainSeg> segs = {
            {M.packs["/h" + LN(l) + "_to_readout"].fwd_map, 0, M.map_plain["h" + LN(l)], Np, 0, d.Hp / 64}};
        build_plain_jobs(pre, d.R, 0, (long long)T * Np, segs, pl);
      }
      sort_by_sample_tile(pre);
      push_table(M, "ln_ro", pre, NT);
      pa.bias = M.dry ? nullptr : M.pp("/att_to_readout.b");
      pa.flags = PF_PLANE_PADDED | PF_ACC;
      std::vector<PlainSeg> segs = {{M.packs["/att_to_readout"].fwd_map, 0, M.map_plain["w"], Np, 0, d.Cp / 64}};
      build_plain_jobs(js, d.R, 0, (long long)T * Np, segs, pa);
      push_table(M, "ln_ro_att", js, NT);
    } else if (train) {
      pa.n_total = T * Np;
      std::vector<PlainSeg> segs;
      for (int l = 0; l < 3; ++l)
        segs.push_back({M.packs["/h" + LN(l) + "_to_readout"].fwd_map, 0, M.map_plain["h" + LN(l)], Np, 0, d.Hp / 64});
      segs.push_back({M.packs["/att_to_readout"].fwd_map, 0, M.map_plain["w"], Np, 0, d.Cp / 64});
      build_plain_jobs(js, d.R, 0, (long long)T * Np, segs, pa);
      push_table(M, "readout", js, NT);
    } else if (d.ln) {
      // sampler with layer_norm (model.py:992-1003): three pre-norm readouts, then the attention readout on top
      pa.n_total = Np;
      std::vector<Job> pre;
      for (int l = 0; l < 3; ++l)
        for (int mt = 0; mt < cdiv(d.R, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_PLAIN; j.row0 = mt * 128; j.m_valid = std::min(128, d.R - mt * 128);
          j.nseg = 1;
          j.seg[0] = mkseg(M.packs["/h" + LN(l) + "_to_readout"].fwd_map, mt * 128, 0, M.map_scan["h" + LN(l)], 0, 0, 1,
                           d.Hp / 64);
          j.pa.out = M.dry ? nullptr : M.fbuf("ropre" + LN(l));
          j.pa.bias = M.dry ? nullptr : M.pp("/h" + LN(l) + "_to_readout.b");
          j.pa.ldo = d.R; j.pa.n_pad = Np; j.pa.n_valid = B; j.pa.n_total = Np;
          pre.push_back(j);
        }
      push_table(M, "ln_ro", pre, Np);
      pa.bias = M.dry ? nullptr : M.pp("/att_to_readout.b");
      pa.flags = PF_PLANE_PADDED | PF_ACC;
      for (int mt = 0; mt < cdiv(d.R, 128); ++mt) {
        Job j = blank_job();
        j.epi = EPI_PLAIN; j.row0 = mt * 128; j.m_valid = std::min(128, d.R - mt * 128);
        j.nseg = 1;
        j.seg[0] = mkseg(M.packs["/att_to_readout"].fwd_map, mt * 128, 0, M.map_scan["w"], 0, 0, 1, d.Cp / 64);
        j.pa = pa;
        js.push_back(j);
      }
      push_table(M, "ln_ro_att", js, Np);
    } else {
      // sampler: one step, operands taken from the scan maps at slot t+1
      pa.n_total = Np;
      for (int mt = 0; mt < cdiv(d.R, 128); ++mt) {
        Job j = blank_job();
        j.epi = EPI_PLAIN; j.row0 = mt * 128; j.m_valid = std::min(128, d.R - mt * 128);
        int ns = 0;
        for (int l = 0; l < 3; ++l)
          j.seg[ns++] = mkseg(M.packs["/h" + LN(l) + "_to_readout"].fwd_map, mt * 128, 0, M.map_scan["h" + LN(l)], 0,
                              0, 1, d.Hp / 64);
        j.seg[ns++] = mkseg(M.packs["/att_to_readout"].fwd_map, mt * 128, 0, M.map_scan["w"], 0, 0, 1, d.Cp / 64);
        j.nseg = ns;
        j.pa = pa;
        js.push_back(j);
      }
      push_table(M, "readout", js, Np, 148);
    }
  }
  if (d.ln) {
    // layer_norm forward side products, per step: q12/q13 (from h1_t), q23 (from h2_t) and, in the sampler, the
    // feedback Forks of the frame just emitted (training: the feedback Forks run over all frames at once)
    const long long tstride = train ? (long long)B * 3 * H : 0;
    auto qjobs = [&](std::vector<Job>& js, const std::string& fork, const std::string& fk_l, const std::string& src,
                     int slot, int nkb, const std::string& dst) {
      for (int part = 0; part < 2; ++part) {
        const std::string pk = fork + "/fork_rnn" + fk_l + (part == 0 ? "_inputs" : "_gates");
        const int rows = part == 0 ? H : 2 * H, f0 = part == 0 ? 0 : H;
        for (int mt = 0; mt < cdiv(rows, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_PLAIN; j.row0 = f0 + mt * 128; j.m_valid = std::min(128, rows - mt * 128);
          j.nseg = 1;
          j.seg[0] = mkseg(M.packs[pk].fwd_map, mt * 128, 0, M.map_scan[src], 0, 0, slot, nkb);
          j.pa.out = M.dry ? nullptr : M.fbuf(dst);
          j.pa.bias = M.dry ? nullptr : M.pp(pk + ".b") - f0;
          j.pa.ldo = 3 * H; j.pa.out_tstride = tstride;
          j.pa.n_pad = Np; j.pa.n_valid = B; j.pa.n_total = Np;
          js.push_back(j);
        }
      }
    };
    {
      std::vector<Job> js;
      qjobs(js, "/h1_to_h2", "2", "h1", 1, d.Hp / 64, "q12");
      qjobs(js, "/h1_to_h3", "3", "h1", 1, d.Hp / 64, "q13");
      push_table(M, "lnQ1", js, Np);
    }
    {
      std::vector<Job> js;
      qjobs(js, "/h2_to_h3", "3", "h2", 1, d.Hp / 64, "q23");
      push_table(M, "lnQ2", js, Np);
    }
    if (!train) {
      std::vector<Job> js;
      for (int l = 0; l < 3; ++l)
        if ((l == 0 && d.weak) || (l > 0 && d.full)) qjobs(js, "/out_to_h" + LN(l), LN(l), "xin", 0, d.Dp / 64, "qfb" + LN(l));
      push_table(M, "ln_fb", js, Np);
    } else {
      std::vector<Job> js;
      for (int l = 0; l < 3; ++l) {
        if (!((l == 0 && d.weak) || (l > 0 && d.full))) continue;
        for (int part = 0; part < 2; ++part) {
          const std::string pk = "/out_to_h" + LN(l) + "/fork_rnn" + LN(l) + (part == 0 ? "_inputs" : "_gates");
          const int rows = part == 0 ? H : 2 * H, f0 = part == 0 ? 0 : H;
          PlainArgs pa;
          memset(&pa, 0, sizeof pa);
          pa.scale = 1.0f;
          pa.out = M.dry ? nullptr : M.fbuf("qfb" + LN(l));
          pa.bias = M.dry ? nullptr : M.pp(pk + ".b") - f0;
          pa.ldo = 3 * H; pa.n_pad = Np; pa.n_valid = B; pa.n_total = T * Np;
          std::vector<PlainSeg> segs = {{M.packs[pk].fwd_map, 0, M.map_plain["xin"], 0, 0, d.Dp / 64}};
          build_plain_jobs(js, rows, f0, (long long)T * Np, segs, pa);
        }
      }
      sort_by_sample_tile(js);
      push_table(M, "ln_fb", js, NT);
    }
  }
  // output layer (model.py:755, 766)
  {
    std::vector<Job> js;
    for (size_t f = 0; f < out_forks.size(); ++f) {
      PlainArgs pa;
      memset(&pa, 0, sizeof pa);
      pa.scale = 1.0f;
      pa.out = M.dry ? nullptr : M.fbuf("pred");
      pa.bias = M.dry ? nullptr : M.fbuf("bias_out");
      pa.ldo = d.Dtot; pa.n_pad = Np; pa.n_valid = B;
      pa.n_total = train ? T * Np : Np;
      pa.flags = d.spk ? PF_ACC : 0;
      std::vector<PlainSeg> segs = {{M.packs[out_forks[f]].fwd_map, 0, M.map_plain["ro"], 0, 0, d.Rp / 64}};
      if (train) build_plain_jobs(js, out_dim[f], out_off[f], (long long)T * Np, segs, pa);
      else {
        for (int mt = 0; mt < cdiv(out_dim[f], 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_PLAIN; j.row0 = out_off[f] + mt * 128; j.m_valid = std::min(128, out_dim[f] - mt * 128);
          j.nseg = 1;
          j.seg[0] = mkseg(M.packs[out_forks[f]].fwd_map, mt * 128, 0, M.map_plain["ro"], 0, 0, NO_SLOT, d.Rp / 64);
          j.pa = pa;
          js.push_back(j);
        }
      }
    }
    if (train) sort_by_sample_tile(js);
    push_table(M, "output", js, train ? NT : Np, train ? 0 : 148);
  }
  if (train) {
    // dread = dpred . Wout^T
    {
      std::vector<Job> js;
      PlainArgs pa;
      memset(&pa, 0, sizeof pa);
      pa.scale = 1.0f;
      pa.out = M.dry ? nullptr : M.fbuf("dread");
      pa.hi = planes_of("dread").hi; pa.lo = planes_of("dread").lo;
      pa.ldo = d.R; pa.ldp = planes_of("dread").pitch;
      pa.n_pad = Np; pa.n_valid = B; pa.n_total = T * Np; pa.flags = PF_PLANE_PADDED;
      std::vector<PlainSeg> segs;
      for (size_t f = 0; f < out_forks.size(); ++f)
        segs.push_back({M.packs[out_forks[f]].bwd_map, 0, M.map_plain["dpred"], 0, out_poff[f], cdiv(out_dim[f], 64)});
      build_plain_jobs(js, d.R, 0, (long long)T * Np, segs, pa);
      push_table(M, "dread", js, NT);
    }
    // dh_l[slot t+1] = dread . W_l^T ; dw[slot t+1] = dread . Wa^T   (fresh stores)
    {
      std::vector<Job> js;
      for (int l = 0; l < 4; ++l) {
        PlainArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.scale = 1.0f;
        const int F = l < 3 ? H : d.C;
        pa.out = M.dry ? nullptr : (l < 3 ? M.ctx.L[l].dh + (long long)B * H : M.ctx.dw + (long long)B * d.C);
        pa.ldo = F; pa.n_pad = Np; pa.n_valid = B; pa.n_total = T * Np;
        const std::string pk = l < 3 ? "/h" + LN(l) + "_to_readout" : "/att_to_readout";
        const std::string src = (d.ln && l < 3) ? "dro" + LN(l) : std::string("dread");
        std::vector<PlainSeg> segs = {{M.packs[pk].bwd_map, 0, M.map_plain[src], 0, 0, d.Rp / 64}};
        build_plain_jobs(js, F, 0, (long long)T * Np, segs, pa);
      }
      sort_by_sample_tile(js);
      push_table(M, "dh_readout", js, NT);
    }
    if (d.ln) {
      // layer_norm backward: one table per layer and product (no wavefront: the norm sits between the layers)
      for (int l = 0; l < 3; ++l) {
        std::vector<Job> js;
        for (int mt = 0; mt < cdiv(H, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_BWD_RH; j.layer = l; j.lag = 0; j.row0 = mt * 128; j.m_valid = std::min(128, H - mt * 128);
          j.nseg = 1;
          j.seg[0] = mkseg(M.packs["/rnn" + LN(l) + ".state_to_state"].bwd_map, mt * 128, 0, M.map_scan["da" + LN(l)],
                           0, 0, 0, d.Hp / 64);
          js.push_back(j);
        }
        push_table(M, "ln_bwd1_" + LN(l), js, Np, 148);
      }
      // (pack, plane holding dY, gate block?) -> destination (aux, slot offset)
      struct Src { std::string pack, plane; };
      auto add = [&](std::vector<Job>& js, int rows, int aux, int slot_off, const std::vector<Src>& src) {
        for (int mt = 0; mt < cdiv(rows, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_BWD_STATE; j.aux = aux; j.lag = 0; j.row0 = mt * 128; j.m_valid = std::min(128, rows - mt * 128);
          j.pa.n_pad = slot_off;
          int ns = 0;
          for (auto& sname : src) {
            const bool is_gate = sname.pack.find("_gates") != std::string::npos;
            j.seg[ns++] = mkseg(M.packs[sname.pack].bwd_map, mt * 128, 0, M.map_scan[sname.plane], 0,
                                is_gate ? d.Hp : 0, 0, (is_gate ? 2 : 1) * d.Hp / 64);
          }
          j.nseg = ns;
          js.push_back(j);
        }
      };
      {
        std::vector<Job> js;
        add(js, H, 2, 0, {{"/rnn3.state_to_gates", "da3"}});
        add(js, d.C, 3, 1, {{"/inp_to_h3/fork_rnn3_inputs", "da3"}, {"/inp_to_h3/fork_rnn3_gates", "da3"}});
        add(js, H, 1, 1, {{"/h2_to_h3/fork_rnn3_inputs", "dq23"}, {"/h2_to_h3/fork_rnn3_gates", "dq23"}});
        add(js, H, 0, 1, {{"/h1_to_h3/fork_rnn3_inputs", "dq13"}, {"/h1_to_h3/fork_rnn3_gates", "dq13"}});
        push_table(M, "ln_bwd2_3", js, Np, 148);
      }
      {
        std::vector<Job> js;
        add(js, H, 1, 0, {{"/rnn2.state_to_gates", "da2"}});
        add(js, d.C, 3, 1, {{"/inp_to_h2/fork_rnn2_inputs", "da2"}, {"/inp_to_h2/fork_rnn2_gates", "da2"}});
        add(js, H, 0, 1, {{"/h1_to_h2/fork_rnn2_inputs", "dq12"}, {"/h1_to_h2/fork_rnn2_gates", "dq12"}});
        push_table(M, "ln_bwd2_2", js, Np, 148);
      }
      {
        std::vector<Job> js;
        add(js, H, 0, 0, {{"/rnn1.state_to_gates", "da1"}});
        add(js, d.C, 3, 0, {{"/inp_to_h1/fork_rnn1_inputs", "da1"}, {"/inp_to_h1/fork_rnn1_gates", "da1"}});
        push_table(M, "ln_bwd2_1", js, Np, 148);
      }
    }
    // backward scan, product 1: d(r*h) = da_c . Ws^T.  Reverse chunk-lagged wavefront: layer 3 leads, layer l runs
    // (2 - l) * Tc steps behind it.
    if (!d.ln) {
      const int Tc = M.Tc;
      std::vector<Job> js;
      for (int l = 0; l < 3; ++l)
        for (int mt = 0; mt < cdiv(H, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_BWD_RH; j.layer = l; j.lag = (2 - l) * Tc; j.row0 = mt * 128; j.m_valid = std::min(128, H - mt * 128);
          j.nseg = 1;
          j.seg[0] = mkseg(M.packs["/rnn" + LN(l) + ".state_to_state"].bwd_map, mt * 128, 0, M.map_scan["da" + LN(l)],
                           0, 0, 0, d.Hp / 64);
          js.push_back(j);
        }
      push_table(M, "bwd1", js, Np, PB_SPLIT_TARGET);
      for (int l = 0; l < 3; ++l) {   // grouped backward scan: layer tables at lag 0
        std::vector<Job> gj;
        for (auto& j : js)
          if (j.layer == l) { gj.push_back(j); gj.back().lag = 0; }
        push_table(M, "hA" + LN(l), gj, Np, M.grp_b[l], 0, true);
      }
    }
    // backward scan, product 2: the dgrads that stay in the recurrence -- da_g . Wg^T into dh_l[slot t] (the state
    // entering step t) and, for layer 1, da_1 . Wi1^T into dw[slot t] (layer 1 consumes w_{t-1}).
    if (!d.ln) {
      const int Tc = M.Tc;
      std::vector<Job> js;
      const int gk = d.Hp;  // column offset of the gate block inside the da planes
      auto add = [&](int rows, int aux, int layer, const std::vector<std::string>& packs_) {
        for (int mt = 0; mt < cdiv(rows, 128); ++mt) {
          Job j = blank_job();
          j.epi = EPI_BWD_STATE; j.aux = aux; j.layer = layer; j.lag = (2 - layer) * Tc;
          j.row0 = mt * 128; j.m_valid = std::min(128, rows - mt * 128);
          j.pa.n_pad = 0;   // destination slot = t
          int ns = 0;
          for (auto& pk : packs_) {
            const bool is_gate = pk.find("_gates") != std::string::npos;
            j.seg[ns++] = mkseg(M.packs[pk].bwd_map, mt * 128, 0, M.map_scan["da" + LN(layer)], 0, is_gate ? gk : 0, 0,
                                (is_gate ? 2 : 1) * d.Hp / 64);
          }
          j.nseg = ns;
          js.push_back(j);
        }
      };
      add(H, 2, 2, {"/rnn3.state_to_gates"});
      add(H, 1, 1, {"/rnn2.state_to_gates"});
      add(H, 0, 0, {"/rnn1.state_to_gates"});
      add(d.C, 3, 0, {"/inp_to_h1/fork_rnn1_inputs", "/inp_to_h1/fork_rnn1_gates"});
      push_table(M, "bwd2", js, Np, PB_SPLIT_TARGET);
      for (int l = 0; l < 3; ++l) {
        std::vector<Job> gj;
        for (auto& j : js)
          if (j.layer == l) {
            gj.push_back(j); gj.back().lag = 0;
            // groups 1 / 2: the finish of the own-layer state dgrad also runs the GRU pre-pass of step t - 1
            if (l > 0 && j.aux == l && H % 4 == 0 && d.Hp % 4 == 0 && !getenv("PARROT_NO_FUSED_PRE")) gj.back().pa.flags |= QF_FUSED_PRE;
          }
        push_table(M, "hB" + LN(l), gj, Np, M.grp_b[l], 0, true);
        assign_resident(M, "hA" + LN(l), "hB" + LN(l));
      }
      // hoisted dgrads, one range of Tc steps at a time (accumulated into slot t + 1 of the consumer's gradient):
      // event e: from da3 of range e into dh2 / dh1 / dw ; from da2 of range e - 1 into dh1 / dw
      std::vector<Job> cj;
      auto chunk = [&](int rows, float* out, int F, int layer, const std::string& fork, int lag) {
        PlainArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.scale = 1.0f;
        pa.out = out;
        pa.ldo = F; pa.n_pad = Np; pa.n_valid = B; pa.n_total = T * Np; pa.flags = PF_ACC;
        const std::string l = LN(layer);
        std::vector<PlainSeg> segs = {
            {M.packs[fork + "/fork_rnn" + l + "_inputs"].bwd_map, 0, M.map_plain["da" + l], 0, 0, d.Hp / 64},
            {M.packs[fork + "/fork_rnn" + l + "_gates"].bwd_map, 0, M.map_plain["da" + l], 0, gk, 2 * d.Hp / 64}};
        const size_t first = cj.size();
        build_plain_jobs(cj, rows, 0, (long long)Tc * Np, segs, pa);
        for (size_t i = first; i < cj.size(); ++i) cj[i].lag = lag;
      };
      float* dh1s = M.dry ? nullptr : M.ctx.L[0].dh + (long long)B * H;
      float* dh2s = M.dry ? nullptr : M.ctx.L[1].dh + (long long)B * H;
      float* dws = M.dry ? nullptr : M.ctx.dw + (long long)B * d.C;
      chunk(H, dh2s, H, 2, "/h2_to_h3", 0);
      chunk(H, dh1s, H, 2, "/h1_to_h3", 0);
      chunk(d.C, dws, d.C, 2, "/inp_to_h3", 0);
      chunk(H, dh1s, H, 1, "/h1_to_h2", 1);
      chunk(d.C, dws, d.C, 1, "/inp_to_h2", 1);
      push_table(M, "chunkB", cj, NT, 0, Tc * Np);
      for (int l = 1; l < 3; ++l) {   // grouped backward scan: chunk dgrads of layer l + 1, event = the group's own range
        std::vector<Job> gj;
        for (auto& j : cj)
          if (j.lag == 2 - l) { gj.push_back(j); gj.back().lag = 0; }
        sort_by_sample_tile(gj);
        push_table(M, "hC" + LN(l), gj, NT, 0, Tc * Np);
      }
    }
    // weight gradients: dW[in][out] = sum_samples X[s][in] * dY[s][out]
    {
      auto wg = [&](const std::string& xt, int a_k, const std::string& dyt, int b_row, const std::string& pname) {
        const PInfo& pi = M.P.get("/parrot" + pname);
        M.wgrads.push_back(WGrad{M.map_tA[xt], a_k, M.map_tB[dyt], b_row, pi.rows, pi.cols, pi.off});
      };
      for (int l = 0; l < 3; ++l) {
        const std::string s = LN(l);
        wg("rh" + s, 0, "da" + s, 0, "/rnn" + s + ".state_to_state");
        wg("h" + s, 0, "da" + s, d.Hp, "/rnn" + s + ".state_to_gates");
        const int wshift = l == 0 ? 0 : Np;
        wg("w", wshift, "da" + s, 0, "/inp_to_h" + s + "/fork_rnn" + s + "_inputs");
        wg("w", wshift, "da" + s, d.Hp, "/inp_to_h" + s + "/fork_rnn" + s + "_gates");
        wg("h" + s, Np, d.ln ? "dro" + s : std::string("dread"), 0, "/h" + s + "_to_readout.W");
        if ((l == 0 && d.weak) || (l > 0 && d.full)) {
          const std::string dfb = d.ln ? "dfb" + s : "da" + s;
          wg("xin", 0, dfb, 0, "/out_to_h" + s + "/fork_rnn" + s + "_inputs");
          wg("xin", 0, dfb, d.Hp, "/out_to_h" + s + "/fork_rnn" + s + "_gates");
        }
      }
      // layer_norm: the gradient reaching a Fork is the one behind its own normalisation
      const std::string g12 = d.ln ? "dq12" : "da2", g13 = d.ln ? "dq13" : "da3", g23 = d.ln ? "dq23" : "da3";
      wg("h1", Np, g12, 0, "/h1_to_h2/fork_rnn2_inputs"); wg("h1", Np, g12, d.Hp, "/h1_to_h2/fork_rnn2_gates");
      wg("h1", Np, g13, 0, "/h1_to_h3/fork_rnn3_inputs"); wg("h1", Np, g13, d.Hp, "/h1_to_h3/fork_rnn3_gates");
      wg("h2", Np, g23, 0, "/h2_to_h3/fork_rnn3_inputs"); wg("h2", Np, g23, d.Hp, "/h2_to_h3/fork_rnn3_gates");
      wg("w", Np, "dread", 0, "/att_to_readout.W");
      for (size_t f = 0; f < out_forks.size(); ++f) wg("ro", 0, "dpred", out_poff[f], out_forks[f] + ".W");
      wg("h1", Np, "datt", 0, "/h1_to_att/fork_alpha.W");
      wg("h1", Np, "datt", d.A, "/h1_to_att/fork_beta.W");
      wg("h1", Np, "datt", 2 * d.A, "/h1_to_att/fork_kappa.W");
      std::vector<Job> js;
      const int nkb = cdiv((long long)T * Np, 64);
      for (auto& g : M.wgrads)
        for (int mt = 0; mt < cdiv(g.in, 128); ++mt)
          for (int nt = 0; nt < cdiv(g.out, WNT); ++nt) {
            Job j = blank_job();
            j.epi = EPI_PLAIN; j.row0 = mt * 128; j.m_valid = std::min(128, g.in - mt * 128); j.n0 = nt * WNT;
            j.nseg = 1;
            j.seg[0] = mkseg(g.xt_map, mt * 128, g.a_k, g.dyt_map, g.b_row + nt * WNT, 0, NO_SLOT, nkb);
            j.pa.out = M.grads ? M.grads + g.goff : nullptr;
            j.pa.ldo = g.out; j.pa.n_total = g.out; j.pa.flags = PF_TRANS; j.pa.scale = 1.0f;
            js.push_back(j);
          }
      push_table(M, "wgrad", js, WNT);
    }
  }
}

// allocations that hold the device copies of the host-built tables (must be last in build order)
static void build_device_tables(parrot_model& M) {
  M.d_maps = (CUtensorMap*)M.alloc("dev_maps", M.maps.size() * sizeof(CUtensorMap));
  M.d_raws = (MapRaw*)M.alloc("dev_raws", M.raws.size() * sizeof(MapRaw));
  M.d_jobs = (Job*)M.alloc("dev_jobs", M.jobs.size() * sizeof(Job));
  M.d_ctx = (ScanCtx*)M.alloc("dev_ctx", sizeof(ScanCtx));
  M.d_split_scratch = (float*)M.alloc("split_scratch", (std::max<size_t>(M.max_split_floats, 1) + M.uniq_split_floats) * 4);
  M.d_split_count = (unsigned int*)M.alloc("split_count", (size_t)(std::max(M.max_groups, 1) + M.uniq_split_groups) * 4);
  M.d_gridbar = (unsigned int*)M.alloc("gridbar", 1024);   // [8] counters, 128 bytes apart (grouped scans: one per group)
}

static void upload_tables(parrot_model& M, cudaStream_t st) {
  CK(cudaMemcpyAsync(M.d_maps, M.maps.data(), M.maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(M.d_raws, M.raws.data(), M.raws.size() * sizeof(MapRaw), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(M.d_jobs, M.jobs.data(), M.jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(M.d_ctx, &M.ctx, sizeof(ScanCtx), cudaMemcpyHostToDevice, st));
}

static void ensure_kernel_attrs() {
  static bool done = false;
  if (done) return;
  CK(cudaFuncSetAttribute(job_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024));
  CK(cudaFuncSetAttribute(scan_fwd_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024 + ATT_SMEM_BYTES));
  CK(cudaFuncSetAttribute(scan_bwd_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024 + ATT_SMEM_BYTES));
  CK(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(attention_proj_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(attention_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(attention_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  CK(cudaFuncSetAttribute(encoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  done = true;
}

// ------------------------------------------------------------------ small helpers
static float* g_sgemm_scratch = nullptr;    // set per model before use (workspace buffer "sgemm_scratch")
static long long g_sgemm_scratch_floats = 0;
static void sgemm(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
                  long long sbn, float* C, long long ldc, int Mr, int N, int K, const float* bias, float beta,
                  const int* gather = nullptr) {
  SGemm g;
  g.A = A; g.sam = sam; g.sak = sak; g.B = B; g.sbk = sbk; g.sbn = sbn; g.C = C; g.ldc = ldc;
  g.bias = bias; g.a_gather = gather; g.M = Mr; g.N = N; g.K = K; g.alpha = 1.0f; g.beta = beta;
  g.k_chunk = K; g.c_zstride = 0;
  dim3 grid(cdiv(N, 32), cdiv(Mr, 32));
  // long reductions with a small output (encoder weight gradients: K = B*U): split K over grid.z into partial
  // products, then add them in chunk order (deterministic)
  const long long mn = (long long)Mr * N;
  if (K >= 2048 && g_sgemm_scratch && !bias && !gather && ldc == N && (beta == 0.0f || beta == 1.0f) &&
      mn * SGEMM_SPLIT <= g_sgemm_scratch_floats) {
    g.k_chunk = rup(cdiv(K, SGEMM_SPLIT), 32);
    const int parts = cdiv(K, g.k_chunk);
    g.C = g_sgemm_scratch; g.c_zstride = mn; g.beta = 0.0f;
    grid.z = parts;
    LAUNCH(sgemm_kernel, grid, 256, 0, st, g);
    LAUNCH(colsum_kernel, cdiv(mn, 32), 256, 0, st, (const float*)g_sgemm_scratch, mn, (long long)parts, (int)mn, C,
           beta != 0.0f ? 1 : 0);
    return;
  }
  LAUNCH(sgemm_kernel, grid, 256, 0, st, g);
}
static float* g_colsum_scratch = nullptr;   // set per model before use (workspace buffer "colsum_scratch")
static void colsum(cudaStream_t st, const float* src, long long ld, long long rows, int cols, float* out, int acc) {
  if (rows >= 4096 && g_colsum_scratch && cols <= 4096) {
    const long long per = (rows + COLSUM_CHUNKS - 1) / COLSUM_CHUNKS;
    dim3 grid(cdiv(cols, 32), COLSUM_CHUNKS);
    LAUNCH(colsum_partial_kernel, grid, 256, 0, st, src, ld, rows, cols, per, g_colsum_scratch);
    LAUNCH(colsum_final_kernel, cdiv(cols, 256), 256, 0, st, g_colsum_scratch, COLSUM_CHUNKS, cols, out, acc);
    return;
  }
  LAUNCH(colsum_kernel, cdiv(cols, 32), 256, 0, st, src, ld, rows, cols, out, acc);
}
static void pack_plane(cudaStream_t st, const float* src, long long src_ld, int rows, int cols, const Plane& dst,
                       int transpose) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  dim3 block(32, 8);
  LAUNCH(pack_planes_kernel, grid, block, 0, st, src, src_ld, rows, cols, dst.hi, dst.lo, (long long)dst.pitch,
         transpose, dst.tiled_nkb);
}
static void transpose_planes(cudaStream_t st, const Plane& src, int feat_cols, const Plane& dst) {
  const long long rows = (long long)src.rows * src.slots;
  dim3 block(32, 8);
  if ((src.pitch & 1) == 0 && (dst.pitch & 1) == 0) {
    dim3 grid(cdiv(feat_cols, 64), cdiv(rows, 64));
    LAUNCH(transpose_plane64_kernel, grid, block, 0, st, src.hi, (long long)src.pitch, rows, feat_cols, dst.hi,
           (long long)dst.pitch);
    LAUNCH(transpose_plane64_kernel, grid, block, 0, st, src.lo, (long long)src.pitch, rows, feat_cols, dst.lo,
           (long long)dst.pitch);
    return;
  }
  dim3 grid(cdiv(feat_cols, 32), cdiv(rows, 32));
  LAUNCH(transpose_plane_kernel, grid, block, 0, st, src.hi, (long long)src.pitch, rows, feat_cols, dst.hi,
         (long long)dst.pitch);
  LAUNCH(transpose_plane_kernel, grid, block, 0, st, src.lo, (long long)src.pitch, rows, feat_cols, dst.lo,
         (long long)dst.pitch);
}

// ------------------------------------------------------------------ weights -> operand planes
static void pack_weights(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  for (auto& kv : M.packs) {
    WPack& w = kv.second;
    const float* src = M.params + M.P.get("/parrot" + w.name).off;
    pack_plane(st, src, w.out, w.in, w.out, w.fwd, 1);
    if (w.need_bwd) pack_plane(st, src, w.out, w.in, w.out, w.bwd, 0);
  }
  // attention projection, transposed fp32 [3A][H]
  const char* an[3] = {"alpha", "beta", "kappa"};
  for (int i = 0; i < 3; ++i) {
    const std::string n = std::string("/h1_to_att/fork_") + an[i];
    LAUNCH(transpose_f32_kernel, gs_blocks((long long)d.H * d.A), 256, 0, st, M.pp(n + ".W"), d.H, d.A,
           M.fbuf("att_wT") + (long long)i * d.A * d.H);
    CK(cudaMemcpyAsync(M.fbuf("att_b") + i * d.A, M.pp(n + ".b"), d.A * 4, cudaMemcpyDeviceToDevice, st));
  }
  // summed Fork biases per layer: [cell | gates]
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l);
    for (int part = 0; part < 2; ++part) {
      const std::string fk = part == 0 ? "_inputs.b" : "_gates.b";
      const float* src[4] = {nullptr, nullptr, nullptr, nullptr};
      int ns = 0;
      src[ns++] = M.pp("/inp_to_h" + s + "/fork_rnn" + s + fk);
      // layer_norm: the other Forks are normalised together with their own bias (model.py:31-34)
      if (l >= 1 && !d.ln) src[ns++] = M.pp("/h1_to_h" + s + "/fork_rnn" + s + fk);
      if (l == 2 && !d.ln) src[ns++] = M.pp("/h2_to_h3/fork_rnn3" + fk);
      if (!d.ln && ((l == 0 && d.weak) || (l > 0 && d.full)))
        src[ns++] = M.pp("/out_to_h" + s + "/fork_rnn" + s + fk);
      const int n = part == 0 ? d.H : 2 * d.H;
      LAUNCH(vec_sum_kernel, gs_blocks(n), 256, 0, st, M.fbuf("bias_l" + s) + (part == 0 ? 0 : d.H), n, src[0], src[1],
             src[2], src[3]);
    }
  }
  LAUNCH(vec_sum_kernel, gs_blocks(d.R), 256, 0, st, M.fbuf("bias_ro"), d.R, M.pp("/h1_to_readout.b"),
         M.pp("/h2_to_readout.b"), M.pp("/h3_to_readout.b"), M.pp("/att_to_readout.b"));
  if (!d.gmm) {
    CK(cudaMemcpyAsync(M.fbuf("bias_out"), M.pp("/readout_to_output.b"), d.D * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    const int DK = d.D * d.K;
    CK(cudaMemcpyAsync(M.fbuf("bias_out"), M.pp("/readout_to_output/fork_gmm_mu.b"), DK * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(M.fbuf("bias_out") + DK, M.pp("/readout_to_output/fork_gmm_sigma.b"), DK * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(M.fbuf("bias_out") + 2 * DK, M.pp("/readout_to_output/fork_gmm_coeff.b"), d.K * 4, cudaMemcpyDeviceToDevice, st));
  }
  // encoder: per-character Fork projections  proj[ch][dir][xi | xg]
  if (d.enc) {
    const int E = d.E;
    const char* dn[2] = {"forward", "backward"};
    for (int dir = 0; dir < 2; ++dir) {
      const std::string base = std::string("/encoder/encoder/") + dn[dir] + "/fork/fork_";
      float* dst = M.fbuf("enc_proj") + dir * 3 * E;
      sgemm(st, M.pp("/encoder/embed_label.W"), d.IN, 1, M.pp(base + "inputs.W"), E, 1, dst, 6 * E, d.NC, E, d.IN,
            M.pp(base + "inputs.b"), 0.0f);
      sgemm(st, M.pp("/encoder/embed_label.W"), d.IN, 1, M.pp(base + "gate_inputs.W"), 2 * E, 1, dst + E, 6 * E, d.NC,
            2 * E, d.IN, M.pp(base + "gate_inputs.b"), 0.0f);
    }
  }
  M.dirty = false;
}

// the two normalisation groups of a (., 3H) pre-activation row: cell block [0,H) and gate block [H,3H)
static NormParts gru_parts(const Dims& d) {
  NormParts P;
  P.off[0] = 0; P.n[0] = d.H; P.poff[0] = 0;
  P.off[1] = d.H; P.n[1] = 2 * d.H; P.poff[1] = d.Hp;
  P.count = 2;
  return P;
}
static NormParts row_parts(int n) {
  NormParts P;
  P.off[0] = 0; P.n[0] = n; P.poff[0] = 0;
  P.off[1] = 0; P.n[1] = 0; P.poff[1] = 0;
  P.count = 1;
  return P;
}

// per-call time-constant inputs of the recurrent layers and of the readout (speaker conditioning)
static void prep_base(parrot_model& M, const int32_t* d_speaker, cudaStream_t st) {
  const Dims& d = M.d;
  for (int l = 0; l < 3; ++l)
    LAUNCH(base_rows_kernel, gs_blocks((long long)d.B * 3 * d.H), 256, 0, st, M.fbuf("bias_l" + LN(l)),
           (const float*)nullptr, d.B, 3 * d.H, M.fbuf("base" + LN(l)));
  if (!d.spk) return;
  REQUIRE(d_speaker != nullptr, "use_speaker=True needs speaker indices (model.py:556-557)");
  float* emb = M.fbuf("spk_emb");
  LAUNCH(gather_rows_kernel, gs_blocks((long long)d.B * d.S), 256, 0, st, M.pp("/lookuptable.W"), d_speaker,
         (long long)d.B, d.S, emb);
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l), pre = "/speaker_to_h" + s + "/fork_rnn" + s;
    float* dst = d.ln ? M.fbuf("spk_pre" + s) : M.fbuf("base" + s);
    const float beta = d.ln ? 0.0f : 1.0f;
    sgemm(st, emb, d.S, 1, M.pp(pre + "_inputs.W"), d.H, 1, dst, 3 * d.H, d.B, d.H, d.S, M.pp(pre + "_inputs.b"), beta);
    sgemm(st, emb, d.S, 1, M.pp(pre + "_gates.W"), 2 * d.H, 1, dst + d.H, 3 * d.H, d.B, 2 * d.H, d.S,
          M.pp(pre + "_gates.b"), beta);
    if (d.ln)   // model.py:619-627: norm(speaker projection), cell and gate halves separately
      LAUNCH(rownorm_fwd_kernel, dim3(d.B, 2), 256, 0, st, dst, (long long)3 * d.H, M.fbuf("base" + s),
             (long long)3 * d.H, gru_parts(d), 1);
  }
  sgemm(st, emb, d.S, 1, M.pp("/speaker_to_readout.W"), d.R, 1, M.fbuf("spk_ro"), d.R, d.B, d.R, d.S,
        M.pp("/speaker_to_readout.b"), 0.0f);
  if (!d.gmm) {
    sgemm(st, emb, d.S, 1, M.pp("/speaker_to_output.W"), d.D, 1, M.fbuf("spk_out"), d.Dtot, d.B, d.D, d.S,
          M.pp("/speaker_to_output.b"), 0.0f);
  } else {
    const int DK = d.D * d.K;
    const char* fk[3] = {"gmm_mu", "gmm_sigma", "gmm_coeff"};
    const int off[3] = {0, DK, 2 * DK}, dim[3] = {DK, DK, d.K};
    for (int f = 0; f < 3; ++f) {
      const std::string n = std::string("/speaker_to_output/fork_") + fk[f];
      sgemm(st, emb, d.S, 1, M.pp(n + ".W"), dim[f], 1, M.fbuf("spk_out") + off[f], d.Dtot, d.B, dim[f], d.S,
            M.pp(n + ".b"), 0.0f);
    }
  }
}

// ------------------------------------------------------------------ encoder
static void enc_args(parrot_model& M, EncArgs& a) {
  const Dims& d = M.d;
  memset(&a, 0, sizeof a);
  const bool ax0 = M.cfg.encoder_time_axis == 0;
  a.L = ax0 ? d.B : d.U;
  a.N = ax0 ? d.U : d.B;
  a.E = d.E;
  const char* dn[2] = {"forward", "backward"};
  for (int dir = 0; dir < 2; ++dir) {
    const std::string s = std::to_string(dir);
    const std::string g = std::string("/encoder/encoder/") + dn[dir] + "/gatedrecurrent";
    a.xi[dir] = M.fbuf("enc_xi" + s); a.xg[dir] = M.fbuf("enc_xg" + s);
    a.Wg[dir] = M.pp(g + ".state_to_gates"); a.Ws[dir] = M.pp(g + ".state_to_state");
    a.s0[dir] = M.pp(g + ".initial_state");
    a.z[dir] = M.fbuf("enc_z" + s); a.r[dir] = M.fbuf("enc_r" + s); a.c[dir] = M.fbuf("enc_c" + s);
    a.sprev[dir] = M.fbuf("enc_sp" + s);
    if (!d.sampling) {
      a.dxi[dir] = M.fbuf("enc_dxi" + s); a.dxg[dir] = M.fbuf("enc_dxg" + s); a.ds0[dir] = M.fbuf("enc_ds0" + s);
    }
  }
  a.out = M.fbuf("enc_out");
  if (!d.sampling) a.dout = M.fbuf("enc_dout");
}

static void encoder_fwd(parrot_model& M, const int32_t* d_labels, const float* d_lmask, cudaStream_t st) {
  const Dims& d = M.d;
  const long long n = (long long)d.B * d.U * d.C;
  if (!d.enc) {
    // encoder_type None: "labels" already are the (B, U, input_dim) context features (model.py:235-236)
    LAUNCH(context_mask_kernel, gs_blocks(n), 256, 0, st, (const float*)d_labels, d_lmask, d.B, d.U, d.C, 0,
           M.fbuf("ctx"), 0);
    return;
  }
  EncArgs a;
  enc_args(M, a);
  LAUNCH(enc_gather_kernel, gs_blocks((long long)a.L * a.N * 6 * d.E), 256, 0, st, M.fbuf("enc_proj"), d_labels, a.L,
         a.N, d.U, M.cfg.encoder_time_axis, d.E, (float*)a.xi[0], (float*)a.xg[0], (float*)a.xi[1], (float*)a.xg[1],
         (int*)M.fbuf("enc_lab"));
  dim3 grid(cdiv(a.N, ENC_ROWS), 2);
  size_t smem = (size_t)ENC_ROWS * d.E * 4 * 4;
  const size_t wbytes = (size_t)3 * d.E * d.E * 4;
  a.w_in_smem = (smem + wbytes <= 220 * 1024) ? 1 : 0;
  if (a.w_in_smem) smem += wbytes;
  LAUNCH(encoder_fwd_kernel, grid, 256, smem, st, a);
  LAUNCH(context_mask_kernel, gs_blocks(n), 256, 0, st, M.fbuf("enc_out"), d_lmask, d.B, d.U, d.C,
         M.cfg.encoder_time_axis, M.fbuf("ctx"), 0);
}

static void encoder_bwd(parrot_model& M, const float* d_lmask, cudaStream_t st) {
  const Dims& d = M.d;
  if (!d.enc) return;
  const int E = d.E;
  EncArgs a;
  enc_args(M, a);
  g_sgemm_scratch = M.fbuf("sgemm_scratch");
  g_sgemm_scratch_floats = (long long)SGEMM_SPLIT * E * 2 * E;
  const long long n = (long long)d.B * d.U * d.C;
  // enc_dout <- dctx * mask (in the encoder's own layout)
  LAUNCH(context_mask_kernel, gs_blocks(n), 256, 0, st, M.fbuf("enc_dout"), d_lmask, d.B, d.U, d.C,
         M.cfg.encoder_time_axis, M.fbuf("dctx"), 1);
  dim3 grid(cdiv(a.N, ENC_ROWS), 2);
  size_t smem = (size_t)ENC_ROWS * E * 5 * 4;
  const size_t wbytes = (size_t)3 * E * E * 4;
  a.w_in_smem = (smem + wbytes <= 220 * 1024) ? 1 : 0;
  if (a.w_in_smem) smem += wbytes;
  LAUNCH(encoder_bwd_kernel, grid, 256, smem, st, a);
  const long long LNn = (long long)a.L * a.N;
  const char* dn[2] = {"forward", "backward"};
  CK(cudaMemsetAsync(M.fbuf("enc_dproj"), 0, (size_t)d.NC * 6 * E * 4, st));
  for (int dir = 0; dir < 2; ++dir) {
    const std::string s = std::to_string(dir);
    const std::string g = std::string("/encoder/encoder/") + dn[dir] + "/gatedrecurrent";
    colsum(st, a.ds0[dir], E, a.N, E, M.gp(g + ".initial_state"), 0);
    float* rs = M.fbuf("enc_rs" + s);
    LAUNCH(mul_kernel, gs_blocks(LNn * E), 256, 0, st, rs, a.sprev[dir], a.r[dir], LNn * E);
    sgemm(st, rs, 1, E, a.dxi[dir], E, 1, M.gp(g + ".state_to_state"), E, E, E, (int)LNn, nullptr, 0.0f);
    sgemm(st, a.sprev[dir], 1, E, a.dxg[dir], 2 * E, 1, M.gp(g + ".state_to_gates"), 2 * E, E, 2 * E, (int)LNn,
          nullptr, 0.0f);
    float* dp = M.fbuf("enc_dproj") + dir * 3 * E;
    LAUNCH(scatter_rows_kernel, dim3(cdiv(E, 32), d.NC), 1024, 0, st, a.dxi[dir], (long long)E,
           (const int*)M.fbuf("enc_lab"), (int)LNn, E, d.NC, dp, (long long)6 * E);
    LAUNCH(scatter_rows_kernel, dim3(cdiv(2 * E, 32), d.NC), 1024, 0, st, a.dxg[dir], (long long)2 * E,
           (const int*)M.fbuf("enc_lab"), (int)LNn, 2 * E, d.NC, dp + E, (long long)6 * E);
  }
  const float* Wemb = M.pp("/encoder/embed_label.W");
  float* dWemb = M.gp("/encoder/embed_label.W");
  for (int dir = 0; dir < 2; ++dir) {
    const std::string base = std::string("/encoder/encoder/") + dn[dir] + "/fork/fork_";
    const float* dp = M.fbuf("enc_dproj") + dir * 3 * E;
    const char* fk[2] = {"inputs", "gate_inputs"};
    const int off[2] = {0, E}, dim[2] = {E, 2 * E};
    for (int f = 0; f < 2; ++f) {
      // dW_f (IN x dim) = Wemb^T (IN x NC) . dproj (NC x dim)
      sgemm(st, Wemb, 1, d.IN, dp + off[f], 6 * E, 1, M.gp(base + fk[f] + ".W"), dim[f], d.IN, dim[f], d.NC, nullptr,
            0.0f);
      colsum(st, dp + off[f], 6 * E, d.NC, dim[f], M.gp(base + fk[f] + ".b"), 0);
      // dWemb (NC x IN) += dproj (NC x dim) . W_f^T (dim x IN)
      sgemm(st, dp + off[f], 6 * E, 1, M.pp(base + fk[f] + ".W"), 1, dim[f], dWemb, d.IN, d.NC, d.IN, dim[f], nullptr,
            1.0f);
    }
  }
  g_sgemm_scratch = nullptr;
  g_sgemm_scratch_floats = 0;
}

// ------------------------------------------------------------------ decoder scan
static void init_slots(parrot_model& M, bool use_initial, cudaStream_t st) {
  const Dims& d = M.d;
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l);
    const float* src = use_initial ? M.pp("/rnn" + s + ".initial_state") : M.fbuf("last_h" + s);
    const Plane& p = M.planes.at("h" + s);
    LAUNCH(state_to_slot_kernel, gs_blocks((long long)d.B * d.H), 256, 0, st, src, (long long)(use_initial ? 0 : d.H),
           d.B, d.H, d.Np, p.pitch, M.ctx.L[l].h, p.hi, p.lo);
  }
  const Plane& pw = M.planes.at("w");
  LAUNCH(state_to_slot_kernel, gs_blocks((long long)d.B * d.C), 256, 0, st,
         use_initial ? M.pp(".initial_w") : M.fbuf("last_w"), (long long)(use_initial ? 0 : d.C), d.B, d.C, d.Np,
         pw.pitch, M.fbuf("w"), pw.hi, pw.lo);
  if (use_initial) CK(cudaMemsetAsync(M.fbuf("kappa"), 0, (size_t)d.B * d.A * 4, st));
  else CK(cudaMemcpyAsync(M.fbuf("kappa"), M.fbuf("last_k"), (size_t)d.B * d.A * 4, cudaMemcpyDeviceToDevice, st));
}

static AttnFwdArgs attn_fwd_args(parrot_model& M, int t, bool sampling) {
  const Dims& d = M.d;
  AttnFwdArgs a;
  a.B = d.B; a.U = d.U; a.C = d.C; a.A = d.A; a.H = d.H; a.Np = d.Np;
  const Plane& pw = M.planes.at("w");
  a.Cp = pw.pitch;
  a.type = M.cfg.attention_type;
  a.eps = M.cfg.epsilon; a.align = M.cfg.attention_alignment;
  a.sharp = sampling ? M.cfg.sharpening_coeff : 1.0f;
  a.timing = sampling ? M.cfg.timing_coeff : 1.0f;
  a.h1 = M.ctx.L[0].h + (long long)(t + 1) * d.B * d.H;
  a.wT = M.fbuf("att_wT"); a.batt = M.fbuf("att_b"); a.ctx = M.fbuf("ctx");
  a.k_prev = M.fbuf("kappa") + (long long)t * d.B * d.A;
  a.k_out = M.fbuf("kappa") + (long long)(t + 1) * d.B * d.A;
  a.w_out = M.fbuf("w") + (long long)(t + 1) * d.B * d.C;
  a.w_hi = pw.hi + (long long)(t + 1) * d.Np * pw.pitch;
  a.w_lo = pw.lo + (long long)(t + 1) * d.Np * pw.pitch;
  a.phi_out = M.fbuf("phi") + (long long)t * d.B * d.U;
  a.ab_out = M.fbuf("ab") + (long long)t * d.B * 2 * d.A;
  a.e_out = M.fbuf("att_e") + (long long)t * d.B * 3 * d.A;
  a.hat = M.fbuf("att_hat");
  a.hat_part = M.fbuf("att_hat_part");
  return a;
}
static size_t att_proj_smem(const Dims& d) { return (size_t)(d.B + 3 * d.A) * (ATT_KS + 4) * 4; }
static size_t att_window_smem(const Dims& d, int nparts) {
  return (size_t)(2 * rup(3 * d.A, 4) + rup(d.U, 4) + 8 * (d.C / nparts)) * 4;
}
static void attention_step(parrot_model& M, int t, bool sampling, cudaStream_t st) {
  const Dims& d = M.d;
  AttnFwdArgs a = attn_fwd_args(M, t, sampling);
  cudaEvent_t pe = M.prof_begin("attn_fwd", st);
  const int nparts = attention_nparts(d.B, d.C, 148);
  if (3 * d.A <= 32 && d.H % 4 == 0) {
    // one launch: every (batch row, context-column part) CTA computes the row's projection itself
    LAUNCH(attention_step_kernel, d.B * nparts, 256, att_window_smem(d, nparts) + 8 * 32 * 4, st, a, nparts);
  } else {
    // stage 1: K-sliced partial projections h1 . Watt^T ; stage 2: window + context, nparts CTAs per batch row
    LAUNCH(attention_proj_slice_kernel, M.att_slices, 256, att_proj_smem(d), st, a);
    LAUNCH(attention_window_kernel, d.B * nparts, 256, att_window_smem(d, nparts), st, a, nparts, M.att_slices);
  }
  parrot_model::prof_end(pe, st);
}

static int prefetch_enabled() {
  const char* e = getenv("PARROT_NO_PREFETCH");
  return (e && e[0] && e[0] != '0') ? 0 : 1;
}
static bool use_persistent(parrot_model& M) {
  if (M.d.ln) return false;   // the normalisations sit between the layers: one launch per phase
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PARROT_NO_PERSISTENT");
    env = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  const Dims& d = M.d;
  if (!env || !M.persistent_ok || M.cfg.gemm_impl == 1 || M.profiling >= 2 || M.sm_count < 148) return false;
  // the scan finishes work on quads of features (128-bit stash / plane accesses): odd sizes take the per-phase path
  if ((d.H & 3) || (d.Hp & 3) || (d.C & 3)) return false;
  const size_t att_f = std::max(att_proj_smem(d), att_window_smem(d, attention_nparts(d.B, d.C, M.grp_f[0])));
  const size_t att_b = ((size_t)d.C + d.U + 3 * d.A * 16 + 10 * d.A) * 4;
  if (std::max(att_f, att_b) > (size_t)ATT_SMEM_BYTES) return false;
  for (const char* nm : {"fwdA", "fwdB", "bwd1", "bwd2"}) {
    auto it = M.tables.find(nm);
    if (it != M.tables.end() && it->second.count > 148) return false;
  }
  for (int l = 0; l < 3; ++l)
      for (const char* nm : {"gA", "gB", "hA", "hB"}) {
        auto it = M.tables.find(std::string(nm) + LN(l));
        if (it != M.tables.end() && it->second.count > (nm[0] == 'g' ? M.grp_f[l] : M.grp_b[l])) return false;
      }
  return true;
}

static EngineParams table_params(parrot_model& M, const std::string& name, int reverse, int tl_slot = -1) {
  const Table& t = M.tables.at(name);
  EngineParams P;
  P.jobs = M.d_jobs + t.off; P.njobs = t.count; P.maps = M.d_maps; P.raws = M.d_raws; P.ctx = M.d_ctx;
  P.tick = 0; P.T = M.d.T; P.n_cols = t.n_cols; P.reverse = reverse;
  P.chunk_samples = t.chunk_samples;
  P.split_scratch = M.d_split_scratch; P.split_count = M.d_split_count;
  if (t.uniq_floats >= 0) {
    P.split_scratch += std::max<size_t>(M.max_split_floats, 1) + t.uniq_floats;
    P.split_count += std::max(M.max_groups, 1) + t.uniq_groups;
  }
  // debug: intra-phase milestones of persistent tick M.tl_tick, one [148][16] block per phase
  P.timeline = (M.timeline && tl_slot >= 0) ? M.timeline + (size_t)tl_slot * 148 * 16 : nullptr;
  P.tl_tick = M.tl_tick;
  P.coop_epilogue = 1;
  P.cta0 = 0; P.ncta = 148;
  {
    const char* e = getenv("PARROT_DEBUG_FLAGS");
    P.debug_flags = e ? atoi(e) : 0;
  }
  return P;
}

static AttnFwdArgs attn_fwd_args(parrot_model& M, int t, bool sampling);
static AttnBwdArgs attn_bwd_args(parrot_model& M, int t);

// ---- grouped persistent scans (kernels.cuh): three layer groups with their own barrier domains
static GroupSched group_sched(parrot_model& M, const char* a, const char* b, const char* c, int l, int reverse,
                              const int* grp) {
  GroupSched G;
  int cta0 = 0;
  for (int i = 0; i < l; ++i) cta0 += grp[i];
  G.cta0 = cta0; G.ncta = grp[l];
  const std::string names[3] = {std::string(a) + LN(l), std::string(b) + LN(l), std::string(c) + LN(l)};
  for (int k = 0; k < 3; ++k) {
    if (M.tables.count(names[k])) {
      G.ph[k] = table_params(M, names[k], reverse, k < 2 ? k : -1);
    } else {
      G.ph[k] = table_params(M, names[0], reverse);
      G.ph[k].njobs = 0;
    }
    G.ph[k].cta0 = G.cta0; G.ph[k].ncta = G.ncta;
    REQUIRE(k == 2 || G.ph[k].njobs <= G.ncta, "grouped scan: more split jobs than CTAs in the group");
  }
  return G;
}
static bool scan_fwd_grouped_launch(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  ScanFwdGParams S;
  for (int l = 0; l < 3; ++l) S.g[l] = group_sched(M, "gA", "gB", "gC", l, 0, M.grp_f);
  S.Tc = M.Tc; S.T = d.T;
  S.att_parts = attention_nparts(d.B, d.C, M.grp_f[0]); S.att_slices = M.att_slices;
  S.att = attn_fwd_args(M, 0, false);
  S.s_h1 = (long long)d.B * d.H; S.s_k = (long long)d.B * d.A; S.s_w = (long long)d.B * d.C;
  S.s_wp = (long long)d.Np * M.planes.at("w").pitch; S.s_phi = (long long)d.B * d.U;
  S.s_ab = (long long)d.B * 2 * d.A; S.s_e = (long long)d.B * 3 * d.A;
  S.bars = M.d_gridbar;
  S.stamps = M.stamps; S.stamp_bars = M.stamp_bars;
  S.prefetch = prefetch_enabled();
  CK(cudaMemsetAsync(M.d_gridbar, 0, 1024, st));
  void* args[] = {&S};
  g_ctx = "scan_fwd_grouped";
  const int grid = M.grp_f[0] + M.grp_f[1] + M.grp_f[2];
  cudaError_t le = cudaLaunchCooperativeKernel((void*)scan_fwd_grouped, dim3(grid), dim3(ENGINE_THREADS), args,
                                               SMEM_BYTES + 1024 + ATT_SMEM_BYTES, st);
  if (le != cudaSuccess) {
    cudaGetLastError();
    M.persistent_ok = false;
    return false;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (debug_sync()) CK(cudaStreamSynchronize(st));
  return true;
}
static bool scan_bwd_grouped_launch(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  ScanBwdGParams S;
  for (int l = 0; l < 3; ++l) S.g[l] = group_sched(M, "hA", "hB", "hC", l, 1, M.grp_b);
  S.Tc = M.Tc; S.T = d.T;
  S.att = attn_bwd_args(M, 0);
  S.s_dw = (long long)d.B * d.C; S.s_ab = (long long)d.B * 2 * d.A; S.s_e = (long long)d.B * 3 * d.A;
  S.s_k = (long long)d.B * d.A; S.s_dh1 = (long long)d.B * d.H; S.s_datt = (long long)d.B * 3 * d.A;
  S.s_dattp = (long long)d.Np * M.planes.at("datt").pitch;
  S.ctx = M.d_ctx; S.bars = M.d_gridbar;
  S.fused_pre = (d.H % 4 == 0 && d.Hp % 4 == 0) ? 1 : 0;   // (same condition as QF_FUSED_PRE in the hB tables)
  if (const char* e = getenv("PARROT_NO_FUSED_PRE")) { if (e[0] && e[0] != '0') S.fused_pre = 0; }
  S.stamps = M.stamps_bwd; S.stamp_bars = M.stamps_bwd ? M.stamp_bars : 0;
  S.tl_buf = M.timeline; S.tl_tick = M.tl_tick;
  S.prefetch = prefetch_enabled();
  CK(cudaMemsetAsync(M.d_gridbar, 0, 1024, st));
  void* args[] = {&S};
  g_ctx = "scan_bwd_grouped";
  const int grid = M.grp_b[0] + M.grp_b[1] + M.grp_b[2];
  cudaError_t le = cudaLaunchCooperativeKernel((void*)scan_bwd_grouped, dim3(grid), dim3(ENGINE_THREADS), args,
                                               SMEM_BYTES + 1024 + ATT_SMEM_BYTES, st);
  if (le != cudaSuccess) {
    cudaGetLastError();
    M.persistent_ok = false;
    return false;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (debug_sync()) CK(cudaStreamSynchronize(st));
  return true;
}

// layer_norm=True forward scan (model.py:571-603, 692-722 with _apply_norm active): every Fork output except
// inp_to_h* is normalised before it is summed, so the layers cannot share one accumulator.  preT_l[t] collects
// bias + norm(speaker) + norm(feedback) (+ norm(q12/q13/q23) as the lower layers finish step t).
static void scan_fwd_ln(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  const long long row = (long long)3 * d.H, step = (long long)d.B * row;
  const NormParts gp = gru_parts(d);
  run_table(M, "ln_fb", 0, 1, 0, st);
  for (int l = 0; l < 3; ++l) {
    LAUNCH(bcast_rows_kernel, gs_blocks((long long)d.T * step), 256, 0, st, M.fbuf("preT" + LN(l)),
           M.fbuf("base" + LN(l)), d.T, step);
    if ((l == 0 && d.weak) || (l > 0 && d.full))
      LAUNCH(rownorm_fwd_kernel, dim3(d.T * d.B, 2), 256, 0, st, M.fbuf("qfb" + LN(l)), row,
             M.fbuf("preT" + LN(l)), row, gp, 1);
  }
  auto norm_into = [&](const char* q, const char* pre, int t) {
    LAUNCH(rownorm_fwd_kernel, dim3(d.B, 2), 256, 0, st, M.fbuf(q) + t * step, row, M.fbuf(pre) + t * step, row, gp,
           1);
  };
  for (int t = 0; t < d.T; ++t) {
    run_table(M, "lnA1", t, d.T, 0, st);
    run_table(M, "lnB1", t, d.T, 0, st);
    attention_step(M, t, false, st);
    run_table(M, "lnQ1", t, d.T, 0, st);
    norm_into("q12", "preT2", t);
    norm_into("q13", "preT3", t);
    run_table(M, "lnA2", t, d.T, 0, st);
    run_table(M, "lnB2", t, d.T, 0, st);
    run_table(M, "lnQ2", t, d.T, 0, st);
    norm_into("q23", "preT3", t);
    run_table(M, "lnA3", t, d.T, 0, st);
    run_table(M, "lnB3", t, d.T, 0, st);
  }
}

static void scan_fwd(parrot_model& M, const float* d_features, const float* d_noise, float level, float start_flag,
                     cudaStream_t st) {
  const Dims& d = M.d;
  if (d.weak) {
    const Plane& px = M.planes.at("xin");
    const bool noisy = d_noise != nullptr && level != 0.0f;
    LAUNCH(frames_to_planes_kernel, gs_blocks((long long)d.T * d.B * d.D), 256, 0, st, d_features,
           noisy ? d_noise : (const float*)nullptr, level, d.T, d.B, d.D, d.Np, px.pitch, px.hi, px.lo,
           (float*)nullptr);
  }
  init_slots(M, start_flag != 0.0f, st);
  M.last_start_flag = start_flag;
  if (d.ln) { scan_fwd_ln(M, st); return; }
  if (d.weak) run_table(M, "hoist1", 0, 1, 0, st);   // pre1 = x_{t-1} . out_to_h1 for all frames
  if (use_persistent(M) && scan_fwd_grouped_launch(M, st)) return;
  // chunk-lagged layer wavefront, one launch per phase: tick tau runs layer 1 at step tau, layer 2 at tau - Tc,
  // layer 3 at tau - 2 Tc; every Tc ticks the hoisted products of the chunk just finished
  const int Tc = M.Tc;
  for (int tick = 0; tick < d.T + 2 * Tc; ++tick) {
    run_table(M, "fwdA", tick, d.T, 0, st);
    run_table(M, "fwdB", tick, d.T, 0, st);
    if (tick < d.T) attention_step(M, tick, false, st);
    if ((tick + 1) % Tc == 0) run_table(M, "chunkF", (tick + 1) / Tc - 1, d.T, 0, st);
  }
}

// ------------------------------------------------------------------ readout / emitter
static EmitArgs emit_args(parrot_model& M, const float* d_features, const float* d_fmask, int unnormalised) {
  const Dims& d = M.d;
  EmitArgs e;
  memset(&e, 0, sizeof e);
  e.N = d.T * d.B; e.D = d.D; e.k = d.K; e.which = d.gmm ? 1 : 0; e.Dtot = d.Dtot; e.eps = M.cfg.epsilon;
  e.pred = M.fbuf("pred");
  e.target = d_features + (long long)d.B * d.D;   // features[1:]  (model.py:559)
  e.mask = d_fmask + d.B;                         // features_mask[1:] (model.py:560)
  e.cost_tb = M.fbuf("cost_tb");
  if (!d.sampling) {
    e.dpred = M.fbuf("dpred");
    const Plane& p = M.planes.at("dpred");
    e.dpred_hi = p.hi; e.dpred_lo = p.lo; e.B = d.B; e.Np = d.Np; e.Dp = p.pitch;
    e.DK = d.D * d.K; e.poff1 = rup(d.D * d.K, 64); e.poff2 = 2 * rup(d.D * d.K, 64);
    e.scale = M.fbuf("cost") + (unnormalised ? 4 : 3);
  }
  return e;
}

static void readout_emit_fwd(parrot_model& M, const float* d_features, const float* d_fmask, float* d_cost,
                             cudaStream_t st) {
  const Dims& d = M.d;
  if (d.spk) {
    LAUNCH(bcast_rows_kernel, gs_blocks((long long)d.T * d.B * d.R), 256, 0, st, M.fbuf("ro"), M.fbuf("spk_ro"), d.T,
           (long long)d.B * d.R);
    LAUNCH(bcast_rows_kernel, gs_blocks((long long)d.T * d.B * d.Dtot), 256, 0, st, M.fbuf("pred"), M.fbuf("spk_out"),
           d.T, (long long)d.B * d.Dtot);
  }
  if (d.ln) {
    // model.py:743-753: norm(h_l readout) summed, then speaker and attention readouts un-normalised
    run_table(M, "ln_ro", 0, 1, 0, st);
    for (int l = 0; l < 3; ++l)
      LAUNCH(rownorm_fwd_kernel, dim3(d.T * d.B, 1), 256, 0, st, M.fbuf("ropre" + LN(l)), (long long)d.R,
             M.fbuf("ro"), (long long)d.R, row_parts(d.R), (l > 0 || d.spk) ? 1 : 0);
    run_table(M, "ln_ro_att", 0, 1, 0, st);
  } else {
    run_table(M, "readout", 0, 1, 0, st);
  }
  run_table(M, "output", 0, 1, 0, st);
  EmitArgs e = emit_args(M, d_features, d_fmask, 0);
  LAUNCH(emit_cost_kernel, cdiv((long long)e.N * 32, 256), 256, 0, st, e);
  LAUNCH(masked_mean_kernel, 1, 1024, 0, st, M.fbuf("cost_tb"), e.mask, (long long)e.N, M.fbuf("cost"));
  if (d_cost) CK(cudaMemcpyAsync(d_cost, M.fbuf("cost"), 16, cudaMemcpyDeviceToDevice, st));
}

static void apply_updates(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  for (int l = 0; l < 3; ++l)
    CK(cudaMemcpyAsync(M.fbuf("last_h" + LN(l)), M.ctx.L[l].h + (long long)d.T * d.B * d.H, (size_t)d.B * d.H * 4,
                       cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(M.fbuf("last_k"), M.fbuf("kappa") + (long long)d.T * d.B * d.A, (size_t)d.B * d.A * 4,
                     cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(M.fbuf("last_w"), M.fbuf("w") + (long long)d.T * d.B * d.C, (size_t)d.B * d.C * 4,
                     cudaMemcpyDeviceToDevice, st));
}

// ------------------------------------------------------------------ backward
struct FwdInputs {
  const float* features = nullptr;
  const float* fmask = nullptr;
  const int32_t* labels = nullptr;
  const float* lmask = nullptr;
  const int32_t* speaker = nullptr;
};
static std::map<parrot_model*, FwdInputs> g_inputs;

static void readout_emit_bwd(parrot_model& M, int unnormalised, cudaStream_t st) {
  const Dims& d = M.d;
  const FwdInputs& in = g_inputs[&M];
  EmitArgs e = emit_args(M, in.features, in.fmask, unnormalised);
  LAUNCH(emit_grad_kernel, cdiv((long long)e.N * 32, 256), 256, 0, st, e);
  run_table(M, "dread", 0, 1, 0, st);
  if (d.ln)
    for (int l = 0; l < 3; ++l) {
      const Plane& pl = M.planes.at("dro" + LN(l));
      LAUNCH(rownorm_bwd_kernel, dim3(d.T * d.B, 1), 256, 0, st, M.fbuf("dread"), (long long)d.R,
             M.fbuf("ropre" + LN(l)), (long long)d.R, M.fbuf("dro" + LN(l)), (long long)d.R, pl.hi, pl.lo,
             (long long)pl.pitch, d.B, d.Np, row_parts(d.R));
    }
  for (int l = 0; l < 3; ++l) CK(cudaMemsetAsync(M.ctx.L[l].dh, 0, (size_t)d.B * d.H * 4, st));
  CK(cudaMemsetAsync(M.ctx.dw, 0, (size_t)d.B * d.C * 4, st));
  run_table(M, "dh_readout", 0, 1, 0, st);
}

static AttnBwdArgs attn_bwd_args(parrot_model& M, int t) {
  const Dims& d = M.d;
  AttnBwdArgs a;
  a.B = d.B; a.U = d.U; a.C = d.C; a.A = d.A; a.H = d.H; a.Np = d.Np; a.Ap = d.Ap;
  a.type = M.cfg.attention_type; a.eps = M.cfg.epsilon; a.align = M.cfg.attention_alignment;
  a.dw = M.ctx.dw + (long long)(t + 1) * d.B * d.C;
  a.ctx = M.fbuf("ctx");
  a.ab = M.fbuf("ab") + (long long)t * d.B * 2 * d.A;
  a.e = M.fbuf("att_e") + (long long)t * d.B * 3 * d.A;
  a.kappa = M.fbuf("kappa") + (long long)(t + 1) * d.B * d.A;
  a.dk_carry = M.fbuf("dk_carry");
  a.watt = M.fbuf("att_wT");
  a.dh1 = M.ctx.L[0].dh + (long long)(t + 1) * d.B * d.H;
  a.datt = M.fbuf("datt") + (long long)t * d.B * 3 * d.A;
  const Plane& p = M.planes.at("datt");
  a.datt_hi = p.hi + (long long)t * d.Np * p.pitch;
  a.datt_lo = p.lo + (long long)t * d.Np * p.pitch;
  a.dbg = nullptr;
  return a;
}
static void attention_bwd_step(parrot_model& M, int t, cudaStream_t st) {
  const Dims& d = M.d;
  AttnBwdArgs a = attn_bwd_args(M, t);
  const size_t smem = (size_t)(d.C + d.U + 3 * d.A * 16 + 10 * d.A) * 4;
  cudaEvent_t pe = M.prof_begin("attn_bwd", st);
  LAUNCH(attention_bwd_kernel, d.B, 256, smem, st, a);
  parrot_model::prof_end(pe, st);
}

// layer_norm=True reverse sweep: per step, layer 3 -> layer 2 -> attention -> layer 1; the gradients that cross a
// normalised Fork (q23, q13, q12) go through rownorm_bwd before the transposed products.
static void scan_bwd_ln(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  const int blocks = std::min(gs_blocks((long long)d.B * d.H), 148);
  const long long row = (long long)3 * d.H, step = (long long)d.B * row;
  const NormParts gp = gru_parts(d);
  auto norm_bwd = [&](int layer, const char* q, const char* dq, int t) {
    const Plane& pl = M.planes.at(dq);
    const long long pstep = (long long)d.Np * pl.pitch;
    LAUNCH(rownorm_bwd_kernel, dim3(d.B, 2), 256, 0, st, M.ctx.L[layer].da + t * step, row, M.fbuf(q) + t * step, row,
           M.fbuf(dq) + t * step, row, pl.hi + t * pstep, pl.lo + t * pstep, (long long)pl.pitch, d.B, d.Np, gp);
  };
  for (int t = d.T - 1; t >= 0; --t) {
    for (int l = 2; l >= 0; --l) {
      if (l == 0) attention_bwd_step(M, t, st);
      PreArgs pa;
      pa.n = 1; pa.layer[0] = l; pa.t[0] = t;
      cudaEvent_t pe = M.prof_begin("gru_bwd_pre", st);
      LAUNCH(gru_bwd_pre_kernel, dim3(blocks, 1), 256, 0, st, M.d_ctx, pa);
      parrot_model::prof_end(pe, st);
      run_table(M, "ln_bwd1_" + LN(l), t, d.T, 0, st);
      if (l == 2) { norm_bwd(2, "q23", "dq23", t); norm_bwd(2, "q13", "dq13", t); }
      if (l == 1) norm_bwd(1, "q12", "dq12", t);
      run_table(M, "ln_bwd2_" + LN(l), t, d.T, 0, st);
    }
  }
}

static void scan_bwd(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  CK(cudaMemsetAsync(M.fbuf("dk_carry"), 0, (size_t)d.B * d.A * 4, st));
  if (d.ln) { scan_bwd_ln(M, st); return; }
  if (use_persistent(M) && scan_bwd_grouped_launch(M, st)) return;
  const int blocks = std::min(gs_blocks((long long)d.B * d.H), 148);
  // reverse chunk-lagged wavefront: tick tau -> layer 3 at s = T-1-tau, layer 2 at s + Tc, attention + layer 1 at
  // s + 2 Tc; every Tc ticks the hoisted dgrads of the ranges just finished
  const int Tc = M.Tc;
  for (int tick = 0; tick < d.T + 2 * Tc; ++tick) {
    const int s = d.T - 1 - tick;
    if (s + 2 * Tc >= 0 && s + 2 * Tc < d.T) attention_bwd_step(M, s + 2 * Tc, st);
    PreArgs pa;
    pa.n = 0;
    for (int l = 2; l >= 0; --l) {
      const int t = s + (2 - l) * Tc;
      if (t >= 0 && t < d.T) { pa.layer[pa.n] = l; pa.t[pa.n] = t; ++pa.n; }
    }
    if (pa.n > 0) {
      cudaEvent_t pe = M.prof_begin("gru_bwd_pre", st);
      LAUNCH(gru_bwd_pre_kernel, dim3(blocks, pa.n), 256, 0, st, M.d_ctx, pa);
      parrot_model::prof_end(pe, st);
    }
    run_table(M, "bwd1", tick, d.T, 1, st);
    run_table(M, "bwd2", tick, d.T, 1, st);
    if ((tick + 1) % Tc == 0) run_table(M, "chunkB", (tick + 1) / Tc - 1, d.T, 1, st);
  }
}

static void add_to(parrot_model& M, cudaStream_t st, const std::string& pname, const float* src, int n) {
  LAUNCH(add_vec_kernel, gs_blocks(n), 256, 0, st, M.gp(pname), src, (long long)n);
}

static void weight_grads(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  g_colsum_scratch = M.fbuf("colsum_scratch");
  const int T = d.T, B = d.B, H = d.H;
  // operand transposes [samples][features] -> [features][samples]
  auto tp = [&](const std::string& nm, int feat) { transpose_planes(st, M.planes.at(nm), feat, M.planes.at("T." + nm)); };
  for (int l = 0; l < 3; ++l) {
    tp("h" + LN(l), d.H); tp("rh" + LN(l), d.H); tp("da" + LN(l), 3 * d.Hp);
  }
  tp("w", d.C);
  if (d.weak) tp("xin", d.D);
  tp("ro", d.R); tp("dread", d.R); tp("dpred", d.Dtp); tp("datt", 3 * d.A);
  if (d.ln) {
    const NormParts gp = gru_parts(d);
    for (int l = 0; l < 3; ++l) {
      tp("dro" + LN(l), d.R);
      if ((l == 0 && d.weak) || (l > 0 && d.full)) {
        // gradient behind the normalised feedback Fork (model.py:571-603)
        const Plane& pl = M.planes.at("dfb" + LN(l));
        LAUNCH(rownorm_bwd_kernel, dim3(T * B, 2), 256, 0, st, M.ctx.L[l].da, (long long)3 * H,
               M.fbuf("qfb" + LN(l)), (long long)3 * H, M.fbuf("dfb" + LN(l)), (long long)3 * H, pl.hi, pl.lo,
               (long long)pl.pitch, B, d.Np, gp);
        tp("dfb" + LN(l), 3 * d.Hp);
      }
    }
    tp("dq12", 3 * d.Hp); tp("dq13", 3 * d.Hp); tp("dq23", 3 * d.Hp);
  }
  { cudaEvent_t p2 = M.prof_begin("tail_transposes_done", st); parrot_model::prof_end(p2, st); }
  run_table(M, "wgrad", 0, 1, 0, st);
  cudaEvent_t pb = M.prof_begin("tail_bias_grads", st);
  // bias gradients: column sums of the pre-activation gradients
  float* scratch = M.fbuf("bias_scratch");
  auto fork_bias = [&](const float* src, const std::string& fork, const std::string& s) {
    add_to(M, st, fork + "/fork_rnn" + s + "_inputs.b", src, H);
    add_to(M, st, fork + "/fork_rnn" + s + "_gates.b", src + H, 2 * H);
  };
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l);
    colsum(st, M.ctx.L[l].da, 3 * H, (long long)T * B, 3 * H, scratch, 0);
    fork_bias(scratch, "/inp_to_h" + s, s);
    if (d.ln) continue;
    if (l >= 1) fork_bias(scratch, "/h1_to_h" + s, s);
    if (l == 2) fork_bias(scratch, "/h2_to_h3", s);
    if ((l == 0 && d.weak) || (l > 0 && d.full)) fork_bias(scratch, "/out_to_h" + s, s);
    if (d.spk) fork_bias(scratch, "/speaker_to_h" + s, s);
  }
  if (d.ln) {
    // biases that sit inside a normalisation receive the column sums of the gradient behind it
    auto through = [&](const char* dq, const std::string& fork, const std::string& s) {
      colsum(st, M.fbuf(dq), 3 * H, (long long)T * B, 3 * H, scratch, 0);
      fork_bias(scratch, fork, s);
    };
    through("dq12", "/h1_to_h2", "2"); through("dq13", "/h1_to_h3", "3"); through("dq23", "/h2_to_h3", "3");
    for (int l = 0; l < 3; ++l)
      if ((l == 0 && d.weak) || (l > 0 && d.full))
        through(("dfb" + LN(l)).c_str(), "/out_to_h" + LN(l), LN(l));
    for (int l = 0; l < 3; ++l) {
      colsum(st, M.fbuf("dro" + LN(l)), d.R, (long long)T * B, d.R, scratch, 0);
      add_to(M, st, "/h" + LN(l) + "_to_readout.b", scratch, d.R);
    }
  }
  colsum(st, M.fbuf("dread"), d.R, (long long)T * B, d.R, scratch, 0);
  if (!d.ln) for (int l = 0; l < 3; ++l) add_to(M, st, "/h" + LN(l) + "_to_readout.b", scratch, d.R);
  add_to(M, st, "/att_to_readout.b", scratch, d.R);
  if (d.spk) add_to(M, st, "/speaker_to_readout.b", scratch, d.R);
  float* sc2 = scratch + std::max(3 * H, d.R);  // dpred sums
  colsum(st, M.fbuf("dpred"), d.Dtot, (long long)T * B, d.Dtot, sc2, 0);
  if (!d.gmm) {
    add_to(M, st, "/readout_to_output.b", sc2, d.D);
    if (d.spk) add_to(M, st, "/speaker_to_output.b", sc2, d.D);
  } else {
    const int DK = d.D * d.K;
    const char* fk[3] = {"gmm_mu", "gmm_sigma", "gmm_coeff"};
    const int off[3] = {0, DK, 2 * DK}, dim[3] = {DK, DK, d.K};
    for (int f = 0; f < 3; ++f) {
      add_to(M, st, std::string("/readout_to_output/fork_") + fk[f] + ".b", sc2 + off[f], dim[f]);
      if (d.spk) add_to(M, st, std::string("/speaker_to_output/fork_") + fk[f] + ".b", sc2 + off[f], dim[f]);
    }
  }
  colsum(st, M.fbuf("datt"), 3 * d.A, (long long)T * B, 3 * d.A, scratch, 0);
  add_to(M, st, "/h1_to_att/fork_alpha.b", scratch, d.A);
  add_to(M, st, "/h1_to_att/fork_beta.b", scratch + d.A, d.A);
  add_to(M, st, "/h1_to_att/fork_kappa.b", scratch + 2 * d.A, d.A);
  parrot_model::prof_end(pb, st);
}

static void speaker_grads(parrot_model& M, cudaStream_t st) {
  const Dims& d = M.d;
  if (!d.spk) return;
  const FwdInputs& in = g_inputs[&M];
  const int T = d.T, B = d.B, H = d.H, S = d.S;
  const float* emb = M.fbuf("spk_emb");
  float* demb = M.fbuf("spk_demb");
  CK(cudaMemsetAsync(demb, 0, (size_t)B * S * 4, st));
  auto through = [&](const float* dproj, long long ld, int N, const std::string& wname) {
    // dW (S x N) = emb^T . dproj ; demb += dproj . W^T
    sgemm(st, emb, 1, S, dproj, ld, 1, M.gp(wname + ".W"), N, S, N, B, nullptr, 0.0f);
    sgemm(st, dproj, ld, 1, M.pp(wname + ".W"), 1, N, demb, S, B, S, N, nullptr, 1.0f);
  };
  float* dp = M.fbuf("spk_dproj");
  for (int l = 0; l < 3; ++l) {
    const std::string s = LN(l), pre = "/speaker_to_h" + s + "/fork_rnn" + s;
    LAUNCH(timesum_kernel, gs_blocks((long long)B * 3 * H), 256, 0, st, M.ctx.L[l].da, T, (long long)B * 3 * H, dp);
    if (d.ln) {
      // norm(speaker projection) is constant over time: sum the gradient first, then one norm backward
      float* tmp = M.fbuf("spk_tmp");
      LAUNCH(rownorm_bwd_kernel, dim3(B, 2), 256, 0, st, dp, (long long)3 * H, M.fbuf("spk_pre" + s), (long long)3 * H,
             tmp, (long long)3 * H, (bf16*)nullptr, (bf16*)nullptr, 0LL, B, d.Np, gru_parts(d));
      CK(cudaMemcpyAsync(dp, tmp, (size_t)B * 3 * H * 4, cudaMemcpyDeviceToDevice, st));
      colsum(st, dp, 3 * H, B, 3 * H, M.fbuf("bias_scratch"), 0);
      add_to(M, st, pre + "_inputs.b", M.fbuf("bias_scratch"), H);
      add_to(M, st, pre + "_gates.b", M.fbuf("bias_scratch") + H, 2 * H);
    }
    through(dp, 3 * H, H, pre + "_inputs");
    through(dp + H, 3 * H, 2 * H, pre + "_gates");
  }
  float* dro = M.fbuf("spk_dro");
  LAUNCH(timesum_kernel, gs_blocks((long long)B * d.R), 256, 0, st, M.fbuf("dread"), T, (long long)B * d.R, dro);
  through(dro, d.R, d.R, "/speaker_to_readout");
  float* dout = M.fbuf("spk_dout");
  LAUNCH(timesum_kernel, gs_blocks((long long)B * d.Dtot), 256, 0, st, M.fbuf("dpred"), T, (long long)B * d.Dtot, dout);
  if (!d.gmm) through(dout, d.Dtot, d.D, "/speaker_to_output");
  else {
    const int DK = d.D * d.K;
    through(dout, d.Dtot, DK, "/speaker_to_output/fork_gmm_mu");
    through(dout + DK, d.Dtot, DK, "/speaker_to_output/fork_gmm_sigma");
    through(dout + 2 * DK, d.Dtot, d.K, "/speaker_to_output/fork_gmm_coeff");
  }
  LAUNCH(scatter_rows_kernel, dim3(cdiv(S, 32), M.cfg.num_speakers), 1024, 0, st, demb, (long long)S, in.speaker,
         B, S, M.cfg.num_speakers, M.gp("/lookuptable.W"), (long long)S);
}

static void backward(parrot_model& M, int unnormalised, cudaStream_t st) {
  const Dims& d = M.d;
  REQUIRE(M.have_fwd, "parrot_backward called before parrot_compute_cost");
  REQUIRE(!d.sampling, "sampling handles have no backward");
  const FwdInputs& in = g_inputs[&M];
  CK(cudaMemsetAsync(M.grads, 0, (size_t)(M.P.total + 1) * 4, st));
  CK(cudaMemcpyAsync(M.grads + M.P.total, M.fbuf("cost") + 2, 4, cudaMemcpyDeviceToDevice, st));
  cudaEvent_t pe = M.prof_begin("sec_readout_emit_bwd", st);
  readout_emit_bwd(M, unnormalised, st);
  parrot_model::prof_end(pe, st);
  pe = M.prof_begin("sec_scan_bwd", st);
  scan_bwd(M, st);
  parrot_model::prof_end(pe, st);
  pe = M.prof_begin("sec_grads_tail", st);
  if (M.last_start_flag != 0.0f) {
    for (int l = 0; l < 3; ++l) colsum(st, M.ctx.L[l].dh, d.H, d.B, d.H, M.gp("/rnn" + LN(l) + ".initial_state"), 0);
    colsum(st, M.ctx.dw, d.C, d.B, d.C, M.gp(".initial_w"), 0);
  }
  // dctx = sum_t phi_t (x) dw_t
  {
    cudaEvent_t p2 = M.prof_begin("tail_dctx", st);
    dim3 grid(cdiv(d.C, 32), cdiv(d.U, 32), d.B);
    LAUNCH(dctx_kernel, grid, 256, 0, st, M.fbuf("phi"), M.ctx.dw + (long long)d.B * d.C, d.T, d.B, d.U, d.C,
           M.fbuf("dctx"));
    parrot_model::prof_end(p2, st);
  }
  {
    cudaEvent_t p2 = M.prof_begin("tail_encoder_bwd", st);
    encoder_bwd(M, in.lmask, st);
    parrot_model::prof_end(p2, st);
    p2 = M.prof_begin("tail_weight_grads", st);
    weight_grads(M, st);
    parrot_model::prof_end(p2, st);
    p2 = M.prof_begin("tail_speaker", st);
    speaker_grads(M, st);
    parrot_model::prof_end(p2, st);
  }
  parrot_model::prof_end(pe, st);
}

// ------------------------------------------------------------------ sampling
static void sample_scan_body(parrot_model& M, const int32_t* d_labels, const float* d_lmask, const int32_t* d_speaker,
                             const float* d_unis, const float* d_normals, uint64_t seed,
                             const unsigned long long* d_seed, cudaStream_t st) {
  const Dims& d = M.d;
  prep_base(M, d_speaker, st);
  encoder_fwd(M, d_labels, d_lmask, st);
  init_slots(M, true, st);   // model.py:830-832: always the initial states
  const Plane* px = d.weak ? &M.planes.at("xin") : nullptr;
  if (px) {
    CK(cudaMemsetAsync(px->hi, 0, (size_t)d.Np * px->pitch * 2, st));  // x_0 = 0 (model.py:834-835)
    CK(cudaMemsetAsync(px->lo, 0, (size_t)d.Np * px->pitch * 2, st));
  }
  const long long row = (long long)3 * d.H;
  const NormParts gp = gru_parts(d);
  auto norm_into = [&](const std::string& q, const std::string& pre) {   // layer_norm: pre += norm(q), one step
    LAUNCH(rownorm_fwd_kernel, dim3(d.B, 2), 256, 0, st, M.fbuf(q), row, M.fbuf(pre), row, gp, 1);
  };
  for (int t = 0; t < d.T; ++t) {
    if (d.ln) {
      for (int l = 0; l < 3; ++l)
        CK(cudaMemcpyAsync(M.fbuf("preT" + LN(l)), M.fbuf("base" + LN(l)), (size_t)d.B * row * 4,
                           cudaMemcpyDeviceToDevice, st));
      if (d.weak) {   // model.py:899-924: norm(Fork(x_{t-1})) per layer
        run_table(M, "ln_fb", t, d.T, 0, st);
        for (int l = 0; l < 3; ++l)
          if (l == 0 || d.full) norm_into("qfb" + LN(l), "preT" + LN(l));
      }
    }
    run_table(M, "sampA1", t, d.T, 0, st);
    run_table(M, "sampB1", t, d.T, 0, st);
    attention_step(M, t, true, st);
    if (d.ln) {
      run_table(M, "lnQ1", t, d.T, 0, st);
      norm_into("q12", "preT2"); norm_into("q13", "preT3");
    }
    run_table(M, "sampA2", t, d.T, 0, st);
    run_table(M, "sampB2", t, d.T, 0, st);
    if (d.ln) {
      run_table(M, "lnQ2", t, d.T, 0, st);
      norm_into("q23", "preT3");
    }
    run_table(M, "sampA3", t, d.T, 0, st);
    run_table(M, "sampB3", t, d.T, 0, st);
    if (d.spk) {
      CK(cudaMemcpyAsync(M.fbuf("ro"), M.fbuf("spk_ro"), (size_t)d.B * d.R * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(M.fbuf("pred"), M.fbuf("spk_out"), (size_t)d.B * d.Dtot * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (d.ln) {
      run_table(M, "ln_ro", t, d.T, 0, st);
      for (int l = 0; l < 3; ++l)
        LAUNCH(rownorm_fwd_kernel, dim3(d.B, 1), 256, 0, st, M.fbuf("ropre" + LN(l)), (long long)d.R, M.fbuf("ro"),
               (long long)d.R, row_parts(d.R), (l > 0 || d.spk) ? 1 : 0);
      run_table(M, "ln_ro_att", t, d.T, 0, st);
    } else
    run_table(M, "readout", t, d.T, 0, st);
    run_table(M, "output", t, d.T, 0, st);
    SampleArgs a;
    memset(&a, 0, sizeof a);
    a.B = d.B; a.D = d.D; a.k = d.K; a.Dtot = d.Dtot; a.which = d.gmm ? 1 : 0;
    a.eps = M.cfg.epsilon; a.bias = M.cfg.sampling_bias;
    a.pred = M.fbuf("pred");
    a.unis = d_unis ? d_unis + (long long)t * d.B : nullptr;
    a.normals = d_normals ? d_normals + (long long)t * d.B * d.D : nullptr;
    a.seed = seed; a.step = t; a.seed_ptr = d_seed;
    a.x_out = M.fbuf("samp_x") + (long long)t * d.B * d.D;
    a.pi_out = M.fbuf("samp_pi") + (long long)t * d.B * (d.gmm ? d.K : d.D);
    if (px) {
      a.x_hi = px->hi + (long long)(t + 1) * d.Np * px->pitch;
      a.x_lo = px->lo + (long long)(t + 1) * d.Np * px->pitch;
      a.Np = d.Np; a.Dp = px->pitch;
    }
    LAUNCH(sample_emit_kernel, cdiv((long long)d.B * 32, 128), 128, 0, st, a);
  }
}

// parrot_sample_scan: the T-step loop above is ~11 dependent launches per step (each step consumes the frame the
// previous one emitted), i.e. launch-latency bound.  The whole loop is captured ONCE per handle into a CUDA graph whose
// nodes read staged copies of the call's inputs (labels, masks, speaker ids, injected noise, Philox seed) from fixed
// workspace buffers; a call then costs the input copies + one graph launch.  PARROT_NO_GRAPH=1 (or per-launch
// profiling / debug sync) runs the plain loop.
static bool sample_graph_enabled(parrot_model& M) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PARROT_NO_GRAPH");
    env = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return env && !M.profiling && !debug_sync();
}
static void sample_scan(parrot_model& M, const int32_t* d_labels, const float* d_lmask, const int32_t* d_speaker,
                        const float* d_unis, const float* d_normals, uint64_t seed, cudaStream_t st) {
  const Dims& d = M.d;
  REQUIRE(d.sampling, "parrot_sample_scan needs a handle created with cfg.sampling = 1");
  if (M.dirty) pack_weights(M, st);
  if (!sample_graph_enabled(M)) {
    sample_scan_body(M, d_labels, d_lmask, d_speaker, d_unis, d_normals, seed, nullptr, st);
    return;
  }
  int32_t* s_lab = (int32_t*)(M.ws + M.bufs.at("samp_in_labels").off);
  int32_t* s_spk = (int32_t*)(M.ws + M.bufs.at("samp_in_speaker").off);
  unsigned long long* s_seed = (unsigned long long*)(M.ws + M.bufs.at("samp_in_seed").off);
  float* s_lm = M.fbuf("samp_in_lmask");
  float* s_un = M.fbuf("samp_in_unis");
  float* s_no = M.fbuf("samp_in_normals");
  CK(cudaMemcpyAsync(s_lab, d_labels, (size_t)d.B * d.U * 4, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(s_lm, d_lmask, (size_t)d.B * d.U * 4, cudaMemcpyDeviceToDevice, st));
  if (d_speaker) CK(cudaMemcpyAsync(s_spk, d_speaker, (size_t)d.B * 4, cudaMemcpyDeviceToDevice, st));
  if (d_unis) CK(cudaMemcpyAsync(s_un, d_unis, (size_t)d.T * d.B * 4, cudaMemcpyDeviceToDevice, st));
  if (d_normals) CK(cudaMemcpyAsync(s_no, d_normals, (size_t)d.T * d.B * d.D * 4, cudaMemcpyDeviceToDevice, st));
  const unsigned long long seed64 = seed;
  CK(cudaMemcpyAsync(s_seed, &seed64, 8, cudaMemcpyHostToDevice, st));   // (pageable source: copied before return)
  const int key = (d_speaker ? 1 : 0) | (d_unis ? 2 : 0) | (d_normals ? 4 : 0);
  auto it = M.samp_graphs.find(key);
  if (it == M.samp_graphs.end()) {
    // captured on a private stream (the caller's may be the legacy default stream, which cannot be captured)
    cudaGraph_t graph = nullptr;
    cudaStream_t cap = nullptr;
    CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaError_t be = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (be != cudaSuccess) { cudaStreamDestroy(cap); CK(be); }
    try {
      sample_scan_body(M, s_lab, s_lm, d_speaker ? s_spk : nullptr, d_unis ? s_un : nullptr,
                       d_normals ? s_no : nullptr, 0, s_seed, cap);
    } catch (...) {
      cudaStreamEndCapture(cap, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaStreamDestroy(cap);
      throw;
    }
    cudaError_t ee = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    CK(ee);
    cudaGraphExec_t exec = nullptr;
    CK(cudaGraphInstantiate(&exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    it = M.samp_graphs.emplace(key, exec).first;
  }
  CK(cudaGraphLaunch(it->second, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

// =========================================================================== C ABI
extern "C" {

const char* parrot_last_error(void) { return g_err.c_str(); }
int parrot_abi_version(void) { return 1; }
int64_t parrot_launch_count(void) { return g_launches.load(); }

int parrot_param_count(const parrot_config* cfg, int32_t* count, int64_t* total_floats) {
  return guard([&] {
    PReg P = make_params(*cfg);
    *count = (int32_t)P.v.size();
    *total_floats = P.total;
  });
}
int parrot_param_info(const parrot_config* cfg, int32_t i, char* name, int32_t name_cap, int64_t* offset,
                      int32_t* rows, int32_t* cols) {
  return guard([&] {
    PReg P = make_params(*cfg);
    REQUIRE(i >= 0 && i < (int)P.v.size(), "parameter index out of range");
    snprintf(name, name_cap, "%s", P.v[i].name.c_str());
    *offset = P.v[i].off; *rows = P.v[i].rows; *cols = P.v[i].cols;
  });
}

int parrot_workspace_bytes(const parrot_config* cfg, size_t* bytes) {
  return guard([&] {
    check_cfg(*cfg);
    parrot_model M;
    M.cfg = *cfg; M.d = make_dims(*cfg); M.P = make_params(*cfg); M.dry = true;
    build(M);
    build_device_tables(M);
    *bytes = M.ws_used + 4096;
  });
}

int parrot_create(const parrot_config* cfg, float* d_params, float* d_grads, void* d_workspace,
                  size_t workspace_bytes, void* stream, parrot_model** out) {
  return guard([&] {
    check_cfg(*cfg);
    cudaStream_t st = (cudaStream_t)stream;
    parrot_model* M = new parrot_model;
    try {
      M->cfg = *cfg; M->d = make_dims(*cfg); M->P = make_params(*cfg);
      M->params = d_params; M->grads = d_grads;
      M->ws = (uint8_t*)d_workspace; M->ws_bytes = workspace_bytes; M->dry = false;
      REQUIRE(((uintptr_t)d_workspace & 1023) == 0, "workspace must be 1024-byte aligned");
      CK(cudaMemsetAsync(d_workspace, 0, workspace_bytes, st));
      build(*M);
      build_device_tables(*M);
      upload_tables(*M, st);
      // split-K scratch of the grouped scans: every slot starts as SPLIT_SENTINEL (engine.cuh q_finish); the readers
      // restore the sentinel after consuming a slot, so this is needed once
      if (M->uniq_split_floats > 0)
        CK(cudaMemsetAsync(M->d_split_scratch + std::max<size_t>(M->max_split_floats, 1), 0xFF,
                           (size_t)M->uniq_split_floats * 4, st));
      // constant 1.0 for unnormalised backward: cost[4]
      const float one = 1.0f;
      CK(cudaMemcpyAsync(M->fbuf("cost") + 4, &one, 4, cudaMemcpyHostToDevice, st));
      ensure_kernel_attrs();
      {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&M->sm_count, cudaDevAttrMultiProcessorCount, dev));
      }
      CK(cudaStreamSynchronize(st));  // host-side tables are read by the copies above
      M->dirty = true;
    } catch (...) {
      delete M;
      throw;
    }
    *out = M;
  });
}
int parrot_destroy(parrot_model* m) {
  return guard([&] {
    for (auto& kv : m->samp_graphs) cudaGraphExecDestroy(kv.second);
    g_inputs.erase(m);
    delete m;
  });
}
int parrot_buffer_info(parrot_model* m, const char* name, int64_t* byte_offset, int64_t* numel) {
  return guard([&] {
    auto it = m->bufs.find(name);
    REQUIRE(it != m->bufs.end(), std::string("unknown buffer ") + name);
    *byte_offset = (int64_t)it->second.off;
    *numel = (int64_t)(it->second.bytes / 4);
  });
}
int parrot_pack_weights(parrot_model* m, void* stream) {
  return guard([&] { pack_weights(*m, (cudaStream_t)stream); });
}
int parrot_mark_params_dirty(parrot_model* m) {
  m->dirty = true;
  return 0;
}
/* debug: launch table `name` `reps` times back to back at scan tick `tick`; returns the average milliseconds per
 * launch.  If d_timeline is not null, the LAST launch writes [cta][16] globaltimer stamps there. */
int parrot_debug_time_table(parrot_model* m, const char* name, int tick, int reverse, int reps, float* avg_ms,
                            unsigned long long* d_timeline, void* stream) {
  return guard([&] {
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const bool plain = m->tables.at(name).n_cols == NT;
    const int T = plain ? 1 : m->d.T;
    run_table(*m, name, tick, T, reverse, st);
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) {
      if (i == reps - 1) m->timeline = d_timeline;
      run_table(*m, name, tick, T, reverse, st);
    }
    m->timeline = nullptr;
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *avg_ms = ms / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  });
}
int parrot_debug_set_stamps(parrot_model* m, unsigned long long* d_stamps, int bars) {
  // bars < 0: d_stamps is instead a [2][148][16] intra-phase timeline buffer for forward tick (-bars)
  if (!d_stamps) { m->timeline = nullptr; m->tl_tick = -1; m->stamps = nullptr; m->stamps_bwd = nullptr; m->stamp_bars = 0; return 0; }
  if (bars < 0) { m->timeline = d_stamps; m->tl_tick = -bars; return 0; }
  if (bars >= (1 << 20)) { m->stamps_bwd = d_stamps; m->stamp_bars = bars - (1 << 20); m->stamps = nullptr; return 0; }   // backward sweep
  m->stamps = d_stamps; m->stamp_bars = bars; m->stamps_bwd = nullptr;
  return 0;
}
int parrot_set_profiling(parrot_model* m, int enable) {
  return guard([&] {
    m->profiling = enable;
    for (auto& r : m->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    m->prof.clear();
  });
}
int parrot_get_profile(parrot_model* m, const char* key, double* total_ms, int64_t* launches) {
  return guard([&] {
    CK(cudaDeviceSynchronize());
    double tot = 0.0;
    int64_t n = 0;
    for (auto& r : m->prof)
      if (r.key == key) {
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        tot += ms;
        ++n;
      }
    *total_ms = tot;
    *launches = n;
  });
}

int parrot_encoder_fwd(parrot_model* m, const int32_t* d_labels, const float* d_labels_mask, void* stream) {
  return guard([&] {
    if (m->dirty) pack_weights(*m, (cudaStream_t)stream);
    g_inputs[m].labels = d_labels; g_inputs[m].lmask = d_labels_mask;
    encoder_fwd(*m, d_labels, d_labels_mask, (cudaStream_t)stream);
  });
}
int parrot_encoder_bwd(parrot_model* m, void* stream) {
  return guard([&] { encoder_bwd(*m, g_inputs[m].lmask, (cudaStream_t)stream); });
}
int parrot_decoder_scan_fwd(parrot_model* m, const float* d_features, const float* d_feedback_noise,
                            float noise_level, float start_flag, void* stream) {
  return guard([&] {
    if (m->dirty) pack_weights(*m, (cudaStream_t)stream);
    g_inputs[m].features = d_features;
    scan_fwd(*m, d_features, d_feedback_noise, noise_level, start_flag, (cudaStream_t)stream);
  });
}
int parrot_decoder_scan_bwd(parrot_model* m, void* stream) {
  return guard([&] { scan_bwd(*m, (cudaStream_t)stream); });
}
int parrot_readout_emit_fwd(parrot_model* m, const float* d_features, const float* d_features_mask, float* d_cost,
                            void* stream) {
  return guard([&] {
    g_inputs[m].features = d_features; g_inputs[m].fmask = d_features_mask;
    readout_emit_fwd(*m, d_features, d_features_mask, d_cost, (cudaStream_t)stream);
  });
}
int parrot_readout_emit_bwd(parrot_model* m, int unnormalised, void* stream) {
  return guard([&] { readout_emit_bwd(*m, unnormalised, (cudaStream_t)stream); });
}

int parrot_attention_step(const parrot_config* cfg, const float* d_h1, const float* d_wT, const float* d_batt,
                          const float* d_ctx, const float* d_k_prev, float* d_k_out, float* d_w_out,
                          float* d_phi_out, float* d_ab_out, float* d_e_out, int training, void* stream) {
  return guard([&] {
    const Dims d = make_dims(*cfg);
    AttnFwdArgs a;
    memset(&a, 0, sizeof a);
    a.B = d.B; a.U = d.U; a.C = d.C; a.A = d.A; a.H = d.H; a.Np = d.Np; a.Cp = d.Cp;
    a.type = cfg->attention_type; a.eps = cfg->epsilon; a.align = cfg->attention_alignment;
    a.sharp = training ? 1.0f : cfg->sharpening_coeff;
    a.timing = training ? 1.0f : cfg->timing_coeff;
    a.h1 = d_h1; a.wT = d_wT; a.batt = d_batt; a.ctx = d_ctx; a.k_prev = d_k_prev; a.k_out = d_k_out;
    a.w_out = d_w_out; a.phi_out = d_phi_out; a.ab_out = d_ab_out; a.e_out = d_e_out;
    const size_t smem = (size_t)(rup(d.H, 4) + 2 * rup(3 * d.A, 4) + rup(d.U, 4) + 8 * d.C) * 4;
    ensure_kernel_attrs();
    // the caller's e_out buffer ([B][3A]) doubles as the projection scratch: it is consumed before it is rewritten
    a.hat = d_e_out;
    (void)smem;
    const int nparts = attention_nparts(d.B, d.C, 148);
    if (3 * d.A <= 32 && d.H % 4 == 0) {
      LAUNCH(attention_step_kernel, d.B * nparts, 256, att_window_smem(d, nparts) + 8 * 32 * 4, (cudaStream_t)stream, a,
             nparts);
    } else {
      LAUNCH(attention_proj_kernel, 148, 256, 0, (cudaStream_t)stream, a);
      LAUNCH(attention_window_kernel, d.B * nparts, 256, att_window_smem(d, nparts), (cudaStream_t)stream, a, nparts, 0);
    }
  });
}

int parrot_compute_cost(parrot_model* m, const float* d_features, const float* d_features_mask,
                        const int32_t* d_labels, const float* d_labels_mask, const int32_t* d_speaker,
                        float start_flag, const float* d_feedback_noise, float noise_level,
                        const float* d_gmm_unis, const float* d_gmm_normals, float* d_cost, void* stream) {
  return guard([&] {
    cudaStream_t st = (cudaStream_t)stream;
    parrot_model& M = *m;
    REQUIRE(!M.d.sampling, "parrot_compute_cost needs a training handle (cfg.sampling = 0)");
    FwdInputs& in = g_inputs[m];
    in.features = d_features; in.fmask = d_features_mask; in.labels = d_labels; in.lmask = d_labels_mask;
    in.speaker = d_speaker;
    cudaEvent_t pe = M.prof_begin("sec_pack_prep_encoder", st);
    if (M.dirty) pack_weights(M, st);
    prep_base(M, d_speaker, st);
    encoder_fwd(M, d_labels, d_labels_mask, st);
    parrot_model::prof_end(pe, st);
    pe = M.prof_begin("sec_scan_fwd", st);
    scan_fwd(M, d_features, d_feedback_noise, noise_level, start_flag, st);
    parrot_model::prof_end(pe, st);
    pe = M.prof_begin("sec_readout_emit_fwd", st);
    readout_emit_fwd(M, d_features, d_features_mask, d_cost, st);
    parrot_model::prof_end(pe, st);
    if (M.d.gmm && d_gmm_unis && d_gmm_normals) {
      // next_x = sample_gmm(mu, sigma, coeff) (model.py:782), all frames at once, no sampling bias
      SampleArgs a;
      memset(&a, 0, sizeof a);
      a.B = M.d.T * M.d.B; a.D = M.d.D; a.k = M.d.K; a.Dtot = M.d.Dtot; a.which = 1;
      a.eps = M.cfg.epsilon; a.bias = 0.0f;
      a.pred = M.fbuf("pred"); a.unis = d_gmm_unis; a.normals = d_gmm_normals;
      a.x_out = M.fbuf("next_x"); a.pi_out = M.fbuf("dpred");  // coeff parked in dpred until backward
      LAUNCH(sample_emit_kernel, cdiv((long long)a.B * 32, 128), 128, 0, st, a);
    }
    apply_updates(M, st);
    M.have_fwd = true;
  });
}
int parrot_backward(parrot_model* m, int unnormalised, void* stream) {
  return guard([&] { backward(*m, unnormalised, (cudaStream_t)stream); });
}
int parrot_sample_scan(parrot_model* m, const int32_t* d_labels, const float* d_labels_mask,
                       const int32_t* d_speaker, const float* d_unis, const float* d_normals, uint64_t seed,
                       void* stream) {
  return guard([&] { sample_scan(*m, d_labels, d_labels_mask, d_speaker, d_unis, d_normals, seed, (cudaStream_t)stream); });
}

int parrot_adam_clip_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                          float grad_scale, const float* d_mask_sum, float threshold, float learning_rate,
                          float beta1, float beta2, float epsilon, int64_t time_step, float* d_stats,
                          double* d_scratch, void* stream) {
  return guard([&] {
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = 592;
    LAUNCH(sumsq_partial_kernel, parts, 1024, 0, st, d_grads, (long long)n, d_scratch);
    LAUNCH(clip_finalize_kernel, 1, 32, 0, st, d_scratch, parts, grad_scale, d_mask_sum, threshold, d_stats);
    // blocks Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
    const double t1 = (double)time_step;
    const float lr_t = (float)(learning_rate * std::sqrt(1.0 - std::pow((double)beta2, t1)) /
                               (1.0 - std::pow((double)beta1, t1)));
    LAUNCH(adam_kernel, 148 * 8, 256, 0, st, d_params, d_grads, d_m, d_v, (long long)n, d_stats, lr_t, beta1, beta2,
           epsilon);
  });
}

// ------------------------------------------------------------------ data-parallel collective (NCCL, run-time bound)
}  // extern "C"
#include <dlfcn.h>
namespace {
struct NcclId { char internal[PARROT_COMM_ID_BYTES]; };
typedef void* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& nccl_api() {
  static NcclApi api;
  if (api.handle) return api;
  // the copy the process already uses (torch bundles one) before anything else: two NCCL instances in one process
  // would each build their own transports
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  REQUIRE(h != nullptr, std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""));
  auto sym = [&](const char* n) {
    void* p = dlsym(h, n);
    REQUIRE(p != nullptr, std::string("NCCL symbol missing: ") + n);
    return p;
  };
  api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
  api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
  api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
  api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
  api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  api.handle = h;
  return api;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    const char* es = nccl_api().GetErrorString ? nccl_api().GetErrorString(rc) : "?";
    throw std::runtime_error(std::string("parrot_b200: ") + what + " failed: " + es);
  }
}
}  // namespace
struct parrot_comm {
  int nranks = 1, rank = 0, version = 0;
  NcclComm comm = nullptr;
};
extern "C" {

int parrot_comm_unique_id(void* id128) {
  return guard([&] {
    REQUIRE(id128 != nullptr, "id buffer is null");
    NcclId id;
    nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
  });
}
int parrot_comm_init(int32_t nranks, int32_t rank, const void* id128, parrot_comm** out) {
  return guard([&] {
    REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    parrot_comm* c = new parrot_comm;
    c->nranks = nranks; c->rank = rank;
    if (nranks > 1) {
      try {
        REQUIRE(id128 != nullptr, "a communicator of more than one rank needs the unique id of rank 0");
        NcclId id;
        memcpy(&id, id128, sizeof id);
        nccl_check(nccl_api().GetVersion(&c->version), "ncclGetVersion");
        nccl_check(nccl_api().CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
      } catch (...) {
        delete c;
        throw;
      }
    }
    *out = c;
  });
}
int parrot_comm_allreduce(parrot_comm* comm, float* d_buf, int64_t count, void* stream) {
  return guard([&] {
    REQUIRE(comm != nullptr, "communicator is null");
    if (comm->nranks == 1) return;   // identity
    const int kFloat = 7, kSum = 0;  // ncclFloat32, ncclSum
    nccl_check(nccl_api().AllReduce(d_buf, d_buf, (size_t)count, kFloat, kSum, comm->comm, (cudaStream_t)stream),
               "ncclAllReduce");
  });
}
int parrot_comm_info(parrot_comm* comm, int32_t* nranks, int32_t* rank, int32_t* nccl_version) {
  return guard([&] {
    REQUIRE(comm != nullptr, "communicator is null");
    if (nranks) *nranks = comm->nranks;
    if (rank) *rank = comm->rank;
    if (nccl_version) *nccl_version = comm->version;
  });
}
int parrot_comm_destroy(parrot_comm* comm) {
  return guard([&] {
    if (!comm) return;
    if (comm->comm) nccl_check(nccl_api().CommDestroy(comm->comm), "ncclCommDestroy");
    delete comm;
  });
}

size_t parrot_gemm_nt_workspace_bytes(int32_t Mr, int32_t N, int32_t K) {
  const size_t kp = (size_t)rup(K, 64);
  size_t b = 0;
  b += 2 * ((size_t)rup(Mr, 128) * kp * 2 + 1024);
  b += 2 * ((size_t)3 * rup(N, 256) * kp * 2 + 1024);
  b += (size_t)cdiv(Mr, 128) * cdiv(N, NT) * sizeof(Job) + 16 * sizeof(CUtensorMap) + 16 * sizeof(MapRaw) +
       sizeof(ScanCtx) + 16 * 1024;
  return b;
}

int parrot_gemm_nt(const float* d_A, const float* d_B, float* d_C, int32_t Mr, int32_t N, int32_t K, int32_t impl_flags,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
  return guard([&] {
    cudaStream_t st = (cudaStream_t)stream;
    // impl_flags: bit 0 = SIMT twin; bits 8..15 = sample tile (0 -> 128); bit 16 = 3-D (slot-indexed) B map;
    // bit 17 = split K into two segments
    const int impl = impl_flags & 1;
    int ntile = (impl_flags >> 8) & 0xff;
    if (ntile == 0) ntile = NT;
    if (ntile == 255) ntile = 256;
    const bool use3d = (impl_flags >> 16) & 1, twoseg = (impl_flags >> 17) & 1;
    parrot_model M;
    memset(&M.cfg, 0, sizeof M.cfg);
    M.cfg.gemm_impl = impl;
    M.ws = (uint8_t*)d_workspace; M.ws_bytes = workspace_bytes; M.dry = false;
    REQUIRE(((uintptr_t)d_workspace & 1023) == 0, "workspace must be 1024-byte aligned");
    CK(cudaMemsetAsync(d_workspace, 0, workspace_bytes, st));
    Plane pa = M.make_plane("A", rup(Mr, 128), K, 1);
    // 3-D variant: B is one slot of rup(N, ntile) rows inside a 3-slot plane, addressed as slot 1 (t=0, b_slot=1)
    const int brows = use3d ? rup(N, ntile) : rup(N, 128);
    Plane pb = M.make_plane("B", brows, K, use3d ? 3 : 1);
    REQUIRE(!use3d || N <= ntile, "3-D test needs N <= tile");
    const int ma = M.make_map(pa, 2, 128);
    const int mb = M.make_map(pb, use3d ? 3 : 2, ntile);
    pack_plane(st, d_A, K, Mr, K, pa, 0);
    Plane pb1 = pb;
    if (use3d) { pb1.hi += (long long)brows * pb.pitch; pb1.lo += (long long)brows * pb.pitch; }
    pack_plane(st, d_B, K, N, K, pb1, 0);
    std::vector<Job> js;
    PlainArgs pargs;
    memset(&pargs, 0, sizeof pargs);
    pargs.out = d_C; pargs.ldo = N; pargs.n_total = N; pargs.flags = PF_TRANS; pargs.scale = 1.0f;
    const int nkb = cdiv(K, 64);
    for (int mt = 0; mt < cdiv(Mr, 128); ++mt)
      for (int nt = 0; nt < cdiv(N, ntile); ++nt) {
        Job j = blank_job();
        j.epi = EPI_PLAIN; j.row0 = mt * 128; j.m_valid = std::min(128, Mr - mt * 128); j.n0 = nt * ntile;
        const int slot = use3d ? 1 : NO_SLOT;
        if (twoseg && nkb >= 2) {
          const int k1 = nkb / 2;
          j.nseg = 2;
          j.seg[0] = mkseg(ma, mt * 128, 0, mb, nt * ntile, 0, slot, k1);
          j.seg[1] = mkseg(ma, mt * 128, k1 * 64, mb, nt * ntile, k1 * 64, slot, nkb - k1);
        } else {
          j.nseg = 1;
          j.seg[0] = mkseg(ma, mt * 128, 0, mb, nt * ntile, 0, slot, nkb);
        }
        j.pa = pargs;
        js.push_back(j);
      }
    push_table(M, "g", js, ntile);
    build_device_tables(M);
    memset(&M.ctx, 0, sizeof M.ctx);
    upload_tables(M, st);
    ensure_kernel_attrs();
    run_table(M, "g", 0, 1, 0, st);
    CK(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
