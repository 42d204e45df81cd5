// myfm_b200 engine: device-resident Gibbs trainer, prediction datasets and the C ABI
// (include/myfm_b200.h).  One CUDA stream per trainer; data is uploaded once and a whole
// regression sweep runs without any host round trip.
#include "../../include/myfm_b200.h"

#include "common.cuh"
#include "host_data.hpp"
#include "kernels.cuh"
#include "field_sweep.cuh"
#include "tile_sweep.cuh"
#include "latent_device.cuh"
#include "eval_device.cuh"
#include "prep_device.cuh"
#include "mt_device.cuh"
#include "mt_jump.hpp"
#include "oprobit.cuh"
#include "rng.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <memory>
#include <thread>

namespace myfm {

namespace {
thread_local std::string g_last_error;

constexpr int REDUCE_BLOCKS = 592; // 4 x 148 SMs

// NCCL is bound at run time (dlopen by SONAME): a process that already carries an NCCL — the
// one bundled with torch, when the host program uses torch.distributed next to this library —
// must not end up with a second copy, and single-GPU use needs none at all.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;

  static NcclApi &get() {
    static NcclApi api = load();
    return api;
  }
  static NcclApi load() {
    NcclApi a;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle)
        break;
    }
    if (!a.handle)
      throw std::runtime_error("row-sharded training needs NCCL, but libnccl.so.2 could not be loaded.");
    auto sym = [&](const char *n) {
      void *p = dlsym(a.handle, n);
      if (!p)
        throw std::runtime_error(std::string("NCCL symbol missing: ") + n);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
  }
  void check(ncclResult_t r, const char *what) const {
    if (r != ncclSuccess)
      throw CudaError(std::string("NCCL error in ") + what + ": " + GetErrorString(r));
  }
};

int pow2_ceil_clamped(double x, int lo, int hi) {
  int v = lo;
  while (v < hi && v < x)
    v <<= 1;
  return v;
}
} // namespace

// ------------------------------------------------------------------------------------------------
template <typename Real> struct DevCs {
  int64_t n_major = 0, n_minor = 0, nnz = 0;
  DevBuf<int> ptr, idx;
  DevBuf<Real> val;
  void upload(const HostCs<Real> &h, cudaStream_t s) {
    n_major = h.n_major, n_minor = h.n_minor, nnz = h.nnz();
    ptr.upload(h.ptr, s);
    idx.upload(h.idx, s);
    val.upload(h.val, s);
  }
  CsView<Real> view() const { return CsView<Real>{ptr.p, idx.p, val.p}; }
  double avg_len() const { return n_major ? static_cast<double>(nnz) / n_major : 0.0; }
};

// Kernel timing by family with CUDA events on the launching stream (bench.py's roofline leg).
struct KernelTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Span {
    int family;
    size_t a, b;
  };
  std::vector<Span> spans;
  static constexpr int FAMILIES = 5; // 0 column sweeps, 1 q_init, 2 e_refresh, 3 streaming level, 4 gather level
  double ms[FAMILIES] = {0, 0, 0, 0, 0};
  int64_t launches[FAMILIES] = {0, 0, 0, 0, 0};
  ~KernelTimer() {
    for (auto e : pool)
      cudaEventDestroy(e);
  }
  size_t mark(cudaStream_t s) {
    if (used == pool.size()) {
      cudaEvent_t e;
      MYFM_CUDA(cudaEventCreate(&e));
      pool.push_back(e);
    }
    MYFM_CUDA(cudaEventRecord(pool[used], s));
    return used++;
  }
  void collect() { // caller has synchronised the stream
    for (auto &sp : spans) {
      float t = 0;
      MYFM_CUDA(cudaEventElapsedTime(&t, pool[sp.a], pool[sp.b]));
      ms[sp.family] += t;
      launches[sp.family]++;
    }
    spans.clear();
    used = 0;
  }
};

// Design matrix + relation blocks on the device, with the per-block tables of the forward pass.
template <typename Real> struct DevRelationData {
  int64_t S = 0, F = 0, offset = 0;
  DevCs<Real> B; // CSR of the block
  DevBuf<int> map;
  DevBuf<Real> tab_lin, tab_q, tab_qs;
};

struct DatasetBase {
  virtual ~DatasetBase() = default;
  int dtype = MYFM_DTYPE_F32;
  int device = 0;
  int64_t n_rows = 0, dim_main = 0, dim_all = 0;
};

// Running state of a per-iteration evaluation callback on the device (eval_device.cuh).
struct EvalState {
  DatasetBase *dataset = nullptr; // not owned
  int task = MYFM_TASK_REGRESSION, width = 1, n_samples = 0;
  int64_t n = 0;
  double clip_min = 0, clip_max = 0, eps = -1;
  DevBuf<double> sum, late, y, cutp, partial, out;
};

template <typename Real> struct Dataset : DatasetBase {
  DevCs<Real> X;
  std::vector<DevRelationData<Real>> rels;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int64_t *launch_counter = nullptr;
  int64_t own_counter = 0;
  // scratch for host-supplied samples
  DevBuf<Real> w, Vt, w0, score, accum, cutp;

  ~Dataset() override {
    if (own_stream && stream)
      cudaStreamDestroy(stream);
  }

  void count(int n = 1) { (launch_counter ? *launch_counter : own_counter) += n; }
  int row_len = 0;   // > 0: every row of X holds exactly this many entries
  bool unit = false; // every stored value of X is 1

  // Shape checks of util.hpp:147-165 / definitions.hpp:38-41, then upload.
  // `perm` (trainer only): device row i holds the caller's row perm[i]; Xh is already permuted.
  void build(const HostCs<Real> &Xh, int n_rel, const myfm_relation_t *relations, cudaStream_t s,
             const std::vector<int> *perm = nullptr) {
    stream = s;
    n_rows = Xh.n_major;
    dim_main = Xh.n_minor;
    dim_all = dim_main;
    X.upload(Xh, s);
    row_len = Xh.n_major ? Xh.ptr[1] - Xh.ptr[0] : 0;
    for (int64_t i = 0; i < Xh.n_major && row_len > 0; i++)
      if (Xh.ptr[i + 1] - Xh.ptr[i] != row_len)
        row_len = 0;
    unit = std::all_of(Xh.val.begin(), Xh.val.end(), [](Real v) { return v == Real(1); });
    if (n_rel > MAX_REL)
      throw std::invalid_argument("too many relation blocks (at most 8 are supported).");
    rels.resize(n_rel);
    for (int b = 0; b < n_rel; b++) {
      const myfm_relation_t &r = relations[b];
      if (n_rows != r.mapper_size) {
        std::ostringstream ss;
        ss << "main table has size " << n_rows << " but the relation[" << b << "] has size "
           << r.mapper_size;
        throw std::runtime_error(ss.str());
      }
      HostCs<Real> Bh = host_from_api<Real>(r.block, "relation block");
      std::vector<int> map(r.mapper_size);
      for (int64_t i = 0; i < r.mapper_size; i++) {
        const int64_t v = r.original_to_block[perm ? (*perm)[i] : i];
        if (v < 0 || v >= Bh.n_major)
          throw std::runtime_error("index mapping points to non-existing row.");
        map[i] = static_cast<int>(v);
      }
      DevRelationData<Real> &d = rels[b];
      d.S = Bh.n_major, d.F = Bh.n_minor, d.offset = dim_all;
      d.B.upload(Bh, s);
      d.map.upload(map, s);
      dim_all += d.F;
    }
    MYFM_CUDA(cudaStreamSynchronize(s)); // host staging vectors die here
  }

  void ensure_tables(int K) {
    for (auto &d : rels) {
      if (d.tab_lin.n < static_cast<size_t>(d.S))
        d.tab_lin.alloc(d.S);
      if (d.tab_q.n < static_cast<size_t>(d.S) * K) {
        d.tab_q.alloc(static_cast<size_t>(d.S) * K);
        d.tab_qs.alloc(static_cast<size_t>(d.S) * K);
      }
    }
  }

  // out = predict_score(X, rels) [- y]; FM.hpp:54-136 with all factors fused in one CSR pass.
  void predict(const Real *w_dev, const Real *Vt_dev, int K, const Real *w0_dev, const Real *y,
               Real *out, int out_stride = 1) {
    ensure_tables(K);
    RelPredictPack<Real> pack;
    pack.n = static_cast<int>(rels.size());
    for (int b = 0; b < pack.n; b++) {
      auto &d = rels[b];
      if (d.S) {
        k_block_tables<Real><<<ceil_div(d.S * 32, 256), 256, 0, stream>>>(
            static_cast<int>(d.S), d.B.view(), w_dev, Vt_dev, K, static_cast<int>(d.offset),
            d.tab_lin.p, d.tab_q.p, d.tab_qs.p);
        count();
      }
      pack.r[b] = RelPredictView<Real>{d.map.p, d.tab_lin.p, d.tab_q.p, d.tab_qs.p};
    }
    if (!n_rows)
      return;
    const int lpr = pow2_ceil_clamped(K / 4.0, 1, 32);
    const int n = static_cast<int>(n_rows);
    if (K >= 16 && K <= 64 && pack.n == 0 && row_len >= 1 && row_len <= 4 && (out_stride == 1 || out_stride == 2)) {
      // fixed-length rows, wide factors: warp per tile of 32 rows (k_predict_tile)
      int sms = 0, dev = 0;
      MYFM_CUDA(cudaGetDevice(&dev));
      MYFM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const int grid = std::min(ceil_div(static_cast<int64_t>(ceil_div(n, 32)) * 32, 256), sms * 8);
#define MYFM_TILE5(LL, U, KPL, PAIR, FULL)                                                         \
  k_predict_tile<Real, LL, U, KPL, PAIR, FULL><<<grid, 256, 0, stream>>>(n, X.idx.p, X.val.p, w_dev, Vt_dev, K, w0_dev, \
                                                                         y, out, out_stride);
#define MYFM_TILE4(LL, U, KPL)                                                                     \
  if (out_stride == 2) {                                                                           \
    if (K == 32 * KPL) {                                                                           \
      MYFM_TILE5(LL, U, KPL, true, true)                                                           \
    } else {                                                                                       \
      MYFM_TILE5(LL, U, KPL, true, false)                                                          \
    }                                                                                              \
  } else {                                                                                         \
    if (K == 32 * KPL) {                                                                           \
      MYFM_TILE5(LL, U, KPL, false, true)                                                          \
    } else {                                                                                       \
      MYFM_TILE5(LL, U, KPL, false, false)                                                         \
    }                                                                                              \
  }
#define MYFM_TILE3(LL, U)                                                                          \
  if (K <= 32) {                                                                                   \
    MYFM_TILE4(LL, U, 1)                                                                           \
  } else {                                                                                         \
    MYFM_TILE4(LL, U, 2)                                                                           \
  }
#define MYFM_TILE2(LL)                                                                             \
  case LL:                                                                                         \
    if (unit) {                                                                                    \
      MYFM_TILE3(LL, true)                                                                         \
    } else {                                                                                       \
      MYFM_TILE3(LL, false)                                                                        \
    }                                                                                              \
    break;
      switch (row_len) {
        MYFM_TILE2(1)
        MYFM_TILE2(2)
        MYFM_TILE2(3)
        MYFM_TILE2(4)
      }
#undef MYFM_TILE2
#undef MYFM_TILE3
#undef MYFM_TILE4
#undef MYFM_TILE5
      count();
      MYFM_CUDA(cudaGetLastError());
      return;
    }
    if (K >= 16 && K <= 64 && X.avg_len() <= 16) { // short rows, wide factors: warp per row tile
      const int warps = ceil_div(n, PREDICT_ROWS_PER_WARP);
      const int grid = ceil_div(static_cast<int64_t>(warps) * 32, 256);
      if (pack.n)
        k_predict_warp<Real, true><<<grid, 256, 0, stream>>>(n, X.view(), w_dev, Vt_dev, K, w0_dev, pack, y, out,
                                                             out_stride);
      else
        k_predict_warp<Real, false><<<grid, 256, 0, stream>>>(n, X.view(), w_dev, Vt_dev, K, w0_dev, pack, y, out,
                                                              out_stride);
      count();
      MYFM_CUDA(cudaGetLastError());
      return;
    }
#define MYFM_PREDICT(L)                                                                            \
  case L:                                                                                          \
    k_predict<Real, L><<<ceil_div(static_cast<int64_t>(n) * L, 256), 256, 0, stream>>>(            \
        n, X.view(), w_dev, Vt_dev, K, w0_dev, pack, y, out, out_stride);                          \
    break;
    switch (lpr) {
      MYFM_PREDICT(1)
      MYFM_PREDICT(2)
      MYFM_PREDICT(4)
      MYFM_PREDICT(8)
      MYFM_PREDICT(16)
      MYFM_PREDICT(32)
    }
#undef MYFM_PREDICT
    count();
    MYFM_CUDA(cudaGetLastError());
  }

  // Stage one host sample (float64, V row-major) into the scratch buffers.
  void stage_sample(double w0_h, const double *w_h, const double *V_h, int K) {
    std::vector<Real> wv(dim_all), Vv(static_cast<size_t>(dim_all) * K);
    for (int64_t i = 0; i < dim_all; i++)
      wv[i] = static_cast<Real>(w_h[i]);
    for (size_t i = 0; i < Vv.size(); i++)
      Vv[i] = static_cast<Real>(V_h[i]);
    Real w0v = static_cast<Real>(w0_h);
    w.upload(wv, stream);
    Vt.upload(Vv, stream);
    w0.upload(&w0v, 1, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
  }

  void check_dim(int64_t given) const { // FM.hpp:66-72 / predictor.hpp:24-33
    if (given != dim_all) {
      std::ostringstream ss;
      ss << "Total feature size mismatch. Should be " << given << ", but got " << dim_all << ".";
      throw std::invalid_argument(ss.str());
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Trainer
// ------------------------------------------------------------------------------------------------
// One posterior sample kept on the device (the FM copy GibbsFMTrainer keeps per kept iteration,
// FMTrainer.hpp:74-76): w0, w and the feature-major V, in the trainer's compute dtype.
struct SampleBase {
  virtual ~SampleBase() = default;
  virtual void get(double *w0, double *w, double *V) = 0;
  int dtype = MYFM_DTYPE_F32, device = 0, K = 0;
  int64_t dim_all = 0;
};
template <typename Real> struct Sample : SampleBase {
  DevBuf<Real> w0, w, Vt;
  void get(double *w0_out, double *w_out, double *V_out) override {
    MYFM_CUDA(cudaSetDevice(device));
    auto pull = [&](const DevBuf<Real> &buf, size_t n, double *out) {
      if (!out || !n)
        return;
      std::vector<Real> h(n);
      MYFM_CUDA(cudaMemcpy(h.data(), buf.p, n * sizeof(Real), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < n; i++)
        out[i] = h[i];
    };
    pull(w0, 1, w0_out);
    pull(w, dim_all, w_out);
    pull(Vt, static_cast<size_t>(dim_all) * K, V_out);
  }
};

struct TrainerBase {
  virtual ~TrainerBase() = default;
  virtual void init_fm(int rank, double init_std) = 0;
  virtual void step(int n) = 0;
  virtual void sync() = 0;
  virtual double timed_steps(int n) = 0;
  virtual void dims(int64_t *n_train, int64_t *dim_all, int32_t *rank, int32_t *n_groups) const = 0;
  virtual void get_fm(double *w0, double *w, double *V) = 0;
  virtual void get_cutpoints(int g, double *out) = 0;
  virtual void get_hyper(double *alpha, double *mu_w, double *lambda_w, double *mu_V,
                         double *lambda_V) = 0;
  virtual void get_e(double *e) = 0;
  virtual void get_q(double *q) = 0;
  virtual int64_t mh_accept(int g) = 0;
  virtual void set_state(const double *w0, const double *w, const double *V, const double *alpha,
                         const double *mu_w, const double *lambda_w, const double *mu_V,
                         const double *lambda_V, const double *e) = 0;
  virtual int64_t launch_count() const = 0;
  virtual int sweep_path() const = 0;
  virtual std::unique_ptr<SampleBase> snapshot() = 0;
  virtual void kernel_ms(int family, double *ms, int64_t *launches) = 0;
  virtual void set_profiling(bool on) = 0;
  virtual void predict_score(DatasetBase *d, double *out) = 0;
  virtual void evaluate(EvalState &ev, int iteration, const double *cutpoints, int n_cpt, double *terms) = 0;
  virtual int64_t get_variates(double *out, int64_t capacity) = 0;
  int dtype = MYFM_DTYPE_F32;
};

// NCCL communicators by (unique id, world, rank); see the Trainer constructor.
inline std::mutex &shared_comm_mutex() {
  static std::mutex m;
  return m;
}
inline std::map<std::string, ncclComm_t> &shared_comms() {
  static std::map<std::string, ncclComm_t> *cache = new std::map<std::string, ncclComm_t>(); // outlives static destruction
  return *cache;
}

template <typename Real> struct DevRelationTrain {
  DevCs<Real> Bt; // CSC of the block
  DevBuf<int> seg_ptr, seg_rows;
  DevBuf<Real> card, q, q_S, c, c_S, e, e_q;
  int n_levels = 0;
  DevBuf<int> level_ptr, level_cols;
  DevBuf<int> run_end; // [n_levels] end of the run of single-column levels a level belongs to (k_rel_sweep_smem)
  DevBuf<int4> level_rec;
  RelCache<Real> cache() { return RelCache<Real>{card.p, q.p, q_S.p, c.p, c_S.p, e.p, e_q.p}; }
};

template <typename Real> struct Trainer : TrainerBase {
  Config cfg;
  myfm_engine_options_t opt;
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  KernelTimer timer;

  Dataset<Real> data; // main table CSR + relation blocks (CSR, map, forward tables), device row order
  DevCs<Real> Xt;     // CSC of the main table (device row order)
  SweepPlan plan;
  std::vector<int> perm; // device row i holds the caller's row perm[i] (host_data.hpp)
  DevBuf<int> perm_dev;
  DevBuf<int> latent_row; // device row -> global row index (the key of the row's Philox stream)
  DevBuf<SweepItem> items;
  DevBuf<int> seg_count;
  DevBuf<Real> seg_partial, seg_theta_old;
  DevBuf<Real> dense_tmp; // [N] staging for boundary copies of e / q
  std::vector<DevRelationTrain<Real>> rel_train;

  // field path (field_sweep.cuh): main table = stack of position-aligned fields
  bool field_path = false;
  int f_tail = 0, f_last_base = 0, f_tab = 0, f_sm_count = 0, f_launch = 0;
  int f_nCC = 0, f_nCR = 0, f_nG = 0, f_nW = 0; // level-0 columns by length class (k_field_stream)
  bool f_pending_valid = false; // the last level's draw of the previous vector awaits the next pass
  SweepLevel f_level0, f_levelL;
  DevBuf<SweepItem> f_items0, f_itemsL;
  DevBuf<int> f_seg_countL, f_tail_idx, f_sched, f_chunk_done, f_item_slot0, f_item_slotL, f_colsL;
  DevBuf<Real> f_colstat; // row shards: column statistics of a level, summed over the ranks
  // Row shards whose level-0 columns are rank-exclusive (every such column has all its rows on one
  // rank, e.g. rows partitioned by user): level 0 needs no exchange at all — each rank sweeps the
  // columns it owns in the fused single pass, and the owners' draws are merged once per sweep.
  bool f_exclusive = false;
  std::vector<SweepItem> f_items0_host;
  std::vector<int> f_slot0_host, f_level_host;
  DevBuf<int> f_owner; // [dim_all] rank that holds the authoritative copy of w[j], V[j, :]
  // peer-memory exchange of those statistics (field_sweep.cuh: PeerView); falls back to NCCL
  bool peer_ok = false;
  int my_rank = 0;
  unsigned char *peer_local = nullptr;       // this rank's buffer: header, statistics buffer 0 and 1
  std::vector<unsigned char *> peer_base;    // every rank's buffer as mapped here (own: peer_local)
  size_t peer_stat_elems = 0;                // Reals per statistics buffer

  int f_ncols0 = 0, f_ncolsL = 0;
  DevBuf<Real> f_tail_val, f_own_val, f_pend_told, f_pend_tnew, f_partial;

  int64_t N = 0, D = 0, D_all = 0;
  int64_t N_global = 0; // training rows over all ranks (== N on one GPU)
  int K = -1, G = 0;
  // row-sharded data parallelism: one rank per GPU, statistics summed with NCCL
  int world = 1;
  ncclComm_t comm = nullptr;
  DevBuf<int> item_slot;
  DevBuf<Real> colstat, told;
  std::vector<Real> y_host;
  DevBuf<Real> y;
  DevBuf<Real> eq_buf;         // interleaved {e_i, q_i}, [2 N]
  DevBuf<Real> w, V, Vt;       // V column-major [D_all x K]; Vt feature-major mirror
  DevBuf<Real> hyper;          // alpha, w0, mu_w[G], lambda_w[G], mu_V[G*K], lambda_V[G*K]
  DevBuf<Real> scal;           // [0] = w0 delta
  DevBuf<Real> partial;        // grid-reduction partials
  DevBuf<int> group, feat_ptr, feat_idx;
  int hyper_chunks = 0;
  DevBuf<int> hyper_chunk_group, hyper_chunk_begin;
  DevBuf<Real> hyper_chunk_sums;
  DevBuf<unsigned int> hyper_done;

  MtStream<Real> rng;
  // MYFM_RNG_PHILOX: the per-row latent draws of classification / ordered probit run on the device
  // from counter-based streams (latent_device.cuh); the sweep's Gaussian / Gamma variates still come
  // from the (device-side) mt19937 stream, which the latent draws then no longer touch.
  bool philox_latents = false;
  uint64_t latent_seed = 0;
  SweepLayout layout;
  DevBuf<Real> z_dev;
  PinnedBuf<Real> z_pinned[2];
  cudaEvent_t z_copied[2] = {nullptr, nullptr};
  int64_t sweep_index = 0;
  // device-side mt19937 stream (regression; mt_device.cuh): runs one sweep ahead on its own stream
  bool device_rng = false;
  cudaStream_t rng_stream = nullptr;
  DevBuf<MtControl> mt_ctl;
  DevBuf<uint32_t> mt_ring;      // tempered words of the stream, ring[n & mt_mask]
  unsigned long long mt_mask = 0, mt_ahead = 0;
  int mt_final_stage = 0;
  DevBuf<int> mt_tile_count;
  DevBuf<long long> mt_tile_offset;
  DevBuf<Real> mt_consts; // a1[G+1], a2[G+1]
  DevBuf<Real> z_slot[2];
  cudaEvent_t z_ready[2] = {nullptr, nullptr}, z_free[2] = {nullptr, nullptr};
  int64_t gen_index = 0;
  const Real *z_last = nullptr;
  std::vector<Real> shapes_lw; // gamma shape per group
  Real shape_alpha = 0;
  std::vector<Real> e_host;
  // ordered probit: one cut-point sampler per cut-point group (FMTrainer.hpp:101-116, 513-521)
  struct CutGroup {
    int n_class = 0;
    std::vector<int64_t> rows;         // caller's row ids in the order given (the latent draws follow it)
    DevBuf<int> class_ptr, class_rows; // device rows of the group, grouped by label
    std::unique_ptr<CutpointSampler<Real>> sampler;
    std::vector<Real> cutpoints;
  };
  std::vector<CutGroup> cut_groups;
  DevBuf<Real> op_gamma, op_partial, op_sums;

  Trainer(const myfm_csr_t &X_api, int n_rel, const myfm_relation_t *relations, const double *y_api,
          int64_t n_y, int seed, const myfm_config_t &cfg_api, const myfm_engine_options_t &o)
      : cfg(cfg_api), opt(o), device(o.device), rng(seed) {
    dtype = o.dtype;
    if (o.rng != MYFM_RNG_MT19937 && o.rng != MYFM_RNG_PHILOX)
      throw std::invalid_argument("unknown rng.");
    philox_latents = o.rng == MYFM_RNG_PHILOX;
    latent_seed = (static_cast<uint64_t>(static_cast<uint32_t>(seed)) << 32) ^ 0x9E3779B97F4A7C15ull;
    {
      std::seed_seq seq{static_cast<uint32_t>(seed), 0x6d79666du, 0x63757470u}; // "myfm", "cutp"
      cut_gen.seed(seq);
    }
    world = o.world_size > 1 ? o.world_size : 1;
    if (world > 1) {
      if (!o.nccl_unique_id)
        throw std::invalid_argument("world_size > 1 needs the ncclUniqueId shared by all ranks.");
      if (n_rel > 0)
        throw std::runtime_error("relation blocks are not supported with row-sharded training yet.");
      if (cfg.task_type != MYFM_TASK_REGRESSION && !philox_latents)
        throw std::runtime_error("row-sharded classification / ordered probit needs rng=philox: the reference's "
                                 "latent draws consume the mt19937 stream row by row.");
    }
    const char *trace_env = std::getenv("MYFM_TRACE_SETUP");
    const bool trace = trace_env && trace_env[0] == '1';
    auto t_last = std::chrono::steady_clock::now();
    auto tick = [&](const char *what) {
      if (!trace)
        return;
      auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[myfm_b200 setup] %-28s %8.1f ms\n", what,
                   std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    if (X_api.n_rows != n_y) { // BaseFMTrainer.hpp:69-76
      std::ostringstream ss;
      ss << "Shape mismatch: X has size " << X_api.n_rows << " and y has size " << n_y;
      throw std::runtime_error(ss.str());
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= device)
      throw CudaError("no usable CUDA device: the myfm_b200 engine has no CPU fallback.");
    MYFM_CUDA(cudaSetDevice(device));
    MYFM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    tick("CUDA context + streams");
    MYFM_CUDA(cudaStreamCreateWithFlags(&rng_stream, cudaStreamNonBlocking));
    for (auto &ev : z_copied)
      MYFM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : z_ready)
      MYFM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : z_free)
      MYFM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    data.launch_counter = &launches;

    // Field-shaped main tables are prepared on the device (prep_device.cuh); everything else — and
    // any input that path declines — on the host (host_data.hpp), which also owns the error messages.
    HostCs<Real> Xh;
    const bool device_prepared = prepare_on_device(X_api, n_rel, o, tick);
    if (!device_prepared) {
      Xh = host_from_api<Real>(X_api, "X");
      tick("copy + validate CSR");
      if (has_duplicate_entries(Xh))
        throw std::invalid_argument(
            "X lists the same (row, column) twice; sum duplicates first (X.sum_duplicates()).");
      tick("duplicate check");
      // dependency levels, device row order, sweep work items (host_data.hpp)
      HostCs<Real> Xth0 = host_transpose(Xh);
      tick("transpose");
      int n_levels = 0, primary = -1;
      std::vector<int> level;
      if (o.column_level) { // agreed between the ranks (myfm_level_relax + max all-reduce)
        if (o.n_column_level != Xth0.n_major)
          throw std::invalid_argument("column_level must have one entry per main-table column.");
        level.assign(o.column_level, o.column_level + o.n_column_level);
        for (int lv : level) {
          if (lv < 0)
            throw std::invalid_argument("column_level must be non-negative.");
          n_levels = std::max(n_levels, lv + 1);
        }
        int check_levels = 0;
        if (compute_levels(Xth0, &check_levels, level.data()) != level)
          throw std::invalid_argument("column_level is not a valid schedule for this shard.");
      } else {
        if (world > 1)
          throw std::invalid_argument("row-sharded training needs column_level agreed between the ranks.");
        if (!compute_levels_by_rows(Xh, level, &n_levels)) // rows with unsorted columns: serial, by column
          level = compute_levels(Xth0, &n_levels);
      }
      tick("dependency levels");
      perm = primary_row_order(Xth0, level, n_levels, &primary);
      Xh = permute_rows(Xh, perm);
      Xth0 = HostCs<Real>(); // not needed any more (hundreds of MB to GB on the large configurations)
      tick("row order + permute");
      main_unit = std::all_of(Xh.val.begin(), Xh.val.end(), [](Real v) { return v == Real(1); });
      main_row_len = Xh.n_major ? Xh.ptr[1] - Xh.ptr[0] : 0;
      for (int64_t i = 0; i < Xh.n_major && main_row_len > 0; i++)
        if (Xh.ptr[i + 1] - Xh.ptr[i] != main_row_len)
          main_row_len = 0;
      tick("unit / row-length scan");
      HostCs<Real> Xth = host_transpose(Xh);
      tick("transpose (device order)");
      plan = make_sweep_plan(Xth, level, n_levels, SWEEP_WARP_MAX, SWEEP_CHUNK);
      plan.primary_level = primary;
      tick("sweep plan");
      Xt.upload(Xth, stream);
      setup_field_path(&Xh, Xth, level, n_levels, n_rel);
      MYFM_CUDA(cudaStreamSynchronize(stream));
      tick("CSC upload + field path");
      setup_tile_path(Xh);
      tick("tile path");
    }
    if (!device_prepared)
      perm_dev.upload(perm, stream);
    { // key of every device row's latent stream (PHILOX): its global row index
      if (o.row_ids && o.n_row_ids != static_cast<int64_t>(perm.size()))
        throw std::invalid_argument("row_ids must have one entry per training row of this shard.");
      if (cfg.task_type != MYFM_TASK_REGRESSION && philox_latents) { // only the device-side latent draws read it
        std::vector<int> key(perm.size());
        parallel_parts(parts_for(static_cast<int64_t>(perm.size())), [&](int t, int n_parts) {
          auto [i0, i1] = part_range(static_cast<int64_t>(perm.size()), t, n_parts);
          for (int64_t i = i0; i < i1; i++)
            key[i] = static_cast<int>(o.row_ids ? o.row_ids[perm[i]] : o.row_offset + perm[i]);
        });
        latent_row.upload(key, stream);
      }
    }
    items.upload(plan.items, stream);
    seg_count.upload(plan.seg_count, stream);
    if (world > 1) {
      item_slot.upload(plan.item_slot, stream);
      colstat.alloc(2 * static_cast<size_t>(std::max(1, plan.max_level_cols)));
      told.alloc(std::max(1, plan.max_level_cols));
    }
    seg_partial.alloc(2 * static_cast<size_t>(std::max(1, plan.max_seg_items)));
    seg_theta_old.alloc(std::max(1, plan.max_seg_items));

    if (!device_prepared)
      data.build(Xh, n_rel, relations, stream, &perm);
    tick("CSR upload");
    N = data.n_rows, D = data.dim_main, D_all = data.dim_all;
    if (static_cast<int64_t>(cfg.group_index.size()) != D_all)
      throw std::invalid_argument("group_index must have one entry per feature.");
    G = cfg.n_groups;

    rel_train.resize(n_rel);
    for (int b = 0; b < n_rel; b++) {
      const myfm_relation_t &r = relations[b];
      HostCs<Real> Bh = host_from_api<Real>(r.block, "relation block");
      if (has_duplicate_entries(Bh))
        throw std::invalid_argument("a relation block lists the same (row, column) twice.");
      HostCs<Real> Bth = host_transpose(Bh);
      DevRelationTrain<Real> &t = rel_train[b];
      t.Bt.upload(Bth, stream);
      const int64_t S = Bh.n_major;
      // rows grouped by block row, ascending (segment sums replace the reference's serial
      // accumulation, definitions.hpp:64-66 / FMTrainer.hpp:270-274)
      std::vector<int> seg_ptr(S + 1, 0), seg_rows(N);
      std::vector<Real> card(S, Real(0));
      for (int64_t i = 0; i < N; i++) {
        seg_ptr[r.original_to_block[i] + 1]++;
        card[r.original_to_block[i]]++;
      }
      for (int64_t s = 0; s < S; s++)
        seg_ptr[s + 1] += seg_ptr[s];
      std::vector<int> cur(seg_ptr.begin(), seg_ptr.end() - 1);
      for (int64_t i = 0; i < N; i++) // i: device row
        seg_rows[cur[r.original_to_block[perm[i]]]++] = static_cast<int>(i);
      t.seg_ptr.upload(seg_ptr, stream);
      t.seg_rows.upload(seg_rows, stream);
      t.card.upload(card, stream);
      for (DevBuf<Real> *buf : {&t.q, &t.q_S, &t.c, &t.c_S, &t.e, &t.e_q}) {
        buf->alloc(S);
        buf->zero(stream);
      }
      LevelPlan bp = make_level_plan(Bth, std::numeric_limits<int>::max());
      t.n_levels = bp.n_levels;
      t.level_ptr.upload(bp.level_ptr, stream);
      t.level_cols.upload(bp.cols, stream);
      {
        std::vector<int> run_end(std::max(1, bp.n_levels), 0);
        for (int lv = bp.n_levels - 1; lv >= 0; lv--) {
          const bool single = bp.level_ptr[lv + 1] - bp.level_ptr[lv] == 1;
          const bool next_single = lv + 1 < bp.n_levels && bp.level_ptr[lv + 2] - bp.level_ptr[lv + 1] == 1;
          run_end[lv] = (single && next_single) ? run_end[lv + 1] : lv + 1;
        }
        t.run_end.upload(run_end, stream);
        // {column, first entry, end entry, group} of the first column of every level
        std::vector<int4> rec(std::max(1, bp.n_levels), make_int4(0, 0, 0, 0));
        const int64_t off = data.rels[b].offset;
        for (int lv = 0; lv < bp.n_levels; lv++) {
          const int l = bp.cols[bp.level_ptr[lv]];
          rec[lv] = make_int4(l, Bth.ptr[l], Bth.ptr[l + 1], static_cast<int>(cfg.group_index[off + l]));
        }
        t.level_rec.upload(rec, stream);
      }
      MYFM_CUDA(cudaStreamSynchronize(stream));
    }

    const bool host_targets = cfg.task_type != MYFM_TASK_REGRESSION; // labels in the caller's row order: latent draws, cut points
    if (host_targets)
      y_host.resize(N);
    {
      std::vector<Real> y_dev(N);
      parallel_parts(parts_for(N), [&](int t, int n_parts) { // a random gather over N targets: 50 ms on one thread at 10 M rows
        auto [i0, i1] = part_range(N, t, n_parts);
        for (int64_t i = i0; i < i1; i++) {
          if (host_targets)
            y_host[i] = static_cast<Real>(y_api[i]);
          y_dev[i] = static_cast<Real>(y_api[perm[i]]);
        }
      });
      y.upload(y_dev, stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
    }
    eq_buf.alloc(2 * static_cast<size_t>(N) + 2); // one spare pair: the tile path's bulk copies move 16 bytes at a time
    eq_buf.zero(stream);
    dense_tmp.alloc(N);
    if (cfg.task_type == MYFM_TASK_ORDERED) { // BaseFMTrainer.hpp:79-104
      std::vector<char> seen(N, 0);
      for (auto &grp : cfg.cutpoint_groups)
        for (int64_t k : grp.second) {
          if (k < 0 || k >= N)
            throw std::invalid_argument("out of range for cutpoint group config.");
          if (seen[k]) {
            std::ostringstream ss;
            ss << "index " << k << " overlapping in cutpoint config.";
            throw std::invalid_argument(ss.str());
          }
          seen[k] = 1;
        }
      for (int64_t i = 0; i < N; i++)
        if (!seen[i]) {
          std::ostringstream ss;
          ss << "cutpoint group not specified for " << i << ".";
          throw std::invalid_argument(ss.str());
        }
    }
    group.upload(cfg.group_index, stream);
    feat_ptr.upload(cfg.feat_ptr, stream);
    feat_idx.upload(cfg.feat_idx, stream);
    { // chunks of the group hyper-parameter reduction (k_group_hyper)
      std::vector<int> cg, cb;
      for (int g = 0; g < G; g++) {
        const int b = cfg.feat_ptr[g], en = cfg.feat_ptr[g + 1];
        for (int p = b; p < en || p == b; p += HYPER_CHUNK)
          cg.push_back(g), cb.push_back(p);
      }
      hyper_chunks = static_cast<int>(cg.size());
      hyper_chunk_group.upload(cg, stream);
      hyper_chunk_begin.upload(cb, stream);
    }
    partial.alloc(2 * REDUCE_BLOCKS);
    scal.alloc(4);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    tick("targets, groups, buffers");

    N_global = world > 1 ? o.n_rows_global : N;
    if (world > 1) {
      if (N_global < N)
        throw std::invalid_argument("n_rows_global is smaller than this shard.");
      NcclApi &nccl = NcclApi::get();
      ncclUniqueId id;
      std::memcpy(&id, o.nccl_unique_id, sizeof(id));
      // One communicator per (ncclUniqueId, world, rank) and process: trainers made from the same id (the same
      // ShardContext) share it — creating one and running its first collective costs 0.6 s on 2 GPUs and more on
      // 8, several seconds the first time in a process.  Shared communicators live until the process exits.
      {
        std::string key(reinterpret_cast<const char *>(o.nccl_unique_id), sizeof(id));
        key += ":" + std::to_string(world) + ":" + std::to_string(o.rank);
        std::lock_guard<std::mutex> lock(shared_comm_mutex());
        auto &cache = shared_comms();
        auto it = cache.find(key);
        if (it == cache.end()) {
          nccl.check(nccl.CommInitRank(&comm, world, id, o.rank), "ncclCommInitRank");
          cache.emplace(key, comm);
        } else {
          comm = it->second;
        }
      }
      tick("ncclCommInitRank (or the shared communicator)");
      // every rank must take the same schedule: the field path only if every shard qualifies
      DevBuf<int> flag(1);
      const int mine = field_path ? 1 : 0;
      flag.upload(&mine, 1, stream);
      nccl.check(nccl.AllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, comm, stream), "ncclAllReduce");
      int all = 0;
      flag.download(&all, 1, stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
      field_path = all != 0;
      my_rank = o.rank;
      tick("first all-reduce (schedule agreement)");
      if (field_path) {
        setup_exclusive_level0();
        tick("rank-exclusive first field");
        setup_peer_exchange();
        tick("peer-memory exchange (cudaIpc)");
      }
    }
    // Gamma shapes are data independent (FMTrainer.hpp:140,157)
    shape_alpha = (static_cast<Real>(cfg.alpha_0) + N_global) / 2;
    shapes_lw.resize(G);
    for (int g = 0; g < G; g++) {
      Real a = static_cast<Real>(cfg.alpha_0) + static_cast<size_t>(cfg.feat_ptr[g + 1] - cfg.feat_ptr[g]);
      shapes_lw[g] = a / 2;
    }
  }

  ~Trainer() override {
    if (stream)
      cudaStreamSynchronize(stream);
    reset_graphs();
    if (rng_stream)
      cudaStreamSynchronize(rng_stream);
    dump_peer_trace();
    if (!peer_base.empty() || peer_local)
      close_peer_exchange(peer_ok);
    comm = nullptr; // shared with the process (shared_comms): never destroyed here
    for (auto *evs : {z_copied, z_ready, z_free})
      for (int k = 0; k < 2; k++)
        if (evs[k])
          cudaEventDestroy(evs[k]);
    if (rng_stream)
      cudaStreamDestroy(rng_stream);
    if (stream)
      cudaStreamDestroy(stream);
  }

  HyperView<Real> hv() {
    Real *h = hyper.p;
    return HyperView<Real>{h, h + 1, h + 2, h + 2 + G, h + 2 + 2 * G,
                           h + 2 + 2 * G + static_cast<size_t>(G) * K};
  }
  size_t hyper_size() const { return 2 + 2 * static_cast<size_t>(G) + 2 * static_cast<size_t>(G) * K; }

  void launched(int n = 1) { launches += n; }
  Pair<Real> *eq() { return reinterpret_cast<Pair<Real> *>(eq_buf.p); }
  Real *e_ptr() { return eq_buf.p; }     // stride 2
  Real *q_ptr() { return eq_buf.p + 1; } // stride 2

  // create_FM + create_Hyper (BaseFMTrainer.hpp:107-115), initialize_hyper + initialize_e
  // (FMTrainer.hpp:89-119)
  void init_fm(int rank, double init_std) override {
    if (rank < 0)
      throw std::invalid_argument("rank must be non-negative.");
    MYFM_CUDA(cudaSetDevice(device));
    K = rank;
    std::vector<Real> Vh(static_cast<size_t>(D_all) * K), wh(D_all);
    Real w0h = 0;
    rng.init_weights(Vh.data(), Vh.size(), wh.data(), wh.size(), &w0h, static_cast<Real>(init_std));
    V.alloc(Vh.size());
    Vt.alloc(Vh.size());
    hyper_chunk_sums.alloc(2 * static_cast<size_t>(std::max(1, hyper_chunks)) * std::max(1, K));
    hyper_done.alloc(static_cast<size_t>(std::max(1, G)) * std::max(1, K));
    hyper_done.zero(stream);
    V.upload(Vh, stream);
    w.upload(wh, stream);
    if (Vh.size()) {
      k_transpose_V<Real><<<ceil_div(Vh.size(), 256), 256, 0, stream>>>(D_all, K, V.p, Vt.p);
      launched();
    }
    std::vector<Real> hh(hyper_size());
    hh[0] = static_cast<Real>(1); // alpha
    hh[1] = w0h;
    for (int g = 0; g < G; g++)
      hh[2 + g] = static_cast<Real>(0), hh[2 + G + g] = static_cast<Real>(1e-5);
    for (size_t i = 0; i < static_cast<size_t>(G) * K; i++)
      hh[2 + 2 * G + i] = static_cast<Real>(0),
                     hh[2 + 2 * G + static_cast<size_t>(G) * K + i] = static_cast<Real>(1e-5);
    hyper.upload(hh, stream);
    layout = SweepLayout::make(cfg.task_type == MYFM_TASK_REGRESSION, cfg.fit_w0, cfg.fit_linear, G,
                               K, D_all);
    reset_graphs();
    {
      const char *no_graph = std::getenv("MYFM_NO_GRAPH");
      use_graphs = !(no_graph && no_graph[0] == '1');
    }
    if (std::getenv("MYFM_PEER_TRACE") && !phase_trace_buf.p) { // sweep-phase stamps (any number of GPUs)
      phase_trace_buf.alloc(static_cast<size_t>(PHASE_TRACE_RECORDS) * PEER_TRACE_SLOTS + 1);
      phase_trace_buf.zero(stream);
    }
    if (field_path || tile_path)
      f_sched.alloc(2 * (static_cast<size_t>(K) + 2));
    setup_rng();
    const bool ordered = cfg.task_type == MYFM_TASK_ORDERED;
    data.predict(w.p, Vt.p, K, hv().w0, ordered ? nullptr : y.p, e_ptr(), 2);
    if (ordered)
      ordered_update(true);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    sweep_index = 0;
  }

  // Chooses where the sweep variates come from.  Regression consumes the mt19937 stream in a
  // data-independent pattern, so the stream itself moves to the device (mt_device.cuh); the
  // host keeps it for classification (data-dependent truncated normals) and for Gamma shapes
  // below 1 (libstdc++ takes another branch there).  MYFM_HOST_RNG=1 forces the host path.
  void setup_rng() {
    MYFM_CUDA(cudaStreamSynchronize(rng_stream));
    const char *force_host = std::getenv("MYFM_HOST_RNG");
    device_rng = (cfg.task_type == MYFM_TASK_REGRESSION || philox_latents) && shape_alpha >= 1 &&
                 !(force_host && force_host[0] == '1');
    for (Real sh : shapes_lw)
      device_rng = device_rng && sh >= 1;
    gen_index = 0, z_last = nullptr;
    if (!device_rng) {
      z_dev.alloc(layout.total);
      for (auto &pb : z_pinned)
        pb.alloc(layout.total);
      return;
    }
    for (auto &zb : z_slot)
      zb.alloc(layout.total);
    // hand the generator over where create_FM left it: operator<< prints x[0..623] and p
    std::ostringstream os;
    os << rng.gen;
    std::istringstream is(os.str());
    std::vector<MtControl> ctl(1);
    std::memset(&ctl[0], 0, sizeof(MtControl));
    std::vector<uint32_t> first_words(MT_N);
    for (int i = 0; i < MT_N; i++) {
      is >> ctl[0].window[i];
      first_words[i] = mt_temper(ctl[0].window[i]);
    }
    unsigned long long p = 0;
    is >> p;
    ctl[0].produced = MT_N;
    // words one sweep consumes at most: every bulk segment is given 1 % + 4096 attempts more than
    // its expectation (acceptance pi/4; the standard deviation is below 0.05 % beyond 1e6
    // attempts), every scalar draw a budget of 64 words
    constexpr int WPA = MtTraits<Real>::WPA;
    const long long n_scalar = 2 + 2 * static_cast<long long>(G) * (K + 1);
    const long long need = (bulk_attempts(D_all) + bulk_attempts(static_cast<long long>(K) * D_all)) * WPA +
                           64 * n_scalar + MT_SCALAR_WINDOW;
    mt_ahead = static_cast<unsigned long long>(need);
    // Large sweeps: the word stream comes from MT_FARM_LANES lanes at once (k_mt_farm), one block of
    // lanes x seg words per sweep; small ones keep the single serial CTA (its 100 ns per 227 words
    // are hidden behind the sweep anyway).  MYFM_MT_FARM="lanes,seg" forces a shape (tests).
    mt_farm_lanes = 0, mt_farm_seg = 0;
    {
      const char *env = std::getenv("MYFM_MT_FARM");
      long long lanes = 0, seg = 0;
      if (env && std::sscanf(env, "%lld,%lld", &lanes, &seg) == 2) {
        if (lanes > 1 && seg > MT_JUMP_SPAN)
          mt_farm_lanes = static_cast<int>(lanes), mt_farm_seg = static_cast<unsigned long long>(seg);
      } else if (need >= (1ll << 20)) {
        mt_farm_lanes = MT_FARM_LANES;
        mt_farm_seg = static_cast<unsigned long long>(ceil_div(ceil_div(need, MT_FARM_LANES), MT_LAG)) * MT_LAG;
      }
    }
    const unsigned long long farm_block = mt_farm_seg * static_cast<unsigned long long>(mt_farm_lanes);
    unsigned long long cap = 1 << 16;
    while (cap < 2 * mt_ahead + 4 * MT_N || cap < 3 * farm_block + mt_ahead + 4 * MT_N)
      cap <<= 1;
    mt_mask = cap - 1;
    mt_ring.alloc(cap);
    mt_ring.upload(first_words.data(), MT_N, stream);
    if (mt_farm_lanes) {
      const std::vector<uint16_t> &taps =
          mt_jump_taps(farm_block - mt_farm_seg + static_cast<unsigned long long>(MT_JUMP_SPAN));
      mt_farm_taps.upload(taps, stream);
      mt_farm_n_taps = static_cast<int>(taps.size());
      std::vector<unsigned long long> blocks(mt_farm_lanes, 1ull); // block 0 comes from k_mt_generate
      mt_farm_blocks.upload(blocks, stream);
    }
    // stages of a sweep: head scalars, w bulk, V-hyper scalars, V bulk (absent ones are skipped)
    mt_final_stage = 1 + (layout.z_w >= 0 && D_all > 0 ? 1 : 0) + (G > 0 && K > 0 ? 1 : 0) +
                     (K > 0 && D_all > 0 ? 1 : 0);
    ctl[0].pos[mt_final_stage] = p;
    mt_ctl.upload(ctl, stream);
    const long long max_tiles = ceil_div(bulk_attempts(static_cast<long long>(std::max(K, 1)) * D_all), MT_BULK_TILE);
    mt_tile_count.alloc(max_tiles + 1);
    mt_tile_offset.alloc(max_tiles + 2);
    // gamma_distribution::param_type::_M_initialize (bits/random.tcc:2338-2345), shape >= 1
    std::vector<Real> consts(2 * (G + 1));
    for (int g = 0; g <= G; g++) {
      const Real malpha = g < G ? shapes_lw[g] : shape_alpha;
      const Real a1 = malpha - Real(1.0) / Real(3.0);
      consts[g] = a1;
      consts[G + 1 + g] = Real(1.0) / std::sqrt(Real(9.0) * a1);
    }
    mt_consts.upload(consts, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    if (mt_farm_lanes) { // block 0 of the farm: the serial generator, once
      k_mt_generate<<<1, MT_GEN_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, farm_block - p,
                                                           mt_final_stage);
      launched();
      const size_t smem = mt_farm_n_taps * sizeof(uint16_t);
      MYFM_CUDA(cudaFuncSetAttribute(k_mt_farm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(std::max<size_t>(smem, 48 * 1024))));
      MYFM_CUDA(cudaGetLastError());
    }
  }
  static constexpr int MT_FARM_LANES = 32;
  int mt_farm_lanes = 0;                  // 0: serial generator (k_mt_generate) every sweep
  unsigned long long mt_farm_seg = 0;     // words per lane per block
  DevBuf<uint16_t> mt_farm_taps;
  int mt_farm_n_taps = 0;
  DevBuf<unsigned long long> mt_farm_blocks;

  // polar attempts examined for `count` bulk normals (mt_device.cuh)
  static long long bulk_attempts(long long count) {
    if (count <= 0)
      return 0;
    const long long a = static_cast<long long>(std::ceil(count * 1.2733 * 1.01)) + 4096;
    return (a + MT_BULK_TILE - 1) / MT_BULK_TILE * MT_BULK_TILE;
  }

  void launch_bulk(int &stage, long long count, Real *out, long long out0) {
    if (count <= 0)
      return;
    const long long n_attempts = bulk_attempts(count);
    const int n_tiles = static_cast<int>(n_attempts / MT_BULK_TILE);
    k_mt_bulk_count<Real><<<n_tiles, MT_BULK_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, stage, n_attempts,
                                                                        mt_tile_count.p);
    k_mt_bulk_scan<<<1, 1024, 0, rng_stream>>>(n_tiles, mt_tile_count.p, mt_tile_offset.p);
    k_mt_bulk_emit<Real><<<n_tiles, MT_BULK_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, stage, stage + 1,
                                                                       n_attempts, count, mt_tile_offset.p, n_tiles, out,
                                                                       out0);
    launched(3);
    stage++;
  }

  void launch_variates(int64_t index) {
    const int slot = static_cast<int>(index & 1);
    if (index >= 2)
      MYFM_CUDA(cudaStreamWaitEvent(rng_stream, z_free[slot], 0));
    const SweepLayout &L = layout;
    Real *out = z_slot[slot].p;
    if (mt_farm_lanes) {
      MtFarm farm;
      farm.lanes = mt_farm_lanes, farm.seg = mt_farm_seg;
      farm.n_taps = mt_farm_n_taps, farm.taps = mt_farm_taps.p;
      farm.lane_blocks = mt_farm_blocks.p;
      k_mt_farm<<<mt_farm_lanes, MT_FARM_THREADS, mt_farm_n_taps * sizeof(uint16_t), rng_stream>>>(
          mt_ctl.p, mt_ring.p, mt_mask, mt_ahead, mt_final_stage, farm);
    } else {
      k_mt_generate<<<1, MT_GEN_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, mt_ahead, mt_final_stage);
    }
    launched();
    int stage = 0;
    { // alpha, w0, lambda_w, mu_w
      MtScalarProgram prog;
      prog.a1 = mt_consts.p, prog.a2 = mt_consts.p + (G + 1);
      int n = 0;
      if (L.g_alpha >= 0)
        prog.seg[n++] = MtScalarSegment{1, 1, L.g_alpha, G, 0};
      if (L.z_w0 >= 0)
        prog.seg[n++] = MtScalarSegment{0, 1, L.z_w0, 0, 0};
      if (G > 0) {
        prog.seg[n++] = MtScalarSegment{1, G, L.g_lw, 0, G};
        prog.seg[n++] = MtScalarSegment{0, G, L.z_mw, 0, 0};
      }
      prog.n_segments = n;
      k_mt_scalars<Real><<<1, MT_SCALAR_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, stage, stage + 1, prog,
                                                                   out);
      launched();
      stage++;
    }
    if (L.z_w >= 0)
      launch_bulk(stage, D_all, out, L.z_w);
    if (G > 0 && K > 0) { // lambda_V (factor-major, group-minor), then mu_V
      MtScalarProgram prog;
      prog.a1 = mt_consts.p, prog.a2 = mt_consts.p + (G + 1);
      prog.seg[0] = MtScalarSegment{1, K * G, L.g_lV, 0, G};
      prog.seg[1] = MtScalarSegment{0, K * G, L.z_mV, 0, 0};
      prog.n_segments = 2;
      k_mt_scalars<Real><<<1, MT_SCALAR_THREADS, 0, rng_stream>>>(mt_ctl.p, mt_ring.p, mt_mask, stage, stage + 1, prog,
                                                                   out);
      launched();
      stage++;
    }
    launch_bulk(stage, static_cast<long long>(K) * D_all, out, L.z_V);
    if (stage != mt_final_stage)
      throw std::logic_error("device RNG: stage count mismatch.");
    MYFM_CUDA(cudaGetLastError());
    MYFM_CUDA(cudaEventRecord(z_ready[slot], rng_stream));
  }

  // Device-side error flags (the word ring of the mt19937 stream ran dry; a peer rank never published
  // its column statistics) are copied into pinned memory together with whatever the caller fetches
  // next and examined after that fetch's synchronisation: sync() and every get_hyper() — i.e. once
  // per sweep of create_train_fm — see them at no extra synchronisation.
  PinnedBuf<int> err_host;
  void queue_error_flags() {
    if (!err_host.p) {
      err_host.alloc(2);
      err_host.p[0] = err_host.p[1] = 0;
    }
    if (peer_ok)
      MYFM_CUDA(cudaMemcpyAsync(err_host.p, peer_error.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (device_rng)
      MYFM_CUDA(cudaMemcpyAsync(err_host.p + 1, &mt_ctl.p->error, sizeof(int), cudaMemcpyDeviceToHost, stream));
  }
  void throw_on_error_flags() { // the stream has been synchronised since queue_error_flags()
    if (!err_host.p)
      return;
    if (err_host.p[0])
      throw std::runtime_error("row-sharded training: a peer rank never published its column statistics.");
    if (err_host.p[1])
      throw std::runtime_error("device mt19937 stream: the word ring ran dry or a bulk segment found too few accepted "
                               "attempts (set MYFM_HOST_RNG=1).");
  }
  void check_rng_error() {
    queue_error_flags();
    MYFM_CUDA(cudaStreamSynchronize(stream));
    throw_on_error_flags();
  }

  // Standardised variates of one sweep, in the reference's consumption order.
  void draw_sweep_variates(Real *out) {
    const SweepLayout &L = layout;
    if (L.g_alpha >= 0)
      out[L.g_alpha] = rng.gamma(shape_alpha);
    if (L.z_w0 >= 0)
      out[L.z_w0] = rng.normal();
    for (int g = 0; g < G; g++)
      out[L.g_lw + g] = rng.gamma(shapes_lw[g]);
    for (int g = 0; g < G; g++)
      out[L.z_mw + g] = rng.normal();
    if (L.z_w >= 0)
      for (int64_t j = 0; j < D_all; j++)
        out[L.z_w + j] = rng.normal();
    for (int r = 0; r < K; r++)
      for (int g = 0; g < G; g++)
        out[L.g_lV + static_cast<int64_t>(r) * G + g] = rng.gamma(shapes_lw[g]);
    for (int64_t i = 0; i < static_cast<int64_t>(K) * G; i++)
      out[L.z_mV + i] = rng.normal();
    for (int64_t i = 0; i < static_cast<int64_t>(K) * D_all; i++)
      out[L.z_V + i] = rng.normal();
  }

  struct TimedSpan {
    KernelTimer &t;
    cudaStream_t s;
    int family;
    size_t a = 0;
    TimedSpan(KernelTimer &t_, cudaStream_t s_, int f) : t(t_), s(s_), family(f) {
      if (t.enabled)
        a = t.mark(s);
    }
    ~TimedSpan() {
      if (t.enabled) {
        size_t b = t.mark(s);
        t.spans.push_back({family, a, b});
      }
    }
  };

  void allreduce_sum(Real *buf, size_t n) {
    NcclApi &nccl = NcclApi::get();
    nccl.check(nccl.AllReduce(buf, buf, n, sizeof(Real) == 4 ? ncclFloat : ncclDouble, ncclSum, comm, stream),
               "ncclAllReduce");
  }

  // One level-ordered sweep over the main-table columns (w or one factor of V).
  template <bool IS_V>
  void sweep_main(Real *theta, Real *theta_t, int64_t t_stride, const Real *z, const Real *lambda,
                  const Real *mu, size_t first_level = 0, size_t end_level = static_cast<size_t>(-1)) {
    SweepArgs<Real> a;
    a.idx = Xt.idx.p, a.val = Xt.val.p;
    a.eq = eq();
    a.theta = theta, a.theta_t = theta_t, a.t_stride = t_stride;
    a.z = z, a.group = group.p;
    a.alpha = hv().alpha, a.lambda = lambda, a.mu = mu;
    a.partial = seg_partial.p, a.theta_old_buf = seg_theta_old.p;
    for (size_t li = first_level; li < std::min(end_level, plan.levels.size()); li++) {
      const SweepLevel &L = plan.levels[li];
      a.item = reinterpret_cast<const int4 *>(items.p + L.s0), a.seg_count = seg_count.p + L.s0;
      a.nS = L.c0 - L.s0, a.nC = L.w0 - L.c0, a.nW = L.end - L.w0;
      const int grid = a.nS + a.nC + ceil_div(a.nW, SWEEP_WARPS);
      if (!grid)
        continue;
      TimedSpan span(timer, stream, 0);
      if (world > 1) { // statistics -> sum over ranks -> identical draw everywhere -> local update
        a.item_slot = item_slot.p + L.s0, a.colstat = colstat.p, a.told = told.p;
        const int n_cols = L.end - L.s0 - (a.nS ? count_chunk_surplus(L) : 0);
#define MYFM_DIST(U, C, UPD) k_level_dist<Real, IS_V, U, C, UPD><<<grid, SWEEP_THREADS, 0, stream>>>(a)
#define MYFM_DIST_PHASE(UPD)                                                                       \
  if (L.unit && L.contig)                                                                          \
    MYFM_DIST(true, true, UPD);                                                                    \
  else if (L.unit)                                                                                 \
    MYFM_DIST(true, false, UPD);                                                                   \
  else if (L.contig)                                                                               \
    MYFM_DIST(false, true, UPD);                                                                   \
  else                                                                                             \
    MYFM_DIST(false, false, UPD);
        MYFM_DIST_PHASE(false)
        if (a.nS) {
          k_level_chunk_fold<Real><<<ceil_div(a.nS, 128), 128, 0, stream>>>(a);
          launched();
        }
        allreduce_sum(colstat.p, 2 * static_cast<size_t>(n_cols));
        MYFM_DIST_PHASE(true)
#undef MYFM_DIST_PHASE
#undef MYFM_DIST
        launched(2);
        continue;
      }
#define MYFM_LEVEL(U, C)                                                                           \
  {                                                                                                \
    k_level_sweep<Real, IS_V, U, C><<<grid, SWEEP_THREADS, 0, stream>>>(a);                        \
    if (a.nS)                                                                                      \
      k_level_seg_update<Real, IS_V, U, C><<<a.nS, SWEEP_THREADS, 0, stream>>>(a);                 \
  }
      if (L.unit && L.contig)
        MYFM_LEVEL(true, true)
      else if (L.unit)
        MYFM_LEVEL(true, false)
      else if (L.contig)
        MYFM_LEVEL(false, true)
      else
        MYFM_LEVEL(false, false)
#undef MYFM_LEVEL
      launched(a.nS ? 2 : 1);
    }
  }

  // ---- field path (field_sweep.cuh) ---------------------------------------------------------------
  // Decides whether the main table is a stack of position-aligned fields and, if so, builds the
  // extra arrays of that path.  Xh: CSR in device row order, Xth its transpose.
  // Field-shaped main tables: row order, permuted CSR, CSC and the field arrays are built on the
  // device (prep_device.cuh); the host checks the row pointers, plans the work items from the
  // column pointers and keeps the permutation.  Returns false (nothing changed) when the input is
  // not of that shape — the host path then prepares it, or reports what is wrong with it.
  template <typename Tick>
  bool prepare_on_device(const myfm_csr_t &X, int n_rel, const myfm_engine_options_t &o, Tick &tick) {
    const char *off = std::getenv("MYFM_HOST_SETUP"), *tile = std::getenv("MYFM_TILE_PATH");
    if ((off && off[0] == '1') || (tile && tile[0] == '1') || n_rel > 0)
      return false;
    const int64_t n = X.n_rows, n_cols = X.n_cols;
    if (n < 1024 || n_cols <= 0 || n_cols >= std::numeric_limits<int>::max() || !X.indptr || !X.indices || !X.data)
      return false;
    const int64_t nnz = X.indptr[n];
    if (X.indptr[0] != 0 || nnz <= 0 || nnz >= std::numeric_limits<int>::max() || nnz % n)
      return false;
    const int L = static_cast<int>(nnz / n);
    if (L < 2 || L > PREP_MAX_FIELDS)
      return false;
    if (o.column_level && o.n_column_level != n_cols)
      return false;
    // every row holds exactly L entries; are all values 1?
    std::atomic<bool> regular{true}, unit{true};
    parallel_parts(parts_for(nnz), [&](int t, int n_parts) {
      auto [r0, r1] = part_range(n, t, n_parts);
      for (int64_t i = r0; i < r1; i++)
        if (X.indptr[i + 1] != (i + 1) * L) {
          regular = false;
          return;
        }
      auto [p0, p1] = part_range(nnz, t, n_parts);
      bool ones = true;
      for (int64_t p = p0; p < p1 && ones; p++)
        ones = X.data[p] == 1.0;
      if (!ones)
        unit = false;
    });
    if (!regular)
      return false;
    tick("row pointers / unit values (host)");
    DevBuf<int> idx_in(nnz), given_level;
    DevBuf<double> val_in;
    idx_in.upload(X.indices, nnz, stream);
    if (!unit)
      val_in.upload(X.data, nnz, stream);
    if (o.column_level)
      given_level.upload(o.column_level, n_cols, stream);
    std::vector<int> stat(129, 0);
    for (int k = 0; k < 64; k++)
      stat[k] = std::numeric_limits<int>::max(), stat[64 + k] = -1;
    DevBuf<int> stat_dev;
    stat_dev.upload(stat, stream);
    k_prep_scan<<<592, 256, 0, stream>>>(n, L, static_cast<int>(n_cols), idx_in.p, given_level.p, stat_dev.p);
    launched();
    stat_dev.download(stat.data(), stat.size(), stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    tick("upload + field scan");
    if (stat[128])
      return false;
    for (int k = 0; k + 1 < L; k++)
      if (stat[64 + k] >= stat[k + 1])
        return false; // the fields' column ranges must ascend with the position
    {
      int dev_smem = 0;
      MYFM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
      const int64_t tab = static_cast<int64_t>(stat[64 + L - 1]) - stat[L - 1] + 1;
      if (tab * FieldTab<Real>::BYTES_PER_COLUMN > dev_smem - 4096 && L == 2 && world == 1)
        return false; // two fields, table too large for the field path: the tile path (host-built B order) takes it
    }
    // rows by their first-field column
    int bits = 1;
    while ((int64_t(1) << bits) < n_cols)
      bits++;
    PrepSorter sorter;
    DevBuf<int> keys(n), keys_sorted(n), rows(n);
    perm_dev.alloc(n);
    const int grid_n = ceil_div(n, 256);
    k_prep_keys<<<grid_n, 256, 0, stream>>>(n, L, 0, idx_in.p, keys.p, rows.p);
    sorter.sort(keys.p, keys_sorted.p, rows.p, perm_dev.p, n, bits, stream);
    // permuted CSR, field arrays, column counts
    data.stream = stream;
    data.n_rows = n, data.dim_main = n_cols, data.dim_all = n_cols;
    data.X.n_major = n, data.X.n_minor = n_cols, data.X.nnz = nnz;
    data.X.ptr.alloc(n + 1), data.X.idx.alloc(nnz), data.X.val.alloc(nnz);
    data.row_len = L, data.unit = unit;
    main_unit = unit, main_row_len = L;
    f_tail_idx.alloc(static_cast<size_t>(L - 1) * n + 4); // (+4: bulk copies of the last field's indices end on 16 bytes)
    if (!unit)
      f_tail_val.alloc(static_cast<size_t>(L - 1) * n), f_own_val.alloc(n);
    DevBuf<int> col_count(n_cols + 1);
    col_count.zero(stream);
    k_prep_permute<Real><<<ceil_div(n + 1, 256), 256, 0, stream>>>(
        n, L, perm_dev.p, idx_in.p, unit ? nullptr : val_in.p, data.X.ptr.p, data.X.idx.p, data.X.val.p, f_tail_idx.p,
        unit ? nullptr : f_tail_val.p, unit ? nullptr : f_own_val.p, col_count.p);
    // CSC: column pointers, then the entries field by field
    Xt.n_major = n_cols, Xt.n_minor = n, Xt.nnz = nnz;
    Xt.ptr.alloc(n_cols + 1), Xt.idx.alloc(nnz), Xt.val.alloc(nnz);
    sorter.exclusive_sum(col_count.p, Xt.ptr.p, n_cols + 1, stream);
    k_prep_csc_first<Real><<<grid_n, 256, 0, stream>>>(n, unit ? nullptr : f_own_val.p, Xt.idx.p, Xt.val.p);
    launched(4);
    for (int k = 1; k < L; k++) {
      const int *field_cols = f_tail_idx.p + static_cast<size_t>(k - 1) * n;
      k_prep_keys<<<grid_n, 256, 0, stream>>>(n, 1, 0, field_cols, keys.p, rows.p);
      sorter.sort(keys.p, keys_sorted.p, rows.p, Xt.idx.p + static_cast<size_t>(k) * n, n, bits, stream);
      k_prep_csc_val<Real><<<grid_n, 256, 0, stream>>>(
          n, Xt.idx.p + static_cast<size_t>(k) * n, unit ? nullptr : f_tail_val.p + static_cast<size_t>(k - 1) * n,
          Xt.val.p + static_cast<size_t>(k) * n);
      launched(3);
    }
    MYFM_CUDA(cudaGetLastError());
    // to the host: the permutation and the column pointers
    perm.resize(n);
    HostCs<Real> Xth; // column pointers only
    Xth.n_major = n_cols, Xth.n_minor = n;
    Xth.ptr.resize(n_cols + 1);
    perm_dev.download(perm.data(), n, stream);
    Xt.ptr.download(Xth.ptr.data(), n_cols + 1, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    tick("device: sorts, CSR, CSC");
    std::vector<int> level;
    const int n_levels = L;
    if (o.column_level) {
      level.assign(o.column_level, o.column_level + n_cols);
      for (int lv : level)
        if (lv < 0 || lv >= L)
          return false; // (the host path reports what is wrong)
    } else {
      level.assign(n_cols, 0); // columns without rows conflict with nothing: level 0, as compute_levels has it
      for (int k = 1; k < L; k++)
        for (int j = stat[k]; j <= stat[64 + k]; j++)
          if (Xth.ptr[j + 1] > Xth.ptr[j])
            level[j] = k;
    }
    LevelFlags flags;
    flags.unit.assign(L, unit ? 1 : 0), flags.contig.assign(L, 0);
    flags.contig[0] = 1;
    plan = make_sweep_plan(Xth, level, n_levels, SWEEP_WARP_MAX, SWEEP_CHUNK, -1, &flags);
    plan.primary_level = 0;
    setup_field_path(nullptr, Xth, level, n_levels, 0, &flags);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    tick("work items + field path");
    return true;
  }

  // Xh: CSR in device row order, or nullptr when the table was prepared on the device (then the
  // field structure is established, the field arrays exist already, Xth carries column pointers
  // only and a first-field column's rows are [ptr[j], ptr[j + 1]) themselves).
  void setup_field_path(const HostCs<Real> *Xh, const HostCs<Real> &Xth, const std::vector<int> &level,
                        int n_levels, int n_rel, const LevelFlags *known = nullptr) {
    field_path = false, f_structure = false;
    const char *off = std::getenv("MYFM_NO_FIELD_PATH");
    if (off && off[0] == '1')
      return;
    const int L = main_row_len;
    const int64_t n = Xth.n_minor;
    if (n_rel > 0 || n == 0 || L < 2 || L != n_levels || plan.primary_level != 0 || !plan.levels[0].contig)
      return;
    if (Xh)
      for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < L; k++)
          if (level[Xh->idx[i * L + k]] != k)
            return;
    int lo = std::numeric_limits<int>::max(), hi = -1, longest0 = 0;
    for (int64_t j = 0; j < Xth.n_major; j++) {
      if (level[j] == L - 1)
        lo = std::min<int>(lo, j), hi = std::max<int>(hi, j);
      if (level[j] == 0)
        longest0 = std::max(longest0, Xth.ptr[j + 1] - Xth.ptr[j]);
    }
    if (hi < 0 || longest0 > FIELD_CTA_MAX)
      return;
    int dev_smem = 0;
    MYFM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    MYFM_CUDA(cudaDeviceGetAttribute(&f_sm_count, cudaDevAttrMultiProcessorCount, device));
    const int64_t tab = static_cast<int64_t>(hi) - lo + 1;
    // the field STRUCTURE is established here; the field path's kernels also need the last field's
    // {theta_old, theta_new, theta_next} table in shared memory (the tile path only theta_next)
    const bool table_fits = tab * FieldTab<Real>::BYTES_PER_COLUMN <= dev_smem - 4096;
    f_last_base = lo, f_tab = static_cast<int>(tab), f_tail = L - 1;
    {
      // Opt-in (MYFM_STAGING=1): measured equal to the direct-load kernel on ml10m (1.90 vs 1.88 ms per sweep
      // for the streaming level: one column of prefetch does not cover the DRAM latency, and there is no room
      // for deeper buffers next to the table), and the larger shared-memory carve-out slows the kernels
      // that follow, so it is not the default.
      const char *staging = std::getenv("MYFM_STAGING");
      f_staged_smem = (static_cast<size_t>(tab) * FieldTab<Real, true>::BYTES_PER_COLUMN + 15) / 16 * 16 +
                      static_cast<size_t>(FIELD_WARPS) * FieldStageSize<Real>::BYTES;
      // (the last field's index array must start on a 16-byte boundary: it follows L - 2 arrays of n ints)
      f_staged = sizeof(Real) == 4 && f_staged_smem + 2048 <= static_cast<size_t>(dev_smem) &&
                 (static_cast<int64_t>(L - 2) * n) % 4 == 0 && staging && staging[0] == '1';
    }

    SweepPlan p0 = make_sweep_plan(Xth, level, n_levels, FIELD_CTA_MAX, FIELD_CTA_MAX, 0, known); // longest first
    SweepPlan pL = make_sweep_plan(Xth, level, n_levels, STATS_WARP_MAX, STATS_CHUNK, L - 1, known);
    f_level0 = p0.levels[0], f_levelL = pL.levels[L - 1];
    // level-0 items carry row ranges: a contiguous column's rows are idx[lo] .. idx[lo] + len
    const int rows_per_warp = 32 * (sizeof(Real) == 8 ? 4 : 8);
    f_nCC = f_nCR = f_nG = f_nW = 0;
    for (SweepItem &it : p0.items) {
      const int len = it.hi - it.lo;
      const int first_row = len ? (Xh ? Xth.idx[it.lo] : it.lo) : 0;
      it.lo = first_row, it.hi = first_row + len;
      if (len > rows_per_warp * FIELD_WARPS)
        f_nCC++;
      else if (len > rows_per_warp * FIELD_GROUP_WARPS)
        f_nCR++;
      else if (len > rows_per_warp)
        f_nG++;
      else
        f_nW++;
    }
    // Inside the group and the warp class the columns are taken in MEMORY order (by first row), not
    // by length: neighbouring warps then stream neighbouring rows, and DRAM pages are used whole
    // (length order makes every column a random 1-2 KB access).  The classes themselves stay
    // longest first, and the warp class is scheduled dynamically, so balance does not suffer.
    {
      std::vector<size_t> order(p0.items.size());
      for (size_t k = 0; k < order.size(); k++)
        order[k] = k;
      const size_t g0 = static_cast<size_t>(f_nCC + f_nCR), w0 = g0 + f_nG;
      auto by_row = [&](size_t x, size_t y) { return p0.items[x].lo < p0.items[y].lo; };
      const char *keep = std::getenv("MYFM_FIELD_LENGTH_ORDER");
      if (!(keep && keep[0] == '1')) {
        std::sort(order.begin() + g0, order.begin() + w0, by_row);
        std::sort(order.begin() + w0, order.end(), by_row);
      }
      std::vector<SweepItem> items(order.size());
      std::vector<int> slots(order.size());
      for (size_t k = 0; k < order.size(); k++)
        items[k] = p0.items[order[k]], slots[k] = p0.item_slot[order[k]];
      p0.items.swap(items), p0.item_slot.swap(slots);
    }
    f_items0_host = p0.items, f_slot0_host = p0.item_slot;
    f_level_host = level;
    f_items0.upload(p0.items, stream);
    f_itemsL.upload(pL.items, stream);
    f_seg_countL.upload(pL.seg_count, stream);
    f_partial.alloc(2 * static_cast<size_t>(std::max(1, f_levelL.c0 - f_levelL.s0)));
    f_chunk_done.alloc(std::max(1, f_levelL.c0 - f_levelL.s0));
    f_chunk_done.zero(stream);
    std::vector<int> tail;
    std::vector<Real> tv, ov;
    if (Xh) {
      tail.resize(static_cast<size_t>(f_tail) * n);
      for (int64_t i = 0; i < n; i++)
        for (int k = 1; k < L; k++)
          tail[static_cast<size_t>(k - 1) * n + i] = Xh->idx[i * L + k];
      tail.resize(tail.size() + 4, 0); // bulk copies of the last field's indices end on 16-byte boundaries
      f_tail_idx.upload(tail, stream);
      if (!main_unit) {
        tv.resize(static_cast<size_t>(f_tail) * n), ov.resize(n);
        for (int64_t i = 0; i < n; i++) {
          ov[i] = Xh->val[i * L];
          for (int k = 1; k < L; k++)
            tv[static_cast<size_t>(k - 1) * n + i] = Xh->val[i * L + k];
        }
        f_tail_val.upload(tv, stream);
        f_own_val.upload(ov, stream);
      }
    }
    if (world > 1) { // column slots and the statistics buffer of the two-pass schedule
      f_item_slot0.upload(p0.item_slot, stream);
      f_item_slotL.upload(pL.item_slot, stream);
      std::vector<int> colsL;
      f_ncols0 = 0;
      for (int64_t j = 0; j < Xth.n_major; j++) {
        if (level[j] == 0)
          f_ncols0++;
        if (level[j] == L - 1)
          colsL.push_back(static_cast<int>(j));
      }
      f_ncolsL = static_cast<int>(colsL.size());
      f_colsL.upload(colsL, stream);
      f_colstat.alloc(2 * static_cast<size_t>(std::max(f_ncols0, f_ncolsL)));
    }
    f_pend_told.alloc(f_tab);
    f_pend_tnew.alloc(f_tab);
    f_pend_told.zero(stream);
    f_pend_tnew.zero(stream);
    MYFM_CUDA(cudaStreamSynchronize(stream)); // host staging vectors die here
    f_structure = true;
    field_path = table_fits;
  }
  bool f_structure = false; // the main table is a stack of position-aligned fields (field / tile path arrays exist)

  // Collective: decides whether every level-0 column lives on one rank only and, if so, restricts
  // this rank's level-0 work items to the columns it owns (columns without rows anywhere are dealt
  // round-robin).
  void setup_exclusive_level0() {
    f_exclusive = false;
    const char *off = std::getenv("MYFM_NO_EXCLUSIVE");
    NcclApi &nccl = NcclApi::get();
    const size_t n0 = f_items0_host.size();
    // per level-0 column slot: (1 << 16 | rank + 1) when this rank holds rows of it
    std::vector<int> mine(std::max<size_t>(1, n0), 0);
    for (size_t k = 0; k < n0; k++)
      if (f_items0_host[k].hi > f_items0_host[k].lo)
        mine[f_slot0_host[k]] = (1 << 16) | (my_rank + 1);
    if (off && off[0] == '1')
      mine[0] = (2 << 16); // vetoes the mode on every rank
    DevBuf<int> buf(mine.size());
    buf.upload(mine, stream);
    nccl.check(nccl.AllReduce(buf.p, buf.p, mine.size(), ncclInt, ncclSum, comm, stream), "ncclAllReduce");
    std::vector<int> all(mine.size());
    buf.download(all.data(), all.size(), stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    for (size_t sl = 0; sl < n0; sl++)
      if ((all[sl] >> 16) > 1)
        return; // some column is split between ranks: the two-pass schedule handles that
    std::vector<int> owner(std::max<int64_t>(1, D_all_main()), 0); // everything else: rank 0's copy counts
    std::vector<SweepItem> kept;
    const int rows_per_warp = 32 * (sizeof(Real) == 8 ? 4 : 8);
    f_nCC = f_nCR = f_nG = f_nW = 0;
    for (size_t k = 0; k < n0; k++) { // items are sorted longest first: filtering keeps the class order
      const int sl = f_slot0_host[k];
      const int who = (all[sl] >> 16) ? (all[sl] & 0xffff) - 1 : sl % world;
      owner[f_items0_host[k].col] = who;
      if (who != my_rank)
        continue;
      kept.push_back(f_items0_host[k]);
      const int len = f_items0_host[k].hi - f_items0_host[k].lo;
      if (len > rows_per_warp * FIELD_WARPS)
        f_nCC++;
      else if (len > rows_per_warp * FIELD_GROUP_WARPS)
        f_nCR++;
      else if (len > rows_per_warp)
        f_nG++;
      else
        f_nW++;
    }
    if (kept.empty())
      kept.push_back(SweepItem{0, 0, 0, 0}); // never read (all counts are zero)
    f_items0.upload(kept, stream);
    owner.resize(std::max<int64_t>(1, D_all), 0);
    f_owner.upload(owner, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    f_exclusive = true;
  }
  int64_t D_all_main() const { return static_cast<int64_t>(f_level_host.size()); }

  // Exclusive level 0, end of the column sweeps: every rank keeps only the entries it owns, the sum
  // over the ranks (zeros elsewhere: exact) is the complete, identical sample on every rank.
  void merge_owned() {
    if (!f_exclusive)
      return;
    const int64_t n_all = D_all * static_cast<int64_t>(K + 1);
    if (merge_pack.n < static_cast<size_t>(n_all))
      merge_pack.alloc(static_cast<size_t>(n_all));
    k_merge_pack<Real><<<ceil_div(n_all, 256), 256, 0, stream>>>(D_all, K, f_owner.p, my_rank, w.p, V.p, merge_pack.p);
    allreduce_sum(merge_pack.p, static_cast<size_t>(n_all));
    k_merge_unpack<Real><<<ceil_div(std::max<int64_t>(D_all, D_all * K), 256), 256, 0, stream>>>(D_all, K, merge_pack.p,
                                                                                                   w.p, V.p, Vt.p);
    launched(3);
  }
  DevBuf<Real> merge_pack; // [w | V] of the owned columns, summed over the ranks in place

  // Maps every rank's statistics buffer into this process (cudaIpc over the NVLink / NVSwitch
  // fabric of one node).  Collective; any failure on any rank leaves every rank on NCCL.
  void setup_peer_exchange() {
    peer_ok = false;
    const char *off = std::getenv("MYFM_NO_PEER");
    NcclApi &nccl = NcclApi::get();
    int ok = !(off && off[0] == '1') && world <= PEER_MAX_RANKS;
    peer_stat_elems = 2 * static_cast<size_t>(std::max(f_ncols0, f_ncolsL));
    const size_t bytes = PEER_HEADER_BYTES + 2 * peer_stat_elems * sizeof(Real);
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok) {
      if (cudaMalloc(&peer_local, bytes) != cudaSuccess || cudaMemset(peer_local, 0, bytes) != cudaSuccess ||
          cudaIpcGetMemHandle(&mine, peer_local) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
    }
    // handles of all ranks, and whether every rank got this far
    DevBuf<unsigned char> send(sizeof(mine) + 4), recv((sizeof(mine) + 4) * static_cast<size_t>(world));
    std::vector<unsigned char> h(sizeof(mine) + 4);
    std::memcpy(h.data(), &mine, sizeof(mine));
    std::memcpy(h.data() + sizeof(mine), &ok, 4);
    send.upload(h, stream);
    nccl.check(nccl.AllGather(send.p, recv.p, h.size(), ncclChar, comm, stream), "ncclAllGather");
    std::vector<unsigned char> all(h.size() * world);
    recv.download(all.data(), all.size(), stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    for (int r = 0; r < world; r++) {
      int theirs = 0;
      std::memcpy(&theirs, all.data() + r * h.size() + sizeof(mine), 4);
      ok = ok && theirs;
    }
    peer_base.assign(world, nullptr);
    if (ok) {
      for (int r = 0; r < world && ok; r++) {
        if (r == my_rank) {
          peer_base[r] = peer_local;
          continue;
        }
        cudaIpcMemHandle_t hd;
        std::memcpy(&hd, all.data() + r * h.size(), sizeof(hd));
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
        }
        peer_base[r] = static_cast<unsigned char *>(p);
      }
    }
    // second agreement: every mapping succeeded everywhere
    DevBuf<int> flag(1);
    flag.upload(&ok, 1, stream);
    nccl.check(nccl.AllReduce(flag.p, flag.p, 1, ncclInt, ncclMin, comm, stream), "ncclAllReduce");
    flag.download(&ok, 1, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
    peer_ok = ok != 0;
    peer_error.alloc(1);
    peer_error.zero(stream);
    if (peer_ok && std::getenv("MYFM_PEER_TRACE")) {
      peer_trace_buf.alloc(static_cast<size_t>(PEER_TRACE_RECORDS) * PEER_TRACE_SLOTS);
      peer_trace_buf.zero(stream);
    }
    MYFM_CUDA(cudaStreamSynchronize(stream));
    if (!peer_ok)
      close_peer_exchange();
  }
  // No collective here (a destructor runs whenever the host language drops the trainer, in any
  // order across ranks, possibly after a peer has died): a rank frees its buffer only after every
  // peer has said, by a flag written INTO that buffer, that its kernels no longer read it.  A
  // peer that does not say so within a few seconds costs a leaked buffer, never a hang or a fault.
  static constexpr size_t PEER_CLOSING_OFFSET = 64; // int[PEER_MAX_RANKS] inside the 256-byte header
  void close_peer_exchange(bool handshake = false) {
    bool all_done = true;
    if (handshake && peer_local) {
      const int one = 1;
      for (int r = 0; r < static_cast<int>(peer_base.size()); r++) // this rank's kernels are done (streams synced)
        if (peer_base[r] && r != my_rank)
          cudaMemcpy(peer_base[r] + PEER_CLOSING_OFFSET + sizeof(int) * my_rank, &one, sizeof(int),
                     cudaMemcpyHostToDevice);
      const auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        int flags[PEER_MAX_RANKS] = {0};
        if (cudaMemcpy(flags, peer_local + PEER_CLOSING_OFFSET, sizeof(flags), cudaMemcpyDeviceToHost) != cudaSuccess)
          break;
        all_done = true;
        for (int r = 0; r < world; r++)
          all_done = all_done && (r == my_rank || flags[r] != 0);
        if (all_done || std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 5.0)
          break;
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
      cudaGetLastError();
    }
    for (int r = 0; r < static_cast<int>(peer_base.size()); r++)
      if (peer_base[r] && r != my_rank)
        cudaIpcCloseMemHandle(peer_base[r]);
    peer_base.clear();
    if (peer_local && all_done)
      cudaFree(peer_local); // otherwise leaked on purpose: a peer may still be reading it
    peer_local = nullptr;
    peer_ok = false;
  }
  Real *peer_stat(int r) const { return reinterpret_cast<Real *>(peer_base[r] + PEER_HEADER_BYTES); }
  // header of a rank's buffer: [0] sequence number visible to the peers, [1] this rank's own counter
  unsigned long long *peer_counter() const { return reinterpret_cast<unsigned long long *>(peer_local) + 1; }
  PeerView<Real> peer_view() const {
    PeerView<Real> pv;
    pv.world = world, pv.counter = peer_counter(), pv.elems = peer_stat_elems, pv.error = peer_error.p;
    {
      static const double seconds = [] {
        const char *env = std::getenv("MYFM_PEER_TIMEOUT_S");
        const double v = env ? std::atof(env) : 0.0;
        return v > 0 ? v : 120.0;
      }();
      pv.timeout_ns = static_cast<unsigned long long>(seconds * 1e9);
    }
    pv.my_posted = reinterpret_cast<unsigned long long *>(peer_local);
    pv.done = reinterpret_cast<unsigned int *>(peer_local + 16);
    pv.trace = peer_trace_buf.p, pv.trace_cap = peer_trace_buf.p ? PEER_TRACE_RECORDS : 0;
    for (int r = 0; r < world; r++) {
      pv.stat[r] = peer_stat(r);
      pv.posted[r] = reinterpret_cast<const unsigned long long *>(peer_base[r]);
    }
    return pv;
  }
  DevBuf<int> peer_error;
  // MYFM_PEER_TRACE=<path prefix>: device time stamps of the last PEER_TRACE_RECORDS collectives, written to
  // <prefix>.rank<r>.csv when the trainer is destroyed (tools/peer_timeline.py reads them).
  static constexpr int PEER_TRACE_RECORDS = 4096;
  static constexpr int PHASE_TRACE_RECORDS = 512;
  DevBuf<unsigned long long> peer_trace_buf;
  DevBuf<unsigned long long> phase_trace_buf; // [PHASE_TRACE_RECORDS][PEER_TRACE_SLOTS] + the sweep counter
  // phases: 0 sweep starts, 1 alpha / w0 drawn, 2 lambda_w / mu_w drawn, 3 w swept, 4 lambda_V / mu_V drawn,
  // 5 V swept, 6 owned columns merged, 7 e refreshed
  void stamp_phase(int slot) {
    if (!phase_trace_buf.p)
      return;
    k_trace_stamp<<<1, 1, 0, stream>>>(phase_trace_buf.p, phase_trace_buf.p + static_cast<size_t>(PHASE_TRACE_RECORDS) * PEER_TRACE_SLOTS,
                                      PHASE_TRACE_RECORDS, slot);
  }
  void dump_peer_trace() {
    const char *prefix = std::getenv("MYFM_PEER_TRACE");
    if (!prefix)
      return;
    if (phase_trace_buf.p) {
      std::vector<unsigned long long> ph(static_cast<size_t>(PHASE_TRACE_RECORDS) * PEER_TRACE_SLOTS);
      if (cudaMemcpy(ph.data(), phase_trace_buf.p, ph.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
        const std::string ppath = std::string(prefix) + ".phases.rank" + std::to_string(my_rank) + ".csv";
        if (FILE *f = std::fopen(ppath.c_str(), "w")) {
          std::fprintf(f, "sweep_start,alpha_w0,hyper_w,w_swept,hyper_V,V_swept,merged,e_refreshed\n");
          for (int r = 0; r < PHASE_TRACE_RECORDS; r++) {
            const unsigned long long *q = ph.data() + static_cast<size_t>(r) * PEER_TRACE_SLOTS;
            if (q[0] == 0 || q[7] == 0)
              continue;
            for (int k = 0; k < 8; k++)
              std::fprintf(f, k ? ",%llu" : "%llu", q[k]);
            std::fprintf(f, "\n");
          }
          std::fclose(f);
        }
      }
    }
    if (!peer_trace_buf.p)
      return;
    std::vector<unsigned long long> t(static_cast<size_t>(PEER_TRACE_RECORDS) * PEER_TRACE_SLOTS);
    if (cudaMemcpy(t.data(), peer_trace_buf.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
      return;
    const std::string path = std::string(prefix) + ".rank" + std::to_string(my_rank) + ".csv";
    if (FILE *f = std::fopen(path.c_str(), "w")) {
      std::fprintf(f, "seq,stats_start,stats_posted,draw_start,draw_peers_seen,draw_done,stream_start,stream_done\n");
      for (int r = 0; r < PEER_TRACE_RECORDS; r++) {
        const unsigned long long *q = t.data() + static_cast<size_t>(r) * PEER_TRACE_SLOTS;
        if (q[0] == 0)
          continue;
        std::fprintf(f, "%llu", q[7]);
        for (int k = 0; k < 7; k++)
          std::fprintf(f, ",%llu", q[k]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
  }

  // ---- tile path (tile_sweep.cuh): two fields, row tiles staged in shared memory -----------------
  bool tile_path = false;
  int t_n_tiles = 0, t_n_colsL = 0;
  uint32_t t_eq_bytes = 0;
  size_t t_smem = 0;
  DevBuf<int> t_tile_row, t_item_ptr, t_n_cta, t_n_warp, t_b_ptr, t_colsL;
  DevBuf<SweepItem> t_items;
  DevBuf<unsigned> t_b_ent, t_c_ent;
  DevBuf<int> t_cls_ptr, t_long_ptr;
  DevBuf<int4> t_long_run;
  DevBuf<Real> t_b_val, t_c_val, t_part, t_pend;

  // Cuts the rows into tiles of whole first-field columns and builds, per tile, the first-field work
  // items and the sliced-ELL "B order" of the last field.  Xh: CSR in device row order.
  void setup_tile_path(const HostCs<Real> &Xh) {
    tile_path = false;
    // Opt-in (MYFM_TILE_PATH=1) where the field path applies: parity-green, but measured at 112 us per vector
    // against the field path's 107 us on the ml10m workload (DESIGN.md section 3c has the profile).
    // It also takes over, unasked, when the field path's table does not fit the shared memory but the tile
    // path's does (f64 with ~10^4 last-field columns: ml10m in double), where the alternative is the general
    // read-modify-write level kernels.
    const char *on = std::getenv("MYFM_TILE_PATH"), *never = std::getenv("MYFM_NO_TILE_PATH");
    const bool asked = on && on[0] == '1';
    if ((never && never[0] == '1') || !f_structure || (!asked && field_path) || f_tail != 1 || world > 1)
      return;
    const int64_t n = Xh.n_major;
    int dev_smem = 0;
    MYFM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t tab_bytes = (static_cast<size_t>(f_tab) * sizeof(Real) + 15) / 16 * 16;
    const size_t reserve = 2048; // the kernel's static shared memory
    if (static_cast<size_t>(dev_smem) < tab_bytes + reserve + 32 * 1024)
      return;
    int64_t cap = static_cast<int64_t>((dev_smem - reserve - tab_bytes) / (2 * sizeof(Real))) - 2;
    {
      const char *env = std::getenv("MYFM_TILE_ROWS"); // smaller tiles (tests: many tiles on small data)
      if (env && std::atoll(env) >= 64)
        cap = std::min<int64_t>(cap, std::atoll(env));
    }
    cap = std::min<int64_t>(cap, TILE_MAX_ROWS) & ~int64_t(1);
    // first-field columns in row order; empty ones are dealt out afterwards
    std::vector<SweepItem> cols, empty;
    int longest = 0;
    for (const SweepItem &it : f_items0_host) {
      if (it.hi > it.lo)
        cols.push_back(it), longest = std::max(longest, it.hi - it.lo);
      else
        empty.push_back(it);
    }
    if (longest > cap || cols.empty())
      return;
    std::sort(cols.begin(), cols.end(), [](const SweepItem &x, const SweepItem &y) { return x.lo < y.lo; });
    // as many tiles as a whole number of waves over the SMs needs, each near `goal` rows
    int64_t n_target = static_cast<int64_t>(f_sm_count) * ceil_div(n, static_cast<int64_t>(0.97 * cap) * f_sm_count);
    if (n / n_target < 2048)
      n_target = std::max<int64_t>(1, n / 2048);
    std::vector<int> tile_row{0}, tile_first_col{0};
    {
      int64_t remaining_rows = n, remaining_tiles = n_target, rows = 0;
      int64_t goal = std::min<int64_t>(cap, ceil_div(remaining_rows, remaining_tiles));
      for (size_t c = 0; c < cols.size(); c++) {
        const int len = cols[c].hi - cols[c].lo;
        if (rows > 0 && (rows + len > cap || rows + len / 2 > goal)) {
          tile_row.push_back(cols[c].lo), tile_first_col.push_back(static_cast<int>(c));
          remaining_rows -= rows, remaining_tiles = std::max<int64_t>(1, remaining_tiles - 1), rows = 0;
          goal = std::min<int64_t>(cap, ceil_div(remaining_rows, remaining_tiles));
        }
        rows += len;
      }
      tile_row.push_back(static_cast<int>(n)), tile_first_col.push_back(static_cast<int>(cols.size()));
    }
    const int n_tiles = static_cast<int>(tile_row.size()) - 1;
    // first-field items per tile, longest first; empty columns round-robin
    std::vector<std::vector<SweepItem>> per_tile(n_tiles);
    for (int t = 0; t < n_tiles; t++)
      per_tile[t].assign(cols.begin() + tile_first_col[t], cols.begin() + tile_first_col[t + 1]);
    for (size_t k = 0; k < empty.size(); k++)
      per_tile[k % n_tiles].push_back(empty[k]);
    std::vector<int> item_ptr{0}, n_cta(n_tiles, 0), n_warp(n_tiles, 0);
    std::vector<SweepItem> items;
    for (int t = 0; t < n_tiles; t++) {
      auto &v = per_tile[t];
      std::stable_sort(v.begin(), v.end(),
                       [](const SweepItem &x, const SweepItem &y) { return x.hi - x.lo > y.hi - y.lo; });
      for (const SweepItem &it : v) {
        n_cta[t] += (it.hi - it.lo > TILE_CTA_MIN);
        n_warp[t] += (it.hi - it.lo > TILE_TEAM_MAX && it.hi - it.lo <= TILE_CTA_MIN);
      }
      for (SweepItem it : v) {
        it.first = static_cast<int>(cfg.group_index[it.col]); // the 4th word carries the column's group
        items.push_back(it);
      }
      item_ptr.push_back(static_cast<int>(items.size()));
    }
    // B order of every tile: its rows sorted by (last-field column, row), one word per row
    if (f_tab > TILE_MAX_TAB)
      return;
    const int L = main_row_len;
    std::vector<int> b_ptr(n_tiles + 1, 0); // every tile's B range starts at a multiple of TILE_VEC words
    for (int t = 0; t < n_tiles; t++)
      b_ptr[t + 1] = b_ptr[t] + (tile_row[t + 1] - tile_row[t] + TILE_VEC - 1) / TILE_VEC * TILE_VEC;
    std::vector<unsigned> b_ent(b_ptr[n_tiles], TILE_NO_KEY);
    std::vector<Real> b_val(main_unit ? 0 : b_ptr[n_tiles], Real(0));
    struct TileRuns { // the statistics layout of one tile (tile_sweep.cuh: TileArgs::c_ent)
      std::vector<unsigned> ent;
      std::vector<Real> val;
      int cls[7] = {0, 0, 0, 0, 0, 0, 0}; // slot offsets of the classes inside ent, [5] .. [6]: long runs
      std::vector<int4> longs;          // first slot relative to ent
    };
    std::vector<TileRuns> tr(n_tiles);
    parallel_parts(std::min(parts_for(n * 4), n_tiles), [&](int part, int n_parts) {
      std::vector<int> cnt(f_tab + 1, 0), touched, run_start, run_len;
      for (int t = part; t < n_tiles; t += n_parts) { // counting sort by column, stable in the row
        const int r0 = tile_row[t], r1 = tile_row[t + 1];
        touched.clear();
        for (int i = r0; i < r1; i++) {
          const int c = Xh.idx[static_cast<size_t>(i) * L + (L - 1)] - f_last_base;
          if (cnt[c]++ == 0)
            touched.push_back(c);
        }
        std::sort(touched.begin(), touched.end());
        const int e0 = b_ptr[t]; // the tile's first B word
        int at = e0;
        for (int c : touched) {
          const int k = cnt[c];
          cnt[c] = at, at += k;
        }
        for (int i = r0; i < r1; i++) {
          const size_t p = static_cast<size_t>(i) * L + (L - 1);
          const int c = Xh.idx[p] - f_last_base;
          const int dst = cnt[c]++;
          b_ent[dst] = (static_cast<unsigned>(i - r0) << 16) | static_cast<unsigned>(c);
          if (!main_unit)
            b_val[dst] = Xh.val[p];
        }
        // runs in column order: run k = rows b_ent[run_start[k] .. + run_len[k])
        run_start.clear(), run_len.clear();
        {
          int at2 = e0;
          for (int c : touched) {
            const int len = cnt[c] - at2; // cnt[c] now points behind the run
            run_start.push_back(at2), run_len.push_back(len);
            at2 += len;
          }
        }
        TileRuns &o = tr[t];
        size_t slots = 0;
        for (int k = 0; k < 5; k++) {
          const int c = 1 << k;
          size_t in_class = 0;
          for (int len : run_len)
            if (len <= c && (k == 0 || len > c / 2))
              in_class += c;
          o.cls[k] = static_cast<int>(slots);
          slots += (in_class + TILE_VEC - 1) / TILE_VEC * TILE_VEC;
        }
        o.cls[5] = static_cast<int>(slots);
        size_t long_rows = 0;
        for (int len : run_len)
          if (len > 16)
            long_rows += (len + 3) / 4 * 4;
        o.cls[6] = static_cast<int>(slots + long_rows);
        o.ent.assign(slots + long_rows, TILE_NO_KEY);
        if (!main_unit)
          o.val.assign(slots + long_rows, Real(0));
        int fill_at[5];
        for (int k = 0; k < 5; k++)
          fill_at[k] = o.cls[k];
        size_t long_at = slots;
        for (size_t k = 0; k < run_len.size(); k++) {
          const int len = run_len[k];
          size_t dst;
          if (len > 16) {
            dst = long_at;
            o.longs.push_back(make_int4(touched[k], static_cast<int>(long_at), (len + 3) / 4 * 4, 0));
            long_at += (len + 3) / 4 * 4;
          } else {
            int cls = 0;
            while ((1 << cls) < len)
              cls++;
            dst = fill_at[cls], fill_at[cls] += 1 << cls;
          }
          for (int i = 0; i < len; i++) {
            o.ent[dst + i] = b_ent[run_start[k] + i];
            if (!main_unit)
              o.val[dst + i] = b_val[run_start[k] + i];
          }
        }
        std::stable_sort(o.longs.begin(), o.longs.end(), [](const int4 &x, const int4 &y) { return x.z > y.z; });
        for (int c : touched)
          cnt[c] = 0;
      }
    });
    size_t n_slots = 0, n_long = 0;
    for (const TileRuns &o : tr)
      n_slots += (o.ent.size() + TILE_VEC - 1) / TILE_VEC * TILE_VEC, n_long += o.longs.size();
    if (n_slots >= static_cast<size_t>(std::numeric_limits<int>::max()))
      return;
    std::vector<unsigned> c_ent(n_slots, TILE_NO_KEY);
    std::vector<Real> c_val(main_unit ? 0 : n_slots);
    std::vector<int> cls_ptr(static_cast<size_t>(n_tiles) * 8, 0), long_ptr{0};
    std::vector<int4> long_run;
    long_run.reserve(n_long);
    {
      size_t at = 0;
      for (int t = 0; t < n_tiles; t++) {
        const TileRuns &o = tr[t];
        std::copy(o.ent.begin(), o.ent.end(), c_ent.begin() + at);
        if (!main_unit)
          std::copy(o.val.begin(), o.val.end(), c_val.begin() + at);
        for (int k = 0; k < 7; k++)
          cls_ptr[static_cast<size_t>(t) * 8 + k] = static_cast<int>(at) + o.cls[k];
        for (int4 r : o.longs) {
          r.y += static_cast<int>(at);
          long_run.push_back(r);
        }
        long_ptr.push_back(static_cast<int>(long_run.size()));
        at += (o.ent.size() + TILE_VEC - 1) / TILE_VEC * TILE_VEC;
      }
    }
    if (long_run.empty())
      long_run.push_back(make_int4(0, 0, 0, 0)); // never read
    std::vector<int> colsL;
    for (size_t j = 0; j < f_level_host.size(); j++)
      if (f_level_host[j] == L - 1)
        colsL.push_back(static_cast<int>(j));
    t_n_tiles = n_tiles, t_n_colsL = static_cast<int>(colsL.size());
    t_eq_bytes = static_cast<uint32_t>((static_cast<size_t>(cap) + 2) * 2 * sizeof(Real) + 15) / 16 * 16;
    t_smem = t_eq_bytes + tab_bytes;
    t_tile_row.upload(tile_row, stream), t_item_ptr.upload(item_ptr, stream), t_n_cta.upload(n_cta, stream);
    t_n_warp.upload(n_warp, stream), t_b_ptr.upload(b_ptr, stream);
    t_items.upload(items, stream);
    t_b_ent.upload(b_ent, stream);
    t_c_ent.upload(c_ent, stream);
    t_cls_ptr.upload(cls_ptr, stream), t_long_ptr.upload(long_ptr, stream), t_long_run.upload(long_run, stream);
    if (!main_unit)
      t_b_val.upload(b_val, stream), t_c_val.upload(c_val, stream);
    t_colsL.upload(colsL, stream);
    t_part.alloc(2 * static_cast<size_t>(f_tab) * n_tiles); // [tile][column]; absent (tile, column) pairs stay zero
    t_part.zero(stream);
    t_pend.alloc(2 * static_cast<size_t>(f_tab));
    t_pend.zero(stream);
    MYFM_CUDA(cudaStreamSynchronize(stream)); // host staging vectors die here
    tile_path = true;
  }

  template <bool IS_V, bool UNIT, int PEND> void launch_tile_sweep(const TileArgs<Real> &a) {
    auto kernel = k_tile_sweep<Real, IS_V, UNIT, PEND>;
    static size_t configured = 0; // per instantiation
    if (t_smem > configured) {
      MYFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(t_smem)));
      configured = t_smem;
    }
    kernel<<<t_n_tiles, TILE_THREADS, t_smem, stream>>>(a);
    launched();
  }

  // One vector over a two-field table on the tile path: k_tile_sweep, then k_tile_fold.
  template <bool IS_V>
  void sweep_tile(Real *theta, Real *theta_t, int64_t t_stride, const Real *z, const Real *lambda, const Real *mu) {
    TimedSpan span(timer, stream, 0);
    const int pend = !f_pending_valid ? PEND_NONE : (IS_V && f_pending_is_v ? PEND_V : PEND_W);
    {
      TimedSpan span_stream(timer, stream, 3);
      TileArgs<Real> a;
      a.tile_row = t_tile_row.p, a.tile_item_ptr = t_item_ptr.p, a.tile_n_cta = t_n_cta.p;
      a.tile_n_warp = t_n_warp.p, a.tile_b_ptr = t_b_ptr.p;
      a.item = reinterpret_cast<const int4 *>(t_items.p);
      a.b_ent = t_b_ent.p, a.b_val = t_b_val.p;
      a.c_ent = t_c_ent.p, a.c_val = t_c_val.p;
      a.tile_cls_ptr = t_cls_ptr.p, a.tile_long_ptr = t_long_ptr.p, a.long_run = t_long_run.p;
      a.n_tiles = t_n_tiles, a.eq_bytes = t_eq_bytes;
      a.eq = eq();
      a.own_val = f_own_val.p;
      a.theta = theta, a.theta_t = theta_t, a.t_stride = t_stride;
      a.z = z, a.group = group.p, a.alpha = hv().alpha, a.lambda = lambda, a.mu = mu;
      a.last_base = f_last_base, a.n_tab = f_tab;
      a.pend = reinterpret_cast<const Pair<Real> *>(t_pend.p);
      a.part = reinterpret_cast<Pair<Real> *>(t_part.p);
#define MYFM_TS(V, P)                                                                              \
  if (main_unit)                                                                                   \
    launch_tile_sweep<V, true, P>(a);                                                              \
  else                                                                                             \
    launch_tile_sweep<V, false, P>(a);
      if (!IS_V) {
        if (pend != PEND_NONE)
          throw std::logic_error("tile path: the w sweep must not find a pending update.");
        MYFM_TS(false, PEND_NONE)
      } else if (pend == PEND_NONE) {
        MYFM_TS(true, PEND_NONE)
      } else if (pend == PEND_W) {
        MYFM_TS(true, PEND_W)
      } else {
        MYFM_TS(true, PEND_V)
      }
#undef MYFM_TS
    }
    {
      TimedSpan span_fold(timer, stream, 4);
      TileFoldArgs<Real> f;
      f.cols = t_colsL.p, f.n_cols = t_n_colsL, f.n_tiles = t_n_tiles, f.last_base = f_last_base, f.n_tab = f_tab;
      f.part = reinterpret_cast<const Pair<Real> *>(t_part.p);
      f.theta = theta, f.theta_t = theta_t, f.t_stride = t_stride;
      f.z = z, f.group = group.p, f.alpha = hv().alpha, f.lambda = lambda, f.mu = mu;
      f.pend = reinterpret_cast<Pair<Real> *>(t_pend.p);
      f.to_peer = 0, f.peer.world = 0, f.peer_local = nullptr, f.colstat = nullptr;
      if (t_n_colsL) {
        k_tile_fold<Real, IS_V><<<ceil_div(t_n_colsL, 32), 32 * FOLD_GROUPS, 0, stream>>>(f);
        launched();
      }
    }
    f_pending_valid = true, f_pending_is_v = IS_V;
  }

  template <bool IS_V, bool UNIT, int PEND> void launch_field_stream(const FieldStreamArgs<Real> &a, int mode) {
    if (IS_V && f_tail > 1)
      launch_field_stream_mid<IS_V, UNIT, true, PEND>(a, mode);
    else
      launch_field_stream_mid<IS_V, UNIT, false, PEND>(a, mode);
  }
  template <bool IS_V, bool UNIT, bool HAS_MID, int PEND>
  void launch_field_stream_mid(const FieldStreamArgs<Real> &a, int mode) {
    if (mode == FIELD_FUSED)
      launch_field_stream_as<IS_V, UNIT, HAS_MID, PEND, FIELD_FUSED>(a);
    else if (mode == FIELD_STATS)
      launch_field_stream_as<IS_V, UNIT, HAS_MID, PEND, FIELD_STATS>(a);
    else
      launch_field_stream_as<IS_V, UNIT, HAS_MID, PEND, FIELD_UPDATE>(a);
  }
  template <bool IS_V, bool UNIT, bool HAS_MID, int PEND, int MODE>
  void launch_field_stream_as(const FieldStreamArgs<Real> &a) {
    // UNIT tables, fused pass: the warp class stages its rows through TMA bulk copies when the compact table
    // and the 32 staging buffers fit the shared memory (opt-in: MYFM_STAGING=1)
    if (UNIT && MODE == FIELD_FUSED && f_staged) {
      launch_field_stream_kernel(k_field_stream<Real, IS_V, UNIT, HAS_MID, PEND, MODE, true>, a, f_staged_smem);
      return;
    }
    launch_field_stream_kernel(k_field_stream<Real, IS_V, UNIT, HAS_MID, PEND, MODE, false>, a,
                               static_cast<size_t>(f_tab) * FieldTab<Real>::BYTES_PER_COLUMN);
  }
  template <typename Kernel> void launch_field_stream_kernel(Kernel kernel, const FieldStreamArgs<Real> &a, size_t smem) {
    static std::map<const void *, size_t> configured; // dynamic shared memory granted per kernel
    size_t &have = configured[reinterpret_cast<const void *>(kernel)];
    if (smem > have) {
      MYFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      have = smem;
    }
    kernel<<<f_sm_count, FIELD_THREADS, smem, stream>>>(a);
    launched();
  }
  bool f_staged = false;
  size_t f_staged_smem = 0;

  // One vector (w or a factor column) over the main table: streaming pass, middle levels, gather.
  template <bool IS_V>
  void sweep_field(Real *theta, Real *theta_t, int64_t t_stride, const Real *z, const Real *lambda,
                   const Real *mu) {
    TimedSpan span(timer, stream, 0);
    const int pend = !f_pending_valid ? PEND_NONE : (IS_V && f_pending_is_v ? PEND_V : PEND_W);
    {
      TimedSpan span_stream(timer, stream, 3);
      FieldStreamArgs<Real> a;
      a.item = reinterpret_cast<const int4 *>(f_items0.p + f_level0.s0);
      a.nCC = f_nCC, a.nCR = f_nCR, a.nG = f_nG, a.nW = f_nW;
      { // about four scheduling steps per warp: large batches amortise the counter, small ones the tail
        const char *env = std::getenv("MYFM_FIELD_BATCH");
        const int warps = f_sm_count * FIELD_WARPS;
        int b = env ? std::atoi(env) : ceil_div(f_nW, 4 * static_cast<int64_t>(warps));
        a.batch = std::max(1, std::min(b, FIELD_BATCH_MAX));
      }
      a.eq = eq(), a.n_rows = N, a.n_tail = f_tail;
      a.tail_idx = f_tail_idx.p, a.tail_val = f_tail_val.p, a.own_val = f_own_val.p;
      a.tail_last = f_tail_idx.p + static_cast<int64_t>(f_tail - 1) * N;
      a.tval_last = f_tail_val.p ? f_tail_val.p + static_cast<int64_t>(f_tail - 1) * N : nullptr;
      a.theta = theta, a.theta_t = theta_t, a.t_stride = t_stride;
      a.z = z, a.group = group.p, a.alpha = hv().alpha, a.lambda = lambda, a.mu = mu;
      a.last_base = f_last_base, a.n_tab = f_tab;
      a.pend_told = f_pend_told.p, a.pend_tnew = f_pend_tnew.p;
      a.item_slot = f_item_slot0.p + f_level0.s0, a.colstat = f_colstat.p;
#define MYFM_FS(V, P)                                                                              \
  if (main_unit)                                                                                   \
    launch_field_stream<V, true, P>(a, mode);                                                      \
  else                                                                                             \
    launch_field_stream<V, false, P>(a, mode);
      // one GPU: one fused pass.  Row shards: statistics, sum over the ranks (peer memory inside the
      // update kernel, or an NCCL all-reduce between the two), update.
      a.peer.world = 0;
      if (peer_ok)
        a.peer = peer_view(), a.peer_local = peer_stat(my_rank);
      const bool two_pass = world > 1 && !f_exclusive;
      for (int pass = 0; pass < (two_pass ? 2 : 1); pass++) {
        const int mode = two_pass ? (pass == 0 ? FIELD_STATS : FIELD_UPDATE) : FIELD_FUSED;
        a.sched = f_sched.p + (f_launch++);
        if (!IS_V) {
          if (pend != PEND_NONE)
            throw std::logic_error("field path: the w sweep must not find a pending update.");
          MYFM_FS(false, PEND_NONE)
        } else if (pend == PEND_NONE) {
          MYFM_FS(true, PEND_NONE)
        } else if (pend == PEND_W) {
          MYFM_FS(true, PEND_W)
        } else {
          MYFM_FS(true, PEND_V)
        }
        if (mode == FIELD_STATS) {
          if (!peer_ok) // with the peer exchange the producing kernel's last CTA has published already
            allreduce_sum(f_colstat.p, 2 * static_cast<size_t>(f_ncols0));
        }
      }
#undef MYFM_FS
    }
    if (f_tail > 1)
      sweep_main<IS_V>(theta, theta_t, t_stride, z, lambda, mu, 1, static_cast<size_t>(f_tail));
    {
      TimedSpan span_stats(timer, stream, 4);
      FieldStatsArgs<Real> a;
      a.idx = Xt.idx.p, a.val = Xt.val.p;
      a.item = reinterpret_cast<const int4 *>(f_itemsL.p + f_levelL.s0);
      a.seg_count = f_seg_countL.p + f_levelL.s0;
      a.nS = f_levelL.c0 - f_levelL.s0, a.nC = f_levelL.w0 - f_levelL.c0, a.nW = f_levelL.end - f_levelL.w0;
      a.eq = eq();
      a.theta = theta, a.theta_t = theta_t, a.t_stride = t_stride;
      a.z = z, a.group = group.p, a.alpha = hv().alpha, a.lambda = lambda, a.mu = mu;
      a.partial = f_partial.p, a.chunk_done = f_chunk_done.p, a.last_base = f_last_base;
      a.item_slot = f_item_slotL.p + f_levelL.s0;
      a.colstat = world > 1 && !peer_ok ? f_colstat.p : nullptr;
      a.peer.world = 0, a.peer_local = nullptr;
      if (peer_ok)
        a.peer = peer_view(), a.peer_local = peer_stat(my_rank);
      a.pend_told = f_pend_told.p, a.pend_tnew = f_pend_tnew.p;
      const int grid = a.nS + a.nC + ceil_div(a.nW, STATS_THREADS / 32);
      if (grid) {
        if (main_unit)
          k_field_stats<Real, IS_V, true><<<grid, STATS_THREADS, 0, stream>>>(a);
        else
          k_field_stats<Real, IS_V, false><<<grid, STATS_THREADS, 0, stream>>>(a);
        launched();
      }
      if (world > 1) { // statistics of this rank's rows -> sum over the ranks -> identical draw everywhere
        if (!peer_ok)
          allreduce_sum(f_colstat.p, 2 * static_cast<size_t>(f_ncolsL));
        k_field_draw_last<Real, IS_V><<<ceil_div(f_ncolsL, 256), 256, 0, stream>>>(a, a.peer, f_colsL.p, f_ncolsL);
        launched();
      }
    }
    f_pending_valid = true, f_pending_is_v = IS_V;
  }
  bool f_pending_is_v = false;

  // q_init of the main table when it is one-hot shaped (all values 1, rows of equal length)
  bool main_unit = false;
  int main_row_len = 0;
  // columns of a level = items minus the extra chunks of its long columns
  int count_chunk_surplus(const SweepLevel &L) const {
    int surplus = 0;
    for (int i = L.s0; i < L.c0; i++)
      if (plan.items[i].first == i - L.s0)
        surplus += plan.seg_count[i] - 1;
    return surplus;
  }

  void spmv(const DevCs<Real> &A, const Real *x, Real *out, bool squared, int out_stride = 1) {
    if (!A.n_major)
      return;
    if (&A == &data.X && main_unit && main_row_len > 0 && main_row_len <= 4 && !squared) {
      const int n = static_cast<int>(A.n_major);
      k_spmv<Real, 1, false, true><<<ceil_div(n, 256), 256, 0, stream>>>(n, A.view(), x, out, out_stride,
                                                                          main_row_len);
      launched();
      return;
    }
    const int lpr = pow2_ceil_clamped(A.avg_len() / 4.0, 1, 32);
    const int n = static_cast<int>(A.n_major);
    const int grid = ceil_div(static_cast<int64_t>(n) * lpr, 256);
#define MYFM_SPMV(L)                                                                               \
  case L:                                                                                          \
    if (squared)                                                                                   \
      k_spmv<Real, L, true><<<grid, 256, 0, stream>>>(n, A.view(), x, out, out_stride);            \
    else                                                                                           \
      k_spmv<Real, L, false><<<grid, 256, 0, stream>>>(n, A.view(), x, out, out_stride);           \
    break;
    switch (lpr) {
      MYFM_SPMV(1)
      MYFM_SPMV(2)
      MYFM_SPMV(4)
      MYFM_SPMV(8)
      MYFM_SPMV(16)
      MYFM_SPMV(32)
    }
#undef MYFM_SPMV
    launched();
  }

  template <bool IS_V>
  void rel_sweep(int b, Real *theta, Real *theta_t, int64_t t_stride, const Real *z,
                 const Real *lambda, const Real *mu) {
    auto &d = data.rels[b];
    auto &t = rel_train[b];
    if (!d.F)
      return;
    RelSweepArgs<Real> a;
    a.Bt = t.Bt.view();
    a.n_levels = t.n_levels, a.level_ptr = t.level_ptr.p, a.level_cols = t.level_cols.p;
    a.cache = t.cache();
    a.theta = theta + d.offset;
    a.theta_t = theta_t ? theta_t + d.offset * t_stride : nullptr;
    a.t_stride = t_stride;
    a.z = z + d.offset, a.group = group.p + d.offset;
    a.alpha = hv().alpha, a.lambda = lambda, a.mu = mu;
    // block caches in shared memory when they fit (7 arrays of S reals), else in global memory
    int dev_smem = 0;
    MYFM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t smem = 7 * static_cast<size_t>(d.S) * sizeof(Real);
    const char *no_smem = std::getenv("MYFM_REL_GLOBAL");
    TimedSpan span(timer, stream, 0); // the block's column sweep belongs to the column-sweep family
    if (smem + 1024 <= static_cast<size_t>(dev_smem) && !(no_smem && no_smem[0] == '1')) {
      auto kernel = k_rel_sweep_smem<Real, IS_V>;
      static size_t configured = 0; // per instantiation
      if (smem > configured) {
        MYFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = smem;
      }
      kernel<<<1, 512, smem, stream>>>(a, static_cast<int>(d.S), t.run_end.p, t.level_rec.p);
    } else {
      k_rel_sweep<Real, IS_V><<<1, 1024, 0, stream>>>(a);
    }
    launched();
  }

  // update_w (FMTrainer.hpp:231-314)
  void update_w(const Real *z) {
    HyperView<Real> h = hv();
    if (!cfg.fit_linear) {
      w.zero(stream); // e keeps the stale contribution until update_e, as in the reference
      return;
    }
    if (tile_path)
      sweep_tile<false>(w.p, nullptr, 0, z, h.lambda_w, h.mu_w);
    else if (field_path)
      sweep_field<false>(w.p, nullptr, 0, z, h.lambda_w, h.mu_w);
    else
      sweep_main<false>(w.p, nullptr, 0, z, h.lambda_w, h.mu_w);
    const int n = static_cast<int>(N);
    for (size_t b = 0; b < data.rels.size(); b++) {
      auto &d = data.rels[b];
      auto &t = rel_train[b];
      spmv(d.B, w.p + d.offset, t.q.p, false);
      if (d.S) {
        k_rel_gather_w<Real><<<ceil_div(d.S * 32, 256), 256, 0, stream>>>(
            static_cast<int>(d.S), t.seg_ptr.p, t.seg_rows.p, t.cache(), eq());
        launched();
      }
      rel_sweep<false>(static_cast<int>(b), w.p, nullptr, 0, z, h.lambda_w, h.mu_w);
      spmv(d.B, w.p + d.offset, t.q.p, false);
      if (n) {
        k_rel_add_rows<Real><<<ceil_div(n, 256), 256, 0, stream>>>(n, d.map.p, t.q.p, e_ptr());
        launched();
      }
    }
  }

  // update_V (FMTrainer.hpp:316-486)
  void update_V(const Real *z_all) {
    HyperView<Real> h = hv();
    const int n = static_cast<int>(N);
    for (int r = 0; r < K; r++) {
      Real *Vr = V.p + static_cast<size_t>(D_all) * r;
      const Real *z = z_all + static_cast<size_t>(D_all) * r;
      const Real *lam = h.lambda_V + static_cast<size_t>(G) * r, *mu = h.mu_V + static_cast<size_t>(G) * r;
      if (tile_path) {
        sweep_tile<true>(Vr, Vt.p + r, K, z, lam, mu);
        continue;
      }
      if (field_path) {
        sweep_field<true>(Vr, Vt.p + r, K, z, lam, mu);
        continue;
      }
      {
        TimedSpan span(timer, stream, 1);
        if (D)
          spmv(data.X, Vr, q_ptr(), false, 2);
        else if (n) {
          k_fill_strided<Real><<<ceil_div(n, 256), 256, 0, stream>>>(n, q_ptr(), 2, Real(0));
          launched();
        }
        for (size_t b = 0; b < data.rels.size(); b++) {
          auto &d = data.rels[b];
          auto &t = rel_train[b];
          spmv(d.B, Vr + d.offset, t.q.p, false);
          if (n) {
            k_rel_add_rows<Real><<<ceil_div(n, 256), 256, 0, stream>>>(n, d.map.p, t.q.p, q_ptr());
            launched();
          }
        }
      }
      sweep_main<true>(Vr, Vt.p + r, K, z, lam, mu);
      for (size_t b = 0; b < data.rels.size(); b++) {
        auto &d = data.rels[b];
        auto &t = rel_train[b];
        spmv(d.B, Vr + d.offset, t.q_S.p, true);
        if (d.S) {
          k_rel_gather_v<Real><<<ceil_div(d.S * 32, 256), 256, 0, stream>>>(
              static_cast<int>(d.S), t.seg_ptr.p, t.seg_rows.p, t.cache(), eq());
          launched();
        }
        rel_sweep<true>(static_cast<int>(b), Vr, Vt.p + r, K, z, lam, mu);
        if (n) {
          k_rel_resync_v<Real><<<ceil_div(n, 256), 256, 0, stream>>>(n, d.map.p, t.cache(), eq());
          launched();
        }
      }
    }
  }

  // update_e for classification (FMTrainer.hpp:498-512): the truncated-normal draws consume the
  // mt19937 stream row by row, data dependently, so in MT19937 mode they run on the host.
  void classification_latent() {
    if (philox_latents) {
      if (N) {
        k_latent_classification<Real><<<ceil_div(N, 256), 256, 0, stream>>>(
            N, eq(), y.p, latent_row.p, latent_seed, static_cast<uint32_t>(sweep_index + 1));
        launched();
      }
      return;
    }
    e_host.resize(N);
    export_component(0, e_host.data()); // caller's row order: the stream is consumed row by row
    const Real zero = 0, sd = 1;
    for (int64_t i = 0; i < N; i++) {
      Real pred = e_host[i];
      Real n = y_host[i] > 0 ? rng.tn_left(pred, sd, zero) : rng.tn_right(pred, sd, zero);
      e_host[i] -= n;
    }
    import_component(0, e_host.data());
  }

  // Builds the samplers of the cut-point groups (OProbitSampler.hpp:25-49: label checks).
  void build_cut_groups() {
    cut_groups.clear();
    cut_groups.reserve(cfg.cutpoint_groups.size()); // the samplers' callbacks keep pointers into it
    std::vector<int> inv(N);
    for (int64_t i = 0; i < N; i++)
      inv[perm[i]] = static_cast<int>(i);
    int max_class = 1;
    for (auto &grp : cfg.cutpoint_groups) {
      cut_groups.emplace_back();
      CutGroup &cg = cut_groups.back();
      cg.n_class = grp.first;
      if (cg.n_class < 2)
        throw std::invalid_argument("a cutpoint group needs at least two classes.");
      cg.rows = grp.second;
      std::vector<int> class_ptr(cg.n_class + 1, 0), class_rows(cg.rows.size());
      for (int64_t i : cg.rows) {
        const Real yi = y_host[i];
        const int label = static_cast<int>(yi);
        if (std::abs(label - yi) > 1e-3)
          throw std::invalid_argument("y has a floating-point element.");
        if (label < 0)
          throw std::invalid_argument("y has a negative element.");
        if (label >= cg.n_class) {
          std::ostringstream ss;
          ss << "y[ " << i << "] is greater than " << (cg.n_class - 1) << ".";
          throw std::invalid_argument(ss.str());
        }
        class_ptr[label + 1]++;
      }
      for (int c = 0; c < cg.n_class; c++)
        class_ptr[c + 1] += class_ptr[c];
      std::vector<int> cur(class_ptr.begin(), class_ptr.end() - 1);
      for (int64_t i : cg.rows)
        class_rows[cur[static_cast<int>(y_host[i])]++] = inv[i];
      cg.class_ptr.upload(class_ptr, stream);
      cg.class_rows.upload(class_rows, stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
      cg.cutpoints.assign(cg.n_class - 1, Real(0));
      CutGroup *self = &cg;
      cg.sampler = std::make_unique<CutpointSampler<Real>>(
          cg.n_class, static_cast<Real>(cfg.reg_0), static_cast<Real>(cfg.nu_oprobit),
          [this, self](const std::vector<Real> &gamma, std::vector<Real> &sums) {
            cutpoint_row_sums(*self, gamma, sums);
          });
      max_class = std::max(max_class, cg.n_class);
    }
    if (world > 1) { // the samplers of all ranks must walk the same Newton / MH steps
      std::vector<int> sizes;
      for (const CutGroup &cg : cut_groups)
        sizes.push_back(cg.n_class), sizes.push_back(-cg.n_class);
      sizes.push_back(static_cast<int>(cut_groups.size())), sizes.push_back(-static_cast<int>(cut_groups.size()));
      DevBuf<int> buf(sizes.size());
      std::vector<int> got(sizes.size());
      buf.upload(sizes, stream);
      NcclApi &nccl = NcclApi::get();
      nccl.check(nccl.AllReduce(buf.p, buf.p, sizes.size(), ncclInt, ncclMax, comm, stream), "ncclAllReduce");
      buf.download(got.data(), got.size(), stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
      for (size_t k = 0; k < sizes.size(); k += 2)
        if (got[k] != sizes[k] || got[k + 1] != sizes[k + 1])
          throw std::invalid_argument("cutpoint groups (their number and class counts) must be the same on every rank.");
    }
    op_gamma.alloc(max_class);
    op_partial.alloc(static_cast<size_t>(max_class) * OP_BLOCKS_PER_CLASS * OP_TERMS);
    op_sums.alloc(static_cast<size_t>(max_class) * OP_TERMS);
  }

  // Label-wise sums of the cut-point log-posterior terms over the group's rows, on the device;
  // e holds the current scores while the cut-points are being sampled (FMTrainer.hpp:494,513-521).
  void cutpoint_row_sums(CutGroup &cg, const std::vector<Real> &gamma, std::vector<Real> &sums) {
    const int Kc = cg.n_class;
    op_gamma.upload(gamma.data(), gamma.size(), stream);
    k_oprobit_terms<Real><<<Kc * OP_BLOCKS_PER_CLASS, OP_THREADS, 0, stream>>>(
        Kc, cg.class_ptr.p, cg.class_rows.p, e_ptr(), 2, op_gamma.p, op_partial.p);
    k_oprobit_finish<Real><<<ceil_div(Kc * OP_TERMS, 128), 128, 0, stream>>>(Kc, op_partial.p, op_sums.p);
    launched(2);
    MYFM_CUDA(cudaGetLastError());
    if (world > 1) // OProbitSampler.hpp:389-463 sums over ALL rows: every rank gets the same totals
      allreduce_sum(op_sums.p, static_cast<size_t>(Kc) * OP_TERMS);
    sums.resize(static_cast<size_t>(Kc) * OP_TERMS);
    op_sums.download(sums.data(), sums.size(), stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
  }

  // ORDERED part of initialize_e (start = true, FMTrainer.hpp:101-116) and of update_e
  // (:513-521): cut-point move, then the latent z of every row from a normal truncated to the
  // label's interval (OProbitSampler.hpp:238-272); e <- score - z.  The truncated normals consume
  // the mt19937 stream row by row, so they are drawn on the host.
  void ordered_update(bool start) {
    if (start)
      build_cut_groups();
    if (philox_latents) { // cut-point move on the host (device row sums), latent z on the device
      for (CutGroup &cg : cut_groups) {
        if (start)
          cg.sampler->start();
        else
          cg.sampler->step(cut_gen); // not rng.gen: the device stream took that state over in setup_rng
        cg.cutpoints = cg.sampler->gamma_now;
        const int n_group = static_cast<int>(cg.rows.size());
        if (!n_group)
          continue;
        MYFM_CUDA(cudaStreamSynchronize(stream)); // lat_gamma may still be read by the previous group's kernel
        lat_gamma.upload(cg.cutpoints, stream);
        k_latent_ordered<Real><<<ceil_div(n_group, 256), 256, 0, stream>>>(
            cg.n_class, cg.class_ptr.p, cg.class_rows.p, lat_gamma.p, eq(), latent_row.p, latent_seed,
            start ? 0u : static_cast<uint32_t>(sweep_index + 1));
        launched();
      }
      MYFM_CUDA(cudaStreamSynchronize(stream)); // host staging of the cut-points dies with this call
      return;
    }
    e_host.resize(N);
    export_component(0, e_host.data());
    for (CutGroup &cg : cut_groups) {
      if (start)
        cg.sampler->start();
      else
        cg.sampler->step(rng.gen);
      cg.cutpoints = cg.sampler->gamma_now;
      const std::vector<Real> &gamma = cg.cutpoints;
      const int Kc = cg.n_class;
      const Real deviation = 1;
      for (int64_t i : cg.rows) {
        const int label = static_cast<int>(y_host[i]);
        const Real pred = e_host[i];
        Real z;
        if (label == 0)
          z = deviation * rng.tn_right((gamma[label] - pred) / deviation) + pred;
        else if (label == Kc - 1)
          z = deviation * rng.tn_left((gamma[Kc - 2] - pred) / deviation) + pred;
        else
          z = deviation * rng.tn_twoside((gamma[label - 1] - pred) / deviation,
                                         (gamma[label] - pred) / deviation) + pred;
        e_host[i] -= z;
      }
    }
    import_component(0, e_host.data());
  }
  DevBuf<Real> lat_gamma;
  // PHILOX mode: the Metropolis-Hastings move of the cut-points (normals, chi-square, acceptance
  // uniform) draws from its own generator.  The sweep's Gaussian / Gamma variates continue the
  // mt19937(seed) stream on the device from the state the host generator had after create_FM, so
  // stepping the host copy here would replay the words of the first sweeps.
  std::mt19937 cut_gen;

  // eq component (0 = e, 1 = q) <-> a dense host vector in the caller's row order
  void export_component(int comp, Real *host) {
    if (!N)
      return;
    k_eq_export<Real><<<ceil_div(N, 256), 256, 0, stream>>>(N, eq_buf.p, comp, perm_dev.p, dense_tmp.p);
    launched();
    dense_tmp.download(host, N, stream);
    MYFM_CUDA(cudaStreamSynchronize(stream));
  }
  void import_component(int comp, const Real *host) {
    if (!N)
      return;
    dense_tmp.upload(host, N, stream);
    k_eq_import<Real><<<ceil_div(N, 256), 256, 0, stream>>>(N, eq_buf.p, comp, perm_dev.p, dense_tmp.p);
    launched();
    MYFM_CUDA(cudaStreamSynchronize(stream));
  }

  // Row-sharded: the block partials of a grid reduction become one scalar summed over the ranks.
  int fold_over_ranks() {
    if (world == 1)
      return REDUCE_BLOCKS;
    k_fold_partials<Real><<<1, 256, 0, stream>>>(REDUCE_BLOCKS, partial.p);
    launched();
    allreduce_sum(partial.p, 1);
    return 1;
  }

  // The device work of one update_all after the variates are in place: kernel launches and memsets
  // on `stream` only (capturable).
  cudaGraphExec_t sweep_graph[2] = {nullptr, nullptr};
  int64_t sweep_graph_launches[2] = {0, 0};
  bool use_graphs = true;
  void reset_graphs() {
    for (auto &g : sweep_graph) {
      if (g)
        cudaGraphExecDestroy(g);
      g = nullptr;
    }
  }
  void sweep_body(const Real *z) {
    HyperView<Real> h = hv();
    const SweepLayout &L = layout;
    if (field_path || tile_path) { // work counters of this sweep's streaming passes; nothing is pending after update_e
      f_sched.zero(stream);
      f_launch = 0, f_pending_valid = false;
    }

    stamp_phase(0);
    const bool alpha_w0_fused = cfg.task_type == MYFM_TASK_REGRESSION && cfg.fit_w0;
    if (alpha_w0_fused) { // update_alpha + update_w0: one pass over e for both sums, one exchange, one draw kernel
      k_reduce_e_both<Real><<<REDUCE_BLOCKS, 512, 0, stream>>>(N, eq(), h.w0, partial.p);
      int n_partial = REDUCE_BLOCKS;
      if (world > 1 && peer_ok) { // the two sums over the ranks through the peer-memory exchange
        k_fold_partials2_peer<Real><<<1, 256, 0, stream>>>(REDUCE_BLOCKS, partial.p, peer_view(), peer_stat(my_rank));
        k_finish_alpha_w0_peer<Real><<<1, 32, 0, stream>>>(peer_view(), static_cast<Real>(cfg.beta_0), z + L.g_alpha,
                                                            h.alpha, static_cast<int>(N_global),
                                                            static_cast<Real>(cfg.reg_0), z + L.z_w0, h.w0, scal.p);
        launched();
      } else {
        if (world > 1) {
          k_fold_partials2<Real><<<1, 256, 0, stream>>>(REDUCE_BLOCKS, partial.p);
          launched();
          allreduce_sum(partial.p, 2);
          n_partial = 1;
        }
        k_finish_alpha_w0<Real><<<1, 256, 0, stream>>>(n_partial, partial.p, partial.p + n_partial,
                                                        static_cast<Real>(cfg.beta_0), z + L.g_alpha, h.alpha,
                                                        static_cast<int>(N_global), static_cast<Real>(cfg.reg_0),
                                                        z + L.z_w0, h.w0, scal.p);
      }
      k_add_scalar<Real><<<REDUCE_BLOCKS, 256, 0, stream>>>(N, eq(), scal.p);
      launched(3);
    } else if (cfg.task_type == MYFM_TASK_REGRESSION) { // update_alpha
      k_reduce_e<Real, 0><<<REDUCE_BLOCKS, 512, 0, stream>>>(N, eq(), h.w0, partial.p);
      const int n_partial = fold_over_ranks();
      k_finish_alpha<Real><<<1, 256, 0, stream>>>(n_partial, partial.p,
                                                   static_cast<Real>(cfg.beta_0), z + L.g_alpha, h.alpha);
      launched(2);
    }
    if (alpha_w0_fused) { // done above
    } else if (cfg.fit_w0) { // update_w0
      k_reduce_e<Real, 1><<<REDUCE_BLOCKS, 512, 0, stream>>>(N, eq(), h.w0, partial.p);
      const int n_partial = fold_over_ranks();
      k_finish_w0<Real><<<1, 256, 0, stream>>>(n_partial, partial.p, static_cast<int>(N_global),
                                                static_cast<Real>(cfg.reg_0), h.alpha, z + L.z_w0, h.w0,
                                                scal.p);
      k_add_scalar<Real><<<REDUCE_BLOCKS, 256, 0, stream>>>(N, eq(), scal.p);
      launched(3);
    } else {
      MYFM_CUDA(cudaMemsetAsync(h.w0, 0, sizeof(Real), stream));
    }
    stamp_phase(1);
    if (G) { // update_lambda_w, update_mu_w
      k_group_hyper<Real><<<hyper_chunks, HYPER_THREADS, 0, stream>>>(
          G, hyper_chunks, hyper_chunk_group.p, hyper_chunk_begin.p, feat_ptr.p, feat_idx.p, w.p, 0, h.mu_w, h.lambda_w,
          z + L.g_lw, z + L.z_mw, static_cast<Real>(cfg.beta_0), static_cast<Real>(cfg.gamma_0),
          static_cast<Real>(cfg.mu_0), hyper_chunk_sums.p, hyper_done.p);
      launched();
    }
    stamp_phase(2);
    update_w(L.z_w >= 0 ? z + L.z_w : nullptr);
    stamp_phase(3);
    if (G && K) { // update_lambda_V, update_mu_V
      k_group_hyper<Real><<<hyper_chunks * K, HYPER_THREADS, 0, stream>>>(
          G, hyper_chunks, hyper_chunk_group.p, hyper_chunk_begin.p, feat_ptr.p, feat_idx.p, V.p, D_all, h.mu_V,
          h.lambda_V, z + L.g_lV, z + L.z_mV, static_cast<Real>(cfg.beta_0), static_cast<Real>(cfg.gamma_0),
          static_cast<Real>(cfg.mu_0), hyper_chunk_sums.p, hyper_done.p);
      launched();
    }
    stamp_phase(4);
    update_V(z + L.z_V);
    stamp_phase(5);
    merge_owned();
    stamp_phase(6);
    f_pending_valid = false; // update_e overwrites e: the last vector's pending update is dropped
    { // update_e
      TimedSpan span(timer, stream, 2);
      data.predict(w.p, Vt.p, K, h.w0, cfg.task_type == MYFM_TASK_REGRESSION ? y.p : nullptr, e_ptr(), 2);
    }
    stamp_phase(7);
  }

  // update_all (BaseFMTrainer.hpp:135-152)
  void sweep() {
    const int slot = static_cast<int>(sweep_index & 1);
    const Real *z = nullptr;
    if (device_rng) {
      while (gen_index <= sweep_index + 1) // this sweep's variates, and the next one's ahead of time
        launch_variates(gen_index++);
      MYFM_CUDA(cudaStreamWaitEvent(stream, z_ready[slot], 0));
      z = z_slot[slot].p;
    } else {
      if (sweep_index >= 2)
        MYFM_CUDA(cudaEventSynchronize(z_copied[slot]));
      draw_sweep_variates(z_pinned[slot].p);
      MYFM_CUDA(cudaMemcpyAsync(z_dev.p, z_pinned[slot].p, layout.total * sizeof(Real),
                                cudaMemcpyHostToDevice, stream));
      MYFM_CUDA(cudaEventRecord(z_copied[slot], stream));
      z = z_dev.p;
    }
    z_last = z;
    // Regression with the device RNG never needs the host inside a sweep: the launch sequence of a
    // slot is captured once into a CUDA graph and replayed (the very first sweep runs directly:
    // first-use allocations and function attributes are not capturable).
    const bool graphable = device_rng && (world == 1 || (field_path && peer_ok)) &&
                           cfg.task_type == MYFM_TASK_REGRESSION && !timer.enabled && use_graphs && sweep_index >= 1;
    if (graphable) {
      if (!sweep_graph[slot]) {
        const int64_t before = launches;
        cudaGraph_t g = nullptr;
        MYFM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        try {
          sweep_body(z);
        } catch (...) {
          cudaStreamEndCapture(stream, &g);
          if (g)
            cudaGraphDestroy(g);
          throw;
        }
        MYFM_CUDA(cudaStreamEndCapture(stream, &g));
        MYFM_CUDA(cudaGraphInstantiate(&sweep_graph[slot], g, 0));
        MYFM_CUDA(cudaGraphDestroy(g));
        sweep_graph_launches[slot] = launches - before;
        launches = before;
      }
      MYFM_CUDA(cudaGraphLaunch(sweep_graph[slot], stream));
      launches += sweep_graph_launches[slot];
    } else {
      sweep_body(z);
    }
    if (cfg.task_type == MYFM_TASK_CLASSIFICATION)
      classification_latent();
    if (cfg.task_type == MYFM_TASK_ORDERED)
      ordered_update(false);
    if (device_rng)
      MYFM_CUDA(cudaEventRecord(z_free[slot], stream));
    MYFM_CUDA(cudaGetLastError());
    sweep_index++;
  }

  void require_fm() const {
    if (K < 0)
      throw std::runtime_error("myfm_trainer_init_fm must be called first.");
  }

  void step(int n) override {
    require_fm();
    MYFM_CUDA(cudaSetDevice(device));
    for (int i = 0; i < n; i++)
      sweep();
  }
  void sync() override {
    MYFM_CUDA(cudaStreamSynchronize(stream));
    if (timer.enabled)
      timer.collect();
    check_rng_error();
  }
  // the standardised variates the most recent sweep consumed (diagnostics / RNG tests)
  int64_t get_variates(double *out, int64_t capacity) override {
    require_fm();
    if (!z_last || capacity < layout.total)
      return layout.total;
    auto h = fetch(z_last, layout.total);
    for (int64_t i = 0; i < layout.total; i++)
      out[i] = h[i];
    return layout.total;
  }
  double timed_steps(int n) override {
    require_fm();
    MYFM_CUDA(cudaSetDevice(device));
    cudaEvent_t a, b;
    MYFM_CUDA(cudaEventCreate(&a));
    MYFM_CUDA(cudaEventCreate(&b));
    sync();
    MYFM_CUDA(cudaEventRecord(a, stream));
    for (int i = 0; i < n; i++)
      sweep();
    MYFM_CUDA(cudaEventRecord(b, stream));
    sync();
    float ms = 0;
    MYFM_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a), cudaEventDestroy(b);
    return ms;
  }
  void dims(int64_t *n_train, int64_t *dim_all, int32_t *rank, int32_t *n_groups) const override {
    *n_train = N, *dim_all = D_all, *rank = K, *n_groups = G;
  }

  template <typename T> std::vector<T> fetch(const T *dev, size_t n) {
    std::vector<T> h(n);
    if (n)
      MYFM_CUDA(cudaMemcpyAsync(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
    MYFM_CUDA(cudaStreamSynchronize(stream));
    return h;
  }

  void get_fm(double *w0_out, double *w_out, double *V_out) override { // any of the three may be NULL
    require_fm();
    if (w0_out)
      *w0_out = fetch(hyper.p, 2)[1];
    // widened on the device, then one copy straight into the caller's buffer
    auto fetch_double = [&](const Real *dev, size_t n, double *out) {
      if (!n)
        return;
      if (wide_tmp.n < n)
        wide_tmp.alloc(n);
      k_to_double<Real><<<ceil_div(n, 256), 256, 0, stream>>>(static_cast<int64_t>(n), dev, wide_tmp.p);
      launched();
      MYFM_CUDA(cudaMemcpyAsync(out, wide_tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
      MYFM_CUDA(cudaStreamSynchronize(stream));
    };
    if (w_out)
      fetch_double(w.p, D_all, w_out);
    if (V_out)
      fetch_double(Vt.p, static_cast<size_t>(D_all) * K, V_out);
  }
  DevBuf<double> wide_tmp;
  void get_cutpoints(int g, double *out) override {
    if (g < 0 || g >= static_cast<int>(cut_groups.size()))
      throw std::runtime_error("No cutpoint available for this FM.");
    for (size_t c = 0; c < cut_groups[g].cutpoints.size(); c++)
      out[c] = cut_groups[g].cutpoints[c];
  }
  void get_hyper(double *alpha, double *mu_w, double *lambda_w, double *mu_V, double *lambda_V) override {
    require_fm();
    queue_error_flags();
    auto hh = fetch(hyper.p, hyper_size());
    throw_on_error_flags();
    *alpha = hh[0];
    for (int g = 0; g < G; g++)
      mu_w[g] = hh[2 + g], lambda_w[g] = hh[2 + G + g];
    const Real *mv = hh.data() + 2 + 2 * G, *lv = mv + static_cast<size_t>(G) * K;
    for (int g = 0; g < G; g++)
      for (int r = 0; r < K; r++) // device: g + G*r  ->  boundary: g*K + r
        mu_V[static_cast<size_t>(g) * K + r] = mv[g + static_cast<size_t>(G) * r],
                                    lambda_V[static_cast<size_t>(g) * K + r] = lv[g + static_cast<size_t>(G) * r];
  }
  void get_e(double *out) override {
    std::vector<Real> h(N);
    export_component(0, h.data());
    for (int64_t i = 0; i < N; i++)
      out[i] = h[i];
  }
  void get_q(double *out) override {
    if ((field_path || tile_path) && K > 0) // the field / tile path leave q undefined between sweeps: q = X V[:, K-1]
      spmv(data.X, V.p + static_cast<size_t>(D_all) * (K - 1), q_ptr(), false, 2);
    std::vector<Real> h(N);
    export_component(1, h.data());
    for (int64_t i = 0; i < N; i++)
      out[i] = h[i];
  }
  int64_t mh_accept(int g) override {
    if (g < 0 || g >= static_cast<int>(cut_groups.size()))
      throw std::invalid_argument("cutpoint group index out of range.");
    return cut_groups[g].sampler->accept_count;
  }

  // Overwrites parts of the chain state (NULL = keep).  Layouts as in the getters.
  void set_state(const double *w0_in, const double *w_in, const double *V_in, const double *alpha_in,
                 const double *mu_w_in, const double *lambda_w_in, const double *mu_V_in,
                 const double *lambda_V_in, const double *e_in) override {
    require_fm();
    MYFM_CUDA(cudaStreamSynchronize(stream));
    std::vector<Real> hh = fetch(hyper.p, hyper_size());
    if (alpha_in)
      hh[0] = static_cast<Real>(*alpha_in);
    if (w0_in)
      hh[1] = static_cast<Real>(*w0_in);
    for (int g = 0; g < G; g++) {
      if (mu_w_in)
        hh[2 + g] = static_cast<Real>(mu_w_in[g]);
      if (lambda_w_in)
        hh[2 + G + g] = static_cast<Real>(lambda_w_in[g]);
      for (int r = 0; r < K; r++) {
        if (mu_V_in)
          hh[2 + 2 * G + g + static_cast<size_t>(G) * r] = static_cast<Real>(mu_V_in[static_cast<size_t>(g) * K + r]);
        if (lambda_V_in)
          hh[2 + 2 * G + static_cast<size_t>(G) * K + g + static_cast<size_t>(G) * r] =
              static_cast<Real>(lambda_V_in[static_cast<size_t>(g) * K + r]);
      }
    }
    hyper.upload(hh, stream);
    std::vector<Real> tmp;
    if (w_in) {
      tmp.assign(w_in, w_in + D_all);
      w.upload(tmp, stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
    }
    if (V_in) {
      std::vector<Real> vt(static_cast<size_t>(D_all) * K), vc(vt.size());
      for (int64_t j = 0; j < D_all; j++)
        for (int r = 0; r < K; r++) {
          Real v = static_cast<Real>(V_in[static_cast<size_t>(j) * K + r]);
          vt[static_cast<size_t>(j) * K + r] = v;
          vc[j + static_cast<size_t>(D_all) * r] = v;
        }
      Vt.upload(vt, stream);
      V.upload(vc, stream);
      MYFM_CUDA(cudaStreamSynchronize(stream));
    }
    if (e_in) {
      tmp.assign(e_in, e_in + N);
      import_component(0, tmp.data());
    }
    MYFM_CUDA(cudaStreamSynchronize(stream));
  }
  int64_t launch_count() const override { return launches; }
  int sweep_path() const override {
    if (tile_path)
      return 5;
    return field_path ? (peer_ok ? 2 : 1) + (f_exclusive ? 2 : 0) : 0;
  }
  std::unique_ptr<SampleBase> snapshot() override { // device-to-device copy of the current sample
    require_fm();
    MYFM_CUDA(cudaSetDevice(device));
    auto sm = std::make_unique<Sample<Real>>();
    sm->dtype = dtype, sm->device = device, sm->K = K, sm->dim_all = D_all;
    sm->w0.alloc(1), sm->w.alloc(D_all), sm->Vt.alloc(static_cast<size_t>(D_all) * K);
    MYFM_CUDA(cudaMemcpyAsync(sm->w0.p, hv().w0, sizeof(Real), cudaMemcpyDeviceToDevice, stream));
    if (D_all)
      MYFM_CUDA(cudaMemcpyAsync(sm->w.p, w.p, D_all * sizeof(Real), cudaMemcpyDeviceToDevice, stream));
    if (sm->Vt.n)
      MYFM_CUDA(cudaMemcpyAsync(sm->Vt.p, Vt.p, sm->Vt.n * sizeof(Real), cudaMemcpyDeviceToDevice, stream));
    MYFM_CUDA(cudaStreamSynchronize(stream));
    return sm;
  }
  void kernel_ms(int family, double *ms, int64_t *n) override {
    sync();
    if (family < 0 || family >= KernelTimer::FAMILIES)
      throw std::invalid_argument("kernel family must be 0 .. 4.");
    *ms = timer.ms[family], *n = timer.launches[family];
    timer.ms[family] = 0, timer.launches[family] = 0;
  }
  void set_profiling(bool on) override {
    sync();
    timer.enabled = on;
  }
  // One step of a LibFM-style callback (libfm.py:44-54, :82-113, :150-182, :226-262) with the live sample:
  // forward pass, link, running sums and metric terms on the device; EVAL_TERMS sums come back.
  void evaluate(EvalState &ev, int iteration, const double *cutpoints, int n_cpt, double *terms) override {
    require_fm();
    if (ev.dataset->dtype != dtype)
      throw std::invalid_argument("dataset and trainer use different compute dtypes.");
    auto *d = static_cast<Dataset<Real> *>(ev.dataset);
    d->check_dim(D_all);
    if (ev.task == MYFM_TASK_ORDERED && n_cpt + 1 != ev.width)
      throw std::invalid_argument("the number of cut-points does not match the evaluator's class count.");
    MYFM_CUDA(cudaSetDevice(device));
    MYFM_CUDA(cudaStreamSynchronize(stream)); // the sample must be final before another stream reads it
    if (d->score.n < static_cast<size_t>(d->n_rows))
      d->score.alloc(d->n_rows);
    d->predict(w.p, Vt.p, K, hv().w0, nullptr, d->score.p);
    ev.n_samples++;
    EvalArgs a;
    a.n = static_cast<int>(ev.n), a.width = ev.width, a.task = ev.task, a.iteration = iteration;
    a.n_samples = ev.n_samples, a.burn_in = 5;
    a.clip_min = ev.clip_min, a.clip_max = ev.clip_max, a.eps = ev.eps;
    a.sum = ev.sum.p, a.late = ev.late.p, a.y = ev.y.p, a.cutpoints = ev.cutp.p, a.partial = ev.partial.p;
    if (ev.task == MYFM_TASK_ORDERED && n_cpt > 0)
      ev.cutp.upload(cutpoints, n_cpt, d->stream);
    if (ev.n) {
      k_eval<Real><<<EVAL_BLOCKS, EVAL_THREADS, 0, d->stream>>>(a, d->score.p);
      k_eval_finish<<<1, 32, 0, d->stream>>>(EVAL_BLOCKS, ev.partial.p, ev.out.p);
      d->count(2);
      MYFM_CUDA(cudaGetLastError());
      ev.out.download(terms, EVAL_TERMS, d->stream);
    } else {
      std::fill(terms, terms + EVAL_TERMS, 0.0);
    }
    MYFM_CUDA(cudaStreamSynchronize(d->stream));
  }

  void predict_score(DatasetBase *db, double *out) override {
    require_fm();
    if (db->dtype != dtype)
      throw std::invalid_argument("dataset and trainer use different compute dtypes.");
    auto *d = static_cast<Dataset<Real> *>(db);
    d->check_dim(D_all);
    MYFM_CUDA(cudaStreamSynchronize(stream)); // the sample must be final before another stream reads it
    if (d->score.n < static_cast<size_t>(d->n_rows))
      d->score.alloc(d->n_rows);
    d->predict(w.p, Vt.p, K, hv().w0, nullptr, d->score.p);
    std::vector<Real> h(d->n_rows);
    d->score.download(h.data(), d->n_rows, d->stream);
    MYFM_CUDA(cudaStreamSynchronize(d->stream));
    for (int64_t i = 0; i < d->n_rows; i++)
      out[i] = h[i];
  }
};

// ------------------------------------------------------------------------------------------------
// prediction over host-supplied samples
// ------------------------------------------------------------------------------------------------
template <typename Real>
void dataset_predict_score(Dataset<Real> &d, double w0, const double *w, const double *V, int64_t dim_all,
                           int K, double *out) {
  d.check_dim(dim_all);
  MYFM_CUDA(cudaSetDevice(d.device));
  d.stage_sample(w0, w, V, K);
  if (d.score.n < static_cast<size_t>(d.n_rows))
    d.score.alloc(d.n_rows);
  d.predict(d.w.p, d.Vt.p, K, d.w0.p, nullptr, d.score.p);
  std::vector<Real> h(d.n_rows);
  d.score.download(h.data(), d.n_rows, d.stream);
  MYFM_CUDA(cudaStreamSynchronize(d.stream));
  for (int64_t i = 0; i < d.n_rows; i++)
    out[i] = h[i];
}

// predictor.hpp:126-147 (and :35-76, same arithmetic) / :78-124 with n_cpt >= 0
template <typename Real>
void dataset_predict_mean(Dataset<Real> &d, int task, int n_samples, const double *w0s, const double *ws,
                          const double *Vs, const double *cutpoints, int n_cpt, int64_t dim_all, int K,
                          double *out) {
  d.check_dim(dim_all);
  if (n_samples <= 0)
    throw std::runtime_error("Told to predict but no sample available.");
  MYFM_CUDA(cudaSetDevice(d.device));
  const bool ordered = n_cpt >= 0;
  const size_t width = ordered ? n_cpt + 1 : 1;
  const int n = static_cast<int>(d.n_rows);
  if (d.score.n < static_cast<size_t>(n))
    d.score.alloc(n);
  d.accum.alloc(static_cast<size_t>(n) * width);
  d.accum.zero(d.stream);
  for (int s = 0; s < n_samples; s++) {
    d.stage_sample(w0s[s], ws + static_cast<size_t>(s) * dim_all, Vs + static_cast<size_t>(s) * dim_all * K, K);
    d.predict(d.w.p, d.Vt.p, K, d.w0.p, nullptr, d.score.p);
    if (!n)
      continue;
    if (ordered) {
      std::vector<Real> cp(n_cpt);
      for (int c = 0; c < n_cpt; c++)
        cp[c] = static_cast<Real>(cutpoints[static_cast<size_t>(s) * n_cpt + c]);
      d.cutp.upload(cp, d.stream);
      MYFM_CUDA(cudaStreamSynchronize(d.stream));
      k_accumulate_oprobit<Real><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.cutp.p, n_cpt, d.accum.p);
    } else if (task == MYFM_TASK_CLASSIFICATION) {
      k_accumulate<Real, 1><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.accum.p);
    } else if (task == MYFM_TASK_REGRESSION) {
      k_accumulate<Real, 0><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.accum.p);
    } // ORDERED through Predictor::predict leaves zeros (predictor.hpp:136-143)
    d.count();
  }
  if (n) {
    k_scale<Real><<<ceil_div(static_cast<int64_t>(n) * width, 256), 256, 0, d.stream>>>(
        static_cast<int64_t>(n) * width, d.accum.p, static_cast<Real>(n_samples));
    d.count();
  }
  std::vector<Real> h(static_cast<size_t>(n) * width);
  d.accum.download(h.data(), h.size(), d.stream);
  MYFM_CUDA(cudaStreamSynchronize(d.stream));
  for (size_t i = 0; i < h.size(); i++)
    out[i] = h[i];
}

// The same mean over samples that already live on the device: no staging.
template <typename Real>
void dataset_predict_mean_samples(Dataset<Real> &d, int task, int n_samples, SampleBase *const *samples,
                                  const double *cutpoints, int n_cpt, double *out) {
  if (n_samples <= 0)
    throw std::runtime_error("Told to predict but no sample available.");
  MYFM_CUDA(cudaSetDevice(d.device));
  const bool ordered = n_cpt >= 0;
  const size_t width = ordered ? n_cpt + 1 : 1;
  const int n = static_cast<int>(d.n_rows);
  if (d.score.n < static_cast<size_t>(n))
    d.score.alloc(n);
  d.accum.alloc(static_cast<size_t>(n) * width);
  d.accum.zero(d.stream);
  for (int s = 0; s < n_samples; s++) {
    if (!samples[s] || samples[s]->dtype != d.dtype || samples[s]->device != d.device)
      throw std::invalid_argument("sample and dataset differ in compute dtype or device.");
    auto *sm = static_cast<Sample<Real> *>(samples[s]);
    d.check_dim(sm->dim_all);
    d.predict(sm->w.p, sm->Vt.p, sm->K, sm->w0.p, nullptr, d.score.p);
    if (!n)
      continue;
    if (ordered) {
      std::vector<Real> cp(n_cpt);
      for (int c = 0; c < n_cpt; c++)
        cp[c] = static_cast<Real>(cutpoints[static_cast<size_t>(s) * n_cpt + c]);
      MYFM_CUDA(cudaStreamSynchronize(d.stream)); // the previous sample's kernel still reads cutp
      d.cutp.upload(cp, d.stream);
      MYFM_CUDA(cudaStreamSynchronize(d.stream));
      k_accumulate_oprobit<Real><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.cutp.p, n_cpt, d.accum.p);
    } else if (task == MYFM_TASK_CLASSIFICATION) {
      k_accumulate<Real, 1><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.accum.p);
    } else if (task == MYFM_TASK_REGRESSION) {
      k_accumulate<Real, 0><<<ceil_div(n, 256), 256, 0, d.stream>>>(n, d.score.p, d.accum.p);
    }
    d.count();
  }
  if (n) {
    k_scale<Real><<<ceil_div(static_cast<int64_t>(n) * width, 256), 256, 0, d.stream>>>(
        static_cast<int64_t>(n) * width, d.accum.p, static_cast<Real>(n_samples));
    d.count();
  }
  std::vector<Real> h(static_cast<size_t>(n) * width);
  d.accum.download(h.data(), h.size(), d.stream);
  MYFM_CUDA(cudaStreamSynchronize(d.stream));
  for (size_t i = 0; i < h.size(); i++)
    out[i] = h[i];
}

} // namespace myfm

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace myfm;

struct myfm_trainer {
  std::unique_ptr<TrainerBase> impl;
};
struct myfm_dataset {
  std::unique_ptr<DatasetBase> impl;
};
struct myfm_sample {
  std::unique_ptr<SampleBase> impl;
};

#define MYFM_API_BEGIN try {
#define MYFM_API_END                                                                               \
  return MYFM_OK;                                                                                  \
  }                                                                                                \
  catch (const std::invalid_argument &ex) {                                                        \
    g_last_error = ex.what();                                                                      \
    return MYFM_ERR_INVALID_ARGUMENT;                                                              \
  }                                                                                                \
  catch (const CudaError &ex) {                                                                    \
    g_last_error = ex.what();                                                                      \
    return MYFM_ERR_CUDA;                                                                          \
  }                                                                                                \
  catch (const std::exception &ex) {                                                               \
    g_last_error = ex.what();                                                                      \
    return MYFM_ERR_RUNTIME;                                                                       \
  }

static void require(const void *p, const char *what) {
  if (!p)
    throw std::invalid_argument(std::string(what) + " must not be NULL.");
}

extern "C" {

const char *myfm_last_error(void) { return g_last_error.c_str(); }

int myfm_device_count(int32_t *count) {
  MYFM_API_BEGIN
  require(count, "count");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  MYFM_API_END
}

int myfm_config_validate(const myfm_config_t *cfg, int32_t *n_groups) {
  MYFM_API_BEGIN
  require(cfg, "cfg");
  Config c(*cfg);
  if (n_groups)
    *n_groups = c.n_groups;
  MYFM_API_END
}

int myfm_trainer_create(myfm_trainer_t **out, const myfm_csr_t *X, int32_t n_relations,
                        const myfm_relation_t *relations, const double *y, int64_t n_y,
                        int32_t random_seed, const myfm_config_t *cfg, const myfm_engine_options_t *opt) {
  MYFM_API_BEGIN
  require(out, "out"), require(X, "X"), require(cfg, "cfg"), require(opt, "opt");
  if (n_y > 0)
    require(y, "y");
  if (n_relations > 0)
    require(relations, "relations");
  auto holder = std::make_unique<myfm_trainer>();
  if (opt->dtype == MYFM_DTYPE_F32)
    holder->impl = std::make_unique<Trainer<float>>(*X, n_relations, relations, y, n_y, random_seed, *cfg, *opt);
  else if (opt->dtype == MYFM_DTYPE_F64)
    holder->impl = std::make_unique<Trainer<double>>(*X, n_relations, relations, y, n_y, random_seed, *cfg, *opt);
  else
    throw std::invalid_argument("unknown dtype.");
  *out = holder.release();
  MYFM_API_END
}

void myfm_trainer_destroy(myfm_trainer_t *t) { delete t; }

int myfm_trainer_init_fm(myfm_trainer_t *t, int32_t rank, double init_std) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->init_fm(rank, init_std);
  MYFM_API_END
}
int myfm_trainer_step(myfm_trainer_t *t, int32_t n_sweeps) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->step(n_sweeps);
  MYFM_API_END
}
int myfm_trainer_sync(myfm_trainer_t *t) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->sync();
  MYFM_API_END
}
int myfm_trainer_timed_steps(myfm_trainer_t *t, int32_t n_sweeps, double *ms) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(ms, "ms");
  *ms = t->impl->timed_steps(n_sweeps);
  MYFM_API_END
}
int myfm_trainer_dims(const myfm_trainer_t *t, int64_t *n_train, int64_t *dim_all, int32_t *rank, int32_t *n_groups) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->dims(n_train, dim_all, rank, n_groups);
  MYFM_API_END
}
int myfm_trainer_get_fm(myfm_trainer_t *t, double *w0, double *w, double *V) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->get_fm(w0, w, V);
  MYFM_API_END
}
int myfm_trainer_get_cutpoints(myfm_trainer_t *t, int32_t g, double *out) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->get_cutpoints(g, out);
  MYFM_API_END
}
int myfm_trainer_get_hyper(myfm_trainer_t *t, double *alpha, double *mu_w, double *lambda_w, double *mu_V,
                           double *lambda_V) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->get_hyper(alpha, mu_w, lambda_w, mu_V, lambda_V);
  MYFM_API_END
}
int myfm_trainer_get_e(myfm_trainer_t *t, double *e) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->get_e(e);
  MYFM_API_END
}
int myfm_trainer_get_q(myfm_trainer_t *t, double *q) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->get_q(q);
  MYFM_API_END
}
int myfm_trainer_mh_accept(myfm_trainer_t *t, int32_t g, int64_t *count) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(count, "count");
  *count = t->impl->mh_accept(g);
  MYFM_API_END
}
int myfm_trainer_set_state(myfm_trainer_t *t, const double *w0, const double *w, const double *V,
                           const double *alpha, const double *mu_w, const double *lambda_w,
                           const double *mu_V, const double *lambda_V, const double *e) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->set_state(w0, w, V, alpha, mu_w, lambda_w, mu_V, lambda_V, e);
  MYFM_API_END
}
int myfm_trainer_get_variates(myfm_trainer_t *t, double *out, int64_t capacity, int64_t *n) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(n, "n");
  *n = t->impl->get_variates(out, capacity);
  MYFM_API_END
}
int myfm_trainer_launch_count(const myfm_trainer_t *t, int64_t *count) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(count, "count");
  *count = t->impl->launch_count();
  MYFM_API_END
}
int myfm_trainer_sweep_path(const myfm_trainer_t *t, int32_t *path) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(path, "path");
  *path = t->impl->sweep_path();
  MYFM_API_END
}
int myfm_trainer_set_profiling(myfm_trainer_t *t, int32_t on) {
  MYFM_API_BEGIN
  require(t, "trainer");
  t->impl->set_profiling(on != 0);
  MYFM_API_END
}
int myfm_trainer_kernel_ms(myfm_trainer_t *t, int32_t family, double *ms, int64_t *launches) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(ms, "ms"), require(launches, "launches");
  t->impl->kernel_ms(family, ms, launches);
  MYFM_API_END
}

int myfm_dataset_create(myfm_dataset_t **out, const myfm_csr_t *X, int32_t n_relations,
                        const myfm_relation_t *relations, int32_t dtype, int32_t device) {
  MYFM_API_BEGIN
  require(out, "out"), require(X, "X");
  if (n_relations > 0)
    require(relations, "relations");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= device)
    throw CudaError("no usable CUDA device: the myfm_b200 engine has no CPU fallback.");
  MYFM_CUDA(cudaSetDevice(device));
  auto holder = std::make_unique<myfm_dataset>();
  auto make = [&](auto tag) {
    using Real = decltype(tag);
    auto d = std::make_unique<Dataset<Real>>();
    d->dtype = dtype, d->device = device;
    cudaStream_t s;
    MYFM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    d->own_stream = true;
    d->stream = s;
    d->build(host_from_api<Real>(*X, "X"), n_relations, relations, s);
    holder->impl = std::move(d);
  };
  if (dtype == MYFM_DTYPE_F32)
    make(float{});
  else if (dtype == MYFM_DTYPE_F64)
    make(double{});
  else
    throw std::invalid_argument("unknown dtype.");
  *out = holder.release();
  MYFM_API_END
}
void myfm_dataset_destroy(myfm_dataset_t *d) { delete d; }

int myfm_predict_score(const myfm_dataset_t *d, double w0, const double *w, const double *V, int64_t dim_all,
                       int32_t rank, double *out) {
  MYFM_API_BEGIN
  require(d, "dataset");
  if (d->impl->dtype == MYFM_DTYPE_F32)
    dataset_predict_score(*static_cast<Dataset<float> *>(d->impl.get()), w0, w, V, dim_all, rank, out);
  else
    dataset_predict_score(*static_cast<Dataset<double> *>(d->impl.get()), w0, w, V, dim_all, rank, out);
  MYFM_API_END
}

int myfm_predict_mean(const myfm_dataset_t *d, int32_t task_type, int32_t n_samples, const double *w0s,
                      const double *ws, const double *Vs, int64_t dim_all, int32_t rank, double *out) {
  MYFM_API_BEGIN
  require(d, "dataset");
  if (d->impl->dtype == MYFM_DTYPE_F32)
    dataset_predict_mean(*static_cast<Dataset<float> *>(d->impl.get()), task_type, n_samples, w0s, ws, Vs, nullptr,
                         -1, dim_all, rank, out);
  else
    dataset_predict_mean(*static_cast<Dataset<double> *>(d->impl.get()), task_type, n_samples, w0s, ws, Vs, nullptr,
                         -1, dim_all, rank, out);
  MYFM_API_END
}

int myfm_predict_oprobit_mean(const myfm_dataset_t *d, int32_t n_samples, const double *w0s, const double *ws,
                              const double *Vs, const double *cutpoints, int32_t n_cpt, int64_t dim_all,
                              int32_t rank, double *out) {
  MYFM_API_BEGIN
  require(d, "dataset");
  if (n_cpt < 0)
    throw std::invalid_argument("n_cpt must be non-negative.");
  if (d->impl->dtype == MYFM_DTYPE_F32)
    dataset_predict_mean(*static_cast<Dataset<float> *>(d->impl.get()), MYFM_TASK_ORDERED, n_samples, w0s, ws, Vs,
                         cutpoints, n_cpt, dim_all, rank, out);
  else
    dataset_predict_mean(*static_cast<Dataset<double> *>(d->impl.get()), MYFM_TASK_ORDERED, n_samples, w0s, ws, Vs,
                         cutpoints, n_cpt, dim_all, rank, out);
  MYFM_API_END
}

int myfm_trainer_snapshot(myfm_trainer_t *t, myfm_sample_t **out) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(out, "out");
  auto holder = std::make_unique<myfm_sample>();
  holder->impl = t->impl->snapshot();
  *out = holder.release();
  MYFM_API_END
}
void myfm_sample_destroy(myfm_sample_t *s) { delete s; }
int myfm_sample_get(const myfm_sample_t *s, double *w0, double *w, double *V) {
  MYFM_API_BEGIN
  require(s, "sample");
  s->impl->get(w0, w, V);
  MYFM_API_END
}
int myfm_predict_samples_mean(const myfm_dataset_t *d, int32_t task_type, int32_t n_samples,
                              myfm_sample_t *const *samples, const double *cutpoints, int32_t n_cpt, double *out) {
  MYFM_API_BEGIN
  require(d, "dataset"), require(samples, "samples");
  if (task_type == MYFM_TASK_ORDERED && n_cpt >= 0)
    require(cutpoints, "cutpoints");
  std::vector<SampleBase *> impls(std::max(0, n_samples));
  for (int i = 0; i < n_samples; i++) {
    require(samples[i], "sample");
    impls[i] = samples[i]->impl.get();
  }
  if (d->impl->dtype == MYFM_DTYPE_F32)
    dataset_predict_mean_samples(*static_cast<Dataset<float> *>(d->impl.get()), task_type, n_samples, impls.data(),
                                 cutpoints, n_cpt, out);
  else
    dataset_predict_mean_samples(*static_cast<Dataset<double> *>(d->impl.get()), task_type, n_samples, impls.data(),
                                 cutpoints, n_cpt, out);
  MYFM_API_END
}

int myfm_trainer_predict_score(myfm_trainer_t *t, const myfm_dataset_t *d, double *out) {
  MYFM_API_BEGIN
  require(t, "trainer"), require(d, "dataset");
  t->impl->predict_score(d->impl.get(), out);
  MYFM_API_END
}

struct myfm_evaluator {
  EvalState state;
};

int myfm_evaluator_create(myfm_evaluator_t **out, const myfm_dataset_t *d, const double *y_test, int64_t n_test,
                          int32_t task_type, int32_t n_class, double clip_min, double clip_max, double eps) {
  MYFM_API_BEGIN
  require(out, "out"), require(d, "dataset");
  if (n_test > 0)
    require(y_test, "y_test");
  if (n_test != d->impl->n_rows)
    throw std::invalid_argument("y_test must have one entry per row of the dataset.");
  if (task_type == MYFM_TASK_ORDERED && n_class < 2)
    throw std::invalid_argument("ordered probit needs at least two classes.");
  MYFM_CUDA(cudaSetDevice(d->impl->device));
  auto holder = std::make_unique<myfm_evaluator>();
  EvalState &ev = holder->state;
  ev.dataset = d->impl.get();
  ev.task = task_type, ev.width = task_type == MYFM_TASK_ORDERED ? n_class : 1, ev.n = n_test;
  ev.clip_min = clip_min, ev.clip_max = clip_max, ev.eps = eps;
  const size_t cells = static_cast<size_t>(n_test) * ev.width;
  ev.sum.alloc(cells), ev.late.alloc(cells);
  ev.sum.zero(), ev.late.zero();
  ev.y.upload(y_test, n_test);
  ev.cutp.alloc(std::max(1, ev.width));
  ev.partial.alloc(static_cast<size_t>(EVAL_BLOCKS) * EVAL_TERMS), ev.out.alloc(EVAL_TERMS);
  MYFM_CUDA(cudaDeviceSynchronize());
  *out = holder.release();
  MYFM_API_END
}
void myfm_evaluator_destroy(myfm_evaluator_t *e) { delete e; }
int myfm_evaluator_step(myfm_evaluator_t *e, myfm_trainer_t *t, int32_t iteration, const double *cutpoints,
                        int32_t n_cpt, double *terms) {
  MYFM_API_BEGIN
  require(e, "evaluator"), require(t, "trainer"), require(terms, "terms");
  if (n_cpt > 0)
    require(cutpoints, "cutpoints");
  t->impl->evaluate(e->state, iteration, cutpoints, n_cpt, terms);
  MYFM_API_END
}
int myfm_evaluator_get_sums(myfm_evaluator_t *e, double *sum, double *late) {
  MYFM_API_BEGIN
  require(e, "evaluator");
  MYFM_CUDA(cudaSetDevice(e->state.dataset->device));
  const size_t cells = static_cast<size_t>(e->state.n) * e->state.width;
  if (sum && cells)
    MYFM_CUDA(cudaMemcpy(sum, e->state.sum.p, cells * sizeof(double), cudaMemcpyDeviceToHost));
  if (late && cells)
    MYFM_CUDA(cudaMemcpy(late, e->state.late.p, cells * sizeof(double), cudaMemcpyDeviceToHost));
  MYFM_API_END
}

// kinds[i]: 0 = fresh standard normal (FMTrainer.hpp:122-125), 1 = fresh Gamma(shapes[i], 1)
// (FMTrainer.hpp:143,165); the first n_skip draws of one PERSISTENT normal_distribution are
// consumed beforehand (FM.hpp:34-45).
int myfm_rng_fill(int32_t dtype, int32_t seed, int64_t n_skip_normals_persistent, const int32_t *kinds,
                  const double *shapes, int64_t n, double *out) {
  MYFM_API_BEGIN
  auto run = [&](auto tag) {
    using Real = decltype(tag);
    MtStream<Real> s(seed);
    {
      std::vector<Real> sink(n_skip_normals_persistent + 1);
      if (n_skip_normals_persistent > 0) {
        Real w0;
        s.init_weights(sink.data(), n_skip_normals_persistent - 1, sink.data(), 0, &w0, Real(1));
      }
    }
    for (int64_t i = 0; i < n; i++)
      out[i] = kinds[i] == 0 ? s.normal() : s.gamma(static_cast<Real>(shapes[i]));
  };
  if (dtype == MYFM_DTYPE_F32)
    run(float{});
  else
    run(double{});
  MYFM_API_END
}

int myfm_host_transpose(const myfm_csr_t *X, int64_t *indptr_out, int32_t *indices_out, double *data_out) {
  MYFM_API_BEGIN
  require(X, "X"), require(indptr_out, "indptr_out");
  HostCs<double> t = host_transpose(host_from_api<double>(*X, "X"));
  for (size_t i = 0; i < t.ptr.size(); i++)
    indptr_out[i] = t.ptr[i];
  for (size_t i = 0; i < t.idx.size(); i++)
    indices_out[i] = t.idx[i], data_out[i] = t.val[i];
  MYFM_API_END
}
int myfm_set_host_threads(int32_t n) {
  MYFM_API_BEGIN
  host_threads_override().store(n > 0 ? n : 0);
  MYFM_API_END
}

int myfm_mt_jump_taps(uint64_t n, uint16_t *taps_out, int32_t capacity, int32_t *n_taps) {
  MYFM_API_BEGIN
  require(n_taps, "n_taps");
  const std::vector<uint16_t> &taps = mt_jump_taps(n);
  *n_taps = static_cast<int32_t>(taps.size());
  if (taps_out && capacity >= *n_taps)
    std::memcpy(taps_out, taps.data(), taps.size() * sizeof(uint16_t));
  MYFM_API_END
}

int myfm_nccl_unique_id(void *out128) {
  MYFM_API_BEGIN
  require(out128, "out128");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi &nccl = NcclApi::get();
  ncclUniqueId id;
  nccl.check(nccl.GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, &id, sizeof(id));
  MYFM_API_END
}

int myfm_level_relax(const myfm_csr_t *X, int32_t *level, int32_t *n_levels, int32_t *changed) {
  MYFM_API_BEGIN
  require(X, "X"), require(n_levels, "n_levels"), require(changed, "changed");
  HostCs<float> Xt = host_transpose(host_from_api<float>(*X, "X"));
  if (Xt.n_major > 0)
    require(level, "level");
  int nl = 0;
  std::vector<int> lv = compute_levels(Xt, &nl, level);
  *changed = 0;
  for (size_t j = 0; j < lv.size(); j++) {
    if (lv[j] != level[j])
      *changed = 1;
    level[j] = lv[j];
  }
  *n_levels = nl;
  MYFM_API_END
}

int myfm_level_schedule(const myfm_csr_t *X, int32_t *level, int32_t *n_levels) {
  MYFM_API_BEGIN
  require(X, "X"), require(n_levels, "n_levels");
  HostCs<float> Xh = host_from_api<float>(*X, "X");
  int nl = 0;
  std::vector<int> lv;
  if (!compute_levels_by_rows(Xh, lv, &nl)) // what the trainer does: row-parallel when rows are sorted
    lv = compute_levels(host_transpose(Xh), &nl);
  for (size_t j = 0; j < lv.size(); j++)
    level[j] = lv[j];
  *n_levels = nl;
  MYFM_API_END
}

} // extern "C"
