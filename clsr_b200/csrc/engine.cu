// clsr_b200 engine: host-side orchestration of one CLSR step on one B200 and the C ABI of
// include/clsr_b200.h.  Replaces the TensorFlow session executor under
// CLSRModel.train / eval / infer (clsr.py:383-408, base_model.py:366-392).
//
// The step is a fixed sequence of kernel launches on one stream over a workspace allocated at
// creation; no allocation, no host synchronisation until the scalar losses are read back.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/clsr_b200.h"
#include "attn.cuh"
#include "embed.cuh"
#include "gemm.cuh"
#include "head.cuh"
#include "head_mlp.cuh"
#include "rnn.cuh"
#include "rnn_tc.cuh"
#include "tc_gemm.cuh"

using namespace clsr;

namespace {

thread_local std::string g_create_error;

struct DenseEntry {
  std::string name;
  long long off;
  long long n;
  int rows, cols;
  int trainable;
};

struct BnLayer {
  int N;
  long long gamma, beta, mmean, mvar;  // offsets into P
  double* stat_f;                      // [2N] forward sums
  double* stat_b;                      // [2N] backward sums
  float *scale, *shift, *mean, *rstd, *al, *be, *ga;
};

struct Mlp {  // one _fcn_net: two hidden layers with BN + ReLU and a scalar output unit
  long long w0, b0, w1, b1, wo, bo;  // offsets into P
  int in, n0, n1;
  BnLayer bn0, bn1;
};

}  // namespace

struct clsr_engine {
  clsr_config cfg;
  std::string err;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t aux[2] = {nullptr, nullptr};   // side streams: the three recurrences run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr}, ev_memset = nullptr;
  bool debug_sync = false;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char*> prof_names;
  size_t prof_used = 0;
  std::map<std::string, std::pair<double, long long>> prof_agg;
  long long launches = 0;
  long long adam_step = 0;
  // fused _fcn_net kernels (head_mlp.cuh): grid-barrier counter, rows per CTA the shared-memory layouts are sized for
  unsigned long long* coop_bar = nullptr;
  int coop_rpc = 0;
  bool coop_alpha = false, coop_logit = false;
  float* d_lr = nullptr;    // [1] this step's Adam step size (written by set_scalar_kernel before the step body)
  // Captured training steps (single GPU): one executable graph per (S, B, G, flags) shape.  The first step of a
  // shape runs eagerly (lazy allocations, attribute calls), the second is captured, later ones are one graph launch.
  struct StepGraph {
    int S, B, G;
    uint32_t flags;
    int seen;
    long long launches;
    cudaGraphExec_t exec;
  };
  std::vector<StepGraph> graphs;
  bool graphs_on = true;
  long long graph_replays = 0;
  // The legacy default stream cannot be captured: an engine bound to it records and launches its step graphs on this
  // (blocking) stream and orders it against the default stream with two events per step.
  cudaStream_t gstream = nullptr;
  cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;
  int num_sms = 148;
  bool gather_attr_set = false;
  int rnn_wglob = 0;   // SIMT recurrences read their weights from global memory (they do not fit shared memory)
  int smem_optin = 49152;
  int tc_smem_max = 49152;
  int tc_dw_smem_max = 49152;
  // pending problems of a grouped weight-gradient launch (tc_dw_group_kernel)
  bool dw_grouping = false;
  tc::DwGroup dwg;
  long long dwg_weight[tc::kDwGroupMax];
  int dwg_tiles[tc::kDwGroupMax];
  int dwg_smem = 0;

  int T, Di, Dc, D, U, H, Q, A0, A1, L0, L1, CA, NX;
  // graph variants (clsr.py:159-274): which optional blocks exist
  bool has_gru1 = true;    // short_term_intention GRU (interest_evolve)
  bool has_gru2 = true;    // causal2 GRU (predict_long_short and not manual_alpha)
  bool has_alpha = true;   // fcn_alpha MLP (not manual_alpha)
  int Hf = 0;              // width of the causal2 final state at the front of concat_all (H or 0)
  bool plain_lstm = false; // sequential_model = 'lstm': the Time4LSTM kernels with both time gates pinned at 1 and no o-gate time terms
  int oG1, oC1, oG2, oC2, oL, oO, oTN, oTL;
  int Bmax, Smax;

  // dense variables
  std::vector<DenseEntry> dense;
  std::map<std::string, int> dense_ix;
  long long Ptot = 0;
  float *P = nullptr, *Pm = nullptr, *Pv = nullptr, *Pg = nullptr;
  DenseVar* d_vars = nullptr;
  float* d_norms = nullptr;
  // folded / transposed weights and their gradients
  std::map<std::string, long long> wd_off;
  long long Wtot = 0;
  float *Wd = nullptr, *dWd = nullptr;
  BlockOp *ops_prep = nullptr, *ops_unprep = nullptr;
  int n_prep = 0, n_unprep = 0;

  Mlp mlp_long, mlp_short, mlp_alpha, mlp_logit;
  long long p_wattl, p_watts;
  long long p_tw1, p_tb1, p_tw2, p_tb2;

  // tables
  float* tab[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  float* tab_m[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  float* tab_v[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  long long tab_rows[CLSR_NUM_TABLES];
  int tab_dim[CLSR_NUM_TABLES];
  int32_t* slot[3] = {nullptr, nullptr, nullptr};  // item, cate, user
  int32_t* uniq[3] = {nullptr, nullptr, nullptr};
  long long uniq_cap[3];
  float* cg[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  int32_t* counts = nullptr;  // [0] contrastive rows, [1] uniq items, [2] uniq cates, [3] uniq users
  double* sumsq = nullptr;    // [4]
  double* acc = nullptr;      // [16]
  float* d_losses = nullptr;  // [16]: 5 losses, 4 table-gradient norms
  float* h_losses = nullptr;  // pinned
  unsigned long long* d_clip_steps = nullptr;  // [1] shared-history steps with an active table clip

  // staged inputs
  int32_t *in_users, *in_items, *in_cates, *in_ih, *in_ch, *in_mask, *d_len;
  float *in_tfa, *in_ttn, *in_labels;
  char* h_stage = nullptr;  // pinned staging
  size_t h_stage_bytes = 0;
  char* in_raw = nullptr;   // device landing block of feeds copied straight from pinned caller memory (stage_inputs)
  cudaEvent_t h2d_done = nullptr;  // the staging buffer may be rewritten once this has fired
  float* h_out = nullptr;  // pinned [2*Bmax]
  int staged_S = 0, staged_G = 1;   // shape of the batch clsr_build_batch left in the staged-feed block
  int32_t* d_err = nullptr;   // [1] id-range flags of the last device-resident feed (stage_device_feed_kernel)
  int32_t* h_err = nullptr;   // pinned mirror, copied after every step

  // data-parallel state (NCCL is dlopen'ed by clsr_comm_init)
  int world = 1, rank = 0;
  void* nccl_lib = nullptr;
  void* comm = nullptr;
  int (*ncclAllReduce_)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ncclAllGather_)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*ncclCommDestroy_)(void*) = nullptr;
  const char* (*ncclGetErrorString_)(int) = nullptr;
  // peer-memory communication (one-shot all-reduce / barrier kernels, embed.cuh) and row-sharded tables
  bool peer_ready = false;
  PeerComm pc;
  unsigned long long peer_seq = 0;
  double* comm_data = nullptr;
  unsigned long long* comm_flags = nullptr;
  std::vector<void*> peer_opened;               // cudaIpcOpenMemHandle'd pointers (closed at destroy)
  bool sharded = false;
  long long tab_local_rows[CLSR_NUM_TABLES] = {0, 0, 0, 0};
  float* sh_val[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};   // engine-owned shards (sharded mode)
  float* sh_m[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  float* sh_v[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  float* sh_g[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  int32_t* sh_touched[CLSR_NUM_TABLES] = {nullptr, nullptr, nullptr, nullptr};
  GradView gview[CLSR_NUM_TABLES];
  TabView tview[CLSR_NUM_TABLES];
  int32_t *g_ih = nullptr, *g_ch = nullptr, *g_items = nullptr, *g_cates = nullptr, *g_users = nullptr;
  int32_t *l_ih = nullptr, *l_ch = nullptr, *l_users = nullptr;
  float *g_dX = nullptr, *g_dtgt = nullptr, *g_dul = nullptr, *g_dus = nullptr;

  // workspace
  std::vector<void*> allocs;
  std::map<std::string, std::pair<float*, long long>> bufs;
  long long ws_bytes = 0;

  float* B(const char* n) { return bufs.at(n).first; }
};

namespace {

int fail(clsr_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _c = (call);                                                                       \
    if (_c != cudaSuccess)                                                                         \
      return fail(e, CLSR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_c), __FILE__, \
                  __LINE__);                                                                       \
  } while (0)

// Per-kernel device timing: one event after every launch on the engine stream; the interval
// between consecutive events is the kernel's duration (the stream serialises them).
void prof_mark(clsr_engine* e, const char* name) {
  if (e->prof_used == e->prof_events.size()) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    e->prof_events.push_back(ev);
    e->prof_names.push_back(name);
  }
  e->prof_names[e->prof_used] = name;
  cudaEventRecord(e->prof_events[e->prof_used], e->stream);
  e->prof_used++;
}
#define MARK(name) do { if (e->profiling) prof_mark(e, name); } while (0)

// Launch bookkeeping: count, optional sync-and-check per kernel (debug).
#define POST(name)                                                                                 \
  do {                                                                                             \
    e->launches++;                                                                                 \
    if (e->profiling) prof_mark(e, name);                                                          \
    cudaError_t _c = cudaGetLastError();                                                           \
    if (_c == cudaSuccess && e->debug_sync) _c = cudaStreamSynchronize(e->stream);                 \
    if (_c != cudaSuccess)                                                                         \
      return fail(e, CLSR_ERR_CUDA, "kernel %s failed: %s (%s:%d)", name, cudaGetErrorString(_c),  \
                  __FILE__, __LINE__);                                                             \
  } while (0)

// Fork the two side streams off the engine stream / join them back (the recurrent kernels are
// latency-bound with few resident warps; running the three of them side by side fills the SMs).
int fork_aux(clsr_engine* e) {
  CK(cudaEventRecord(e->ev_fork, e->stream));
  for (int i = 0; i < 2; ++i) CK(cudaStreamWaitEvent(e->aux[i], e->ev_fork, 0));
  return 0;
}
int join_aux(clsr_engine* e, const char* name) {
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventRecord(e->ev_join[i], e->aux[i]));
    CK(cudaStreamWaitEvent(e->stream, e->ev_join[i], 0));
  }
  cudaError_t c = cudaGetLastError();
  if (c == cudaSuccess && e->debug_sync) {
    for (int i = 0; i < 2 && c == cudaSuccess; ++i) c = cudaStreamSynchronize(e->aux[i]);
    if (c == cudaSuccess) c = cudaStreamSynchronize(e->stream);
  }
  if (c != cudaSuccess) return fail(e, CLSR_ERR_CUDA, "%s failed: %s", name, cudaGetErrorString(c));
  MARK(name);
  return 0;
}

enum { kNcclInt32 = 2, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };

// One exchange over peer memory for up to three small arrays (doubles, doubles, int32); n = 0 everywhere = barrier.
int peer_reduce(clsr_engine* e, double* a, int na, double* b, int nb, int32_t* c, int nc) {
  if (e->world <= 1) return 0;
  if (!e->peer_ready) return fail(e, CLSR_ERR_STATE, "peer communication not set up (clsr_peer_setup_*)");
  if (na + nb + nc > kPeerSlots) return fail(e, CLSR_ERR_ARG, "peer_reduce: %d values exceed %d", na + nb + nc, kPeerSlots);
  peer_allreduce_kernel<<<1, 256, 0, e->stream>>>(e->pc, ++e->peer_seq, a, na, b, nb, c, nc);
  POST(na + nb + nc ? "peer_allreduce" : "peer_barrier");
  return 0;
}

int allreduce(clsr_engine* e, void* buf, size_t count, int dtype) {
  if (e->world <= 1) return 0;
  if (e->peer_ready && count <= (size_t)kPeerSlots && (dtype == kNcclFloat64 || dtype == kNcclInt32))
    return dtype == kNcclFloat64 ? peer_reduce(e, (double*)buf, (int)count, nullptr, 0, nullptr, 0)
                                 : peer_reduce(e, nullptr, 0, nullptr, 0, (int32_t*)buf, (int)count);
  int r = e->ncclAllReduce_(buf, buf, count, dtype, kNcclSum, e->comm, e->stream);
  if (r != 0) return fail(e, CLSR_ERR_NCCL, "ncclAllReduce failed: %s", e->ncclGetErrorString_(r));
  MARK("nccl_allreduce");
  return 0;
}
int allgather(clsr_engine* e, const void* src, void* dst, size_t count, int dtype) {
  int r = e->ncclAllGather_(src, dst, count, dtype, e->comm, e->stream);
  if (r != 0) return fail(e, CLSR_ERR_NCCL, "ncclAllGather failed: %s", e->ncclGetErrorString_(r));
  MARK("nccl_allgather");
  return 0;
}

template <typename Tp>
int dalloc(clsr_engine* e, Tp** p, long long n, bool zero = true) {
  size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(Tp);
  bytes = (bytes + 255) & ~(size_t)255;
  CK(cudaMalloc((void**)p, bytes));
  e->allocs.push_back((void*)*p);
  e->ws_bytes += (long long)bytes;
  if (zero) CK(cudaMemset(*p, 0, bytes));
  return 0;
}

int fbuf(clsr_engine* e, const char* name, long long n) {
  float* p = nullptr;
  int rc = dalloc(e, &p, n);
  if (rc) return rc;
  e->bufs[name] = std::make_pair(p, n);
  return 0;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int grid1d(clsr_engine* e, long long n, int block, int per_sm = 8) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)e->num_sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// item rows pre-summed per CTA in shared memory by the scatter-add (16 KB of rows, at most 64)
int scatter_hot_rows(const clsr_engine* e) {
  int hot = (16 * 1024) / (e->Di * 4);
  if (hot > 64) hot = 64;
  if ((long long)hot > e->tab_rows[0]) hot = (int)e->tab_rows[0];
  return hot < 1 ? 1 : hot;
}

// K1 + K3: history gather.  Bulk-copy (TMA) kernel when two [128, D] tiles fit shared memory (always for the
// model widths in use); CLSR_GATHER_LDG=1 selects the register-path kernel (A/B).
int launch_gather_hist(clsr_engine* e, const int32_t* ih, const int32_t* ch, int seq_stride, float* out, long long npos) {
  static const bool ldg = getenv("CLSR_GATHER_LDG") != nullptr;
  const int D = e->D;
  int hot = (8 * 1024) / (e->Di * 4);   // lowest item ids kept in shared memory per CTA (8 KB)
  if (hot > 64) hot = 64;
  if ((long long)hot > e->tab_rows[0]) hot = (int)e->tab_rows[0];
  if (hot < 1) hot = 1;
  const size_t smem = 2 * (size_t)kGatherTile * D * 4 + ((size_t)hot * e->Di + ((e->Dc + 3) & ~3)) * 4 + 64;
  if (!ldg && smem <= (size_t)e->smem_optin - 1024) {
    int per_sm = (int)((size_t)(e->smem_optin) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long long tiles = (npos + kGatherTile - 1) / kGatherTile;
    long long grid = (long long)e->num_sms * per_sm;
    if (grid > tiles) grid = tiles;
    if (!e->gather_attr_set) {
      CK(cudaFuncSetAttribute(gather_hist_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      e->gather_attr_set = true;
    }
    gather_hist_tma_kernel<<<(int)grid, kGatherTile, smem, e->stream>>>(ih, ch, seq_stride, e->T, e->tview[CLSR_TABLE_ITEM],
                                                                      e->tview[CLSR_TABLE_CATE], e->Di, e->Dc, hot, out, npos);
  } else {
    const long long nvec = npos * (D / 4);
    gather_hist_kernel<4><<<grid1d(e, cdiv(nvec, 4), 256, 8), 256, 0, e->stream>>>(
        ih, ch, seq_stride, e->T, e->tview[CLSR_TABLE_ITEM], e->tview[CLSR_TABLE_CATE], e->Di, e->Dc, out, npos);
  }
  POST("gather_hist");
  return 0;
}

// ---- dense variable inventory (mirrors clsr_b200/params.py::dense_spec) ---------------------------
void add_dense(clsr_engine* e, const std::string& name, int rows, int cols, int trainable) {
  DenseEntry d;
  d.name = name;
  d.off = e->Ptot;
  d.rows = rows;
  d.cols = cols;
  d.n = (long long)rows * cols;
  d.trainable = trainable;
  e->dense_ix[name] = (int)e->dense.size();
  e->dense.push_back(d);
  e->Ptot += (d.n + 3) & ~3LL;  // keep every variable 16-byte aligned
}
long long poff(clsr_engine* e, const std::string& name) { return e->dense[e->dense_ix.at(name)].off; }

void add_fcn(clsr_engine* e, const std::string& scope, int in, int n0, int n1, Mlp* m) {
  int sizes[2] = {n0, n1};
  int last = in;
  for (int i = 0; i < 2; ++i) {
    std::string li = std::to_string(i);
    add_dense(e, scope + "w_nn_layer" + li, last, sizes[i], 1);
    add_dense(e, scope + "b_nn_layer" + li, 1, sizes[i], 1);
    std::string bn = scope + (i == 0 ? "batch_normalization/" : "batch_normalization_1/");
    add_dense(e, bn + "gamma", 1, sizes[i], 1);
    add_dense(e, bn + "beta", 1, sizes[i], 1);
    add_dense(e, bn + "moving_mean", 1, sizes[i], 0);
    add_dense(e, bn + "moving_variance", 1, sizes[i], 0);
    last = sizes[i];
  }
  add_dense(e, scope + "w_nn_output", last, 1, 1);
  add_dense(e, scope + "b_nn_output", 1, 1, 1);
  m->in = in; m->n0 = n0; m->n1 = n1;
  m->w0 = poff(e, scope + "w_nn_layer0"); m->b0 = poff(e, scope + "b_nn_layer0");
  m->w1 = poff(e, scope + "w_nn_layer1"); m->b1 = poff(e, scope + "b_nn_layer1");
  m->wo = poff(e, scope + "w_nn_output"); m->bo = poff(e, scope + "b_nn_output");
  BnLayer* bl[2] = {&m->bn0, &m->bn1};
  for (int i = 0; i < 2; ++i) {
    std::string bn = scope + (i == 0 ? "batch_normalization/" : "batch_normalization_1/");
    bl[i]->N = sizes[i];
    bl[i]->gamma = poff(e, bn + "gamma"); bl[i]->beta = poff(e, bn + "beta");
    bl[i]->mmean = poff(e, bn + "moving_mean"); bl[i]->mvar = poff(e, bn + "moving_variance");
  }
}

const char* kSC = "sequential/clsr/";

void build_inventory(clsr_engine* e) {
  const int D = e->D, U = e->U, H = e->H, Q = e->Q;
  std::string sc = kSC;
  std::string lt = sc + "long_term/attention_fcn/";
  add_dense(e, lt + "attention_mat", D, U, 1);
  add_fcn(e, lt + "att_fcn/nn_part/", 4 * U, e->A0, e->A1, &e->mlp_long);
  std::string st = sc + "short_term/";
  add_dense(e, st + "attention_fcn/attention_mat", H, Q, 1);
  add_fcn(e, st + "attention_fcn/att_fcn/nn_part/", 4 * Q, e->A0, e->A1, &e->mlp_short);
  const std::string gs[2] = {st + "short_term_intention/gru_cell/", sc + "causal2/causal2/gru_cell/"};
  const int gu[2] = {U, H};
  for (int i = 0; i < 2; ++i) {
    if (!(i == 0 ? e->has_gru1 : e->has_gru2)) continue;   // the reference graph does not create these variables then
    add_dense(e, gs[i] + "gates/kernel", D + gu[i], 2 * gu[i], 1);
    add_dense(e, gs[i] + "gates/bias", 1, 2 * gu[i], 1);
    add_dense(e, gs[i] + "candidate/kernel", D + gu[i], gu[i], 1);
    add_dense(e, gs[i] + "candidate/bias", 1, gu[i], 1);
  }
  std::string tl = st + (e->plain_lstm ? "simple_lstm/lstm_cell/" : "time4lstm/time4lstm_cell/");
  add_dense(e, tl + "kernel", D + H, 4 * H, 1);
  add_dense(e, tl + "bias", 1, 4 * H, 1);
  if (!e->plain_lstm) {
    const char* v1[] = {"_time_input_w1", "_time_input_bias1", "_time_input_w2", "_time_input_bias2",
                        "_time_bias1", "_time_bias2"};
    for (const char* n : v1) add_dense(e, tl + n, 1, H, 1);
    add_dense(e, tl + "_time_kernel_w1", D, H, 1);
    add_dense(e, tl + "_time_kernel_w2", D, H, 1);
    const char* v2[] = {"_time_kernel_t1", "_time_kernel_t2", "_o_kernel_t1", "_o_kernel_t2"};
    for (const char* n : v2) add_dense(e, tl + n, H, H, 1);
  }
  if (e->has_alpha) add_fcn(e, sc + "fcn_alpha/nn_part/", e->CA, e->A0, e->A1, &e->mlp_alpha);
  else { memset(&e->mlp_alpha, 0, sizeof e->mlp_alpha); }
  add_fcn(e, "sequential/logit_fcn/nn_part/", H + D, e->L0, e->L1, &e->mlp_logit);
  e->p_wattl = poff(e, lt + "attention_mat");
  e->p_watts = poff(e, st + "attention_fcn/attention_mat");
  if (!e->plain_lstm) {
    e->p_tw1 = poff(e, tl + "_time_input_w1"); e->p_tb1 = poff(e, tl + "_time_input_bias1");
    e->p_tw2 = poff(e, tl + "_time_input_w2"); e->p_tb2 = poff(e, tl + "_time_input_bias2");
  } else {
    e->p_tw1 = e->p_tb1 = e->p_tw2 = e->p_tb2 = 0;
  }
}

long long wd_add(clsr_engine* e, const char* name, long long n) {
  long long o = e->Wtot;
  e->wd_off[name] = o;
  e->Wtot += (n + 3) & ~3LL;
  return o;
}

BlockOp mk(long long dst, int ldd, long long s1, int lds1, int rows, int cols, float c1 = 1.f,
           long long s2 = -1, int lds2 = 0, float c2 = 0.f, int transpose = 0) {
  BlockOp o;
  o.dst = dst; o.src1 = s1; o.src2 = s2; o.rows = rows; o.cols = cols; o.ldd = ldd; o.lds1 = lds1;
  o.lds2 = lds2; o.c1 = c1; o.c2 = c2; o.transpose = transpose; o.accumulate = 0;
  return o;
}

// Folded weights (forward) and the inverse map of their gradients onto the TF variables.
int build_weight_maps(clsr_engine* e) {
  const int D = e->D, U = e->U, H = e->H, Q = e->Q, A0 = e->A0, A1 = e->A1, NX = e->NX, CA = e->CA;
  const int L0 = e->L0, L1 = e->L1;
  std::vector<BlockOp> prep, unprep;
  std::string sc = kSC, st = sc + "short_term/";
  std::string tl = st + (e->plain_lstm ? "simple_lstm/lstm_cell/" : "time4lstm/time4lstm_cell/");
  auto W = [&](const char* n) { return e->wd_off.at(n); };

  // ---- long attention: W0 rows [a | b | c | d], each U rows: feat = [a, q, a-q, a*q] ----
  {
    long long w0 = e->mlp_long.w0;
    long long wl0 = wd_add(e, "Wl0", 2LL * U * A0), wl0q = wd_add(e, "Wl0q", (long long)U * A0);
    long long wl0T = wd_add(e, "Wl0T", 2LL * U * A0), wl0qT = wd_add(e, "Wl0qT", (long long)U * A0);
    long long wattT = wd_add(e, "WattlT", (long long)U * D), w1T = wd_add(e, "W1lT", (long long)A1 * A0);
    prep.push_back(mk(wl0, A0, w0, A0, U, A0, 1.f, w0 + 2LL * U * A0, A0, 1.f));          // a + c
    prep.push_back(mk(wl0 + (long long)U * A0, A0, w0 + 3LL * U * A0, A0, U, A0));          // d
    prep.push_back(mk(wl0q, A0, w0 + (long long)U * A0, A0, U, A0, 1.f, w0 + 2LL * U * A0, A0, -1.f));
    prep.push_back(mk(wl0T, 2 * U, w0, A0, U, A0, 1.f, w0 + 2LL * U * A0, A0, 1.f, 1));
    prep.push_back(mk(wl0T + U, 2 * U, w0 + 3LL * U * A0, A0, U, A0, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(wl0qT, U, w0 + (long long)U * A0, A0, U, A0, 1.f, w0 + 2LL * U * A0, A0, -1.f, 1));
    prep.push_back(mk(wattT, D, e->p_wattl, U, D, U, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(w1T, A0, e->mlp_long.w1, A1, A0, A1, 1.f, -1, 0, 0.f, 1));
    // gradients: dW0[a] = dWl0[0:U]; dW0[b] = dWl0q; dW0[c] = dWl0[0:U] - dWl0q; dW0[d] = dWl0[U:2U]
    unprep.push_back(mk(w0, A0, wl0, A0, U, A0));
    unprep.push_back(mk(w0 + (long long)U * A0, A0, wl0q, A0, U, A0));
    unprep.push_back(mk(w0 + 2LL * U * A0, A0, wl0, A0, U, A0, 1.f, wl0q, A0, -1.f));
    unprep.push_back(mk(w0 + 3LL * U * A0, A0, wl0 + (long long)U * A0, A0, U, A0));
  }
  // ---- short attention: W0 rows [a | b | c | d], each Q rows; d splits into sti (U) / target (D) ----
  {
    long long w0 = e->mlp_short.w0;
    long long wi = wd_add(e, "Ws0i", (long long)(Q + U) * A0), wt = wd_add(e, "Ws0t", (long long)D * A0);
    long long wq = wd_add(e, "Ws0q", (long long)Q * A0);
    long long wiT = wd_add(e, "Ws0iT", (long long)(Q + U) * A0), wtT = wd_add(e, "Ws0tT", (long long)D * A0);
    long long wqT = wd_add(e, "Ws0qT", (long long)Q * A0);
    long long wattT = wd_add(e, "WattsT", (long long)Q * H), w1T = wd_add(e, "W1sT", (long long)A1 * A0);
    long long a = w0, b = w0 + (long long)Q * A0, c = w0 + 2LL * Q * A0, d = w0 + 3LL * Q * A0;
    prep.push_back(mk(wi, A0, a, A0, Q, A0, 1.f, c, A0, 1.f));
    prep.push_back(mk(wi + (long long)Q * A0, A0, d, A0, U, A0));
    prep.push_back(mk(wt, A0, d + (long long)U * A0, A0, D, A0));
    prep.push_back(mk(wq, A0, b, A0, Q, A0, 1.f, c, A0, -1.f));
    prep.push_back(mk(wiT, Q + U, a, A0, Q, A0, 1.f, c, A0, 1.f, 1));
    prep.push_back(mk(wiT + Q, Q + U, d, A0, U, A0, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(wtT, D, d + (long long)U * A0, A0, D, A0, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(wqT, Q, b, A0, Q, A0, 1.f, c, A0, -1.f, 1));
    prep.push_back(mk(wattT, H, e->p_watts, Q, H, Q, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(w1T, A0, e->mlp_short.w1, A1, A0, A1, 1.f, -1, 0, 0.f, 1));
    unprep.push_back(mk(a, A0, wi, A0, Q, A0));
    unprep.push_back(mk(b, A0, wq, A0, Q, A0));
    unprep.push_back(mk(c, A0, wi, A0, Q, A0, 1.f, wq, A0, -1.f));
    unprep.push_back(mk(d, A0, wi + (long long)Q * A0, A0, U, A0));
    unprep.push_back(mk(d + (long long)U * A0, A0, wt, A0, D, A0));
  }
  // ---- recurrences: input halves of every kernel side by side in Wx_all [D, NX] ----
  {
    // Wx_all [D + 2H, NX]: rows D.. hold [0 | Wt] under the last 3H columns, so that PX[:, oO:] = [X | TNL] . Wx_all[:, oO:]
    // is ONE product (the rows above are zero-initialised and never written)
    long long wx = wd_add(e, "Wx_all", (long long)(D + 2 * H) * NX), bx = wd_add(e, "bx_all", NX);
    long long wxT = wd_add(e, "Wx_allT", (long long)D * NX);
    long long wt = wd_add(e, "Wt", 2LL * H * 3 * H), wtT = wd_add(e, "WtT", 2LL * H * 3 * H);
    const std::string gs[2] = {st + "short_term_intention/gru_cell/", sc + "causal2/causal2/gru_cell/"};
    const int gu[2] = {U, H};
    const int og[2] = {e->oG1, e->oG2}, oc[2] = {e->oC1, e->oC2};
    const char* ng[2][4] = {{"Wgh1", "Wch1", "Wgh1T", "Wch1T"}, {"Wgh2", "Wch2", "Wgh2T", "Wch2T"}};
    for (int i = 0; i < 2; ++i) {
      if (!(i == 0 ? e->has_gru1 : e->has_gru2)) continue;   // its PX columns stay zero and are never read
      int u = gu[i];
      long long gk = poff(e, gs[i] + "gates/kernel"), gb = poff(e, gs[i] + "gates/bias");
      long long ck = poff(e, gs[i] + "candidate/kernel"), cb = poff(e, gs[i] + "candidate/bias");
      long long wgh = wd_add(e, ng[i][0], (long long)u * 2 * u), wch = wd_add(e, ng[i][1], (long long)u * u);
      long long wghT = wd_add(e, ng[i][2], (long long)u * 2 * u), wchT = wd_add(e, ng[i][3], (long long)u * u);
      prep.push_back(mk(wx + og[i], NX, gk, 2 * u, D, 2 * u));
      prep.push_back(mk(wx + oc[i], NX, ck, u, D, u));
      prep.push_back(mk(wxT + (long long)og[i] * D, D, gk, 2 * u, D, 2 * u, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wxT + (long long)oc[i] * D, D, ck, u, D, u, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(bx + og[i], NX, gb, 2 * u, 1, 2 * u));
      prep.push_back(mk(bx + oc[i], NX, cb, u, 1, u));
      prep.push_back(mk(wgh, 2 * u, gk + (long long)D * 2 * u, 2 * u, u, 2 * u));
      prep.push_back(mk(wch, u, ck + (long long)D * u, u, u, u));
      prep.push_back(mk(wghT, u, gk + (long long)D * 2 * u, 2 * u, u, 2 * u, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wchT, u, ck + (long long)D * u, u, u, u, 1.f, -1, 0, 0.f, 1));
      unprep.push_back(mk(gk, 2 * u, wx + og[i], NX, D, 2 * u));
      unprep.push_back(mk(ck, u, wx + oc[i], NX, D, u));
      unprep.push_back(mk(gb, 2 * u, bx + og[i], NX, 1, 2 * u));
      unprep.push_back(mk(cb, u, bx + oc[i], NX, 1, u));
      unprep.push_back(mk(gk + (long long)D * 2 * u, 2 * u, wgh, 2 * u, u, 2 * u));
      unprep.push_back(mk(ck + (long long)D * u, u, wch, u, u, u));
    }
    long long lk = poff(e, tl + "kernel"), lb = poff(e, tl + "bias");
    long long km = wd_add(e, "Km", (long long)H * 4 * H), kmT = wd_add(e, "KmT", (long long)H * 4 * H);
    prep.push_back(mk(wx + e->oL, NX, lk, 4 * H, D, 4 * H));
    prep.push_back(mk(wxT + (long long)e->oL * D, D, lk, 4 * H, D, 4 * H, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(bx + e->oL, NX, lb, 4 * H, 1, 4 * H));
    prep.push_back(mk(km, 4 * H, lk + (long long)D * 4 * H, 4 * H, H, 4 * H));
    prep.push_back(mk(kmT, H, lk + (long long)D * 4 * H, 4 * H, H, 4 * H, 1.f, -1, 0, 0.f, 1));
    unprep.push_back(mk(lk, 4 * H, wx + e->oL, NX, D, 4 * H));
    unprep.push_back(mk(lk + (long long)D * 4 * H, 4 * H, km, 4 * H, H, 4 * H));
    unprep.push_back(mk(lb, 4 * H, bx + e->oL, NX, 1, 4 * H));
    if (!e->plain_lstm) {
      long long k1 = poff(e, tl + "_time_kernel_w1"), k2 = poff(e, tl + "_time_kernel_w2");
      long long tb1 = poff(e, tl + "_time_bias1"), tb2 = poff(e, tl + "_time_bias2");
      long long t1 = poff(e, tl + "_time_kernel_t1"), t2 = poff(e, tl + "_time_kernel_t2");
      long long o1 = poff(e, tl + "_o_kernel_t1"), o2 = poff(e, tl + "_o_kernel_t2");
      prep.push_back(mk(wx + e->oTN, NX, k1, H, D, H));
      prep.push_back(mk(wx + e->oTL, NX, k2, H, D, H));
      prep.push_back(mk(wxT + (long long)e->oTN * D, D, k1, H, D, H, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wxT + (long long)e->oTL * D, D, k2, H, D, H, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(bx + e->oTN, NX, tb1, H, 1, H));
      prep.push_back(mk(bx + e->oTL, NX, tb2, H, 1, H));
      // Wt [2H, 3H]: rows [tanh-now | tanh-last], cols [o | time-now gate | time-last gate]
      prep.push_back(mk(wt, 3 * H, o1, H, H, H));
      prep.push_back(mk(wt + (long long)H * 3 * H, 3 * H, o2, H, H, H));
      prep.push_back(mk(wt + H, 3 * H, t1, H, H, H));
      prep.push_back(mk(wt + (long long)H * 3 * H + 2 * H, 3 * H, t2, H, H, H));
      {
        const long long wxe = wx + (long long)D * NX + e->oO;   // same four blocks, row pitch NX
        prep.push_back(mk(wxe, NX, o1, H, H, H));
        prep.push_back(mk(wxe + (long long)H * NX, NX, o2, H, H, H));
        prep.push_back(mk(wxe + H, NX, t1, H, H, H));
        prep.push_back(mk(wxe + (long long)H * NX + 2 * H, NX, t2, H, H, H));
      }
      prep.push_back(mk(wtT, 2 * H, o1, H, H, H, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wtT + H, 2 * H, o2, H, H, H, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wtT + (long long)H * 2 * H, 2 * H, t1, H, H, H, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(wtT + 2LL * H * 2 * H + H, 2 * H, t2, H, H, H, 1.f, -1, 0, 0.f, 1));
      unprep.push_back(mk(k1, H, wx + e->oTN, NX, D, H));
      unprep.push_back(mk(k2, H, wx + e->oTL, NX, D, H));
      unprep.push_back(mk(tb1, H, bx + e->oTN, NX, 1, H));
      unprep.push_back(mk(tb2, H, bx + e->oTL, NX, 1, H));
      unprep.push_back(mk(o1, H, wt, 3 * H, H, H));
      unprep.push_back(mk(o2, H, wt + (long long)H * 3 * H, 3 * H, H, H));
      unprep.push_back(mk(t1, H, wt + H, 3 * H, H, H));
      unprep.push_back(mk(t2, H, wt + (long long)H * 3 * H + 2 * H, 3 * H, H, H));
    }
  }
  // ---- alpha / logit MLPs: transposed copies for the data-gradient GEMMs ----
  {
    long long a0T = wd_add(e, "Wa0T", (long long)CA * A0), a1T = wd_add(e, "Wa1T", (long long)A0 * A1);
    long long g0T = wd_add(e, "Wg0T", (long long)(H + D) * L0), g1T = wd_add(e, "Wg1T", (long long)L0 * L1);
    if (e->has_alpha) {
      prep.push_back(mk(a0T, CA, e->mlp_alpha.w0, A0, CA, A0, 1.f, -1, 0, 0.f, 1));
      prep.push_back(mk(a1T, A0, e->mlp_alpha.w1, A1, A0, A1, 1.f, -1, 0, 0.f, 1));
    }
    prep.push_back(mk(g0T, H + D, e->mlp_logit.w0, L0, H + D, L0, 1.f, -1, 0, 0.f, 1));
    prep.push_back(mk(g1T, L0, e->mlp_logit.w1, L1, L0, L1, 1.f, -1, 0, 0.f, 1));
  }
  (void)W;
  e->n_prep = (int)prep.size();
  e->n_unprep = (int)unprep.size();
  int rc;
  if ((rc = dalloc(e, &e->Wd, e->Wtot))) return rc;
  if (e->plain_lstm) {
    // both time gates pinned at 1: their pre-activation columns of PX are a constant bias of 1e4 (sigmoid = 1 exactly in
    // fp32, derivative exactly 0), their weights stay zero; no o-gate time terms
    std::vector<float> big(2 * H, 1.0e4f);
    CK(cudaMemcpy(e->Wd + e->wd_off.at("bx_all") + e->oTN, big.data(), big.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if ((rc = dalloc(e, &e->Pg, e->Ptot + e->Wtot))) return rc;
  e->dWd = e->Pg + e->Ptot;   // Ptot is a multiple of 4 floats: dWd stays 16-byte aligned
  if ((rc = dalloc(e, &e->ops_prep, e->n_prep, false))) return rc;
  if ((rc = dalloc(e, &e->ops_unprep, e->n_unprep, false))) return rc;
  CK(cudaMemcpy(e->ops_prep, prep.data(), prep.size() * sizeof(BlockOp), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->ops_unprep, unprep.data(), unprep.size() * sizeof(BlockOp), cudaMemcpyHostToDevice));
  return 0;
}

int alloc_bn(clsr_engine* e, BnLayer* b) {
  int rc;
  if ((rc = dalloc(e, &b->stat_f, 2 * b->N))) return rc;
  if ((rc = dalloc(e, &b->stat_b, 2 * b->N))) return rc;
  float** ps[] = {&b->scale, &b->shift, &b->mean, &b->rstd, &b->al, &b->be, &b->ga};
  for (float** p : ps)
    if ((rc = dalloc(e, p, b->N))) return rc;
  return 0;
}

// ---- launch helpers --------------------------------------------------------------------------------
inline int round16(int x) { return (x + 15) & ~15; }

// ---- TMA tensor maps for the operand loads of the tcgen05 kernels -----------------------------------
// cuTensorMapEncodeTiled is fetched through the runtime (no link against libcuda).  One map describes an
// fp32 matrix [rows, cols] with row pitch ld; the kernels load boxes of {8 columns, 128 rows} = one plane.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// L2 promotion of the TMA requests (a box row is only 32 bytes: the promotion size decides how much of the surrounding
// line(s) one request brings into L2).  CLSR_TMA_L2 = 0 | 64 | 128 | 256 overrides the default for A/B runs.
CUtensorMapL2promotion l2_promotion() {
  static const CUtensorMapL2promotion v = [] {
    const char* s = getenv("CLSR_TMA_L2");
    const int b = s ? atoi(s) : 128;
    return b >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                    : (b >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                : (b >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE));
  }();
  return v;
}
bool make_tmap(CUtensorMap* m, const float* base, int cols, long long rows, int ld, bool swizzle32 = false) {
  memset(m, 0, sizeof *m);
  EncodeTiledFn enc = encode_tiled();
  if (!enc || !base) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {8, (cuuint32_t)tc::kTileM};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
             l2_promotion(),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// Number of operand streams the TMA can load for this prologue (0: register loads): the rows must be plain
// matrix rows, 16-byte aligned, and K a whole number of planes.
int tma_streams(const AOp& a, int K) {
  static const bool off = getenv("CLSR_NO_TMA") != nullptr;
  if (off || !encode_tiled() || (K & 7) || K <= 0) return 0;
  auto ok = [](const float* p, int ld) { return p && !((uintptr_t)p & 15) && !(ld & 3); };
  if (a.mode == A_PLAIN || a.mode == A_BNRELU) return ok(a.A, a.lda) ? 1 : 0;
  if (a.mode == A_AFFINE2) return (ok(a.A, a.lda) && ok(a.A2, a.lda2)) ? 2 : 0;
  // concatenation [rows | rows[:, off:] * per-sequence row]: both parts are column ranges of the same plain rows
  // [A | A2], two plain matrices side by side (A_CAT2ROW with G == 1): one stream, two tensor maps
  if (a.mode == A_CAT2ROW)
    return (a.G == 1 && ok(a.A, a.lda) && ok(a.A2, a.lda2) && !(a.W1 & 7) && K > a.W1) ? 1 : 0;
  static const bool cat_off = getenv("CLSR_NO_TMA_CATMUL") != nullptr;
  if (a.mode == A_CATMUL && !cat_off)
    return (ok(a.A, a.lda) && ok(a.A2, a.lda2) && !(a.W1 & 7) && !(a.off & 3) && K > a.W1) ? 1 : 0;
  return 0;
}
// Columns of the first operand stream the tensor map has to cover.
int tmap_cols(const AOp& a, int K) {
  if (a.mode == A_CAT2ROW) return a.W1;
  if (a.mode != A_CATMUL) return K;
  const int second = a.off + (K - a.W1);
  return a.W1 > second ? a.W1 : second;
}

// One tcgen05 launch: K <= 160 (W resident in shared memory), N <= 256 (one UMMA, one TMEM accumulator).
int tc_gemm_one(clsr_engine* e, const char* name, int M, int N, int K, const AOp& a, const float* W, int ldw,
                const EpiOp& ep, bool stats, int kchunks = 1, bool probe = false) {
  // kchunks > 1: K is the width of one chunk, the contraction runs over kchunks * K columns of the (plain) operand in
  // ONE launch (accumulator resident in TMEM across the chunks); probe: only report whether that configuration fits
  // with two operand stages.
  const int kpad = round16(K), npad = round16(N);
  int nstages = 2;
  // per-element epilogue operand prefetched by bulk copy: needs contiguous, 16-byte aligned rows
  int eop = 0;
  auto al16p = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if ((N & 3) == 0) {
    if ((ep.flags & (E_RELUMASK | E_STAT_XHAT)) && ep.ldh == N && al16p(ep.hpre)) eop = 1;
    else if ((ep.flags & E_GROUPADD) && ep.ldga == N && al16p(ep.ga)) eop = 2;
    else if ((ep.flags & E_ACCUM) && ep.ldc == N && al16p(ep.C)) eop = 3;
  }
  const int st = stats ? 1 : 0;
  int tma = tma_streams(a, K);
  // TMA-store epilogue: the output tile goes through a shared-memory image (the kernel re-checks the operand
  // alignments its vector epilogue needs)
  static const bool no_tstore = getenv("CLSR_NO_TMA_STORE") != nullptr;
  int tstore = (!no_tstore && encode_tiled() && (N & 7) == 0 && (ep.ldc & 3) == 0 && al16p(ep.C) && M >= tc::kTileM) ? 1 : 0;
  // first configuration that fits shared memory: prefer two stages + operand prefetch + TMA
  // (two operand stages matter more than the prefetched epilogue operand, which matters more than the store image)
  // (three operand stages when everything else still fits: more loads in flight per SM; CLSR_TC_STAGES=2 caps it)
  static const int max_st = getenv("CLSR_TC_STAGES") ? atoi(getenv("CLSR_TC_STAGES")) : 3;
  const int cand[][4] = {{max_st >= 3 ? 3 : 2, eop, tma, tstore}, {2, eop, tma, tstore}, {2, eop, tma, 0}, {2, eop, tma == 2 ? 0 : tma, 0}, {2, 0, tma, 0}, {1, eop, tma, 0},
                         {2, 0, 0, 0},          {1, eop, 0, 0},   {1, 0, 0, 0}};
  tc::Smem L;
  bool fits = false;
  for (const auto& c : cand) {
    L = tc::smem_layout(kpad, npad, N, c[0], c[1], st, c[2], c[3], kchunks);
    if (L.total <= e->tc_smem_max) { nstages = c[0]; eop = c[1]; tma = c[2]; tstore = c[3]; fits = true; break; }
  }
  if (probe) return (fits && nstages >= 2) ? 0 : 1;
  if (!fits) return fail(e, CLSR_ERR_ARG, "tc_gemm %s: K=%d N=%d does not fit shared memory", name, K, N);
  CUtensorMap tmA, tmA2, tmC;
  memset(&tmA2, 0, sizeof tmA2);
  memset(&tmC, 0, sizeof tmC);
  if (tma >= 1 && !make_tmap(&tmA, a.A, kchunks > 1 ? kchunks * K : tmap_cols(a, K), M, a.lda)) tma = 0;
  if (tma == 2 && !make_tmap(&tmA2, a.A2, K, M, a.lda2)) tma = 0;
  if (tma == 1 && a.mode == A_CAT2ROW && !make_tmap(&tmA2, a.A2, K - a.W1, M, a.lda2)) tma = 0;
  if (tstore && !make_tmap(&tmC, ep.C, N, M, ep.ldc, true)) tstore = 0;
  if (!tma) memset(&tmA, 0, sizeof tmA);
  L = tc::smem_layout(kpad, npad, N, nstages, eop, st, tma, tstore, kchunks);
  uint32_t cols = 32;
  while ((int)cols < (stats ? 4 : 2) * npad) cols <<= 1;   // two accumulators (+ two statistic regions)
  int per_sm = e->smem_optin / (L.total + 6 * 1024);
  if (per_sm > 1) per_sm = 1;  // 544 threads: one resident CTA per SM
  if (per_sm * (int)cols > 512) per_sm = 512 / (int)cols;
  if (per_sm < 1) per_sm = 1;
  int tiles = cdiv(M, tc::kTileM);
  int grid = tiles < e->num_sms * per_sm ? tiles : e->num_sms * per_sm;
  if (stats) tc::tc_gemm_kernel<true><<<grid, tc::kThreads, L.total, e->stream>>>(M, N, K, kpad, npad, nstages, cols, eop, tma, tstore, kchunks, a, W, ldw, ep, tmA, tmA2, tmC);
  else tc::tc_gemm_kernel<false><<<grid, tc::kThreads, L.total, e->stream>>>(M, N, K, kpad, npad, nstages, cols, eop, tma, tstore, kchunks, a, W, ldw, ep, tmA, tmA2, tmC);
  POST(name);
  return 0;
}

// tensor-core recurrences: instantiated for the model's 40-wide states (H == U == item_dim + cate_dim)
bool rnn_tc_ok(const clsr_engine* e) {
  if (getenv("CLSR_RNN_SIMT")) return false;
  return e->cfg.math_mode == 1 && e->U == 40 && e->H == 40;
}

bool tc_eligible(clsr_engine* e, const char* name, int M, int N, int K, const AOp& a, bool stats) {
  if (e->cfg.math_mode != 1 || M < 1024) return false;
  if (const char* only = getenv("CLSR_TC_ONLY")) {  // developer bisection aid: comma-separated GEMM names
    std::string list = std::string(",") + only + ",";
    if (list.find(std::string(",") + name + ",") == std::string::npos) return false;
  }
  if (N > 128 && stats) return false;   // accumulators + per-row statistics must fit 512 TMEM columns
  if (K > 160 && a.mode != A_PLAIN) return false;
  return true;
}

int tc_gemm(clsr_engine* e, const char* name, int M, int N, int K, const AOp& a, const float* W, int ldw,
            const EpiOp& ep, bool stats) {
  int rc;
  static const bool no_kloop = getenv("CLSR_NO_KLOOP") != nullptr;
  if (!no_kloop && K > 160 && N <= 256 && a.mode == A_PLAIN && tma_streams(a, 8) == 1) {
    // deep contraction of a plain operand (dX: K = 480): one launch that walks K in chunks, if W and two stages fit
    for (int kc = cdiv(K, 160); kc <= 16; ++kc) {
      if (K % kc || (K / kc) % 8) continue;
      if (tc_gemm_one(e, name, M, N, K / kc, a, W, ldw, ep, stats, kc, true) == 0)
        return tc_gemm_one(e, name, M, N, K / kc, a, W, ldw, ep, stats, kc);
    }
  }
  for (int n0 = 0; n0 < N; n0 += 240) {          // column slabs (one UMMA is at most 256 wide)
    int nn = N - n0 < 240 ? N - n0 : 240;
    if (N <= 256) nn = N;
    // K slabs accumulate through the epilogue; wide outputs take 128-deep slabs so that W (K x N, split) and one
    // operand stage still fit shared memory
    const int ksl = (K > 160 && nn > 128) ? 128 : 160;
    for (int k0 = 0; k0 < K; k0 += ksl) {
      int kk = K - k0 < ksl ? K - k0 : ksl;
      AOp a2 = a;
      EpiOp p2 = ep;
      // later K slabs accumulate onto C: bias / broadcast terms were added by the first slab
      if (k0) { a2.A = a.A + k0; p2.flags = (p2.flags | E_ACCUM) & ~(E_ROWBIAS | E_GROUPADD); p2.bias = nullptr; }
      p2.C = ep.C + n0;
      if (ep.bias && !k0) p2.bias = ep.bias + n0;
      // column statistics are those of the finished sum: only the last K slab accumulates them (counting every
      // slab's partial result put the sums of BOTH partial and final values into the alpha gate's first
      // BatchNorm whenever its K = 2H+2D+1 = 161 > 160 input took this path, i.e. for >= 1024 rows)
      const bool last_k = k0 + ksl >= K;
      if ((rc = tc_gemm_one(e, name, M, nn, kk, a2, W + (size_t)k0 * ldw + n0, ldw, p2, stats && last_k))) return rc;
    }
    if (N <= 256) break;
  }
  return 0;
}

int gemm(clsr_engine* e, const char* name, int M, int N, int K, const AOp& a, const float* W, int ldw,
         const EpiOp& ep, bool stats) {
  if (M <= 0) return 0;
  if (tc_eligible(e, name, M, N, K, a, stats)) return tc_gemm(e, name, M, N, K, a, W, ldw, ep, stats);
  auto launch = [&](auto kern, int BM, int BN) {
    int tiles = cdiv(M, BM);
    int gx = tiles < e->num_sms * 4 ? tiles : e->num_sms * 4;
    dim3 grid(gx, cdiv(N, BN));
    kern<<<grid, 256, 0, e->stream>>>(M, N, K, a, W, ldw, ep);
  };
  if (N <= 40) {
    if (stats) launch(gemm_kernel<8, 5, 8, true>, 256, 40); else launch(gemm_kernel<8, 5, 8, false>, 256, 40);
  } else if (N <= 64) {
    if (stats) launch(gemm_kernel<8, 4, 16, true>, 128, 64); else launch(gemm_kernel<8, 4, 16, false>, 128, 64);
  } else if (N <= 80) {
    if (stats) launch(gemm_kernel<8, 5, 16, true>, 128, 80); else launch(gemm_kernel<8, 5, 16, false>, 128, 80);
  } else {
    if (stats) launch(gemm_kernel<8, 4, 32, true>, 64, 128); else launch(gemm_kernel<8, 4, 32, false>, 64, 128);
  }
  POST(name);
  return 0;
}

// dW on tensor cores; N is cut into column slabs of at most 240.
// One weight-gradient problem: `la` (Kl columns) on the MMA's lanes, `cb` (Nc columns, + a constant-one column when
// ones_b) on its N side.  transposed = 0: dW[Kl, Nc] (+ colsum row from a ones column on the lane side);
// transposed = 1: the operands were exchanged by the caller and the accumulator is dW^T (tc_gemm.cuh, dw_body).
int tc_dw_emit(clsr_engine* e, const char* name, int M, int Kl, int acols, int Nc, int transposed, const AOp& la, const AOp& cb,
               float* dW, int lddw, float* colsum) {
  const int ncols = Nc + ((transposed && colsum) ? 1 : 0);   // columns of the N-side stage (with the ones column)
  const int npad = round16(ncols);
  int nstages = 2;
  int ta = tma_streams(la, Kl), tb = tma_streams(cb, Nc);
  // first configuration that fits: two stages beat TMA with one stage (measured on dWs0t: 0.25 vs 0.28 ms)
  // (three stages when they fit with full TMA: more loads in flight per SM; CLSR_DW_STAGES=2 caps it)
  static const int max_st = getenv("CLSR_DW_STAGES") ? atoi(getenv("CLSR_DW_STAGES")) : 3;
  const int cand[][3] = {{max_st >= 3 ? 3 : 2, ta, tb}, {2, ta, tb}, {2, ta == 2 ? 0 : ta, tb}, {2, ta == 2 ? 0 : ta, tb == 2 ? 0 : tb},
                         {1, ta, tb}, {2, 0, 0}, {1, 0, 0}};
  tc::DwSmem L;
  bool fits = false;
  for (int ci = 0; ci < 7; ++ci) {
    const int* c = cand[ci];
    L = tc::dw_smem_layout(Kl, acols, Nc, npad, c[0], c[1], c[2]);
    if (L.total <= e->tc_dw_smem_max) { nstages = c[0]; ta = c[1]; tb = c[2]; fits = true; break; }
  }
  if (!fits) return fail(e, CLSR_ERR_ARG, "tc_dwgemm %s: %d x %d does not fit", name, Kl, Nc);
  CUtensorMap tmA, tmA2, tmB, tmB2;
  memset(&tmA, 0, sizeof tmA); memset(&tmA2, 0, sizeof tmA2); memset(&tmB, 0, sizeof tmB); memset(&tmB2, 0, sizeof tmB2);
  bool okm = true;
  if (ta >= 1) okm = okm && make_tmap(&tmA, la.A, tmap_cols(la, Kl), M, la.lda);
  if (ta == 2) okm = okm && make_tmap(&tmA2, la.A2, Kl, M, la.lda2);
  if (tb >= 1) okm = okm && make_tmap(&tmB, cb.A, tmap_cols(cb, Nc), M, cb.lda);
  if (tb == 2) okm = okm && make_tmap(&tmB2, cb.A2, Nc, M, cb.lda2);
  if (!okm) { ta = tb = 0; L = tc::dw_smem_layout(Kl, acols, Nc, npad, nstages, 0, 0); }
  uint32_t cols = 32;
  while ((int)cols < npad) cols <<= 1;
  const int tiles = cdiv(M, tc::kTileM);
  const int grid = tiles < e->num_sms ? tiles : e->num_sms;
  const int planes_a = (acols + 7) / 8, planes_b = (ncols + 7) / 8;
  const int octa = tc::dw_split(tc::kDwProducers / 8, planes_a, planes_b, tc::piece_cost(la.mode), tc::piece_cost(cb.mode));
  if (e->dw_grouping && e->dwg.n < tc::kDwGroupMax) {
    // deferred: becomes one problem of the next grouped launch (dw_group_flush)
    tc::DwProblem& P = e->dwg.p[e->dwg.n];
    P.M = M; P.K = Kl; P.acols = acols; P.N = Nc; P.npad = npad; P.nstages = nstages; P.tma_a = ta; P.tma_b = tb; P.lddw = lddw;
    P.octa = octa; P.transposed = transposed;
    P.tmem_cols = cols; P.a = la; P.b = cb; P.dW = dW; P.colsum = colsum;
    P.tmA = tmA; P.tmA2 = tmA2; P.tmB = tmB; P.tmB2 = tmB2;
    e->dwg_weight[e->dwg.n] = (long long)tiles * (planes_a + planes_b);
    e->dwg_tiles[e->dwg.n] = tiles;
    if (L.total > e->dwg_smem) e->dwg_smem = L.total;
    e->dwg.n++;
    return 0;
  }
  tc::tc_dw_kernel<<<grid, tc::kDwThreads, L.total, e->stream>>>(M, Kl, acols, Nc, npad, nstages, cols, ta, tb, octa, transposed, la, cb,
                                                               dW, lddw, colsum, tmA, tmA2, tmB, tmB2);
  POST(name);
  return 0;
}

int tc_dwgemm(clsr_engine* e, const char* name, int M, int K, int N, const AOp& a, const AOp& b, float* dW,
              int lddw, float* colsum) {
  int rc;
  // Wide gradients (N >= 2 K): put B on the lanes in chunks of <= 128 columns and A (+ the ones column) on the N side.
  // The MMA's N drops from up to 240 to round16(K + 1), no lane is wasted and every chunk double-buffers.
  static const bool no_tr = getenv("CLSR_DW_NO_TRANSPOSE") != nullptr;
  const int kones = K + (colsum ? 1 : 0);
  // (single chunk only: with several chunks A is re-read and re-converted per chunk -- dWx as 4 x 120 lanes measured 13 % slower
  //  than as 3 slabs of 160 columns; dWs0t, 80 lanes: 0.215 -> 0.167 ms)
  if (!no_tr && N >= 2 * K && kones <= 64 && N <= 128) {
    const int nch = cdiv(N, 128);
    const int per = ((cdiv(N, nch) + 7) / 8) * 8;
    for (int n0 = 0; n0 < N; n0 += per) {
      const int nn = N - n0 < per ? N - n0 : per;
      AOp b2 = b;
      if (n0) b2.A = b.A + n0;
      if ((rc = tc_dw_emit(e, name, M, nn, nn, K, 1, b2, a, dW + n0, lddw, colsum ? colsum + n0 : nullptr))) return rc;
    }
    return 0;
  }
  // columns of A that are stored (the MMA's remaining lanes alias what follows; CLSR_DW_FULL_A=1 stores all 128: A/B switch)
  static const bool full_a = getenv("CLSR_DW_FULL_A") != nullptr;
  const int acols = full_a ? 128 : kones;
  // Column slabs of B: as few as possible (<= 240 columns: one TMEM accumulator), but when B is a plain matrix take
  // more of them if that is what lets TWO stages with TMA loads fit shared memory -- a single-staged slab loop
  // serialises load, conversion and MMA.
  int nslab = cdiv(N, 240);
  if (b.mode == A_PLAIN) {
    static const int slab_st = getenv("CLSR_DW_SLAB_STAGES") ? atoi(getenv("CLSR_DW_SLAB_STAGES")) : 2;
    for (int ns = nslab; ns <= nslab + 3; ++ns) {
      const int p8 = ((cdiv(N, ns) + 7) / 8) * 8;
      if (tc::dw_smem_layout(K, acols, p8, round16(p8), slab_st, tma_streams(a, K), tma_streams(b, p8)).total <= e->tc_dw_smem_max) { nslab = ns; break; }
    }
  }
  const int per = ((cdiv(N, nslab) + 7) / 8) * 8;
  for (int n0 = 0; n0 < N; n0 += per) {
    const int nn = N - n0 < per ? N - n0 : per;
    AOp b2 = b;
    if (n0) {
      if (b.mode != A_PLAIN) return fail(e, CLSR_ERR_ARG, "tc_dwgemm %s: column slabs need a plain B operand", name);
      b2.A = b.A + n0;
    }
    if ((rc = tc_dw_emit(e, name, M, K, acols, nn, 0, a, b2, dW + n0, lddw, colsum ? colsum + n0 : nullptr))) return rc;
  }
  return 0;
}

int dwgemm(clsr_engine* e, const char* name, int M, int K, int N, const AOp& a, const AOp& b, float* dW,
           int lddw, float* colsum) {
  if (M <= 0) return 0;
  if (tc_eligible(e, name, M, N, K, a, false) && K + (colsum ? 1 : 0) <= 128 && (N <= 240 || b.mode == A_PLAIN))
    return tc_dwgemm(e, name, M, K, N, a, b, dW, lddw, colsum);
  if (tc_eligible(e, name, M, N, 120, a, false) && a.mode == A_PLAIN && (N <= 240 || b.mode == A_PLAIN)) {
    // wide models: dW rows are independent -> chunks of <= 120 rows of dW (columns of A), one accumulator each
    for (int k0 = 0; k0 < K; k0 += 120) {
      const int kk = K - k0 < 120 ? K - k0 : 120;
      AOp a2 = a;
      a2.A = a.A + k0;
      int rc = tc_dwgemm(e, name, M, kk, N, a2, b, dW + (size_t)k0 * lddw, lddw, k0 == 0 ? colsum : nullptr);
      if (rc) return rc;
    }
    return 0;
  }
  int ty = cdiv(K, 64), tz = cdiv(N, 64);
  int want = (e->num_sms * 4) / (ty * tz);
  if (want < 1) want = 1;
  int rows_per = cdiv(M, want);
  rows_per = ((rows_per + 31) / 32) * 32;
  dim3 grid(cdiv(M, rows_per), ty, tz);
  if (colsum) dw_kernel<true><<<grid, 256, 0, e->stream>>>(M, K, N, a, b, dW, lddw, colsum, rows_per);
  else dw_kernel<false><<<grid, 256, 0, e->stream>>>(M, K, N, a, b, dW, lddw, nullptr, rows_per);
  POST(name);
  return 0;
}

// Launch the deferred weight-gradient problems as one grid partitioned in proportion to their work.
int dw_group_flush(clsr_engine* e, const char* name) {
  e->dw_grouping = false;
  tc::DwGroup& g = e->dwg;
  if (g.n == 0) return 0;
  long long wsum = 0;
  for (int i = 0; i < g.n; ++i) wsum += e->dwg_weight[i];
  int left = e->num_sms;
  for (int i = 0; i < g.n; ++i) {
    int c = (int)((long long)e->num_sms * e->dwg_weight[i] / wsum);
    if (c < 1) c = 1;
    if (c > e->dwg_tiles[i]) c = e->dwg_tiles[i];
    g.p[i].ncta = c;
    left -= c;
  }
  for (int guard = 0; left != 0 && guard < 4 * e->num_sms; ++guard) {   // hand the rounding remainder to the busiest problems
    int best = -1;
    double bl = -1.0;
    for (int i = 0; i < g.n; ++i) {
      const double load = (double)e->dwg_weight[i] / g.p[i].ncta;
      if (left > 0 ? (g.p[i].ncta < e->dwg_tiles[i] && load > bl) : (g.p[i].ncta > 1 && (bl < 0 || load < bl))) { bl = load; best = i; }
    }
    if (best < 0) break;
    g.p[best].ncta += left > 0 ? 1 : -1;
    left += left > 0 ? -1 : 1;
  }
  int c0 = 0;
  for (int i = 0; i < g.n; ++i) { g.p[i].cta0 = c0; c0 += g.p[i].ncta; }
  tc::tc_dw_group_kernel<<<c0, tc::kDwThreads, e->dwg_smem, e->stream>>>(g);
  POST(name);
  g.n = 0;
  e->dwg_smem = 0;
  return 0;
}
void dw_group_begin(clsr_engine* e) {
  static const bool off = getenv("CLSR_DW_NO_GROUP") != nullptr;   // developer switch: one launch per weight gradient
  e->dw_grouping = !off && e->cfg.math_mode == 1;
  e->dwg.n = 0;
  e->dwg_smem = 0;
}

AOp a_plain(const float* A, int lda) {
  AOp a; a.mode = A_PLAIN; a.A = A; a.lda = lda; return a;
}
AOp a_bnrelu(const float* A, int lda, const BnLayer& b) {
  AOp a; a.mode = A_BNRELU; a.A = A; a.lda = lda; a.v0 = b.scale; a.v1 = b.shift; return a;
}
AOp a_affine2(const float* dy, const float* h, int ld, const BnLayer& b) {
  AOp a; a.mode = A_AFFINE2; a.A = dy; a.A2 = h; a.lda = ld; a.lda2 = ld; a.v0 = b.al; a.v1 = b.be; a.v2 = b.ga;
  return a;
}
AOp a_catmul(const float* A, int lda, int W1, int off, const float* V, int ldv, int T) {
  AOp a; a.mode = A_CATMUL; a.A = A; a.lda = lda; a.W1 = W1; a.off = off; a.A2 = V; a.lda2 = ldv; a.T = T;
  return a;
}
AOp a_mulrow(const float* A, int lda, int off, const float* V, int ldv, int T, int G) {
  AOp a; a.mode = A_MULROW; a.A = A; a.lda = lda; a.off = off; a.A2 = V; a.lda2 = ldv; a.T = T; a.G = G;
  return a;
}
AOp a_cat2row(const float* P, int ldp, int W1, const float* V, int ldv, int G) {
  AOp a; a.mode = A_CAT2ROW; a.A = P; a.lda = ldp; a.W1 = W1; a.A2 = V; a.lda2 = ldv; a.G = G; return a;
}
EpiOp e_store(float* C, int ldc, const float* bias = nullptr, int flags = 0) {
  EpiOp p; p.C = C; p.ldc = ldc; p.bias = bias; p.flags = flags; return p;
}

int bn_fwd(clsr_engine* e, BnLayer& b, double count, int train, int update) {
  if (train && e->world > 1) {  // single-device semantics: statistics over the global batch
    count *= e->world;
    if (e->peer_ready && 2 * b.N <= kPeerSlots) {   // all-reduce over NVLink peer memory inside the finalize kernel
      bn_fwd_finalize_peer_kernel<<<1, 256, 0, e->stream>>>(
          e->pc, ++e->peer_seq, b.stat_f, b.N, count, e->P + b.gamma, e->P + b.beta, e->cfg.bn_eps, e->cfg.bn_momentum,
          e->P + b.mmean, e->P + b.mvar, update, b.scale, b.shift, b.mean, b.rstd);
      POST("bn_fwd_finalize_peer");
      return 0;
    }
    int rc = allreduce(e, b.stat_f, 2 * b.N, kNcclFloat64);
    if (rc) return rc;
  }
  bn_fwd_finalize_kernel<<<cdiv(b.N, 128), 128, 0, e->stream>>>(
      b.stat_f, b.N, count, e->P + b.gamma, e->P + b.beta, e->cfg.bn_eps, e->cfg.bn_momentum,
      e->P + b.mmean, e->P + b.mvar, train, update, b.scale, b.shift, b.mean, b.rstd);
  POST("bn_fwd_finalize");
  return 0;
}
int bn_bwd(clsr_engine* e, BnLayer& b, double count) {
  // gamma/beta gradients come from the already global sums: only rank 0 contributes them to the
  // dense-gradient all-reduce
  const float add_scale = e->rank == 0 ? 1.f : 0.f;
  if (e->world > 1) {
    count *= e->world;
    if (e->peer_ready && 2 * b.N <= kPeerSlots) {
      bn_bwd_finalize_peer_kernel<<<1, 256, 0, e->stream>>>(e->pc, ++e->peer_seq, b.stat_b, b.N, count, e->P + b.gamma, b.mean,
                                                            b.rstd, b.al, b.be, b.ga, e->Pg + b.gamma, e->Pg + b.beta, add_scale);
      POST("bn_bwd_finalize_peer");
      return 0;
    }
    int rc = allreduce(e, b.stat_b, 2 * b.N, kNcclFloat64);
    if (rc) return rc;
  }
  bn_bwd_finalize_kernel<<<cdiv(b.N, 128), 128, 0, e->stream>>>(
      b.stat_b, b.N, count, e->P + b.gamma, b.mean, b.rstd, b.al, b.be, b.ga, e->Pg + b.gamma, e->Pg + b.beta, add_scale);
  POST("bn_bwd_finalize");
  return 0;
}

// Forward of a row-wise _fcn_net (alpha gate, logit): in [rows, m.in] -> h0, h1, out[rows].
// Fill the argument block of the fused _fcn_net kernels.
void coop_args(clsr_engine* e, Mlp& m, const float* in, int rows, float* h0, float* h1, CoopMlp* a) {
  memset(a, 0, sizeof *a);
  a->rows = rows; a->K = m.in; a->n0 = m.n0; a->n1 = m.n1; a->ld_in = m.in;
  a->in = in; a->w0 = e->P + m.w0; a->b0 = e->P + m.b0; a->w1 = e->P + m.w1; a->b1 = e->P + m.b1;
  a->wo = e->P + m.wo; a->bo = e->P + m.bo;
  a->h0 = h0; a->h1 = h1;
  a->dw0 = e->Pg + m.w0; a->dw1 = e->Pg + m.w1;
  a->dwo = e->Pg + m.wo; a->dbo = e->Pg + m.bo;
  BnLayer* bl[2] = {&m.bn0, &m.bn1};
  CoopBn* cb[2] = {&a->bn0, &a->bn1};
  for (int i = 0; i < 2; ++i) {
    BnLayer& b = *bl[i];
    *cb[i] = CoopBn{e->P + b.gamma, e->P + b.beta, e->P + b.mmean, e->P + b.mvar, b.scale, b.shift, b.mean, b.rstd,
                    b.al, b.be, b.ga, e->Pg + b.gamma, e->Pg + b.beta, b.stat_f, b.stat_b};
  }
  a->count = (double)rows * e->world;
  a->eps = e->cfg.bn_eps; a->momentum = e->cfg.bn_momentum;
  a->add_scale = e->rank == 0 ? 1.f : 0.f;
  a->bar = e->coop_bar;
  a->world = (e->world > 1 && e->peer_ready) ? e->world : 1;
  a->pc = e->pc;
}
bool coop_usable(const clsr_engine* e, const Mlp& m, int rows) {
  const bool on = (&m == &e->mlp_alpha) ? e->coop_alpha : ((&m == &e->mlp_logit) ? e->coop_logit : false);
  // data parallel without the peer-memory exchange (plain NCCL mode before clsr_peer_setup_*): layer-by-layer path
  return on && rows <= (long long)e->coop_rpc * e->num_sms && (e->world == 1 || e->peer_ready);
}
constexpr int kCoopUnavailable = -1000;
template <typename Kern>
int coop_launch(clsr_engine* e, Kern kern, const CoopMlp& a, size_t smem, const char* name) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(e->num_sms); cfg.blockDim = dim3(kCoopThreads); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;    // all CTAs co-resident: the kernels synchronise through a grid barrier
  cfg.attrs = at; cfg.numAttrs = 1;
  const cudaError_t lc = cudaLaunchKernelEx(&cfg, kern, a, e->coop_rpc);
  if (lc == cudaErrorCooperativeLaunchTooLarge || lc == cudaErrorNotSupported) {
    // the device cannot co-schedule one CTA per SM right now (MPS / MIG partitions, SM limits): use the
    // layer-by-layer kernels from here on -- they read and write the same buffers
    cudaGetLastError();
    e->coop_alpha = e->coop_logit = false;
    return kCoopUnavailable;
  }
  CK(lc);
  POST(name);
  return 0;
}

int mlp_fwd(clsr_engine* e, Mlp& m, const float* in, int rows, float* h0, float* h1, float* out, int train,
            int update) {
  int rc;
  if (train && coop_usable(e, m, rows)) {
    CoopMlp a;
    coop_args(e, m, in, rows, h0, h1, &a);
    a.out = out; a.update_moving = update;
    const CoopFwdSmem L = coop_fwd_smem(m.in, m.n0, m.n1, e->coop_rpc);
    rc = coop_launch(e, mlp_fwd_coop_kernel, a, (size_t)L.total * 4, "mlp_fwd_fused");
    if (rc != kCoopUnavailable) return rc;
  }
  EpiOp ep = e_store(h0, m.n0, e->P + m.b0);
  ep.stat = m.bn0.stat_f;
  if ((rc = gemm(e, "mlp_l0", rows, m.n0, m.in, a_plain(in, m.in), e->P + m.w0, m.n0, ep, train != 0))) return rc;
  if ((rc = bn_fwd(e, m.bn0, rows, train, update))) return rc;
  ep = e_store(h1, m.n1, e->P + m.b1);
  ep.stat = m.bn1.stat_f;
  if ((rc = gemm(e, "mlp_l1", rows, m.n1, m.n0, a_bnrelu(h0, m.n0, m.bn0), e->P + m.w1, m.n1, ep, train != 0)))
    return rc;
  if ((rc = bn_fwd(e, m.bn1, rows, train, update))) return rc;
  rowdot_kernel<<<grid1d(e, (long long)rows * 32, 256), 256, 0, e->stream>>>(
      h1, m.n1, m.bn1.scale, m.bn1.shift, e->P + m.wo, e->P + m.bo, rows, out);
  POST("rowdot");
  return 0;
}

// Backward of a row-wise _fcn_net from dout[rows]: weight/BN gradients into Pg, d(in) into din.
int mlp_bwd(clsr_engine* e, Mlp& m, const float* in, int rows, const float* h0, const float* h1,
            const float* dout, float* dy1, float* dy0, const float* w0T, const float* w1T, float* din) {
  int rc;
  if (coop_usable(e, m, rows)) {
    CoopMlp a;
    coop_args(e, m, in, rows, const_cast<float*>(h0), const_cast<float*>(h1), &a);
    a.dout = dout; a.din = din; a.w0T = w0T; a.w1T = w1T;
    const CoopBwdSmem L = coop_bwd_smem(m.in, m.n0, m.n1, e->coop_rpc);
    rc = coop_launch(e, mlp_bwd_coop_kernel, a, (size_t)L.total * 4, "mlp_bwd_fused");
    if (rc != kCoopUnavailable) return rc;
  }
  {
    int nx = ((m.n1 + 31) / 32) * 32;
    dim3 blk(nx, 256 / nx > 0 ? 256 / nx : 1);
    rowdot_bwd_kernel<<<grid1d(e, rows, blk.y, 2), blk, 0, e->stream>>>(
        dout, h1, m.n1, m.bn1.scale, m.bn1.shift, m.bn1.mean, m.bn1.rstd, e->P + m.wo, rows, dy1,
        m.bn1.stat_b, e->Pg + m.wo, e->Pg + m.bo);
    POST("rowdot_bwd");
  }
  if ((rc = bn_bwd(e, m.bn1, rows))) return rc;
  EpiOp ep = e_store(dy0, m.n0, nullptr, E_RELUMASK | E_STAT_XHAT);
  ep.hpre = h0; ep.ldh = m.n0; ep.scale = m.bn0.scale; ep.shift = m.bn0.shift; ep.mean = m.bn0.mean;
  ep.rstd = m.bn0.rstd; ep.stat = m.bn0.stat_b;
  if ((rc = gemm(e, "mlp_dy0", rows, m.n0, m.n1, a_affine2(dy1, h1, m.n1, m.bn1), w1T, m.n0, ep, true))) return rc;
  if ((rc = dwgemm(e, "mlp_dw1", rows, m.n0, m.n1, a_bnrelu(h0, m.n0, m.bn0), a_affine2(dy1, h1, m.n1, m.bn1),
                   e->Pg + m.w1, m.n1, nullptr)))
    return rc;
  if ((rc = bn_bwd(e, m.bn0, rows))) return rc;
  if ((rc = gemm(e, "mlp_din", rows, m.in, m.n0, a_affine2(dy0, h0, m.n0, m.bn0), w0T, m.in,
                 e_store(din, m.in), false)))
    return rc;
  if ((rc = dwgemm(e, "mlp_dw0", rows, m.in, m.n0, a_plain(in, m.in), a_affine2(dy0, h0, m.n0, m.bn0),
                   e->Pg + m.w0, m.n0, nullptr)))
    return rc;
  return 0;
}

struct StepCtx {
  int B, G, S, T;
  int seq_stride;   // elements between consecutive sequences in the [*,T] input arrays
  int user_stride;  // elements between consecutive sequences in users[]
  const int32_t *users, *items, *cates, *ih, *ch, *mask;
  const float *tfa, *ttn, *labels;
};

// Copy (host) or alias (device) the feed arrays.
int stage_inputs(clsr_engine* e, const clsr_batch* b, StepCtx* c, bool need_labels, bool sync_call) {
  const int T = e->T;
  c->B = b->rows; c->G = b->group; c->S = b->rows / b->group; c->T = T;
  // device layout of the staged block: [ih | ch | mask | tfa | ttn | users | items | cates | labels]
  const int S = c->S, B = c->B, G = c->G;
  const size_t seq_i = (size_t)S * T * 4;
  {
    char* d = (char*)e->in_ih;
    c->ih = (const int32_t*)d; d += seq_i;
    c->ch = (const int32_t*)d; d += seq_i;
    c->mask = (const int32_t*)d; d += seq_i;
    c->tfa = (const float*)d; d += seq_i;
    c->ttn = (const float*)d; d += seq_i;
    c->users = (const int32_t*)d; d += (size_t)S * 4;
    c->items = (const int32_t*)d; d += (size_t)B * 4;
    c->cates = (const int32_t*)d; d += (size_t)B * 4;
    c->labels = (const float*)d;
    c->seq_stride = T;
    c->user_stride = 1;
  }
  if (b->on_device) {
    // Device-resident feed: one kernel compacts it into the same staging block (one row per sequence) and
    // validates every id against its table; the step itself only ever reads engine-owned, validated arrays.
    StageFeed a;
    a.users = b->users; a.items = b->items; a.cates = b->cates; a.ih = b->item_history; a.ch = b->cate_history;
    a.mask = b->mask; a.tfa = b->time_from_first_action; a.ttn = b->time_to_now; a.labels = need_labels ? b->labels : nullptr;
    a.o_ih = const_cast<int32_t*>(c->ih); a.o_ch = const_cast<int32_t*>(c->ch); a.o_mask = const_cast<int32_t*>(c->mask);
    a.o_tfa = const_cast<float*>(c->tfa); a.o_ttn = const_cast<float*>(c->ttn); a.o_users = const_cast<int32_t*>(c->users);
    a.o_items = const_cast<int32_t*>(c->items); a.o_cates = const_cast<int32_t*>(c->cates);
    a.o_labels = const_cast<float*>(c->labels);
    a.S = S; a.G = G; a.T = T; a.B = B; a.seq_g = G;
    a.n_items = e->cfg.n_items; a.n_cates = e->cfg.n_cates; a.n_users = e->cfg.n_users;
    a.err = e->d_err;
    CK(cudaMemsetAsync(e->d_err, 0, sizeof(int32_t), e->stream));
    stage_device_feed_kernel<<<grid1d(e, (long long)S * T, 256, 4), 256, 0, e->stream>>>(a);
    POST("stage_device_feed");
    CK(cudaMemcpyAsync(e->h_err, e->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    return 0;
  }
  // Host feed in pinned (page-locked / registered) caller memory, synchronous call: the copy engine picks every
  // G-th row itself (2-D copies, source pitch G*T*4) -- no host-side packing pass over the feed; the ids are
  // validated on the device by the same kernel that stages device-resident feeds.  Only when the caller waits for
  // the step (its buffers are then guaranteed to outlive the copies).
  static const bool no_direct = getenv("CLSR_NO_DIRECT_H2D") != nullptr;
  if (sync_call && !no_direct && e->in_raw) {
    auto pinned = [](const void* p) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
      return at.type == cudaMemoryTypeHost;
    };
    bool all = pinned(b->item_history) && pinned(b->cate_history) && pinned(b->mask) && pinned(b->time_from_first_action) &&
               pinned(b->time_to_now) && pinned(b->users) && pinned(b->items) && pinned(b->cates) &&
               (!need_labels || !b->labels || pinned(b->labels));
    if (all) {
      char* r = e->in_raw;
      const void* seqs[5] = {b->item_history, b->cate_history, b->mask, b->time_from_first_action, b->time_to_now};
      for (int k = 0; k < 5; ++k)
        CK(cudaMemcpy2DAsync(r + k * seq_i, (size_t)T * 4, seqs[k], (size_t)G * T * 4, (size_t)T * 4, S, cudaMemcpyHostToDevice,
                             e->stream));
      char* rb = r + 5 * seq_i;
      const size_t bb = (size_t)B * 4;
      CK(cudaMemcpyAsync(rb, b->users, bb, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(rb + bb, b->items, bb, cudaMemcpyHostToDevice, e->stream));
      CK(cudaMemcpyAsync(rb + 2 * bb, b->cates, bb, cudaMemcpyHostToDevice, e->stream));
      const bool lab = need_labels && b->labels;
      if (lab) CK(cudaMemcpyAsync(rb + 3 * bb, b->labels, bb, cudaMemcpyHostToDevice, e->stream));
      StageFeed a;
      a.ih = (const int32_t*)r; a.ch = (const int32_t*)(r + seq_i); a.mask = (const int32_t*)(r + 2 * seq_i);
      a.tfa = (const float*)(r + 3 * seq_i); a.ttn = (const float*)(r + 4 * seq_i);
      a.users = (const int32_t*)rb; a.items = (const int32_t*)(rb + bb); a.cates = (const int32_t*)(rb + 2 * bb);
      a.labels = lab ? (const float*)(rb + 3 * bb) : nullptr;
      a.o_ih = const_cast<int32_t*>(c->ih); a.o_ch = const_cast<int32_t*>(c->ch); a.o_mask = const_cast<int32_t*>(c->mask);
      a.o_tfa = const_cast<float*>(c->tfa); a.o_ttn = const_cast<float*>(c->ttn); a.o_users = const_cast<int32_t*>(c->users);
      a.o_items = const_cast<int32_t*>(c->items); a.o_cates = const_cast<int32_t*>(c->cates);
      a.o_labels = const_cast<float*>(c->labels);
      a.S = S; a.G = G; a.T = T; a.B = B; a.seq_g = 1;
      a.n_items = e->cfg.n_items; a.n_cates = e->cfg.n_cates; a.n_users = e->cfg.n_users;
      a.err = e->d_err;
      CK(cudaMemsetAsync(e->d_err, 0, sizeof(int32_t), e->stream));
      stage_device_feed_kernel<<<grid1d(e, (long long)S * T, 256, 4), 256, 0, e->stream>>>(a);
      POST("stage_host_feed");
      CK(cudaMemcpyAsync(e->h_err, e->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      return 0;
    }
  }
  // Host feed in pageable memory (or an asynchronous call): pack only what the step reads (one row per sequence) into
  // pinned staging, validate the ids against the table sizes, then a single async H2D copy.
  size_t need = 5 * seq_i + (size_t)S * 4 + (size_t)B * 4 * 3;
  if (need > e->h_stage_bytes) return fail(e, CLSR_ERR_ARG, "batch exceeds staging capacity");
  CK(cudaEventSynchronize(e->h2d_done));
  char* hp = e->h_stage;
  auto pack_rows = [&](const void* src, size_t row_bytes) {
    const char* sp = (const char*)src;
    if (G == 1) memcpy(hp, sp, row_bytes * S);
    else for (int i = 0; i < S; ++i) memcpy(hp + row_bytes * i, sp + row_bytes * (size_t)i * G, row_bytes);
    char* r = hp;
    hp += row_bytes * S;
    return r;
  };
  // any id outside [0, limit)?  (branch-free min/max scan: vectorises)
  auto out_of_range = [](const int32_t* p, size_t n, long long limit) {
    uint32_t mx = 0;
    for (size_t i = 0; i < n; ++i) { const uint32_t v = (uint32_t)p[i]; mx = v > mx ? v : mx; }
    return n > 0 && (long long)mx >= limit;   // negative ids wrap to >= 2^31
  };
  size_t rowb = (size_t)T * 4;
  char* base = hp;
  const int32_t* p_ih = (const int32_t*)pack_rows(b->item_history, rowb);
  const int32_t* p_ch = (const int32_t*)pack_rows(b->cate_history, rowb);
  pack_rows(b->mask, rowb);
  pack_rows(b->time_from_first_action, rowb);
  pack_rows(b->time_to_now, rowb);
  const int32_t* p_us = (const int32_t*)pack_rows(b->users, 4);
  const int32_t* p_it = (const int32_t*)hp;
  memcpy(hp, b->items, (size_t)B * 4); hp += (size_t)B * 4;
  const int32_t* p_ct = (const int32_t*)hp;
  memcpy(hp, b->cates, (size_t)B * 4); hp += (size_t)B * 4;
  if (need_labels && b->labels) memcpy(hp, b->labels, (size_t)B * 4);
  hp += (size_t)B * 4;
  if (out_of_range(p_ih, (size_t)S * T, e->cfg.n_items)) return fail(e, CLSR_ERR_ARG, "item_history holds an id outside [0, %lld)", (long long)e->cfg.n_items);
  if (out_of_range(p_ch, (size_t)S * T, e->cfg.n_cates)) return fail(e, CLSR_ERR_ARG, "item_cate_history holds an id outside [0, %lld)", (long long)e->cfg.n_cates);
  if (out_of_range(p_us, (size_t)S, e->cfg.n_users)) return fail(e, CLSR_ERR_ARG, "users holds an id outside [0, %lld)", (long long)e->cfg.n_users);
  if (out_of_range(p_it, (size_t)B, e->cfg.n_items)) return fail(e, CLSR_ERR_ARG, "items holds an id outside [0, %lld)", (long long)e->cfg.n_items);
  if (out_of_range(p_ct, (size_t)B, e->cfg.n_cates)) return fail(e, CLSR_ERR_ARG, "cates holds an id outside [0, %lld)", (long long)e->cfg.n_cates);
  CK(cudaMemcpyAsync(e->in_ih, base, (size_t)(hp - base), cudaMemcpyHostToDevice, e->stream));
  CK(cudaEventRecord(e->h2d_done, e->stream));
  return 0;
}

// Raise the id-range error of a device-resident feed (flags written by stage_device_feed_kernel) once the
// stream has been synchronised.
int check_feed_flags(clsr_engine* e) {
  const int f = e->h_err ? *e->h_err : 0;
  if (!f) return 0;
  *e->h_err = 0;
  static const char* names[5] = {"users", "items", "cates", "item_history", "item_cate_history"};
  std::string which;
  for (int i = 0; i < 5; ++i)
    if (f & (1 << i)) which += std::string(which.empty() ? "" : ", ") + names[i];
  return fail(e, CLSR_ERR_ARG, "device-resident feed holds ids outside their tables (%s); they were read as id 0",
              which.c_str());
}

int check_batch(clsr_engine* e, const clsr_batch* b, bool train) {
  if (!b || b->rows <= 0 || b->group <= 0 || b->rows % b->group) return fail(e, CLSR_ERR_ARG, "bad rows/group");
  if (b->rows > e->Bmax) return fail(e, CLSR_ERR_ARG, "rows %d exceed max_rows %d", b->rows, e->Bmax);
  if (b->rows / b->group > e->Smax)
    return fail(e, CLSR_ERR_ARG, "sequences %d exceed max_seqs %d", b->rows / b->group, e->Smax);
  if (!b->users || !b->items || !b->cates || !b->item_history || !b->cate_history || !b->mask ||
      !b->time_from_first_action || !b->time_to_now)
    return fail(e, CLSR_ERR_ARG, "null feed array");
  if (train) {
    if (!b->labels) return fail(e, CLSR_ERR_ARG, "labels required for training");
    if (b->rows % e->cfg.train_group) return fail(e, CLSR_ERR_ARG, "rows not a multiple of train_group");
  }
  for (int t = 0; t < CLSR_NUM_TABLES; ++t)
    if (!e->tab[t]) return fail(e, CLSR_ERR_STATE, "table %d not bound", t);
  return 0;
}

int forward(clsr_engine* e, const StepCtx& c, int train, int update_bn) {
  const int B = c.B, G = c.G, S = c.S, T = c.T;
  const int D = e->D, U = e->U, H = e->H, Q = e->Q, A0 = e->A0, A1 = e->A1, NX = e->NX, Di = e->Di, Dc = e->Dc;
  const long long M = (long long)S * T, MB = (long long)B * T;
  cudaStream_t st = e->stream;
  int rc;
  auto W = [&](const char* n) { return e->Wd + e->wd_off.at(n); };

  seq_prep_kernel<<<grid1d(e, S, 128), 128, 0, st>>>(c.mask, c.seq_stride, T, S, e->cfg.contrastive_len_threshold,
                                                     e->d_len, e->counts);
  POST("seq_prep");
  if (train && (rc = allreduce(e, e->counts, 1, kNcclInt32))) return rc;
  blockop_kernel<<<e->n_prep, 256, 0, st>>>(e->ops_prep, e->Wd, e->P);
  POST("prep_weights");

  // ---- K1+K3: embedding gathers ----
  float* X = e->B("X");
  if ((rc = launch_gather_hist(e, c.ih, c.ch, c.seq_stride, X, M))) return rc;
  float* tgt = e->B("tgt");
  float *ul = e->B("ul"), *us = e->B("us");
  {
    // target item / category rows and the two user rows: one launch
    GatherRowsMulti gm;
    gm.n = 4;
    gm.s[0] = GatherRowsSeg{c.items, 1, e->tview[CLSR_TABLE_ITEM], Di, tgt, D, 0, B};
    gm.s[1] = GatherRowsSeg{c.cates, 1, e->tview[CLSR_TABLE_CATE], Dc, tgt, D, Di, B};
    gm.s[2] = GatherRowsSeg{c.users, c.user_stride, e->tview[CLSR_TABLE_USER_LONG], U, ul, U, 0, S};
    gm.s[3] = GatherRowsSeg{c.users, c.user_stride, e->tview[CLSR_TABLE_USER_SHORT], U, us, U, 0, S};
    gather_rows_multi_kernel<<<grid1d(e, (long long)B * Di / 4, 256), 256, 0, st>>>(gm);
    POST("gather_rows(tgt|users)");
  }

  // ---- hoisted input projections of the three recurrences ----
  float *TNL = e->B("TNL"), *PX = e->B("PX");
  if (e->plain_lstm) {
    // no time features: one product x . Wx (+ biases, incl. the constant gate columns)
    if ((rc = gemm(e, "px", (int)M, NX, D, a_plain(X, D), W("Wx_all"), NX, e_store(PX, NX, W("bx_all")), false))) return rc;
  } else {
  if ((H & 3) == 0)
    time_feat_v4_kernel<<<grid1d(e, M * 2 * H / 4, 256, 16), 256, 0, st>>>(c.ttn, c.tfa, c.seq_stride, T, e->P + e->p_tw1,
                                                                          e->P + e->p_tb1, e->P + e->p_tw2, e->P + e->p_tb2, H, TNL, M);
  else
    time_feat_kernel<<<grid1d(e, M * 2 * H, 256), 256, 0, st>>>(c.ttn, c.tfa, c.seq_stride, T, e->P + e->p_tw1, e->P + e->p_tb1,
                                                               e->P + e->p_tw2, e->P + e->p_tb2, H, TNL, M);
  POST("time_feat");
  static const bool px_split = getenv("CLSR_PX_SPLIT") != nullptr;   // A/B: separate time product accumulated onto PX
  if (!px_split && e->oO + 3 * H == NX && (D & 7) == 0 && (H & 3) == 0 && D + 2 * H <= 160) {
    // columns [0, oO): x . Wx;  columns [oO, NX) (o gate, two time gates): [x | time features] . [Wx ; Wt] in one product
    if ((rc = gemm(e, "px", (int)M, e->oO, D, a_plain(X, D), W("Wx_all"), NX, e_store(PX, NX, W("bx_all")), false))) return rc;
    if ((rc = gemm(e, "px_time", (int)M, 3 * H, D + 2 * H, a_cat2row(X, D, D, TNL, 2 * H, 1), W("Wx_all") + e->oO, NX,
                   e_store(PX + e->oO, NX, W("bx_all") + e->oO), false)))
      return rc;
  } else {
    if ((rc = gemm(e, "px", (int)M, NX, D, a_plain(X, D), W("Wx_all"), NX, e_store(PX, NX, W("bx_all")), false))) return rc;
    if ((rc = gemm(e, "px_time", (int)M, 3 * H, 2 * H, a_plain(TNL, 2 * H), W("Wt"), 3 * H,
                   e_store(PX + e->oO, NX, nullptr, E_ACCUM), false)))
      return rc;
  }
  }

  // ---- recurrences: three independent kernels, one per stream ----
  const int nblk = cdiv(S, RNN_NSEQ);
  {
    if ((rc = fork_aux(e))) return rc;
    // interest_evolve = False: short_term_intention is the user's short embedding itself (clsr.py:169-170)
    if (!e->has_gru1) CK(cudaMemcpyAsync(e->B("sti"), us, (size_t)S * U * 4, cudaMemcpyDeviceToDevice, e->aux[0]));
    if (rnn_tc_ok(e)) {
      // mma.sync kernels: weights in registers, states thread-local (rnn_tc.cuh)
      if (e->has_gru1) {
        rtc::gru_fwd_tc_kernel<40><<<nblk, 160, 0, e->aux[0]>>>(PX, NX, e->oG1, e->oC1, us, W("Wgh1"), W("Wch1"), e->d_len, S, T,
                                                              e->B("g1"), e->B("c1"), e->B("hp1"), e->B("rh1"), e->B("sti"));
        e->launches++;
      }
      if (e->has_gru2) {
        rtc::gru_fwd_tc_kernel<40><<<nblk, 160, 0, e->aux[1]>>>(PX, NX, e->oG2, e->oC2, nullptr, W("Wgh2"), W("Wch2"), e->d_len, S,
                                                              T, e->B("g2"), e->B("c2"), e->B("hp2"), e->B("rh2"), e->B("fs"));
        e->launches++;
      }
      rtc::lstm_fwd_tc_kernel<40><<<nblk, 160, 0, st>>>(PX, NX, e->oL, e->oTN, e->oTL, W("Km"), e->d_len, S, T, e->B("G4"),
                                                      e->B("cp"), e->B("mp"), e->B("R"));
      e->launches++;
    } else {
    // shared memory: resident recurrent weights + state tiles, or (wide states) the state tiles only
    const int wg = e->rnn_wglob;
    size_t smg = (size_t)((wg ? 0 : U * 2 * U + U * U) + 3 * U * RNN_LD) * 4;
    if (e->has_gru1) {
      gru_fwd_kernel<<<nblk, RNN_THREADS, smg, e->aux[0]>>>(PX, NX, e->oG1, e->oC1, us, W("Wgh1"), W("Wch1"), e->d_len, S, T, U,
                                                         e->B("g1"), e->B("c1"), e->B("hp1"), e->B("rh1"), e->B("sti"), wg);
      e->launches++;
    }
    size_t smg2 = (size_t)((wg ? 0 : H * 2 * H + H * H) + 3 * H * RNN_LD) * 4;
    if (e->has_gru2) {
      gru_fwd_kernel<<<nblk, RNN_THREADS, smg2, e->aux[1]>>>(PX, NX, e->oG2, e->oC2, nullptr, W("Wgh2"), W("Wch2"), e->d_len, S, T, H,
                                                          e->B("g2"), e->B("c2"), e->B("hp2"), e->B("rh2"), e->B("fs"), wg);
      e->launches++;
    }
    size_t sml = (size_t)((wg ? 0 : H * 4 * H) + 2 * H * RNN_LD + 4 * H * RNN_LD) * 4;
    lstm_fwd_kernel<<<nblk, RNN_THREADS, sml, st>>>(PX, NX, e->oL, e->oTN, e->oTL, W("Km"), e->d_len, S, T, H,
                                                    e->B("G4"), e->B("cp"), e->B("mp"), e->B("R"), wg);
    e->launches++;
    }
    if ((rc = join_aux(e, "rnn_fwd(gru_sti|gru_causal2|time4lstm)"))) return rc;
  }

  // ---- long-term attention (all per sequence) ----
  Mlp& ml = e->mlp_long;
  float *al = e->B("al"), *qbl = e->B("qbl"), *h0l = e->B("h0l"), *h1l = e->B("h1l");
  if ((rc = gemm(e, "al", (int)M, U, D, a_plain(X, D), e->P + e->p_wattl, U, e_store(al, U), false))) return rc;
  if ((rc = gemm(e, "qbl", S, A0, U, a_plain(ul, U), W("Wl0q"), A0, e_store(qbl, A0, e->P + ml.b0), false))) return rc;
  {
    EpiOp ep = e_store(h0l, A0, nullptr, E_ROWBIAS);
    ep.rb = qbl; ep.ldrb = A0; ep.rbT = T; ep.stat = ml.bn0.stat_f;
    if ((rc = gemm(e, "h0l", (int)M, A0, 2 * U, a_catmul(al, U, U, 0, ul, U, T), W("Wl0"), A0, ep, train != 0))) return rc;
    if ((rc = bn_fwd(e, ml.bn0, (double)M, train, update_bn))) return rc;
    ep = e_store(h1l, A1, e->P + ml.b1);
    ep.stat = ml.bn1.stat_f;
    if ((rc = gemm(e, "h1l", (int)M, A1, A0, a_bnrelu(h0l, A0, ml.bn0), e->P + ml.w1, A1, ep, train != 0))) return rc;
    if ((rc = bn_fwd(e, ml.bn1, (double)M, train, update_bn))) return rc;
  }
  {
    int wpb = 4;
    size_t sm = (size_t)(3 * A1 + wpb * T) * 4;
    pool_fwd_kernel<<<grid1d(e, cdiv(S, wpb), 1, 8), 128, sm, st>>>(
        h1l, A1, ml.bn1.scale, ml.bn1.shift, e->P + ml.wo, e->P + ml.bo, X, D, e->d_len, S, T, 1, e->B("wl"),
        e->B("afl"), e->B("hm"), e->B("hr"), e->cfg.contrastive_recent_k);
    POST("pool_fwd_long");
  }

  // ---- short-term attention ----
  Mlp& ms = e->mlp_short;
  float *R = e->B("R"), *as = e->B("as"), *invs = e->B("invs"), *qbs = e->B("qbs"), *h0s = e->B("h0s"),
        *h1s = e->B("h1s"), *sti = e->B("sti");
  if ((rc = gemm(e, "as", (int)M, Q, H, a_plain(R, H), e->P + e->p_watts, Q, e_store(as, Q), false))) return rc;
  if ((rc = gemm(e, "invs", (int)M, A0, Q + U, a_catmul(as, Q, Q, 0, sti, U, T), W("Ws0i"), A0, e_store(invs, A0), false)))
    return rc;
  if ((rc = gemm(e, "qbs", B, A0, Q, a_cat2row(sti, U, U, tgt, D, G), W("Ws0q"), A0, e_store(qbs, A0, e->P + ms.b0), false)))
    return rc;
  {
    EpiOp ep = e_store(h0s, A0, nullptr, E_ROWBIAS | E_GROUPADD);
    ep.rb = qbs; ep.ldrb = A0; ep.rbT = T; ep.ga = invs; ep.ldga = A0; ep.T = T; ep.G = G; ep.stat = ms.bn0.stat_f;
    if ((rc = gemm(e, "h0s", (int)MB, A0, D, a_mulrow(as, Q, U, tgt, D, T, G), W("Ws0t"), A0, ep, train != 0))) return rc;
    if ((rc = bn_fwd(e, ms.bn0, (double)MB, train, update_bn))) return rc;
    ep = e_store(h1s, A1, e->P + ms.b1);
    ep.stat = ms.bn1.stat_f;
    if ((rc = gemm(e, "h1s", (int)MB, A1, A0, a_bnrelu(h0s, A0, ms.bn0), e->P + ms.w1, A1, ep, train != 0))) return rc;
    if ((rc = bn_fwd(e, ms.bn1, (double)MB, train, update_bn))) return rc;
  }
  {
    int wpb = 4;
    size_t sm = (size_t)(3 * A1 + wpb * T) * 4;
    pool_fwd_kernel<<<grid1d(e, cdiv(B, wpb), 1, 8), 128, sm, st>>>(
        h1s, A1, ms.bn1.scale, ms.bn1.shift, e->P + ms.wo, e->P + ms.bo, R, H, e->d_len, B, T, G, e->B("ws"),
        e->B("afs"), nullptr, nullptr, e->cfg.contrastive_recent_k);
    POST("pool_fwd_short");
  }

  // ---- alpha gate, fusion, prediction MLP ----
  float *ca = e->B("ca"), *mo = e->B("mo");
  if (e->has_alpha) {
    concat_alpha_kernel<<<grid1d(e, (long long)B * e->CA, 256), 256, 0, st>>>(
        e->B("fs"), tgt, e->B("afl"), e->B("afs"), c.ttn, c.seq_stride, T, H, D, G, B, ca, e->Hf);
    POST("concat_alpha");
    if ((rc = mlp_fwd(e, e->mlp_alpha, ca, B, e->B("ha0"), e->B("ha1"), e->B("alogit"), train, update_bn))) return rc;
  }
  // (manual_alpha: the "alogit" buffer holds logit(manual_alpha_value) in every row since clsr_create)
  head_mid_kernel<<<grid1d(e, (long long)B * (H + D), 256), 256, 0, st>>>(e->B("alogit"), e->B("afl"), e->B("afs"), tgt, H,
                                                                        D, G, B, e->B("alpha"), mo);
  POST("head_mid");
  if ((rc = mlp_fwd(e, e->mlp_logit, mo, B, e->B("hl0"), e->B("hl1"), e->B("logit"), train, update_bn))) return rc;
  return 0;
}

// dinv / dqb reductions of the BatchNorm-backward input gradient (attn.cuh): vectorised stream when the
// layer width allows it.
int h0_reduce(clsr_engine* e, const char* name, const float* dy0, const float* h0, const BnLayer& bn, int S, int T,
              int G, float* dinv, float* dqb) {
  const int A0 = bn.N, nx4 = A0 / 4;
  int ny = cdiv(T, 4);
  if (nx4 > 0 && ny * nx4 > 1024) ny = 1024 / nx4;
  if (ny > 32) ny = 32;
  if (ny < 1) ny = 1;
  const int tp = cdiv(T, ny);
  const size_t smv = (size_t)G * ny * nx4 * 16;
  const bool vec = (A0 & 3) == 0 && tp <= 8 && smv <= 96 * 1024 &&
                   !(((uintptr_t)dy0 | (uintptr_t)h0 | (uintptr_t)dinv | (uintptr_t)dqb | (uintptr_t)bn.al |
                      (uintptr_t)bn.be | (uintptr_t)bn.ga) & 15);
  if (vec && tp <= 4)
    h0_reduce_v4_kernel<4><<<S, dim3(nx4, ny), smv, e->stream>>>(dy0, h0, A0, bn.al, bn.be, bn.ga, T, G, dinv, dqb);
  else if (vec)
    h0_reduce_v4_kernel<8><<<S, dim3(nx4, ny), smv, e->stream>>>(dy0, h0, A0, bn.al, bn.be, bn.ga, T, G, dinv, dqb);
  else {
    int nx = ((A0 + 31) / 32) * 32;
    dim3 blk(nx, 4);
    h0_reduce_kernel<<<S, blk, (size_t)4 * A0 * 4, e->stream>>>(dy0, h0, A0, bn.al, bn.be, bn.ga, T, G, dinv, dqb);
  }
  POST(name);
  return 0;
}

int backward(clsr_engine* e, const StepCtx& c) {
  const int B = c.B, G = c.G, S = c.S, T = c.T;
  const int D = e->D, U = e->U, H = e->H, Q = e->Q, A0 = e->A0, A1 = e->A1, NX = e->NX, Di = e->Di, Dc = e->Dc;
  const int CA = e->CA;
  const long long M = (long long)S * T, MB = (long long)B * T;
  cudaStream_t st = e->stream;
  int rc;
  auto W = [&](const char* n) { return e->Wd + e->wd_off.at(n); };
  auto dW = [&](const char* n) { return e->dWd + e->wd_off.at(n); };
  float *tgt = e->B("tgt"), *X = e->B("X");

  // dPX (every position's pre-activation gradients; the BPTT kernels only write the live ones) is cleared on a
  // side stream while the head and the short attention run: the fill is pure HBM traffic, they are not
  float *PX = e->B("PX"), *dPX = e->B("dPX");
  CK(cudaEventRecord(e->ev_fork, st));
  CK(cudaStreamWaitEvent(e->aux[0], e->ev_fork, 0));
  CK(cudaMemsetAsync(dPX, 0, (size_t)M * NX * 4, e->aux[0]));
  CK(cudaEventRecord(e->ev_memset, e->aux[0]));

  // ---- data loss + prediction MLP ----
  softmax_loss_kernel<<<grid1d(e, B / e->cfg.train_group, 128), 128, 0, st>>>(
      e->B("logit"), c.labels, e->cfg.train_group, B, B * e->world, e->B("dlogit"), e->B("pred"), e->acc);
  POST("softmax_loss");
  if ((rc = mlp_bwd(e, e->mlp_logit, e->B("mo"), B, e->B("hl0"), e->B("hl1"), e->B("dlogit"), e->B("dhl1"), e->B("dhl0"),
                    W("Wg0T"), W("Wg1T"), e->B("dmo"))))
    return rc;
  head_mid_bwd_kernel<<<grid1d(e, (long long)B * 32, 256), 256, 0, st>>>(e->B("dmo"), H + D, e->B("afl"), e->B("afs"),
                                                                       e->B("alpha"), H, G, B, e->B("dalogit"));
  POST("head_mid_bwd");
  // (manual_alpha: no alpha MLP; "dca" is zero since clsr_create and nothing writes it)
  if (e->has_alpha && (rc = mlp_bwd(e, e->mlp_alpha, e->B("ca"), B, e->B("ha0"), e->B("ha1"), e->B("dalogit"), e->B("dha1"),
                                    e->B("dha0"), W("Wa0T"), W("Wa1T"), e->B("dca"))))
    return rc;
  const float* bpr_gs = nullptr;
  if (e->cfg.contrastive_kind == 1) {
    bpr_dots_kernel<<<grid1d(e, (long long)B * 32, 256), 256, 0, st>>>(e->B("afl"), e->B("afs"), e->B("hm"), e->B("hr"), e->d_len,
                                                                     e->cfg.contrastive_len_threshold, D, G, B, e->B("bpr_gs"),
                                                                     e->acc);
    POST("bpr_dots");
    bpr_gs = e->B("bpr_gs");
  }
  head_final_bwd_kernel<<<grid1d(e, (long long)S * D, 128), 128, 0, st>>>(
      e->B("dmo"), e->B("dca"), e->B("alpha"), e->B("afl"), e->B("afs"), e->B("hm"), e->B("hr"), e->d_len, e->counts,
      e->cfg.contrastive_len_threshold, e->cfg.triplet_margin, e->cfg.contrastive_weight, bpr_gs, H, D, G, S, e->B("dfs"),
      e->B("dtgt"), e->B("dafl"), e->B("dafs"), e->B("dhm"), e->B("dhr"), e->acc, e->Hf);
  POST("head_final_bwd");

  // ---- short-term attention ----
  Mlp& ms = e->mlp_short;
  float *R = e->B("R"), *as = e->B("as"), *h0s = e->B("h0s"), *h1s = e->B("h1s"), *sti = e->B("sti");
  float *dy1s = e->B("dy1s"), *dy0s = e->B("dy0s"), *dR = e->B("dR");
  {
    int wpb = 4;
    size_t sm = (size_t)(5 * A1 + 3 * A1 + 4 + wpb * (2 * T + H)) * 4;
    // one warp per ROW (5x the parallelism of one warp per group); the group sum of the values
    // gradient is a separate small kernel
    pool_bwd_kernel<<<grid1d(e, cdiv(B, wpb), 1, 16), 128, sm, st>>>(
        e->B("dafs"), e->B("ws"), R, H, h1s, A1, ms.bn1.scale, ms.bn1.shift, ms.bn1.mean, ms.bn1.rstd, e->P + ms.wo,
        e->d_len, B, T, 1, G, dy1s, ms.bn1.stat_b, e->Pg + ms.wo, e->Pg + ms.bo, nullptr, 0, nullptr, nullptr,
        e->cfg.contrastive_recent_k);
    POST("pool_bwd_short");
    if ((H & 3) == 0 && !(((uintptr_t)e->B("dafs") | (uintptr_t)dR) & 15))
      pool_dv_v4_kernel<<<grid1d(e, M * H / 4, 256, 16), 256, 0, st>>>(e->B("ws"), e->B("dafs"), e->d_len, S, T, G, H, dR);
    else
      pool_dv_kernel<<<grid1d(e, M * H, 256), 256, 0, st>>>(e->B("ws"), e->B("dafs"), e->d_len, S, T, G, H, dR);
    POST("pool_dv_short");
  }
  if ((rc = bn_bwd(e, ms.bn1, (double)MB))) return rc;
  {
    EpiOp ep = e_store(dy0s, A0, nullptr, E_RELUMASK | E_STAT_XHAT);
    ep.hpre = h0s; ep.ldh = A0; ep.scale = ms.bn0.scale; ep.shift = ms.bn0.shift; ep.mean = ms.bn0.mean;
    ep.rstd = ms.bn0.rstd; ep.stat = ms.bn0.stat_b;
    if ((rc = gemm(e, "dy0s", (int)MB, A0, A1, a_affine2(dy1s, h1s, A1, ms.bn1), W("W1sT"), A0, ep, true))) return rc;
  }
  if ((rc = dwgemm(e, "dW1s", (int)MB, A0, A1, a_bnrelu(h0s, A0, ms.bn0), a_affine2(dy1s, h1s, A1, ms.bn1),
                   e->Pg + ms.w1, A1, nullptr)))
    return rc;
  if ((rc = bn_bwd(e, ms.bn0, (double)MB))) return rc;
  if ((rc = h0_reduce(e, "h0_reduce_short", dy0s, h0s, ms.bn0, S, T, G, e->B("dinvs"), e->B("dqbs")))) return rc;
  AOp dh0s = a_affine2(dy0s, h0s, A0, ms.bn0);
  if ((rc = gemm(e, "dP", (int)MB, D, A0, dh0s, W("Ws0tT"), D, e_store(e->B("dP"), D), false))) return rc;
  if ((rc = dwgemm(e, "dWs0t", (int)MB, D, A0, a_mulrow(as, Q, U, tgt, D, T, G), dh0s, dW("Ws0t"), A0, nullptr))) return rc;
  {
    const int nx4 = D / 4;
    int ny = cdiv(T, 4);
    if (ny * nx4 > 1024) ny = 1024 / nx4;
    if (ny > 32) ny = 32;
    const int tp = cdiv(T, ny);
    const size_t smv = (size_t)G * ny * nx4 * 16;
    float *dPb = e->B("dP"), *da2 = e->B("da2"), *dtg = e->B("dtgt");
    const bool vec = (D & 3) == 0 && (Q & 3) == 0 && (U & 3) == 0 && tp <= 8 && smv <= 96 * 1024 &&
                     !(((uintptr_t)dPb | (uintptr_t)as | (uintptr_t)tgt | (uintptr_t)da2 | (uintptr_t)dtg) & 15);
    if (vec && tp <= 4) mulrow_bwd_v4_kernel<4><<<S, dim3(nx4, ny), smv, st>>>(dPb, D, as, Q, U, tgt, D, T, G, da2, dtg, D);
    else if (vec) mulrow_bwd_v4_kernel<8><<<S, dim3(nx4, ny), smv, st>>>(dPb, D, as, Q, U, tgt, D, T, G, da2, dtg, D);
    else {
      int nx = ((D + 31) / 32) * 32;
      dim3 blk(nx, 4);
      mulrow_bwd_kernel<<<S, blk, (size_t)4 * D * 4, st>>>(dPb, D, as, Q, U, tgt, D, T, G, da2, dtg, D);
    }
    POST("mulrow_bwd");
  }
  if ((rc = gemm(e, "dFs", (int)M, Q + U, A0, a_plain(e->B("dinvs"), A0), W("Ws0iT"), Q + U, e_store(e->B("dFs"), Q + U), false)))
    return rc;
  dw_group_begin(e);   // dWs0i, dWs0q, dWatts: launched together once the last one's operands exist
  if ((rc = dwgemm(e, "dWs0i", (int)M, Q + U, A0, a_catmul(as, Q, Q, 0, sti, U, T), a_plain(e->B("dinvs"), A0), dW("Ws0i"),
                   A0, nullptr)))
    return rc;
  {
    int nx = ((Q + 31) / 32) * 32;
    dim3 blk(nx, 256 / nx > 0 ? 256 / nx : 1);
    catmul_bwd_kernel<<<S, blk, (size_t)blk.y * U * 4, st>>>(e->B("dFs"), Q, U, as, Q, sti, U, e->B("da2"), D, U, T,
                                                            e->B("das"), Q, e->B("dsti"), U);
    POST("catmul_bwd_short");
  }
  if ((rc = gemm(e, "dqs", B, Q, A0, a_plain(e->B("dqbs"), A0), W("Ws0qT"), Q, e_store(e->B("dqs"), Q), false))) return rc;
  if ((rc = dwgemm(e, "dWs0q", B, Q, A0, a_cat2row(sti, U, U, tgt, D, G), a_plain(e->B("dqbs"), A0), dW("Ws0q"), A0, nullptr)))
    return rc;
  qs_bwd_kernel<<<grid1d(e, (long long)S * Q, 256), 256, 0, st>>>(e->B("dqs"), U, D, G, S, e->B("dsti"), e->B("dtgt"));
  POST("qs_bwd");
  if ((rc = gemm(e, "dR", (int)M, H, Q, a_plain(e->B("das"), Q), W("WattsT"), H, e_store(dR, H, nullptr, E_ACCUM), false)))
    return rc;
  if ((rc = dwgemm(e, "dWatts", (int)M, H, Q, a_plain(R, H), a_plain(e->B("das"), Q), e->Pg + e->p_watts, Q, nullptr)))
    return rc;
  if ((rc = dw_group_flush(e, "dW_short_group"))) return rc;

  // ---- BPTT through the three recurrences ----
  CK(cudaStreamWaitEvent(st, e->ev_memset, 0));
  const int nblk = cdiv(S, RNN_NSEQ);
  {
    // the three BPTT kernels write disjoint column ranges of dPX: run them side by side
    if ((rc = fork_aux(e))) return rc;
    // interest_evolve = False: the gradient of short_term_intention IS the gradient of the short user rows
    if (!e->has_gru1) CK(cudaMemcpyAsync(e->B("dus"), e->B("dsti"), (size_t)S * U * 4, cudaMemcpyDeviceToDevice, e->aux[0]));
    if (rnn_tc_ok(e)) {
      rtc::lstm_bwd_tc_kernel<40><<<nblk, 160, 0, st>>>(PX, NX, e->oL, e->oTN, e->oTL, e->B("G4"), e->B("cp"), W("KmT"), dR,
                                                      e->d_len, S, T, dPX);
      e->launches++;
      if (e->has_gru1) {
        rtc::gru_bwd_tc_kernel<40><<<nblk, 160, 0, e->aux[0]>>>(e->B("g1"), e->B("c1"), e->B("hp1"), W("Wgh1T"), W("Wch1T"),
                                                              e->B("dsti"), e->d_len, S, T, dPX, NX, e->oG1, e->oC1, e->B("dus"));
        e->launches++;
      }
      if (e->has_gru2) {
        rtc::gru_bwd_tc_kernel<40><<<nblk, 160, 0, e->aux[1]>>>(e->B("g2"), e->B("c2"), e->B("hp2"), W("Wgh2T"), W("Wch2T"),
                                                              e->B("dfs"), e->d_len, S, T, dPX, NX, e->oG2, e->oC2, nullptr);
        e->launches++;
      }
    } else {
    const int wg = e->rnn_wglob;
    size_t sml = (size_t)((wg ? 0 : H * 4 * H) + 2 * H * RNN_LD + 4 * H * RNN_LD) * 4;
    lstm_bwd_kernel<<<nblk, RNN_THREADS, sml, st>>>(PX, NX, e->oL, e->oTN, e->oTL, e->B("G4"), e->B("cp"), W("KmT"), dR,
                                                    e->d_len, S, T, H, dPX, wg);
    e->launches++;
    size_t smg = (size_t)((wg ? 0 : 2 * U * U + U * U) + 5 * U * RNN_LD) * 4;
    if (e->has_gru1) {
      gru_bwd_kernel<<<nblk, RNN_THREADS, smg, e->aux[0]>>>(e->B("g1"), e->B("c1"), e->B("hp1"), W("Wgh1T"), W("Wch1T"), e->B("dsti"),
                                                         e->d_len, S, T, U, dPX, NX, e->oG1, e->oC1, e->B("dus"), wg);
      e->launches++;
    }
    size_t smg2 = (size_t)((wg ? 0 : 2 * H * H + H * H) + 5 * H * RNN_LD) * 4;
    if (e->has_gru2) {
      gru_bwd_kernel<<<nblk, RNN_THREADS, smg2, e->aux[1]>>>(e->B("g2"), e->B("c2"), e->B("hp2"), W("Wgh2T"), W("Wch2T"), e->B("dfs"),
                                                          e->d_len, S, T, H, dPX, NX, e->oG2, e->oC2, nullptr, wg);
      e->launches++;
    }
    }
    if ((rc = join_aux(e, "rnn_bwd(time4lstm|gru_sti|gru_causal2)"))) return rc;
  }
  float *dX = e->B("dX"), *TNL = e->B("TNL"), *dTNL = e->B("dTNL");
  if ((rc = gemm(e, "dX", (int)M, D, NX, a_plain(dPX, NX), W("Wx_allT"), D, e_store(dX, D), false))) return rc;
  // the seven weight gradients fed by dPX are independent of everything up to the optimizer: one grouped launch
  dw_group_begin(e);
  if ((rc = dwgemm(e, "dWx", (int)M, D, NX, a_plain(X, D), a_plain(dPX, NX), dW("Wx_all"), NX, dW("bx_all")))) return rc;
  if (!e->plain_lstm) {
  if ((rc = gemm(e, "dTNL", (int)M, 2 * H, 3 * H, a_plain(dPX + e->oO, NX), W("WtT"), 2 * H, e_store(dTNL, 2 * H), false)))
    return rc;
  if ((rc = dwgemm(e, "dWt", (int)M, 2 * H, 3 * H, a_plain(TNL, 2 * H), a_plain(dPX + e->oO, NX), dW("Wt"), 3 * H, nullptr)))
    return rc;
  {
    int ppc = cdiv(M, (long long)e->num_sms * 4);
    if ((H & 3) == 0 && 2 * H / 4 <= 32) {
      const int q4 = 2 * H / 4, ny = 1024 / q4 > 32 ? 32 : 1024 / q4;
      time_feat_bwd_v4_kernel<<<cdiv(M, ppc), dim3(q4, ny), (size_t)2 * ny * 2 * H * 4, st>>>(
          dTNL, TNL, c.ttn, c.tfa, c.seq_stride, T, H, M, ppc, e->Pg + e->p_tw1, e->Pg + e->p_tb1, e->Pg + e->p_tw2,
          e->Pg + e->p_tb2);
    } else {
      int nx = ((2 * H + 31) / 32) * 32, ny = 1024 / nx > 8 ? 8 : (1024 / nx > 0 ? 1024 / nx : 1);
      time_feat_bwd_kernel<<<cdiv(M, ppc), dim3(nx, ny), (size_t)2 * ny * 2 * H * 4, st>>>(
          dTNL, TNL, c.ttn, c.tfa, c.seq_stride, T, H, M, ppc, e->Pg + e->p_tw1, e->Pg + e->p_tb1, e->Pg + e->p_tw2,
          e->Pg + e->p_tb2);
    }
    POST("time_feat_bwd");
  }
  }
  if (e->has_gru1) {
    if ((rc = dwgemm(e, "dWgh1", (int)M, U, 2 * U, a_plain(e->B("hp1"), U), a_plain(dPX + e->oG1, NX), dW("Wgh1"), 2 * U, nullptr))) return rc;
    if ((rc = dwgemm(e, "dWch1", (int)M, U, U, a_plain(e->B("rh1"), U), a_plain(dPX + e->oC1, NX), dW("Wch1"), U, nullptr))) return rc;
  }
  if (e->has_gru2) {
    if ((rc = dwgemm(e, "dWgh2", (int)M, H, 2 * H, a_plain(e->B("hp2"), H), a_plain(dPX + e->oG2, NX), dW("Wgh2"), 2 * H, nullptr))) return rc;
    if ((rc = dwgemm(e, "dWch2", (int)M, H, H, a_plain(e->B("rh2"), H), a_plain(dPX + e->oC2, NX), dW("Wch2"), H, nullptr))) return rc;
  }
  if ((rc = dwgemm(e, "dKm", (int)M, H, 4 * H, a_plain(e->B("mp"), H), a_plain(dPX + e->oL, NX), dW("Km"), 4 * H, nullptr))) return rc;
  if ((rc = dw_group_flush(e, "dW_bptt_group"))) return rc;

  // ---- long-term attention ----
  Mlp& ml = e->mlp_long;
  float *al = e->B("al"), *ul = e->B("ul"), *h0l = e->B("h0l"), *h1l = e->B("h1l");
  float *dy1l = e->B("dy1l"), *dy0l = e->B("dy0l");
  {
    int wpb = 4;
    size_t sm = (size_t)(5 * A1 + 3 * A1 + 4 + wpb * (2 * T + D)) * 4;
    pool_bwd_kernel<<<grid1d(e, cdiv(S, wpb), 1, 8), 128, sm, st>>>(
        e->B("dafl"), e->B("wl"), X, D, h1l, A1, ml.bn1.scale, ml.bn1.shift, ml.bn1.mean, ml.bn1.rstd, e->P + ml.wo,
        e->d_len, S, T, 1, 1, dy1l, ml.bn1.stat_b, e->Pg + ml.wo, e->Pg + ml.bo, dX, 1, e->B("dhm"), e->B("dhr"),
        e->cfg.contrastive_recent_k);
    POST("pool_bwd_long");
  }
  if ((rc = bn_bwd(e, ml.bn1, (double)M))) return rc;
  {
    EpiOp ep = e_store(dy0l, A0, nullptr, E_RELUMASK | E_STAT_XHAT);
    ep.hpre = h0l; ep.ldh = A0; ep.scale = ml.bn0.scale; ep.shift = ml.bn0.shift; ep.mean = ml.bn0.mean;
    ep.rstd = ml.bn0.rstd; ep.stat = ml.bn0.stat_b;
    if ((rc = gemm(e, "dy0l", (int)M, A0, A1, a_affine2(dy1l, h1l, A1, ml.bn1), W("W1lT"), A0, ep, true))) return rc;
  }
  dw_group_begin(e);   // dW1l, dWl0, dWl0q, dWattl
  if ((rc = dwgemm(e, "dW1l", (int)M, A0, A1, a_bnrelu(h0l, A0, ml.bn0), a_affine2(dy1l, h1l, A1, ml.bn1),
                   e->Pg + ml.w1, A1, nullptr)))
    return rc;
  if ((rc = bn_bwd(e, ml.bn0, (double)M))) return rc;
  if ((rc = h0_reduce(e, "h0_reduce_long", dy0l, h0l, ml.bn0, S, T, 1, nullptr, e->B("dqbl")))) return rc;
  AOp dh0l = a_affine2(dy0l, h0l, A0, ml.bn0);
  if ((rc = gemm(e, "dFl", (int)M, 2 * U, A0, dh0l, W("Wl0T"), 2 * U, e_store(e->B("dFl"), 2 * U), false))) return rc;
  if ((rc = dwgemm(e, "dWl0", (int)M, 2 * U, A0, a_catmul(al, U, U, 0, ul, U, T), dh0l, dW("Wl0"), A0, nullptr))) return rc;
  {
    int nx = ((U + 31) / 32) * 32;
    dim3 blk(nx, 256 / nx > 0 ? 256 / nx : 1);
    catmul_bwd_kernel<<<S, blk, (size_t)blk.y * U * 4, st>>>(e->B("dFl"), U, U, al, U, ul, U, nullptr, 0, 0, T, e->B("dal"),
                                                            U, e->B("dul"), U);
    POST("catmul_bwd_long");
  }
  if ((rc = gemm(e, "dul", S, U, A0, a_plain(e->B("dqbl"), A0), W("Wl0qT"), U, e_store(e->B("dul"), U, nullptr, E_ACCUM), false)))
    return rc;
  if ((rc = dwgemm(e, "dWl0q", S, U, A0, a_plain(ul, U), a_plain(e->B("dqbl"), A0), dW("Wl0q"), A0, nullptr))) return rc;
  if ((rc = gemm(e, "dX_al", (int)M, D, U, a_plain(e->B("dal"), U), W("WattlT"), D, e_store(dX, D, nullptr, E_ACCUM), false)))
    return rc;
  if ((rc = dwgemm(e, "dWattl", (int)M, D, U, a_plain(X, D), a_plain(e->B("dal"), U), e->Pg + e->p_wattl, U, nullptr)))
    return rc;
  if ((rc = dw_group_flush(e, "dW_long_group"))) return rc;
  (void)Di; (void)Dc; (void)CA;
  return 0;
}

// Replicated-table data parallelism only: staging for the all-gathered sparse-gradient inputs and room for the
// global unique sets (allocated on first use; the row-sharded mode needs none of it).
int alloc_replicated_staging(clsr_engine* e) {
  if (e->g_ih) return 0;
  const long long W = e->world, Bm = e->Bmax, Sm = e->Smax, T = e->T, M = Sm * T;
  int rc;
  if ((rc = dalloc(e, &e->l_ih, M)) || (rc = dalloc(e, &e->l_ch, M)) || (rc = dalloc(e, &e->l_users, Sm))) return rc;
  if ((rc = dalloc(e, &e->g_ih, W * M)) || (rc = dalloc(e, &e->g_ch, W * M)) || (rc = dalloc(e, &e->g_users, W * Sm)) ||
      (rc = dalloc(e, &e->g_items, W * Bm)) || (rc = dalloc(e, &e->g_cates, W * Bm)))
    return rc;
  if ((rc = dalloc(e, &e->g_dX, W * M * e->D)) || (rc = dalloc(e, &e->g_dtgt, W * Bm * e->D)) ||
      (rc = dalloc(e, &e->g_dul, W * Sm * e->U)) || (rc = dalloc(e, &e->g_dus, W * Sm * e->U)))
    return rc;
  const long long rows3[3] = {e->cfg.n_items, e->cfg.n_cates, e->cfg.n_users};
  const long long want[3] = {W * (M + Bm), W * (M + Bm), W * Sm};
  for (int i = 0; i < 3; ++i) {
    long long cap = want[i] < rows3[i] ? want[i] : rows3[i];
    if (cap > e->uniq_cap[i]) {
      e->uniq_cap[i] = cap;
      if ((rc = dalloc(e, &e->uniq[i], cap))) return rc;
    }
  }
  if ((rc = dalloc(e, &e->cg[0], e->uniq_cap[0] * e->Di)) || (rc = dalloc(e, &e->cg[1], e->uniq_cap[1] * e->Dc)) ||
      (rc = dalloc(e, &e->cg[2], e->uniq_cap[2] * e->U)) || (rc = dalloc(e, &e->cg[3], e->uniq_cap[2] * e->U)))
    return rc;
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// K2 + K13: unique ids per table and scatter-add of every slice gradient into compact rows.
// Data-parallel runs first all-gather the raw ids and per-position gradients of every rank, so each
// replica builds the identical global (unique ids, summed rows) set and applies the identical update.
int sparse_grads(clsr_engine* e, const StepCtx& c) {
  const int T = c.T, U = e->U, D = e->D, Di = e->Di, Dc = e->Dc;
  int B = c.B, S = c.S;
  long long M = (long long)S * T;
  cudaStream_t st = e->stream;
  const int32_t *ih = c.ih, *ch = c.ch, *items = c.items, *cates = c.cates, *users = c.users;
  int seq_stride = c.seq_stride, user_stride = c.user_stride;
  const float *dX = e->B("dX"), *dtgt = e->B("dtgt"), *dul = e->B("dul"), *dus = e->B("dus");
  int rc;
  if (e->world > 1 && !e->sharded) {
    if ((rc = alloc_replicated_staging(e))) return rc;
    compact_ids_kernel<<<grid1d(e, M, 256), 256, 0, st>>>(c.ih, T, c.seq_stride, M, e->l_ih);
    POST("compact_ids");
    compact_ids_kernel<<<grid1d(e, M, 256), 256, 0, st>>>(c.ch, T, c.seq_stride, M, e->l_ch);
    POST("compact_ids");
    compact_ids_kernel<<<grid1d(e, S, 256), 256, 0, st>>>(c.users, 1, c.user_stride, S, e->l_users);
    POST("compact_ids");
    if ((rc = allgather(e, e->l_ih, e->g_ih, (size_t)M, kNcclInt32))) return rc;
    if ((rc = allgather(e, e->l_ch, e->g_ch, (size_t)M, kNcclInt32))) return rc;
    if ((rc = allgather(e, e->l_users, e->g_users, (size_t)S, kNcclInt32))) return rc;
    if ((rc = allgather(e, c.items, e->g_items, (size_t)B, kNcclInt32))) return rc;
    if ((rc = allgather(e, c.cates, e->g_cates, (size_t)B, kNcclInt32))) return rc;
    if ((rc = allgather(e, dX, e->g_dX, (size_t)M * D, kNcclFloat32))) return rc;
    if ((rc = allgather(e, dtgt, e->g_dtgt, (size_t)B * D, kNcclFloat32))) return rc;
    if ((rc = allgather(e, dul, e->g_dul, (size_t)S * U, kNcclFloat32))) return rc;
    if ((rc = allgather(e, dus, e->g_dus, (size_t)S * U, kNcclFloat32))) return rc;
    ih = e->g_ih; ch = e->g_ch; items = e->g_items; cates = e->g_cates; users = e->g_users;
    dX = e->g_dX; dtgt = e->g_dtgt; dul = e->g_dul; dus = e->g_dus;
    seq_stride = T; user_stride = 1;
    B *= e->world; S *= e->world; M *= e->world;
  }
  {
    // tf.unique of the five id arrays (history + target items, history + target categories, users): one launch
    UniqueMulti um;
    um.n = 5;
    um.s[0] = UniqueSeg{ih, M, T, seq_stride, e->slot[0], e->uniq[0], e->counts + 1};
    um.s[1] = UniqueSeg{items, B, 1, 1, e->slot[0], e->uniq[0], e->counts + 1};
    um.s[2] = UniqueSeg{ch, M, T, seq_stride, e->slot[1], e->uniq[1], e->counts + 2};
    um.s[3] = UniqueSeg{cates, B, 1, 1, e->slot[1], e->uniq[1], e->counts + 2};
    um.s[4] = UniqueSeg{users, S, 1, user_stride, e->slot[2], e->uniq[2], e->counts + 3};
    unique_plan(&um);
    mark_unique_multi_kernel<<<um.blk0[um.n], kUniqueBlock, 0, st>>>(um);
    POST("unique(items|cates|users)");
    CompactMulti cm;
    cm.n = 4;
    const int cnt_ix0[4] = {1, 2, 3, 3};
    for (int t = 0; t < 4; ++t) cm.s[t] = CompactSeg{e->cg[t], e->counts + cnt_ix0[t], e->tab_dim[t], nullptr, nullptr};
    zero_compact_multi_kernel<<<e->num_sms * 2, 256, 0, st>>>(cm);
    POST("zero_compact");
  }
  {
    const int hot = scatter_hot_rows(e);
    scatter_hist_kernel<<<e->num_sms * 4, 256, (size_t)(hot * Di + Dc) * 4, st>>>(dX, ih, ch, seq_stride, T, e->slot[0], e->slot[1],
                                                                                e->cg[0], e->cg[1], Di, Dc, M, e->sumsq, hot);
    POST("scatter_hist");
  }
  {
    ScatterRowsMulti sm;
    sm.n = 4;
    sm.s[0] = ScatterRowsSeg{dtgt, D, 0, Di, items, 1, e->slot[0], e->cg[0], B, e->sumsq + 0};
    sm.s[1] = ScatterRowsSeg{dtgt, D, Di, Dc, cates, 1, e->slot[1], e->cg[1], B, e->sumsq + 1};
    sm.s[2] = ScatterRowsSeg{dul, U, 0, U, users, user_stride, e->slot[2], e->cg[2], S, e->sumsq + 2};
    sm.s[3] = ScatterRowsSeg{dus, U, 0, U, users, user_stride, e->slot[2], e->cg[3], S, e->sumsq + 3};
    scatter_rows_multi_kernel<<<grid1d(e, (long long)B * Di / 4, 256), 256, 0, st>>>(sm);
    POST("scatter_rows(tgt|users)");
  }
  const float l2 = e->cfg.embed_l2, dw = e->cfg.discrepancy_weight;
  if (e->sharded) {
    // Row-sharded tables: every unique compact row goes ONCE to its owner (NVLink reductions into the owner's
    // dense gradient shard + a touched mark); after a barrier the owner counts its touched user rows (the global
    // tf.unique count of the discrepancy mean is their sum over ranks) and adds the involved-row terms.
    const int cnt_ix[4] = {1, 2, 3, 3}, slot_ix[4] = {0, 1, 2, 2};
    for (int t = 0; t < 4; ++t) {
      push_compact_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->uniq[slot_ix[t]], e->counts + cnt_ix[t], e->cg[t], e->tab_dim[t],
                                                          e->gview[t]);
      POST("push_compact");
    }
    if ((rc = peer_reduce(e, nullptr, 0, nullptr, 0, nullptr, 0))) return rc;   // every rank's pushes have landed
    shard_count_touched_kernel<<<e->num_sms, 256, 0, st>>>(e->sh_touched[2], e->tab_local_rows[2], e->counts + 4);
    POST("shard_count_touched");
    if ((rc = peer_reduce(e, nullptr, 0, nullptr, 0, e->counts + 4, 1))) return rc;
    shard_involved_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->sh_val[0], nullptr, Di, e->sh_touched[0], e->tab_local_rows[0],
                                                          e->sh_g[0], l2, 0.f, e->counts + 4, 0, e->sumsq + 0, e->acc + 5);
    POST("involved_item");
    shard_involved_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->sh_val[1], nullptr, Dc, e->sh_touched[1], e->tab_local_rows[1],
                                                          e->sh_g[1], l2, 0.f, e->counts + 4, 0, e->sumsq + 1, e->acc + 6);
    POST("involved_cate");
    shard_involved_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->sh_val[2], e->sh_val[3], U, e->sh_touched[2], e->tab_local_rows[2],
                                                          e->sh_g[2], l2, dw, e->counts + 4, 0, e->sumsq + 2, e->acc + 7);
    POST("involved_ul");
    shard_involved_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->sh_val[3], e->sh_val[2], U, e->sh_touched[2], e->tab_local_rows[3],
                                                          e->sh_g[3], l2, dw, e->counts + 4, 1, e->sumsq + 3, e->acc + 8);
    POST("involved_us");
    return 0;
  }
  {
    // acc[5] item rows^2, acc[6] cate rows^2, acc[7] long rows^2, acc[8] short rows^2, acc[9] sum (long-short)^2
    // (written by the short-table segment only: its acc base is acc+8, its discrepancy slot acc+8+1)
    InvolvedMulti im;
    im.n = 4;
    im.s[0] = InvolvedSeg{e->tab[0], nullptr, Di, e->uniq[0], e->counts + 1, e->cg[0], l2, 0.f, 0, e->sumsq + 0, e->acc + 5};
    im.s[1] = InvolvedSeg{e->tab[1], nullptr, Dc, e->uniq[1], e->counts + 2, e->cg[1], l2, 0.f, 0, e->sumsq + 1, e->acc + 6};
    im.s[2] = InvolvedSeg{e->tab[2], e->tab[3], U, e->uniq[2], e->counts + 3, e->cg[2], l2, dw, 0, e->sumsq + 2, e->acc + 7};
    im.s[3] = InvolvedSeg{e->tab[3], e->tab[2], U, e->uniq[2], e->counts + 3, e->cg[3], l2, dw, 1, e->sumsq + 3, e->acc + 8};
    involved_multi_kernel<<<e->num_sms * 2, 256, 0, st>>>(im);
    POST("involved(4 tables)");
  }
  return 0;
}

// Sharded mode: zero the touched marks (and, after a gradient-only step, the gradient rows) of the owner's shards.
int shard_cleanup(clsr_engine* e, bool grads_too) {
  AdamHyper hz = {0.f, 0.f, 0.f, 1.f, 0.f, nullptr};
  for (int t = 0; t < 4; ++t) {
    const int tt = t == 3 ? 2 : t;   // the two user tables share one touched array
    if (grads_too) {
      adam_lazy_shard_kernel<<<e->num_sms * 4, 256, 0, e->stream>>>(e->sh_val[t], e->sh_m[t], e->sh_v[t], e->sh_g[t], e->sh_touched[tt],
                                                                   e->tab_dim[t], e->tab_local_rows[t], hz, e->sumsq + t, 0);
      POST("shard_zero_grad");
    }
  }
  for (int t = 0; t < 3; ++t) CK(cudaMemsetAsync(e->sh_touched[t], 0, (size_t)e->tab_local_rows[t] * 4, e->stream));
  return 0;
}

int optimizer_step(clsr_engine* e) {
  cudaStream_t st = e->stream;
  const clsr_config& cf = e->cfg;
  // the step size of this step sits in e->d_lr (step_scalars)
  const float lr_t = 0.f;
  AdamDense hd = {lr_t, cf.beta1, cf.beta2, cf.adam_eps, cf.clip_norm ? cf.max_grad_norm : 0.f, e->d_lr};
  dense_adam_kernel<<<(int)e->dense.size(), 1024, 0, st>>>(e->d_vars, e->P, e->Pm, e->Pv, e->Pg, e->d_norms, hd);
  POST("dense_adam");
  AdamHyper hp = {lr_t, cf.beta1, cf.beta2, cf.adam_eps, cf.clip_norm ? cf.max_grad_norm : 0.f, e->d_lr};
  const int slot_ix[4] = {0, 1, 2, 2}, cnt_ix[4] = {1, 2, 3, 3};
  for (int tb = 0; tb < 4; ++tb)
    if (!e->tab_m[tb] || !e->tab_v[tb]) return fail(e, CLSR_ERR_STATE, "Adam slots of table %d not bound", tb);
  if (!e->sharded && cf.optimizer == 0) {
    // TF's non-lazy sweep of all four tables in one launch (UN = 2 vectors in flight per thread, 16 CTAs per SM:
    // 0.9 of the measured HBM peak)
    SweepMulti sw;
    sw.n = 4;
    for (int tb = 0; tb < 4; ++tb)
      sw.s[tb] = SweepSeg{e->tab[tb], e->tab_m[tb], e->tab_v[tb], e->slot[slot_ix[tb]], e->cg[tb], e->tab_dim[tb], e->tab_rows[tb],
                          e->sumsq + tb};
    adam_sweep_multi_kernel<2><<<e->num_sms * 16, 256, 0, st>>>(sw, hp);
    POST("adam_sweep");
    return 0;
  }
  for (int tb = 0; tb < 4; ++tb) {
    if (e->sharded) {   // the owner updates its 1/world of the rows from its dense gradient shard
      const int tt = tb == 3 ? 2 : tb;
      if (cf.optimizer == 1) {
        adam_lazy_shard_kernel<<<e->num_sms * 4, 256, 0, st>>>(e->sh_val[tb], e->sh_m[tb], e->sh_v[tb], e->sh_g[tb], e->sh_touched[tt],
                                                             e->tab_dim[tb], e->tab_local_rows[tb], hp, e->sumsq + tb, 1);
        POST("adam_lazy");
      } else {
        adam_sweep_shard_kernel<2><<<e->num_sms * 16, 256, 0, st>>>(e->sh_val[tb], e->sh_m[tb], e->sh_v[tb], e->sh_g[tb],
                                                                   e->sh_touched[tt], e->tab_dim[tb], e->tab_local_rows[tb], hp,
                                                                   e->sumsq + tb);
        POST("adam_sweep");
      }
      continue;
    }
    if (cf.optimizer == 1) {
      adam_lazy_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->tab[tb], e->tab_m[tb], e->tab_v[tb], e->uniq[slot_ix[tb]],
                                                       e->counts + cnt_ix[tb], e->cg[tb], e->tab_dim[tb], hp, e->sumsq + tb);
      POST("adam_lazy");
    } else {
      // launch shape measured on B200 (UN = 2 vectors in flight per thread, 16 CTAs per SM): 0.89 of the HBM peak
      adam_sweep_kernel<2><<<e->num_sms * 16, 256, 0, st>>>(e->tab[tb], e->tab_m[tb], e->tab_v[tb], e->slot[slot_ix[tb]],
                                                           e->cg[tb], e->tab_dim[tb], e->tab_rows[tb], hp, e->sumsq + tb);
      POST("adam_sweep");
    }
  }
  return 0;
}

int reset_slots(clsr_engine* e) {
  CompactMulti cm;
  cm.n = 3;
  for (int i = 0; i < 3; ++i) cm.s[i] = CompactSeg{nullptr, e->counts + 1 + i, 0, e->uniq[i], e->slot[i]};
  reset_slots_multi_kernel<<<e->num_sms, 256, 0, e->stream>>>(cm);
  POST("reset_slots");
  return 0;
}

int zero_step_state(clsr_engine* e) {
  // counters, loss / norm accumulators, the dense-gradient block (Pg | dWd) and the 16 BatchNorm sum vectors: one launch
  // instead of 20 memset nodes
  ZeroMulti z;
  z.n = 0;
  auto add = [&](void* p, size_t bytes) { z.p[z.n] = p; z.bytes[z.n] = (long long)bytes; z.n++; };
  add(e->counts, 8 * sizeof(int32_t));
  add(e->sumsq, 4 * sizeof(double));
  add(e->acc, 16 * sizeof(double));
  add(e->Pg, (size_t)(e->Ptot + e->Wtot) * 4);
  Mlp* ms[4] = {&e->mlp_long, &e->mlp_short, &e->mlp_alpha, &e->mlp_logit};
  for (Mlp* m : ms) {
    BnLayer* bl[2] = {&m->bn0, &m->bn1};
    for (BnLayer* b : bl) {
      add(b->stat_f, 2 * b->N * sizeof(double));
      add(b->stat_b, 2 * b->N * sizeof(double));
    }
  }
  zero_multi_kernel<<<e->num_sms, 256, 0, e->stream>>>(z);
  e->launches++;
  {
    cudaError_t c = cudaGetLastError();
    if (c != cudaSuccess) return fail(e, CLSR_ERR_CUDA, "zero_multi_kernel failed: %s", cudaGetErrorString(c));
  }
  MARK("zero_state");
  return 0;
}

void drop_graphs(clsr_engine* e) {
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* clsr_last_error(const clsr_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int clsr_create(const clsr_config* cfg, clsr_engine** out) {
  clsr_engine* e = nullptr;
  if (!cfg || !out) return fail(e, CLSR_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->hidden != cfg->item_dim + cfg->cate_dim)
    return fail(e, CLSR_ERR_ARG, "hidden (%d) must equal item_dim + cate_dim (%d)", cfg->hidden, cfg->item_dim + cfg->cate_dim);
  if (cfg->item_dim % 4 || cfg->cate_dim % 4 || cfg->user_dim % 4)
    return fail(e, CLSR_ERR_ARG, "embedding dims must be multiples of 4 (16-byte rows)");
  if (cfg->max_rows <= 0 || cfg->seq_len <= 0 || cfg->train_group <= 0) return fail(e, CLSR_ERR_ARG, "bad sizes");
  if (cfg->att1 > 128) return fail(e, CLSR_ERR_ARG, "att_fcn_layer_sizes[1] > 128 unsupported");
  if (cfg->contrastive_kind != 0 && cfg->contrastive_kind != 1) return fail(e, CLSR_ERR_ARG, "contrastive_kind must be 0 (triplet) or 1 (bpr)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(e, CLSR_ERR_CUDA, "no CUDA device: clsr_b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(e, CLSR_ERR_ARG, "bad device %d", cfg->device);
  e = new clsr_engine();
  e->cfg = *cfg;
#define CKC(call)                                                                  \
  do {                                                                             \
    int _rc = (call);                                                              \
    if (_rc) { g_create_error = e->err; clsr_destroy(e); return _rc; }             \
  } while (0)
#define CKCU(call)                                                                                       \
  do {                                                                                                   \
    cudaError_t _c = (call);                                                                             \
    if (_c != cudaSuccess) {                                                                             \
      fail(nullptr, CLSR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_c));                      \
      clsr_destroy(e);                                                                                   \
      return CLSR_ERR_CUDA;                                                                              \
    }                                                                                                    \
  } while (0)
  CKCU(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CKCU(cudaGetDeviceProperties(&prop, cfg->device));
  e->num_sms = prop.multiProcessorCount;
  e->smem_optin = (int)prop.sharedMemPerBlockOptin;
  CKCU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  e->own_stream = true;
  for (int i = 0; i < 2; ++i) {
    CKCU(cudaStreamCreateWithFlags(&e->aux[i], cudaStreamNonBlocking));
    CKCU(cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming));
  }
  CKCU(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CKCU(cudaEventCreateWithFlags(&e->ev_memset, cudaEventDisableTiming));

  e->T = cfg->seq_len; e->Di = cfg->item_dim; e->Dc = cfg->cate_dim; e->D = e->Di + e->Dc; e->U = cfg->user_dim;
  e->H = cfg->hidden; e->Q = e->U + e->D; e->A0 = cfg->att0; e->A1 = cfg->att1; e->L0 = cfg->fc0; e->L1 = cfg->fc1;
  e->plain_lstm = cfg->sequential_model == 1;
  e->has_gru1 = !cfg->no_interest_evolve;
  e->has_alpha = !cfg->manual_alpha;
  e->has_gru2 = e->has_alpha && !cfg->no_predict_long_short;
  e->Hf = e->has_gru2 ? e->H : 0;
  e->CA = e->Hf + 2 * e->D + e->H + 1;
  const int U = e->U, H = e->H, D = e->D, T = e->T;
  e->oG1 = 0; e->oC1 = 2 * U; e->oG2 = 3 * U; e->oC2 = 3 * U + 2 * H; e->oL = 3 * U + 3 * H;
  e->oO = e->oL + 3 * H; e->oTN = e->oL + 4 * H; e->oTL = e->oL + 5 * H; e->NX = e->oL + 6 * H;
  e->Bmax = cfg->max_rows;
  // per-sequence buffers: sized for ungrouped batches of max_rows unless the caller bounds the sequences
  e->Smax = (cfg->max_seqs > 0 && cfg->max_seqs < cfg->max_rows) ? cfg->max_seqs : cfg->max_rows;
  e->tab_rows[0] = cfg->n_items; e->tab_rows[1] = cfg->n_cates; e->tab_rows[2] = cfg->n_users; e->tab_rows[3] = cfg->n_users;
  e->tab_dim[0] = e->Di; e->tab_dim[1] = e->Dc; e->tab_dim[2] = U; e->tab_dim[3] = U;

  {
    const int UH = U > H ? U : H;
    size_t sml = (size_t)(H * 4 * H + 6 * H * RNN_LD) * 4;
    size_t smg = (size_t)(3 * UH * UH + 5 * UH * RNN_LD) * 4;
    size_t smax = sml > smg ? sml : smg;
    if (smax > (size_t)prop.sharedMemPerBlockOptin) {
      // wide states (BASELINE configs 4-5): the recurrent weights stay in global memory (L1 / L2 resident)
      e->rnn_wglob = 1;
      sml = (size_t)(6 * H * RNN_LD) * 4;
      smg = (size_t)(5 * UH * RNN_LD) * 4;
    }
    CKCU(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml));
    CKCU(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml));
    CKCU(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smg));
    CKCU(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smg));
    size_t smp = (size_t)(8 * e->A1 + 4 + 4 * (2 * T + D)) * 4;
    CKCU(cudaFuncSetAttribute(pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smp > 49152 ? smp : 49152)));
    size_t smf = (size_t)(3 * e->A1 + 4 * T) * 4;
    CKCU(cudaFuncSetAttribute(pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smf > 49152 ? smf : 49152)));
  }

  CKCU(cudaFuncSetAttribute(h0_reduce_v4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CKCU(cudaFuncSetAttribute(h0_reduce_v4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CKCU(cudaFuncSetAttribute(mulrow_bwd_v4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CKCU(cudaFuncSetAttribute(mulrow_bwd_v4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  {
    cudaFuncAttributes fa;
    CKCU(cudaFuncGetAttributes(&fa, tc::tc_gemm_kernel<true>));
    e->tc_smem_max = e->smem_optin - (int)fa.sharedSizeBytes - 256;
    CKCU(cudaFuncSetAttribute(tc::tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->tc_smem_max));
    CKCU(cudaFuncSetAttribute(tc::tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->tc_smem_max));
    CKCU(cudaFuncGetAttributes(&fa, tc::tc_dw_kernel));
    e->tc_dw_smem_max = e->smem_optin - (int)fa.sharedSizeBytes - 256;
    CKCU(cudaFuncSetAttribute(tc::tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, e->tc_dw_smem_max));
    CKCU(cudaFuncSetAttribute(tc::tc_dw_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, e->tc_dw_smem_max));
  }
  build_inventory(e);
  CKC(dalloc(e, &e->P, e->Ptot));
  CKC(dalloc(e, &e->Pm, e->Ptot));
  CKC(dalloc(e, &e->Pv, e->Ptot));
  // Pg is allocated together with dWd (build_weight_maps): [Pg | dWd] is one block for memset / all-reduce
  {
    std::vector<DenseVar> dv;
    for (auto& d : e->dense) dv.push_back(DenseVar{d.off, (int)d.n, d.trainable});
    CKC(dalloc(e, &e->d_vars, (long long)dv.size(), false));
    CKCU(cudaMemcpy(e->d_vars, dv.data(), dv.size() * sizeof(DenseVar), cudaMemcpyHostToDevice));
    CKC(dalloc(e, &e->d_norms, (long long)dv.size()));
  }
  CKC(build_weight_maps(e));
  Mlp* mls[4] = {&e->mlp_long, &e->mlp_short, &e->mlp_alpha, &e->mlp_logit};
  for (Mlp* m : mls) { CKC(alloc_bn(e, &m->bn0)); CKC(alloc_bn(e, &m->bn1)); }

  // slot tables and compact gradient rows
  const long long Bm = e->Bmax, Sm = e->Smax, M = Sm * T, MB = Bm * T;
  const long long rows3[3] = {cfg->n_items, cfg->n_cates, cfg->n_users};
  e->uniq_cap[0] = M + Bm; e->uniq_cap[1] = M + Bm; e->uniq_cap[2] = Sm;
  for (int i = 0; i < 3; ++i) {
    if (e->uniq_cap[i] > rows3[i]) e->uniq_cap[i] = rows3[i];
    CKC(dalloc(e, &e->slot[i], rows3[i], false));
    CKCU(cudaMemset(e->slot[i], 0xFF, (size_t)rows3[i] * 4));
    CKC(dalloc(e, &e->uniq[i], e->uniq_cap[i]));
  }
  CKC(dalloc(e, &e->cg[0], e->uniq_cap[0] * e->Di));
  CKC(dalloc(e, &e->cg[1], e->uniq_cap[1] * e->Dc));
  CKC(dalloc(e, &e->cg[2], e->uniq_cap[2] * U));
  CKC(dalloc(e, &e->cg[3], e->uniq_cap[2] * U));
  CKC(dalloc(e, &e->counts, 8));
  CKC(dalloc(e, &e->sumsq, 4));
  CKC(dalloc(e, &e->acc, 16));
  CKC(dalloc(e, &e->d_losses, 16));
  CKC(dalloc(e, &e->d_lr, 4));
  {
    // fused _fcn_net kernels: usable when both layouts of a network fit the opt-in shared memory
    CKC(dalloc(e, &e->coop_bar, 2));
    e->coop_rpc = cdiv(e->Bmax, e->num_sms);
    cudaFuncAttributes ff, fb;
    CKCU(cudaFuncGetAttributes(&ff, mlp_fwd_coop_kernel));
    CKCU(cudaFuncGetAttributes(&fb, mlp_bwd_coop_kernel));
    int coop_ok = 0;
    CKCU(cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, e->cfg.device));
    const bool off = getenv("CLSR_NO_COOP_HEAD") != nullptr;
    size_t need_f = 0, need_b = 0;
    Mlp* ms[2] = {&e->mlp_alpha, &e->mlp_logit};
    bool* flag[2] = {&e->coop_alpha, &e->coop_logit};
    for (int i = 0; i < 2; ++i) {
      const Mlp& m = *ms[i];
      const bool shape = m.n0 % 4 == 0 && m.n1 % 4 == 0 && m.n0 <= kCoopMaxN && m.n1 <= kCoopMaxN && m.n0 >= 4 && m.n1 >= 4 &&
                         (m.in + 3) / 4 <= kCoopThreads && 2 * m.n0 <= kPeerSlots;
      if (!shape) { *flag[i] = false; continue; }   // (also the absent alpha MLP of manual_alpha: all sizes zero)
      const size_t sf = (size_t)coop_fwd_smem(m.in, m.n0, m.n1, e->coop_rpc).total * 4;
      const size_t sb = (size_t)coop_bwd_smem(m.in, m.n0, m.n1, e->coop_rpc).total * 4;
      const bool fits = sf + ff.sharedSizeBytes + 512 <= (size_t)e->smem_optin && sb + fb.sharedSizeBytes + 512 <= (size_t)e->smem_optin;
      *flag[i] = !off && coop_ok && shape && fits;
      if (*flag[i]) { need_f = sf > need_f ? sf : need_f; need_b = sb > need_b ? sb : need_b; }
    }
    if (need_f) CKCU(cudaFuncSetAttribute(mlp_fwd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need_f));
    if (need_b) CKCU(cudaFuncSetAttribute(mlp_bwd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need_b));
  }
  CKC(dalloc(e, &e->d_clip_steps, 1));
  CKCU(cudaMallocHost((void**)&e->h_losses, 16 * sizeof(float)));
  CKCU(cudaMallocHost((void**)&e->h_out, (size_t)2 * Bm * sizeof(float)));
  CKC(dalloc(e, &e->d_err, 1));
  CKCU(cudaMallocHost((void**)&e->h_err, sizeof(int32_t)));
  *e->h_err = 0;

  // staged inputs: one contiguous block [ih | ch | mask | tfa | ttn | users | items | cates | labels]
  {
    size_t bytes = (size_t)5 * M * 4 + (size_t)Sm * 4 + (size_t)Bm * 4 * 3;
    char* blk = nullptr;
    CKC(dalloc(e, &blk, (long long)bytes));
    e->in_ih = (int32_t*)blk;
    e->h_stage_bytes = bytes;
    CKCU(cudaMallocHost((void**)&e->h_stage, bytes));
    CKC(dalloc(e, &e->in_raw, (long long)((size_t)5 * M * 4 + (size_t)Bm * 4 * 4), false));
    CKCU(cudaEventCreateWithFlags(&e->h2d_done, cudaEventDisableTiming));
    CKC(dalloc(e, &e->d_len, Sm));
  }

  // activations
  const int A0 = e->A0, A1 = e->A1, Q = e->Q, NX = e->NX, CA = e->CA, L0 = e->L0, L1 = e->L1;
  struct { const char* n; long long sz; } specs[] = {
      {"X", M * D}, {"TNL", M * 2 * H}, {"PX", M * NX}, {"dPX", M * NX},
      {"g1", M * 2 * U}, {"c1", M * U}, {"hp1", M * U}, {"rh1", M * U},
      {"g2", M * 2 * H}, {"c2", M * H}, {"hp2", M * H}, {"rh2", M * H},
      {"G4", M * 4 * H}, {"cp", M * H}, {"mp", M * H}, {"R", M * H},
      {"al", M * U}, {"h0l", M * A0}, {"h1l", M * A1}, {"wl", M}, {"dy1l", M * A1}, {"dy0l", M * A0},
      {"dFl", M * 2 * U}, {"dal", M * U},
      {"as", M * Q}, {"invs", M * A0}, {"dinvs", M * A0}, {"dFs", M * (Q + U)}, {"da2", M * D}, {"das", M * Q},
      {"dR", M * H}, {"dTNL", M * 2 * H}, {"dX", M * D},
      {"h0s", MB * A0}, {"h1s", MB * A1}, {"dy1s", MB * A1}, {"dy0s", MB * A0}, {"dP", MB * D}, {"ws", MB},
      {"tgt", Bm * D}, {"ul", Sm * U}, {"us", Sm * U}, {"sti", Sm * U}, {"fs", Sm * H}, {"afl", Sm * D},
      {"hm", Sm * D}, {"hr", Sm * D}, {"afs", Bm * H}, {"qbl", Sm * A0}, {"qbs", Bm * A0}, {"ca", Bm * CA},
      {"ha0", Bm * A0}, {"ha1", Bm * A1}, {"alogit", Bm}, {"alpha", Bm}, {"mo", Bm * (H + D)}, {"hl0", Bm * L0},
      {"hl1", Bm * L1}, {"logit", Bm}, {"pred", Bm}, {"dlogit", Bm}, {"dhl1", Bm * L1}, {"dhl0", Bm * L0},
      {"dmo", Bm * (H + D)}, {"dalogit", Bm}, {"dha1", Bm * A1}, {"dha0", Bm * A0}, {"dca", Bm * CA},
      {"dfs", Sm * H}, {"dtgt", Bm * D}, {"dafl", Sm * D}, {"dafs", Bm * H}, {"dhm", Sm * D}, {"dhr", Sm * D},
      {"dsti", Sm * U}, {"dus", Sm * U}, {"dul", Sm * U}, {"dqbs", Bm * A0}, {"dqs", Bm * Q}, {"dqbl", Sm * A0},
      {"bpr_gs", Bm * 4},
  };
  for (auto& s : specs) CKC(fbuf(e, s.n, s.sz));
  if (cfg->manual_alpha) {
    // alpha is the constant manual_alpha_value (clsr.py:272-274): the fusion kernels read sigmoid("alogit"), so the buffer
    // is filled once with its logit (+-inf at the ends: sigmoid gives exactly 0 / 1)
    const double a = (double)cfg->manual_alpha_value;
    const float lg = a <= 0.0 ? -INFINITY : (a >= 1.0 ? INFINITY : (float)log(a / (1.0 - a)));
    std::vector<float> h((size_t)Bm, lg);
    CKCU(cudaMemcpy(e->B("alogit"), h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  CKCU(cudaDeviceSynchronize());
  *out = e;
  return CLSR_OK;
}

void clsr_destroy(clsr_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  drop_graphs(e);
  if (e->comm && e->ncclCommDestroy_) e->ncclCommDestroy_(e->comm);
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  for (void* p : e->peer_opened) cudaIpcCloseMemHandle(p);
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_losses) cudaFreeHost(e->h_losses);
  if (e->h_out) cudaFreeHost(e->h_out);
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->h2d_done) cudaEventDestroy(e->h2d_done);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  for (int i = 0; i < 2; ++i) {
    if (e->aux[i]) { cudaStreamSynchronize(e->aux[i]); cudaStreamDestroy(e->aux[i]); }
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_memset) cudaEventDestroy(e->ev_memset);
  if (e->gstream) cudaStreamDestroy(e->gstream);
  if (e->ev_g0) cudaEventDestroy(e->ev_g0);
  if (e->ev_g1) cudaEventDestroy(e->ev_g1);
  delete e;
}

int clsr_set_stream(clsr_engine* e, void* s) {
  if (!e) return CLSR_ERR_ARG;
  if (e->stream) cudaStreamSynchronize(e->stream);
  drop_graphs(e);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)s;
  e->own_stream = false;
  return CLSR_OK;
}

int64_t clsr_workspace_bytes(const clsr_engine* e) { return e ? e->ws_bytes : 0; }
int32_t clsr_dense_count(const clsr_engine* e) { return e ? (int32_t)e->dense.size() : 0; }
const char* clsr_dense_name(const clsr_engine* e, int32_t i) {
  return (e && i >= 0 && i < (int)e->dense.size()) ? e->dense[i].name.c_str() : nullptr;
}
int64_t clsr_dense_offset(const clsr_engine* e, int32_t i) {
  return (e && i >= 0 && i < (int)e->dense.size()) ? e->dense[i].off : -1;
}
int64_t clsr_dense_numel(const clsr_engine* e, int32_t i) {
  return (e && i >= 0 && i < (int)e->dense.size()) ? e->dense[i].n : -1;
}
int32_t clsr_dense_trainable(const clsr_engine* e, int32_t i) {
  return (e && i >= 0 && i < (int)e->dense.size()) ? e->dense[i].trainable : 0;
}
int64_t clsr_dense_total(const clsr_engine* e) { return e ? e->Ptot : 0; }

static float* dense_which(clsr_engine* e, int which) {
  switch (which) {
    case 0: return e->P;
    case 1: return e->Pm;
    case 2: return e->Pv;
    case 3: return e->Pg;
  }
  return nullptr;
}
int clsr_dense_read(clsr_engine* e, int32_t which, float* host_dst) {
  if (!e || !host_dst || !dense_which(e, which)) return fail(e, CLSR_ERR_ARG, "bad argument");
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(host_dst, dense_which(e, which), (size_t)e->Ptot * 4, cudaMemcpyDeviceToHost));
  return CLSR_OK;
}
int clsr_dense_write(clsr_engine* e, int32_t which, const float* host_src) {
  if (!e || !host_src || !dense_which(e, which)) return fail(e, CLSR_ERR_ARG, "bad argument");
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(dense_which(e, which), host_src, (size_t)e->Ptot * 4, cudaMemcpyHostToDevice));
  return CLSR_OK;
}

int clsr_bind_table(clsr_engine* e, int32_t t, float* values, float* m, float* v) {
  if (!e || t < 0 || t >= CLSR_NUM_TABLES || !values) return fail(e, CLSR_ERR_ARG, "bad table binding");
  if (((uintptr_t)values | (uintptr_t)m | (uintptr_t)v) & 15) return fail(e, CLSR_ERR_ARG, "table pointers must be 16-byte aligned");
  if (e->sharded) return fail(e, CLSR_ERR_STATE, "tables are row-sharded and engine-owned (clsr_table_local)");
  if (e->tab[t] != values || e->tab_m[t] != m || e->tab_v[t] != v) drop_graphs(e);   // captured steps hold the old pointers
  e->tab[t] = values; e->tab_m[t] = m; e->tab_v[t] = v;
  memset(&e->tview[t], 0, sizeof(TabView));
  e->tview[t].p[0] = values;
  return CLSR_OK;
}
int clsr_set_graphs(clsr_engine* e, int32_t on) {
  if (!e) return CLSR_ERR_ARG;
  if (e->stream) CK(cudaStreamSynchronize(e->stream));
  if (!on) drop_graphs(e);
  e->graphs_on = on != 0;
  return CLSR_OK;
}
int64_t clsr_graph_replays(const clsr_engine* e) { return e ? e->graph_replays : 0; }
int clsr_set_adam_step(clsr_engine* e, int64_t s) { if (!e) return CLSR_ERR_ARG; e->adam_step = s; return CLSR_OK; }
int64_t clsr_get_adam_step(const clsr_engine* e) { return e ? e->adam_step : -1; }
int clsr_set_debug_sync(clsr_engine* e, int32_t on) { if (!e) return CLSR_ERR_ARG; e->debug_sync = on != 0; return CLSR_OK; }
int64_t clsr_kernel_launches(const clsr_engine* e) { return e ? e->launches : 0; }

// Clip diagnostics of the last synchronised training step: the four table-gradient norms the clip used, and
// the number of shared-history (group > 1) steps so far in which one of them exceeded max_grad_norm.
int clsr_clip_report(clsr_engine* e, float* norms4, int64_t* grouped_active_steps) {
  if (!e) return CLSR_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  if (norms4) for (int t = 0; t < 4; ++t) norms4[t] = e->h_losses[5 + t];
  if (grouped_active_steps) {
    unsigned long long n = 0;
    CK(cudaMemcpy(&n, e->d_clip_steps, sizeof n, cudaMemcpyDeviceToHost));
    *grouped_active_steps = (int64_t)n;
  }
  return CLSR_OK;
}

int clsr_synchronize(clsr_engine* e) {
  if (!e) return CLSR_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  return check_feed_flags(e);
}

static int train_after_stage(clsr_engine* e, const StepCtx& c, uint32_t flags, clsr_losses* out);

int clsr_train_step(clsr_engine* e, const clsr_batch* batch, uint32_t flags, clsr_losses* out) {
  if (!e) return CLSR_ERR_ARG;
  int rc;
  CK(cudaSetDevice(e->cfg.device));
  if ((rc = check_batch(e, batch, true))) return rc;
  e->launches = 0;
  StepCtx c;
  MARK("(step begin)");
  if ((rc = stage_inputs(e, batch, &c, true, out != nullptr))) return rc;
  MARK("h2d_inputs");
  return train_after_stage(e, c, flags, out);
}

// The values that change every step and must therefore stay out of a captured graph: Adam's step count.
static int step_scalars(clsr_engine* e, uint32_t flags) {
  if (flags & CLSR_STEP_NO_OPTIMIZER) return 0;
  const clsr_config& cf = e->cfg;
  e->adam_step += 1;
  const double t = (double)e->adam_step;
  const float lr_t = (float)(cf.learning_rate * sqrt(1.0 - pow((double)cf.beta2, t)) / (1.0 - pow((double)cf.beta1, t)));
  set_scalar_kernel<<<1, 1, 0, e->stream>>>(e->d_lr, lr_t);
  POST("step_scalars");
  return 0;
}

// The step proper: only engine-owned buffers, shape-constant launch parameters -> capturable.
static int step_body(clsr_engine* e, const StepCtx& c, uint32_t flags) {
  int rc;
  if ((rc = zero_step_state(e))) return rc;
  // sharded tables: peers must have finished the previous step's optimizer (and its touched / gradient clean-up)
  // before this step's gathers read their rows or its pushes mark them
  if (e->sharded && (rc = peer_reduce(e, nullptr, 0, nullptr, 0, nullptr, 0))) return rc;
  if ((rc = forward(e, c, 1, (flags & CLSR_STEP_NO_BN_UPDATE) ? 0 : 1))) return rc;
  if ((rc = backward(e, c))) return rc;
  if (e->world > 1) {
    // dense gradients + folded-weight gradients live in one block: one NCCL all-reduce
    if ((rc = allreduce(e, e->Pg, (size_t)(e->Ptot + e->Wtot), kNcclFloat32))) return rc;
    if (!e->sharded && (rc = allreduce(e, e->acc, 5, kNcclFloat64))) return rc;  // data + contrastive partial sums
  }
  if ((rc = sparse_grads(e, c))) return rc;
  blockop_kernel<<<e->n_unprep, 256, 0, e->stream>>>(e->ops_unprep, e->Pg, e->dWd);
  POST("unprep_grads");
  // replicas hold identical dense variables: in sharded mode only rank 0 contributes their regular-loss term
  dense_l2_norm_kernel<<<(int)e->dense.size(), 1024, 0, e->stream>>>(e->d_vars, e->P, e->Pg, e->cfg.layer_l2, e->d_norms, e->acc,
                                                                    (e->sharded && e->rank != 0) ? 0.0 : 1.0);
  POST("dense_l2_norm");
  // sharded: loss partial sums (data, contrastive, owner-side row norms) and the clip norms in one exchange
  if (e->sharded && (rc = peer_reduce(e, e->acc, 11, e->sumsq, 4, nullptr, 0))) return rc;
  loss_finalize_kernel<<<1, 32, 0, e->stream>>>(e->acc, e->counts, e->counts + (e->sharded ? 4 : 3), c.G, e->U, e->cfg.embed_l2,
                                                e->cfg.contrastive_weight, e->cfg.discrepancy_weight, e->sumsq,
                                                e->cfg.clip_norm ? e->cfg.max_grad_norm : 0.f, c.G, e->d_clip_steps,
                                                e->d_losses);
  POST("loss_finalize");
  if (!(flags & CLSR_STEP_NO_OPTIMIZER)) {
    if ((rc = optimizer_step(e))) return rc;
  }
  if (e->sharded && (rc = shard_cleanup(e, (flags & CLSR_STEP_NO_OPTIMIZER) != 0))) return rc;
  if ((rc = reset_slots(e))) return rc;
  CK(cudaMemcpyAsync(e->h_losses, e->d_losses, 9 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  return 0;
}

// Run the step body eagerly, or (single GPU, no per-kernel events) as one graph launch.
static int run_step(clsr_engine* e, const StepCtx& c, uint32_t flags) {
  static const bool env_off = getenv("CLSR_NO_GRAPH") != nullptr;
  // (data parallel: only with row-sharded tables -- its exchanges are peer-memory kernels with device-side sequence
  //  numbers plus one NCCL all-reduce, all capturable; every rank replays or launches the same kernel sequence)
  const bool can = e->graphs_on && !env_off && (e->world == 1 || e->sharded) && !e->profiling && !e->debug_sync;
  if (!can) return step_body(e, c, flags);
  clsr_engine::StepGraph* g = nullptr;
  for (auto& x : e->graphs)
    if (x.S == c.S && x.B == c.B && x.G == c.G && x.flags == flags) { g = &x; break; }
  if (!g) {
    if (e->graphs.size() >= 16) drop_graphs(e);   // a stream of odd shapes: start over rather than grow
    e->graphs.push_back({c.S, c.B, c.G, flags, 0, 0, nullptr});
    g = &e->graphs.back();
  }
  cudaStream_t user = e->stream, cap = e->stream;
  if (user == nullptr || user == cudaStreamLegacy) {
    if (!e->gstream) {
      bool okc = cudaStreamCreate(&e->gstream) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_g0, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_g1, cudaEventDisableTiming) == cudaSuccess;
      if (!okc) { cudaGetLastError(); e->graphs_on = false; return step_body(e, c, flags); }
    }
    cap = e->gstream;
  } else if (user == cudaStreamPerThread) {
    return step_body(e, c, flags);
  }
  auto launch = [&](cudaGraphExec_t x) -> int {
    if (cap != user) {   // default-stream engine: staged feed / scalars -> graph -> whatever the caller enqueues next
      CK(cudaEventRecord(e->ev_g0, user));
      CK(cudaStreamWaitEvent(cap, e->ev_g0, 0));
    }
    CK(cudaGraphLaunch(x, cap));
    if (cap != user) {
      CK(cudaEventRecord(e->ev_g1, cap));
      CK(cudaStreamWaitEvent(user, e->ev_g1, 0));
    }
    e->graph_replays++;
    return 0;
  };
  if (g->exec) {
    e->launches += g->launches;
    return launch(g->exec);
  }
  if (g->seen++ == 0) return step_body(e, c, flags);   // first step of this shape: eager
  // second step of this shape: capture it (relaxed mode: the body may call non-stream APIs such as tensor-map
  // encoding), then launch the instantiated graph.  Any failure turns graphs off for this engine and runs eagerly.
  const long long l0 = e->launches;
  cudaGraph_t graph = nullptr;
  bool ok = cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) == cudaSuccess;
  int rc = 0;
  if (ok) {
    e->stream = cap;
    rc = step_body(e, c, flags);
    e->stream = user;
    ok = cudaStreamEndCapture(cap, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
  }
  if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok) {
    cudaGetLastError();
    g->exec = nullptr;
    e->graphs_on = false;
    e->launches = l0;
    return step_body(e, c, flags);
  }
  g->launches = e->launches - l0;
  return launch(g->exec);
}

// Everything after the feed is staged (by stage_inputs or by clsr_build_batch).
static int train_after_stage(clsr_engine* e, const StepCtx& c, uint32_t flags, clsr_losses* out) {
  int rc;
  if ((rc = step_scalars(e, flags))) return rc;
  if ((rc = run_step(e, c, flags))) return rc;
  if (out) {
    CK(cudaStreamSynchronize(e->stream));
    if ((rc = check_feed_flags(e))) return rc;
    out->loss = e->h_losses[0]; out->data_loss = e->h_losses[1]; out->regular_loss = e->h_losses[2];
    out->contrastive_loss = e->h_losses[3]; out->discrepancy_loss = e->h_losses[4];
    for (int t = 0; t < 4; ++t) out->table_grad_norm[t] = e->h_losses[5 + t];
  }
  return CLSR_OK;
}

// Inference forward on the staged feed; leaves sigmoid(logit) in the "pred" buffer and alpha in "alpha".
static int predict_forward(clsr_engine* e, const clsr_batch* batch, StepCtx* c, bool sync_call) {
  int rc;
  CK(cudaSetDevice(e->cfg.device));
  if ((rc = check_batch(e, batch, false))) return rc;
  e->launches = 0;
  if ((rc = stage_inputs(e, batch, c, false, sync_call))) return rc;
  CK(cudaMemsetAsync(e->counts, 0, 8 * sizeof(int32_t), e->stream));
  // sharded tables: a collective call -- peers must have finished updating the rows this rank is about to read
  if (e->sharded && (rc = peer_reduce(e, nullptr, 0, nullptr, 0, nullptr, 0))) return rc;
  if ((rc = forward(e, *c, 0, 0))) return rc;
  sigmoid_kernel<<<grid1d(e, c->B, 256), 256, 0, e->stream>>>(e->B("logit"), c->B, e->B("pred"));
  POST("sigmoid");
  return 0;
}

int clsr_predict(clsr_engine* e, const clsr_batch* batch, float* pred, float* alpha) {
  if (!e) return CLSR_ERR_ARG;
  if (!pred) return fail(e, CLSR_ERR_ARG, "null pred buffer");
  int rc;
  StepCtx c;
  if ((rc = predict_forward(e, batch, &c, true))) return rc;
  const int B = c.B;
  CK(cudaMemcpyAsync(e->h_out, e->B("pred"), (size_t)B * 4, cudaMemcpyDeviceToHost, e->stream));
  if (alpha) CK(cudaMemcpyAsync(e->h_out + e->Bmax, e->B("alpha"), (size_t)B * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if ((rc = check_feed_flags(e))) return rc;
  memcpy(pred, e->h_out, (size_t)B * 4);
  if (alpha) memcpy(alpha, e->h_out + e->Bmax, (size_t)B * 4);
  return CLSR_OK;
}

int clsr_predict_device(clsr_engine* e, const clsr_batch* batch, float* dev_pred, float* dev_alpha) {
  if (!e) return CLSR_ERR_ARG;
  if (!dev_pred) return fail(e, CLSR_ERR_ARG, "null pred buffer");
  int rc;
  StepCtx c;
  if ((rc = predict_forward(e, batch, &c, false))) return rc;
  CK(cudaMemcpyAsync(dev_pred, e->B("pred"), (size_t)c.B * 4, cudaMemcpyDeviceToDevice, e->stream));
  if (dev_alpha) CK(cudaMemcpyAsync(dev_alpha, e->B("alpha"), (size_t)c.B * 4, cudaMemcpyDeviceToDevice, e->stream));
  return CLSR_OK;
}

// ---- batch construction on the GPU --------------------------------------------------------
// Reference: SASequentialIterator._convert_data (sequential_iterator.py:519-704) pads, replicates x(1+num_ngs)
// and samples in-batch negatives row by row in Python between two sess.run calls.  Here a parsed file is uploaded
// once as columnar device arrays and a batch is built by one kernel straight into the staged-feed block the step
// reads; the host only sends the line indices of the batch.
struct clsr_dataset {
  int device = 0, T = 0;
  long long n = 0;
  float *label = nullptr, *tfa = nullptr, *ttn = nullptr;
  int32_t *user = nullptr, *item = nullptr, *cate = nullptr, *length = nullptr, *ih = nullptr, *ch = nullptr;
  int32_t* lines = nullptr;   // device copy of the current batch's line indices
  long long lines_cap = 0;
};

namespace {

// staged-block pointers for S sequences / B rows (same layout as stage_inputs)
void staged_ctx(clsr_engine* e, int S, int G, StepCtx* c) {
  const int T = e->T, B = S * G;
  const size_t seq_i = (size_t)S * T * 4;
  c->B = B; c->G = G; c->S = S; c->T = T;
  char* d = (char*)e->in_ih;
  c->ih = (const int32_t*)d; d += seq_i;
  c->ch = (const int32_t*)d; d += seq_i;
  c->mask = (const int32_t*)d; d += seq_i;
  c->tfa = (const float*)d; d += seq_i;
  c->ttn = (const float*)d; d += seq_i;
  c->users = (const int32_t*)d; d += (size_t)S * 4;
  c->items = (const int32_t*)d; d += (size_t)B * 4;
  c->cates = (const int32_t*)d; d += (size_t)B * 4;
  c->labels = (const float*)d;
  c->seq_stride = T;
  c->user_stride = 1;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void build_batch_kernel(const float* __restrict__ dl, const int32_t* __restrict__ du, const int32_t* __restrict__ di,
                                   const int32_t* __restrict__ dc, const int32_t* __restrict__ dlen,
                                   const int32_t* __restrict__ dih, const int32_t* __restrict__ dch,
                                   const float* __restrict__ dtfa, const float* __restrict__ dttn,
                                   const int32_t* __restrict__ lines, int S, int G, int T, unsigned long long seed,
                                   int32_t* __restrict__ o_ih, int32_t* __restrict__ o_ch, int32_t* __restrict__ o_mask,
                                   float* __restrict__ o_tfa, float* __restrict__ o_ttn, int32_t* __restrict__ o_users,
                                   int32_t* __restrict__ o_items, int32_t* __restrict__ o_cates, float* __restrict__ o_labels) {
  const long long M = (long long)S * T, B = (long long)S * G;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const int s = (int)(i / T), t = (int)(i - (long long)s * T);
    const size_t src = (size_t)lines[s] * T + t;
    o_ih[i] = dih[src]; o_ch[i] = dch[src]; o_tfa[i] = dtfa[src]; o_ttn[i] = dttn[src];
    o_mask[i] = t < dlen[lines[s]] ? 1 : 0;
  }
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
    const int s = (int)(b / G), g = (int)(b - (long long)s * G);
    const int line = lines[s];
    if (g == 0) {
      o_users[s] = du[line];
      o_items[b] = di[line]; o_cates[b] = dc[line];
      o_labels[b] = G > 1 ? 1.0f : dl[line];
    } else {
      // a negative = the positive item of another line of this batch, never this line's own item
      // (sequential_iterator.py:612-634: random.randint until it differs); counter-based draws
      const int own = di[line];
      int pick = -1;
      for (int k = 0; k < 64 && pick < 0; ++k) {
        const int j = (int)(mix64(seed ^ ((unsigned long long)b << 8) ^ (unsigned long long)k) % (unsigned long long)S);
        if (di[lines[j]] != own) pick = j;
      }
      for (int j = 0; j < S && pick < 0; ++j)   // 64 misses: the batch is dominated by one item; scan for any other
        if (di[lines[(s + 1 + j) % S]] != own) pick = (s + 1 + j) % S;
      if (pick < 0) pick = s;                   // every line has the same item (the reference would not terminate)
      o_items[b] = di[lines[pick]]; o_cates[b] = dc[lines[pick]];
      o_labels[b] = 0.0f;
    }
  }
}

}  // namespace

int clsr_dataset_create(clsr_engine* e, int64_t n_lines, const float* label, const int32_t* user, const int32_t* item,
                        const int32_t* cate, const int32_t* length, const int32_t* item_hist, const int32_t* cate_hist,
                        const float* tfa, const float* ttn, clsr_dataset** out) {
  if (!e || !out || n_lines <= 0 || !label || !user || !item || !cate || !length || !item_hist || !cate_hist || !tfa || !ttn)
    return fail(e, CLSR_ERR_ARG, "bad argument");
  *out = nullptr;
  const int T = e->T;
  // ids are validated once, here: batches built from the cache need no per-step check
  auto bad = [](const int32_t* p, size_t n, long long limit) {
    uint32_t mx = 0;
    for (size_t i = 0; i < n; ++i) { const uint32_t v = (uint32_t)p[i]; mx = v > mx ? v : mx; }
    return (long long)mx >= limit;
  };
  if (bad(user, n_lines, e->cfg.n_users) || bad(item, n_lines, e->cfg.n_items) || bad(cate, n_lines, e->cfg.n_cates) ||
      bad(item_hist, (size_t)n_lines * T, e->cfg.n_items) || bad(cate_hist, (size_t)n_lines * T, e->cfg.n_cates))
    return fail(e, CLSR_ERR_ARG, "dataset holds an id outside its table");
  for (int64_t i = 0; i < n_lines; ++i)
    if (length[i] < 0 || length[i] > T) return fail(e, CLSR_ERR_ARG, "history length outside [0, %d]", T);
  CK(cudaSetDevice(e->cfg.device));
  clsr_dataset* d = new clsr_dataset();
  d->device = e->cfg.device; d->T = T; d->n = n_lines;
  const size_t n = (size_t)n_lines, nt = n * T;
  struct { void** p; const void* src; size_t bytes; } cols[] = {
      {(void**)&d->label, label, n * 4}, {(void**)&d->user, user, n * 4}, {(void**)&d->item, item, n * 4},
      {(void**)&d->cate, cate, n * 4}, {(void**)&d->length, length, n * 4}, {(void**)&d->ih, item_hist, nt * 4},
      {(void**)&d->ch, cate_hist, nt * 4}, {(void**)&d->tfa, tfa, nt * 4}, {(void**)&d->ttn, ttn, nt * 4}};
  for (auto& c : cols) {
    cudaError_t err = cudaMalloc(c.p, c.bytes);
    if (err == cudaSuccess) err = cudaMemcpy(*c.p, c.src, c.bytes, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
      clsr_dataset_destroy(d);
      return fail(e, CLSR_ERR_CUDA, "dataset upload failed: %s", cudaGetErrorString(err));
    }
  }
  *out = d;
  return CLSR_OK;
}

void clsr_dataset_destroy(clsr_dataset* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  void* ps[] = {d->label, d->user, d->item, d->cate, d->length, d->ih, d->ch, d->tfa, d->ttn, d->lines};
  for (void* p : ps)
    if (p) cudaFree(p);
  delete d;
}

int clsr_build_batch(clsr_engine* e, clsr_dataset* d, const int32_t* host_lines, int32_t count, int32_t num_ngs, uint64_t seed) {
  if (!e || !d || !host_lines || count <= 0 || num_ngs < 0) return fail(e, CLSR_ERR_ARG, "bad argument");
  const int G = num_ngs + 1, S = count;
  if ((long long)S * G > e->Bmax) return fail(e, CLSR_ERR_ARG, "rows %lld exceed max_rows %d", (long long)S * G, e->Bmax);
  if (S > e->Smax) return fail(e, CLSR_ERR_ARG, "sequences %d exceed max_seqs %d", S, e->Smax);
  if (d->T != e->T || d->device != e->cfg.device) return fail(e, CLSR_ERR_ARG, "dataset belongs to another engine shape / device");
  for (int i = 0; i < count; ++i)
    if (host_lines[i] < 0 || host_lines[i] >= d->n) return fail(e, CLSR_ERR_ARG, "line index %d outside the dataset", host_lines[i]);
  CK(cudaSetDevice(e->cfg.device));
  if (d->lines_cap < count) {
    if (d->lines) CK(cudaFree(d->lines));
    d->lines = nullptr;
    CK(cudaMalloc((void**)&d->lines, (size_t)e->Bmax * 4));
    d->lines_cap = e->Bmax;
  }
  CK(cudaEventSynchronize(e->h2d_done));
  memcpy(e->h_stage, host_lines, (size_t)count * 4);
  CK(cudaMemcpyAsync(d->lines, e->h_stage, (size_t)count * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaEventRecord(e->h2d_done, e->stream));
  StepCtx c;
  staged_ctx(e, S, G, &c);
  build_batch_kernel<<<grid1d(e, (long long)S * e->T, 256, 4), 256, 0, e->stream>>>(
      d->label, d->user, d->item, d->cate, d->length, d->ih, d->ch, d->tfa, d->ttn, d->lines, S, G, e->T, seed,
      const_cast<int32_t*>(c.ih), const_cast<int32_t*>(c.ch), const_cast<int32_t*>(c.mask), const_cast<float*>(c.tfa),
      const_cast<float*>(c.ttn), const_cast<int32_t*>(c.users), const_cast<int32_t*>(c.items), const_cast<int32_t*>(c.cates),
      const_cast<float*>(c.labels));
  POST("build_batch");
  e->staged_S = S; e->staged_G = G;
  return CLSR_OK;
}

int clsr_train_step_staged(clsr_engine* e, uint32_t flags, clsr_losses* out) {
  if (!e) return CLSR_ERR_ARG;
  if (e->staged_S <= 0) return fail(e, CLSR_ERR_STATE, "no staged batch (clsr_build_batch)");
  if ((e->staged_S * e->staged_G) % e->cfg.train_group) return fail(e, CLSR_ERR_ARG, "rows not a multiple of train_group");
  for (int t = 0; t < CLSR_NUM_TABLES; ++t)
    if (!e->tab[t]) return fail(e, CLSR_ERR_STATE, "table %d not bound", t);
  CK(cudaSetDevice(e->cfg.device));
  StepCtx c;
  staged_ctx(e, e->staged_S, e->staged_G, &c);
  return train_after_stage(e, c, flags, out);
}

// Inference on the staged batch; pred / alpha / users / labels: device buffers of `rows` elements (any may be NULL).
int clsr_predict_staged(clsr_engine* e, float* dev_pred, float* dev_alpha, int32_t* dev_users, float* dev_labels) {
  if (!e) return CLSR_ERR_ARG;
  if (e->staged_S <= 0) return fail(e, CLSR_ERR_STATE, "no staged batch (clsr_build_batch)");
  CK(cudaSetDevice(e->cfg.device));
  int rc;
  StepCtx c;
  staged_ctx(e, e->staged_S, e->staged_G, &c);
  CK(cudaMemsetAsync(e->counts, 0, 8 * sizeof(int32_t), e->stream));
  if (e->sharded && (rc = peer_reduce(e, nullptr, 0, nullptr, 0, nullptr, 0))) return rc;
  if ((rc = forward(e, c, 0, 0))) return rc;
  sigmoid_kernel<<<grid1d(e, c.B, 256), 256, 0, e->stream>>>(e->B("logit"), c.B, e->B("pred"));
  POST("sigmoid");
  if (dev_pred) CK(cudaMemcpyAsync(dev_pred, e->B("pred"), (size_t)c.B * 4, cudaMemcpyDeviceToDevice, e->stream));
  if (dev_alpha) CK(cudaMemcpyAsync(dev_alpha, e->B("alpha"), (size_t)c.B * 4, cudaMemcpyDeviceToDevice, e->stream));
  if (dev_labels) CK(cudaMemcpyAsync(dev_labels, c.labels, (size_t)c.B * 4, cudaMemcpyDeviceToDevice, e->stream));
  if (dev_users) {
    if (c.G != 1) return fail(e, CLSR_ERR_ARG, "users are per sequence: only for ungrouped (evaluation) batches");
    CK(cudaMemcpyAsync(dev_users, c.users, (size_t)c.S * 4, cudaMemcpyDeviceToDevice, e->stream));
  }
  return CLSR_OK;
}

// Copy one array of the staged feed to the host (tests): which = 0 item_history, 1 cate_history, 2 mask,
// 3 time_from_first_action, 4 time_to_now, 5 users, 6 items, 7 cates, 8 labels.
int clsr_staged_feed_read(clsr_engine* e, int32_t which, void* host_dst, int64_t bytes) {
  if (!e || !host_dst || which < 0 || which > 8) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (e->staged_S <= 0) return fail(e, CLSR_ERR_STATE, "no staged batch");
  StepCtx c;
  staged_ctx(e, e->staged_S, e->staged_G, &c);
  const void* src[9] = {c.ih, c.ch, c.mask, c.tfa, c.ttn, c.users, c.items, c.cates, c.labels};
  const size_t seq = (size_t)c.S * c.T * 4, row = (size_t)c.B * 4;
  const size_t have[9] = {seq, seq, seq, seq, seq, (size_t)c.S * 4, row, row, row};
  if ((size_t)bytes > have[which]) return fail(e, CLSR_ERR_ARG, "array %d holds %zu bytes", which, have[which]);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(host_dst, src[which], (size_t)bytes, cudaMemcpyDeviceToHost));
  return CLSR_OK;
}

int clsr_gather_history(clsr_engine* e, const int32_t* ih, const int32_t* ch, int64_t positions, float* out) {
  if (!e || !ih || !ch || !out || positions <= 0) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (!e->tab[0] || !e->tab[1]) return fail(e, CLSR_ERR_STATE, "tables not bound");
  int rc = launch_gather_hist(e, ih, ch, e->T, out, positions);
  if (rc) return rc;
  return CLSR_OK;
}

int clsr_scatter_history_grad(clsr_engine* e, const int32_t* ih, const int32_t* ch, int64_t positions, const float* d_hist) {
  if (!e || !ih || !ch || !d_hist || positions <= 0) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (positions > e->uniq_cap[0] && e->uniq_cap[0] < e->tab_rows[0]) return fail(e, CLSR_ERR_ARG, "too many positions");
  cudaStream_t st = e->stream;
  CK(cudaMemsetAsync(e->counts, 0, 8 * sizeof(int32_t), st));
  CK(cudaMemsetAsync(e->sumsq, 0, 4 * sizeof(double), st));
  mark_unique_kernel<<<grid1d(e, positions, 256), 256, 0, st>>>(ih, positions, e->T, e->T, e->slot[0], e->uniq[0], e->counts + 1);
  POST("unique_item_hist");
  mark_unique_kernel<<<grid1d(e, positions, 256), 256, 0, st>>>(ch, positions, e->T, e->T, e->slot[1], e->uniq[1], e->counts + 2);
  POST("unique_cate_hist");
  zero_compact_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->cg[0], e->counts + 1, e->Di);
  POST("zero_compact");
  zero_compact_kernel<<<e->num_sms * 2, 256, 0, st>>>(e->cg[1], e->counts + 2, e->Dc);
  POST("zero_compact");
  {
    const int hot = scatter_hot_rows(e);
    scatter_hist_kernel<<<e->num_sms * 4, 256, (size_t)(hot * e->Di + e->Dc) * 4, st>>>(
        d_hist, ih, ch, e->T, e->T, e->slot[0], e->slot[1], e->cg[0], e->cg[1], e->Di, e->Dc, positions, e->sumsq, hot);
    POST("scatter_hist");
  }
  for (int i = 0; i < 2; ++i) {
    reset_slots_kernel<<<e->num_sms, 256, 0, st>>>(e->uniq[i], e->counts + 1 + i, e->slot[i]);
    POST("reset_slots");
  }
  return CLSR_OK;
}

int clsr_sparse_grad_view(clsr_engine* e, int32_t t, const int32_t** ids, const float** rows, const int32_t** count) {
  if (!e || t < 0 || t >= CLSR_NUM_TABLES) return fail(e, CLSR_ERR_ARG, "bad table");
  const int slot_ix[4] = {0, 1, 2, 2}, cnt_ix[4] = {1, 2, 3, 3};
  if (ids) *ids = e->uniq[slot_ix[t]];
  if (rows) *rows = e->cg[t];
  if (count) *count = e->counts + cnt_ix[t];
  return CLSR_OK;
}

int clsr_debug_buffer(clsr_engine* e, const char* name, const float** p, int64_t* n) {
  if (!e || !name) return CLSR_ERR_ARG;
  auto it = e->bufs.find(name);
  if (it != e->bufs.end()) {
    if (p) *p = it->second.first;
    if (n) *n = it->second.second;
    return CLSR_OK;
  }
  // "bn/<mlp><layer>/<field>": BatchNorm vectors of the last step, e.g. bn/short0/scale (field: scale, shift, mean, rstd)
  if (!strncmp(name, "bn/", 3)) {
    const char* mlps[4] = {"long", "short", "alpha", "logit"};
    Mlp* ms[4] = {&e->mlp_long, &e->mlp_short, &e->mlp_alpha, &e->mlp_logit};
    for (int i = 0; i < 4; ++i) {
      const size_t ln = strlen(mlps[i]);
      if (strncmp(name + 3, mlps[i], ln) || (name[3 + ln] != '0' && name[3 + ln] != '1') || name[4 + ln] != '/') continue;
      BnLayer& b = name[3 + ln] == '0' ? ms[i]->bn0 : ms[i]->bn1;
      const char* f = name + 5 + ln;
      float* v = !strcmp(f, "scale") ? b.scale : !strcmp(f, "shift") ? b.shift : !strcmp(f, "mean") ? b.mean
                 : !strcmp(f, "rstd") ? b.rstd : nullptr;
      if (!v) break;
      if (p) *p = v;
      if (n) *n = b.N;
      return CLSR_OK;
    }
    return fail(e, CLSR_ERR_ARG, "no buffer named %s", name);
  }
  auto wi = e->wd_off.find(name);
  if (wi != e->wd_off.end()) {
    if (p) *p = e->Wd + wi->second;
    if (n) *n = e->Wtot - wi->second;
    return CLSR_OK;
  }
  return fail(e, CLSR_ERR_ARG, "no buffer named %s", name);
}

int clsr_debug_read(clsr_engine* e, const void* src, void* dst, int64_t bytes) {
  if (!e || !src || !dst || bytes < 0) return fail(e, CLSR_ERR_ARG, "bad argument");
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return CLSR_OK;
}

// Standalone linear layer C = A.W + bias on device pointers (mode 0: fp32 SIMT, 1: tcgen05 split-bf16).
int clsr_debug_gemm(clsr_engine* e, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* W,
                    int32_t ldw, const float* bias, float* C, int32_t ldc, int32_t mode) {
  if (!e || !A || !W || !C) return fail(e, CLSR_ERR_ARG, "bad argument");
  int saved = e->cfg.math_mode;
  e->cfg.math_mode = mode;
  int rc = mode == 1 ? tc_gemm(e, "debug_tc_gemm", M, N, K, a_plain(A, lda), W, ldw, e_store(C, ldc, bias), false)
                     : gemm(e, "debug_gemm", M, N, K, a_plain(A, lda), W, ldw, e_store(C, ldc, bias), false);
  e->cfg.math_mode = saved;
  return rc;
}

// Standalone weight-gradient product dW[K,N] += A[M,K]^T . B[M,N] (+ column sums of B) on device pointers.
int clsr_debug_dwgemm(clsr_engine* e, int32_t M, int32_t K, int32_t N, const float* A, int32_t lda, const float* B,
                      int32_t ldb, float* dW, int32_t lddw, float* colsum, int32_t mode) {
  if (!e || !A || !B || !dW) return fail(e, CLSR_ERR_ARG, "bad argument");
  int saved = e->cfg.math_mode;
  e->cfg.math_mode = mode;
  int rc = dwgemm(e, "debug_dwgemm", M, K, N, a_plain(A, lda), a_plain(B, ldb), dW, lddw, colsum);
  e->cfg.math_mode = saved;
  return rc;
}

int clsr_set_profiling(clsr_engine* e, int32_t on) {
  if (!e) return CLSR_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  e->profiling = on != 0;
  e->prof_used = 0;
  e->prof_agg.clear();
  return CLSR_OK;
}

// Fold the recorded events into per-kernel totals; returns the number of distinct names.
int clsr_profile_collect(clsr_engine* e) {
  if (!e) return CLSR_ERR_ARG;
  CK(cudaStreamSynchronize(e->stream));
  for (size_t i = 1; i < e->prof_used; ++i) {
    if (!strcmp(e->prof_names[i], "(step begin)")) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_events[i - 1], e->prof_events[i]) != cudaSuccess) continue;
    auto& a = e->prof_agg[e->prof_names[i]];
    a.first += ms;
    a.second += 1;
  }
  e->prof_used = 0;
  return (int)e->prof_agg.size();
}

int clsr_profile_entry(clsr_engine* e, int32_t i, char* name, int32_t name_cap, double* total_ms, int64_t* calls) {
  if (!e || i < 0 || i >= (int)e->prof_agg.size() || !name || name_cap <= 0) return CLSR_ERR_ARG;
  auto it = e->prof_agg.begin();
  std::advance(it, i);
  snprintf(name, (size_t)name_cap, "%s", it->first.c_str());
  if (total_ms) *total_ms = it->second.first;
  if (calls) *calls = it->second.second;
  return CLSR_OK;
}

// ---- peer-memory communication and row-sharded tables ------------------------------------------
// Blob exchanged between the ranks (through any host channel): a 64-byte header and 13 CUDA IPC handles
// (communication buffer; per table: values, dense gradient shard, touched marks).
namespace {
constexpr int kBlobBytes = 1024;
constexpr size_t kCommBytes = 2u << 20;   // [data: 2 x world x kPeerSlots doubles | flags at +1 MB]
struct PeerBlob {
  int32_t magic, sharded, rank, world;
  char pad[48];
  cudaIpcMemHandle_t comm;
  cudaIpcMemHandle_t tab[CLSR_NUM_TABLES][3];
};
static_assert(sizeof(PeerBlob) <= kBlobBytes, "blob size");
}  // namespace

int clsr_peer_setup_begin(clsr_engine* e, int32_t shard_tables, void* blob_out) {
  if (!e || !blob_out) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (e->world <= 1) return fail(e, CLSR_ERR_STATE, "clsr_comm_init first (world > 1)");
  if (e->comm_data) return fail(e, CLSR_ERR_STATE, "peer communication already set up");
  if (e->world > kMaxWorld || (e->world & (e->world - 1))) return fail(e, CLSR_ERR_ARG, "world %d: need a power of two <= %d", e->world, kMaxWorld);
  if ((size_t)2 * e->world * kPeerSlots * 8 > (1u << 20)) return fail(e, CLSR_ERR_ARG, "world too large for the communication buffer");
  CK(cudaSetDevice(e->cfg.device));
  char* cb = nullptr;
  int rc = dalloc(e, &cb, (long long)kCommBytes);
  if (rc) return rc;
  e->comm_data = (double*)cb;
  e->comm_flags = (unsigned long long*)(cb + (1u << 20));
  PeerBlob b;
  memset(&b, 0, sizeof b);
  b.magic = 0x434C5352; b.sharded = shard_tables ? 1 : 0; b.rank = e->rank; b.world = e->world;
  CK(cudaIpcGetMemHandle(&b.comm, cb));
  if (shard_tables) {
    // engine-owned shards: global row r lives on rank r % world at local row r / world
    for (int t = 0; t < CLSR_NUM_TABLES; ++t) {
      long long lr = (e->tab_rows[t] - e->rank + e->world - 1) / e->world;
      if (lr < 1) lr = 1;
      e->tab_local_rows[t] = lr;
      const long long n = lr * e->tab_dim[t];
      if ((rc = dalloc(e, &e->sh_val[t], n)) || (rc = dalloc(e, &e->sh_m[t], n)) || (rc = dalloc(e, &e->sh_v[t], n)) ||
          (rc = dalloc(e, &e->sh_g[t], n)))
        return rc;
      if (t < 3 && (rc = dalloc(e, &e->sh_touched[t], lr))) return rc;
      if (t == 3) e->sh_touched[3] = e->sh_touched[2];   // the two user tables are indexed by the same ids
      e->tab[t] = e->sh_val[t]; e->tab_m[t] = e->sh_m[t]; e->tab_v[t] = e->sh_v[t];
      CK(cudaIpcGetMemHandle(&b.tab[t][0], e->sh_val[t]));
      CK(cudaIpcGetMemHandle(&b.tab[t][1], e->sh_g[t]));
      if (t < 3) CK(cudaIpcGetMemHandle(&b.tab[t][2], e->sh_touched[t]));
    }
    e->sharded = true;
  }
  memset(blob_out, 0, kBlobBytes);
  memcpy(blob_out, &b, sizeof b);
  CK(cudaDeviceSynchronize());
  return CLSR_OK;
}

int clsr_peer_setup_finish(clsr_engine* e, const void* all_blobs) {
  if (!e || !all_blobs) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (!e->comm_data) return fail(e, CLSR_ERR_STATE, "clsr_peer_setup_begin first");
  if (e->peer_ready) return CLSR_OK;
  CK(cudaSetDevice(e->cfg.device));
  int shift = 0;
  while ((1 << shift) < e->world) ++shift;
  drop_graphs(e);
  memset(&e->pc, 0, sizeof e->pc);
  e->pc.rank = e->rank; e->pc.world = e->world;
  {
    int rc0 = dalloc(e, &e->pc.seq_ctr, 2);   // zeroed: exchange counter (see PeerComm)
    if (rc0) return rc0;
  }
  for (int t = 0; t < CLSR_NUM_TABLES; ++t) {
    memset(&e->gview[t], 0, sizeof(GradView));
    if (e->sharded) { memset(&e->tview[t], 0, sizeof(TabView)); e->tview[t].shift = shift; e->tview[t].mask = e->world - 1; }
    e->gview[t].shift = shift; e->gview[t].mask = e->world - 1;
  }
  auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> int {
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    e->peer_opened.push_back(*out);
    return 0;
  };
  for (int r = 0; r < e->world; ++r) {
    PeerBlob b;
    memcpy(&b, (const char*)all_blobs + (size_t)r * kBlobBytes, sizeof b);
    if (b.magic != 0x434C5352 || b.rank != r || b.world != e->world || b.sharded != (e->sharded ? 1 : 0))
      return fail(e, CLSR_ERR_ARG, "peer blob %d does not match this group (rank %d world %d sharded %d)", r, b.rank, b.world, b.sharded);
    void* cb = nullptr;
    int rc;
    if (r == e->rank) cb = e->comm_data;
    else if ((rc = open(b.comm, &cb))) return rc;
    e->pc.data[r] = (double*)cb;
    e->pc.flag[r] = (unsigned long long*)((char*)cb + (1u << 20));
    if (!e->sharded) continue;
    for (int t = 0; t < CLSR_NUM_TABLES; ++t) {
      void *pv = e->sh_val[t], *pg = e->sh_g[t], *pt = e->sh_touched[t];
      if (r != e->rank) {
        if ((rc = open(b.tab[t][0], &pv)) || (rc = open(b.tab[t][1], &pg))) return rc;
        if (t < 3) { if ((rc = open(b.tab[t][2], &pt))) return rc; }
        else pt = e->gview[2].touched[r];
      }
      e->tview[t].p[r] = (const float*)pv;
      e->gview[t].g[r] = (float*)pg;
      e->gview[t].touched[r] = (int32_t*)pt;
    }
  }
  e->peer_ready = true;
  return CLSR_OK;
}

// Local shard of a row-sharded table: which = 0 values, 1 Adam m, 2 Adam v (device pointers, [rows, dim] fp32).
int clsr_table_local(clsr_engine* e, int32_t t, int32_t which, float** ptr, int64_t* rows) {
  if (!e || t < 0 || t >= CLSR_NUM_TABLES || which < 0 || which > 2) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (!e->sharded) return fail(e, CLSR_ERR_STATE, "tables are not sharded");
  if (ptr) *ptr = which == 0 ? e->sh_val[t] : which == 1 ? e->sh_m[t] : e->sh_v[t];
  if (rows) *rows = e->tab_local_rows[t];
  return CLSR_OK;
}

static void* open_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return h;
}

int clsr_nccl_unique_id(void* out128) {
  void* h = open_nccl();
  if (!h || !out128) return CLSR_ERR_NCCL;
  typedef int (*fn_t)(void*);
  fn_t f = (fn_t)dlsym(h, "ncclGetUniqueId");
  if (!f) return CLSR_ERR_NCCL;
  return f(out128) == 0 ? CLSR_OK : CLSR_ERR_NCCL;
}

// Data-parallel training over `world` ranks (one process per GPU): replicated variables, the batch
// sharded by user-sequences.  Collectives (all on the engine stream): BatchNorm statistics in both
// directions, the contrastive row count, dense gradients + loss partial sums, and an all-gather of the
// sparse-gradient inputs.  Every rank must feed the same number of rows per step.
int clsr_comm_init(clsr_engine* e, int32_t rank, int32_t world, const void* id128) {
  if (e) drop_graphs(e);
  if (!e || !id128 || world < 1 || rank < 0 || rank >= world) return fail(e, CLSR_ERR_ARG, "bad argument");
  if (world == 1) return CLSR_OK;
  if (e->comm) return fail(e, CLSR_ERR_STATE, "communicator already initialised");
  e->nccl_lib = open_nccl();
  if (!e->nccl_lib) return fail(e, CLSR_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
  struct Id { char b[128]; };
  typedef int (*init_t)(void**, int, Id, int);
  init_t init = (init_t)dlsym(e->nccl_lib, "ncclCommInitRank");
  *(void**)(&e->ncclAllReduce_) = dlsym(e->nccl_lib, "ncclAllReduce");
  *(void**)(&e->ncclAllGather_) = dlsym(e->nccl_lib, "ncclAllGather");
  *(void**)(&e->ncclCommDestroy_) = dlsym(e->nccl_lib, "ncclCommDestroy");
  *(void**)(&e->ncclGetErrorString_) = dlsym(e->nccl_lib, "ncclGetErrorString");
  if (!init || !e->ncclAllReduce_ || !e->ncclAllGather_ || !e->ncclGetErrorString_)
    return fail(e, CLSR_ERR_NCCL, "libnccl is missing required symbols");
  CK(cudaSetDevice(e->cfg.device));
  Id id;
  memcpy(id.b, id128, 128);
  int r = init(&e->comm, world, id, rank);
  if (r != 0) return fail(e, CLSR_ERR_NCCL, "ncclCommInitRank failed: %s", e->ncclGetErrorString_(r));
  e->world = world;
  e->rank = rank;
  CK(cudaDeviceSynchronize());
  return CLSR_OK;
}

}  // extern "C"
