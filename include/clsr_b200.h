/* clsr_b200 -- C ABI of the B200-native CLSR training/inference step.
 *
 * The reference has no FFI of its own: its one process/device boundary is
 * `sess.run(fetches, feed_dict)` issued from
 *   reco_utils/recommender/deeprec/models/sequential/clsr.py:383-408        (train)
 *   reco_utils/recommender/deeprec/models/base_model.py:366-392             (eval / infer)
 *   reco_utils/recommender/deeprec/models/sequential/sequential_base_model.py:294-324
 *                                                       (eval_with_user[_and_alpha])
 * Each entry point below names the sess.run it replaces.  Conventions: every call
 * returns 0 on success or a negative clsr_status; the message is available from
 * clsr_last_error(); no exceptions cross the boundary; all device work is enqueued on
 * the engine's stream; an engine is used from one host thread at a time.
 */
#ifndef CLSR_B200_H
#define CLSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clsr_engine clsr_engine;

enum clsr_status {
  CLSR_OK = 0,
  CLSR_ERR_ARG = -1,     /* bad argument / unsupported configuration */
  CLSR_ERR_CUDA = -2,    /* a CUDA runtime call or kernel failed */
  CLSR_ERR_STATE = -3,   /* call sequence error (e.g. tables not bound) */
  CLSR_ERR_NCCL = -4
};

enum clsr_table {
  CLSR_TABLE_ITEM = 0,        /* sequential/embedding/item_embedding        [n_items, item_dim] */
  CLSR_TABLE_CATE = 1,        /* sequential/embedding/cate_embedding        [n_cates, cate_dim] */
  CLSR_TABLE_USER_LONG = 2,   /* sequential/embedding/user_long_embedding   [n_users, user_dim] */
  CLSR_TABLE_USER_SHORT = 3,  /* sequential/embedding/user_short_embedding  [n_users, user_dim] */
  CLSR_NUM_TABLES = 4
};

/* Hyper-parameters: clsr.yaml + the flag overrides of examples/00_quick_start/sequential.py:36-68. */
typedef struct clsr_config {
  int32_t device;            /* CUDA device ordinal */
  int32_t max_rows;          /* capacity in graph rows B (file lines x (1 + train_num_ngs)) */
  int32_t seq_len;           /* max_seq_length T */
  int32_t item_dim, cate_dim, user_dim, hidden;   /* hidden must equal item_dim + cate_dim */
  int32_t att0, att1;        /* att_fcn_layer_sizes */
  int32_t fc0, fc1;          /* layer_sizes */
  int64_t n_items, n_cates, n_users;
  int32_t train_group;       /* 1 + train_num_ngs (grouped softmax width) */
  float embed_l2, layer_l2;
  int32_t contrastive_kind;  /* 0 = triplet, 1 = bpr */
  float triplet_margin, contrastive_weight, discrepancy_weight;
  int32_t contrastive_len_threshold, contrastive_recent_k;
  int32_t optimizer;         /* 0 = adam (TF non-lazy sparse sweep), 1 = lazyadam */
  float learning_rate, beta1, beta2, adam_eps;
  int32_t clip_norm;         /* is_clip_norm */
  float max_grad_norm;
  float bn_momentum, bn_eps; /* 0.95 / 1e-4 (base_model.py:676-677) */
  int32_t math_mode;         /* 0 = fp32 SIMT everywhere, 1 = bf16 tcgen05 for the large GEMMs */
  int32_t max_seqs;          /* capacity in sequences (rows / group); 0 = max_rows (ungrouped batches of max_rows fit) */
  /* Graph variants of _build_seq_graph (clsr.py:159-274); all zero = the CLSR configuration the reference ships. */
  int32_t no_interest_evolve;    /* hparams.interest_evolve = False: short_term_intention = user_short_embedding (no GRU) */
  int32_t no_predict_long_short; /* hparams.predict_long_short = False: no causal2 GRU, alpha MLP input without its final state */
  int32_t manual_alpha;          /* hparams.manual_alpha = True: alpha = manual_alpha_value, no causal2 GRU, no alpha MLP */
  float manual_alpha_value;
  int32_t sequential_model;      /* hparams.sequential_model: 0 = time4lstm (the shipped configuration), 1 = lstm
                                  * (tf.nn.rnn_cell.LSTMCell, scope simple_lstm; clsr.py:209-216) */
} clsr_config;

/* One feed_dict (sequential_iterator.py:718-731), as plain arrays.  `group` declares that
 * rows s*group .. s*group+group-1 share users / histories / mask / time features (the
 * iterator replicates each training line 1+num_ngs times, :588-610); the engine then reads
 * those arrays at row s*group only.  group = 1 makes no assumption. */
typedef struct clsr_batch {
  int32_t rows;                 /* B */
  int32_t group;                /* G, must divide rows */
  int32_t on_device;            /* 0: host pointers (staged by the call), 1: device pointers */
  const int32_t* users;         /* [B] */
  const int32_t* items;         /* [B] */
  const int32_t* cates;         /* [B] */
  const int32_t* item_history;  /* [B, T] */
  const int32_t* cate_history;  /* [B, T] */
  const int32_t* mask;          /* [B, T] 0/1, left-aligned */
  const float* time_from_first_action; /* [B, T] */
  const float* time_to_now;     /* [B, T] */
  const float* labels;          /* [B] (train only; may be NULL for inference) */
} clsr_batch;

/* Outputs of one training step = the fetch list of CLSRModel.train (clsr.py:396-406). */
typedef struct clsr_losses {
  float loss, data_loss, regular_loss, contrastive_loss, discrepancy_loss;
  float table_grad_norm[4];  /* L2 norm of each table's gradient as clipped (diagnostic; see clsr_clip_report) */
} clsr_losses;

#define CLSR_STEP_NO_OPTIMIZER 1u /* stop after gradients (+ scatter-add); do not touch variables */
#define CLSR_STEP_NO_BN_UPDATE 2u /* do not move the BN moving statistics */

/* ---- lifetime ------------------------------------------------------------------- */
int clsr_create(const clsr_config* cfg, clsr_engine** out);     /* BaseModel.__init__ (base_model.py:18-71) */
void clsr_destroy(clsr_engine* e);
const char* clsr_last_error(const clsr_engine* e);              /* e may be NULL (creation errors) */
int clsr_set_stream(clsr_engine* e, void* cuda_stream);
int64_t clsr_workspace_bytes(const clsr_engine* e);

/* ---- variables -------------------------------------------------------------------
 * Dense (non-table) variables live in one engine-owned fp32 device buffer; entry i has
 * TF name clsr_dense_name(i), offset and element count in floats. */
int32_t clsr_dense_count(const clsr_engine* e);
const char* clsr_dense_name(const clsr_engine* e, int32_t i);
int64_t clsr_dense_offset(const clsr_engine* e, int32_t i);
int64_t clsr_dense_numel(const clsr_engine* e, int32_t i);
int32_t clsr_dense_trainable(const clsr_engine* e, int32_t i);
int64_t clsr_dense_total(const clsr_engine* e);
/* which: 0 = values, 1 = Adam m, 2 = Adam v, 3 = last gradient (before clipping) */
int clsr_dense_read(clsr_engine* e, int32_t which, float* host_dst);
int clsr_dense_write(clsr_engine* e, int32_t which, const float* host_src);
/* Tables are caller-owned device memory (fp32, row-major, 16-byte aligned).  m / v are the
 * Adam slots (may be NULL when the engine is used for inference only). */
int clsr_bind_table(clsr_engine* e, int32_t table, float* values, float* adam_m, float* adam_v);
int clsr_set_adam_step(clsr_engine* e, int64_t step);  /* number of completed updates */
int64_t clsr_get_adam_step(const clsr_engine* e);

/* ---- steps ---------------------------------------------------------------------- */
/* sess.run([update, update_ops, loss, ...]) of CLSRModel.train (clsr.py:383-408). */
int clsr_train_step(clsr_engine* e, const clsr_batch* batch, uint32_t flags, clsr_losses* out);
/* sess.run([pred, ...]) of eval / eval_with_user / eval_with_user_and_alpha / infer with
 * is_train_stage=False.  pred / alpha: host buffers of `rows` floats (alpha may be NULL). */
int clsr_predict(clsr_engine* e, const clsr_batch* batch, float* pred, float* alpha);
int clsr_synchronize(clsr_engine* e);
/* Clip diagnostics (base_model.py:289-297: per-variable tf.clip_by_norm; for a table the norm runs over the
 * concatenated, not yet de-duplicated IndexedSlices).  norms4: the four table-gradient norms the last
 * training step clipped with.  grouped_active_steps: number of steps so far that ran with group > 1 AND
 * had a table norm above max_grad_norm -- in those the norm was taken over group-summed history / user
 * slices and the update differs from TF's; group = 1 reproduces TF exactly. */
int clsr_clip_report(clsr_engine* e, float* norms4, int64_t* grouped_active_steps);

/* ---- standalone hot-path operators (benchmarks / parity tests) -------------------- */
/* K1+K3: hist_input[r,t,:] = concat(item_table[ih[r,t]], cate_table[ch[r,t]]) (clsr.py:145-147).
 * All pointers are device pointers; out is [rows, T, item_dim+cate_dim]. */
int clsr_gather_history(clsr_engine* e, const int32_t* item_hist, const int32_t* cate_hist,
                        int64_t positions, float* out);
/* K13: sparse-gradient scatter-add of d_hist [positions, D] into compact per-unique-id rows.
 * Returns the unique ids / compact rows through device pointers owned by the engine. */
int clsr_scatter_history_grad(clsr_engine* e, const int32_t* item_hist, const int32_t* cate_hist,
                              int64_t positions, const float* d_hist);
int clsr_sparse_grad_view(clsr_engine* e, int32_t table, const int32_t** unique_ids,
                          const float** rows, const int32_t** count);

/* ---- multi-GPU (data parallel over sequences; NCCL over NVLink) -------------------- */
int clsr_nccl_unique_id(void* out128);                        /* 128 bytes */
int clsr_comm_init(clsr_engine* e, int32_t rank, int32_t world, const void* id128);

/* Peer-memory communication for the data-parallel step (after clsr_comm_init, world a power of two): every
 * rank exports a blob, the blobs of all ranks (rank-major, CLSR_PEER_BLOB_BYTES each) are exchanged through any
 * host channel and attached.  From then on the step's small reductions (BatchNorm sums, loss partial sums,
 * counts, barriers) are one-shot all-reduces over NVLink peer memory inside the consuming kernels.
 * shard_tables != 0 additionally row-shards the four tables over the ranks (global row r on rank r % world at
 * local row r / world, engine-owned memory: fill / read the local shard through clsr_table_local): history and
 * target gathers read straight from the owning GPU, every rank pushes its de-duplicated gradient rows to
 * their owners (16-byte NVLink reductions), and each owner applies the optimizer to its 1/world of the rows. */
#define CLSR_PEER_BLOB_BYTES 1024
int clsr_peer_setup_begin(clsr_engine* e, int32_t shard_tables, void* blob_out);
int clsr_peer_setup_finish(clsr_engine* e, const void* all_blobs);
int clsr_table_local(clsr_engine* e, int32_t table, int32_t which, float** dev_ptr, int64_t* rows);

/* ---- row-sharded embedding tables (BASELINE configs 4-5) ------------------
 * For tables that outgrow one GPU: global row r lives on rank r % world at local row r / world.
 * The reference has no counterpart (tf.nn.embedding_lookup on one device,
 * sequential_base_model.py:381-437); these entry points replace that lookup and its
 * IndexedSlices gradient for a sharded table.  Peers map each other's shards through CUDA IPC
 * (NVLink / NVSwitch peer memory); gather and scatter-add are single kernels that read /
 * reduce straight into the owning GPU's memory - there is no id or row exchange step.
 * The caller orders steps across ranks (one barrier between an owner's update and its peers'
 * gathers, one between the peers' scatter-adds and the owner's optimizer). */
#define CLSR_SHARD_MAX_WORLD 16
typedef struct clsr_shard_table clsr_shard_table;
int clsr_shard_create(int32_t device, int32_t rank, int32_t world, int64_t n_rows, int32_t dim, int32_t with_grad,
                      clsr_shard_table** out);
void clsr_shard_destroy(clsr_shard_table* t);
const char* clsr_shard_last_error(const clsr_shard_table* t);   /* t may be NULL (creation errors) */
int64_t clsr_shard_local_rows(const clsr_shard_table* t);
float* clsr_shard_local_values(clsr_shard_table* t);            /* device pointer [local_rows, dim] */
float* clsr_shard_local_grad(clsr_shard_table* t);              /* device pointer or NULL */
int clsr_shard_export(clsr_shard_table* t, void* handles_out);  /* 2 x 64 bytes: values, gradient */
int clsr_shard_attach(clsr_shard_table* t, const void* all_handles); /* world x 128 bytes, rank-major */
int clsr_shard_zero_grad(clsr_shard_table* t, void* cuda_stream);
/* K1+K3 on sharded tables: out[p,:] = concat(item[ih[p]], cate[ch[p]]); device pointers. */
int clsr_shard_gather_history(clsr_shard_table* item, clsr_shard_table* cate, const int32_t* item_hist,
                              const int32_t* cate_hist, int64_t positions, float* out, void* cuda_stream);
/* K13 on sharded tables: owner's gradient shard row += d_hist[p, item | cate columns]. */
int clsr_shard_scatter_add_history(clsr_shard_table* item, clsr_shard_table* cate, const int32_t* item_hist,
                                   const int32_t* cate_hist, int64_t positions, const float* d_hist,
                                   void* cuda_stream);

/* ---- introspection for parity tests ------------------------------------------------ */
/* Named intermediate buffers of the last step (device pointers, fp32). */
int clsr_debug_buffer(clsr_engine* e, const char* name, const float** dev_ptr, int64_t* numel);
/* Synchronise the engine stream, then copy `bytes` from device memory to the host. */
int clsr_debug_read(clsr_engine* e, const void* dev_src, void* host_dst, int64_t bytes);
int clsr_set_debug_sync(clsr_engine* e, int32_t on);  /* sync + check after every kernel */
int64_t clsr_kernel_launches(const clsr_engine* e);   /* kernels launched by the last step */
/* Single-GPU training steps are replayed as one CUDA graph per batch shape (first step of a shape eager, second
 * captured, later ones one cudaGraphLaunch; per-step scalars such as Adam's step size live in device memory).
 * clsr_set_graphs(e, 0) turns that off (every step launches its kernels one by one), clsr_graph_replays counts the
 * steps that ran as a graph launch so far. */
int clsr_set_graphs(clsr_engine* e, int32_t on);
int64_t clsr_graph_replays(const clsr_engine* e);
/* Standalone linear layer C[M,N] = A[M,K].W[K,N] + bias on device pointers; mode 0 = fp32 SIMT kernel,
 * 1 = tcgen05 split-bf16 kernel.  Used by the kernel-level parity tests. */
int clsr_debug_gemm(clsr_engine* e, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* W,
                    int32_t ldw, const float* bias, float* C, int32_t ldc, int32_t mode);
/* Standalone weight-gradient product dW[K,N] += A[M,K]^T . B[M,N], colsum[n] += sum_m B[m,n] (may be NULL). */
int clsr_debug_dwgemm(clsr_engine* e, int32_t M, int32_t K, int32_t N, const float* A, int32_t lda, const float* B,
                      int32_t ldb, float* dW, int32_t lddw, float* colsum, int32_t mode);
/* Per-kernel device time: with profiling on, a CUDA event is recorded on the engine stream after
 * every launch; collect() folds the intervals into per-kernel totals (ms, calls) by name. */
int clsr_set_profiling(clsr_engine* e, int32_t on);
int clsr_profile_collect(clsr_engine* e);
int clsr_profile_entry(clsr_engine* e, int32_t i, char* name, int32_t name_cap, double* total_ms,
                       int64_t* calls);

/* ---- batch construction on the GPU -----------------------------------------------
 * Replaces SASequentialIterator._convert_data (sequential_iterator.py:519-704: pad, replicate x(1+num_ngs), draw
 * in-batch negatives, row by row in Python).  clsr_dataset_create uploads one parsed file as columnar arrays
 * (host pointers: per line label, user, item, cate, history length and the last T events of the item / category /
 * time-feature histories, left-aligned, [n_lines, T]); ids are validated against the tables here, once.
 * clsr_build_batch builds the feed of `count` lines (host array of line indices) on the device, straight into the
 * block the step reads: num_ngs = 0 gives the evaluation batch (one row per line, the file's labels); num_ngs > 0
 * the training batch (row 0 of each group = the line's positive, rows 1.. = positive items of other lines of the
 * batch, never the line's own item, drawn by a counter-based generator from `seed`).  clsr_train_step_staged /
 * clsr_predict_staged then run on that batch with no host feed at all. */
typedef struct clsr_dataset clsr_dataset;
int clsr_dataset_create(clsr_engine* e, int64_t n_lines, const float* label, const int32_t* user, const int32_t* item,
                        const int32_t* cate, const int32_t* length, const int32_t* item_hist, const int32_t* cate_hist,
                        const float* time_from_first_action, const float* time_to_now, clsr_dataset** out);
void clsr_dataset_destroy(clsr_dataset* d);
int clsr_build_batch(clsr_engine* e, clsr_dataset* d, const int32_t* host_lines, int32_t count, int32_t num_ngs,
                     uint64_t seed);
int clsr_train_step_staged(clsr_engine* e, uint32_t flags, clsr_losses* out);
int clsr_predict_staged(clsr_engine* e, float* dev_pred, float* dev_alpha, int32_t* dev_users, float* dev_labels);
int clsr_staged_feed_read(clsr_engine* e, int32_t which, void* host_dst, int64_t bytes);

/* ---- evaluation on the GPU ----------------------------------------------------
 * clsr_predict_device: clsr_predict with the predictions left in device memory (pred / alpha: device buffers of
 * `rows` floats, alpha may be NULL; asynchronous on the engine stream) -- run_eval / run_weighted_eval
 * (sequential_base_model.py:204-292) accumulate a whole file there instead of in Python lists.
 * clsr_eval_metrics_compute: cal_metric + cal_weighted_metric (deeprec_utils.py:554-810) over n rows held on the
 * device: auc and logloss over all rows; mean_mrr, ndcg@k, hit@k, group_auc over impressions of `group`
 * consecutive rows (0: skip); wauc = per-user AUC weighted by the user's share of the rows (users may be NULL:
 * skip).  status: bit 0 / 1 / 2 = all rows / a user / an impression hold a single label class (sklearn raises
 * ValueError there and so does the Python wrapper). */
typedef struct clsr_eval_metrics {
  double auc, logloss, mean_mrr, group_auc, wauc;
  double ndcg[8], hit[8];
  int64_t n, n_pos, n_groups, n_users;
  int32_t status;
} clsr_eval_metrics;
int clsr_predict_device(clsr_engine* e, const clsr_batch* batch, float* dev_pred, float* dev_alpha);
int clsr_eval_metrics_compute(int32_t device, const float* preds, const float* labels, const int32_t* users, int64_t n,
                              int32_t group, const int32_t* ks, int32_t nk, clsr_eval_metrics* out, void* cuda_stream);

/* ---- host utility ------------------------------------------------------------------- */
/* CRC-32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start).  Used by the tensor-bundle
 * checkpoint writer: tf.train.Saver.restore (base_model.py:394-410) verifies the per-tensor crc32c. */
uint32_t clsr_crc32c(const void* data, uint64_t bytes, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif
