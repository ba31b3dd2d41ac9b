/*
 * recbox_b200.h -- C ABI of librecbox_b200.so: the sm_100a (B200) implementation of RecBox's
 * embedding + feature-interaction hot path.
 *
 * The reference (reczoo/RecBox) is pure Python; it has no FFI of its own.  The boundary this
 * library replaces is the arithmetic its nn.Module layers hand to PyTorch ATen.  Every entry
 * point below names the reference code (file:line under /root/reference/recbox) whose work it
 * performs; INTEGRATION.md shows the ctypes binding a RecBox maintainer adds on their side.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no torch types.  Pointers marked DEVICE are device memory
 *     of the current CUDA device, pointers marked HOST are small host arrays read during the
 *     call (they are copied into the kernel's parameter space; nothing is retained).
 *   - all device work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises, nothing allocates, no library-owned threads.
 *   - return 0 on success, <0 on error (RBX_ERR_*); rbx_last_error() returns the message of the
 *     calling thread's last failure.
 *   - fp32 rows, int32 global row ids.  "Global row" = id + row offset of the slot's table inside
 *     the fused table (all per-feature nn.Embedding weights live back to back in one [R, D]
 *     allocation, see DESIGN.md "Data layout").
 *   - gradient outputs ACCUMULATE (+=); the caller zero-fills (this is what lets autograd's
 *     zero_grad(set_to_none=False) semantics and multi-call accumulation work unchanged).
 */
#ifndef RECBOX_B200_H
#define RECBOX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBX_OK               0
#define RBX_ERR_ARG         -1   /* bad argument (null pointer, unsupported size ...) */
#define RBX_ERR_CUDA        -2   /* a CUDA runtime call or launch failed */
#define RBX_ERR_UNSUPPORTED -3   /* shape outside what the kernels cover */

#define RBX_MAX_SLOTS      192   /* max categorical (resp. numeric) slots per call */
#define RBX_MAX_DIM        512   /* max embedding dim */
#define RBX_MAX_WORLD        8   /* GPUs of one NVSwitch box a table can be sharded over */
#define RBX_PEER_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

typedef void* rbx_stream_t;      /* cudaStream_t */

int         rbx_version(void);            /* 10000*major + 100*minor + patch */
const char* rbx_last_error(void);         /* thread-local, never NULL */
int         rbx_device_sm_count(void);    /* SM count of the current device (<0 on error) */
/* Reserve `bytes` of the current device's L2 for persisting (evict_last) lines -- what the fused
 * kernels' table / gradient-table hints live in.  Clamped to the device maximum; returns the size
 * in effect, <0 on error.  Process-wide device setting: the host decides, the kernels only hint. */
long long   rbx_l2_set_persisting_bytes(long long bytes);

/* ------------------------------------------------------------------------------------------
 * a1  batch matrix -> slot blocks
 * Replaces RankingModel.get_inputs / get_labels (ranking/pytorch/models/ranking_model.py:106-122:
 * one column slice + .to(device) per feature) and the per-feature casts `.long()` / `.float()`
 * of FeatureEmbeddingDict.forward (ranking/pytorch/layers/embeddings/feature_embedding.py:201,204).
 * One pass over the [B, n_cols] float64 batch (h5_dataloader.py:46 builds it with np.hstack):
 *   col_kind[c] = 0 ignore | 1 categorical | 2 numeric | 3 label ; col_slot[c] = slot index.
 *   rows[b, s]   = (int32)trunc(batch[b,c]) + field_off[s]      (kind 1)
 *   dense_x[b,s] = (float)batch[b,c]                            (kind 2)
 *   label[b]     = (float)batch[b,c]                            (kind 3)
 * field_rows[s] (nullable) is the vocabulary size of slot s: an id outside [0, field_rows[s]) -- where
 * nn.Embedding raises IndexError -- becomes row -1 (reads as a zero row, receives no gradient) instead of
 * aliasing a neighbouring feature's rows in the fused table, and is counted in n_bad[0] (DEVICE int32,
 * nullable, accumulated; the caller zeroes it).
 * ------------------------------------------------------------------------------------------ */
int rbx_split_batch_f64(const double* batch /*DEVICE [B, ld]*/, int64_t B, int n_cols, int64_t ld,
                        const int8_t* col_kind /*HOST [n_cols]*/,
                        const int16_t* col_slot /*HOST [n_cols]*/,
                        const int64_t* field_off /*HOST [F]*/,
                        const int64_t* field_rows /*HOST [F] | NULL*/, int F, int Fn,
                        int32_t* rows /*DEVICE [B,F] | NULL*/,
                        float* dense_x /*DEVICE [B,Fn] | NULL*/,
                        float* label /*DEVICE [B] | NULL*/,
                        int32_t* n_bad /*DEVICE [1] | NULL*/,
                        rbx_stream_t stream);

/* Same conversion for the dict-of-columns form of X (feature_embedding.py:188-214 receives
 * {feature: Tensor[B]}): up to RBX_MAX_SLOTS separate column pointers with element strides.
 * dtype codes: 0 = float64, 1 = float32, 2 = int64, 3 = int32.  Writes out[b, s] for every
 * column s as int32 (+ add[s]) when as_rows != 0, else as float32.  vocab[s] / n_bad: as field_rows / n_bad above. */
int rbx_pack_columns(const void* const* cols /*HOST [n] of DEVICE ptrs*/,
                     const int64_t* strides /*HOST [n], in elements*/,
                     const int8_t* dtypes /*HOST [n]*/,
                     const int64_t* add /*HOST [n] | NULL*/,
                     const int64_t* vocab /*HOST [n] | NULL*/,
                     int n, int64_t B, int as_rows,
                     void* out /*DEVICE [B, n] int32 or float32*/,
                     int32_t* n_bad /*DEVICE [1] | NULL*/,
                     rbx_stream_t stream);

/* Compact form of the id block (recbox_b200.loader.PackedDataset, every vocabulary < 65 536 -- the
 * Criteo config of BASELINE configs[1]): uint16 local ids [B,F] cross PCIe (2 B instead of the 8 B the
 * reference's float64 batch matrix spends, h5_dataloader.py:46); rows[b,f] = ids[b,f] + field_off[f]. */
int rbx_unpack_ids_u16(const uint16_t* ids /*DEVICE [B,F]*/, int64_t B, int F, const int64_t* field_off /*HOST [F] | NULL*/,
                       int32_t* rows /*DEVICE [B,F]*/, rbx_stream_t stream);

/* optimizer.zero_grad() of the fused gradient buffer (ranking_model.py:192): one streaming fill, so the
 * zero-fill can run on a side stream under the forward instead of on the backward's critical path. */
int rbx_zero_f32(float* ptr /*DEVICE [n]*/, int64_t n, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a2+a3+a6+a7+a8  fused multi-slot gather + FM + LR, forward      (kernel K1/K2)
 * One launch does what the reference spreads over
 *   FeatureEmbeddingDict.forward + dict2tensor   feature_embedding.py:169-214 (26 x aten::embedding,
 *                                                13 x Linear(1,D), torch.stack)
 *   LogisticRegression.forward                   blocks/logistic_regression.py:30-35 (2nd, D=1 lookup)
 *   InnerProductInteraction("product_sum")       interactions/inner_product.py:40-48
 *   FactorizationMachine.forward                 blocks/factorization_machine.py:30-34
 * For sample b, slot f (categorical) e = table[rows[b,f], :]; slot n (numeric)
 * e = dense_x[b,n] * dense_w[n, :].  Ft = F + Fn, slot output positions cat_pos / num_pos.
 *   E[b, pos, :] = e                                       (nullable)
 *   S[b, :]      = sum over slots of e                      (nullable; saved for backward)
 *   fm_out[b]    = sum_d 0.5 * (S_d^2 - sum_slots e_d^2)    (nullable)
 *   lr_out[b]    = sum_f table_lr[rows[b,f]] + sum_n dense_x[b,n]*dense_w_lr[n] + lr_bias[0]
 *                                                           (nullable; needs table_lr)
 * With E, S and fm_out all NULL the call is the first-order term alone (LogisticRegression
 * without a D-dim table): table / dense_w may be NULL and D is ignored.
 * R = rows in the fused table.  A row id outside [0, R) reads as a zero row and receives no
 * gradient (the reference raises IndexError on CPU / device-asserts on CUDA; a library must not
 * fault the device).
 * ------------------------------------------------------------------------------------------ */
int rbx_embed_fm_fwd(const float* table /*DEVICE [R,D]*/,
                     const float* table_lr /*DEVICE [R] | NULL*/,
                     const int32_t* rows /*DEVICE [B,F]*/,
                     const int32_t* cat_pos /*HOST [F]*/,
                     const int32_t* lr_delta /*HOST [F] | NULL: row in table_lr = rows[b,f] + lr_delta[f]*/,
                     const float* dense_x /*DEVICE [B,Fn] | NULL*/,
                     const float* dense_w /*DEVICE [*,D] | NULL*/,
                     const float* dense_w_lr /*DEVICE [*] | NULL*/,
                     const int32_t* num_pos /*HOST [Fn]*/,
                     const int32_t* num_widx /*HOST [Fn] | NULL: row of slot n in dense_w / dense_w_lr (identity)*/,
                     const float* lr_bias /*DEVICE [1] | NULL*/,
                     float* E /*DEVICE [B,Ft,D] | NULL*/,
                     float* S /*DEVICE [B,D] | NULL*/,
                     float* fm_out /*DEVICE [B] | NULL*/,
                     float* lr_out /*DEVICE [B] | NULL*/,
                     int64_t B, int64_t R, int F, int Fn, int D,
                     int n_slots /* slots of E (row stride n_slots*D); 0 = F + Fn.  Slots not named by
                                    cat_pos / num_pos are left untouched (pooled sequence slots) */,
                     rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5+a6+a7 backward: grad scatter-add                                (kernel K3)
 * Replaces aten::embedding_dense_backward x52 and the Pow/Sum backward chain autograd builds
 * for inner_product.py:42-48 (loss.backward(), ranking_model.py:194).
 *   g_e[b,slot,:] = dE[b,pos,:] + d_fm[b] * (S[b,:] - e[b,slot,:])
 *   g_table[rows[b,f], :] += g_e        unless rows[b,f] == pad_row[f]   (padding_idx: zero grad)
 *   g_table_lr[rows[b,f]] += d_lr[b]    unless padding row
 *   g_dense_w[n,:]  += sum_b dense_x[b,n] * g_e[b,n,:] ;  g_dense_w_lr[n] += sum_b dense_x[b,n]*d_lr[b]
 *   g_lr_bias[0]    += sum_b d_lr[b]
 * e is re-read from E (streaming) when E != NULL, else re-gathered from `table`.
 * dE, d_fm, d_lr are each nullable (treated as zero); S is required when d_fm != NULL.
 * Atomic (red.global.add) accumulation: order is not deterministic; see rbx_scatter_add_sorted
 * for the deterministic path.
 * ------------------------------------------------------------------------------------------ */
int rbx_embed_fm_bwd(const float* table /*DEVICE [R,D] | NULL if E given*/,
                     const int32_t* rows /*DEVICE [B,F]*/,
                     const int32_t* cat_pos /*HOST [F]*/,
                     const int32_t* pad_row /*HOST [F], -1 = none*/,
                     const int32_t* lr_delta /*HOST [F] | NULL*/,
                     const float* dense_x /*DEVICE [B,Fn] | NULL*/,
                     const float* dense_w /*DEVICE [*,D] | NULL*/,
                     const int32_t* num_pos /*HOST [Fn]*/,
                     const int32_t* num_widx /*HOST [Fn] | NULL*/,
                     const float* E /*DEVICE [B,Ft,D] | NULL*/,
                     const float* S /*DEVICE [B,D] | NULL*/,
                     const float* dE /*DEVICE [B,Ft,D] | NULL*/,
                     const float* d_fm /*DEVICE [B] | NULL*/,
                     const float* d_lr /*DEVICE [B] | NULL*/,
                     float* g_table /*DEVICE [R,D] | NULL*/,
                     float* g_table_lr /*DEVICE [R] | NULL*/,
                     float* g_dense_w /*DEVICE [Fn,D] | NULL*/,
                     float* g_dense_w_lr /*DEVICE [Fn] | NULL*/,
                     float* g_lr_bias /*DEVICE [1] | NULL*/,
                     int64_t B, int64_t R, int F, int Fn, int D, int n_slots /* of E and dE; 0 = F + Fn */,
                     rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5  plain row gather / scatter-add (un-pooled sequence features, SASRec's three shared-table
 * lookups third_party/rechub/models/matching/sasrec.py:99-100; aten::embedding and
 * aten::embedding_dense_backward).  out[i,:] = table[ids[i],:]; g_table[ids[i],:] += g[i,:]
 * unless ids[i] == pad_row.
 * ------------------------------------------------------------------------------------------ */
int rbx_gather_rows(const float* table /*DEVICE [R,D]*/, const int32_t* ids /*DEVICE [N]*/,
                    float* out /*DEVICE [N,D]*/, int64_t N, int D, rbx_stream_t stream);
int rbx_scatter_add_rows(const float* g /*DEVICE [N,D]*/, const int32_t* ids /*DEVICE [N]*/,
                         int32_t pad_row, float* g_table /*DEVICE [R,D]*/,
                         int64_t N, int D, rbx_stream_t stream);
/* Deterministic form: the caller sorts the ids STABLY (sorted_ids ascending, order[i] = original
 * position of sorted entry i); every run of equal ids is summed in position order by one warp and added to its row with a
 * plain read-modify-write -- no atomics, bit-identical results from run to run. */
int rbx_segment_sum_rows(const float* g /*DEVICE [N,D]*/, const int64_t* order /*DEVICE [N]*/,
                         const int32_t* sorted_ids /*DEVICE [N]*/, int32_t pad_row, float* g_table /*DEVICE [R,D]*/,
                         int64_t N, int D, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a9  pooled sequence gather  (nn.Embedding on [B,L] ids followed by MaskedSumPooling /
 * MaskedAveragePooling: core/pytorch/layers/sequence.py:4-20, ranking/pytorch/layers/pooling.py:22-40)
 * without materialising [B,L,D].  mode 0 = sum, 1 = masked average: divide by
 * (#positions l whose gathered row sums to non-zero) + 1e-12, the reference's own mask rule.
 * ids has row stride ids_ld (a sequence feature is L consecutive columns of the [B, n] int32 slot
 * block rbx_split_batch_f64 writes); out has row stride out_ld floats (so it can be a slot of E).  cnt[b] (nullable) receives the
 * divisor's count for backward.
 * Backward: g_table[ids[b,l],:] += g[b,:] * (mode ? 1/(cnt[b]+1e-12) : 1), padding rows skipped.
 * ------------------------------------------------------------------------------------------ */
int rbx_pooled_gather_fwd(const float* table, const int32_t* ids /*DEVICE [B, ids_ld]*/, int64_t ids_ld,
                          float* out /*DEVICE [B, out_ld]*/, int64_t out_ld,
                          float* cnt /*DEVICE [B] | NULL*/,
                          int64_t B, int L, int D, int mode, rbx_stream_t stream);
int rbx_pooled_gather_bwd(const float* g /*DEVICE [B, g_ld]*/, int64_t g_ld,
                          const int32_t* ids, int64_t ids_ld, const float* cnt /*DEVICE [B] | NULL (mode 0)*/,
                          int32_t pad_row, float* g_table,
                          int64_t B, int L, int D, int mode, rbx_stream_t stream);

/* The same two pooling layers applied to an already materialised emb [B,L,D] (the modules called
 * directly: core/pytorch/layers/sequence.py:4-20, ranking/pytorch/layers/pooling.py:22-40).
 * mask (uint8 [B,L], nullable) replaces the row-sum != 0 rule (pooling.py:27-29).  cnt[b] receives
 * the mask count (mode 1).  Backward: d_emb[b,l,:] = g[b,:] / (cnt[b] + 1e-12) (mode 1) or g[b,:]
 * (mode 0) for EVERY l -- the sum's gradient reaches masked positions too, as in autograd. */
int rbx_pool_fwd(const float* emb /*DEVICE [B,L,D]*/, const uint8_t* mask /*DEVICE [B,L] | NULL*/,
                 float* out /*DEVICE [B,D]*/, float* cnt /*DEVICE [B] | NULL*/,
                 int64_t B, int L, int D, int mode, rbx_stream_t stream);
int rbx_pool_bwd(const float* g /*DEVICE [B,D]*/, const float* cnt /*DEVICE [B] | NULL (mode 0)*/,
                 float* d_emb /*DEVICE [B,L,D]*/, int64_t B, int L, int D, int mode, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a10  two-tower scores  y[b,k] = <u[b,:], v[b,k,:]>   (layout fixed by
 * matching/pytorch/models/match_model.py:71-75; rechub dssm.py:48 is K = 1)
 * Backward: du[b,:] = sum_k dy[b,k] v[b,k,:] ; dv[b,k,:] = dy[b,k] u[b,:]   (overwrite, not +=)
 * ------------------------------------------------------------------------------------------ */
int rbx_rowdot_fwd(const float* u /*DEVICE [B,D]*/, const float* v /*DEVICE [B,K,D]*/,
                   float* y /*DEVICE [B,K]*/, int64_t B, int K, int D, rbx_stream_t stream);
int rbx_rowdot_bwd(const float* u, const float* v, const float* dy,
                   float* du /*DEVICE [B,D] | NULL*/, float* dv /*DEVICE [B,K,D] | NULL*/,
                   int64_t B, int K, int D, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6  InnerProductInteraction on a materialised E [B,F,D]  (interactions/inner_product.py:40-56)
 * mode 0 product_sum -> [B] ; 1 bi_interaction -> [B,D] ; 2 inner_product -> [B,F(F-1)/2]
 * (row-major upper triangle, i<j) ; 3 elementwise_product -> [B,F(F-1)/2,D].
 * Backward writes dE (overwrite).
 * ------------------------------------------------------------------------------------------ */
int rbx_interact_fwd(const float* E, float* out, int64_t B, int F, int D, int mode,
                     rbx_stream_t stream);
int rbx_interact_bwd(const float* E, const float* dout, float* dE, int64_t B, int F, int D,
                     int mode, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (e)  row-shard routing around the all-to-all (no reference counterpart --
 * the reference has no sharded embedding).  owner(r) = r % world, local row = r / world.
 * rbx_shard_route  : STABLE bucket by owner.  send[pos[i]] = rows[i] / world with pos[i] = start of
 *                    owner's bucket + rank of i among the ids of that owner in input order;
 *                    counts[w] = size of bucket w.  Deterministic; equals a stable argsort by owner.
 *                    ws = rbx_shard_ws_bytes(N, world) bytes of scratch.
 * rbx_shard_permute: out[pos[i],:] = in[i,:]     (payload into send order: grads going to owners)
 * rbx_shard_unroute: out[i,:] = recv[pos[i],:]   (rows coming back in send order)
 * ------------------------------------------------------------------------------------------ */
size_t rbx_shard_ws_bytes(int64_t N, int world);
int rbx_shard_route(const int32_t* rows /*DEVICE [N]*/, int64_t N, int world,
                    void* ws /*DEVICE*/, size_t ws_bytes,
                    int32_t* send /*DEVICE [N]*/, int32_t* pos /*DEVICE [N]*/,
                    int32_t* counts /*DEVICE [world]*/, rbx_stream_t stream);
int rbx_shard_permute(const float* in /*DEVICE [N,D]*/, const int32_t* pos /*DEVICE [N]*/,
                      float* out /*DEVICE [N,D]*/, int64_t N, int D, rbx_stream_t stream);
int rbx_shard_unroute(const float* recv /*DEVICE [N,D]*/, const int32_t* pos /*DEVICE [N]*/,
                      float* out /*DEVICE [N,D]*/, int64_t N, int D, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (e)  row-sharded table over NVLink peer memory: the fused kernels with the exchange inside.
 * No reference counterpart (the reference has no sharded embedding).
 * Global row r lives on shard r % world (world a power of two <= RBX_MAX_WORLD) at local row
 * r / world.  shard_tables[w] is a pointer valid on the CALLING device: this rank's own
 * allocation for w == rank, a peer's allocation mapped with rbx_peer_open otherwise.  The forward
 * reads remote rows with plain loads over NVLink/NVSwitch; the backward issues its reductions
 * (red.global.add) straight into the owner's gradient table.  Everything else is
 * rbx_embed_fm_fwd / rbx_embed_fm_bwd (rows are GLOBAL row ids; pad_row too; the first-order
 * table shares the row numbering).  The caller separates steps with a cross-rank barrier: remote
 * reductions of step t must land before the owner's optimizer reads its gradient shard.
 * Covers D in {4,8,16,32,64,128} with 16-byte aligned buffers; RBX_ERR_UNSUPPORTED otherwise.
 * ------------------------------------------------------------------------------------------ */
/* Tell the library which shard this process owns (one process per GPU; -1 = unknown, the default).  Only a hint: the
 * sharded backward may route rows it owns and rows a peer owns differently (RBX_BWD_BULK=2 build: TMA bulk
 * reductions for peer rows).  Results do not depend on it. */
int rbx_shard_set_rank(int rank);

int rbx_embed_fm_fwd_sharded(const float* const* shard_tables /*HOST [world] of DEVICE [R_w, D]*/,
                             const float* const* shard_tables_lr /*HOST [world] of DEVICE [R_w] | NULL*/,
                             int world,
                             const int32_t* rows /*DEVICE [B,F] global rows*/, const int32_t* cat_pos,
                             const float* dense_x, const float* dense_w, const float* dense_w_lr,
                             const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias,
                             float* E, float* S, float* fm_out, float* lr_out,
                             int64_t B, int64_t R /*global rows*/, int F, int Fn, int D, int n_slots,
                             rbx_stream_t stream);
int rbx_embed_fm_bwd_sharded(const float* const* shard_tables /*HOST [world] | NULL when E is given*/,
                             float* const* shard_g_tables /*HOST [world] of DEVICE [R_w, D] | NULL*/,
                             float* const* shard_g_tables_lr /*HOST [world] of DEVICE [R_w] | NULL*/,
                             int world,
                             const int32_t* rows, const int32_t* cat_pos, const int32_t* pad_row,
                             const float* dense_x, const float* dense_w, const int32_t* num_pos,
                             const int32_t* num_widx, const float* E, const float* S, const float* dE,
                             const float* d_fm, const float* d_lr,
                             float* g_dense_w, float* g_dense_w_lr, float* g_lr_bias,
                             int64_t B, int64_t R, int F, int Fn, int D, int n_slots, rbx_stream_t stream);

/* ROW+LR shard layout: every shard is [cap, 2 D] floats, physical row = [e_0 .. e_{D-1} | w_lr | 0 ...], and the
 * gradient shards mirror it, so an embedding row and its first-order weight (resp. their gradients) cross NVLink
 * in ONE request instead of two -- remote traffic of 64-byte rows is bound by request count.
 * D in {4, 8, 16}.  Same outputs as rbx_embed_fm_fwd/bwd_sharded. */
int rbx_embed_fm_fwd_sharded_rowlr(const float* const* shard_tables /*HOST [world] of DEVICE [cap,2D]*/, int world,
                                   const int32_t* rows, const int32_t* cat_pos /*HOST*/,
                                   const float* dense_x, const float* dense_w, const float* dense_w_lr,
                                   const int32_t* num_pos /*HOST*/, const int32_t* num_widx /*HOST | NULL*/,
                                   const float* lr_bias, float* E, float* S, float* fm_out, float* lr_out,
                                   int64_t B, int64_t R, int F, int Fn, int D, int n_slots, rbx_stream_t stream);
int rbx_embed_fm_bwd_sharded_rowlr(const float* const* shard_tables /*| NULL when E is given*/,
                                   float* const* shard_g_tables /*HOST [world] of DEVICE [cap,2D]*/, int world,
                                   const int32_t* rows, const int32_t* cat_pos, const int32_t* pad_row,
                                   const float* dense_x, const float* dense_w, const int32_t* num_pos,
                                   const int32_t* num_widx, const float* E, const float* S, const float* dE,
                                   const float* d_fm, const float* d_lr,
                                   float* g_dense_w, float* g_dense_w_lr, float* g_lr_bias,
                                   int64_t B, int64_t R, int F, int Fn, int D, int n_slots, rbx_stream_t stream);

/* Push-based exchange (the path used for tables too large for remote gathers, see
 * csrc/shard_push.cu): only dense, contiguous buffers cross NVLink; every random access is local
 * to the row's owner.  All `*const*` arguments are HOST arrays of `world` DEVICE pointers valid on
 * the calling device (own allocation or rbx_peer_open mapping); `cap` = slot capacity in ids per
 * (owner, requester) pair; inbox_meta is int32 [world][4] = {count, offset, overflow flag, 0}.
 *   rbx_shard_push_ids   requester: inbox_ids[w][rank*cap + j] = send[start_w + j] for every owner w
 *                        (send / counts from rbx_shard_route), plus (count, start_w) into w's meta
 *   rbx_shard_serve_rows owner: out_rows[q][(off_q + j)*D ..] = table[inbox_ids[q*cap + j]] (and the
 *                        first-order value into out_lr[q]) -- rows land in q's buffer in q's send order
 *   rbx_shard_push_grads requester: ginbox[w][(rank*cap + j)*D ..] = gsend[(start_w + j)*D ..]
 *   rbx_shard_apply_grads owner: g_table[inbox_ids[q*cap + j]] += ginbox[q*cap + j]; then clears the
 *                        owned padding rows (pad_local: DEVICE int32 [n_pad] local row numbers)
 * The caller separates the phases with cross-rank barriers. */
int rbx_shard_push_ids(const int32_t* send /*DEVICE [N]*/, const int32_t* counts /*DEVICE [world]*/,
                       int32_t* const* inbox_ids, int32_t* const* inbox_meta,
                       int rank, int world, int64_t cap, int64_t N, rbx_stream_t stream);
int rbx_shard_serve_rows(const float* table /*DEVICE [R_w,D]*/, const float* table_lr /*DEVICE [R_w] | NULL*/, int D,
                         const int32_t* inbox_ids /*DEVICE [world,cap]*/, const int32_t* inbox_meta /*DEVICE [world,4]*/,
                         float* const* out_rows, float* const* out_lr /*| NULL*/,
                         int world, int64_t cap, rbx_stream_t stream);
int rbx_shard_push_grads(const float* gsend /*DEVICE [N,D]*/, const float* gsend_lr /*DEVICE [N] | NULL*/,
                         const int32_t* counts /*DEVICE [world]*/,
                         float* const* ginbox, float* const* ginbox_lr /*| NULL*/,
                         int rank, int world, int64_t cap, int D, int64_t N, rbx_stream_t stream);
int rbx_shard_apply_grads(const float* ginbox /*DEVICE [world,cap,D]*/, const float* ginbox_lr /*DEVICE [world,cap] | NULL*/,
                          const int32_t* inbox_ids, const int32_t* inbox_meta,
                          float* g_table, float* g_table_lr /*| NULL*/,
                          int world, int64_t cap, int D,
                          const int32_t* pad_local /*DEVICE [n_pad] | NULL*/, int n_pad, rbx_stream_t stream);

/* Streamed exchange (csrc/shard_stream.cu; the default data path of the row-sharded table): the all-to-all of
 * SURVEY.md section 8e done by the kernels themselves so that every random access is local to the row's owner and
 * only contiguous runs cross NVLink.  The reference has no counterpart (no sharded embedding anywhere); the contract
 * is equality with rbx_embed_fm_fwd / rbx_embed_fm_bwd on the concatenated table.  world must be a power of two
 * <= RBX_MAX_WORLD; all `void* const*` arguments are HOST arrays of `world` DEVICE pointers valid on the calling
 * device (own block first-class, peers' through rbx_peer_open / CUDA VMM); `cap` = slots per (owner, requester)
 * pair.  A step is: route, barrier(counts), serve, barrier, consume | grad_push, barrier, apply.  Inbox buffers
 * (ids, counts) are double-buffered by the caller (step parity), so apply needs no trailing barrier.
 *   rbx_xs_tile_samples  samples per tile T for (F, D) (0 = unsupported shape); tiles index tile_base / tile_cnt
 *                        ([n_tiles, RBX_MAX_WORLD] int32) and pair_sorted ([n_tiles, T*F] uint16)
 *   rbx_xs_route         requester: buckets rows[B,F] by owner (row % world) tile by tile; the ids (local rows,
 *                        row / world) of one tile for one owner form one contiguous run of that owner's
 *                        inbox_ids[w][rank*cap + slot]; cursor[w] (DEVICE int32 [RBX_MAX_WORLD], zero before the
 *                        first call) counts the slots used; overflow[0] is set if a lane would exceed cap
 *   rbx_xs_barrier       all ranks: publishes cursor[w] into meta[w][rank] (then clears it) when cursor != NULL,
 *                        writes `epoch` into every rank's flags[w][rank] (st.release.sys) and waits until its own
 *                        flags[rank][*] reached `epoch` -- a stream-ordered barrier with no host involvement
 *   rbx_xs_serve         owner: rowbuf[q][rank*cap + j] = table[inbox_ids[q*cap + j]] (row stride `row_stride`
 *                        floats), rowbuf_lr[q][rank*cap + j] = lr[id * lr_stride]
 *   rbx_xs_consume       requester: E / S / fm_out / lr_out of rbx_embed_fm_fwd from its row buffer
 *   rbx_xs_grad_push     requester: ginbox[w][rank*cap + slot] = dE + d_fm (S - e) (zero for padding rows),
 *                        ginbox_lr likewise = d_lr; the numeric-slot / bias batch reductions of rbx_embed_fm_bwd
 *                        (g_dense_w, g_dense_w_lr, g_lr_bias) ride in the same launch (Fn * D / 4 <= 256)
 *   rbx_xs_apply         owner: g_table[inbox_ids[q*cap + j]] += ginbox[q*cap + j] (red.global.add) */
int rbx_xs_tile_samples(int F, int D);
int rbx_xs_route(const int32_t* rows /*DEVICE [B,F]*/, int64_t B, int F, int64_t R, int D,
                 int rank, int world, int64_t cap,
                 int32_t* cursor, int32_t* tile_base, int32_t* tile_cnt, uint16_t* pair_sorted, int32_t* overflow,
                 void* const* inbox_ids, rbx_stream_t stream);
int rbx_xs_barrier(void* const* flags /*each DEVICE uint32 [RBX_MAX_WORLD]*/, void* const* meta /*each DEVICE int32 [RBX_MAX_WORLD] | NULL*/,
                   int32_t* cursor /*| NULL*/, int rank, int world, uint32_t epoch, rbx_stream_t stream);
int rbx_xs_serve(const float* table, int64_t row_stride, const float* lr /*| NULL*/, int64_t lr_stride, int D,
                 const int32_t* inbox_ids /*DEVICE [world,cap]*/, const int32_t* meta /*DEVICE [RBX_MAX_WORLD]*/,
                 int64_t cap, int rank, int world, void* const* rowbuf, void* const* rowbuf_lr /*| NULL*/,
                 rbx_stream_t stream);
int rbx_xs_consume(const float* rowbuf /*DEVICE [world,cap,D]*/, const float* rowbuf_lr /*DEVICE [world,cap] | NULL*/,
                   const int32_t* tile_base, const int32_t* tile_cnt, const uint16_t* pair_sorted,
                   const int32_t* cat_pos /*HOST [F]*/, const float* dense_x, const float* dense_w,
                   const float* dense_w_lr, const int32_t* num_pos /*HOST [Fn]*/, const int32_t* num_widx /*HOST | NULL*/,
                   const float* lr_bias, float* E, float* S, float* fm_out, float* lr_out,
                   int64_t B, int64_t cap, int F, int Fn, int D, int n_slots, int world, rbx_stream_t stream);
int rbx_xs_grad_push(const float* E /*| NULL: e re-read from rowbuf*/, const float* rowbuf, const float* S,
                     const float* dE, const float* d_fm, const float* d_lr,
                     const int32_t* rows /*DEVICE [B,F]*/, const int32_t* pad_row /*HOST [F] | NULL*/,
                     const int32_t* tile_base, const int32_t* tile_cnt, const uint16_t* pair_sorted,
                     const int32_t* cat_pos /*HOST [F]*/,
                     const float* dense_x, const float* dense_w, const int32_t* num_pos /*HOST [Fn]*/,
                     const int32_t* num_widx /*HOST | NULL*/, int Fn,
                     float* g_dense_w /*[Fn_total,D] | NULL*/, float* g_dense_w_lr /*| NULL*/, float* g_lr_bias /*| NULL*/,
                     int64_t B, int64_t cap, int F, int D, int n_slots,
                     int rank, int world, void* const* ginbox, void* const* ginbox_lr /*| NULL*/, rbx_stream_t stream);
int rbx_xs_apply(const float* ginbox /*DEVICE [world,cap,D]*/, const float* ginbox_lr /*| NULL*/,
                 const int32_t* inbox_ids, const int32_t* meta, int64_t cap, int world,
                 float* g_table, int64_t row_stride, float* g_lr /*| NULL*/, int64_t lr_stride, int D,
                 rbx_stream_t stream);

/* Peer-visible device memory (CUDA IPC; one process per GPU, all on one box).
 * rbx_peer_alloc  : cudaMalloc'd block (IPC-exportable, unlike a caching-allocator sub-block)
 * rbx_peer_export : 64-byte handle the owner sends to its peers (any byte transport)
 * rbx_peer_open   : map a peer's block into this process / device; enables peer access lazily
 * rbx_peer_close / rbx_peer_free : unmap / release
 * rbx_peer_can_access : 1 if `device` can address `peer_device` memory directly (NVLink / PCIe P2P) */
int rbx_peer_alloc(size_t bytes, void** ptr /*HOST out*/);
int rbx_peer_free(void* ptr);
int rbx_peer_export(const void* ptr, unsigned char handle[RBX_PEER_HANDLE_BYTES] /*HOST out*/);
int rbx_peer_open(const unsigned char handle[RBX_PEER_HANDLE_BYTES], void** ptr /*HOST out*/);
int rbx_peer_close(void* ptr);
int rbx_peer_can_access(int device, int peer_device);

/* ------------------------------------------------------------------------------------------
 * a12  clip + optimizer on the fused table  (RankingModel.train_step ranking_model.py:191-197:
 * clip_grad_norm_(all params, max_norm) then torch.optim.Adam.step, dense over every row)
 * rbx_sqnorm     : out[0] += sum g[i]^2    (float64 accumulator, the global-norm partial)
 * rbx_clip_coef  : coef[0] = min(1, max_norm / (sqrt(sqnorm[0]) + 1e-6)); norm_out[0] = sqrt(sqnorm[0])
 * rbx_adam_dense : torch.optim.Adam single-tensor update in its operation order, with the
 *                  clip coefficient read from DEVICE memory (no host sync):
 *                  g' = g * clip[0];  m = lerp(m, g', 1-b1);  v = b2 v + (1-b2) g'^2;
 *                  w -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
 * ------------------------------------------------------------------------------------------ */
int rbx_sqnorm(const float* g, int64_t n, double* out /*DEVICE [1]*/, rbx_stream_t stream);
int rbx_clip_coef(const double* sqnorm /*DEVICE [1]*/, float max_norm, float* coef /*DEVICE [1]*/,
                  float* norm_out /*DEVICE [1] | NULL*/, rbx_stream_t stream);
int rbx_adam_dense(float* w, const float* g, float* m, float* v, int64_t n,
                   const float* clip /*DEVICE [1] | NULL*/,
                   float lr, float beta1, float beta2, float eps, int step,
                   rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f1  touched-rows clip + optimizer (same reference lines as a12).
 * `rows` = sorted-unique table rows the batch touched (rbx_unique_ids_i32 over its global row ids),
 * n_rows_dev = its length in DEVICE memory (NULL: use max_rows), max_rows bounds the launch.
 * rbx_sqnorm_rows : out[0] += sum over touched rows |g[row,:]|^2  (equals the dense sum)
 * rbx_optim_rows  : kind 0 SGD (exact), 1 Adagrad (exact; v = state sum), 2 Adam -- the dense formula
 *                   of rbx_adam_dense on touched rows only ("lazy": untouched rows keep their moments,
 *                   which is NOT the reference's dense Adam), 3 torch.optim.SparseAdam's formula.
 *                   g' = g * clip[0].  zero_grad != 0 clears the consumed gradient rows, so the dense
 *                   gradient table stays all-zero between steps with no O(table) memset.
 * ------------------------------------------------------------------------------------------ */
int rbx_sqnorm_rows(const float* g /*DEVICE [R,D]*/, const int32_t* rows /*DEVICE [max_rows]*/,
                    const int64_t* n_rows_dev /*DEVICE [1] | NULL*/, int64_t max_rows, int D,
                    double* out /*DEVICE [1]*/, rbx_stream_t stream);
int rbx_optim_rows(float* w, float* g, float* m /*| NULL*/, float* v /*| NULL*/,
                   const int32_t* rows, const int64_t* n_rows_dev, int64_t max_rows, int D,
                   const float* clip /*DEVICE [1] | NULL*/, int kind,
                   float lr, float beta1, float beta2, float eps, int step, int zero_grad,
                   rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a14  sorted unique + inverse + first occurrence of a bag of ids bounded by a vocabulary
 * (collate_fn_unique, recbox/matching/pytorch/dataloaders/h5_generator.py:45-53:
 *  torch.unique(item_indexes.flatten(), return_inverse=True, sorted=True) + the flip/scatter_
 *  "return_index").  No sort: a vocab-bit bitmap (byte map for dense batches) + popcount prefix (csrc/dedup.cu).
 *   uniq[0..U)   ascending distinct ids            (capacity min(n, vocab))
 *   first[u]     smallest flat position i with ids[i] == uniq[u]   (int64; NULL to skip)
 *   inverse[i]   u with uniq[u] == ids[i]; -1 for an id outside [0, vocab)   (NULL to skip)
 *   n_out[0] = U, n_out[1] = number of out-of-range ids          (DEVICE int64[2])
 * ws = rbx_unique_ws_bytes(vocab) bytes of 16-byte aligned DEVICE scratch.  Also used to list the
 * table rows a batch touched (vocab = R) for the f1 optimizer.
 * ------------------------------------------------------------------------------------------ */
size_t rbx_unique_ws_bytes(int64_t vocab);
int rbx_unique_ids_i64(const int64_t* ids /*DEVICE [n]*/, int64_t n, int64_t vocab,
                       void* ws /*DEVICE*/, size_t ws_bytes,
                       int64_t* uniq, int64_t* first, int64_t* inverse, int64_t* n_out,
                       rbx_stream_t stream);
int rbx_unique_ids_i32(const int32_t* ids /*DEVICE [n]*/, int64_t n, int64_t vocab,
                       void* ws /*DEVICE*/, size_t ws_bytes,
                       int32_t* uniq, int64_t* first, int32_t* inverse, int64_t* n_out,
                       rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f4  power sums over the fields: P[b,k-1,:] = sum_f E[b,f,:]^k, k = 1..order (order <= 5)
 * The single pass over [B,F,D] behind InteractionMachine.forward
 * (ranking/pytorch/layers/interactions/interaction_machine.py:44-70: p_k = (Q *= X).sum(dim=1), `order` passes
 * and `order` temporaries in the reference); the polynomial combinations of p_1..p_order (:29-42) act on [B,D].
 * Backward: dE[b,f,:] = sum_k k * E[b,f,:]^(k-1) * dP[b,k-1,:].
 * ------------------------------------------------------------------------------------------ */
int rbx_power_sums_fwd(const float* E /*DEVICE [B,F,D]*/, float* P /*DEVICE [B,order,D]*/,
                       int64_t B, int F, int D, int order, rbx_stream_t stream);
int rbx_power_sums_bwd(const float* E, const float* dP /*DEVICE [B,order,D]*/, float* dE /*DEVICE [B,F,D]*/,
                       int64_t B, int F, int D, int order, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f2  per-epoch negative sampling (csrc/sample.cu)
 * Replaces sampling_block + the hstack of TrainGenerator.negative_sampling
 * (recbox/matching/pytorch/dataloaders/h5_generator.py:72-95, 144-181): uniform draws with
 * replacement over [0, num_items); with pos_ptr/pos_items (CSR of each user's interacted items,
 * sorted ascending per user) draws that hit one of the query user's items are redrawn
 * (= ignore_pos_items=True: probabilities of those items zeroed and renormalised).
 *   out [n_queries, (pos ? 1 : 0) + num_negs] int64: pos[q] first when given, then the negatives
 *   user_of_query [n_queries] | NULL (= identity): row of the CSR for query q
 *   gave_up DEVICE int32[1]: += number of elements still colliding after 4096 redraws
 * Counter-based RNG keyed by (seed, element index): reproducible, geometry-independent; not numpy's
 * MT19937 stream (parity is distributional).
 * ------------------------------------------------------------------------------------------ */
int rbx_sample_negatives(int64_t n_queries, int num_negs, int64_t num_items, uint64_t seed,
                         const int64_t* pos /*DEVICE [n_queries] | NULL*/,
                         const int64_t* user_of_query /*DEVICE | NULL*/,
                         const int64_t* pos_ptr /*DEVICE [n_users+1] | NULL*/,
                         const int64_t* pos_items /*DEVICE | NULL*/,
                         int64_t* out /*DEVICE*/, int* gave_up /*DEVICE [1] | NULL*/,
                         rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a13  dense MLP tail (csrc/gemm.cu): fp32 GEMM on the tcgen05 tensor cores, fused epilogue
 * Replaces nn.Linear + activation of MLP_Block (recbox/ranking/pytorch/layers/blocks/mlp_block.py:43-61;
 * core/pytorch/layers/mlp.py) and their autograd backward inside loss.backward()
 * (recbox/ranking/pytorch/models/ranking_model.py:191-197):
 *     C[M,N] (+)= relu?( A[M,K] * B[N,K]^T + bias[N] ) * (mask[M,N] > 0)?
 * A is logically [M,K], B logically [N,K]; a_mn / b_mn say how they lie in memory:
 *     a_mn = 0: A stored row-major [M,K] (K contiguous, pitch lda)    a_mn = 1: stored [K,M] (M contiguous)
 *     b_mn = 0: B stored row-major [N,K] (nn.Linear.weight)           b_mn = 1: stored [K,N] (N contiguous)
 *   forward  H' = relu(H W^T + b):  A = H, B = W,            a_mn = 0, b_mn = 0, bias, act = 1
 *   backward dH = (dZ W) * (H > 0): A = dZ, B = W as [K=N_lin, N=K_lin],  b_mn = 1, mask = H
 *            dW = dZ^T H:           A = dZ as [K=batch, M=N_lin], B = H as [K=batch, N=K_lin], a_mn = b_mn = 1
 * precision 3: every operand split into two TF32 words in shared memory, three MMAs per k-step, fp32 accumulation in
 * TMEM -> fp32-level results (the 1e-5 contract); precision 1: plain TF32 (~1e-3 rel).  act: 0 none, 1 ReLU.
 * accumulate != 0 adds into C.  Outputs with few tiles and a long K are split over K (red.add into C; not with act / mask).
 * lda, ldb multiples of 4 floats and A, B 16-byte aligned on the tensor-core path; N <= 8, K <= 8 and M <= 8 shapes
 * (the Linear(hidden, 1) head) run memory-bound SIMT kernels.
 * rbx_colsum_f32: out[n] (+)= sum_m X[m,n]  (bias gradient).
 * ------------------------------------------------------------------------------------------ */
int rbx_gemm_f32(const float* A /*DEVICE*/, int64_t lda, int a_mn, const float* B /*DEVICE*/, int64_t ldb, int b_mn,
                 float* C /*DEVICE [M,N] pitch ldc*/, int64_t ldc, int64_t M, int64_t N, int64_t K,
                 const float* bias /*DEVICE [N] | NULL*/, int act, const float* mask /*DEVICE [M,N] pitch ldmask | NULL*/,
                 int64_t ldmask, int precision, int accumulate, rbx_stream_t stream);
int rbx_colsum_f32(const float* X /*DEVICE [M,N] pitch ldx*/, int64_t ldx, float* out /*DEVICE [N]*/, int64_t M, int64_t N,
                   int accumulate, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * e  replicas: in-switch all-reduce of the fused gradient buffer (csrc/allreduce.cu)
 * The reference's multi-device mode is DistributedDataParallel over replicas (third_party/recbole/trainer/trainer.py:48-64,
 * third_party/rechub/trainers/ctr_trainer.py:43): the dense gradients are all-reduced once per step.  mc is the MULTICAST
 * address of a symmetric buffer of n floats (n % 4 == 0) that every rank of the NVSwitch domain has mapped; rank r sums
 * floats [r n / world, (r+1) n / world) across all copies with multimem.ld_reduce and writes the sums back to all copies
 * with multimem.st.  The caller brackets the call with cross-rank barriers.
 * ------------------------------------------------------------------------------------------ */
int rbx_nvls_allreduce_f32(float* mc /*DEVICE multicast address*/, int64_t n, int rank, int world, rbx_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f3  retrieval evaluation (csrc/topk.cu)
 * rbx_topk_ip replaces FaissIndex.search = faiss.IndexFlatIP(dim).search(query, topk)
 * (recbox/utils/ann/faiss.py:3-14; called from evaluate_block, recbox/core/metrics.py:52-54):
 * exact inner-product top-k of every query row against the whole corpus, descending, ties to the
 * smaller index; rows of a corpus smaller than k are padded with (-inf, -1).  fp32-level accuracy
 * (3xTF32 split products on the tensor cores, fp32 accumulation; RBX_TOPK_MMA=0 builds plain fp32
 * FMA chains).  D a multiple of 4 in [4,128]; k <= 1024; q, items 16-byte aligned.
 *   chunk = items per pass (rounded up to 128); ws = rbx_topk_ws_bytes(U, k, chunk) bytes of
 *   256-byte aligned DEVICE scratch (U * chunk * 8 B candidate queue + the running lists).
 * rbx_rank_metrics replaces the rest of evaluate_block (core/metrics.py:55-68) and the metric
 * classes (core/metrics.py:71-200): candidates the user clicked in train sink behind all others
 * (scores += -1e9 * mask; argsort), the first kmax are `ranked`, hit[u,r] = ranked[u,r] is a valid
 * item of u, and out[u,m] = metric kinds[m] at cut-off ks[m]:
 *   0 Recall 1 nRecall 2 Precision 3 F1 4 DCG 5 NDCG 6 MRR 7 HitRate 8 MAP   (float64, as Python)
 * train_* / valid_*: CSR over the U query rows, items sorted ascending per row (duplicates kept:
 * len(true_items) counts them, core/metrics.py:78).
 * ------------------------------------------------------------------------------------------ */
size_t rbx_topk_ws_bytes(int64_t U, int k, int64_t chunk);
int rbx_topk_ip(const float* q /*DEVICE [U,D]*/, const float* items /*DEVICE [N,D]*/,
                int64_t U, int64_t N, int D, int k, int64_t chunk,
                float* out_scores /*DEVICE [U,k] | NULL*/, int64_t* out_idx /*DEVICE [U,k] | NULL*/,
                void* ws /*DEVICE*/, size_t ws_bytes, rbx_stream_t stream);
int rbx_rank_metrics(const int64_t* cand /*DEVICE [U,T] top-T item ids, descending*/, int T, int64_t U,
                     const int64_t* train_ptr /*DEVICE [U+1] | NULL*/, const int64_t* train_items /*DEVICE | NULL*/,
                     const int64_t* valid_ptr /*DEVICE [U+1]*/, const int64_t* valid_items /*DEVICE*/,
                     int kmax, const int* kinds /*DEVICE [M]*/, const int* ks /*DEVICE [M]*/, int M,
                     int64_t* ranked /*DEVICE [U,kmax]*/, uint8_t* hit /*DEVICE [U,kmax]*/,
                     double* out /*DEVICE [U,M] | NULL when M == 0*/, rbx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RECBOX_B200_H */
