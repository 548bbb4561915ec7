/*
 * nasr_ctc.h — C-ABI of libnasr_ctc.so: the B200-native CTC training-loss / greedy-decode /
 * label-error-rate path behind NeuralASR's loss(), decoding and label_error_rate.
 *
 * The reference has no FFI of its own on this path: its three helpers are Python methods that
 * forward to TensorFlow 1.x CPU ops,
 *     create_loss   networks/tfnetwork.py:58-59  tf.reduce_mean(tf.nn.ctc_loss(labels, logits, seq_len))
 *     create_model  networks/tfnetwork.py:61-64  tf.nn.ctc_greedy_decoder(logits, seq_len) (line 63)
 *     create_metric networks/tfnetwork.py:66-70  tf.reduce_mean(tf.edit_distance(cast(model), labels))
 * so the entry points below are what a ctypes binding placed behind those three methods binds to
 * (INTEGRATION.md shows the stub).  Each entry cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns an int status: NASR_OK or a NASR_ERR_* code; nasr_last_error() gives the
 *     thread-local message of the last failure on the calling thread;
 *   - all tensor pointers are DEVICE pointers unless the name says host; the caller owns every buffer
 *     (outputs and workspace included); the library never allocates, frees or retains them, except
 *     inside an explicit nasr_host_ctx (the HOST-buffer convenience path);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream)
 *     and the call returns without synchronising, so it can be stream-captured into a CUDA graph;
 *   - logits are float32, time-major [T, B, C], contiguous (what every model tail produces after its
 *     transpose: networks/bilstm_ctc_net.py:48, lstm_ctc_net.py:43, deepspeech.py:126-127,
 *     wavenet.py:171); blank is C-1 in the reference (preprocess_mfcc.py:81-92) and is passed
 *     explicitly here;
 *   - labels are CSR: label_values int32[N] (the `values` of the sparse triple utils.py:44-58 builds)
 *     and label_offsets int32[B+1] (prefix sum of the rows of its `indices`);
 *   - TF's InvalidArgument exceptions become per-utterance bit flags in status[B] (NASR_ST_*),
 *     so a bad utterance does not abort the batch and no host sync is forced.
 */
#ifndef NASR_CTC_H_
#define NASR_CTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NASR_ABI_VERSION 1

/* return codes */
#define NASR_OK 0
#define NASR_ERR_INVALID_ARGUMENT 1
#define NASR_ERR_WORKSPACE_TOO_SMALL 2
#define NASR_ERR_CUDA 3
#define NASR_ERR_UNSUPPORTED 4

/* per-utterance status bits (status[b]); 0 = OK.  TF raises InvalidArgumentError for the first three. */
#define NASR_ST_LABEL_OUT_OF_RANGE 1   /* "Saw a non-null label (index >= num_classes - 1)..." */
#define NASR_ST_SEQ_LEN_OUT_OF_RANGE 2 /* sequence_length(b) > max_time (or negative)          */
#define NASR_ST_NOT_ENOUGH_TIME 4      /* "Not enough time for target transition sequence"     */
#define NASR_ST_NO_VALID_PATH 8        /* log p = -inf: loss = +inf, gradient = softmax        */

/* Minimal DLPack (v0.8 ABI) declarations so DLPack-capsule callers need no extra header.  Layout is
 * identical to dlpack.h; include that header first if you have it. */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType; /* code: 0 int, 1 uint, 2 float */
typedef struct {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides; /* NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

int nasr_abi_version(void);

/* Message for the last non-OK return on this thread ("" if none). */
const char* nasr_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * CTC loss + gradient.  Replaces tf.nn.ctc_loss + _CTCLossGrad behind create_loss
 * (networks/tfnetwork.py:58-59) with the defaults that call uses: time-major, ctc_merge_repeated=True,
 * preprocess_collapse_repeated=False, ignore_longer_outputs_than_inputs=False.
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of device workspace nasr_ctc_loss_grad_f32 needs for these shapes
 * (max_label_len = longest transcript in the batch; the extended state count is 2*max_label_len+1). */
int nasr_ctc_workspace_bytes(int T, int B, int C, int max_label_len, size_t* out_bytes);

/*
 * loss[b]      = -log p(labels_b | softmax(logits[:seq_len[b], b, :]))          float32[B]
 * grad[t,b,c]  = grad_loss[b] * (softmax(logits)[t,b,c] - occupancy[t,b,c])     float32[T,B,C]
 *                exactly 0 for t >= seq_len[b]; grad may be NULL (loss only);
 *                grad_loss may be NULL (= 1 for every b).  Under the reference's reduce_mean
 *                (tfnetwork.py:59) the upstream gradient is 1/B.
 * status[b]    = NASR_ST_* flags                                                int32[B]
 * max_label_len must be >= every row length of the CSR labels (it sizes shared memory and workspace).
 */
int nasr_ctc_loss_grad_f32(const float* logits, int T, int B, int C,
                           const int32_t* label_values, const int32_t* label_offsets,
                           int max_label_len, const int32_t* seq_len, int blank,
                           float* loss, float* grad, const float* grad_loss, int32_t* status,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Same call for logits / grad whose frames and utterances are strided: element (t, b, c) lives at
 * logits[t*stride_t + b*stride_b + c] (grad uses the same strides).  stride_t = B*C, stride_b = C is the call
 * above; stride_t = C, stride_b = T*C reads a batch-major [B, T, C] tensor in place, which removes the
 * transpose every model tail of the reference performs before create_loss (networks/bilstm_ctc_net.py:48,
 * lstm_ctc_net.py:43, wavenet.py:171); a block of utterances b0.. of a [T, B, C] tensor is the pointer
 * logits + b0*C with the parent's strides. */
int nasr_ctc_loss_grad_strided_f32(const float* logits, int T, int B, int C, long long stride_t,
                                   long long stride_b, const int32_t* label_values,
                                   const int32_t* label_offsets, int max_label_len, const int32_t* seq_len,
                                   int blank, float* loss, float* grad, const float* grad_loss,
                                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with the tensors passed as DLPack tensors (the north-star "ctypes + DLPack buffers" form):
 * device, dtype, rank and layout of every argument are validated here instead of in the caller.
 * logits (and grad, with identical strides) may be any [T, B, C] view whose innermost stride is 1, e.g. the
 * transposed view of a batch-major tensor; everything else must be compact.  grad and grad_loss may be NULL. */
int nasr_ctc_loss_grad_dl(const DLTensor* logits, const DLTensor* label_values,
                          const DLTensor* label_offsets, int max_label_len, const DLTensor* seq_len,
                          int blank, const DLTensor* loss, const DLTensor* grad,
                          const DLTensor* grad_loss, const DLTensor* status,
                          const DLTensor* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Greedy decode.  Replaces tf.nn.ctc_greedy_decoder(logits, seq_len, merge_repeated) — the `decoding`
 * of the north star, networks/tfnetwork.py:63.
 *   hyp            int64[B, T]  row b holds hyp_len[b] label ids (first-index argmax of the RAW logits
 *                               per frame, blank dropped, repeats merged); the rest of the row is undefined
 *                               (wide rows use it as scratch)
 *   hyp_len        int32[B]
 *   neg_sum_logits float32[B]   -(sum over frames of the max logit)   (TF's second output, [B,1])
 * Frames t >= seq_len[b] are ignored; seq_len is clamped to [0, T].
 * ---------------------------------------------------------------------------------------------- */
int nasr_ctc_greedy_decode_i64(const float* logits, int T, int B, int C, const int32_t* seq_len,
                               int blank, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                               float* neg_sum_logits, void* stream);

/* Same for strided logits (see nasr_ctc_loss_grad_strided_f32). */
int nasr_ctc_greedy_decode_strided_i64(const float* logits, int T, int B, int C, long long stride_t,
                                       long long stride_b, const int32_t* seq_len, int blank,
                                       int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                                       float* neg_sum_logits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Beam search decode.  Replaces tf.nn.ctc_beam_search_decoder(logits, seq_len) with its defaults
 * (beam_width=100, top_paths=1, merge_repeated=True) — what create_model runs today,
 * networks/tfnetwork.py:62,64.
 *   hyp       int64[B, top_paths, T]  row (b, p) holds hyp_len[b, p] label ids of the p-th best prefix (repeats
 *                                     collapsed in the output when merge_repeated, like TF)
 *   hyp_len   int32[B, top_paths]     0 for paths beyond the number of prefixes the beam holds
 *   log_prob  float32[B, top_paths]   log P(prefix | logits) under the log-softmax of the logits (-inf when absent)
 * Scores are computed in float64; ties are broken by (kept prefix before new extension, prefix hash).
 * workspace: nasr_ctc_beam_workspace_bytes(T, B, C, beam_width) bytes (the per-utterance prefix trie).
 * Supported: beam_width <= 512, C <= 8192, beam_width * C < 2^20 while C <= 4095, and a shared-memory footprint
 * under 200 KB (beam_width 100: every C <= 8192); otherwise NASR_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------------- */
int nasr_ctc_beam_workspace_bytes(int T, int B, int C, int beam_width, size_t* out_bytes);

int nasr_ctc_beam_search_i64(const float* logits, int T, int B, int C, const int32_t* seq_len, int blank,
                             int beam_width, int top_paths, int merge_repeated, int64_t* hyp,
                             int32_t* hyp_len, float* log_prob, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Same for strided logits (see nasr_ctc_loss_grad_strided_f32). */
int nasr_ctc_beam_search_strided_i64(const float* logits, int T, int B, int C, long long stride_t,
                                     long long stride_b, const int32_t* seq_len, int blank, int beam_width,
                                     int top_paths, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                                     float* log_prob, void* workspace, size_t workspace_bytes, void* stream);

/* Device-side label ingest: the COO triple sparse_tuple_from builds (reference utils.py:44-58: indices int64 [N,2]
 * row-major with rows in order, values int32 [N]) -> the CSR the loss / edit-distance entries take, without touching
 * the host.  Rows [row0, row0 + B) are taken and re-based: a tower's block under tf.sparse_split(axis=0)
 * (tfnetwork.py:97-99) is a row window.  offsets int32 [B+1]; values_out int32 [>= entries of the window] or NULL
 * (with row0 = 0 the input values ARE the CSR values); info int32 [3] on the device: [0] = 1 if the rows are not in
 * order or a position is not the running index of its row (TF would reject the SparseTensor), [1] = longest row,
 * [2] = first entry of the window.  max_label_len of the loss call can be the triple's dense_shape[1], known on the
 * host without a sync. */
int nasr_labels_coo_to_csr_i32(const int64_t* indices, const int32_t* values, int N, int row0, int B,
                               int32_t* offsets, int32_t* values_out, int32_t* info, void* stream);

/* Dense hypotheses -> the SparseTensor triple TF returns as decoded[0] (tfnetwork.py:64):
 * hyp_offsets int32[B+1] must hold the exclusive prefix sum of hyp_len (M = hyp_offsets[B]);
 * indices int64[M,2] row-major (b, position), values int64[M], dense_shape int64[2] = [B, max len]. */
int nasr_hyp_to_sparse_i64(const int64_t* hyp, int hyp_stride, const int32_t* hyp_offsets, int B,
                           int64_t* indices, int64_t* values, int64_t* dense_shape, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Label error rate.  Replaces tf.edit_distance(tf.cast(model, tf.int32), labels, normalize) behind
 * create_metric (networks/tfnetwork.py:66-70).  Hypotheses are the dense rows greedy decode wrote
 * (row stride hyp_stride elements); truth is CSR.
 *   dist[b] = Levenshtein(truth_b, hyp_b)                       int32[B]   (bit-exact integers)
 *   ler[b]  = dist / |truth_b| if normalize else dist           float32[B]
 *             (|truth_b| = 0: +inf if dist != 0 else 0)
 * max_truth_len >= every truth row length (sizes shared memory).
 * ---------------------------------------------------------------------------------------------- */
int nasr_edit_distance_i64(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len,
                           const int32_t* truth_values, const int32_t* truth_offsets,
                           int max_truth_len, int B, int normalize, int32_t* dist, float* ler,
                           void* stream);

/* Same, hypotheses given as CSR int64 values + int32 offsets (a decoded SparseTensor's values and the
 * prefix sum of its row counts); max_hyp_len >= every hypothesis row length. */
int nasr_edit_distance_csr_i64(const int64_t* hyp_values, const int32_t* hyp_offsets,
                               int max_hyp_len, const int32_t* truth_values, const int32_t* truth_offsets,
                               int max_truth_len, int B, int normalize, int32_t* dist, float* ler,
                               void* stream);

/* Batch reductions the reference wraps around the ops (tf.reduce_mean, tfnetwork.py:59,69) and the
 * tower means of tfnetwork.py:135-136: sums[0] = sum loss, sums[1] = sum ler, sums[2] = sum dist,
 * sums[3] = B, as float64[4] on the device (the 4-element vector each rank all-reduces). loss/ler/dist
 * may be NULL (that slot is 0). Deterministic (single-block tree in fixed order). */
int nasr_batch_sums_f64(const float* loss, const float* ler, const int32_t* dist, int B,
                        double* sums, void* stream);

/* The one collective of the path: the tower means of tfnetwork.py:135-136 are an all-reduce(sum) of the float64
 * vector nasr_batch_sums_f64 builds (n = 4), 32 bytes over NCCL / NVLink.  `nccl_comm` is the caller's ncclComm_t;
 * the library resolves ncclAllReduce from the NCCL the process has loaded (libnccl.so.2), it does not link NCCL.
 * Asynchronous on `stream`, in place.  A caller that lives in torch.distributed uses towers.all_reduce_sums instead. */
int nasr_allreduce_scalars(void* nccl_comm, double* vec, int n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HOST-buffer path: what a caller that holds numpy arrays (the reference's feed_dict world,
 * tfnetwork.py:183-190) uses.  The context owns device buffers, pinned staging and one stream, sized
 * for the maxima given at creation.  One call = H2D of logits/labels/seq_len, loss+grad, greedy decode,
 * edit distance, D2H of the results, and a stream synchronise.  Large batches are processed in eight blocks
 * of utterances on two compute streams (a workspace each) so that the H2D copy of one block, the kernels of the
 * previous ones and the D2H copy of the ones before overlap (environment NASR_HOST_BLOCKS / NASR_HOST_STREAMS,
 * read at context creation, change the split: measured 1.78 ms unsplit, 1.30 ms 4x1, 1.20 ms 8x2 at B=256, T=1000,
 * C=38).
 * ---------------------------------------------------------------------------------------------- */
typedef struct nasr_host_ctx nasr_host_ctx;

int nasr_host_ctx_create(int device, int max_T, int max_B, int max_C, int max_label_len,
                         nasr_host_ctx** out);
void nasr_host_ctx_destroy(nasr_host_ctx* ctx);
/* Which decoder nasr_host_ctc_step runs: 0 = greedy (tfnetwork.py:63, the default), 1 = beam search of width
 * beam_width, top path, merge_repeated (tfnetwork.py:62, what train()/evaluate() fetch in the snapshot); with the beam
 * decoder the step's neg_sum_logits output carries the top path's log probability.  Allocates the beam workspace. */
int nasr_host_ctx_set_decoder(nasr_host_ctx* ctx, int decoder, int beam_width);
/* Pinned host staging the caller may fill directly to skip one host copy: logits float[max_T*max_B*max_C],
 * grad float[same]. */
float* nasr_host_ctx_pinned_logits(nasr_host_ctx* ctx);
float* nasr_host_ctx_pinned_grad(nasr_host_ctx* ctx);

/* All pointers are HOST pointers. logits/grad may be the ctx's pinned buffers (then no extra host copy is
 * made). grad, hyp, hyp_len, neg_sum_logits, dist, ler may each be NULL to skip that output
 * (decode runs only if hyp_len/dist/ler is wanted). */
int nasr_host_ctc_step(nasr_host_ctx* ctx, const float* logits, int T, int B, int C,
                       const int32_t* label_values, const int32_t* label_offsets,
                       const int32_t* seq_len, int blank, const float* grad_loss,
                       float* loss, float* grad, int32_t* status,
                       int64_t* hyp /*[B,T]*/, int32_t* hyp_len, float* neg_sum_logits,
                       int32_t* dist, float* ler);

/* ------------------------------------------------------------------------------------------------
 * The model tails' affine projection.  Replaces
 *     outputs = tf.reshape(outputs, [-1, num_hidden]); logits = tf.matmul(outputs, W) + b
 * of networks/bilstm_ctc_net.py:33-45 and lstm_ctc_net.py:28-40 (num_hidden = 500, W [K, C], b [C]) and, for
 * training, the three gradients TensorFlow derives from it.  float32 in and out; the products run on the tensor
 * cores as 3xTF32 (hi*hi + hi*lo + lo*hi, float32 accumulation), i.e. within float32 rounding of a float32 matmul.
 *   H       float32 [rows, K], row r at H + r*ldh   (rows = B*T in the callers' batch-major order)
 *   logits  float32 [rows, C], row r at logits + r*ldl; viewed as [B, T, C] it is the tensor the reference then
 *           transposes -- nasr_ctc_loss_grad_strided_f32 (stride_t = C, stride_b = T*C) reads it in place.
 * bias may be NULL.  Not fused into the CTC kernel on purpose: H is 13 times the logits in bytes and the loss reads
 * every frame twice, so a fused producer would move more bytes than writing the logits once.
 * ---------------------------------------------------------------------------------------------- */
int nasr_affine_logits_f32(const float* H, long long rows, int K, long long ldh, const float* W,
                           const float* bias, int C, float* logits, long long ldl, void* stream);

/* Bytes of device workspace nasr_affine_backward_f32 needs for dW / db (per-CTA partial sums, added in a fixed
 * order: the result does not depend on scheduling). */
int nasr_affine_workspace_bytes(long long rows, int K, int C, size_t* out_bytes);

/* dH[r, k] = sum_c dlogits[r, c] * W[k, c];  dW[k, c] = sum_r H[r, k] * dlogits[r, c];  db[c] = sum_r dlogits[r, c].
 * Each of dH, dW, db may be NULL to skip it (H may be NULL when only dH is wanted, the workspace when neither dW
 * nor db is). */
int nasr_affine_backward_f32(const float* H, long long rows, int K, long long ldh, const float* W, int C,
                             const float* dlogits, long long ldd, float* dH, long long lddh, float* dW,
                             float* db, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernel launches this library has enqueued since load (for bench.py's gpu_launches). */
uint64_t nasr_launch_count(void);

/* Test hook (process-wide, not thread-safe; production code never calls it).
 * nasr_ctc_loss_grad_* normally runs the throughput kernel and then redoes, with the robust kernel, every
 * utterance the first one flagged (the flags are the first B int32 of the workspace, 256-byte aligned).
 *   path 0: that default;  path 1: robust kernel for everything;  path 2: throughput kernel only —
 *           flagged utterances are left unwritten so a test can see which ones it would have handed over.
 *   split_frames: 0 = the throughput kernel splits each utterance in the middle; otherwise the number of
 *           frames its forward half covers (rounded to the chunk size), to exercise uneven splits.
 * Change it only between calls, and query nasr_ctc_workspace_bytes again afterwards. */
int nasr_debug_config(int path, int split_frames);

/* Tuning hook, active only in a library built with -DNASR_TUNING=1 (NASR_TUNING=1 python -m neuralasr_b200._build;
 * production builds compile it out: it costs the throughput kernel 14 %): when device_buffer is non-NULL the throughput kernel writes, per utterance and warp,
 * int64[4] = {cycles of work before the meeting, cycles of work after it, total cycles, warp role} to
 * device_buffer[(b*16 + warp)*4 ...] (B*16*4 int64, then a per-iteration trace of the first four utterances:
 * 4*200*8*2 int64).  NULL switches it off (the default). */
int nasr_debug_profile(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* NASR_CTC_H_ */
