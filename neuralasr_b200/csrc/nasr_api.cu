// extern "C" surface of libnasr_ctc.so (declared in include/nasr_ctc.h): argument checking,
// DLPack unwrapping, the HOST-buffer context, and dispatch into the kernels of ctc_loss.cu /
// ctc_decode.cu.  No torch types cross this boundary.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <thread>

#include <dlfcn.h>

#include "nasr_common.cuh"

namespace nasr {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

// implemented in ctc_loss.cu / ctc_decode.cu
int ctc_workspace_bytes(int T, int B, int C, int Lmax, size_t* out);
int ctc_loss_grad(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                  const int32_t* label_values, const int32_t* label_offsets, int Lmax, const int32_t* seq_len,
                  int blank, float* loss, float* grad, const float* grad_loss, int32_t* status, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
int greedy_decode(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                  const int32_t* seq_len, int blank, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                  float* neg_sum_logits, cudaStream_t stream);
int edit_distance_dense(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len,
                        const int32_t* truth_values, const int32_t* truth_offsets, int max_truth_len,
                        int B, int normalize, int32_t* dist, float* ler, cudaStream_t stream);
int edit_distance_csr(const int64_t* hyp_values, const int32_t* hyp_offsets, int max_hyp_len,
                      const int32_t* truth_values, const int32_t* truth_offsets, int max_truth_len,
                      int B, int normalize, int32_t* dist, float* ler, cudaStream_t stream);
int labels_coo_to_csr(const int64_t* indices, const int32_t* values, int N, int row0, int B, int32_t* offsets,
                      int32_t* values_out, int32_t* info, cudaStream_t stream);
int hyp_to_sparse(const int64_t* hyp, int hyp_stride, const int32_t* hyp_offsets, int B,
                  int64_t* indices, int64_t* values, int64_t* dense_shape, cudaStream_t stream);
int batch_sums(const float* loss, const float* ler, const int32_t* dist, int B, double* sums,
               cudaStream_t stream);
// ctc_beam.cu
int ctc_beam_workspace_bytes(int T, int B, int C, int W, size_t* out);
int ctc_beam_search(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                    const int32_t* seq_len, int blank, int W, int P, int merge_repeated, int64_t* hyp,
                    int32_t* hyp_len, float* log_prob, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream);
extern int g_debug_path;   // ctc_loss.cu
extern int g_debug_split;  // ctc_fast.cu
extern long long* g_debug_prof;
extern int g_debug_ablate;

namespace {

// DLPack validation: CUDA device, expected dtype, expected rank, compact row-major.
bool dl_ok(const DLTensor* t, const char* name, int code, int bits, int ndim, int device_id,
           bool check_contiguous = true) {
  if (!t) {
    set_error("%s: NULL DLTensor", name);
    return false;
  }
  if (t->device.device_type != kDLCUDA) {
    set_error("%s: not a CUDA tensor (device_type=%d)", name, (int)t->device.device_type);
    return false;
  }
  if (device_id >= 0 && t->device.device_id != device_id) {
    set_error("%s: on cuda:%d, expected cuda:%d", name, t->device.device_id, device_id);
    return false;
  }
  if (t->dtype.code != code || t->dtype.bits != bits || t->dtype.lanes != 1) {
    set_error("%s: dtype (code=%d,bits=%d) but expected (code=%d,bits=%d)", name, t->dtype.code,
              t->dtype.bits, code, bits);
    return false;
  }
  if (ndim >= 0 && t->ndim != ndim) {
    set_error("%s: rank %d, expected %d", name, t->ndim, ndim);
    return false;
  }
  if (t->strides && check_contiguous) {
    int64_t expect = 1;
    for (int i = t->ndim - 1; i >= 0; i--) {
      if (t->shape[i] != 1 && t->strides[i] != expect) {
        set_error("%s: not contiguous (stride[%d]=%lld, expected %lld)", name, i,
                  (long long)t->strides[i], (long long)expect);
        return false;
      }
      expect *= t->shape[i];
    }
  }
  return true;
}

// [T,B,C] float32 CUDA tensor whose innermost dimension is dense; frames and utterances may be strided
// (e.g. the transposed view of a batch-major [B,T,C] tensor).  Returns the element strides.
bool dl_logits_ok(const DLTensor* t, const char* name, int device_id, long long* st_t, long long* st_b) {
  if (!dl_ok(t, name, 2, 32, 3, device_id, /*check_contiguous=*/false)) return false;
  const int64_t B = t->shape[1], C = t->shape[2];
  long long s0 = (long long)B * C, s1 = C, s2 = 1;
  if (t->strides) {
    s0 = t->strides[0];
    s1 = t->strides[1];
    s2 = t->strides[2];
  }
  if ((C != 1 && s2 != 1) || s0 < 0 || s1 < 0) {
    set_error("%s: innermost stride must be 1 and outer strides non-negative (got %lld, %lld, %lld)", name, s0, s1,
              s2);
    return false;
  }
  *st_t = s0;
  *st_b = s1;
  return true;
}

template <typename T>
T* dl_ptr(const DLTensor* t) {
  return reinterpret_cast<T*>(static_cast<char*>(t->data) + t->byte_offset);
}

int64_t dl_numel(const DLTensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; i++) n *= t->shape[i];
  return n;
}

}  // namespace
}  // namespace nasr

using namespace nasr;

// Host copy between a caller's pageable buffer and the pinned staging: one thread moves about 20 GB/s on this box,
// so large copies are split over a few (a per-block, row-by-row staging that overlaps the DMA was measured and is
// slower: 8000 copies of 4.8 KB cost more than the overlap gains).
static void host_copy(void* dst, const void* src, size_t bytes) {
  const size_t kMin = (size_t)4 << 20;
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)(bytes / kMin);
  nt = nt > 8 ? 8 : nt;
  if (hw && nt > (int)hw) nt = (int)hw;
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::thread th[8];
  const size_t chunk = ((bytes / nt) + 4095) & ~(size_t)4095;
  int started = 0;
  for (int i = 1; i < nt; i++) {
    const size_t o = (size_t)i * chunk;
    if (o >= bytes) break;
    const size_t n = bytes - o < chunk ? bytes - o : chunk;
    th[started++] = std::thread([=]() { memcpy((char*)dst + o, (const char*)src + o, n); });
  }
  memcpy(dst, src, chunk < bytes ? chunk : bytes);
  for (int i = 0; i < started; i++) th[i].join();
}

struct nasr_host_ctx {
  int device;
  int max_T, max_B, max_C, max_L;
  cudaStream_t stream;             // compute (block k runs on s_k[k % n_streams]; stream == s_k[0])
  cudaStream_t s_k[4];
  int n_streams, n_blocks;         // NASR_HOST_STREAMS (1..4, default 2), NASR_HOST_BLOCKS (1..8, default 8)
  cudaStream_t s_in, s_out;        // H2D / D2H copies of the utterance blocks
  cudaEvent_t ev_in[8], ev_done[8];
  void* d_ws_blk[8];               // one workspace per block: blocks on different streams run concurrently
  size_t ws_blk_bytes;
  int blk_B;                       // most utterances a block can hold
  int decoder, beam_width;         // nasr_host_ctx_set_decoder: 0 greedy (default), 1 beam search
  cudaStream_t s_beam;
  cudaEvent_t ev_beam;
  void* d_ws_beam;
  size_t ws_beam_bytes;
  // device
  float *d_logits, *d_grad, *d_loss, *d_grad_loss, *d_nsl, *d_ler;
  int32_t *d_lab_vals, *d_lab_offs, *d_seq, *d_status, *d_hyp_len, *d_dist;
  int64_t* d_hyp;
  void* d_ws;
  size_t ws_bytes;
  // pinned host
  float *h_logits, *h_grad;
  char* h_small;  // labels, offsets, seq_len, grad_loss in; loss, status, hyp_len, nsl, dist, ler out
  int64_t* h_hyp;
  size_t small_bytes;
};

extern "C" {

int nasr_abi_version(void) { return NASR_ABI_VERSION; }

const char* nasr_last_error(void) { return t_err; }

uint64_t nasr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nasr_debug_config(int path, int split_frames) {
  NASR_CHECK_ARG(path >= 0 && (path & 0xff) <= 2 && split_frames >= 0, "nasr_debug_config: bad arguments");
  g_debug_ablate = path >> 8;  // tuning only: switch warp roles off to see what they cost (results are then wrong)
  path &= 0xff;
  g_debug_path = path;
  g_debug_split = split_frames;
  return NASR_OK;
}

int nasr_allreduce_scalars(void* nccl_comm, double* vec, int n, void* stream) {
  NASR_CHECK_ARG(nccl_comm && vec && n > 0, "nasr_allreduce_scalars: bad arguments");
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static std::atomic<allreduce_fn> fn{nullptr};
  allreduce_fn f = fn.load(std::memory_order_acquire);
  if (!f) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    f = h ? reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce")) : nullptr;
    if (!f) {
      set_error("nasr_allreduce_scalars: NCCL (libnccl.so.2) is not available in this process");
      return NASR_ERR_UNSUPPORTED;
    }
    fn.store(f, std::memory_order_release);
  }
  // ncclFloat64 = 8, ncclSum = 0 (nccl.h)
  const int rc = f(vec, vec, (size_t)n, 8, 0, nccl_comm, static_cast<cudaStream_t>(stream));
  if (rc != 0) {
    set_error("nasr_allreduce_scalars: ncclAllReduce failed with code %d", rc);
    return NASR_ERR_CUDA;
  }
  return NASR_OK;
}

int nasr_debug_profile(void* device_buffer) {
  g_debug_prof = static_cast<long long*>(device_buffer);
  return NASR_OK;
}

int nasr_ctc_workspace_bytes(int T, int B, int C, int max_label_len, size_t* out_bytes) {
  NASR_CHECK_ARG(out_bytes, "nasr_ctc_workspace_bytes: out_bytes is NULL");
  NASR_CHECK_ARG(T >= 0 && B >= 0 && C >= 1 && max_label_len >= 0,
                 "nasr_ctc_workspace_bytes: bad shape T=%d B=%d C=%d L=%d", T, B, C, max_label_len);
  return ctc_workspace_bytes(T, B, C, max_label_len, out_bytes);
}

int nasr_ctc_loss_grad_f32(const float* logits, int T, int B, int C, const int32_t* label_values,
                           const int32_t* label_offsets, int max_label_len, const int32_t* seq_len,
                           int blank, float* loss, float* grad, const float* grad_loss,
                           int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return ctc_loss_grad(logits, T, B, C, (long long)B * C, C, label_values, label_offsets, max_label_len, seq_len,
                       blank, loss, grad, grad_loss, status, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int nasr_ctc_loss_grad_strided_f32(const float* logits, int T, int B, int C, long long stride_t,
                                   long long stride_b, const int32_t* label_values,
                                   const int32_t* label_offsets, int max_label_len, const int32_t* seq_len,
                                   int blank, float* loss, float* grad, const float* grad_loss,
                                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return ctc_loss_grad(logits, T, B, C, stride_t, stride_b, label_values, label_offsets, max_label_len, seq_len,
                       blank, loss, grad, grad_loss, status, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int nasr_ctc_loss_grad_dl(const DLTensor* logits, const DLTensor* label_values,
                          const DLTensor* label_offsets, int max_label_len, const DLTensor* seq_len,
                          int blank, const DLTensor* loss, const DLTensor* grad,
                          const DLTensor* grad_loss, const DLTensor* status,
                          const DLTensor* workspace, void* stream) {
  long long st_t = 0, st_b = 0;
  if (!dl_logits_ok(logits, "logits", -1, &st_t, &st_b)) return NASR_ERR_INVALID_ARGUMENT;
  const int dev = logits->device.device_id;
  const int64_t T = logits->shape[0], B = logits->shape[1], C = logits->shape[2];
  NASR_CHECK_ARG(T < (1 << 30) && B < (1 << 30) && C < (1 << 30), "logits: dimension too large");
  if (!dl_ok(label_values, "label_values", 0, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  if (!dl_ok(label_offsets, "label_offsets", 0, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  if (!dl_ok(seq_len, "seq_len", 0, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  if (!dl_ok(loss, "loss", 2, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  if (!dl_ok(status, "status", 0, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  if (!dl_ok(workspace, "workspace", 1, 8, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
  NASR_CHECK_ARG(label_offsets->shape[0] == B + 1, "label_offsets: length %lld, expected B+1=%lld",
                 (long long)label_offsets->shape[0], (long long)(B + 1));
  NASR_CHECK_ARG(seq_len->shape[0] == B, "seq_len: length %lld, expected B=%lld",
                 (long long)seq_len->shape[0], (long long)B);
  NASR_CHECK_ARG(loss->shape[0] == B && status->shape[0] == B, "loss/status: length must be B=%lld",
                 (long long)B);
  float* g = nullptr;
  if (grad) {
    long long gt = 0, gb = 0;
    if (!dl_logits_ok(grad, "grad", dev, &gt, &gb)) return NASR_ERR_INVALID_ARGUMENT;
    NASR_CHECK_ARG(grad->shape[0] == T && grad->shape[1] == B && grad->shape[2] == C,
                   "grad: shape differs from logits");
    NASR_CHECK_ARG(gt == st_t && gb == st_b, "grad: strides differ from the logits' (%lld,%lld vs %lld,%lld)", gt,
                   gb, st_t, st_b);
    g = dl_ptr<float>(grad);
  }
  const float* gl = nullptr;
  if (grad_loss) {
    if (!dl_ok(grad_loss, "grad_loss", 2, 32, 1, dev)) return NASR_ERR_INVALID_ARGUMENT;
    NASR_CHECK_ARG(grad_loss->shape[0] == B, "grad_loss: length must be B");
    gl = dl_ptr<const float>(grad_loss);
  }
  return ctc_loss_grad(dl_ptr<const float>(logits), (int)T, (int)B, (int)C, st_t, st_b,
                       dl_ptr<const int32_t>(label_values), dl_ptr<const int32_t>(label_offsets),
                       max_label_len, dl_ptr<const int32_t>(seq_len), blank, dl_ptr<float>(loss), g, gl,
                       dl_ptr<int32_t>(status), dl_ptr<void>(workspace), (size_t)dl_numel(workspace),
                       static_cast<cudaStream_t>(stream));
}

int nasr_ctc_greedy_decode_i64(const float* logits, int T, int B, int C, const int32_t* seq_len,
                               int blank, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                               float* neg_sum_logits, void* stream) {
  return greedy_decode(logits, T, B, C, (long long)B * C, C, seq_len, blank, merge_repeated, hyp, hyp_len,
                       neg_sum_logits, static_cast<cudaStream_t>(stream));
}

int nasr_ctc_greedy_decode_strided_i64(const float* logits, int T, int B, int C, long long stride_t,
                                       long long stride_b, const int32_t* seq_len, int blank,
                                       int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                                       float* neg_sum_logits, void* stream) {
  return greedy_decode(logits, T, B, C, stride_t, stride_b, seq_len, blank, merge_repeated, hyp, hyp_len,
                       neg_sum_logits, static_cast<cudaStream_t>(stream));
}

int nasr_ctc_beam_workspace_bytes(int T, int B, int C, int beam_width, size_t* out_bytes) {
  NASR_CHECK_ARG(out_bytes, "nasr_ctc_beam_workspace_bytes: out_bytes is NULL");
  NASR_CHECK_ARG(T >= 0 && B >= 0 && C >= 1 && beam_width >= 1,
                 "nasr_ctc_beam_workspace_bytes: bad shape T=%d B=%d C=%d beam_width=%d", T, B, C, beam_width);
  return ctc_beam_workspace_bytes(T, B, C, beam_width, out_bytes);
}

int nasr_ctc_beam_search_i64(const float* logits, int T, int B, int C, const int32_t* seq_len, int blank,
                             int beam_width, int top_paths, int merge_repeated, int64_t* hyp,
                             int32_t* hyp_len, float* log_prob, void* workspace, size_t workspace_bytes,
                             void* stream) {
  return ctc_beam_search(logits, T, B, C, (long long)B * C, C, seq_len, blank, beam_width, top_paths,
                         merge_repeated, hyp, hyp_len, log_prob, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int nasr_ctc_beam_search_strided_i64(const float* logits, int T, int B, int C, long long stride_t,
                                     long long stride_b, const int32_t* seq_len, int blank, int beam_width,
                                     int top_paths, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                                     float* log_prob, void* workspace, size_t workspace_bytes, void* stream) {
  return ctc_beam_search(logits, T, B, C, stride_t, stride_b, seq_len, blank, beam_width, top_paths,
                         merge_repeated, hyp, hyp_len, log_prob, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

int nasr_labels_coo_to_csr_i32(const int64_t* indices, const int32_t* values, int N, int row0, int B,
                               int32_t* offsets, int32_t* values_out, int32_t* info, void* stream) {
  return labels_coo_to_csr(indices, values, N, row0, B, offsets, values_out, info, static_cast<cudaStream_t>(stream));
}

int nasr_hyp_to_sparse_i64(const int64_t* hyp, int hyp_stride, const int32_t* hyp_offsets, int B,
                           int64_t* indices, int64_t* values, int64_t* dense_shape, void* stream) {
  return hyp_to_sparse(hyp, hyp_stride, hyp_offsets, B, indices, values, dense_shape,
                       static_cast<cudaStream_t>(stream));
}

int nasr_edit_distance_i64(const int64_t* hyp, int hyp_stride, const int32_t* hyp_len,
                           const int32_t* truth_values, const int32_t* truth_offsets,
                           int max_truth_len, int B, int normalize, int32_t* dist, float* ler,
                           void* stream) {
  return edit_distance_dense(hyp, hyp_stride, hyp_len, truth_values, truth_offsets, max_truth_len, B,
                             normalize, dist, ler, static_cast<cudaStream_t>(stream));
}

int nasr_edit_distance_csr_i64(const int64_t* hyp_values, const int32_t* hyp_offsets, int max_hyp_len,
                               const int32_t* truth_values, const int32_t* truth_offsets,
                               int max_truth_len, int B, int normalize, int32_t* dist, float* ler,
                               void* stream) {
  return edit_distance_csr(hyp_values, hyp_offsets, max_hyp_len, truth_values, truth_offsets,
                           max_truth_len, B, normalize, dist, ler, static_cast<cudaStream_t>(stream));
}

int nasr_batch_sums_f64(const float* loss, const float* ler, const int32_t* dist, int B,
                        double* sums, void* stream) {
  return batch_sums(loss, ler, dist, B, sums, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------
// HOST-buffer context
// ------------------------------------------------------------------------------------------------
static size_t small_layout(int B, int N, size_t* o_vals, size_t* o_offs, size_t* o_seq, size_t* o_gl,
                           size_t* o_loss, size_t* o_status, size_t* o_hl, size_t* o_nsl,
                           size_t* o_dist, size_t* o_ler) {
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = (o + bytes + 63) & ~(size_t)63;
    return r;
  };
  *o_vals = take(sizeof(int32_t) * (size_t)(N + 1));
  *o_offs = take(sizeof(int32_t) * (size_t)(B + 1));
  *o_seq = take(sizeof(int32_t) * (size_t)B);
  *o_gl = take(sizeof(float) * (size_t)B);
  *o_loss = take(sizeof(float) * (size_t)B);
  *o_status = take(sizeof(int32_t) * (size_t)B);
  *o_hl = take(sizeof(int32_t) * (size_t)B);
  *o_nsl = take(sizeof(float) * (size_t)B);
  *o_dist = take(sizeof(int32_t) * (size_t)B);
  *o_ler = take(sizeof(float) * (size_t)B);
  return o;
}

void nasr_host_ctx_destroy(nasr_host_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->s_in) cudaStreamSynchronize(c->s_in);
  if (c->s_out) cudaStreamSynchronize(c->s_out);
  for (int i = 0; i < 8; i++) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  }
  for (int i = 1; i < 4; i++)
    if (c->s_k[i]) {
      cudaStreamSynchronize(c->s_k[i]);
      cudaStreamDestroy(c->s_k[i]);
    }
  for (int i = 0; i < 8; i++) cudaFree(c->d_ws_blk[i]);
  if (c->s_beam) {
    cudaStreamSynchronize(c->s_beam);
    cudaStreamDestroy(c->s_beam);
  }
  if (c->ev_beam) cudaEventDestroy(c->ev_beam);
  cudaFree(c->d_ws_beam);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  cudaFree(c->d_logits); cudaFree(c->d_grad); cudaFree(c->d_loss); cudaFree(c->d_grad_loss);
  cudaFree(c->d_nsl); cudaFree(c->d_ler); cudaFree(c->d_lab_vals); cudaFree(c->d_lab_offs);
  cudaFree(c->d_seq); cudaFree(c->d_status); cudaFree(c->d_hyp_len); cudaFree(c->d_dist);
  cudaFree(c->d_hyp); cudaFree(c->d_ws);
  cudaFreeHost(c->h_logits); cudaFreeHost(c->h_grad); cudaFreeHost(c->h_small); cudaFreeHost(c->h_hyp);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int nasr_host_ctx_create(int device, int max_T, int max_B, int max_C, int max_label_len,
                         nasr_host_ctx** out) {
  NASR_CHECK_ARG(out, "nasr_host_ctx_create: out is NULL");
  NASR_CHECK_ARG(max_T >= 1 && max_B >= 1 && max_C >= 1 && max_label_len >= 0,
                 "nasr_host_ctx_create: bad maxima");
  NASR_CUDA(cudaSetDevice(device));
  size_t ws = 0;
  int rc = ctc_workspace_bytes(max_T, max_B, max_C, max_label_len, &ws);
  if (rc != NASR_OK) return rc;
  nasr_host_ctx* c = new (std::nothrow) nasr_host_ctx();
  NASR_CHECK_ARG(c, "nasr_host_ctx_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->device = device; c->max_T = max_T; c->max_B = max_B; c->max_C = max_C; c->max_L = max_label_len;
  c->ws_bytes = ws;
  const size_t nlog = (size_t)max_T * max_B * max_C;
  const size_t N = (size_t)max_B * max_label_len;
  size_t o[10];
  c->small_bytes = small_layout(max_B, (int)N, &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7], &o[8], &o[9]);
#define NASR_CTX_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      set_error("nasr_host_ctx_create: %s failed: %s", #expr, cudaGetErrorString(e_));  \
      nasr_host_ctx_destroy(c);                                                         \
      return NASR_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)
  NASR_CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->s_k[0] = c->stream;
  for (int i = 1; i < 4; i++) NASR_CTX_TRY(cudaStreamCreateWithFlags(&c->s_k[i], cudaStreamNonBlocking));
  {
    const char* e = getenv("NASR_HOST_STREAMS");
    c->n_streams = e ? atoi(e) : 2;
    c->n_streams = c->n_streams < 1 ? 1 : (c->n_streams > 4 ? 4 : c->n_streams);
    e = getenv("NASR_HOST_BLOCKS");
    c->n_blocks = e ? atoi(e) : 8;
    c->n_blocks = c->n_blocks < 1 ? 1 : (c->n_blocks > 8 ? 8 : c->n_blocks);
    c->blk_B = (max_B + c->n_blocks - 1) / c->n_blocks + 1;
    rc = ctc_workspace_bytes(max_T, c->blk_B, max_C, max_label_len, &c->ws_blk_bytes);
    if (rc != NASR_OK) {
      nasr_host_ctx_destroy(c);
      return rc;
    }
    for (int i = 0; i < c->n_blocks; i++) NASR_CTX_TRY(cudaMalloc(&c->d_ws_blk[i], c->ws_blk_bytes));
  }
  NASR_CTX_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  NASR_CTX_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 8; i++) {
    NASR_CTX_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    NASR_CTX_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  }
  NASR_CTX_TRY(cudaMalloc(&c->d_logits, sizeof(float) * nlog));
  NASR_CTX_TRY(cudaMalloc(&c->d_grad, sizeof(float) * nlog));
  NASR_CTX_TRY(cudaMalloc(&c->d_loss, sizeof(float) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_grad_loss, sizeof(float) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_nsl, sizeof(float) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_ler, sizeof(float) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_lab_vals, sizeof(int32_t) * (N + 1)));
  NASR_CTX_TRY(cudaMalloc(&c->d_lab_offs, sizeof(int32_t) * (max_B + 1)));
  NASR_CTX_TRY(cudaMalloc(&c->d_seq, sizeof(int32_t) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_status, sizeof(int32_t) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_hyp_len, sizeof(int32_t) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_dist, sizeof(int32_t) * max_B));
  NASR_CTX_TRY(cudaMalloc(&c->d_hyp, sizeof(int64_t) * (size_t)max_B * max_T));
  NASR_CTX_TRY(cudaMalloc(&c->d_ws, ws));
  NASR_CTX_TRY(cudaMallocHost(&c->h_logits, sizeof(float) * nlog));
  NASR_CTX_TRY(cudaMallocHost(&c->h_grad, sizeof(float) * nlog));
  NASR_CTX_TRY(cudaMallocHost(&c->h_small, c->small_bytes));
  NASR_CTX_TRY(cudaMallocHost(&c->h_hyp, sizeof(int64_t) * (size_t)max_B * max_T));
#undef NASR_CTX_TRY
  *out = c;
  return NASR_OK;
}

int nasr_host_ctx_set_decoder(nasr_host_ctx* c, int decoder, int beam_width) {
  NASR_CHECK_ARG(c, "nasr_host_ctx_set_decoder: ctx is NULL");
  NASR_CHECK_ARG(decoder == 0 || decoder == 1, "nasr_host_ctx_set_decoder: decoder must be 0 (greedy) or 1 (beam)");
  if (decoder == 1) {
    NASR_CHECK_ARG(beam_width >= 1, "nasr_host_ctx_set_decoder: beam_width must be >= 1");
    NASR_CUDA(cudaSetDevice(c->device));
    size_t need = 0;
    int rc = ctc_beam_workspace_bytes(c->max_T, c->max_B, c->max_C, beam_width, &need);
    if (rc != NASR_OK) return rc;
    if (need > c->ws_beam_bytes) {
      cudaFree(c->d_ws_beam);
      c->d_ws_beam = nullptr;
      c->ws_beam_bytes = 0;
      NASR_CUDA(cudaMalloc(&c->d_ws_beam, need));
      c->ws_beam_bytes = need;
    }
    if (!c->s_beam) NASR_CUDA(cudaStreamCreateWithFlags(&c->s_beam, cudaStreamNonBlocking));
    if (!c->ev_beam) NASR_CUDA(cudaEventCreateWithFlags(&c->ev_beam, cudaEventDisableTiming));
  }
  c->decoder = decoder;
  c->beam_width = beam_width;
  return NASR_OK;
}

// An error return from the middle of a step: wait for the copies and kernels already queued on the context's streams,
// so that the caller may reuse or free its buffers (the status of the waits is ignored: rc is what is reported).
static int host_drain(nasr_host_ctx* c, int rc) {
  if (c->s_in) cudaStreamSynchronize(c->s_in);
  for (int i = 0; i < 4; i++)
    if (c->s_k[i]) cudaStreamSynchronize(c->s_k[i]);
  if (c->s_beam) cudaStreamSynchronize(c->s_beam);
  if (c->s_out) cudaStreamSynchronize(c->s_out);
  return rc;
}

float* nasr_host_ctx_pinned_logits(nasr_host_ctx* c) { return c ? c->h_logits : nullptr; }
float* nasr_host_ctx_pinned_grad(nasr_host_ctx* c) { return c ? c->h_grad : nullptr; }

int nasr_host_ctc_step(nasr_host_ctx* c, const float* logits, int T, int B, int C,
                       const int32_t* label_values, const int32_t* label_offsets,
                       const int32_t* seq_len, int blank, const float* grad_loss, float* loss,
                       float* grad, int32_t* status, int64_t* hyp, int32_t* hyp_len,
                       float* neg_sum_logits, int32_t* dist, float* ler) {
  NASR_CHECK_ARG(c, "nasr_host_ctc_step: ctx is NULL");
  NASR_CHECK_ARG(T >= 1 && B >= 1 && C >= 1 && T <= c->max_T && B <= c->max_B && C <= c->max_C &&
                     (size_t)T * B * C <= (size_t)c->max_T * c->max_B * c->max_C,
                 "nasr_host_ctc_step: shape T=%d B=%d C=%d exceeds the context maxima", T, B, C);
  NASR_CHECK_ARG(logits && label_offsets && seq_len && loss && status, "nasr_host_ctc_step: NULL argument");
  const int N = label_offsets[B];
  NASR_CHECK_ARG(label_offsets[0] == 0 && N >= 0 && (size_t)N <= (size_t)c->max_B * c->max_L,
                 "nasr_host_ctc_step: label_offsets inconsistent with the context maxima");
  int Lmax = 0;
  for (int b = 0; b < B; b++) {
    const int l = label_offsets[b + 1] - label_offsets[b];
    NASR_CHECK_ARG(l >= 0, "nasr_host_ctc_step: label_offsets not monotone at row %d", b);
    Lmax = l > Lmax ? l : Lmax;
  }
  NASR_CHECK_ARG(Lmax <= c->max_L, "nasr_host_ctc_step: transcript of %d labels exceeds max_label_len=%d",
                 Lmax, c->max_L);
  NASR_CUDA(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const size_t nlog = (size_t)T * B * C;
  size_t o_vals, o_offs, o_seq, o_gl, o_loss, o_status, o_hl, o_nsl, o_dist, o_ler;
  small_layout(c->max_B, c->max_B * c->max_L, &o_vals, &o_offs, &o_seq, &o_gl, &o_loss, &o_status,
               &o_hl, &o_nsl, &o_dist, &o_ler);
  // stage inputs in pinned memory (skipped for logits when the caller filled the pinned buffer)
  if (logits != c->h_logits) host_copy(c->h_logits, logits, sizeof(float) * nlog);
  if (N) memcpy(c->h_small + o_vals, label_values, sizeof(int32_t) * N);
  memcpy(c->h_small + o_offs, label_offsets, sizeof(int32_t) * (B + 1));
  memcpy(c->h_small + o_seq, seq_len, sizeof(int32_t) * B);
  if (grad_loss) memcpy(c->h_small + o_gl, grad_loss, sizeof(float) * B);
  // Small inputs first, then the batch in blocks of utterances: the H2D copy of block k+1, the kernels of
  // block k and the D2H copy of block k-1 run concurrently (PCIe is full duplex; logits and grad are the two
  // 4*T*B*C-byte transfers that dominate this call).  A block is a [T, Bk, C] column slab of the [T, B, C]
  // tensors: a pitched copy on the bus, a strided launch on the device.  (Alternating the blocks' kernels
  // between two compute streams was measured and gave nothing.)
  cudaStream_t sin = c->s_in, sout = c->s_out;
  if (N) NASR_CUDA(cudaMemcpyAsync(c->d_lab_vals, c->h_small + o_vals, sizeof(int32_t) * N, cudaMemcpyHostToDevice, sin));
  NASR_CUDA(cudaMemcpyAsync(c->d_lab_offs, c->h_small + o_offs, sizeof(int32_t) * (B + 1), cudaMemcpyHostToDevice, sin));
  NASR_CUDA(cudaMemcpyAsync(c->d_seq, c->h_small + o_seq, sizeof(int32_t) * B, cudaMemcpyHostToDevice, sin));
  if (grad_loss) NASR_CUDA(cudaMemcpyAsync(c->d_grad_loss, c->h_small + o_gl, sizeof(float) * B, cudaMemcpyHostToDevice, sin));
  const bool want_decode = hyp || hyp_len || neg_sum_logits || dist || ler;
  // Blocks run on up to four compute streams with a workspace each: a block's kernels are bound by the length of
  // one utterance's recursion, not by the number of utterances, so the kernels of consecutive blocks overlap
  // and the call is left with the H2D copy of the whole batch plus one block's kernels and D2H copy.
  const int nblk = (nlog * sizeof(float) >= ((size_t)8 << 20) && B >= 8 * c->n_blocks) ? c->n_blocks : 1;
  const size_t pitch = sizeof(float) * (size_t)B * C;
  for (int k = 0; k < nblk; k++) {
    const int b0 = (int)((long long)B * k / nblk), b1 = (int)((long long)B * (k + 1) / nblk);
    const int Bk = b1 - b0;
    if (Bk == 0) continue;
    const size_t off = (size_t)b0 * C;
    NASR_CUDA(cudaMemcpy2DAsync(c->d_logits + off, pitch, c->h_logits + off, pitch, sizeof(float) * (size_t)Bk * C,
                                (size_t)T, cudaMemcpyHostToDevice, sin));
    NASR_CUDA(cudaEventRecord(c->ev_in[k], sin));
    s = nblk > 1 ? c->s_k[k % c->n_streams] : c->stream;
    void* ws = nblk > 1 ? c->d_ws_blk[k] : c->d_ws;
    const size_t ws_bytes = nblk > 1 ? c->ws_blk_bytes : c->ws_bytes;
    NASR_CUDA(cudaStreamWaitEvent(s, c->ev_in[k], 0));
    int rc = ctc_loss_grad(c->d_logits + off, T, Bk, C, (long long)B * C, C, c->d_lab_vals, c->d_lab_offs + b0, Lmax,
                           c->d_seq + b0, blank, c->d_loss + b0, grad ? c->d_grad + off : nullptr,
                           grad_loss ? c->d_grad_loss + b0 : nullptr, c->d_status + b0, ws, ws_bytes, s);
    if (rc != NASR_OK) return host_drain(c, rc);
    if (want_decode && c->decoder == 0) {
      rc = greedy_decode(c->d_logits + off, T, Bk, C, (long long)B * C, C, c->d_seq + b0, blank, 1,
                         c->d_hyp + (size_t)b0 * T, c->d_hyp_len + b0, c->d_nsl + b0, s);
      if (rc != NASR_OK) return host_drain(c, rc);
      if (dist || ler) {
        rc = edit_distance_dense(c->d_hyp + (size_t)b0 * T, T, c->d_hyp_len + b0, c->d_lab_vals, c->d_lab_offs + b0,
                                 Lmax, Bk, 1, c->d_dist + b0, c->d_ler + b0, s);
        if (rc != NASR_OK) return host_drain(c, rc);
      }
    }
    NASR_CUDA(cudaEventRecord(c->ev_done[k], s));
    NASR_CUDA(cudaStreamWaitEvent(sout, c->ev_done[k], 0));
    if (grad)
      NASR_CUDA(cudaMemcpy2DAsync(c->h_grad + off, pitch, c->d_grad + off, pitch, sizeof(float) * (size_t)Bk * C,
                                  (size_t)T, cudaMemcpyDeviceToHost, sout));
  }
  if (want_decode && c->decoder == 1) {
    // Beam search: one launch over the whole batch on its own stream once the last block has arrived (a launch per
    // block would put eight latency-bound searches in a row), concurrent with the blocks' loss/gradient kernels.
    NASR_CUDA(cudaStreamWaitEvent(c->s_beam, c->ev_in[nblk - 1], 0));
    int rc = ctc_beam_search(c->d_logits, T, B, C, (long long)B * C, C, c->d_seq, blank, c->beam_width, 1, 1, c->d_hyp,
                             c->d_hyp_len, c->d_nsl, c->d_ws_beam, c->ws_beam_bytes, c->s_beam);
    if (rc != NASR_OK) return host_drain(c, rc);
    if (dist || ler) {
      rc = edit_distance_dense(c->d_hyp, T, c->d_hyp_len, c->d_lab_vals, c->d_lab_offs, Lmax, B, 1, c->d_dist,
                               c->d_ler, c->s_beam);
      if (rc != NASR_OK) return host_drain(c, rc);
    }
    NASR_CUDA(cudaEventRecord(c->ev_beam, c->s_beam));
    NASR_CUDA(cudaStreamWaitEvent(sout, c->ev_beam, 0));
  }
  s = sout;  // everything below is ordered after the last block's kernels
  NASR_CUDA(cudaMemcpyAsync(c->h_small + o_loss, c->d_loss, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
  NASR_CUDA(cudaMemcpyAsync(c->h_small + o_status, c->d_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
  if (want_decode) {
    NASR_CUDA(cudaMemcpyAsync(c->h_small + o_hl, c->d_hyp_len, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
    NASR_CUDA(cudaMemcpyAsync(c->h_small + o_nsl, c->d_nsl, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
    if (hyp) NASR_CUDA(cudaMemcpyAsync(c->h_hyp, c->d_hyp, sizeof(int64_t) * (size_t)B * T, cudaMemcpyDeviceToHost, s));
    if (dist || ler) {
      NASR_CUDA(cudaMemcpyAsync(c->h_small + o_dist, c->d_dist, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, s));
      NASR_CUDA(cudaMemcpyAsync(c->h_small + o_ler, c->d_ler, sizeof(float) * B, cudaMemcpyDeviceToHost, s));
    }
  }
  NASR_CUDA(cudaStreamSynchronize(s));
  memcpy(loss, c->h_small + o_loss, sizeof(float) * B);
  memcpy(status, c->h_small + o_status, sizeof(int32_t) * B);
  if (grad && grad != c->h_grad) host_copy(grad, c->h_grad, sizeof(float) * nlog);
  if (hyp_len) memcpy(hyp_len, c->h_small + o_hl, sizeof(int32_t) * B);
  if (neg_sum_logits) memcpy(neg_sum_logits, c->h_small + o_nsl, sizeof(float) * B);
  if (hyp) memcpy(hyp, c->h_hyp, sizeof(int64_t) * (size_t)B * T);
  if (dist) memcpy(dist, c->h_small + o_dist, sizeof(int32_t) * B);
  if (ler) memcpy(ler, c->h_small + o_ler, sizeof(float) * B);
  return NASR_OK;
}

}  // extern "C"
