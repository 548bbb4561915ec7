/*
 * CPU oracle (plain C) for the CTC loss+grad / greedy decode / edit distance path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product library
 * (neuralasr_b200/csrc) never links, loads or calls it.
 *
 * Restates the algorithm behind networks/tfnetwork.py:58-59 (tf.nn.ctc_loss +
 * _CTCLossGrad), :63 (tf.nn.ctc_greedy_decoder) and :68 (tf.edit_distance) of the
 * reference.  That algorithm lives in TensorFlow 1.x (version unpinned, absent
 * from /root/reference and from this image); what is restated is TF's published
 * CPU algorithm (core/util/ctc/ctc_loss_calculator.{h,cc}, ctc_loss_util.h,
 * ctc_decoder.h, core/kernels/edit_distance_op.cc) as summarised in SURVEY.md
 * Appendix A: softmax -> log-alpha -> log-beta -> gradient per utterance, scalar
 * log1p/exp log-sum-exp, utterances sharded over a thread pool (pthreads here, the
 * intra-op pool in TF).
 *
 * PARITY UNPINNED by the reference (it has no tests / golden vectors); pinned in
 * tests/test_oracle.py against oracle/ctc_oracle.py, torch's independent CPU CTC
 * (fixtures in tests/golden) and closed-form cases.
 *
 * Two precisions are compiled from one body: *_f32 computes in float like TF's
 * kernel does (this is the timed CPU baseline), *_f64 computes in double and is
 * the numerical truth the CUDA kernels are compared with.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NASR_ST_LABEL_OUT_OF_RANGE 1
#define NASR_ST_SEQ_LEN_OUT_OF_RANGE 2
#define NASR_ST_NOT_ENOUGH_TIME 4
#define NASR_ST_NO_VALID_PATH 8

int oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* Minimal thread pool: utterances are handed out one at a time from an atomic counter,
 * which is how TF's Shard() spreads CTC batch elements over its intra-op pool. */
typedef void (*item_fn)(int b, void* ctx);
typedef struct {
  item_fn fn;
  void* ctx;
  int n;
  int next;
} pf_job;

static void* pf_worker(void* p) {
  pf_job* j = (pf_job*)p;
  for (;;) {
    int b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->n) break;
    j->fn(b, j->ctx);
  }
  return NULL;
}

static void parallel_for(int n, item_fn fn, void* ctx, int num_threads) {
  if (num_threads <= 0) num_threads = oracle_num_threads();
  if (num_threads > n) num_threads = n;
  pf_job job = {fn, ctx, n, 0};
  if (num_threads <= 1) {
    pf_worker(&job);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * num_threads);
  int started = 0;
  for (int i = 0; i < num_threads - 1; i++)
    if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
  pf_worker(&job);
  for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
  free(th);
}

#define DEFINE_CTC(SUFFIX, real, EXP, LOG, LOG1P)                                          \
  static inline real lse_##SUFFIX(real a, real b) {                                        \
    if (a == -INFINITY) return b;                                                          \
    if (b == -INFINITY) return a;                                                          \
    return a > b ? a + LOG1P(EXP(b - a)) : b + LOG1P(EXP(a - b));                          \
  }                                                                                        \
  /* one utterance; x points at logits[0,b,0], frame stride = stride elements */           \
  static int ctc_one_##SUFFIX(const float* x, long stride, int Tb, int C, const int32_t* lab, \
                              int L, int blank, real* loss_out, float* grad, long gstride,  \
                              real grad_scale) {                                           \
    int status = 0;                                                                        \
    for (int i = 0; i < L; i++)                                                            \
      if (lab[i] < 0 || lab[i] >= C || lab[i] == blank) {                                              \
        *loss_out = INFINITY;                                                              \
        return NASR_ST_LABEL_OUT_OF_RANGE;                                                 \
      }                                                                                    \
    if (Tb == 0) {                                                                         \
      *loss_out = 0;                                                                       \
      return 0;                                                                            \
    }                                                                                      \
    int need = L;                                                                          \
    for (int i = 1; i < L; i++) need += (lab[i] == lab[i - 1]);                            \
    if (Tb < need) status |= NASR_ST_NOT_ENOUGH_TIME;                                      \
    const int U = 2 * L + 1;                                                               \
    int* lp = (int*)malloc(sizeof(int) * U);                                               \
    for (int u = 0; u < U; u++) lp[u] = (u & 1) ? lab[u >> 1] : blank;                     \
    real* y = (real*)malloc(sizeof(real) * (size_t)Tb * C);                                \
    real* ly = (real*)malloc(sizeof(real) * (size_t)Tb * C);                               \
    real* al = (real*)malloc(sizeof(real) * (size_t)Tb * U);                               \
    real* be = (real*)malloc(sizeof(real) * (size_t)Tb * U);                               \
    for (int t = 0; t < Tb; t++) {                                                         \
      const float* r = x + (long)t * stride;                                               \
      real m = r[0];                                                                       \
      for (int c = 1; c < C; c++) m = r[c] > m ? r[c] : m;                                 \
      real s = 0;                                                                          \
      for (int c = 0; c < C; c++) {                                                        \
        y[(size_t)t * C + c] = EXP((real)r[c] - m);                                        \
        s += y[(size_t)t * C + c];                                                         \
      }                                                                                    \
      for (int c = 0; c < C; c++) {                                                        \
        y[(size_t)t * C + c] /= s;                                                         \
        ly[(size_t)t * C + c] = LOG(y[(size_t)t * C + c]);                                 \
      }                                                                                    \
    }                                                                                      \
    for (size_t i = 0; i < (size_t)Tb * U; i++) al[i] = be[i] = -INFINITY;                 \
    al[0] = ly[blank];                                                                     \
    if (U > 1) al[1] = ly[lp[1]];                                                          \
    for (int t = 1; t < Tb; t++) {                                                         \
      int lo = U - 2 * (Tb - t);                                                           \
      if (lo < 0) lo = 0;                                                                  \
      int hi = 2 * (t + 1);                                                                \
      if (hi > U) hi = U;                                                                  \
      const real* ap = al + (size_t)(t - 1) * U;                                           \
      real* ac = al + (size_t)t * U;                                                       \
      const real* l = ly + (size_t)t * C;                                                  \
      for (int u = lo; u < hi; u++) {                                                      \
        real s = ap[u];                                                                    \
        if (u > 0) s = lse_##SUFFIX(s, ap[u - 1]);                                         \
        if (u > 1 && lp[u] != blank && lp[u] != lp[u - 2]) s = lse_##SUFFIX(s, ap[u - 2]); \
        ac[u] = l[lp[u]] + s;                                                              \
      }                                                                                    \
    }                                                                                      \
    for (int u = (U - 2 > 0 ? U - 2 : 0); u < U; u++) be[(size_t)(Tb - 1) * U + u] = 0;    \
    for (int t = Tb - 2; t >= 0; t--) {                                                    \
      int lo = U - 2 * (Tb - t);                                                           \
      if (lo < 0) lo = 0;                                                                  \
      int hi = 2 * (t + 1);                                                                \
      if (hi > U) hi = U;                                                                  \
      const real* bn = be + (size_t)(t + 1) * U;                                           \
      real* bc = be + (size_t)t * U;                                                       \
      const real* l = ly + (size_t)(t + 1) * C;                                            \
      for (int u = lo; u < hi; u++) {                                                      \
        real s = bn[u] + l[lp[u]];                                                         \
        if (u + 1 < U) s = lse_##SUFFIX(s, bn[u + 1] + l[lp[u + 1]]);                      \
        if (u + 2 < U && lp[u] != blank && lp[u] != lp[u + 2])                             \
          s = lse_##SUFFIX(s, bn[u + 2] + l[lp[u + 2]]);                                   \
        bc[u] = s;                                                                         \
      }                                                                                    \
    }                                                                                      \
    real logp = -INFINITY;                                                                 \
    for (int u = 0; u < U; u++) logp = lse_##SUFFIX(logp, al[u] + be[u]);                  \
    if (logp == -INFINITY) {                                                               \
      status |= NASR_ST_NO_VALID_PATH;                                                     \
      *loss_out = INFINITY;                                                                \
      if (grad)                                                                            \
        for (int t = 0; t < Tb; t++)                                                       \
          for (int c = 0; c < C; c++)                                                      \
            grad[(long)t * gstride + c] = (float)(grad_scale * y[(size_t)t * C + c]);      \
    } else {                                                                               \
      *loss_out = -logp;                                                                   \
      if (grad) {                                                                          \
        real* occ = (real*)malloc(sizeof(real) * C);                                       \
        for (int t = 0; t < Tb; t++) {                                                     \
          for (int c = 0; c < C; c++) occ[c] = -INFINITY;                                  \
          for (int u = 0; u < U; u++)                                                      \
            occ[lp[u]] = lse_##SUFFIX(occ[lp[u]], al[(size_t)t * U + u] + be[(size_t)t * U + u]); \
          for (int c = 0; c < C; c++)                                                      \
            grad[(long)t * gstride + c] =                                                  \
                (float)(grad_scale * (y[(size_t)t * C + c] - EXP(occ[c] - logp)));         \
        }                                                                                  \
        free(occ);                                                                         \
      }                                                                                    \
    }                                                                                      \
    free(lp);                                                                              \
    free(y);                                                                               \
    free(ly);                                                                              \
    free(al);                                                                              \
    free(be);                                                                              \
    return status;                                                                         \
  }                                                                                        \
  typedef struct {                                                                         \
    const float* logits;                                                                   \
    int T, B, C;                                                                           \
    const int32_t *label_values, *label_offsets, *seq_len;                                 \
    int blank;                                                                             \
    real* loss;                                                                            \
    float* grad;                                                                           \
    const float* grad_loss;                                                                \
    int32_t* status;                                                                       \
  } ctc_args_##SUFFIX;                                                                     \
  static void ctc_item_##SUFFIX(int b, void* p) {                                          \
    ctc_args_##SUFFIX* a = (ctc_args_##SUFFIX*)p;                                          \
    int Tb = a->seq_len[b];                                                                \
    if (Tb < 0 || Tb > a->T) {                                                             \
      a->status[b] = NASR_ST_SEQ_LEN_OUT_OF_RANGE;                                         \
      a->loss[b] = INFINITY;                                                               \
      return;                                                                              \
    }                                                                                      \
    a->status[b] = ctc_one_##SUFFIX(                                                       \
        a->logits + (long)b * a->C, (long)a->B * a->C, Tb, a->C,                           \
        a->label_values + a->label_offsets[b], a->label_offsets[b + 1] - a->label_offsets[b], \
        a->blank, &a->loss[b], a->grad ? a->grad + (long)b * a->C : NULL, (long)a->B * a->C, \
        a->grad_loss ? (real)a->grad_loss[b] : (real)1);                                   \
  }                                                                                        \
  /* logits [T,B,C] f32, CSR labels, seq_len[B]; loss[B] (real), grad [T,B,C] f32 or NULL, \
   * grad_loss[B] or NULL (=1), status[B].  num_threads<=0 -> all cores. */                \
  int oracle_ctc_loss_grad_##SUFFIX(const float* logits, int T, int B, int C,              \
                                    const int32_t* label_values, const int32_t* label_offsets, \
                                    const int32_t* seq_len, int blank, real* loss, float* grad, \
                                    const float* grad_loss, int32_t* status, int num_threads) { \
    if (grad) memset(grad, 0, sizeof(float) * (size_t)T * B * C);                          \
    ctc_args_##SUFFIX a = {logits, T, B, C, label_values, label_offsets, seq_len, blank,   \
                           loss, grad, grad_loss, status};                                 \
    parallel_for(B, ctc_item_##SUFFIX, &a, num_threads);                                   \
    return 0;                                                                              \
  }

DEFINE_CTC(f32, float, expf, logf, log1pf)
DEFINE_CTC(f64, double, exp, log, log1p)

/* tf.nn.ctc_greedy_decoder: first argmax of raw logits, drop blank, merge repeats.
 * hyp is dense [B, T] int64 (row b holds hyp_len[b] ids), neg_sum_logits f32[B]. */
typedef struct {
  const float* logits;
  int T, B, C;
  const int32_t* seq_len;
  int blank, merge_repeated;
  int64_t* hyp;
  int32_t* hyp_len;
  float* neg_sum_logits;
} greedy_args;

static void greedy_item(int b, void* p) {
  greedy_args* a = (greedy_args*)p;
  int Tb = a->seq_len[b];
  if (Tb < 0) Tb = 0;
  if (Tb > a->T) Tb = a->T;
  int prev = -1, n = 0;
  float acc = 0.f;
  for (int t = 0; t < Tb; t++) {
    const float* r = a->logits + ((long)t * a->B + b) * a->C;
    int am = 0;
    float m = r[0];
    for (int c = 1; c < a->C; c++)
      if (r[c] > m) {
        m = r[c];
        am = c;
      }
    acc -= m;
    if (am != a->blank && !(a->merge_repeated && am == prev)) a->hyp[(long)b * a->T + n++] = am;
    prev = am;
  }
  a->hyp_len[b] = n;
  a->neg_sum_logits[b] = acc;
}

int oracle_greedy_decode(const float* logits, int T, int B, int C, const int32_t* seq_len,
                         int blank, int merge_repeated, int64_t* hyp, int32_t* hyp_len,
                         float* neg_sum_logits, int num_threads) {
  greedy_args a = {logits, T, B, C, seq_len, blank, merge_repeated, hyp, hyp_len, neg_sum_logits};
  parallel_for(B, greedy_item, &a, num_threads);
  return 0;
}

static int lev(const int32_t* a, int n, const int32_t* b, int m) {
  if (n == 0) return m;
  if (m == 0) return n;
  int* row = (int*)malloc(sizeof(int) * (m + 1));
  for (int j = 0; j <= m; j++) row[j] = j;
  for (int i = 1; i <= n; i++) {
    int diag = row[0];
    row[0] = i;
    for (int j = 1; j <= m; j++) {
      int up = row[j];
      int best = diag + (a[i - 1] != b[j - 1]);
      if (up + 1 < best) best = up + 1;
      if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
      diag = up;
      row[j] = best;
    }
  }
  int d = row[m];
  free(row);
  return d;
}

/* tf.edit_distance(hyp, truth, normalize): CSR hyp (int32 after the reference's cast) and truth. */
typedef struct {
  const int32_t *hyp_values, *hyp_offsets, *truth_values, *truth_offsets;
  int normalize;
  int32_t* dist;
  float* ler;
} ed_args;

static void ed_item(int b, void* p) {
  ed_args* a = (ed_args*)p;
  int n = a->truth_offsets[b + 1] - a->truth_offsets[b];
  int m = a->hyp_offsets[b + 1] - a->hyp_offsets[b];
  int d = lev(a->truth_values + a->truth_offsets[b], n, a->hyp_values + a->hyp_offsets[b], m);
  a->dist[b] = d;
  if (!a->normalize)
    a->ler[b] = (float)d;
  else if (n == 0)
    a->ler[b] = d ? INFINITY : 0.f;
  else
    a->ler[b] = (float)d / (float)n;
}

int oracle_edit_distance(const int32_t* hyp_values, const int32_t* hyp_offsets,
                         const int32_t* truth_values, const int32_t* truth_offsets, int B,
                         int normalize, int32_t* dist, float* ler, int num_threads) {
  ed_args a = {hyp_values, hyp_offsets, truth_values, truth_offsets, normalize, dist, ler};
  parallel_for(B, ed_item, &a, num_threads);
  return 0;
}
