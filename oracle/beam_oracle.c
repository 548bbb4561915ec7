/*
 * CPU oracle (plain C) of TensorFlow's CTC beam search decoder — TEST INFRASTRUCTURE ONLY: used by tests/,
 * __graft_entry__.smoke() and the CPU legs of bench.py; the product library never links, loads or calls it.
 *
 * Restates tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width=100, top_paths=1,
 * merge_repeated=True), the op behind the reference's create_model (networks/tfnetwork.py:62,64).  The algorithm
 * lives in TensorFlow 1.x (version unpinned, absent from /root/reference and from this image); what is
 * restated is the CONTROL FLOW of TF's core/util/ctc/ctc_beam_search.h, CTCBeamSearchDecoder::Step / TopPaths:
 *   - a prefix trie of beam entries {parent, label, oldp, newp, children}; an entry is Active() while
 *     newp.total != log 0; children are created when their parent is first expanded;
 *   - leaves_ is a TopN of beam_width entries ordered by newp.total with a peek at its worst ("bottom");
 *   - Step: extract the leaves; oldp = newp; every leaf keeps itself (label term joins the parent's when the
 *     parent is Active, blank term = oldp.total + P(blank)) and is pushed back; then every leaf that is still a
 *     candidate against the current bottom expands: an inactive child gets newp = (log 0, P(l) + (l == label ?
 *     oldp.blank : oldp.total)), and if it is a candidate it replaces the bottom (whose newp is reset, i.e. it
 *     leaves the beam), else it is deactivated (one order-dependent detail of TF is NOT reproduced, see the
 *     comment in the growth loop);
 *   - TopPaths: the top_paths best leaves, label sequences read towards the root, repeats merged in the output.
 * This is deliberately a different program from oracle/beam_oracle.py (a dictionary of prefixes re-ranked per
 * frame) so that the two check each other; both use float64 where TF uses float32, the per-frame update
 *   label' = LSE(label + lp, parent_term + lp),  total' = log(e^blank' + e^(label+lp) + e^(parent_term+lp))
 * and the tie-break (score, kept prefix before new extension, 64-bit prefix hash) where TF has whatever its
 * heap does.  PARITY WITH TF ITSELF IS UNPINNED (no TF here, no tests or vectors in the reference).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef uint64_t u64;
#define LOG0 (-INFINITY)
#define ROOT_HASH 0x243f6a8885a308d3ull

static u64 mix64(u64 x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
static u64 child_hash(u64 h, int label) { return mix64(h + 0x9e3779b97f4a7c15ull * (u64)(label + 1)); }

static double lse2(double a, double b) {
  if (a == LOG0) return b;
  if (b == LOG0) return a;
  double m = a > b ? a : b, n = a > b ? b : a;
  return m + log1p(exp(n - m));
}

typedef struct {
  double blank, label, total;
} prob_t;

typedef struct {
  int parent, label, len, is_new;
  u64 hash;
  prob_t oldp, newp;
  int* children; /* [C] node index per label, -1 = never created; NULL until first expansion */
} entry_t;

typedef struct {
  entry_t* e;
  int n, cap;
  int* heap; /* indices; heap[0] is the WORST entry of the beam (TF: leaves_.peek_bottom()) */
  int hn;
} beam_t;

/* a is worse than b: lower total; ties: a new extension is worse than a kept prefix; then the larger hash */
static int worse(const entry_t* a, const entry_t* b) {
  if (a->newp.total != b->newp.total) return a->newp.total < b->newp.total;
  if (a->is_new != b->is_new) return a->is_new > b->is_new;
  return a->hash > b->hash;
}

static void heap_push(beam_t* s, int idx) {
  int i = s->hn++;
  s->heap[i] = idx;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!worse(&s->e[s->heap[i]], &s->e[s->heap[p]])) break;
    int t = s->heap[i];
    s->heap[i] = s->heap[p];
    s->heap[p] = t;
    i = p;
  }
}

static int heap_pop(beam_t* s) {
  int top = s->heap[0];
  s->heap[0] = s->heap[--s->hn];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, w = i;
    if (l < s->hn && worse(&s->e[s->heap[l]], &s->e[s->heap[w]])) w = l;
    if (r < s->hn && worse(&s->e[s->heap[r]], &s->e[s->heap[w]])) w = r;
    if (w == i) break;
    int t = s->heap[i];
    s->heap[i] = s->heap[w];
    s->heap[w] = t;
    i = w;
  }
  return top;
}

static int new_entry(beam_t* s, int parent, int label) {
  if (s->n == s->cap) {
    s->cap *= 2;
    s->e = (entry_t*)realloc(s->e, sizeof(entry_t) * (size_t)s->cap);
  }
  entry_t* x = &s->e[s->n];
  x->parent = parent;
  x->label = label;
  x->len = parent < 0 ? 0 : s->e[parent].len + 1;
  x->hash = parent < 0 ? ROOT_HASH : child_hash(s->e[parent].hash, label);
  x->is_new = 0;
  x->oldp.blank = x->oldp.label = x->oldp.total = LOG0;
  x->newp = x->oldp;
  x->children = NULL;
  return s->n++;
}

static void reset_prob(prob_t* p) { p->blank = p->label = p->total = LOG0; }

/* one utterance: x rows at stride st_t, Tb frames */
/* margin (may be NULL): margin[0] = smallest NONZERO |difference of totals| over all decisions taken -- a candidate
 * tested against the beam's worst entry, and the order of the paths returned; margin[1] = number of decisions between
 * bitwise-equal totals.  A result whose margin[0] is within rounding of zero depends on the last bit of exp/log and
 * may legitimately differ between two correct implementations.  Bitwise-equal totals are of two kinds: twins -- the
 * same operations on equal inputs (two labels with equal logits in a frame, and everything that grows from such a
 * pair once their parents have left the beam), which every implementation ties and the hash decides identically --
 * and, with quantised logits, different prefixes that happen to be equal here and one ulp apart elsewhere. */
static void beam_one(const float* x, long long st_t, int Tb, int C, int blank, int W, int P, int merge,
                     long long* hyp /*[P][T]*/, int T, int* hyp_len /*[P]*/, double* log_prob /*[P]*/,
                     double* margin) {
  double mg = INFINITY, nzero = 0.0;
#define NASR_MARGIN(dm_) do { const double d__ = (dm_); if (d__ == 0.0) nzero += 1.0; else if (d__ < mg) mg = d__; } while (0)
  beam_t s;
  s.cap = 1024;
  s.n = 0;
  s.e = (entry_t*)malloc(sizeof(entry_t) * (size_t)s.cap);
  s.heap = (int*)malloc(sizeof(int) * (size_t)(W + 1));
  s.hn = 0;
  int* branches = (int*)malloc(sizeof(int) * (size_t)(W + 1));
  double* lp = (double*)malloc(sizeof(double) * (size_t)C);
  int root = new_entry(&s, -1, -1);
  s.e[root].newp.blank = 0.0; /* log 1 */
  s.e[root].newp.total = 0.0;
  heap_push(&s, root);

  for (int t = 0; t < Tb; t++) {
    const float* row = x + (size_t)t * st_t;
    double m = row[0];
    for (int c = 1; c < C; c++) m = row[c] > m ? row[c] : m;
    double sum = 0.0;
    for (int c = 0; c < C; c++) sum += exp((double)row[c] - m);
    const double lse = log(sum);
    for (int c = 0; c < C; c++) lp[c] = ((double)row[c] - m) - lse;

    /* branches = leaves_.Extract(); leaves_.Reset() */
    int nb = s.hn;
    memcpy(branches, s.heap, sizeof(int) * (size_t)nb);
    s.hn = 0;
    for (int i = 0; i < nb; i++) {
      entry_t* b = &s.e[branches[i]];
      b->oldp = b->newp;
      b->is_new = 0;
    }
    for (int i = 0; i < nb; i++) {
      entry_t* b = &s.e[branches[i]];
      double b1 = LOG0, b2 = LOG0;
      if (b->parent >= 0) {
        const entry_t* par = &s.e[b->parent];
        b1 = b->oldp.label + lp[b->label];
        if (par->newp.total != LOG0) /* parent Active(): its oldp is this frame's too */
          b2 = (b->label == par->label ? par->oldp.blank : par->oldp.total) + lp[b->label];
      }
      const double a = b->oldp.total + lp[blank];
      b->newp.blank = a;
      b->newp.label = lse2(b1, b2);
      double M = a > b1 ? a : b1;
      M = M > b2 ? M : b2;
      b->newp.total = M == LOG0 ? LOG0 : M + log(exp(a - M) + exp(b1 - M) + exp(b2 - M));
    }
    for (int i = 0; i < nb; i++)
      if (s.e[branches[i]].newp.total != LOG0) heap_push(&s, branches[i]); /* (TF pushes log-0 leaves too) */

    /* grow new leaves */
    for (int i = 0; i < nb; i++) {
      const int bi = branches[i];
      if (!(s.e[bi].oldp.total > LOG0)) continue;
      /* TF: skip b unless oldp.total > bottom.total; '<' here so that exact ties reach the per-child test */
      if (s.hn == W && s.e[bi].oldp.total < s.e[s.heap[0]].newp.total) {
        /* the skipped prefix's best extension is at most oldp.total + max lp: its distance to the bottom counts */
        NASR_MARGIN(fabs(s.e[s.heap[0]].newp.total - s.e[bi].oldp.total));
        continue;
      }
      if (!s.e[bi].children) {
        int* ch = (int*)malloc(sizeof(int) * (size_t)C);
        for (int c = 0; c < C; c++) ch[c] = -1;
        s.e[bi].children = ch;
      }
      for (int l = 0; l < C; l++) {
        if (l == blank) continue;
        int ci = s.e[bi].children[l];
        if (ci >= 0 && s.e[ci].newp.total != LOG0) continue; /* already Active */
        const double v = lp[l] + (l == s.e[bi].label ? s.e[bi].oldp.blank : s.e[bi].oldp.total);
        if (!(v > LOG0)) continue;
        entry_t probe;
        probe.newp.total = v;
        probe.is_new = 1;
        probe.hash = child_hash(s.e[bi].hash, l);
        const int cand = s.hn < W || worse(&s.e[s.heap[0]], &probe);
        if (s.hn == W) {
          NASR_MARGIN(fabs(v - s.e[s.heap[0]].newp.total));
        }
        if (cand) {
          if (ci < 0) {
            ci = new_entry(&s, bi, l); /* may move s.e */
            s.e[bi].children[l] = ci;
          }
          entry_t* c = &s.e[ci];
          c->newp.blank = LOG0;
          c->newp.label = v;
          c->newp.total = v;
          c->is_new = 1;
          if (s.hn == W) {
            const int bottom = heap_pop(&s);
            reset_prob(&s.e[bottom].newp); /* bottom is no longer in the beam search */
          }
          heap_push(&s, ci);
        }
        /* else TF deactivates the child: c.oldp.Reset(); c.newp.Reset().  newp is log 0 already (the child is
         * inactive); oldp is deliberately NOT reset here: the child may be a leaf that this frame's growth has
         * just evicted and whose own turn to expand (from its oldp) is still to come.  TF's outcome in that
         * case depends on the iteration order of its heap; the semantics defined here (and in beam_oracle.py
         * and the CUDA kernel) are order-independent: every prefix active at the start of the frame expands. */
      }
    }
  }

  /* TopPaths: best first by (total, smaller hash) */
  int nl = s.hn;
  int* order = (int*)malloc(sizeof(int) * (size_t)(nl + 1));
  memcpy(order, s.heap, sizeof(int) * (size_t)nl);
  for (int i = 1; i < nl; i++) { /* insertion sort, nl <= W */
    int k = order[i], j = i - 1;
    while (j >= 0) {
      const entry_t *a = &s.e[order[j]], *b = &s.e[k];
      const int b_first = b->newp.total > a->newp.total || (b->newp.total == a->newp.total && b->hash < a->hash);
      if (!b_first) break;
      order[j + 1] = order[j];
      j--;
    }
    order[j + 1] = k;
  }
  for (int p = 0; p + 1 < nl && p < P; p++) {
    NASR_MARGIN(fabs(s.e[order[p]].newp.total - s.e[order[p + 1]].newp.total));
  }
#undef NASR_MARGIN
  if (margin) {
    margin[0] = mg;
    margin[1] = nzero;
  }
  for (int p = 0; p < P; p++) {
    long long* out = hyp + (size_t)p * T;
    if (p >= nl) {
      hyp_len[p] = 0;
      log_prob[p] = LOG0;
      continue;
    }
    const entry_t* e = &s.e[order[p]];
    log_prob[p] = e->newp.total;
    int len = e->len;
    for (int i = len - 1, k = order[p]; i >= 0; i--) {
      out[i] = s.e[k].label;
      k = s.e[k].parent;
    }
    if (merge) {
      int o = 0;
      for (int i = 0; i < len; i++)
        if (i == 0 || out[i] != out[i - 1]) out[o++] = out[i];
      len = o;
    }
    hyp_len[p] = len;
  }
  for (int i = 0; i < s.n; i++) free(s.e[i].children);
  free(order);
  free(lp);
  free(branches);
  free(s.heap);
  free(s.e);
}

typedef struct {
  const float* logits;
  int T, B, C;
  long long st_t, st_b;
  const int32_t* seq_len;
  int blank, W, P, merge;
  long long* hyp;
  int* hyp_len;
  double* log_prob;
  double* margin;
  int next;
} job_t;

static void* worker(void* p) {
  job_t* j = (job_t*)p;
  for (;;) {
    int b = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (b >= j->B) break;
    int Tb = j->seq_len[b];
    Tb = Tb < 0 ? 0 : (Tb > j->T ? j->T : Tb);
    beam_one(j->logits + (size_t)b * j->st_b, j->st_t, Tb, j->C, j->blank, j->W, j->P, j->merge,
             j->hyp + (size_t)b * j->P * j->T, j->T, j->hyp_len + (size_t)b * j->P, j->log_prob + (size_t)b * j->P,
             j->margin ? j->margin + 2 * (size_t)b : NULL);
  }
  return NULL;
}

/* logits float32 [T,B,C] at element strides (st_t, st_b, 1); hyp int64 [B,P,T]; hyp_len int32 [B,P];
 * log_prob float64 [B,P].  Utterances are spread over num_threads threads (<= 0: all cores). */
/* margin: NULL or float64 [B][2] (see beam_one). */
int oracle_beam_search_margin(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                              const int32_t* seq_len, int blank, int W, int P, int merge, long long* hyp,
                              int* hyp_len, double* log_prob, double* margin, int num_threads) {
  if (num_threads <= 0) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    num_threads = n > 0 ? (int)n : 1;
  }
  if (num_threads > B) num_threads = B;
  job_t job = {logits, T, B, C, st_t, st_b, seq_len, blank, W, P, merge, hyp, hyp_len, log_prob, margin, 0};
  if (num_threads <= 1) {
    worker(&job);
    return 0;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)num_threads);
  int started = 0;
  for (int i = 0; i < num_threads - 1; i++)
    if (pthread_create(&th[started], NULL, worker, &job) == 0) started++;
  worker(&job);
  for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
  free(th);
  return 0;
}

int oracle_beam_search(const float* logits, int T, int B, int C, long long st_t, long long st_b,
                       const int32_t* seq_len, int blank, int W, int P, int merge, long long* hyp, int* hyp_len,
                       double* log_prob, int num_threads) {
  return oracle_beam_search_margin(logits, T, B, C, st_t, st_b, seq_len, blank, W, P, merge, hyp, hyp_len, log_prob,
                                   NULL, num_threads);
}
