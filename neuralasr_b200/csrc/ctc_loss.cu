// CTC loss + d(loss)/d(logits) for sm_100a: one CTA per utterance, one launch per batch.
//
// Replaces tf.nn.ctc_loss + _CTCLossGrad behind create_loss (reference networks/tfnetwork.py:58-59);
// semantics per SURVEY.md Appendix A.1 (blank passed in, ctc_merge_repeated=True).
//
// Number representation ("log-space in the exponent, linear in the mantissa"):
//   alpha and beta are carried as fp64 values times a power of two that is tracked as an exact
//   integer (S): stored = true * 2^S.  Every kRescale frames the row is renormalised so its largest
//   state lies in [1,2) and S absorbs the shift.  The recursion is then 2 adds + 1 multiply per state
//   (no exp/log on the dependent chain), the dynamic range between states of one frame is 2^-1074..1
//   (e^-744), and the rounding error does not grow with |loss| the way an fp32 log-space recursion's
//   does (tests/test_oracle.py shows TF-style fp32 log-space drifting past 1e-4 at T=300).
//
// Memory plan (the [T,U] alpha lattice never goes to HBM):
//   pass 1 (alpha, t ascending)  keeps two rows in shared memory and writes a checkpoint row every K
//          frames to the workspace (B * ceil(T/K) * U doubles: 51 MB at B=256,T=1000,U=401 — L2 resident);
//   pass 2 (beta, t descending, by segments of K frames) reloads the segment's checkpoint, recomputes
//          its K alpha rows into shared memory, then walks beta backwards through the segment and
//          emits grad[t,b,:] = grad_loss * (softmax - alpha*beta/p) row by row, coalesced.
//   The softmax of a segment's K frames is computed once per pass into shared memory, so the logits
//   are read from HBM twice (once per pass) and the gradient is written once: 12*C bytes per frame.
#include <math.h>

#include "nasr_common.cuh"

namespace nasr {

namespace {

constexpr int kRescale = 8;           // frames between renormalisations (K is a multiple of it)
constexpr int kSkipBit = 1 << 30;     // meta[u]: state may take the u-2 transition
constexpr int kLabelMask = kSkipBit - 1;

struct Plan {
  int K;            // frames per segment
  int nseg;         // ceil(T / K)
  int Upad;         // padded extended-state count (even)
  int Cpad;
  int threads;
  size_t smem;      // dynamic shared memory bytes
  size_t ws_ckpt;   // bytes of checkpoint rows
  size_t ws_scale;  // bytes of checkpoint scales
};

struct SmemLayout {
  size_t seg, row, post, yseg, meta, lab, cls_off, cls_idx, seg_s, red, total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline SmemLayout smem_layout(int K, int Upad, int Cpad, int Lmax) {
  SmemLayout s;
  size_t o = 0;
  s.seg = o;      o = align16(o + sizeof(double) * (size_t)K * Upad);
  s.row = o;      o = align16(o + sizeof(double) * 2 * (size_t)Upad);
  s.post = o;     o = align16(o + sizeof(float) * 2 * (size_t)Upad);
  s.yseg = o;     o = align16(o + sizeof(float) * (size_t)K * Cpad);
  s.meta = o;     o = align16(o + sizeof(int) * (size_t)Upad);
  s.lab = o;      o = align16(o + sizeof(int) * (size_t)(Lmax + 1));
  s.cls_off = o;  o = align16(o + sizeof(int) * (size_t)(Cpad + 2));
  s.cls_idx = o;  o = align16(o + sizeof(int) * (size_t)(Lmax + 1));
  s.seg_s = o;    o = align16(o + sizeof(int) * (size_t)K);
  s.red = o;      o = align16(o + sizeof(double) * 40);
  s.total = o;
  return s;
}

constexpr size_t kSmemBudget = 200 * 1024;

bool make_plan(int T, int B, int C, int Lmax, Plan* p) {
  int U = 2 * Lmax + 1;
  p->Upad = (U + 1) & ~1;
  p->Cpad = (C + 3) & ~3;
  p->K = 16;
  if (smem_layout(p->K, p->Upad, p->Cpad, Lmax).total > kSmemBudget) p->K = 8;
  SmemLayout s = smem_layout(p->K, p->Upad, p->Cpad, Lmax);
  if (s.total > kSmemBudget) return false;
  p->smem = s.total;
  p->nseg = (T + p->K - 1) / p->K;
  if (p->nseg < 1) p->nseg = 1;
  p->threads = U > 256 ? 512 : (U > 128 ? 256 : 128);
  if (C > 512 && p->threads < 256) p->threads = 256;
  p->ws_ckpt = sizeof(double) * (size_t)B * p->nseg * p->Upad;
  p->ws_scale = align16(sizeof(int) * (size_t)B * p->nseg);
  return true;
}

struct Params {
  const float* logits;
  int T, B, C;
  const int32_t* lab_vals;
  const int32_t* lab_offs;
  const int32_t* seq_len;
  int blank;
  int Lmax;
  float* loss;
  float* grad;
  const float* grad_loss;
  int32_t* status;
  double* ckpt;
  int* ckpt_s;
  int K, nseg, Upad, Cpad;
};

__device__ __forceinline__ double pow2_double(int e) {  // 2^e for e in [-1022, 1023]
  return __hiloint2double((1023 + e) << 20, 0);
}

// softmax rows of frames [t0, t1) of utterance b into yseg[(t-t0)*Cpad + c]; one warp per frame.
__device__ void softmax_rows(const Params& p, int b, int t0, int t1, float* yseg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int C = p.C;
  for (int t = t0 + warp; t < t1; t += nw) {
    const float* x = p.logits + ((size_t)t * p.B + b) * C;
    float* y = yseg + (size_t)(t - t0) * p.Cpad;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) {
      float v = __ldg(x + c);
      y[c] = v;
      m = fmaxf(m, v);
    }
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      float e = expf(y[c] - m);
      y[c] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) y[c] = y[c] / s;
  }
}

// Renormalise a row so that its largest element lies in [1,2); returns the exponent removed.
// Contains one __syncthreads(); the caller must sync again before other threads read the row.
__device__ int block_rescale(double* row, int U, int* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int mx = 0;
  for (int u = threadIdx.x; u < U; u += blockDim.x) mx = max(mx, __double2hiint(row[u]));
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[lane < nw ? lane : 0];
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (mx == 0) return 0;  // all zero (or denormal dust): nothing to normalise
  int e = ((mx >> 20) & 0x7ff) - 1023;
  if (e == 0) return 0;
  e = max(-1022, min(1022, e));
  const double sc = pow2_double(-e);
  for (int u = threadIdx.x; u < U; u += blockDim.x) row[u] *= sc;
  return e;
}

__device__ __forceinline__ void alpha_step(const double* __restrict__ prev, double* __restrict__ cur,
                                           const float* __restrict__ yrow,
                                           const int* __restrict__ meta, int U) {
  for (int u = threadIdx.x; u < U; u += blockDim.x) {
    const int m = meta[u];
    double s = prev[u];
    if (u >= 1) s += prev[u - 1];
    if (m & kSkipBit) s += prev[u - 2];
    cur[u] = s * (double)yrow[m & kLabelMask];
  }
}

__global__ void __launch_bounds__(512) ctc_loss_grad_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout sl = smem_layout(p.K, p.Upad, p.Cpad, p.Lmax);
  double* seg = reinterpret_cast<double*>(smem + sl.seg);
  double* row = reinterpret_cast<double*>(smem + sl.row);
  float* post = reinterpret_cast<float*>(smem + sl.post);
  float* yseg = reinterpret_cast<float*>(smem + sl.yseg);
  int* meta = reinterpret_cast<int*>(smem + sl.meta);
  int* lab = reinterpret_cast<int*>(smem + sl.lab);
  int* cls_off = reinterpret_cast<int*>(smem + sl.cls_off);
  int* cls_idx = reinterpret_cast<int*>(smem + sl.cls_idx);
  int* seg_s = reinterpret_cast<int*>(smem + sl.seg_s);
  int* red = reinterpret_cast<int*>(smem + sl.red);
  double* red_d = reinterpret_cast<double*>(smem + sl.red) + 20;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int T = p.T, B = p.B, C = p.C, K = p.K, Upad = p.Upad, blank = p.blank;
  const int Tb = p.seq_len[b];
  const int l0 = p.lab_offs[b];
  const int L = p.lab_offs[b + 1] - l0;
  const int U = 2 * L + 1;
  const float gs = p.grad_loss ? p.grad_loss[b] : 1.0f;
  const size_t rstride = (size_t)B * C;          // elements between consecutive frames of one utterance
  float* gbase = p.grad ? p.grad + (size_t)b * C : nullptr;

  // ---- validation (TF: InvalidArgument) -------------------------------------------------------
  int st = 0;
  if (Tb < 0 || Tb > T) st |= NASR_ST_SEQ_LEN_OUT_OF_RANGE;
  int bad = 0;
  for (int i = tid; i < L; i += nt) {
    const int v = p.lab_vals[l0 + i];
    lab[i] = v;
    bad |= (v < 0 || v >= blank);
  }
  if (L > p.Lmax) bad = 1;  // caller lied about max_label_len: shared memory is not sized for it
  if (__syncthreads_or(bad)) st |= NASR_ST_LABEL_OUT_OF_RANGE;
  const int Tz = st ? 0 : Tb;  // first frame whose gradient row is all zero
  if (gbase) {
    const size_t n = (size_t)(T - Tz) * C;
    for (size_t i = tid; i < n; i += nt) {
      const size_t t = Tz + i / C;
      gbase[t * rstride + (i % C)] = 0.f;
    }
  }
  if (st) {
    if (tid == 0) {
      p.status[b] = st;
      p.loss[b] = INFINITY;
    }
    return;
  }
  if (Tb == 0) {
    if (tid == 0) {
      p.status[b] = 0;
      p.loss[b] = 0.f;
    }
    return;
  }

  // ---- per-utterance tables: state metadata, class -> label positions (for the gradient gather) ----
  for (int c = tid; c < C + 2; c += nt) cls_off[c] = 0;
  if (tid == 0) red[32] = 0;
  __syncthreads();
  int rep = 0;
  for (int i = tid; i < L; i += nt) {
    atomicAdd(&cls_off[lab[i] + 1], 1);
    if (i > 0 && lab[i] == lab[i - 1]) rep++;
  }
  for (int u = tid; u < U; u += nt) {
    int m = blank;
    if (u & 1) {
      m = lab[u >> 1];
      if (u >= 3 && m != lab[(u >> 1) - 1]) m |= kSkipBit;
    }
    meta[u] = m;
  }
  if (rep) atomicAdd(&red[32], rep);
  __syncthreads();
  if (Tb < L + red[32]) st |= NASR_ST_NOT_ENOUGH_TIME;
  if (warp == 0) {
    // in-place inclusive prefix sum over cls_off[0..C] (cls_off[0] = 0, cls_off[c+1] = count of class c)
    const int per = (C + 1 + 31) / 32;
    const int lo = min(C + 1, lane * per), hi = min(C + 1, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += cls_off[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - s;
    for (int i = lo; i < hi; i++) {
      run += cls_off[i];
      cls_off[i] = run;
    }
  }
  __syncthreads();
  // now cls_off[c] = number of labels with class < c  (cls_off[0] = 0), cls_off[C] = L
  for (int i = tid; i < L; i += nt) {
    const int v = lab[i];
    int r = 0;
    for (int j = 0; j < i; j++) r += (lab[j] == v);
    cls_idx[cls_off[v] + r] = i;
  }

  double* ckpt = p.ckpt + (size_t)b * p.nseg * Upad;
  int* ckpt_s = p.ckpt_s + (size_t)b * p.nseg;

  // ---- pass 1: alpha, ascending, checkpoints every K frames -----------------------------------
  double* prev = row;
  double* cur = row + Upad;
  int S = 0;  // stored = true * 2^S
  const int nseg = (Tb + K - 1) / K;
  for (int s = 0; s < nseg; s++) {
    const int t0 = s * K, t1 = min(Tb, t0 + K);
    __syncthreads();  // previous segment's readers of yseg are done
    softmax_rows(p, b, t0, t1, yseg);
    __syncthreads();
    for (int t = t0; t < t1; t++) {
      const float* yrow = yseg + (size_t)(t - t0) * p.Cpad;
      if (t == 0) {
        for (int u = tid; u < U; u += nt)
          cur[u] = u < 2 ? (double)yrow[meta[u] & kLabelMask] : 0.0;
      } else {
        alpha_step(prev, cur, yrow, meta, U);
      }
      __syncthreads();
      if ((t % kRescale) == kRescale - 1) {
        S -= block_rescale(cur, U, red);
        __syncthreads();
      }
      if (t == t0) {  // checkpoint = the row of the segment's first frame (after any rescale: none at t0)
        for (int u = tid; u < U; u += nt) ckpt[(size_t)s * Upad + u] = cur[u];
        if (tid == 0) ckpt_s[s] = S;
      }
      double* tmp = prev;
      prev = cur;
      cur = tmp;
    }
  }
  // prev = alpha row of frame Tb-1
  const double phat = prev[U - 1] + (U > 1 ? prev[U - 2] : 0.0);
  const int S_T = S;
  __syncthreads();
  if (!(phat > 0.0)) {
    // no valid path: loss = +inf, gradient = softmax
    if (tid == 0) {
      p.status[b] = st | NASR_ST_NO_VALID_PATH;
      p.loss[b] = INFINITY;
    }
    if (gbase) {
      for (int s = 0; s < nseg; s++) {
        const int t0 = s * K, t1 = min(Tb, t0 + K);
        __syncthreads();
        softmax_rows(p, b, t0, t1, yseg);
        __syncthreads();
        const int n = (t1 - t0) * C;
        for (int i = tid; i < n; i += nt) {
          const int tt = i / C, c = i % C;
          gbase[(size_t)(t0 + tt) * rstride + c] = gs * yseg[(size_t)tt * p.Cpad + c];
        }
      }
    }
    return;
  }
  int ep;
  const double mp = frexp(phat, &ep);  // phat = mp * 2^ep, mp in [0.5, 1)
  if (tid == 0) {
    p.status[b] = st;
    p.loss[b] = (float)(-(log(mp) + (double)(ep - S_T) * 0.6931471805599453));
  }
  if (!gbase) return;
  const double inv_mp = 1.0 / mp;

  // ---- pass 2: beta, descending by segments, with alpha recompute and gradient emission -----------
  double* bprev = row;
  double* bcur = row + Upad;
  int Sb = 0;
  for (int s = nseg - 1; s >= 0; s--) {
    const int t0 = s * K, t1 = min(Tb, t0 + K);
    __syncthreads();
    softmax_rows(p, b, t0, t1, yseg);
    for (int u = tid; u < U; u += nt) seg[u] = ckpt[(size_t)s * Upad + u];
    if (tid == 0) seg_s[0] = ckpt_s[s];
    __syncthreads();
    {
      int Sa = seg_s[0];
      for (int t = t0 + 1; t < t1; t++) {
        double* ar = seg + (size_t)(t - t0) * Upad;
        alpha_step(ar - Upad, ar, yseg + (size_t)(t - t0) * p.Cpad, meta, U);
        __syncthreads();
        if ((t % kRescale) == kRescale - 1) {
          Sa -= block_rescale(ar, U, red);
          __syncthreads();
        }
        if (tid == 0) seg_s[t - t0] = Sa;
      }
      __syncthreads();
    }
    for (int t = t1 - 1; t >= t0; t--) {
      const float* yrow = yseg + (size_t)(t - t0) * p.Cpad;
      const double* ar = seg + (size_t)(t - t0) * Upad;
      float* pt = post + (size_t)(t & 1) * Upad;
      // factor = 2^(S_T - Sa(t) - Sb) / phat, exponent clamped to the representable range
      int q = S_T - seg_s[t - t0] - Sb - ep;
      q = max(-1000, min(1000, q));
      const double factor = inv_mp * pow2_double(q);
      for (int u = tid; u < U; u += nt) {
        double bv;
        if (t == Tb - 1) {
          bv = (u >= U - 2) ? 1.0 : 0.0;
        } else {
          bv = bprev[u];
          if (u + 1 < U) bv += bprev[u + 1];
          if (u + 2 < U && (meta[u + 2] & kSkipBit)) bv += bprev[u + 2];
        }
        pt[u] = (float)(ar[u] * bv * factor);
        bcur[u] = bv * (double)yrow[meta[u] & kLabelMask];
      }
      // gradient row of the previous iteration's frame (t+1): its posteriors are complete
      if (t < t1 - 1) {
        const int tg = t + 1;
        const float* pg = post + (size_t)(tg & 1) * Upad;
        const float* yg = yseg + (size_t)(tg - t0) * p.Cpad;
        float* g = gbase + (size_t)tg * rstride;
        for (int c = tid; c < C; c += nt) {
          if (c == blank) continue;
          float occ = 0.f;
          for (int k = cls_off[c]; k < cls_off[c + 1]; k++) occ += pg[2 * cls_idx[k] + 1];
          g[c] = gs * (yg[c] - occ);
        }
        if (warp == nw - 1) {
          float occ = 0.f;
          for (int u = 2 * lane; u < U; u += 64) occ += pg[u];
          occ = warp_sum(occ);
          if (lane == 0) g[blank] = gs * (yg[blank] - occ);
        }
      }
      __syncthreads();
      if (((Tb - 1 - t) % kRescale) == kRescale - 1) {
        Sb -= block_rescale(bcur, U, red);
        __syncthreads();
      }
      double* tmp = bprev;
      bprev = bcur;
      bcur = tmp;
    }
    {  // flush the gradient row of frame t0 before yseg is overwritten
      const int tg = t0;
      const float* pg = post + (size_t)(tg & 1) * Upad;
      const float* yg = yseg;
      float* g = gbase + (size_t)tg * rstride;
      for (int c = tid; c < C; c += nt) {
        if (c == blank) continue;
        float occ = 0.f;
        for (int k = cls_off[c]; k < cls_off[c + 1]; k++) occ += pg[2 * cls_idx[k] + 1];
        g[c] = gs * (yg[c] - occ);
      }
      if (warp == nw - 1) {
        float occ = 0.f;
        for (int u = 2 * lane; u < U; u += 64) occ += pg[u];
        occ = warp_sum(occ);
        if (lane == 0) g[blank] = gs * (yg[blank] - occ);
      }
    }
  }
  (void)red_d;
}

}  // namespace

int ctc_workspace_bytes(int T, int B, int C, int Lmax, size_t* out) {
  Plan pl;
  if (!make_plan(T, B, C, Lmax, &pl)) {
    set_error("nasr_ctc: shapes T=%d C=%d max_label_len=%d need more shared memory than one SM has",
              T, C, Lmax);
    return NASR_ERR_UNSUPPORTED;
  }
  *out = pl.ws_ckpt + pl.ws_scale + 256;
  return NASR_OK;
}

int ctc_loss_grad(const float* logits, int T, int B, int C, const int32_t* label_values,
                  const int32_t* label_offsets, int Lmax, const int32_t* seq_len, int blank,
                  float* loss, float* grad, const float* grad_loss, int32_t* status, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
  NASR_CHECK_ARG(T >= 0 && B >= 0 && C >= 1 && Lmax >= 0, "nasr_ctc_loss_grad: bad shape T=%d B=%d C=%d L=%d", T, B, C, Lmax);
  NASR_CHECK_ARG(blank >= 0 && blank < C, "nasr_ctc_loss_grad: blank=%d outside [0,%d)", blank, C);
  NASR_CHECK_ARG(C < kSkipBit, "nasr_ctc_loss_grad: C too large");
  if (B == 0) return NASR_OK;
  NASR_CHECK_ARG(logits || T == 0, "nasr_ctc_loss_grad: logits is NULL");
  NASR_CHECK_ARG(label_offsets && seq_len && loss && status, "nasr_ctc_loss_grad: NULL argument");
  NASR_CHECK_ARG(label_values || Lmax == 0, "nasr_ctc_loss_grad: label_values is NULL");
  Plan pl;
  if (!make_plan(T, B, C, Lmax, &pl)) {
    set_error("nasr_ctc_loss_grad: shapes T=%d C=%d max_label_len=%d exceed shared memory", T, C, Lmax);
    return NASR_ERR_UNSUPPORTED;
  }
  const size_t need = pl.ws_ckpt + pl.ws_scale + 256;
  if (workspace_bytes < need || !workspace) {
    set_error("nasr_ctc_loss_grad: workspace too small (%zu < %zu)", workspace_bytes, need);
    return NASR_ERR_WORKSPACE_TOO_SMALL;
  }
  uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  Params p;
  p.logits = logits; p.T = T; p.B = B; p.C = C;
  p.lab_vals = label_values; p.lab_offs = label_offsets; p.seq_len = seq_len;
  p.blank = blank; p.Lmax = Lmax;
  p.loss = loss; p.grad = grad; p.grad_loss = grad_loss; p.status = status;
  p.ckpt_s = reinterpret_cast<int*>(base);
  p.ckpt = reinterpret_cast<double*>(base + pl.ws_scale);
  p.K = pl.K; p.nseg = pl.nseg; p.Upad = pl.Upad; p.Cpad = pl.Cpad;
  static bool attr_set = false;
  if (!attr_set) {
    NASR_CUDA(cudaFuncSetAttribute(ctc_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kSmemBudget));
    attr_set = true;
  }
  ctc_loss_grad_kernel<<<B, pl.threads, pl.smem, stream>>>(p);
  count_launch();
  NASR_CUDA(cudaGetLastError());
  return NASR_OK;
}

}  // namespace nasr
