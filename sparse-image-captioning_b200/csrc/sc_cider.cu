// CIDEr-D reward of self-critical sequence training on the device.
// Reference: sparse_caption/scst/cider/pyciderevalcap/ciderD/ciderD_scorer.py:133-212 (compute_cider: counts2vec + sim) as
// called by CaptionScorer (sparse_caption/scst/scorers.py:47-114) on decoded sample / baseline captions.
//
// The reference detokenises every rollout to a string on the host, splits it into words and scores it in Python dictionaries
// (B x (samples + 1) captions per step).  Here the rollouts stay on the device as word ids: an n-gram (n <= 4) of ids below
// 65536 is packed exactly into one 64-bit key; the document-frequency table (coco-train-words.p in the reference) and the
// tf-idf vectors of the reference captions - both static over training - are uploaded once as sorted key arrays.
// One CTA scores one hypothesis against the references of its image.  All arithmetic is double precision in the reference's
// own summation order (reductions are sequential in first-occurrence n-gram order, like the dict iteration of the reference),
// so scores agree with the Python scorer to the last bits.
#include "sc_common.cuh"

namespace {

constexpr int kMaxWords = 64;
constexpr int kMaxNgrams = 4 * kMaxWords;
constexpr int kThreads = 128;

struct CiderArgs {
  const int* hyp; int L, H;
  const int* hyp_img;
  int eos, pad;
  const unsigned long long* df_keys; const double* df_log; long n_df;   // sorted; df_log = log(max(1, df))
  double ref_len, sigma;
  const long* img_ref_off;      // [B + 1] -> reference captions of an image
  const long* ref_ng_off;       // [R + 1] -> n-gram entries of a reference
  const unsigned long long* ref_keys; const double* ref_vec;   // per reference: keys ascending, tf-idf weights
  const double* ref_norm;       // [R, 4]
  const int* ref_length;        // [R]   (the reference scorer's `length`: its bigram count)
  double* out;                  // [H]
};

__device__ __forceinline__ long lower_bound(const unsigned long long* a, long n, unsigned long long key) {
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads) ciderd_kernel(const CiderArgs a) {
  __shared__ int s_words[kMaxWords];
  __shared__ unsigned long long s_key[kMaxNgrams];
  __shared__ int s_tf[kMaxNgrams];      // 0 = not the first occurrence of its n-gram
  __shared__ int s_ord[kMaxNgrams];     // n-gram order - 1
  __shared__ double s_vec[kMaxNgrams], s_ref[kMaxNgrams];
  __shared__ double s_norm[4], s_score[4];
  __shared__ int s_nw, s_T;
  const int hI = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    // words of the caption: ids before the first <eos>, pads dropped (what tokenizer.decode + str.split leave)
    int n = 0;
    for (int i = 0; i < a.L && n < kMaxWords; ++i) {
      const int t = a.hyp[(size_t)hI * a.L + i];
      if (t == a.eos) break;
      if (t != a.pad) s_words[n++] = t;
    }
    s_nw = n;
    int T = 0;
    for (int k = 1; k <= 4; ++k) T += max(0, n - k + 1);
    s_T = T;
    for (int k = 0; k < 4; ++k) { s_norm[k] = 0.0; s_score[k] = 0.0; }
  }
  __syncthreads();
  const int nw = s_nw, T = s_T;
  // entries in the reference's dict insertion order: k = 1..4, position ascending
  for (int e = tid; e < T; e += kThreads) {
    int k = 1, base = 0;
    while (e >= base + (nw - k + 1)) { base += nw - k + 1; ++k; }
    const int i = e - base;
    unsigned long long key = 0;
    for (int j = 0; j < k; ++j) key |= (unsigned long long)(s_words[i + j] & 0xffff) << (16 * j);
    s_key[e] = key;
    s_ord[e] = k - 1;
  }
  __syncthreads();
  for (int e = tid; e < T; e += kThreads) {
    const unsigned long long key = s_key[e];
    int tf = 0; bool first = true;
    for (int j = 0; j < T; ++j) if (s_key[j] == key) { ++tf; if (j < e) first = false; }
    s_tf[e] = first ? tf : 0;
    double v = 0.0;
    if (first) {
      const long p = lower_bound(a.df_keys, a.n_df, key);
      const double dl = (p < a.n_df && a.df_keys[p] == key) ? a.df_log[p] : 0.0;
      v = (double)tf * (a.ref_len - dl);
    }
    s_vec[e] = v;
  }
  __syncthreads();
  if (tid == 0) {
    for (int e = 0; e < T; ++e) if (s_tf[e]) s_norm[s_ord[e]] += s_vec[e] * s_vec[e];
    for (int k = 0; k < 4; ++k) s_norm[k] = sqrt(s_norm[k]);
  }
  __syncthreads();
  const int len_h = max(0, nw - 1);   // the reference counts the hypothesis' bigrams (ciderD_scorer.py:157-158)
  const int img = a.hyp_img[hI];
  const long r0 = a.img_ref_off[img], r1 = a.img_ref_off[img + 1];
  for (long r = r0; r < r1; ++r) {
    const long g0 = a.ref_ng_off[r], gn = a.ref_ng_off[r + 1] - g0;
    for (int e = tid; e < T; e += kThreads) {
      double v = 0.0;
      if (s_tf[e]) {
        const long p = lower_bound(a.ref_keys + g0, gn, s_key[e]);
        if (p < gn && a.ref_keys[g0 + p] == s_key[e]) v = a.ref_vec[g0 + p];
      }
      s_ref[e] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double val[4] = {0.0, 0.0, 0.0, 0.0};
      for (int e = 0; e < T; ++e) if (s_tf[e]) val[s_ord[e]] += fmin(s_vec[e], s_ref[e]) * s_ref[e];
      const double delta = (double)(len_h - a.ref_length[r]);
      const double pen = pow(2.718281828459045, -(delta * delta) / (2.0 * a.sigma * a.sigma));
      for (int k = 0; k < 4; ++k) {
        const double nr = a.ref_norm[r * 4 + k];
        if (s_norm[k] != 0.0 && nr != 0.0) val[k] /= s_norm[k] * nr;
        val[k] *= pen;
        s_score[k] += val[k];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    double s = ((s_score[0] + s_score[1]) + s_score[2]) + s_score[3];
    s /= 4.0;
    s /= (double)(r1 - r0);
    s *= 10.0;
    a.out[hI] = s;
  }
}

}  // namespace

extern "C" {

int sc_ciderd_score(const int* hyp, int H, int L, const int* hyp_img, int eos, int pad, const unsigned long long* df_keys,
                    const double* df_log, long n_df, double ref_len, double sigma, const long* img_ref_off, const long* ref_ng_off,
                    const unsigned long long* ref_keys, const double* ref_vec, const double* ref_norm, const int* ref_length,
                    double* out, cudaStream_t stream) {
  SC_CHECK(H > 0 && L > 0 && L <= kMaxWords, SC_ERR_SHAPE, "sc_ciderd_score: H=%d L=%d (L <= %d)", H, L, kMaxWords);
  SC_CHECK(hyp && hyp_img && img_ref_off && ref_ng_off && ref_norm && ref_length && out, SC_ERR_SHAPE, "sc_ciderd_score: null argument");
  CiderArgs a;
  a.hyp = hyp; a.L = L; a.H = H; a.hyp_img = hyp_img; a.eos = eos; a.pad = pad; a.df_keys = df_keys; a.df_log = df_log; a.n_df = n_df;
  a.ref_len = ref_len; a.sigma = sigma; a.img_ref_off = img_ref_off; a.ref_ng_off = ref_ng_off; a.ref_keys = ref_keys;
  a.ref_vec = ref_vec; a.ref_norm = ref_norm; a.ref_length = ref_length; a.out = out;
  ciderd_kernel<<<H, kThreads, 0, stream>>>(a);
  SC_LAUNCH_CHECK("sc_ciderd_score");
  return SC_OK;
}

}  // extern "C"
