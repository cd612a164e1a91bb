/* oracle/kgw_oracle.c -- TEST INFRASTRUCTURE ONLY.  See kgw_oracle.h.
 *
 * Every function cites the reference lines it restates
 * (paths relative to msesia/knockoffgwas, snpknock2/src/).  The restatement
 * follows the reference's evaluation order exactly: `x * y * z` is (x*y)*z,
 * sums over k run k = 0..K-1, divisions are true divisions.
 */
#include "kgw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* chaplotype::operator[] (chaplotype.h:25): LSB-first bit in a byte row */
static inline int hap_bit(const kgw_oracle_problem *P, int i, int j) {
  return (P->H[(int64_t)i * P->row_bytes + (j >> 3)] >> (j & 7)) & 1;
}

/* Knockoffs::mutate (knockoffs.cpp:1626-1629) */
static inline double mutate(int h, double mut_rate) {
  return (h == 1) ? (1.0 - mut_rate) : mut_rate;
}

double kgw_oracle_next_uniform(kgw_oracle_rng *rng, int hap, int stream, int locus, int p, int w) {
  rng->n_draws++;
  double u;
  switch (rng->mode) {
    case KGW_ORACLE_RNG_INJECTED:
      u = rng->injected[(((int64_t)hap * (rng->n_windows > 1 ? rng->n_windows : 1) + (rng->n_windows > 1 ? w : 0)) * 3 + stream) * p + locus];
      break;
    case KGW_ORACLE_RNG_TAUS88:
      u = kgw_taus88_uniform(&rng->taus);
      break;
    default:
      /* window pass w in the upper bits of the stream word: overlapping padded passes of
       * knockoffs.cpp:1222-1226 never reuse a uniform (include/kgw.h) */
      u = kgw_oracle_philox_uniform(rng->seed, (uint32_t)hap, (uint32_t)stream | ((uint32_t)w << 2), (uint32_t)locus);
  }
  if (rng->record) rng->record[(((int64_t)hap * (rng->n_windows > 1 ? rng->n_windows : 1) + (rng->n_windows > 1 ? w : 0)) * 3 + stream) * p + locus] = u;
  return u;
}

/* weighted_choice (utils.cpp:690-698): S = sum left to right from 0.0, R = u*S,
 * first i with R < w[i] after successive subtraction, else the last index. */
int kgw_oracle_weighted_choice(const double *w, int n, double u) {
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += w[i];
  double R = u * sum;
  for (int i = 0; i < n; i++) {
    if (R < w[i]) return i;
    R -= w[i];
  }
  return n - 1;
}

/* compute_hmm_backward (knockoffs.cpp:1041-1075); backward is [p][K] */
static void hmm_backward(const kgw_oracle_problem *P, int i, const int32_t *ref, double *backward,
                         int j_start, int j_end) {
  const int K = P->K;
  const double alpha = 1.0 / K; /* knockoffs.cpp:75 */
  for (int k = 0; k < K; k++) backward[(int64_t)(j_end - 1) * K + k] = 1.0 / K; /* :1046 */
  for (int j = j_end - 2; j >= j_start; j--) {
    double backward_const = 0.0;
    const int h = hap_bit(P, i, j + 1);
    const double bj = P->b[j + 1];
    const double mut_j = P->mu[j + 1];
    const double *next = backward + (int64_t)(j + 1) * K;
    double *cur = backward + (int64_t)j * K;
    for (int k = 0; k < K; k++) { /* :1052-1060 */
      const double a_j_k = (1.0 - bj) * alpha;
      const int h_ref = hap_bit(P, ref[k], j + 1);
      if (h == h_ref) backward_const += next[k] * a_j_k * (1.0 - mut_j);
      else backward_const += next[k] * a_j_k * mut_j;
    }
    double backwardSum = 0.0;
    for (int k = 0; k < K; k++) { /* :1062-1068 */
      const int h_ref = hap_bit(P, ref[k], j + 1);
      if (h == h_ref) cur[k] = backward_const + bj * next[k] * (1.0 - mut_j);
      else cur[k] = backward_const + bj * next[k] * mut_j;
      backwardSum += cur[k];
    }
    for (int k = 0; k < K; k++) cur[k] /= backwardSum; /* :1069-1073 */
  }
}

static void window_pad(const kgw_oracle_problem *P, int w, int *j_start_pad, int *j_end_pad) {
  /* knockoffs.cpp:1222-1226 and :1276-1280 */
  *j_start_pad = P->win_start[w];
  *j_end_pad = P->win_end[w];
  if (w > 0) *j_start_pad = P->win_start[w - 1];
  if (w < P->W - 1) *j_end_pad = P->win_end[w + 1];
}

/* sample_posterior_MC (knockoffs.cpp:1215-1261) */
static void sample_posterior_MC(const kgw_oracle_problem *P, kgw_oracle_rng *rng, int i, int w,
                                int32_t *z, double *backward, double *weights) {
  const int K = P->K;
  const double alpha = 1.0 / K;
  const int32_t *ref = P->ref_local + ((int64_t)i * P->W + w) * K;
  int js, je;
  window_pad(P, w, &js, &je);
  hmm_backward(P, i, ref, backward, js, je);

  for (int k = 0; k < K; k++) { /* :1233-1242 */
    const double a_0_k = (1.0 - P->b[js]) * alpha;
    const double theta_0_k = mutate(hap_bit(P, ref[k], js), P->mu[js]);
    if (hap_bit(P, i, js) == 1) weights[k] = a_0_k * theta_0_k * backward[(int64_t)js * K + k];
    else weights[k] = a_0_k * (1.0 - theta_0_k) * backward[(int64_t)js * K + k];
  }
  z[js] = kgw_oracle_weighted_choice(weights, K, kgw_oracle_next_uniform(rng, i, KGW_STREAM_Z, js, P->p, w));

  for (int j = js + 1; j < je; j++) { /* :1245-1260 */
    const double bj = P->b[j];
    const int h = hap_bit(P, i, j);
    const double mut_j = P->mu[j];
    for (int k = 0; k < K; k++) {
      const double a_j_k = (1.0 - bj) * alpha;
      const double theta_j_k = mutate(hap_bit(P, ref[k], j), mut_j);
      if (h == 1)
        weights[k] = (a_j_k + bj * (double)(k == z[j - 1])) * theta_j_k * backward[(int64_t)j * K + k];
      else
        weights[k] = (a_j_k + bj * (double)(k == z[j - 1])) * (1.0 - theta_j_k) * backward[(int64_t)j * K + k];
    }
    z[j] = kgw_oracle_weighted_choice(weights, K, kgw_oracle_next_uniform(rng, i, KGW_STREAM_Z, j, P->p, w));
  }
}

/* sample_knockoff_MC (knockoffs.cpp:1270-1428).
 * grp_first[g] / grp_last[g] = elements[g].front() / .back() (groups are contiguous runs,
 * set_groups knockoffs.cpp:96-107). */
static void sample_knockoff_MC(const kgw_oracle_problem *P, kgw_oracle_rng *rng, int i, int w,
                               const int32_t *z, int32_t *zk, const int32_t *grp_first,
                               const int32_t *grp_last, double *weights, double *N, double *N_old,
                               double *vstar, double *vbar) {
  const int K = P->K;
  const double alpha = 1.0 / K;
  const double *b = P->b;
  int js, je;
  window_pad(P, w, &js, &je);
  const int g_start = P->groups[js];
  const int g_end = P->groups[je - 1] + 1;

  for (int k = 0; k < K; k++) { N[k] = 1.0; N_old[k] = 1.0; } /* :1293-1294 */
  const double N_min = 1.0e-10;

  for (int g = g_start; g < g_end; g++) {
    const int e0 = grp_first[g];
    const int groupSize = grp_last[g] - e0 + 1;
    /* vstar (:1298-1313); elements[g][t] = e0 + t */
    for (int t = groupSize - 1; t >= 0; t--) {
      if (t < groupSize - 1) vstar[t] = vstar[t + 1] * b[e0 + t + 1];
      else vstar[t] = (g < g_end - 1) ? b[e0] : 0.0;
    }
    /* vbar (:1315-1339); vbar is [groupSize][K] */
    for (int t = groupSize - 1; t >= 0; t--) {
      double sum_a = 0;
      if (t < groupSize - 1) {
        for (int l = 0; l < K; l++) sum_a += (1.0 - b[e0 + t + 1]) * alpha;
      }
      for (int x = 0; x < K; x++) {
        if (t < groupSize - 1) {
          vbar[(int64_t)t * K + x] = vbar[(int64_t)(t + 1) * K + x] * (sum_a + b[e0 + t + 1]);
          vbar[(int64_t)t * K + x] += vstar[t + 1] * (1.0 - b[e0 + t + 1]) * alpha;
        } else {
          if (g < g_end - 1) vbar[(int64_t)t * K + x] = (1.0 - b[grp_first[g + 1]]) * alpha;
          else vbar[(int64_t)t * K + x] = 1.0;
        }
      }
    }

    const int jg0 = e0;
    /* N_sum (:1345-1358) */
    double N_sum = 0;
    for (int k = 0; k < K; k++) {
      if (g == g_start) {
        N_sum += (1.0 - b[jg0]) * alpha;
      } else {
        const int z0 = z[grp_last[g - 1]];
        const int z0k = zk[grp_last[g - 1]];
        double tmp = ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0));
        tmp = tmp * ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0k));
        N_sum += tmp / N_old[k];
      }
    }
    /* partition function update (:1360-1376) */
    for (int k = 0; k < K; k++) {
      if (g < g_end - 1) {
        N[k] += vbar[k] * N_sum; /* vbar[0][k] */
        if (g == 0) {
          N[k] += vstar[0] * (1.0 - b[jg0]) * alpha;
        } else {
          const int z0 = z[grp_last[g - 1]];
          const int z0k = zk[grp_last[g - 1]];
          double tmp = ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0));
          tmp = tmp * ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0k));
          N[k] += vstar[0] * tmp / N_old[k];
        }
      }
    }
    /* normalise + clamp (:1378-1390) */
    double N_norm = 0;
    for (int k = 0; k < K; k++) {
      if (N[k] < N_min) N[k] = N_min;
      N_norm += N[k];
    }
    for (int k = 0; k < K; k++) {
      N[k] /= N_norm;
      if (N[k] < N_min) N[k] = N_min;
    }
    /* sampling weights (:1392-1422) */
    for (int t = 0; t < groupSize; t++) {
      const int j = e0 + t;
      for (int k = 0; k < K; k++) {
        double wk = 1.0;
        if (t > 0) {
          const int z0k = zk[j - 1];
          wk *= ((1.0 - b[j]) * alpha + b[j] * (double)(k == z0k));
        } else {
          if (g == 0) {
            wk *= (1.0 - b[jg0]) * alpha;
          } else {
            const int z0 = z[grp_last[g - 1]];
            const int z0k = zk[grp_last[g - 1]];
            wk *= ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0));
            wk *= ((1.0 - b[jg0]) * alpha + b[jg0] * (double)(k == z0k));
          }
          wk /= N_old[k];
        }
        if (g == g_end - 1) {
          wk *= vbar[(int64_t)t * K + 0];
        } else {
          const int z1 = z[grp_first[g + 1]];
          wk *= (vbar[(int64_t)t * K + z1] + vstar[t] * (double)(k == z1));
        }
        weights[k] = wk;
      }
      zk[j] = kgw_oracle_weighted_choice(weights, K, kgw_oracle_next_uniform(rng, i, KGW_STREAM_ZK, j, P->p, w));
    }
    for (int k = 0; k < K; k++) N_old[k] = N[k]; /* :1424-1426 */
  }
}

/* sample_emission_copula (knockoffs.cpp:1469-1502), lambda = 0.5 at the call site (:677) */
static void sample_emission_copula(const kgw_oracle_problem *P, kgw_oracle_rng *rng, int i, int w,
                                   const int32_t *z, const int32_t *zk, const int32_t *ref,
                                   int32_t *out, double lambda, int j_start, int j_end) {
  double wb[2];
  for (int j = j_start; j < j_end; j++) {
    const int h_predicted = hap_bit(P, ref[z[j]], j);
    const int h_predicted_ko = hap_bit(P, ref[zk[j]], j);
    const double p_orig1 = mutate(h_predicted, P->mu[j]);
    const double p_ko1 = mutate(h_predicted_ko, P->mu[j]);
    double temp = p_orig1 - (1.0 - p_ko1);
    double temp2 = 0.0;
    if (temp < 0) { temp2 = -temp; temp = 0.0; }
    if (hap_bit(P, i, j) == 0) {
      wb[1] = 1.0 - temp2 / (1.0 - p_orig1);
      wb[0] = temp2 / (1.0 - p_orig1);
    } else {
      wb[1] = temp / (p_orig1);
      wb[0] = 1.0 - temp / (p_orig1);
    }
    wb[0] = lambda * wb[0] + (1 - lambda) * (1 - p_ko1);
    wb[1] = lambda * wb[1] + (1 - lambda) * p_ko1;
    out[j] = kgw_oracle_weighted_choice(wb, 2, kgw_oracle_next_uniform(rng, i, KGW_STREAM_HK, j, P->p, w));
  }
}

int kgw_oracle_unrelated(const kgw_oracle_problem *P, kgw_oracle_rng *rng, const int32_t *haps,
                         int32_t n, int32_t *z_out, int32_t *zk_out, uint8_t *hk_out, double *beta_out) {
  const int p = P->p, K = P->K;
  if (p <= 0 || K <= 0 || P->W <= 0) return -1;
  const int G = P->groups[p - 1] + 1;
  int32_t *grp_first = (int32_t *)malloc(sizeof(int32_t) * G);
  int32_t *grp_last = (int32_t *)malloc(sizeof(int32_t) * G);
  int max_group = 0;
  for (int g = 0; g < G; g++) { grp_first[g] = -1; grp_last[g] = -1; }
  for (int j = 0; j < p; j++) {
    const int g = P->groups[j];
    if (g < 0 || g >= G || (j > 0 && (g < P->groups[j - 1] || g > P->groups[j - 1] + 1))) {
      free(grp_first); free(grp_last);
      return -2; /* groups must be contiguous, non-decreasing, dense */
    }
    if (grp_first[g] < 0) grp_first[g] = j;
    grp_last[g] = j;
  }
  for (int g = 0; g < G; g++) {
    const int m = grp_last[g] - grp_first[g] + 1;
    if (m > max_group) max_group = m;
  }
  /* worker scratch (knockoffs.cpp:645-652): allocated once, persists across haplotypes */
  int32_t *z = (int32_t *)calloc(p, sizeof(int32_t));
  int32_t *zk = (int32_t *)calloc(p, sizeof(int32_t));
  int32_t *hk_tmp = (int32_t *)calloc(p, sizeof(int32_t));
  int32_t *hk = (int32_t *)calloc(p, sizeof(int32_t));
  double *backward = (double *)malloc(sizeof(double) * (size_t)p * K);
  for (int64_t t = 0; t < (int64_t)p * K; t++) backward[t] = 1.0 / K;
  double *weights = (double *)malloc(sizeof(double) * K);
  double *N = (double *)malloc(sizeof(double) * K);
  double *N_old = (double *)malloc(sizeof(double) * K);
  double *vstar = (double *)malloc(sizeof(double) * max_group);
  double *vbar = (double *)malloc(sizeof(double) * (size_t)max_group * K);

  for (int32_t t = 0; t < n; t++) {
    const int i = haps[t];
    for (int w = 0; w < P->W; w++) { /* :654-681 */
      const int32_t *ref = P->ref_local + ((int64_t)i * P->W + w) * K;
      sample_posterior_MC(P, rng, i, w, z, backward, weights);
      sample_knockoff_MC(P, rng, i, w, z, zk, grp_first, grp_last, weights, N, N_old, vstar, vbar);
      sample_emission_copula(P, rng, i, w, z, zk, ref, hk_tmp, 0.5, P->win_start[w], P->win_end[w]);
      for (int j = P->win_start[w]; j < P->win_end[w]; j++) hk[j] = hk_tmp[j];
    }
    if (z_out) memcpy(z_out + (int64_t)t * p, z, sizeof(int32_t) * p);
    if (zk_out) memcpy(zk_out + (int64_t)t * p, zk, sizeof(int32_t) * p);
    if (beta_out) memcpy(beta_out + (int64_t)t * p * K, backward, sizeof(double) * (size_t)p * K);
    if (hk_out) { /* set_Hk -> chaplotype(vector<int>) (knockoffs.cpp:299, chaplotype.cpp:25-29) */
      uint8_t *row = hk_out + (int64_t)t * P->row_bytes;
      memset(row, 0, (size_t)P->row_bytes);
      for (int j = 0; j < p; j++)
        if (hk[j]) row[j >> 3] |= (uint8_t)(1u << (j & 7));
    }
  }
  free(grp_first); free(grp_last); free(z); free(zk); free(hk_tmp); free(hk); free(backward);
  free(weights); free(N); free(N_old); free(vstar); free(vbar);
  return 0;
}

/* ---- --estimate-hmm (restated ahead of the CUDA work) ---------------------------------------
 * Knockoffs::EM / EM_step / EM_worker (knockoffs.cpp:706-960), single-threaded.  As the reference stands, the
 * recombination scale is NOT updated (`new_rho = rho`, :929-930; EM_solve_rho is commented out) and the Xi
 * accumulation (:803-830) feeds nothing, so one EM step is: for every haplotype, backward and forward passes
 * (compute_hmm_backward :1041-1075, compute_hmm_forward :1087-1125), posterior marginals (:777-787), expected
 * mismatches gamma_j (:789-798); then mu_j := clamp(mean gamma_j, 1e-6, 1e-3) (:917-926, update_hmm :1550-1561).
 * Iterations stop when the root-mean-square relative change of mu falls below 5e-2, or after 20 (:716-736).
 * mu_init: what Knockoffs::init_hmm(K, rho) leaves (:1527-1548; compute_maf, haplotypes.cpp:101-111, adds 1.0 for
 * every haplotype whatever its allele, so maf_j = min(1, 1 - 1) and every mu_j starts at the lower clamp 1e-5). */
static void hmm_forward(const kgw_oracle_problem *P, int i, const int32_t *ref, double *forward, int j_start, int j_end) {
  const int K = P->K;
  const double alpha = 1.0 / K;
  double forwardSum_0 = 0.0;
  const int h0 = hap_bit(P, i, j_start);
  double *f0 = forward + (int64_t)j_start * K;
  for (int k = 0; k < K; k++) {
    const int h0_ref = hap_bit(P, ref[k], j_start);
    if (h0 == h0_ref) f0[k] = alpha * (1.0 - P->mu[j_start]);
    else f0[k] = alpha * P->mu[j_start];
    forwardSum_0 += f0[k];
  }
  for (int k = 0; k < K; k++) f0[k] /= forwardSum_0;
  for (int j = j_start + 1; j < j_end; j++) {
    const int h = hap_bit(P, i, j);
    const double bj = P->b[j];
    const double mut_j = P->mu[j];
    double forwardSum = 0.0;
    const double *prev = forward + (int64_t)(j - 1) * K;
    double *cur = forward + (int64_t)j * K;
    for (int k = 0; k < K; k++) {
      const double a_j_k = (1.0 - bj) * alpha;
      cur[k] = a_j_k + bj * prev[k];
      const int h_ref = hap_bit(P, ref[k], j);
      if (h == h_ref) cur[k] *= (1.0 - mut_j);
      else cur[k] *= mut_j;
      forwardSum += cur[k];
    }
    for (int k = 0; k < K; k++) cur[k] /= forwardSum;
  }
}

int kgw_oracle_em(const kgw_oracle_problem *P0, const double *mu_init, double *mu_out, int32_t *iterations, double *deltas) {
  const int N = P0->N, p = P0->p, K = P0->K;
  kgw_oracle_problem P = *P0;
  double *mu = (double *)malloc(sizeof(double) * (size_t)p);
  double *gamma = (double *)malloc(sizeof(double) * (size_t)p);
  double *backward = (double *)malloc(sizeof(double) * (size_t)p * K);
  double *forward = (double *)malloc(sizeof(double) * (size_t)p * K);
  double *marginal = (double *)malloc(sizeof(double) * (size_t)K);
  if (!mu || !gamma || !backward || !forward || !marginal) return -1;
  memcpy(mu, mu_init, sizeof(double) * (size_t)p);
  P.mu = mu;
  const double delta_min = 5e-2;
  const int it_max = 20;
  const double n_inv = 1.0 / (double)N;
  int it = 0, converged = 0;
  for (it = 0; it < it_max; it++) {
    for (int j = 0; j < p; j++) gamma[j] = 0.0;
    for (int i = 0; i < N; i++) { /* every haplotype, related or not (:862-864) */
      for (int w = 0; w < P.W; w++) {
        const int32_t *ref = P.ref_local + ((int64_t)i * P.W + w) * K;
        int js, je;
        window_pad(&P, w, &js, &je);
        hmm_backward(&P, i, ref, backward, js, je);
        hmm_forward(&P, i, ref, forward, js, je);
        for (int j = P.win_start[w]; j < P.win_end[w]; j++) {
          double marginal_sum = 0;
          for (int k = 0; k < K; k++) {
            marginal[k] = backward[(int64_t)j * K + k] * forward[(int64_t)j * K + k];
            marginal_sum += marginal[k];
          }
          for (int k = 0; k < K; k++) marginal[k] /= marginal_sum;
          const int h_real = hap_bit(&P, i, j);
          for (int k = 0; k < K; k++)
            if (h_real != hap_bit(&P, ref[k], j)) gamma[j] += marginal[k];
        }
      }
    }
    double delta_mutation = 0;
    for (int j = 0; j < p; j++) {
      double g = 0.0 + gamma[j] * n_inv; /* :908-914 with one worker */
      g = (g > 1e-6) ? g : 1e-6;         /* std::max(1e-6, gamma) */
      g = (g < 1e-3) ? g : 1e-3;         /* std::min(1e-3, gamma) */
      const double denom = (mu[j] > 1e-5) ? mu[j] : 1e-5;
      const double delta_j = fabs(g - mu[j]) / denom;
      delta_mutation += delta_j * delta_j;
      gamma[j] = g;
    }
    delta_mutation = sqrt(delta_mutation / (double)p);
    const double delta_parameters = (0.0 > delta_mutation) ? 0.0 : delta_mutation; /* max(delta_recomb = 0, .) */
    for (int j = 0; j < p; j++) mu[j] = gamma[j];
    if (deltas) deltas[it] = delta_parameters;
    if (delta_parameters < delta_min) { converged = 1; break; }
  }
  memcpy(mu_out, mu, sizeof(double) * (size_t)p);
  if (iterations) *iterations = converged ? it + 1 : it;
  free(mu); free(gamma); free(backward); free(forward); free(marginal);
  return converged ? 0 : 1;
}
