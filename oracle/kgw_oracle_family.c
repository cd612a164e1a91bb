/* oracle/kgw_oracle_family.c -- TEST INFRASTRUCTURE ONLY.  See kgw_oracle.h.
 *
 * CPU restatement of the IBD-conditioned branch of snpknock2's knockoff sampler:
 *   Knockoffs::generate_related_cluster   snpknock2/src/knockoffs.cpp:465-571
 *   FamilyBP                              snpknock2/src/family_bp.cpp:8-626
 *   FamilyKnockoffs                       snpknock2/src/family_knockoffs.cpp:8-454
 *   Knockoffs::sample_emission            snpknock2/src/knockoffs.cpp:1430-1441
 * IbdCluster::tidy (ibd.cpp:99-189) is host pre-processing that stays with the unchanged
 * reference code: this file takes the TIDIED segments of a family as input, in the iteration
 * order of the reference's std::set<IbdSeg> (ibd.cpp:482-509).
 * Strict IEEE double, reference evaluation order, every sum over k sequential.
 */
#include "kgw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int hap_bit(const kgw_oracle_problem *P, int i, int j) {
  return (P->H[(int64_t)i * P->row_bytes + (j >> 3)] >> (j & 7)) & 1;
}
static inline double mutate(int h, double mut_rate) { return (h == 1) ? (1.0 - mut_rate) : mut_rate; }
/* emission_prob (family_bp.cpp:628-631) */
static inline double emission_prob(int h1, int h2, double mut_rate) { return (h1 == h2) ? 1.0 - mut_rate : mut_rate; }

typedef struct {
  const kgw_oracle_problem *P;
  kgw_oracle_rng *rng;
  int n, p, K;
  const int32_t *idx;   /* absolute ids */
  int NB;               /* capacity of a neighbour list */
  int32_t *nb_cnt;      /* [n][p] */
  int32_t *nb;          /* [n][p][NB]   neighbors[i][j], push_back order (family_bp.cpp:41-61) */
  uint8_t *inactive;    /* [n][p]       inactive_nodes (family_bp.cpp:561-567) */
  double *mr, *ml;      /* [n][p][K]    messages_right / messages_left */
  double *message, *f_j, *weights;
  int64_t bp_sweeps;
  int bp_runs;          /* run_bp calls so far (audit only) */
} fam_ctx;

#define NBC(c, i, j) ((c)->nb_cnt[(int64_t)(i) * (c)->p + (j)])
#define NBV(c, i, j) ((c)->nb + ((int64_t)(i) * (c)->p + (j)) * (c)->NB)
#define MR(c, i, j) ((c)->mr + ((int64_t)(i) * (c)->p + (j)) * (c)->K)
#define ML(c, i, j) ((c)->ml + ((int64_t)(i) * (c)->p + (j)) * (c)->K)

static int local_index(const kgw_oracle_family *F, int id) { /* std::find over `indices` */
  for (int i = 0; i < F->n_members; i++)
    if (F->members[i] == id) return i;
  return F->n_members;
}

/* neighbors[i1][j] for every segment, member and locus (family_bp.cpp:41-61, family_knockoffs.cpp:40-58) */
static int build_neighbors(fam_ctx *c, const kgw_oracle_family *F) {
  const int n = c->n, p = c->p;
  c->nb_cnt = (int32_t *)calloc((size_t)n * p, sizeof(int32_t));
  for (int s = 0; s < F->n_segs; s++) {
    const int m = F->seg_off[s + 1] - F->seg_off[s];
    for (int a = 0; a < m; a++) {
      const int i1 = local_index(F, F->seg_ids[F->seg_off[s] + a]);
      if (i1 >= n) return -3;
      for (int j = F->seg_jmin[s]; j <= F->seg_jmax[s]; j++) c->nb_cnt[(int64_t)i1 * p + j] += m - 1;
    }
  }
  int NB = 1;
  for (int64_t t = 0; t < (int64_t)n * p; t++)
    if (c->nb_cnt[t] > NB) NB = c->nb_cnt[t];
  c->NB = NB;
  c->nb = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * p * NB);
  memset(c->nb_cnt, 0, sizeof(int32_t) * (size_t)n * p);
  for (int s = 0; s < F->n_segs; s++) {
    const int m = F->seg_off[s + 1] - F->seg_off[s];
    for (int a = 0; a < m; a++) {
      const int i1 = local_index(F, F->seg_ids[F->seg_off[s] + a]);
      for (int j = F->seg_jmin[s]; j <= F->seg_jmax[s]; j++) {
        for (int bq = 0; bq < m; bq++) {
          const int i2 = local_index(F, F->seg_ids[F->seg_off[s] + bq]);
          if (i1 != i2) NBV(c, i1, j)[NBC(c, i1, j)++] = i2;
        }
      }
    }
  }
  return 0;
}

static int nb_differ(const fam_ctx *c, int i, int ja, int jb) { /* vector<int> operator!= */
  const int na = NBC(c, i, ja), nbn = NBC(c, i, jb);
  if (na != nbn) return 1;
  return memcmp(NBV(c, i, ja), NBV(c, i, jb), sizeof(int32_t) * (size_t)na) != 0;
}

/* std::set_difference(A, B) exactly as the algorithm is specified (inputs used as they are) */
static int set_difference(const int32_t *A, int na, const int32_t *B, int nbn, int32_t *out) {
  int ia = 0, ib = 0, no = 0;
  while (ia < na && ib < nbn) {
    if (A[ia] < B[ib]) out[no++] = A[ia++];
    else if (B[ib] < A[ia]) ib++;
    else { ia++; ib++; }
  }
  while (ia < na) out[no++] = A[ia++];
  return no;
}

/* junction test + neighbour differences shared by run_bp (family_bp.cpp:352-378) and
 * sample_posterior (:168-193).  Returns is_junction as computed BEFORE the "really a junction" check. */
static int neighbor_diffs(const fam_ctx *c, int i, int j, int32_t *ndl, int *nl, int32_t *ndr, int *nr) {
  const int p = c->p;
  int is_junction = 0;
  *nl = 0;
  *nr = 0;
  if (j < p - 1 && nb_differ(c, i, j, j + 1)) is_junction = 1;
  if (j > 0 && nb_differ(c, i, j, j - 1)) is_junction = 1;
  if (is_junction) {
    if (j > 0 && nb_differ(c, i, j, j - 1))
      *nl = set_difference(NBV(c, i, j), NBC(c, i, j), NBV(c, i, j - 1), NBC(c, i, j - 1), ndl);
    if (j < p - 1 && nb_differ(c, i, j, j + 1))
      *nr = set_difference(NBV(c, i, j), NBC(c, i, j), NBV(c, i, j + 1), NBC(c, i, j + 1), ndr);
  }
  return is_junction;
}

/* emission of node (i1, j) for every state: geometric mean over the node and its neighbours
 * (family_bp.cpp:437-454, :509-526) */
static void node_emission(const fam_ctx *c, int i1, int j, double *f_j) {
  const kgw_oracle_problem *P = c->P;
  const int K = c->K;
  const int i1_abs = c->idx[i1];
  const int32_t *ref = P->ref_global + (int64_t)i1_abs * K;
  const int h_real = hap_bit(P, i1_abs, j);
  const double mu = P->mu[j];
  for (int k = 0; k < K; k++) f_j[k] = emission_prob(h_real, hap_bit(P, ref[k], j), mu);
  const int nn = NBC(c, i1, j);
  for (int t = 0; t < nn; t++) {
    const int h2_real = hap_bit(P, c->idx[NBV(c, i1, j)[t]], j);
    for (int k = 0; k < K; k++) f_j[k] *= emission_prob(h2_real, hap_bit(P, ref[k], j), mu);
  }
  if (nn > 0) {
    const double num_neighbors = nn + 1;
    for (int k = 0; k < K; k++) f_j[k] = pow(f_j[k], 1.0 / num_neighbors);
  }
}

/* send_messages_right (family_bp.cpp:429-499) */
static double send_messages_right(fam_ctx *c, int i1, int j, const int32_t *ndl, int nl, const int32_t *ndr, int nr) {
  const int K = c->K, p = c->p;
  if (j >= p - 1) return 0;
  double *message = c->message, *f_j = c->f_j;
  node_emission(c, i1, j, f_j);
  const double b_j = c->P->b[j + 1];
  const double a_j = (1.0 - b_j) / (double)K;
  double const_msg = 0;
  for (int l = 0; l < K; l++) {
    double v = a_j * f_j[l];
    if (j > 0) v *= MR(c, i1, j - 1)[l];
    for (int t = 0; t < nl; t++) v *= MR(c, ndl[t], j - 1)[l];
    for (int t = 0; t < nr; t++) v *= ML(c, ndr[t], j + 1)[l];
    const_msg += v;
  }
  double message_sum = 0;
  for (int k = 0; k < K; k++) {
    double v = b_j * f_j[k];
    if (j > 0) v *= MR(c, i1, j - 1)[k];
    for (int t = 0; t < nl; t++) v *= MR(c, ndl[t], j - 1)[k];
    for (int t = 0; t < nr; t++) v *= ML(c, ndr[t], j + 1)[k];
    message[k] = const_msg + v;
    message_sum += message[k];
  }
  double diff = 0;
  const double inv = 1.0 / message_sum;
  for (int k = 0; k < K; k++) {
    message[k] = message[k] * inv;
    diff += fabs(message[k] - MR(c, i1, j)[k]);
  }
  return diff;
}

/* send_messages_left (family_bp.cpp:501-559) */
static double send_messages_left(fam_ctx *c, int i1, int j, const int32_t *ndl, int nl, const int32_t *ndr, int nr) {
  const int K = c->K, p = c->p;
  if (j <= 0) return 0;
  double *message = c->message, *f_j = c->f_j;
  node_emission(c, i1, j, f_j);
  const double b_j = c->P->b[j];
  const double a_j = (1.0 - b_j) / (double)K;
  double const_msg = 0;
  for (int l = 0; l < K; l++) {
    double v = a_j * f_j[l];
    if (j < p - 1) v *= ML(c, i1, j + 1)[l];
    for (int t = 0; t < nr; t++) v *= ML(c, ndr[t], j + 1)[l];
    for (int t = 0; t < nl; t++) v *= MR(c, ndl[t], j - 1)[l];
    const_msg += v;
  }
  double message_sum = 0;
  for (int k = 0; k < K; k++) {
    double v = b_j * f_j[k];
    if (j < p - 1) v *= ML(c, i1, j + 1)[k];
    for (int t = 0; t < nr; t++) v *= ML(c, ndr[t], j + 1)[k];
    for (int t = 0; t < nl; t++) v *= MR(c, ndl[t], j - 1)[k];
    message[k] = const_msg + v;
    message_sum += message[k];
  }
  double diff = 0;
  const double inv = 1.0 / message_sum;
  for (int k = 0; k < K; k++) {
    message[k] = message[k] * inv;
    diff += fabs(message[k] - ML(c, i1, j)[k]);
  }
  return diff;
}

/* ---- audit of the CUDA kernel's work-skipping rule (knockoffgwas_b200/csrc/kgw_family.cuh) --------------------
 * The kernel recomputes a BP message only (1) in the first iteration of the first run_bp, (2) at the "special"
 * loci [b-2, b+1] around every locus b where some member's neighbour list changes, (3) when the message one locus
 * back on the same chain changed bits earlier in the same sweep.  With the audit on, run_bp below still computes
 * EVERY message exactly as the reference does, and counts the messages the rule would have skipped and, among
 * them, those whose recomputation did not return the stored bits ("violations": must stay 0).  Results of the
 * oracle are unaffected. */
static int g_audit = 0;
static int64_t g_audit_stats[3]; /* messages, skippable, violations */
void kgw_oracle_family_audit(int enable) {
  g_audit = enable;
  g_audit_stats[0] = g_audit_stats[1] = g_audit_stats[2] = 0;
}
void kgw_oracle_family_audit_get(int64_t out[3]) {
  out[0] = g_audit_stats[0];
  out[1] = g_audit_stats[1];
  out[2] = g_audit_stats[2];
}

/* run_bp (family_bp.cpp:318-419) */
static int run_bp(fam_ctx *c) {
  const int it_max = 20;
  const int n = c->n, p = c->p, K = c->K;
  int32_t *ndl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->NB + 1));
  int32_t *ndr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->NB + 1));
  uint8_t *special = NULL;
  if (g_audit) {
    special = (uint8_t *)calloc((size_t)p, 1);
    for (int b = 1; b < p; b++) {
      int boundary = 0;
      for (int i = 0; i < n; i++) boundary |= nb_differ(c, i, b, b - 1);
      if (boundary)
        for (int j = b - 2; j <= b + 1; j++)
          if (j >= 0 && j < p) special[j] = 1;
    }
  }
  const int first_run = (c->bp_runs++ == 0);
  int converged = 0, it = 0;
  while (!converged && (it++ < it_max)) {
    double diff = 0;
    const int full = first_run && it == 1;
    for (int dir = 0; dir < 2; dir++) { /* "right" then "left" */
      for (int k = 0; k < K; k++) { c->message[k] = 1.0; c->f_j[k] = 1.0; } /* :350-351 */
      double diff_tmp = 0;
      unsigned changed_prev = 0; /* audit: members whose message one locus back changed bits in this sweep */
      for (int jj = 0; jj < p; jj++) {
        const int j = (dir == 0) ? jj : p - 1 - jj;
        unsigned changed_cur = 0;
        for (int i = 0; i < n; i++) {
          if (c->inactive[(int64_t)i * p + j]) continue;
          int nl, nr;
          neighbor_diffs(c, i, j, ndl, &nl, ndr, &nr);
          double *dst = NULL;
          double d = 0;
          if (dir == 0) {
            d = send_messages_right(c, i, j, ndl, nl, ndr, nr);
            /* the reference copies `message` even when send_messages_right returned early (j == p-1) */
            dst = MR(c, i, j);
          } else {
            if (j <= 0) continue;
            d = send_messages_left(c, i, j, ndl, nl, ndr, nr);
            dst = ML(c, i, j);
          }
          if (g_audit && !(dir == 0 && j >= p - 1)) {
            const int same = memcmp(dst, c->message, sizeof(double) * (size_t)K) == 0;
            const int rule_computes = full || special[j] || (i < 32 && ((changed_prev >> i) & 1u));
            g_audit_stats[0]++;
            if (!rule_computes) {
              g_audit_stats[1]++;
              if (!same || d != 0.0) g_audit_stats[2]++;
            }
            if (!same && i < 32) changed_cur |= 1u << i;
          }
          diff_tmp += d;
          memcpy(dst, c->message, sizeof(double) * (size_t)K);
        }
        changed_prev = changed_cur;
      }
      diff_tmp /= (double)((size_t)p * (size_t)n);
      diff += diff_tmp;
      c->bp_sweeps++;
    }
    if (diff <= 1e-3) converged = 1;
  }
  free(ndl);
  free(ndr);
  free(special);
  return converged ? it : -1;
}

/* condition_on (family_bp.cpp:236-255) */
static void condition_on(fam_ctx *c, int i, int j, int k, int forward) {
  const int K = c->K, p = c->p;
  const double *b = c->P->b;
  const double alpha = 1.0 / K; /* alpha_global[i_abs][l], knockoffs.cpp:77 */
  for (int l = 0; l < K; l++) {
    if (j < p - 1) {
      const double a_jl = (1.0 - b[j + 1]) * alpha;
      MR(c, i, j)[l] = a_jl + b[j + 1] * (double)(l == k);
    }
    if (!forward) {
      if (j > 0) {
        const double a_jk = (1.0 - b[j]) * alpha;
        ML(c, i, j)[l] = a_jk + b[j] * (double)(l == k);
      }
    }
  }
  c->inactive[(int64_t)i * p + j] = 1;
}

/* sample_posterior_joint (family_bp.cpp:569-601) */
static int sample_posterior_joint(fam_ctx *c, int i, int j, const int32_t *ndl, int nl, const int32_t *ndr, int nr) {
  const kgw_oracle_problem *P = c->P;
  const int K = c->K, p = c->p;
  const int i_abs = c->idx[i];
  const int32_t *ref = P->ref_global + (int64_t)i_abs * K;
  const int h_real = hap_bit(P, i_abs, j);
  const double mu = P->mu[j];
  const int nn = NBC(c, i, j);
  double *w = c->weights;
  for (int k = 0; k < K; k++) {
    w[k] = emission_prob(h_real, hap_bit(P, ref[k], j), mu);
    for (int t = 0; t < nn; t++) {
      const int h2_real = hap_bit(P, c->idx[NBV(c, i, j)[t]], j);
      w[k] *= emission_prob(h2_real, hap_bit(P, ref[k], j), mu);
    }
    if (nn > 0) {
      const double num_neighbors = nn + 1;
      w[k] = pow(w[k], 1.0 / num_neighbors);
    }
    if (j > 0) {
      w[k] *= MR(c, i, j - 1)[k];
      for (int t = 0; t < nl; t++) w[k] *= MR(c, ndl[t], j - 1)[k];
    }
    if (j < p - 1) {
      w[k] *= ML(c, i, j + 1)[k];
      for (int t = 0; t < nr; t++) w[k] *= ML(c, ndr[t], j + 1)[k];
    }
  }
  return kgw_oracle_weighted_choice(w, K, kgw_oracle_next_uniform(c->rng, i_abs, KGW_STREAM_Z, j, p, 0));
}

/* FamilyBP::sample_posterior (family_bp.cpp:160-230) */
static void sample_posterior(fam_ctx *c, int32_t *Z) {
  const int n = c->n, p = c->p;
  int32_t *ndl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->NB + 1));
  int32_t *ndr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c->NB + 1));
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < p; j++) {
      if (c->inactive[(int64_t)i * p + j]) continue;
      int nl, nr;
      int is_junction = neighbor_diffs(c, i, j, ndl, &nl, ndr, &nr);
      if (is_junction && nl == 0 && nr == 0) is_junction = 0;
      if (is_junction) {
        const int k = sample_posterior_joint(c, i, j, ndl, nl, ndr, nr);
        Z[(int64_t)i * p + j] = k;
        condition_on(c, i, j, k, 0);
        for (int t = 0; t < NBC(c, i, j); t++) {
          const int i2 = NBV(c, i, j)[t];
          Z[(int64_t)i2 * p + j] = k;
          condition_on(c, i2, j, k, 0);
        }
        run_bp(c);
      }
    }
  }
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < p; j++) {
      if (c->inactive[(int64_t)i * p + j]) continue;
      const int k = sample_posterior_joint(c, i, j, ndl, 0, ndr, 0);
      Z[(int64_t)i * p + j] = k;
      condition_on(c, i, j, k, 1);
      for (int t = 0; t < NBC(c, i, j); t++) {
        const int i2 = NBV(c, i, j)[t];
        Z[(int64_t)i2 * p + j] = k;
        condition_on(c, i2, j, k, 1);
      }
    }
  }
  free(ndl);
  free(ndr);
}

/* FamilyKnockoffs::sample_knockoff_MC (family_knockoffs.cpp:295-454); groups g_begin..g_end inclusive */
static void family_knockoff_MC(fam_ctx *c, int i, const int32_t *z, int32_t *zk, int g_begin, int g_end,
                               const int32_t *grp_first, const int32_t *grp_last, int num_groups,
                               double *N, double *N_old, double *vstar, double *vbar) {
  const int K = c->K, p = c->p;
  const double *b = c->P->b;
  const double alpha = 1.0 / K;
  double *weights = c->weights;
  const int i_abs = c->idx[i];
  for (int k = 0; k < K; k++) { N[k] = 1.0; N_old[k] = 1.0; }
  const double N_min = 1.0e-10;
  for (int g = g_begin; g <= g_end; g++) {
    const int e0 = grp_first[g];
    const int groupSize = grp_last[g] - e0 + 1;
    for (int t = groupSize - 1; t >= 0; t--) { /* vstar :315-328 */
      if (t < groupSize - 1) vstar[t] = vstar[t + 1] * b[e0 + t + 1];
      else vstar[t] = (g < num_groups - 1) ? b[e0] : 0.0;
    }
    for (int t = groupSize - 1; t >= 0; t--) { /* vbar :331-355 */
      double sum_a = 0;
      if (t < groupSize - 1)
        for (int l = 0; l < K; l++) sum_a += (1.0 - b[e0 + t + 1]) * alpha;
      for (int x = 0; x < K; x++) {
        if (t < groupSize - 1) {
          vbar[(int64_t)t * K + x] = vbar[(int64_t)(t + 1) * K + x] * (sum_a + b[e0 + t + 1]);
          vbar[(int64_t)t * K + x] += vstar[t + 1] * (1.0 - b[e0 + t + 1]) * alpha;
        } else {
          if (g < num_groups - 1) vbar[(int64_t)t * K + x] = (1.0 - b[grp_first[g + 1]]) * alpha;
          else vbar[(int64_t)t * K + x] = 1.0;
        }
      }
    }
    double N_sum = 0; /* :358-375 */
    for (int k = 0; k < K; k++) {
      if (g == 0) {
        N_sum += (1.0 - b[e0]) * alpha;
      } else if (g == g_begin) {
        const int z0 = z[grp_last[g - 1]];
        const double tmp = ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
        N_sum += tmp / N_old[k];
      } else {
        const int z0 = z[grp_last[g - 1]];
        const int z0k = zk[grp_last[g - 1]];
        double tmp = ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
        tmp = tmp * ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0k));
        N_sum += tmp / N_old[k];
      }
    }
    for (int k = 0; k < K; k++) { /* :378-398 */
      if (g < num_groups - 1) {
        N[k] += vbar[k] * N_sum;
        if (g == 0) {
          N[k] += vstar[0] * (1.0 - b[e0]) * alpha;
        } else if (g == g_begin) {
          const int z0 = z[grp_last[g - 1]];
          const double tmp = ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
          N[k] += vstar[0] * tmp / N_old[k];
        } else {
          const int z0 = z[grp_last[g - 1]];
          const int z0k = zk[grp_last[g - 1]];
          double tmp = ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
          tmp = tmp * ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0k));
          N[k] += vstar[0] * tmp / N_old[k];
        }
      }
    }
    double N_norm = 0; /* :401-413 */
    for (int k = 0; k < K; k++) {
      if (N[k] < N_min) N[k] = N_min;
      N_norm += N[k];
    }
    for (int k = 0; k < K; k++) {
      N[k] /= N_norm;
      if (N[k] < N_min) N[k] = N_min;
    }
    for (int t = 0; t < groupSize; t++) { /* :416-446 */
      const int j = e0 + t;
      for (int k = 0; k < K; k++) {
        double wk = 1.0;
        if (t > 0) {
          const int z0k = zk[j - 1];
          wk *= ((1.0 - b[j]) * alpha + b[j] * (double)(k == z0k));
        } else {
          if (g == 0) {
            wk *= (1.0 - b[e0]) * alpha;
          } else if (g == g_begin) {
            const int z0 = z[grp_last[g - 1]];
            wk *= ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
          } else {
            const int z0 = z[grp_last[g - 1]];
            const int z0k = zk[grp_last[g - 1]];
            wk *= ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0));
            wk *= ((1.0 - b[e0]) * alpha + b[e0] * (double)(k == z0k));
          }
          wk /= N_old[k];
        }
        if (g == num_groups - 1) {
          wk *= vbar[(int64_t)t * K + 0];
        } else {
          const int z1 = z[grp_first[g + 1]];
          wk *= (vbar[(int64_t)t * K + z1] + vstar[t] * (double)(k == z1));
        }
        weights[k] = wk;
      }
      zk[j] = kgw_oracle_weighted_choice(weights, K, kgw_oracle_next_uniform(c->rng, i_abs, KGW_STREAM_ZK, j, p, 0));
    }
    for (int k = 0; k < K; k++) N_old[k] = N[k];
  }
}

/* FamilyKnockoffs::run (family_knockoffs.cpp:77-293).  Returns 0, -10 "Missing knockoff variants",
 * -11 "Inconsistency found on Zk" (the reference aborts, :272-291). */
static int family_knockoffs_run(fam_ctx *c, const kgw_oracle_family *F, const int32_t *Z, int32_t *Zk,
                                const int32_t *grp_first, const int32_t *grp_last, int num_groups, int max_group) {
  const int n = c->n, p = c->p, K = c->K;
  const int32_t *groups = c->P->groups;
  for (int64_t t = 0; t < (int64_t)n * p; t++) Zk[t] = -1;
  double *N = (double *)malloc(sizeof(double) * K), *N_old = (double *)malloc(sizeof(double) * K);
  double *vstar = (double *)malloc(sizeof(double) * max_group);
  double *vbar = (double *)malloc(sizeof(double) * (size_t)max_group * K);
  /* (the boundary_groups / boundary_neighbors lists of :113-139 are computed by the reference but never used) */
  /* trivial knockoffs in the groups holding the segment end points (:155-168) */
  for (int s = 0; s < F->n_segs; s++) {
    for (int a = F->seg_off[s]; a < F->seg_off[s + 1]; a++) {
      const int i = local_index(F, F->seg_ids[a]);
      const int g_min = groups[F->seg_jmin[s]], g_max = groups[F->seg_jmax[s]];
      for (int j = grp_first[g_min]; j <= grp_last[g_min]; j++) Zk[(int64_t)i * p + j] = Z[(int64_t)i * p + j];
      for (int j = grp_first[g_max]; j <= grp_last[g_max]; j++) Zk[(int64_t)i * p + j] = Z[(int64_t)i * p + j];
    }
  }
  /* one draw inside every segment for its first member, copied to the others (:172-193) */
  for (int s = 0; s < F->n_segs; s++) {
    const int i0 = local_index(F, F->seg_ids[F->seg_off[s]]);
    const int j_min = F->seg_jmin[s], j_max = F->seg_jmax[s];
    int g_begin = groups[j_min] + 1;
    if (groups[j_min] == 0) g_begin = 0;
    int g_end = groups[j_max] - 1;
    if (groups[j_max] == num_groups - 1) g_end = groups[j_max];
    if (g_end >= g_begin) {
      family_knockoff_MC(c, i0, Z + (int64_t)i0 * p, Zk + (int64_t)i0 * p, g_begin, g_end, grp_first, grp_last,
                         num_groups, N, N_old, vstar, vbar);
      for (int a = F->seg_off[s] + 1; a < F->seg_off[s + 1]; a++) {
        const int i = local_index(F, F->seg_ids[a]);
        for (int j = grp_first[g_begin]; j <= grp_last[g_end]; j++) Zk[(int64_t)i * p + j] = Zk[(int64_t)i0 * p + j];
      }
    }
  }
  /* complement ranges of every member (:196-270) */
  int32_t *gl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F->n_segs + 1));
  int32_t *gr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F->n_segs + 1));
  for (int i = 0; i < n; i++) {
    int m = 0;
    for (int s = 0; s < F->n_segs; s++) {
      int member = 0;
      for (int a = F->seg_off[s]; a < F->seg_off[s + 1]; a++)
        if (local_index(F, F->seg_ids[a]) == i) member = 1;
      if (member) { gl[m] = groups[F->seg_jmin[s]]; gr[m] = groups[F->seg_jmax[s]]; m++; }
    }
    /* std::sort of both lists independently (:210-213) */
    for (int a = 1; a < m; a++) {
      int v = gl[a], t = a;
      while (t > 0 && gl[t - 1] > v) { gl[t] = gl[t - 1]; t--; }
      gl[t] = v;
      v = gr[a]; t = a;
      while (t > 0 && gr[t - 1] > v) { gr[t] = gr[t - 1]; t--; }
      gr[t] = v;
    }
    int g_begin = 0, g_end = num_groups - 1;
    for (int q = 0; q < m; q++) {
      g_end = gl[q] - 1;
      if (g_end >= 0 && g_begin <= g_end)
        family_knockoff_MC(c, i, Z + (int64_t)i * p, Zk + (int64_t)i * p, g_begin, g_end, grp_first, grp_last,
                           num_groups, N, N_old, vstar, vbar);
      g_begin = gr[q] + 1;
    }
    if (g_end < num_groups) {
      g_end = num_groups - 1;
      if (g_begin < num_groups && g_begin <= g_end)
        family_knockoff_MC(c, i, Z + (int64_t)i * p, Zk + (int64_t)i * p, g_begin, g_end, grp_first, grp_last,
                           num_groups, N, N_old, vstar, vbar);
    }
  }
  free(gl); free(gr); free(N); free(N_old); free(vstar); free(vbar);
  for (int64_t t = 0; t < (int64_t)n * p; t++)
    if (Zk[t] == -1) return -10;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < p; j++)
      for (int t = 0; t < NBC(c, i, j); t++)
        if (Zk[(int64_t)i * p + j] != Zk[(int64_t)NBV(c, i, j)[t] * p + j]) return -11;
  return 0;
}

int kgw_oracle_family_run(const kgw_oracle_problem *P, kgw_oracle_rng *rng, const kgw_oracle_family *F,
                          int32_t *z_out, int32_t *zk_out, uint8_t *hk_out, int64_t *bp_sweeps) {
  const int p = P->p, K = P->K, n = F->n_members;
  if (p <= 0 || K <= 0 || n <= 0 || !P->ref_global) return -1;
  const int G = P->groups[p - 1] + 1;
  int32_t *grp_first = (int32_t *)malloc(sizeof(int32_t) * G), *grp_last = (int32_t *)malloc(sizeof(int32_t) * G);
  for (int g = 0; g < G; g++) grp_first[g] = grp_last[g] = -1;
  for (int j = 0; j < p; j++) {
    const int g = P->groups[j];
    if (grp_first[g] < 0) grp_first[g] = j;
    grp_last[g] = j;
  }
  int max_group = 0;
  for (int g = 0; g < G; g++)
    if (grp_last[g] - grp_first[g] + 1 > max_group) max_group = grp_last[g] - grp_first[g] + 1;

  fam_ctx c;
  memset(&c, 0, sizeof(c));
  c.P = P; c.rng = rng; c.n = n; c.p = p; c.K = K; c.idx = F->members;
  int rc = build_neighbors(&c, F);
  if (rc != 0) { free(grp_first); free(grp_last); free(c.nb_cnt); return rc; }
  c.inactive = (uint8_t *)calloc((size_t)n * p, 1);
  c.mr = (double *)malloc(sizeof(double) * (size_t)n * p * K);
  c.ml = (double *)malloc(sizeof(double) * (size_t)n * p * K);
  c.message = (double *)malloc(sizeof(double) * K);
  c.f_j = (double *)malloc(sizeof(double) * K);
  c.weights = (double *)malloc(sizeof(double) * K);
  const double p0 = 1.0 / ((double)K); /* init_messages, family_bp.cpp:421-427 */
  for (int64_t t = 0; t < (int64_t)n * p * K; t++) { c.mr[t] = p0; c.ml[t] = p0; }

  int32_t *Z = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * p);
  int32_t *Zk = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * p);
  for (int64_t t = 0; t < (int64_t)n * p; t++) Z[t] = -1;
  run_bp(&c);                 /* FamilyBP::run, family_bp.cpp:79-100 */
  sample_posterior(&c, Z);
  rc = family_knockoffs_run(&c, F, Z, Zk, grp_first, grp_last, G, max_group);
  if (rc == 0) {
    /* plain emission per member with its global references (knockoffs.cpp:505-510, :1430-1441) */
    double wb[2];
    for (int i = 0; i < n; i++) {
      const int i_abs = F->members[i];
      const int32_t *ref = P->ref_global + (int64_t)i_abs * K;
      uint8_t *row = hk_out ? hk_out + (int64_t)i * P->row_bytes : NULL;
      if (row) memset(row, 0, (size_t)P->row_bytes);
      for (int j = 0; j < p; j++) {
        const int h_predicted = hap_bit(P, ref[Zk[(int64_t)i * p + j]], j);
        wb[1] = mutate(h_predicted, P->mu[j]);
        wb[0] = 1.0 - wb[1];
        const int hk = kgw_oracle_weighted_choice(wb, 2, kgw_oracle_next_uniform(rng, i_abs, KGW_STREAM_HK, j, p, 0));
        if (row && hk) row[j >> 3] |= (uint8_t)(1u << (j & 7));
      }
    }
    if (z_out) memcpy(z_out, Z, sizeof(int32_t) * (size_t)n * p);
    if (zk_out) memcpy(zk_out, Zk, sizeof(int32_t) * (size_t)n * p);
  }
  if (bp_sweeps) *bp_sweeps = c.bp_sweeps;
  free(grp_first); free(grp_last); free(c.nb_cnt); free(c.nb); free(c.inactive); free(c.mr); free(c.ml);
  free(c.message); free(c.f_j); free(c.weights); free(Z); free(Zk);
  return rc;
}
