/* oracle/kgw_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of snpknock2's haplotype knockoff sampler
 * (msesia/knockoffgwas, snpknock2/src/knockoffs.cpp:311-700, :1041-1075,
 * :1215-1502; utils.cpp:690-698), in plain C with strict IEEE-754 double
 * arithmetic (compile with -O2 -ffp-contract=off, no fast-math), every sum over
 * k sequential and every product in the reference's left-to-right evaluation
 * order.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Parity pin: tests/test_oracle_vs_reference.py replays the reference's own
 * taus88 stream through this code and requires the haplotype-level knockoffs
 * dumped from the unmodified reference binary (oracle/_ref/snpknock2_probe) and
 * its traced latent paths to be reproduced bit-for-bit (tests/golden/).
 */
#ifndef KGW_ORACLE_H
#define KGW_ORACLE_H
#include <stdint.h>
#include "kgw_rng.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t N, p, K, W;
  int64_t row_bytes;          /* bytes per packed haplotype row, >= ceil(p/8)            */
  const uint8_t *H;           /* [N][row_bytes], bit j of row i at byte j>>3, bit j&7    */
  const int32_t *ref_local;   /* [N][W][K]  references_local  (knockoffs.h:139)          */
  const int32_t *ref_global;  /* [N][K]     references_global (knockoffs.h:141)          */
  const double *b;            /* [p] b_j = exp(-rho d_j) (knockoffs.cpp:1508-1517)       */
  const double *mu;           /* [p] mutation rates                                      */
  const int32_t *groups;      /* [p] group of each SNP, contiguous, non-decreasing       */
  const int32_t *win_start;   /* [W] metadata.windows.start                              */
  const int32_t *win_end;     /* [W] metadata.windows.end (exclusive)                    */
} kgw_oracle_problem;

enum { KGW_ORACLE_RNG_INJECTED = 0, KGW_ORACLE_RNG_TAUS88 = 1, KGW_ORACLE_RNG_PHILOX = 2 };
enum { KGW_STREAM_Z = 0, KGW_STREAM_ZK = 1, KGW_STREAM_HK = 2, KGW_STREAM_FAMILY = 3 };

typedef struct {
  int32_t mode;
  const double *injected;   /* mode 0: U[hap][3][p] (hap = absolute haplotype index)      */
  kgw_taus88 taus;          /* mode 1: sequential stream, consumed in reference call order */
  uint64_t seed;            /* mode 2: Philox key                                          */
  int64_t n_draws;          /* bookkeeping                                                 */
  double *record;           /* optional U[N][3][p]: every uniform handed out is also written at
                               its (haplotype, stream, locus) key (lets a sequential taus88 replay
                               be re-fed to the CUDA path as injected uniforms)                     */
  int32_t n_windows;        /* > 1: `injected` / `record` are U[N][W][3][p], one block per padded window pass
                               (the passes of knockoffs.cpp:1222-1226 overlap); 0 or 1: U[N][3][p]           */
  int32_t reserved;
} kgw_oracle_rng;

/* Unrelated haplotypes (generate_unrelated_worker, knockoffs.cpp:636-700).
 * haps[n] are processed in the given order (the reference: ascending, :583-591).
 * Outputs are indexed by position in `haps`; any may be NULL.
 *   z, zk : int32 [n][p]      latent path and its knockoff
 *   hk    : uint8 [n][row_bytes] packed knockoff alleles (chaplotype layout)
 *   beta  : double [n][p][K]  normalised backward messages of the LAST window pass
 * Returns 0, or a negative error code. */
int kgw_oracle_unrelated(const kgw_oracle_problem *P, kgw_oracle_rng *rng,
                         const int32_t *haps, int32_t n,
                         int32_t *z, int32_t *zk, uint8_t *hk, double *beta);

/* --estimate-hmm as the reference runs it (Knockoffs::EM, knockoffs.cpp:706-960): mutation rates only, rho fixed.
 * mu_init, mu_out: double [p]; deltas (may be NULL): double [20], the change reported by every step; *iterations =
 * steps run.  Returns 0 when converged, 1 when the 20 steps ran out, negative on error. */
int kgw_oracle_em(const kgw_oracle_problem *P, const double *mu_init, double *mu_out, int32_t *iterations, double *deltas);

/* One IBD-sharing family as generate_related_cluster() sees it (knockoffs.cpp:465-481):
 * member list + TIDIED segments (the unchanged host code runs IbdCluster::tidy, ibd.cpp:99-189). */
typedef struct {
  int32_t n_members;
  const int32_t *members;    /* absolute haplotype ids, order of metadata.related_families[f]    */
  int32_t n_segs;
  const int32_t *seg_jmin;   /* [n_segs] first SNP of the segment                                 */
  const int32_t *seg_jmax;   /* [n_segs] last SNP (inclusive)                                     */
  const int32_t *seg_off;    /* [n_segs+1] CSR offsets into seg_ids                               */
  const int32_t *seg_ids;    /* absolute haplotype ids of each segment, ascending (std::set<int>) */
} kgw_oracle_family;

/* Related haplotypes of one family (generate_related_cluster, knockoffs.cpp:465-571: FamilyBP::run,
 * FamilyKnockoffs::run, sample_emission with references_global).  Outputs indexed by position in
 * F->members; any may be NULL.  z, zk: int32 [n_members][p]; hk: uint8 [n_members][row_bytes].
 * Draws are keyed (absolute haplotype, stream, locus) exactly like the unrelated branch; in
 * taus88 mode they are consumed in the reference's call order.
 * Returns 0, -10 / -11 where the reference aborts ("Missing knockoff variants" /
 * "Inconsistency found on Zk", family_knockoffs.cpp:272-291), other negatives on bad input. */
/* audit of the CUDA kernel's BP work-skipping rule (see kgw_oracle_family.c); stats = messages, skippable, violations */
void kgw_oracle_family_audit(int enable);
void kgw_oracle_family_audit_get(int64_t out[3]);

int kgw_oracle_family_run(const kgw_oracle_problem *P, kgw_oracle_rng *rng, const kgw_oracle_family *F,
                          int32_t *z, int32_t *zk, uint8_t *hk, int64_t *bp_sweeps);

/* next uniform of the selected source (shared by both branches) */
double kgw_oracle_next_uniform(kgw_oracle_rng *rng, int hap, int stream, int locus, int p, int w);

/* weighted_choice restated (utils.cpp:690-698), exposed for unit tests. */
int kgw_oracle_weighted_choice(const double *w, int n, double u);

#ifdef __cplusplus
}
#endif
#endif
