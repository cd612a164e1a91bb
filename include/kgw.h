/* include/kgw.h -- C ABI of the B200-native haplotype knockoff sampler.
 *
 * Drop-in boundary: the reference (msesia/knockoffgwas, snpknock2) has no plugin /
 * FFI interface; the seam is the C++ method
 *
 *     void Knockoffs::generate(const int num_threads)     snpknock2/src/knockoffs.h:56
 *                                                          snpknock2/src/knockoffs.cpp:311-327
 *
 * called once per (chromosome, resolution) from KnockoffGenerator::generate()
 * (generator_master.cpp:109-115, main.cpp:159-181).  Everything it reads is host
 * state that the unchanged snpknock2 code has already prepared; everything it
 * writes is the packed knockoff matrix Hk.  This header is that seam as plain C:
 * flat caller-owned buffers in, flat caller-owned buffer out, int status.
 * INTEGRATION.md shows the 40-line body a maintainer puts in Knockoffs::generate().
 *
 * All pointers in kgw_problem are HOST pointers unless a function says otherwise.
 * The callee never retains a caller pointer past return (kgw_plan_* copies what it
 * needs to the device).  No torch / C++ types cross this boundary.
 */
#ifndef KGW_H
#define KGW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGW_ABI_VERSION 4

/* status codes (negative = error; kgw_last_error() has the text) */
enum {
  KGW_OK = 0,
  KGW_ERR_INVALID = -1,      /* bad argument / inconsistent problem description          */
  KGW_ERR_UNSUPPORTED = -2,  /* valid for the reference but outside this build (e.g. K)   */
  KGW_ERR_CUDA = -3,         /* CUDA runtime failure or no sm_100 device                  */
  KGW_ERR_NOMEM = -4
};

/* random streams of the counter-based generator (replaces the per-thread sequential
 * boost::random::taus88 of knockoffs.cpp:401-406,596-612 -- results no longer depend
 * on --n_threads).  u(seed, hap, stream, locus) =
 *     Philox4x32-10(key = {seed lo, seed hi}, ctr = {locus >> 2, hap, stream, 0})[locus & 3] / 2^32
 * i.e. 32-bit resolution like the reference's uniform_real_distribution over taus88
 * (uniform_real_distribution.hpp:56-64).  `hap` is the absolute haplotype index.  With
 * --windows > 0 the padded window passes (knockoffs.cpp:1222-1226) overlap; pass w uses
 * stream word `stream | (w << 2)` so no uniform is consumed twice. */
enum { KGW_STREAM_Z = 0, KGW_STREAM_ZK = 1, KGW_STREAM_HK = 2, KGW_STREAM_FAMILY = 3 };

/* IBD-sharing families for the related branch (Knockoffs::generate_related, knockoffs.cpp:329-571).
 * The unchanged host code builds, per family, the member list (metadata.related_families,
 * metadata.cpp:242-293) and its TIDIED segments (IbdCluster(ibd_segments), knockoffs.cpp:474-481,
 * ibd.cpp:99-189) and passes them flat; segment order = iteration order of the reference's
 * std::set<IbdSeg> (IbdCluster::get(s), s = 0..size-1). */
typedef struct kgw_families {
  int32_t n_families;
  int32_t reserved;
  const int32_t *fam_off;      /* [n_families+1] CSR into fam_members                                */
  const int32_t *fam_members;  /* absolute haplotype ids, order of metadata.related_families[f]      */
  const int32_t *seg_off;      /* [n_families+1] CSR into seg_jmin / seg_jmax / seg_ids_off          */
  const int32_t *seg_jmin;     /* [n_segments] IbdSeg::j_min                                         */
  const int32_t *seg_jmax;     /* [n_segments] IbdSeg::j_max (inclusive)                             */
  const int32_t *seg_ids_off;  /* [n_segments+1] CSR into seg_ids                                    */
  const int32_t *seg_ids;      /* IbdSeg::indices (absolute haplotype ids, ascending)                */
} kgw_families;

/* What Knockoffs::generate() reads. */
typedef struct kgw_problem {
  int32_t abi_version;        /* = KGW_ABI_VERSION                                                   */
  int32_t N;                  /* Haplotypes::num_haps                       haplotypes.h:38          */
  int32_t p;                  /* Haplotypes::num_snps                       haplotypes.h:37          */
  int32_t K;                  /* Knockoffs::K                               knockoffs.h:138
                                 Bound of this build: K <= 256 (latent states are stored as bytes), K <= 255
                                 when `families` is given; larger K returns KGW_ERR_UNSUPPORTED.  The
                                 reference's default is 100 (arguments.cpp), module_2_knockoffs.sh uses 50.  */
  int32_t W;                  /* metadata.windows.num_windows               windows.h:17             */
  int32_t reserved0;
  int64_t row_bytes;          /* bytes per packed row of H / Hk, >= ceil(p/8)                        */
  const uint8_t *H;           /* [N][row_bytes] Haplotypes::H, chaplotype bit layout
                                 (bit j at byte j>>3, mask 1<<(j&7))        chaplotype.h:25          */
  const int32_t *ref_local;   /* [N][W][K] Knockoffs::references_local (sorted) knockoffs.h:139      */
  const int32_t *ref_global;  /* [N][K]    Knockoffs::references_global     knockoffs.h:141; may be
                                 NULL when there are no related haplotypes                          */
  const double *b;            /* [p] Knockoffs::b            knockoffs.h:76, knockoffs.cpp:1508-1517 */
  const double *mu;           /* [p] Knockoffs::mut_rates    knockoffs.h:77                          */
  const int32_t *groups;      /* [p] Knockoffs::groups (contiguous runs, 0-based) knockoffs.h:151    */
  const int32_t *win_start;   /* [W] metadata.windows.start                 windows.h:18             */
  const int32_t *win_end;     /* [W] metadata.windows.end (exclusive)       windows.h:18             */
  const int32_t *unrelated;   /* [n_unrelated] metadata.unrelated, ascending metadata.h:101          */
  int32_t n_unrelated;
  int32_t reserved1;
  uint64_t seed;              /* Knockoffs::seed                            knockoffs.h:71           */
  /* test hook: if non-NULL, U[N][W][3][p] doubles in [0,1) replace the generator (window pass w, stream order
   * KGW_STREAM_Z, _ZK, _HK; indexed by absolute haplotype and absolute locus -- the padded passes of
   * knockoffs.cpp:1222-1226 overlap, so every pass has its own block; the related branch reads block 0). */
  const double *injected_u;
  /* related haplotypes (NULL or n_families == 0: none).  Needs ref_global.  Their draws use the same
   * keyed generator: (member haplotype, KGW_STREAM_Z / _ZK / _HK, locus). */
  const kgw_families *families;
} kgw_problem;

typedef struct kgw_options {
  int32_t device;        /* CUDA device ordinal (one process per GPU: pass LOCAL_RANK)        */
  int32_t block_loci;    /* 0 = default (4); 2, 4 or 8: spacing of the stored backward-message rows */
  int32_t lanes_per_hap; /* 0 = auto; 1, 2, 4 or 8 lanes cooperate on one haplotype (K <= 32 lanes) */
  int32_t ctas_per_sm;   /* 0 = auto (occupancy query)                                        */
  void *stream;          /* cudaStream_t to launch on (caller-owned); NULL = the plan creates
                            its own non-blocking stream                                       */
  int32_t reserved;
  int32_t segment_loci;  /* 0 = auto (whole chromosome when the scratch fits); otherwise a multiple of
                            32: loci per segment of the backward/forward sweeps                */
  /* Several GPUs in one process -- the counterpart of the reference's worker threads inside generate(num_threads)
   * (knockoffs.cpp:573-634 unrelated slices, :329-431 families).  n_devices > 1: the sorted unrelated list is cut
   * into contiguous, cost-balanced slices and the IBD families are dealt out whole (largest first, least loaded
   * device); every device holds H and the references, runs on its own stream from its own host thread and copies
   * its rows of Hk straight into the caller's hk_out (disjoint rows, no exchange between devices).  The result
   * does not depend on n_devices (keyed generator).  `devices` lists the CUDA ordinals (NULL = 0 .. n_devices-1);
   * `device` and `stream` are ignored / must be NULL then.  n_devices 0 or 1: one GPU, `device`. */
  int32_t n_devices;
  int32_t reserved2;
  const int32_t *devices;
} kgw_options;

/* Optional debug outputs (any pointer may be NULL); rows follow `unrelated` order. */
typedef struct kgw_debug {
  int32_t *z;            /* [n_unrelated][p] sampled latent path (knockoffs.cpp:1243,1259)     */
  int32_t *zk;           /* [n_unrelated][p] knockoff path       (knockoffs.cpp:1421)          */
  double *beta_ckpt;     /* [n_unrelated][n_ckpt][K] backward messages at the checkpoint loci of
                            the LAST window pass, normalised to sum 1 (knockoffs.cpp:1069-1073) */
  int32_t *ckpt_locus;   /* [n_ckpt] the loci those rows belong to                             */
  int32_t n_ckpt;        /* in: capacity of the two arrays above; out: rows written            */
  int32_t reserved;
  int32_t *fam_z;        /* [total family members][p] latent paths of the related haplotypes, rows in
                            fam_members order (FamilyBP::run, family_bp.cpp:79)                  */
  int32_t *fam_zk;       /* [total family members][p] their knockoff paths (family_knockoffs.cpp:77) */
} kgw_debug;

typedef struct kgw_plan kgw_plan;

int32_t kgw_abi_version(void);
/* CUDA devices visible to the process (0 when there is none: every compute entry point then fails, there is no
 * CPU fallback) */
int32_t kgw_device_count(void);
/* The split kgw_options.n_devices applies, for callers that shard across processes themselves (one process per
 * GPU): unrelated_cut[n_devices+1] are positions in the sorted unrelated list (device d takes
 * [cut[d], cut[d+1])), family_device[f] the device of family f.  A family member counts as 7 unrelated haplotypes
 * (measured cost of the related branch); families go largest first to the least loaded device, as the reference
 * deals them to its worker threads in descending size (knockoffs.cpp:364-372).  Host only, no device needed. */
int32_t kgw_partition_work(int32_t n_unrelated, int32_t n_families, const int32_t *fam_off, int32_t n_devices,
                           int32_t *unrelated_cut, int32_t *family_device);
const char *kgw_last_error(void);

/* One-shot drop-in for Knockoffs::generate(): host buffers in, host buffer out.
 * hk_out is [N][row_bytes]; the rows listed in `unrelated` (generate_unrelated, knockoffs.cpp:573-700)
 * and the members of `families` (generate_related, :329-571) are written.  H2D, kernels and D2H all
 * inside.  Returns KGW_ERR_INVALID with the reference's own message where the reference aborts
 * ("Missing knockoff variants" / "Inconsistency found on Zk", family_knockoffs.cpp:272-291). */
int32_t kgw_generate(const kgw_problem *prob, const kgw_options *opt, uint8_t *hk_out, kgw_debug *dbg);

/* Audit mode of the unrelated branch (hard part 1): the same sampler with every sum over the K
 * states taken left to right and every product in the reference's evaluation order, correctly rounded operation by
 * operation (backward_const / backwardSum knockoffs.cpp:1052-1073, N_sum / N_norm :1345-1390, weights_sum
 * utils.cpp:690-698) -- the arithmetic a strict-IEEE build of the reference performs, where kgw_generate sums over
 * lanes as a tree with fused multiply-adds.  One thread per haplotype with its p x K tables in HBM: slow, meant for
 * checking the fast path on a sample of haplotypes (prob->unrelated lists them; families must be absent).  Same
 * keyed generator, same injected_u hook, same outputs as kgw_generate: the listed rows of hk_out, dbg->z / zk, and
 * in dbg->beta_ckpt EVERY backward row (n_ckpt >= p on entry, = p on return; the table the reference's worker holds
 * after the last window pass).  Latent states are int32 here as in the reference: no bound on K. */
int32_t kgw_generate_strict(const kgw_problem *prob, int32_t device, uint8_t *hk_out, kgw_debug *dbg);

/* Resident-data path (bench / repeated resolutions): upload once, run many times. */
int32_t kgw_plan_create(const kgw_problem *prob, const kgw_options *opt, kgw_plan **plan_out);
/* Replace the partition (Knockoffs::set_partition, knockoffs.cpp:91-107) without re-uploading H. */
int32_t kgw_plan_set_groups(kgw_plan *plan, const int32_t *groups);
/* Launch on the plan's stream; returns after the kernels have been enqueued. */
int32_t kgw_plan_run(kgw_plan *plan);
/* Run `iters` times, timed with CUDA events on the launching stream; ms_out = mean ms per run. */
int32_t kgw_plan_run_timed(kgw_plan *plan, int32_t iters, float *ms_out);
int32_t kgw_plan_sync(kgw_plan *plan);
/* D2H of the packed knockoff rows (same layout as kgw_generate's hk_out). */
int32_t kgw_plan_download(kgw_plan *plan, uint8_t *hk_out, kgw_debug *dbg);
/* Device pointer of the packed knockoff matrix, rows kgw_plan_info().row_stride_bytes apart (for NCCL gathers
 * etc.).  Multi-device plan: the rows of the other devices are first brought to the first device's matrix (packed
 * on the source device, one peer copy per device over NVLink, scattered on arrival) and that pointer is returned. */
void *kgw_plan_device_hk(kgw_plan *plan);
/* Introspection for bench.py. */
typedef struct kgw_plan_info_t {
  int32_t launches_per_run;   /* kernel launches per kgw_plan_run()                            */
  int32_t grid, block;        /* persistent grid / CTA size of the sampler kernel              */
  int32_t lanes_per_hap;      /* lanes cooperating on one haplotype                            */
  int32_t states_per_lane;    /* HMM states held in registers per lane                         */
  int32_t block_loci;         /* spacing of the stored backward-message rows                   */
  int32_t segment_loci;       /* loci per segment of sweeps 1-2 (>= p: one segment)            */
  int32_t reserved;
  int64_t scratch_bytes;      /* checkpoint + latent-path scratch in HBM                       */
  int64_t row_stride_bytes;   /* bytes between rows of the DEVICE Hk matrix (kgw_plan_device_hk) */
  int64_t smem_bytes;         /* dynamic shared memory per CTA                                 */
} kgw_plan_info_t;
int32_t kgw_plan_info(kgw_plan *plan, kgw_plan_info_t *out);
void kgw_plan_destroy(kgw_plan *plan);

/* ---- PLINK .bed body ---------------------------------------------------
 * Replaces the data loops of writeBED / writeBED_variant (plink_interface.cpp:10-35, 67-103), called from
 * Knockoffs::write (knockoffs.cpp:181-186) through plink::write_binary (:255-284) right after generate():
 * SNP-major 2-bit genotype columns of kgw_bed_column_bytes(N) = ceil(N/2/4) bytes each, sample s =
 * haplotypes 2s, 2s+1.  The caller writes the magic 6c 1b 01 and then this buffer.
 *   include_genotypes != 0:  2p columns, SNP j -> (H, Hk), or (Hk, H) where swap[j] != 0; `swap` = the
 *       coin flips of plink::write_binary (mt19937_64(2019), uniform_int_distribution<int>(0,1), :267-274),
 *       drawn by the unchanged host code, which also writes the matching .bim / .fam;
 *   include_genotypes == 0:  p columns of Hk (the "<out>_tilde" file), swap ignored.
 * kgw_plan_encode_bed uses the matrices resident in the plan (H as uploaded, Hk as last generated);
 * kernel_ms (may be NULL) receives the device time of the transpose.  kgw_bed_encode is the one-shot form
 * from host matrices [N][row_bytes]; pass H == NULL for the Hk-only file. */
int64_t kgw_bed_column_bytes(int32_t N);
int32_t kgw_plan_encode_bed(kgw_plan *plan, const uint8_t *swap, int32_t include_genotypes, uint8_t *bed_out, float *kernel_ms);
int32_t kgw_bed_encode(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, const uint8_t *Hk, const uint8_t *swap,
                       int32_t device, uint8_t *bed_out);

/* ---- reference selection ------------------------------------------------
 * Replaces the worker loop of KinshipComputer::assign_references (kinship_computer.cpp:856-881 ->
 * findNeighbors :174-233) for one window (--windows 0, and the global references): for every haplotype the K
 * members of its kinship cluster with the smallest Hamming distance on the thinned chromosome (every
 * `compression`-th SNP, metadata.cpp:563-580; main.cpp:101 uses 10), IBD relatives pushed to distance 1e6,
 * ties in cluster order.  ref_out[i][0..K) is in (distance, cluster position) order, -1 where the cluster
 * has fewer than K+1 members; the unchanged host code then runs combine_references_families (:765-835) and
 * its sanity checks (:887-928) on it.
 *   clust_off [n_clusters+1], clust_members: the kinship clusters in the order of clusters[chr][c] (:693-711)
 *   family_of [N] or NULL: a label shared by the haplotypes of metadata.related_families[i], negative for
 *   haplotypes without relatives.
 * (With --windows > 0 the reference applies full-resolution window bounds to the thinned rows and reads past
 * their end, distance_hap_window :78-104; that mode has no defined result and is not offered.) */
int32_t kgw_assign_references(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, int32_t K, int32_t compression,
                              int32_t n_clusters, const int32_t *clust_off, const int32_t *clust_members,
                              const int32_t *family_of, int32_t device, int32_t *ref_out, float *kernel_ms);

/* ---- kinship clusters: the data passes of the bifurcating k-means ----------
 * KinshipComputer::bifurcating_kmeans (kinship_computer.cpp:511-691) splits the haplotypes into clusters of at most
 * cluster_size_max by repeated 2-means on the thinned genotypes; the clusters feed assign_references above.  Its two
 * data passes run on the device, on a matrix that is uploaded once per run:
 *   kgw_kinship_create      H [N][row_bytes]: the thinned rows of ALL chromosomes side by side (packed, LSB first,
 *                           p loci in chromosome order), N even (haplotypes 2s, 2s+1 = sample s)
 *   kgw_kinship_distances   update_assignments' distances (:314-336 with distance :239-250): for member t (a haplotype
 *                           index i) d0[t] = sum_j (H[i0][j]-mu0[j])^2 + sum_j (H[i1][j]-mu0[j])^2 over the loci with
 *                           use[j] != 0 (NULL = all; zeros on the chromosome left out), i0 = 2*(i/2), i1 = i0+1; d1 with
 *                           mu1.  Sequential fp64 sums in locus order, no fused multiply-add: the numbers a strict-IEEE
 *                           build of the reference compares in `dist_0 < dist_1` (:331).
 *   kgw_kinship_centroids   update_centroids (:368-405): column means of the members with assign 0 / 1 (exact counts,
 *                           one division per locus as the reference divides); n0_out / n1_out (may be NULL): the sizes.
 * The control flow around them -- initial centroids from the reference's random stream (:252-312), the size
 * rebalancing (:338-362), acceptance of a split and the list of unfinished clusters (:620-676) -- is integer bookkeeping
 * that stays on the host (integration/snpknock2_kgw_shim.cpp keeps the reference's order of operations). */
typedef struct kgw_kinship kgw_kinship;
int32_t kgw_kinship_create(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, int32_t device, kgw_kinship **out);
void kgw_kinship_destroy(kgw_kinship *k);
int32_t kgw_kinship_distances(kgw_kinship *k, const int32_t *members, int32_t n, const double *mu0, const double *mu1,
                              const uint8_t *use, double *d0, double *d1);
int32_t kgw_kinship_centroids(kgw_kinship *k, const int32_t *members, const int32_t *assign, int32_t n, double *mu0, double *mu1,
                              int32_t *n0_out, int32_t *n1_out);

/* ---- phased BGEN -> packed haplotype matrix -----------------------------
 * Replaces the data loop of BgenReader::read_worker (bgen_reader.cpp:139-261) behind BgenReader::read
 * (:66-80), the producer of Haplotypes::H (dataset.cpp).  `file` is the whole .bgen image (v1.2, layout 2,
 * zstd / zlib / no compression); the unchanged host code keeps the sample and variant bookkeeping:
 *   sample_idx     [n_samples] row of the BGEN sample block for output sample i (match_indices, :167-168)
 *   variant_of_col [p]         position in file order of the variant stored in output column j
 *                              (the two std::find on rsid, :183,191-201)
 * H_out [2 n_samples][row_bytes]: chaplotype rows (bit j at byte j>>3, mask 1<<(j&7)), haplotypes 2i, 2i+1 of
 * sample i as H[2*i].set / H[2*i+1].set fill them (:207-210).  A probability that is not exactly 0 or 1, a
 * missing sample or a block that is not phased / diploid / biallelic returns KGW_ERR_INVALID -- the
 * reference throws "Incorrect input" from prob_to_hap (:48-63).  zlib blocks are inflated on the device (one warp
 * per block; a stream that does not inflate to its stated length or fails its Adler-32 returns KGW_ERR_INVALID,
 * the reference's uncompress() != Z_OK), zstd blocks on `host_threads` host threads (0 = all; system libzstd.so.1,
 * loaded on first use; KGW_BGEN_HOST_INFLATE=1 sends zlib blocks the same way through libz.so.1); all are decoded
 * and transposed on the device; kernel_ms (may be NULL) receives the device time of the kernels. */
int32_t kgw_bgen_variant_count(const uint8_t *file, int64_t file_bytes, int32_t *n_variants, int32_t *n_samples);
/* rsid of every variant in file order as (offset into `file`, length): what BgenParser::read_variant hands to the
 * std::find calls of read_worker (bgen_reader.cpp:176-183); the caller builds its id -> position map from it */
int32_t kgw_bgen_rsids(const uint8_t *file, int64_t file_bytes, int32_t n_variants, int64_t *rsid_off, int32_t *rsid_len);
int32_t kgw_bgen_read(const uint8_t *file, int64_t file_bytes, int32_t n_samples, const int32_t *sample_idx, int32_t p,
                      const int32_t *variant_of_col, int32_t host_threads, int32_t device, int64_t row_bytes, uint8_t *H_out,
                      float *kernel_ms);

/* ---- `.haps` text -> packed haplotype matrix (text input) -------------------
 * Replaces the data loop of HapsReader::read (haps_reader.cpp:93-124).  `text` is the whole .haps file: one row per
 * SNP (rows as std::getline splits them), tokens separated by blanks / tabs as sutils::tokenize (utils.cpp:234-246)
 * splits them, a trailing '\r' ignored on the last token of a row; haplotype a of file sample s is token 2s + a
 * (:79-81) and its allele is 1 exactly when that token is "1" (:109).
 *   sample_idx [n_samples]  row of the .sample file for output sample i (the std::find of :73-85)
 *   row_of_col [p]          file row of the SNP stored in output column j (the std::find of :98-104 on legend ids)
 * H_out [2 n_samples][row_bytes] chaplotype rows.  A requested row with fewer tokens than a requested column
 * returns KGW_ERR_INVALID (the reference indexes past the end of `tokens`). */
int32_t kgw_haps_read(const uint8_t *text, int64_t text_bytes, int32_t n_samples, const int32_t *sample_idx, int32_t p,
                      const int32_t *row_of_col, int32_t device, int64_t row_bytes, uint8_t *H_out, float *kernel_ms);

/* ---- `--estimate-hmm` ---------------------------------------------------------
 * Replaces Knockoffs::EM (knockoffs.cpp:706-748 -> EM_step :830-960 -> EM_worker :750-828), called from main.cpp
 * when --estimate-hmm is given, after init_hmm(K, rho) (:1527-1548).  As the reference stands only the mutation
 * rates are re-estimated (new_rho = rho, :929-930): per step and haplotype the backward and forward messages over
 * the local references of every window pass, the posterior marginals, the expected mismatches gamma_j; then
 * mu_j = clamp(mean gamma_j, 1e-6, 1e-3); at most 20 steps, stopping when the root mean square relative change
 * falls below 5e-2.  Reads prob->H, ref_local, b, the windows (groups / mu / unrelated / families are ignored: every
 * haplotype takes part, :862-864).
 *   mu_init [p]   the rates to start from (init_hmm leaves 1e-5 everywhere, see oracle/kgw_oracle.c)
 *   mu_out  [p]   the rates after the last step (what update_hmm, :1550-1561, installs)
 *   iterations    steps run;  deltas [20] (may be NULL): the change printed per step (:724-728)
 * Returns KGW_OK when converged, 1 when 20 steps did not converge (the reference warns and goes on, :739-748),
 * negative on error.  Floating-point sums over haplotypes are order dependent at the last bits: results agree
 * with the reference to ~1e-12 relative, not bit for bit. */
int32_t kgw_estimate_hmm(const kgw_problem *prob, const kgw_options *opt, const double *mu_init, double *mu_out,
                         int32_t *iterations, double *deltas);

/* A destroyed plan parks its scratch slab (tens of GB) for the next plan on the same device, because
 * Knockoffs::generate() runs once per resolution (main.cpp:159-181); this frees the parked slabs. */
void kgw_release_workspace(void);

/* Host replay of the generator (bit-identical to the device code). */
double kgw_uniform(uint64_t seed, uint32_t hap, uint32_t stream, uint32_t locus);

#ifdef __cplusplus
}
#endif
#endif /* KGW_H */
