// integration/snpknock2_kgw_shim.cpp -- the reference-side binding of include/kgw.h, compiled.
//
// INTEGRATION.md section 2 shows the body a maintainer pastes into Knockoffs::generate
// (snpknock2/src/knockoffs.cpp:311).  This file is the same code as a link-time replacement, so that the
// drop-in can be built and run without touching a reference source file: GNU ld --wrap puts
// wrap_generate() below in front of Knockoffs::generate in the UNMODIFIED reference objects
// (integration/build_dropin.sh -> oracle/_ref/snpknock2_kgw).  Everything else -- the CLI, the BGEN / HAPS
// readers, metadata, kinship clusters and reference selection, IBD tidying, partitions, the PLINK / text
// writers -- is the reference's own code, unchanged; only the sampler runs in libkgw.so on the B200.
//
// KGW_SHIM_DUMP=<dir>: additionally writes the kgw_problem the shim assembled as raw little-endian arrays
// (tests/test_dropin_shim.py compares them with the inputs the probe recorded from the reference).
// every standard header the reference pulls in comes first: `private` is redefined for the reference's own headers only
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <numeric>
#include <cstring>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>

#define private public
#define protected public
#include "knockoffs.h"
#undef private
#undef protected

#include "kgw.h"

#include "bgen_reader.h"
#include "plink_interface.h"
#define private public
#include "haps_reader.h"
#include "kinship_computer.h"
#undef private
#include <random>
#include <unordered_map>

#if !defined(KGW_GEN_SYM) || !defined(KGW_BGEN_SYM) || !defined(KGW_WB_SYM) || !defined(KGW_HAPS_SYM) || !defined(KGW_AR_SYM)
#error "KGW_GEN_SYM / KGW_BGEN_SYM / KGW_WB_SYM / KGW_HAPS_SYM / KGW_AR_SYM (mangled Knockoffs::generate, BgenReader::read, plink::write_binary, HapsReader::read, KinshipComputer::assign_references) must be defined by the build script"
#endif
#define KGW_STR2(x) #x
#define KGW_STR(x) KGW_STR2(x)

void wrap_generate(Knockoffs *self, int num_threads) asm("__wrap_" KGW_STR(KGW_GEN_SYM));
// the two neighbours of the path; KGW_SHIM_IO=0 leaves them to the reference
void real_bgen_read(BgenReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose, int nthreads) asm("__real_" KGW_STR(KGW_BGEN_SYM));
void wrap_bgen_read(BgenReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose, int nthreads) asm("__wrap_" KGW_STR(KGW_BGEN_SYM));
void real_write_binary(const string &basename, const Metadata &metadata, const vector<chaplotype> &H, const vector<chaplotype> &Hk,
                       bool random) asm("__real_" KGW_STR(KGW_WB_SYM));
void wrap_write_binary(const string &basename, const Metadata &metadata, const vector<chaplotype> &H, const vector<chaplotype> &Hk,
                       bool random) asm("__wrap_" KGW_STR(KGW_WB_SYM));
void real_haps_read(HapsReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose) asm("__real_" KGW_STR(KGW_HAPS_SYM));
void wrap_haps_read(HapsReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose) asm("__wrap_" KGW_STR(KGW_HAPS_SYM));
ivector3d real_assign_references(const KinshipComputer *self, const int K, const int chr, const Windows &windows) asm("__real_" KGW_STR(KGW_AR_SYM));
ivector3d wrap_assign_references(const KinshipComputer *self, const int K, const int chr, const Windows &windows) asm("__wrap_" KGW_STR(KGW_AR_SYM));
#ifdef KGW_KC_SYM
// KinshipComputer::KinshipComputer(metadata, compression, cluster_size_min, cluster_size_max, num_threads)
// (kinship_computer.cpp:12-36), the caller of the bifurcating k-means; its call of cluster() stays inside
// kinship_computer.o, so the seam that can be wrapped is the constructor itself
void real_kc_ctor(KinshipComputer *self, const vector<Metadata> &metadata, int compression, int cluster_size_min, int cluster_size_max,
                  int num_threads) asm("__real_" KGW_STR(KGW_KC_SYM));
void wrap_kc_ctor(KinshipComputer *self, const vector<Metadata> &metadata, int compression, int cluster_size_min, int cluster_size_max,
                  int num_threads) asm("__wrap_" KGW_STR(KGW_KC_SYM));
#endif
void writeBIM(const string &basename, const Metadata &metadata, bool include_genotypes, const vector<bool> &swap);   // plink_interface.cpp:105
void writeFAM(const string &basename, const Metadata &metadata);                                                      // plink_interface.cpp:162

namespace {
bool shim_io() {
  const char *e = getenv("KGW_SHIM_IO");
  return !(e && e[0] == '0');
}
}  // namespace

// BgenReader::read (bgen_reader.cpp:66-80 -> read_mt -> read_worker :139-261): same bookkeeping, data loop on the device
void wrap_bgen_read(BgenReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose, int nthreads) {
  if (!shim_io()) { real_bgen_read(self, sample_ids, snp_ids, H, verbose, nthreads); return; }
  std::ifstream fd(self->filename, std::ios::binary | std::ios::ate);
  if (!fd.is_open()) { std::cout << "Exception occurred while trying to open BGEN file " << self->filename << std::endl; exit(1); }
  std::vector<uint8_t> image((size_t)fd.tellg());
  fd.seekg(0);
  fd.read((char *)image.data(), (std::streamsize)image.size());
  const int num_samples = (int)sample_ids.size(), num_snps = (int)snp_ids.size();
  vector<int> sample_idx(num_samples, -1);
  match_indices(self->sample.ID, sample_ids, sample_idx);            // unchanged (:167-168)
  int32_t M = 0, Nfile = 0;
  if (kgw_bgen_variant_count(image.data(), (int64_t)image.size(), &M, &Nfile) != KGW_OK) { std::cerr << kgw_last_error() << std::endl; exit(1); }
  std::vector<int64_t> roff(M);
  std::vector<int32_t> rlen(M);
  if (kgw_bgen_rsids(image.data(), (int64_t)image.size(), M, roff.data(), rlen.data()) != KGW_OK) { std::cerr << kgw_last_error() << std::endl; exit(1); }
  std::unordered_map<std::string, int32_t> where;                    // rsid -> position in file order (first occurrence, as std::find)
  for (int32_t v = 0; v < M; v++) where.emplace(std::string((const char *)image.data() + roff[v], (size_t)rlen[v]), v);
  std::vector<int32_t> variant_of_col(num_snps), sidx(sample_idx.begin(), sample_idx.end());
  for (int j = 0; j < num_snps; j++) {
    auto it = where.find(snp_ids[j]);
    if (it == where.end()) { cerr << "Error loading BGEN file. Some requested variants were not found."; exit(1); }   // :238-241
    variant_of_col[j] = it->second;
  }
  const int64_t row_bytes = (num_snps + 7) / 8;
  std::vector<uint8_t> Hflat((size_t)2 * num_samples * row_bytes);
  if (kgw_bgen_read(image.data(), (int64_t)image.size(), num_samples, sidx.data(), num_snps, variant_of_col.data(), nthreads,
                    /*device=*/0, row_bytes, Hflat.data(), nullptr) != KGW_OK) {
    std::cerr << "Error: " << kgw_last_error() << std::endl;         // incl. prob_to_hap's "Incorrect input" (:48-63)
    exit(1);
  }
  H.clear();
  H = vector<chaplotype>(2 * num_samples, chaplotype(num_snps));     // as BgenReader::read (:76-77)
  for (int i = 0; i < 2 * num_samples; i++)
    std::copy(Hflat.begin() + (size_t)i * row_bytes, Hflat.begin() + (size_t)(i + 1) * row_bytes, H[i].data.begin());
}

// HapsReader::read (haps_reader.cpp:37-132): same sample / variant look-ups, the text is tokenized and transposed on the device
void wrap_haps_read(HapsReader *self, const vector<string> &sample_ids, const vector<string> &snp_ids, vector<chaplotype> &H,
                    bool verbose) {
  if (!shim_io()) { real_haps_read(self, sample_ids, snp_ids, H, verbose); return; }
  std::ifstream fd(self->filename, std::ios::binary | std::ios::ate);
  if (!fd.is_open()) { real_haps_read(self, sample_ids, snp_ids, H, verbose); return; }   // the reference's own error path
  std::vector<uint8_t> text((size_t)fd.tellg());
  fd.seekg(0);
  fd.read((char *)text.data(), (std::streamsize)text.size());
  if (text.size() >= 2 && text[0] == 0x1f && text[1] == 0x8b) { real_haps_read(self, sample_ids, snp_ids, H, verbose); return; }   // gzip: the reference's ifile inflates
  const int num_samples = (int)sample_ids.size(), num_snps = (int)snp_ids.size();
  std::vector<int32_t> sidx(num_samples);
  for (int i = 0; i < num_samples; i++) {                              // :73-85
    auto it = std::find(self->sample.ID.begin(), self->sample.ID.end(), sample_ids[i]);
    if (it == self->sample.ID.end()) { cerr << "Error parsing HAPS file. Could not find sample " << sample_ids[i] << endl; exit(1); }
    sidx[i] = (int32_t)std::distance(self->sample.ID.begin(), it);
  }
  // a file row goes to the column of the first requested id equal to its legend id (:98-104); rows are visited in file
  // order, so the last such row wins a column
  std::unordered_map<std::string, int32_t> col_of;
  for (int j = num_snps - 1; j >= 0; j--) col_of[snp_ids[j]] = j;
  std::vector<int32_t> row_of_col(num_snps, -1);
  for (size_t r = 0; r < self->legend.ID.size(); r++) {
    auto it = col_of.find(self->legend.ID[r]);
    if (it != col_of.end()) row_of_col[it->second] = (int32_t)r;
  }
  for (int j = 0; j < num_snps; j++)
    if (row_of_col[j] < 0) { real_haps_read(self, sample_ids, snp_ids, H, verbose); return; }   // a column no row fills stays 0 in the reference
  const int64_t row_bytes = (num_snps + 7) / 8;
  std::vector<uint8_t> Hflat((size_t)2 * num_samples * row_bytes);
  if (kgw_haps_read(text.data(), (int64_t)text.size(), num_samples, sidx.data(), num_snps, row_of_col.data(), /*device=*/0, row_bytes,
                    Hflat.data(), nullptr) != KGW_OK) {
    std::cerr << "Error: " << kgw_last_error() << std::endl;
    exit(1);
  }
  H.clear();
  H = vector<chaplotype>(2 * num_samples, chaplotype(num_snps));
  for (int i = 0; i < 2 * num_samples; i++)
    std::copy(Hflat.begin() + (size_t)i * row_bytes, Hflat.begin() + (size_t)(i + 1) * row_bytes, H[i].data.begin());
}

#ifdef KGW_KC_SYM
// ---- `--compute-references`, clustering half: the bifurcating k-means (kinship_computer.cpp:511-691) with its two data
// passes -- update_assignments' distances and update_centroids (:239-250, :314-405) -- on the device (kgw_kinship_*).
// Everything else keeps the reference's order of operations: the 100 candidate pairs of init_centroids drawn from the
// reference's own putils::getRandom stream (:252-312), the size rebalancing through multimaps (:338-362), the acceptance
// of a split and the list of unfinished clusters (:620-676), the convergence test on the centroids (:593-617).  With
// several chromosomes the reference solves one leave-one-out problem per chromosome on racing threads that share
// that random stream; here they run one after the other (the reference's result with one thread).
namespace {
struct KinDevice {
  kgw_kinship *h = nullptr;
  std::vector<int> off;   // first concatenated locus of every chromosome (+ end)
  int P = 0;
};

void kin_fail() { std::cerr << "Error: " << kgw_last_error() << std::endl; exit(-1); }

void kin_use_mask(const KinDevice &D, int chr_out, std::vector<uint8_t> &use) {
  use.assign(D.P, 1);
  if (chr_out >= 0) std::fill(use.begin() + D.off[chr_out], use.begin() + D.off[chr_out + 1], 0);
}

// init_centroids (:252-312), genetic data only
void kin_init_centroids(const KinshipComputer *self, const KinDevice &D, int c, int chr_out, const vector<vector<int>> &clust_chr,
                        std::vector<double> &mu0, std::vector<double> &mu1) {
  const int clust_size = clust_chr[c].size();
  int dil_max = 0, i_best = 0, l_best = 0;
  for (int pr = 0; pr < 100; pr++) {
    int i = clust_chr[c][putils::getRandom(clust_size)];
    int i0 = 2 * (i / 2), i1 = 2 * (i / 2) + 1;
    int l = i;
    while ((l == i0) || (l == i1)) l = clust_chr[c][putils::getRandom(clust_size)];
    int result = 0;   // distance_leaveout (:106-124): an int accumulator
    for (int chr = 0; chr < self->num_chrs; chr++)
      if (chr != chr_out)
        result += self->chromosomes[chr]->distance(2 * (i / 2), 2 * (l / 2)) + self->chromosomes[chr]->distance(2 * (i / 2) + 1, 2 * (l / 2) + 1);
    int dil = result;
    if (dil > dil_max) { i_best = i; l_best = l; dil_max = dil; }
  }
  mu0.assign(D.P, 0.0);
  mu1.assign(D.P, 0.0);
  for (int chr = 0; chr < self->num_chrs; chr++) {
    if (chr == chr_out) continue;
    const int num_snps = self->chromosomes[chr]->get_num_snps();
    for (int j = 0; j < num_snps; j++) {
      mu0[D.off[chr] + j] = (double)(self->chromosomes[chr]->H[i_best][j]);
      mu1[D.off[chr] + j] = (double)(self->chromosomes[chr]->H[l_best][j]);
    }
  }
}

// distance(mu, mu_old) (:156-170): chromosome by chromosome, then their sum
double kin_centroid_shift(const KinDevice &D, int num_chrs, const std::vector<double> &a, const std::vector<double> &b) {
  double dist = 0.0;
  for (int chr = 0; chr < num_chrs; chr++) {
    double d = 0.0;
    for (int j = D.off[chr]; j < D.off[chr + 1]; j++) d += std::pow(a[j] - b[j], 2);
    dist += d;
  }
  return dist;
}

// kmeans_core (:593-617) with update_assignments (:314-362) and update_centroids (:368-405)
void kin_kmeans_core(const KinshipComputer *self, const KinDevice &D, int c, int chr_out, vector<vector<int>> &clust_chr,
                     vector<int> &new_assignments) {
  std::vector<double> mu0, mu1;
  kin_init_centroids(self, D, c, chr_out, clust_chr, mu0, mu1);
  std::vector<double> mu0_old = mu0, mu1_old = mu1;
  std::vector<uint8_t> use;
  kin_use_mask(D, chr_out, use);
  const int n = clust_chr[c].size();
  std::vector<int32_t> members(clust_chr[c].begin(), clust_chr[c].end()), asg(n);
  std::vector<double> d0(n), d1(n);
  const int cluster_size_min = self->cluster_size_min;
  for (int it = 0; it < 50; it++) {
    if (kgw_kinship_distances(D.h, members.data(), n, mu0.data(), mu1.data(), use.data(), d0.data(), d1.data()) != KGW_OK) kin_fail();
    std::multimap<double, int> distances0, distances1;
    int csize0 = 0, csize1 = 0;
    for (int idx = 0; idx < n; idx++) {
      const double dist_0 = d0[idx], dist_1 = d1[idx];
      if (dist_0 < dist_1) { new_assignments[idx] = 0; csize0++; distances0.emplace(dist_0, idx); }
      else { new_assignments[idx] = 1; csize1++; distances1.emplace(dist_1, idx); }
    }
    if (csize0 < cluster_size_min) {
      int n_missing = cluster_size_min - csize0;
      for (auto itr = distances1.rbegin(); itr != std::next(distances1.rbegin(), n_missing); ++itr) new_assignments[(*itr).second] = 0;
    }
    if (csize1 < cluster_size_min) {
      int n_missing = cluster_size_min - csize1;
      for (auto itr = distances0.rbegin(); itr != std::next(distances0.rbegin(), n_missing); ++itr) new_assignments[(*itr).second] = 1;
    }
    for (int idx = 0; idx < n; idx++) asg[idx] = new_assignments[idx];
    if (kgw_kinship_centroids(D.h, members.data(), asg.data(), n, mu0.data(), mu1.data(), nullptr, nullptr) != KGW_OK) kin_fail();
    if (chr_out >= 0) {   // the chromosome left out keeps zero centroids (:375-379)
      std::fill(mu0.begin() + D.off[chr_out], mu0.begin() + D.off[chr_out + 1], 0.0);
      std::fill(mu1.begin() + D.off[chr_out], mu1.begin() + D.off[chr_out + 1], 0.0);
    }
    const double delta_mu0 = kin_centroid_shift(D, self->num_chrs, mu0, mu0_old), delta_mu1 = kin_centroid_shift(D, self->num_chrs, mu1, mu1_old);
    mu0_old = mu0;
    mu1_old = mu1;
    if ((delta_mu0 < 1e-3) && (delta_mu1 < 1e-3)) break;
  }
}

// kmeans (:620-676)
void kin_kmeans(const KinshipComputer *self, const KinDevice &D, int c, int chr_out, vector<vector<int>> &clust_chr, vector<int> &assign_chr,
                vector<int> &unfinished_clusters) {
  const int clust_size = clust_chr[c].size();
  vector<int> new_assign_chr(clust_size, 0);
  kin_kmeans_core(self, D, c, chr_out, clust_chr, new_assign_chr);
  int clust_size_0 = 0, clust_size_1 = 0;
  for (int idx = 0; idx < clust_size; idx++) { if (new_assign_chr[idx] == 0) clust_size_0++; else clust_size_1++; }
  const int cluster_size_min = self->cluster_size_min;
  if (std::min(clust_size_0, clust_size_1) >= cluster_size_min) {
    int num_clust = clust_chr.size();
    vector<int> new_cluster_0, new_cluster_1;
    for (int idx = 0; idx < clust_size; idx++) {
      int i = clust_chr[c][idx];
      if (new_assign_chr[idx] == 0) { new_assign_chr[idx] = c; new_cluster_0.push_back(i); }
      else { new_assign_chr[idx] = num_clust; new_cluster_1.push_back(i); }
      assign_chr[i] = new_assign_chr[idx];
    }
    clust_chr.erase(clust_chr.begin() + c);
    clust_chr.insert(clust_chr.begin() + c, new_cluster_0);
    clust_chr.push_back(new_cluster_1);
    if (clust_size_0 <= cluster_size_min) {
      auto it = find(unfinished_clusters.begin(), unfinished_clusters.end(), c);
      unfinished_clusters.erase(it);
    }
    if (clust_size_1 >= cluster_size_min) unfinished_clusters.push_back(num_clust);
  }
}

// bifurcating_kmeans (:511-591)
void kin_bifurcating_kmeans(const KinshipComputer *self, const KinDevice &D, int chr_out, vector<int> &assign_chr, vector<vector<int>> &clust_chr) {
  const int num_haps = self->num_haps;
  assign_chr.resize(num_haps);
  std::fill(assign_chr.begin(), assign_chr.end(), 0);
  vector<int> clust_chr_tmp(num_haps);
  std::iota(clust_chr_tmp.begin(), clust_chr_tmp.end(), 0);
  clust_chr.clear();
  clust_chr.push_back(clust_chr_tmp);
  vector<int> unfinished_clusters;
  unfinished_clusters.push_back(0);
  int step = 0;
  const int max_steps = 10000;
  while ((unfinished_clusters.size() > 0) && (step++ < max_steps)) {
    int c = unfinished_clusters[0];
    if ((int)clust_chr[c].size() <= self->cluster_size_max) {
      auto it = find(unfinished_clusters.begin(), unfinished_clusters.end(), c);
      unfinished_clusters.erase(it);
      continue;
    }
    kin_kmeans(self, D, c, chr_out, clust_chr, assign_chr, unfinished_clusters);
  }
  cerr << "Bifurcating K-means on the GPU";
  if (chr_out != -1) cerr << " (leaving out chromosome " << chr_out + 1 << ")";
  cerr << ": " << clust_chr.size() << " clusters after " << step - 1 << " steps." << endl;
}
}  // namespace

void wrap_kc_ctor(KinshipComputer *self, const vector<Metadata> &metadata, int compression, int cluster_size_min, int cluster_size_max,
                  int num_threads) {
  if (!shim_io()) { real_kc_ctor(self, metadata, compression, cluster_size_min, cluster_size_max, num_threads); return; }
  new (self) KinshipComputer(metadata);                 // the members, default-constructed (:9-10)
  self->covariates_available = false;
  self->genomes = Dataset(metadata, num_threads, compression);
  self->num_chrs = self->genomes.num_chrs();
  for (int chr = 0; chr < self->num_chrs; chr++) self->chromosomes.push_back(&(self->genomes.chromosomes[chr]));
  self->num_haps = self->chromosomes[0]->get_num_haps();
  self->cluster_size_min = cluster_size_min;
  self->cluster_size_max = cluster_size_max;
  self->num_threads = num_threads;
  // the thinned rows of all chromosomes side by side, uploaded once
  KinDevice D;
  D.off.assign(1, 0);
  for (int chr = 0; chr < self->num_chrs; chr++) D.off.push_back(D.off.back() + self->chromosomes[chr]->get_num_snps());
  D.P = D.off.back();
  const int num_haps = self->num_haps;
  const int64_t row_bytes = (D.P + 7) / 8;
  std::vector<uint8_t> Hflat((size_t)num_haps * row_bytes, 0);
  for (int chr = 0; chr < self->num_chrs; chr++) {
    const Haplotypes *hp = self->chromosomes[chr];
    const int ps = hp->get_num_snps();
    for (int i = 0; i < num_haps; i++)
      for (int j = 0; j < ps; j++)
        if (hp->H[i][j]) { const int g = D.off[chr] + j; Hflat[(size_t)i * row_bytes + (g >> 3)] |= (uint8_t)(1u << (g & 7)); }
  }
  if (kgw_kinship_create(num_haps, D.P, row_bytes, Hflat.data(), /*device=*/0, &D.h) != KGW_OK) kin_fail();
  // cluster(true) (:407-478)
  self->assignments.resize(self->num_chrs);
  self->clusters.resize(self->num_chrs);
  if (self->num_chrs > 1) {
    cout << "Solving " << self->num_chrs << " bifurcating K-means problems on the GPU." << endl;
    for (int chr = 0; chr < self->num_chrs; chr++) kin_bifurcating_kmeans(self, D, chr, self->assignments[chr], self->clusters[chr]);
  } else {
    kin_bifurcating_kmeans(self, D, -1, self->assignments[0], self->clusters[0]);
  }
  cout << endl;
  kgw_kinship_destroy(D.h);
}
#endif  // KGW_KC_SYM

// KinshipComputer::assign_references (kinship_computer.cpp:836-931): the worker loop over findNeighbors (:174-233, all-pairs
// Hamming distances inside the kinship cluster and the K smallest) runs on the device; the merge inside IBD families
// (combine_references_families, :765-835) and the sanity checks stay the reference's.  With --windows > 0 the reference
// path is kept (include/kgw.h says why).
ivector3d wrap_assign_references(const KinshipComputer *self, const int K, const int chr, const Windows &windows) {
  if (!shim_io() || windows.num_windows > 1) return real_assign_references(self, K, chr, windows);
  const int num_haps = self->num_haps;
  const Haplotypes *hp = self->chromosomes[chr];                      // the thinned chromosome the distances are taken on
  const int p = hp->get_num_snps();
  const int64_t row_bytes = (p + 7) / 8;
  std::vector<uint8_t> Hflat((size_t)num_haps * row_bytes);
  for (int i = 0; i < num_haps; i++)
    std::copy(hp->H[i].data.begin(), hp->H[i].data.begin() + row_bytes, Hflat.begin() + (size_t)i * row_bytes);
  std::vector<int32_t> clust_off{0}, clust_members, family_of(num_haps, -1);
  for (auto &c : self->clusters[chr]) {
    clust_members.insert(clust_members.end(), c.begin(), c.end());
    clust_off.push_back((int32_t)clust_members.size());
  }
  for (int i = 0; i < num_haps; i++) {
    const vector<int> &fam = self->metadata[chr].related_families[i];
    if (fam.size() > 1) family_of[i] = *std::min_element(fam.begin(), fam.end());
  }
  cout << "Assigning references for the whole chromosome on the GPU: " << endl;
  std::vector<int32_t> nn((size_t)num_haps * K, -1);
  if (kgw_assign_references(num_haps, p, row_bytes, Hflat.data(), K, /*compression=*/1, (int32_t)self->clusters[chr].size(), clust_off.data(),
                            clust_members.data(), family_of.data(), /*device=*/0, nn.data(), nullptr) != KGW_OK) {
    std::cerr << "Error: " << kgw_last_error() << std::endl;
    exit(-1);
  }
  ivector3d ref(num_haps, ivector2d(1, vector<int>(K, -1)));
  for (int i = 0; i < num_haps; i++)
    for (int k = 0; k < K; k++) ref[i][0][k] = nn[(size_t)i * K + k];
  ref = self->combine_references_families(chr, ref);                   // unchanged (:765-835)
  for (int i1 = 0; i1 < num_haps; i1++) {                              // the reference's sanity checks (:883-928)
    const vector<int> &ibd_family = self->metadata[chr].related_families[i1];
    if (ibd_family.size() > 1)
      for (int i2 : ibd_family)
        if (ref[i1][0] != ref[i2][0]) { cerr << "Error: references for haplotypes " << i1 << " and " << i2 << " do not match!" << endl; exit(-1); }
    for (int k = 0; k < K; k++) {
      const int j = ref[i1][0][k];
      if (j < 0) { cerr << "Error: could not assign all references for lack of sufficient samples. Try decreasing K." << endl; exit(-1); }
      if (ibd_family.size() > 1 && std::count(ibd_family.begin(), ibd_family.end(), j) > 0) {
        cerr << "Error: could not assign all references for lack of sufficient unrelated samples. Try decreasing K." << endl;
        exit(-1);
      }
    }
  }
  return ref;
}

// plink::write_binary(basename, metadata, H, Hk, random) (plink_interface.cpp:255-284): the .bed body comes from the device
void wrap_write_binary(const string &basename, const Metadata &metadata, const vector<chaplotype> &H, const vector<chaplotype> &Hk,
                       bool random) {
  if (!shim_io()) { real_write_binary(basename, metadata, H, Hk, random); return; }
  const int num_haps = (int)H.size(), num_snps = (int)H[0].size();
  int seed = 2019;                                                   // coin flips as before (:262-274)
  std::mt19937_64 rng(seed);
  std::uniform_int_distribution<int> coin_flip(0, 1);
  vector<bool> swap(num_snps, 0);
  std::vector<uint8_t> swap8(num_snps);
  for (int j = 0; j < num_snps; j++) { swap[j] = coin_flip(rng); swap8[j] = swap[j] ? 1 : 0; }
  const int64_t row_bytes = (num_snps + 7) / 8;
  std::vector<uint8_t> Hflat((size_t)num_haps * row_bytes), Hkflat(Hflat.size());
  for (int i = 0; i < num_haps; i++) {
    std::copy(H[i].data.begin(), H[i].data.begin() + row_bytes, Hflat.begin() + (size_t)i * row_bytes);
    std::copy(Hk[i].data.begin(), Hk[i].data.begin() + row_bytes, Hkflat.begin() + (size_t)i * row_bytes);
  }
  std::vector<uint8_t> body((size_t)2 * num_snps * kgw_bed_column_bytes(num_haps));
  if (kgw_bed_encode(num_haps, num_snps, row_bytes, Hflat.data(), Hkflat.data(), swap8.data(), /*device=*/0, body.data()) != KGW_OK) {
    std::cerr << "Error: " << kgw_last_error() << std::endl;
    exit(-1);
  }
  std::ofstream file(basename + ".bed", std::ios_base::binary);
  if (!file.is_open()) {
    cerr << "Problem creating the output file: " << basename << ".bed";
    cerr << "Either the directory does not exist or you do not have write permissions." << endl;
    exit(1);
  }
  const unsigned char magic[3] = {0x6c, 0x1b, 0x01};               // :56
  file.write((const char *)magic, 3);
  file.write((const char *)body.data(), (std::streamsize)body.size());
  file.close();
  writeBIM(basename, metadata, true, swap);                          // unchanged text files
  writeFAM(basename, metadata);
  cout << "Output (binary) written to:" << endl;
  cout << "  " << basename << ".bed" << endl;
  cout << "  " << basename << ".bim" << endl;
  cout << "  " << basename << ".fam" << endl;
}

namespace {

template <typename T>
void dump(const std::string &dir, const char *name, const std::vector<T> &v) {
  FILE *f = fopen((dir + "/" + name).c_str(), "wb");
  if (!f) return;
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

int g_call = 0;

}  // namespace

void wrap_generate(Knockoffs *self, int /*num_threads*/) {
  Knockoffs &k = *self;
  k.reset_knockoffs();                                   // unchanged: alpha := 1/K, Hk := 0  (knockoffs.cpp:72)
  const int num_haps = k.num_haps, num_snps = k.num_snps, K = k.K;
  const int num_windows = k.metadata.windows.num_windows;
  const size_t row_bytes = (size_t)(num_snps + 7) / 8;

  // chaplotype rows -> one flat bit matrix (same LSB-first byte layout, chaplotype.h:25)
  std::vector<uint8_t> Hflat((size_t)num_haps * row_bytes, 0), Hkflat(Hflat.size(), 0);
  for (int i = 0; i < num_haps; i++)
    std::copy(k.H[i].data.begin(), k.H[i].data.begin() + row_bytes, Hflat.begin() + (size_t)i * row_bytes);

  std::vector<int32_t> rl((size_t)num_haps * num_windows * K), rg((size_t)num_haps * K);
  for (int i = 0; i < num_haps; i++) {
    for (int w = 0; w < num_windows; w++)
      std::copy(k.references_local[i][w].begin(), k.references_local[i][w].end(), rl.begin() + ((size_t)i * num_windows + w) * K);
    std::copy(k.references_global[i].begin(), k.references_global[i].end(), rg.begin() + (size_t)i * K);
  }
  std::vector<int32_t> grp(k.groups.begin(), k.groups.end());
  std::vector<int32_t> ws(k.metadata.windows.start.begin(), k.metadata.windows.start.end());
  std::vector<int32_t> we(k.metadata.windows.end.begin(), k.metadata.windows.end.end());
  std::vector<int32_t> unrel(k.metadata.unrelated.begin(), k.metadata.unrelated.end());   // std::set: ascending
  std::vector<double> b(k.b.begin(), k.b.end()), mu(k.mut_rates.begin(), k.mut_rates.end());

  kgw_problem P = {};
  P.abi_version = KGW_ABI_VERSION;
  P.N = num_haps;  P.p = num_snps;  P.K = K;  P.W = num_windows;
  P.row_bytes = (int64_t)row_bytes;
  P.H = Hflat.data();  P.ref_local = rl.data();  P.ref_global = rg.data();
  P.b = b.data();      P.mu = mu.data();  P.groups = grp.data();
  P.win_start = ws.data();  P.win_end = we.data();
  P.unrelated = unrel.data();  P.n_unrelated = (int32_t)unrel.size();
  P.seed = (uint64_t)k.seed;

  // related haplotypes (generate_related, knockoffs.cpp:329-571): the host still dedups the families
  // (:331-342) and tidies their IBD segments (IbdCluster, :474-481, ibd.cpp:99-189); the device runs
  // FamilyBP, FamilyKnockoffs and sample_emission on them
  std::set<std::vector<int>> fams;
  for (auto &f : k.metadata.related_families)
    if (f.size() > 1) fams.insert(std::vector<int>(f.begin(), f.end()));
  std::vector<int32_t> fam_off{0}, fam_mem, seg_off{0}, seg_jmin, seg_jmax, seg_ids_off{0}, seg_ids;
  for (auto &fam : fams) {
    fam_mem.insert(fam_mem.end(), fam.begin(), fam.end());
    fam_off.push_back((int32_t)fam_mem.size());
    std::set<IbdSeg> segs;
    for (int i : fam)
      for (auto &sg : k.metadata.ibd_segments[i]) segs.insert(sg);
    IbdCluster cluster(segs);                            // constructor runs tidy()
    for (int s = 0; s < cluster.size(); s++) {
      IbdSeg sg;
      cluster.get(s, sg);
      seg_jmin.push_back(sg.j_min);
      seg_jmax.push_back(sg.j_max);
      seg_ids.insert(seg_ids.end(), sg.indices.begin(), sg.indices.end());
      seg_ids_off.push_back((int32_t)seg_ids.size());
    }
    seg_off.push_back((int32_t)seg_jmin.size());
  }
  kgw_families F = {};
  F.n_families = (int32_t)fams.size();
  F.fam_off = fam_off.data();  F.fam_members = fam_mem.data();  F.seg_off = seg_off.data();
  F.seg_jmin = seg_jmin.data();  F.seg_jmax = seg_jmax.data();
  F.seg_ids_off = seg_ids_off.data();  F.seg_ids = seg_ids.data();
  if (F.n_families > 0) P.families = &F;

  if (const char *dd = getenv("KGW_SHIM_DUMP")) {
    const std::string dir = std::string(dd) + "/call" + std::to_string(g_call);
    mkdir(dd, 0777);
    mkdir(dir.c_str(), 0777);
    std::vector<int64_t> dims = {num_haps, num_snps, K, num_windows, (int64_t)row_bytes, (int64_t)k.seed};
    dump(dir, "dims.i64", dims);
    dump(dir, "H.u8", Hflat);
    dump(dir, "ref_local.i32", rl);
    dump(dir, "ref_global.i32", rg);
    dump(dir, "b.f64", b);
    dump(dir, "mu.f64", mu);
    dump(dir, "groups.i32", grp);
    dump(dir, "win_start.i32", ws);
    dump(dir, "win_end.i32", we);
    dump(dir, "unrelated.i32", unrel);
    dump(dir, "fam_off.i32", fam_off);
    dump(dir, "fam_members.i32", fam_mem);
    dump(dir, "seg_off.i32", seg_off);
    dump(dir, "seg_jmin.i32", seg_jmin);
    dump(dir, "seg_jmax.i32", seg_jmax);
    dump(dir, "seg_ids_off.i32", seg_ids_off);
    dump(dir, "seg_ids.i32", seg_ids);
  }
  g_call++;

  // One resident plan per Knockoffs object: main.cpp:159-181 calls set_partition(r) + generate() once per resolution
  // on the same object, so H, the references and the scratch stay in HBM and only the per-locus knockoff tables
  // change (kgw_plan_set_groups).  The plan is keyed by a fingerprint of everything except the partition.
  // Devices: every visible GPU (the counterpart of --n_threads), or KGW_DEVICES="0,2,3".
  uint64_t fp = 1469598103934665603ull;
  auto mix = [&](const void *ptr, size_t bytes) {
    const uint64_t *w = (const uint64_t *)ptr;
    for (size_t i = 0; i < bytes / 8; i++) { fp ^= w[i]; fp *= 1099511628211ull; }
    const uint8_t *t = (const uint8_t *)ptr + (bytes / 8) * 8;
    for (size_t i = 0; i < bytes % 8; i++) { fp ^= t[i]; fp *= 1099511628211ull; }
  };
  const int64_t dims[6] = {num_haps, num_snps, K, num_windows, (int64_t)k.seed, (int64_t)F.n_families};
  mix(dims, sizeof(dims));
  mix(Hflat.data(), Hflat.size());
  mix(rl.data(), rl.size() * 4);  mix(rg.data(), rg.size() * 4);
  mix(b.data(), b.size() * 8);    mix(mu.data(), mu.size() * 8);
  mix(ws.data(), ws.size() * 4);  mix(we.data(), we.size() * 4);
  mix(unrel.data(), unrel.size() * 4);
  mix(fam_mem.data(), fam_mem.size() * 4);  mix(seg_jmin.data(), seg_jmin.size() * 4);  mix(seg_jmax.data(), seg_jmax.size() * 4);
  mix(seg_ids.data(), seg_ids.size() * 4);
  static kgw_plan *plan = nullptr;
  static uint64_t plan_fp = 0;
  static Knockoffs *plan_owner = nullptr;
  int rc = KGW_OK;
  if (plan && plan_owner == self && plan_fp == fp) {
    rc = kgw_plan_set_groups(plan, grp.data());
  } else {
    if (plan) { kgw_plan_destroy(plan); plan = nullptr; }
    std::vector<int32_t> devs;
    if (const char *dv = getenv("KGW_DEVICES")) {
      std::stringstream ss(dv);
      std::string tok;
      while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
    } else {
      for (int d = 0; d < kgw_device_count(); d++) devs.push_back(d);
    }
    kgw_options O = {};
    O.n_devices = (int32_t)devs.size();
    O.devices = devs.data();
    rc = kgw_plan_create(&P, &O, &plan);
    plan_fp = fp;
    plan_owner = self;
    static bool registered = false;
    if (!registered) { registered = true; atexit([]() { if (plan) { kgw_plan_destroy(plan); plan = nullptr; } kgw_release_workspace(); }); }
  }
  if (rc == KGW_OK) rc = kgw_plan_run(plan);
  if (rc == KGW_OK) rc = kgw_plan_download(plan, Hkflat.data(), nullptr);
  if (rc != KGW_OK) {
    std::cerr << "Error: " << kgw_last_error() << std::endl;      // same convention as throw_error_bug (utils.cpp:700)
    exit(-1);
  }
  // KGW_AUDIT_HAPS=n: n unrelated haplotypes, evenly spaced, are sampled again in the reference's own summation order
  // (kgw_generate_strict: sequential sums over k, no fused multiply-adds) and must equal the rows just generated
  if (const char *au = getenv("KGW_AUDIT_HAPS")) {
    const int want = std::atoi(au);
    if (want > 0 && !unrel.empty()) {
      const int n = std::min<int>(want, (int)unrel.size());
      std::vector<int32_t> pick(n);
      for (int t = 0; t < n; t++) pick[t] = unrel[(size_t)t * unrel.size() / n];
      kgw_problem A = P;
      A.unrelated = pick.data();
      A.n_unrelated = n;
      A.families = nullptr;
      std::vector<uint8_t> Hs(Hflat.size(), 0);
      const int dev0 = getenv("KGW_DEVICES") ? std::atoi(getenv("KGW_DEVICES")) : 0;
      if (kgw_generate_strict(&A, dev0, Hs.data(), nullptr) != KGW_OK) {
        std::cerr << "Error: " << kgw_last_error() << std::endl;
        exit(-1);
      }
      int differing = 0;
      for (int i : pick)
        differing += std::memcmp(Hs.data() + (size_t)i * row_bytes, Hkflat.data() + (size_t)i * row_bytes, row_bytes) != 0;
      std::cout << "[kgw] audit: " << n << " haplotypes sampled again in the reference's summation order, " << differing
                << " differ" << std::endl;
      if (differing) {
        std::cerr << "Error: the audit of the knockoff sampler found differing haplotypes" << std::endl;
        exit(-1);
      }
    }
  }
  auto take_row = [&](int i) {                           // == set_Hk(i, hk)  (knockoffs.cpp:299)
    k.Hk[i].data.assign(Hkflat.begin() + (size_t)i * row_bytes, Hkflat.begin() + (size_t)(i + 1) * row_bytes);
  };
  for (int i : k.metadata.unrelated) take_row(i);
  for (int i : fam_mem) take_row(i);
}
