// oracle/ref_shims/wrap_probe.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Linked (via GNU ld --wrap, no source patch) in front of two functions of the
// UNMODIFIED reference objects built by oracle/build_ref.sh:
//
//   Knockoffs::generate(int)        snpknock2/src/knockoffs.cpp:311
//   weighted_choice(weights, rng)   snpknock2/src/utils.cpp:690
//
// The generate() wrapper (a) times the real call with std::chrono -- this is
// the CPU-baseline clock BASELINE.md section 4 asks for -- and (b), when
// KGW_PROBE_DIR is set, dumps everything the unchanged host code hands to the
// hot path at full binary precision before the call
// and the haplotype-level knockoffs Hk after it.  The reference itself only
// ever writes genotype-level BED, so this dump is how golden haplotype-level
// vectors are obtained from the real implementation.
//
// The weighted_choice() wrapper forwards to the real function and, when
// KGW_PROBE_TRACE=1 (single-threaded runs only), records (size, result) of
// every categorical draw: p draws of size K (Z), p of size K (Zk), p of size 2
// (Hk) per unrelated haplotype (knockoffs.cpp:1243,1259,1421,1500).
//
// This file contains no reference code: it only reads public/private members
// through the reference's own headers.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>
#include <vector>
#include <sstream>
#include <fstream>
#include <iostream>
#include <set>
#include <map>
#include <algorithm>
#include <numeric>
#include <sys/stat.h>

#define private public
#define protected public
#include "knockoffs.h"
#undef private
#undef protected

#ifndef KGW_GEN_SYM
#error "KGW_GEN_SYM (mangled Knockoffs::generate) must be defined by the build script"
#endif
#if !defined(KGW_WC_SYM) && !defined(KGW_TIMER_ONLY)
#error "KGW_WC_SYM (mangled weighted_choice) must be defined by the build script"
#endif
#define KGW_STR2(x) #x
#define KGW_STR(x) KGW_STR2(x)

void real_generate(Knockoffs *self, int num_threads) asm("__real_" KGW_STR(KGW_GEN_SYM));
void wrap_generate(Knockoffs *self, int num_threads) asm("__wrap_" KGW_STR(KGW_GEN_SYM));
#ifndef KGW_TIMER_ONLY   // snpknock2_timer (the binary bench.py times) wraps Knockoffs::generate only
int real_weighted_choice(const std::vector<double> &w, boost::random::taus88 &rng) asm("__real_" KGW_STR(KGW_WC_SYM));
int wrap_weighted_choice(const std::vector<double> &w, boost::random::taus88 &rng) asm("__wrap_" KGW_STR(KGW_WC_SYM));
#endif

namespace {

bool g_trace = false;
std::vector<int32_t> g_trace_buf;
int g_call_index = 0;

template <class T>
void write_raw(const std::string &path, const std::vector<T> &v) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "wrap_probe: cannot write %s\n", path.c_str()); exit(2); }
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

void dump_rows(const std::string &path, const std::vector<chaplotype> &rows, int p) {
  const size_t row_bytes = (size_t)(p + 7) / 8;
  std::vector<uint8_t> out(rows.size() * row_bytes, 0);
  for (size_t i = 0; i < rows.size(); i++) {
    const std::vector<unsigned char> &d = rows[i].data;
    for (size_t c = 0; c < row_bytes && c < d.size(); c++) out[i * row_bytes + c] = d[c];
  }
  write_raw(path, out);
}

}  // namespace

#ifndef KGW_TIMER_ONLY
int wrap_weighted_choice(const std::vector<double> &w, boost::random::taus88 &rng) {
  const int r = real_weighted_choice(w, rng);
  if (g_trace) {
    g_trace_buf.push_back((int32_t)w.size());
    g_trace_buf.push_back((int32_t)r);
  }
  return r;
}
#endif

void wrap_generate(Knockoffs *self, int num_threads) {
  const char *dir_env = getenv("KGW_PROBE_DIR");
  const char *trace_env = getenv("KGW_PROBE_TRACE");
  std::string dir;
  if (dir_env && *dir_env) {
    std::ostringstream ss;
    ss << dir_env << "/call" << g_call_index;
    dir = ss.str();
    mkdir(dir_env, 0755);
    mkdir(dir.c_str(), 0755);
  }
  g_trace = (!dir.empty() && trace_env && trace_env[0] == '1' && num_threads == 1);
  g_trace_buf.clear();

  const int N = self->num_haps, p = self->num_snps, K = self->K, W = self->num_windows;

  if (!dir.empty()) {
    dump_rows(dir + "/H.u8", self->H, p);
    std::vector<int32_t> rl((size_t)N * W * K), rg((size_t)N * K);
    for (int i = 0; i < N; i++) {
      for (int w = 0; w < W; w++)
        for (int k = 0; k < K; k++) rl[((size_t)i * W + w) * K + k] = self->references_local[i][w][k];
      for (int k = 0; k < K; k++) rg[(size_t)i * K + k] = self->references_global[i][k];
    }
    write_raw(dir + "/ref_local.i32", rl);
    write_raw(dir + "/ref_global.i32", rg);
    write_raw(dir + "/b.f64", self->b);
    write_raw(dir + "/mu.f64", self->mut_rates);
    std::vector<int32_t> groups(self->groups.begin(), self->groups.end());
    write_raw(dir + "/groups.i32", groups);
    const Windows &win = self->metadata.windows;
    write_raw(dir + "/win_start.i32", std::vector<int32_t>(win.start.begin(), win.start.end()));
    write_raw(dir + "/win_end.i32", std::vector<int32_t>(win.end.begin(), win.end.end()));
    write_raw(dir + "/win_of.i32", std::vector<int32_t>(win.window.begin(), win.window.end()));
    write_raw(dir + "/unrelated.i32",
              std::vector<int32_t>(self->metadata.unrelated.begin(), self->metadata.unrelated.end()));
    write_raw(dir + "/related.i32",
              std::vector<int32_t>(self->metadata.related.begin(), self->metadata.related.end()));
    std::vector<int32_t> fam_off(1, 0), fam_mem;
    for (size_t f = 0; f < self->metadata.related_families.size(); f++) {
      for (int i : self->metadata.related_families[f]) fam_mem.push_back(i);
      fam_off.push_back((int32_t)fam_mem.size());
    }
    write_raw(dir + "/fam_offsets.i32", fam_off);
    write_raw(dir + "/fam_members.i32", fam_mem);
    // IBD segments exactly as metadata holds them (per haplotype, pre-tidy):
    // record = hap, j_min, j_max, bp_min, bp_max, n, idx[0..n)
    std::vector<int32_t> ibd;
    for (size_t i = 0; i < self->metadata.ibd_segments.size(); i++) {
      for (const IbdSeg &s : self->metadata.ibd_segments[i]) {
        ibd.push_back((int32_t)i);
        ibd.push_back(s.j_min); ibd.push_back(s.j_max);
        ibd.push_back(s.bp_min); ibd.push_back(s.bp_max);
        ibd.push_back((int32_t)s.indices.size());
        for (int id : s.indices) ibd.push_back(id);
      }
    }
    write_raw(dir + "/ibd_segments.i32", ibd);
    std::vector<int32_t> bp(self->metadata.legend_filter.bp.begin(), self->metadata.legend_filter.bp.end());
    write_raw(dir + "/bp.i32", bp);
    // Families in the order generate_related() processes them with one thread (knockoffs.cpp:331-372:
    // std::set of member lists -> vector -> std::sort by size, descending), and for each family the
    // tidied IBD segments exactly as generate_related_cluster() hands them to FamilyBP / FamilyKnockoffs
    // (knockoffs.cpp:474-481): built here by calling the reference's own IbdCluster constructor.
    // record per family: n_members, members..., n_segments, then per segment
    //   j_min, j_max, bp_min, bp_max, n_ids, ids...   (segment order = std::set<IbdSeg> iteration order)
    {
      std::set<std::vector<int> > famset;
      for (size_t f = 0; f < self->metadata.related_families.size(); f++) {
        if (self->metadata.related_families[f].size() <= 1) continue;
        famset.insert(self->metadata.related_families[f]);
      }
      std::vector<std::vector<int> > famvec(famset.begin(), famset.end());
      std::sort(famvec.begin(), famvec.end(),
                [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
      std::vector<int32_t> out;
      out.push_back((int32_t)famvec.size());
      for (const std::vector<int> &fam : famvec) {
        out.push_back((int32_t)fam.size());
        for (int i : fam) out.push_back(i);
        std::set<IbdSeg> segs;
        for (int i : fam)
          for (const IbdSeg &sg : self->metadata.ibd_segments[i]) segs.insert(sg);
        IbdCluster cluster(segs);
        out.push_back((int32_t)cluster.size());
        for (int t = 0; t < cluster.size(); t++) {
          IbdSeg sg;
          cluster.get(t, sg);
          out.push_back(sg.j_min); out.push_back(sg.j_max);
          out.push_back(sg.bp_min); out.push_back(sg.bp_max);
          out.push_back((int32_t)sg.indices.size());
          for (int id : sg.indices) out.push_back(id);
        }
      }
      write_raw(dir + "/fam_clusters.i32", out);
    }
  }

  const auto t0 = std::chrono::steady_clock::now();
  real_generate(self, num_threads);
  const auto t1 = std::chrono::steady_clock::now();
  const double secs = std::chrono::duration<double>(t1 - t0).count();
  fprintf(stderr, "KGW_PROBE generate call=%d chr=%s N=%d p=%d K=%d W=%d groups=%d threads=%d seconds=%.6f\n",
          g_call_index, self->metadata.get_chr_id().c_str(), N, p, K, W, self->num_groups(), num_threads, secs);

  if (!dir.empty()) {
    dump_rows(dir + "/Hk.u8", self->Hk, p);
    if (g_trace) write_raw(dir + "/trace.i32", g_trace_buf);
    std::ofstream meta((dir + "/meta.txt").c_str());
    meta << "N " << N << "\np " << p << "\nK " << K << "\nW " << W << "\nseed " << self->seed
         << "\nnum_threads " << num_threads << "\nchr " << self->metadata.get_chr_id()
         << "\nnum_groups " << self->num_groups() << "\nn_unrelated " << self->metadata.unrelated.size()
         << "\nn_related " << self->metadata.related.size() << "\ntraced " << (g_trace ? 1 : 0)
         << "\nseconds " << secs << "\n";
  }
  g_trace = false;
  g_trace_buf.clear();
  g_call_index++;
}
