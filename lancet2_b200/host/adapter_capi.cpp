// adapter_capi.cpp — C hooks around lancet_gpu::GpuGenotyper so that the pytest suite can
// drive the C++ adapter (ctypes) without a C++ test runner.  Test support only; the product
// boundary is include/lancet_gpu_realign.h + host/gpu_genotyper.h.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <deque>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "gpu_genotyper.h"

namespace {

void AppI(std::string& s, long long v) { char b[32]; std::snprintf(b, sizeof(b), "%lld,", v); s += b; }
void AppD(std::string& s, double v) { char b[48]; std::snprintf(b, sizeof(b), "%a,", v); s += b; }

// plain snprintf/std::string on purpose (no iostreams)
void DumpAllele(std::string& s, const lancet_gpu::PerAlleleData& d) {
  s += "fwdbq:";
  for (auto v : d.mFwdBaseQuals) AppI(s, v);
  s += "|revbq:";
  for (auto v : d.mRevBaseQuals) AppI(s, v);
  s += "|mapq:";
  for (auto v : d.mMapQuals) AppI(s, v);
  s += "|aln:";
  for (auto v : d.mAlnScores) AppD(s, v);
  s += "|isz:";
  for (auto v : d.mProperPairIsizes) AppD(s, v);
  s += "|fold:";
  for (auto v : d.mFoldedReadPositions) AppD(s, v);
  s += "|refnm:";
  for (auto v : d.mRefNmValues) AppD(s, v);
  s += "|ownnm:";
  for (auto v : d.mOwnHapNmValues) AppD(s, v);
  s += "|starts:";
  for (auto v : d.mAlignmentStarts) AppI(s, v);
  s += "|hapids:";
  for (auto v : d.mHaplotypeIds) AppI(s, v);
  s += "|sc:" + std::to_string(d.mSoftClipCount) + "|hashes:";
  std::vector<std::pair<std::uint32_t, int>> hs;
  for (auto& kv : d.mNameHashes) hs.emplace_back(kv.first, (int)(kv.second == lancet_gpu::Strand::REV));
  std::sort(hs.begin(), hs.end());
  for (auto& kv : hs) s += std::to_string(kv.first) + ":" + std::to_string(kv.second) + ",";
}

int WriteOut(const std::string& s, char* out, long long cap) {
  if ((long long)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

}  // namespace

extern "C" {

// Feed one evidence stream through lancet_gpu::VariantSupport::AddEvidence and dump the
// per-allele vectors (same text format as oracle/ref_shim/ref_capi.cpp:ref_add_evidence_dump).
int lgr_adapter_add_evidence_dump(int n, const long long* isize, const long long* start, const double* aln,
                                  const double* fold, const unsigned* hash, const unsigned* ref_nm, const unsigned* own_nm,
                                  const unsigned* hap_id, const unsigned char* allele, const unsigned char* rev,
                                  const unsigned char* bq, const unsigned char* mapq, const unsigned char* softclip,
                                  const unsigned char* proper, char* out, long long cap) {
  lancet_gpu::VariantSupport vs;
  for (int i = 0; i < n; ++i) {
    lancet_gpu::ReadEvidence ev;
    ev.mInsertSize = isize[i], ev.mAlignmentStart = start[i], ev.mAlnScore = aln[i], ev.mFoldedReadPos = fold[i];
    ev.mRnameHash = hash[i], ev.mRefNm = ref_nm[i], ev.mOwnHapNm = own_nm[i], ev.mAssignedHaplotypeId = hap_id[i];
    ev.mAllele = allele[i], ev.mStrand = rev[i] ? lancet_gpu::Strand::REV : lancet_gpu::Strand::FWD;
    ev.mBaseQual = bq[i], ev.mMapQual = mapq[i], ev.mIsSoftClipped = softclip[i] != 0, ev.mIsProperPair = proper[i] != 0;
    vs.AddEvidence(ev);
  }
  std::string os;
  for (std::size_t a = 0; a < vs.AlleleData().size(); ++a) {
    os += "A" + std::to_string(a) + "|";
    DumpAllele(os, vs.AlleleData()[a]);
    os += "\n";
  }
  return WriteOut(os, out, cap);
}

}  // extern "C"

namespace {

// Rebuilds the C++-side view (strings, ReadIn, VariantIn, one GenotypeJob per group) of a batch
// given in the C-ABI's SoA form plus the per-read metadata AddToTable needs.
struct JobSet {
  std::vector<std::string_view> qn, sn;
  std::vector<std::string> haps;
  std::vector<lancet_gpu::ReadIn> reads;
  std::vector<lancet_gpu::VariantIn> vars;
  std::vector<lancet_gpu::GenotypeJob> jobs;
};

void BuildJobs(const lgr_batch_in* in, const char* names, const char* samples, const int* sample_id, const long long* start0,
               const long long* isize, const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
               JobSet& js) {
  for (const char* p = names; (int)js.qn.size() < in->n_reads; p += std::strlen(p) + 1) js.qn.emplace_back(p);
  int n_samples = 0;
  for (int r = 0; r < in->n_reads; ++r) n_samples = std::max(n_samples, sample_id[r] + 1);
  for (const char* p = samples; (int)js.sn.size() < n_samples; p += std::strlen(p) + 1) js.sn.emplace_back(p);
  js.haps.resize(in->n_haps);
  for (int h = 0; h < in->n_haps; ++h)
    js.haps[h].assign(reinterpret_cast<const char*>(in->hap_bases) + in->hap_off[h], (size_t)(in->hap_off[h + 1] - in->hap_off[h]));
  js.reads.resize(in->n_reads);
  for (int r = 0; r < in->n_reads; ++r) {
    const size_t len = (size_t)(in->read_off[r + 1] - in->read_off[r]);
    js.reads[r] = lancet_gpu::ReadIn{js.qn[r], std::string_view(reinterpret_cast<const char*>(in->read_bases) + in->read_off[r], len),
                                     in->read_quals + in->read_off[r], js.sn[sample_id[r]], start0[r], isize[r], sam_flag[r],
                                     mapq[r], softclip[r] != 0};
  }
  js.vars.resize(in->n_vars);
  for (int g = 0; g < in->n_groups; ++g) {
    const int h0 = in->grp_hap_begin[g], P = in->grp_hap_begin[g + 1] - h0;
    for (int v = in->grp_var_begin[g]; v < in->grp_var_begin[g + 1]; ++v) {
      lancet_gpu::VariantIn& vi = js.vars[v];
      vi.key = &js.vars[v];
      const long long o = in->var_hap_off[v];
      vi.local_ref_start0 = (size_t)in->var_start[o], vi.ref_allele_len = (size_t)in->var_len[o];
      int max_al = 0;
      for (int h = 1; h < P; ++h) max_al = std::max(max_al, (int)in->var_allele[o + h]);
      vi.alts.resize(max_al);
      for (int h = 1; h < P; ++h) {
        const int al = in->var_allele[o + h];
        if (al <= 0) continue;
        vi.alts[al - 1].seq_len = (size_t)in->var_len[o + h];
        vi.alts[al - 1].length = (long long)in->var_len[o + h] - (long long)vi.ref_allele_len;
        vi.alts[al - 1].hap_start0.emplace_back((size_t)h, (size_t)in->var_start[o + h]);
      }
    }
    js.jobs.push_back(lancet_gpu::GenotypeJob{js.haps.data() + h0, (size_t)P, js.reads.data() + in->grp_read_begin[g],
                                              (size_t)(in->grp_read_begin[g + 1] - in->grp_read_begin[g]),
                                              js.vars.data() + in->grp_var_begin[g],
                                              (size_t)(in->grp_var_begin[g + 1] - in->grp_var_begin[g])});
  }
}

std::string DumpResults(const lgr_batch_in* in, const JobSet& js, const std::vector<lancet_gpu::Result>& res) {
  std::string os;
  for (int g = 0; g < in->n_groups; ++g) {
    for (int v = in->grp_var_begin[g]; v < in->grp_var_begin[g + 1]; ++v) {
      auto it = res[g].find(&js.vars[v]);
      if (it == res[g].end()) continue;
      for (const auto& ns : it->second) {
        const auto& ad = ns.mData->AlleleData();
        for (std::size_t a = 0; a < ad.size(); ++a) {
          os += "G" + std::to_string(g) + " V" + std::to_string(v - in->grp_var_begin[g]) + " S" + std::string(ns.mSampleName) +
                " A" + std::to_string(a) + "|";
          DumpAllele(os, ad[a]);
          os += "\n";
        }
      }
    }
  }
  return os;
}

std::uint32_t X31OfView(std::string_view q) { return lgr_x31_hash(std::string(q).c_str()); }

}  // namespace

extern "C" {

// Run GpuGenotyper::GenotypeMany on the batch; dump every (group, variant, sample, allele)
// evidence block.  names: NUL-separated read names; samples: NUL-separated sample names,
// sample_id per read.  The name hash is X31 of the name (deterministic stand-in for
// absl::HashOf in tests).
int lgr_adapter_genotype_dump(int device, const lgr_batch_in* in, const char* names, const char* samples,
                              const int* sample_id, const long long* start0, const long long* isize,
                              const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
                              char* out, long long cap) {
  try {
    JobSet js;
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    lancet_gpu::GpuGenotyper gt(device);
    std::vector<lancet_gpu::Result> res = gt.GenotypeMany(js.jobs, X31OfView);
    return WriteOut(DumpResults(in, js, res), out, cap);
  } catch (const std::exception& e) {
    std::snprintf(out, (size_t)cap, "EXCEPTION: %s", e.what());
    return -2;
  }
}

// Host logic without a GPU (CPU test suite): pack the batch's groups through PackedJob/PackedBatch
// and compare the resulting C-ABI SoA arrays with the arrays the batch came from (a round trip
// through the C++ view), then run AddToTable (PackedBatch::BuildResult) on caller-provided
// lgr_assign records (the oracle's, in the tests) and dump the evidence.  Returns the number of
// mismatching array elements of the packing round trip in *pack_mismatches.
int lgr_adapter_host_logic_dump(const lgr_batch_in* in, const char* names, const char* samples, const int* sample_id,
                                const long long* start0, const long long* isize, const unsigned short* sam_flag,
                                const unsigned char* mapq, const unsigned char* softclip, const lgr_assign* assign,
                                long long* pack_mismatches, char* out, long long cap) {
  try {
    JobSet js;
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    lancet_gpu::PackedBatch pb;
    pb.Pack(js.jobs.data(), js.jobs.size(), 0);
    const lgr_batch_in& p = pb.In();
    long long bad = 0;
    bad += p.n_groups != in->n_groups || p.n_haps != in->n_haps || p.n_reads != in->n_reads || p.n_vars != in->n_vars;
    if (!bad) {
      for (int g = 0; g <= in->n_groups; ++g)
        bad += p.grp_hap_begin[g] != in->grp_hap_begin[g] || p.grp_read_begin[g] != in->grp_read_begin[g] ||
               p.grp_var_begin[g] != in->grp_var_begin[g];
      for (int h = 0; h <= in->n_haps; ++h) bad += p.hap_off[h] != in->hap_off[h];
      for (int r = 0; r <= in->n_reads; ++r) bad += p.read_off[r] != in->read_off[r];
      for (int v = 0; v <= in->n_vars; ++v) bad += p.var_hap_off[v] != in->var_hap_off[v];
      bad += std::memcmp(p.hap_bases, in->hap_bases, (size_t)in->hap_off[in->n_haps]) != 0;
      bad += std::memcmp(p.read_bases, in->read_bases, (size_t)in->read_off[in->n_reads]) != 0;
      bad += std::memcmp(p.read_quals, in->read_quals, (size_t)in->read_off[in->n_reads]) != 0;
      for (int r = 0; r < in->n_reads; ++r) bad += p.read_name_hash[r] != in->read_name_hash[r];
      for (long long x = 0; x < in->var_hap_off[in->n_vars]; ++x)
        bad += p.var_start[x] != in->var_start[x] || p.var_len[x] != in->var_len[x] || p.var_allele[x] != in->var_allele[x];
      bad += p.grp_mid_occ != nullptr;  // latched value 0 → per-group derivation left to the device
    }
    if (pack_mismatches) *pack_mismatches = bad;
    std::vector<lancet_gpu::Result> res(js.jobs.size());
    long long off = 0;
    for (std::size_t g = 0; g < js.jobs.size(); ++g) {
      res[g] = lancet_gpu::PackedBatch::BuildResult(js.jobs[g], assign + off, X31OfView);
      off += (long long)(js.jobs[g].n_reads * js.jobs[g].n_variants);
    }
    return WriteOut(DumpResults(in, js, res), out, cap);
  } catch (const std::exception& e) {
    std::snprintf(out, (size_t)cap, "EXCEPTION: %s", e.what());
    return -2;
  }
}

// Same batch, but every group is a separate blocking GenotypeBatcher::Genotype() call issued
// from `n_threads` worker threads (round-robin over the groups, `rounds` times), the way
// Lancet2's workers would call it.  counters[9] = batches, jobs, pairs, max jobs in one batch,
// wall nanoseconds of the worker phase, batcher-thread nanoseconds in pack / submit / wait / deliver.  cap <= 0 skips the dump (timing runs).  window > 1:
// every worker keeps that many groups enqueued (Enqueue/Collect) instead of blocking per group.
// n_devices > 1: a GenotypeDispatcher over devices device .. device+n_devices-1; counters[9 + d] =
// payloads that went to device d (counters must hold 17 entries).
static int BatcherRun(int device, const lgr_batch_in* in, const char* names, const char* samples,
                             const int* sample_id, const long long* start0, const long long* isize,
                             const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
                             int n_threads, int rounds, int window, int n_devices, unsigned long long* counters, int n_counters,
                             char* out, long long cap, bool dump) {
  try {
    JobSet js;
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    std::vector<lancet_gpu::Result> res(js.jobs.size());
    std::vector<std::string> errors((size_t)n_threads);
    {
      std::vector<int> devices;
      for (int d = 0; d < (n_devices > 1 ? n_devices : 1); ++d) devices.push_back(device + d);
      lancet_gpu::GenotypeDispatcher batcher(devices, X31OfView);
      std::vector<std::thread> workers;
      const auto t0 = std::chrono::steady_clock::now();
      for (int t = 0; t < n_threads; ++t) {
        workers.emplace_back([&, t] {
          try {
            // split ProcessWindow: up to `window` groups enqueued per worker before the oldest is collected.  The
            // window rolls across rounds (a worker of the caller has thousands of windows queued, it never drains
            // its pipeline between two of them); everything is collected before the clock stops.
            std::deque<std::pair<std::size_t, lancet_gpu::GenotypeDispatcher::Ticket>> open_t;
            for (int round = 0; round < rounds; ++round) {
              if (window <= 1) {  // the reference's call shape: one blocking call per group
                for (std::size_t g = (size_t)t; g < js.jobs.size(); g += (size_t)n_threads) {
                  const lancet_gpu::GenotypeJob& j = js.jobs[g];
                  res[g] = batcher.Genotype(j.haps, j.n_haps, j.reads, j.n_reads, j.variants, j.n_variants);
                }
              } else {
                for (std::size_t g = (size_t)t; g < js.jobs.size(); g += (size_t)n_threads) {
                  if ((int)open_t.size() == window) {
                    res[open_t.front().first] = batcher.Collect(open_t.front().second);
                    open_t.pop_front();
                  }
                  open_t.emplace_back(g, batcher.Enqueue(js.jobs[g]));
                }
              }
            }
            for (auto& ot : open_t) res[ot.first] = batcher.Collect(ot.second);
          } catch (const std::exception& e) {
            errors[(size_t)t] = e.what();
          }
        });
      }
      for (auto& w : workers) w.join();
      const auto t1 = std::chrono::steady_clock::now();
      lancet_gpu::GenotypeBatcher::Counters c;  // summed over the devices; counters[9 + d] = jobs of device d
      for (std::size_t d = 0; d < batcher.Devices(); ++d) {
        const auto cd = batcher.Stats(d);
        c.batches += cd.batches, c.jobs += cd.jobs, c.pairs += cd.pairs;
        c.max_jobs_in_batch = std::max(c.max_jobs_in_batch, cd.max_jobs_in_batch);
        c.ns_pack += cd.ns_pack, c.ns_submit += cd.ns_submit, c.ns_wait += cd.ns_wait, c.ns_deliver += cd.ns_deliver;
        c.h2d_bytes += cd.h2d_bytes, c.d2h_bytes += cd.d2h_bytes, c.retried_alone += cd.retried_alone, c.ns_seal_wait += cd.ns_seal_wait;
        if (counters && d < 8) counters[9 + d] = cd.jobs;
      }
      if (counters && n_counters >= 20) counters[17] = c.h2d_bytes, counters[18] = c.d2h_bytes, counters[19] = c.retried_alone;
      if (counters && n_counters >= 21) counters[20] = c.ns_seal_wait;
      if (counters) {
        counters[0] = c.batches, counters[1] = c.jobs, counters[2] = c.pairs, counters[3] = c.max_jobs_in_batch;
        counters[4] = (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
        counters[5] = c.ns_pack, counters[6] = c.ns_submit, counters[7] = c.ns_wait, counters[8] = c.ns_deliver;
      }
    }
    for (const auto& e : errors)
      if (!e.empty()) throw std::runtime_error(e);
    if (!dump || cap <= 0) return 0;
    return WriteOut(DumpResults(in, js, res), out, cap);
  } catch (const std::exception& e) {
    if (out && cap > 0) std::snprintf(out, (size_t)cap, "EXCEPTION: %s", e.what());
    return -2;
  }
}

int lgr_adapter_batcher_dump(int device, const lgr_batch_in* in, const char* names, const char* samples,
                             const int* sample_id, const long long* start0, const long long* isize,
                             const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
                             int n_threads, int rounds, int window, int n_devices, unsigned long long* counters, char* out,
                             long long cap) {
  return BatcherRun(device, in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, n_threads, rounds, window,
                    n_devices, counters, 17, out, cap, true);
}

// the same without the dump, for bench.py's e2e arm: counters[21] adds [17] bytes copied host→device, [18] device→host,
// [19] payloads re-run alone, [20] batcher-thread ns spent waiting for workers still packing when a slab was sealed
int lgr_adapter_batcher_bench(int device, const lgr_batch_in* in, const char* names, const char* samples,
                              const int* sample_id, const long long* start0, const long long* isize,
                              const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
                              int n_threads, int rounds, int window, int n_devices, unsigned long long* counters, char* err,
                              long long err_cap) {
  const int rc = BatcherRun(device, in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, n_threads, rounds,
                            window, n_devices, counters, 21, err, err_cap, false);
  return rc;
}

// Failure isolation of the batcher: ONE worker enqueues every group before it collects the first (so they
// share device batches), catching per payload.  job_status[g] = 0 ok, 1 Enqueue threw, 2 Collect threw; the dump
// holds the evidence of the groups that succeeded.  mid_occ > 0 overrides lgr_params::mid_occ (to provoke a
// device-side cap).  counters[0] = payloads re-run alone.
int lgr_adapter_isolation_dump(int device, const lgr_batch_in* in, const char* names, const char* samples,
                               const int* sample_id, const long long* start0, const long long* isize,
                               const unsigned short* sam_flag, const unsigned char* mapq, const unsigned char* softclip,
                               int mid_occ, int* job_status, unsigned long long* counters, char* out, long long cap) {
  try {
    JobSet js;
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    std::vector<lancet_gpu::Result> res(js.jobs.size());
    lgr_params prm;
    lgr_default_params(&prm);
    if (mid_occ > 0) prm.mid_occ = mid_occ;
    lancet_gpu::GenotypeBatcher::Options opt;
    opt.device = device, opt.params = &prm, opt.linger_us = 50000;  // all payloads of the probe share one device batch
    {
      lancet_gpu::GenotypeBatcher batcher(opt, X31OfView);
      std::vector<lancet_gpu::GenotypeBatcher::Ticket> tickets(js.jobs.size());
      for (std::size_t g = 0; g < js.jobs.size(); ++g) {
        job_status[g] = 0;
        try {
          tickets[g] = batcher.Enqueue(js.jobs[g]);
        } catch (const std::exception&) {
          job_status[g] = 1;
        }
      }
      for (std::size_t g = 0; g < js.jobs.size(); ++g) {
        if (job_status[g]) continue;
        try {
          res[g] = batcher.Collect(tickets[g]);
        } catch (const std::exception&) {
          job_status[g] = 2;
        }
      }
      if (counters) counters[0] = batcher.Stats().retried_alone, counters[1] = batcher.Stats().batches;
    }
    return WriteOut(DumpResults(in, js, res), out, cap);
  } catch (const std::exception& e) {
    if (out && cap > 0) std::snprintf(out, (size_t)cap, "EXCEPTION: %s", e.what());
    return -2;
  }
}

// SURVEY.md §8f #2 — EvidenceColumns::AppendJob over every group of the batch on caller-provided
// lgr_assign records (no device involved).  Returns the columns as a lgr_evidence_in that stays
// valid until the next call on this thread; key_out[3 s + 0..2] = group, variant (index within
// the group) and sample id of support s (at most key_cap supports are described).
const lgr_evidence_in* lgr_adapter_evidence_columns(const lgr_batch_in* in, const char* names, const char* samples,
                                                    const int* sample_id, const long long* start0, const long long* isize,
                                                    const unsigned short* sam_flag, const unsigned char* mapq,
                                                    const unsigned char* softclip, const lgr_assign* assign, int* key_out,
                                                    int key_cap) {
  static thread_local lancet_gpu::EvidenceColumns cols;
  static thread_local JobSet js;
  try {
    js = JobSet();
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    cols.Clear();
    long long off = 0;
    std::vector<int> job_of_support;
    for (std::size_t g = 0; g < js.jobs.size(); ++g) {
      cols.AppendJob(js.jobs[g], assign + off, X31OfView);
      off += (long long)(js.jobs[g].n_reads * js.jobs[g].n_variants);
      job_of_support.resize(cols.NumSupports(), (int)g);
    }
    for (std::size_t s = 0; s < cols.NumSupports() && (int)s < key_cap; ++s) {
      const auto& k = cols.Keys()[s];
      const int v = (int)(static_cast<const lancet_gpu::VariantIn*>(k.variant) - js.vars.data());
      int sid = -1;
      for (std::size_t i = 0; i < js.sn.size(); ++i)
        if (js.sn[i] == k.sample) sid = (int)i;
      key_out[3 * s] = job_of_support[s], key_out[3 * s + 1] = v - in->grp_var_begin[job_of_support[s]], key_out[3 * s + 2] = sid;
    }
    return &cols.In();
  } catch (const std::exception&) {
    return nullptr;
  }
}

// The same columns through lancet_gpu::GpuFormatMetrics::Compute on `device`: out must hold one
// lgr_format per support (n_out of them); returns the number of supports, < 0 on error (message in err).
int lgr_adapter_format_metrics(int device, const lgr_batch_in* in, const char* names, const char* samples, const int* sample_id,
                               const long long* start0, const long long* isize, const unsigned short* sam_flag,
                               const unsigned char* mapq, const unsigned char* softclip, const lgr_assign* assign,
                               lgr_format* out, int n_out, char* err, int err_cap) {
  try {
    JobSet js;
    BuildJobs(in, names, samples, sample_id, start0, isize, sam_flag, mapq, softclip, js);
    lancet_gpu::EvidenceColumns cols;
    long long off = 0;
    for (std::size_t g = 0; g < js.jobs.size(); ++g) {
      cols.AppendJob(js.jobs[g], assign + off, X31OfView);
      off += (long long)(js.jobs[g].n_reads * js.jobs[g].n_variants);
    }
    if ((int)cols.NumSupports() > n_out) throw std::runtime_error("output too small");
    lancet_gpu::GpuFormatMetrics fm(device);
    const std::vector<lgr_format> res = fm.Compute(cols);
    std::memcpy(out, res.data(), res.size() * sizeof(lgr_format));
    return (int)res.size();
  } catch (const std::exception& e) {
    std::snprintf(err, (size_t)err_cap, "%s", e.what());
    return -2;
  }
}

}  // extern "C"
