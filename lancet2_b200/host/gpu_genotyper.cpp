// gpu_genotyper.cpp — see gpu_genotyper.h.  Host logic only: batch building, the C-ABI call,
// and AddToTable/AddEvidence in the reference's read order.
#include "gpu_genotyper.h"

#include <cstring>
#include <stdexcept>

namespace lancet_gpu {

// variant_support.cpp:23-67 (EnsureAlleleSlot + first-seen-wins dedup on the name hash)
void VariantSupport::AddEvidence(ReadEvidence const& ev) {
  if (mAlleleData.size() <= ev.mAllele) mAlleleData.resize(static_cast<std::size_t>(ev.mAllele) + 1);
  PerAlleleData& data = mAlleleData[ev.mAllele];
  if (!data.mNameHashes.try_emplace(ev.mRnameHash, ev.mStrand).second) return;
  if (ev.mStrand == Strand::FWD) data.mFwdBaseQuals.push_back(ev.mBaseQual);
  else data.mRevBaseQuals.push_back(ev.mBaseQual);
  data.mMapQuals.push_back(ev.mMapQual);
  data.mAlnScores.push_back(ev.mAlnScore);
  if (ev.mIsSoftClipped) ++data.mSoftClipCount;
  if (ev.mIsProperPair && ev.mInsertSize != 0) data.mProperPairIsizes.push_back(static_cast<double>(ev.mInsertSize));
  data.mFoldedReadPositions.push_back(ev.mFoldedReadPos);
  data.mRefNmValues.push_back(static_cast<double>(ev.mRefNm));
  data.mAlignmentStarts.push_back(ev.mAlignmentStart);
  data.mOwnHapNmValues.push_back(static_cast<double>(ev.mOwnHapNm));
  data.mHaplotypeIds.push_back(ev.mAssignedHaplotypeId);
}

// support_array.cpp:19-28
VariantSupport& SupportArray::FindOrCreate(std::string_view sample_name) {
  for (auto& item : mItems)
    if (item.mSampleName == sample_name) return *item.mData;
  mItems.push_back(NamedSupport{sample_name, std::make_unique<VariantSupport>()});
  return *mItems.back().mData;
}

static void Check(lgr_ctx* ctx, int rc) {
  if (rc != LGR_OK) {
    std::string msg = lgr_strerror(rc);
    const char* detail = lgr_last_error(ctx);
    if (detail && *detail) msg += std::string(": ") + detail;
    throw std::runtime_error("lancet_gpu::GpuGenotyper: " + msg);
  }
}

GpuGenotyper::GpuGenotyper(int device_ordinal, const lgr_params* params) {
  if (params) mParams = *params;
  else lgr_default_params(&mParams);
  Check(nullptr, lgr_create(device_ordinal, &mParams, &mCtx));
}

GpuGenotyper::~GpuGenotyper() { lgr_destroy(mCtx); }

Result GpuGenotyper::Genotype(const std::string* haps, std::size_t n_haps, const ReadIn* reads, std::size_t n_reads,
                              const VariantIn* variants, std::size_t n_variants, const NameHashFn& name_hash) {
  std::vector<GenotypeJob> jobs{GenotypeJob{haps, n_haps, reads, n_reads, variants, n_variants}};
  return std::move(GenotypeMany(jobs, name_hash)[0]);
}

std::vector<Result> GpuGenotyper::GenotypeMany(const std::vector<GenotypeJob>& jobs, const NameHashFn& name_hash) {
  const int G = static_cast<int>(jobs.size());
  std::vector<Result> results(jobs.size());
  if (G == 0) return results;
  // mm_mapopt_update latches mid_occ from the first index this Genotyper ever builds
  // (genotyper.cpp:263-266); later haplotypes never refresh it.
  if (mParams.mid_occ <= 0 && mLatchedMidOcc <= 0) {
    for (const GenotypeJob& j : jobs) {
      if (j.n_haps == 0) continue;
      Check(mCtx, lgr_hap_mid_occ(mCtx, reinterpret_cast<const std::uint8_t*>(j.haps[0].data()),
                                  static_cast<std::int32_t>(j.haps[0].size()), &mLatchedMidOcc));
      break;
    }
  }
  // ---- SoA batch ----
  std::vector<std::int32_t> ghb(G + 1, 0), grb(G + 1, 0), gvb(G + 1, 0), gmid(G, mLatchedMidOcc);
  std::vector<std::int64_t> hap_off{0}, read_off{0}, var_hap_off{0};
  std::vector<std::uint8_t> hap_bases, read_bases, read_quals;
  std::vector<std::uint32_t> x31;
  std::vector<std::int32_t> var_start, var_len;
  std::vector<std::int8_t> var_allele;
  for (int g = 0; g < G; ++g) {
    const GenotypeJob& j = jobs[g];
    for (std::size_t h = 0; h < j.n_haps; ++h) {
      hap_bases.insert(hap_bases.end(), j.haps[h].begin(), j.haps[h].end());
      hap_off.push_back(static_cast<std::int64_t>(hap_bases.size()));
    }
    for (std::size_t r = 0; r < j.n_reads; ++r) {
      const ReadIn& rd = j.reads[r];
      read_bases.insert(read_bases.end(), rd.seq.begin(), rd.seq.end());
      read_quals.insert(read_quals.end(), rd.qual, rd.qual + rd.seq.size());
      read_off.push_back(static_cast<std::int64_t>(read_bases.size()));
      const std::string qn(rd.qname);  // mm_map receives the NUL-terminated QnamePtr()
      x31.push_back(lgr_x31_hash(qn.c_str()));
    }
    for (std::size_t v = 0; v < j.n_variants; ++v) {
      const VariantIn& var = j.variants[v];
      for (std::size_t h = 0; h < j.n_haps; ++h) {  // ExtractHapBounds (genotyper.cpp:329-352)
        std::int32_t st = -1, ln = 0;
        std::int8_t al = -1;
        if (h == 0) {
          st = static_cast<std::int32_t>(var.local_ref_start0), ln = static_cast<std::int32_t>(var.ref_allele_len), al = 0;
        } else {
          for (std::size_t a = 0; a < var.alts.size() && al < 0; ++a)
            for (const auto& kv : var.alts[a].hap_start0)
              if (kv.first == h) {
                st = static_cast<std::int32_t>(kv.second), ln = static_cast<std::int32_t>(var.alts[a].seq_len);
                al = static_cast<std::int8_t>(a + 1);
                break;
              }
        }
        var_start.push_back(st), var_len.push_back(ln), var_allele.push_back(al);
      }
      var_hap_off.push_back(static_cast<std::int64_t>(var_start.size()));
    }
    ghb[g + 1] = ghb[g] + static_cast<std::int32_t>(j.n_haps);
    grb[g + 1] = grb[g] + static_cast<std::int32_t>(j.n_reads);
    gvb[g + 1] = gvb[g] + static_cast<std::int32_t>(j.n_variants);
  }
  // keep pointers valid for empty vectors
  hap_bases.push_back(0), read_bases.push_back(0), read_quals.push_back(0), x31.push_back(0);
  var_start.push_back(0), var_len.push_back(0), var_allele.push_back(0);
  lgr_batch_in in;
  std::memset(&in, 0, sizeof(in));
  in.n_groups = G, in.n_haps = ghb[G], in.n_reads = grb[G], in.n_vars = gvb[G];
  in.grp_hap_begin = ghb.data(), in.grp_read_begin = grb.data(), in.grp_var_begin = gvb.data();
  in.hap_off = hap_off.data(), in.hap_bases = hap_bases.data();
  in.read_off = read_off.data(), in.read_bases = read_bases.data(), in.read_quals = read_quals.data();
  in.read_name_hash = x31.data();
  in.var_hap_off = var_hap_off.data(), in.var_start = var_start.data(), in.var_len = var_len.data();
  in.var_allele = var_allele.data();
  in.grp_mid_occ = mLatchedMidOcc > 0 ? gmid.data() : nullptr;
  std::vector<std::int64_t> pair_off(in.n_reads + 1), asg_off(in.n_reads + 1);
  Check(mCtx, lgr_pair_offsets(&in, pair_off.data(), asg_off.data()));
  std::vector<lgr_assign> assign(static_cast<std::size_t>(asg_off[in.n_reads]) + 1);
  lgr_batch_out out;
  std::memset(&out, 0, sizeof(out));
  out.n_assign = asg_off[in.n_reads];
  out.assign = assign.data();  // out.aln stays NULL: the adapter only needs the assignments
  Check(mCtx, lgr_genotype_batch(mCtx, &in, &out, &mStats));
  // ---- AddToTable (genotyper.cpp:423-456), reads in the caller's order ----
  for (int g = 0; g < G; ++g) {
    const GenotypeJob& j = jobs[g];
    Result& table = results[g];
    for (std::size_t r = 0; r < j.n_reads; ++r) {
      const ReadIn& rd = j.reads[r];
      const std::int64_t base = asg_off[grb[g] + static_cast<std::int64_t>(r)];
      std::uint32_t rname_hash = 0;
      bool hashed = false;
      for (std::size_t v = 0; v < j.n_variants; ++v) {
        const lgr_assign& a = assign[static_cast<std::size_t>(base) + v];
        if (!a.assigned) continue;
        if (!hashed) rname_hash = name_hash(rd.qname), hashed = true;
        VariantSupport& support = table[j.variants[v].key].FindOrCreate(rd.sample_name);
        ReadEvidence ev;
        ev.mInsertSize = rd.insert_size;
        ev.mAlignmentStart = rd.start0;
        ev.mAlnScore = static_cast<double>(a.global_score) + (a.local_score * a.local_identity);  // CombinedScore()
        ev.mFoldedReadPos = a.folded_read_pos;
        ev.mRnameHash = rname_hash;
        ev.mRefNm = a.ref_nm, ev.mOwnHapNm = a.own_hap_nm, ev.mAssignedHaplotypeId = a.hap_id;
        ev.mAllele = static_cast<AlleleIndex>(a.allele);
        ev.mStrand = (rd.sam_flag & 0x10) ? Strand::REV : Strand::FWD;
        ev.mBaseQual = a.base_qual, ev.mMapQual = rd.map_qual;
        ev.mIsSoftClipped = rd.is_soft_clipped;
        ev.mIsProperPair = (rd.sam_flag & 0x2) != 0;
        support.AddEvidence(ev);
      }
    }
  }
  return results;
}

}  // namespace lancet_gpu
