// gpu_genotyper.cpp — see gpu_genotyper.h.  Host logic only: batch building, the C-ABI call,
// and AddToTable/AddEvidence in the reference's read order.
#include "gpu_genotyper.h"

#include <chrono>
#include <cstdlib>
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

// ---------------------------------------------------------------------------------------------
// PackedBatch
// ---------------------------------------------------------------------------------------------
template <typename T>
void PackedBatch::Free(Buf<T>& b) {
  if (b.p) {
    if (b.pinned) lgr_free_pinned(b.p);
    else std::free(b.p);
  }
  b.p = nullptr, b.n = b.cap = 0;
}

template <typename T>
void PackedBatch::Reserve(Buf<T>& b, std::size_t n) {
  if (n <= b.cap) return;
  std::size_t cap = b.cap ? b.cap : 256;
  while (cap < n) cap = cap * 2 + 64;  // pinned allocations are slow: grow rarely
  bool pinned = true;
  T* p = static_cast<T*>(lgr_alloc_pinned(cap * sizeof(T)));
  if (!p) {
    pinned = false;
    p = static_cast<T*>(std::malloc(cap * sizeof(T)));
    if (!p) throw std::bad_alloc();
  }
  if (b.n) std::memcpy(p, b.p, b.n * sizeof(T));
  const std::size_t keep = b.n;
  Free(b);
  b.p = p, b.n = keep, b.cap = cap, b.pinned = pinned;
}

template <typename T>
void PackedBatch::Push(Buf<T>& b, T v) {
  if (b.n == b.cap) Reserve(b, b.n + 1);
  b.p[b.n++] = v;
}

template <typename T>
void PackedBatch::Append(Buf<T>& b, const T* src, std::size_t n) {
  Reserve(b, b.n + n + 1);
  if (n) std::memcpy(b.p + b.n, src, n * sizeof(T));
  b.n += n;
}

PackedBatch::~PackedBatch() {
  Free(mGhb), Free(mGrb), Free(mGvb), Free(mGmid), Free(mVarStart), Free(mVarLen);
  Free(mHapOff), Free(mReadOff), Free(mVarHapOff);
  Free(mHapBases), Free(mReadBases), Free(mReadQuals);
  Free(mX31), Free(mVarAllele), Free(mAssign);
}

void PackedJob::Build(const GenotypeJob& j) {
  n_haps = j.n_haps, n_reads = j.n_reads, n_variants = j.n_variants;
  hap_bases.clear(), read_bases.clear(), read_quals.clear(), x31.clear();
  var_start.clear(), var_len.clear(), var_allele.clear();
  hap_off.assign(1, 0), read_off.assign(1, 0), var_hap_off.assign(1, 0);
  std::size_t hap_total = 0, read_total = 0;
  for (std::size_t h = 0; h < j.n_haps; ++h) hap_total += j.haps[h].size();
  for (std::size_t r = 0; r < j.n_reads; ++r) read_total += j.reads[r].seq.size();
  hap_bases.reserve(hap_total), read_bases.reserve(read_total), read_quals.reserve(read_total);
  x31.reserve(j.n_reads), read_off.reserve(j.n_reads + 1), hap_off.reserve(j.n_haps + 1);
  for (std::size_t h = 0; h < j.n_haps; ++h) {
    hap_bases.insert(hap_bases.end(), j.haps[h].begin(), j.haps[h].end());
    hap_off.push_back(static_cast<std::int64_t>(hap_bases.size()));
  }
  std::string qn;
  for (std::size_t r = 0; r < j.n_reads; ++r) {
    const ReadIn& rd = j.reads[r];
    read_bases.insert(read_bases.end(), rd.seq.begin(), rd.seq.end());
    read_quals.insert(read_quals.end(), rd.qual, rd.qual + rd.seq.size());
    read_off.push_back(static_cast<std::int64_t>(read_bases.size()));
    qn.assign(rd.qname);  // mm_map receives the NUL-terminated QnamePtr()
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
}

void PackedBatch::Pack(const GenotypeJob* jobs, std::size_t n_jobs, std::int32_t latched_mid_occ) {
  std::vector<PackedJob> packed(n_jobs);
  std::vector<const PackedJob*> ptrs(n_jobs);
  for (std::size_t g = 0; g < n_jobs; ++g) packed[g].Build(jobs[g]), ptrs[g] = &packed[g];
  PackPrepared(ptrs.data(), n_jobs, latched_mid_occ);
}

void PackedBatch::PackPrepared(const PackedJob* const* jobs, std::size_t n_jobs, std::int32_t latched_mid_occ) {
  for (auto* b : {&mGhb, &mGrb, &mGvb, &mGmid, &mVarStart, &mVarLen}) b->n = 0;
  for (auto* b : {&mHapOff, &mReadOff, &mVarHapOff}) b->n = 0;
  for (auto* b : {&mHapBases, &mReadBases, &mReadQuals}) b->n = 0;
  mX31.n = 0, mVarAllele.n = 0;
  mJobAsg.clear();
  // sizes first: one reservation per array, then straight block copies
  std::size_t nh = 0, nr = 0, nv = 0, hb = 0, rb = 0, vh = 0;
  for (std::size_t g = 0; g < n_jobs; ++g) {
    const PackedJob& j = *jobs[g];
    nh += j.n_haps, nr += j.n_reads, nv += j.n_variants;
    hb += j.hap_bases.size(), rb += j.read_bases.size(), vh += j.var_start.size();
  }
  Reserve(mGhb, n_jobs + 2), Reserve(mGrb, n_jobs + 2), Reserve(mGvb, n_jobs + 2), Reserve(mGmid, n_jobs + 1);
  Reserve(mHapOff, nh + 2), Reserve(mReadOff, nr + 2), Reserve(mVarHapOff, nv + 2);
  Reserve(mHapBases, hb + 1), Reserve(mReadBases, rb + 1), Reserve(mReadQuals, rb + 1);
  Reserve(mX31, nr + 1), Reserve(mVarStart, vh + 1), Reserve(mVarLen, vh + 1), Reserve(mVarAllele, vh + 1);
  Push<std::int32_t>(mGhb, 0), Push<std::int32_t>(mGrb, 0), Push<std::int32_t>(mGvb, 0);
  Push<std::int64_t>(mHapOff, 0), Push<std::int64_t>(mReadOff, 0), Push<std::int64_t>(mVarHapOff, 0);
  std::int64_t n_assign = 0;
  mPairs = 0;
  auto rebase = [](Buf<std::int64_t>& dst, const std::vector<std::int64_t>& rel, std::int64_t base) {
    for (std::size_t i = 1; i < rel.size(); ++i) dst.p[dst.n++] = rel[i] + base;
  };
  for (std::size_t g = 0; g < n_jobs; ++g) {
    const PackedJob& j = *jobs[g];
    rebase(mHapOff, j.hap_off, static_cast<std::int64_t>(mHapBases.n));
    rebase(mReadOff, j.read_off, static_cast<std::int64_t>(mReadBases.n));
    rebase(mVarHapOff, j.var_hap_off, static_cast<std::int64_t>(mVarStart.n));
    Append(mHapBases, j.hap_bases.data(), j.hap_bases.size());
    Append(mReadBases, j.read_bases.data(), j.read_bases.size());
    Append(mReadQuals, j.read_quals.data(), j.read_quals.size());
    Append(mX31, j.x31.data(), j.x31.size());
    Append(mVarStart, j.var_start.data(), j.var_start.size());
    Append(mVarLen, j.var_len.data(), j.var_len.size());
    Append(mVarAllele, j.var_allele.data(), j.var_allele.size());
    Push(mGhb, mGhb.p[g] + static_cast<std::int32_t>(j.n_haps));
    Push(mGrb, mGrb.p[g] + static_cast<std::int32_t>(j.n_reads));
    Push(mGvb, mGvb.p[g] + static_cast<std::int32_t>(j.n_variants));
    Push(mGmid, latched_mid_occ);
    mJobAsg.push_back(n_assign);
    n_assign += static_cast<std::int64_t>(j.n_reads) * static_cast<std::int64_t>(j.n_variants);
    mPairs += static_cast<std::int64_t>(j.n_reads) * static_cast<std::int64_t>(j.n_haps);
  }
  Reserve(mAssign, static_cast<std::size_t>(n_assign) + 1);
  mAssign.n = static_cast<std::size_t>(n_assign);
  const int G = static_cast<int>(n_jobs);
  std::memset(&mIn, 0, sizeof(mIn));
  mIn.n_groups = G, mIn.n_haps = mGhb.p[G], mIn.n_reads = mGrb.p[G], mIn.n_vars = mGvb.p[G];
  mIn.grp_hap_begin = mGhb.p, mIn.grp_read_begin = mGrb.p, mIn.grp_var_begin = mGvb.p;
  mIn.hap_off = mHapOff.p, mIn.hap_bases = mHapBases.p;
  mIn.read_off = mReadOff.p, mIn.read_bases = mReadBases.p, mIn.read_quals = mReadQuals.p;
  mIn.read_name_hash = mX31.p;
  mIn.var_hap_off = mVarHapOff.p, mIn.var_start = mVarStart.p, mIn.var_len = mVarLen.p, mIn.var_allele = mVarAllele.p;
  mIn.grp_mid_occ = latched_mid_occ > 0 ? mGmid.p : nullptr;
  std::memset(&mOut, 0, sizeof(mOut));
  mOut.n_assign = n_assign;
  mOut.assign = mAssign.p;  // out.aln stays NULL: the adapter only needs the assignments
}

Result PackedBatch::BuildResult(const GenotypeJob& j, const lgr_assign* assign, const NameHashFn& name_hash) {
  Result table;
  for (std::size_t r = 0; r < j.n_reads; ++r) {
    const ReadIn& rd = j.reads[r];
    std::uint32_t rname_hash = 0;
    bool hashed = false;
    for (std::size_t v = 0; v < j.n_variants; ++v) {
      const lgr_assign& a = assign[r * j.n_variants + v];
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
  return table;
}

// ---------------------------------------------------------------------------------------------
// EvidenceColumns / GpuFormatMetrics
// ---------------------------------------------------------------------------------------------
void EvidenceColumns::Clear() {
  mKeys.clear();
  mSupBegin.assign(1, 0);
  for (auto* v : {&mInsertSize, &mAlnStart}) v->clear();
  for (auto* v : {&mNumAlleles, &mVariantLen, &mTotalHaps}) v->clear();
  for (auto* v : {&mAlnScore, &mFoldedPos}) v->clear();
  for (auto* v : {&mRnameHash, &mRefNm, &mOwnHapNm, &mHapId}) v->clear();
  for (auto* v : {&mAllele, &mFlags, &mBaseQual, &mMapQual}) v->clear();
}

void EvidenceColumns::AppendJob(const GenotypeJob& j, const lgr_assign* assign, const NameHashFn& name_hash) {
  // name hashes once per read that has any assignment (AddToTable hashes once per read, genotyper.cpp:426)
  std::vector<std::uint32_t> hash(j.n_reads, 0);
  for (std::size_t r = 0; r < j.n_reads; ++r)
    for (std::size_t v = 0; v < j.n_variants; ++v)
      if (assign[r * j.n_variants + v].assigned) {
        hash[r] = name_hash(j.reads[r].qname);
        break;
      }
  std::vector<std::string_view> samples;  // FindOrCreate's creation order for this variant
  for (std::size_t v = 0; v < j.n_variants; ++v) {
    const VariantIn& var = j.variants[v];
    samples.clear();
    for (std::size_t r = 0; r < j.n_reads; ++r) {
      if (!assign[r * j.n_variants + v].assigned) continue;
      const std::string_view s = j.reads[r].sample_name;
      bool seen = false;
      for (const auto& x : samples) seen = seen || x == s;
      if (!seen) samples.push_back(s);
    }
    std::int64_t max_len = 0;  // variant_call.cpp:179-182: max |mLength| over the ALTs
    for (const auto& alt : var.alts) max_len = std::max<std::int64_t>(max_len, alt.length < 0 ? -alt.length : alt.length);
    for (const auto& sample : samples) {
      for (std::size_t r = 0; r < j.n_reads; ++r) {
        const lgr_assign& a = assign[r * j.n_variants + v];
        const ReadIn& rd = j.reads[r];
        if (!a.assigned || rd.sample_name != sample) continue;
        mInsertSize.push_back(rd.insert_size);
        mAlnStart.push_back(rd.start0);
        mAlnScore.push_back(static_cast<double>(a.global_score) + (a.local_score * a.local_identity));  // CombinedScore()
        mFoldedPos.push_back(a.folded_read_pos);
        mRnameHash.push_back(hash[r]);
        mRefNm.push_back(a.ref_nm), mOwnHapNm.push_back(a.own_hap_nm), mHapId.push_back(a.hap_id);
        mAllele.push_back(static_cast<std::uint8_t>(a.allele));
        mFlags.push_back(static_cast<std::uint8_t>(((rd.sam_flag & 0x10) ? LGR_EV_REV : 0u) | (rd.is_soft_clipped ? LGR_EV_SOFTCLIP : 0u) |
                                                   ((rd.sam_flag & 0x2) ? LGR_EV_PROPER_PAIR : 0u)));
        mBaseQual.push_back(a.base_qual), mMapQual.push_back(rd.map_qual);
      }
      mKeys.push_back(SupportKey{var.key, sample});
      mSupBegin.push_back(static_cast<std::int64_t>(mAllele.size()));
      mNumAlleles.push_back(static_cast<std::int32_t>(1 + var.alts.size()));
      mVariantLen.push_back(static_cast<std::int32_t>(max_len));
      mTotalHaps.push_back(static_cast<std::int32_t>(j.n_haps));
    }
  }
}

const lgr_evidence_in& EvidenceColumns::In() {
  mIn = lgr_evidence_in{};
  mIn.n_supports = static_cast<std::int32_t>(mKeys.size());
  mIn.n_evidence = static_cast<std::int64_t>(mAllele.size());
  mIn.sup_begin = mSupBegin.data(), mIn.sup_n_alleles = mNumAlleles.data();
  mIn.sup_variant_len = mVariantLen.data(), mIn.sup_total_haps = mTotalHaps.data();
  mIn.insert_size = mInsertSize.data(), mIn.aln_start = mAlnStart.data();
  mIn.aln_score = mAlnScore.data(), mIn.folded_pos = mFoldedPos.data();
  mIn.rname_hash = mRnameHash.data(), mIn.ref_nm = mRefNm.data(), mIn.own_hap_nm = mOwnHapNm.data(), mIn.hap_id = mHapId.data();
  mIn.allele = mAllele.data(), mIn.flags = mFlags.data(), mIn.base_qual = mBaseQual.data(), mIn.map_qual = mMapQual.data();
  return mIn;
}

GpuFormatMetrics::GpuFormatMetrics(int device_ordinal) {
  const int rc = lgr_format_create(device_ordinal, &mCtx);
  if (rc != LGR_OK)
    throw std::runtime_error(std::string("lancet_gpu::GpuFormatMetrics: ") + lgr_strerror(rc) + ": " + lgr_format_last_error(nullptr));
}

GpuFormatMetrics::~GpuFormatMetrics() { lgr_format_destroy(mCtx); }

std::vector<lgr_format> GpuFormatMetrics::Compute(EvidenceColumns& columns, float* ms_kernels) {
  std::vector<lgr_format> out(columns.NumSupports());
  const int rc = lgr_format_metrics(mCtx, &columns.In(), out.data(), ms_kernels);
  if (rc != LGR_OK)
    throw std::runtime_error(std::string("lancet_gpu::GpuFormatMetrics: ") + lgr_strerror(rc) + ": " + lgr_format_last_error(mCtx));
  return out;
}

// ---------------------------------------------------------------------------------------------
// GpuGenotyper
// ---------------------------------------------------------------------------------------------
GpuGenotyper::GpuGenotyper(int device_ordinal, const lgr_params* params) {
  if (params) mParams = *params;
  else lgr_default_params(&mParams);
  Check(nullptr, lgr_create(device_ordinal, &mParams, &mCtx));
  mBatch = std::make_unique<PackedBatch>();
}

GpuGenotyper::~GpuGenotyper() {
  mBatch.reset();
  lgr_destroy(mCtx);
}

Result GpuGenotyper::Genotype(const std::string* haps, std::size_t n_haps, const ReadIn* reads, std::size_t n_reads,
                              const VariantIn* variants, std::size_t n_variants, const NameHashFn& name_hash) {
  std::vector<GenotypeJob> jobs{GenotypeJob{haps, n_haps, reads, n_reads, variants, n_variants}};
  return std::move(GenotypeMany(jobs, name_hash)[0]);
}

// mm_mapopt_update latches mid_occ from the first index a Genotyper ever builds
// (genotyper.cpp:263-266); later haplotypes never refresh it.
static void LatchMidOcc(lgr_ctx* ctx, const lgr_params& prm, const GenotypeJob* jobs, std::size_t n, std::int32_t* latched) {
  if (prm.mid_occ > 0 || *latched > 0) return;
  for (std::size_t i = 0; i < n; ++i) {
    if (jobs[i].n_haps == 0) continue;
    Check(ctx, lgr_hap_mid_occ(ctx, reinterpret_cast<const std::uint8_t*>(jobs[i].haps[0].data()),
                               static_cast<std::int32_t>(jobs[i].haps[0].size()), latched));
    return;
  }
}

std::vector<Result> GpuGenotyper::GenotypeMany(const std::vector<GenotypeJob>& jobs, const NameHashFn& name_hash) {
  std::vector<Result> results(jobs.size());
  if (jobs.empty()) return results;
  LatchMidOcc(mCtx, mParams, jobs.data(), jobs.size(), &mLatchedMidOcc);
  mBatch->Pack(jobs.data(), jobs.size(), mLatchedMidOcc);
  Check(mCtx, lgr_genotype_batch(mCtx, &mBatch->In(), &mBatch->Out(), &mStats));
  for (std::size_t g = 0; g < jobs.size(); ++g) results[g] = PackedBatch::BuildResult(jobs[g], mBatch->JobAssign(g), name_hash);
  return results;
}

// ---------------------------------------------------------------------------------------------
// GenotypeBatcher
// ---------------------------------------------------------------------------------------------
GenotypeBatcher::GenotypeBatcher(const Options& opt, NameHashFn name_hash) : mOpt(opt), mNameHash(std::move(name_hash)) {
  if (mOpt.depth < 1) mOpt.depth = 1;
  if (mOpt.depth > LGR_MAX_INFLIGHT) mOpt.depth = LGR_MAX_INFLIGHT;
  if (opt.params) mParams = *opt.params;
  else lgr_default_params(&mParams);
  mOpt.params = nullptr;
  Check(nullptr, lgr_create(mOpt.device, &mParams, &mCtx));
  for (int i = 0; i < mOpt.depth; ++i) mSlots.push_back(std::make_unique<Slot>());
  mThread = std::thread([this] { Run(); });
}

GenotypeBatcher::~GenotypeBatcher() {
  {
    std::lock_guard<std::mutex> lk(mMu);
    mStop = true;
  }
  mCv.notify_all();
  if (mThread.joinable()) mThread.join();
  mSlots.clear();
  lgr_destroy(mCtx);
}

GenotypeBatcher::Ticket GenotypeBatcher::Enqueue(const GenotypeJob& job) {
  Pending p;
  p.job = job;
  p.packed = std::make_unique<PackedJob>();
  p.packed->Build(job);  // on the enqueuing worker: the batcher thread only concatenates
  Ticket t;
  t.job = job;
  t.done = p.done.get_future();
  {
    std::lock_guard<std::mutex> lk(mMu);
    if (mStop) throw std::runtime_error("lancet_gpu::GenotypeBatcher: shut down");
    mQueue.push_back(std::move(p));
  }
  mCv.notify_all();
  return t;
}

Result GenotypeBatcher::Collect(Ticket& ticket) {
  const std::vector<lgr_assign> assign = ticket.done.get();  // rethrows a device error on this thread
  // AddToTable on the collecting thread: the serial batcher thread only moves bytes
  return PackedBatch::BuildResult(ticket.job, assign.data(), mNameHash);
}

Result GenotypeBatcher::Genotype(const std::string* haps, std::size_t n_haps, const ReadIn* reads, std::size_t n_reads,
                                 const VariantIn* variants, std::size_t n_variants) {
  Ticket t = Enqueue(GenotypeJob{haps, n_haps, reads, n_reads, variants, n_variants});
  return Collect(t);
}

GenotypeBatcher::Counters GenotypeBatcher::Stats() {
  std::lock_guard<std::mutex> lk(mMu);
  return mCounters;
}

static std::uint64_t NowNs() {
  return static_cast<std::uint64_t>(
      std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

void GenotypeBatcher::Complete(Slot& s) {
  std::exception_ptr err;
  const std::uint64_t t0 = NowNs();
  try {
    Check(mCtx, lgr_wait(mCtx, s.ticket, nullptr));
  } catch (...) {
    err = std::current_exception();
  }
  const std::uint64_t t1 = NowNs();
  s.ticket = -1;
  for (std::size_t j = 0; j < s.jobs.size(); ++j) {
    Pending& p = s.jobs[j];
    if (err) {
      p.done.set_exception(err);
      continue;
    }
    const lgr_assign* a = s.pb.JobAssign(j);
    p.done.set_value(std::vector<lgr_assign>(a, a + p.job.n_reads * p.job.n_variants));
  }
  s.jobs.clear();
  const std::uint64_t t2 = NowNs();
  std::lock_guard<std::mutex> lk(mMu);
  mCounters.ns_wait += t1 - t0, mCounters.ns_deliver += t2 - t1;
}

void GenotypeBatcher::Run() {
  std::size_t head = 0, tail = 0, inflight = 0;  // slots [tail, head) hold submitted batches
  for (;;) {
    std::vector<Pending> take;
    {
      std::unique_lock<std::mutex> lk(mMu);
      if (inflight == 0) {
        mCv.wait(lk, [&] { return mStop || !mQueue.empty(); });
        if (mQueue.empty()) break;  // stop requested and nothing left
        // the GPU is idle: give the other workers a moment to arrive so they share the launch
        if (mOpt.linger_us > 0 && !mStop)
          mCv.wait_for(lk, std::chrono::microseconds(mOpt.linger_us), [&] { return mStop || mQueue.size() >= mOpt.max_jobs; });
      }
      std::int64_t pairs = 0;
      while (!mQueue.empty() && take.size() < mOpt.max_jobs) {
        const GenotypeJob& j = mQueue.front().job;
        const std::int64_t p = static_cast<std::int64_t>(j.n_reads) * static_cast<std::int64_t>(j.n_haps);
        if (!take.empty() && pairs + p > mOpt.max_pairs) break;
        pairs += p;
        take.push_back(std::move(mQueue.front()));
        mQueue.pop_front();
      }
      if (!take.empty()) {
        mCounters.batches += 1, mCounters.jobs += take.size(), mCounters.pairs += static_cast<std::uint64_t>(pairs);
        if (take.size() > mCounters.max_jobs_in_batch) mCounters.max_jobs_in_batch = take.size();
      }
    }
    if (take.empty()) {  // nothing new to pack: hand the oldest batch back as soon as it is done
      Complete(*mSlots[tail % mSlots.size()]);
      ++tail, --inflight;
      continue;
    }
    if (inflight == mSlots.size()) {
      Complete(*mSlots[tail % mSlots.size()]);
      ++tail, --inflight;
    }
    Slot& s = *mSlots[head % mSlots.size()];
    s.jobs = std::move(take);
    try {
      std::vector<GenotypeJob> jobs;
      jobs.reserve(s.jobs.size());
      for (const Pending& p : s.jobs) jobs.push_back(p.job);
      LatchMidOcc(mCtx, mParams, jobs.data(), jobs.size(), &mLatchedMidOcc);
      const std::uint64_t t0 = NowNs();
      std::vector<const PackedJob*> packed;
      packed.reserve(s.jobs.size());
      for (const Pending& p : s.jobs) packed.push_back(p.packed.get());
      s.pb.PackPrepared(packed.data(), packed.size(), mLatchedMidOcc);
      for (Pending& p : s.jobs) p.packed.reset();
      const std::uint64_t t1 = NowNs();
      Check(mCtx, lgr_submit(mCtx, &s.pb.In(), &s.pb.Out(), &s.ticket));
      const std::uint64_t t2 = NowNs();
      ++head, ++inflight;
      std::lock_guard<std::mutex> lk(mMu);
      mCounters.ns_pack += t1 - t0, mCounters.ns_submit += t2 - t1;
    } catch (...) {
      for (Pending& p : s.jobs) p.done.set_exception(std::current_exception());
      s.jobs.clear();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GenotypeDispatcher
// ---------------------------------------------------------------------------------------------
GenotypeDispatcher::GenotypeDispatcher(const std::vector<int>& devices, NameHashFn name_hash, GenotypeBatcher::Options base) {
  if (devices.empty()) throw std::runtime_error("lancet_gpu::GenotypeDispatcher: no devices given (there is no CPU fallback)");
  mOutstanding = std::make_unique<std::atomic<std::int64_t>[]>(devices.size());
  for (std::size_t i = 0; i < devices.size(); ++i) {
    base.device = devices[i];
    mOutstanding[i].store(0);
    mBatchers.push_back(std::make_unique<GenotypeBatcher>(base, name_hash));
  }
}

std::int64_t GenotypeDispatcher::Cost(const GenotypeJob& job) {
  std::int64_t hap_total = 0;
  for (std::size_t h = 0; h < job.n_haps; ++h) hap_total += static_cast<std::int64_t>(job.haps[h].size());
  return static_cast<std::int64_t>(job.n_reads) * hap_total + 1;
}

GenotypeDispatcher::Ticket GenotypeDispatcher::Enqueue(const GenotypeJob& job) {
  std::size_t best = 0;
  std::int64_t best_load = mOutstanding[0].load(std::memory_order_relaxed);
  for (std::size_t i = 1; i < mBatchers.size(); ++i) {
    const std::int64_t load = mOutstanding[i].load(std::memory_order_relaxed);
    if (load < best_load) best = i, best_load = load;
  }
  Ticket t;
  t.device_slot = best;
  t.cost = Cost(job);
  mOutstanding[best].fetch_add(t.cost, std::memory_order_relaxed);
  try {
    t.inner = mBatchers[best]->Enqueue(job);
  } catch (...) {
    mOutstanding[best].fetch_sub(t.cost, std::memory_order_relaxed);
    throw;
  }
  return t;
}

Result GenotypeDispatcher::Collect(Ticket& ticket) {
  struct Release {
    std::atomic<std::int64_t>& load;
    std::int64_t cost;
    ~Release() { load.fetch_sub(cost, std::memory_order_relaxed); }
  } release{mOutstanding[ticket.device_slot], ticket.cost};
  return mBatchers[ticket.device_slot]->Collect(ticket.inner);
}

Result GenotypeDispatcher::Genotype(const std::string* haps, std::size_t n_haps, const ReadIn* reads, std::size_t n_reads,
                                    const VariantIn* variants, std::size_t n_variants) {
  Ticket t = Enqueue(GenotypeJob{haps, n_haps, reads, n_reads, variants, n_variants});
  return Collect(t);
}

}  // namespace lancet_gpu
