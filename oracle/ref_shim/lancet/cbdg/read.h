// TEST INFRASTRUCTURE shim: caller/genotyper.h only needs cbdg::Read as a name
// (the real header drags in htslib/abseil-status/spdlog).  Nothing on the scoring
// path touches it.
#ifndef SHIM_LANCET_CBDG_READ_H_
#define SHIM_LANCET_CBDG_READ_H_
namespace lancet::cbdg { class Read; }
#endif
