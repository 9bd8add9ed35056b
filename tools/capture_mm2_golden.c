/* capture_mm2_golden.c — dump what the REAL minimap2 returns for the Genotyper's option set, so that
 * the from-memory restatement in oracle/mm2_restate.cpp (and with it every "bit-exact" claim of this
 * repository) can be pinned.  SURVEY.md §8c(iii), VERDICT r1 "Next round" #6.
 *
 * minimap2 is NOT part of this repository or of /root/reference (Lancet2 downloads v2.30 at CMake
 * time, cmake/dependencies.cmake:163-166), and the build container has no network: this file cannot be
 * compiled here.  On any machine with a minimap2 2.30 checkout:
 *
 *     make -C tools capture MM2=/path/to/minimap2        # needs minimap.h + libminimap2.a
 *     python tools/export_mm2_cases.py cases > /tmp/mm2_cases.txt
 *     tools/_build/capture_mm2_golden /tmp/mm2_cases.txt | gzip > tests/golden/mm2_golden.jsonl.gz
 *     python -m pytest tests/test_mm2_golden.py          # oracle (CPU) and CUDA path (GPU) against it
 *
 * It does exactly what lancet::caller::Genotyper does (reference src/lancet/caller/genotyper.cpp):
 *   options        :89-191  mm_set_opt(0) + the overrides, k = 11, w = 5
 *   per group      :243-267 mm_idx_str(w, k, 0, bucket_bits, 1, &seq, NULL) per haplotype, then
 *                           mm_mapopt_update for every index (latches mid_occ from the first one ever)
 *   per read x hap :385-404 mm_map(idx, len, seq, &n, tbuf, opt, qname), regs[0]
 * One options struct lives for the whole run, like one Genotyper on one worker thread; the latched
 * mid_occ is printed per group so that the comparison can hand the same value to lgr_batch_in::grp_mid_occ.
 *
 * Input (tools/export_mm2_cases.py):  "G <n_haps> <n_reads>", then n_haps lines "H <seq>", then n_reads
 * lines "R <qname> <seq>".  Output: one JSON object per (group, read, haplotype) line; the field names are
 * those of mm_reg1_t / mm_extra_t (minimap.h), cigar = the raw uint32 ops (len << 4 | op).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "minimap.h"

static char* read_line(FILE* f, size_t* cap, char** buf) {
  ssize_t n = getline(buf, cap, f);
  if (n <= 0) return NULL;
  while (n > 0 && ((*buf)[n - 1] == '\n' || (*buf)[n - 1] == '\r')) (*buf)[--n] = 0;
  return *buf;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s cases.txt > golden.jsonl\n", argv[0]);
    return 2;
  }
  FILE* f = fopen(argv[1], "r");
  if (!f) { perror(argv[1]); return 1; }
  mm_idxopt_t iopt;
  mm_mapopt_t mopt;
  mm_verbose = 1;
  mm_set_opt(0, &iopt, &mopt);                    /* genotyper.cpp:94 */
  mopt.flag |= MM_F_CIGAR | MM_F_SR;              /* :109 */
  mopt.best_n = 1;                                /* :110 */
  mopt.a = 1, mopt.b = 4, mopt.q = 12, mopt.e = 3, mopt.q2 = 12, mopt.e2 = 3; /* :118-123, scoring_constants.h:17-20 */
  mopt.zdrop = 100000, mopt.zdrop_inv = 100000;   /* :131-132 */
  mopt.bw = 10000;                                /* :140 */
  mopt.max_gap = 200, mopt.max_gap_ref = 5000;    /* :158-159 */
  mopt.end_bonus = 10000;                         /* :181 */
  iopt.k = 11, iopt.w = 5;                        /* :189-190 */
  mm_tbuf_t* tbuf = mm_tbuf_init();
  size_t cap = 0;
  char* line = NULL;
  long g = 0;
  while (read_line(f, &cap, &line)) {
    int n_haps = 0, n_reads = 0;
    if (sscanf(line, "G %d %d", &n_haps, &n_reads) != 2) { fprintf(stderr, "bad group header: %s\n", line); return 1; }
    mm_idx_t** idx = (mm_idx_t**)calloc((size_t)n_haps, sizeof(*idx));
    for (int h = 0; h < n_haps; ++h) {
      if (!read_line(f, &cap, &line) || line[0] != 'H') { fprintf(stderr, "expected H line\n"); return 1; }
      const char* seq = line + 2;
      idx[h] = mm_idx_str(iopt.w, iopt.k, 0, iopt.bucket_bits, 1, &seq, NULL);  /* :250 */
    }
    for (int h = 0; h < n_haps; ++h) mm_mapopt_update(&mopt, idx[h]);            /* :263-266 */
    printf("{\"g\":%ld,\"mid_occ\":%d}\n", g, mopt.mid_occ);
    for (int r = 0; r < n_reads; ++r) {
      if (!read_line(f, &cap, &line) || line[0] != 'R') { fprintf(stderr, "expected R line\n"); return 1; }
      char* name = line + 2;
      char* seq = strchr(name, ' ');
      if (!seq) { fprintf(stderr, "bad R line\n"); return 1; }
      *seq++ = 0;
      const int len = (int)strlen(seq);
      for (int h = 0; h < n_haps; ++h) {
        int n_regs = 0;
        mm_reg1_t* regs = mm_map(idx[h], len, seq, &n_regs, tbuf, &mopt, name);  /* :387-388 */
        printf("{\"g\":%ld,\"r\":%d,\"h\":%d,\"n_regs\":%d", g, r, h, n_regs);
        if (regs && n_regs > 0) {
          const mm_reg1_t* t = &regs[0];                                          /* :396 */
          printf(",\"score\":%d,\"rs\":%d,\"re\":%d,\"qs\":%d,\"qe\":%d,\"rev\":%d,\"cnt\":%d,\"mlen\":%d,\"blen\":%d", t->score, t->rs,
                 t->re, t->qs, t->qe, (int)t->rev, t->cnt, t->mlen, t->blen);
          if (t->p) {
            printf(",\"dp_score\":%d,\"dp_max\":%d,\"n_ambi\":%u,\"cigar\":[", t->p->dp_score, t->p->dp_max, (unsigned)t->p->n_ambi);
            for (uint32_t k = 0; k < t->p->n_cigar; ++k) printf("%s%u", k ? "," : "", t->p->cigar[k]);
            printf("]");
          }
        }
        printf("}\n");
        for (int i = 0; i < n_regs; ++i) free(regs[i].p);
        free(regs);
      }
    }
    for (int h = 0; h < n_haps; ++h) mm_idx_destroy(idx[h]);
    free(idx);
    ++g;
  }
  mm_tbuf_destroy(tbuf);
  free(line);
  fclose(f);
  return 0;
}
