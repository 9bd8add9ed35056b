/* TEST INFRASTRUCTURE shim: opaque minimap2 types so caller/genotyper.h parses.
 * No minimap2 function is ever called through this shim. */
#ifndef SHIM_MINIMAP_H_
#define SHIM_MINIMAP_H_
typedef struct mm_idx_s { int unused; } mm_idx_t;
typedef struct mm_mapopt_s { int unused; } mm_mapopt_t;
typedef struct mm_idxopt_s { int unused; } mm_idxopt_t;
typedef struct mm_tbuf_s mm_tbuf_t;
static inline void mm_idx_destroy(mm_idx_t*) {}
static inline void mm_tbuf_destroy(mm_tbuf_t*) {}
static inline mm_tbuf_t* mm_tbuf_init(void) { return 0; }
#endif
