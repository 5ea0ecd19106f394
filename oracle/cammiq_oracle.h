/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement of CAMMiQ's query-time read-matching path, written from the
 * reference's behaviour (citations = /root/reference/src/<file>:<lines>):
 *
 *   index codec      binaryio.cpp:141-214 (BitReader), hashtrie.cpp:425-507 (decodeTrie_p, loadIdx64_p)
 *   prefix hash      hashtrie.cpp:132-137 (computeHashVal64), query.cpp:482-495 (rolling form)
 *   lookup           hashtrie.cpp:350-369 (find64_p)
 *   reverse compl.   query.cpp:447-450 (getRC), tables query.cpp:1860-1883
 *   scan             query.cpp:480-527
 *   classification   query.cpp:529-636 (query64_p / query64mt_p), 964-1067 (query64_sc)
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF: oracle/_ref/ref_harness (the
 * unmodified reference sources compiled by oracle/Makefile) on seeded inputs, live in
 * tests/test_oracle_vs_ref.py when oracle/_ref exists and through the committed dumps in
 * tests/golden/ otherwise.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use this library, and only as the checker.
 */
#ifndef CAMMIQ_ORACLE_H
#define CAMMIQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CQO_NONE UINT64_MAX

/* Per-read decision classes (shared vocabulary with include/cammiq_gpu.h). */
enum {
	CQO_CLASS_UNLABELED = 0, /* nundet++                                    query.cpp:544-545 */
	CQO_CLASS_CONFLICT = 1,  /* nconf++                                                       */
	CQO_CLASS_U = 2,         /* u[a]++                                      query.cpp:547-551 */
	CQO_CLASS_D_PAIR = 3,    /* d[a]++, d[b]++ (sc: pair map too)           query.cpp:557-563 */
	CQO_CLASS_UD = 4,        /* u[a]++, d[a]++                     query.cpp:573-576, 597-600 */
	CQO_CLASS_D_INTER = 5    /* p: d[a]++ ; sc: u[a]++ and d[a]++ query.cpp:624-629, 1055-1059 */
};

enum { CQO_MODE_P = 0, CQO_MODE_SC = 1 };

typedef struct cqo_index cqo_index;

/* Decode <path> (INT stream) and <path>.aux (bit stream).  NULL on I/O or format error. */
cqo_index *cqo_index_load(const char *path);
void cqo_index_free(cqo_index *idx);
uint32_t cqo_index_hash_len(const cqo_index *idx);
int cqo_index_is_doubly_unique(const cqo_index *idx);
uint64_t cqo_index_num_buckets(const cqo_index *idx);
uint64_t cqo_index_num_leaves(const cqo_index *idx);
/* Leaf fields in FILE order (the order decodeTrie_p creates them). */
const uint32_t *cqo_leaf_ref1(const cqo_index *idx);
const uint32_t *cqo_leaf_ref2(const cqo_index *idx);
const uint16_t *cqo_leaf_ucount1(const cqo_index *idx);
const uint16_t *cqo_leaf_ucount2(const cqo_index *idx);
const uint8_t *cqo_leaf_depth(const cqo_index *idx);
/* file-order leaf id -> rank of its key in lexicographic order (ref_harness' canonical id). */
void cqo_canonical_ids(const cqo_index *idx, uint64_t *out);
/* map_sp as CSR: for rid in 1..G the file-order leaf ids carrying rid (hashtrie.cpp:452-453,476).
   offsets has G+2 entries (offsets[rid]..offsets[rid+1]); returns total entries; ids may be NULL
   to size the array. */
uint64_t cqo_map_sp(const cqo_index *idx, uint32_t G, uint64_t *offsets, uint64_t *ids);

/* Hash::computeHashVal64 and Hash::find64_p.  Returns file-order leaf id or CQO_NONE. */
uint64_t cqo_hash(const uint8_t *key, uint32_t h);
uint64_t cqo_find(const cqo_index *idx, uint64_t bucket, const uint8_t *cand, size_t len);

typedef struct {
	uint64_t *cnt_u;    /* [G+1], index 0 unused: Genome::read_cnts_u */
	uint64_t *cnt_d;    /* [G+1]: Genome::read_cnts_d */
	uint64_t nundet, nconf;
	uint64_t n_invalid; /* reads outside the reference's defined domain (see cqo_query) */
	uint32_t *rcount_u; /* [nU] pleafNode::rcount, file order; untouched in SC mode; may be NULL */
	uint32_t *rcount_d; /* [nD] */
	/* SC mode: read_cnts_b as parallel arrays sorted by (a,b); capacity pairs_cap. */
	uint32_t *pair_a, *pair_b;
	uint64_t *pair_cnt;
	uint64_t n_pairs, pairs_cap;
	/* Optional per-read outputs (NULL to skip). */
	uint8_t *read_class;  /* [n_reads] */
	uint32_t *read_rid_a; /* [n_reads] */
	uint32_t *read_rid_b; /* [n_reads] */
	/* Optional per-read distinct leaf sets: up to leaf_cap ids per read per table, sorted
	   ascending (file-order ids); counts are the true set sizes. */
	uint32_t leaf_cap;
	uint32_t *read_nleaf_u, *read_nleaf_d; /* [n_reads] */
	uint32_t *read_leaf_u, *read_leaf_d;   /* [n_reads * leaf_cap] */
} cqo_result;

/*
 * query64_p / query64mt_p (mode P) or query64_sc (mode SC) over n_reads reads stored as
 * ASCII in bases[offsets[i] .. offsets[i]+lengths[i]).  Counters are ACCUMULATED into out.
 * Reads the reference cannot process are defined here the way the product defines them and
 * are counted in n_invalid as well as nundet: length < h (query.cpp:486 underflows) and
 * reads holding a byte outside ACGTacgt (symbolIdx = -1, undefined behaviour).
 * Returns 0, or -1 on bad arguments (hash lengths differ, refID out of 1..G, pair overflow).
 */
int cqo_query(const cqo_index *u, const cqo_index *d, int mode, uint32_t G, const uint8_t *bases,
		const uint64_t *offsets, const uint8_t *lengths, uint64_t n_reads, cqo_result *out);

#ifdef __cplusplus
}
#endif
#endif
