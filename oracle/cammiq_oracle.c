/*
 * TEST INFRASTRUCTURE ONLY -- see cammiq_oracle.h for the scope statement, the reference
 * citations and how this restatement is pinned against the reference itself.
 * Plain C99, single-threaded, written for clarity: it mirrors the reference's data flow
 * (pointer trie -> index arrays, std::set -> sorted arrays) rather than the GPU design.
 */
#include "cammiq_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- base tables (query.cpp:1860-1883; hashtrie.cpp:701-714) ------------------------ */

static int sym_code(uint8_t c) {
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return -1;
	}
}

/* rcIdx: complement, upper-case output (query.cpp rcIdx[128]). */
static uint8_t rc_base(uint8_t c) {
	switch (c) {
	case 'A': case 'a': return 'T';
	case 'C': case 'c': return 'G';
	case 'G': case 'g': return 'C';
	case 'T': case 't': return 'A';
	default: return 0;
	}
}

/* ---- index container ----------------------------------------------------------------- */

struct cqo_index {
	int is_d;
	uint32_t h;
	uint64_t n_buckets, n_nodes, n_leaves;
	uint64_t cap_buckets, cap_nodes, cap_leaves;
	uint64_t *bucket_key; /* file order */
	int64_t *bucket_root; /* node id, -1 = NULL root */
	int64_t *child;       /* 4 per node, -1 = NULL (trieNode::children, hashtrie.cpp:8-13) */
	uint8_t *is_end;      /* trieNode::isEnd */
	uint64_t *node_leaf;  /* leaf id for isEnd nodes */
	uint32_t *ref1, *ref2;
	uint16_t *uc1, *uc2;
	uint8_t *depth;
	uint64_t *order; /* bucket indices sorted by key (stands in for map64.find) */
	/* BitReader state (binaryio.hpp:44-52) */
	unsigned char *buf_int, *buf_aux;
	size_t cur_int, cur_aux, size_int, size_aux;
	int cur_bits, cur_byte;
	int error;
};

static unsigned char *slurp(const char *fn, size_t *size) {
	FILE *f = fopen(fn, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	unsigned char *b = (unsigned char *) malloc((size_t) n + 16);
	if (b && n > 0 && fread(b, 1, (size_t) n, f) != (size_t) n) {
		free(b);
		b = NULL;
	}
	fclose(f);
	*size = (size_t) n;
	return b;
}

/* BitReader::readBit (binaryio.cpp:141-155): MSB first, all-ones past EOF. */
static uint32_t read_bit(cqo_index *x) {
	if (x->cur_bits == 0) {
		x->cur_bits = 8;
		x->cur_byte = (x->cur_aux < x->size_aux) ? x->buf_aux[x->cur_aux++] : 0xFF;
	}
	uint32_t v = ((uint32_t) x->cur_byte >> (x->cur_bits - 1)) & 1u;
	x->cur_bits--;
	return v;
}

static uint32_t read_bits(cqo_index *x, int count) {
	uint32_t v = 0;
	for (int i = 0; i < count; i++) v = (v << 1) + read_bit(x);
	return v;
}

/* BitReader::readBits16/32/64 (binaryio.cpp:164-182): big-endian bytes of the INT stream. */
static uint64_t read_be(cqo_index *x, int nbytes) {
	uint64_t v = 0;
	for (int i = 0; i < nbytes; i++) {
		unsigned char b = 0xFF;
		if (x->cur_int < x->size_int) b = x->buf_int[x->cur_int];
		else x->error = 1;
		x->cur_int++;
		v = (v << 8) | b;
	}
	return v;
}

static int grow(void **p, uint64_t *cap, uint64_t need, size_t elem) {
	if (need <= *cap) return 0;
	uint64_t nc = *cap ? *cap * 2 : 1024;
	while (nc < need) nc *= 2;
	void *q = realloc(*p, (size_t) nc * elem);
	if (!q) return -1;
	*p = q;
	*cap = nc;
	return 0;
}

static int64_t new_node(cqo_index *x) {
	uint64_t cap = x->cap_nodes;
	if (x->n_nodes + 1 > cap) {
		uint64_t c1 = cap, c2 = cap, c3 = cap;
		if (grow((void **) &x->child, &c1, (x->n_nodes + 1), 4 * sizeof(int64_t)) ||
			grow((void **) &x->is_end, &c2, x->n_nodes + 1, 1) ||
			grow((void **) &x->node_leaf, &c3, x->n_nodes + 1, sizeof(uint64_t))) {
			x->error = 1;
			return -1;
		}
		x->cap_nodes = c1;
	}
	int64_t id = (int64_t) x->n_nodes++;
	for (int i = 0; i < 4; i++) x->child[4 * id + i] = -1;
	x->is_end[id] = 0;
	x->node_leaf[id] = CQO_NONE;
	return id;
}

static uint64_t new_leaf(cqo_index *x) {
	if (x->n_leaves + 1 > x->cap_leaves) {
		uint64_t c = x->cap_leaves, c1 = c, c2 = c, c3 = c, c4 = c, c5 = c;
		if (grow((void **) &x->ref1, &c1, x->n_leaves + 1, 4) || grow((void **) &x->ref2, &c2, x->n_leaves + 1, 4) ||
			grow((void **) &x->uc1, &c3, x->n_leaves + 1, 2) || grow((void **) &x->uc2, &c4, x->n_leaves + 1, 2) ||
			grow((void **) &x->depth, &c5, x->n_leaves + 1, 1)) {
			x->error = 1;
			return CQO_NONE;
		}
		x->cap_leaves = c1;
	}
	return x->n_leaves++;
}

/* Hash::decodeTrie_p (hashtrie.cpp:425-484).  `depth` wraps as uint8_t like the reference. */
static int64_t decode_trie(cqo_index *x, uint8_t depth, int guard) {
	if (x->error || guard > 4096) {
		x->error = 1;
		return -1;
	}
	if (read_bit(x) == 0) return -1;
	int64_t root = new_node(x);
	if (root < 0) return -1;
	int is_leaf = 1;
	for (int i = 0; i < 4; i++) {
		int64_t c = decode_trie(x, (uint8_t) (depth + 1), guard + 1);
		x->child[4 * root + i] = c;
		if (c >= 0) is_leaf = 0;
	}
	if (is_leaf) {
		uint64_t l = new_leaf(x);
		if (l == CQO_NONE) return -1;
		x->depth[l] = (uint8_t) (depth + x->h);
		if (x->is_d) {
			x->ref1[l] = (uint32_t) read_be(x, 4);
			x->ref2[l] = (uint32_t) read_be(x, 4);
			if (x->ref1[l] == 0 || x->ref2[l] == 0) x->error = 1; /* assert, hashtrie.cpp:446 */
			x->uc1[l] = (uint16_t) read_be(x, 2);
			x->uc2[l] = (uint16_t) read_be(x, 2);
		} else {
			x->ref1[l] = (uint32_t) read_be(x, 4);
			x->ref2[l] = 0;
			x->uc1[l] = (uint16_t) read_be(x, 2);
			x->uc2[l] = 0;
		}
		x->is_end[root] = 1; /* the reference swaps in a pleafNode, isEnd = true */
		x->node_leaf[root] = l;
	}
	return root;
}

static const uint64_t *g_sort_keys;
static int cmp_order(const void *a, const void *b) {
	uint64_t ka = g_sort_keys[*(const uint64_t *) a], kb = g_sort_keys[*(const uint64_t *) b];
	if (ka != kb) return ka < kb ? -1 : 1;
	/* duplicate bucket keys: map64[bucket] = root keeps the LAST one (hashtrie.cpp:500) */
	uint64_t ia = *(const uint64_t *) a, ib = *(const uint64_t *) b;
	return ia < ib ? -1 : (ia > ib);
}

cqo_index *cqo_index_load(const char *path) {
	cqo_index *x = (cqo_index *) calloc(1, sizeof(cqo_index));
	if (!x) return NULL;
	size_t n = strlen(path);
	char *aux = (char *) malloc(n + 8);
	sprintf(aux, "%s.aux", path);
	x->buf_int = slurp(path, &x->size_int);
	x->buf_aux = slurp(aux, &x->size_aux);
	free(aux);
	if (!x->buf_int || !x->buf_aux) {
		cqo_index_free(x);
		return NULL;
	}
	/* Hash::loadIdx64_p (hashtrie.cpp:486-507) */
	x->is_d = (int) read_bit(x);
	uint32_t option = read_bits(x, 7);
	x->h = read_bits(x, 8);
	if (option != 64 || x->h < 1 || x->h > 32) x->error = 1;
	uint64_t bucket = read_be(x, 8);
	while (!x->error && bucket != UINT64_MAX) {
		uint64_t c1 = x->cap_buckets, c2 = x->cap_buckets;
		if (grow((void **) &x->bucket_key, &c1, x->n_buckets + 1, 8) ||
			grow((void **) &x->bucket_root, &c2, x->n_buckets + 1, 8)) {
			x->error = 1;
			break;
		}
		x->cap_buckets = c1;
		int64_t root = decode_trie(x, 0, 0);
		x->bucket_key[x->n_buckets] = bucket;
		x->bucket_root[x->n_buckets] = root;
		x->n_buckets++;
		bucket = read_be(x, 8);
	}
	free(x->buf_int);
	free(x->buf_aux);
	x->buf_int = x->buf_aux = NULL;
	if (x->error) {
		cqo_index_free(x);
		return NULL;
	}
	x->order = (uint64_t *) malloc((size_t) (x->n_buckets + 1) * 8);
	for (uint64_t i = 0; i < x->n_buckets; i++) x->order[i] = i;
	g_sort_keys = x->bucket_key;
	qsort(x->order, (size_t) x->n_buckets, 8, cmp_order);
	return x;
}

void cqo_index_free(cqo_index *x) {
	if (!x) return;
	free(x->bucket_key); free(x->bucket_root); free(x->child); free(x->is_end); free(x->node_leaf);
	free(x->ref1); free(x->ref2); free(x->uc1); free(x->uc2); free(x->depth); free(x->order);
	free(x->buf_int); free(x->buf_aux);
	free(x);
}

uint32_t cqo_index_hash_len(const cqo_index *x) { return x->h; }
int cqo_index_is_doubly_unique(const cqo_index *x) { return x->is_d; }
uint64_t cqo_index_num_buckets(const cqo_index *x) { return x->n_buckets; }
uint64_t cqo_index_num_leaves(const cqo_index *x) { return x->n_leaves; }
const uint32_t *cqo_leaf_ref1(const cqo_index *x) { return x->ref1; }
const uint32_t *cqo_leaf_ref2(const cqo_index *x) { return x->ref2; }
const uint16_t *cqo_leaf_ucount1(const cqo_index *x) { return x->uc1; }
const uint16_t *cqo_leaf_ucount2(const cqo_index *x) { return x->uc2; }
const uint8_t *cqo_leaf_depth(const cqo_index *x) { return x->depth; }

/* map64.find(bucket): the entry that survived insertion (last duplicate wins). */
static int64_t find_root(const cqo_index *x, uint64_t bucket, int *found) {
	uint64_t lo = 0, hi = x->n_buckets;
	while (lo < hi) {
		uint64_t mid = (lo + hi) / 2;
		if (x->bucket_key[x->order[mid]] <= bucket) lo = mid + 1;
		else hi = mid;
	}
	if (lo == 0 || x->bucket_key[x->order[lo - 1]] != bucket) {
		*found = 0;
		return -1;
	}
	*found = 1;
	return x->bucket_root[x->order[lo - 1]];
}

static void number_leaves(const cqo_index *x, int64_t node, uint64_t *out, uint64_t *next) {
	if (node < 0) return;
	if (x->is_end[node]) {
		out[x->node_leaf[node]] = (*next)++;
		return;
	}
	for (int c = 0; c < 4; c++) number_leaves(x, x->child[4 * node + c], out, next);
}

void cqo_canonical_ids(const cqo_index *x, uint64_t *out) {
	uint64_t next = 0;
	for (uint64_t i = 0; i < x->n_leaves; i++) out[i] = CQO_NONE;
	for (uint64_t i = 0; i < x->n_buckets; i++) {
		/* skip shadowed duplicates: only the last bucket with a given key is reachable */
		if (i + 1 < x->n_buckets && x->bucket_key[x->order[i + 1]] == x->bucket_key[x->order[i]]) continue;
		number_leaves(x, x->bucket_root[x->order[i]], out, &next);
	}
}

uint64_t cqo_map_sp(const cqo_index *x, uint32_t G, uint64_t *offsets, uint64_t *ids) {
	uint64_t *cnt = (uint64_t *) calloc((size_t) G + 2, 8);
	for (uint64_t l = 0; l < x->n_leaves; l++) {
		if (x->ref1[l] >= 1 && x->ref1[l] <= G) cnt[x->ref1[l]]++;
		if (x->is_d && x->ref2[l] >= 1 && x->ref2[l] <= G) cnt[x->ref2[l]]++;
	}
	uint64_t total = 0;
	offsets[0] = 0;
	for (uint32_t r = 0; r <= G; r++) {
		offsets[r] = total;
		total += cnt[r];
	}
	offsets[G + 1] = total;
	if (ids) {
		uint64_t *fill = (uint64_t *) calloc((size_t) G + 2, 8);
		for (uint64_t l = 0; l < x->n_leaves; l++) {
			uint32_t a = x->ref1[l], b = x->ref2[l];
			if (a >= 1 && a <= G) ids[offsets[a] + fill[a]++] = l;
			if (x->is_d && b >= 1 && b <= G) ids[offsets[b] + fill[b]++] = l;
		}
		free(fill);
	}
	free(cnt);
	return total;
}

/* Hash::computeHashVal64 (hashtrie.cpp:132-137). */
uint64_t cqo_hash(const uint8_t *key, uint32_t h) {
	uint64_t res = 0;
	for (uint32_t i = 0; i < h; i++) res = (res << 2) | (uint64_t) sym_code(key[i]);
	return res;
}

/* Hash::find64_p (hashtrie.cpp:350-369). */
uint64_t cqo_find(const cqo_index *x, uint64_t bucket, const uint8_t *cand, size_t len) {
	int found;
	int64_t cur = find_root(x, bucket, &found);
	if (!found || cur < 0) return CQO_NONE; /* NULL root: the reference would fault; never written */
	for (size_t i = 0; i < len; i++) {
		int index = sym_code(cand[i]);
		if (x->is_end[cur]) return x->node_leaf[cur];
		if (x->child[4 * cur + index] < 0) return CQO_NONE;
		cur = x->child[4 * cur + index];
	}
	if (x->is_end[cur]) return x->node_leaf[cur];
	return CQO_NONE;
}

/* ---- tiny sorted sets (std::set stand-ins) ------------------------------------------- */

typedef struct { uint64_t *v; size_t n, cap; } set64;

static void set_insert(set64 *s, uint64_t key) {
	size_t lo = 0, hi = s->n;
	while (lo < hi) {
		size_t mid = (lo + hi) / 2;
		if (s->v[mid] < key) lo = mid + 1;
		else hi = mid;
	}
	if (lo < s->n && s->v[lo] == key) return;
	if (s->n == s->cap) {
		s->cap = s->cap ? s->cap * 2 : 64;
		s->v = (uint64_t *) realloc(s->v, s->cap * 8);
	}
	memmove(s->v + lo + 1, s->v + lo, (s->n - lo) * 8);
	s->v[lo] = key;
	s->n++;
}

static void add_pair(cqo_result *out, uint32_t a, uint32_t b, int *err) {
	uint64_t lo = 0, hi = out->n_pairs;
	while (lo < hi) {
		uint64_t mid = (lo + hi) / 2;
		if (out->pair_a[mid] < a || (out->pair_a[mid] == a && out->pair_b[mid] < b)) lo = mid + 1;
		else hi = mid;
	}
	if (lo < out->n_pairs && out->pair_a[lo] == a && out->pair_b[lo] == b) {
		out->pair_cnt[lo]++;
		return;
	}
	if (out->n_pairs >= out->pairs_cap) {
		*err = 1;
		return;
	}
	for (uint64_t i = out->n_pairs; i > lo; i--) {
		out->pair_a[i] = out->pair_a[i - 1];
		out->pair_b[i] = out->pair_b[i - 1];
		out->pair_cnt[i] = out->pair_cnt[i - 1];
	}
	out->pair_a[lo] = a;
	out->pair_b[lo] = b;
	out->pair_cnt[lo] = 1;
	out->n_pairs++;
}

#define TABLE_D (1ull << 63)

int cqo_query(const cqo_index *u, const cqo_index *d, int mode, uint32_t G, const uint8_t *bases,
		const uint64_t *offsets, const uint8_t *lengths, uint64_t n_reads, cqo_result *out) {
	if (!u || !d || !out || u->h != d->h) return -1; /* assert(hash_len_u == hash_len_d), query.cpp:460 */
	const uint32_t h = u->h;
	for (uint64_t l = 0; l < u->n_leaves; l++)
		if (u->ref1[l] < 1 || u->ref1[l] > G) return -1;
	for (uint64_t l = 0; l < d->n_leaves; l++)
		if (d->ref1[l] < 1 || d->ref1[l] > G || d->ref2[l] < 1 || d->ref2[l] > G) return -1;

	set64 pnodes = {0, 0, 0}, rids = {0, 0, 0}, rid_pairs = {0, 0, 0};
	uint8_t rc_read[256];
	int err = 0;

	for (uint64_t r = 0; r < n_reads; r++) {
		const uint8_t *read = bases + offsets[r];
		size_t rl = lengths[r];
		pnodes.n = rids.n = rid_pairs.n = 0;
		uint8_t cls = CQO_CLASS_UNLABELED;
		uint32_t rid_a = 0, rid_b = 0;

		int valid = rl >= h;
		for (size_t i = 0; i < rl && valid; i++)
			if (sym_code(read[i]) < 0) valid = 0;
		if (!valid) {
			out->n_invalid++;
			out->nundet++;
			goto record;
		}

		/* query.cpp:480-527: forward strand, then the reverse complement (getRC, 447-450). */
		for (int strand = 0; strand < 2; strand++) {
			const uint8_t *s = read;
			if (strand == 1) {
				for (size_t i = 0; i < rl; i++) rc_read[i] = rc_base(read[rl - i - 1]);
				s = rc_read;
			}
			uint32_t hs = 2 * h - 2;
			uint64_t hv = 0;
			for (size_t i = 0; i < h; i++) hv = (hv << 2) | (uint64_t) sym_code(s[i]);
			for (size_t i = 0; i < rl - h; i++) {
				uint64_t l = cqo_find(u, hv, s + i + h, rl - h - i);
				if (l != CQO_NONE) set_insert(&pnodes, l);
				l = cqo_find(d, hv, s + i + h, rl - h - i);
				if (l != CQO_NONE) set_insert(&pnodes, l | TABLE_D);
				hv = hv - ((uint64_t) sym_code(s[i]) << hs);
				hv = (hv << 2) | (uint64_t) sym_code(s[i + h]);
			}
			uint64_t l = cqo_find(u, hv, s, 0);
			if (l != CQO_NONE) set_insert(&pnodes, l);
			l = cqo_find(d, hv, s, 0);
			if (l != CQO_NONE) set_insert(&pnodes, l | TABLE_D);
		}

		/* query.cpp:529-540 */
		for (size_t i = 0; i < pnodes.n; i++) {
			uint64_t k = pnodes.v[i];
			if (!(k & TABLE_D)) {
				set_insert(&rids, u->ref1[k]); /* refID2 == 0 */
			} else {
				uint32_t a = d->ref1[k & ~TABLE_D], b = d->ref2[k & ~TABLE_D];
				if (a < b) set_insert(&rid_pairs, ((uint64_t) a << 32) | b);
				else set_insert(&rid_pairs, ((uint64_t) b << 32) | a);
			}
		}

		/* query.cpp:542-636 (mode P) / 977-1067 (mode SC) */
		switch (rid_pairs.n) {
		case 0:
			if (rids.n == 0) cls = CQO_CLASS_UNLABELED;
			else if (rids.n == 1) { cls = CQO_CLASS_U; rid_a = (uint32_t) rids.v[0]; }
			else cls = CQO_CLASS_CONFLICT;
			break;
		case 1: {
			uint32_t a = (uint32_t) (rid_pairs.v[0] >> 32), b = (uint32_t) rid_pairs.v[0];
			if (rids.n == 0) { cls = CQO_CLASS_D_PAIR; rid_a = a; rid_b = b; }
			else if (rids.n > 1) cls = CQO_CLASS_CONFLICT;
			else {
				uint32_t rid = (uint32_t) rids.v[0];
				if (a != rid && b != rid) cls = CQO_CLASS_CONFLICT;
				else { cls = CQO_CLASS_UD; rid_a = rid; }
			}
			break;
		}
		default:
			if (rids.n > 0) {
				if (rids.n > 1) { cls = CQO_CLASS_CONFLICT; break; }
				uint32_t rid = (uint32_t) rids.v[0];
				int conf = 0;
				for (size_t i = 0; i < rid_pairs.n; i++) {
					uint32_t a = (uint32_t) (rid_pairs.v[i] >> 32), b = (uint32_t) rid_pairs.v[i];
					if (a != rid && b != rid) { conf = 1; break; }
				}
				if (conf) cls = CQO_CLASS_CONFLICT;
				else { cls = CQO_CLASS_UD; rid_a = rid; }
			} else {
				uint32_t inter[2];
				int ni = 0;
				for (size_t i = 0; i < rid_pairs.n; i++) {
					uint32_t a = (uint32_t) (rid_pairs.v[i] >> 32), b = (uint32_t) rid_pairs.v[i];
					if (i == 0) {
						inter[ni++] = a;
						if (b != a) inter[ni++] = b;
					} else {
						int k = 0;
						for (int j = 0; j < ni; j++)
							if (!(a != inter[j] && b != inter[j])) inter[k++] = inter[j];
						ni = k;
					}
				}
				if (ni == 1) { cls = CQO_CLASS_D_INTER; rid_a = inter[0]; }
				else cls = CQO_CLASS_CONFLICT;
			}
			break;
		}

		switch (cls) {
		case CQO_CLASS_UNLABELED: out->nundet++; break;
		case CQO_CLASS_CONFLICT: out->nconf++; break;
		case CQO_CLASS_U: out->cnt_u[rid_a]++; break;
		case CQO_CLASS_D_PAIR:
			out->cnt_d[rid_a]++;
			out->cnt_d[rid_b]++;
			if (mode == CQO_MODE_SC) add_pair(out, rid_a, rid_b, &err);
			break;
		case CQO_CLASS_UD: out->cnt_u[rid_a]++; out->cnt_d[rid_a]++; break;
		case CQO_CLASS_D_INTER:
			if (mode == CQO_MODE_SC) out->cnt_u[rid_a]++;
			out->cnt_d[rid_a]++;
			break;
		}
		if (mode == CQO_MODE_P && cls >= CQO_CLASS_U) {
			for (size_t i = 0; i < pnodes.n; i++) {
				uint64_t k = pnodes.v[i];
				if (!(k & TABLE_D)) { if (out->rcount_u) out->rcount_u[k]++; }
				else if (out->rcount_d) out->rcount_d[k & ~TABLE_D]++;
			}
		}

record:
		if (out->read_class) out->read_class[r] = cls;
		if (out->read_rid_a) out->read_rid_a[r] = rid_a;
		if (out->read_rid_b) out->read_rid_b[r] = rid_b;
		if (out->read_nleaf_u && out->read_nleaf_d) {
			uint32_t nu = 0, nd = 0;
			for (size_t i = 0; i < pnodes.n; i++) {
				uint64_t k = pnodes.v[i];
				if (!(k & TABLE_D)) {
					if (nu < out->leaf_cap && out->read_leaf_u) out->read_leaf_u[r * out->leaf_cap + nu] = (uint32_t) k;
					nu++;
				} else {
					if (nd < out->leaf_cap && out->read_leaf_d) out->read_leaf_d[r * out->leaf_cap + nd] = (uint32_t) (k & ~TABLE_D);
					nd++;
				}
			}
			out->read_nleaf_u[r] = nu;
			out->read_nleaf_d[r] = nd;
		}
	}
	free(pnodes.v);
	free(rids.v);
	free(rid_pairs.v);
	return err ? -1 : 0;
}
