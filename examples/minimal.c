/*
 * Smallest C program over include/cammiq_gpu.h: load an index pair, put it on GPU 0, scan a few
 * reads held as ASCII in host memory, print the counters.  Plain C99, no CUDA headers needed.
 *
 *   gcc -std=c99 -Iinclude examples/minimal.c -Lcammiq_b200 -lcammiq_gpu -Wl,-rpath,$PWD/cammiq_b200 -o minimal
 *   ./minimal index_u.bin1 index_d.bin2 <n_genomes> ACGT... [more reads]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cammiq_gpu.h"

static int die(const char *what) {
	fprintf(stderr, "%s: %s\n", what, cq_last_error());
	return 1;
}

int main(int argc, char **argv) {
	if (argc < 5) {
		fprintf(stderr, "usage: %s index_u.bin1 index_d.bin2 n_genomes read [read ...]\n", argv[0]);
		return 2;
	}
	printf("cammiq_gpu ABI %d, host packer: %s\n", cq_abi_version(), cq_pack_isa());
	const uint32_t n_genomes = (uint32_t) atoi(argv[3]);
	const uint64_t n_reads = (uint64_t) (argc - 4);

	/* reads back to back in one buffer + offsets + uint8 lengths: what cq_query consumes */
	size_t total = 0;
	for (int i = 4; i < argc; i++)
		total += strlen(argv[i]);
	uint8_t *bases = (uint8_t *) malloc(total + 1);
	uint64_t *offsets = (uint64_t *) malloc(n_reads * sizeof(uint64_t));
	uint8_t *lengths = (uint8_t *) malloc(n_reads);
	size_t at = 0;
	for (uint64_t r = 0; r < n_reads; r++) {
		size_t len = strlen(argv[4 + r]);
		if (len > 255)
			len = 255; /* read lengths are uint8_t, as in the reference (query.cpp:387) */
		memcpy(bases + at, argv[4 + r], len);
		offsets[r] = at;
		lengths[r] = (uint8_t) len;
		at += len;
	}

	cq_index *idx = NULL;
	cq_ctx *ctx = NULL;
	if (cq_index_load(argv[1], argv[2], 0.0, &idx) != CQ_OK)
		return die("cq_index_load");
	cq_index_info info;
	cq_index_get_info(idx, &info);
	printf("h = %u, %llu + %llu leaves, %llu keys\n", info.hash_len, (unsigned long long) info.n_leaves_u,
		(unsigned long long) info.n_leaves_d, (unsigned long long) info.n_keys);
	if (cq_ctx_create(0, NULL, &ctx) != CQ_OK)
		return die("cq_ctx_create"); /* no GPU: CQ_ENODEV, there is no CPU path */
	if (cq_index_upload(ctx, idx, n_genomes) != CQ_OK)
		return die("cq_index_upload");

	uint64_t *cnt_u = (uint64_t *) calloc(n_genomes + 1, sizeof(uint64_t));
	uint64_t *cnt_d = (uint64_t *) calloc(n_genomes + 1, sizeof(uint64_t));
	uint8_t *cls = (uint8_t *) calloc(n_reads, 1);
	uint32_t *rid_a = (uint32_t *) calloc(n_reads, sizeof(uint32_t));
	uint32_t *rid_b = (uint32_t *) calloc(n_reads, sizeof(uint32_t));
	cq_result res;
	memset(&res, 0, sizeof(res));
	res.cnt_u = cnt_u;
	res.cnt_d = cnt_d;
	res.read_class = cls;
	res.read_rid_a = rid_a;
	res.read_rid_b = rid_b;
	if (cq_query(ctx, CQ_MODE_P, bases, offsets, 0, lengths, n_reads, &res) != CQ_OK)
		return die("cq_query");
	printf("unlabeled %llu, conflicting %llu, invalid %llu\n", (unsigned long long) res.nundet,
		(unsigned long long) res.nconf, (unsigned long long) res.n_invalid);
	for (uint64_t r = 0; r < n_reads; r++)
		printf("read %llu: class %u, genome(s) %u %u\n", (unsigned long long) r, cls[r], rid_a[r], rid_b[r]);
	for (uint32_t g = 1; g <= n_genomes; g++)
		if (cnt_u[g] || cnt_d[g])
			printf("genome %u: %llu unique, %llu doubly-unique reads\n", g, (unsigned long long) cnt_u[g],
				(unsigned long long) cnt_d[g]);
	cq_ctx_destroy(ctx);
	cq_index_free(idx);
	free(bases); free(offsets); free(lengths); free(cnt_u); free(cnt_d); free(cls); free(rid_a); free(rid_b);
	return 0;
}
